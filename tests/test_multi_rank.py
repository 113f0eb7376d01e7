"""World-size-2 test of the multi-GPU host logic on CPU (gloo): block sharding, the all-gather of
per-block byte counts, the global offset table and the sharded write must reproduce the file the
unmodified reference wrote (tests/golden/chr20_small_default.xsi, 3 blocks), byte for byte."""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "chr20_small_default.xsi")

WORKER = r'''
import ctypes, os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import torch, torch.distributed as dist
import xsqueezeit_b200 as xb
from xsqueezeit_b200 import sharded
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
L = xb.lib()
r = ctypes.c_void_p()
assert L.xsi_reader_open(%(gold)r.encode(), ctypes.byref(r)) == 0
acc = xb.Accessor.__new__(xb.Accessor)
ns, hs, pl, aet, nb, bl = (ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32())
ent, nv, z, rt, dp = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_int32(), ctypes.c_uint64(), ctypes.c_int32()
L.xsi_reader_info(r, ctypes.byref(ns), ctypes.byref(hs), ctypes.byref(pl), ctypes.byref(aet), ctypes.byref(nb), ctypes.byref(bl),
                  ctypes.byref(ent), ctypes.byref(nv), ctypes.byref(z), ctypes.byref(rt), ctypes.byref(dp))
names = [L.xsi_reader_sample_name(r, i).decode() for i in range(ns.value)]
b0, b1 = sharded.shard_range(nb.value, rank, world)
blocks = []
for b in range(b0, b1):
    p, s = ctypes.c_void_p(), ctypes.c_uint64()
    assert L.xsi_reader_gt_block(r, b, ctypes.byref(p), ctypes.byref(s)) == 0
    blocks.append(ctypes.string_at(p.value, s.value))
# records / variants of this shard: KEY_BCF_LINES (0) / KEY_BINARY_LINES (1) of each block's dictionary
recs = vars_ = 0
for blk in blocks:
    n = int(np.frombuffer(blk[4:8], "<u4")[0])
    d = dict(np.frombuffer(blk[8:8 + 8 * n], "<u4").reshape(n, 2).tolist())
    recs += d[0]; vars_ += d[1]
sharded.write_sharded(%(out)r, rank, world, dist, "cpu", blocks, nb.value, ns.value, names, bl.value, rt.value, dp.value,
                      recs, vars_, pl.value)
dist.destroy_process_group()
'''


def test_sharded_write_two_ranks(tmp_path):
    out = str(tmp_path / "sharded.xsi")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "gold": GOLD, "out": out})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29731", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r))) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    assert open(out, "rb").read() == open(GOLD, "rb").read()


def test_shard_ranges_cover_blocks():
    from xsqueezeit_b200 import sharded
    for nb in (1, 3, 25, 123):
        for world in (1, 2, 4, 8):
            got = []
            for g in range(world):
                b0, b1 = sharded.shard_range(nb, g, world)
                got.extend(range(b0, b1))
            assert got == list(range(nb))
    idx, end = sharded.offset_table([sharded.disk_size(10), sharded.disk_size(17)])
    assert list(idx) == [256, 256 + 28] and end == 256 + 28 + 36
