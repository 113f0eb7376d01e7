"""Multi-GPU host logic: PBWT blocks are independent (a fresh GtBlock per block, reference
include/xsi_factory.hpp:527-539), so rank g of G encodes the contiguous block range
[g*B/G, (g+1)*B/G).  The only exchange is an all-gather of per-block on-disk byte counts (NCCL over
NVLink on GPUs, gloo in the CPU tests) from which every rank derives the global block offset table
(`indices[]`, xsi_factory.hpp:533,554,575); each rank then writes its own blocks at their final
offsets and rank 0 adds the index, the sample names and the header.  The file is byte-identical
to what the single writer (`xsi_writer_*`, reference XsiFactoryExt) produces."""
import os
import struct

import numpy as np

HEADER_BYTES = 256
KEY_GT_ENTRY = 256
_OUTER = struct.pack("<4I", 0xFFFFFFFF, 1, KEY_GT_ENTRY, 16)  # outer dictionary {KEY_GT_ENTRY: 16}, interfaces.hpp:187-221


def shard_range(n_blocks, rank, world):
    """Contiguous block range of `rank`."""
    return (n_blocks * rank) // world, (n_blocks * (rank + 1)) // world


def disk_size(gt_block_bytes):
    """Bytes one (non-zstd) block occupies in the file: outer dictionary + GT block, padded to 4
    (interfaces.hpp:254-263)."""
    return (16 + int(gt_block_bytes) + 3) // 4 * 4


def offset_table(all_disk_sizes):
    """Absolute file offset of every block (first block right after the 256-byte header) and the end."""
    off = np.zeros(len(all_disk_sizes) + 1, dtype=np.uint64)
    off[0] = HEADER_BYTES
    if len(all_disk_sizes):
        off[1:] = HEADER_BYTES + np.cumsum(np.asarray(all_disk_sizes, dtype=np.uint64))
    return off[:-1], int(off[-1])


def exchange(values, dist, device, world, reduce_max=False):
    """all-gather of a small int64 vector of identical length on every rank -> [world, len] numpy."""
    import torch
    t = torch.as_tensor(np.asarray(values, dtype=np.int64), device=device)
    if world == 1:
        return t.cpu().numpy()[None, :]
    out = torch.empty(world * t.numel(), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, t)
    return out.cpu().numpy().reshape(world, -1)


def write_sharded(path, rank, world, dist, device, my_blocks, n_blocks_total, n_samples, sample_names, block_len,
                  mac_threshold, default_phasing, my_records, my_variants, my_max_ploidy):
    """Every rank calls this with the GT blocks (bytes) of its shard_range, in order."""
    b0, b1 = shard_range(n_blocks_total, rank, world)
    assert len(my_blocks) == b1 - b0
    per = max(1, -(-n_blocks_total // world) + 1)
    mine = np.zeros(per + 3, dtype=np.int64)
    mine[:len(my_blocks)] = [disk_size(len(b)) for b in my_blocks]
    mine[per:per + 3] = (my_records, my_variants, my_max_ploidy)
    allv = exchange(mine, dist, device, world)
    sizes = []
    for g in range(world):
        g0, g1 = shard_range(n_blocks_total, g, world)
        sizes.extend(int(x) for x in allv[g, :g1 - g0])
    indices, end = offset_table(sizes)
    if rank == 0:
        with open(path, "wb") as f:
            f.truncate(end)
    if world > 1:
        dist.barrier()
    fd = os.open(path, os.O_WRONLY)
    try:
        for k, blk in enumerate(my_blocks):
            body = _OUTER + bytes(blk)
            body += b"\0" * (disk_size(len(blk)) - len(body))
            os.pwrite(fd, body, int(indices[b0 + k]))
    finally:
        os.close(fd)
    if world > 1:
        dist.barrier()
    if rank == 0:
        records = int(allv[:, per].sum())
        variants = int(allv[:, per + 1].sum())
        max_ploidy = int(allv[:, per + 2].max())
        # index, sample names and header: the single writer's own finalize code (csrc/xsi_container.cpp)
        import ctypes
        from . import lib
        blob = None if sample_names is None else b"".join(s.encode() + b"\0" for s in sample_names)
        idx = np.ascontiguousarray(indices, dtype=np.uint64)
        rc = lib().xsi_writer_finalize_sharded(path.encode(), int(n_samples), blob, int(block_len), int(mac_threshold), int(default_phasing),
                                               max_ploidy, len(idx), idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), int(end), records, variants)
        if rc != 0:
            raise RuntimeError("xsi_writer_finalize_sharded failed (%d)" % rc)
    if world > 1:
        dist.barrier()
    return indices
