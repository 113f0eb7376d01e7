"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol that
include/xsi_b200.h declares; the container layer (writer/reader) round-trips and matches the
reference header layout; no compute call works without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def xb():
    import xsqueezeit_b200 as m
    if not os.path.exists(m.SO_PATH):
        from xsqueezeit_b200 import build
        build.build()
    return m


def test_exports_every_declared_symbol(xb):
    hdr = open(os.path.join(ROOT, "include", "xsi_b200.h")).read()
    names = sorted(set(re.findall(r"\b(xsi_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    L = ctypes.CDLL(xb.SO_PATH)
    for n in names:
        assert hasattr(L, n), n


def test_no_cpu_fallback_without_device(xb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(xb.XsiError) as e:
        xb.Context(0)
    assert e.value.code == -1


def test_file_level_parameters(xb):
    import xsi_oracle as xo
    rng = np.random.default_rng(0)
    for dp in (0, 1):
        rows = [((rng.integers(1, 3, 40) << 1) | (rng.random(40) < (0.8 if dp else 0.2))).astype(np.int32) for _ in range(5)]
        gt = np.concatenate(rows)
        ngt = np.full(5, 40, np.int32)
        assert xb.seek_default_phased((r, 2) for r in rows) == xo.default_phased(gt, xo.row_offsets(ngt), ngt, 20)
    assert xb.seek_default_phased([(np.zeros(10, np.int32), 1)]) == 0
    for ns, pl, maf in ((2504, 2, 0.001), (32488, 2, 0.001), (10, 2, 0.002), (500000, 2, 0.001), (90, 1, 0.05)):
        assert xb.mac_threshold(ns, pl, maf) == xo.mac_threshold(ns, pl, maf)
    nal = rng.integers(2, 5, 3000)
    assert np.array_equal(xb.bm_positions(nal, 256), xo.bm_positions(nal, 256))


def test_container_roundtrip_and_reference_reader(xb, tmp_path):
    """Write GT blocks taken from a reference-written file through xsi_writer, get the same file."""
    import json
    G = os.path.join(ROOT, "tests", "golden")
    man = json.load(open(os.path.join(G, "manifest.json")))
    for name in ("micro_missing_non_uniform_phasing_ploidy", "test_region_target", "chr20_small_default"):
        gold_path = os.path.join(G, name + ".xsi")
        gold = open(gold_path, "rb").read()
        L = xb.lib()
        r = ctypes.c_void_p()
        assert L.xsi_reader_open(gold_path.encode(), ctypes.byref(r)) == 0
        acc = xb.Accessor.__new__(xb.Accessor)  # reader only, no GPU context
        ns, hs, pl, aet, nb, bl = (ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint32(), ctypes.c_uint32(),
                                   ctypes.c_uint32(), ctypes.c_uint32())
        ent, nv, z, rt, dp = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_int32(), ctypes.c_uint64(), ctypes.c_int32()
        L.xsi_reader_info(r, ctypes.byref(ns), ctypes.byref(hs), ctypes.byref(pl), ctypes.byref(aet), ctypes.byref(nb),
                          ctypes.byref(bl), ctypes.byref(ent), ctypes.byref(nv), ctypes.byref(z), ctypes.byref(rt), ctypes.byref(dp))
        names = b"".join(L.xsi_reader_sample_name(r, i) + b"\0" for i in range(hs.value // pl.value))
        for zstd in (0, 1):
            out = str(tmp_path / (name + "_%d.xsi" % zstd))
            w = ctypes.c_void_p()
            assert L.xsi_writer_open(out.encode(), ns.value, names, bl.value, rt.value, dp.value, zstd, 7, ctypes.byref(w)) == 0
            for b in range(nb.value):
                p, s = ctypes.c_void_p(), ctypes.c_uint64()
                assert L.xsi_reader_gt_block(r, b, ctypes.byref(p), ctypes.byref(s)) == 0
                blk = ctypes.string_at(p.value, s.value)
                # strip the <=3 alignment bytes the file adds after the block
                nxt = np.frombuffer(gold, np.uint64, nb.value, int(np.frombuffer(gold, np.uint64, 1, 72)[0]))
                end = int(nxt[b + 1]) if b + 1 < nb.value else None
                ptr = (ctypes.c_void_p * 1)(ctypes.cast(ctypes.c_char_p(blk), ctypes.c_void_p).value)
                size = (ctypes.c_uint64 * 1)(len(blk))
                recs = min(bl.value, ent.value - b * bl.value)
                assert L.xsi_writer_add_blocks(w, 1, ptr, size, recs, 0) == 0
            assert L.xsi_writer_close(w, pl.value) == 0
            got = open(out, "rb").read()
            r2 = ctypes.c_void_p()
            assert L.xsi_reader_open(out.encode(), ctypes.byref(r2)) == 0
            for b in range(nb.value):
                p, s, p2, s2 = ctypes.c_void_p(), ctypes.c_uint64(), ctypes.c_void_p(), ctypes.c_uint64()
                L.xsi_reader_gt_block(r, b, ctypes.byref(p), ctypes.byref(s))
                assert L.xsi_reader_gt_block(r2, b, ctypes.byref(p2), ctypes.byref(s2)) == 0
                assert ctypes.string_at(p.value, s.value) == ctypes.string_at(p2.value, s2.value)
            L.xsi_reader_close(r2)
            if not zstd:
                # everything but num_variants (not passed above) must equal the reference's bytes
                a = bytearray(got); g = bytearray(gold)
                a[40:48] = g[40:48]
                assert bytes(a) == bytes(g), name
        L.xsi_reader_close(r)


def test_reader_rejects_garbage(xb, tmp_path):
    p = tmp_path / "bad.xsi"
    p.write_bytes(b"\0" * 300)
    r = ctypes.c_void_p()
    assert xb.lib().xsi_reader_open(str(p).encode(), ctypes.byref(r)) == -6
    assert xb.lib().xsi_reader_open(b"/nonexistent/file.xsi", ctypes.byref(r)) == -8


def test_host_int8_transport_conversions(xb):
    """int32 <-> BCF int8 transport encoding (csrc/host_narrow.cpp) against numpy, every ISA level it dispatches to."""
    L = xb.lib()
    rng = np.random.default_rng(7)
    n = 3 * (1 << 18) + 77  # several pool tasks + a ragged tail
    src = rng.integers(0, 128, n).astype(np.int32)
    src[rng.integers(0, n, 1000)] = np.int32(-2**31)        # bcf_int32_missing
    src[rng.integers(0, n, 1000)] = np.int32(-2**31 + 1)    # bcf_int32_vector_end
    want = np.where(src < 0, 0x80 | (src & 1), src).astype(np.uint8)
    for m in (0, 1, 15, 16, 63, 64, 65, 4097, n):
        dst = np.zeros(m + 8, np.uint8)
        dst[m:] = 0xEE
        assert L.xsi_host_narrow_i32_i8(src.ctypes.data, dst.ctypes.data, m) == 1
        assert np.array_equal(dst[:m], want[:m]) and (dst[m:] == 0xEE).all()
    # values without an int8 encoding are reported, wherever they sit
    for bad in (128, 255, 256, 1 << 20, -1, -2**31 + 2, -2**31 + 0x80, 2**31 - 1):
        for at in (0, 17, n - 1, n // 2):
            s2 = src.copy()
            s2[at] = np.int32(bad)
            dst = np.zeros(n, np.uint8)
            assert L.xsi_host_narrow_i32_i8(s2.ctypes.data, dst.ctypes.data, n) == 0, (bad, at)
    # widen: ragged rows, untouched tails, unaligned destinations
    rows, s8, s32 = 37, 4099, 4111
    b = rng.integers(0, 128, (rows, s8)).astype(np.uint8)
    b[rng.random((rows, s8)) < 0.01] = 0x80
    b[rng.random((rows, s8)) < 0.01] = 0x81
    ln = rng.integers(0, s8 + 1, rows).astype(np.uint32)
    ln[0], ln[1] = 0, s8
    out = np.full((rows, s32), 0x5A5A5A5A, np.int32)
    L.xsi_host_widen_i8_i32(b.ctypes.data, s8, out.ctypes.data, s32, ln.ctypes.data, rows)
    wide = np.where(b >= 0x80, np.int64(-2**31) + (b & 1), b).astype(np.int32)
    for r in range(rows):
        assert np.array_equal(out[r, :ln[r]], wide[r, :ln[r]]) and (out[r, ln[r]:] == 0x5A5A5A5A).all()
    # one long row (split over several tasks) and the round trip
    big = np.zeros(n, np.int32)
    L.xsi_host_widen_i8_i32(want.ctypes.data, n, big.ctypes.data, n, np.array([n], np.uint32).ctypes.data, 1)
    assert np.array_equal(big, src)
    assert L.xsi_host_threads() >= 1
