"""CPU tests: the plain-C oracle against the LIVE unmodified reference (oracle/_ref/libxsi_ref.so,
built by `make -C oracle ref` where /root/reference exists) on seeded synthetic matrices that
exercise what the shipped fixtures do not: wide multi-allelic records, negated sparse lines,
bcf_int32_missing, the uint32 index path, small blocks, all-haploid records."""
import os
import tempfile

import numpy as np
import pytest

import synth
import xsi_oracle as xo
import xsi_ref

pytestmark = pytest.mark.skipif(not xsi_ref.available(), reason="oracle/_ref not built (no /root/reference here)")


def both(ds, block_len, maf, zstd=False):
    gt, ngt, nal, ns = ds["gt"], ds["ngt"], ds["n_allele"], ds["n_samples"]
    off = xo.row_offsets(ngt)
    dp = xo.default_phased(gt, off, ngt, ns)
    thr = xo.mac_threshold(ns, int(ngt[0]) // ns, maf)
    img = xo.encode(gt, off, ngt, nal, ns, block_len, thr, dp)
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "r.xsi")
        xsi_ref.encode_file(p, gt, off, ngt, nal, ns, block_len, thr, dp)
        ref = open(p, "rb").read()
        assert img == ref
        acc = xsi_ref.RefAccessor(p)
        rd = xo.Reader(img)
        pos = xo.bm_positions(nal, block_len)
        for r in range(len(nal)):
            a, na = acc.fill_genotype_array(int(nal[r]), int(pos[r]))
            b, nb = rd.fill_genotype_array(int(nal[r]), int(pos[r]))
            assert na == nb and np.array_equal(a[:na], b[:nb]), r
            assert np.array_equal(acc.allele_counts(), rd.allele_counts()), r
        acc.close()
        # counts only (Accessor::fill_allele_counts), fresh cursors on both sides, records in file order
        acc = xsi_ref.RefAccessor(p)
        rd = xo.Reader(img)
        for r in range(len(nal)):
            assert np.array_equal(acc.fill_allele_counts(int(nal[r]), int(pos[r])),
                                  rd.fill_allele_counts(int(nal[r]), int(pos[r]))), r
        acc.close()
    return img


def test_biallelic_ld():
    both(synth.make_dataset(700, 301, seed=1), block_len=256, maf=0.01)


def test_multiallelic_missing_eov_unphased():
    ds = synth.make_dataset(600, 257, seed=2, max_alt=4, multi_frac=0.2, missing=0.01, unphased=0.02,
                            haploid_samples=0.4)
    both(ds, block_len=128, maf=0.02)
    both(ds, block_len=8192, maf=0.3)   # everything sparse


def test_negated_sparse_and_int32_missing():
    ds = synth.make_dataset(200, 100, seed=3)
    gt = ds["gt"].reshape(200, 200)
    gt[5, :] = synth.encode_gt(np.ones(200, np.int8))     # ALT fixed -> negated sparse, empty list
    gt[6, :] = synth.encode_gt(np.ones(200, np.int8))
    gt[6, 17] = synth.encode_gt(np.zeros(1, np.int8))[0]   # one REF carrier
    gt[7, 3] = synth.I32_MISSING
    gt[8, 0::2] = 0                                       # unphased missing on first alleles
    ds["gt"] = np.ascontiguousarray(gt.reshape(-1))
    both(ds, block_len=64, maf=0.05)


def test_unphased_default():
    both(synth.make_dataset(300, 64, seed=4, phased=0, unphased=0.03, missing=0.01), block_len=100, maf=0.01)


def test_uint32_index_path():
    ds = synth.make_dataset(12, 66000, seed=5, n_founders=16, fmin=0.001)
    both(ds, block_len=5, maf=0.001)


def test_all_haploid_records():
    rng = np.random.default_rng(6)
    ns, nrec = 90, 150
    al = (rng.random((nrec, ns)) < rng.uniform(0.0, 0.6, size=(nrec, 1))).astype(np.int8)
    gt = synth.encode_gt(al, 0)
    ds = dict(gt=np.ascontiguousarray(gt.reshape(-1)), ngt=np.full(nrec, ns, np.int32),
              n_allele=np.full(nrec, 2, np.int32), n_samples=ns)
    both(ds, block_len=40, maf=0.05)


def test_mixed_ploidy_records_biallelic():
    # haploid and diploid RECORDS in one block (biallelic only: the reference does not round-trip
    # haploid + multi-allelic in one block) -- bytes and decode still agree
    rng = np.random.default_rng(7)
    ns = 120
    rows, ngt = [], []
    for r in range(200):
        p = 1 if r % 7 == 3 else 2
        al = (rng.random(ns * p) < 0.3).astype(np.int8)
        rows.append(synth.encode_gt(al, 1 if p == 2 else 0))
        ngt.append(ns * p)
    ds = dict(gt=np.concatenate(rows).astype(np.int32), ngt=np.array(ngt, np.int32),
              n_allele=np.full(200, 2, np.int32), n_samples=ns)
    both(ds, block_len=64, maf=0.01)


def both_wah_missing(ds, block_len, maf):
    """--wah-encode-missing (WS_WAH): missing / end-of-vector lines as natural-order WAH instead of index lists."""
    gt, ngt, nal, ns = ds["gt"], ds["ngt"], ds["n_allele"], ds["n_samples"]
    off = xo.row_offsets(ngt)
    dp = xo.default_phased(gt, off, ngt, ns)
    thr = xo.mac_threshold(ns, int(ngt[0]) // ns, maf)
    img = xo.encode(gt, off, ngt, nal, ns, block_len, thr, dp, wah_encode_missing=True)
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "r.xsi")
        xsi_ref.encode_file(p, gt, off, ngt, nal, ns, block_len, thr, dp, wah_encode_missing=True)
        assert img == open(p, "rb").read()
        acc = xsi_ref.RefAccessor(p)
        rd = xo.Reader(img)
        pos = xo.bm_positions(nal, block_len)
        for r in range(len(nal)):
            a, na = acc.fill_genotype_array(int(nal[r]), int(pos[r]))
            b, nb = rd.fill_genotype_array(int(nal[r]), int(pos[r]))
            assert na == nb and np.array_equal(a[:na], b[:nb]), r
            assert np.array_equal(acc.allele_counts(), rd.allele_counts()), r
        acc.close()
    return img


def test_wah_encode_missing():
    ds = synth.make_dataset(400, 257, seed=41, max_alt=3, multi_frac=0.2, missing=0.02, unphased=0.02, haploid_samples=0.4)
    img = both_wah_missing(ds, 128, 0.02)
    off = xo.row_offsets(ds["ngt"])
    plain = xo.encode(ds["gt"], off, ds["ngt"], ds["n_allele"], ds["n_samples"], 128, xo.mac_threshold(ds["n_samples"], 2, 0.02),
                      xo.default_phased(ds["gt"], off, ds["ngt"], ds["n_samples"]))
    assert img != plain
    both_wah_missing(synth.make_dataset(60, 66000, seed=42, n_founders=16, fmin=0.001, missing=0.001), 20, 0.001)


def test_haploid_and_multiallelic_block_bytes():
    """The reference WRITES a block with all-haploid and multi-allelic records (and the oracle writes the same bytes) but its
    reader corrupts the heap on it (probed in a child process: abort in munmap_chunk), so only the bytes are compared here;
    the product refuses to decode such a block (tests/test_gpu_parity.py::test_haploid_and_multiallelic_block)."""
    ds = synth.haploid_multiallelic_block()
    gt, ngt, nal, ns = ds["gt"], ds["ngt"], ds["n_allele"], ds["n_samples"]
    off = xo.row_offsets(ngt)
    dp = xo.default_phased(gt, off, ngt, ns)
    thr = xo.mac_threshold(ns, int(ngt[0]) // ns, 0.01)
    img = xo.encode(gt, off, ngt, nal, ns, 64, thr, dp)
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "r.xsi")
        xsi_ref.encode_file(p, gt, off, ngt, nal, ns, 64, thr, dp)
        assert img == open(p, "rb").read()
