"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI
(include/xsi_b200.h via xsqueezeit_b200.Context / Compressor / Accessor), against the oracle
(oracle/xsi_oracle.c, itself pinned to the reference in test_oracle_golden.py) and against the
committed golden .xsi files produced by the unmodified reference.  Bit-exact everywhere."""
import hashlib
import json
import os

import numpy as np
import pytest

import synth
import xsi_oracle as xo

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MAN = json.load(open(os.path.join(G, "manifest.json")))
SMALL = sorted(k for k in MAN if k != "chr20_small")


@pytest.fixture(scope="module")
def ctx():
    import xsqueezeit_b200 as xb
    c = xb.Context(0)
    yield c
    c.close()


def gpu_encode(ctx, tmp_path, ds, block_len, maf, names=None, elem=4, zstd=False, blocks_per_batch=8):
    import xsqueezeit_b200 as xb
    comp = xb.Compressor(ctx, maf=maf, reset_sort_block_length=block_len, zstd_compression_on=zstd,
                         blocks_per_batch=blocks_per_batch)
    p = str(tmp_path / "gpu.xsi")
    gt = ds["gt"]
    if elem == 1:
        g8 = gt.astype(np.int64)
        g8 = np.where(gt == synth.I32_MISSING, -128, np.where(gt == synth.EOV, -127, g8)).astype(np.int8)
        gt = g8
    comp.compress_to_file(p, gt, ds["ngt"], ds["n_allele"], ds["n_samples"], sample_names=names, gt_elem_bytes=elem)
    return p


def oracle_image(ds, block_len, maf, names=None):
    gt, ngt, nal, ns = ds["gt"], ds["ngt"], ds["n_allele"], ds["n_samples"]
    off = xo.row_offsets(ngt)
    dp = xo.default_phased(gt, off, ngt, ns)
    thr = xo.mac_threshold(ns, int(ngt[0]) // ns, maf)
    return xo.encode(gt, off, ngt, nal, ns, block_len, thr, dp, names)


def check_decode(ctx, path, image, n_allele, block_len, one_by_one=False):
    import xsqueezeit_b200 as xb
    acc = xb.Accessor(path, ctx)
    rd = xo.Reader(image)
    pos = xo.bm_positions(n_allele, block_len)
    assert np.array_equal(pos, xb.bm_positions(n_allele, block_len))
    if one_by_one:
        for r in range(len(n_allele)):
            a, na = acc.fill_genotype_array(int(n_allele[r]), int(pos[r]))
            b, nb = rd.fill_genotype_array(int(n_allele[r]), int(pos[r]))
            assert na == nb and np.array_equal(a[:na], b[:nb]), r
            assert np.array_equal(acc.get_allele_counts(), rd.allele_counts()), r
    else:
        out, filled, counts = acc.fill_genotype_arrays(n_allele, pos, want_counts=True)
        for r in range(len(n_allele)):
            b, nb = rd.fill_genotype_array(int(n_allele[r]), int(pos[r]))
            assert filled[r] == nb and np.array_equal(out[r, :nb], b[:nb]), r
            assert np.array_equal(counts[r, :int(n_allele[r])], rd.allele_counts()), r
        # counts only (xsi_decode_allele_counts == Accessor::fill_allele_counts), oracle in file order on a fresh cursor
        ac = acc.fill_allele_counts_batch(n_allele, pos)
        rd2 = xo.Reader(image)
        for r in range(len(n_allele)):
            assert np.array_equal(ac[r, :int(n_allele[r])], rd2.fill_allele_counts(int(n_allele[r]), int(pos[r]))), r
        rd2.close()
        if int(np.max(n_allele)) <= 63:
            # raw BCF int8 rows (xsi_decode_records_i8): the int32 values narrowed the way htslib stores them
            out8, filled8, _ = acc.fill_genotype_arrays(n_allele, pos, elem_bytes=1)
            assert np.array_equal(filled8, filled)
            for r in range(len(n_allele)):
                nb = int(filled[r])
                assert np.array_equal(out8[r, :nb], narrow_i8(out[r, :nb])), r
    acc.close()


def narrow_i8(row):
    """int32 genotype values -> BCF int8 payload (vector end 0x81, htslib vcf.h bcf_int8_vector_end)."""
    return np.where(row == synth.EOV, -127, row).astype(np.int8)


def roundtrip(ctx, tmp_path, ds, block_len, maf, elem=4, one_by_one=False, blocks_per_batch=8):
    p = gpu_encode(ctx, tmp_path, ds, block_len, maf, elem=elem, blocks_per_batch=blocks_per_batch)
    img = oracle_image(ds, block_len, maf)
    got = open(p, "rb").read()
    assert len(got) == len(img)
    assert got == img
    check_decode(ctx, p, img, ds["n_allele"], block_len, one_by_one=one_by_one)


# ---- golden fixtures written by the unmodified reference --------------------------------------
@pytest.mark.parametrize("name", SMALL)
def test_golden_small_fixture(ctx, tmp_path, name):
    d = np.load(os.path.join(G, name + ".npz"))
    ds = dict(gt=d["gt"], ngt=d["ngt"], n_allele=d["n_allele"], n_samples=int(d["n_samples"]))
    names = [str(x) for x in d["names"]]
    p = gpu_encode(ctx, tmp_path, ds, int(d["block_len"]), float(d["maf"]), names=names)
    got = open(p, "rb").read()
    gold = open(os.path.join(G, name + ".xsi"), "rb").read()
    assert hashlib.sha256(got).hexdigest() == MAN[name]["xsi_sha256"]
    assert got == gold
    # decode the REFERENCE's file and compare with the reference Accessor's own output
    import xsqueezeit_b200 as xb
    acc = xb.Accessor(os.path.join(G, name + ".xsi"), ctx)
    pos = xb.bm_positions(d["n_allele"], int(d["block_len"]))
    out, filled, _ = acc.fill_genotype_arrays(d["n_allele"], pos)
    dec = np.concatenate([out[i, :filled[i]] for i in range(len(pos))])
    assert list(filled) == list(d["ref_decoded_ngt"])
    assert np.array_equal(dec, d["ref_decoded"])
    assert acc.get_sample_list() == names


def test_golden_chr20_small(ctx, tmp_path):
    import xsqueezeit_b200 as xb
    man = MAN["chr20_small"]
    d = np.load(os.path.join(G, "chr20_small_meta.npz"))
    ns, nal, ngt = int(d["n_samples"]), d["n_allele"].astype(np.uint32), d["ngt"]
    names = [str(x) for x in d["names"]]
    gold_path = os.path.join(G, "chr20_small_default.xsi")
    acc = xb.Accessor(gold_path, ctx)
    pos = xb.bm_positions(nal, 8192)
    out, filled, _ = acc.fill_genotype_arrays(nal, pos)
    assert (filled == ngt).all()
    gt = np.ascontiguousarray(out[:, :2 * ns]).reshape(-1)
    # == bcf_get_genotypes on the original BCF == reference Accessor decode
    assert hashlib.sha256(gt.tobytes()).hexdigest() == man["gt_sha256"]
    ds = dict(gt=gt, ngt=ngt, n_allele=nal, n_samples=ns)
    for key, o in man["options"].items():
        argv = o["argv"]
        maf = float(argv[argv.index("--maf") + 1]) if "--maf" in argv else 0.001
        bl = int(argv[argv.index("--variant-block-length") + 1]) if "--variant-block-length" in argv else 8192
        p = gpu_encode(ctx, tmp_path, ds, bl, maf, names=names)
        got = open(p, "rb").read()
        assert len(got) == o["xsi_size"], key
        assert hashlib.sha256(got).hexdigest() == o["xsi_sha256"], key


# ---- seeded synthetic matrices against the oracle -----------------------------------------------
def test_biallelic_ld(ctx, tmp_path):
    roundtrip(ctx, tmp_path, synth.make_dataset(700, 301, seed=1), 256, 0.01, one_by_one=True)


def test_int8_input_equals_int32(ctx, tmp_path):
    ds = synth.make_dataset(500, 333, seed=11, missing=0.01, haploid_samples=0.3, unphased=0.02)
    roundtrip(ctx, tmp_path, ds, 128, 0.01, elem=1)


def test_multiallelic_missing_eov_unphased(ctx, tmp_path):
    ds = synth.make_dataset(600, 257, seed=2, max_alt=4, multi_frac=0.2, missing=0.01, unphased=0.02, haploid_samples=0.4)
    roundtrip(ctx, tmp_path, ds, 128, 0.02)
    roundtrip(ctx, tmp_path, ds, 8192, 0.3)


def test_many_alleles(ctx, tmp_path):
    ds = synth.make_dataset(120, 200, seed=12, max_alt=9, multi_frac=0.5)
    roundtrip(ctx, tmp_path, ds, 50, 0.02)


def test_negated_sparse_and_int32_missing(ctx, tmp_path):
    ds = synth.make_dataset(200, 100, seed=3)
    gt = ds["gt"].reshape(200, 200)
    gt[5, :] = synth.encode_gt(np.ones(200, np.int8))
    gt[6, :] = synth.encode_gt(np.ones(200, np.int8))
    gt[6, 17] = synth.encode_gt(np.zeros(1, np.int8))[0]
    gt[7, 3] = synth.I32_MISSING
    gt[8, 0::2] = 0
    ds["gt"] = np.ascontiguousarray(gt.reshape(-1))
    roundtrip(ctx, tmp_path, ds, 64, 0.05, one_by_one=True)


def test_unphased_default(ctx, tmp_path):
    roundtrip(ctx, tmp_path, synth.make_dataset(300, 64, seed=4, phased=0, unphased=0.03, missing=0.01), 100, 0.01)


def test_all_haploid_records(ctx, tmp_path):
    rng = np.random.default_rng(6)
    ns, nrec = 90, 150
    al = (rng.random((nrec, ns)) < rng.uniform(0.0, 0.6, size=(nrec, 1))).astype(np.int8)
    ds = dict(gt=np.ascontiguousarray(synth.encode_gt(al, 0).reshape(-1)), ngt=np.full(nrec, ns, np.int32),
              n_allele=np.full(nrec, 2, np.int32), n_samples=ns)
    roundtrip(ctx, tmp_path, ds, 40, 0.05)


def test_mixed_ploidy_records(ctx, tmp_path):
    rng = np.random.default_rng(7)
    ns = 1200
    rows, ngt = [], []
    for r in range(200):
        p = 1 if r % 7 == 3 else 2
        al = (rng.random(ns * p) < 0.3).astype(np.int8)
        rows.append(synth.encode_gt(al, 1 if p == 2 else 0))
        ngt.append(ns * p)
    ds = dict(gt=np.concatenate(rows).astype(np.int32), ngt=np.array(ngt, np.int32), n_allele=np.full(200, 2, np.int32),
              n_samples=ns)
    roundtrip(ctx, tmp_path, ds, 64, 0.01)


def test_ragged_last_block_and_single_record(ctx, tmp_path):
    roundtrip(ctx, tmp_path, synth.make_dataset(1, 17, seed=8), 8192, 0.0)
    roundtrip(ctx, tmp_path, synth.make_dataset(257, 1000, seed=9), 128, 0.001, blocks_per_batch=1)


def test_kgp_shape_blocks(ctx, tmp_path):
    # 1KGP3 shape: 2,504 samples, two full 8192-record blocks + a partial one would be 100M genotypes;
    # 3000 records in blocks of 1024 keeps the oracle in seconds and still crosses batch boundaries
    ds = synth.make_dataset(3000, 2504, seed=13, n_founders=128)
    roundtrip(ctx, tmp_path, ds, 1024, 0.001, blocks_per_batch=2)


def test_hrc_shape_uint16_limit(ctx, tmp_path):
    # HRC shape: 32,488 samples = 64,976 haplotypes (uint16 permutation in shared memory)
    ds = synth.make_dataset(160, 32488, seed=14, n_founders=128, fmin=0.0005)
    roundtrip(ctx, tmp_path, ds, 64, 0.001)


def test_max_uint16_samples(ctx, tmp_path):
    ds = synth.make_dataset(40, 32767, seed=15, n_founders=64, fmin=0.001)
    roundtrip(ctx, tmp_path, ds, 16, 0.001)


def test_uint32_index_path(ctx, tmp_path):
    ds = synth.make_dataset(12, 66000, seed=5, n_founders=16, fmin=0.001)
    roundtrip(ctx, tmp_path, ds, 5, 0.001)


def test_biobank_width(ctx, tmp_path, monkeypatch):
    # S5 shape: 500,000 samples = 1,000,000 haplotypes (uint32 indices, WAH lines longer than 65,535 words)
    # encode: grid-wide cooperative PBWT kernel (6 blocks = groups of 4 + 2 at 32 haplotypes per thread, and one
    # block at a time at 8); decode: the wide inverse-permutation kernel with L2-resident tables
    ds = synth.make_dataset(48, 500000, seed=17, n_founders=16, fmin=0.0005)
    roundtrip(ctx, tmp_path, ds, 8, 0.001)
    monkeypatch.setenv("XSI_PBWT_KH", "8")
    monkeypatch.setenv("XSI_UNPERM_KH", "32")
    monkeypatch.setenv("XSI_UNPERM_WINDOW", "3")  # several launches per block: positions travel through pos_state
    roundtrip(ctx, tmp_path, ds, 20, 0.001)


def test_zstd_layer_roundtrip(ctx, tmp_path):
    import xsqueezeit_b200 as xb
    ds = synth.make_dataset(400, 150, seed=16, missing=0.01)
    p = gpu_encode(ctx, tmp_path, ds, 100, 0.01, zstd=True)
    img = oracle_image(ds, 100, 0.01)
    acc = xb.Accessor(p, ctx)
    assert acc.zstd
    acc.close()
    check_decode(ctx, p, img, ds["n_allele"], 100)


def test_int8_rows_aligned_fast_path(ctx, tmp_path):
    """int8 egress through the TMA-store kernel (rows a multiple of 16 bytes) incl. sparse and negated-sparse lines."""
    ds = synth.make_dataset(400, 1024, seed=21)
    roundtrip(ctx, tmp_path, ds, 128, 0.01)
    ds = synth.make_dataset(300, 1000, seed=22, missing=0.002)  # diploid rows 2000 B: 16-byte multiple; odd ones are not
    roundtrip(ctx, tmp_path, ds, 128, 0.01)


def test_int8_rows_reject_wide_alleles(ctx, tmp_path):
    import xsqueezeit_b200 as xb
    ds = synth.make_dataset(20, 64, seed=23, max_alt=70, multi_frac=1.0)
    if int(ds["n_allele"].max()) <= 63:
        pytest.skip("generator produced no record with more than 63 alleles")
    p = gpu_encode(ctx, tmp_path, ds, 128, 0.01)
    acc = xb.Accessor(p, ctx)
    pos = xb.bm_positions(ds["n_allele"], 128)
    with pytest.raises(xb.XsiError) as e:
        acc.fill_genotype_arrays(ds["n_allele"], pos, elem_bytes=1)
    assert e.value.code == -5  # XSI_E_UNSUPPORTED
    acc.close()


def test_host_rows_int8_transport_and_int32_fallbacks(ctx, tmp_path, monkeypatch):
    """Host int32 rows cross PCIe as int8 (csrc/host_narrow.cpp) when every value fits, as int32 otherwise or when
    XSI_HOST_NARROW=0; the .xsi bytes and the decoded rows are the same on all three routes."""
    ds = synth.make_dataset(500, 520, seed=31, max_alt=3, multi_frac=0.1, missing=0.01, unphased=0.02, haploid_samples=0.3)
    ds["gt"][7] = synth.I32_MISSING
    h0, d0 = ctx.transport_stats
    roundtrip(ctx, tmp_path, ds, 128, 0.01)
    h1, d1 = ctx.transport_stats
    assert h1 - h0 == ds["gt"].size and d1 > d0
    monkeypatch.setenv("XSI_HOST_NARROW", "0")
    roundtrip(ctx, tmp_path, ds, 128, 0.01)
    assert ctx.transport_stats == (h1, d1)
    monkeypatch.delenv("XSI_HOST_NARROW")
    # alleles above 62 have no int8 encoding: the upload and the download fall back to int32 on their own
    wide = synth.make_dataset(40, 64, seed=23, max_alt=70, multi_frac=1.0)
    if int(wide["n_allele"].max()) <= 64 or int((wide["gt"] > 127).sum()) == 0:
        pytest.skip("generator produced no genotype value above 127")
    roundtrip(ctx, tmp_path, wide, 16, 0.01)
    assert ctx.transport_stats[0] == h1


def test_error_codes(ctx):
    import xsqueezeit_b200 as xb
    gt = synth.encode_gt(np.zeros((4, 20), np.int8)).reshape(-1).copy()
    gt[5] = synth.encode_gt(np.array([3], np.int8))[0]  # allele 3 with n_allele 2
    with pytest.raises(xb.XsiError) as e:
        ctx.encode_launch(gt, [2, 2, 2, 2], 10, 8192, 0, 1)
    assert e.value.code == -3
    with pytest.raises(xb.XsiError) as e:
        ctx.encode_launch(np.zeros(30, np.int32), [2], 10, 8192, 0, 1, ploidy=[3])
    assert e.value.code == -4
    with pytest.raises(xb.XsiError) as e:
        ctx.encode_launch(np.zeros(2 * 40000, np.int32), [2], 40000, 8192, 0, 1)
    assert e.value.code == -5


def test_concurrent_contexts(tmp_path, monkeypatch):
    """Four host threads, one xsi_ctx each, encode and decode HRC-width blocks at the same time (one context per thread
    is the C ABI's threading model).  XSI_UNPERM_NC=128 multiplies the CTAs of the inverse-PBWT kernel, which is what
    exposed a stage of its table ring being refilled while reads of it were still queued."""
    import threading
    import xsqueezeit_b200 as xb
    monkeypatch.setenv("XSI_UNPERM_NC", "128")
    ds = synth.make_dataset(192, 32488, seed=33, n_founders=128, fmin=0.0005)
    img = oracle_image(ds, 64, 0.001)
    rd = xo.Reader(img)
    pos = xo.bm_positions(ds["n_allele"], 64)
    want = [rd.fill_genotype_array(2, int(pos[r]))[0].copy() for r in range(192)]
    errs = []

    def work(w):
        try:
            c = xb.Context(0)
            for it in range(3):
                p = str(tmp_path / ("t%d_%d.xsi" % (w, it)))
                xb.Compressor(c, maf=0.001, reset_sort_block_length=64).compress_to_file(p, ds["gt"], ds["ngt"], ds["n_allele"], ds["n_samples"])
                if open(p, "rb").read() != img:
                    errs.append("thread %d: .xsi bytes differ" % w)
                acc = xb.Accessor(p, c)
                out, filled, _ = acc.fill_genotype_arrays(ds["n_allele"], pos)
                for r in range(192):
                    if filled[r] != want[r].size or not np.array_equal(out[r, :filled[r]], want[r]):
                        errs.append("thread %d: row %d differs" % (w, r))
                        break
                acc.close()
            c.close()
        except Exception as ex:  # surfaced after join
            errs.append(repr(ex))

    ts = [threading.Thread(target=work, args=(w,)) for w in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs[:3]


def test_pinned_host_rows_two_routes(ctx, tmp_path):
    """PINNED host int32 rows: chunks are split between the host conversion (worker pool + int8 DMA) and the plain int32
    DMA + device conversion kernel, whichever route is free.  Same .xsi bytes, same rows."""
    import torch
    import xsqueezeit_b200 as xb
    ds = synth.make_dataset(1100, 32488, seed=35, n_founders=128, fmin=0.0005)
    n_el = ds["gt"].size
    assert n_el >= 4 * (16 << 20)  # enough 16 Mi-genotype chunks for both routes
    gt = torch.from_numpy(ds["gt"]).pin_memory()
    img = oracle_image(ds, 8192, 0.001)
    h0, d0 = ctx.transport_stats
    p = str(tmp_path / "pinned.xsi")
    xb.Compressor(ctx, maf=0.001, reset_sort_block_length=8192, blocks_per_batch=8).compress_to_file(
        p, gt.numpy(), ds["ngt"], ds["n_allele"], ds["n_samples"])
    assert open(p, "rb").read() == img
    h1, d1 = ctx.transport_stats
    assert 0 < h1 - h0 < n_el, "both upload routes should have carried chunks"
    acc = xb.Accessor(p, ctx)
    pos = xb.bm_positions(ds["n_allele"], 8192)
    off = (pos & np.uint64(0x7FFF)).astype(np.uint32)
    acc._load(0, 1)
    out = torch.empty((1100, 64976), dtype=torch.int32).pin_memory()
    out.fill_(-7)
    o, filled, _ = ctx.decode_records(np.zeros(1100, np.uint32), off, ds["n_allele"], out=out.numpy(), out_stride=64976)
    h2, d2 = ctx.transport_stats
    assert 0 < d2 - d1 < n_el, "both download routes should have carried rows"
    rd = xo.Reader(img)
    res = out.numpy()
    for r in range(1100):
        b, nb = rd.fill_genotype_array(2, int(pos[r]))
        assert filled[r] == nb and np.array_equal(res[r, :nb], b[:nb]), r
    acc.close()


@pytest.mark.parametrize("name", SMALL)
def test_decode_reference_wah_encode_missing_files(ctx, name):
    """Files the reference CLI wrote with --wah-encode-missing (missing / end-of-vector lines as natural-order WAH,
    KEY_WEIRDNESS_STRATEGY = WS_WAH) decode to the same rows as the default files."""
    import xsqueezeit_b200 as xb
    d = np.load(os.path.join(G, name + ".npz"))
    nal = d["n_allele"]
    acc = xb.Accessor(os.path.join(G, name + "_wah_missing.xsi"), ctx)
    pos = xb.bm_positions(nal, int(d["block_len"]))
    out, filled, counts = acc.fill_genotype_arrays(nal, pos, want_counts=True)
    rd = xo.Reader(open(os.path.join(G, name + "_wah_missing.xsi"), "rb").read())
    got = np.concatenate([out[r, :filled[r]] for r in range(len(nal))])
    assert np.array_equal(got, d["ref_decoded"])
    for r in range(len(nal)):
        rd.fill_genotype_array(int(nal[r]), int(pos[r]))
        assert np.array_equal(counts[r, :int(nal[r])], rd.allele_counts()), r
    acc.close()


def test_decode_wah_encode_missing_synthetic(ctx, tmp_path):
    """WS_WAH images written by the oracle (pinned to the reference for this mode): wide multi-allelic mixed-ploidy rows."""
    import xsqueezeit_b200 as xb
    ds = synth.make_dataset(500, 520, seed=43, max_alt=3, multi_frac=0.1, missing=0.01, unphased=0.02, haploid_samples=0.3)
    gt, ngt, nal, ns = ds["gt"], ds["ngt"], ds["n_allele"], ds["n_samples"]
    off = xo.row_offsets(ngt)
    img = xo.encode(gt, off, ngt, nal, ns, 128, xo.mac_threshold(ns, 2, 0.01), xo.default_phased(gt, off, ngt, ns), wah_encode_missing=True)
    p = str(tmp_path / "wm.xsi")
    open(p, "wb").write(img)
    check_decode(ctx, p, img, nal, 128)


def test_encode_wah_encode_missing(ctx, tmp_path):
    """--wah-encode-missing on the writer side (xsi_encode_desc.wah_encode_missing / Compressor(wah_encode_missing=True)):
    byte-exact against the oracle, which is pinned to the reference CLI for this mode; fixtures and synthetic rows."""
    import xsqueezeit_b200 as xb
    for name in SMALL:
        d = np.load(os.path.join(G, name + ".npz"))
        ns, nal, ngt, gt = int(d["n_samples"]), d["n_allele"], d["ngt"], d["gt"]
        names = [str(x) for x in d["names"]]
        p = str(tmp_path / (name + "_wm.xsi"))
        xb.Compressor(ctx, maf=float(d["maf"]), reset_sort_block_length=int(d["block_len"]), wah_encode_missing=True).compress_to_file(
            p, gt, ngt, nal, ns, sample_names=names)
        gold = open(os.path.join(G, name + "_wah_missing.xsi"), "rb").read()
        assert open(p, "rb").read() == gold, name
    ds = synth.make_dataset(500, 520, seed=44, max_alt=3, multi_frac=0.1, missing=0.01, unphased=0.02, haploid_samples=0.3)
    gt, ngt, nal, ns = ds["gt"], ds["ngt"], ds["n_allele"], ds["n_samples"]
    off = xo.row_offsets(ngt)
    img = xo.encode(gt, off, ngt, nal, ns, 128, xo.mac_threshold(ns, 2, 0.01), xo.default_phased(gt, off, ngt, ns), wah_encode_missing=True)
    p = str(tmp_path / "wm_synth.xsi")
    xb.Compressor(ctx, maf=0.01, reset_sort_block_length=128, wah_encode_missing=True).compress_to_file(p, gt, ngt, nal, ns)
    assert open(p, "rb").read() == img
    check_decode(ctx, p, img, nal, 128)


def test_sample_subset(ctx, tmp_path):
    """xsi_decode_records_subset against fill_selected_genotypes restated on oracle-decoded rows: arbitrary order, repeats,
    mixed ploidy (end-of-vector), missing, multi-allelic; all-haploid records (ploidy 1 rows)."""
    import xsqueezeit_b200 as xb
    rng = np.random.default_rng(5)
    al = (rng.random((150, 90)) < rng.uniform(0.0, 0.6, size=(150, 1))).astype(np.int8)
    haploid = dict(gt=np.ascontiguousarray(synth.encode_gt(al, 0).reshape(-1)), ngt=np.full(150, 90, np.int32),
                   n_allele=np.full(150, 2, np.int32), n_samples=90)
    for ds, bl in ((synth.make_dataset(300, 257, seed=51, max_alt=4, multi_frac=0.3, missing=0.02, unphased=0.02, haploid_samples=0.4), 128),
                   (synth.make_dataset(200, 1000, seed=52), 64), (haploid, 40)):
        ns = ds["n_samples"]
        p = gpu_encode(ctx, tmp_path, ds, bl, 0.01)
        img = open(p, "rb").read()
        acc = xb.Accessor(p, ctx)
        rd = xo.Reader(img)
        pos = xb.bm_positions(ds["n_allele"], bl)
        sel = rng.permutation(ns)[: max(3, ns // 3)].astype(np.uint32)
        sel[1] = sel[0]  # a repeated sample is allowed
        nb = (len(pos) + bl - 1) // bl
        acc._load(0, nb)
        blk = (pos >> np.uint64(15)).astype(np.uint32)
        off = (pos & np.uint64(0x7FFF)).astype(np.uint32)
        out, filled, ac = ctx.decode_records_subset(blk, off, ds["n_allele"], sel)
        for r in range(len(pos)):
            row, n = rd.fill_genotype_array(int(ds["n_allele"][r]), int(pos[r]))
            want, want_ac = xo.select_samples(row, n, ns, sel, int(ds["n_allele"][r]))
            assert filled[r] == want.size and np.array_equal(out[r, :filled[r]], want), r
            assert np.array_equal(ac[r, :int(ds["n_allele"][r]) - 1], want_ac), r
        acc.close()


# ---- hygiene: orders, shared contexts, malformed input, the reference's own limits ----------------
import ctypes  # noqa: E402


def ctypes_block(acc, b):
    p, s = ctypes.c_void_p(), ctypes.c_uint64()
    assert acc._L.xsi_reader_gt_block(acc.r, b, ctypes.byref(p), ctypes.byref(s)) == 0
    return p.value, s.value


def test_device_resident_reencode_is_idempotent(ctx, tmp_path):
    """decode every record of a file into DEVICE rows (raw BCF int8 and int32, rows a stride apart) and encode those rows again with
    xsi_encode_launch_strided: the GT blocks are the file's own blocks, byte for byte (mixed ploidy, missing, multi-allelic)"""
    import xsqueezeit_b200 as xb
    ds = synth.make_dataset(330, 413, seed=71, max_alt=4, multi_frac=0.25, missing=0.02, unphased=0.03, haploid_samples=0.3)
    bl, maf = 100, 0.01
    ns, nal = ds["n_samples"], ds["n_allele"]
    p = gpu_encode(ctx, tmp_path, ds, bl, maf)
    acc = xb.Accessor(p, ctx)
    pos = xb.bm_positions(nal, bl)
    R = len(pos)
    nb = (R + bl - 1) // bl
    acc._load(0, nb)
    blk = (pos >> np.uint64(15)).astype(np.uint32)
    off = (pos & np.uint64(0x7FFF)).astype(np.uint32)
    na = np.ascontiguousarray(nal, dtype=np.uint32)
    o = xo.row_offsets(ds["ngt"])
    dp = xo.default_phased(ds["gt"], o, ds["ngt"], ns)
    thr = xo.mac_threshold(ns, int(ds["ngt"][0]) // ns, maf)
    want = [ctypes.string_at(*ctypes_block(acc, b)) for b in range(nb)]
    for elem, stride in ((4, 2 * ns + 6), (1, (2 * ns + 15) // 16 * 16)):
        dev = ctx.device_alloc(R * stride * elem)
        try:
            filled = np.zeros(R, dtype=np.uint32)
            fn = ctx._L.xsi_decode_records if elem == 4 else ctx._L.xsi_decode_records_i8
            ctx._check(fn(ctx.h, R, blk.ctypes.data, off.ctypes.data, na.ctypes.data, ctypes.c_void_p(dev), stride, 1, filled.ctypes.data, None, 0))
            assert np.array_equal(filled, ds["ngt"].astype(np.uint32))
            ctx.encode_launch(dev, nal, ns, bl, thr, dp, ploidy=(ds["ngt"] // ns).astype(np.uint8), gt_elem_bytes=elem, gt_on_device=True, row_stride=stride)
            blocks = ctx.encode_collect()
        finally:
            ctx.device_free(dev)
        assert len(blocks) == nb
        for b, got in enumerate(blocks):
            assert 0 <= len(want[b]) - len(got) < 16 and got == want[b][:len(got)] and not any(want[b][len(got):]), (elem, b)
    acc.close()


def test_subset_reencode_on_device(ctx, tmp_path):
    """The extractor's XSI -> XSI path with a sample subset (-s ... -Ox, gt_decompressor_new.hpp:241-273) without the rows leaving the
    device: xsi_decode_records_subset into device rows, xsi_encode_launch_strided from them.  The GT blocks equal the oracle's
    encoding of the rows fill_selected_genotypes would have handed to XsiFactoryExt::append (mixed ploidy, all-haploid records)."""
    import xsqueezeit_b200 as xb
    rng = np.random.default_rng(15)
    al = (rng.random((150, 90)) < rng.uniform(0.0, 0.6, size=(150, 1))).astype(np.int8)
    haploid = dict(gt=np.ascontiguousarray(synth.encode_gt(al, 0).reshape(-1)), ngt=np.full(150, 90, np.int32),
                   n_allele=np.full(150, 2, np.int32), n_samples=90)
    for k, (ds, bl) in enumerate(((synth.make_dataset(300, 257, seed=61, max_alt=4, multi_frac=0.3, missing=0.02, unphased=0.02, haploid_samples=0.4), 128),
                                  (synth.make_dataset(200, 1000, seed=62), 64), (haploid, 40))):
        ns, nal = ds["n_samples"], ds["n_allele"]
        p = gpu_encode(ctx, tmp_path, ds, bl, 0.01)
        acc = xb.Accessor(p, ctx)
        rd = xo.Reader(open(p, "rb").read())
        pos = xb.bm_positions(nal, bl)
        sel = rng.permutation(ns)[: max(5, ns // 3)].astype(np.uint32)
        R, n_sel = len(pos), len(sel)
        acc._load(0, (R + bl - 1) // bl)
        blk = (pos >> np.uint64(15)).astype(np.uint32)
        off = (pos & np.uint64(0x7FFF)).astype(np.uint32)
        dev = ctx.device_alloc(R * 2 * n_sel * 4)
        try:
            _, filled, _ = ctx.decode_records_subset(blk, off, nal, sel, out_device=dev)
            rows = []
            for r in range(R):
                row, n = rd.fill_genotype_array(int(nal[r]), int(pos[r]))
                rows.append(xo.select_samples(row, n, ns, sel, int(nal[r]))[0])
            ngt = np.array([x.size for x in rows], dtype=np.int32)
            assert np.array_equal(filled, ngt)
            gt = np.ascontiguousarray(np.concatenate(rows), dtype=np.int32)
            o = xo.row_offsets(ngt)
            dp = xo.default_phased(gt, o, ngt, n_sel)
            thr = xo.mac_threshold(n_sel, int(ngt[0]) // n_sel, 0.01)
            want = str(tmp_path / ("want%d.xsi" % k))
            open(want, "wb").write(xo.encode(gt, o, ngt, nal, n_sel, bl, thr, dp, None))
            ctx.encode_launch(dev, nal, n_sel, bl, thr, dp, ploidy=(ngt // n_sel).astype(np.uint8), gt_on_device=True, row_stride=2 * n_sel)
            blocks = ctx.encode_collect()
        finally:
            ctx.device_free(dev)
        acc.close()
        ref = xb.Accessor(want, ctx)
        assert len(blocks) == ref.n_blocks
        for b, got in enumerate(blocks):
            a, n = ctypes_block(ref, b)
            want_b = ctypes.string_at(a, n)  # the reader's view runs to the next block / the index: file padding (zeros) follows
            assert 0 <= n - len(got) < 16 and got == want_b[:len(got)] and not any(want_b[len(got):]), (k, b)
        ref.close()
        with pytest.raises(xb.XsiError):  # strided rows are device rows
            ctx.encode_launch(gt, nal, n_sel, bl, thr, dp, ploidy=(ngt // n_sel).astype(np.uint8), row_stride=2 * n_sel)


def test_positions_in_any_order(ctx, tmp_path):
    """seek (accessor_internals_new.hpp:154-196) replays or resets the cursor; here any (block, line) is addressable:
    shuffled, backward and repeated positions through fill_genotype_arrays, the one-record call and the subset call"""
    import xsqueezeit_b200 as xb
    ds = synth.make_dataset(700, 301, seed=61, max_alt=3, multi_frac=0.2, missing=0.01, unphased=0.02, haploid_samples=0.3)
    bl = 128
    p = gpu_encode(ctx, tmp_path, ds, bl, 0.01)
    img = open(p, "rb").read()
    nal, ns = ds["n_allele"], ds["n_samples"]
    pos = xb.bm_positions(nal, bl)
    rd = xo.Reader(img)
    want = [rd.fill_genotype_array(int(nal[r]), int(pos[r])) for r in range(len(nal))]
    want = [(row[:n].copy(), n) for row, n in want]
    rng = np.random.default_rng(3)
    orders = {"backward": np.arange(len(nal))[::-1], "shuffled": rng.permutation(len(nal)),
              "repeated": rng.integers(0, len(nal), size=400), "ping_pong": np.array([0, len(nal) - 1] * 20 + [len(nal) // 2] * 3)}
    for blocks_resident in (1, 3):
        acc = xb.Accessor(p, ctx, blocks_resident=blocks_resident)
        for name, order in orders.items():
            out, filled, counts = acc.fill_genotype_arrays(nal[order], pos[order], want_counts=True)
            for i, r in enumerate(order):
                assert filled[i] == want[r][1] and np.array_equal(out[i, :filled[i]], want[r][0]), (name, i, r)
            for r in order[:25]:  # one record per call, the reference's calling pattern
                a, n = acc.fill_genotype_array(int(nal[r]), int(pos[r]))
                assert n == want[r][1] and np.array_equal(a[:n], want[r][0]), (name, r)
        acc.close()
    # subset call in shuffled order
    acc = xb.Accessor(p, ctx)
    nb = (len(pos) + bl - 1) // bl
    acc._load(0, nb)
    order = orders["shuffled"]
    sel = rng.permutation(ns)[:40].astype(np.uint32)
    blk = (pos[order] >> np.uint64(15)).astype(np.uint32)
    off = (pos[order] & np.uint64(0x7FFF)).astype(np.uint32)
    out, filled, ac = ctx.decode_records_subset(blk, off, nal[order], sel)
    for i, r in enumerate(order):
        w, w_ac = xo.select_samples(want[r][0], want[r][1], ns, sel, int(nal[r]))
        assert filled[i] == w.size and np.array_equal(out[i, :filled[i]], w), (i, r)
        assert np.array_equal(ac[i, :int(nal[r]) - 1], w_ac), (i, r)
    acc.close()


def test_two_accessors_share_a_context(ctx, tmp_path):
    """the loaded block set belongs to the context: an Accessor must notice that another one replaced it"""
    import xsqueezeit_b200 as xb
    d1 = synth.make_dataset(300, 200, seed=71)
    d2 = synth.make_dataset(300, 200, seed=72, missing=0.01)
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    p1 = gpu_encode(ctx, tmp_path / "a", d1, 128, 0.01)
    p2 = gpu_encode(ctx, tmp_path / "b", d2, 128, 0.01)
    a1, a2 = xb.Accessor(p1, ctx), xb.Accessor(p2, ctx)
    r1, r2 = xo.Reader(open(p1, "rb").read()), xo.Reader(open(p2, "rb").read())
    pos = xb.bm_positions(d1["n_allele"], 128)
    for r in (5, 200, 6, 201, 130, 7):
        for acc, rd in ((a1, r1), (a2, r2), (a1, r1)):
            a, n = acc.fill_genotype_array(2, int(pos[r]))
            b, nb = rd.fill_genotype_array(2, int(pos[r]))
            assert n == nb and np.array_equal(a[:n], b[:nb]), r
    a1.close()
    a2.close()


def test_malformed_blocks_are_refused(ctx, tmp_path):
    """truncated and corrupted GT blocks end in XSI_E_FORMAT (-6) / XSI_E_UNSUPPORTED (-5), never in a device fault:
    the context must still work afterwards"""
    import xsqueezeit_b200 as xb
    ds = synth.make_dataset(300, 200, seed=81, max_alt=3, multi_frac=0.2, missing=0.02, haploid_samples=0.3)
    p = gpu_encode(ctx, tmp_path, ds, 128, 0.3)  # maf 0.3: most lines are sparse index lists
    acc = xb.Accessor(p, ctx)
    ptr, size = ctypes_block(acc, 0)
    good = bytes((ctypes.c_uint8 * size).from_address(ptr))
    acc.close()
    rng = np.random.default_rng(9)
    refused = 0
    trials = [good[:n] for n in (4, 12, 40, size // 3, size // 2, size - 7, size - 1)]
    for _ in range(40):  # overwrite a few 16-bit words inside the sections with large counts / garbage
        b = bytearray(good)
        for _k in range(3):
            at = int(rng.integers(100, size - 2)) & ~1
            b[at:at + 2] = int(rng.choice([0x7FFF, 0xFFFF, 0x3FFF, int(rng.integers(0, 65536))])).to_bytes(2, "little")
        trials.append(bytes(b))
    for t in trials:
        try:
            ctx.decode_load_blocks([t], ds["n_samples"], 2)
            # a corruption may leave a well-formed block: then every line must still decode without a fault
            n = min(20, len(ds["n_allele"]))
            ctx.decode_records(np.zeros(n, np.uint32), np.arange(n, dtype=np.uint32), np.full(n, 2, np.uint32))
        except xb.XsiError as e:
            assert e.code in (-6, -5, -2), e.code
            refused += 1
    assert refused >= 7
    roundtrip(ctx, tmp_path, synth.make_dataset(100, 64, seed=82), 50, 0.01)  # the context survived


def test_record_without_alt_is_refused(ctx):
    import xsqueezeit_b200 as xb
    gt = synth.encode_gt(np.zeros((3, 20), np.int8)).reshape(-1).copy()
    for bad in (1, 255):
        with pytest.raises(xb.XsiError) as e:
            ctx.encode_launch(gt, [2, bad, 2], 10, 8192, 0, 1)
        assert e.value.code == -5


def test_haploid_and_multiallelic_block(ctx, tmp_path):
    """A block holding an all-haploid record and a multi-allelic record: written byte for byte like the reference (the oracle
    and the live reference agree on the bytes, tests/test_oracle_vs_reference.py), but the reference's reader corrupts its heap
    on it (LINE_HAPLOID per BCF line vs per binary line) -- the decode call answers XSI_E_UNSUPPORTED."""
    import xsqueezeit_b200 as xb
    ds = synth.haploid_multiallelic_block()
    p = gpu_encode(ctx, tmp_path, ds, 64, 0.01)
    assert open(p, "rb").read() == oracle_image(ds, 64, 0.01)
    acc = xb.Accessor(p, ctx)
    with pytest.raises(xb.XsiError) as e:
        acc.fill_genotype_array(2, 0)
    assert e.value.code == -5
    acc.close()


def test_async_encode_beside_decode(ctx, tmp_path):
    """xsi_encode_async: a batch encodes on the library's thread and stream while the caller decodes the previous one on the
    same context; blocks of a collect stay valid while the next launch runs.  Bytes and rows equal the synchronous path."""
    import torch
    import xsqueezeit_b200 as xb
    dsA = synth.make_dataset(900, 1500, seed=91, max_alt=2, multi_frac=0.1, missing=0.003)
    dsB = synth.make_dataset(900, 1500, seed=92)
    bl, ns = 128, 1500
    thr = xo.mac_threshold(ns, 2, 0.01)

    def oracle_blocks(ds):
        img = oracle_image(ds, bl, 0.01)
        rd = xo.Reader(img)
        pos = xb.bm_positions(ds["n_allele"], bl)
        rows = [rd.fill_genotype_array(int(ds["n_allele"][r]), int(pos[r])) for r in range(len(pos))]
        return pos, [(row[:n].copy(), n) for row, n in rows]

    dev = {k: torch.as_tensor(d["gt"], device="cuda") for k, d in (("A", dsA), ("B", dsB))}
    sync_blocks = {}
    for k, d in (("A", dsA), ("B", dsB)):
        ctx.encode_launch(dev[k].data_ptr(), d["n_allele"], ns, bl, thr, 1, gt_on_device=True)
        sync_blocks[k] = ctx.encode_collect()
    want = {k: oracle_blocks(d) for k, d in (("A", dsA), ("B", dsB))}
    ctx.encode_async(True)
    try:
        ctx.encode_launch(dev["A"].data_ptr(), dsA["n_allele"], ns, bl, thr, 1, gt_on_device=True)
        prev, prev_k = ctx.encode_collect(), "A"
        for it in range(6):
            k = "B" if it % 2 == 0 else "A"
            d = dsB if k == "B" else dsA
            ctx.encode_launch(dev[k].data_ptr(), d["n_allele"], ns, bl, thr, 1, gt_on_device=True)  # returns at once
            assert prev == sync_blocks[prev_k]
            dprev = dsA if prev_k == "A" else dsB
            ctx.decode_load_blocks(prev, ns, 2)
            pos, rows = want[prev_k]
            out, filled, _ = ctx.decode_records((pos >> np.uint64(15)).astype(np.uint32), (pos & np.uint64(0x7FFF)).astype(np.uint32), dprev["n_allele"])
            for r in range(0, len(pos), 7):
                assert filled[r] == rows[r][1] and np.array_equal(out[r, :filled[r]], rows[r][0]), (it, r)
            prev, prev_k = ctx.encode_collect(), k
            assert prev == sync_blocks[k]
        # an error inside an asynchronous launch is reported by the collect
        bad = dsA["gt"].copy()
        bad[5] = synth.encode_gt(np.array([7], np.int8))[0]
        tb = torch.as_tensor(bad, device="cuda")
        ctx.encode_launch(tb.data_ptr(), np.full(len(dsA["n_allele"]), 2, np.int32), ns, bl, thr, 1, gt_on_device=True)
        with pytest.raises(xb.XsiError) as e:
            ctx.encode_collect()
        assert e.value.code == -3
    finally:
        ctx.encode_async(False)
    roundtrip(ctx, tmp_path, synth.make_dataset(50, 40, seed=93), 16, 0.01)  # back in the synchronous mode


@pytest.mark.parametrize("env", [{"XSI_PBWT_V": "5"}, {"XSI_PBWT_V": "4"}, {"XSI_PBWT_V": "1"},
                                 {"XSI_PBWT_V": "5", "XSI_PBWT_CLUSTER": "2"}, {"XSI_PBWT_V": "5", "XSI_PBWT_CLUSTER": "1", "XSI_PBWT_KH": "16"},
                                 {"XSI_PBWT_V": "5", "XSI_PBWT_CLUSTER": "8"}, {"XSI_PBWT_V": "4", "XSI_PBWT_CLUSTER": "2"},
                                 {"XSI_SCAN_V1": "1"}, {"XSI_SCAN_NT": "256", "XSI_COMPOSE_NT": "256"}, {"XSI_COMPOSE_V1": "1"},
                                 {"XSI_UNPERM_FENCE": "1"}, {"XSI_UNPERM_KH": "32", "XSI_UNPERM_NC": "160"}, {"XSI_UNPERM_KH": "8", "XSI_UNPERM_NC": "512"},
                                 {"XSI_UNPERM_KH": "16", "XSI_UNPERM_NC": "320"}])
def test_kernel_variants_are_byte_exact(ctx, tmp_path, monkeypatch, env):
    """every kernel variant that ships (the two-line and the one-line cluster kernels at several cluster sizes, the general
    shared-memory kernel, the ballot scan, the fixed-width CTAs) against the oracle: odd and even numbers of WAH lines per block,
    rows of several widths, blocks that end on a sparse line"""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    cases = [(synth.make_dataset(333, 2504, seed=101), 111, 0.001),            # 1KGP3 width, 3 blocks of 111 records
             (synth.make_dataset(150, 4096, seed=105), 50, 0.001),             # 8192 haplotypes: the widest row of the short-row decode split
             (synth.make_dataset(150, 4097, seed=106), 75, 0.001),             # ... and the first one past it
             (synth.make_dataset(90, 20000, seed=102, n_founders=32), 45, 0.001),  # 40,000 haplotypes
             (synth.make_dataset(64, 32488, seed=103, n_founders=32), 32, 0.001),  # HRC width
             (synth.make_dataset(41, 700, seed=104, max_alt=3, multi_frac=0.3, missing=0.01), 41, 0.0)]  # every line WAH
    for ds, bl, maf in cases:
        roundtrip(ctx, tmp_path, ds, bl, maf)


def test_haploid_records_in_some_blocks_only(ctx, tmp_path):
    """one batch, five blocks: all-haploid records only in blocks 1 and 3 -- those two take the general PBWT kernel, the other
    three the cluster kernel, in the same launch (the haploid flag demotes a block, not the batch)"""
    rng = np.random.default_rng(17)
    ns, bl = 1300, 60
    rows, ngt = [], []
    for r in range(5 * bl - 7):
        blk = r // bl
        p = 1 if (blk in (1, 3) and r % 9 == 2) else 2
        al = (rng.random(ns * p) < rng.uniform(0.02, 0.5)).astype(np.int8)
        rows.append(synth.encode_gt(al, 1 if p == 2 else 0))
        ngt.append(ns * p)
    ds = dict(gt=np.concatenate(rows).astype(np.int32), ngt=np.array(ngt, np.int32), n_allele=np.full(len(ngt), 2, np.int32), n_samples=ns)
    roundtrip(ctx, tmp_path, ds, bl, 0.01, blocks_per_batch=8)
    roundtrip(ctx, tmp_path, ds, bl, 0.01, blocks_per_batch=2)


def test_lazy_chain(ctx, tmp_path):
    """xsi_decode_load_blocks_lazy: the inverse-PBWT chain stops early and continues on demand; records before, at and after the
    frontier, in any order, equal the oracle; allele counts need no chain at all"""
    import xsqueezeit_b200 as xb
    ds = synth.make_dataset(1500, 3000, seed=111, max_alt=2, multi_frac=0.1, missing=0.002)
    bl = 700
    p = gpu_encode(ctx, tmp_path, ds, bl, 0.002)
    img = open(p, "rb").read()
    nal = ds["n_allele"]
    pos = xb.bm_positions(nal, bl)
    rd = xo.Reader(img)
    want = [rd.fill_genotype_array(int(nal[r]), int(pos[r])) for r in range(len(nal))]
    want = [(row[:n].copy(), n) for row, n in want]
    acc = xb.Accessor(p, ctx)
    blocks = [ctypes_block(acc, b) for b in range(acc.n_blocks)]
    os.environ["XSI_LAZY_WINDOW"] = "64"
    try:
        ctx.decode_load_blocks(blocks, ds["n_samples"], 2, lazy_lines=0)
        assert ctx.decode_lines_ready(0) < 50 and ctx.decode_lines_ready(1) < 50  # nothing un-permuted yet (sparse lines are always final)
        blk = (pos >> np.uint64(15)).astype(np.uint32)
        off = (pos & np.uint64(0x7FFF)).astype(np.uint32)
        # counts without any chain
        rd2 = xo.Reader(img)
        sel = np.arange(0, len(nal), 37)
        ac = ctx.decode_allele_counts(blk[sel], off[sel], nal[sel])
        for i, r in enumerate(sel):
            assert np.array_equal(ac[i, :int(nal[r])], rd2.fill_allele_counts(int(nal[r]), int(pos[r]))), r
        assert ctx.decode_lines_ready(0) < 50
        rng = np.random.default_rng(5)
        for r in [300, 301, 10, 650, 305, 1400, 1399, 699, 700, 0, len(nal) - 1]:
            out, filled, _ = ctx.decode_records(blk[r:r + 1], off[r:r + 1], nal[r:r + 1])
            assert filled[0] == want[r][1] and np.array_equal(out[0, :filled[0]], want[r][0]), r
            assert ctx.decode_lines_ready(int(blk[r])) > off[r]
        assert ctx.decode_lines_ready(0) >= off[650]
        order = rng.permutation(len(nal))
        out, filled, _ = ctx.decode_records(blk[order], off[order], nal[order])
        for i, r in enumerate(order):
            assert filled[i] == want[r][1] and np.array_equal(out[i, :filled[i]], want[r][0]), r
        # explicit extension, then a full lazy load (initial_lines = everything) equals the eager one
        ctx.decode_load_blocks(blocks, ds["n_samples"], 2, lazy_lines=0)
        ctx.decode_extend(1, 400)
        assert ctx.decode_lines_ready(1) >= 400 and ctx.decode_lines_ready(0) < 50
    finally:
        del os.environ["XSI_LAZY_WINDOW"]
    acc.close()


def test_dot_products_on_encoded_lines(ctx, tmp_path):
    """xsi_decode_dot_products (the reference's dot_prod consumer of InternalGtAccess) against a float64 sum over the oracle's
    decoded rows: WAH and sparse lines, multi-allelic records, missing, haploid samples, all-haploid records and negated sparse
    lines (composed-row path).  Floating point: different summation order, tolerance 1e-10 relative (stated in include/xsi_b200.h)."""
    import xsqueezeit_b200 as xb
    rng = np.random.default_rng(23)
    neg = synth.make_dataset(120, 100, seed=121)
    g = neg["gt"].reshape(120, 200)
    g[5, :] = synth.encode_gt(np.ones(200, np.int8))
    g[6, :] = synth.encode_gt(np.ones(200, np.int8))
    g[6, 17] = synth.encode_gt(np.zeros(1, np.int8))[0]
    g[6, 40] = 0  # a missing entry on a negated line
    neg["gt"] = np.ascontiguousarray(g.reshape(-1))
    al = (rng.random((150, 90)) < rng.uniform(0.0, 0.6, size=(150, 1))).astype(np.int8)
    haploid = dict(gt=np.ascontiguousarray(synth.encode_gt(al, 0).reshape(-1)), ngt=np.full(150, 90, np.int32),
                   n_allele=np.full(150, 2, np.int32), n_samples=90)
    for ds, bl, maf in ((synth.make_dataset(400, 1500, seed=122, max_alt=3, multi_frac=0.3, missing=0.01, haploid_samples=0.3), 128, 0.01),
                        (neg, 64, 0.05), (haploid, 40, 0.05), (synth.make_dataset(64, 32488, seed=123, n_founders=32), 32, 0.001)):
        ns = ds["n_samples"]
        p = gpu_encode(ctx, tmp_path, ds, bl, maf)
        img = open(p, "rb").read()
        acc = xb.Accessor(p, ctx)
        nal = ds["n_allele"]
        pos = xb.bm_positions(nal, bl)
        y = rng.normal(0, 10, ns)
        rd = xo.Reader(img)
        nb = (len(pos) + bl - 1) // bl
        blocks = [ctypes_block(acc, b) for b in range(nb)]
        ctx.decode_load_blocks(blocks, ns, 2 if 2 * ns <= 65535 else 4, lazy_lines=0)
        blk = (pos >> np.uint64(15)).astype(np.uint32)
        off = (pos & np.uint64(0x7FFF)).astype(np.uint32)
        got = ctx.decode_dot_products(blk, off, nal, y)
        for r in range(len(nal)):
            row, n = rd.fill_genotype_array(int(nal[r]), int(pos[r]))
            row = row[:n].astype(np.int64)
            ys = y if n == ns else np.repeat(y, 2)
            for a in range(1, int(nal[r])):
                want = float(ys[(row >> 1) - 1 == a].sum()) if True else 0.0
                # vector end (INT32_MIN + 1) and bcf_int32_missing shift to large negatives: never equal to an allele
                assert abs(got[r, a - 1] - want) <= 1e-10 * max(1.0, abs(want)), (r, a, got[r, a - 1], want)
        acc.close()
