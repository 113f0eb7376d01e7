"""CPU tests: pin the plain-C oracle (oracle/xsi_oracle.c) to the reference's own outputs.

tests/golden/ was produced by tests/golden/make_golden.py by running the UNMODIFIED reference
(oracle/_ref) on /root/reference/test/test_files; SHA-256 values equal SURVEY.md section 8(c)."""
import hashlib
import json
import os

import numpy as np
import pytest

import xsi_oracle as xo

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MAN = json.load(open(os.path.join(G, "manifest.json")))
SMALL = sorted(k for k in MAN if k != "chr20_small")

# inner-dictionary key order KAT, SURVEY.md section 8(c) (hex keys as emitted by libstdc++ 13)
BASE = [0, 1, 2, 3, 4, 0x10, 0x11, 0x20, 0x21]
MISS, EOVK, PH, HAP = [0x16, 0x26, 0x36], [0x18, 0x28, 0x38], [0x17, 0x27], [0x12]
KAT = [
    (BASE, "21 20 11 4 10 3 2 1 0"),
    (BASE + HAP, "12 21 20 11 4 10 3 2 1 0"),
    (BASE + PH, "17 21 20 11 4 10 3 2 1 27 0"),
    (BASE + MISS, "26 16 21 20 11 4 10 3 36 2 1 0"),
    (BASE + EOVK, "18 21 20 38 11 4 10 3 2 28 1 0"),
    (BASE + MISS + PH, "27 0 1 2 36 10 11 3 20 4 21 16 26 17"),
    (BASE + MISS + EOVK + PH, "27 17 38 28 0 1 2 36 10 11 3 20 4 21 16 26 18"),
    (BASE + MISS + EOVK + PH + HAP, "12 27 17 38 28 0 1 2 36 10 11 3 20 4 21 16 26 18"),
]


@pytest.mark.parametrize("keys,expect", KAT)
def test_unordered_map_order_kat(keys, expect):
    got = " ".join("%x" % k for k in xo.unordered_order(keys))
    assert got == expect


def test_wah_rules():
    # literal, zero run, one run, tail padding (SURVEY 8(c) normative rules)
    bits = np.zeros(15 * 5 + 3, dtype=np.uint8)
    bits[0] = 1                      # group0 literal 0x0001
    bits[30:45] = 1                  # group2 all ones
    bits[75:78] = 1                  # tail group literal 0b111
    w = xo.wah_encode_bits(bits)
    assert list(w) == [0x0001, 0x8001, 0xC001, 0x8002, 0x0007]
    back, used, ones = xo.wah_decode_bits(w, bits.size)
    assert used == len(w) and np.array_equal(back, bits) and ones == 1 + 15 + 3


def test_wah_counter_saturation():
    n = 15 * (16383 * 2 + 5)
    for val, full, tag in ((0, 0xBFFF, 0x8000), (1, 0xFFFF, 0xC000)):
        bits = np.full(n, val, dtype=np.uint8)
        w = xo.wah_encode_bits(bits)
        assert list(w) == [full, full, tag | 5]
        back, used, ones = xo.wah_decode_bits(w, n)
        assert np.array_equal(back, bits) and ones == val * n


@pytest.mark.parametrize("name", SMALL)
def test_small_fixture_bytes_and_decode(name):
    d = np.load(os.path.join(G, name + ".npz"))
    ns, nal, ngt, gt = int(d["n_samples"]), d["n_allele"], d["ngt"], d["gt"]
    names = [str(x) for x in d["names"]]
    off = xo.row_offsets(ngt)
    dp = xo.default_phased(gt, off, ngt, ns)
    thr = xo.mac_threshold(ns, int(ngt[0]) // ns, float(d["maf"]))
    img = xo.encode(gt, off, ngt, nal, ns, int(d["block_len"]), thr, dp, names)
    gold = open(os.path.join(G, name + ".xsi"), "rb").read()
    assert hashlib.sha256(gold).hexdigest() == MAN[name]["xsi_sha256"]
    assert img == gold
    r = xo.Reader(gold)
    pos = xo.bm_positions(nal, int(d["block_len"]))
    rows = [r.fill_genotype_array(int(nal[i]), int(pos[i])) for i in range(len(nal))]
    dec = np.concatenate([o[:n] for o, n in rows])
    assert np.array_equal(dec, d["ref_decoded"])          # == reference Accessor output
    assert [n for _, n in rows] == list(d["ref_decoded_ngt"])


@pytest.mark.parametrize("name", SMALL)
def test_small_fixture_wah_encode_missing(name):
    """--wah-encode-missing (WS_WAH): oracle bytes equal the reference CLI's, oracle decode of it equals the default decode."""
    d = np.load(os.path.join(G, name + ".npz"))
    ns, nal, ngt, gt = int(d["n_samples"]), d["n_allele"], d["ngt"], d["gt"]
    names = [str(x) for x in d["names"]]
    off = xo.row_offsets(ngt)
    dp = xo.default_phased(gt, off, ngt, ns)
    thr = xo.mac_threshold(ns, int(ngt[0]) // ns, float(d["maf"]))
    img = xo.encode(gt, off, ngt, nal, ns, int(d["block_len"]), thr, dp, names, wah_encode_missing=True)
    gold = open(os.path.join(G, name + "_wah_missing.xsi"), "rb").read()
    assert hashlib.sha256(gold).hexdigest() == MAN[name]["wah_missing"]["xsi_sha256"]
    assert img == gold
    assert MAN[name]["wah_missing"]["ref_decode_equals_default"]
    r = xo.Reader(gold)
    pos = xo.bm_positions(nal, int(d["block_len"]))
    rows = [r.fill_genotype_array(int(nal[i]), int(pos[i])) for i in range(len(nal))]
    dec = np.concatenate([o[:n] for o, n in rows])
    assert np.array_equal(dec, d["ref_decoded"])


def test_chr20_small_decode_then_all_option_hashes():
    man = MAN["chr20_small"]
    d = np.load(os.path.join(G, "chr20_small_meta.npz"))
    ns, nal, ngt = int(d["n_samples"]), d["n_allele"].astype(np.int32), d["ngt"]
    names = [str(x) for x in d["names"]]
    gold = open(os.path.join(G, "chr20_small_default.xsi"), "rb").read()
    assert hashlib.sha256(gold).hexdigest() == man["options"]["default"]["xsi_sha256"]
    r = xo.Reader(gold)
    pos = xo.bm_positions(nal, 8192)
    off = xo.row_offsets(ngt)
    gt = np.empty(int(ngt.sum()), np.int32)
    for i in range(len(nal)):
        o = int(off[i])
        _, n = r.fill_genotype_array(int(nal[i]), int(pos[i]), gt[o:o + int(ngt[i])])
        assert n == ngt[i]
    # equals bcf_get_genotypes on the original file AND the reference Accessor's decode
    assert hashlib.sha256(gt.tobytes()).hexdigest() == man["gt_sha256"] == man["ref_decode_sha256"]
    dp = xo.default_phased(gt, off, ngt, ns)
    for key, o in man["options"].items():
        argv = o["argv"]
        maf = float(argv[argv.index("--maf") + 1]) if "--maf" in argv else 0.001
        bl = int(argv[argv.index("--variant-block-length") + 1]) if "--variant-block-length" in argv else 8192
        img = xo.encode(gt, off, ngt, nal, ns, bl, xo.mac_threshold(ns, 2, maf), dp, names)
        assert len(img) == o["xsi_size"] and hashlib.sha256(img).hexdigest() == o["xsi_sha256"], key
