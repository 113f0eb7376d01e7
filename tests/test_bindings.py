"""The reference's OWN programs on the B200 path (bindings/): `xsqueezeit -c / -x`, Accessor, c_api.h and the
reference's lockstep_loader, built from the unmodified reference sources with the two adapters of bindings/ injected
(bindings/Makefile).  Golden values come from the unmodified CPU reference (tests/golden/make_cli_golden.py,
tests/golden/make_golden.py); the cases are the reference's own test list (test/cukinia_v4.conf:4-20) plus the 14
compress runs of SURVEY.md 8(c).

-m gpu: the parity tests proper.  Without a GPU: the binaries exist, link, and refuse to work (no CPU fallback)."""
import hashlib
import json
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
G = os.path.join(HERE, "golden")
INP = os.path.join(G, "inputs")
OUT = os.path.join(ROOT, "bindings", "_out")
CLI = os.path.join(OUT, "xsqueezeit_b200")
LOCKSTEP = os.path.join(OUT, "lockstep_loader_b200")
CAPI = os.path.join(OUT, "capi_decode_b200")
CLIMAN = json.load(open(os.path.join(G, "cli_manifest.json")))
MAN = json.load(open(os.path.join(G, "manifest.json")))

needs_bindings = pytest.mark.skipif(not os.path.exists(CLI), reason="bindings/_out not built (make -C bindings needs /root/reference)")


def sha(b):
    return hashlib.sha256(b).hexdigest()


def body_sha(text):
    return sha(b"".join(l for l in text.splitlines(True) if not l.startswith(b"##")))


def run(argv, check=True, timeout=600):
    p = subprocess.run(argv, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout)
    if check and p.returncode != 0:
        raise AssertionError("%s failed (%d):\n%s\n%s" % (" ".join(argv), p.returncode, p.stdout.decode()[-2000:], p.stderr.decode()[-2000:]))
    return p


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


# ------------------------------------------------------------------ CPU side ------------------------------------------------------------------
@needs_bindings
def test_binaries_present_and_linked():
    for b in ("xsqueezeit_b200", "lockstep_loader_b200", "c_api_test_b200", "capi_decode_b200", "libxsqueezeit_b200.so"):
        path = os.path.join(OUT, b)
        assert os.path.exists(path), b
        ldd = run(["ldd", path]).stdout.decode()
        assert "libxsi_b200.so" in ldd and "not found" not in ldd, ldd


@needs_bindings
def test_capi_library_exports_the_reference_c_api():
    """bindings/_out/libxsqueezeit_b200.so exports every function of the reference's include/c_api.h:38-93"""
    import ctypes
    lib = ctypes.CDLL(os.path.join(OUT, "libxsqueezeit_b200.so"))
    for sym in ("c_xcf_new", "c_xcf_add_readers", "c_xcf_update_readers", "c_xcf_sample_name", "c_xcf_nsamples",
                "__c__xcf__get__genotypes__void", "c_xcf_delete"):
        assert hasattr(lib, sym), sym


@needs_bindings
@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(tmp_path):
    """without a device the adapters throw (the reference CLI exits non-zero): nothing is computed on the CPU"""
    out = str(tmp_path / "x.xsi")
    p = run([CLI, "-c", "--maf", "0.002", "-f", os.path.join(INP, "micro_missing.vcf"), "-o", out], check=False)
    assert p.returncode != 0
    assert b"xsi_b200 rc -1" in p.stderr
    ref = os.path.join(G, "micro_missing.xsi")
    # decode side: needs the companion _var.bcf, which only a compress run makes -> use the reference CLI when it is here
    ref_cli = os.path.join(ROOT, "oracle", "_ref", "xsqueezeit_ref")
    if os.path.exists(ref_cli):
        run([ref_cli, "-c", "--maf", "0.002", "-f", os.path.join(INP, "micro_missing.vcf"), "-o", out])
        p = run([CAPI, out + "_var.bcf"], check=False)
        assert p.returncode != 0 and b"no CPU fallback" in p.stderr
    assert os.path.exists(ref)


@needs_bindings
@pytest.mark.skipif(not os.path.exists(os.path.join(OUT, "lockstep_loader_ref")), reason="reference lockstep_loader not built")
def test_reference_lockstep_harness(tmp_path):
    """the checker itself: the reference's lockstep_loader on the reference's own output agrees (CPU only)"""
    ref_cli = os.path.join(ROOT, "oracle", "_ref", "xsqueezeit_ref")
    if not os.path.exists(ref_cli):
        pytest.skip("reference CLI not built")
    out = str(tmp_path / "m.xsi")
    run([ref_cli, "-c", "--maf", "0.002", "-f", os.path.join(INP, "micro_mixed_ploidy.vcf"), "-o", out])
    p = run([os.path.join(OUT, "lockstep_loader_ref"), "--file1", os.path.join(INP, "micro_mixed_ploidy.bgz.vcf"), "--file2", out + "_var.bcf"])
    assert b"Files have the same GT data" in p.stderr


# ------------------------------------------------------------------ GPU side ------------------------------------------------------------------
def compress(tmp_path, name, case):
    out = str(tmp_path / (name + ".xsi"))
    run([CLI] + case["compress_argv"] + ["-f", os.path.join(INP, case["input"]), "-o", out])
    return out


@pytest.mark.gpu
@needs_bindings
@pytest.mark.parametrize("name", sorted(CLIMAN))
def test_reference_cli_on_gpu(tmp_path, name):
    """`xsqueezeit -c` then `-x` of the reference CLI with both adapters: byte-identical .xsi, identical extracted
    records for every case of test/cukinia_v4.conf, identical c_xcf_get_genotypes stream"""
    case = CLIMAN[name]
    xsi = compress(tmp_path, name, case)
    data = open(xsi, "rb").read()
    if "xsi_sha256" in case:
        assert len(data) == case["xsi_size"]
        assert sha(data) == case["xsi_sha256"], "xsqueezeit -c on the GPU path: .xsi bytes differ from the reference's"
    if "x" in case["extract_argv"]:  # -Ox: decode on the GPU, gather, re-encode on the GPU
        out = str(tmp_path / (name + "_out.xsi"))
        run([CLI, "-x"] + case["extract_argv"] + ["-f", xsi, "-o", out])
        assert sha(open(out, "rb").read()) == case["out_xsi_sha256"]
        assert body_sha(run([CLI, "-x", "-O", "v", "-f", out, "-o", "-"]).stdout) == case["out_vcf_body_sha256"]
        var = out + "_var.bcf"
    else:
        text = run([CLI, "-x"] + case["extract_argv"] + ["-O", "v", "-f", xsi, "-o", "-"]).stdout
        assert body_sha(text) == case["vcf_body_sha256"], "xsqueezeit -x on the GPU path: records differ from the reference's"
        var = xsi + "_var.bcf"
    line = run([CAPI, var]).stdout.decode().split()
    got = {"records": int(line[1]), "genotypes": int(line[3]), "checksum": line[7]}
    assert got == case["capi_decode"], "c_xcf_get_genotypes stream differs from the reference's"


@pytest.mark.gpu
@needs_bindings
@pytest.mark.parametrize("key", sorted(MAN["chr20_small"]["options"]))
def test_reference_cli_chr20_options(tmp_path, key):
    """the six --maf / --variant-block-length runs of SURVEY.md 8(c): SHA-256 of the .xsi"""
    o = MAN["chr20_small"]["options"][key]
    out = str(tmp_path / "c.xsi")
    run([CLI, "-c"] + o["argv"] + ["-f", os.path.join(INP, "chr20_small.bcf"), "-o", out])
    assert sha(open(out, "rb").read()) == o["xsi_sha256"]


@pytest.mark.gpu
@needs_bindings
@pytest.mark.parametrize("name", sorted(k for k in MAN if k != "chr20_small"))
def test_reference_cli_small_fixtures(tmp_path, name):
    """the seven micro fixtures and test_region_target at their SURVEY 8(c) options, default and --wah-encode-missing"""
    src = os.path.join(INP, name + (".vcf" if name.startswith("micro") else ".bcf"))
    opts = ["--maf", "0.002"] if name.startswith("micro") else []
    out = str(tmp_path / "s.xsi")
    run([CLI, "-c"] + opts + ["-f", src, "-o", out])
    assert sha(open(out, "rb").read()) == MAN[name]["xsi_sha256"]
    run([CLI, "-c"] + opts + ["--wah-encode-missing", "-f", src, "-o", out])
    assert sha(open(out, "rb").read()) == MAN[name]["wah_missing"]["xsi_sha256"]


@pytest.mark.gpu
@needs_bindings
@pytest.mark.parametrize("name", ["chr20_small"] + sorted(k for k in MAN if k.startswith("micro")))
def test_lockstep_loader_on_gpu(tmp_path, name):
    """the reference's own parity definition (lockstep_loader/gt_lockstep_loader.hpp:83-157): the source file through
    bcf_get_genotypes and the .xsi through the GPU Accessor, in one synced reader, element for element"""
    src = os.path.join(INP, "chr20_small.bcf" if name == "chr20_small" else name + ".bgz.vcf")
    out = str(tmp_path / "l.xsi")
    run([CLI, "-c", "--maf", "0.002", "-f", os.path.join(INP, "chr20_small.bcf" if name == "chr20_small" else name + ".vcf"), "-o", out])
    p = run([LOCKSTEP, "--file1", src, "--file2", out + "_var.bcf"])
    assert b"Files have the same GT data" in p.stderr, p.stderr.decode()[-1000:]
    if name == "chr20_small":
        assert b"Checked 109635136 GT entries" in p.stderr


@pytest.mark.gpu
@needs_bindings
def test_reference_c_api_test_program(tmp_path):
    """the reference's minimal C consumer (c_api_test/main.c), linked against the GPU Accessor"""
    out = str(tmp_path / "c.xsi")
    run([CLI, "-c", "-f", os.path.join(INP, "chr20_small.bcf"), "-o", out])
    p = run([os.path.join(OUT, "c_api_test_b200"), out + "_var.bcf"])
    assert b"is 2504" in p.stdout and b"21892 records" in p.stdout, p.stdout


@pytest.mark.gpu
@needs_bindings
@pytest.mark.parametrize("name", ["chr20_small", "chr20_small_zstd_b1024", "test_region_target"] + sorted(k for k in CLIMAN if k.startswith("micro")))
def test_bcf_ingest_tool(tmp_path, name):
    """bindings/xsi_b200_bcf.cpp (threaded BGZF, raw int8 FORMAT/GT rows, several blocks per launch): the same file pair
    as the reference CLI, byte for byte (zstd runs: same decoded records)"""
    case = CLIMAN[name]
    out = str(tmp_path / (name + ".xsi"))
    argv = [a for a in case["compress_argv"] if a != "-c"]
    run([os.path.join(OUT, "xsi_b200_bcf"), "compress", os.path.join(INP, case["input"]), out, "--threads", "4", "--batch-blocks", "2"] + argv)
    if "xsi_sha256" in case:
        assert sha(open(out, "rb").read()) == case["xsi_sha256"]
    assert sha(open(out + "_var.bcf", "rb").read()) == case["var_sha256"], "single-pass _var.bcf differs from the reference's"
    assert os.path.exists(out + "_var.bcf.csi")
    if "x_bcf_sha256" in case:  # egress: rows spliced in as int8 FORMAT/GT, threaded deflate
        bcf = str(tmp_path / "o.bcf")
        run([os.path.join(OUT, "xsi_b200_bcf"), "extract", out, bcf, "--threads", "4", "--window-bytes", "4000000"])
        assert sha(open(bcf, "rb").read()) == case["x_bcf_sha256"], "xsi_b200_bcf extract: BCF differs from the reference's -x output"
    line = run([CAPI, out + "_var.bcf"]).stdout.decode().split()
    assert {"records": int(line[1]), "genotypes": int(line[3]), "checksum": line[7]} == case["capi_decode"]


SUBMAN = json.load(open(os.path.join(G, "subset_manifest.json")))


@pytest.mark.gpu
@needs_bindings
@pytest.mark.parametrize("name", sorted(SUBMAN["cases"]))
def test_subset_tool_matches_reference_Ox(tmp_path, name):
    """`xsi_b200_bcf subset` = `xsqueezeit -x -O x [-s/-S]` (gt_decompressor_new.hpp:241-273) with the rows kept on the device between
    xsi_decode_records_subset and xsi_encode_launch_strided: the new .xsi AND its companion (new BM, AC / AN of the selected
    samples) equal the unmodified reference's, byte for byte (tests/golden/make_subset_golden.py)"""
    case = SUBMAN["cases"][name]
    xsi = str(tmp_path / "in.xsi")
    run([CLI, "-c", "--maf", "0.002", "-f", os.path.join(INP, case["input"]), "-o", xsi])
    lst = str(tmp_path / "samples.txt")
    open(lst, "w").write(SUBMAN["sample_file"])
    out = str(tmp_path / "sub.xsi")
    for batch in ("1", "3"):
        run([os.path.join(OUT, "xsi_b200_bcf"), "subset", xsi, out, "--batch-blocks", batch] + [lst if a == "@LIST" else a for a in case["subset_argv"]])
        data = open(out, "rb").read()
        assert len(data) == case["xsi_size"]
        assert sha(data) == case["xsi_sha256"], "subset: .xsi differs from the reference's -Ox output"
        assert sha(open(out + "_var.bcf", "rb").read()) == case["var_sha256"], "subset: companion differs from the reference's"


# ---- the in-memory door (oracle/ref_shim.cpp: XsiFactoryExt + Accessor fed from arrays) built twice: CPU reference vs both adapters ----
SHIM = os.path.join(OUT, "libxsi_shim_b200.so")


def _plugin_cases():
    import numpy as np
    import synth
    rng = np.random.default_rng(7)
    ns = 120
    rows, ngt = [], []
    for r in range(200):
        p = 1 if r % 7 == 3 else 2
        al = (rng.random(ns * p) < 0.3).astype(np.int8)
        rows.append(synth.encode_gt(al, 1 if p == 2 else 0))
        ngt.append(ns * p)
    mixed = dict(gt=np.concatenate(rows).astype(np.int32), ngt=np.array(ngt, np.int32), n_allele=np.full(200, 2, np.int32), n_samples=ns)
    return [("biallelic", synth.make_dataset(700, 301, seed=1), 256, 0.01),
            ("multiallelic_missing_eov", synth.make_dataset(600, 257, seed=2, max_alt=4, multi_frac=0.2, missing=0.01, unphased=0.02, haploid_samples=0.4), 128, 0.02),
            ("all_sparse", synth.make_dataset(300, 257, seed=3, max_alt=3, multi_frac=0.2, missing=0.01), 8192, 0.3),
            ("mixed_ploidy_records", mixed, 64, 0.01),
            ("uint32_indices", synth.make_dataset(12, 66000, seed=5, n_founders=16, fmin=0.001), 5, 0.001)]


@pytest.mark.gpu
@needs_bindings
@pytest.mark.skipif(not os.path.exists(SHIM), reason="bindings/_out/libxsi_shim_b200.so not built")
@pytest.mark.parametrize("case", range(5))
def test_plugin_interfaces_against_the_live_reference(tmp_path, case):
    """The reference's writer (XsiFactoryExt) with GtBlockB200 and its reader (Accessor) with AccessorInternalsB200, fed from
    memory exactly like the CPU reference in tests/test_oracle_vs_reference.py: same file bytes; fill_genotype_array /
    get_allele_counts / fill_allele_counts / get_internal_access equal record by record, forward, backward and shuffled."""
    import numpy as np
    import xsi_oracle as xo
    import xsi_ref
    if not xsi_ref.available():
        pytest.skip("oracle/_ref/libxsi_ref.so not built")
    name, ds, bl, maf = _plugin_cases()[case]
    gt, ngt, nal, ns = ds["gt"], ds["ngt"], ds["n_allele"], ds["n_samples"]
    off = xo.row_offsets(ngt)
    dp = xo.default_phased(gt, off, ngt, ns)
    thr = xo.mac_threshold(ns, int(ngt[0]) // ns, maf)
    B = xsi_ref.open_lib(SHIM)
    pr, pb = str(tmp_path / "ref.xsi"), str(tmp_path / "b200.xsi")
    xsi_ref.encode_file(pr, gt, off, ngt, nal, ns, bl, thr, dp)
    xsi_ref.encode_file(pb, gt, off, ngt, nal, ns, bl, thr, dp, L=B)
    assert open(pr, "rb").read() == open(pb, "rb").read(), "XsiFactoryExt + GtBlockB200 wrote different bytes"
    pos = xo.bm_positions(nal, bl)
    R = len(nal)
    ref = xsi_ref.RefAccessor(pr)
    want = []
    for r in range(R):
        row, n = ref.fill_genotype_array(int(nal[r]), int(pos[r]))
        want.append((row[:n].copy(), n, ref.allele_counts()))
    ref.close()
    acc = xsi_ref.RefAccessor(pb, L=B)
    rng = np.random.default_rng(case)
    for order in (range(R), range(R - 1, -1, -1), rng.permutation(R)[:150]):
        for r in order:
            row, n = acc.fill_genotype_array(int(nal[r]), int(pos[r]))
            assert n == want[r][1] and np.array_equal(row[:n], want[r][0]), (name, r)
            assert np.array_equal(acc.allele_counts(), want[r][2]), (name, r)
    acc.close()
    # counts only: fresh cursors on both sides, file order (the reference's own call pattern)
    ref, acc = xsi_ref.RefAccessor(pr), xsi_ref.RefAccessor(pb, L=B)
    for r in range(R):
        assert np.array_equal(acc.fill_allele_counts(int(nal[r]), int(pos[r])), ref.fill_allele_counts(int(nal[r]), int(pos[r]))), (name, r)
    ref.close()
    acc.close()
    # InternalGtAccess: sparse flags, the bytes the pointers point at, the default allele, the arrangement a[]
    if name != "mixed_ploidy_records":  # haploid lines: the arrangement is only kept by the lazy chain kernel (diploid lines)
        a_bytes = 2 if 2 * ns <= 65535 else 4
        ref, acc = xsi_ref.RefAccessor(pr), xsi_ref.RefAccessor(pb, L=B)
        if a_bytes == 2:
            seq = list(range(0, R, 3)) + [R // 2, 1, R - 1]  # forward, then requests behind the chain (block reload)
            for r in seq:
                if int(nal[r]) > 3:
                    # the reference's loop seeks to (cursor + i) while the cursor itself advances (accessor_internals_new.hpp:456-458):
                    # with three or more ALT alleles it skips lines (P, P+1, P+3, ...); the adapter returns the consecutive lines
                    continue
                ra = ref.internal_access(int(nal[r]), int(pos[r]), a_bytes)
                ba = acc.internal_access(int(nal[r]), int(pos[r]), a_bytes)
                assert np.array_equal(ra[1], ba[1]) and np.array_equal(ra[2], ba[2]) and ra[3] == ba[3], (name, r)
                assert np.array_equal(ra[0], ba[0]), (name, r, "arrangement")
        ref.close()
        acc.close()
