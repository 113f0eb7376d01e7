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
