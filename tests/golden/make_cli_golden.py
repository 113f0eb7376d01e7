#!/usr/bin/env python
"""Golden vectors for the reference's OWN programs (runs only where /root/reference exists).

The bindings under bindings/ make the reference CLI, Accessor and C API run on the B200 path; their tests compare
against what the UNMODIFIED reference (oracle/_ref/xsqueezeit_ref, bindings/_out/capi_decode_ref) produces for the
17 cases of the reference's test configuration (test/cukinia_v4.conf:4-20: compress with --maf 0.002, extract with
optional -r / -t / -s), plus the 14 compress runs whose .xsi hashes are already in manifest.json.

Writes:
  tests/golden/inputs/     the input fixtures of test/test_files (DATA files: VCF/BCF + index; the micro VCFs also
                           bgzipped + tabix-indexed, which lockstep_loader needs, gt_lockstep_loader.hpp:87-99)
  tests/golden/cli_manifest.json
       per case: argv of -c and -x, SHA-256 of the .xsi, SHA-256 of the `-x -Ov` record lines (header lines that
       start with '##' carry command lines and are left out, like verify_v4.sh:112-129 tolerates them), and the
       c_xcf_get_genotypes checksum of bindings/capi_decode.c.
Usage: make -C oracle ref && make -C bindings && python tests/golden/make_cli_golden.py
"""
import hashlib
import json
import os
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
FIX = "/root/reference/test/test_files"
REF = os.path.join(ROOT, "oracle", "_ref")
CLI = os.path.join(REF, "xsqueezeit_ref")
CAPI = os.path.join(ROOT, "bindings", "_out", "capi_decode_ref")
INPUTS = os.path.join(HERE, "inputs")

MICRO = ["micro_missing", "micro_eov", "micro_haploid", "micro_mixed_ploidy", "micro_non_uniform_phase",
         "micro_missing_non_uniform_phasing", "micro_missing_non_uniform_phasing_ploidy"]
# test/cukinia_v4.conf:4-20 (verify_v4.sh:98-99: -c always with --maf 0.002)
CASES = [(m, m + ".vcf", [], []) for m in MICRO] + [
    ("chr20_small", "chr20_small.bcf", [], []),
    ("chr20_small_zstd", "chr20_small.bcf", ["--zstd"], []),
    ("chr20_small_zstd_b4096", "chr20_small.bcf", ["--zstd", "--variant-block-length", "4096"], []),
    ("chr20_small_zstd_b1024", "chr20_small.bcf", ["--zstd", "--variant-block-length", "1024"], []),
    ("chr20_small_region", "chr20_small.bcf", [], ["-r", "20:100000-200000"]),
    ("chr20_small_samples", "chr20_small.bcf", [], ["-s", "NA12878,HG00110,HG00112"]),
    ("chr20_small_samples_order", "chr20_small.bcf", [], ["-s", "HG00112,HG00110,NA12878"]),
    ("chr20_small_samples_neg", "chr20_small.bcf", [], ["-s", "^NA12878,HG00110"]),
    ("chr20_small_region_samples", "chr20_small.bcf", [], ["-r", "20:100000-200000", "-s", "NA12878,HG00110,HG00112"]),
    ("test_region_target", "test_region_target.bcf", [], ["-t", "chr17:117980-117999"]),
    # beyond the reference's list: the XSI -> XSI path (-Ox, gt_decompressor_new.hpp:130-143,241-273) with a subset
    ("chr20_small_samples_Ox", "chr20_small.bcf", [], ["-s", "NA12878,HG00110,HG00112", "-O", "x"]),
]


def sha(b):
    return hashlib.sha256(b).hexdigest()


def body_sha(text):
    return sha(b"".join(l for l in text.splitlines(True) if not l.startswith(b"##")))


def run(argv, **kw):
    return subprocess.run(argv, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, **kw).stdout


def capi(var_bcf):
    line = run([CAPI, var_bcf]).decode().split()
    return {"records": int(line[1]), "genotypes": int(line[3]), "checksum": line[7]}


def main():
    os.makedirs(INPUTS, exist_ok=True)
    for f in sorted(os.listdir(FIX)):
        shutil.copyfile(os.path.join(FIX, f), os.path.join(INPUTS, f))
        os.chmod(os.path.join(INPUTS, f), 0o644)
    for m in MICRO:  # indexed copies for lockstep_loader
        gz = os.path.join(INPUTS, m + ".bgz.vcf")
        with open(gz, "wb") as out:
            out.write(run([os.path.join(REF, "bgzip"), "-c", os.path.join(INPUTS, m + ".vcf")]))
        run([os.path.join(REF, "tabix"), "-f", "-p", "vcf", gz])
    manifest = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, src, copts, xopts in CASES:
            xsi = os.path.join(tmp, name + ".xsi")
            c_argv = ["-c"] + copts + ["--maf", "0.002"]
            run([CLI] + c_argv + ["-f", os.path.join(INPUTS, src), "-o", xsi])
            entry = {"input": src, "compress_argv": c_argv, "extract_argv": xopts, "xsi_size": os.path.getsize(xsi)}
            if "--zstd" not in copts:  # zstd frames depend on the libzstd build; compare by content there
                entry["xsi_sha256"] = sha(open(xsi, "rb").read())
            entry["var_sha256"] = sha(open(xsi + "_var.bcf", "rb").read())  # companion (holds ##XSI=<basename>: tests keep the name)
            if not xopts:
                bcf = os.path.join(tmp, name + "_x.bcf")
                run([CLI, "-x", "-f", xsi, "-o", bcf])
                entry["x_bcf_sha256"] = sha(open(bcf, "rb").read())
            if "x" in xopts:  # -Ox writes a new .xsi + _var.bcf pair: record that file and its decoded content
                out = os.path.join(tmp, name + "_out.xsi")
                run([CLI, "-x"] + xopts + ["-f", xsi, "-o", out])
                entry["out_xsi_sha256"] = sha(open(out, "rb").read())
                entry["out_vcf_body_sha256"] = body_sha(run([CLI, "-x", "-O", "v", "-f", out, "-o", "-"]))
                entry["capi_decode"] = capi(out + "_var.bcf")
            else:
                entry["vcf_body_sha256"] = body_sha(run([CLI, "-x"] + xopts + ["-O", "v", "-f", xsi, "-o", "-"]))
                entry["capi_decode"] = capi(xsi + "_var.bcf")
            manifest[name] = entry
            print(name, entry)
    json.dump(manifest, open(os.path.join(HERE, "cli_manifest.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
