#!/usr/bin/env python
"""Golden vectors for the XSI -> XSI path of the reference's extractor (`xsqueezeit -x -O x [-s/-S]`,
include/gt_decompressor_new.hpp:130-143,241-273); runs only where the reference is built (oracle/_ref).

Per case: the compress argv, the extractor argv and the SHA-256 of BOTH files the UNMODIFIED reference writes
(`sub.xsi` and `sub.xsi_var.bcf`; the companion holds `##XSI=sub.xsi`, so the tests keep that name).
`bindings/_out/xsi_b200_bcf subset` (rows stay on the device between decode and encode) must reproduce them.
Usage: make -C oracle ref && python tests/golden/make_subset_golden.py   -> tests/golden/subset_manifest.json
"""
import hashlib
import json
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CLI = os.path.join(ROOT, "oracle", "_ref", "xsqueezeit_ref")
INPUTS = os.path.join(HERE, "inputs")

# name: (input, reference -x options, xsi_b200_bcf subset options)
CASES = {
    "chr20_three": ("chr20_small.bcf", ["-s", "NA12878,HG00110,HG00112"], ["--samples", "NA12878,HG00110,HG00112"]),
    "chr20_exclude_two": ("chr20_small.bcf", ["-s", "^NA12878,HG00110"], ["--samples", "^NA12878,HG00110"]),
    "chr20_all_maf01": ("chr20_small.bcf", ["--maf", "0.01"], ["--maf", "0.01"]),
    "chr20_file": ("chr20_small.bcf", ["-S", "@LIST"], ["--samples-file", "@LIST"]),
    "mixed_ploidy_three": ("micro_mixed_ploidy.vcf", ["-s", "HG00119,HG00111,HG00113"], ["--samples", "HG00119,HG00111,HG00113"]),
    "missing_phasing_ploidy_exclude": ("micro_missing_non_uniform_phasing_ploidy.vcf", ["-s", "^HG00110"], ["--samples", "^HG00110"]),
    "haploid_three": ("micro_haploid.vcf", ["-s", "HG00119,HG00111,HG00113"], ["--samples", "HG00119,HG00111,HG00113"]),
    "eov_two": ("micro_eov.vcf", ["-s", "HG00114,HG00112"], ["--samples", "HG00114,HG00112"]),
    "missing_all": ("micro_missing.vcf", [], []),
}
LIST = "HG00096\nHG00100\tignored\nNA12878\nHG00112\n"  # -S: first tab-separated field of every line


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    man = {"sample_file": LIST, "cases": {}}
    with tempfile.TemporaryDirectory() as tmp:
        lst = os.path.join(tmp, "samples.txt")
        open(lst, "w").write(LIST)
        for name, (src, xopts, topts) in CASES.items():
            d = os.path.join(tmp, name)
            os.makedirs(d)
            xsi = os.path.join(d, "in.xsi")
            subprocess.run([CLI, "-c", "--maf", "0.002", "-f", os.path.join(INPUTS, src), "-o", xsi], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            out = os.path.join(d, "sub.xsi")
            subprocess.run([CLI, "-x"] + [lst if a == "@LIST" else a for a in xopts] + ["-O", "x", "-f", xsi, "-o", out], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            man["cases"][name] = {"input": src, "reference_argv": xopts + ["-O", "x"], "subset_argv": topts,
                                  "xsi_sha256": sha(out), "xsi_size": os.path.getsize(out), "var_sha256": sha(out + "_var.bcf")}
            print(name, man["cases"][name])
    json.dump(man, open(os.path.join(HERE, "subset_manifest.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
