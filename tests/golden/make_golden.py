#!/usr/bin/env python
"""Regenerates tests/golden/ from the UNMODIFIED reference (runs only where /root/reference exists).

For every fixture of /root/reference/test/test_files it
  1. dumps what the reference compressor sees (oracle/_ref/gtdump: bcf_get_genotypes rows, n_allele),
  2. runs the reference CLI (oracle/_ref/xsqueezeit_ref -c) with the options of SURVEY.md section 8(c),
  3. decodes the result with the reference Accessor (oracle/_ref/libxsi_ref.so) and checks it equals (1),
and stores: small fixtures in full (<name>.npz + <name>.xsi), chr20_small as the reference .xsi for the
default options + record metadata + SHA-256 of the genotype stream and of every option combo's .xsi.
Usage: make -C oracle ref && python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import xsi_ref  # noqa: E402

FIX = "/root/reference/test/test_files"
GTDUMP = os.path.join(ROOT, "oracle", "_ref", "gtdump")

MICRO = ["micro_eov", "micro_haploid", "micro_missing", "micro_missing_non_uniform_phasing",
         "micro_missing_non_uniform_phasing_ploidy", "micro_mixed_ploidy", "micro_non_uniform_phase"]
CHR20_OPTS = {
    "default": [],
    "maf0.002": ["--maf", "0.002"],
    "maf0.002_b1024": ["--maf", "0.002", "--variant-block-length", "1024"],
    "maf0.002_b4096": ["--maf", "0.002", "--variant-block-length", "4096"],
    "maf0.01": ["--maf", "0.01"],
    "maf0": ["--maf", "0"],
}


def sha(b):
    return hashlib.sha256(b).hexdigest()


def run_cli(src, out, opts):
    subprocess.run([xsi_ref.REF_CLI, "-c", "-f", src, "-o", out] + opts, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return open(out, "rb").read()


def ref_decode_all(xsi_path, n_allele, ngt, block_len):
    acc = xsi_ref.RefAccessor(xsi_path)
    rows = []
    blk_off = 0
    for r in range(len(n_allele)):
        if r % block_len == 0:
            blk_off = 0
        pos = ((r // block_len) << 15) | blk_off
        out, n = acc.fill_genotype_array(int(n_allele[r]), pos)
        rows.append(out[:n].copy())
        blk_off += int(n_allele[r]) - 1
    acc.close()
    return rows


def main():
    manifest = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name in MICRO + ["test_region_target"]:
            src = os.path.join(FIX, name + (".vcf" if name.startswith("micro") else ".bcf"))
            pre = os.path.join(tmp, name)
            subprocess.run([GTDUMP, src, pre], check=True, stderr=subprocess.DEVNULL)
            ns, nal, ngt, gt, names = xsi_ref.load_gtdump(pre)
            opts = ["--maf", "0.002"] if name.startswith("micro") else []
            maf = 0.002 if name.startswith("micro") else 0.001
            xsi = run_cli(src, pre + ".xsi", opts)
            rows = ref_decode_all(pre + ".xsi", nal, ngt, 8192)
            dec = np.concatenate(rows) if rows else np.zeros(0, np.int32)
            dec_ngt = np.array([len(x) for x in rows], dtype=np.int32)
            np.savez_compressed(os.path.join(HERE, name + ".npz"), n_samples=ns, n_allele=nal, ngt=ngt, gt=gt,
                                names=np.array(names), maf=maf, block_len=8192, ref_decoded=dec,
                                ref_decoded_ngt=dec_ngt)
            open(os.path.join(HERE, name + ".xsi"), "wb").write(xsi)
            manifest[name] = {"xsi_sha256": sha(xsi), "xsi_size": len(xsi), "maf": maf, "block_len": 8192,
                              "ref_decode_equals_input": bool(dec.size == gt.size and np.array_equal(dec, gt))}
            # the same fixture through --wah-encode-missing (WS_WAH): the .xsi is stored too (the GPU decode tests read it)
            xw = run_cli(src, pre + "_wm.xsi", opts + ["--wah-encode-missing"])
            rows_w = ref_decode_all(pre + "_wm.xsi", nal, ngt, 8192)
            dec_w = np.concatenate(rows_w) if rows_w else np.zeros(0, np.int32)
            open(os.path.join(HERE, name + "_wah_missing.xsi"), "wb").write(xw)
            manifest[name]["wah_missing"] = {"xsi_sha256": sha(xw), "xsi_size": len(xw),
                                             "ref_decode_equals_default": bool(np.array_equal(dec_w, dec))}
            print(name, manifest[name])

        # chr20_small: too big to store raw (438 MB of int32) -> reference .xsi + metadata + hashes
        src = os.path.join(FIX, "chr20_small.bcf")
        pre = os.path.join(tmp, "chr20_small")
        subprocess.run([GTDUMP, src, pre], check=True, stderr=subprocess.DEVNULL)
        ns, nal, ngt, gt, names = xsi_ref.load_gtdump(pre)
        entry = {"n_samples": ns, "n_records": int(len(nal)), "gt_sha256": sha(gt.tobytes()), "options": {}}
        for key, opts in CHR20_OPTS.items():
            xsi = run_cli(src, pre + "_" + key + ".xsi", opts)
            entry["options"][key] = {"argv": opts, "xsi_sha256": sha(xsi), "xsi_size": len(xsi)}
            if key == "default":
                open(os.path.join(HERE, "chr20_small_default.xsi"), "wb").write(xsi)
                rows = ref_decode_all(pre + "_default.xsi", nal, ngt, 8192)
                dec = np.concatenate(rows)
                entry["ref_decode_sha256"] = sha(dec.tobytes())
                entry["ref_decode_equals_input"] = bool(np.array_equal(dec, gt))
        np.savez_compressed(os.path.join(HERE, "chr20_small_meta.npz"), n_samples=ns, n_allele=nal.astype(np.uint8),
                            ngt=ngt, names=np.array(names))
        manifest["chr20_small"] = entry
        print("chr20_small", json.dumps(entry, indent=1))
    json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
