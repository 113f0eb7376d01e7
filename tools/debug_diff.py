"""Debug helper (GPU box): encode a dataset on the GPU and with the oracle, show where bytes differ."""
import os, sys, struct, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import synth, xsi_oracle as xo, xsqueezeit_b200 as xb

NAMES = {0: "BCF_LINES", 1: "BIN_LINES", 2: "MAX_PLOIDY", 3: "DEF_PHASING", 4: "WEIRD", 0x10: "L_SORT", 0x11: "L_SELECT",
         0x12: "L_HAPLOID", 0x16: "L_MISSING", 0x17: "L_PHASE", 0x18: "L_EOV", 0x20: "M_WAH", 0x21: "M_SPARSE",
         0x26: "M_MISSING", 0x27: "M_PHASE", 0x28: "M_EOV", 0x36: "M_MISS_SP", 0x38: "M_EOV_SP"}


def parse(img):
    idx_off, smp_off = struct.unpack_from("<QQ", img, 72)
    nb = (smp_off - idx_off) // 8
    offs = struct.unpack_from("<%dQ" % nb, img, idx_off)
    blocks = []
    for b in range(nb):
        o = offs[b] + 16
        end = (offs[b + 1] if b + 1 < nb else idx_off)
        n = struct.unpack_from("<I", img, o + 4)[0]
        d = [struct.unpack_from("<II", img, o + 8 + 8 * i) for i in range(n)]
        blocks.append((o, end, d))
    return blocks


def compare(a, g, label):
    print("==", label, "sizes", len(a), len(g), "equal", a == g)
    if a == g:
        return True
    if a[:256] != g[:256]:
        for i in range(0, 256, 4):
            if a[i:i + 4] != g[i:i + 4]:
                print("  header differs at", i, a[i:i + 8].hex(), g[i:i + 8].hex())
    ba, bg = parse(a), parse(g)
    print("  blocks", len(ba), len(bg))
    for b, (x, y) in enumerate(zip(ba, bg)):
        da, dg = x[2], y[2]
        if [k for k, _ in da] != [k for k, _ in dg]:
            print("  block", b, "dict key order differs", [hex(k) for k, _ in da], [hex(k) for k, _ in dg])
        va, vg = dict(da), dict(dg)
        secs = sorted((v, k) for k, v in vg.items() if k >= 0x10 and v != 0xFFFFFFFF)
        for k in sorted(set(va) | set(vg)):
            if va.get(k) != vg.get(k):
                print("  block", b, "key", NAMES.get(k, hex(k)), "gpu", va.get(k), "oracle", vg.get(k))
        # compare sections by oracle layout
        for i, (off, k) in enumerate(secs):
            end = secs[i + 1][0] if i + 1 < len(secs) else y[1] - y[0]
            sg = g[y[0] + off:y[0] + end]
            oa = va.get(k)
            if oa is None or oa == 0xFFFFFFFF:
                print("  block", b, NAMES.get(k), "missing on gpu"); continue
            sa = a[x[0] + oa:x[0] + oa + len(sg)]
            if sa != sg:
                w = next(i for i in range(min(len(sa), len(sg))) if sa[i] != sg[i]) if len(sa) and len(sg) else 0
                print("  block", b, "section", NAMES.get(k), "len", len(sg), "first diff at byte", w,
                      "gpu", sa[w:w + 16].hex(), "oracle", sg[w:w + 16].hex())
        if b > 3:
            break
    return False


def run(ds, block_len, maf, label, elem=4):
    ctx = xb.Context(0)
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "x.xsi")
        gt = ds["gt"]
        xb.Compressor(ctx, maf=maf, reset_sort_block_length=block_len).compress_to_file(p, gt, ds["ngt"], ds["n_allele"], ds["n_samples"], gt_elem_bytes=elem)
        a = open(p, "rb").read()
    off = xo.row_offsets(ds["ngt"])
    g = xo.encode(ds["gt"], off, ds["ngt"], ds["n_allele"], ds["n_samples"], block_len,
                  xo.mac_threshold(ds["n_samples"], int(ds["ngt"][0]) // ds["n_samples"], maf),
                  xo.default_phased(ds["gt"], off, ds["ngt"], ds["n_samples"]))
    ok = compare(a, g, label)
    ctx.close()
    return ok


if __name__ == "__main__":
    run(synth.make_dataset(20, 10, seed=1), 8192, 0.0, "tiny biallelic all-wah")
    run(synth.make_dataset(20, 10, seed=1), 8192, 0.9, "tiny biallelic all-sparse")
    run(synth.make_dataset(64, 100, seed=2), 32, 0.01, "small 2 blocks")
    run(synth.make_dataset(300, 1252, seed=42), 128, 0.001, "mid biallelic")
    run(synth.make_dataset(100, 100, seed=3, missing=0.02), 8192, 0.01, "missing")
    run(synth.make_dataset(100, 100, seed=3, haploid_samples=0.3), 8192, 0.01, "eov")
    run(synth.make_dataset(100, 100, seed=3, unphased=0.05), 8192, 0.01, "phase")
    run(synth.make_dataset(100, 100, seed=3, max_alt=3, multi_frac=0.3), 8192, 0.01, "multi")
