"""Debug helper (GPU box): bench generator at biobank width, GPU encode/decode vs the oracle."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import numpy as np, torch
import xsi_oracle as xo, xsqueezeit_b200 as xb
import bench, debug_diff

S = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 256
BL = int(sys.argv[3]) if len(sys.argv) > 3 else 128
gen = bench.HrcSynth(S, 1002, "cuda:0")
g = torch.empty((R, 2 * S), dtype=torch.int32, device="cuda:0")
gen.fill(g)
gt = g.cpu().numpy().reshape(-1)
del g
ngt = np.full(R, 2 * S, np.int32); nal = np.full(R, 2, np.int32)
ds = dict(gt=gt, ngt=ngt, n_allele=nal, n_samples=S)
off = xo.row_offsets(ngt)
thr = xo.mac_threshold(S, 2, 0.001)
dp = xo.default_phased(gt, off, ngt, S)
img = xo.encode(gt, off, ngt, nal, S, BL, thr, dp)
ctx = xb.Context(0)
with tempfile.TemporaryDirectory() as tmp:
    p = os.path.join(tmp, "x.xsi")
    xb.Compressor(ctx, maf=0.001, reset_sort_block_length=BL).compress_to_file(p, gt, ngt, nal, S)
    a = open(p, "rb").read()
    ok = debug_diff.compare(a, img, "biobank S=%d R=%d BL=%d" % (S, R, BL))
    # decode the ORACLE image on the GPU and compare rows with the input
    q = os.path.join(tmp, "o.xsi")
    open(q, "wb").write(img)
    acc = xb.Accessor(q, ctx)
    pos = xb.bm_positions(nal, BL)
    out, filled, _ = acc.fill_genotype_arrays(nal, pos)
    bad = 0
    for r in range(R):
        row = gt[r * 2 * S:(r + 1) * 2 * S]
        if not np.array_equal(out[r, :filled[r]], row):
            d = np.nonzero(out[r, :filled[r]] != row)[0]
            cnt = int((row >> 1 == 2).sum())
            if bad < 8:
                print("decode row", r, "carriers", cnt, "wah" if min(cnt, 2 * S - cnt) > thr else "sparse", "ndiff", len(d), d[:6], out[r, d[:6]], row[d[:6]])
            bad += 1
    print("decode bad rows", bad, "of", R)
