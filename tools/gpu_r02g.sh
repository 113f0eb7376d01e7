# experiment: haplotypes per thread (KH) / cluster size of pbwt_permute_v4 on the short-row shapes (more, smaller CTAs per SM)
mkdir -p gpurun_out
T=${T:-r02g}
show='import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print("value %.1f enc %.1f dec %.1f verified %s permute %.2f ms unpermute %.2f ms" % (d["value"], d["compress_ggts"], d["decompress_ggts"], d["verified"], d["kernels"]["pbwt_permute"]["ms_per_step"], d["kernels"]["pbwt_unpermute"]["ms_per_step"]))'
for kh in 8 16 32 64; do
  echo "== 1KGP3 KH=$kh"; XSI_PBWT_KH=$kh timeout 300 python bench.py --sub --steps 2 --warmup 1 --samples 2504 --blocks 220 2>/dev/null | python -c "$show"
done 2>&1 | tee gpurun_out/${T}_kh.txt
for c in 8 4 2 1; do for kh in 8 16; do
  echo "== chrX C=$c KH=$kh"; XSI_PBWT_CLUSTER=$c XSI_PBWT_KH=$kh timeout 300 python bench.py --sub --steps 2 --warmup 1 --samples 2504 --blocks 24 --shape chrx 2>/dev/null | python -c "$show"
done; done 2>&1 | tee -a gpurun_out/${T}_kh.txt
