# round 2: the two-lines-per-round PBWT kernel (v5): parity, then v5 vs v4 timing on the HRC / 1KGP3 / chrX shapes
mkdir -p gpurun_out
T=${T:-r02i}
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "variants or golden or biallelic or hrc_shape or kgp_shape or max_uint16" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -25 gpurun_out/${T}_pytest.log
show='import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
p=d.get("one_context_pipelined") or {}
print("one ctx %.1f (%.2f ms) enc %.1f dec %.1f pipelined %.1f verified %s  permute %.2f ms" % (d["value_one_context"], d["ms_per_step_one_context"], d["compress_ggts"], d["decompress_ggts"], p.get("value",0), d["verified"], d["kernels"]["pbwt_permute"]["ms_per_step"]))'
for v in 5 4; do
  for shape in "--blocks 32" "--samples 2504 --blocks 220" "--samples 2504 --blocks 24 --shape chrx"; do
    echo "== XSI_PBWT_V=$v $shape"
    XSI_PBWT_V=$v timeout 600 python bench.py --no-e2e --no-cpu-baseline --bcf-records 0 --no-shapes --resident-contexts 0 --steps 4 --warmup 2 $shape 2>gpurun_out/${T}_err.txt | python -c "$show" || tail -5 gpurun_out/${T}_err.txt
  done
done 2>&1 | tee gpurun_out/${T}_v5.txt
