# experiment: pbwt_permute_v4 compiled with a register cap (48 / 40) so that HBM-bound CTAs of the decode stream can share
# its SMs in the pipelined one-context leg; plus the dead-warp skip / KH=16 changes on the 1KGP3 shape
mkdir -p gpurun_out
T=${T:-r02h}
show='import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
p=d.get("one_context_pipelined") or {}
m=d.get("resident_multi_context") or {}
print("one ctx %.1f (%.2f ms)  pipelined %.1f (%.2f ms) ok %s  mt %s  verified %s  permute %.2f" % (d["value_one_context"], d["ms_per_step_one_context"], p.get("value",0), p.get("ms_per_step",0), p.get("verified"), m.get("value"), d["verified"], d["kernels"]["pbwt_permute"]["ms_per_step"]))'
for so in "" xsqueezeit_b200/libxsi_b200_r48.so xsqueezeit_b200/libxsi_b200_r40.so; do
  echo "== HRC 32 blocks, lib=${so:-default}"
  XSI_B200_SO=${so:+$PWD/$so} timeout 600 python bench.py --no-e2e --no-cpu-baseline --bcf-records 0 --no-shapes --resident-contexts 3 --steps 5 --warmup 3 2>gpurun_out/${T}_err.txt | python -c "$show" || tail -3 gpurun_out/${T}_err.txt
done 2>&1 | tee gpurun_out/${T}_regcap.txt
echo "== 1KGP3 220 blocks (dead-warp skip, KH=16)"
timeout 600 python bench.py --no-e2e --no-cpu-baseline --bcf-records 0 --no-shapes --resident-contexts 3 --steps 4 --warmup 2 --samples 2504 --blocks 220 2>/dev/null | python -c "$show" | tee -a gpurun_out/${T}_regcap.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
