# round 2: async encode (xsi_encode_async) parity + the pipelined one-context leg on HRC / 1KGP3 / chrX shapes
mkdir -p gpurun_out
T=${T:-r02f}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -5 gpurun_out/${T}_pytest.log
show='import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
p=d.get("one_context_pipelined") or {}
m=d.get("resident_multi_context") or {}
print("value %.1f (%s)  one ctx %.1f (%.2f ms)  pipelined %s (%s ms) ok %s  mt %s  verified %s" % (d["value"], d["value_mode"], d["value_one_context"], d["ms_per_step_one_context"], p.get("value"), p.get("ms_per_step"), p.get("verified"), m.get("value"), d["verified"]))
print("  " + "  ".join("%s %.2f" % (k, v["ms_per_step"]) for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:8]))'
for shape in "--blocks 32" "--samples 2504 --blocks 220" "--samples 2504 --blocks 24 --shape chrx"; do
  echo "== $shape"
  timeout 600 python bench.py --no-e2e --no-cpu-baseline --bcf-records 0 --no-shapes --resident-contexts 3 --steps 5 --warmup 3 $shape 2>gpurun_out/${T}_err.txt | python -c "$show" || tail -5 gpurun_out/${T}_err.txt
done 2>&1 | tee gpurun_out/${T}_pipelined.txt
