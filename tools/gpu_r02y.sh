# round 2: decode request arrays packed into pinned memory by the pool, checks in parallel
mkdir -p gpurun_out
T=${T:-r02y}
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log
show='import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
p=d.get("one_context_pipelined") or {}
print("value %.1f pipelined %.1f one-ctx %.1f enc %.1f dec %.1f ms/step %.2f verified %s" % (d["value"], p.get("value",0), d["value_one_context"], d["compress_ggts"], d["decompress_ggts"], d["ms_per_step"], d["verified"]))
print("  " + "  ".join("%s %.2f" % (k, v["ms_per_step"]) for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:9]))
print("  wall", {k: round(v,2) for k,v in d["call_wall_ms_per_step"].items()}, "host", {k: round(v["ms_per_step"],2) for k,v in d["host_phases"].items()})'
run() { echo "== $1 | $2"; env $1 timeout 600 python bench.py --sub --warmup 2 --steps 4 $2 2>/dev/null | python -c "$show"; }
{
run "XSI_X=0" "--samples 2504 --blocks 220"
run "XSI_X=0" "--samples 2504 --blocks 24 --shape chrx"
run "XSI_X=0" "--blocks 32"
} 2>&1 | tee gpurun_out/${T}_shapes.txt
