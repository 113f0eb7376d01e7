mkdir -p gpurun_out
T=${T:-r01aa}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
run() {  # name args...
  n=$1; shift
  timeout 900 python bench.py "$@" --no-e2e --no-cpu-baseline > gpurun_out/${T}_bench_$n.json 2> gpurun_out/${T}_bench_$n.err; echo "$n rc=$?"
  python - <<P
import json
try:
    d=json.loads(open('gpurun_out/${T}_bench_$n.json').read().strip().splitlines()[-1]); k=d["kernels"]
    print("$n", "value %.1f (one ctx %.1f) enc %.1f dec %.1f verified %s | " % (d["value"], d["value_one_context"], d["compress_ggts"], d["decompress_ggts"], d["verified"]) + " ".join("%s %.2f" % (a, v["ms_per_step"]) for a, v in k.items() if v["ms_per_step"] > 0.3))
except Exception as e:
    print("$n failed", e)
P
}
run kgp --samples 2504 --blocks 220 --steps 3 --warmup 2 --resident-contexts 0
run chrx --shape chrx --samples 2504 --blocks 24 --steps 3 --warmup 2
run hrc --steps 3 --warmup 3 --resident-contexts 0
