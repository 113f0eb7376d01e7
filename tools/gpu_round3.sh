# int8 ingest/egress validation: tests, default bench (with e2e_bcf_int8), int8-resident sweep
mkdir -p gpurun_out
T=${T:-r01d}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
cat gpurun_out/${T}_bench.json
for cfg in "1 32" "1 64" "1 128"; do
  set -- $cfg
  timeout 600 python bench.py --elem $1 --blocks $2 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/${T}_bench_e$1_b$2.json 2> gpurun_out/${T}_bench_e$1_b$2.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${T}_bench_e$1_b$2.json")); k=d["kernels"]
    print("elem $1 blocks $2: value %.1f enc %.1f dec %.1f verified %s | " % (d["value"], d["compress_ggts"], d["decompress_ggts"], d["verified"]) + " ".join("%s %.2f" % (n, v["ms_per_step"]) for n, v in k.items() if v["ms_per_step"] > 0.5), d["call_wall_ms_per_step"])
except Exception as e:
    print("elem $1 blocks $2 failed", e)
PY
done
