mkdir -p gpurun_out
T=${T:-r01s}
timeout 600 python -m pytest tests -m gpu -x -q -k "biobank or uint32" 2>&1 | tail -15
timeout 900 python bench.py --samples 500000 --blocks 2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${T}_bench_biobank.json 2> gpurun_out/${T}_bench_biobank.err; echo "biobank rc=$?"
tail -n 3 gpurun_out/${T}_bench_biobank.err
python - <<P
import json
for n in ("biobank",):
    try:
        d=json.loads(open('gpurun_out/${T}_bench_%s.json'%n).read().strip().splitlines()[-1]); k=d["kernels"]
        print(n, "value %.1f enc %.1f dec %.1f verified %s | " % (d["value"], d["compress_ggts"], d["decompress_ggts"], d["verified"]) + " ".join("%s %.2f" % (a, v["ms_per_step"]) for a, v in k.items() if v["ms_per_step"] > 0.3), d["call_wall_ms_per_step"])
    except Exception as e:
        print(n, "failed", e)
P
