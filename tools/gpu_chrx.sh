mkdir -p gpurun_out
T=${T:-r01y}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --shape chrx --samples 2504 --blocks 24 --steps 3 --warmup 2 > gpurun_out/${T}_bench_chrx.json 2> gpurun_out/${T}_bench_chrx.err; echo "chrx rc=$?"
tail -n 3 gpurun_out/${T}_bench_chrx.err
python - <<P
import json
d=json.loads(open('gpurun_out/${T}_bench_chrx.json').read().strip().splitlines()[-1]); k=d["kernels"]
print("chrx value %.1f enc %.1f dec %.1f verified %s | " % (d["value"], d["compress_ggts"], d["decompress_ggts"], d["verified"]) + " ".join("%s %.2f" % (a, v["ms_per_step"]) for a, v in k.items() if v["ms_per_step"] > 0.05))
print("   wall", {a: round(b, 2) for a, b in d["call_wall_ms_per_step"].items()})
P
