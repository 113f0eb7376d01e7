# r01l: host-side parallelisation (decode parse/tables, encode layout): tests + HRC + 1KGP3 + biobank resident legs
mkdir -p gpurun_out
T=${T:-r01l}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
run() {  # name args...
  n=$1; shift
  timeout 900 python bench.py "$@" --no-e2e --no-cpu-baseline > gpurun_out/${T}_bench_$n.json 2> gpurun_out/${T}_bench_$n.err; echo "$n rc=$?"
  python - <<P
import json
try:
    d=json.loads(open('gpurun_out/${T}_bench_$n.json').read().strip().splitlines()[-1]); k=d["kernels"]
    print("$n", "value %.1f enc %.1f dec %.1f verified %s | " % (d["value"], d["compress_ggts"], d["decompress_ggts"], d["verified"]) + " ".join("%s %.2f" % (a, v["ms_per_step"]) for a, v in k.items() if v["ms_per_step"] > 0.3))
    print("   wall", {a: round(b, 2) for a, b in d["call_wall_ms_per_step"].items()}, "host", {a: round(b["ms_per_step"], 2) for a, b in d.get("host_phases", {}).items()})
except Exception as e:
    print("$n failed", e)
P
}
run hrc --steps 5 --warmup 3
run kgp --samples 2504 --blocks 220 --steps 3 --warmup 2
run biobank --samples 500000 --blocks 2 --steps 2 --warmup 1
