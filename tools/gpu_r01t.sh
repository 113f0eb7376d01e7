# r01t: full GPU tests, default bench (HRC), biobank 2 blocks int32 and 4 blocks int8
mkdir -p gpurun_out
T=${T:-r01t}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/${T}_pytest.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
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
run biobank2 --samples 500000 --blocks 2 --steps 2 --warmup 1 --resident-contexts 0
run biobank4_i8 --samples 500000 --blocks 4 --elem 1 --steps 2 --warmup 1 --resident-contexts 0
python - <<P
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'one ctx',d['value_one_context'],'verified',d['verified'], "enc %.1f dec %.1f" % (d["compress_ggts"], d["decompress_ggts"]))
print("mt", d.get("resident_multi_context"))
for kk in ('e2e','e2e_bcf_int8'):
    e=dict(d[kk]); e.pop('int32_over_pcie',None); e.pop('serial',None); print(kk, json.dumps(e)[:300])
print("cpu", d["cpu_baseline"])
P
