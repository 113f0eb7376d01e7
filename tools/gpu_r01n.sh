# r01n: ring-release fix in the inverse-PBWT kernel, host-side parallelisation, multi-context resident leg
mkdir -p gpurun_out
T=${T:-r01n}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/${T}_pytest.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 --resident-contexts 3 > gpurun_out/${T}_bench_b32.json 2> gpurun_out/${T}_bench_b32.err; echo "bench rc=$?"
tail -n 3 gpurun_out/${T}_bench_b32.err
python - <<P
import json
d=json.loads(open('gpurun_out/${T}_bench_b32.json').read().strip().splitlines()[-1]); k=d["kernels"]
print('value',d['value'],'verified',d['verified'], "enc %.1f dec %.1f" % (d["compress_ggts"], d["decompress_ggts"]))
print(" ".join("%s %.2f" % (a, v["ms_per_step"]) for a, v in k.items() if v["ms_per_step"] > 0.3))
print("wall", {a: round(b, 2) for a, b in d["call_wall_ms_per_step"].items()})
print("mt", d.get("resident_multi_context"))
for kk in ('e2e','e2e_bcf_int8'):
    e=dict(d[kk]); e.pop('int32_over_pcie',None); print(kk, json.dumps(e)[:600])
print(json.dumps(d['roofline']))
P
