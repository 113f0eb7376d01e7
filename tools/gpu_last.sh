mkdir -p gpurun_out
T=${T:-r01last}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/${T}_pytest.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'one ctx',d['value_one_context'],'verified',d['verified'], "enc %.1f dec %.1f" % (d["compress_ggts"], d["decompress_ggts"]))
for kk in ('e2e','e2e_bcf_int8'):
    e=dict(d[kk]); i=e.pop('int32_over_pcie',None); s=e.pop('serial',None); print(kk, round(e['value'],2), e['verified'])
print('cpu', d['cpu_baseline']['value'], 'launches', d['gpu_launches'], d['clocks'])
P
