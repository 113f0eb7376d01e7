# r01v: two-route (host conversion + plain DMA) transport of pinned int32 rows: tests + default bench
mkdir -p gpurun_out
T=${T:-r01v}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -6 gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -n 3 gpurun_out/${T}_bench.err
python - <<P
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'one ctx',d['value_one_context'],'verified',d['verified'])
for kk in ('e2e','e2e_bcf_int8'):
    e=dict(d[kk]); i=e.pop('int32_over_pcie',None); s=e.pop('serial',None); print(kk, json.dumps(e)[:500]); print('   serial', s); print('   i32bus', i and i['value'])
P
