# r01g: int8 transport of host int32 rows: tests + default bench
mkdir -p gpurun_out
T=${T:-r01g}
nproc > gpurun_out/${T}_host.txt; lscpu | grep -i "model name\|^CPU(s)\|socket\|thread" >> gpurun_out/${T}_host.txt; free -g | head -2 >> gpurun_out/${T}_host.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/${T}_pytest.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_b32.json 2> gpurun_out/${T}_bench_b32.err; echo "bench rc=$?"
tail -5 gpurun_out/${T}_bench_b32.err
cat gpurun_out/${T}_host.txt
python - <<P
import json
d=json.loads(open('gpurun_out/${T}_bench_b32.json').read().strip().splitlines()[-1])
print('value',d['value'],'verified',d['verified'])
for k in ('e2e','e2e_bcf_int8'):
    print(k, json.dumps(d[k]))
print(json.dumps(d['roofline']))
P
