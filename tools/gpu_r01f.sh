# r01f: multi-context (full-duplex PCIe) host-buffer leg
mkdir -p gpurun_out
T=${T:-r01f}
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_b32.json 2> gpurun_out/${T}_bench_b32.err; echo "bench rc=$?"
tail -5 gpurun_out/${T}_bench_b32.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r01f_bench_b32.json').read().strip().splitlines()[-1])
print('value',d['value'],'verified',d['verified'])
for k in ('e2e','e2e_bcf_int8'):
    print(k, json.dumps(d[k]))
print(json.dumps(d['roofline']))
P
