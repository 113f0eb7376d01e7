mkdir -p gpurun_out
T=${T:-e2e}
for envs in "XSI_DMA_SLOTS=2" "XSI_DMA_SLOTS=4" "XSI_DMA_SLOTS=8"; do
env $envs timeout 600 python bench.py --blocks 8 --steps 3 --warmup 2 --resident-contexts 0 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$? ($envs)"
python - <<P
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
e=dict(d['e2e']); i=e.pop('int32_over_pcie',None); s=e.pop('serial',None)
print('e2e', round(e['value'],2), 'bus GB/step', round(e['h2d_bytes_per_step']/1e9,2), round(e['d2h_bytes_per_step']/1e9,2), 'verified', e['verified'])
print('   serial', {k: round(v,2) if isinstance(v,float) else v for k,v in s.items()})
P
done
