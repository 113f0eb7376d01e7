# r01o: official numbers after the ring-release fix: bench (driver defaults), reference arm, ncu launch list (HRC), ncu full of the biobank kernels
mkdir -p gpurun_out
T=${T:-r01o}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_smi.txt
timeout 900 python bench.py --resident-contexts 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:scan_rows|build_wah|pbwt_|wah_|scan_u32|sparse_|pack_wah|compose_' --csv --log-file gpurun_out/${T}_launches.csv python bench.py --profile-only --blocks 32 > gpurun_out/${T}_launches.out 2>&1; echo "launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:pbwt_permute_grid|pbwt_unpermute_wide|wah_expand' -c 3 -o gpurun_out/${T}_biobank python bench.py --profile-only --samples 500000 --blocks 2 > gpurun_out/${T}_biobank.out 2>&1; echo "ncu biobank rc=$?"
ls -la gpurun_out | head -30
python - <<P
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]); k=d["kernels"]
print('value',d['value'],'verified',d['verified'], "enc %.1f dec %.1f" % (d["compress_ggts"], d["decompress_ggts"]))
print("mt", d.get("resident_multi_context"))
for kk in ('e2e','e2e_bcf_int8'):
    e=dict(d[kk]); e.pop('int32_over_pcie',None); e.pop('serial',None); print(kk, json.dumps(e)[:400])
print(open('gpurun_out/${T}_bench_ref.json').read()[:600])
P
