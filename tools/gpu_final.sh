# round-1 final: tests, smoke, default bench, reference arm, ncu launch list and full captures of the top HRC kernels
mkdir -p gpurun_out
T=${T:-r01final}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/${T}_pytest.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:scan_rows|build_wah|pbwt_|wah_|scan_u32|sparse_|pack_wah|compose_' --csv --log-file gpurun_out/${T}_launches.csv python bench.py --profile-only --blocks 32 > gpurun_out/${T}_launches.out 2>&1; echo "launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:pbwt_permute|pbwt_unpermute|scan_rows|compose_simple' -s 4 -c 4 -o gpurun_out/${T}_top python bench.py --profile-only --blocks 32 > gpurun_out/${T}_top.out 2>&1; echo "ncu top rc=$?"
ls -la gpurun_out | grep ${T}
python - <<P
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'one ctx',d['value_one_context'],'verified',d['verified'], "enc %.1f dec %.1f" % (d["compress_ggts"], d["decompress_ggts"]))
for kk in ('e2e','e2e_bcf_int8'):
    e=dict(d[kk]); i=e.pop('int32_over_pcie',None); s=e.pop('serial',None); print(kk, round(e['value'],2), 'i32bus', i and round(i['value'],2))
print('ref', json.loads(open('gpurun_out/${T}_bench_ref.json').read().strip().splitlines()[-1])['value'])
P
