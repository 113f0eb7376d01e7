# round 2: whole GPU suite + smoke + default bench (all legs) + reference arm
mkdir -p gpurun_out
T=${T:-r02c}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_smi.txt
free -g | head -2 >> gpurun_out/${T}_smi.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -15 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/${T}_pytest.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_pytest.log
tail -2 gpurun_out/${T}_pytest.log
( time timeout 1500 python bench.py ) > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/${T}_bench.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"; tail -4 gpurun_out/${T}_bench_ref.err
python - <<P
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'one ctx',d['value_one_context'],'verified',d['verified'], "enc %.1f dec %.1f" % (d["compress_ggts"], d["decompress_ggts"]))
print('e2e', d['e2e'] and d['e2e']['value'], 'roofline', d['roofline'] and (d['roofline']['kernel'], d['roofline']['frac']))
print('e2e_bcf', json.dumps(d.get('e2e_bcf'))[:1800])
print('shapes', json.dumps(d.get('shapes'))[:2500])
r=json.loads(open('gpurun_out/${T}_bench_ref.json').read().strip().splitlines()[-1])
print('ref', r['value'], r['cpu_baseline'], json.dumps(r.get('e2e_bcf'))[:600])
P
