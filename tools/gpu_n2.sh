# two ranks on one box: the multi-GPU path of bench.py (NCCL all-gather of block sizes, one ordered file written by both ranks and
# compared with the single writer's, NUMA binding, the biobank sub-run on every rank)
mkdir -p gpurun_out
T=${T:-r02n2}
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1; lscpu | grep -i "numa\|socket\|^CPU(s)" >> gpurun_out/${T}_topo.txt
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus ${NG:-2} --steps 5 --warmup 3 ) > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/${T}_bench.err
python - <<P
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],d['value_mode'],'one ctx',d['value_one_context'],'verified',d['verified'])
print('e2e', d['e2e'] and d['e2e']['value'], 'sharded', d['sharded_file'], 'numa', d['numa_binding'])
print('shapes', json.dumps(d.get('shapes'))[:900])
P
cat gpurun_out/${T}_topo.txt | head -20
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus ${NG:-2} --steps 2 --warmup 1 ) > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"; tail -3 gpurun_out/${T}_bench_ref.err; cut -c1-300 gpurun_out/${T}_bench_ref.json
