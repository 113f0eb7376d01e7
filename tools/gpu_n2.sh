# two-GPU sanity of the bench contract (weak scaling, NCCL all-gather of block sizes)
mkdir -p gpurun_out
T=${T:-r01u}
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/${T}_smi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/${T}_bench_n2.err
cat gpurun_out/${T}_bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref_n2.json 2> gpurun_out/${T}_bench_ref_n2.err; echo "ref n2 rc=$?"
head -c 400 gpurun_out/${T}_bench_ref_n2.json; nproc
