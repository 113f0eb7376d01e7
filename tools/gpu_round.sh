# Round-end measurement pass (one GPU): tests, smoke, both bench arms, launch list, one full capture of the top kernels.
# usage: gpurun --timeout 3000 -- 'bash tools/gpu_round.sh r02final'   (FULL=1: also the ncu --set full capture, ~8 GPU-minutes)
T=${1:-r02final}
O=gpurun_out/$T
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
nproc > $O/host.txt; free -g >> $O/host.txt; numactl -H >> $O/host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 1200 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
# launch list of the hot path only (-k: the synthetic-input generator is 35,000 torch launches that are not ours)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:scan_rows|build_wah|pbwt_|wah_|scan_u32|sparse_|pack_wah|compose_|select_samples|allele_counts|dot_' --csv --log-file $O/launches.csv python bench.py --profile-only --blocks 32 > $O/launches.out 2>&1
python tools/launches_summary.py $O/launches.csv > $O/launches_summary.txt 2>&1
[ "${FULL:-0}" = 1 ] && timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:pbwt_permute|pbwt_unpermute|scan_rows|compose_simple|wah_expand|wah_encode' -s 6 -c 6 -o $O/top python bench.py --profile-only --blocks 32 > $O/top.out 2>&1
tail -3 $O/pytest.log; tail -2 $O/smoke.log; cat $O/bench.json; cat $O/bench_ref.json; head -12 $O/launches_summary.txt
