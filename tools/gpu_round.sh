mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r01_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r01_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_ref.json 2> gpurun_out/r01_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches.csv python bench.py --profile-only --blocks 8 > gpurun_out/r01_launches.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:pbwt_permute|pbwt_unpermute|scan_rows|compose_simple|wah_expand|sparse_index' -s 6 -c 6 -o gpurun_out/r01_top python bench.py --profile-only --blocks 8 > gpurun_out/r01_top.out 2>&1
ls -la gpurun_out
tail -3 gpurun_out/r01_pytest.log; cat gpurun_out/r01_smoke.log | tail -2; cat gpurun_out/r01_bench.json
