# round-1 refresh after pbwt_permute_v4: tests, smoke, bench, reference arm, ncu launch list + full captures
mkdir -p gpurun_out
T=${T:-r01c}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:scan_rows|build_wah|pbwt_|wah_|scan_u32|sparse_|pack_wah|compose_' --csv --log-file gpurun_out/${T}_launches.csv python bench.py --profile-only --blocks 32 > gpurun_out/${T}_launches.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:pbwt_permute|pbwt_unpermute|scan_rows|compose_simple' -s 4 -c 4 -o gpurun_out/${T}_top python bench.py --profile-only --blocks 32 > gpurun_out/${T}_top.out 2>&1
ls -la gpurun_out
tail -3 gpurun_out/${T}_pytest.log; tail -2 gpurun_out/${T}_smoke.log; cat gpurun_out/${T}_bench.json; cat gpurun_out/${T}_bench_ref.json
