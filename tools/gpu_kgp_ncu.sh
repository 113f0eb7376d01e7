mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:scan_rows|compose_simple' -s 2 -c 2 -o gpurun_out/kgp_top python bench.py --profile-only --samples 2504 --blocks 220 > gpurun_out/kgp_top.out 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/kgp_top.ncu-rep
