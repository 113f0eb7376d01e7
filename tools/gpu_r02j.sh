mkdir -p gpurun_out
T=${T:-r02j}
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:pbwt_permute_v5' -s 1 -c 1 -o gpurun_out/${T}_v5 python bench.py --sub --profile-only --blocks 32 > gpurun_out/${T}_ncu.out 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | grep ${T}
