# round 2: source-level ncu capture of pbwt_permute_v4 where it is latency-bound (short rows: 1KGP3 220 blocks, chrX 24 blocks)
mkdir -p gpurun_out
T=${T:-r02t}
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:pbwt_permute_v4' -s 1 -c 1 -o gpurun_out/${T}_kgp python bench.py --profile-only --samples 2504 --blocks 220 > gpurun_out/${T}_kgp.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:pbwt_permute_v4' -s 1 -c 1 -o gpurun_out/${T}_chrx python bench.py --profile-only --samples 2504 --blocks 24 --shape chrx > gpurun_out/${T}_chrx.out 2>&1
ls -la gpurun_out/${T}_*
