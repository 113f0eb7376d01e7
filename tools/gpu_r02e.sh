# round 2: short-row CTAs (scan_rows_v2 / compose_simple as wide as the row): parity + before/after on the 1KGP3 and chrX shapes
mkdir -p gpurun_out
T=${T:-r02e}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log
show='import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print("value %.1f enc %.1f dec %.1f ms/step %.2f verified %s" % (d["value"], d["compress_ggts"], d["decompress_ggts"], d["ms_per_step"], d["verified"]))
print("  " + "  ".join("%s %.2f" % (k, v["ms_per_step"]) for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:8]))
print("  frac " + "  ".join("%s %.2f" % (r["kernel"], r["frac"]) for r in d["roofline_kernels"]))
print("  host", {k: round(v, 1) for k, v in d["call_wall_ms_per_step"].items()})'
for nt in 256 0; do
  for shape in "--samples 2504 --blocks 220" "--samples 2504 --blocks 24 --shape chrx"; do
    echo "== NT=$nt (0 = row-wide CTAs) $shape"
    if [ $nt = 256 ]; then export XSI_SCAN_NT=256 XSI_COMPOSE_NT=256; else unset XSI_SCAN_NT XSI_COMPOSE_NT; fi
    timeout 600 python bench.py --sub --steps 3 --warmup 2 $shape 2>/dev/null | python -c "$show"
  done
done 2>&1 | tee gpurun_out/${T}_short_rows.txt
unset XSI_SCAN_NT XSI_COMPOSE_NT
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:scan_rows_v2|compose_simple' -s 2 -c 2 -o gpurun_out/${T}_kgp python bench.py --sub --profile-only --samples 2504 --blocks 220 > gpurun_out/${T}_ncu.out 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | grep ${T}
