# round 2: parity + shapes after the scan changes (fused aux words for one-tile rows, trimmed epilogue), then memcheck of the new paths
mkdir -p gpurun_out
T=${T:-r02l}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -4 gpurun_out/${T}_pytest.log
show='import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print("value %.1f enc %.1f dec %.1f ms/step %.2f verified %s" % (d["value"], d["compress_ggts"], d["decompress_ggts"], d["ms_per_step"], d["verified"]))
print("  " + "  ".join("%s %.2f" % (k, v["ms_per_step"]) for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:8]))
print("  frac " + "  ".join("%s %.2f" % (r["kernel"], r["frac"]) for r in d["roofline_kernels"]))'
for shape in "--blocks 32" "--samples 2504 --blocks 220" "--samples 2504 --blocks 24 --shape chrx"; do
  echo "== $shape"
  timeout 600 python bench.py --sub --steps 4 --warmup 2 $shape 2>/dev/null | python -c "$show"
done 2>&1 | tee gpurun_out/${T}_shapes.txt
( time timeout 600 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "multiallelic or mixed_ploidy or golden_small or variants or haploid or lazy or dot_products or async" ) > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${T}_memcheck.log | tail -3
