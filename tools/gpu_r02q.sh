# round 2: one-barrier epilogue of the narrow scan CTAs (short rows): parity, racecheck of the scan-heavy tests, 1KGP3 / chrX / HRC timing
mkdir -p gpurun_out
T=${T:-r02q}
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log
timeout 900 python -m pytest tests/test_bindings.py -m gpu -q -x > gpurun_out/${T}_bind.log 2>&1; echo "bind rc=$?" >> gpurun_out/${T}_bind.log; tail -3 gpurun_out/${T}_bind.log
show='import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
p=d.get("one_context_pipelined") or {}
print("value %.1f (%s) pipelined %.1f one-ctx %.1f enc %.1f dec %.1f ms/step %.2f verified %s" % (d["value"], d.get("value_mode","")[:24], p.get("value",0), d["value_one_context"], d["compress_ggts"], d["decompress_ggts"], d["ms_per_step"], d["verified"]))
print("  " + "  ".join("%s %.2f" % (k, v["ms_per_step"]) for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:9]))
print("  frac " + "  ".join("%s %.2f" % (r["kernel"], r["frac"]) for r in d["roofline_kernels"][:6]))'
run() { echo "== $1 | $2"; env $1 timeout 600 python bench.py --sub --warmup 2 --steps 4 $2 2>/dev/null | python -c "$show"; }
{
run "XSI_X=0" "--samples 2504 --blocks 220"
run "XSI_X=0" "--samples 2504 --blocks 24 --shape chrx"
run "XSI_X=0" "--blocks 32 --steps 6"
} 2>&1 | tee gpurun_out/${T}_shapes.txt
SEL="golden or multiallelic or missing or ploidy or haploid or ragged or single_record or kgp or int8 or variants"
for tool in racecheck memcheck; do
  ( time timeout 1200 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" ) > gpurun_out/${T}_${tool}.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/${T}_${tool}.log
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|real" gpurun_out/${T}_${tool}.log | tail -4
done
grep -c "scan_rows" gpurun_out/${T}_racecheck.log
