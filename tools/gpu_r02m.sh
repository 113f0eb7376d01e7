# round 2: bindings suite incl. the plugin-interface tests (shim built with both adapters), parity suite, HRC check of the small-kernel prefetches
mkdir -p gpurun_out
T=${T:-r02m}
timeout 1500 python -m pytest tests/test_bindings.py -m gpu -q -x -k "plugin" > gpurun_out/${T}_plugin.log 2>&1; echo "plugin rc=$?" >> gpurun_out/${T}_plugin.log; tail -15 gpurun_out/${T}_plugin.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log
show='import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print("value %.1f enc %.1f dec %.1f ms/step %.2f verified %s" % (d["value"], d["compress_ggts"], d["decompress_ggts"], d["ms_per_step"], d["verified"]))
print("  " + "  ".join("%s %.2f" % (k, v["ms_per_step"]) for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:9]))'
for shape in "--blocks 32" "--samples 2504 --blocks 220"; do
  echo "== $shape"
  timeout 600 python bench.py --sub --steps 4 --warmup 2 $shape 2>/dev/null | python -c "$show"
done 2>&1 | tee gpurun_out/${T}_shapes.txt
