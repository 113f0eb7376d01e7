# compute-sanitizer over the GPU parity suite (final build of the round): memcheck, racecheck, synccheck, initcheck
# usage (GPU box): bash tools/gpu_sanitize.sh [tag]
mkdir -p gpurun_out
T=${1:-r02san}
SEL=${SEL:-"not biobank and not concurrent_contexts and not pinned_host and not max_uint16"}
for tool in memcheck racecheck synccheck; do
  ( time timeout 1500 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" ) > gpurun_out/${T}_${tool}.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/${T}_${tool}.log
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|real" gpurun_out/${T}_${tool}.log | tail -5
done
