# experiment: PBWT permute on fewer SMs (cluster size 2 / 1) with several contexts, so that HBM-bound kernels of one context
# run beside the chain of another
mkdir -p gpurun_out
T=${T:-r02d}
for c in 4 2 1; do for w in 3 4; do
  echo "== XSI_PBWT_CLUSTER=$c contexts=$w"
  XSI_PBWT_CLUSTER=$c timeout 600 python bench.py --no-e2e --no-cpu-baseline --bcf-records 0 --no-shapes --resident-contexts $w --steps 4 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('one ctx %.1f (%.1f ms)  mt %.1f (%.1f ms) verified %s  permute %.2f ms unpermute %.2f ms' % (d['value_one_context'], d['ms_per_step_one_context'], d['resident_multi_context']['value'], d['resident_multi_context']['ms_per_step'], d['verified'], d['kernels']['pbwt_permute']['ms_per_step'], d['kernels']['pbwt_unpermute']['ms_per_step']))"
done; done 2>&1 | tee gpurun_out/${T}_cluster_contexts.txt
