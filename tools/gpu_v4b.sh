mkdir -p gpurun_out
unset XSI_PBWT_CLUSTER
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for cfg in ${CFGS:-"4 16" "8 16" "2 32"}; do
  set -- $cfg
  unset XSI_PBWT_CLUSTER XSI_PBWT_KH XSI_PBWT_V3
  if [ "$1" = "v3" ]; then export XSI_PBWT_V3=1; elif [ "$1" != "0" ]; then export XSI_PBWT_CLUSTER=$1; fi
  if [ "$2" != "0" ]; then export XSI_PBWT_KH=$2; fi
  timeout 300 python bench.py --blocks ${BLOCKS:-32} --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>gpurun_out/v4_err_$1_$2.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); k=d['kernels']; print('cfg $cfg', 'value %.1f enc %.1f dec %.1f permute %.2f ms verified %s' % (d['value'], d['compress_ggts'], d['decompress_ggts'], k['pbwt_permute']['ms_per_step'], d['verified']))
except Exception as e: print('cfg $cfg failed', e)"
done
