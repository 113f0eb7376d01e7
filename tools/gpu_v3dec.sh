# decode v3 (register-resident positions) parity + sweep; encode v4 cluster sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() {
  echo "== $*"
  env "$@" timeout 400 python bench.py --blocks ${BLOCKS:-32} --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>gpurun_out/sweep_err.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); k=d['kernels']; print('value %.1f enc %.1f dec %.1f permute %.2f unpermute %.2f ms verified %s' % (d['value'], d['compress_ggts'], d['decompress_ggts'], k['pbwt_permute']['ms_per_step'], k['pbwt_unpermute']['ms_per_step'], d['verified']))
except Exception as e: print('failed', e)"
  tail -2 gpurun_out/sweep_err.log
}
run XSI_UNPERM_KH=16 XSI_PBWT_CLUSTER=4 XSI_PBWT_KH=16
run XSI_UNPERM_KH=32 XSI_PBWT_CLUSTER=4 XSI_PBWT_KH=32
run XSI_UNPERM_KH=8 XSI_PBWT_CLUSTER=8 XSI_PBWT_KH=32
run XSI_UNPERM_KH=16 XSI_UNPERM_NC=256 XSI_PBWT_CLUSTER=2 XSI_PBWT_KH=32
run XSI_UNPERM_V2=1 XSI_PBWT_CLUSTER=4 XSI_PBWT_KH=8
BLOCKS=8 run XSI_UNPERM_KH=8
BLOCKS=8 run XSI_UNPERM_KH=16
