for c in 1 2 4 8; do echo "== cluster $c"; XSI_PBWT_CLUSTER=$c timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3; done
for cfg in "32 0" "32 1" "32 2" "32 4" "8 1" "8 4" "8 8" "8 0" "4 0" "16 0"; do
  set -- $cfg
  if [ "$2" = "0" ]; then unset XSI_PBWT_CLUSTER; else export XSI_PBWT_CLUSTER=$2; fi
  timeout 600 python bench.py --blocks $1 --steps 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('$cfg', 'value %.1f enc %.1f permute %.2f ms verified %s' % (d['value'], d['compress_ggts'], k['pbwt_permute']['ms_per_step'], d['verified']))"
done
