(timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2)
for cfg in "32 32 16" "32 16 16" "32 8 16" "32 16 8" "8 16 16" "8 8 16" "8 8 8"; do
  set -- $cfg
  XSI_UNPERM_WPW=$2 XSI_UNPERM_WARPS=$3 timeout 600 python bench.py --blocks $1 --steps 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('$cfg', 'value %.1f dec %.1f unperm %.2f ms expand %.2f verified %s' % (d['value'], d['decompress_ggts'], k['pbwt_unpermute']['ms_per_step'], k['wah_expand']['ms_per_step'], d['verified']))"
done
