mkdir -p gpurun_out
for cfg in "512 4" "2048 2" "4096 1" "8192 1"; do
  set -- $cfg
  timeout 600 python bench.py --samples 500000 --block-len $1 --blocks $2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bis_$1.json 2> gpurun_out/bis_$1.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/bis_$1.json").read().strip().splitlines()[-1])
    print("block_len $1 blocks $2 verified", d["verified"], "value %.1f"%d["value"], d["config"]["wah_lines"], d["config"]["binary_lines"])
except Exception as e:
    print("block_len $1 failed", e)
P
done
