# round 2: device timeline of the pipelined leg (encode of batch i+1 beside the decode of batch i), ordered and unordered
mkdir -p gpurun_out
T=${T:-r02r}
for o in 1 0 L; do
  export XSI_BENCH_PIPE_ORDER=load_first; if [ $o = L ]; then export XSI_BENCH_PIPE_ORDER=launch_first; fi
  rm -f /tmp/tl.txt
  XSI_OVERLAP_ORDER=$([ $o = 0 ] && echo 0 || echo 1) XSI_TIMELINE=/tmp/tl.txt timeout 600 python bench.py --sub --warmup 2 --steps 4 --blocks 32 > gpurun_out/${T}_bench_order$o.json 2>/dev/null
  cp /tmp/tl.txt gpurun_out/${T}_raw_order$o.txt
  python - <<PY > gpurun_out/${T}_timeline_order$o.txt
import json
d=json.loads(open("gpurun_out/${T}_bench_order$o.json").read().strip().splitlines()[-1])
print("pipelined", d["one_context_pipelined"]["value"], d["one_context_pipelined"]["ms_per_step"])
reads=open("/tmp/tl.txt").read().split("# read\n")
# the pipelined leg is the read with the most spans that holds both scan_rows and compose_simple
best=[r for r in reads if "scan_rows" in r and "compose_simple" in r][-1]  # the last such read is the pipelined leg
rows=[l.split() for l in best.strip().splitlines()]
big=[(n,float(a),float(b)) for n,a,b in rows if float(b)-float(a)>0.3]
big.sort(key=lambda x:x[1])
for n,a,b in big: print("%-18s %9.2f -> %9.2f  (%6.2f ms)" % (n,a,b,b-a))
PY
  cat gpurun_out/${T}_timeline_order$o.txt | head -26
done
