# round 2: small-row permute kernel with batched lookups, unpermute KH/NC sweep, subset tool, capi_decode setup/teardown
mkdir -p gpurun_out
T=${T:-r02o}
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log
timeout 900 python -m pytest tests/test_bindings.py -m gpu -q -k "subset or plugin or ingest" > gpurun_out/${T}_bind.log 2>&1; echo "bind rc=$?" >> gpurun_out/${T}_bind.log; tail -5 gpurun_out/${T}_bind.log
show='import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print("value %.1f enc %.1f dec %.1f ms/step %.2f verified %s" % (d["value"], d["compress_ggts"], d["decompress_ggts"], d["ms_per_step"], d["verified"]))
print("  " + "  ".join("%s %.2f" % (k, v["ms_per_step"]) for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:9]))'
run() { echo "== $1 | $2"; env $1 timeout 600 python bench.py --sub --warmup 2 --steps 4 $2 2>/dev/null | python -c "$show"; }
{
run "XSI_X=0" "--samples 2504 --blocks 220"
run "XSI_PBWT_SMALL=0" "--samples 2504 --blocks 220"
run "XSI_X=0" "--samples 2504 --blocks 24 --shape chrx"
run "XSI_PBWT_SMALL=0" "--samples 2504 --blocks 24 --shape chrx"
run "XSI_UNPERM_KH=32 XSI_UNPERM_NC=256" "--blocks 32"
run "XSI_UNPERM_KH=32 XSI_UNPERM_NC=384" "--blocks 32"
run "XSI_UNPERM_KH=32 XSI_UNPERM_NC=512" "--blocks 32"
run "XSI_UNPERM_KH=16 XSI_UNPERM_NC=192" "--samples 2504 --blocks 220"
run "XSI_UNPERM_KH=16 XSI_UNPERM_NC=256" "--samples 2504 --blocks 220"
run "XSI_UNPERM_KH=8 XSI_UNPERM_NC=160" "--samples 2504 --blocks 220"
run "XSI_UNPERM_KH=8 XSI_UNPERM_NC=320" "--samples 2504 --blocks 24 --shape chrx"
run "XSI_UNPERM_KH=8 XSI_UNPERM_NC=160" "--samples 2504 --blocks 24 --shape chrx"
} 2>&1 | tee gpurun_out/${T}_shapes.txt
{
D=/dev/shm/capi; mkdir -p $D
bindings/_out/synth_bcf $D/in.bcf hrc 32488 16384 1002 16 > /dev/null
bindings/_out/xsi_b200_bcf compress $D/in.bcf $D/d.xsi --threads 16 --batch-blocks 1
for i in 1 2 3; do s=$(date +%s.%N); XSI_CAPI_NO_CHECKSUM=1 bindings/_out/capi_decode_b200 $D/d.xsi_var.bcf; e=$(date +%s.%N); echo "wall $(echo "$e - $s" | bc) s"; done
python -c "
import ctypes,time
t=time.time(); L=ctypes.CDLL('xsqueezeit_b200/libxsi_b200.so'); c=ctypes.c_void_p(); print('dlopen',time.time()-t)
t=time.time(); L.xsi_create(0,ctypes.byref(c)); print('create',time.time()-t)
t=time.time(); L.xsi_destroy(c); print('destroy',time.time()-t)"
rm -rf $D
} 2>&1 | tee gpurun_out/${T}_capi.txt
# priorities: who gets the SMs that free up when batch i+1 encodes beside the decode of batch i (pipelined leg = `value` here)
{
for rep in 1 2; do
run "XSI_X=0" "--blocks 32 --steps 8"
run "XSI_PERMUTE_PRIORITY=1" "--blocks 32 --steps 8"
run "XSI_ENC_STREAM_PRIORITY=1" "--blocks 32 --steps 8"
run "XSI_DEC_STREAM_PRIORITY=1" "--blocks 32 --steps 8"
run "XSI_PERMUTE_PRIORITY=1 XSI_UNPERM_KH=32" "--blocks 32 --steps 8"
done
} 2>&1 | tee gpurun_out/${T}_priority.txt
