# round 2, first GPU call: the reference's own programs on the GPU path (bindings tests), the existing parity suite,
# decode timing through c_api.h (CPU reference vs GPU adapter), a baseline bench line
mkdir -p gpurun_out
T=${T:-r02a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_smi.txt
nproc > gpurun_out/${T}_host.txt; lscpu | head -20 >> gpurun_out/${T}_host.txt
timeout 1200 python -m pytest tests/test_bindings.py -m gpu -q > gpurun_out/${T}_bindings.log 2>&1; echo "bindings rc=$?" >> gpurun_out/${T}_bindings.log
tail -25 gpurun_out/${T}_bindings.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
# c_api.h decode loop: CPU reference vs GPU adapter on chr20_small (2,504 samples x 21,892 records)
D=/tmp/capi; mkdir -p $D
bindings/_out/xsqueezeit_b200 -c -f tests/golden/inputs/chr20_small.bcf -o $D/c.xsi > /dev/null 2>&1
for i in 1 2 3; do bindings/_out/capi_decode_ref $D/c.xsi_var.bcf; bindings/_out/capi_decode_b200 $D/c.xsi_var.bcf; done > gpurun_out/${T}_capi.txt 2>&1
( time bindings/_out/xsqueezeit_b200 -c -f tests/golden/inputs/chr20_small.bcf -o $D/c2.xsi ) >> gpurun_out/${T}_capi.txt 2>&1
( time oracle/_ref/xsqueezeit_ref -c -f tests/golden/inputs/chr20_small.bcf -o $D/c3.xsi ) >> gpurun_out/${T}_capi.txt 2>&1
( time bindings/_out/xsqueezeit_b200 -x -f $D/c.xsi -o $D/o1.bcf ) >> gpurun_out/${T}_capi.txt 2>&1
( time oracle/_ref/xsqueezeit_ref -x -f $D/c.xsi -o $D/o2.bcf ) >> gpurun_out/${T}_capi.txt 2>&1
cat gpurun_out/${T}_capi.txt | grep -v "^$"
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/${T}_bench.json
