# round 2: whole GPU suite after the lazy chain / dot products / per-block haploid split / kernel clean-up
mkdir -p gpurun_out
T=${T:-r02k}
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -30 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/${T}_pytest.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_pytest.log; tail -2 gpurun_out/${T}_pytest.log
D=/dev/shm/xsi; mkdir -p $D
bindings/_out/synth_bcf $D/hrc.bcf hrc 32488 16384 1002 16 2>/dev/null
bindings/_out/xsi_b200_bcf compress $D/hrc.bcf $D/h.xsi --threads 16 --batch-blocks 1 | tail -1
for i in 1 2; do XSI_CAPI_NO_CHECKSUM=1 bindings/_out/capi_decode_b200 $D/h.xsi_var.bcf; done
( time bindings/_out/xsqueezeit_b200 -x -r 20:10000-90000 -f $D/h.xsi -o $D/r.bcf ) 2>&1 | grep real
( time oracle/_ref/xsqueezeit_ref -x -r 20:10000-90000 -f $D/h.xsi -o $D/r2.bcf ) 2>&1 | grep real
cmp $D/r.bcf $D/r2.bcf && echo "region extract identical"
rm -rf $D
