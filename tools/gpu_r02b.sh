# round 2, BCF legs on an HRC-shaped synthetic BCF: reference CLI / C API on the CPU vs the adapters and the ingest/egress tool
mkdir -p gpurun_out
T=${T:-r02b}
R=${R:-16384}
D=/dev/shm/xsi; mkdir -p $D
timeout 900 python -m pytest tests/test_bindings.py -m gpu -q -k "bcf_ingest" > gpurun_out/${T}_ingest_tests.log 2>&1; tail -12 gpurun_out/${T}_ingest_tests.log
( time bindings/_out/synth_bcf $D/hrc.bcf hrc 32488 $R 1002 16 ) 2>&1 | tail -4
ls -la $D
{
echo "== reference CLI -c"; ( time oracle/_ref/xsqueezeit_ref -c -f $D/hrc.bcf -o $D/ref.xsi ) 2>&1 | grep -v "^$" | tail -4
echo "== adapter CLI -c"; ( time bindings/_out/xsqueezeit_b200 -c -f $D/hrc.bcf -o $D/ada.xsi ) 2>&1 | grep -v "^$" | tail -4
for t in 4 8 16; do for k in 1 2; do echo "== ingest tool threads $t batch $k"; ( time bindings/_out/xsi_b200_bcf compress $D/hrc.bcf $D/ref.xsi.ing --threads $t --batch-blocks $k ) 2>&1 | grep "xsi_b200\|real"; done; done
cmp $D/ref.xsi $D/ada.xsi && echo "adapter .xsi identical"; cmp $D/ref.xsi $D/ref.xsi.ing && echo "ingest .xsi identical"
mkdir -p $D/i; bindings/_out/xsi_b200_bcf compress $D/hrc.bcf $D/i/ref.xsi --threads 8 > /dev/null; cmp $D/ref.xsi_var.bcf $D/i/ref.xsi_var.bcf && echo "var.bcf identical"
echo "== C API decode (ref, b200 x2, b200 without checksum, ref without checksum)"; bindings/_out/capi_decode_ref $D/ref.xsi_var.bcf; bindings/_out/capi_decode_b200 $D/ref.xsi_var.bcf; bindings/_out/capi_decode_b200 $D/ref.xsi_var.bcf
XSI_CAPI_NO_CHECKSUM=1 bindings/_out/capi_decode_b200 $D/ref.xsi_var.bcf; XSI_CAPI_NO_CHECKSUM=1 bindings/_out/capi_decode_ref $D/ref.xsi_var.bcf
echo "== -x to BCF: reference, adapter CLI, egress tool (4/8/16 threads)"; ( time oracle/_ref/xsqueezeit_ref -x -f $D/ref.xsi -o $D/o_ref.bcf ) 2>&1 | grep real
( time bindings/_out/xsqueezeit_b200 -x -f $D/ref.xsi -o $D/o_ada.bcf ) 2>&1 | grep real
for t in 4 8 16; do ( time bindings/_out/xsi_b200_bcf extract $D/ref.xsi $D/o_tool.bcf --threads $t ) 2>&1 | grep "xsi_b200\|real"; done
cmp $D/o_ref.bcf $D/o_ada.bcf && echo "-x adapter output identical"; cmp $D/o_ref.bcf $D/o_tool.bcf && echo "-x tool output identical"
( time bindings/_out/xsi_b200_bcf extract $D/ref.xsi $D/o_tool_u.bcf --threads 8 -O u ) 2>&1 | grep "xsi_b200\|real"
} > gpurun_out/${T}_bcf.txt 2>&1
cat gpurun_out/${T}_bcf.txt
rm -rf $D
