mkdir -p gpurun_out
T=${T:-san3}
export PYTHONPATH=$PWD:$PWD/tests:$PWD/oracle
timeout 420 compute-sanitizer --tool memcheck --print-limit 30 --log-file gpurun_out/${T}_memcheck.log python -m pytest tests/test_gpu_parity.py -x -q -k "hrc_shape or kgp_shape or biallelic_ld or multiallelic" > gpurun_out/${T}_memcheck.out 2>&1
echo "memcheck rc=$?"; tail -n 2 gpurun_out/${T}_memcheck.out
grep -E "Invalid|at xsi::|at void xsi|ERROR SUMMARY|by thread" gpurun_out/${T}_memcheck.log | cut -c1-260 | head -40
ls -la gpurun_out/${T}_memcheck.log
