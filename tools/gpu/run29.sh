cd $GRAFT_REPO_ROOT
timeout 130 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "test_update_V_then_U and ragged" > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -6 gpurun_out/san_memcheck.log
timeout 130 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "test_update_V_then_U and ragged" > gpurun_out/san_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -6 gpurun_out/san_racecheck.log
