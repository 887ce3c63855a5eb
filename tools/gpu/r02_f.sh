cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
bash tools/gpu/r02_yahoo_n.sh 4
bash tools/gpu/r02_yahoo_n.sh 2
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -k "powerlaw001 or sharded or own_cli" > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02f_pytest.log; grep -n "^E  " gpurun_out/r02f_pytest.log | head -5
