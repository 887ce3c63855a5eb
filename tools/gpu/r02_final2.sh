cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02final2_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02final2_pytest.log
