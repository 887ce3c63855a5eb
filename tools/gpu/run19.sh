cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "level_count" > gpurun_out/pytest25.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest25.log
