cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest29.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest29.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s1c.json 2> gpurun_out/s1c.err; echo rc=$?; grep "e2e phases" gpurun_out/s1c.err
