cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest26.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest26.log
python tools/eval_probe.py --scale 1.0 > gpurun_out/eval_probe2.json 2> gpurun_out/eval_probe2.err; echo rc=$?
cat gpurun_out/eval_probe2.json
