cd $GRAFT_REPO_ROOT
python tools/eval_probe.py --scale 1.0 > gpurun_out/eval_probe.json 2> gpurun_out/eval_probe.err; echo rc=$?
cat gpurun_out/eval_probe.json; tail -3 gpurun_out/eval_probe.err
