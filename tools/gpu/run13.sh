cd $GRAFT_REPO_ROOT
rm -f gpurun_out/trace1.txt
PRIMALCR_TRACE=$PWD/gpurun_out/trace1.txt python tools/ab.py --tag trace --warmup 2 --iters 1 > gpurun_out/ab13.jsonl 2> gpurun_out/ab13.err
wc -l gpurun_out/trace1.txt
