# usage: bash tools/gpu/r02_yahoo_n.sh N  (run under gpurun --gpus N)
cd $GRAFT_REPO_ROOT
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --workload yahoo --steps 5 --warmup 3 --no-cpu-baseline --no-shim-e2e > gpurun_out/r02_bench_yahoo_k100_1gpu.json 2> gpurun_out/r02y_1.err; echo "yahoo1 rc=$?"
  timeout 1500 python bench.py --workload powerlaw --k 200 --steps 3 --warmup 2 --no-cpu-baseline --no-shim-e2e > gpurun_out/r02_bench_powerlaw_k200_1gpu.json 2> gpurun_out/r02p_1.err; echo "powerlaw1 rc=$?"
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --workload yahoo --steps 5 --warmup 3 > gpurun_out/r02_bench_yahoo_k100_${N}gpu.json 2> gpurun_out/r02y_$N.err; echo "yahoo$N rc=$?"
fi
tail -c 300 gpurun_out/r02y_$N.err
