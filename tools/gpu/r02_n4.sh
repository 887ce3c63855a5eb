cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for N in 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2962$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_netflix_k100_${N}gpu.json 2> gpurun_out/r02n4_$N.err; echo "netflix $N rc=$?"
done
python - <<'PY'
import json
for n in (4,2):
    d=json.loads(open("gpurun_out/r02_bench_netflix_k100_%dgpu.json"%n).read().strip().splitlines()[-1]); print(n, d["value"], d["objective"][-1])
PY
