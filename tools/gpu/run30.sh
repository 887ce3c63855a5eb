cd $GRAFT_REPO_ROOT
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --workload yahoo --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/yahoo_4gpu.json 2> gpurun_out/yahoo_4gpu.err
echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/yahoo_4gpu.json').read().strip().splitlines()[-1])
print('yahoo N=4 value',d['value'],'e2e',d['e2e']['value'], d['config']['workload'][:60], d['objective'])
for k in d['roofline']['kernels'][:10]: print('  ',k['name'],round(k['ms_per_step'],2),k['launches_per_step'])
PY
