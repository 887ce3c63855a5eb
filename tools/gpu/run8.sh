cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s8.json 2> gpurun_out/s8.err
echo rc=$?
tail -3 gpurun_out/s8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s8.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'])
for k in d['roofline']['kernels']: print('  ',k['name'],round(k['ms_per_step'],3),k['launches_per_step'])
for r in d['per_rank']: print(r['nnz'],r['users'],round(r['sec'],4),{n:v for n,v in r['kernels'].items() if n in ('nccl_allreduce','rowsum_items','dots','hv_hv_look','tile_prepare','lm_sweep_hv')})
PY
