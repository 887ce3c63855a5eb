cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r01f_8gpu.json 2> gpurun_out/r01f_8gpu.err
echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r01f_8gpu.json').read().strip().splitlines()[-1])
print('N=8 value',d['value'],'e2e',d['e2e']['value'],d['e2e']['phases_s'], d['objective'][-1])
PY
