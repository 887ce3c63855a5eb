cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_multi2.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_multi2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s2.json 2> gpurun_out/s2.err
echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'], d['objective'][-1])
PY
