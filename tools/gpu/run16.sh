cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_multi3.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_multi3.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s2b.json 2> gpurun_out/s2b.err
echo rc=$?; grep "e2e phases" gpurun_out/s2b.err
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s1b.json 2> gpurun_out/s1b.err
echo rc=$?; grep "e2e phases" gpurun_out/s1b.err
python - <<'PY'
import json
for f in ('gpurun_out/s2b.json','gpurun_out/s1b.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f,'value',d['value'],'e2e',d['e2e']['value'], d['objective'][-1])
PY
