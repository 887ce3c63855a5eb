cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest30.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest30.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r01f_1gpu.json 2> gpurun_out/r01f_1gpu.err; echo rc=$?; grep "e2e phases" gpurun_out/r01f_1gpu.err
python -c "
import json
d=json.loads(open('gpurun_out/r01f_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['cpu_baseline']['value'], d['clocks'])"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
