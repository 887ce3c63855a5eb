cd $GRAFT_REPO_ROOT
python bench.py --steps 3 --warmup 3 > gpurun_out/r01f_1gpu.json 2> gpurun_out/r01f_1gpu.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/r01f_1gpu.json').read().strip().splitlines()[-1])
r=d['roofline']; print(d['value'], d['e2e']['value'], r['kernel'], r['achieved'], r['frac'], r.get('traffic'), r.get('dram_achieved'), r.get('dram_frac'), d['cpu_baseline']['value'])"
