cd $GRAFT_REPO_ROOT
python bench.py --steps 3 --warmup 3 > gpurun_out/r01f_1gpu.json 2> gpurun_out/r01f_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r01f_reference.json 2> gpurun_out/r01f_reference.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01f_bench_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r01f_bench_under_ncu.json 2> gpurun_out/r01f_bench_under_ncu.err; echo "ncu rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r01f_1gpu.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline'],'clocks',d['clocks'])
r=json.loads(open('gpurun_out/r01f_reference.json').read().strip().splitlines()[-1])
print('ref',r['value'])
PY
wc -l gpurun_out/r01f_bench_launches.csv
