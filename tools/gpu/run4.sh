cd $GRAFT_REPO_ROOT
python tools/run_shapes.py --which powerlaw --powerlaw-scale 0.05 > gpurun_out/shapes_p05.jsonl 2> gpurun_out/shapes_p05.err
python tools/run_shapes.py --which powerlaw --powerlaw-scale 0.2 > gpurun_out/shapes_p20.jsonl 2> gpurun_out/shapes_p20.err
python - <<'PY'
import json
for f in ('gpurun_out/shapes_p05.jsonl','gpurun_out/shapes_p20.jsonl'):
    for l in open(f):
        d=json.loads(l)
        print(d['shape'],d['scale'],'nnz',d['nnz'],'d1',d['d1'],'d2',d['d2'],'s/iter',[round(x,3) for x in d['sec_per_iter']],'setup',d['setup_s'])
        for k in d['kernels_last_iter']: print('   ',k)
PY
