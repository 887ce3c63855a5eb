cd $GRAFT_REPO_ROOT
python tools/run_shapes.py --which yahoo,powerlaw --yahoo-scale 1.0 --powerlaw-scale 1.0 > gpurun_out/shapes_full2.jsonl 2> gpurun_out/shapes_full2.err; echo rc=$?
python tools/run_shapes.py --which ml1m_pcr,ml1m_pcrpp,yahoo,powerlaw > gpurun_out/shapes2.jsonl 2> gpurun_out/shapes2.err; echo rc=$?
python - <<'PY'
import json
for f in ('gpurun_out/shapes_full2.jsonl','gpurun_out/shapes2.jsonl'):
    for l in open(f):
        d=json.loads(l)
        print(d['shape'],d['scale'],'k',d['k'],'nnz',d['nnz'],'s/iter',[round(x,4) for x in d['sec_per_iter']],'GB',round(d['device_gb'],1),'mono',d['monotone'],'recomp',d['recomputed_rel_err'], d.get('other_solver_rel_err'))
PY
