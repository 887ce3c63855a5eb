cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -q -k "powerlaw001" > gpurun_out/r02e_pytest_powerlaw.log 2>&1; echo "pytest powerlaw rc=$?"; tail -5 gpurun_out/r02e_pytest_powerlaw.log; grep -n "^E  " gpurun_out/r02e_pytest_powerlaw.log | head -5
for cfg in "hint0:PRIMALCR_ROWSUM_L2HINT=0" "hint1:PRIMALCR_ROWSUM_L2HINT=1" "hint2:PRIMALCR_ROWSUM_L2HINT=2"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python tools/stage_bench.py --side V --reps 6 --tag $tag >> gpurun_out/r02e_stage.jsonl 2>> gpurun_out/r02e_stage.err; echo "stage $tag rc=$?"
done
python - <<'PY'
import json
for l in open('gpurun_out/r02e_stage.jsonl'):
    d=json.loads(l); print(d['tag'], d['ms_per_launch'].get('rowsum_items'))
PY
python tools/dump_csr.py --workload netflix --dir /dev/shm/nf > gpurun_out/r02e_dump.log 2>&1
( cd /dev/shm && PRIMALCR_VERBOSE_SETUP=1 $GRAFT_REPO_ROOT/primalcr_b200/bin/primalcr-train -s 2 -k 100 -l 5000 -t 10 -p 0 -n 16 nf nf.model ) > gpurun_out/r02_cli_full_size.log 2>&1; echo "cli rc=$?"
ls -la /dev/shm/nf /dev/shm/nf.model /dev/shm/U.txt /dev/shm/V.txt >> gpurun_out/r02_cli_full_size.log 2>&1
rm -rf /dev/shm/nf /dev/shm/nf.model /dev/shm/U.txt /dev/shm/V.txt
grep "primalcr host\|Wall\|Iter 10" gpurun_out/r02_cli_full_size.log
bash tools/gpu/r02_yahoo_n.sh 1
