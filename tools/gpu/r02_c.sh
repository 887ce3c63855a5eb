cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -k "not powerlaw001" > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02c_pytest.log
./tools/gpu/k6_ceiling > gpurun_out/r02c_k6_ceiling.jsonl 2>&1; echo "k6 rc=$?"; cat gpurun_out/r02c_k6_ceiling.jsonl
for cfg in "base:" "l2hint:PRIMALCR_ROWSUM_L2HINT=1"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python tools/stage_bench.py --side V --tag $tag >> gpurun_out/r02c_stage.jsonl 2>> gpurun_out/r02c_stage.err; echo "stage $tag rc=$?"
done
cat gpurun_out/r02c_stage.jsonl
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02c_bench2.json 2> gpurun_out/r02c_bench2.err; echo "bench2 rc=$?"
tail -c 300 gpurun_out/r02c_bench2.err
