cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "not powerlaw001" > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02a_bench.err
for cfg in "base:" "nowidx:PRIMALCR_EXP_NOWIDX=1" "ub12:PRIMALCR_UBLOCK_MB=12" "ub48:PRIMALCR_UBLOCK_MB=48" "nowidx_ub48:PRIMALCR_EXP_NOWIDX=1 PRIMALCR_UBLOCK_MB=48" "nowidx_ub96:PRIMALCR_EXP_NOWIDX=1 PRIMALCR_UBLOCK_MB=96"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python tools/stage_bench.py --side V --tag $tag >> gpurun_out/r02a_stage.jsonl 2>> gpurun_out/r02a_stage.err; echo "stage $tag rc=$?"
done
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"rowsum_kernel|dots_units_kernel" -s 2 -c 4 -o gpurun_out/r02a_yahoo04 python tools/profile_step.py --workload yahoo --scale 0.4 > gpurun_out/r02a_yahoo04.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/ | tail -12
