cd $GRAFT_REPO_ROOT
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"rowsum_kernel|dots_units_kernel|tile_lm_sweep_kernel" -s 4 -c 5 -o gpurun_out/r01f_top python tools/profile_step.py --scale 1.0 > gpurun_out/r01f_top.log 2>&1; echo "ncu1 rc=$?"
ncu --profile-from-start off --set full --clock-control none -k regex:"tile_prepare_kernel" -c 3 -o gpurun_out/r01f_prep python tools/profile_step.py --scale 1.0 > gpurun_out/r01f_prep.log 2>&1; echo "ncu2 rc=$?"
ncu --profile-from-start off --set full --clock-control none -k regex:"hv_chunk|hv_lookup|rowsum_finalize|u_cg_step" -s 3 -c 6 -o gpurun_out/r01f_hv python tools/profile_step.py --scale 1.0 > gpurun_out/r01f_hv.log 2>&1; echo "ncu3 rc=$?"
ls -la gpurun_out/*.ncu-rep
