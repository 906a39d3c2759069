set -x
cd $GRAFT_REPO_ROOT
python bench/band_diag.py 1.0 > gpurun_out/r2b_diag.log 2>&1
python bench/walk_sweep.py 1.0 > gpurun_out/r2b_sweep.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:locate_eval_kernel -s 2 -c 1 -o gpurun_out/r2b_seg -f python bench/ncu_case.py 1.0 0 > gpurun_out/r2b_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:locate_eval_rows_kernel -s 2 -c 1 -o gpurun_out/r2b_rows -f python bench/ncu_case.py 1.0 1 > gpurun_out/r2b_ncu2.log 2>&1
cat gpurun_out/r2b_diag.log gpurun_out/r2b_sweep.log; tail -3 gpurun_out/r2b_ncu1.log gpurun_out/r2b_ncu2.log
