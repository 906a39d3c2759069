set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c_pytest.log
python bench/walk_sweep.py 1.0 > gpurun_out/r2c_sweep.log 2>&1
python bench/band_diag.py 1.0 > gpurun_out/r2c_diag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:locate_eval_kernel -s 2 -c 1 -o gpurun_out/r2c_le -f python bench/ncu_case.py 1.0 > gpurun_out/r2c_ncu1.log 2>&1
tail -4 gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_sweep.log gpurun_out/r2c_diag.log
