cd $GRAFT_REPO_ROOT
python bench/walk_sweep.py 1.0 2>&1 | head -2 > gpurun_out/r2o_sweep.log 2>&1
cat gpurun_out/r2o_sweep.log
GI_BARY_RASTER=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2o_pytest1.log
cat gpurun_out/r2o_pytest1.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2o_pytest.log
cat gpurun_out/r2o_pytest.log
