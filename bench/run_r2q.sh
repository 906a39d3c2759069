cd $GRAFT_REPO_ROOT
python bench/walk_sweep.py 1.0 2>&1 | head -2 > gpurun_out/r2q_sweep.log 2>&1
GI_BARY_RASTER=1 python bench/walk_sweep.py 1.0 2>&1 | head -1 >> gpurun_out/r2q_sweep.log 2>&1
python bench/walk_sweep.py 0.5 2>&1 | head -1 >> gpurun_out/r2q_sweep.log 2>&1
GI_BARY_RASTER=1 python bench/walk_sweep.py 0.5 2>&1 | head -1 >> gpurun_out/r2q_sweep.log 2>&1
cat gpurun_out/r2q_sweep.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2q_pytest.log
cat gpurun_out/r2q_pytest.log
