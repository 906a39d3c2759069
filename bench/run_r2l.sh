cd $GRAFT_REPO_ROOT
for v in 0 1; do echo "RASTER=$v"; GI_BARY_RASTER=$v python bench/walk_sweep.py 1.0 2>&1 | head -2; done > gpurun_out/r2l_sweep.log 2>&1
cat gpurun_out/r2l_sweep.log
GI_BARY_RASTER=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2l_pytest.log
cat gpurun_out/r2l_pytest.log
