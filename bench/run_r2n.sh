cd $GRAFT_REPO_ROOT
GI_BARY_RASTER=1 python bench/walk_sweep.py 1.0 2>&1 | head -2 > gpurun_out/r2n_sweep.log 2>&1
cat gpurun_out/r2n_sweep.log
GI_BARY_RASTER=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_fuzz.py tests/test_gpu_f90_golden.py tests/test_gpu_pipeline.py -x -q 2>&1 | tail -5 > gpurun_out/r2n_pytest.log
cat gpurun_out/r2n_pytest.log
