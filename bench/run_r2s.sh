cd $GRAFT_REPO_ROOT
free -g | head -2; nproc
timeout 1200 python -m pytest tests/test_gpu_large.py tests/test_gpu_edge_cases.py -x -q --durations=8 2>&1 | tail -20 > gpurun_out/r2s_pytest.log
cat gpurun_out/r2s_pytest.log
