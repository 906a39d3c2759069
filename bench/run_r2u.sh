cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "morton" 2>&1 | tail -12 > gpurun_out/r2u_pytest.log
cat gpurun_out/r2u_pytest.log
timeout 900 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_c5.json 2> gpurun_out/r2u_c5.err
tail -c 1500 gpurun_out/r2u_c5.json; tail -5 gpurun_out/r2u_c5.err
