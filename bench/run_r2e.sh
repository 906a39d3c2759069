set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -30 > gpurun_out/r2e_pytest.log
ncu --set full --clock-control none --import-source on -k regex:locate_eval_kernel -s 2 -c 1 -o gpurun_out/r2e_band0 -f python bench/ncu_band.py 0 8 > gpurun_out/r2e_ncu0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:locate_eval_kernel -s 2 -c 1 -o gpurun_out/r2e_band4 -f python bench/ncu_band.py 4 8 > gpurun_out/r2e_ncu4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/r2e_band0_launches.csv python bench/ncu_band.py 0 8 > gpurun_out/r2e_ncu0b.log 2>&1
tail -5 gpurun_out/r2e_pytest.log
