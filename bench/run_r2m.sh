cd $GRAFT_REPO_ROOT
export GI_BARY_RASTER=1
ncu --set full --clock-control none --import-source on -k regex:raster_eval_kernel -s 2 -c 1 -o gpurun_out/r2m_raster -f python bench/ncu_case.py 1.0 > gpurun_out/r2m_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/r2m_launches.csv python bench/ncu_case.py 1.0 > gpurun_out/r2m_ncu2.log 2>&1
tail -3 gpurun_out/r2m_ncu1.log
