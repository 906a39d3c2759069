set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/fin_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin_smoke.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/fin_c4.json 2> gpurun_out/fin_c4.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/fin_c4_ref.json 2> gpurun_out/fin_c4_ref.err
for w in c2 c3 c5; do python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/fin_$w.json 2> gpurun_out/fin_$w.err; done
python bench/c1_latency.py > gpurun_out/fin_c1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/fin_c4_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/fin_ncu_c4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/fin_c2_launches.csv python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/fin_ncu_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/fin_c3_launches.csv python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/fin_ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:esrg_tile_kernel -s 2 -c 1 -o gpurun_out/prof_esrg4 python bench/phase_probe.py c3 1.0 > gpurun_out/fin_ncu_esrg.log 2>&1
cat gpurun_out/fin_pytest.log gpurun_out/fin_smoke.log gpurun_out/fin_c1.log
