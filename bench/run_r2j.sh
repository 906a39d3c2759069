cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2j_scale8.json 2> gpurun_out/r2j_scale8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2j_scale4.json 2> gpurun_out/r2j_scale4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2j_ref8.json 2> gpurun_out/r2j_ref8.err
tail -c 3500 gpurun_out/r2j_scale8.json; tail -5 gpurun_out/r2j_scale8.err; tail -c 1500 gpurun_out/r2j_scale4.json; tail -c 900 gpurun_out/r2j_ref8.json
