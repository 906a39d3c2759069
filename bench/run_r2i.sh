cd $GRAFT_REPO_ROOT
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2i_scale2.json 2> gpurun_out/r2i_scale2.err
tail -c 4000 gpurun_out/r2i_scale2.json; tail -20 gpurun_out/r2i_scale2.err
