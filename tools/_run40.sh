cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/multi_gpu_check.py > gpurun_out/mg40.log 2>&1; echo "check rc=$?"
tail -5 gpurun_out/mg40.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/b40_2gpu.json 2> gpurun_out/b40.err; echo "bench rc=$?"
tail -c 600 gpurun_out/b40_2gpu.json
