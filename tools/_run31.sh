set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q --timeout 300 --timeout-method=thread -k "run_host or pipeline" > gpurun_out/t31.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t31.log
timeout 600 python -m pytest tests/test_gpu_api.py -m gpu -x -q --timeout 300 --timeout-method=thread -k "transport or graph or network" >> gpurun_out/t31.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t31.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/b31.json 2> gpurun_out/b31.err; echo "bench rc=$?" >> gpurun_out/t31.log
tail -5 gpurun_out/t31.log
