cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q --timeout 300 --timeout-method=thread > gpurun_out/t53.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t53.log
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_fullsize.py -m gpu -x -q --timeout 600 --timeout-method=thread >> gpurun_out/t53.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t53.log
grep -a "passed\|failed\|rc=\|Error" gpurun_out/t53.log | head -20
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/b53.json 2> gpurun_out/b53.err; echo "bench rc=$?"
tail -5 gpurun_out/b53.err
