cd $GRAFT_REPO_ROOT
timeout 600 python tools/_exp_k3.py > gpurun_out/exp_k3.log 2>&1; echo rc=$?
cat gpurun_out/exp_k3.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/b39.json 2> gpurun_out/b39.err; echo "bench rc=$?"
