cd $GRAFT_REPO_ROOT
timeout 1200 python tools/class_api_bench.py > gpurun_out/class_api.json 2> gpurun_out/class_api.err; echo rc=$?
cat gpurun_out/class_api.json; tail -5 gpurun_out/class_api.err
