cd $GRAFT_REPO_ROOT
timeout 600 python tools/timeline.py > gpurun_out/timeline.md 2> gpurun_out/timeline.err; echo rc=$?
cat gpurun_out/timeline.md; tail -3 gpurun_out/timeline.err
