cd $GRAFT_REPO_ROOT
timeout 900 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke62.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke62.log
timeout 1200 python -m pytest tests -x -q -m gpu --timeout 600 --timeout-method=thread > gpurun_out/t62.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t62.log
