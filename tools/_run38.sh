cd $GRAFT_REPO_ROOT
timeout 600 python tools/_exp_k3.py > gpurun_out/exp_k3.log 2>&1; echo rc=$?
HOPB_K3_PIPE=1 timeout 600 python tools/_exp_k3.py > gpurun_out/exp_k3_pipe.log 2>&1; echo rc=$?
cat gpurun_out/exp_k3.log; echo ---- pipe; cat gpurun_out/exp_k3_pipe.log
