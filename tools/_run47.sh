cd $GRAFT_REPO_ROOT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c5.csv python tools/config_sweep.py --configs 5 > gpurun_out/ncu47.log 2>&1; echo "rc=$?"
