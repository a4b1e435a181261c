set -x
cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --import-source on --clock-control none -k regex:^hoptraj_kernel -c 1 -f -o gpurun_out/k3_r33 python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu33.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/
