cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu_r02l.log
python tools/ab_stages.py fuse_gather=1 conv_cluster=2 conv_cluster=4 fuse_gather=0 --check > gpurun_out/ab_r02l.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:gnm:: -s 54 -c 90 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 5 --warmup 3 --no-module --cpu-sample 0 > gpurun_out/ncu_bench_r02.log 2>&1
python bench.py --steps 100 > gpurun_out/bench_r02_v4.json 2> gpurun_out/bench_r02_v4.err
tail -3 gpurun_out/pytest_gpu_r02l.log; cut -c1-260 gpurun_out/ab_r02l.log; head -c 300 gpurun_out/bench_r02_v4.json
