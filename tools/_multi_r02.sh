cd $GRAFT_REPO_ROOT
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/multigpu_check.py > gpurun_out/r02_multigpu_check_$N.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/bench_r02_n$N.json 2> gpurun_out/bench_r02_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 30 --warmup 3 --config 3 > gpurun_out/bench_r02_cfg3_n$N.json 2> gpurun_out/bench_r02_cfg3_n$N.err
grep -v "^W\|Warning\|warn\|^\*\|OMP_NUM" gpurun_out/r02_multigpu_check_$N.log | tail -12; tail -2 gpurun_out/bench_r02_n$N.err; head -c 250 gpurun_out/bench_r02_n$N.json; echo; head -c 250 gpurun_out/bench_r02_cfg3_n$N.json
