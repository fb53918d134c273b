#!/bin/bash
# The GPU-box command lines behind the round-2 evidence in profiles/ (run with `gpurun [--gpus N] -- 'bash tools/gpurun_suite.sh <what> [N]'`).
#   validate : GPU tests, smoke, compute-sanitizer memcheck + racecheck, default bench line
#   profile  : ncu launch list of 5 forward steps + `--set full` capture of the three big kernels
#   ab       : same-box A/B of library options with per-stage CUDA events (tools/ab_stages.py)
#   configs  : parity sweep, batch sweep (config 5), module runs of configs 1 and 4
#   multi N  : torchrun x N: sharded-vs-single-GPU parity check, bench configs 2 and 3
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
what=${1:-validate}; N=${2:-2}
case "$what" in
  validate)
    python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log
    python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
    timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > gpurun_out/sanitizer_memcheck.log 2>&1
    timeout 900 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/sanitizer_racecheck.log 2>&1
    python bench.py --steps 100 > gpurun_out/bench.json 2> gpurun_out/bench.err
    tail -2 gpurun_out/pytest_gpu.log; grep -hE "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok" gpurun_out/sanitizer_*.log gpurun_out/smoke.log; head -c 300 gpurun_out/bench.json ;;
  profile)
    ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:gnm:: -s 54 -c 90 --csv \
        --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-module --cpu-sample 0 > gpurun_out/ncu_bench.log 2>&1
    ncu --set full --clock-control none --import-source on -k regex:"wv_gather|conv_t_kernel|embed_conv1" -s 6 -c 5 -o gpurun_out/prof \
        python tools/ab_stages.py fuse_gather=1 --steps 2 > gpurun_out/ncu_full.log 2>&1
    tail -2 gpurun_out/ncu_full.log ;;
  ab)
    python tools/ab_stages.py fuse_gather=1 fuse_gather=0 --check --steps 30 --wvg-cycles > gpurun_out/ab.log 2>&1; cut -c1-260 gpurun_out/ab.log ;;
  configs)
    python tools/parity_sweep.py > gpurun_out/parity_sweep.log 2>&1
    python tools/batch_sweep.py > gpurun_out/batch_sweep.md 2> gpurun_out/batch_sweep.err
    python bench.py --config 1 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err
    python bench.py --config 4 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
    tail -3 gpurun_out/parity_sweep.log; cut -c1-200 gpurun_out/batch_sweep.md; head -c 400 gpurun_out/bench_cfg1.json; echo; head -c 400 gpurun_out/bench_cfg4.json ;;
  multi)
    TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
    $TR --master-port 29517 tools/multigpu_check.py > gpurun_out/multigpu_check_$N.log 2>&1
    $TR --master-port 29518 bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
    $TR --master-port 29519 bench.py --gpus $N --steps 30 --warmup 3 --config 3 > gpurun_out/bench_cfg3_n$N.json 2> gpurun_out/bench_cfg3_n$N.err
    grep -v "^W\|Warning\|warn\|^\*\|OMP_NUM" gpurun_out/multigpu_check_$N.log | tail -12; head -c 250 gpurun_out/bench_n$N.json; echo; head -c 250 gpurun_out/bench_cfg3_n$N.json ;;
  *) echo "usage: $0 validate|profile|ab|configs|multi [N]"; exit 2 ;;
esac
