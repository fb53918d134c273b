cd $GRAFT_REPO_ROOT
python tools/parity_sweep.py > gpurun_out/parity_sweep_r02c.log 2>&1
python tools/batch_sweep.py > gpurun_out/r02_batch_sweep.md 2> gpurun_out/batch_sweep.err
python bench.py --config 1 > gpurun_out/bench_r02_cfg1.json 2> gpurun_out/bench_r02_cfg1.err
python bench.py --config 4 > gpurun_out/bench_r02_cfg4.json 2> gpurun_out/bench_r02_cfg4.err
tail -3 gpurun_out/parity_sweep_r02c.log; cat gpurun_out/r02_batch_sweep.md | cut -c1-200; head -c 400 gpurun_out/bench_r02_cfg1.json; echo; head -c 400 gpurun_out/bench_r02_cfg4.json; tail -2 gpurun_out/batch_sweep.err gpurun_out/bench_r02_cfg4.err
