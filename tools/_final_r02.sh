cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu_r02z.log
python __graft_entry__.py smoke > gpurun_out/smoke_r02.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > gpurun_out/sanitizer_memcheck_r02.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/sanitizer_racecheck_r02.log 2>&1
python bench.py --steps 100 > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err
tail -2 gpurun_out/pytest_gpu_r02z.log; tail -2 gpurun_out/smoke_r02.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok" gpurun_out/sanitizer_*_r02.log; head -c 300 gpurun_out/bench_r02_final.json
