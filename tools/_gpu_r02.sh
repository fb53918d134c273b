cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu_r02n.log
python tools/ab_stages.py fuse_gather=1 fuse_gather=0 --check --steps 40 > gpurun_out/ab_r02n.log 2>&1
tail -3 gpurun_out/pytest_gpu_r02n.log; cut -c1-260 gpurun_out/ab_r02n.log
