cd $GRAFT_REPO_ROOT
ncu --set full --clock-control none --import-source on -k regex:wv_gather -s 4 -c 1 -o gpurun_out/prof_wvg python tools/ab_stages.py fuse_gather=1 --steps 2 > gpurun_out/ncu_wvg.log 2>&1
tail -3 gpurun_out/ncu_wvg.log
