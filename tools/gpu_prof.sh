mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 2 -c 1 -o gpurun_out/prof_reach -f python tools/prof_one.py reach 8192 4 > gpurun_out/ncu_reach.log 2>&1
tail -3 gpurun_out/ncu_reach.log
