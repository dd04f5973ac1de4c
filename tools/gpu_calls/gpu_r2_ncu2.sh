mkdir -p gpurun_out
timeout 300 python tools/steady_time.py block_stack:2048 block_stack:256 push:4096 pick_and_place:4096 2>&1 | grep -v "Task id" | tee gpurun_out/r2_ninth_timing.txt
timeout 900 ncu --set full --import-source on --clock-control none -k regex:step_kernel_coop_multi --launch-skip 61 --launch-count 1 -f -o gpurun_out/r02_stack_full2 python tools/prof_steady.py block_stack 2048 3 2>&1 | tail -2
timeout 900 ncu --set full --import-source on --clock-control none -k regex:step_kernel_coop_block --launch-skip 61 --launch-count 1 -f -o gpurun_out/r02_push_full python tools/prof_steady.py push 4096 3 2>&1 | tail -2
