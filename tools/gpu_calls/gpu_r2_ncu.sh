# Round 2 ncu captures (1 GPU): one steady-state launch of the block_stack(4) B=2048 kernel and of the Reach B=8192 kernel, --set full with source.
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:step_kernel_coop_multi --launch-skip 61 --launch-count 1 -f -o gpurun_out/r02_stack_full python tools/prof_steady.py block_stack 2048 3 2>&1 | tail -3
timeout 900 ncu --set full --import-source on --clock-control none -k regex:step_kernel_coop_reach --launch-skip 61 --launch-count 1 -f -o gpurun_out/r02_reach_full python tools/prof_steady.py reach 8192 3 2>&1 | tail -3
ls -la gpurun_out/*.ncu-rep
