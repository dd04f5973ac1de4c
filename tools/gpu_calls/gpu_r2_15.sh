mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -x 2>&1 | tail -6
timeout 600 python tools/steady_time.py reach:8192 push:4096 pick_and_place:4096 slide:4096 block_stack:2048 2>&1 | grep -v "Task id" | tee gpurun_out/r2_15_timing.txt
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum
for tb in reach:8192 push:4096 block_stack:2048; do
  t=${tb%%:*}; b=${tb##*:}
  timeout 600 ncu --metrics $M --clock-control none -k regex:step_kernel -s 62 -c 3 --csv --log-file gpurun_out/ncu_traffic2_${t}_$b.csv python tools/prof_steady.py $t $b 5 > /dev/null 2>&1
done
