mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -x 2>&1 | tail -6
{
echo "== state tiles (4 envs)"; timeout 600 python tools/steady_time.py reach:8192 push:4096 pick_and_place:4096 slide:4096 block_stack:2048 reach:8190 2>&1 | grep -v "Task id"
echo "== plain [word][env] arrays"; PMG_STATE_TILE=0 timeout 600 python tools/steady_time.py reach:8192 push:4096 block_stack:2048 2>&1 | grep -v "Task id"
echo "== thread-per-env kernels (32-env tiles)"; PMG_COOP=0 PMG_COOP_BLOCK=0 PMG_COOP_STACK=0 timeout 600 python tools/steady_time.py reach:8192 push:4096 2>&1 | grep -v "Task id"
} | tee gpurun_out/r2_14_timing.txt
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum
for tb in reach:8192 block_stack:2048; do
  t=${tb%%:*}; b=${tb##*:}
  timeout 600 ncu --metrics $M --clock-control none -k regex:step_kernel -s 62 -c 3 --csv --log-file gpurun_out/ncu_traffic_${t}_$b.csv python tools/prof_steady.py $t $b 5 > /dev/null 2>&1
  PMG_STATE_TILE=0 timeout 600 ncu --metrics $M --clock-control none -k regex:step_kernel -s 62 -c 3 --csv --log-file gpurun_out/ncu_traffic_plain_${t}_$b.csv python tools/prof_steady.py $t $b 5 > /dev/null 2>&1
done
grep -h "dram__bytes\|lts__t_sectors" gpurun_out/ncu_traffic_*.csv | cut -d, -f5,13-15 | head -60
