mkdir -p gpurun_out
for cfg in "push 4096" "push 512"; do
  set -- $cfg; n=coop_$1$2
  ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -o /tmp/prof_$n -f python tools/prof_one.py $1 $2 8 > gpurun_out/ncu_$n.log 2>&1; tail -1 gpurun_out/ncu_$n.log
  ncu -i /tmp/prof_$n.ncu-rep --page raw --csv > gpurun_out/prof_${n}_raw.csv
  ncu -i /tmp/prof_$n.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/prof_${n}_source.csv
  ls -la /tmp/prof_$n.ncu-rep; cp /tmp/prof_$n.ncu-rep gpurun_out/
done
