mkdir -p gpurun_out
TASK=${1:-reach}; B=${2:-8192}; SKIP=${3:-2}; N=${4:-4}
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s $SKIP -c 1 -o gpurun_out/prof_$TASK -f python tools/prof_one.py $TASK $B $N $5 > gpurun_out/ncu_$TASK.log 2>&1
tail -2 gpurun_out/ncu_$TASK.log
