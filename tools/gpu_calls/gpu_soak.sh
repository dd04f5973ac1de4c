mkdir -p gpurun_out
python tools/soak.py 2048 3 2>&1 | grep -v "Task id" | tee gpurun_out/soak.log
for t in reach push block_stack; do
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/prof_one.py $t 64 3 > gpurun_out/sanitizer_$t.log 2>&1; echo "sanitizer $t rc=$?"; tail -2 gpurun_out/sanitizer_$t.log
done
