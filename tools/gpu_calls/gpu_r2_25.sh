mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_all.py 37 6 > gpurun_out/memcheck_all.log 2>&1; grep -E "finite|ERROR SUMMARY" gpurun_out/memcheck_all.log | tail -14
PMG_COOP=0 PMG_COOP_BLOCK=0 PMG_COOP_STACK=0 timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_all.py 37 5 > gpurun_out/memcheck_thread_per_env.log 2>&1; grep -E "ERROR SUMMARY" gpurun_out/memcheck_thread_per_env.log | tail -2
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_all.py 13 5 > gpurun_out/racecheck_all.log 2>&1; grep -E "RACECHECK SUMMARY" gpurun_out/racecheck_all.log | tail -2
