mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/r02_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1; grep -c "Race reported" gpurun_out/r02_sanitizer_racecheck.log; grep "Race reported" -A3 gpurun_out/r02_sanitizer_racecheck.log | grep "at \|Race" | sort | uniq -c | sort -rn | head -12; tail -2 gpurun_out/r02_sanitizer_racecheck.log
