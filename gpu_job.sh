mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
tail -4 gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/sanitizer_initcheck.log 2>&1; echo "initcheck rc=$?" >> gpurun_out/sanitizer_initcheck.log
tail -4 gpurun_out/sanitizer_initcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.log
tail -4 gpurun_out/sanitizer_racecheck.log
