set -x
export PATH=$PATH:/usr/local/cuda/bin
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/sanitize_case.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/sanitize_case.py > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -8 gpurun_out/sanitize_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 3 python profiles/sanitize_case.py > gpurun_out/sanitize_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -5 gpurun_out/sanitize_synccheck.log
