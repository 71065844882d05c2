timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python scripts/san_case.py > gpurun_out/san_racecheck.log 2>&1; tail -4 gpurun_out/san_racecheck.log
timeout 600 compute-sanitizer --tool memcheck python scripts/san_case.py > gpurun_out/san_memcheck.log 2>&1; tail -3 gpurun_out/san_memcheck.log
