mkdir -p gpurun_out
timeout 600 python tools/fixture_timing.py > gpurun_out/r2q_fixture_timing.txt 2>&1; tail -25 gpurun_out/r2q_fixture_timing.txt
timeout 600 python tools/fixture_exhaustive.py > gpurun_out/r2q_fixture_exhaustive.txt 2>&1; tail -12 gpurun_out/r2q_fixture_exhaustive.txt
