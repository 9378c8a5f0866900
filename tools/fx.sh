mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2m_pytest.log; tail -25 gpurun_out/r2m_pytest.log
timeout 900 python tools/fixture_paths.py > gpurun_out/r2m_fixture_paths.txt 2>gpurun_out/r2m_fixture_paths.err; cat gpurun_out/r2m_fixture_paths.txt; tail -3 gpurun_out/r2m_fixture_paths.err
