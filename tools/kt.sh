mkdir -p gpurun_out
for f in 1 0; do KT_FOLD=$f python tools/kerneltiming.py C3 > gpurun_out/r2d_kt_fold$f.txt 2>&1; tail -20 gpurun_out/r2d_kt_fold$f.txt; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2d_pytest.log; tail -30 gpurun_out/r2d_pytest.log
