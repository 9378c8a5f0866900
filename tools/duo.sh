#!/bin/bash
# CTA-pair walk (smax_duo.cu): bring-up parity, A/B timing against the pair walk at C3, then the GPU suite with the walk forced
tag=${1:-duo}
mkdir -p gpurun_out
timeout 180 python tools/duo_bringup.py > gpurun_out/${tag}_bringup.txt 2>&1; echo "bringup rc $?" | tee -a gpurun_out/${tag}_bringup.txt
tail -5 gpurun_out/${tag}_bringup.txt
if grep -q "bringup rc 0" gpurun_out/${tag}_bringup.txt; then
  DIAG_PAIR=1,2 timeout 300 python tools/diag_ab.py C3 0 > gpurun_out/${tag}_ab.txt 2>&1; tail -8 gpurun_out/${tag}_ab.txt
  if [ "$2" == "tests" ]; then
    SSRLCV_TEST_PAIR=2 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${tag}_pytest.txt
  fi
fi
