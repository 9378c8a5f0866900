#!/usr/bin/env bash
# Builds the UNMODIFIED reference MatchFactory (CUDA) for sm_100a, straight from the
# sources under /root/reference, plus oracle/ref_harness.cu, into oracle/_ref/libssrlcv_ref.so.
# Outputs only go to oracle/_ref/ (git-ignored; still shipped to the GPU box by gpurun).
# Recipe notes:
#   * png.h / jpeglib.h / jerror.h are only #included by include/io_util.hpp -> empty stubs
#   * -DFIX_THRUST_WARNING_H skips include/fix_thrust_warning.h (breaks with Thrust 2.x)
#   * -dc is mandatory: the reference relies on cross-TU device calls (ref: Makefile:126)
#   * same preprocessor configuration as the reference's CI: LOG_LEVEL=3 GEO_ORBIT=1
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${SSRLCV_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
  echo "build_ref: $REF not present; keeping whatever is in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/obj" "$OUT/stubs"
: > "$OUT/stubs/png.h"; : > "$OUT/stubs/jpeglib.h"; : > "$OUT/stubs/jerror.h"
ARCH="-gencode arch=compute_100a,code=sm_100a"
NVFLAGS="-std=c++14 -DFIX_THRUST_WARNING_H -I$OUT/stubs -I$REF/include -DLOG_LEVEL=3 -DGEO_ORBIT=1 $ARCH -Xcompiler -fPIC -w"
pids=()
for f in MatchFactory Feature matrix_util cuda_vec_util cuda_util Image; do
  if [ ! -f "$OUT/obj/$f.o" ] || [ "$REF/src/$f.cu" -nt "$OUT/obj/$f.o" ]; then
    nvcc $NVFLAGS -dc "$REF/src/$f.cu" -o "$OUT/obj/$f.o" &
    pids+=($!)
  fi
done
if [ ! -f "$OUT/obj/ref_harness.o" ] || [ "$HERE/ref_harness.cu" -nt "$OUT/obj/ref_harness.o" ]; then
  nvcc $NVFLAGS -dc "$HERE/ref_harness.cu" -o "$OUT/obj/ref_harness.o" &
  pids+=($!)
fi
if [ ! -f "$OUT/obj/Logger.o" ]; then
  g++ -std=c++14 -fPIC -w -I"$REF/include" -I/usr/local/cuda/include -DLOG_LEVEL=3 -c "$REF/src/Logger.cpp" -o "$OUT/obj/Logger.o" &
  pids+=($!)
fi
# the drop-in: our integration/MatchFactory_b200.cu compiled against the reference's headers, in place of src/MatchFactory.cu
ROOT="$(cd "$HERE/.." && pwd)"
if [ ! -f "$OUT/obj/MatchFactory_b200.o" ] || [ "$ROOT/integration/MatchFactory_b200.cu" -nt "$OUT/obj/MatchFactory_b200.o" ] || [ "$ROOT/include/ssrlcv_b200.h" -nt "$OUT/obj/MatchFactory_b200.o" ]; then
  # (the reference's Makefile compiles host code without -O, :21-27; the drop-in's own translation unit asks for -O2, as INTEGRATION.md says)
  nvcc $NVFLAGS -Xcompiler -O2 -I"$ROOT/include" -dc "$ROOT/integration/MatchFactory_b200.cu" -o "$OUT/obj/MatchFactory_b200.o" &
  pids+=($!)
fi
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
COMMON="$OUT/obj/Feature.o $OUT/obj/matrix_util.o $OUT/obj/cuda_vec_util.o $OUT/obj/cuda_util.o $OUT/obj/Image.o $OUT/obj/Logger.o $OUT/obj/ref_harness.o"
nvcc $ARCH -shared -Xcompiler -fPIC "$OUT/obj/MatchFactory.o" $COMMON -lcusolver -lcublas -lcuda -o "$OUT/libssrlcv_ref.so"
echo "built $OUT/libssrlcv_ref.so"
if [ -f "$ROOT/ssrlcv_b200/libssrlcv_match_b200.so" ]; then
  nvcc $ARCH -shared -Xcompiler -fPIC "$OUT/obj/MatchFactory_b200.o" $COMMON -L"$ROOT/ssrlcv_b200" -lssrlcv_match_b200 \
       -Xlinker -rpath -Xlinker '$ORIGIN/../../ssrlcv_b200' -lcusolver -lcublas -lcuda -o "$OUT/libssrlcv_dropin.so"
  echo "built $OUT/libssrlcv_dropin.so"
fi
