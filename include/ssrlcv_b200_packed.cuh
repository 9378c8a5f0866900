/*
 * ssrlcv_b200_packed.cuh -- device-side writer for the packed feature layout of libssrlcv_match_b200.so.
 *
 * The reference's SIFT feature factory ends by storing each descriptor into a 152-byte AoS record
 * (fillDescriptors, src/SIFT_FeatureFactory.cu:533-548: one 4x4x8 block per feature, thread (x, y, z) owns byte
 * (y * 4 + x) * 8 + z), which the matcher cannot read with TMA (stride 152 = 8 mod 16) and therefore repacks.  A producer
 * that includes this header writes the packed arrays directly and the pack pass disappears:
 *
 *   host:    ssrlcv_b200_set_alloc(ctx, numFeatures, &set);  ssrlcv_b200_set_arrays_of(set, &arrays);
 *            fillDescriptors<<<grid, block>>>(..., arrays, ...);          // patched kernel, see INTEGRATION.md
 *            ssrlcv_b200_set_commit(ctx, set);                            // norms, fold digits, window bounds: one launch
 *   device:  ssrlcv_b200_packed_store_value(arrays, blockId, (threadIdx.y * 4 + threadIdx.x) * 8 + threadIdx.z, value);
 *            if (first thread) ssrlcv_b200_packed_store_keypoint(arrays, blockId, kp.loc * pixelWidth, kp.sigma, kp.theta);
 *
 * Rows are 128 bytes, so the 128 threads of a block store one fully coalesced line instead of 128 bytes scattered inside a
 * 152-byte record.  Plain C++/CUDA, no dependency on the library's internals.
 */
#ifndef SSRLCV_B200_PACKED_CUH
#define SSRLCV_B200_PACKED_CUH

#include <cuda_runtime.h>

#include "ssrlcv_b200.h"

/* byte `k` (0..127) of the descriptor of feature `row` */
__device__ __forceinline__ void ssrlcv_b200_packed_store_value(const ssrlcv_b200_set_arrays& a, unsigned long row, unsigned int k,
                                                               unsigned char value) {
  a.desc[row * 128ul + k] = value;
}
/* location (absolute pixel coordinates) and sigma / theta of feature `row` (Feature<SIFT_Descriptor>::loc,
 * SIFT_Descriptor::sigma / theta, include/Feature.cuh:30-38,68-94) */
__device__ __forceinline__ void ssrlcv_b200_packed_store_keypoint(const ssrlcv_b200_set_arrays& a, unsigned long row, float2 loc,
                                                                  float sigma, float theta) {
  reinterpret_cast<float2*>(a.loc)[row] = loc;
  reinterpret_cast<float2*>(a.sigma_theta)[row] = make_float2(sigma, theta);
}

#endif /* SSRLCV_B200_PACKED_CUH */
