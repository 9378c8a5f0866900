// TEST INFRASTRUCTURE: a stand-in for the tail of the reference's fillDescriptors (src/SIFT_FeatureFactory.cu:533-548) written
// against include/ssrlcv_b200_packed.cuh, exactly as the documented patch of INTEGRATION.md does: one 4 x 4 x 8 block per
// feature, thread (x, y, z) owns descriptor byte (y * 4 + x) * 8 + z, the first thread stores location, sigma and theta.
// The values come from an AoS Feature<SIFT_Descriptor> array so that the same features can also go through the pack kernel.
#include "ssrlcv_b200_packed.cuh"

struct Feature152 { int parent; int pad; float2 loc; float sigma, theta; unsigned char values[128]; };
static_assert(sizeof(Feature152) == 152, "Feature<SIFT_Descriptor>");

__global__ void fill_descriptors_tail(unsigned int numFeatures, ssrlcv_b200_set_arrays packed, const Feature152* source) {
  const unsigned long blockId = blockIdx.y * gridDim.x + blockIdx.x;
  if (blockId >= numFeatures) return;
  const unsigned int k = (threadIdx.y * 4 + threadIdx.x) * 8 + threadIdx.z;
  ssrlcv_b200_packed_store_value(packed, blockId, k, source[blockId].values[k]);
  if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0)
    ssrlcv_b200_packed_store_keypoint(packed, blockId, source[blockId].loc, source[blockId].sigma, source[blockId].theta);
}

extern "C" int packed_writer_fill(const ssrlcv_b200_set_arrays* packed, const void* features_dev, unsigned long n, void* stream) {
  if (n == 0) return 0;
  const unsigned int gx = n < 65535 ? static_cast<unsigned int>(n) : 65535u;
  const dim3 grid(gx, static_cast<unsigned int>((n + gx - 1) / gx), 1), block(4, 4, 8);
  fill_descriptors_tail<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<unsigned int>(n), *packed,
                                                                              static_cast<const Feature152*>(features_dev));
  return static_cast<int>(cudaGetLastError());
}
