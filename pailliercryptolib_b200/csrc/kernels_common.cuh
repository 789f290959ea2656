// kernels_common.cuh -- what every batch kernel of the library shares.
#pragma once
#include <cstdint>

#include "mont_core.cuh"

namespace ipclb200 {

constexpr int kBlockThreads = 128;

// Dynamic work distribution: a warp claims the next chunk of 32/T consecutive
// elements from a global counter (zeroed by the host before the launch).  The
// batch is only a few waves of the resident groups, so a static split leaves
// whole SMs idle during the last wave; with claims the tail is shared.
__device__ __forceinline__ unsigned int claim_chunk(unsigned int* counter) {
  unsigned int w = 0;
  if ((threadIdx.x & 31) == 0) w = atomicAdd(counter, 1u);
  return __shfl_sync(IPCLB200_FULL_MASK, w, 0);
}

}  // namespace ipclb200
