// hensel_launch.hpp -- host-side launcher of decrypt_hensel_kernel, compiled in
// its own translation unit (hensel_decrypt.cu).
//
// Why a separate unit: the schedule loop of the two-digit CRT decrypt is ~30 KB
// of straight-line IMAD.WIDE code, close to the 32 KB instruction cache, and
// ptxas produced measurably different code for it (spilled pointers, 92 -> 98 ms
// per 65536 ciphertexts) whenever an unrelated kernel was added to the one big
// translation unit of the library.  Compiled alone, without --split-compile,
// its code depends on this kernel's sources only.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

#include "kernel_hensel_decrypt.cuh"

namespace ipclb200 {

struct HenselDecryptPlan {
  int grid = 0;              // blocks of kBlockThreads threads
  size_t smem = 0;           // dynamic shared memory per block
  size_t groups = 0;         // resident (ciphertext, side) tasks = table slots
  void (*launch)(const DecryptHenselParams&, int grid, size_t smem, cudaStream_t s) = nullptr;
  const char* name = "";
};

// layout: 0 default, 1 / 2 lane-spread (small batches), -1 thread-per-task.
// Returns cudaSuccess, or cudaErrorInvalidValue for an unsupported prime width.
cudaError_t hensel_decrypt_plan(int pl_words, int layout, int rows, bool w64, size_t count,
                                int sms, int want_blocks, HenselDecryptPlan* out);

}  // namespace ipclb200
