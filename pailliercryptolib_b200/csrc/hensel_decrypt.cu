// hensel_decrypt.cu -- the instantiations of decrypt_hensel_kernel (kernels.cuh
// K4h, mont_hensel.cuh) and their launch plan; see hensel_launch.hpp for why
// this is a translation unit of its own.  Reference path:
// PrivateKey::decryptCRT, ipcl/pri_key.cpp:114-157.
#include "hensel_launch.hpp"

#include <cstdio>
#include <cstdlib>

namespace ipclb200 {
namespace {

template <int K, int T, int MINB, int ROWS, bool W64, int BT, bool COMPACT>
void launch_one(const DecryptHenselParams& p, int grid, size_t smem, cudaStream_t s) {
  decrypt_hensel_kernel<K, T, MINB, ROWS, W64, BT, COMPACT><<<grid, BT, smem, s>>>(p);
}
// grid = the blocks that are resident at once (persistent kernel, work is
// claimed chunk by chunk), at most max_blocks per SM
template <typename Kern>
cudaError_t plan_kernel(Kern kern, size_t smem, int T, int BT, int max_blocks, size_t count,
                        int sms, HenselDecryptPlan* out) {
  cudaError_t e =
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  // the per-task staging areas want most of the SM's unified L1/shared memory
  e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                           (int)cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BT, smem);
  if (e != cudaSuccess) return e;
  if (getenv("IPCLB200_DEBUG_PLAN")) {
    cudaFuncAttributes fa{};
    cudaFuncGetAttributes(&fa, kern);
    fprintf(stderr, "[plan] %s: regs %d static smem %zu dyn smem %zu BT %d -> %d blocks/SM (cap %d)\n",
            out->name, fa.numRegs, fa.sharedSizeBytes, smem, BT, per_sm, max_blocks);
  }
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  if (per_sm > max_blocks) per_sm = max_blocks;
  const size_t gpb = BT / T;
  const size_t chunks = 2 * ((count + (32 / T) - 1) / (32 / T));
  const size_t wpb = BT / 32;  // warps per block, one chunk at a time each
  const size_t need = (chunks + wpb - 1) / wpb;
  const size_t cap = (size_t)per_sm * (size_t)sms;
  out->grid = (int)(need < cap ? need : cap);
  out->smem = smem;
  out->groups = (size_t)out->grid * gpb;
  return cudaSuccess;
}

template <int K, int T, int MINB, int ROWS, bool W64, bool COMPACT = false>
cudaError_t plan_one(size_t count, int sms, int want_blocks, const char* name,
                     HenselDecryptPlan* out) {
  constexpr int BT = kBlockThreads;
  out->launch = launch_one<K, T, MINB, ROWS, W64, BT, COMPACT>;
  out->name = name;
  return plan_kernel(decrypt_hensel_kernel<K, T, MINB, ROWS, W64, BT, COMPACT>,
                     hensel_smem_bytes<K, T, COMPACT>(BT), T, BT,
                     want_blocks < MINB ? want_blocks : MINB, count, sms, out);
}
}  // namespace

#define PLAN(K_, T_, MINB_, ROWS_, W64_)                                              \
  return plan_one<K_, T_, MINB_, ROWS_, W64_>(count, sms, want_blocks,                \
                                              "decrypt_hensel_kernel<" #K_ "," #T_    \
                                              "," #MINB_ "," #ROWS_ "," #W64_ ">",    \
                                              out)
#define PLANC(K_, T_, MINB_, ROWS_, W64_)                                                \
  return plan_one<K_, T_, MINB_, ROWS_, W64_, true>(                                     \
      count, sms, want_blocks,                                                           \
      "decrypt_hensel_kernel<" #K_ "," #T_ "," #MINB_ "," #ROWS_ "," #W64_ ",compact>", out)

cudaError_t hensel_decrypt_plan(int pl, int layout, int rows, bool w64, size_t count, int sms,
                                int want_blocks, HenselDecryptPlan* out) {
  switch (pl) {
    case 16:
      if (layout >= 1) PLAN(4, 4, 3, 8, false);
      PLAN(8, 2, 3, 8, false);
    case 32:
      if (layout >= 2) PLAN(4, 8, 3, 8, false);
      if (layout == 1) PLAN(8, 4, 3, 8, false);
      if (layout == -2) PLANC(32, 1, 3, 4, true);
      if (layout == -1) PLAN(32, 1, 2, 4, true);  // with the prefetch buffer: 8 warps per SM
      if (rows == 4) PLAN(16, 2, 3, 4, false);
      if (w64) PLAN(16, 2, 3, 8, true);
      PLAN(16, 2, 3, 8, false);
    case 48:
      if (layout >= 1) PLAN(12, 4, 3, 8, false);
      PLAN(24, 2, 2, 8, false);
    case 64:
      if (layout >= 1) PLAN(8, 8, 3, 8, false);
      PLAN(16, 4, 3, 8, false);
    default:
      return cudaErrorInvalidValue;
  }
}
#undef PLAN
#undef PLANC

}  // namespace ipclb200
