// common.hpp -- tuning constants and the host RNG entry point.
// Constants keep the reference's names (ipcl/include/ipcl/utils/common.hpp:
// 15-25) for source compatibility; the QAT/hybrid ones only feed the hybrid
// knobs of mod_exp.hpp, which do not change routing here.
#ifndef IPCL_B200_UTILS_COMMON_HPP_
#define IPCL_B200_UTILS_COMMON_HPP_

#include <vector>

#include "ipcl/bignum.h"

namespace ipcl {

constexpr int IPCL_CRYPTO_MB_SIZE = 8;
constexpr int IPCL_QAT_MODEXP_BATCH_SIZE = 1024;
constexpr int IPCL_WORKLOAD_SIZE_THRESHOLD = 128;
constexpr float IPCL_HYBRID_MODEXP_RATIO_FULL = 1.0;
constexpr float IPCL_HYBRID_MODEXP_RATIO_ENCRYPT = 0.25;
constexpr float IPCL_HYBRID_MODEXP_RATIO_DECRYPT = 0.12;
constexpr float IPCL_HYBRID_MODEXP_RATIO_MULTIPLY = 0.18;

// fills `addr` with random 32-bit words (OS entropy)
void rand32u(std::vector<Ipp32u>& addr);

// uniformly random non-negative BigNumber of at most `bits` bits
// (ipcl/utils/common.cpp:79-101)
BigNumber getRandomBN(int bits);

}  // namespace ipcl
#endif  // IPCL_B200_UTILS_COMMON_HPP_
