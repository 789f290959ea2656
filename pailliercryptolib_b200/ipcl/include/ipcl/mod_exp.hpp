// mod_exp.hpp -- the hot-path entry points, same signatures as the reference
// (ipcl/include/ipcl/mod_exp.hpp:16-115).  Every overload marshals into flat
// limb buffers and submits ONE batch to the sm_100a kernels through the C ABI
// (include/ipcl_b200.h); there is no CPU arithmetic path behind them.
#ifndef IPCL_B200_MOD_EXP_HPP_
#define IPCL_B200_MOD_EXP_HPP_

#include <vector>

#include "ipcl/bignum.h"

namespace ipcl {

// Hybrid CPU/QAT split of the reference.  Kept so callers compile and the
// getters round-trip; the ratio has no effect because the whole batch goes to
// the GPU.
enum class HybridMode {
  OPTIMAL = 95,
  QAT = 100,
  PREF_QAT90 = 90,
  PREF_QAT80 = 80,
  PREF_QAT70 = 70,
  PREF_QAT60 = 60,
  HALF = 50,
  PREF_IPP60 = 40,
  PREF_IPP70 = 30,
  PREF_IPP80 = 20,
  PREF_IPP90 = 10,
  IPP = 0,
  UNDEFINED = -1
};

void setHybridMode(HybridMode mode);
void setHybridRatio(float qat_ratio, bool reset_mode = true);
void setHybridOff();
float getHybridRatio();
HybridMode getHybridMode();
bool isHybridOptimal();

// res[i] = base[i]^exp[i] mod mod[i]
std::vector<BigNumber> modExp(const std::vector<BigNumber>& base,
                              const std::vector<BigNumber>& exp,
                              const std::vector<BigNumber>& mod);
BigNumber modExp(const BigNumber& base, const BigNumber& exp,
                 const BigNumber& mod);

// names kept from the reference; both run on the GPU
std::vector<BigNumber> ippModExp(const std::vector<BigNumber>& base,
                                 const std::vector<BigNumber>& exp,
                                 const std::vector<BigNumber>& mod);
BigNumber ippModExp(const BigNumber& base, const BigNumber& exp,
                    const BigNumber& mod);

// always throws: there is no QAT (mod_exp.cpp:587-595 without IPCL_USE_QAT)
std::vector<BigNumber> qatModExp(const std::vector<BigNumber>& base,
                                 const std::vector<BigNumber>& exp,
                                 const std::vector<BigNumber>& mod);

// a[i]*b[i] mod m for one modulus (b may hold one element: broadcast).
// Not in the reference's header; it is the batch form of BigNumber::ModMul
// that raw_add / applyObfuscator loop over on the host.
std::vector<BigNumber> modMul(const std::vector<BigNumber>& a,
                              const std::vector<BigNumber>& b,
                              const BigNumber& mod);

}  // namespace ipcl
#endif  // IPCL_B200_MOD_EXP_HPP_
