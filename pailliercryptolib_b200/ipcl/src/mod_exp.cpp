// mod_exp.cpp -- ipcl::modExp / ippModExp on the B200 back-end.
//
// Reference: ipcl/mod_exp.cpp.  There, ippModExp(vector...) (:655-678) cuts the
// batch into chunks of 8, copies each chunk into zero-padded u64[8][dwords]
// buffers and calls mbx_exp_mb8 per chunk under OpenMP (:597-636, :446-533),
// and a batch of one goes to the single-buffer ippsMontExp path (:535-585).
// Here the whole batch -- of any size, one included -- is packed once into a
// flat limb buffer and handed to ipclb200_modexp in ONE call.
#include "ipcl/mod_exp.hpp"

#include <type_traits>

#include "ipcl/utils/util.hpp"
#include "ipcl_b200.h"
#include "marshal.hpp"

namespace ipcl {

// hybrid knobs: state only (mod_exp.cpp:22-64); thread_local like the reference
static thread_local struct {
  float ratio;
  HybridMode mode;
} g_hybrid_params = {0.0f, HybridMode::OPTIMAL};

void setHybridRatio(float ratio, bool reset_mode) {
  ERROR_CHECK((ratio <= 1.0f) && (ratio >= 0.0f),
              "setHybridRatio: Hybrid modexp qat ratio is NOT correct");
  g_hybrid_params.ratio = ratio;
  if (reset_mode) g_hybrid_params.mode = HybridMode::UNDEFINED;
}

void setHybridMode(HybridMode mode) {
  int v = static_cast<std::underlying_type<HybridMode>::type>(mode);
  g_hybrid_params.ratio = v < 0 ? 0.0f : v / 100.0f;
  g_hybrid_params.mode = mode;
}

void setHybridOff() {
  g_hybrid_params.ratio = 0.0f;
  g_hybrid_params.mode = HybridMode::UNDEFINED;
}

float getHybridRatio() { return g_hybrid_params.ratio; }
HybridMode getHybridMode() { return g_hybrid_params.mode; }
bool isHybridOptimal() { return g_hybrid_params.mode == HybridMode::OPTIMAL; }

std::vector<BigNumber> ippModExp(const std::vector<BigNumber>& base,
                                 const std::vector<BigNumber>& exp,
                                 const std::vector<BigNumber>& mod) {
  const std::size_t n = base.size();
  ERROR_CHECK((n == exp.size()) && (n == mod.size()),
              "ippModExp: input vector size error");
  if (n == 0) return {};

  for (std::size_t i = 0; i < n; i++) {
    ERROR_CHECK(!exp[i].isNegative(), "ippModExp: negative exponent");
    ERROR_CHECK(!mod[i].isNegative() && mod[i] != BigNumber::Zero(),
                "ippModExp: modulus must be positive");
  }

  unsigned flags = 0;
  const bool shared_mod = detail::allEqual(mod);
  const bool shared_exp = n > 1 && detail::allEqual(exp);
  const int mod_words = detail::maxWords(mod);
  const int exp_words = detail::maxWords(exp);

  // bases must be non-negative and fit mod_words words; the kernel reduces
  // anything in [0, 2^(32 mod_words)) itself, only oversize/negative values
  // are brought into range here
  const std::vector<BigNumber>* bp = &base;
  std::vector<BigNumber> reduced;
  for (std::size_t i = 0; i < n; i++) {
    if (base[i].isNegative() ||
        static_cast<int>(base[i].words().size()) > mod_words) {
      if (reduced.empty()) reduced = base;
      reduced[i] = base[i] % mod[i];
      bp = &reduced;
    }
  }
  const bool shared_base = n > 1 && detail::allEqual(*bp);

  // flat operands and results in page-locked slabs: the copies to and from the
  // GPU(s) are DMA transfers
  const std::size_t MW = static_cast<std::size_t>(mod_words);
  const std::size_t EW = static_cast<std::size_t>(exp_words);
  detail::ScopedSlab fb((shared_base ? 1 : n) * MW), fe((shared_exp ? 1 : n) * EW),
      fm((shared_mod ? 1 : n) * MW), fo(n * MW);
  if (shared_base) {
    detail::pack({(*bp)[0]}, mod_words, fb.data());
    flags |= IPCLB200_SHARED_BASE;
  } else {
    detail::pack(*bp, mod_words, fb.data());
  }
  if (shared_exp) {
    detail::pack({exp[0]}, exp_words, fe.data());
    flags |= IPCLB200_SHARED_EXP;
  } else {
    detail::pack(exp, exp_words, fe.data());
  }
  if (shared_mod) {
    detail::pack({mod[0]}, mod_words, fm.data());
    flags |= IPCLB200_SHARED_MOD;
  } else {
    detail::pack(mod, mod_words, fm.data());
  }
  DEVICE_CHECK(ipclb200_modexp(fb.data(), fe.data(), fm.data(), mod_words,
                               exp_words, n, flags, fo.data()));
  return detail::unpack(fo.data(), n, mod_words);
}

BigNumber ippModExp(const BigNumber& base, const BigNumber& exp,
                    const BigNumber& mod) {
  return ippModExp(std::vector<BigNumber>{base}, std::vector<BigNumber>{exp},
                   std::vector<BigNumber>{mod})[0];
}

std::vector<BigNumber> modExp(const std::vector<BigNumber>& base,
                              const std::vector<BigNumber>& exp,
                              const std::vector<BigNumber>& mod) {
  // the CPU/QAT split of mod_exp.cpp:688-732 has nothing to split here
  return ippModExp(base, exp, mod);
}

BigNumber modExp(const BigNumber& base, const BigNumber& exp,
                 const BigNumber& mod) {
  return ippModExp(base, exp, mod);
}

std::vector<BigNumber> qatModExp(const std::vector<BigNumber>&,
                                 const std::vector<BigNumber>&,
                                 const std::vector<BigNumber>&) {
  ERROR_CHECK(false, "qatModExp: Need to turn on IPCL_ENABLE_QAT");
  return {};
}

std::vector<BigNumber> modMul(const std::vector<BigNumber>& a,
                              const std::vector<BigNumber>& b,
                              const BigNumber& mod) {
  const std::size_t n = a.size();
  ERROR_CHECK(b.size() == n || b.size() == 1, "modMul: input vector size error");
  ERROR_CHECK(!mod.isNegative() && mod != BigNumber::Zero(),
              "modMul: modulus must be positive");
  if (n == 0) return {};
  const int words = static_cast<int>(mod.words().size());
  auto in_range = [&](const BigNumber& x) {
    return !x.isNegative() && static_cast<int>(x.words().size()) <= words;
  };
  std::vector<BigNumber> ra, rb;
  const std::vector<BigNumber>*pa = &a, *pb = &b;
  for (std::size_t i = 0; i < n; i++)
    if (!in_range(a[i])) {
      if (ra.empty()) ra = a;
      ra[i] = a[i] % mod;
      pa = &ra;
    }
  for (std::size_t i = 0; i < b.size(); i++)
    if (!in_range(b[i])) {
      if (rb.empty()) rb = b;
      rb[i] = b[i] % mod;
      pb = &rb;
    }
  const std::size_t W = static_cast<std::size_t>(words);
  detail::ScopedSlab fa(n * W), fb(pb->size() * W), fm(W), fo(n * W);
  detail::pack(*pa, words, fa.data());
  detail::pack(*pb, words, fb.data());
  detail::pack({mod}, words, fm.data());
  unsigned flags = (b.size() == 1 && n > 1) ? IPCLB200_SHARED_B : 0u;
  DEVICE_CHECK(ipclb200_modmul(fa.data(), fb.data(), fm.data(), words, n, flags,
                               fo.data()));
  return detail::unpack(fo.data(), n, words);
}

}  // namespace ipcl
