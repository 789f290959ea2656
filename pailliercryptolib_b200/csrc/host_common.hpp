// host_common.hpp -- host-side helpers of the C ABI implementation
// (ipcl_b200.cu): error convention, kernel size classes and lane layouts,
// exponent schedules, per-modulus Montgomery constants.  No device code.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ipcl_b200.h"
#include "hostbn.hpp"

namespace ipclb200 {
namespace host {

using hbn::Limbs;

inline std::string& last_error() {
  thread_local std::string e;
  return e;
}
inline int fail(int code, const std::string& msg) {
  last_error() = msg;
  return code;
}

#define CUDA_TRY(expr)                                                         \
  do {                                                                         \
    cudaError_t e_ = (expr);                                                   \
    if (e_ != cudaSuccess)                                                     \
      return ::ipclb200::host::fail(                                           \
          IPCLB200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

#define TRY(expr)             \
  do {                        \
    int rc_ = (expr);         \
    if (rc_ != 0) return rc_; \
  } while (0)

// ---------------------------------------------------------------------------
// size classes: modulus words -> (limbs per lane K, lanes per integer T)
// ---------------------------------------------------------------------------
constexpr int kClasses[] = {16, 32, 48, 64, 96, 128, 192, 256};
// largest number of jobs for which the wide (4 limbs per lane) layout is used
// (measured crossovers, profiles/r01_wide_layout.md)
constexpr size_t kWideMax32 = 6144, kWideMax64 = 3072, kWideMax128 = 1536;
// ... and for the middle layout (8 limbs per lane)
constexpr size_t kMidMax32 = 12288, kMidMax64 = 6144, kMidMax128 = 3072;

inline int class_words(int words) {
  for (int c : kClasses)
    if (words <= c) return c;
  return 0;
}

#define IPCLB200_DISPATCH(L, F)            \
  switch (L) {                             \
    case 16:  F(8, 2); break;              \
    case 32:  F(16, 2); break;             \
    case 48:  F(12, 4); break;             \
    case 64:  F(16, 4); break;             \
    case 96:  F(12, 8); break;             \
    case 128: F(16, 8); break;             \
    case 192: F(12, 16); break;            \
    case 256: F(16, 16); break;            \
    default: return ::ipclb200::host::fail(IPCLB200_ERR_UNSUPPORTED, "unsupported width"); \
  }

// Small batches: the same kernels with one integer spread over four times as
// many lanes (4 limbs per lane).  A batch that cannot fill the 148 SMs is
// latency bound -- a modexp is ~1200-2500 dependent Montgomery products -- and
// the wide layout shortens every product (8 multiplies per row and lane
// instead of 32) at the price of more shuffles per multiply.
#define IPCLB200_DISPATCH_WIDE(L, F)       \
  switch (L) {                             \
    case 32:  F(4, 8); break;              \
    case 64:  F(4, 16); break;             \
    case 128: F(4, 32); break;             \
    default: return ::ipclb200::host::fail(IPCLB200_ERR_UNSUPPORTED, "unsupported width"); \
  }

// in between: 8 limbs per lane, twice the default number of lanes
#define IPCLB200_DISPATCH_MID(L, F)        \
  switch (L) {                             \
    case 32:  F(8, 4); break;              \
    case 64:  F(8, 8); break;              \
    case 128: F(8, 16); break;             \
    default: return ::ipclb200::host::fail(IPCLB200_ERR_UNSUPPORTED, "unsupported width"); \
  }

// 0 = default layout, 1 = wide (4 limbs per lane), 2 = middle (8 limbs per
// lane); tasks = independent big-integer jobs of L words in the launch
inline int pick_layout(size_t tasks, int L) {
  if (!(L == 32 || L == 64 || L == 128)) return 0;
  const char* e = getenv("IPCLB200_WIDE");
  if (e && e[0] == '0') return 0;
  if (e && e[0] == '1') return 1;
  if (e && e[0] == '2') return 2;
  size_t wide_max = 0, mid_max = 0;
  switch (L) {
    case 32: wide_max = kWideMax32; mid_max = kMidMax32; break;
    case 64: wide_max = kWideMax64; mid_max = kMidMax64; break;
    default: wide_max = kWideMax128; mid_max = kMidMax128; break;
  }
  if (const char* m = getenv("IPCLB200_WIDE_MAX")) wide_max = strtoul(m, nullptr, 10);
  if (const char* m = getenv("IPCLB200_MID_MAX")) mid_max = strtoul(m, nullptr, 10);
  if (tasks <= wide_max) return 1;
  if (tasks <= mid_max) return 2;
  return 0;
}

constexpr int kMaxWindowBits = 6;
// fixed-window width minimising (2^w - 2) + bits + bits/w multiplies
inline int pick_window(int ebits) {
  int best = 1;
  long best_cost = -1;
  for (int w = 1; w <= kMaxWindowBits; w++) {
    long cost = ((1L << w) - 2) + ebits + (ebits + w - 1) / w;
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = w;
    }
  }
  return best;
}

constexpr int kSchedWindow = 5;  // 16 odd powers per table
// decrypt_hensel_kernel takes its table size from the schedule: window chosen per
// key in this range (ipclb200_privkey_create)
constexpr int kHenselMinWindow = 4, kHenselMaxWindow = 6;

// left-to-right sliding-window schedule for a fixed exponent (format: see
// modexp_sched_core in kernels.cuh).  e > 0.
inline std::vector<uint8_t> build_schedule(const Limbs& e, int w) {
  std::vector<uint8_t> s;
  s.push_back((uint8_t)(1 << (w - 1)));
  auto bit = [&](int i) { return i >= 0 && ((e[(size_t)i / 32] >> (i % 32)) & 1u); };
  int i = hbn::bitlen(e) - 1;
  bool first = true;
  while (i >= 0) {
    if (!bit(i)) {
      s.push_back(0);
      i--;
      continue;
    }
    int l = i - w + 1;
    if (l < 0) l = 0;
    while (!bit(l)) l++;
    unsigned v = 0;
    for (int k = i; k >= l; k--) v = (v << 1) | (bit(k) ? 1u : 0u);
    if (first) {
      s.push_back((uint8_t)((v - 1) / 2));
      first = false;
    } else {
      for (int k = i; k >= l; k--) s.push_back(0);
      s.push_back((uint8_t)((v - 1) / 2 + 1));
    }
    i = l - 1;
  }
  s.push_back(0xff);
  return s;
}

// decrypt_hensel_kernel's schedule: the byte schedule of build_schedule as 32-bit
// words [nodd, first, (run << 8 | entry)..., (run << 8 | 0xff)]: `run` squarings,
// then a multiply by odd power `entry` (0xff: none, end)
inline std::vector<uint32_t> hensel_schedule(const std::vector<uint8_t>& sched) {
  std::vector<uint32_t> out = {sched[0], sched[1]};
  uint32_t run = 0;
  for (size_t i = 2; sched[i] != 0xff; i++) {
    if (sched[i] == 0) {
      run++;
    } else {
      out.push_back((run << 8) | (uint32_t)(sched[i] - 1));
      run = 0;
    }
  }
  out.push_back((run << 8) | 0xffu);
  return out;
}

// The constant schedule of decrypt_hensel_kernel for a SECRET exponent: fixed
// 4-bit windows over a table of all 16 powers -- every window is four squarings
// and one multiply whatever the exponent's bits are, so the sequence of
// operations (and the run time) depends on the bit length only.  The table
// index of each multiply is still an exponent-dependent address.
inline std::vector<uint32_t> hensel_schedule_fixed(const Limbs& e) {
  const int w = 4;
  const int bits = hbn::bitlen(e);
  const int nwin = (bits + w - 1) / w;
  auto window = [&](int k) {
    uint32_t v = 0;
    for (int b = w - 1; b >= 0; b--) {
      const int i = k * w + b;
      const uint32_t bit = (i < bits) ? ((e[(size_t)i / 32] >> (i % 32)) & 1u) : 0u;
      v = (v << 1) | bit;
    }
    return v;
  };
  std::vector<uint32_t> out = {16u | (1u << 8), window(nwin - 1)};
  for (int k = nwin - 2; k >= 0; k--) out.push_back(((uint32_t)w << 8) | window(k));
  out.push_back(0xffu);
  return out;
}

// per-side constants of the two-digit decrypt (10*pl words):
//   p | pairs (k0_j, kw_j) of R^(j+1) mod p^2 in Montgomery form, j = 0..3 | -hp mod p
// A pair (x0, w) stands for x0 - w*p mod p^2 (mont_hensel.cuh).
inline void hensel_side_block(const Limbs& p, const Limbs& psq, const Limbs& hp, int pl,
                              uint32_t* out) {
  hbn::to_words(p, out, pl);
  const Limbs R = hbn::pow2(32u * (unsigned)pl);
  Limbs t = hbn::mod(R, psq);  // R^1
  for (int j = 0; j < 4; j++) {
    t = hbn::mod(hbn::mul(t, R), psq);  // R^(j+2) = R^(j+1) in Montgomery form
    Limbs hi, lo;
    hbn::divmod(t, p, &hi, &lo);
    Limbs wneg = hbn::mod(hi, p);
    Limbs w = hbn::is_zero(wneg) ? wneg : hbn::sub(p, wneg);
    hbn::to_words(lo, out + (size_t)(1 + 2 * j) * pl, pl);
    hbn::to_words(w, out + (size_t)(2 + 2 * j) * pl, pl);
  }
  Limbs nhp = hbn::is_zero(hp) ? hp : hbn::sub(p, hp);
  hbn::to_words(nhp, out + (size_t)9 * pl, pl);
}

// Lane layout of decrypt_hensel_kernel for a batch.  Layouts: 0/1/2 = one
// (ciphertext, side) task over T0, 2*T0, 4*T0 lanes (fewer limbs per lane: lower
// latency, more instructions per product), -2 = one task per THREAD (T = 1, only
// for 32-word primes: fewest instructions, 12 warps per SM, longest latency).
//
// A launch is a queue of warp-sized chunks over the 12*sms resident warps.
// Measured (profiles/r02_layout_sweep.jsonl, 2048-bit key, ms): the time is a
// staircase in the number of chunks -- `full` complete rounds of the resident
// warps plus a last round that puts w = 1, 2 or 3 warps on every SM
// sub-partition:
//     t = first[w]                                   (no complete round)
//     t = first[3] + (full - 1) * round + inc[w]     (otherwise)
// Only the ratios between layouts matter, so the 2048-bit numbers serve every
// key size.  Layout 2 keeps the simpler max(L, F * fill) model it was fitted
// with (profiles/r02_layout_small_batches.jsonl).
struct HenselStair {
  double first[3], round, inc[3];
};
inline double hensel_stair_time(const HenselStair& m, size_t count, int lanes, int sms) {
  const size_t per = (size_t)(32 / lanes);
  const size_t chunks = 2 * ((count + per - 1) / per);
  const size_t resident = (size_t)12 * (size_t)sms;
  const size_t full = chunks / resident, rem = chunks % resident;
  const size_t w = (rem + (size_t)4 * sms - 1) / ((size_t)4 * sms);  // 0..3
  if (full == 0) return m.first[w ? w - 1 : 0];
  return m.first[2] + (double)(full - 1) * m.round + (w ? m.inc[w - 1] : 0.0);
}
inline int pick_hensel_spread(size_t count, int pl, int sms) {
  static const HenselStair kThread = {{17.8, 23.8, 38.65}, 32.6, {8.9, 20.8, 32.6}};
  static const HenselStair kS0 = {{10.0, 13.67, 21.0}, 20.1, {5.3, 11.5, 19.0}};
  static const HenselStair kS1 = {{5.72, 8.17, 11.5}, 11.0, {3.9, 6.1, 11.3}};
  static const double kL2 = 3.95, kF2 = 6.5;
  const int t0 = pl == 64 ? 4 : 2;          // lanes per task at spread 0
  const int nspread = pl == 32 ? 3 : 2;     // instantiated layouts
  const HenselStair* stairs[2] = {pl == 16 ? &kS1 : &kS0, &kS1};
  int best = 0;
  double best_t = -1;
  auto consider = [&](int layout, double t) {
    if (best_t < 0 || t < best_t) {
      best_t = t;
      best = layout;
    }
  };
  for (int sp = 0; sp < nspread; sp++) {
    const int T = t0 << sp;
    if (sp < 2 && !(pl == 16 && sp == 1)) {
      consider(sp, hensel_stair_time(*stairs[sp], count, T, sms));
    } else {
      const double resident = 12.0 * sms;
      const double chunks = 2.0 * (double)((count + (32 / T) - 1) / (32 / T));
      const double full = (double)(size_t)(chunks / resident);
      const double fill = chunks / resident - full;
      double t = full * kF2;
      if (fill > 0) t += (kL2 > kF2 * fill ? kL2 : kF2 * fill);
      consider(sp, t);
    }
  }
  if (pl == 32) consider(-2, hensel_stair_time(kThread, count, 1, sms));
  // 48-word primes (3072-bit keys): layout 0 is 24 limbs x 2 lanes at 252
  // registers, i.e. 8 warps per SM, and runs latency-bound once a launch is more
  // than one round of those warps; the 12 x 4 layout keeps 12 warps per SM
  // (profiles/r02_layout_other_keys.jsonl: 8192 ciphertexts 42.9 vs 45.7 ms,
  // 32768: 173 vs 152 ms, 65536: 311 vs 301 ms)
  if (pl == 48 && best == 0 && 2 * ((count + 15) / 16) > (size_t)8 * (size_t)sms) best = 1;
  return best;
}

// ---------------------------------------------------------------------------
// per-modulus Montgomery constants (host side)
// ---------------------------------------------------------------------------
struct HostModConst {
  std::vector<uint32_t> n, rr, r3, one;
  uint32_t n0inv;
  uint32_t small_mod;
};

inline void host_mod_const(const Limbs& n, int L, HostModConst* h) {
  Limbs R = hbn::pow2(32u * (unsigned)L);
  Limbs one = hbn::mod(R, n);
  Limbs rr = hbn::mod(hbn::mul(one, one), n);
  Limbs r3 = hbn::mod(hbn::mul(rr, one), n);
  h->n.resize(L);
  h->rr.resize(L);
  h->r3.resize(L);
  h->one.resize(L);
  hbn::to_words(n, h->n.data(), L);
  hbn::to_words(rr, h->rr.data(), L);
  hbn::to_words(r3, h->r3.data(), L);
  hbn::to_words(one, h->one.data(), L);
  h->n0inv = hbn::neg_inv32(n[0]);
  h->small_mod = hbn::bitlen(n) <= 32 * L - 2 ? 1u : 0u;
}

inline int check_modulus(const uint32_t* mod, int words, Limbs* out) {
  Limbs n = hbn::from_words(mod, words);
  if (n.empty()) return fail(IPCLB200_ERR_BAD_ARG, "modulus is zero");
  if (!(n[0] & 1u))
    return fail(IPCLB200_ERR_EVEN_MODULUS,
                "modulus is even (Montgomery arithmetic needs an odd modulus)");
  *out = n;
  return 0;
}

inline int max_bits(const uint32_t* v, int words, size_t count, size_t stride) {
  int best = 0;
  for (size_t i = 0; i < count; i++) {
    const uint32_t* e = v + i * stride;
    for (int w = words - 1; w >= 0; w--) {
      if (e[w]) {
        int b = w * 32 + 32 - __builtin_clz(e[w]);
        if (b > best) best = b;
        break;
      }
      if ((w + 1) * 32 <= best) break;
    }
  }
  return best;
}

}  // namespace host
}  // namespace ipclb200
