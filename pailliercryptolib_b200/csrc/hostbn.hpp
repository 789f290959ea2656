// hostbn.hpp -- small host-side unsigned big-integer helpers (magnitude only).
//
// Used (a) by the C-ABI layer to derive the per-modulus Montgomery constants
// the kernels need (R mod n, R^2 mod n, -n^-1 mod 2^32) and (b) as the engine
// under the from-scratch ::BigNumber value type that replaces the IPP-Crypto
// backed one of the reference (ipcl/bignum.cpp).  Little-endian 32-bit limbs,
// normalised (no leading zero limbs; zero is the empty vector).
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace ipclb200 {
namespace hbn {

using Limbs = std::vector<uint32_t>;

inline void trim(Limbs& a) {
  while (!a.empty() && a.back() == 0) a.pop_back();
}
inline Limbs from_words(const uint32_t* p, size_t n) {
  Limbs r(p, p + n);
  trim(r);
  return r;
}
inline Limbs from_u64(uint64_t v) {
  Limbs r;
  while (v) {
    r.push_back((uint32_t)v);
    v >>= 32;
  }
  return r;
}
// copy into a fixed-width little-endian buffer (zero padded / truncated)
inline void to_words(const Limbs& a, uint32_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) out[i] = i < a.size() ? a[i] : 0u;
}
inline bool is_zero(const Limbs& a) { return a.empty(); }
inline int bitlen(const Limbs& a) {
  if (a.empty()) return 0;
  return (int)(a.size() * 32) - __builtin_clz(a.back());
}
inline int cmp(const Limbs& a, const Limbs& b) {
  if (a.size() != b.size()) return a.size() < b.size() ? -1 : 1;
  for (size_t i = a.size(); i-- > 0;) {
    if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  }
  return 0;
}
inline Limbs add(const Limbs& a, const Limbs& b) {
  const Limbs& x = a.size() >= b.size() ? a : b;
  const Limbs& y = a.size() >= b.size() ? b : a;
  Limbs r(x.size() + 1);
  uint64_t c = 0;
  for (size_t i = 0; i < x.size(); i++) {
    c += (uint64_t)x[i] + (i < y.size() ? y[i] : 0u);
    r[i] = (uint32_t)c;
    c >>= 32;
  }
  r[x.size()] = (uint32_t)c;
  trim(r);
  return r;
}
// a - b, requires a >= b
inline Limbs sub(const Limbs& a, const Limbs& b) {
  if (cmp(a, b) < 0) throw std::runtime_error("hbn::sub: negative result");
  Limbs r(a.size());
  int64_t br = 0;
  for (size_t i = 0; i < a.size(); i++) {
    int64_t t = (int64_t)a[i] - (i < b.size() ? b[i] : 0u) - br;
    br = t < 0;
    r[i] = (uint32_t)t;
  }
  trim(r);
  return r;
}
inline Limbs mul(const Limbs& a, const Limbs& b) {
  if (a.empty() || b.empty()) return {};
  Limbs r(a.size() + b.size(), 0u);
  for (size_t i = 0; i < a.size(); i++) {
    uint64_t c = 0, ai = a[i];
    for (size_t j = 0; j < b.size(); j++) {
      c += ai * b[j] + r[i + j];
      r[i + j] = (uint32_t)c;
      c >>= 32;
    }
    r[i + b.size()] = (uint32_t)c;
  }
  trim(r);
  return r;
}
inline Limbs shl(const Limbs& a, unsigned bits) {
  if (a.empty()) return {};
  size_t ws = bits / 32;
  unsigned bs = bits % 32;
  Limbs r(a.size() + ws + 1, 0u);
  for (size_t i = 0; i < a.size(); i++) {
    r[i + ws] |= a[i] << bs;
    if (bs) r[i + ws + 1] |= a[i] >> (32 - bs);
  }
  trim(r);
  return r;
}
inline Limbs shr(const Limbs& a, unsigned bits) {
  size_t ws = bits / 32;
  unsigned bs = bits % 32;
  if (ws >= a.size()) return {};
  Limbs r(a.size() - ws, 0u);
  for (size_t i = 0; i < r.size(); i++) {
    r[i] = a[i + ws] >> bs;
    if (bs && i + ws + 1 < a.size()) r[i] |= a[i + ws + 1] << (32 - bs);
  }
  trim(r);
  return r;
}
// schoolbook long division with a normalised divisor; q = a / d, r = a % d
inline void divmod(const Limbs& a, const Limbs& d, Limbs* q, Limbs* r) {
  if (d.empty()) throw std::runtime_error("hbn::divmod: division by zero");
  if (cmp(a, d) < 0) {
    if (q) q->clear();
    if (r) *r = a;
    return;
  }
  if (d.size() == 1) {
    Limbs qq(a.size());
    uint64_t rem = 0;
    for (size_t i = a.size(); i-- > 0;) {
      uint64_t cur = (rem << 32) | a[i];
      qq[i] = (uint32_t)(cur / d[0]);
      rem = cur % d[0];
    }
    trim(qq);
    if (q) *q = qq;
    if (r) *r = from_u64(rem);
    return;
  }
  unsigned s = (unsigned)__builtin_clz(d.back());
  Limbs v = shl(d, s);
  Limbs u = shl(a, s);
  const size_t n = v.size();
  if (u.size() == a.size()) u.push_back(0);
  const size_t m = u.size() - n - 1 + 1;  // number of quotient digits
  Limbs qq(m, 0u);
  const uint64_t B = 1ull << 32;
  for (size_t jj = m; jj-- > 0;) {
    uint64_t num = ((uint64_t)u[jj + n] << 32) | u[jj + n - 1];
    uint64_t qh = num / v[n - 1];
    uint64_t rh = num % v[n - 1];
    while (qh >= B || qh * v[n - 2] > ((rh << 32) | u[jj + n - 2])) {
      qh--;
      rh += v[n - 1];
      if (rh >= B) break;
    }
    // u[jj..jj+n] -= qh * v
    uint64_t carry = 0;
    int64_t borrow = 0;
    for (size_t i = 0; i < n; i++) {
      uint64_t p = qh * v[i] + carry;
      carry = p >> 32;
      int64_t t = (int64_t)u[jj + i] - (int64_t)(uint32_t)p - borrow;
      borrow = t < 0;
      u[jj + i] = (uint32_t)t;
    }
    int64_t t = (int64_t)u[jj + n] - (int64_t)carry - borrow;
    u[jj + n] = (uint32_t)t;
    if (t < 0) {
      qh--;
      uint64_t c = 0;
      for (size_t i = 0; i < n; i++) {
        c += (uint64_t)u[jj + i] + v[i];
        u[jj + i] = (uint32_t)c;
        c >>= 32;
      }
      u[jj + n] += (uint32_t)c;
    }
    qq[jj] = (uint32_t)qh;
  }
  trim(qq);
  if (q) *q = qq;
  if (r) {
    u.resize(n);
    trim(u);
    *r = shr(u, s);
  }
}
inline Limbs mod(const Limbs& a, const Limbs& d) {
  Limbs r;
  divmod(a, d, nullptr, &r);
  return r;
}
inline Limbs pow2(unsigned bits) {
  Limbs r(bits / 32 + 1, 0u);
  r[bits / 32] = 1u << (bits % 32);
  return r;
}
// -n^{-1} mod 2^32 for odd n0 (Newton iteration on the 2-adic inverse)
// floor(sqrt(a)) by Newton's iteration; *exact = (root^2 == a)
inline Limbs isqrt(const Limbs& a, bool* exact) {
  if (a.empty()) {
    if (exact) *exact = true;
    return a;
  }
  Limbs x = pow2((unsigned)((bitlen(a) + 1) / 2));  // >= sqrt(a)
  for (;;) {
    Limbs q;
    divmod(a, x, &q, nullptr);
    Limbs y = shr(add(x, q), 1);
    if (cmp(y, x) >= 0) break;
    x = y;
  }
  if (exact) *exact = cmp(mul(x, x), a) == 0;
  return x;
}

inline uint32_t neg_inv32(uint32_t n0) {
  uint32_t x = n0;
  for (int i = 0; i < 5; i++) x *= 2u - n0 * x;
  return 0u - x;
}
inline Limbs gcd(Limbs a, Limbs b) {
  while (!b.empty()) {
    Limbs r = mod(a, b);
    a.swap(b);
    b.swap(r);
  }
  return a;
}
// a^{-1} mod m (extended Euclid on magnitudes with sign tracking).
// returns false if gcd(a, m) != 1
inline bool modinv(const Limbs& a, const Limbs& m, Limbs* out) {
  Limbs r0 = m, r1 = mod(a, m);
  Limbs t0, t1 = from_u64(1);  // coefficients of a
  bool n0 = false, n1 = false; // signs of t0, t1
  while (!r1.empty()) {
    Limbs q, r2;
    divmod(r0, r1, &q, &r2);
    // t2 = t0 - q*t1
    Limbs qt = mul(q, t1);
    Limbs t2;
    bool n2;
    if (n0 == n1) {
      if (cmp(t0, qt) >= 0) {
        t2 = sub(t0, qt);
        n2 = n0;
      } else {
        t2 = sub(qt, t0);
        n2 = !n0;
      }
    } else {
      t2 = add(t0, qt);
      n2 = n0;
    }
    r0.swap(r1);
    r1.swap(r2);
    t0.swap(t1);
    n0 = n1;
    t1.swap(t2);
    n1 = n2;
  }
  if (!(r0.size() == 1 && r0[0] == 1)) return false;
  Limbs res = mod(t0, m);
  if (n0 && !res.empty()) res = sub(m, res);
  *out = res;
  return true;
}

}  // namespace hbn
}  // namespace ipclb200
