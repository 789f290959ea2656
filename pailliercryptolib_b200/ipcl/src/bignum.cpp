// bignum.cpp -- from-scratch BigNumber (see ipcl/bignum.h).  Behaviour follows
// /root/reference/ipcl/bignum.cpp: operator% returns the non-negative residue
// (ippsMod_BN, :304-308), num2hex prints "0x" + lower-case digits without
// leading zeros (:470-494), a zero value has one word / one bit as ippsRef_BN
// reports it, toBin/fromBin are big-endian octet strings (:511-565).
#include "ipcl/bignum.h"

#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "hostbn.hpp"

namespace hbn = ipclb200::hbn;

void BigNumber::normalize() {
  hbn::trim(m_mag);
  if (m_mag.empty()) m_neg = false;
}

BigNumber::BigNumber(Ipp32u value) {
  if (value) m_mag.push_back(value);
}

BigNumber::BigNumber(Ipp32s value) {
  if (value) {
    m_neg = value < 0;
    m_mag.push_back(m_neg ? (Ipp32u)(-(int64_t)value) : (Ipp32u)value);
  }
}

BigNumber::BigNumber(const Ipp32u* pData, int length, IppsBigNumSGN sgn) {
  Set(pData, length, sgn);
}

void BigNumber::Set(const Ipp32u* pData, int length, IppsBigNumSGN sgn) {
  m_mag.clear();
  if (pData && length > 0) m_mag.assign(pData, pData + length);
  m_neg = (sgn == IppsBigNumNEG);
  normalize();
}

BigNumber::BigNumber(const char* s) {
  if (!s) return;
  bool neg = (*s == '-');
  if (neg) s++;
  if (s[0] == '0' && (s[1] == 'x' || s[1] == 'X')) {
    s += 2;
    size_t len = strlen(s);
    m_mag.assign((len + 7) / 8, 0u);
    for (size_t i = 0; i < len; i++) {
      char c = s[len - 1 - i];
      Ipp32u d = (c >= '0' && c <= '9')   ? (Ipp32u)(c - '0')
                 : (c >= 'a' && c <= 'f') ? (Ipp32u)(c - 'a' + 10)
                 : (c >= 'A' && c <= 'F') ? (Ipp32u)(c - 'A' + 10)
                                          : 0u;
      m_mag[i / 8] |= d << (4 * (i % 8));
    }
  } else {
    for (; *s; s++) {
      if (*s < '0' || *s > '9') continue;
      // m_mag = m_mag * 10 + digit
      uint64_t carry = (uint64_t)(*s - '0');
      for (auto& w : m_mag) {
        carry += (uint64_t)w * 10u;
        w = (Ipp32u)carry;
        carry >>= 32;
      }
      if (carry) m_mag.push_back((Ipp32u)carry);
    }
  }
  m_neg = neg;
  normalize();
}

const BigNumber& BigNumber::Zero() {
  static const BigNumber z(0u);
  return z;
}
const BigNumber& BigNumber::One() {
  static const BigNumber o(1u);
  return o;
}
const BigNumber& BigNumber::Two() {
  static const BigNumber t(2u);
  return t;
}

// signed add of (an, a) + (bn, b)
static void signed_add(bool an, const hbn::Limbs& a, bool bn, const hbn::Limbs& b,
                       bool* rn, hbn::Limbs* r) {
  if (an == bn) {
    *r = hbn::add(a, b);
    *rn = an;
  } else {
    int c = hbn::cmp(a, b);
    if (c == 0) {
      r->clear();
      *rn = false;
    } else if (c > 0) {
      *r = hbn::sub(a, b);
      *rn = an;
    } else {
      *r = hbn::sub(b, a);
      *rn = bn;
    }
  }
}

BigNumber& BigNumber::operator+=(const BigNumber& bn) {
  hbn::Limbs r;
  bool rn;
  signed_add(m_neg, m_mag, bn.m_neg, bn.m_mag, &rn, &r);
  m_mag.swap(r);
  m_neg = rn;
  normalize();
  return *this;
}
BigNumber& BigNumber::operator+=(Ipp32u n) { return *this += BigNumber(n); }

BigNumber& BigNumber::operator-=(const BigNumber& bn) {
  hbn::Limbs r;
  bool rn;
  signed_add(m_neg, m_mag, !bn.m_neg, bn.m_mag, &rn, &r);
  m_mag.swap(r);
  m_neg = rn;
  normalize();
  return *this;
}
BigNumber& BigNumber::operator-=(Ipp32u n) { return *this -= BigNumber(n); }

BigNumber& BigNumber::operator*=(const BigNumber& bn) {
  m_mag = hbn::mul(m_mag, bn.m_mag);
  m_neg = (m_neg != bn.m_neg);
  normalize();
  return *this;
}
BigNumber& BigNumber::operator*=(Ipp32u n) { return *this *= BigNumber(n); }

// quotient truncated toward zero (the remainder takes the dividend's sign)
BigNumber& BigNumber::operator/=(const BigNumber& bn) {
  if (bn.m_mag.empty()) throw std::runtime_error("BigNumber: division by zero");
  hbn::Limbs q;
  hbn::divmod(m_mag, bn.m_mag, &q, nullptr);
  m_mag.swap(q);
  m_neg = (m_neg != bn.m_neg);
  normalize();
  return *this;
}
BigNumber& BigNumber::operator/=(Ipp32u n) { return *this /= BigNumber(n); }

// non-negative residue modulo |bn| (ippsMod_BN semantics)
BigNumber& BigNumber::operator%=(const BigNumber& bn) {
  if (bn.m_mag.empty()) throw std::runtime_error("BigNumber: modulo by zero");
  hbn::Limbs r = hbn::mod(m_mag, bn.m_mag);
  if (m_neg && !r.empty()) r = hbn::sub(bn.m_mag, r);
  m_mag.swap(r);
  m_neg = false;
  normalize();
  return *this;
}
BigNumber& BigNumber::operator%=(Ipp32u n) { return *this %= BigNumber(n); }

BigNumber operator+(const BigNumber& a, const BigNumber& b) {
  BigNumber r(a);
  return r += b;
}
BigNumber operator+(const BigNumber& a, Ipp32u n) {
  BigNumber r(a);
  return r += n;
}
BigNumber operator-(const BigNumber& a, const BigNumber& b) {
  BigNumber r(a);
  return r -= b;
}
BigNumber operator-(const BigNumber& a, Ipp32u n) {
  BigNumber r(a);
  return r -= n;
}
BigNumber operator*(const BigNumber& a, const BigNumber& b) {
  BigNumber r(a);
  return r *= b;
}
BigNumber operator*(const BigNumber& a, Ipp32u n) {
  BigNumber r(a);
  return r *= n;
}
BigNumber operator/(const BigNumber& a, const BigNumber& b) {
  BigNumber r(a);
  return r /= b;
}
BigNumber operator/(const BigNumber& a, Ipp32u n) {
  BigNumber r(a);
  return r /= n;
}
BigNumber operator%(const BigNumber& a, const BigNumber& b) {
  BigNumber r(a);
  return r %= b;
}
BigNumber operator%(const BigNumber& a, Ipp32u n) {
  BigNumber r(a);
  return r %= n;
}

BigNumber BigNumber::Modulo(const BigNumber& a) const { return a % *this; }

BigNumber BigNumber::InverseAdd(const BigNumber& a) const {
  BigNumber t = Modulo(a);
  if (t.m_mag.empty()) return t;
  return *this - t;
}

BigNumber BigNumber::InverseMul(const BigNumber& a) const {
  BigNumber ar = Modulo(a);
  BigNumber r;
  if (!hbn::modinv(ar.m_mag, m_mag, &r.m_mag))
    throw std::runtime_error("BigNumber::InverseMul: not invertible");
  r.normalize();
  return r;
}

BigNumber BigNumber::ModAdd(const BigNumber& a, const BigNumber& b) const {
  return Modulo(a + b);
}
BigNumber BigNumber::ModSub(const BigNumber& a, const BigNumber& b) const {
  return Modulo(a + InverseAdd(b));
}
BigNumber BigNumber::ModMul(const BigNumber& a, const BigNumber& b) const {
  return Modulo(a * b);
}

BigNumber BigNumber::gcd(const BigNumber& q) const {
  BigNumber r;
  r.m_mag = hbn::gcd(m_mag, q.m_mag);
  return r;
}

int BigNumber::compare(const BigNumber& bn) const {
  if (m_neg != bn.m_neg) return m_neg ? -1 : 1;
  int c = hbn::cmp(m_mag, bn.m_mag);
  return m_neg ? -c : c;
}

bool operator<(const BigNumber& a, const BigNumber& b) { return a.compare(b) < 0; }
bool operator>(const BigNumber& a, const BigNumber& b) { return a.compare(b) > 0; }
bool operator==(const BigNumber& a, const BigNumber& b) { return a.compare(b) == 0; }
bool operator!=(const BigNumber& a, const BigNumber& b) { return a.compare(b) != 0; }

bool BigNumber::IsOdd() const { return !m_mag.empty() && (m_mag[0] & 1u); }

bool BigNumber::TestBit(int index) const {
  if (index < 0) return false;
  size_t w = (size_t)index / 32;
  if (w >= m_mag.size()) return false;
  return (m_mag[w] >> (index % 32)) & 1u;
}

int BigNumber::MSB() const {
  if (m_mag.empty()) return 0;
  return hbn::bitlen(m_mag) - 1;
}

int BigNumber::LSB() const {
  for (size_t i = 0; i < m_mag.size(); i++)
    if (m_mag[i]) return (int)(i * 32) + __builtin_ctz(m_mag[i]);
  return 0;
}

int Bit(const std::vector<Ipp32u>& v, int n) {
  return 0 != (v[n >> 5] & (1u << (n & 0x1F)));
}

void BigNumber::num2vec(std::vector<Ipp32u>& v) const {
  if (m_mag.empty())
    v.push_back(0u);  // ippsRef_BN reports one word for zero
  else
    v.insert(v.end(), m_mag.begin(), m_mag.end());
}

void BigNumber::num2hex(std::string& s) const {
  static const char digits[] = "0123456789abcdef";
  if (m_neg) s.push_back('-');
  s += "0x";
  bool started = false;
  for (size_t i = m_mag.size(); i-- > 0;) {
    for (int nd = 7; nd >= 0; nd--) {
      char c = digits[(m_mag[i] >> (4 * nd)) & 0xF];
      if (c != '0' || started) {
        started = true;
        s.push_back(c);
      }
    }
  }
}

std::ostream& operator<<(std::ostream& os, const BigNumber& a) {
  std::string s;
  a.num2hex(s);
  return os << s;
}

void BigNumber::num2char(std::vector<Ipp8u>& dest) const {
  int bits = m_mag.empty() ? 1 : hbn::bitlen(m_mag);
  int len = (bits + 7) >> 3;
  dest.clear();
  for (int i = 0; i < len; i++) {
    size_t w = (size_t)i / 4;
    Ipp32u word = w < m_mag.size() ? m_mag[w] : 0u;
    dest.push_back((Ipp8u)(word >> (8 * (i % 4))));
  }
}

bool BigNumber::toWords(Ipp32u* out, std::size_t n) const {
  if (m_mag.size() > n) return false;
  hbn::to_words(m_mag, out, n);
  return true;
}

bool BigNumber::fromBin(BigNumber& bn, const unsigned char* data, int len) {
  if (len <= 0 || !data) return false;
  // the reference consumes len/4 words (bignum.cpp:515)
  int words = len / 4;
  std::vector<Ipp32u> v((size_t)words, 0u);
  for (int i = 0; i < words * 4; i++)
    v[(size_t)i / 4] |= (Ipp32u)data[len - 1 - i] << (8 * (i % 4));
  bn.Set(v.data(), words, IppsBigNumPOS);
  return true;
}

bool BigNumber::toBin(unsigned char* data, int len, const BigNumber& bn) {
  if (len <= 0 || !data) return false;
  int nbytes = (int)(bn.m_mag.empty() ? 1 : bn.m_mag.size()) * 4;
  for (int i = 0; i < nbytes && i < len; i++) {
    size_t w = (size_t)i / 4;
    Ipp32u word = w < bn.m_mag.size() ? bn.m_mag[w] : 0u;
    data[len - 1 - i] = (unsigned char)(word >> (8 * (i % 4)));
  }
  return true;
}

bool BigNumber::toBin(unsigned char** bin, int* len, const BigNumber& bn) {
  if (!bin || !len) return false;
  int nbytes = (int)(bn.m_mag.empty() ? 1 : bn.m_mag.size()) * 4;
  *len = nbytes;
  bin[0] = reinterpret_cast<unsigned char*>(calloc((size_t)nbytes, 1));
  if (!bin[0]) return false;
  return toBin(bin[0], nbytes, bn);
}
