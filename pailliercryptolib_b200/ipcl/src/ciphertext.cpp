// ciphertext.cpp -- ipcl::CipherText homomorphic operations on the B200
// back-end (reference: ipcl/ciphertext.cpp:35-162).
#include "ipcl/ciphertext.hpp"

#include <algorithm>
#include <utility>

#include "device_batch.hpp"
#include "ipcl/mod_exp.hpp"
#include "text_util.hpp"

namespace ipcl {

CipherText::CipherText(const PublicKey& pk, const uint32_t& n)
    : BaseText(n), m_pk(std::make_shared<PublicKey>(pk)) {}

CipherText::CipherText(const PublicKey& pk, const std::vector<uint32_t>& n_v)
    : BaseText(n_v), m_pk(std::make_shared<PublicKey>(pk)) {}

CipherText::CipherText(const PublicKey& pk, const BigNumber& bn)
    : BaseText(bn), m_pk(std::make_shared<PublicKey>(pk)) {}

CipherText::CipherText(const PublicKey& pk, const std::vector<BigNumber>& bn_v)
    : BaseText(bn_v), m_pk(std::make_shared<PublicKey>(pk)) {}

CipherText::CipherText(const PublicKey& pk, std::vector<BigNumber>&& bn_v)
    : BaseText(std::move(bn_v)), m_pk(std::make_shared<PublicKey>(pk)) {}

CipherText::CipherText(const PublicKey& pk,
                       std::shared_ptr<detail::DeviceBatch> dev)
    : BaseText(std::move(dev)), m_pk(std::make_shared<PublicKey>(pk)) {}

CipherText::CipherText(const CipherText& ct) : BaseText(ct), m_pk(ct.m_pk) {}

CipherText& CipherText::operator=(const CipherText& other) {
  BaseText::operator=(other);
  m_pk = other.m_pk;
  return *this;
}

// ct + ct = a*b mod n^2; a size-1 right operand is broadcast (:37,51-59)
CipherText CipherText::operator+(const CipherText& other) const {
  const std::size_t b_size = other.getSize();
  ERROR_CHECK(this->m_size == b_size || b_size == 1,
              "CT + CT error: Size mismatch!");
  ERROR_CHECK(*(m_pk->getN()) == *(other.m_pk->getN()),
              "CT + CT error: 2 different public keys detected!");
  if (m_size == 1)
    return CipherText(*m_pk, raw_add(texts().front(), other.texts().front()));
  // device-resident: one ipclb200_modmul_dev on the batches where they are
  const int W = 2 * static_cast<int>(m_pk->getN()->words().size());
  if (detail::deviceResidentEnabled() && detail::isClassWords(W)) {
    auto a = deviceBatch(W);
    // a size-1 right operand is sharded differently: it goes down as one
    // shared host value instead
    auto b = (a && b_size != 1) ? other.deviceBatch(W) : nullptr;
    if (a && b_size == 1) {
      std::vector<uint32_t> mod(static_cast<std::size_t>(W)), bw(static_cast<std::size_t>(W));
      m_pk->getNSQ()->toWords(mod.data(), mod.size());
      const BigNumber bv = other.texts().front() % *(m_pk->getNSQ());
      bv.toWords(bw.data(), bw.size());
      auto out = std::make_shared<detail::DeviceBatch>(m_size, W);
      DEVICE_CHECK(ipclb200_modmul_batch(a->h, nullptr, bw.data(), mod.data(), W, out->h));
      return CipherText(*m_pk, std::move(out));
    }
    if (a && b) {
      std::vector<uint32_t> mod(static_cast<std::size_t>(W));
      m_pk->getNSQ()->toWords(mod.data(), mod.size());
      auto out = std::make_shared<detail::DeviceBatch>(m_size, W);
      DEVICE_CHECK(ipclb200_modmul_batch(a->h, b->h, nullptr, mod.data(), W, out->h));
      return CipherText(*m_pk, std::move(out));
    }
  }
  return CipherText(*m_pk, modMul(texts(), other.texts(), *(m_pk->getNSQ())));
}

// ct + pt: encode pt without obfuscation, then ct + ct (:75-80)
CipherText CipherText::operator+(const PlainText& other) const {
  CipherText b = this->m_pk->encrypt(other, false);
  return this->operator+(b);
}

// ct * pt = a^b mod n^2; a size-1 plaintext is broadcast (:83-106)
CipherText CipherText::operator*(const PlainText& other) const {
  const std::size_t b_size = other.getSize();
  ERROR_CHECK(this->m_size == b_size || b_size == 1,
              "CT * PT error: Size mismatch!");
  if (m_size == 1)
    return CipherText(*m_pk, raw_mul(texts().front(), other.texts().front()));
  // device-resident: one ipclb200_modexp_dev, base = the ciphertexts where they
  // are, exponent = the plaintexts (uploaded once if they are host values)
  const int W = 2 * static_cast<int>(m_pk->getN()->words().size());
  if (detail::deviceResidentEnabled() && detail::isClassWords(W)) {
    int ew = 0, ebits = 0;
    if (other.isHostMaterialized()) {
      ew = detail::maxWords(other.texts());
      for (const auto& x : other.texts()) ebits = std::max(ebits, x.BitSize());
    } else {
      ew = W / 2;  // came out of decrypt: n-word values
      ebits = 32 * ew;
    }
    auto a = deviceBatch(W);
    if (a && b_size == 1 && !other.texts().front().isNegative()) {
      // one exponent for the whole batch: a shared host value (sliding-window
      // schedule on the device)
      std::vector<uint32_t> mod(static_cast<std::size_t>(W));
      m_pk->getNSQ()->toWords(mod.data(), mod.size());
      const auto& ev = other.texts().front().words();
      std::vector<uint32_t> ex(ev.empty() ? std::vector<uint32_t>(1, 0u) : ev);
      auto out = std::make_shared<detail::DeviceBatch>(m_size, W);
      DEVICE_CHECK(ipclb200_modexp_batch(a->h, nullptr, ex.data(),
                                         static_cast<int>(ex.size()), 0, mod.data(), W,
                                         out->h));
      return CipherText(*m_pk, std::move(out));
    }
    auto e = (a && b_size != 1) ? other.deviceBatch(ew) : nullptr;
    if (a && e) {
      std::vector<uint32_t> mod(static_cast<std::size_t>(W));
      m_pk->getNSQ()->toWords(mod.data(), mod.size());
      auto out = std::make_shared<detail::DeviceBatch>(m_size, W);
      DEVICE_CHECK(ipclb200_modexp_batch(a->h, e->h, nullptr, ew, ebits > 0 ? ebits : 1,
                                         mod.data(), W, out->h));
      return CipherText(*m_pk, std::move(out));
    }
  }
  if (b_size == 1) {
    std::vector<BigNumber> b_v(m_size, other.texts().front());
    return CipherText(*m_pk, raw_mul(texts(), b_v));
  }
  return CipherText(*m_pk, raw_mul(texts(), other.texts()));
}

CipherText CipherText::getCipherText(const size_t& idx) const {
  ERROR_CHECK(idx < m_size, "CipherText::getCipherText index is out of range");
  return CipherText(*m_pk, texts()[idx]);
}

std::shared_ptr<PublicKey> CipherText::getPubKey() const { return m_pk; }

CipherText CipherText::rotate(int shift) const {
  return CipherText(*m_pk, detail::rotated(texts(), shift));
}

BigNumber CipherText::raw_add(const BigNumber& a, const BigNumber& b) const {
  return modMul({a}, {b}, *(m_pk->getNSQ()))[0];
}

BigNumber CipherText::raw_mul(const BigNumber& a, const BigNumber& b) const {
  return modExp(a, b, *(m_pk->getNSQ()));
}

std::vector<BigNumber> CipherText::raw_mul(
    const std::vector<BigNumber>& a, const std::vector<BigNumber>& b) const {
  std::vector<BigNumber> sq(a.size(), *(m_pk->getNSQ()));
  return modExp(a, b, sq);
}

}  // namespace ipcl
