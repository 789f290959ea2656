// ciphertext.cpp -- ipcl::CipherText homomorphic operations on the B200
// back-end (reference: ipcl/ciphertext.cpp:35-162).
#include "ipcl/ciphertext.hpp"

#include <algorithm>
#include <utility>

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
    return CipherText(*m_pk, raw_add(m_texts.front(), other.m_texts.front()));
  return CipherText(*m_pk, modMul(m_texts, other.m_texts, *(m_pk->getNSQ())));
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
    return CipherText(*m_pk, raw_mul(m_texts.front(), other.texts().front()));
  if (b_size == 1) {
    std::vector<BigNumber> b_v(m_size, other.texts().front());
    return CipherText(*m_pk, raw_mul(m_texts, b_v));
  }
  return CipherText(*m_pk, raw_mul(m_texts, other.texts()));
}

CipherText CipherText::getCipherText(const size_t& idx) const {
  ERROR_CHECK(idx < m_size, "CipherText::getCipherText index is out of range");
  return CipherText(*m_pk, m_texts[idx]);
}

std::shared_ptr<PublicKey> CipherText::getPubKey() const { return m_pk; }

CipherText CipherText::rotate(int shift) const {
  return CipherText(*m_pk, detail::rotated(m_texts, shift));
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
