// plaintext.cpp -- ipcl::PlainText (reference: ipcl/plaintext.cpp).
#include "ipcl/plaintext.hpp"

#include <algorithm>
#include <utility>

#include "ipcl/ciphertext.hpp"
#include "ipcl/utils/util.hpp"

namespace ipcl {

PlainText::PlainText(const uint32_t& n) : BaseText(n) {}
PlainText::PlainText(const std::vector<uint32_t>& n_v) : BaseText(n_v) {}
PlainText::PlainText(const BigNumber& bn) : BaseText(bn) {}
PlainText::PlainText(const std::vector<BigNumber>& bn_v) : BaseText(bn_v) {}
PlainText::PlainText(std::vector<BigNumber>&& bn_v) : BaseText(std::move(bn_v)) {}
PlainText::PlainText(const PlainText& pt) : BaseText(pt) {}

PlainText& PlainText::operator=(const PlainText& other) {
  BaseText::operator=(other);
  return *this;
}

CipherText PlainText::operator+(const CipherText& other) const {
  return other.operator+(*this);
}

CipherText PlainText::operator*(const CipherText& other) const {
  return other.operator*(*this);
}

PlainText::operator std::vector<uint32_t>() const {
  ERROR_CHECK(m_size > 0, "PlainText: type conversion to uint32_t vector error");
  std::vector<uint32_t> v;
  m_texts[0].num2vec(v);
  return v;
}

PlainText::operator BigNumber() const {
  ERROR_CHECK(m_size > 0, "PlainText: type conversion to BigNumber error");
  return m_texts[0];
}

PlainText::operator std::vector<BigNumber>() const {
  ERROR_CHECK(m_size > 0, "PlainText: type conversion to BigNumber vector error");
  return m_texts;
}

PlainText PlainText::rotate(int shift) const {
  const int size = static_cast<int>(m_size);
  ERROR_CHECK(m_size != 1, "rotate: Cannot rotate single CipherText");
  ERROR_CHECK(shift >= -size && shift <= size,
              "rotate: Cannot shift more than the test size");
  if (shift == 0 || shift == size || shift == -size) return PlainText(m_texts);
  // positive shift moves elements towards higher indices
  const int left = shift > 0 ? size - shift : -shift;
  std::vector<BigNumber> v(m_texts);
  std::rotate(v.begin(), v.begin() + left, v.end());
  return PlainText(v);
}

}  // namespace ipcl
