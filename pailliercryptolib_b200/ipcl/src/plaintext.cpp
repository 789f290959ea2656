// plaintext.cpp -- ipcl::PlainText (interface:
// /root/reference/ipcl/include/ipcl/plaintext.hpp:18-98).  A PlainText is a
// BaseText plus conversions and the commuted forms of the homomorphic
// operators, which simply forward to CipherText.
#include "ipcl/plaintext.hpp"

#include <utility>

#include "ipcl/ciphertext.hpp"
#include "ipcl/utils/util.hpp"
#include "text_util.hpp"

namespace ipcl {

PlainText::PlainText(const uint32_t& n) : BaseText(n) {}
PlainText::PlainText(const std::vector<uint32_t>& n_v) : BaseText(n_v) {}
PlainText::PlainText(const BigNumber& bn) : BaseText(bn) {}
PlainText::PlainText(const std::vector<BigNumber>& bn_v) : BaseText(bn_v) {}
PlainText::PlainText(std::vector<BigNumber>&& bn_v) : BaseText(std::move(bn_v)) {}
PlainText::PlainText(std::shared_ptr<detail::DeviceBatch> dev)
    : BaseText(std::move(dev)) {}
PlainText::PlainText(const PlainText& pt) : BaseText(pt) {}

PlainText& PlainText::operator=(const PlainText& other) {
  BaseText::operator=(other);
  return *this;
}

// pt + ct and pt * ct commute with the CipherText forms
CipherText PlainText::operator+(const CipherText& ct) const { return ct + *this; }
CipherText PlainText::operator*(const CipherText& ct) const { return ct * *this; }

// conversions of the FIRST element (and of the whole container)
PlainText::operator std::vector<uint32_t>() const {
  ERROR_CHECK(m_size > 0, "PlainText: type conversion to uint32_t vector error");
  return getElementVec(0);
}

PlainText::operator BigNumber() const {
  ERROR_CHECK(m_size > 0, "PlainText: type conversion to BigNumber error");
  return texts().front();
}

PlainText::operator std::vector<BigNumber>() const {
  ERROR_CHECK(m_size > 0, "PlainText: type conversion to BigNumber vector error");
  return texts();
}

PlainText PlainText::rotate(int shift) const {
  return PlainText(detail::rotated(texts(), shift));
}

}  // namespace ipcl
