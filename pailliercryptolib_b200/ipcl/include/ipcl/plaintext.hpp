// plaintext.hpp -- ipcl::PlainText (ipcl/include/ipcl/plaintext.hpp:18-98).
#ifndef IPCL_B200_PLAINTEXT_HPP_
#define IPCL_B200_PLAINTEXT_HPP_

#include <vector>

#include "ipcl/base_text.hpp"

namespace ipcl {

class CipherText;

class PlainText : public BaseText {
 public:
  PlainText() = default;
  ~PlainText() = default;

  explicit PlainText(const uint32_t& n);
  explicit PlainText(const std::vector<uint32_t>& n_v);
  explicit PlainText(const BigNumber& bn);
  explicit PlainText(const std::vector<BigNumber>& bn_v);
  explicit PlainText(std::vector<BigNumber>&& bn_v);
  // back-end internal: a batch that lives in HBM (PrivateKey::decrypt)
  explicit PlainText(std::shared_ptr<detail::DeviceBatch> dev);
  PlainText(const PlainText& pt);
  PlainText& operator=(const PlainText& other);

  operator std::vector<uint32_t>() const;
  operator BigNumber() const;
  operator std::vector<BigNumber>() const;

  CipherText operator+(const CipherText& other) const;
  CipherText operator*(const CipherText& other) const;

  PlainText rotate(int shift) const;
};

}  // namespace ipcl
#endif  // IPCL_B200_PLAINTEXT_HPP_
