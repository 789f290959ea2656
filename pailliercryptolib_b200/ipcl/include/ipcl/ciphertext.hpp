// ciphertext.hpp -- ipcl::CipherText (ipcl/include/ipcl/ciphertext.hpp:15-78).
// ct + ct is one launch of the batched modular-multiply kernel, ct * pt one
// launch of the batched modexp kernel (modulus n^2 shared).
#ifndef IPCL_B200_CIPHERTEXT_HPP_
#define IPCL_B200_CIPHERTEXT_HPP_

#include <memory>
#include <vector>

#include "ipcl/plaintext.hpp"
#include "ipcl/pub_key.hpp"
#include "ipcl/utils/util.hpp"

namespace ipcl {

class CipherText : public BaseText {
 public:
  CipherText() = default;
  ~CipherText() = default;

  CipherText(const PublicKey& pk, const uint32_t& n);
  CipherText(const PublicKey& pk, const std::vector<uint32_t>& n_v);
  CipherText(const PublicKey& pk, const BigNumber& bn);
  CipherText(const PublicKey& pk, const std::vector<BigNumber>& bn_vec);
  CipherText(const PublicKey& pk, std::vector<BigNumber>&& bn_vec);
  CipherText(const CipherText& ct);
  CipherText& operator=(const CipherText& other);

  // homomorphic operations
  CipherText operator+(const CipherText& other) const;
  CipherText operator+(const PlainText& other) const;
  CipherText operator*(const PlainText& other) const;

  CipherText getCipherText(const size_t& idx) const;
  std::shared_ptr<PublicKey> getPubKey() const;
  CipherText rotate(int shift) const;

 private:
  BigNumber raw_add(const BigNumber& a, const BigNumber& b) const;
  BigNumber raw_mul(const BigNumber& a, const BigNumber& b) const;
  std::vector<BigNumber> raw_mul(const std::vector<BigNumber>& a,
                                 const std::vector<BigNumber>& b) const;

  std::shared_ptr<PublicKey> m_pk;
};

}  // namespace ipcl
#endif  // IPCL_B200_CIPHERTEXT_HPP_
