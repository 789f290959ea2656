// ciphertext.hpp -- ipcl::CipherText: a batch of Paillier ciphertexts bound to
// the public key they were produced with, plus the homomorphic operators.
// Same public interface as /root/reference/ipcl/include/ipcl/ciphertext.hpp:15-78.
//
// On this back-end every operator is ONE device submission for the whole batch:
//   ct + ct   batched modular multiply            (ipclb200_modmul)
//   ct + pt   encode pt without obfuscator, then ct + ct
//   ct * pt   batched modexp with the shared modulus n^2   (ipclb200_modexp)
// and operands and results stay in HBM (device-resident texts, base_text.hpp):
// encrypt -> + -> * -> decrypt is four kernel launches and no marshalling.
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

  // Wrap already-encrypted values (no encryption happens here).
  CipherText(const PublicKey& pk, const uint32_t& n);
  CipherText(const PublicKey& pk, const std::vector<uint32_t>& n_v);
  CipherText(const PublicKey& pk, const BigNumber& bn);
  CipherText(const PublicKey& pk, const std::vector<BigNumber>& bn_vec);
  // Takes ownership of a freshly unmarshalled batch (addition to the
  // reference interface: avoids copying 64 Ki BigNumbers a second time).
  CipherText(const PublicKey& pk, std::vector<BigNumber>&& bn_vec);
  // back-end internal: a batch that lives in HBM (PublicKey::encrypt and the
  // operators below); the BigNumbers are built on first access
  CipherText(const PublicKey& pk, std::shared_ptr<detail::DeviceBatch> dev);

  CipherText(const CipherText& ct);
  CipherText& operator=(const CipherText& other);

  // Enc(a) + Enc(b) = Enc(a + b).  `other` holds as many elements as *this,
  // or one element that is applied to all.  Both must share the public key.
  CipherText operator+(const CipherText& other) const;
  // Enc(a) + b = Enc(a + b)
  CipherText operator+(const PlainText& other) const;
  // Enc(a) * b = Enc(a * b); `other` may again be a single element.
  CipherText operator*(const PlainText& other) const;

  // Element `idx` as a CipherText of size one.
  CipherText getCipherText(const size_t& idx) const;

  // The key the texts belong to.
  std::shared_ptr<PublicKey> getPubKey() const;

  // Cyclic rotation: element i moves to i + shift.
  CipherText rotate(int shift) const;

 private:
  std::shared_ptr<PublicKey> m_pk;

  // single-element forms used when the text holds exactly one value
  BigNumber raw_add(const BigNumber& a, const BigNumber& b) const;
  BigNumber raw_mul(const BigNumber& a, const BigNumber& b) const;
  // batch form of ct * pt
  std::vector<BigNumber> raw_mul(const std::vector<BigNumber>& a,
                                 const std::vector<BigNumber>& b) const;
};

}  // namespace ipcl
#endif  // IPCL_B200_CIPHERTEXT_HPP_
