// pri_key.hpp -- ipcl::PrivateKey (ipcl/include/ipcl/pri_key.hpp:29-196).
// decrypt() is one submission to the fused CRT-decrypt kernels (or the RAW
// path when enableCRT(false)).
#ifndef IPCL_B200_PRI_KEY_HPP_
#define IPCL_B200_PRI_KEY_HPP_

#include <memory>
#include <vector>

#include "ipcl/ciphertext.hpp"
#include "ipcl/mod_exp.hpp"
#include "ipcl/plaintext.hpp"

struct ipclb200_privkey;

namespace ipcl {

// least common multiple (pri_key.hpp:23-27)
static inline BigNumber lcm(const BigNumber& p, const BigNumber& q) {
  return p * q / p.gcd(q);
}

class PrivateKey {
 public:
  PrivateKey() = default;
  ~PrivateKey() = default;

  PrivateKey(const PublicKey& pk, const BigNumber& p, const BigNumber& q);
  PrivateKey(const BigNumber& n, const BigNumber& p, const BigNumber& q);

  void enableCRT(bool crt) { m_enable_crt = crt; }
  PlainText decrypt(const CipherText& ciphertext) const;

  const void* addr = static_cast<const void*>(this);

  std::shared_ptr<BigNumber> getN() const { return m_n; }
  std::shared_ptr<BigNumber> getP() const { return m_p; }
  std::shared_ptr<BigNumber> getQ() const { return m_q; }
  BigNumber getLambda() const { return m_lambda; }
  bool isInitialized() { return m_isInitialized; }

 private:
  void init(const BigNumber& p, const BigNumber& q);

  bool m_isInitialized = false;
  bool m_enable_crt = false;
  std::shared_ptr<BigNumber> m_n;
  std::shared_ptr<BigNumber> m_nsquare;
  std::shared_ptr<BigNumber> m_g;
  std::shared_ptr<BigNumber> m_p;
  std::shared_ptr<BigNumber> m_q;
  BigNumber m_pminusone;
  BigNumber m_qminusone;
  BigNumber m_psquare;
  BigNumber m_qsquare;
  BigNumber m_pinverse;
  BigNumber m_hp;
  BigNumber m_hq;
  BigNumber m_lambda;
  BigNumber m_x;

  struct DeviceKey;
  std::shared_ptr<DeviceKey> m_dev;

  BigNumber computeLfun(const BigNumber& a, const BigNumber& b) const;
  BigNumber computeHfun(const BigNumber& a, const BigNumber& b) const;
  BigNumber computeCRT(const BigNumber& mp, const BigNumber& mq) const;
  void decryptRAW(std::vector<BigNumber>& plaintext,
                  const std::vector<BigNumber>& ciphertext) const;
  void decryptCRT(std::vector<BigNumber>& plaintext,
                  const std::vector<BigNumber>& ciphertext) const;
};

}  // namespace ipcl
#endif  // IPCL_B200_PRI_KEY_HPP_
