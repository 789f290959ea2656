// pub_key.hpp -- ipcl::PublicKey (ipcl/include/ipcl/pub_key.hpp:18-193).
// encrypt() is one submission to the fused encrypt kernel; the per-key device
// constants (n^2 Montgomery constants, n*R, hs in Montgomery form, the
// fixed-base table for hs) live behind an ipclb200_pubkey handle that is
// created on first use and shared by copies of the key.
#ifndef IPCL_B200_PUB_KEY_HPP_
#define IPCL_B200_PUB_KEY_HPP_

#include <memory>
#include <vector>

#include "ipcl/bignum.h"
#include "ipcl/plaintext.hpp"

struct ipclb200_pubkey;

namespace ipcl {

class CipherText;

class PublicKey {
 public:
  PublicKey() = default;
  ~PublicKey() = default;

  explicit PublicKey(const BigNumber& n, int bits = 1024,
                     bool enableDJN_ = false);
  explicit PublicKey(const Ipp32u n, int bits = 1024, bool enableDJN_ = false)
      : PublicKey(BigNumber(n), bits, enableDJN_) {}

  void enableDJN();
  void setDJN(const BigNumber& hs, int randbit);

  CipherText encrypt(const PlainText& plaintext, bool make_secure = true) const;

  std::shared_ptr<BigNumber> getN() const { return m_n; }
  std::shared_ptr<BigNumber> getNSQ() const { return m_nsquare; }
  std::shared_ptr<BigNumber> getG() const { return m_g; }
  int getBits() const { return m_bits; }
  int getDwords() const { return m_dwords; }

  void applyObfuscator(std::vector<BigNumber>& ciphertext) const;
  void setRandom(const std::vector<BigNumber>& r);
  void setHS(const BigNumber& hs);

  bool isDJN() const { return m_enable_DJN; }
  BigNumber getHS() const {
    if (m_enable_DJN) return m_hs;
    return BigNumber::Zero();
  }
  int getRandBits() const {
    if (m_enable_DJN) return m_randbits;
    return -1;
  }
  bool isInitialized() { return m_isInitialized; }

  void create(const BigNumber& n, int bits, bool enableDJN_ = false);
  void create(const BigNumber& n, int bits, const BigNumber& hs, int randbits);

  const void* addr = static_cast<const void*>(this);

 private:
  bool m_isInitialized = false;
  std::shared_ptr<BigNumber> m_n;
  std::shared_ptr<BigNumber> m_g;
  std::shared_ptr<BigNumber> m_nsquare;
  int m_bits = 0;
  int m_dwords = 0;
  BigNumber m_hs;
  int m_randbits = 0;
  bool m_enable_DJN = false;
  std::vector<BigNumber> m_r;
  bool m_testv = false;

  // device-side key object; rebuilt when n / hs change
  struct DeviceKey;
  mutable std::shared_ptr<DeviceKey> m_dev;
  ipclb200_pubkey* deviceKey() const;
  void flatRandoms(std::size_t sz, std::vector<uint32_t>& f_r, int& r_words) const;
  void resetDeviceKey();
  bool deviceRandoms() const;

  std::vector<BigNumber> raw_encrypt(const std::vector<BigNumber>& pt,
                                     bool make_secure = true) const;
  std::vector<BigNumber> getDJNObfuscator(std::size_t sz) const;
  std::vector<BigNumber> getNormalObfuscator(std::size_t sz) const;
  // the randoms the obfuscator consumes (injected by setRandom or fresh)
  std::vector<BigNumber> drawRandoms(std::size_t sz) const;
};

}  // namespace ipcl
#endif  // IPCL_B200_PUB_KEY_HPP_
