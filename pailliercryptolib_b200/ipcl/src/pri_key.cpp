// pri_key.cpp -- ipcl::PrivateKey on the B200 back-end.
//
// Reference: ipcl/pri_key.cpp.  decryptCRT there is an OpenMP loop of two
// 4096->2048-bit reductions per ciphertext (:127-130), two modExp batches
// (:133-134) and an OpenMP loop of L-function, hp/hq multiply and CRT
// recombination (:141-145).  Here all of it runs on the device behind one
// ipclb200_decrypt call.
#include "ipcl/pri_key.hpp"

#include "device_batch.hpp"
#include "ipcl/utils/util.hpp"
#include "ipcl_b200.h"
#include "marshal.hpp"

namespace ipcl {

struct PrivateKey::DeviceKey {
  ipclb200_privkey* h = nullptr;
  ~DeviceKey() {
    if (h) ipclb200_privkey_destroy(h);
  }
};

void PrivateKey::init(const BigNumber& p, const BigNumber& q) {
  m_enable_crt = true;
  // p < q (pri_key.cpp:19-22)
  m_p = std::make_shared<BigNumber>((q < p) ? q : p);
  m_q = std::make_shared<BigNumber>((q < p) ? p : q);
  ERROR_CHECK((*m_p) * (*m_q) == *m_n,
              "PrivateKey ctor: Public key does not match p * q.");
  ERROR_CHECK(*m_p != *m_q, "PrivateKey ctor: p and q are same");
  m_pminusone = *m_p - 1;
  m_qminusone = *m_q - 1;
  m_psquare = (*m_p) * (*m_p);
  m_qsquare = (*m_q) * (*m_q);
  m_pinverse = m_q->InverseMul(*m_p);
  m_hp = computeHfun(*m_p, m_psquare);
  m_hq = computeHfun(*m_q, m_qsquare);
  m_lambda = lcm(m_pminusone, m_qminusone);
  m_x = m_n->InverseMul((modExp(*m_g, m_lambda, *m_nsquare) - 1) / (*m_n));

  // device-side key: the same constants, derived inside the C ABI, in the
  // layout the kernels read
  const int pl = static_cast<int>(m_q->words().size());
  std::vector<uint32_t> pw(static_cast<std::size_t>(pl)), qw(pw.size());
  m_p->toWords(pw.data(), pw.size());
  m_q->toWords(qw.data(), qw.size());
  auto dk = std::make_shared<DeviceKey>();
  DEVICE_CHECK(ipclb200_privkey_create(pw.data(), qw.data(), pl, &dk->h));
  m_dev = dk;
  m_isInitialized = true;
}

PrivateKey::PrivateKey(const PublicKey& pk, const BigNumber& p,
                       const BigNumber& q)
    : m_n(pk.getN()), m_nsquare(pk.getNSQ()), m_g(pk.getG()) {
  init(p, q);
}

PrivateKey::PrivateKey(const BigNumber& n, const BigNumber& p,
                       const BigNumber& q)
    : m_n(std::make_shared<BigNumber>(n)),
      m_nsquare(std::make_shared<BigNumber>(n * n)),
      m_g(std::make_shared<BigNumber>(n + 1)) {
  init(p, q);
}

PlainText PrivateKey::decrypt(const CipherText& ct) const {
  ERROR_CHECK(m_isInitialized, "decrypt: Private key is NOT initialized.");
  ERROR_CHECK(*(ct.getPubKey()->getN()) == *(this->getN()),
              "decrypt: The value of N in public key mismatch.");
  const std::size_t ct_size = ct.getSize();
  ERROR_CHECK(ct_size > 0, "decrypt: Cannot decrypt empty CipherText");

  // device-resident: the ciphertexts are (or are put) in HBM, the plaintexts
  // stay there until somebody reads them
  const int pl = static_cast<int>(m_q->words().size());
  if (detail::deviceResidentEnabled() && detail::isClassWords(2 * pl) &&
      detail::isClassWords(4 * pl)) {
    if (auto d_ct = ct.deviceBatch(4 * pl)) {
      auto d_pt = std::make_shared<detail::DeviceBatch>(ct_size, 2 * pl);
      DEVICE_CHECK(ipclb200_decrypt_batch(m_dev->h, d_ct->h, m_enable_crt ? 1 : 0, d_pt->h));
      return PlainText(std::move(d_pt));
    }
  }

  std::vector<BigNumber> pt_bn(ct_size);
  if (m_enable_crt)
    decryptCRT(pt_bn, ct.texts());
  else
    decryptRAW(pt_bn, ct.texts());
  return PlainText(std::move(pt_bn));
}

// both variants: pack the ciphertexts (reduced mod n^2 if a caller built an
// out-of-range CipherText by hand), one device call, unpack
static void device_decrypt(ipclb200_privkey* h, int pl, const BigNumber& nsq,
                           int use_crt, std::vector<BigNumber>& plaintext,
                           const std::vector<BigNumber>& ciphertext) {
  const std::size_t n = ciphertext.size();
  const int cw = 4 * pl;
  const std::vector<BigNumber>* cp = &ciphertext;
  std::vector<BigNumber> reduced;
  for (std::size_t i = 0; i < n; i++) {
    if (ciphertext[i].isNegative() ||
        static_cast<int>(ciphertext[i].words().size()) > cw) {
      if (reduced.empty()) reduced = ciphertext;
      reduced[i] = ciphertext[i] % nsq;
      cp = &reduced;
    }
  }
  detail::ScopedSlab f_ct(n * static_cast<std::size_t>(cw)),
      f_pt(n * 2 * static_cast<std::size_t>(pl));
  detail::pack(*cp, cw, f_ct.data());
  DEVICE_CHECK(ipclb200_decrypt(h, f_ct.data(), n, use_crt, f_pt.data()));
  plaintext = detail::unpack(f_pt.data(), n, 2 * pl);
}

void PrivateKey::decryptRAW(std::vector<BigNumber>& plaintext,
                            const std::vector<BigNumber>& ciphertext) const {
  device_decrypt(m_dev->h, static_cast<int>(m_q->words().size()), *m_nsquare, 0,
                 plaintext, ciphertext);
}

void PrivateKey::decryptCRT(std::vector<BigNumber>& plaintext,
                            const std::vector<BigNumber>& ciphertext) const {
  device_decrypt(m_dev->h, static_cast<int>(m_q->words().size()), *m_nsquare, 1,
                 plaintext, ciphertext);
}

// host forms of the scalar helpers, used by the constructor (pri_key.cpp:148-167)
BigNumber PrivateKey::computeCRT(const BigNumber& mp, const BigNumber& mq) const {
  BigNumber u = (mq - mp) * m_pinverse % (*m_q);
  return mp + (u * (*m_p));
}

BigNumber PrivateKey::computeLfun(const BigNumber& a, const BigNumber& b) const {
  return (a - 1) / b;
}

BigNumber PrivateKey::computeHfun(const BigNumber& a, const BigNumber& b) const {
  BigNumber xm = a - 1;
  BigNumber base = *m_g % b;
  BigNumber pm = modExp(base, xm, b);
  BigNumber lcrt = computeLfun(pm, a);
  return a.InverseMul(lcrt);
}

}  // namespace ipcl
