// pub_key.cpp -- ipcl::PublicKey on the B200 back-end.
//
// Reference: ipcl/pub_key.cpp.  There encrypt() is a serial host loop
// ct = (n*pt + 1) % n^2 (:105), one modExp batch for the obfuscator (:63 DJN
// hs^r, :79 r^n) and a second serial host loop of mul + mod (:88-89).  Here all
// three stages are one fused kernel behind ipclb200_encrypt; the host only
// marshals plaintexts and randoms.
#include "ipcl/pub_key.hpp"

#include <mutex>

#include "device_batch.hpp"
#include "drbg.hpp"
#include "ipcl/ciphertext.hpp"
#include "ipcl/mod_exp.hpp"
#include "ipcl/utils/util.hpp"
#include "ipcl_b200.h"
#include "marshal.hpp"

namespace ipcl {

// The device-side key is built lazily, on the first encrypt.  The holder is
// allocated when the key VALUE is set (constructor, create, setHS ...), so every
// copy of a PublicKey taken afterwards -- each CipherText stores one
// (ipcl/ciphertext.cpp:14) -- shares the one device key and its fixed-base table
// instead of building its own.
struct PublicKey::DeviceKey {
  std::mutex mu;  // guards lazy creation; encrypt() is const and is called
                  // concurrently on one key (test/test_cryptography.cpp:45-57)
  ipclb200_pubkey* h = nullptr;
  ~DeviceKey() {
    if (h) ipclb200_pubkey_destroy(h);
  }
};

void PublicKey::resetDeviceKey() { m_dev = std::make_shared<DeviceKey>(); }

PublicKey::PublicKey(const BigNumber& n, int bits, bool enableDJN_)
    : m_n(std::make_shared<BigNumber>(n)),
      m_g(std::make_shared<BigNumber>(*m_n + 1)),
      m_nsquare(std::make_shared<BigNumber>((*m_n) * (*m_n))),
      m_bits(bits),
      m_dwords(BITSIZE_DWORD(bits * 2)),
      m_hs(0u),
      m_randbits(0),
      m_enable_DJN(false),
      m_testv(false) {
  resetDeviceKey();
  if (enableDJN_) this->enableDJN();
  m_isInitialized = true;
}

// hs = (-x^2)^n mod n^2 for a random x coprime to n (pub_key.cpp:32-49)
void PublicKey::enableDJN() {
  BigNumber rmod;
  for (;;) {
    BigNumber rand = getRandomBN(m_n->BitSize() + 128);
    rmod = rand % (*m_n);
    if (rand.gcd(*m_n) == BigNumber::One()) break;
  }
  BigNumber h = (rmod * rmod * BigNumber(static_cast<Ipp32s>(-1))) % (*m_n);
  m_hs = modExp(h, *m_n, *m_nsquare);
  m_randbits = m_bits >> 1;
  m_enable_DJN = true;
  resetDeviceKey();
}

void PublicKey::setDJN(const BigNumber& hs, int randbit) {
  if (m_enable_DJN) return;
  m_hs = hs;
  m_randbits = randbit;
  m_enable_DJN = true;
  resetDeviceKey();
}

void PublicKey::setRandom(const std::vector<BigNumber>& r) {
  m_r.insert(m_r.end(), r.begin(), r.end());
  m_testv = true;
}

void PublicKey::setHS(const BigNumber& hs) {
  m_hs = hs;
  resetDeviceKey();
}

void PublicKey::create(const BigNumber& n, int bits, bool enableDJN_) {
  m_n = std::make_shared<BigNumber>(n);
  m_g = std::make_shared<BigNumber>(*m_n + 1);
  m_nsquare = std::make_shared<BigNumber>((*m_n) * (*m_n));
  m_bits = bits;
  m_dwords = BITSIZE_DWORD(bits * 2);
  m_enable_DJN = false;
  m_hs = BigNumber::Zero();
  m_randbits = 0;
  m_r.clear();
  m_testv = false;
  resetDeviceKey();
  if (enableDJN_) this->enableDJN();
  m_isInitialized = true;
}

void PublicKey::create(const BigNumber& n, int bits, const BigNumber& hs,
                       int randbits) {
  create(n, bits, false);
  m_enable_DJN = true;
  m_hs = hs;
  m_randbits = randbits;
  resetDeviceKey();
}

ipclb200_pubkey* PublicKey::deviceKey() const {
  ERROR_CHECK(m_dev != nullptr, "encrypt: Public key is NOT initialized.");
  std::lock_guard<std::mutex> lk(m_dev->mu);
  if (!m_dev->h) {
    DeviceKey* dk = m_dev.get();
    const int nl = static_cast<int>(m_n->words().size());
    std::vector<uint32_t> n_w(static_cast<std::size_t>(nl));
    m_n->toWords(n_w.data(), n_w.size());
    std::vector<uint32_t> hs_w;
    if (m_enable_DJN) {
      BigNumber hs = m_hs % (*m_nsquare);
      hs_w.resize(2 * static_cast<std::size_t>(nl));
      hs.toWords(hs_w.data(), hs_w.size());
    }
    DEVICE_CHECK(ipclb200_pubkey_create(n_w.data(), nl,
                                        m_enable_DJN ? hs_w.data() : nullptr,
                                        m_randbits, &dk->h));
  }
  return m_dev->h;
}

std::vector<BigNumber> PublicKey::drawRandoms(std::size_t sz) const {
  std::vector<BigNumber> r;
  if (m_testv) {
    ERROR_CHECK(m_r.size() >= sz,
                "encrypt: fewer injected randoms (setRandom) than plaintexts");
    r.assign(m_r.begin(), m_r.begin() + static_cast<std::ptrdiff_t>(sz));
    return r;
  }
  r.resize(sz);
  if (m_enable_DJN) {
    for (auto& x : r) x = getRandomBN(m_randbits);  // pub_key.cpp:59-61
  } else {
    const BigNumber nm1 = *m_n - 1;
    for (auto& x : r) x = getRandomBN(m_bits) % nm1 + 1;  // pub_key.cpp:74-77
  }
  return r;
}

std::vector<BigNumber> PublicKey::getDJNObfuscator(std::size_t sz) const {
  std::vector<BigNumber> base(sz, m_hs);
  std::vector<BigNumber> sq(sz, *m_nsquare);
  return modExp(base, drawRandoms(sz), sq);
}

std::vector<BigNumber> PublicKey::getNormalObfuscator(std::size_t sz) const {
  std::vector<BigNumber> sq(sz, *m_nsquare);
  std::vector<BigNumber> pown(sz, *m_n);
  return modExp(drawRandoms(sz), pown, sq);
}

void PublicKey::applyObfuscator(std::vector<BigNumber>& ciphertext) const {
  const std::size_t sz = ciphertext.size();
  if (sz == 0) return;
  std::vector<BigNumber> obf =
      m_enable_DJN ? getDJNObfuscator(sz) : getNormalObfuscator(sz);
  ciphertext = modMul(ciphertext, obf, *m_nsquare);
}

// the randoms of one batch as a flat r_words-strided limb buffer
// (pub_key.cpp:51-80: getRandomBN(m_randbits) per element for DJN, a value in
// [1, n-1] otherwise; setRandom() injects fixed ones, :92-95)
void PublicKey::flatRandoms(std::size_t sz, std::vector<uint32_t>& f_r,
                            int& r_words) const {
  const int nl = static_cast<int>(m_n->words().size());
  if (m_enable_DJN && !m_testv) {
    // fresh DJN randoms drawn straight into the flat buffer with one entropy
    // call for the batch
    r_words = (m_randbits + 31) / 32;
    f_r.resize(sz * static_cast<std::size_t>(r_words));
    rand32u(f_r);
    if (m_randbits % 32) {
      const uint32_t mask = (1u << (m_randbits % 32)) - 1u;
      for (std::size_t i = 0; i < sz; i++)
        f_r[i * static_cast<std::size_t>(r_words) + r_words - 1] &= mask;
    }
    return;
  }
  std::vector<BigNumber> r = drawRandoms(sz);
  for (auto& x : r) {
    ERROR_CHECK(!x.isNegative(), "encrypt: negative random");
    if (static_cast<int>(x.words().size()) > 2 * nl) {
      // r is the BASE of r^n for a non-DJN key: r mod n^2 gives the same
      // obfuscator.  For a DJN key it is the EXPONENT of hs^r and must not be
      // reduced (hs^(r mod n^2) != hs^r)
      ERROR_CHECK(!m_enable_DJN,
                  "encrypt: injected DJN random is wider than n^2");
      x = x % (*m_nsquare);
    }
  }
  r_words = detail::maxWords(r);
  detail::pack(r, r_words, f_r);
}

// fresh DJN randoms (not injected by setRandom) are drawn on the device
bool PublicKey::deviceRandoms() const {
  return m_enable_DJN && !m_testv && m_randbits > 0 && detail::deviceRandomEnabled();
}

// (n*pt + 1) mod n^2 only depends on pt mod n: negative or oversize plaintexts
// are brought into [0, n) so they fit the n-word slot
static const std::vector<BigNumber>& reducedPlain(
    const std::vector<BigNumber>& pt, const BigNumber& n, int nl,
    std::vector<BigNumber>& reduced) {
  const std::vector<BigNumber>* pp = &pt;
  for (std::size_t i = 0; i < pt.size(); i++) {
    if (pt[i].isNegative() || static_cast<int>(pt[i].words().size()) > nl) {
      if (reduced.empty()) reduced = pt;
      reduced[i] = pt[i] % n;
      pp = &reduced;
    }
  }
  return *pp;
}

std::vector<BigNumber> PublicKey::raw_encrypt(const std::vector<BigNumber>& pt,
                                              bool make_secure) const {
  const std::size_t sz = pt.size();
  const int nl = static_cast<int>(m_n->words().size());
  ipclb200_pubkey* dev = deviceKey();
  std::vector<BigNumber> reduced;
  const std::vector<BigNumber>& pp = reducedPlain(pt, *m_n, nl, reduced);
  std::vector<uint32_t> f_r;
  detail::ScopedSlab f_pt(sz * static_cast<std::size_t>(nl)),
      f_ct(sz * 2 * static_cast<std::size_t>(nl));
  detail::pack(pp, nl, f_pt.data());
  int r_words = 0;
  if (make_secure && deviceRandoms()) {
    uint32_t key[8], nonce[3];
    detail::freshDrbgSeed(key, nonce);
    DEVICE_CHECK(ipclb200_encrypt_drbg(dev, f_pt.data(), nl, sz, key, nonce, f_ct.data()));
    return detail::unpack(f_ct.data(), sz, 2 * nl);
  }
  if (make_secure) flatRandoms(sz, f_r, r_words);
  DEVICE_CHECK(ipclb200_encrypt(dev, f_pt.data(), nl,
                                make_secure ? f_r.data() : nullptr, r_words, sz,
                                make_secure ? 1 : 0, f_ct.data()));
  return detail::unpack(f_ct.data(), sz, 2 * nl);
}

CipherText PublicKey::encrypt(const PlainText& pt, bool make_secure) const {
  ERROR_CHECK(m_isInitialized, "encrypt: Public key is NOT initialized.");
  const std::size_t pt_size = pt.getSize();
  ERROR_CHECK(pt_size > 0, "encrypt: Cannot encrypt empty PlainText");
  const int nl = static_cast<int>(m_n->words().size());
  if (detail::deviceResidentEnabled() && detail::isClassWords(2 * nl)) {
    // device-resident: plaintexts and randoms go up once, the ciphertexts stay
    // in HBM until somebody reads them
    std::shared_ptr<detail::DeviceBatch> d_pt;
    if (!pt.isHostMaterialized()) {
      d_pt = pt.deviceBatch(nl);
    } else {
      std::vector<BigNumber> reduced;
      d_pt = detail::DeviceBatch::fromHost(
          reducedPlain(pt.texts(), *m_n, nl, reduced), nl);
    }
    if (d_pt) {
      ipclb200_pubkey* dev = deviceKey();
      std::shared_ptr<detail::DeviceBatch> d_r;
      int r_words = 0;
      if (make_secure && deviceRandoms()) {
        // fresh DJN randoms never exist on the host: one OS-entropy key per
        // call, expanded in HBM (drbg.hpp)
        uint32_t key[8], nonce[3];
        detail::freshDrbgSeed(key, nonce);
        r_words = (m_randbits + 31) / 32;
        d_r = std::make_shared<detail::DeviceBatch>(pt_size, r_words);
        DEVICE_CHECK(ipclb200_batch_random(d_r->h, m_randbits, key, nonce));
      } else if (make_secure) {
        std::vector<uint32_t> f_r;
        flatRandoms(pt_size, f_r, r_words);
        d_r = std::make_shared<detail::DeviceBatch>(pt_size, r_words);
        d_r->upload(f_r.data(), r_words);
      }
      auto d_ct = std::make_shared<detail::DeviceBatch>(pt_size, 2 * nl);
      DEVICE_CHECK(ipclb200_encrypt_batch(dev, d_pt->h, make_secure ? d_r->h : nullptr, 0,
                                          make_secure ? 1 : 0, d_ct->h));
      return CipherText(*this, std::move(d_ct));
    }
  }
  return CipherText(*this, raw_encrypt(pt.texts(), make_secure));
}

}  // namespace ipcl
