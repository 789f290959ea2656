// Re-expression of /root/reference/test/test_cryptography.cpp against the
// ipcl:: API of this repository: same shapes (10 x 20 values under 4 threads,
// 20 values, the ISO/IEC 18033-6 known-answer vectors with injected randoms).
#include <omp.h>

#include <climits>
#include <random>
#include <vector>

#include "check.hpp"
#include "ipcl/ipcl.hpp"
#include "iso_vectors.hpp"

constexpr int SELF_DEF_NUM_VALUES = 20;

// test_cryptography.cpp:18-65 -- encrypt/decrypt called concurrently on one key
TEST(CryptoTest, CryptoTest_APPLEVEL_OMP) {
  const uint32_t num_values = SELF_DEF_NUM_VALUES;
  const int num_pt = 10;
  ipcl::KeyPair key = ipcl::generateKeypair(2048, true);
  std::vector<std::vector<uint32_t>> exp_value(num_pt,
                                               std::vector<uint32_t>(num_values));
  std::vector<ipcl::PlainText> pt(num_pt), dt(num_pt);
  std::vector<ipcl::CipherText> ct(num_pt);
  std::mt19937 rng(12345);
  std::uniform_int_distribution<uint32_t> dist(0, UINT_MAX);
  for (int i = 0; i < num_pt; i++) {
    for (uint32_t j = 0; j < num_values; j++) exp_value[i][j] = dist(rng);
    pt[i] = ipcl::PlainText(exp_value[i]);
  }
  ipcl::setHybridOff();
#pragma omp parallel for num_threads(4)
  for (int i = 0; i < num_pt; i++) ct[i] = key.pub_key.encrypt(pt[i]);
#pragma omp parallel for num_threads(4)
  for (int i = 0; i < num_pt; i++) dt[i] = key.priv_key.decrypt(ct[i]);
  for (int i = 0; i < num_pt; i++)
    for (uint32_t j = 0; j < num_values; j++) {
      std::vector<uint32_t> v = dt[i].getElementVec(j);
      EXPECT_EQ(v[0], exp_value[i][j]);
    }
}

// test_cryptography.cpp:67-97
TEST(CryptoTest, CryptoTest) {
  const uint32_t num_values = SELF_DEF_NUM_VALUES;
  ipcl::KeyPair key = ipcl::generateKeypair(2048, true);
  std::vector<uint32_t> exp_value(num_values);
  std::mt19937 rng(777);
  std::uniform_int_distribution<uint32_t> dist(0, UINT_MAX);
  for (auto& v : exp_value) v = dist(rng);
  ipcl::PlainText pt(exp_value);
  ipcl::setHybridRatio(0.5f);
  ipcl::CipherText ct = key.pub_key.encrypt(pt);
  ipcl::PlainText dt = key.priv_key.decrypt(ct);
  for (uint32_t i = 0; i < num_values; i++) {
    std::vector<uint32_t> v = dt.getElementVec(i);
    EXPECT_EQ(v[0], exp_value[i]);
  }
  // non-CRT decrypt gives the same plaintexts (PrivateKey::enableCRT, pri_key.hpp:54)
  key.priv_key.enableCRT(false);
  ipcl::PlainText dr = key.priv_key.decrypt(ct);
  for (uint32_t i = 0; i < num_values; i++) EXPECT_EQ(dr.getElement(i), pt.getElement(i));
}

// 1024-bit key, batch 8 round trip: BASELINE.json configs[0]
TEST(CryptoTest, Key1024Batch8) {
  ipcl::KeyPair key = ipcl::generateKeypair(1024, true);
  std::vector<uint32_t> vals = {0, 1, 2, 0xFFFFFFFFu, 12345, 99, 7, 0x80000000u};
  ipcl::PlainText pt(vals);
  ipcl::CipherText ct = key.pub_key.encrypt(pt);
  ipcl::PlainText dt = key.priv_key.decrypt(ct);
  for (size_t i = 0; i < vals.size(); i++) EXPECT_EQ(dt.getElementVec(i)[0], vals[i]);
  EXPECT_EQ(key.pub_key.getN()->BitSize(), 1024);
}

// test_cryptography.cpp:99-241 -- the reference's only known-answer test:
// non-DJN key from the ISO primes, 21 values, injected randoms; pins
// enc(m0;r0)=c1, enc(m1;r1)=c2, c1*c2 mod n^2, dec(c1*c2)=m0+m1
TEST(CryptoTest, ISO_IEC_18033_6_ComplianceTest) {
  const uint32_t count = SELF_DEF_NUM_VALUES + 1;
  const BigNumber p(iso::kIsoP), q(iso::kIsoQ);
  const BigNumber n = p * q;
  ipcl::PublicKey pk(n, n.BitSize());
  ipcl::PrivateKey sk(pk, p, q);

  std::vector<BigNumber> plain(count, BigNumber(iso::kIsoM0));
  std::vector<BigNumber> randoms(count, BigNumber(iso::kIsoR0));
  plain[1] = BigNumber(iso::kIsoM1);
  randoms[1] = BigNumber(iso::kIsoR1);

  ipcl::setHybridOff();
  pk.setRandom(randoms);
  ipcl::CipherText ct = pk.encrypt(ipcl::PlainText(plain));
  ipcl::PlainText dt = sk.decrypt(ct);
  for (uint32_t i = 0; i < count; i++) EXPECT_EQ(dt.getElement(i), plain[i]);

  auto hex = [](const char* s) {
    std::string h;
    BigNumber(s).num2hex(h);
    return h;
  };
  EXPECT_EQ(hex(iso::kIsoC1), ct.getElementHex(0));
  EXPECT_EQ(hex(iso::kIsoC2), ct.getElementHex(1));
  for (uint32_t i = 2; i < count; i++) EXPECT_EQ(hex(iso::kIsoC1), ct.getElementHex(i));

  ipcl::CipherText sum = ipcl::CipherText(pk, ct.getElement(0)) +
                         ipcl::CipherText(pk, ct.getElement(1));
  EXPECT_EQ(hex(iso::kIsoC1C2), sum.getElementHex(0));
  EXPECT_EQ(hex(iso::kIsoM1M2), sk.decrypt(sum).getElementHex(0));
  sk.enableCRT(false);
  EXPECT_EQ(hex(iso::kIsoM1M2), sk.decrypt(sum).getElementHex(0));
}

// error behaviour at the API boundary (ERROR_CHECK -> std::runtime_error)
TEST(CryptoTest, ErrorConvention) {
  ipcl::PublicKey pk;
  EXPECT_THROW(pk.encrypt(ipcl::PlainText(5u)));
  ipcl::KeyPair key = ipcl::generateKeypair(1024, true);
  EXPECT_THROW(key.pub_key.encrypt(ipcl::PlainText()));
  ipcl::KeyPair other = ipcl::generateKeypair(1024, true);
  ipcl::CipherText ct = key.pub_key.encrypt(ipcl::PlainText(7u));
  EXPECT_THROW(other.priv_key.decrypt(ct));
  EXPECT_THROW(ipcl::qatModExp({BigNumber(2u)}, {BigNumber(3u)}, {BigNumber(5u)}));
  EXPECT_THROW(ipcl::modExp(BigNumber(2u), BigNumber(3u), BigNumber(10u)));  // even modulus
  try {
    ipcl::modExp(BigNumber(2u), BigNumber(3u), BigNumber(10u));
  } catch (const std::runtime_error& e) {
    std::string w = e.what();
    EXPECT_TRUE(w.find("\nFile: ") == 0 && w.find("\nLine: ") != std::string::npos &&
                w.find("\nError: ") != std::string::npos);
  }
  EXPECT_THROW(ipcl::generateKeypair(100, true));
  EXPECT_THROW(ipcl::generateKeypair(1026, true));
}
