// Re-expression of /root/reference/test/test_serialization.cpp:13-106 plus a
// byte-level check of the archive grammar documented in serialize.hpp.
#include <cstring>
#include <sstream>

#include "check.hpp"
#include "ipcl/ipcl.hpp"

TEST(SerialTest, PublicKeyTest) {
  ipcl::KeyPair key = ipcl::generateKeypair(1024);
  ipcl::PublicKey ret_pk;
  std::ostringstream os;
  ipcl::serializer::serialize(os, key.pub_key);
  std::istringstream is(os.str());
  ipcl::serializer::deserialize(is, ret_pk);
  EXPECT_EQ(*ret_pk.getN(), *key.pub_key.getN());
  EXPECT_EQ(ret_pk.getHS(), key.pub_key.getHS());
  EXPECT_EQ(ret_pk.getRandBits(), key.pub_key.getRandBits());
  ipcl::PlainText pt(123u);
  ipcl::CipherText ct = ret_pk.encrypt(pt);
  ipcl::PlainText dt = key.priv_key.decrypt(ct);
  EXPECT_EQ(pt.getElement(0), dt.getElement(0));
}

TEST(SerialTest, PrivateKeyTest) {
  ipcl::KeyPair key = ipcl::generateKeypair(1024);
  ipcl::PrivateKey ret_sk;
  std::ostringstream os;
  ipcl::serializer::serialize(os, key.priv_key);
  std::istringstream is(os.str());
  ipcl::serializer::deserialize(is, ret_sk);
  ipcl::PlainText pt(123u);
  ipcl::CipherText ct = key.pub_key.encrypt(pt);
  ipcl::PlainText dt = ret_sk.decrypt(ct);
  EXPECT_EQ(pt.getElement(0), dt.getElement(0));
  EXPECT_EQ(ret_sk.getLambda(), key.priv_key.getLambda());
}

TEST(SerialTest, PlaintextAndCipherText) {
  std::vector<uint32_t> vals = {0, 1, 0xFFFFFFFFu, 42, 7, 1u << 31};
  ipcl::PlainText pt(vals), pt_after;
  std::ostringstream os;
  ipcl::serializer::serialize(os, pt);
  std::istringstream is(os.str());
  ipcl::serializer::deserialize(is, pt_after);
  EXPECT_EQ(pt_after.getSize(), vals.size());
  for (size_t i = 0; i < vals.size(); i++)
    EXPECT_EQ(pt.getElementVec(i)[0], pt_after.getElementVec(i)[0]);

  ipcl::KeyPair key = ipcl::generateKeypair(1024);
  ipcl::CipherText ct = key.pub_key.encrypt(pt), ct_after;
  std::ostringstream os2;
  ipcl::serializer::serialize(os2, ct);
  std::istringstream is2(os2.str());
  ipcl::serializer::deserialize(is2, ct_after);
  ipcl::PlainText dt = key.priv_key.decrypt(ct_after);
  for (size_t i = 0; i < vals.size(); i++) EXPECT_EQ(dt.getElementVec(i)[0], vals[i]);
  // a CipherText archive read back as PlainText (test_serialization.cpp:95-97)
  ipcl::PlainText as_pt;
  std::istringstream is3(os2.str());
  ipcl::serializer::deserialize(is3, as_pt);
  EXPECT_EQ(as_pt.getElement(2), ct.getElement(2));
}

TEST(SerialTest, ArchiveBytes) {
  // BigNumber 0x1_00000002, positive
  BigNumber x = "0x100000002";
  std::ostringstream os;
  ipcl::serializer::serialize(os, x);
  const unsigned char want[] = {1,                          // little endian
                                0, 0, 0, 0,                 // class version
                                2, 0, 0, 0, 0, 0, 0, 0,     // two words
                                2, 0, 0, 0, 1, 0, 0, 0,     // limbs
                                1, 0, 0, 0};                // IppsBigNumPOS
  std::string got = os.str();
  EXPECT_EQ(got.size(), sizeof(want));
  EXPECT_TRUE(got.size() == sizeof(want) && !memcmp(got.data(), want, sizeof(want)));
  BigNumber y;
  std::istringstream is(got);
  ipcl::serializer::deserialize(is, y);
  EXPECT_EQ(x, y);
  std::istringstream bad(got.substr(0, 10));
  EXPECT_THROW(ipcl::serializer::deserialize(bad, y));
}

// ---- golden archives (tests/golden/make_serial_golden.py): hand-derived from
// cereal's PortableBinary rules for the reference's save hooks.  Load them,
// check the values, save again and compare byte for byte.
#include "serial_golden.hpp"

namespace {
std::string unhex(const char* h) {
  std::string out;
  auto nib = [](char c) { return c <= '9' ? c - '0' : c - 'a' + 10; };
  for (size_t i = 0; h[i] && h[i + 1]; i += 2)
    out.push_back(static_cast<char>(nib(h[i]) * 16 + nib(h[i + 1])));
  return out;
}
BigNumber bn_hex(const char* h) { return BigNumber((std::string("0x") + h).c_str()); }
template <typename T>
std::string resave(const T& obj) {
  std::ostringstream os;
  ipcl::serializer::serialize(os, obj);
  return os.str();
}
}  // namespace

TEST(SerialTest, GoldenArchives) {
  using namespace serial_golden;
  {
    BigNumber x, y;
    std::istringstream a(unhex(k_bignum_pos)), b(unhex(k_bignum_neg));
    ipcl::serializer::deserialize(a, x);
    ipcl::serializer::deserialize(b, y);
    EXPECT_EQ(x, BigNumber("0x1234567890ABCDEF0011223344556677"));
    EXPECT_EQ(y, BigNumber::Zero() - (BigNumber("0x10000000000000000") + BigNumber(5u)));
    EXPECT_TRUE(resave(x) == unhex(k_bignum_pos));
    EXPECT_TRUE(resave(y) == unhex(k_bignum_neg));
  }
  const BigNumber p = bn_hex(k_val_p), q = bn_hex(k_val_q), n = bn_hex(k_val_n),
                  hs = bn_hex(k_val_hs);
  ipcl::PublicKey pk;
  {
    std::istringstream is(unhex(k_pubkey_djn_1024));
    ipcl::serializer::deserialize(is, pk);
    EXPECT_EQ(*pk.getN(), n);
    EXPECT_EQ(pk.getHS(), hs);
    EXPECT_EQ(pk.getBits(), 1024);
    EXPECT_TRUE(pk.isDJN());
    EXPECT_EQ(pk.getRandBits(), 512);
    EXPECT_TRUE(resave(pk) == unhex(k_pubkey_djn_1024));
  }
  ipcl::PrivateKey sk;
  {
    std::istringstream is(unhex(k_privkey_1024));
    ipcl::serializer::deserialize(is, sk);
    EXPECT_EQ(*sk.getP(), p);
    EXPECT_EQ(*sk.getQ(), q);
    EXPECT_TRUE(resave(sk) == unhex(k_privkey_1024));
  }
  {
    ipcl::PlainText pt;
    std::istringstream is(unhex(k_plaintext_3));
    ipcl::serializer::deserialize(is, pt);
    EXPECT_EQ(pt.getSize(), (size_t)3);
    for (size_t i = 0; i < 3; i++) EXPECT_EQ(pt.getElement(i), bn_hex(k_pt_values[i]));
    EXPECT_TRUE(resave(pt) == unhex(k_plaintext_3));
  }
  {
    ipcl::CipherText ct;
    std::istringstream is(unhex(k_ciphertext_2));
    ipcl::serializer::deserialize(is, ct);
    EXPECT_EQ(ct.getSize(), (size_t)2);
    for (size_t i = 0; i < 2; i++) EXPECT_EQ(ct.getElement(i), bn_hex(k_ct_values[i]));
    EXPECT_EQ(*ct.getPubKey()->getN(), n);
    EXPECT_TRUE(resave(ct) == unhex(k_ciphertext_2));
    // ct[0] = (n+1)^3: a valid (unobfuscated) encryption of 3 under the golden key
    ipcl::PlainText dt = sk.decrypt(ct);
    EXPECT_EQ(dt.getElement(0), BigNumber(3u));
  }
}
