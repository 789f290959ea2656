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
