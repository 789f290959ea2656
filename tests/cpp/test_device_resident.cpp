// Device-resident texts: encrypt, the CipherText
// operators and decrypt keep their batches in HBM and build the
// vector<BigNumber> only when a caller reads values.  These cases check that
// (a) a chain encrypt -> + -> * -> decrypt never materialises an intermediate,
// (b) every result equals plain BigNumber arithmetic mod n, (c) texts built on
// the host mix freely with resident ones, (d) copies are independent once one
// of them is written to.  The reference has no such mode; the values are
// pinned by the homomorphic identities its own tests use
// (test/test_ops.cpp:126-608).
#include <random>
#include <vector>

#include "check.hpp"
#include "ipcl/ipcl.hpp"

namespace {

struct Fx {
  ipcl::KeyPair key;
  std::vector<BigNumber> a, b, k;
};

Fx& fx() {
  static Fx* f = [] {
    auto* g = new Fx{ipcl::generateKeypair(2048), {}, {}, {}};
    std::mt19937 rng(8200);
    const int count = 77;
    for (int i = 0; i < count; i++) {
      std::vector<uint32_t> wa(60), wb(60);
      for (auto& w : wa) w = rng();
      for (auto& w : wb) w = rng();
      g->a.emplace_back(wa.data(), 60);
      g->b.emplace_back(wb.data(), 60);
      g->k.emplace_back((uint32_t)rng());
    }
    return g;
  }();
  return *f;
}

}  // namespace

TEST(DeviceResidentTest, ChainStaysInHBM) {
  Fx& f = fx();
  const BigNumber& n = *f.key.pub_key.getN();
  ipcl::PlainText pa(f.a), pb(f.b), pk(f.k);
  ipcl::CipherText ca = f.key.pub_key.encrypt(pa);
  ipcl::CipherText cb = f.key.pub_key.encrypt(pb);
  EXPECT_TRUE(ca.isDeviceResident());
  EXPECT_TRUE(!ca.isHostMaterialized());
  ipcl::CipherText sum = ca + cb;
  ipcl::CipherText prod = sum * pk;
  ipcl::CipherText mixed = prod + pa;  // ct + pt
  EXPECT_TRUE(!sum.isHostMaterialized());
  EXPECT_TRUE(!prod.isHostMaterialized());
  EXPECT_TRUE(!mixed.isHostMaterialized());
  ipcl::PlainText dt = f.key.priv_key.decrypt(mixed);
  EXPECT_TRUE(dt.isDeviceResident());
  EXPECT_TRUE(!dt.isHostMaterialized());
  EXPECT_TRUE(!ca.isHostMaterialized());  // nothing upstream was read back
  EXPECT_EQ(dt.getSize(), f.a.size());
  for (size_t i = 0; i < f.a.size(); i++) {
    BigNumber want = ((f.a[i] + f.b[i]) * f.k[i] + f.a[i]) % n;
    EXPECT_EQ(dt.getElement(i), want);
  }
  // single-element reads come from the flat host image: ONE download, no
  // vector<BigNumber> for the whole batch
  EXPECT_TRUE(!dt.isHostMaterialized());
  EXPECT_EQ(dt.getTexts().size(), f.a.size());
  EXPECT_TRUE(dt.isHostMaterialized());
  // the intermediate values are still readable afterwards and are valid
  // ciphertexts of what they should be
  ipcl::PlainText ds = f.key.priv_key.decrypt(ipcl::CipherText(f.key.pub_key, sum.getTexts()));
  for (size_t i = 0; i < f.a.size(); i++) EXPECT_EQ(ds.getElement(i), (f.a[i] + f.b[i]) % n);
}

TEST(DeviceResidentTest, HostAndResidentOperandsMix) {
  Fx& f = fx();
  const BigNumber& n = *f.key.pub_key.getN();
  ipcl::CipherText ca = f.key.pub_key.encrypt(ipcl::PlainText(f.a));
  ipcl::CipherText cb = f.key.pub_key.encrypt(ipcl::PlainText(f.b));
  // a CipherText rebuilt from host BigNumbers (what a deserialised text is)
  ipcl::CipherText cb_host(f.key.pub_key, cb.getTexts());
  EXPECT_TRUE(!cb_host.isDeviceResident());
  ipcl::PlainText d1 = f.key.priv_key.decrypt(ca + cb_host);
  ipcl::PlainText d2 = f.key.priv_key.decrypt(cb_host + ca);
  for (size_t i = 0; i < f.a.size(); i++) {
    EXPECT_EQ(d1.getElement(i), (f.a[i] + f.b[i]) % n);
    EXPECT_EQ(d2.getElement(i), (f.a[i] + f.b[i]) % n);
  }
  // broadcast of a size-1 operand, resident left side
  ipcl::CipherText one = f.key.pub_key.encrypt(ipcl::PlainText(f.b[0]));
  ipcl::PlainText d3 = f.key.priv_key.decrypt(ca + one);
  ipcl::PlainText d4 = f.key.priv_key.decrypt(ca * ipcl::PlainText(f.k[0]));
  for (size_t i = 0; i < f.a.size(); i++) {
    EXPECT_EQ(d3.getElement(i), (f.a[i] + f.b[0]) % n);
    EXPECT_EQ(d4.getElement(i), (f.a[i] * f.k[0]) % n);
  }
  // a decrypted (resident) PlainText used again as an exponent and re-encrypted
  ipcl::PlainText pk_res = f.key.priv_key.decrypt(f.key.pub_key.encrypt(ipcl::PlainText(f.k)));
  EXPECT_TRUE(!pk_res.isHostMaterialized());
  ipcl::PlainText d5 = f.key.priv_key.decrypt(ca * pk_res);
  ipcl::PlainText d6 = f.key.priv_key.decrypt(f.key.pub_key.encrypt(pk_res));
  for (size_t i = 0; i < f.a.size(); i++) {
    EXPECT_EQ(d5.getElement(i), (f.a[i] * f.k[i]) % n);
    EXPECT_EQ(d6.getElement(i), f.k[i]);
  }
}

TEST(DeviceResidentTest, CopiesAndMutation) {
  Fx& f = fx();
  ipcl::CipherText ca = f.key.pub_key.encrypt(ipcl::PlainText(f.a));
  ipcl::CipherText copy = ca;  // shares the resident batch
  EXPECT_TRUE(copy.isDeviceResident() && !copy.isHostMaterialized());
  ipcl::CipherText other = f.key.pub_key.encrypt(ipcl::PlainText(f.b));
  BigNumber repl = other.getElement(3);
  copy[3] = repl;  // write through: the copy leaves HBM, the original does not
  EXPECT_TRUE(!copy.isDeviceResident());
  EXPECT_TRUE(ca.isDeviceResident());
  ipcl::PlainText d_orig = f.key.priv_key.decrypt(ca);
  ipcl::PlainText d_copy = f.key.priv_key.decrypt(copy);
  for (size_t i = 0; i < f.a.size(); i++) {
    EXPECT_EQ(d_orig.getElement(i), f.a[i]);
    EXPECT_EQ(d_copy.getElement(i), i == 3 ? f.b[3] : f.a[i]);
  }
  // rotate / getChunk / getCipherText / assignment on resident texts
  ipcl::CipherText rot = ca.rotate(5);
  ipcl::PlainText d_rot = f.key.priv_key.decrypt(rot);
  for (size_t i = 0; i < f.a.size(); i++)
    EXPECT_EQ(d_rot.getElement((i + 5) % f.a.size()), f.a[i]);
  ipcl::CipherText assigned;
  assigned = other;
  EXPECT_EQ(assigned.getSize(), f.b.size());
  EXPECT_EQ(f.key.priv_key.decrypt(assigned.getCipherText(7)).getElement(0), f.b[7]);
  EXPECT_EQ(other.getChunk(2, 3).size(), (size_t)3);
  // decryptRAW on a resident batch
  ipcl::PrivateKey raw = f.key.priv_key;
  raw.enableCRT(false);
  ipcl::PlainText d_raw = raw.decrypt(ca);
  for (size_t i = 0; i < f.a.size(); i++) EXPECT_EQ(d_raw.getElement(i), f.a[i]);
}
