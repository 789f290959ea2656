// Homomorphic-operation tests: the twelve OperationTest cases of
// /root/reference/test/test_ops.cpp:126-608 (ct+ct, ct+pt, ct*pt incl. x0,
// pt+ct, pt*ct, element-wise through size-1 texts and as arrays, and
// a + b*2 + b), 14 random 32-bit pairs on a DJN 2048-bit key, checking the low
// 64 bits of the decrypted value as the reference does (:157-163).
#include <climits>
#include <functional>
#include <random>
#include <vector>

#include "check.hpp"
#include "ipcl/ipcl.hpp"

namespace {

constexpr int kNumValues = 14;

struct Fixture {
  ipcl::KeyPair key;
  std::vector<uint32_t> v1, v2;
  ipcl::PlainText pt1, pt2;
  ipcl::CipherText ct1, ct2;
};

Fixture& fx() {
  static Fixture* f = [] {
    auto* g = new Fixture{ipcl::generateKeypair(2048), {}, {}, {}, {}, {}, {}};
    std::mt19937 rng(20231017);
    std::uniform_int_distribution<uint32_t> dist(0, UINT_MAX);
    g->v1.resize(kNumValues);
    g->v2.resize(kNumValues);
    for (int i = 0; i < kNumValues; i++) {
      g->v1[i] = dist(rng);
      g->v2[i] = dist(rng);
    }
    g->pt1 = ipcl::PlainText(g->v1);
    g->pt2 = ipcl::PlainText(g->v2);
    ipcl::setHybridRatio(0.5f);
    g->ct1 = g->key.pub_key.encrypt(g->pt1);
    g->ct2 = g->key.pub_key.encrypt(g->pt2);
    return g;
  }();
  return *f;
}

uint64_t low64(const ipcl::PlainText& pt, size_t i) {
  std::vector<uint32_t> v = pt.getElementVec(i);
  uint64_t r = v[0];
  if (v.size() > 1) r |= (uint64_t)v[1] << 32;
  return r;
}

// applies `op` element by element through size-1 texts (the single-element
// path of the reference, ippSBModExp) and collects the results
ipcl::CipherText elementwise(
    const std::function<ipcl::CipherText(const ipcl::CipherText&, size_t)>& op) {
  Fixture& f = fx();
  std::vector<BigNumber> out(kNumValues);
  for (int i = 0; i < kNumValues; i++) {
    ipcl::CipherText a(f.key.pub_key, f.ct1.getElement(i));
    out[i] = op(a, i).getElement(0);
  }
  return ipcl::CipherText(f.key.pub_key, out);
}

void expect_values(const ipcl::CipherText& ct,
                   const std::function<uint64_t(uint64_t, uint64_t)>& want) {
  Fixture& f = fx();
  ipcl::PlainText dt = f.key.priv_key.decrypt(ct);
  EXPECT_EQ(dt.getSize(), (size_t)kNumValues);
  for (int i = 0; i < kNumValues; i++)
    EXPECT_EQ(low64(dt, i), want(f.v1[i], f.v2[i]));
}

const auto kSum = [](uint64_t a, uint64_t b) { return a + b; };
const auto kProd = [](uint64_t a, uint64_t b) { return a * b; };

}  // namespace

TEST(OperationTest, CtPlusCtTest) {
  Fixture& f = fx();
  expect_values(elementwise([&](const ipcl::CipherText& a, size_t i) {
                  return a + ipcl::CipherText(f.key.pub_key, f.ct2.getElement(i));
                }), kSum);
}
TEST(OperationTest, CtPlusCtArrayTest) { expect_values(fx().ct1 + fx().ct2, kSum); }

TEST(OperationTest, CtPlusPtTest) {
  Fixture& f = fx();
  expect_values(elementwise([&](const ipcl::CipherText& a, size_t i) {
                  return a + ipcl::PlainText(f.pt2.getElement(i));
                }), kSum);
}
TEST(OperationTest, CtPlusPtArrayTest) { expect_values(fx().ct1 + fx().pt2, kSum); }

TEST(OperationTest, CtMultiplyPtTest) {
  Fixture& f = fx();
  expect_values(elementwise([&](const ipcl::CipherText& a, size_t i) {
                  return a * ipcl::PlainText(f.pt2.getElement(i));
                }), kProd);
}
TEST(OperationTest, CtMultiplyPtArrayTest) { expect_values(fx().ct1 * fx().pt2, kProd); }

// exponent zero: ct * 0 must decrypt to 0 (test_ops.cpp:328-367)
TEST(OperationTest, CtMultiplyZeroPtTest) {
  Fixture& f = fx();
  ipcl::PlainText zeros(std::vector<uint32_t>(kNumValues, 0u));
  ipcl::CipherText prod = elementwise([&](const ipcl::CipherText& a, size_t i) {
    return a * ipcl::PlainText(zeros.getElement(i));
  });
  ipcl::PlainText dt = f.key.priv_key.decrypt(prod);
  for (int i = 0; i < kNumValues; i++) EXPECT_EQ(low64(dt, i), (uint64_t)0);
  ipcl::PlainText dt2 = f.key.priv_key.decrypt(f.ct1 * zeros);
  for (int i = 0; i < kNumValues; i++) EXPECT_EQ(low64(dt2, i), (uint64_t)0);
}

// a + b*2 + b = a + 3b (test_ops.cpp:71-86,409-448)
TEST(OperationTest, AddSubTest) {
  Fixture& f = fx();
  expect_values(elementwise([&](const ipcl::CipherText& a0, size_t i) {
                  ipcl::CipherText b(f.key.pub_key, f.ct2.getElement(i));
                  ipcl::CipherText a = a0 + b * ipcl::PlainText(2u);
                  return a + b;
                }),
                [](uint64_t a, uint64_t b) { return a + 3 * b; });
}

TEST(OperationTest, PtPlusCtTest) {
  Fixture& f = fx();
  expect_values(elementwise([&](const ipcl::CipherText& a, size_t i) {
                  return ipcl::PlainText(f.pt2.getElement(i)) + a;
                }), kSum);
}
TEST(OperationTest, PtPlusCtArrayTest) { expect_values(fx().pt2 + fx().ct1, kSum); }

TEST(OperationTest, PtMultiplyCtTest) {
  Fixture& f = fx();
  expect_values(elementwise([&](const ipcl::CipherText& a, size_t i) {
                  return ipcl::PlainText(f.pt2.getElement(i)) * a;
                }), kProd);
}
TEST(OperationTest, PtMultiplyCtArrayTest) { expect_values(fx().pt2 * fx().ct1, kProd); }

// broadcasting of a size-1 right operand (ciphertext.cpp:37,51-59,96-99) and
// rotate / getCipherText / size checks
TEST(OperationTest, BroadcastRotateAndErrors) {
  Fixture& f = fx();
  ipcl::CipherText one = f.ct2.getCipherText(0);
  ipcl::PlainText ds = f.key.priv_key.decrypt(f.ct1 + one);
  for (int i = 0; i < kNumValues; i++)
    EXPECT_EQ(low64(ds, i), (uint64_t)f.v1[i] + f.v2[0]);
  ipcl::PlainText dm = f.key.priv_key.decrypt(f.ct1 * ipcl::PlainText(3u));
  for (int i = 0; i < kNumValues; i++) EXPECT_EQ(low64(dm, i), (uint64_t)f.v1[i] * 3);
  ipcl::PlainText dr = f.key.priv_key.decrypt(f.ct1.rotate(3));
  for (int i = 0; i < kNumValues; i++)
    EXPECT_EQ(low64(dr, (i + 3) % kNumValues), (uint64_t)f.v1[i]);
  ipcl::PlainText pr = f.pt1.rotate(-2);
  for (int i = 0; i < kNumValues; i++)
    EXPECT_EQ(low64(pr, i), (uint64_t)f.v1[(i + 2) % kNumValues]);
  ipcl::CipherText shorter(f.key.pub_key, f.ct2.getChunk(0, 5));
  EXPECT_THROW(f.ct1 + shorter);
  EXPECT_THROW(f.ct1 * ipcl::PlainText(std::vector<uint32_t>(3, 1u)));
  EXPECT_THROW(f.ct1.getCipherText(kNumValues));
  EXPECT_THROW(one.rotate(1));
}

// ipcl::modExp called directly, as benchmark/bench_hybrid.cpp:114 does
TEST(OperationTest, ModExpDirect) {
  BigNumber m = "0xf123456789abcdef0123456789abcdef0123456789abcdef0123456789abcdef1";
  std::vector<BigNumber> base = {BigNumber(2u), BigNumber(3u), m + 5, BigNumber(0u)};
  std::vector<BigNumber> exp = {BigNumber(10u), BigNumber(0u), BigNumber(2u), BigNumber(7u)};
  std::vector<BigNumber> mod(4, m);
  std::vector<BigNumber> r = ipcl::modExp(base, exp, mod);
  EXPECT_EQ(r[0], BigNumber(1024u));
  EXPECT_EQ(r[1], BigNumber(1u));
  EXPECT_EQ(r[2], BigNumber(25u));
  EXPECT_EQ(r[3], BigNumber(0u));
  EXPECT_EQ(ipcl::ippModExp(BigNumber(5u), BigNumber(3u), BigNumber(13u)), BigNumber(8u));
  // heterogeneous moduli in one batch (mod_exp.cpp:479-484)
  std::vector<BigNumber> mods = {BigNumber(13u), BigNumber(1000003u), m};
  std::vector<BigNumber> b3 = {BigNumber(5u), BigNumber(2u), BigNumber(2u)};
  std::vector<BigNumber> e3 = {BigNumber(3u), BigNumber(20u), BigNumber(100u)};
  std::vector<BigNumber> r3 = ipcl::modExp(b3, e3, mods);
  EXPECT_EQ(r3[0], BigNumber(8u));
  EXPECT_EQ(r3[1], BigNumber(1048576u % 1000003u));
  BigNumber two100 = "0x10000000000000000000000000";
  EXPECT_EQ(r3[2], two100);
}
