// keygen.cpp -- ipcl::generateKeypair / getPrimeBN (reference:
// ipcl/keygen.cpp:13-117, which loops over ippsPrimeGen_BN).
//
// Prime search here is batched: a block of random odd candidates is sieved by
// small primes on the host, the survivors take a base-2 Fermat test as ONE
// heterogeneous-modulus modexp batch on the GPU, and the first survivor is
// confirmed by Miller-Rabin rounds whose modexps are again one GPU batch.
#include <algorithm>
#include <vector>

#include "ipcl/ipcl.hpp"
#include "ipcl/utils/util.hpp"

namespace ipcl {

constexpr int N_BIT_SIZE_MAX = 4096;
constexpr int N_BIT_SIZE_MIN = 200;
constexpr int kMillerRabinRounds = 10;  // nTrials of ippsPrimeGen_BN (:33)

static const std::vector<Ipp32u>& smallPrimes() {
  static const std::vector<Ipp32u> primes = [] {
    std::vector<Ipp32u> p;
    std::vector<bool> comp(4096, false);
    for (Ipp32u i = 2; i < 4096; i++) {
      if (comp[i]) continue;
      if (i > 2) p.push_back(i);
      for (Ipp32u j = i * i; j < 4096; j += i) comp[j] = true;
    }
    return p;
  }();
  return primes;
}

static bool passesSieve(const BigNumber& c) {
  for (Ipp32u p : smallPrimes())
    if ((c % p) == BigNumber::Zero()) return false;
  return true;
}

// Miller-Rabin with `rounds` random bases; the a^d mod c of all rounds is one
// modexp batch
static bool millerRabin(const BigNumber& c, int rounds) {
  const BigNumber cm1 = c - 1;
  const int s = cm1.LSB();
  BigNumber d = cm1;
  for (int i = 0; i < s; i++) d /= 2u;
  std::vector<BigNumber> base(static_cast<std::size_t>(rounds)),
      e(base.size(), d), m(base.size(), c);
  for (auto& a : base) a = getRandomBN(c.BitSize() + 64) % (c - 3) + 2;
  std::vector<BigNumber> x = modExp(base, e, m);
  for (auto& xi : x) {
    if (xi == BigNumber::One() || xi == cm1) continue;
    bool witness = true;
    for (int r = 1; r < s && witness; r++) {
      xi = c.ModMul(xi, xi);
      if (xi == cm1) witness = false;
    }
    if (witness) return false;
  }
  return true;
}

BigNumber getPrimeBN(int max_bits) {
  ERROR_CHECK(max_bits >= 16, "getPrimeBN: need at least 16 bits");
  const std::size_t block = 192;
  for (;;) {
    std::vector<BigNumber> cand;
    while (cand.size() < block) {
      BigNumber c = getRandomBN(max_bits);
      // exact bit length and odd
      std::vector<Ipp32u> w;
      c.num2vec(w);
      w.resize(static_cast<std::size_t>((max_bits + 31) / 32), 0u);
      w[0] |= 1u;
      w[static_cast<std::size_t>((max_bits - 1) / 32)] |= 1u << ((max_bits - 1) % 32);
      c = BigNumber(w.data(), static_cast<int>(w.size()));
      if (passesSieve(c)) cand.push_back(c);
    }
    // Fermat base 2 on the whole block: 2^(c-1) mod c
    std::vector<BigNumber> two(cand.size(), BigNumber::Two()), e(cand.size());
    for (std::size_t i = 0; i < cand.size(); i++) e[i] = cand[i] - 1;
    std::vector<BigNumber> f = modExp(two, e, cand);
    for (std::size_t i = 0; i < cand.size(); i++)
      if (f[i] == BigNumber::One() && millerRabin(cand[i], kMillerRabinRounds))
        return cand[i];
  }
}

// 2^(key_size/2 - 100): |p - q| must exceed it (keygen.cpp:43-59)
static BigNumber getPrimeDistance(int64_t key_size) {
  const uint64_t count = static_cast<uint64_t>(key_size / 2 - 100);
  std::vector<Ipp32u> tmp(count / 32 + 1, 0u);
  tmp[count / 32] = 1u << (count & 0x1F);
  return BigNumber(tmp.data(), static_cast<int>(tmp.size()));
}

static bool isClosePrimeBN(const BigNumber& p, const BigNumber& q,
                           const BigNumber& ref_dist) {
  BigNumber real_dist = (p >= q) ? (p - q) : (q - p);
  return !(real_dist > ref_dist);
}

KeyPair generateKeypair(int64_t n_length, bool enable_DJN) {
  ERROR_CHECK(n_length <= N_BIT_SIZE_MAX,
              "generateKeyPair: modulus size in bits should belong to either "
              "1Kb, 2Kb, 3Kb or 4Kb range only, key size exceed the range!!!");
  ERROR_CHECK((n_length >= N_BIT_SIZE_MIN) && (n_length % 4 == 0),
              "generateKeyPair: key size should >=200, and divisible by 4");
  const BigNumber ref_dist = getPrimeDistance(n_length);
  const int half = static_cast<int>(n_length / 2);
  BigNumber p, q, n;
  for (;;) {
    if (enable_DJN) {
      // p = q = 3 (mod 4) and gcd(p-1, q-1) = 2 (keygen.cpp:73-90)
      do {
        p = getPrimeBN(half);
      } while (!p.TestBit(1));
      do {
        q = getPrimeBN(half);
      } while (q == p || !q.TestBit(1));
      if ((p - 1).gcd(q - 1) != BigNumber::Two()) continue;
    } else {
      p = getPrimeBN(half);
      do {
        q = getPrimeBN(half);
      } while (q == p);
    }
    n = p * q;
    if (n.BitSize() != n_length || isClosePrimeBN(p, q, ref_dist)) continue;
    break;
  }
  PublicKey pk(n, static_cast<int>(n_length), enable_DJN);
  PrivateKey sk(pk, p, q);
  return KeyPair{pk, sk};
}

}  // namespace ipcl
