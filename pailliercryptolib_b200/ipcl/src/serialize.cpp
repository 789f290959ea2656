// serialize.cpp -- hand-written cereal-PortableBinary-compatible layout, see
// ipcl/utils/serialize.hpp for the grammar and the reference hooks it mirrors.
#include "ipcl/utils/serialize.hpp"

#include <cstring>
#include <set>
#include <vector>

#include "ipcl/ciphertext.hpp"
#include "ipcl/pri_key.hpp"
#include "ipcl/utils/util.hpp"

namespace ipcl {
namespace serializer {
namespace {

enum ClassId { kBigNumber, kPublicKey, kPrivateKey, kBaseText, kPlainText, kCipherText };

class Writer {
 public:
  explicit Writer(std::ostream& os) : m_os(os) { put<uint8_t>(1); }
  template <typename T>
  void put(T v) {
    unsigned char b[sizeof(T)];
    for (size_t i = 0; i < sizeof(T); i++)
      b[i] = static_cast<unsigned char>(static_cast<uint64_t>(v) >> (8 * i));
    m_os.write(reinterpret_cast<const char*>(b), sizeof(T));
  }
  // cereal emits a class version the first time a versioned type is seen
  void version(ClassId id) {
    if (m_seen.insert(id).second) put<uint32_t>(0);
  }
  void bn(const BigNumber& x) {
    version(kBigNumber);
    std::vector<Ipp32u> v;
    x.num2vec(v);
    put<uint64_t>(v.size());
    for (Ipp32u w : v) put<uint32_t>(w);
    put<int32_t>(x.isNegative() ? IppsBigNumNEG : IppsBigNumPOS);
  }
  void pub(const PublicKey& pk) {
    version(kPublicKey);
    put<int32_t>(pk.getBits());
    put<uint8_t>(pk.isDJN() ? 1 : 0);
    put<int32_t>(pk.isDJN() ? pk.getRandBits() : 0);
    bn(*pk.getN());
    bn(pk.getHS());
  }
  void base(const BaseText& t) {
    version(kBaseText);
    put<uint64_t>(t.getSize());
    put<uint64_t>(t.texts().size());
    for (const auto& x : t.texts()) bn(x);
  }

 private:
  std::ostream& m_os;
  std::set<int> m_seen;
};

class Reader {
 public:
  explicit Reader(std::istream& is) : m_is(is) {
    uint8_t little = get<uint8_t>();
    ERROR_CHECK(little == 1, "deserialize: big-endian archives are not supported");
  }
  template <typename T>
  T get() {
    unsigned char b[sizeof(T)];
    m_is.read(reinterpret_cast<char*>(b), sizeof(T));
    ERROR_CHECK(static_cast<size_t>(m_is.gcount()) == sizeof(T),
                "deserialize: truncated archive");
    uint64_t v = 0;
    for (size_t i = 0; i < sizeof(T); i++) v |= static_cast<uint64_t>(b[i]) << (8 * i);
    return static_cast<T>(v);
  }
  void version(ClassId id) {
    if (m_seen.insert(id).second) {
      uint32_t v = get<uint32_t>();
      ERROR_CHECK(v == 0, "deserialize: unknown class version");
    }
  }
  BigNumber bn() {
    version(kBigNumber);
    uint64_t n = get<uint64_t>();
    ERROR_CHECK(n <= (1u << 20), "deserialize: BigNumber too large");
    std::vector<Ipp32u> v(static_cast<size_t>(n));
    for (auto& w : v) w = get<uint32_t>();
    int32_t sgn = get<int32_t>();
    return BigNumber(v.data(), static_cast<int>(v.size()),
                     sgn == IppsBigNumNEG ? IppsBigNumNEG : IppsBigNumPOS);
  }
  void pub(PublicKey& pk) {
    version(kPublicKey);
    int32_t bits = get<int32_t>();
    bool djn = get<uint8_t>() != 0;
    int32_t randbits = get<int32_t>();
    BigNumber n = bn();
    BigNumber hs = bn();
    if (djn)
      pk.create(n, bits, hs, randbits);  // pub_key.hpp:158-163
    else
      pk.create(n, bits);
  }
  std::vector<BigNumber> base() {
    version(kBaseText);
    uint64_t size = get<uint64_t>();
    uint64_t count = get<uint64_t>();
    ERROR_CHECK(size == count && count <= (1u << 28),
                "deserialize: inconsistent text container");
    std::vector<BigNumber> v(static_cast<size_t>(count));
    for (auto& x : v) x = bn();
    return v;
  }

 private:
  std::istream& m_is;
  std::set<int> m_seen;
};

}  // namespace

void serialize(std::ostream& ss, const BigNumber& obj) { Writer(ss).bn(obj); }

void serialize(std::ostream& ss, const PublicKey& obj) { Writer(ss).pub(obj); }

void serialize(std::ostream& ss, const PrivateKey& obj) {
  Writer w(ss);
  w.version(kPrivateKey);
  w.put<int32_t>(obj.getP()->BitSize());
  w.bn(*obj.getP());
  w.bn(*obj.getQ());
}

void serialize(std::ostream& ss, const PlainText& obj) {
  Writer w(ss);
  w.version(kPlainText);
  w.base(obj);
}

void serialize(std::ostream& ss, const CipherText& obj) {
  Writer w(ss);
  w.version(kCipherText);
  w.base(obj);
  w.pub(*obj.getPubKey());
}

void deserialize(std::istream& ss, BigNumber& obj) { obj = Reader(ss).bn(); }

void deserialize(std::istream& ss, PublicKey& obj) { Reader(ss).pub(obj); }

void deserialize(std::istream& ss, PrivateKey& obj) {
  Reader r(ss);
  r.version(kPrivateKey);
  (void)r.get<int32_t>();
  BigNumber p = r.bn();
  BigNumber q = r.bn();
  obj = PrivateKey(p * q, p, q);  // rebuilds every derived constant (:108-131)
}

void deserialize(std::istream& ss, PlainText& obj) {
  Reader r(ss);
  r.version(kPlainText);
  obj = PlainText(r.base());
}

void deserialize(std::istream& ss, CipherText& obj) {
  Reader r(ss);
  r.version(kCipherText);
  std::vector<BigNumber> texts = r.base();
  PublicKey pk;
  r.pub(pk);
  obj = CipherText(pk, texts);
}

}  // namespace serializer
}  // namespace ipcl
