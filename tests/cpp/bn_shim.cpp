// bn_shim.cpp -- extern "C" window onto ::BigNumber for the host-only pytest
// (tests/test_bignum_host.py): operands and results travel as strings so the
// test can compare every operation against Python integers.
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "ipcl/bignum.h"

static int put(const std::string& s, char* out, int cap) {
  if ((int)s.size() + 1 > cap) return -2;
  memcpy(out, s.c_str(), s.size() + 1);
  return 0;
}

extern "C" int bn_op(const char* op, const char* a_s, const char* b_s,
                     const char* c_s, char* out, int cap) {
  try {
    BigNumber a(a_s), b(b_s), c(c_s);
    std::string o(op), s;
    BigNumber r;
    if (o == "add") r = a + b;
    else if (o == "sub") r = a - b;
    else if (o == "mul") r = a * b;
    else if (o == "div") r = a / b;
    else if (o == "mod") r = a % b;
    else if (o == "modmul") r = c.ModMul(a, b);
    else if (o == "modadd") r = c.ModAdd(a, b);
    else if (o == "modsub") r = c.ModSub(a, b);
    else if (o == "invmul") r = b.InverseMul(a);
    else if (o == "invadd") r = b.InverseAdd(a);
    else if (o == "gcd") r = a.gcd(b);
    else if (o == "addw") r = a + (Ipp32u)std::stoul(b_s);
    else if (o == "mulw") r = a * (Ipp32u)std::stoul(b_s);
    else if (o == "cmp") return put(std::to_string(a.compare(b)), out, cap);
    else if (o == "bits")
      return put(std::to_string(a.BitSize()) + "," + std::to_string(a.MSB()) + "," +
                     std::to_string(a.LSB()) + "," + std::to_string(a.DwordSize()) +
                     "," + std::to_string((int)a.IsOdd()) + "," +
                     std::to_string((int)a.TestBit(std::stoi(b_s))), out, cap);
    else if (o == "vec") {
      std::vector<Ipp32u> v;
      a.num2vec(v);
      std::ostringstream os;
      for (size_t i = 0; i < v.size(); i++) os << (i ? "," : "") << v[i];
      return put(os.str(), out, cap);
    } else if (o == "bin") {  // toBin -> fromBin round trip, returns hex of both
      int len = std::stoi(b_s);
      std::vector<unsigned char> buf((size_t)len, 0);
      BigNumber::toBin(buf.data(), len, a);
      BigNumber back;
      BigNumber::fromBin(back, buf.data(), len);
      std::ostringstream os;
      for (unsigned char ch : buf) {
        static const char* d = "0123456789abcdef";
        os << d[ch >> 4] << d[ch & 15];
      }
      std::string hs;
      back.num2hex(hs);
      return put(os.str() + "," + hs, out, cap);
    } else if (o == "dec") r = a;  // parse (decimal or hex) and print
    else return -3;
    r.num2hex(s);
    return put(s, out, cap);
  } catch (const std::exception& e) {
    put(std::string("EXC:") + e.what(), out, cap);
    return -1;
  }
}
