// bignum.h -- ::BigNumber, the value type of every hot-path argument.
//
// Same public interface as the reference's BigNumber
// (/root/reference/ipcl/include/ipcl/bignum.h:28-158) so that code written
// against ipcl:: compiles unchanged, but written from scratch: the reference
// wraps an opaque IppsBigNumState* from IPP-Crypto; this one is a sign plus a
// little-endian vector of 32-bit limbs (ipclb200::hbn), which is also exactly
// the layout the C ABI (include/ipcl_b200.h) wants, so marshalling a batch is a
// memcpy per element.
//
// Not carried over: the IppsBigNumState* constructor / conversion operator /
// BN() friend (IPP plumbing with no meaning without IPP-Crypto) and the cereal
// save/load hooks (serialisation is SURVEY.md section 8f row 3).
#ifndef IPCL_B200_BIGNUM_H_
#define IPCL_B200_BIGNUM_H_

#include <cstdint>
#include <ostream>
#include <string>
#include <vector>

// the IPP typedefs leak through the reference's public API (bignum.h:30-34)
typedef uint32_t Ipp32u;
typedef int32_t Ipp32s;
typedef uint8_t Ipp8u;
typedef uint64_t Ipp64u;
typedef enum { IppsBigNumNEG = 0, IppsBigNumPOS = 1 } IppsBigNumSGN;

class BigNumber {
 public:
  BigNumber(Ipp32u value = 0);
  BigNumber(Ipp32s value);
  BigNumber(const Ipp32u* pData, int length = 1,
            IppsBigNumSGN sgn = IppsBigNumPOS);
  BigNumber(const BigNumber& bn) = default;
  BigNumber(BigNumber&& bn) = default;
  BigNumber(const char* s);
  virtual ~BigNumber() = default;

  // set value from a word array
  void Set(const Ipp32u* pData, int length = 1,
           IppsBigNumSGN sgn = IppsBigNumPOS);

  static const BigNumber& Zero();
  static const BigNumber& One();
  static const BigNumber& Two();

  BigNumber& operator=(const BigNumber& bn) = default;
  BigNumber& operator=(BigNumber&& bn) = default;
  BigNumber& operator+=(Ipp32u n);
  BigNumber& operator+=(const BigNumber& bn);
  BigNumber& operator-=(Ipp32u n);
  BigNumber& operator-=(const BigNumber& bn);
  BigNumber& operator*=(Ipp32u n);
  BigNumber& operator*=(const BigNumber& bn);
  BigNumber& operator/=(Ipp32u n);
  BigNumber& operator/=(const BigNumber& bn);
  BigNumber& operator%=(Ipp32u n);
  BigNumber& operator%=(const BigNumber& bn);
  friend BigNumber operator+(const BigNumber& a, const BigNumber& b);
  friend BigNumber operator+(const BigNumber& a, Ipp32u);
  friend BigNumber operator-(const BigNumber& a, const BigNumber& b);
  friend BigNumber operator-(const BigNumber& a, Ipp32u);
  friend BigNumber operator*(const BigNumber& a, const BigNumber& b);
  friend BigNumber operator*(const BigNumber& a, Ipp32u);
  friend BigNumber operator%(const BigNumber& a, const BigNumber& b);
  friend BigNumber operator%(const BigNumber& a, Ipp32u);
  friend BigNumber operator/(const BigNumber& a, const BigNumber& b);
  friend BigNumber operator/(const BigNumber& a, Ipp32u);

  // modulo arithmetic; *this is the modulus
  BigNumber Modulo(const BigNumber& a) const;
  BigNumber ModAdd(const BigNumber& a, const BigNumber& b) const;
  BigNumber ModSub(const BigNumber& a, const BigNumber& b) const;
  BigNumber ModMul(const BigNumber& a, const BigNumber& b) const;
  BigNumber InverseAdd(const BigNumber& a) const;
  BigNumber InverseMul(const BigNumber& a) const;
  BigNumber gcd(const BigNumber& q) const;
  int compare(const BigNumber&) const;

  friend bool operator<(const BigNumber& a, const BigNumber& b);
  friend bool operator>(const BigNumber& a, const BigNumber& b);
  friend bool operator==(const BigNumber& a, const BigNumber& b);
  friend bool operator!=(const BigNumber& a, const BigNumber& b);
  friend bool operator<=(const BigNumber& a, const BigNumber& b) {
    return !(a > b);
  }
  friend bool operator>=(const BigNumber& a, const BigNumber& b) {
    return !(a < b);
  }

  bool IsOdd() const;
  bool IsEven() const { return !IsOdd(); }
  bool TestBit(int index) const;

  int MSB() const;
  int LSB() const;
  int BitSize() const { return MSB() + 1; }
  int DwordSize() const { return (BitSize() + 31) >> 5; }
  friend int Bit(const std::vector<Ipp32u>& v, int n);

  void num2hex(std::string& s) const;          // "0x..." lower case
  void num2vec(std::vector<Ipp32u>& v) const;  // appends the 32-bit words
  friend std::ostream& operator<<(std::ostream& os, const BigNumber& a);
  void num2char(std::vector<Ipp8u>& dest) const;

  // big-endian octet strings (the QAT wire format, bignum.cpp:511-565)
  static bool fromBin(BigNumber& bn, const unsigned char* data, int len);
  static bool toBin(unsigned char* data, int len, const BigNumber& bn);
  static bool toBin(unsigned char** data, int* len, const BigNumber& bn);

  // ---- additions for the flat-buffer boundary ------------------------------
  // magnitude words, little endian, no leading zeros (empty for zero)
  const std::vector<Ipp32u>& words() const { return m_mag; }
  bool isNegative() const { return m_neg; }
  // copy |value| into `out[0..n)`, zero padded; false if it does not fit
  bool toWords(Ipp32u* out, std::size_t n) const;

 private:
  void normalize();
  bool m_neg = false;
  std::vector<Ipp32u> m_mag;
};

constexpr int BITSIZE_WORD(int n) { return (((n) + 31) >> 5); }
constexpr int BITSIZE_DWORD(int n) { return (((n) + 63) >> 6); }

#endif  // IPCL_B200_BIGNUM_H_
