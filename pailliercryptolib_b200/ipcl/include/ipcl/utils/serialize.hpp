// serialize.hpp -- ipcl::serializer.
//
// Same entry points as the reference (ipcl/include/ipcl/utils/serialize.hpp:
// 25-63): serialize / deserialize on streams, serializeToFile /
// deserializeFromFile, for BigNumber, PublicKey, PrivateKey, PlainText and
// CipherText.  The reference delegates to cereal's PortableBinary archives;
// cereal is not available in this image, so the byte layout that cereal 1.3.2
// produces for the reference's save/load hooks is written by hand:
//
//   archive   := u8 little_endian_flag(=1) object
//   versioned := the first occurrence of a class type in an archive is
//                preceded by its u32 class version (0); later ones are not
//   BigNumber := [u32 ver] u64 nwords, nwords x u32 (LE limbs), i32 sign(0 neg,1 pos)
//                                            (bignum.h:133-150)
//   PublicKey := [u32 ver] i32 bits, u8 enable_DJN, i32 randbits, BigNumber n,
//                BigNumber hs                 (pub_key.hpp:134-164)
//   PrivateKey:= [u32 ver] i32 bits_of_p, BigNumber p, BigNumber q
//                                            (pri_key.hpp:94-133)
//   BaseText  := [u32 ver] u64 size, u64 count, count x BigNumber
//                                            (base_text.hpp:110-114)
//   PlainText := [u32 ver] BaseText           (plaintext.hpp:92-97)
//   CipherText:= [u32 ver] BaseText PublicKey (ciphertext.hpp:71-75)
//
// The layout is restated from cereal's documented behaviour and has not been
// diffed against bytes produced by the real library (not buildable here).
#ifndef IPCL_B200_UTILS_SERIALIZE_HPP_
#define IPCL_B200_UTILS_SERIALIZE_HPP_

#include <fstream>
#include <istream>
#include <ostream>
#include <string>

#include "ipcl/bignum.h"

namespace ipcl {

class PublicKey;
class PrivateKey;
class PlainText;
class CipherText;

namespace serializer {

void serialize(std::ostream& ss, const BigNumber& obj);
void serialize(std::ostream& ss, const PublicKey& obj);
void serialize(std::ostream& ss, const PrivateKey& obj);
void serialize(std::ostream& ss, const PlainText& obj);
void serialize(std::ostream& ss, const CipherText& obj);

void deserialize(std::istream& ss, BigNumber& obj);
void deserialize(std::istream& ss, PublicKey& obj);
void deserialize(std::istream& ss, PrivateKey& obj);
void deserialize(std::istream& ss, PlainText& obj);
void deserialize(std::istream& ss, CipherText& obj);

template <typename T>
bool serializeToFile(const std::string& fn, const T& obj) {
  std::ofstream ofs(fn, std::ios::out | std::ios::binary);
  if (!ofs.is_open()) return false;
  serialize(ofs, obj);
  return true;
}

template <typename T>
bool deserializeFromFile(const std::string& fn, T& obj) {
  std::ifstream ifs(fn, std::ios::in | std::ios::binary);
  if (!ifs.is_open()) return false;
  deserialize(ifs, obj);
  return true;
}

}  // namespace serializer
}  // namespace ipcl
#endif  // IPCL_B200_UTILS_SERIALIZE_HPP_
