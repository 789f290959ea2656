// text_util.hpp -- helpers shared by PlainText and CipherText.
#ifndef IPCL_B200_SRC_TEXT_UTIL_HPP_
#define IPCL_B200_SRC_TEXT_UTIL_HPP_

#include <vector>

#include "ipcl/bignum.h"

namespace ipcl {
namespace detail {

// rotate(shift) of PlainText / CipherText (reference: ipcl/plaintext.cpp:62-79,
// ipcl/ciphertext.cpp:117-133): element i moves to i + shift (mod size);
// throws for a single-element text or |shift| > size
std::vector<BigNumber> rotated(const std::vector<BigNumber>& v, int shift);

}  // namespace detail
}  // namespace ipcl
#endif  // IPCL_B200_SRC_TEXT_UTIL_HPP_
