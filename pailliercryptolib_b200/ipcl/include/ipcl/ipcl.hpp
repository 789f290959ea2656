// ipcl.hpp -- umbrella header (ipcl/include/ipcl/ipcl.hpp:7-37).
#ifndef IPCL_B200_IPCL_HPP_
#define IPCL_B200_IPCL_HPP_

#include "ipcl/mod_exp.hpp"
#include "ipcl/pri_key.hpp"
#include "ipcl/utils/context.hpp"
#include "ipcl/utils/serialize.hpp"

namespace ipcl {

struct KeyPair {
  PublicKey pub_key;
  PrivateKey priv_key;
};

// random probable prime of exactly maxBitSize bits (keygen.cpp:13-41)
BigNumber getPrimeBN(int maxBitSize);

// n_length: key size in bits, multiple of 4, 200 <= n_length <= 2048 in the
// reference (keygen.cpp:10-11,97-102); this build also accepts up to 4096
// because the kernels cover 8192-bit n^2.
KeyPair generateKeypair(int64_t n_length, bool enable_DJN = true);

}  // namespace ipcl
#endif  // IPCL_B200_IPCL_HPP_
