// drbg.hpp -- seed of the device-side generator of the DJN randoms.
//
// Reference: PublicKey::getDJNObfuscator draws r = getRandomBN(m_randbits) per
// element on the host (ipcl/pub_key.cpp:59-61; ipcl/utils/common.cpp:42-101
// picks RDSEED / RDRAND / ippsPRNGen).  At batch 65536 that is 8 MB from the
// entropy source per encrypt call -- about 20-30 ms of getrandom(2), twice the
// encrypt kernel.  Here every encrypt call takes ONE fresh 256-bit key and a
// 96-bit nonce from the OS and the GPU expands them to the r of the batch with
// the ChaCha20 block function (ipclb200_batch_random / ipclb200_encrypt_drbg,
// include/ipcl_b200.h); nothing is kept between calls, so the generator is
// fork- and thread-safe by construction.  IPCL_B200_DEVICE_RANDOM=0 restores
// host-drawn randoms; setRandom() (injected r) never takes this path.
#ifndef IPCL_B200_SRC_DRBG_HPP_
#define IPCL_B200_SRC_DRBG_HPP_

#include <cstdint>

namespace ipcl {
namespace detail {

void freshDrbgSeed(uint32_t (&key)[8], uint32_t (&nonce)[3]);
bool deviceRandomEnabled();

}  // namespace detail
}  // namespace ipcl
#endif  // IPCL_B200_SRC_DRBG_HPP_
