// common.cpp -- host randomness (reference: ipcl/utils/common.cpp:42-101 picks
// RDSEED/RDRAND/IPP-PRNG; here the OS entropy pool via getrandom(2)).  The RNG
// never sits on the measured path: benchmarks inject r through setRandom.
#include "ipcl/utils/common.hpp"

#include <sys/random.h>

#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "drbg.hpp"

namespace ipcl {

static void fill_random(void* p, std::size_t n) {
  unsigned char* b = static_cast<unsigned char*>(p);
  while (n) {
    // the kernel hands out at most 32 MB - 1 per call and may be interrupted by a
    // signal on requests over 256 bytes
    ssize_t got = getrandom(b, n < (1u << 24) ? n : (1u << 24), 0);
    if (got < 0) {
      if (errno == EINTR) continue;
      throw std::runtime_error("getrandom failed");
    }
    b += got;
    n -= static_cast<std::size_t>(got);
  }
}

void rand32u(std::vector<Ipp32u>& addr) {
  if (!addr.empty()) fill_random(addr.data(), addr.size() * sizeof(Ipp32u));
}

BigNumber getRandomBN(int bits) {
  if (bits <= 0) return BigNumber::Zero();
  std::vector<Ipp32u> w(static_cast<std::size_t>((bits + 31) / 32));
  rand32u(w);
  if (bits % 32) w.back() &= (1u << (bits % 32)) - 1u;
  return BigNumber(w.data(), static_cast<int>(w.size()));
}

namespace detail {

void freshDrbgSeed(uint32_t (&key)[8], uint32_t (&nonce)[3]) {
  uint32_t buf[11];
  fill_random(buf, sizeof(buf));
  std::memcpy(key, buf, sizeof(key));
  std::memcpy(nonce, buf + 8, sizeof(nonce));
}

bool deviceRandomEnabled() {
  static const bool on = [] {
    const char* e = std::getenv("IPCL_B200_DEVICE_RANDOM");
    return !(e && e[0] == '0');
  }();
  return on;
}

}  // namespace detail

}  // namespace ipcl
