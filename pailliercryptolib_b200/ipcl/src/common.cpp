// common.cpp -- host randomness (reference: ipcl/utils/common.cpp:42-101 picks
// RDSEED/RDRAND/IPP-PRNG; here the OS entropy pool via getrandom(2)).  The RNG
// never sits on the measured path: benchmarks inject r through setRandom.
#include "ipcl/utils/common.hpp"

#include <sys/random.h>

#include <cstring>
#include <stdexcept>

namespace ipcl {

static void fill_random(void* p, std::size_t n) {
  unsigned char* b = static_cast<unsigned char*>(p);
  while (n) {
    ssize_t got = getrandom(b, n, 0);
    if (got < 0) throw std::runtime_error("getrandom failed");
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

}  // namespace ipcl
