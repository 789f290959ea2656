// marshal.hpp -- vector<BigNumber> <-> flat little-endian limb buffers, the
// data layout of the C ABI (include/ipcl_b200.h).  One pass, one memcpy per
// element, OpenMP over the elements; replaces the per-chunk sub-vector copies
// and 8 x mod_dwords padding buffers of ippMBModExp
// (ipcl/mod_exp.cpp:490-506,627-632).  The flat side is a per-thread PINNED
// staging buffer, so the copies to and from the GPUs are true DMA transfers
// that overlap across devices.
#ifndef IPCL_B200_SRC_MARSHAL_HPP_
#define IPCL_B200_SRC_MARSHAL_HPP_

#include <cstdint>
#include <cstring>
#include <vector>

#include "ipcl/bignum.h"

namespace ipcl {
namespace detail {

// Page-locked host memory comes from a process-wide pool of slabs that is never
// torn down (at process exit the CUDA runtime may already be gone): acquire a
// slab of at least `words` words, hand it back when done.  Falls back to
// pageable memory if pinning fails.
struct HostSlab {
  uint32_t* p = nullptr;
  std::size_t words = 0;
  bool pinned = false;
};
HostSlab acquireHostSlab(std::size_t words);
void returnHostSlab(HostSlab& s);

// a slab for the duration of a scope (flat operands / results of one call)
struct ScopedSlab {
  HostSlab s;
  explicit ScopedSlab(std::size_t words) : s(acquireHostSlab(words)) {}
  ~ScopedSlab() { returnHostSlab(s); }
  ScopedSlab(const ScopedSlab&) = delete;
  ScopedSlab& operator=(const ScopedSlab&) = delete;
  uint32_t* data() const { return s.p; }
};

// grow-only staging buffer of the calling thread (pack / unpack go through it)
class PinnedBuffer {
 public:
  static PinnedBuffer& forThread();
  uint32_t* get(std::size_t words);
  ~PinnedBuffer();

 private:
  HostSlab m_slab;
};

// the flat host image of a device batch (count x words limbs): what a text
// keeps when a caller reads single elements -- no BigNumber is built for the
// elements nobody looks at
struct FlatImage {
  HostSlab slab;
  std::size_t count = 0;
  int words = 0;
  FlatImage(std::size_t count_, int words_)
      : slab(acquireHostSlab(count_ * static_cast<std::size_t>(words_))),
        count(count_),
        words(words_) {}
  ~FlatImage() { returnHostSlab(slab); }
  FlatImage(const FlatImage&) = delete;
  FlatImage& operator=(const FlatImage&) = delete;
  const uint32_t* element(std::size_t i) const {
    return slab.p + i * static_cast<std::size_t>(words);
  }
};

inline int maxWords(const std::vector<BigNumber>& v) {
  std::size_t w = 1;
  for (const auto& x : v)
    if (x.words().size() > w) w = x.words().size();
  return static_cast<int>(w);
}

// out: v.size() x words, zero padded.  Every element must be non-negative and
// fit (callers reduce first).
inline void pack(const std::vector<BigNumber>& v, int words, uint32_t* out) {
  const std::size_t W = static_cast<std::size_t>(words);
#pragma omp parallel for schedule(static) if (v.size() >= 1024)
  for (std::size_t i = 0; i < v.size(); i++) {
    const auto& w = v[i].words();
    uint32_t* dst = out + i * W;
    if (!w.empty()) std::memcpy(dst, w.data(), w.size() * sizeof(uint32_t));
    if (w.size() < W) std::memset(dst + w.size(), 0, (W - w.size()) * sizeof(uint32_t));
  }
}

inline void pack(const std::vector<BigNumber>& v, int words,
                 std::vector<uint32_t>& out) {
  out.resize(v.size() * static_cast<std::size_t>(words));
  pack(v, words, out.data());
}

inline std::vector<BigNumber> unpack(const uint32_t* flat, std::size_t count,
                                     int words) {
  std::vector<BigNumber> r(count);
#pragma omp parallel for schedule(static) if (count >= 1024)
  for (std::size_t i = 0; i < count; i++)
    r[i].Set(flat + i * static_cast<std::size_t>(words), words, IppsBigNumPOS);
  return r;
}

inline std::vector<BigNumber> unpack(const std::vector<uint32_t>& flat,
                                     std::size_t count, int words) {
  return unpack(flat.data(), count, words);
}

inline bool allEqual(const std::vector<BigNumber>& v) {
  for (std::size_t i = 1; i < v.size(); i++)
    if (v[i].words() != v[0].words() || v[i].isNegative() != v[0].isNegative())
      return false;
  return true;
}

}  // namespace detail
}  // namespace ipcl
#endif  // IPCL_B200_SRC_MARSHAL_HPP_
