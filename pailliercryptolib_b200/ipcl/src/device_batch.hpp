// device_batch.hpp -- a batch of big integers resident in HBM: count x words
// little-endian 32-bit limbs, element-major, sharded in contiguous blocks over
// the GPUs the library was initialised on (ipclb200_batch_*,
// include/ipcl_b200.h).  Owned through shared_ptr by the texts that refer to
// it; immutable once its producer kernel is enqueued.  Allocation, copies,
// kernels and the free are ordered on the library stream of each shard's
// device, so no host synchronisation is needed until a caller asks for the
// values (toHost()).
//
// This is the device-resident CipherText of SURVEY.md section 8f row 2: the
// reference copies vector<BigNumber> at every step (ipcl/base_text.cpp:102,
// ipcl/mod_exp.cpp:627-632).
#ifndef IPCL_B200_SRC_DEVICE_BATCH_HPP_
#define IPCL_B200_SRC_DEVICE_BATCH_HPP_

#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

#include "ipcl/bignum.h"
#include "ipcl/utils/util.hpp"
#include "ipcl_b200.h"
#include "marshal.hpp"

namespace ipcl {
namespace detail {

struct DeviceBatch {
  ipclb200_batch* h = nullptr;
  std::size_t count = 0;
  int words = 0;

  DeviceBatch(std::size_t count_, int words_) : count(count_), words(words_) {
    DEVICE_CHECK(ipclb200_batch_alloc(count, words, &h));
  }
  ~DeviceBatch() {
    if (h) ipclb200_batch_free(h);
  }
  DeviceBatch(const DeviceBatch&) = delete;
  DeviceBatch& operator=(const DeviceBatch&) = delete;

  // every element must be non-negative and at most `words` words wide
  static bool fits(const std::vector<BigNumber>& v, int words) {
    for (const auto& x : v)
      if (x.isNegative() || static_cast<int>(x.words().size()) > words)
        return false;
    return true;
  }

  void upload(const uint32_t* flat, int flat_words) {
    if (count) DEVICE_CHECK(ipclb200_batch_upload(h, flat, flat_words));
  }

  static std::shared_ptr<DeviceBatch> fromHost(const std::vector<BigNumber>& v,
                                               int words) {
    auto b = std::make_shared<DeviceBatch>(v.size(), words);
    PinnedBuffer& stage = PinnedBuffer::forThread();
    uint32_t* flat = stage.get(v.size() * static_cast<std::size_t>(words));
    pack(v, words, flat);
    b->upload(flat, words);
    return b;
  }

  void downloadFlat(uint32_t* dst) const {
    if (count) DEVICE_CHECK(ipclb200_batch_download(h, dst, words));
  }

  std::vector<BigNumber> toHost() const {
    PinnedBuffer& stage = PinnedBuffer::forThread();
    uint32_t* flat = stage.get(count * static_cast<std::size_t>(words));
    if (count) DEVICE_CHECK(ipclb200_batch_download(h, flat, words));
    return unpack(flat, count, words);
  }
};

// the device-resident path needs limb strides that are kernel size classes;
// IPCL_B200_DEVICE_RESIDENT=0 switches it off (every call then marshals
// through the host-pointer entry points, as before)
bool deviceResidentEnabled();
inline bool isClassWords(int words) {
  return words > 0 && ipclb200_class_words(words) == words;
}

}  // namespace detail
}  // namespace ipcl
#endif  // IPCL_B200_SRC_DEVICE_BATCH_HPP_
