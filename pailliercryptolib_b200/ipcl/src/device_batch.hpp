// device_batch.hpp -- a batch of big integers resident in HBM: count x words
// little-endian 32-bit limbs, element-major, the layout every *_dev entry
// point of include/ipcl_b200.h reads and writes.  Owned through shared_ptr by
// the texts that refer to it; immutable once its producer kernel is enqueued.
// Allocation, copies, kernels and the free are all ordered on the library
// stream (ipclb200_stream()), so no host synchronisation is needed until a
// caller asks for the values (toHost()).
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
  void* d = nullptr;
  std::size_t count = 0;
  int words = 0;

  DeviceBatch(std::size_t count_, int words_) : count(count_), words(words_) {
    DEVICE_CHECK(ipclb200_dev_alloc(bytes(), &d));
  }
  ~DeviceBatch() {
    if (d) ipclb200_dev_free(d);
  }
  DeviceBatch(const DeviceBatch&) = delete;
  DeviceBatch& operator=(const DeviceBatch&) = delete;

  std::size_t bytes() const {
    return count * static_cast<std::size_t>(words) * sizeof(uint32_t);
  }
  uint32_t* ptr() const { return static_cast<uint32_t*>(d); }

  // every element must be non-negative and at most `words` words wide
  static bool fits(const std::vector<BigNumber>& v, int words) {
    for (const auto& x : v)
      if (x.isNegative() || static_cast<int>(x.words().size()) > words)
        return false;
    return true;
  }

  static std::shared_ptr<DeviceBatch> fromHost(const std::vector<BigNumber>& v,
                                               int words) {
    auto b = std::make_shared<DeviceBatch>(v.size(), words);
    std::vector<uint32_t> flat;
    pack(v, words, flat);
    if (!flat.empty()) DEVICE_CHECK(ipclb200_dev_upload(b->d, flat.data(), b->bytes()));
    return b;
  }

  std::vector<BigNumber> toHost() const {
    std::vector<uint32_t> flat(count * static_cast<std::size_t>(words));
    if (!flat.empty()) DEVICE_CHECK(ipclb200_dev_download(flat.data(), d, bytes()));
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
