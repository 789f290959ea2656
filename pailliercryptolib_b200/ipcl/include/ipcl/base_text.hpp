// base_text.hpp -- container of BigNumbers at the API boundary
// (ipcl/include/ipcl/base_text.hpp:14-115).
#ifndef IPCL_B200_BASE_TEXT_HPP_
#define IPCL_B200_BASE_TEXT_HPP_

#include <atomic>
#include <memory>
#include <string>
#include <vector>

#include "ipcl/bignum.h"

namespace ipcl {

namespace detail {
// a batch of big integers in HBM, count x words little-endian limbs -- the
// layout of the C ABI (src/device_batch.hpp)
struct DeviceBatch;
// the flat host image of a device batch (src/marshal.hpp)
struct FlatImage;
}  // namespace detail

class BaseText {
 public:
  BaseText() = default;
  ~BaseText() = default;

  explicit BaseText(const uint32_t& n);
  explicit BaseText(const std::vector<uint32_t>& n_v);
  explicit BaseText(const BigNumber& bn);
  explicit BaseText(const std::vector<BigNumber>& bn_v);
  // takes ownership of a freshly unmarshalled batch (no second copy)
  explicit BaseText(std::vector<BigNumber>&& bn_v);
  BaseText(const BaseText& bt);
  BaseText& operator=(const BaseText& other);

  BigNumber& operator[](const std::size_t idx);
  void insert(const std::size_t pos, BigNumber& bn);
  void clear();
  void remove(const std::size_t pos, const std::size_t length = 1);

  BigNumber getElement(const std::size_t& idx) const;
  std::vector<uint32_t> getElementVec(const std::size_t& idx) const;
  std::string getElementHex(const std::size_t& idx) const;
  std::vector<BigNumber> getChunk(const std::size_t& start,
                                  const std::size_t& size) const;
  std::vector<BigNumber> getTexts() const;
  // no-copy view of the container (the reference only has the copying
  // getTexts(), base_text.cpp:102, which its own hot loops pay for)
  const std::vector<BigNumber>& texts() const;
  std::size_t getSize() const;

  // ---- device-resident form (this back-end) ----
  // A text produced by PublicKey::encrypt, the CipherText operators or
  // PrivateKey::decrypt lives in HBM; the vector<BigNumber> is only built when
  // a caller looks at the values (any accessor above).  Chains such as
  // encrypt -> + -> * -> decrypt therefore never leave the GPU.
  // isDeviceResident(): the batch is in HBM.  isHostMaterialized(): the
  // BigNumbers exist on the host.  deviceBatch(words): the batch as
  // count x `words` limbs in HBM, uploading it first if it only exists on the
  // host; nullptr if an element is negative or wider than `words`.
  bool isDeviceResident() const { return static_cast<bool>(m_dev); }
  bool isHostMaterialized() const { return m_host_valid.load(); }
  std::shared_ptr<detail::DeviceBatch> deviceBatch(int words) const;

  const void* addr = static_cast<const void*>(this);

 protected:
  explicit BaseText(std::shared_ptr<detail::DeviceBatch> dev);
  // builds m_texts from the device batch if that has not happened yet
  void ensureHost() const;
  // before a mutation: materialise, then forget the (now stale) device copy
  void hostOnly();
  // single-element reads of a device-resident text: ONE download of the flat
  // limb image, a BigNumber only for the element asked for (the reference
  // copies the whole vector<BigNumber>, base_text.cpp:102)
  const detail::FlatImage* flatImage() const;
  BigNumber elementFromFlat(std::size_t idx) const;

  mutable std::vector<BigNumber> m_texts;
  std::size_t m_size = 0;
  mutable std::shared_ptr<detail::DeviceBatch> m_dev;
  mutable std::shared_ptr<detail::FlatImage> m_flat;
  mutable std::atomic<bool> m_host_valid{true};
};

}  // namespace ipcl
#endif  // IPCL_B200_BASE_TEXT_HPP_
