// base_text.hpp -- container of BigNumbers at the API boundary
// (ipcl/include/ipcl/base_text.hpp:14-115).
#ifndef IPCL_B200_BASE_TEXT_HPP_
#define IPCL_B200_BASE_TEXT_HPP_

#include <string>
#include <vector>

#include "ipcl/bignum.h"

namespace ipcl {

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
  const std::vector<BigNumber>& texts() const { return m_texts; }
  std::size_t getSize() const;

  const void* addr = static_cast<const void*>(this);

 protected:
  std::vector<BigNumber> m_texts;
  std::size_t m_size = 0;
};

}  // namespace ipcl
#endif  // IPCL_B200_BASE_TEXT_HPP_
