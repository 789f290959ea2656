// base_text.cpp -- ipcl::BaseText (reference: ipcl/base_text.cpp).
#include "ipcl/base_text.hpp"

#include <utility>

#include "ipcl/utils/util.hpp"

namespace ipcl {

BaseText::BaseText(const uint32_t& n) : m_texts{BigNumber(n)}, m_size(1) {}

BaseText::BaseText(const std::vector<uint32_t>& n_v) {
  m_texts.reserve(n_v.size());
  for (uint32_t n : n_v) m_texts.emplace_back(n);
  m_size = m_texts.size();
}

BaseText::BaseText(const BigNumber& bn) : m_texts{bn}, m_size(1) {}

BaseText::BaseText(const std::vector<BigNumber>& bn_v)
    : m_texts(bn_v), m_size(bn_v.size()) {}

BaseText::BaseText(std::vector<BigNumber>&& bn_v)
    : m_texts(std::move(bn_v)), m_size(m_texts.size()) {}

BaseText::BaseText(const BaseText& bt)
    : m_texts(bt.m_texts), m_size(bt.m_size) {}

BaseText& BaseText::operator=(const BaseText& other) {
  if (this != &other) {
    m_texts = other.m_texts;
    m_size = other.m_size;
  }
  return *this;
}

BigNumber& BaseText::operator[](const std::size_t idx) {
  ERROR_CHECK(idx < m_size, "BaseText:operator[] index is out of range");
  return m_texts[idx];
}

void BaseText::insert(const std::size_t pos, BigNumber& bn) {
  ERROR_CHECK(pos <= m_size, "BaseText: insert position is out of range");
  m_texts.insert(m_texts.begin() + static_cast<std::ptrdiff_t>(pos), bn);
  m_size++;
}

void BaseText::clear() {
  m_texts.clear();
  m_size = 0;
}

void BaseText::remove(const std::size_t pos, const std::size_t length) {
  ERROR_CHECK(pos + length < m_size, "BaseText: remove position is out of range");
  auto first = m_texts.begin() + static_cast<std::ptrdiff_t>(pos);
  m_texts.erase(first, first + static_cast<std::ptrdiff_t>(length));
  m_size -= length;
}

BigNumber BaseText::getElement(const std::size_t& idx) const {
  ERROR_CHECK(idx < m_size, "BaseText: getElement index is out of range");
  return m_texts[idx];
}

std::vector<uint32_t> BaseText::getElementVec(const std::size_t& idx) const {
  ERROR_CHECK(idx < m_size, "BaseText: getElementVec index is out of range");
  std::vector<uint32_t> v;
  m_texts[idx].num2vec(v);
  return v;
}

std::string BaseText::getElementHex(const std::size_t& idx) const {
  ERROR_CHECK(idx < m_size, "BaseText: getElementHex index is out of range");
  std::string s;
  m_texts[idx].num2hex(s);
  return s;
}

std::vector<BigNumber> BaseText::getChunk(const std::size_t& start,
                                          const std::size_t& size) const {
  ERROR_CHECK((start + size) <= m_size, "BaseText: getChunk parameter is incorrect");
  auto first = m_texts.begin() + static_cast<std::ptrdiff_t>(start);
  return std::vector<BigNumber>(first, first + static_cast<std::ptrdiff_t>(size));
}

std::vector<BigNumber> BaseText::getTexts() const { return m_texts; }

std::size_t BaseText::getSize() const { return m_size; }

}  // namespace ipcl
