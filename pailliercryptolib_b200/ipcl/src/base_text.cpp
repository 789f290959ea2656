// base_text.cpp -- ipcl::BaseText, the container of BigNumbers at the API
// boundary (interface: /root/reference/ipcl/include/ipcl/base_text.hpp:14-115;
// error texts as in /root/reference/ipcl/base_text.cpp so callers that match on
// them keep working).
#include "ipcl/base_text.hpp"

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <utility>

#include "device_batch.hpp"
#include "ipcl/utils/util.hpp"
#include "text_util.hpp"

namespace ipcl {

namespace detail {

std::vector<BigNumber> rotated(const std::vector<BigNumber>& v, int shift) {
  const int size = static_cast<int>(v.size());
  ERROR_CHECK(size != 1, "rotate: Cannot rotate single CipherText");
  ERROR_CHECK(shift >= -size && shift <= size,
              "rotate: Cannot shift more than the test size");
  std::vector<BigNumber> out(v);
  if (size == 0 || shift % size == 0) return out;
  // a positive shift moves element i to i + shift
  const int left = shift > 0 ? size - shift : -shift;
  std::rotate(out.begin(), out.begin() + left, out.end());
  return out;
}

// Page-locked staging buffers are kept in a process-wide pool that is never
// destroyed: a thread borrows one on first use and hands it back when it ends.
// Nothing is freed at process exit, where the CUDA runtime may already be gone.
namespace {
struct PinnedSlab {
  uint32_t* p = nullptr;
  std::size_t words = 0;
  bool pinned = false;
};
struct PinnedPool {
  std::mutex mu;
  std::vector<PinnedSlab> idle;
};
PinnedPool& pinnedPool() {
  static PinnedPool* pool = new PinnedPool;  // intentionally leaked
  return *pool;
}
void releaseSlab(PinnedSlab& s) {
  if (!s.p) return;
  if (s.pinned)
    ipclb200_host_free(s.p);
  else
    std::free(s.p);
  s = PinnedSlab{};
}
}  // namespace

PinnedBuffer& PinnedBuffer::forThread() {
  thread_local PinnedBuffer b;
  return b;
}

uint32_t* PinnedBuffer::get(std::size_t words) {
  if (words <= m_words) return m_p;
  PinnedSlab cur{m_p, m_words, m_pinned};
  m_p = nullptr;
  m_words = 0;
  {
    // a large enough idle slab?
    PinnedPool& pool = pinnedPool();
    std::lock_guard<std::mutex> lk(pool.mu);
    for (std::size_t i = 0; i < pool.idle.size(); i++)
      if (pool.idle[i].words >= words) {
        PinnedSlab got = pool.idle[i];
        pool.idle.erase(pool.idle.begin() + static_cast<std::ptrdiff_t>(i));
        if (cur.p) pool.idle.push_back(cur);
        m_p = got.p;
        m_words = got.words;
        m_pinned = got.pinned;
        return m_p;
      }
  }
  releaseSlab(cur);
  const std::size_t want = words + words / 4 + 1024;
  void* p = nullptr;
  if (ipclb200_host_alloc(want * sizeof(uint32_t), &p) == 0) {
    m_pinned = true;
  } else {
    p = std::malloc(want * sizeof(uint32_t));
    ERROR_CHECK(p != nullptr, "out of host memory");
    m_pinned = false;
  }
  m_p = static_cast<uint32_t*>(p);
  m_words = want;
  return m_p;
}

PinnedBuffer::~PinnedBuffer() {
  if (!m_p) return;
  PinnedPool& pool = pinnedPool();
  std::lock_guard<std::mutex> lk(pool.mu);
  pool.idle.push_back(PinnedSlab{m_p, m_words, m_pinned});
}

bool deviceResidentEnabled() {
  static const bool on = [] {
    const char* e = std::getenv("IPCL_B200_DEVICE_RESIDENT");
    return !(e && e[0] == '0');
  }();
  return on;
}

}  // namespace detail

// lazy host materialisation happens inside const accessors that several
// threads may call on one text; one process-wide lock is enough (it is taken
// once per text)
static std::mutex g_materialize_mutex;

BaseText::BaseText(std::shared_ptr<detail::DeviceBatch> dev)
    : m_size(dev->count), m_dev(std::move(dev)), m_host_valid(false) {}

void BaseText::ensureHost() const {
  if (m_host_valid.load(std::memory_order_acquire)) return;
  std::lock_guard<std::mutex> lk(g_materialize_mutex);
  if (m_host_valid.load(std::memory_order_relaxed)) return;
  m_texts = m_dev->toHost();
  m_host_valid.store(true, std::memory_order_release);
}

void BaseText::hostOnly() {
  ensureHost();
  m_dev.reset();
}

std::shared_ptr<detail::DeviceBatch> BaseText::deviceBatch(int words) const {
  std::lock_guard<std::mutex> lk(g_materialize_mutex);
  if (m_dev && m_dev->words == words) return m_dev;
  if (!m_host_valid.load()) {
    // resident at another width: go through the host form
    m_texts = m_dev->toHost();
    m_host_valid.store(true);
  }
  if (!detail::DeviceBatch::fits(m_texts, words)) return nullptr;
  auto b = detail::DeviceBatch::fromHost(m_texts, words);
  if (!m_dev) m_dev = b;  // keep it for the next operator on this text
  return b;
}

const std::vector<BigNumber>& BaseText::texts() const {
  ensureHost();
  return m_texts;
}

#define TEXT_INDEX_CHECK(i, what) \
  ERROR_CHECK((i) < m_size, "BaseText: " what " index is out of range")

BaseText::BaseText(const uint32_t& n) : m_texts(1, BigNumber(n)), m_size(1) {}

BaseText::BaseText(const std::vector<uint32_t>& n_v) : m_size(n_v.size()) {
  m_texts.reserve(m_size);
  std::transform(n_v.begin(), n_v.end(), std::back_inserter(m_texts),
                 [](uint32_t w) { return BigNumber(w); });
}

BaseText::BaseText(const BigNumber& bn) : m_texts(1, bn), m_size(1) {}

BaseText::BaseText(const std::vector<BigNumber>& bn_v)
    : m_texts(bn_v), m_size(bn_v.size()) {}

BaseText::BaseText(std::vector<BigNumber>&& bn_v)
    : m_texts(std::move(bn_v)), m_size(m_texts.size()) {}

// a copy shares the (immutable) device batch and copies the host form only if
// it exists
BaseText::BaseText(const BaseText& bt) : m_size(bt.m_size) {
  std::lock_guard<std::mutex> lk(g_materialize_mutex);
  m_dev = bt.m_dev;
  const bool valid = bt.m_host_valid.load();
  if (valid) m_texts = bt.m_texts;
  m_host_valid.store(valid);
}

BaseText& BaseText::operator=(const BaseText& other) {
  if (this != &other) {
    std::lock_guard<std::mutex> lk(g_materialize_mutex);
    m_size = other.m_size;
    m_dev = other.m_dev;
    const bool valid = other.m_host_valid.load();
    if (valid)
      m_texts = other.m_texts;
    else
      m_texts.clear();
    m_host_valid.store(valid);
  }
  return *this;
}

BigNumber& BaseText::operator[](const std::size_t idx) {
  ERROR_CHECK(idx < m_size, "BaseText:operator[] index is out of range");
  hostOnly();  // the caller may write through the reference
  return m_texts[idx];
}

BigNumber BaseText::getElement(const std::size_t& idx) const {
  TEXT_INDEX_CHECK(idx, "getElement");
  ensureHost();
  return m_texts[idx];
}

std::vector<uint32_t> BaseText::getElementVec(const std::size_t& idx) const {
  TEXT_INDEX_CHECK(idx, "getElementVec");
  ensureHost();
  std::vector<uint32_t> words;
  m_texts[idx].num2vec(words);
  return words;
}

std::string BaseText::getElementHex(const std::size_t& idx) const {
  TEXT_INDEX_CHECK(idx, "getElementHex");
  ensureHost();
  std::string hex;
  m_texts[idx].num2hex(hex);
  return hex;
}

std::vector<BigNumber> BaseText::getChunk(const std::size_t& start,
                                          const std::size_t& size) const {
  ERROR_CHECK(start + size <= m_size, "BaseText: getChunk parameter is incorrect");
  ensureHost();
  return {m_texts.begin() + static_cast<std::ptrdiff_t>(start),
          m_texts.begin() + static_cast<std::ptrdiff_t>(start + size)};
}

void BaseText::insert(const std::size_t pos, BigNumber& bn) {
  ERROR_CHECK(pos <= m_size, "BaseText: insert position is out of range");
  hostOnly();
  m_texts.insert(m_texts.begin() + static_cast<std::ptrdiff_t>(pos), bn);
  m_size = m_texts.size();
}

void BaseText::remove(const std::size_t pos, const std::size_t length) {
  ERROR_CHECK(pos + length < m_size, "BaseText: remove position is out of range");
  hostOnly();
  const auto first = m_texts.begin() + static_cast<std::ptrdiff_t>(pos);
  m_texts.erase(first, first + static_cast<std::ptrdiff_t>(length));
  m_size = m_texts.size();
}

void BaseText::clear() {
  m_dev.reset();
  m_texts.clear();
  m_host_valid.store(true);
  m_size = 0;
}

std::vector<BigNumber> BaseText::getTexts() const {
  ensureHost();
  return m_texts;
}

std::size_t BaseText::getSize() const { return m_size; }

}  // namespace ipcl
