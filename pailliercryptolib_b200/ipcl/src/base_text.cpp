// base_text.cpp -- ipcl::BaseText, the container of BigNumbers at the API
// boundary (interface: /root/reference/ipcl/include/ipcl/base_text.hpp:14-115;
// error texts as in /root/reference/ipcl/base_text.cpp so callers that match on
// them keep working).
#include "ipcl/base_text.hpp"

#include <algorithm>
#include <cstdint>
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

// ---- pool of page-locked slabs (marshal.hpp) -------------------------------------
namespace {
struct SlabPool {
  std::mutex mu;
  std::vector<HostSlab> idle;
  std::size_t idle_words = 0;
};
SlabPool& slabPool() {
  static SlabPool* pool = new SlabPool;  // intentionally leaked, see marshal.hpp
  return *pool;
}
constexpr std::size_t kMaxIdleWords = (std::size_t)1 << 28;  // 1 GB kept for reuse
void freeSlab(HostSlab& s) {
  if (!s.p) return;
  if (s.pinned)
    ipclb200_host_free(s.p);
  else
    std::free(s.p);
  s = HostSlab{};
}
}  // namespace

HostSlab acquireHostSlab(std::size_t words) {
  if (words == 0) words = 4;
  {
    SlabPool& pool = slabPool();
    std::lock_guard<std::mutex> lk(pool.mu);
    // best fit among the idle slabs (not more than 2x too large)
    std::size_t best = pool.idle.size();
    for (std::size_t i = 0; i < pool.idle.size(); i++)
      if (pool.idle[i].words >= words && pool.idle[i].words <= 2 * words + 4096 &&
          (best == pool.idle.size() || pool.idle[i].words < pool.idle[best].words))
        best = i;
    if (best != pool.idle.size()) {
      HostSlab got = pool.idle[best];
      pool.idle.erase(pool.idle.begin() + static_cast<std::ptrdiff_t>(best));
      pool.idle_words -= got.words;
      return got;
    }
  }
  HostSlab s;
  const std::size_t want = words + words / 8 + 1024;
  void* p = nullptr;
  if (ipclb200_host_alloc(want * sizeof(uint32_t), &p) == 0) {
    s.pinned = true;
  } else {
    p = std::malloc(want * sizeof(uint32_t));
    ERROR_CHECK(p != nullptr, "out of host memory");
    s.pinned = false;
  }
  s.p = static_cast<uint32_t*>(p);
  s.words = want;
  return s;
}

void returnHostSlab(HostSlab& s) {
  if (!s.p) return;
  SlabPool& pool = slabPool();
  {
    std::lock_guard<std::mutex> lk(pool.mu);
    if (pool.idle_words + s.words <= kMaxIdleWords) {
      pool.idle.push_back(s);
      pool.idle_words += s.words;
      s = HostSlab{};
      return;
    }
  }
  freeSlab(s);
}

PinnedBuffer& PinnedBuffer::forThread() {
  thread_local PinnedBuffer b;
  return b;
}

uint32_t* PinnedBuffer::get(std::size_t words) {
  if (words > m_slab.words) {
    returnHostSlab(m_slab);
    m_slab = acquireHostSlab(words);
  }
  return m_slab.p;
}

// thread exit: the slab goes back to the pool (nothing is freed at process exit)
PinnedBuffer::~PinnedBuffer() {
  if (!m_slab.p) return;
  SlabPool& pool = slabPool();
  std::lock_guard<std::mutex> lk(pool.mu);
  pool.idle.push_back(m_slab);
  pool.idle_words += m_slab.words;
}

bool deviceResidentEnabled() {
  static const bool on = [] {
    const char* e = std::getenv("IPCL_B200_DEVICE_RESIDENT");
    return !(e && e[0] == '0');
  }();
  return on;
}

}  // namespace detail

// Lazy host materialisation happens inside const accessors that several
// threads may call on one text.  The lock is held while a download completes,
// so it must not be shared between unrelated texts (four callers working on
// their own batches would queue on it): one of 64 locks chosen by the text's
// address.
static std::mutex& lockFor(const void* text) {
  static std::mutex locks[64];
  return locks[(reinterpret_cast<std::uintptr_t>(text) >> 6) & 63];
}

BaseText::BaseText(std::shared_ptr<detail::DeviceBatch> dev)
    : m_size(dev->count), m_dev(std::move(dev)), m_host_valid(false) {}

void BaseText::ensureHost() const {
  if (m_host_valid.load(std::memory_order_acquire)) return;
  std::lock_guard<std::mutex> lk(lockFor(this));
  if (m_host_valid.load(std::memory_order_relaxed)) return;
  if (m_flat)
    m_texts = detail::unpack(m_flat->slab.p, m_flat->count, m_flat->words);
  else
    m_texts = m_dev->toHost();
  m_host_valid.store(true, std::memory_order_release);
}

// the flat host image of a device-resident text: ONE download, no BigNumber
// objects; single-element accessors read from it
const detail::FlatImage* BaseText::flatImage() const {
  std::lock_guard<std::mutex> lk(lockFor(this));
  if (!m_flat && m_dev) {
    auto img = std::make_shared<detail::FlatImage>(m_dev->count, m_dev->words);
    m_dev->downloadFlat(img->slab.p);
    m_flat = std::move(img);
  }
  return m_flat.get();
}

BigNumber BaseText::elementFromFlat(std::size_t idx) const {
  const detail::FlatImage* f = flatImage();
  BigNumber r;
  r.Set(f->element(idx), f->words, IppsBigNumPOS);
  return r;
}

void BaseText::hostOnly() {
  ensureHost();
  m_dev.reset();
  m_flat.reset();
}

std::shared_ptr<detail::DeviceBatch> BaseText::deviceBatch(int words) const {
  std::lock_guard<std::mutex> lk(lockFor(this));
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
  std::lock_guard<std::mutex> lk(lockFor(&bt));
  m_dev = bt.m_dev;
  m_flat = bt.m_flat;
  const bool valid = bt.m_host_valid.load();
  if (valid) m_texts = bt.m_texts;
  m_host_valid.store(valid);
}

BaseText& BaseText::operator=(const BaseText& other) {
  if (this != &other) {
    // both texts: the source may be materialising, the target may be read
    std::mutex &ma = lockFor(this), &mb = lockFor(&other);
    std::unique_lock<std::mutex> la(ma, std::defer_lock), lb(mb, std::defer_lock);
    if (&ma == &mb) {
      la.lock();
    } else {
      std::lock(la, lb);
    }
    m_size = other.m_size;
    m_dev = other.m_dev;
    m_flat = other.m_flat;
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
  if (!m_host_valid.load(std::memory_order_acquire)) return elementFromFlat(idx);
  return m_texts[idx];
}

std::vector<uint32_t> BaseText::getElementVec(const std::size_t& idx) const {
  TEXT_INDEX_CHECK(idx, "getElementVec");
  std::vector<uint32_t> words;
  if (!m_host_valid.load(std::memory_order_acquire))
    elementFromFlat(idx).num2vec(words);
  else
    m_texts[idx].num2vec(words);
  return words;
}

std::string BaseText::getElementHex(const std::size_t& idx) const {
  TEXT_INDEX_CHECK(idx, "getElementHex");
  std::string hex;
  if (!m_host_valid.load(std::memory_order_acquire))
    elementFromFlat(idx).num2hex(hex);
  else
    m_texts[idx].num2hex(hex);
  return hex;
}

std::vector<BigNumber> BaseText::getChunk(const std::size_t& start,
                                          const std::size_t& size) const {
  ERROR_CHECK(start + size <= m_size, "BaseText: getChunk parameter is incorrect");
  if (!m_host_valid.load(std::memory_order_acquire)) {
    std::vector<BigNumber> out(size);
    for (std::size_t i = 0; i < size; i++) out[i] = elementFromFlat(start + i);
    return out;
  }
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
  m_flat.reset();
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
