// ipcl_b200.cu -- implementation of the C ABI in include/ipcl_b200.h.
//
// Host side of the boundary: argument checking, derivation of the per-modulus
// Montgomery constants (hostbn.hpp), staging of the flat limb buffers, launch
// configuration, sharding of a batch over the GPUs of the box, and the error
// convention.  All arithmetic on batch elements happens in the kernels of
// kernels.cuh; nothing here falls back to the CPU.
//
// Runtime model
//   * one Dev per CUDA device in use (created lazily), with a pool of streams;
//     every host-pointer call takes its own stream from the pool and allocates
//     its temporaries stream-ordered (cudaMallocAsync), so concurrent callers
//     -- the reference calls encrypt/decrypt from 4 OpenMP threads on one key,
//     test/test_cryptography.cpp:45-57 -- overlap instead of queueing on a lock;
//   * keys keep their host-side constants and replicate the device blocks per
//     device on first use;
//   * host-pointer entry points split a batch into contiguous blocks over the
//     active devices (ipclb200_init_devices): the slot of the reference's
//     prefix/suffix split between CPU and accelerator, ipcl/mod_exp.cpp:702-731;
//   * ipclb200_batch_* hold a batch sharded over the active devices in HBM;
//     scatter/gather from/to one device go through NCCL send/recv.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ipcl_b200.h"
#include "host_common.hpp"
#include "hostbn.hpp"
#include "kernels.cuh"
#include "hensel_launch.hpp"

using namespace ipclb200;
using namespace ipclb200::host;
using hbn::Limbs;

namespace {

constexpr int kMaxDevices = 64;

// ---------------------------------------------------------------------------
// per-modulus constants on one device
// ---------------------------------------------------------------------------
struct DevModulus {
  int L = 0;
  int device = -1;
  Limbs n;
  uint32_t* d = nullptr;  // [n | rr | r3 | one] each L words, then n0inv
  const uint32_t* d_n0inv = nullptr;
  ModConst mc{};
  // the modulus is the square of a root that fills its LH = L/2 words (n^2 of a
  // Paillier key): modexp_hensel_kernel applies.  d_root: root | pairs of R^2, R^3
  int root_words = 0;
  uint32_t* d_root = nullptr;
  uint32_t root_n0inv = 0;
  ~DevModulus() {
    cudaSetDevice(device);
    if (d) cudaFree(d);
    if (d_root) cudaFree(d_root);
  }
};

// ---------------------------------------------------------------------------
// devices
// ---------------------------------------------------------------------------
struct Dev {
  int id = -1;
  int sms = 0;
  cudaStream_t stream = nullptr;  // the device's library stream (batches, dev_*)
  cudaStream_t side = nullptr;    // fixed-base table builds
  std::mutex mu;                  // stream pool, modulus cache, table accounting
  std::vector<cudaStream_t> idle;
  std::vector<std::shared_ptr<DevModulus>> mod_cache;
  size_t comb_bytes = 0;  // bytes of wide fixed-base tables alive on this device
};

struct Runtime {
  std::mutex mu;
  std::atomic<Dev*> dev[kMaxDevices];
  std::vector<int> active;  // device ordinals host-pointer batches are split over
  bool active_set = false;
  std::atomic<uint64_t> launches{0};
  std::atomic<uint64_t> zero_copies{0};
  std::atomic<uint64_t> use_clock{0};
  Runtime() {
    for (auto& d : dev) d.store(nullptr);
  }
};
Runtime g;

int device_count_raw() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

// the Dev of CUDA device `id`, created on first use
int dev_get(int id, Dev** out) {
  if (id < 0 || id >= kMaxDevices) return fail(IPCLB200_ERR_NO_DEVICE, "no such CUDA device");
  Dev* d = g.dev[id].load(std::memory_order_acquire);
  if (!d) {
    std::lock_guard<std::mutex> lk(g.mu);
    d = g.dev[id].load(std::memory_order_acquire);
    if (!d) {
      if (id >= device_count_raw())
        return fail(IPCLB200_ERR_NO_DEVICE, "no such CUDA device");
      cudaDeviceProp prop;
      CUDA_TRY(cudaGetDeviceProperties(&prop, id));
      if (prop.major != 10)
        return fail(IPCLB200_ERR_NO_DEVICE,
                    std::string("device is sm_") + std::to_string(prop.major) +
                        std::to_string(prop.minor) +
                        ", this library holds sm_100a code only");
      CUDA_TRY(cudaSetDevice(id));
      std::unique_ptr<Dev> nd(new Dev);
      nd->id = id;
      nd->sms = prop.multiProcessorCount;
      CUDA_TRY(cudaStreamCreateWithFlags(&nd->stream, cudaStreamNonBlocking));
      CUDA_TRY(cudaStreamCreateWithFlags(&nd->side, cudaStreamNonBlocking));
      {
        // temporaries and batches come from the stream-ordered pool: keep freed
        // blocks cached instead of returning them to the OS at every synchronise
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, id) == cudaSuccess) {
          uint64_t keep = ~(uint64_t)0;
          cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
      }
      d = nd.release();
      g.dev[id].store(d, std::memory_order_release);
    }
  }
  *out = d;
  return 0;
}

// the devices a host-pointer batch is split over; the first one is the primary
// device (key set-up scalars, ipclb200_stream(), dev_alloc ...)
int active_devices(std::vector<Dev*>* out) {
  std::vector<int> ids;
  {
    std::lock_guard<std::mutex> lk(g.mu);
    if (!g.active_set) {
      const int ndev = device_count_raw();
      if (ndev == 0) return fail(IPCLB200_ERR_NO_DEVICE, "no CUDA device");
      int dev = 0;
      const char* many = getenv("IPCLB200_DEVICES");
      if (many && *many) {
        // one process, several GPUs without an explicit ipclb200_init_devices
        int n = !strcmp(many, "all") ? ndev : atoi(many);
        if (n <= 0 || n > ndev) n = ndev;
        g.active.clear();
        for (int i = 0; i < n; i++) g.active.push_back(i);
      } else {
        if (const char* lr = getenv("LOCAL_RANK")) {
          dev = atoi(lr) % ndev;  // one process per GPU under torchrun
        } else if (cudaGetDevice(&dev) != cudaSuccess) {
          cudaGetLastError();
          dev = 0;
        }
        g.active = {dev};
      }
      g.active_set = true;
    }
    ids = g.active;
  }
  out->clear();
  for (int id : ids) {
    Dev* d = nullptr;
    TRY(dev_get(id, &d));
    out->push_back(d);
  }
  return 0;
}

int primary_device(Dev** out) {
  std::vector<Dev*> devs;
  TRY(active_devices(&devs));
  *out = devs[0];
  return 0;
}

// the Dev that owns a device pointer
int dev_of_pointer(const void* p, Dev** out) {
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess ||
      (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged)) {
    cudaGetLastError();
    if (device_count_raw() == 0) return fail(IPCLB200_ERR_NO_DEVICE, "no CUDA device");
    return fail(IPCLB200_ERR_BAD_ARG, "not a device pointer");
  }
  return dev_get(at.device, out);
}

// ---------------------------------------------------------------------------
// one operation on one device: a stream plus stream-ordered temporaries
// ---------------------------------------------------------------------------
struct Op {
  Dev* dev = nullptr;
  cudaStream_t s = nullptr;
  bool pooled = false;
  std::vector<void*> tmp;

  Op() = default;
  Op(const Op&) = delete;
  Op& operator=(const Op&) = delete;
  Op(Op&& o) noexcept { *this = std::move(o); }
  Op& operator=(Op&& o) noexcept {
    dev = o.dev;
    s = o.s;
    pooled = o.pooled;
    tmp = std::move(o.tmp);
    o.dev = nullptr;
    o.s = nullptr;
    o.pooled = false;
    return *this;
  }

  // on a stream of the library's pool (host-pointer calls)
  int open(Dev* d) {
    dev = d;
    CUDA_TRY(cudaSetDevice(d->id));
    {
      std::lock_guard<std::mutex> lk(d->mu);
      if (!d->idle.empty()) {
        s = d->idle.back();
        d->idle.pop_back();
      }
    }
    if (!s) CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    pooled = true;
    return 0;
  }
  // on a caller's stream (*_dev calls, batches): only enqueues
  int open_on(Dev* d, cudaStream_t user) {
    dev = d;
    CUDA_TRY(cudaSetDevice(d->id));
    s = user;
    pooled = false;
    return 0;
  }
  int alloc(size_t bytes, void** out) {
    void* p = nullptr;
    CUDA_TRY(cudaMallocAsync(&p, bytes ? bytes : 16, s));
    tmp.push_back(p);
    *out = p;
    return 0;
  }
  int words(size_t n, uint32_t** out) { return alloc(n * sizeof(uint32_t), (void**)out); }
  int sync() {
    CUDA_TRY(cudaSetDevice(dev->id));
    CUDA_TRY(cudaStreamSynchronize(s));
    return 0;
  }
  // releases the temporaries (stream-ordered: after the work enqueued so far)
  void close() {
    if (!dev) return;
    cudaSetDevice(dev->id);
    for (void* p : tmp) cudaFreeAsync(p, s);
    tmp.clear();
    if (pooled && s) {
      std::lock_guard<std::mutex> lk(dev->mu);
      dev->idle.push_back(s);
    }
    dev = nullptr;
    s = nullptr;
  }
  ~Op() { close(); }
};

// The stream batch operations of the calling thread run on (one per thread and
// device, never returned: threads are few).  Batches carry an event, so work
// of different threads on different batches overlaps while every batch sees its
// own operations in order.
struct ThreadStreams {
  cudaStream_t s[kMaxDevices] = {};
  Dev* owner[kMaxDevices] = {};
  ~ThreadStreams();
};

cudaStream_t thread_stream(Dev* d) {
  thread_local ThreadStreams ts;
  if (ts.owner[d->id] != d || !ts.s[d->id]) {
    cudaSetDevice(d->id);
    cudaStream_t s = nullptr;
    {
      // streams of threads that have ended wait in the device's pool (with the
      // pool memory the allocator associates with them)
      std::lock_guard<std::mutex> lk(d->mu);
      if (!d->idle.empty()) {
        s = d->idle.back();
        d->idle.pop_back();
      }
    }
    if (!s && cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) {
      cudaGetLastError();
      return d->stream;
    }
    ts.s[d->id] = s;
    ts.owner[d->id] = d;
  }
  return ts.s[d->id];
}

ThreadStreams::~ThreadStreams() {
  for (int i = 0; i < kMaxDevices; i++) {
    if (!s[i]) continue;
    Dev* d = g.dev[i].load();
    if (d && d == owner[i]) {  // not after ipclb200_shutdown()
      std::lock_guard<std::mutex> lk(d->mu);
      d->idle.push_back(s[i]);
    }
  }
}

// restores the caller's current device when a multi-device call returns
struct DeviceGuard {
  int prev = -1;
  DeviceGuard() {
    if (cudaGetDevice(&prev) != cudaSuccess) {
      cudaGetLastError();
      prev = -1;
    }
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// contiguous block partition of a batch over the active devices.
// A device gets a block only if at least `min_block` elements are left for it,
// so that small batches stay on one GPU.
struct Shard {
  Dev* dev;
  size_t begin, count;
};
constexpr size_t kMinShard = 512;

int plan_shards(size_t count, std::vector<Shard>* out) {
  std::vector<Dev*> devs;
  TRY(active_devices(&devs));
  size_t min_block = kMinShard;
  if (const char* e = getenv("IPCLB200_MIN_SHARD")) min_block = strtoul(e, nullptr, 10);
  if (min_block < 1) min_block = 1;
  size_t use = std::min<size_t>(devs.size(), std::max<size_t>(1, count / min_block));
  out->clear();
  const size_t base = count / use, rem = count % use;
  size_t at = 0;
  for (size_t i = 0; i < use; i++) {
    const size_t c = base + (i < rem ? 1 : 0);
    out->push_back({devs[i], at, c});
    at += c;
  }
  return 0;
}

// ---------------------------------------------------------------------------
// modulus cache
// ---------------------------------------------------------------------------
int make_modulus(Dev* dev, const Limbs& n, int L, std::shared_ptr<DevModulus>* out) {
  {
    std::lock_guard<std::mutex> lk(dev->mu);
    for (auto& m : dev->mod_cache)
      if (m->L == L && m->n == n) {
        *out = m;
        return 0;
      }
  }
  auto m = std::make_shared<DevModulus>();
  HostModConst h;
  host_mod_const(n, L, &h);
  m->L = L;
  m->n = n;
  m->device = dev->id;
  CUDA_TRY(cudaSetDevice(dev->id));
  CUDA_TRY(cudaMalloc(&m->d, sizeof(uint32_t) * (4 * (size_t)L + 4)));
  std::vector<uint32_t> blk;
  blk.insert(blk.end(), h.n.begin(), h.n.end());
  blk.insert(blk.end(), h.rr.begin(), h.rr.end());
  blk.insert(blk.end(), h.r3.begin(), h.r3.end());
  blk.insert(blk.end(), h.one.begin(), h.one.end());
  blk.push_back(h.n0inv);
  CUDA_TRY(cudaMemcpy(m->d, blk.data(), blk.size() * sizeof(uint32_t),
                      cudaMemcpyHostToDevice));
  m->mc.n = m->d;
  m->mc.rr = m->d + L;
  m->mc.r3 = m->d + 2 * L;
  m->mc.one = m->d + 3 * L;
  m->d_n0inv = m->d + 4 * L;
  m->mc.n0inv = h.n0inv;
  m->mc.small_mod = h.small_mod;
  if ((L == 64 || L == 128 || L == 192 || L == 256) && hbn::bitlen(n) > 32 * (L - 1)) {
    // a perfect square whose root fills L/2 words?  (ct * pt works mod n^2)
    bool exact = false;
    Limbs root = hbn::isqrt(n, &exact);
    const int lh = L / 2;
    if (exact && (root[0] & 1u) && hbn::bitlen(root) == 32 * lh) {
      std::vector<uint32_t> rb(5 * (size_t)lh, 0u);
      hbn::to_words(root, rb.data(), lh);
      const Limbs Rh = hbn::pow2(32u * (unsigned)lh);
      Limbs t = hbn::mod(hbn::mul(Rh, Rh), n);  // R^2
      for (int j = 0; j < 2; j++) {
        Limbs hi, lo;
        hbn::divmod(t, root, &hi, &lo);
        Limbs wneg = hbn::mod(hi, root);
        Limbs w = hbn::is_zero(wneg) ? wneg : hbn::sub(root, wneg);
        hbn::to_words(lo, rb.data() + (size_t)(1 + 2 * j) * lh, lh);
        hbn::to_words(w, rb.data() + (size_t)(2 + 2 * j) * lh, lh);
        t = hbn::mod(hbn::mul(t, Rh), n);  // R^3
      }
      CUDA_TRY(cudaMalloc(&m->d_root, rb.size() * sizeof(uint32_t)));
      CUDA_TRY(cudaMemcpy(m->d_root, rb.data(), rb.size() * sizeof(uint32_t),
                          cudaMemcpyHostToDevice));
      m->root_words = lh;
      m->root_n0inv = hbn::neg_inv32(root[0]);
    }
  }
  std::lock_guard<std::mutex> lk(dev->mu);
  if (dev->mod_cache.size() >= 32) dev->mod_cache.erase(dev->mod_cache.begin());
  dev->mod_cache.push_back(m);
  *out = m;
  return 0;
}

// grid for a persistent kernel: enough blocks for `groups` groups, capped at
// what is co-resident (a multiple of the SM count)
template <typename Kern>
int grid_for(Dev* dev, Kern kern, size_t groups, int T, size_t smem, int* grid) {
  int per_sm = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBlockThreads, smem));
  if (per_sm < 1) return fail(IPCLB200_ERR_CUDA, "kernel does not fit an SM");
  size_t gpb = kBlockThreads / T;
  size_t need = (groups + gpb - 1) / gpb;
  size_t cap = (size_t)per_sm * dev->sms;
  if (need < 1) need = 1;
  // persistent grid: all resident blocks (work is claimed dynamically), or
  // just enough blocks when the batch is smaller than one wave
  *grid = (int)(need < cap ? need : cap);
  return 0;
}

// window-table workspace of `words` words plus the zeroed work counter behind it
int table_ws(Op& op, size_t words, uint32_t** ws, unsigned int** counter) {
  words = (words + 3) & ~(size_t)3;
  TRY(op.words(words + 4, ws));
  *counter = reinterpret_cast<unsigned int*>(*ws + words);
  CUDA_TRY(cudaMemsetAsync(*counter, 0, 16, op.s));
  return 0;
}

// host (count x words) -> device (count x L), zero padded
int upload_padded(uint32_t* d, const uint32_t* h, int words, int L, size_t count,
                  cudaStream_t s) {
  if (count == 0) return 0;
  if (words == L) {
    CUDA_TRY(cudaMemcpyAsync(d, h, count * (size_t)L * 4, cudaMemcpyHostToDevice, s));
  } else {
    CUDA_TRY(cudaMemsetAsync(d, 0, count * (size_t)L * 4, s));
    CUDA_TRY(cudaMemcpy2DAsync(d, (size_t)L * 4, h, (size_t)words * 4, (size_t)words * 4,
                               count, cudaMemcpyHostToDevice, s));
  }
  return 0;
}
int download_padded(uint32_t* h, const uint32_t* d, int words, int L, size_t count,
                    cudaStream_t s) {
  if (count == 0) return 0;
  if (words == L) {
    CUDA_TRY(cudaMemcpyAsync(h, d, count * (size_t)L * 4, cudaMemcpyDeviceToHost, s));
  } else {
    CUDA_TRY(cudaMemcpy2DAsync(h, (size_t)words * 4, d, (size_t)L * 4, (size_t)words * 4,
                               count, cudaMemcpyDeviceToHost, s));
  }
  return 0;
}

// Zero-copy: the device-visible alias of a caller's PAGE-LOCKED host buffer
// (cudaHostAlloc / cudaHostRegister memory is mapped under unified addressing),
// or nullptr for pageable memory.  Large host-pointer batches whose operands a
// kernel touches exactly once per element (plaintexts in, ciphertexts in / out)
// are read and written over PCIe by the kernel itself, element by element as
// they are claimed, instead of being staged by a copy before and after the
// launch: the transfer hides behind the arithmetic.  Measured at 65536 elements
// and a 2048-bit key (profiles/r02_zero_copy_probe.jsonl): encrypt 12.67 ->
// 11.83 ms with plaintexts and ciphertexts in place; the CRT decrypt LOSES 0.7 ms
// when it reads its ciphertexts in place (one task per thread: 16-byte pieces
// 512 bytes apart, each ciphertext once per side), so that class is off unless
// asked for.  IPCLB200_ZERO_COPY = bit mask of the classes below (default 3,
// 0 = stage everything).  The device of the shard must be current.
constexpr size_t kZeroCopyMin = 1024;
// operand classes (bits of IPCLB200_ZERO_COPY, default all)
constexpr int kZcEncryptIn = 1, kZcEncryptOut = 2, kZcDecryptIn = 4;
bool zero_copy_enabled(int which) {
  const char* e = getenv("IPCLB200_ZERO_COPY");
  return ((e && *e ? atoi(e) : (kZcEncryptIn | kZcEncryptOut)) & which) != 0;
}
uint32_t* mapped_alias(const void* h, size_t bytes, int which) {
  if (!zero_copy_enabled(which) || !h || bytes == 0) return nullptr;
  cudaPointerAttributes a0{}, a1{};
  if (cudaPointerGetAttributes(&a0, h) != cudaSuccess ||
      cudaPointerGetAttributes(&a1, static_cast<const char*>(h) + bytes - 1) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  if (a0.type != cudaMemoryTypeHost || a1.type != cudaMemoryTypeHost || !a0.devicePointer ||
      !a1.devicePointer)
    return nullptr;
  // one contiguous mapping
  if (static_cast<char*>(a1.devicePointer) - static_cast<char*>(a0.devicePointer) !=
      (ptrdiff_t)(bytes - 1))
    return nullptr;
  g.zero_copies++;
  return static_cast<uint32_t*>(a0.devicePointer);
}

// ---------------------------------------------------------------------------
// launch helpers (device pointers, the Op's stream)
// ---------------------------------------------------------------------------
int launch_modexp(Op& op, ModexpParams p, int L) {
  // a schedule needs 2^(kSchedWindow-1) table entries per group
  p.window = p.sched ? kSchedWindow - 1 : pick_window(p.exp_bits);
  int grid = 0;
#define F(K_, T_)                                                              \
  {                                                                            \
    TRY(grid_for(op.dev, modexp_kernel<K_, T_>, p.count, T_, 0, &grid));       \
    size_t groups = (size_t)grid * (kBlockThreads / T_);                       \
    TRY(table_ws(op, groups * ((size_t)L << p.window), &p.table_ws,            \
                 &p.work_counter));                                            \
    modexp_kernel<K_, T_><<<grid, kBlockThreads, 0, op.s>>>(p);                \
  }
  switch (pick_layout(p.count, L)) {
    case 1: IPCLB200_DISPATCH_WIDE(L, F) break;
    case 2: IPCLB200_DISPATCH_MID(L, F) break;
    default: IPCLB200_DISPATCH(L, F)
  }
#undef F
  g.launches++;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// modexp mod the square of dm's root in two-digit arithmetic (K1h)
constexpr size_t kHenselModexpMin = 2048;
bool modexp_hensel_applies(const DevModulus& dm, size_t count, int exp_bits) {
  if (!dm.root_words || count < kHenselModexpMin || exp_bits < 1) return false;
  const char* e = getenv("IPCLB200_NO_HENSEL_MODEXP");
  return !(e && e[0] == '1');
}

int launch_modexp_hensel(Op& op, const DevModulus& dm, const uint32_t* d_base,
                         const uint32_t* d_exp, size_t exp_stride, int exp_words, int exp_bits,
                         size_t count, uint32_t* d_out) {
  const int lh = dm.root_words;
  ModexpHenselParams p{};
  p.base = d_base;
  p.exp = d_exp;
  p.exp_stride = exp_stride;
  p.exp_words = exp_words;
  p.exp_bits = exp_bits;
  p.window = pick_window(exp_bits);
  p.blk = dm.d_root;
  p.n0inv = dm.root_n0inv;
  p.out = d_out;
  p.count = count;
#define FM(K_, T_, MINB_)                                                                \
  {                                                                                      \
    auto kern = modexp_hensel_kernel<K_, T_, MINB_, 8>;                                  \
    constexpr size_t smem = hensel_smem_bytes<K_, T_>(kBlockThreads);                    \
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                  (int)smem));                                           \
    int grid = 0;                                                                        \
    TRY(grid_for(op.dev, kern, count, T_, smem, &grid));                                 \
    const size_t groups = (size_t)grid * (kBlockThreads / T_);                           \
    TRY(table_ws(op, groups * ((size_t)(2 * lh) << p.window), &p.table_ws,               \
                 &p.work_counter));                                                      \
    kern<<<grid, kBlockThreads, smem, op.s>>>(p);                                        \
  }
  switch (lh) {
    case 32: FM(16, 2, 3) break;
    case 64: FM(16, 4, 3) break;
    case 96: FM(12, 8, 3) break;
    case 128: FM(16, 8, 3) break;
    default: return fail(IPCLB200_ERR_UNSUPPORTED, "hensel modexp: unsupported width");
  }
#undef FM
  g.launches++;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int launch_modmul(Op& op, ModmulParams p, int L) {
  int grid = 0;
  {
    uint32_t* ws = nullptr;
    TRY(table_ws(op, 0, &ws, &p.work_counter));
  }
#define F(K_, T_)                                                         \
  {                                                                       \
    TRY(grid_for(op.dev, modmul_kernel<K_, T_>, p.count, T_, 0, &grid));  \
    modmul_kernel<K_, T_><<<grid, kBlockThreads, 0, op.s>>>(p);           \
  }
  IPCLB200_DISPATCH(L, F)
#undef F
  g.launches++;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

struct CombTable {
  uint32_t* d = nullptr;
  int w = 0, windows = 0;
  bool full = false;  // built at the budget width
  bool pairs = false;  // entries are two-digit pairs (encrypt_hensel_kernel)
  size_t bytes = 0;
  cudaEvent_t ready = nullptr;  // recorded behind the build kernels
};

// device-side state of a public key on one device
struct PubDev {
  Dev* dev = nullptr;
  int dev_id = -1;  // valid after ipclb200_shutdown() has deleted *dev
  std::shared_ptr<DevModulus> msq;
  uint32_t* d_const = nullptr;  // [nR (L) | hs_m (L) | n as exponent (nl)]
  uint8_t* d_sched_n = nullptr;
  uint32_t* d_hensel = nullptr;  // n | R^2 mod n | conversion pairs (hensel_ok keys)
  CombTable cur, next;
  bool next_pending = false;
  std::vector<CombTable> retired;  // replaced tables, freed with the key
};

// dev == nullptr: the runtime was shut down (terminateContext) before the key
// died; the device memory is still valid and is freed, only the accounting is gone
void free_comb(Dev* dev, CombTable& t) {
  if (t.d) {
    cudaFree(t.d);
    if (t.full && dev) {
      std::lock_guard<std::mutex> lk(dev->mu);
      dev->comb_bytes -= std::min(dev->comb_bytes, t.bytes);
    }
  }
  if (t.ready) cudaEventDestroy(t.ready);
  t = CombTable{};
}

// the Dev a key replica was created on, or nullptr if ipclb200_shutdown() has
// deleted it since (the reference's examples call terminateContext() before
// their keys and texts go out of scope)
Dev* live_dev(Dev* dev, int id) {
  if (id < 0 || id >= kMaxDevices) return nullptr;
  return g.dev[id].load() == dev ? dev : nullptr;
}

struct PrivDev {
  Dev* dev = nullptr;
  int dev_id = -1;
  std::shared_ptr<DevModulus> mp2, mq2, mnsq;
  uint32_t* d_const = nullptr;
  uint8_t* d_sched = nullptr;
  uint32_t* d_hensel = nullptr;
#ifdef IPCLB200_EXPERIMENTS
  double* d_fp = nullptr;
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
#endif
};

}  // namespace

// ---------------------------------------------------------------------------
// key objects
// ---------------------------------------------------------------------------
struct ipclb200_pubkey {
  int nl = 0;  // words of n
  int L = 0;   // class words of n^2
  Limbs n, nsq, hs;
  bool djn = false;
  int rand_bits = 0;
  std::vector<uint32_t> h_const;
  std::vector<uint8_t> h_sched_n;
  // two-digit encrypt (mont_hensel.cuh, digits base n): n | R^2 mod n | the pairs
  // of 1, R (Montgomery windows) | the pairs of R^-1, 1 (plain window 0);
  // hensel_ok: n fills its nl words and nl is a layout of encrypt_hensel_kernel
  std::vector<uint32_t> h_hensel;
  bool hensel_ok = false;
  uint32_t n_n0inv = 0;
  // fixed-base table policy (ipclb200_pubkey_set_table_policy)
  size_t comb_max_mb = 4096;
  size_t comb_upgrade_at = 8192;
  std::mutex mu;  // device replicas, table state, counters
  size_t enc_total = 0;
  std::unique_ptr<PubDev> dev[kMaxDevices];
  ~ipclb200_pubkey() {
    for (auto& pd : dev) {
      if (!pd) continue;
      Dev* live = live_dev(pd->dev, pd->dev_id);
      cudaSetDevice(pd->dev_id);
      cudaDeviceSynchronize();
      if (pd->d_const) cudaFree(pd->d_const);
      if (pd->d_sched_n) cudaFree(pd->d_sched_n);
      if (pd->d_hensel) cudaFree(pd->d_hensel);
      free_comb(live, pd->cur);
      free_comb(live, pd->next);
      for (auto& t : pd->retired) free_comb(live, t);
      cudaGetLastError();
    }
  }
};

struct ipclb200_privkey {
  int pl = 0;    // words of p (and q)
  int L = 0;     // class words of p^2
  int Lnsq = 0;  // class words of n^2
  Limbs p, q, n, nsq, psq, qsq, lambda;
  // host images of the device blocks (replicated per device on first use)
  std::vector<uint32_t> h_const;  // p q pm1 qm1 hpR hqR pinvR | n muR | n0inv | lambda
  std::vector<uint8_t> h_sched;   // schedules of p-1, q-1, [experiments], lambda
  size_t off_sched_q = 0, off_prog_p = 0, off_prog_q = 0, off_sched_lambda = 0;
  std::vector<uint32_t> h_hensel;  // 10*pl per side, then the run schedules
  size_t off_hsched_p = 0, off_hsched_q = 0;        // sliding window
  int hensel_entries = 16;                          // odd powers per task table
  size_t off_hfixed_p = 0, off_hfixed_q = 0;        // constant schedule
  std::atomic<int> constant_schedule{0};            // ipclb200_privkey_set_schedule
  bool hensel_ok = false;
  uint32_t p_inv32 = 0, q_inv32 = 0, p_n0inv = 0, q_n0inv = 0, n_inv32 = 0, n_n0inv = 0;
  int pm1_bits = 0, qm1_bits = 0, lambda_bits = 0;
#ifdef IPCLB200_EXPERIMENTS
  uint32_t ninv_p[8] = {}, ninv_q[8] = {};
  int tile_slots = 0;
  std::vector<double> h_fp;
  uint32_t fp_n0inv_p = 0, fp_n0inv_q = 0;
  bool fp_ok = false;
#endif
  std::mutex mu;
  std::unique_ptr<PrivDev> dev[kMaxDevices];
  ~ipclb200_privkey() {
    for (auto& sd : dev) {
      if (!sd) continue;
      cudaSetDevice(sd->dev_id);
      cudaDeviceSynchronize();
      if (sd->d_const) cudaFree(sd->d_const);
      if (sd->d_sched) cudaFree(sd->d_sched);
      if (sd->d_hensel) cudaFree(sd->d_hensel);
#ifdef IPCLB200_EXPERIMENTS
      if (sd->d_fp) cudaFree(sd->d_fp);
      if (sd->aux_stream) cudaStreamDestroy(sd->aux_stream);
      if (sd->ev_fork) cudaEventDestroy(sd->ev_fork);
      if (sd->ev_join) cudaEventDestroy(sd->ev_join);
#endif
    }
  }
};

namespace {

int pub_dev(const ipclb200_pubkey* pk_c, Dev* dev, PubDev** out) {
  ipclb200_pubkey* pk = const_cast<ipclb200_pubkey*>(pk_c);
  std::lock_guard<std::mutex> lk(pk->mu);
  auto& slot = pk->dev[dev->id];
  if (!slot) {
    std::unique_ptr<PubDev> pd(new PubDev);
    pd->dev = dev;
    pd->dev_id = dev->id;
    TRY(make_modulus(dev, pk->nsq, pk->L, &pd->msq));
    CUDA_TRY(cudaSetDevice(dev->id));
    CUDA_TRY(cudaMalloc(&pd->d_const, pk->h_const.size() * sizeof(uint32_t)));
    CUDA_TRY(cudaMemcpy(pd->d_const, pk->h_const.data(),
                        pk->h_const.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    if (pk->hensel_ok) {
      CUDA_TRY(cudaMalloc(&pd->d_hensel, pk->h_hensel.size() * sizeof(uint32_t)));
      CUDA_TRY(cudaMemcpy(pd->d_hensel, pk->h_hensel.data(),
                          pk->h_hensel.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    if (!pk->h_sched_n.empty()) {
      CUDA_TRY(cudaMalloc(&pd->d_sched_n, pk->h_sched_n.size()));
      CUDA_TRY(cudaMemcpy(pd->d_sched_n, pk->h_sched_n.data(), pk->h_sched_n.size(),
                          cudaMemcpyHostToDevice));
    }
    slot = std::move(pd);
  }
  slot->dev = dev;  // a Dev re-created after ipclb200_shutdown(): same device memory
  *out = slot.get();
  return 0;
}

int priv_dev(const ipclb200_privkey* sk_c, Dev* dev, PrivDev** out) {
  ipclb200_privkey* sk = const_cast<ipclb200_privkey*>(sk_c);
  std::lock_guard<std::mutex> lk(sk->mu);
  auto& slot = sk->dev[dev->id];
  if (!slot) {
    std::unique_ptr<PrivDev> sd(new PrivDev);
    sd->dev = dev;
    sd->dev_id = dev->id;
    TRY(make_modulus(dev, sk->psq, sk->L, &sd->mp2));
    TRY(make_modulus(dev, sk->qsq, sk->L, &sd->mq2));
    TRY(make_modulus(dev, sk->nsq, sk->Lnsq, &sd->mnsq));
    CUDA_TRY(cudaSetDevice(dev->id));
    CUDA_TRY(cudaMalloc(&sd->d_const, sk->h_const.size() * sizeof(uint32_t)));
    CUDA_TRY(cudaMemcpy(sd->d_const, sk->h_const.data(),
                        sk->h_const.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&sd->d_sched, sk->h_sched.size()));
    CUDA_TRY(cudaMemcpy(sd->d_sched, sk->h_sched.data(), sk->h_sched.size(),
                        cudaMemcpyHostToDevice));
    if (sk->hensel_ok) {
      CUDA_TRY(cudaMalloc(&sd->d_hensel, sk->h_hensel.size() * sizeof(uint32_t)));
      CUDA_TRY(cudaMemcpy(sd->d_hensel, sk->h_hensel.data(),
                          sk->h_hensel.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
#ifdef IPCLB200_EXPERIMENTS
    if (sk->fp_ok) {
      CUDA_TRY(cudaMalloc(&sd->d_fp, sk->h_fp.size() * sizeof(double)));
      CUDA_TRY(cudaMemcpy(sd->d_fp, sk->h_fp.data(), sk->h_fp.size() * sizeof(double),
                          cudaMemcpyHostToDevice));
    }
#endif
    slot = std::move(sd);
  }
  slot->dev = dev;
  *out = slot.get();
  return 0;
}

// ---------------------------------------------------------------------------
// generic modexp on one device
// ---------------------------------------------------------------------------
// shared modulus: device pointers in, device pointer out
int modexp_shared_dev(Op& op, const uint32_t* d_base, const uint32_t* d_exp,
                      const Limbs& n, int L, int exp_words, int exp_bits, size_t count,
                      unsigned flags, const uint8_t* d_sched, uint32_t* d_out) {
  std::shared_ptr<DevModulus> dm;
  TRY(make_modulus(op.dev, n, L, &dm));
  CUDA_TRY(cudaSetDevice(op.dev->id));
  {
    const int eb = exp_bits > 0 ? exp_bits : exp_words * 32;
    if (!(flags & IPCLB200_SHARED_BASE) && modexp_hensel_applies(*dm, count, eb))
      return launch_modexp_hensel(op, *dm, d_base, d_exp,
                                  (flags & IPCLB200_SHARED_EXP) ? 0 : (size_t)exp_words,
                                  exp_words, eb, count, d_out);
  }
  ModexpParams p{};
  p.base = d_base;
  p.base_stride = (flags & IPCLB200_SHARED_BASE) ? 0 : L;
  p.exp = d_exp;
  p.exp_stride = (flags & IPCLB200_SHARED_EXP) ? 0 : exp_words;
  p.exp_words = exp_words;
  p.exp_bits = exp_bits > 0 ? exp_bits : exp_words * 32;
  p.n = dm->mc.n;
  p.rr = dm->mc.rr;
  p.one = dm->mc.one;
  p.n0inv = dm->d_n0inv;
  p.out = d_out;
  p.count = count;
  p.sched = d_sched;
  return launch_modexp(op, p, L);
}

// host pointers, one device, `count` elements starting at the given pointers
int modexp_host_shard(Op& op, const uint32_t* base, const uint32_t* exp, const uint32_t* mod,
                      int mod_words, int exp_words, size_t count, unsigned flags,
                      int exp_bits, const std::vector<uint8_t>* sched, uint32_t* out) {
  const int L = class_words(mod_words);
  cudaStream_t s = op.s;
  const bool sh_mod = flags & IPCLB200_SHARED_MOD;
  const bool sh_base = flags & IPCLB200_SHARED_BASE;
  const bool sh_exp = flags & IPCLB200_SHARED_EXP;
  ModexpParams p{};
  p.count = count;
  p.exp_words = exp_words;
  p.exp_bits = exp_bits;
  std::shared_ptr<DevModulus> dm;
  if (sh_mod) {
    Limbs n;
    TRY(check_modulus(mod, mod_words, &n));
    TRY(make_modulus(op.dev, n, L, &dm));
    CUDA_TRY(cudaSetDevice(op.dev->id));
    p.n = dm->mc.n;
    p.rr = dm->mc.rr;
    p.one = dm->mc.one;
    p.mod_stride = 0;
    p.n0_stride = 0;
    p.n0inv = dm->d_n0inv;
  } else {
    // heterogeneous moduli (allowed by ippMBModExp, mod_exp.cpp:479-484; no
    // in-tree caller uses it): per-element constants derived on the host
    std::vector<uint32_t> hn(count * (size_t)L), hrr(count * (size_t)L),
        hone(count * (size_t)L), hn0(count);
    HostModConst h;
    Limbs prev;
    for (size_t i = 0; i < count; i++) {
      Limbs n;
      TRY(check_modulus(mod + i * (size_t)mod_words, mod_words, &n));
      if (i == 0 || n != prev) host_mod_const(n, L, &h);
      prev = n;
      memcpy(&hn[i * (size_t)L], h.n.data(), (size_t)L * 4);
      memcpy(&hrr[i * (size_t)L], h.rr.data(), (size_t)L * 4);
      memcpy(&hone[i * (size_t)L], h.one.data(), (size_t)L * 4);
      hn0[i] = h.n0inv;
    }
    uint32_t *dn, *drr, *done, *d_n0;
    TRY(op.words(count * (size_t)L, &dn));
    TRY(op.words(count * (size_t)L, &drr));
    TRY(op.words(count * (size_t)L, &done));
    TRY(op.words(count, &d_n0));
    CUDA_TRY(cudaMemcpyAsync(dn, hn.data(), hn.size() * 4, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(drr, hrr.data(), hrr.size() * 4, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(done, hone.data(), hone.size() * 4, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_n0, hn0.data(), hn0.size() * 4, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s));  // host vectors die at scope end
    p.n = dn;
    p.rr = drr;
    p.one = done;
    p.n0inv = d_n0;
    p.mod_stride = L;
    p.n0_stride = 1;
  }
  uint32_t *d_base, *d_exp, *d_out;
  TRY(op.words((sh_base ? 1 : count) * (size_t)L, &d_base));
  TRY(op.words((sh_exp ? 1 : count) * (size_t)exp_words, &d_exp));
  TRY(op.words(count * (size_t)L, &d_out));
  TRY(upload_padded(d_base, base, mod_words, L, sh_base ? 1 : count, s));
  CUDA_TRY(cudaMemcpyAsync(d_exp, exp, (sh_exp ? 1 : count) * (size_t)exp_words * 4,
                           cudaMemcpyHostToDevice, s));
  p.base = d_base;
  p.base_stride = sh_base ? 0 : L;
  p.exp = d_exp;
  p.exp_stride = sh_exp ? 0 : exp_words;
  p.out = d_out;
  if (sh_mod && !sh_base && mod_words == L && modexp_hensel_applies(*dm, count, exp_bits)) {
    TRY(launch_modexp_hensel(op, *dm, d_base, d_exp, sh_exp ? 0 : (size_t)exp_words,
                             exp_words, exp_bits, count, d_out));
    TRY(download_padded(out, d_out, mod_words, L, count, s));
    return 0;
  }
  if (sched) {
    uint32_t* d_sc;
    TRY(op.words((sched->size() + 3) / 4, &d_sc));
    CUDA_TRY(cudaMemcpyAsync(d_sc, sched->data(), sched->size(), cudaMemcpyHostToDevice, s));
    p.sched = reinterpret_cast<const uint8_t*>(d_sc);
  }
  TRY(launch_modexp(op, p, L));
  TRY(download_padded(out, d_out, mod_words, L, count, s));
  return 0;
}

// single modexp on the primary device through the batch kernel (key-setup
// scalars: what the reference routes to ippSBModExp, ipcl/mod_exp.cpp:535-585)
int modexp_scalar(const Limbs& base, const Limbs& e, const Limbs& mod, Limbs* out) {
  int mw = (int)mod.size();
  int ew = e.empty() ? 1 : (int)e.size();
  std::vector<uint32_t> b(mw), x(ew), m(mw), r(mw);
  Limbs br = hbn::mod(base, mod);
  hbn::to_words(br, b.data(), mw);
  hbn::to_words(e, x.data(), ew);
  hbn::to_words(mod, m.data(), mw);
  Dev* dev = nullptr;
  TRY(primary_device(&dev));
  Op op;
  TRY(op.open(dev));
  TRY(modexp_host_shard(op, b.data(), x.data(), m.data(), mw, ew, 1, IPCLB200_SHARED_MOD,
                        max_bits(x.data(), ew, 1, ew), nullptr, r.data()));
  TRY(op.sync());
  *out = hbn::from_words(r.data(), mw);
  return 0;
}

// ---------------------------------------------------------------------------
// fixed-base comb table of a DJN key (K5)
// ---------------------------------------------------------------------------
constexpr int kStarterWindow = 8;
constexpr int kMaxCombWindow = 18;

size_t comb_table_words(int L, int bits, int w) {
  return (size_t)((bits + w - 1) / w) * ((size_t)L << w);
}

// Widest window (<= kMaxCombWindow bits) whose table fits the key's budget.  B200
// has 180 GB of HBM and the kernel needs one 4L-byte entry per window and
// element, so the table can be large.  1024-bit r at a 2048-bit key and the
// default 4 GB budget: w = 17 -> 61 windows x 131072 entries x 512 B = 3.9 GB and
// 60 two-digit products per encryption (w = 16: 2.1 GB and 63).
int comb_pick_window(const ipclb200_pubkey* pk, int bits) {
  size_t budget_mb = pk->comb_max_mb;
  if (const char* e = getenv("IPCLB200_COMB_MAX_MB")) budget_mb = strtoul(e, nullptr, 10);
  int w = kMaxCombWindow;
  while (w > 4 && comb_table_words(pk->L, bits, w) * 4 > (budget_mb << 20)) w--;
  if (const char* cw = getenv("IPCLB200_COMB_WINDOW")) {
    int v = atoi(cw);
    if (v >= 1 && v <= 20) w = v;
  }
  return w;
}

// enqueue the build of a table for `bits`-bit exponents on stream s.
// small = true: the starter table (8-bit windows, 17 MB at a 2048-bit key);
// false: the widest table the key's budget allows.
int build_comb(const ipclb200_pubkey* pk, PubDev* pd, int bits, bool small, cudaStream_t s,
               CombTable* out) {
  const int L = pk->L;
  int w = comb_pick_window(pk, bits);
  if (small && w > kStarterWindow && !getenv("IPCLB200_COMB_WINDOW")) w = kStarterWindow;
  CombTable t;
  // the wide table of a key whose n fills its words is stored as two-digit pairs
  // (encrypt_hensel_kernel: 5/8 of the multiplies per window); it is built in
  // full-width form first (a stream-ordered temporary) and converted
  const char* no_h = getenv("IPCLB200_NO_HENSEL_ENCRYPT");
  const bool pairs = !small && pk->hensel_ok && !(no_h && no_h[0] == '1');
  uint32_t* d_full = nullptr;
  // back off to narrower windows if the allocation does not fit
  for (;; w--) {
    cudaError_t e = cudaMalloc(&t.d, comb_table_words(L, bits, w) * sizeof(uint32_t));
    if (e == cudaSuccess && pairs) {
      e = cudaMallocAsync(&d_full, comb_table_words(L, bits, w) * sizeof(uint32_t), s);
      if (e != cudaSuccess) {
        cudaFree(t.d);
        t.d = nullptr;
      }
    }
    if (e == cudaSuccess) break;
    cudaGetLastError();
    t.d = nullptr;
    if (w <= 4) return fail(IPCLB200_ERR_CUDA, "cannot allocate the fixed-base table");
  }
  uint32_t* d_build = pairs ? d_full : t.d;
  int windows = (bits + w - 1) / w;
  if (windows < 1) windows = 1;
  CombParams cp{};
  cp.m = pd->msq->mc;
  cp.hs_m = pd->d_const + L;
  cp.comb = d_build;
  cp.w = w;
  cp.w_lo = w > 11 ? (w + 1) / 2 : w;  // two-level build for wide windows
  cp.windows = windows;
  const int nchains = cp.w_lo < w ? 2 * windows : windows;
#define F(K_, T_)                                                            \
  {                                                                          \
    comb_spine_kernel<K_, T_><<<1, 32, 0, s>>>(cp);                          \
    int gpb = kBlockThreads / T_;                                            \
    int grid = (nchains + gpb - 1) / gpb;                                    \
    comb_fill_kernel<K_, T_><<<grid, kBlockThreads, 0, s>>>(cp);             \
    if (cp.w_lo < w) {                                                       \
      int eg = 0;                                                            \
      TRY(grid_for(pd->dev, comb_expand_kernel<K_, T_>, (size_t)windows << w, \
                   T_, 0, &eg));                                             \
      comb_expand_kernel<K_, T_><<<eg, kBlockThreads, 0, s>>>(cp);           \
      g.launches++;                                                          \
    }                                                                        \
  }
  IPCLB200_DISPATCH(L, F)
#undef F
  g.launches += 2;
  CUDA_TRY(cudaGetLastError());
  if (pairs) {
    const int nl = pk->nl;
    CombPairsParams cv{};
    cv.in = d_full;
    cv.out = t.d;
    cv.count = (size_t)windows << w;
    cv.plain_entries = (size_t)1 << w;
    cv.blk = pd->d_hensel;
    cv.consts_mont = pd->d_hensel + 2 * (size_t)nl;
    cv.consts_plain = pd->d_hensel + 6 * (size_t)nl;
    cv.n0inv = pk->n_n0inv;
    uint32_t* cnt = nullptr;
    CUDA_TRY(cudaMallocAsync(&cnt, 16, s));
    CUDA_TRY(cudaMemsetAsync(cnt, 0, 16, s));
    cv.work_counter = reinterpret_cast<unsigned int*>(cnt);
#define FP(K_, T_)                                                                       \
  {                                                                                      \
    auto kern = comb_pairs_kernel<K_, T_>;                                               \
    const size_t smem = (size_t)(kBlockThreads / T_) * HMont<K_, T_, 4>::kStride * 4;    \
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                  (int)smem));                                           \
    int grid = 0;                                                                        \
    TRY(grid_for(pd->dev, kern, cv.count, T_, smem, &grid));                             \
    kern<<<grid, kBlockThreads, smem, s>>>(cv);                                          \
  }
    switch (nl) {
      case 32: FP(16, 2) break;
      case 64: FP(16, 4) break;
      case 96: FP(12, 8) break;
      case 128: FP(16, 8) break;
      default: return fail(IPCLB200_ERR_UNSUPPORTED, "hensel encrypt: unsupported key width");
    }
#undef FP
    g.launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaFreeAsync(cnt, s));
    CUDA_TRY(cudaFreeAsync(d_full, s));
    t.pairs = true;
  }
  CUDA_TRY(cudaEventCreateWithFlags(&t.ready, cudaEventDisableTiming));
  CUDA_TRY(cudaEventRecord(t.ready, s));
  t.w = w;
  t.windows = windows;
  t.full = !small;
  t.bytes = comb_table_words(L, bits, w) * sizeof(uint32_t);
  if (t.full) {
    std::lock_guard<std::mutex> lk(pd->dev->mu);
    pd->dev->comb_bytes += t.bytes;
  }
  *out = t;
  return 0;
}

// A finished wide table was converted from a full-width image of the same size
// that went back to the stream-ordered pool (whose release threshold keeps freed
// blocks cached): hand the unused part of the pool back to the driver, once per
// table, so that a 3.9 GB table does not pin 7.8 GB.
void trim_pool(Dev* dev) {
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev->id) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
  cudaGetLastError();
}

// The table this launch uses.  A key starts with the small table (built on the
// caller's stream, a few ms) and, once it has encrypted `comb_upgrade_at`
// elements, gets the wide one built on the device's side stream: encryptions
// keep using the small table until the wide one is ready -- nobody waits for
// the 35-50 ms build.  IPCLB200_COMB_SYNC=1 waits for it (benchmarks, tests).
int comb_for_launch(const ipclb200_pubkey* pk_c, PubDev* pd, int bits, size_t count,
                    cudaStream_t s, CombTable* use) {
  ipclb200_pubkey* pk = const_cast<ipclb200_pubkey*>(pk_c);
  std::lock_guard<std::mutex> lk(pk->mu);
  pk->enc_total += count;
  size_t upgrade_at = pk->comb_upgrade_at;
  if (const char* e = getenv("IPCLB200_COMB_UPGRADE")) upgrade_at = strtoul(e, nullptr, 10);
  const bool want_full = pk->enc_total >= upgrade_at;
  const char* sync_env = getenv("IPCLB200_COMB_SYNC");
  const bool sync_build = sync_env && sync_env[0] == '1';
  // adopt a finished wide table
  if (pd->next_pending) {
    cudaError_t q = sync_build ? cudaEventSynchronize(pd->next.ready)
                               : cudaEventQuery(pd->next.ready);
    if (q == cudaSuccess) {
      if (pd->cur.d) pd->retired.push_back(pd->cur);
      pd->cur = pd->next;
      pd->next = CombTable{};
      pd->next_pending = false;
      trim_pool(pd->dev);
    } else if (q != cudaErrorNotReady) {
      return fail(IPCLB200_ERR_CUDA, std::string("comb build: ") + cudaGetErrorString(q));
    }
    cudaGetLastError();
  }
  const bool covers = pd->cur.d && pd->cur.w * pd->cur.windows >= bits;
  if (!covers) {
    // first use, or a wider exponent than the table covers: build on the caller's
    // stream (ordered before the kernel that needs it)
    if (pd->cur.d) pd->retired.push_back(pd->cur);
    pd->cur = CombTable{};
    if (pd->next_pending && pd->next.w * pd->next.windows < bits) {
      cudaEventSynchronize(pd->next.ready);
      pd->retired.push_back(pd->next);
      pd->next = CombTable{};
      pd->next_pending = false;
    }
    TRY(build_comb(pk, pd, bits, true, s, &pd->cur));
  }
  if (want_full && !pd->cur.full && !pd->next_pending) {
    // device-wide budget for wide tables: past it the key stays on its small table
    size_t budget_mb = 65536;
    if (const char* e = getenv("IPCLB200_COMB_DEVICE_MB")) budget_mb = strtoul(e, nullptr, 10);
    size_t alive;
    {
      std::lock_guard<std::mutex> dl(pd->dev->mu);
      alive = pd->dev->comb_bytes;
    }
    const int tbits = std::max(bits, pk->rand_bits);
    if (alive + comb_table_words(pk->L, tbits, comb_pick_window(pk, tbits)) * 4 <=
        (budget_mb << 20)) {
      // the side stream must see the constants the caller's stream may still be
      // uploading: key blocks are uploaded synchronously, nothing to wait for
      TRY(build_comb(pk, pd, tbits, false, pd->dev->side, &pd->next));
      pd->next_pending = true;
      if (sync_build) {
        CUDA_TRY(cudaEventSynchronize(pd->next.ready));
        if (pd->cur.d) pd->retired.push_back(pd->cur);
        pd->cur = pd->next;
        pd->next = CombTable{};
        pd->next_pending = false;
        trim_pool(pd->dev);
      }
    }
  }
  // the launch stream waits for the build (no-op once it has completed)
  CUDA_TRY(cudaStreamWaitEvent(s, pd->cur.ready, 0));
  *use = pd->cur;
  return 0;
}

// DJN encrypt in two-digit arithmetic from a table of pairs (K3h)
int encrypt_hensel_launch(Op& op, const ipclb200_pubkey* pk, PubDev* pd,
                          const EncryptParams& e) {
  const int nl = pk->nl;
  EncryptHenselParams p{};
  p.pt = e.pt;
  p.pt_words = e.pt_words;
  p.r = e.r;
  p.r_words = e.r_words;
  p.blk = pd->d_hensel;
  p.n0inv = pk->n_n0inv;
  p.comb = e.comb;
  p.comb_w = e.comb_w;
  p.comb_windows = e.comb_windows;
  p.ct = e.ct;
  p.count = e.count;
  {
    uint32_t* ws = nullptr;
    TRY(table_ws(op, 0, &ws, &p.work_counter));
  }
#define FE(K_, T_, MINB_)                                                                \
  {                                                                                      \
    auto kern = encrypt_hensel_kernel<K_, T_, MINB_, 8>;                                 \
    constexpr size_t smem = encrypt_hensel_smem_bytes<K_, T_>(kBlockThreads);            \
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                  (int)smem));                                           \
    int grid = 0;                                                                        \
    TRY(grid_for(op.dev, kern, e.count, T_, smem, &grid));                               \
    kern<<<grid, kBlockThreads, smem, op.s>>>(p);                                        \
  }
  switch (nl) {
    case 32: FE(16, 2, 3) break;
    case 64: FE(16, 4, 3) break;
    case 96: FE(12, 8, 3) break;
    case 128: FE(16, 8, 3) break;
    default: return fail(IPCLB200_ERR_UNSUPPORTED, "hensel encrypt: unsupported key width");
  }
#undef FE
  g.launches++;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int encrypt_dev_impl(Op& op, const ipclb200_pubkey* pk, const uint32_t* d_pt, int pt_words,
                     const uint32_t* d_r, int r_words, int r_bits, size_t count,
                     int make_secure, uint32_t* d_ct) {
  PubDev* pd = nullptr;
  TRY(pub_dev(pk, op.dev, &pd));
  CUDA_TRY(cudaSetDevice(op.dev->id));
  const int L = pk->L;
  EncryptParams p{};
  p.pt = d_pt;
  p.pt_words = pt_words;
  p.r = d_r;
  p.r_words = r_words;
  p.r_bits = r_bits;
  p.m = pd->msq->mc;
  p.nR = pd->d_const;
  p.hs_m = pd->d_const + L;
  p.n_exp = pd->d_const + 2 * L;
  p.n_exp_words = pk->nl;
  p.ct = d_ct;
  p.count = count;
  p.window = 1;
  if (!make_secure) {
    p.mode = 0;
  } else if (pk->djn) {
    const char* no_comb = getenv("IPCLB200_NO_COMB");
    // DJN: fixed-base comb (IPCLB200_NO_COMB=1 keeps the generic windowed path)
    if (!(no_comb && no_comb[0] == '1')) {
      CombTable t;
      TRY(comb_for_launch(pk, pd, r_bits > pk->rand_bits ? r_bits : pk->rand_bits, count,
                          op.s, &t));
      p.mode = 1;
      p.comb = t.d;
      p.comb_w = t.w;
      p.comb_windows = (r_bits + t.w - 1) / t.w;
      if (p.comb_windows < 1) p.comb_windows = 1;
      if (t.pairs) return encrypt_hensel_launch(op, pk, pd, p);
    } else {
      p.mode = 2;
      p.window = pick_window(r_bits);
    }
  } else if (modexp_hensel_applies(*pd->msq, count, hbn::bitlen(pk->n)) && d_r &&
             r_words <= L) {
    // non-DJN: obf = r^n mod n^2 (ipcl/pub_key.cpp:66-80) by the two-digit ladder
    // (K1h: half the multiplies of the full-width kernel), ct = (n*m + 1) * obf.
    // r zero-extended to the 2*nl words of a residue mod n^2; n as the one exponent.
    uint32_t *d_base = nullptr, *d_obf = nullptr;
    TRY(op.words(count * (size_t)L, &d_base));
    TRY(op.words(count * (size_t)L, &d_obf));
    if (r_words < L) {
      CUDA_TRY(cudaMemsetAsync(d_base, 0, count * (size_t)L * 4, op.s));
      CUDA_TRY(cudaMemcpy2DAsync(d_base, (size_t)L * 4, d_r, (size_t)r_words * 4,
                                 (size_t)r_words * 4, count, cudaMemcpyDeviceToDevice, op.s));
    } else {
      CUDA_TRY(cudaMemcpyAsync(d_base, d_r, count * (size_t)L * 4, cudaMemcpyDeviceToDevice,
                               op.s));
    }
    TRY(launch_modexp_hensel(op, *pd->msq, d_base, p.n_exp, 0, pk->nl, hbn::bitlen(pk->n), count,
                             d_obf));
    TRY(encrypt_dev_impl(op, pk, d_pt, pt_words, nullptr, 0, 0, count, 0, d_ct));  // n*m + 1
    ModmulParams mm{};
    mm.a = d_ct;
    mm.b = d_obf;
    mm.b_stride = L;
    mm.m = pd->msq->mc;
    mm.out = d_ct;
    mm.count = count;
    return launch_modmul(op, mm, L);
  } else {
    p.mode = 3;
    const char* ns = getenv("IPCLB200_NO_SCHED");
    if (pd->d_sched_n && !(ns && ns[0] == '1')) {
      p.sched_n = pd->d_sched_n;
      p.window = kSchedWindow - 1;
    } else {
      p.window = pick_window(pk->nl * 32);
    }
  }
  int grid = 0;
#define F(K_, T_)                                                             \
  {                                                                           \
    TRY(grid_for(op.dev, encrypt_kernel<K_, T_>, count, T_, 0, &grid));       \
    size_t groups = (size_t)grid * (kBlockThreads / T_);                      \
    TRY(table_ws(op, groups * ((size_t)L << p.window), &p.table_ws,           \
                 &p.work_counter));                                           \
    encrypt_kernel<K_, T_><<<grid, kBlockThreads, 0, op.s>>>(p);              \
  }
  switch (pick_layout(count, L)) {
    case 1: IPCLB200_DISPATCH_WIDE(L, F) break;
    case 2: IPCLB200_DISPATCH_MID(L, F) break;
    default: IPCLB200_DISPATCH(L, F)
  }
#undef F
  g.launches++;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// decrypt
// ---------------------------------------------------------------------------
struct PrivPtrs {
  const uint32_t *p, *q, *hpR, *hqR, *pinvR, *n, *muR, *lambda;
  const uint8_t *sched_p, *sched_q, *sched_lambda;
};
PrivPtrs priv_ptrs(const ipclb200_privkey* sk, const PrivDev* sd) {
  const int pl = sk->pl;
  const uint32_t* d = sd->d_const;
  PrivPtrs r{};
  r.p = d;
  r.q = d + pl;
  r.hpR = d + 4 * pl;
  r.hqR = d + 5 * pl;
  r.pinvR = d + 6 * pl;
  r.n = d + 7 * pl;
  r.muR = d + 9 * pl;
  r.lambda = d + 11 * pl + 4;
  r.sched_p = sd->d_sched;
  r.sched_q = sd->d_sched + sk->off_sched_q;
  r.sched_lambda = sd->d_sched + sk->off_sched_lambda;
  return r;
}

// Schedule of the secret exponents (p-1, q-1, lambda) of a private key:
// ipclb200_privkey_set_schedule, default from IPCLB200_CONSTANT_SCHEDULE (0).
// Constant: every kernel runs a fixed-window ladder whose operation sequence
// depends on the exponent's bit length only (as mbx_exp_mb8 does); otherwise
// the host-built sliding-window schedule (about 8 % fewer products).
bool secret_schedule_is_constant(const ipclb200_privkey* sk) {
  const int v = sk->constant_schedule.load();
  if (v == 1) return true;
  if (v == 2) return false;
  const char* e = getenv("IPCLB200_CONSTANT_SCHEDULE");
  if (e && e[0] == '1') return true;
  const char* ns = getenv("IPCLB200_NO_SCHED");
  return ns && ns[0] == '1';
}

// CRT decrypt in two-digit arithmetic (decrypt_hensel_kernel + crt_combine_kernel)
int decrypt_hensel_impl(Op& op, const ipclb200_privkey* sk, const PrivDev* sd,
                        const uint32_t* d_ct, size_t count, uint32_t* d_pt) {
  const int pl = sk->pl;
  const PrivPtrs pp = priv_ptrs(sk, sd);
  uint32_t* d_mpq;
  TRY(op.words(count * (size_t)(2 * pl), &d_mpq));
  DecryptHenselParams p{};
  p.ct = d_ct;
  p.s0.blk = sd->d_hensel;
  // secret exponents p-1, q-1: sliding window (fastest) or the constant schedule
  const bool fixed = secret_schedule_is_constant(sk);
  p.s0.sched = sd->d_hensel + (fixed ? sk->off_hfixed_p : sk->off_hsched_p);
  p.s0.n0inv = sk->p_n0inv;
  p.s1.blk = sd->d_hensel + 10 * (size_t)pl;
  p.s1.sched = sd->d_hensel + (fixed ? sk->off_hfixed_q : sk->off_hsched_q);
  p.s1.n0inv = sk->q_n0inv;
  p.mpq = d_mpq;
  p.count = count;
  p.table_entries = sk->hensel_entries;
  // blocks of 128 threads per SM: 3 (12 warps) saturate the multiplier pipe
  // (measured: 1/2/3 blocks -> 139.6/97.0/92.4 ms per 65536 at a 2048-bit key)
  int want_blocks = 3;
  if (const char* e = getenv("IPCLB200_HENSEL_BLOCKS")) want_blocks = atoi(e);
  if (want_blocks < 1 || want_blocks > 4) want_blocks = 3;
  int rows = 8;
  if (const char* e = getenv("IPCLB200_HENSEL_ROWS")) rows = atoi(e);
  bool w64 = false;
  if (const char* e = getenv("IPCLB200_HENSEL_W64")) w64 = e[0] == '1';
  // Lane layout by batch size: a small batch spreads every (ciphertext, side)
  // task over more lanes (fewer limbs per lane) so that the launch fills the
  // GPU and a task's latency shrinks -- the idea of the wide layouts of the
  // generic kernels, chosen by the cost model of pick_hensel_spread().
  int spread = pick_hensel_spread(count, pl, op.dev->sms);
  {
    if (const char* e = getenv("IPCLB200_HENSEL_SPREAD")) spread = atoi(e);
    const char* w = getenv("IPCLB200_WIDE");
    if (w && w[0] == '0') spread = 0;
    if (w && w[0] == '1') spread = 2;
    if (w && w[0] == '2') spread = 1;
  }
  // the kernels live in hensel_decrypt.cu (a translation unit of their own)
  HenselDecryptPlan plan;
  {
    const cudaError_t e =
        hensel_decrypt_plan(pl, spread, rows, w64, count, op.dev->sms, want_blocks, &plan);
    if (e == cudaErrorInvalidValue)
      return fail(IPCLB200_ERR_UNSUPPORTED, "hensel: unsupported prime width");
    if (e == cudaErrorLaunchOutOfResources)
      return fail(IPCLB200_ERR_CUDA, "hensel kernel does not fit an SM");
    CUDA_TRY(e);
  }
  TRY(table_ws(op, plan.groups * 2 * pl * p.table_entries, &p.table_ws, &p.work_counter));
  plan.launch(p, plan.grid, plan.smem, op.s);
  g.launches++;
  CUDA_TRY(cudaGetLastError());
  CrtCombineParams f{};
  f.mpq = d_mpq;
  f.p = pp.p;
  f.q = pp.q;
  f.pinvR = pp.pinvR;
  f.q_n0inv = sk->q_n0inv;
  f.pl = pl;
  f.pt = d_pt;
  f.count = count;
  crt_combine_kernel<<<(unsigned)((count + 63) / 64), 64, 0, op.s>>>(f);
  g.launches++;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace

#ifdef IPCLB200_EXPERIMENTS
#include "experiments/host_exp.inc"
#endif

namespace {

// full-width CRT modexps: x[i] = (ct^(p-1) mod p^2, ct^(q-1) mod q^2), L words each
int crt_residues_impl(Op& op, const ipclb200_privkey* sk, PrivDev* sd, const uint32_t* d_ct,
                      size_t count, uint32_t* d_x) {
  const int L = sk->L;
  const PrivPtrs pp = priv_ptrs(sk, sd);
  DecryptCrtParams p{};
  p.ct = d_ct;
  p.m0 = sd->mp2->mc;
  p.m1 = sd->mq2->mc;
  p.sched0 = pp.sched_p;
  p.sched1 = pp.sched_q;
  p.x = d_x;
  p.count = count;
  p.table_entries = 1 << (kSchedWindow - 1);
#ifdef IPCLB200_EXPERIMENTS
  {
    bool handled = false;
    TRY(decrypt_crt_experiment(op, sk, sd, p, getenv("IPCLB200_DECRYPT"), &handled));
    if (handled) return 0;
  }
#endif
  int grid = 0;
#define F(K_, T_)                                                                 \
  {                                                                               \
    TRY(grid_for(op.dev, decrypt_crt_kernel<K_, T_>, 2 * count, T_, 0, &grid));   \
    size_t groups = (size_t)grid * (kBlockThreads / T_);                          \
    TRY(table_ws(op, groups * ((size_t)L * p.table_entries), &p.table_ws,         \
                 &p.work_counter));                                               \
    decrypt_crt_kernel<K_, T_><<<grid, kBlockThreads, 0, op.s>>>(p);              \
  }
  switch (pick_layout(2 * count, L)) {
    case 1: IPCLB200_DISPATCH_WIDE(L, F) break;
    case 2: IPCLB200_DISPATCH_MID(L, F) break;
    default: IPCLB200_DISPATCH(L, F)
  }
#undef F
  g.launches++;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// d_ct: count x 4*pl words when the key widths are kernel size classes (the
// *_dev contract), else count x 2L (CRT) / Lnsq (RAW) zero-padded words
int decrypt_dev_impl(Op& op, const ipclb200_privkey* sk, const uint32_t* d_ct, size_t count,
                     int use_crt, uint32_t* d_pt) {
  PrivDev* sd = nullptr;
  TRY(priv_dev(sk, op.dev, &sd));
  CUDA_TRY(cudaSetDevice(op.dev->id));
  const int pl = sk->pl;
  const PrivPtrs pp = priv_ptrs(sk, sd);
  const char* force = getenv("IPCLB200_DECRYPT");
  if (use_crt) {
    // default: two-digit (Hensel) arithmetic, half the multiplies of the
    // full-width kernel; IPCLB200_DECRYPT=int keeps the full-width kernel
    if (sk->hensel_ok && sk->L == 2 * pl && (!force || !strcmp(force, "hensel")))
      return decrypt_hensel_impl(op, sk, sd, d_ct, count, d_pt);
    const int L = sk->L;
    uint32_t* d_x;
    TRY(op.words(count * (size_t)(2 * L), &d_x));
    TRY(crt_residues_impl(op, sk, sd, d_ct, count, d_x));
    CrtFinishParams f{};
    f.x = d_x;
    f.p = pp.p;
    f.q = pp.q;
    f.hpR = pp.hpR;
    f.hqR = pp.hqR;
    f.pinvR = pp.pinvR;
    f.p_inv32 = sk->p_inv32;
    f.q_inv32 = sk->q_inv32;
    f.p_n0inv = sk->p_n0inv;
    f.q_n0inv = sk->q_n0inv;
    f.pl = pl;
    f.xl = L;
    f.pt = d_pt;
    f.count = count;
    crt_finish_kernel<<<(unsigned)((count + 63) / 64), 64, 0, op.s>>>(f);
    g.launches++;
    CUDA_TRY(cudaGetLastError());
  } else {
    const int L = sk->Lnsq;
    uint32_t* d_x;
    TRY(op.words(count * (size_t)L, &d_x));
    ModexpParams p{};
    p.base = d_ct;
    p.base_stride = L;
    p.exp = pp.lambda;
    p.exp_stride = 0;
    p.exp_words = 2 * pl;
    p.exp_bits = sk->lambda_bits;
    p.n = sd->mnsq->mc.n;
    p.rr = sd->mnsq->mc.rr;
    p.one = sd->mnsq->mc.one;
    p.n0inv = sd->mnsq->d_n0inv;
    p.mod_stride = 0;
    p.n0_stride = 0;
    p.out = d_x;
    p.count = count;
    if (!secret_schedule_is_constant(sk)) p.sched = pp.sched_lambda;
    if (modexp_hensel_applies(*sd->mnsq, count, sk->lambda_bits) && L == 4 * pl) {
      // ct^lambda mod n^2 by the two-digit ladder (fixed windows: the operation
      // sequence depends on the bit length of lambda only)
      TRY(launch_modexp_hensel(op, *sd->mnsq, d_ct, pp.lambda, 0, 2 * pl, sk->lambda_bits, count,
                               d_x));
    } else {
      TRY(launch_modexp(op, p, L));
    }
    RawFinishParams f{};
    f.x = d_x;
    f.n = pp.n;
    f.muR = pp.muR;
    f.n_inv32 = sk->n_inv32;
    f.n_n0inv = sk->n_n0inv;
    f.nl = 2 * pl;
    f.xl = L;
    f.pt = d_pt;
    f.count = count;
    raw_finish_kernel<<<(unsigned)((count + 63) / 64), 64, 0, op.s>>>(f);
    g.launches++;
    CUDA_TRY(cudaGetLastError());
  }
  return 0;
}

// ---------------------------------------------------------------------------
// NCCL (loaded on first use: the library itself does not link against it)
// ---------------------------------------------------------------------------
struct Nccl {
  typedef struct ncclComm* comm_t;
  void* handle = nullptr;
  bool tried = false;
  int (*CommInitAll)(comm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(comm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::vector<int> ids;       // device ordinals of the communicator clique
  std::vector<comm_t> comms;  // one per device, rank = index
  std::mutex mu;
};
Nccl g_nccl;
constexpr int kNcclUint32 = 3;  // ncclUint32 in nccl.h

int nccl_load() {
  if (g_nccl.tried)
    return g_nccl.handle ? 0 : fail(IPCLB200_ERR_NCCL, "libnccl.so.2 not found");
  g_nccl.tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) return fail(IPCLB200_ERR_NCCL, "libnccl.so.2 not found");
#define SYM(field, name)                                            \
  *(void**)(&g_nccl.field) = dlsym(g_nccl.handle, name);            \
  if (!g_nccl.field) {                                              \
    g_nccl.handle = nullptr;                                        \
    return fail(IPCLB200_ERR_NCCL, std::string("missing ") + name); \
  }
  SYM(CommInitAll, "ncclCommInitAll")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  return 0;
}

#define NCCL_TRY(expr)                                                       \
  do {                                                                       \
    int r_ = (expr);                                                         \
    if (r_ != 0)                                                             \
      return fail(IPCLB200_ERR_NCCL,                                         \
                  std::string(#expr) + ": " + g_nccl.GetErrorString(r_));    \
  } while (0)

// communicators over the devices of `shards` (rank = shard index)
int nccl_comms(const std::vector<Shard>& shards) {
  TRY(nccl_load());
  std::vector<int> ids;
  for (auto& sh : shards) ids.push_back(sh.dev->id);
  if (ids == g_nccl.ids) return 0;
  for (auto c : g_nccl.comms)
    if (c) g_nccl.CommDestroy(c);
  g_nccl.comms.assign(ids.size(), nullptr);
  g_nccl.ids.clear();
  NCCL_TRY(g_nccl.CommInitAll(g_nccl.comms.data(), (int)ids.size(), ids.data()));
  g_nccl.ids = ids;
  return 0;
}

}  // namespace

// a batch of big integers sharded over the active devices
struct ipclb200_batch {
  size_t count = 0;
  int words = 0;
  std::vector<Shard> shards;
  std::vector<uint32_t*> d;
  std::vector<int> dev_ids;
  // per shard: recorded behind the last operation that touched the shard (as
  // input or output), on whatever stream that was; the next operation waits for it
  std::vector<cudaEvent_t> last;
};

namespace {

// the calling thread's stream for shard i, ordered behind everything enqueued
// on the operand batches so far
int batch_stream(std::initializer_list<const ipclb200_batch*> operands, size_t i, Dev* dev,
                 cudaStream_t* out) {
  cudaStream_t s = thread_stream(dev);
  CUDA_TRY(cudaSetDevice(dev->id));
  for (const ipclb200_batch* b : operands)
    if (b) CUDA_TRY(cudaStreamWaitEvent(s, b->last[i], 0));
  *out = s;
  return 0;
}
int batch_touch(std::initializer_list<const ipclb200_batch*> operands, size_t i,
                cudaStream_t s) {
  for (const ipclb200_batch* b : operands)
    if (b) CUDA_TRY(cudaEventRecord(b->last[i], s));
  return 0;
}

bool same_plan(const ipclb200_batch* a, const ipclb200_batch* b) {
  if (a->count != b->count || a->shards.size() != b->shards.size()) return false;
  for (size_t i = 0; i < a->shards.size(); i++)
    if (a->shards[i].dev != b->shards[i].dev || a->shards[i].begin != b->shards[i].begin ||
        a->shards[i].count != b->shards[i].count)
      return false;
  return true;
}

int modmul_on(Op& op, const uint32_t* d_a, const uint32_t* d_b, const Limbs& n, int L,
              size_t count, unsigned flags, uint32_t* d_out) {
  std::shared_ptr<DevModulus> dm;
  TRY(make_modulus(op.dev, n, L, &dm));
  CUDA_TRY(cudaSetDevice(op.dev->id));
  ModmulParams p{};
  p.a = d_a;
  p.b = d_b;
  p.b_stride = (flags & IPCLB200_SHARED_B) ? 0 : L;
  p.m = dm->mc;
  p.out = d_out;
  p.count = count;
  return launch_modmul(op, p, L);
}


// r of a batch as ChaCha20 keystream (chacha20_fill_kernel, kernels.cuh K6)
int random_fill(Op& op, uint32_t* d_out, size_t count, int words, int bits, const uint32_t* key,
                const uint32_t* nonce, uint64_t first) {
  if (count == 0) return 0;
  ChachaParams p{};
  for (int i = 0; i < 8; i++) p.key[i] = key[i];
  for (int i = 0; i < 3; i++) p.nonce[i] = nonce[i];
  p.out = d_out;
  p.count = count;
  p.first = first;
  p.words = words;
  p.bits = bits;
  p.bpe = (words + 15) / 16;
  if ((first + count) * (uint64_t)p.bpe > 0xffffffffull)
    return fail(IPCLB200_ERR_BAD_ARG, "random: block counter exceeds 32 bits");
  CUDA_TRY(cudaSetDevice(op.dev->id));
  const size_t blocks = count * (size_t)p.bpe;
  chacha20_fill_kernel<<<(unsigned)((blocks + 255) / 256), 256, 0, op.s>>>(p);
  g.launches++;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int check_random_args(const char* who, int words, int bits, const uint32_t* key,
                      const uint32_t* nonce) {
  if (!key || !nonce) return fail(IPCLB200_ERR_BAD_ARG, std::string(who) + ": null key or nonce");
  if (words <= 0 || words > 2 * IPCLB200_MAX_MOD_WORDS || bits <= 0 || bits > 32 * words)
    return fail(IPCLB200_ERR_BAD_ARG, std::string(who) + ": words/bits out of range");
  return 0;
}

}  // namespace

// ===========================================================================
// exported C ABI
// ===========================================================================
extern "C" {

const char* ipclb200_version(void) { return "ipcl_b200 0.2 (sm_100a)"; }
const char* ipclb200_last_error(void) { return last_error().c_str(); }

int ipclb200_has_experiments(void) {
#ifdef IPCLB200_EXPERIMENTS
  return 1;
#else
  return 0;
#endif
}

int ipclb200_device_count(void) { return device_count_raw(); }

int ipclb200_init(int device) {
  const int ndev = device_count_raw();
  if (ndev == 0) return fail(IPCLB200_ERR_NO_DEVICE, "no CUDA device");
  {
    std::lock_guard<std::mutex> lk(g.mu);
    if (device >= 0) {
      if (device >= ndev) return fail(IPCLB200_ERR_NO_DEVICE, "no such CUDA device");
      if (g.active_set &&
          std::find(g.active.begin(), g.active.end(), device) == g.active.end())
        return fail(IPCLB200_ERR_BAD_ARG, "already initialised on another device");
      if (!g.active_set) {
        g.active = {device};
        g.active_set = true;
      }
    }
  }
  Dev* d = nullptr;
  return primary_device(&d);
}

int ipclb200_init_devices(int n_devices) {
  const int ndev = device_count_raw();
  if (ndev == 0) return fail(IPCLB200_ERR_NO_DEVICE, "no CUDA device");
  if (n_devices <= 0 || n_devices > ndev) n_devices = ndev;
  {
    std::lock_guard<std::mutex> lk(g.mu);
    g.active.clear();
    for (int i = 0; i < n_devices; i++) g.active.push_back(i);
    g.active_set = true;
  }
  DeviceGuard guard;
  std::vector<Dev*> devs;
  TRY(active_devices(&devs));
  // peer access helps NCCL's P2P transport; failures are not fatal
  for (Dev* a : devs)
    for (Dev* b : devs)
      if (a != b) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, a->id, b->id) == cudaSuccess && can) {
          cudaSetDevice(a->id);
          cudaDeviceEnablePeerAccess(b->id, 0);
        }
        cudaGetLastError();
      }
  return 0;
}

int ipclb200_active_devices(void) {
  std::lock_guard<std::mutex> lk(g.mu);
  return g.active_set ? (int)g.active.size() : 0;
}

void ipclb200_shutdown(void) {
  std::lock_guard<std::mutex> lk(g.mu);
  {
    std::lock_guard<std::mutex> nl(g_nccl.mu);
    for (auto c : g_nccl.comms)
      if (c) g_nccl.CommDestroy(c);
    g_nccl.comms.clear();
    g_nccl.ids.clear();
  }
  for (auto& slot : g.dev) {
    Dev* d = slot.load();
    if (!d) continue;
    cudaSetDevice(d->id);
    cudaDeviceSynchronize();
    d->mod_cache.clear();
    for (cudaStream_t s : d->idle) cudaStreamDestroy(s);
    if (d->stream) cudaStreamDestroy(d->stream);
    if (d->side) cudaStreamDestroy(d->side);
    delete d;
    slot.store(nullptr);
  }
  g.active.clear();
  g.active_set = false;
}

uint64_t ipclb200_launch_count(void) { return g.launches.load(); }
uint64_t ipclb200_zero_copy_count(void) { return g.zero_copies.load(); }

// ---- modexp ---------------------------------------------------------------
int ipclb200_modexp(const uint32_t* base, const uint32_t* exp, const uint32_t* mod,
                    int mod_words, int exp_words, size_t count, unsigned flags,
                    uint32_t* out) {
  if (!base || !exp || !mod || !out) return fail(IPCLB200_ERR_BAD_ARG, "modexp: null pointer");
  if (mod_words <= 0 || exp_words <= 0)
    return fail(IPCLB200_ERR_BAD_ARG, "modexp: non-positive width");
  if (mod_words > IPCLB200_MAX_MOD_WORDS)
    return fail(IPCLB200_ERR_UNSUPPORTED, "modexp: modulus wider than 8192 bits");
  if (count == 0) return 0;
  DeviceGuard guard;
  const bool sh_mod = flags & IPCLB200_SHARED_MOD;
  const bool sh_base = flags & IPCLB200_SHARED_BASE;
  const bool sh_exp = flags & IPCLB200_SHARED_EXP;
  std::vector<Shard> shards;
  TRY(plan_shards(sh_mod ? count : 1, &shards));
  if (!sh_mod) shards[0].count = count;  // heterogeneous moduli: one device
  const int exp_bits = max_bits(exp, exp_words, sh_exp ? 1 : count, exp_words);
  // one exponent for the whole batch (ct * scalar, ipcl/ciphertext.cpp:97-99;
  // Miller-Rabin rounds): sliding-window schedule instead of scanning
  std::vector<uint8_t> sched;
  if (sh_exp && count >= 64 && exp_bits > 64) {
    const char* ns = getenv("IPCLB200_NO_SCHED");
    if (!(ns && ns[0] == '1'))
      sched = build_schedule(hbn::from_words(exp, exp_words), kSchedWindow);
  }
  std::vector<Op> ops(shards.size());
  for (size_t i = 0; i < shards.size(); i++) {
    const Shard& sh = shards[i];
    TRY(ops[i].open(sh.dev));
    TRY(modexp_host_shard(ops[i], sh_base ? base : base + sh.begin * (size_t)mod_words,
                          sh_exp ? exp : exp + sh.begin * (size_t)exp_words,
                          sh_mod ? mod : mod + sh.begin * (size_t)mod_words, mod_words,
                          exp_words, sh.count, flags, exp_bits,
                          sched.empty() ? nullptr : &sched,
                          out + sh.begin * (size_t)mod_words));
  }
  for (auto& op : ops) TRY(op.sync());
  return 0;
}

int ipclb200_modexp_dev(const uint32_t* d_base, const uint32_t* d_exp, const uint32_t* h_mod,
                        int mod_words, int exp_words, int exp_bits, size_t count,
                        unsigned flags, uint32_t* d_out, void* stream) {
  if (!d_base || !d_exp || !h_mod || !d_out)
    return fail(IPCLB200_ERR_BAD_ARG, "modexp_dev: null pointer");
  if (!(flags & IPCLB200_SHARED_MOD))
    return fail(IPCLB200_ERR_UNSUPPORTED, "modexp_dev: needs IPCLB200_SHARED_MOD");
  if (mod_words <= 0 || exp_words <= 0 || class_words(mod_words) != mod_words)
    return fail(IPCLB200_ERR_UNSUPPORTED,
                "modexp_dev: mod_words must be one of 16,32,48,64,96,128,192,256");
  if (count == 0) return 0;
  Dev* dev = nullptr;
  TRY(dev_of_pointer(d_out, &dev));
  Limbs n;
  TRY(check_modulus(h_mod, mod_words, &n));
  Op op;
  TRY(op.open_on(dev, (cudaStream_t)stream));
  return modexp_shared_dev(op, d_base, d_exp, n, mod_words, exp_words, exp_bits, count, flags,
                           nullptr, d_out);
}

// ---- modmul -----------------------------------------------------------------
int ipclb200_modmul(const uint32_t* a, const uint32_t* b, const uint32_t* mod, int mod_words,
                    size_t count, unsigned flags, uint32_t* out) {
  if (!a || !b || !mod || !out) return fail(IPCLB200_ERR_BAD_ARG, "modmul: null pointer");
  if (mod_words <= 0) return fail(IPCLB200_ERR_BAD_ARG, "modmul: non-positive width");
  const int L = class_words(mod_words);
  if (!L) return fail(IPCLB200_ERR_UNSUPPORTED, "modmul: modulus wider than 8192 bits");
  if (count == 0) return 0;
  Limbs n;
  TRY(check_modulus(mod, mod_words, &n));
  DeviceGuard guard;
  const bool sh_b = flags & IPCLB200_SHARED_B;
  std::vector<Shard> shards;
  TRY(plan_shards(count, &shards));
  std::vector<Op> ops(shards.size());
  for (size_t i = 0; i < shards.size(); i++) {
    const Shard& sh = shards[i];
    Op& op = ops[i];
    TRY(op.open(sh.dev));
    uint32_t *d_a, *d_b, *d_out;
    TRY(op.words(sh.count * (size_t)L, &d_a));
    TRY(op.words((sh_b ? 1 : sh.count) * (size_t)L, &d_b));
    TRY(op.words(sh.count * (size_t)L, &d_out));
    TRY(upload_padded(d_a, a + sh.begin * (size_t)mod_words, mod_words, L, sh.count, op.s));
    TRY(upload_padded(d_b, sh_b ? b : b + sh.begin * (size_t)mod_words, mod_words, L,
                      sh_b ? 1 : sh.count, op.s));
    TRY(modmul_on(op, d_a, d_b, n, L, sh.count, flags, d_out));
    TRY(download_padded(out + sh.begin * (size_t)mod_words, d_out, mod_words, L, sh.count,
                        op.s));
  }
  for (auto& op : ops) TRY(op.sync());
  return 0;
}

int ipclb200_modmul_dev(const uint32_t* d_a, const uint32_t* d_b, const uint32_t* h_mod,
                        int mod_words, size_t count, unsigned flags, uint32_t* d_out,
                        void* stream) {
  if (!d_a || !d_b || !h_mod || !d_out)
    return fail(IPCLB200_ERR_BAD_ARG, "modmul_dev: null pointer");
  if (mod_words <= 0 || class_words(mod_words) != mod_words)
    return fail(IPCLB200_ERR_UNSUPPORTED,
                "modmul_dev: mod_words must be one of 16,32,48,64,96,128,192,256");
  if (count == 0) return 0;
  Dev* dev = nullptr;
  TRY(dev_of_pointer(d_out, &dev));
  Limbs n;
  TRY(check_modulus(h_mod, mod_words, &n));
  Op op;
  TRY(op.open_on(dev, (cudaStream_t)stream));
  return modmul_on(op, d_a, d_b, n, mod_words, count, flags, d_out);
}

// ---- public key -----------------------------------------------------------
int ipclb200_pubkey_create(const uint32_t* n, int n_words, const uint32_t* hs, int rand_bits,
                           ipclb200_pubkey** out) {
  if (!n || !out || n_words <= 0)
    return fail(IPCLB200_ERR_BAD_ARG, "pubkey_create: bad argument");
  if (2 * n_words > IPCLB200_MAX_MOD_WORDS)
    return fail(IPCLB200_ERR_UNSUPPORTED, "pubkey_create: key wider than 4096 bits");
  if (device_count_raw() == 0) return fail(IPCLB200_ERR_NO_DEVICE, "no CUDA device");
  DeviceGuard guard;
  std::unique_ptr<ipclb200_pubkey> pk(new ipclb200_pubkey);
  pk->nl = n_words;
  pk->n = hbn::from_words(n, n_words);
  if (pk->n.empty()) return fail(IPCLB200_ERR_BAD_ARG, "pubkey_create: n is zero");
  if (!(pk->n[0] & 1u)) return fail(IPCLB200_ERR_EVEN_MODULUS, "pubkey_create: n is even");
  pk->nsq = hbn::mul(pk->n, pk->n);
  pk->L = class_words(2 * n_words);
  const int L = pk->L;
  // n*R mod n^2, hs*R mod n^2, n (exponent)
  Limbs R = hbn::pow2(32u * (unsigned)L);
  Limbs nR = hbn::mod(hbn::mul(pk->n, R), pk->nsq);
  pk->h_const.assign((size_t)L + n_words + L, 0u);
  hbn::to_words(nR, pk->h_const.data(), L);
  hbn::to_words(pk->n, pk->h_const.data() + 2 * L, n_words);
  if (hs) {
    pk->djn = true;
    pk->rand_bits = rand_bits > 0 ? rand_bits : n_words * 16;
    pk->hs = hbn::mod(hbn::from_words(hs, 2 * n_words), pk->nsq);
    Limbs hsm = hbn::mod(hbn::mul(pk->hs, R), pk->nsq);
    hbn::to_words(hsm, pk->h_const.data() + L, L);
  } else {
    // non-DJN obfuscator r^n: every element has the exponent n
    pk->h_sched_n = build_schedule(pk->n, kSchedWindow);
  }
  pk->n_n0inv = hbn::neg_inv32(pk->n[0]);
  pk->hensel_ok = pk->djn && hbn::bitlen(pk->n) == 32 * n_words &&
                  (n_words == 32 || n_words == 64 || n_words == 96 || n_words == 128);
  if (pk->hensel_ok) {
    const int nl = n_words;
    const Limbs Rh = hbn::pow2(32u * (unsigned)nl);
    pk->h_hensel.assign(10 * (size_t)nl, 0u);
    uint32_t* hb = pk->h_hensel.data();
    hbn::to_words(pk->n, hb, nl);
    hbn::to_words(hbn::mod(hbn::mul(hbn::mod(Rh, pk->n), hbn::mod(Rh, pk->n)), pk->n), hb + nl, nl);
    // raw pair (x0, w) of an integer t mod n^2: t = x0 - w*n
    auto raw_pair = [&](const Limbs& t, uint32_t* dst) {
      Limbs hi, lo;
      hbn::divmod(hbn::mod(t, pk->nsq), pk->n, &hi, &lo);
      Limbs wneg = hbn::mod(hi, pk->n);
      Limbs w = hbn::is_zero(wneg) ? wneg : hbn::sub(pk->n, wneg);
      hbn::to_words(lo, dst, nl);
      hbn::to_words(w, dst + nl, nl);
    };
    Limbs one = hbn::from_u64(1), rinv;
    if (!hbn::modinv(hbn::mod(Rh, pk->nsq), pk->nsq, &rinv))
      return fail(IPCLB200_ERR_BAD_ARG, "pubkey_create: R not invertible mod n^2");
    raw_pair(one, hb + 2 * (size_t)nl);   // Montgomery windows: chunk 0 * 1
    raw_pair(Rh, hb + 4 * (size_t)nl);    //                     chunk 1 * R
    raw_pair(rinv, hb + 6 * (size_t)nl);  // plain window:       chunk 0 * R^-1
    raw_pair(one, hb + 8 * (size_t)nl);   //                     chunk 1 * 1
  }
  Dev* dev = nullptr;
  TRY(primary_device(&dev));
  PubDev* pd = nullptr;
  TRY(pub_dev(pk.get(), dev, &pd));
  *out = pk.release();
  return 0;
}

void ipclb200_pubkey_destroy(ipclb200_pubkey* pk) {
  if (!pk) return;
  DeviceGuard guard;
  delete pk;
}

int ipclb200_pubkey_set_table_policy(ipclb200_pubkey* pk, long max_table_mb,
                                     long upgrade_after) {
  if (!pk) return fail(IPCLB200_ERR_BAD_ARG, "set_table_policy: null key");
  std::lock_guard<std::mutex> lk(pk->mu);
  if (max_table_mb >= 0) pk->comb_max_mb = (size_t)max_table_mb;
  if (upgrade_after >= 0) pk->comb_upgrade_at = (size_t)upgrade_after;
  return 0;
}

int ipclb200_encrypt(const ipclb200_pubkey* pk, const uint32_t* pt, int pt_words,
                     const uint32_t* r, int r_words, size_t count, int make_secure,
                     uint32_t* ct) {
  if (!pk || !pt || !ct) return fail(IPCLB200_ERR_BAD_ARG, "encrypt: null pointer");
  if (make_secure && !r) return fail(IPCLB200_ERR_BAD_ARG, "encrypt: randoms missing");
  if (pt_words <= 0 || pt_words > pk->nl)
    return fail(IPCLB200_ERR_BAD_ARG, "encrypt: pt_words out of range");
  if (make_secure && (r_words <= 0 || r_words > 2 * pk->nl))
    return fail(IPCLB200_ERR_BAD_ARG, "encrypt: r_words out of range");
  if (count == 0) return 0;
  DeviceGuard guard;
  const int L = pk->L, CW = 2 * pk->nl;
  const int r_bits = make_secure ? max_bits(r, r_words, count, r_words) : 0;
  const char* no_comb = getenv("IPCLB200_NO_COMB");
  const bool zc = make_secure && pk->djn && L == CW && !(no_comb && no_comb[0] == '1');
  std::vector<Shard> shards;
  TRY(plan_shards(count, &shards));
  std::vector<Op> ops(shards.size());
  for (size_t i = 0; i < shards.size(); i++) {
    const Shard& sh = shards[i];
    Op& op = ops[i];
    TRY(op.open(sh.dev));
    uint32_t *d_pt = nullptr, *d_r = nullptr, *d_ct = nullptr;
    const uint32_t* h_pt = pt + sh.begin * (size_t)pt_words;
    uint32_t* h_ct = ct + sh.begin * (size_t)CW;
    // The DJN fixed-base kernels read a plaintext and write a ciphertext once per
    // element: page-locked caller buffers are used in place (mapped_alias above).
    // The randoms are read window by window and always staged.
    const bool in_place = zc && sh.count >= kZeroCopyMin;
    if (in_place) d_pt = mapped_alias(h_pt, sh.count * (size_t)pt_words * 4, kZcEncryptIn);
    if (!d_pt) {
      TRY(op.words(sh.count * (size_t)pt_words, &d_pt));
      CUDA_TRY(cudaMemcpyAsync(d_pt, h_pt, sh.count * (size_t)pt_words * 4,
                               cudaMemcpyHostToDevice, op.s));
    }
    if (make_secure) {
      TRY(op.words(sh.count * (size_t)r_words, &d_r));
      CUDA_TRY(cudaMemcpyAsync(d_r, r + sh.begin * (size_t)r_words,
                               sh.count * (size_t)r_words * 4, cudaMemcpyHostToDevice, op.s));
    }
    if (in_place) d_ct = mapped_alias(h_ct, sh.count * (size_t)CW * 4, kZcEncryptOut);
    const bool ct_in_place = d_ct != nullptr;
    if (!ct_in_place) TRY(op.words(sh.count * (size_t)L, &d_ct));
    TRY(encrypt_dev_impl(op, pk, d_pt, pt_words, d_r, r_words, r_bits, sh.count, make_secure,
                         d_ct));
    if (!ct_in_place) TRY(download_padded(h_ct, d_ct, CW, L, sh.count, op.s));
  }
  for (auto& op : ops) TRY(op.sync());
  return 0;
}

int ipclb200_encrypt_dev(const ipclb200_pubkey* pk, const uint32_t* d_pt, int pt_words,
                         const uint32_t* d_r, int r_words, size_t count, int make_secure,
                         uint32_t* d_ct, void* stream) {
  if (!pk || !d_pt || !d_ct) return fail(IPCLB200_ERR_BAD_ARG, "encrypt_dev: null pointer");
  if (make_secure && !d_r) return fail(IPCLB200_ERR_BAD_ARG, "encrypt_dev: randoms missing");
  if (pk->L != 2 * pk->nl)
    return fail(IPCLB200_ERR_UNSUPPORTED,
                "encrypt_dev: 2*n_words must be one of 16,32,48,64,96,128,192,256");
  if (pt_words <= 0 || pt_words > pk->nl)
    return fail(IPCLB200_ERR_BAD_ARG, "encrypt_dev: pt_words out of range");
  if (make_secure && (r_words <= 0 || r_words > 2 * pk->nl))
    return fail(IPCLB200_ERR_BAD_ARG, "encrypt_dev: r_words out of range");
  if (count == 0) return 0;
  Dev* dev = nullptr;
  TRY(dev_of_pointer(d_ct, &dev));
  Op op;
  TRY(op.open_on(dev, (cudaStream_t)stream));
  return encrypt_dev_impl(op, pk, d_pt, pt_words, d_r, r_words, r_words * 32, count,
                          make_secure, d_ct);
}

// ---- DJN randoms drawn on the device ----------------------------------------
int ipclb200_random_dev(uint32_t* d_out, size_t count, int words, int bits,
                        const uint32_t* key, const uint32_t* nonce, uint64_t first_element,
                        void* stream) {
  if (!d_out) return fail(IPCLB200_ERR_BAD_ARG, "random_dev: null pointer");
  TRY(check_random_args("random_dev", words, bits, key, nonce));
  if (count == 0) return 0;
  Dev* dev = nullptr;
  TRY(dev_of_pointer(d_out, &dev));
  Op op;
  TRY(op.open_on(dev, (cudaStream_t)stream));
  return random_fill(op, d_out, count, words, bits, key, nonce, first_element);
}

int ipclb200_batch_random(ipclb200_batch* out, int bits, const uint32_t* key,
                          const uint32_t* nonce) {
  if (!out) return fail(IPCLB200_ERR_BAD_ARG, "batch_random: null pointer");
  TRY(check_random_args("batch_random", out->words, bits, key, nonce));
  DeviceGuard guard;
  for (size_t i = 0; i < out->shards.size(); i++) {
    const Shard& sh = out->shards[i];
    if (sh.count == 0) continue;
    Op op;
    cudaStream_t bs = nullptr;
    TRY(batch_stream({out}, i, sh.dev, &bs));
    TRY(op.open_on(sh.dev, bs));
    TRY(random_fill(op, out->d[i], sh.count, out->words, bits, key, nonce, sh.begin));
    TRY(batch_touch({out}, i, bs));
  }
  return 0;
}

int ipclb200_encrypt_drbg(const ipclb200_pubkey* pk, const uint32_t* pt, int pt_words,
                          size_t count, const uint32_t* key, const uint32_t* nonce,
                          uint32_t* ct) {
  if (!pk || !pt || !ct) return fail(IPCLB200_ERR_BAD_ARG, "encrypt_drbg: null pointer");
  if (!pk->djn || pk->rand_bits <= 0)
    return fail(IPCLB200_ERR_UNSUPPORTED, "encrypt_drbg: needs a DJN key");
  if (pt_words <= 0 || pt_words > pk->nl)
    return fail(IPCLB200_ERR_BAD_ARG, "encrypt_drbg: pt_words out of range");
  const int r_words = (pk->rand_bits + 31) / 32;
  TRY(check_random_args("encrypt_drbg", r_words, pk->rand_bits, key, nonce));
  if (count == 0) return 0;
  DeviceGuard guard;
  const int L = pk->L, CW = 2 * pk->nl;
  std::vector<Shard> shards;
  TRY(plan_shards(count, &shards));
  std::vector<Op> ops(shards.size());
  for (size_t i = 0; i < shards.size(); i++) {
    const Shard& sh = shards[i];
    Op& op = ops[i];
    TRY(op.open(sh.dev));
    uint32_t *d_pt, *d_r, *d_ct;
    TRY(op.words(sh.count * (size_t)pt_words, &d_pt));
    TRY(op.words(sh.count * (size_t)r_words, &d_r));
    TRY(op.words(sh.count * (size_t)L, &d_ct));
    CUDA_TRY(cudaMemcpyAsync(d_pt, pt + sh.begin * (size_t)pt_words,
                             sh.count * (size_t)pt_words * 4, cudaMemcpyHostToDevice, op.s));
    TRY(random_fill(op, d_r, sh.count, r_words, pk->rand_bits, key, nonce, sh.begin));
    TRY(encrypt_dev_impl(op, pk, d_pt, pt_words, d_r, r_words, pk->rand_bits, sh.count, 1,
                         d_ct));
    TRY(download_padded(ct + sh.begin * (size_t)CW, d_ct, CW, L, sh.count, op.s));
  }
  for (auto& op : ops) TRY(op.sync());
  return 0;
}

// ---- private key ------------------------------------------------------------
int ipclb200_privkey_create(const uint32_t* p_in, const uint32_t* q_in, int p_words,
                            ipclb200_privkey** out) {
  if (!p_in || !q_in || !out || p_words <= 0)
    return fail(IPCLB200_ERR_BAD_ARG, "privkey_create: bad argument");
  if (2 * p_words > kMaxPrimeWords)
    return fail(IPCLB200_ERR_UNSUPPORTED, "privkey_create: primes wider than 2048 bits");
  if (device_count_raw() == 0) return fail(IPCLB200_ERR_NO_DEVICE, "no CUDA device");
  DeviceGuard guard;
  std::unique_ptr<ipclb200_privkey> sk(new ipclb200_privkey);
  const int pl = p_words;
  sk->pl = pl;
  Limbs p = hbn::from_words(p_in, pl), q = hbn::from_words(q_in, pl);
  if (hbn::cmp(q, p) < 0) p.swap(q);  // pri_key.cpp:19-22
  if (p.empty() || !(p[0] & 1u) || !(q[0] & 1u))
    return fail(IPCLB200_ERR_EVEN_MODULUS, "privkey_create: p and q must be odd");
  if (hbn::cmp(p, q) == 0) return fail(IPCLB200_ERR_BAD_ARG, "privkey_create: p and q are same");
  sk->p = p;
  sk->q = q;
  sk->n = hbn::mul(p, q);
  sk->nsq = hbn::mul(sk->n, sk->n);
  Limbs one = hbn::from_u64(1);
  Limbs psq = hbn::mul(p, p), qsq = hbn::mul(q, q);
  sk->psq = psq;
  sk->qsq = qsq;
  Limbs pm1 = hbn::sub(p, one), qm1 = hbn::sub(q, one);
  Limbs gg = hbn::add(sk->n, one);
  sk->L = class_words(2 * pl);
  sk->Lnsq = class_words(4 * pl);
  // computeHfun (pri_key.cpp:159-167)
  Limbs hp, hq, pinv, tmp, lq;
  TRY(modexp_scalar(hbn::mod(gg, psq), pm1, psq, &tmp));
  hbn::divmod(hbn::sub(tmp, one), p, &lq, nullptr);
  if (!hbn::modinv(lq, p, &hp))
    return fail(IPCLB200_ERR_BAD_ARG, "privkey_create: hp not invertible (p not prime?)");
  TRY(modexp_scalar(hbn::mod(gg, qsq), qm1, qsq, &tmp));
  hbn::divmod(hbn::sub(tmp, one), q, &lq, nullptr);
  if (!hbn::modinv(lq, q, &hq))
    return fail(IPCLB200_ERR_BAD_ARG, "privkey_create: hq not invertible (q not prime?)");
  if (!hbn::modinv(p, q, &pinv))
    return fail(IPCLB200_ERR_BAD_ARG, "privkey_create: p not invertible mod q");
  // lambda = lcm(p-1, q-1); mu = (L(g^lambda mod n^2))^-1 mod n (:34-37)
  Limbs gc = hbn::gcd(pm1, qm1), lam;
  hbn::divmod(hbn::mul(pm1, qm1), gc, &lam, nullptr);
  sk->lambda = lam;
  Limbs mu;
  TRY(modexp_scalar(gg, lam, sk->nsq, &tmp));
  hbn::divmod(hbn::sub(tmp, one), sk->n, &lq, nullptr);
  if (!hbn::modinv(lq, sk->n, &mu))
    return fail(IPCLB200_ERR_BAD_ARG, "privkey_create: mu not invertible");
  // constants pre-multiplied by the Montgomery radix of their modulus
  Limbs Rp = hbn::pow2(32u * (unsigned)pl), Rn = hbn::pow2(64u * (unsigned)pl);
  Limbs hpR = hbn::mod(hbn::mul(hp, Rp), p);
  Limbs hqR = hbn::mod(hbn::mul(hq, Rp), q);
  Limbs pinvR = hbn::mod(hbn::mul(pinv, Rp), q);
  Limbs muR = hbn::mod(hbn::mul(mu, Rn), sk->n);
  // device block: p q pm1 qm1 hpR hqR pinvR (pl each) | n muR (2pl each) |
  // n0inv(n^2) (1, padded to 4) | lambda (2pl)
  sk->h_const.assign(7 * (size_t)pl + 4 * (size_t)pl + 4 + 2 * (size_t)pl, 0u);
  uint32_t* b = sk->h_const.data();
  hbn::to_words(p, b, pl);
  hbn::to_words(q, b + pl, pl);
  hbn::to_words(pm1, b + 2 * pl, pl);
  hbn::to_words(qm1, b + 3 * pl, pl);
  hbn::to_words(hpR, b + 4 * pl, pl);
  hbn::to_words(hqR, b + 5 * pl, pl);
  hbn::to_words(pinvR, b + 6 * pl, pl);
  hbn::to_words(sk->n, b + 7 * pl, 2 * pl);
  hbn::to_words(muR, b + 9 * pl, 2 * pl);
  b[11 * pl] = hbn::neg_inv32(sk->nsq[0]);
  hbn::to_words(lam, b + 11 * pl + 4, 2 * pl);
  sk->p_n0inv = hbn::neg_inv32(p[0]);
  sk->q_n0inv = hbn::neg_inv32(q[0]);
  sk->n_n0inv = hbn::neg_inv32(sk->n[0]);
  sk->p_inv32 = 0u - sk->p_n0inv;
  sk->q_inv32 = 0u - sk->q_n0inv;
  sk->n_inv32 = 0u - sk->n_n0inv;
  sk->pm1_bits = hbn::bitlen(pm1);
  sk->qm1_bits = hbn::bitlen(qm1);
  sk->lambda_bits = hbn::bitlen(lam);
  {
    std::vector<uint8_t> sp = build_schedule(pm1, kSchedWindow);
    std::vector<uint8_t> sq = build_schedule(qm1, kSchedWindow);
    std::vector<uint8_t> sl = build_schedule(lam, kSchedWindow);
    std::vector<uint8_t> pp, pq;
#ifdef IPCLB200_EXPERIMENTS
    pp = build_tile_program(sp);
    pq = build_tile_program(sq);
    sk->tile_slots = sp[0] + 1;
    Limbs two256 = hbn::pow2(256), inv;
    hbn::modinv(hbn::mod(psq, two256), two256, &inv);
    hbn::to_words(hbn::sub(two256, inv), sk->ninv_p, 8);
    hbn::modinv(hbn::mod(qsq, two256), two256, &inv);
    hbn::to_words(hbn::sub(two256, inv), sk->ninv_q, 8);
#endif
    sk->h_sched = sp;
    sk->off_sched_q = sk->h_sched.size();
    sk->h_sched.insert(sk->h_sched.end(), sq.begin(), sq.end());
    sk->off_prog_p = sk->h_sched.size();
    sk->h_sched.insert(sk->h_sched.end(), pp.begin(), pp.end());
    sk->off_prog_q = sk->h_sched.size();
    sk->h_sched.insert(sk->h_sched.end(), pq.begin(), pq.end());
    sk->off_sched_lambda = sk->h_sched.size();
    sk->h_sched.insert(sk->h_sched.end(), sl.begin(), sl.end());
    // two-digit decrypt: needs primes that fill their words (so that a digit
    // < R is < 2p) and a prime width decrypt_hensel_kernel is instantiated for
    sk->hensel_ok = hbn::bitlen(p) == 32 * pl && hbn::bitlen(q) == 32 * pl &&
                    (pl == 16 || pl == 32 || pl == 48 || pl == 64);
    if (sk->hensel_ok) {
      // the two-digit kernel takes its table size from the schedule: its window
      // may differ from the generic kernels' (IPCLB200_HENSEL_WINDOW, 4..6)
      // and is chosen per key: the width with the fewest half-width products
      // (a squaring is 4, a multiply 5, 2^(w-1) - 1 multiplies build the table):
      // 5 for 512-bit exponents, 6 from 1024 bits on (measured at a 2048-bit key:
      // 80.4 -> 79.6 ms per 65536, profiles/r02_window_probe.jsonl)
      auto cost = [](const std::vector<uint32_t>& hs) {
        long c = 5L * ((long)hs[0] - 1);
        for (size_t i = 2; i < hs.size(); i++)
          c += 4L * (long)(hs[i] >> 8) + ((hs[i] & 0xffu) != 0xffu ? 5L : 0L);
        return c;
      };
      int hw = 0;
      if (const char* e = getenv("IPCLB200_HENSEL_WINDOW")) hw = atoi(e);
      if (hw < kHenselMinWindow || hw > kHenselMaxWindow) {
        long best = -1;
        for (int w = kHenselMinWindow; w <= kHenselMaxWindow; w++) {
          const long c = cost(hensel_schedule(build_schedule(pm1, w))) +
                         cost(hensel_schedule(build_schedule(qm1, w)));
          if (best < 0 || c < best) {
            best = c;
            hw = w;
          }
        }
      }
      sk->hensel_entries = std::max(16, 1 << (hw - 1));  // the constant schedule needs 16
      std::vector<uint32_t> hs_p = hensel_schedule(build_schedule(pm1, hw)),
                            hs_q = hensel_schedule(build_schedule(qm1, hw));
      std::vector<uint32_t> hf_p = hensel_schedule_fixed(pm1), hf_q = hensel_schedule_fixed(qm1);
      sk->h_hensel.assign(20 * (size_t)pl + hs_p.size() + hs_q.size() + hf_p.size() +
                              hf_q.size(), 0u);
      hensel_side_block(p, psq, hp, pl, sk->h_hensel.data());
      hensel_side_block(q, qsq, hq, pl, sk->h_hensel.data() + 10 * (size_t)pl);
      sk->off_hsched_p = 20 * (size_t)pl;
      sk->off_hsched_q = sk->off_hsched_p + hs_p.size();
      sk->off_hfixed_p = sk->off_hsched_q + hs_q.size();
      sk->off_hfixed_q = sk->off_hfixed_p + hf_p.size();
      std::copy(hs_p.begin(), hs_p.end(), sk->h_hensel.begin() + sk->off_hsched_p);
      std::copy(hs_q.begin(), hs_q.end(), sk->h_hensel.begin() + sk->off_hsched_q);
      std::copy(hf_p.begin(), hf_p.end(), sk->h_hensel.begin() + sk->off_hfixed_p);
      std::copy(hf_q.begin(), hf_q.end(), sk->h_hensel.begin() + sk->off_hfixed_q);
    }
  }
#ifdef IPCLB200_EXPERIMENTS
  privkey_experiment_constants(sk.get(), psq, qsq);
#endif
  Dev* dev = nullptr;
  TRY(primary_device(&dev));
  PrivDev* sd = nullptr;
  TRY(priv_dev(sk.get(), dev, &sd));
  *out = sk.release();
  return 0;
}

int ipclb200_privkey_set_schedule(ipclb200_privkey* sk, int constant_schedule) {
  if (!sk) return fail(IPCLB200_ERR_BAD_ARG, "set_schedule: null key");
  sk->constant_schedule.store(constant_schedule ? 1 : 2);
  return 0;
}

void ipclb200_privkey_destroy(ipclb200_privkey* sk) {
  if (!sk) return;
  DeviceGuard guard;
  delete sk;
}

int ipclb200_decrypt(const ipclb200_privkey* sk, const uint32_t* ct, size_t count,
                     int use_crt, uint32_t* pt) {
  if (!sk || !ct || !pt) return fail(IPCLB200_ERR_BAD_ARG, "decrypt: null pointer");
  if (count == 0) return 0;
  DeviceGuard guard;
  const int pl = sk->pl;
  const int CW = 4 * pl;  // caller's ciphertext words
  // the kernels read a ciphertext as 2L (CRT, L = class of p^2) or Lnsq (RAW)
  // words; zero padding keeps the value
  const int ctw = use_crt ? 2 * sk->L : sk->Lnsq;
  const bool zc = use_crt && sk->hensel_ok && sk->L == 2 * pl && CW == ctw &&
                  !getenv("IPCLB200_DECRYPT");
  std::vector<Shard> shards;
  TRY(plan_shards(count, &shards));
  std::vector<Op> ops(shards.size());
  for (size_t i = 0; i < shards.size(); i++) {
    const Shard& sh = shards[i];
    Op& op = ops[i];
    TRY(op.open(sh.dev));
    uint32_t *d_ct = nullptr, *d_pt;
    const uint32_t* h_ct = ct + sh.begin * (size_t)CW;
    // the two-digit CRT kernel reads a ciphertext once per side, in its prologue:
    // a page-locked caller buffer can be read in place (mapped_alias above; off by
    // default, the staged copy is faster)
    if (zc && sh.count >= kZeroCopyMin)
      d_ct = mapped_alias(h_ct, sh.count * (size_t)CW * 4, kZcDecryptIn);
    if (!d_ct) {
      TRY(op.words(sh.count * (size_t)ctw, &d_ct));
      TRY(upload_padded(d_ct, h_ct, CW, ctw, sh.count, op.s));
    }
    TRY(op.words(sh.count * (size_t)(2 * pl), &d_pt));
    TRY(decrypt_dev_impl(op, sk, d_ct, sh.count, use_crt, d_pt));
    CUDA_TRY(cudaMemcpyAsync(pt + sh.begin * (size_t)(2 * pl), d_pt,
                             sh.count * (size_t)(2 * pl) * 4, cudaMemcpyDeviceToHost, op.s));
  }
  for (auto& op : ops) TRY(op.sync());
  return 0;
}

int ipclb200_crt_residues(const ipclb200_privkey* sk, const uint32_t* ct, size_t count,
                          uint32_t* x, int* x_words) {
  if (!sk || !ct || !x) return fail(IPCLB200_ERR_BAD_ARG, "crt_residues: null pointer");
  if (x_words) *x_words = sk->L;
  if (count == 0) return 0;
  DeviceGuard guard;
  Dev* dev = nullptr;
  TRY(primary_device(&dev));
  PrivDev* sd = nullptr;
  TRY(priv_dev(sk, dev, &sd));
  Op op;
  TRY(op.open(dev));
  const int pl = sk->pl;
  uint32_t *d_ct, *d_x;
  TRY(op.words(count * (size_t)(2 * sk->L), &d_ct));
  TRY(op.words(count * (size_t)(2 * sk->L), &d_x));
  TRY(upload_padded(d_ct, ct, 4 * pl, 2 * sk->L, count, op.s));
  TRY(crt_residues_impl(op, sk, sd, d_ct, count, d_x));
  CUDA_TRY(cudaMemcpyAsync(x, d_x, count * (size_t)(2 * sk->L) * 4, cudaMemcpyDeviceToHost,
                           op.s));
  return op.sync();
}

int ipclb200_decrypt_dev(const ipclb200_privkey* sk, const uint32_t* d_ct, size_t count,
                         int use_crt, uint32_t* d_pt, void* stream) {
  if (!sk || !d_ct || !d_pt) return fail(IPCLB200_ERR_BAD_ARG, "decrypt_dev: null pointer");
  if (count == 0) return 0;
  const int pl = sk->pl;
  if ((use_crt && sk->L != 2 * pl) || (!use_crt && sk->Lnsq != 4 * pl))
    return fail(IPCLB200_ERR_UNSUPPORTED, "decrypt_dev: key width is not a kernel size class");
  Dev* dev = nullptr;
  TRY(dev_of_pointer(d_pt, &dev));
  Op op;
  TRY(op.open_on(dev, (cudaStream_t)stream));
  // every intermediate of this call is allocated on the caller's stream: calls
  // on different streams share nothing
  return decrypt_dev_impl(op, sk, d_ct, count, use_crt, d_pt);
}

// ---- device-resident buffers on the primary device ----------------------------
void* ipclb200_stream(void) {
  Dev* d = nullptr;
  if (primary_device(&d) != 0) return nullptr;
  return (void*)d->stream;
}

int ipclb200_dev_alloc(size_t bytes, void** d_out) {
  if (!d_out) return fail(IPCLB200_ERR_BAD_ARG, "dev_alloc: null pointer");
  Dev* d = nullptr;
  TRY(primary_device(&d));
  CUDA_TRY(cudaSetDevice(d->id));
  CUDA_TRY(cudaMallocAsync(d_out, bytes ? bytes : 4, d->stream));
  return 0;
}

int ipclb200_dev_free(void* p) {
  if (!p) return 0;
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return 0;  // the context (and its pool) is already gone
  }
  Dev* d = at.device >= 0 && at.device < kMaxDevices ? g.dev[at.device].load() : nullptr;
  if (!d) return 0;
  CUDA_TRY(cudaSetDevice(d->id));
  CUDA_TRY(cudaFreeAsync(p, d->stream));
  return 0;
}

int ipclb200_dev_upload(void* d, const void* h, size_t bytes) {
  if (!d || !h) return fail(IPCLB200_ERR_BAD_ARG, "dev_upload: null pointer");
  Dev* dev = nullptr;
  TRY(dev_of_pointer(d, &dev));
  CUDA_TRY(cudaSetDevice(dev->id));
  // pageable source: the call returns once the data is staged, so the caller
  // may reuse h; pinned source: wait for the copy
  CUDA_TRY(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, dev->stream));
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, h) == cudaSuccess && at.type == cudaMemoryTypeHost)
    CUDA_TRY(cudaStreamSynchronize(dev->stream));
  cudaGetLastError();
  return 0;
}

int ipclb200_dev_download(void* h, const void* d, size_t bytes) {
  if (!d || !h) return fail(IPCLB200_ERR_BAD_ARG, "dev_download: null pointer");
  Dev* dev = nullptr;
  TRY(dev_of_pointer(d, &dev));
  CUDA_TRY(cudaSetDevice(dev->id));
  CUDA_TRY(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, dev->stream));
  CUDA_TRY(cudaStreamSynchronize(dev->stream));
  return 0;
}

int ipclb200_dev_copy(void* d_dst, const void* d_src, size_t bytes) {
  if (!d_dst || !d_src) return fail(IPCLB200_ERR_BAD_ARG, "dev_copy: null pointer");
  Dev* dev = nullptr;
  TRY(dev_of_pointer(d_dst, &dev));
  CUDA_TRY(cudaSetDevice(dev->id));
  CUDA_TRY(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, dev->stream));
  return 0;
}

int ipclb200_sync(void) {
  if (device_count_raw() == 0) return fail(IPCLB200_ERR_NO_DEVICE, "no CUDA device");
  DeviceGuard guard;
  for (auto& slot : g.dev) {
    Dev* d = slot.load();
    if (!d) continue;
    CUDA_TRY(cudaSetDevice(d->id));
    CUDA_TRY(cudaStreamSynchronize(d->stream));
  }
  return 0;
}

int ipclb200_class_words(int words) { return words > 0 ? class_words(words) : 0; }

// page-locked host memory for callers that are not CUDA programs themselves
// (the ipcl:: layer stages vector<BigNumber> batches through it)
int ipclb200_host_alloc(size_t bytes, void** out) {
  if (!out) return fail(IPCLB200_ERR_BAD_ARG, "host_alloc: null pointer");
  if (device_count_raw() == 0) return fail(IPCLB200_ERR_NO_DEVICE, "no CUDA device");
  CUDA_TRY(cudaHostAlloc(out, bytes ? bytes : 16, cudaHostAllocPortable));
  return 0;
}

int ipclb200_host_free(void* p) {
  if (!p) return 0;
  cudaFreeHost(p);
  cudaGetLastError();
  return 0;
}

// ---- batches sharded over the active devices -------------------------------------
int ipclb200_batch_alloc(size_t count, int words, ipclb200_batch** out) {
  if (!out || words <= 0) return fail(IPCLB200_ERR_BAD_ARG, "batch_alloc: bad argument");
  DeviceGuard guard;
  std::unique_ptr<ipclb200_batch> b(new ipclb200_batch);
  b->count = count;
  b->words = words;
  TRY(plan_shards(count, &b->shards));
  for (auto& sh : b->shards) {
    uint32_t* p = nullptr;
    CUDA_TRY(cudaSetDevice(sh.dev->id));
    cudaStream_t s = thread_stream(sh.dev);
    CUDA_TRY(cudaMallocAsync(&p, std::max<size_t>(16, sh.count * (size_t)words * 4), s));
    cudaEvent_t ev = nullptr;
    CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(ev, s));
    b->d.push_back(p);
    b->dev_ids.push_back(sh.dev->id);
    b->last.push_back(ev);
  }
  *out = b.release();
  return 0;
}

void ipclb200_batch_free(ipclb200_batch* b) {
  if (!b) return;
  DeviceGuard guard;
  for (size_t i = 0; i < b->d.size(); i++) {
    Dev* d = live_dev(b->shards[i].dev, b->dev_ids[i]);
    cudaSetDevice(b->dev_ids[i]);
    if (d) {
      cudaStream_t s = thread_stream(d);
      cudaStreamWaitEvent(s, b->last[i], 0);
      cudaFreeAsync(b->d[i], s);
    } else {
      cudaFree(b->d[i]);  // after ipclb200_shutdown(): no stream left to order on
    }
    cudaEventDestroy(b->last[i]);
    cudaGetLastError();
  }
  delete b;
}

size_t ipclb200_batch_count(const ipclb200_batch* b) { return b ? b->count : 0; }
int ipclb200_batch_words(const ipclb200_batch* b) { return b ? b->words : 0; }
int ipclb200_batch_num_shards(const ipclb200_batch* b) { return b ? (int)b->shards.size() : 0; }

int ipclb200_batch_shard(const ipclb200_batch* b, int shard, int* device, void** d_ptr,
                         size_t* begin, size_t* count, void** stream) {
  if (!b || shard < 0 || shard >= (int)b->shards.size())
    return fail(IPCLB200_ERR_BAD_ARG, "batch_shard: bad argument");
  if (device) *device = b->shards[shard].dev->id;
  if (d_ptr) *d_ptr = b->d[shard];
  if (begin) *begin = b->shards[shard].begin;
  if (count) *count = b->shards[shard].count;
  if (stream) {
    // the calling thread's stream, ordered behind what was enqueued on the shard;
    // call ipclb200_batch_touch after enqueuing work of your own on it
    cudaStream_t s = nullptr;
    TRY(batch_stream({b}, (size_t)shard, b->shards[shard].dev, &s));
    *stream = (void*)s;
  }
  return 0;
}

// host (count x h_words, h_words <= words) -> shards, zero padded
int ipclb200_batch_upload(ipclb200_batch* b, const uint32_t* h, int h_words) {
  if (!b || !h || h_words <= 0 || h_words > b->words)
    return fail(IPCLB200_ERR_BAD_ARG, "batch_upload: bad argument");
  DeviceGuard guard;
  bool pinned = false;
  {
    cudaPointerAttributes at{};
    pinned = cudaPointerGetAttributes(&at, h) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
  }
  std::vector<cudaStream_t> ss(b->shards.size());
  for (size_t i = 0; i < b->shards.size(); i++) {
    const Shard& sh = b->shards[i];
    TRY(batch_stream({b}, i, sh.dev, &ss[i]));
    TRY(upload_padded(b->d[i], h + sh.begin * (size_t)h_words, h_words, b->words, sh.count,
                      ss[i]));
    TRY(batch_touch({b}, i, ss[i]));
  }
  if (pinned)
    for (size_t i = 0; i < b->shards.size(); i++) {
      CUDA_TRY(cudaSetDevice(b->shards[i].dev->id));
      CUDA_TRY(cudaStreamSynchronize(ss[i]));
    }
  return 0;
}

int ipclb200_batch_download(const ipclb200_batch* b, uint32_t* h, int h_words) {
  if (!b || !h || h_words <= 0 || h_words > b->words)
    return fail(IPCLB200_ERR_BAD_ARG, "batch_download: bad argument");
  DeviceGuard guard;
  std::vector<cudaStream_t> ss(b->shards.size());
  for (size_t i = 0; i < b->shards.size(); i++) {
    const Shard& sh = b->shards[i];
    TRY(batch_stream({b}, i, sh.dev, &ss[i]));
    TRY(download_padded(h + sh.begin * (size_t)h_words, b->d[i], h_words, b->words, sh.count,
                        ss[i]));
    TRY(batch_touch({b}, i, ss[i]));
  }
  for (size_t i = 0; i < b->shards.size(); i++) {
    CUDA_TRY(cudaSetDevice(b->shards[i].dev->id));
    CUDA_TRY(cudaStreamSynchronize(ss[i]));
  }
  return 0;
}

int ipclb200_batch_sync(const ipclb200_batch* b) {
  if (!b) return 0;
  DeviceGuard guard;
  for (size_t i = 0; i < b->shards.size(); i++) {
    CUDA_TRY(cudaSetDevice(b->shards[i].dev->id));
    CUDA_TRY(cudaEventSynchronize(b->last[i]));
  }
  return 0;
}

// after enqueuing work of your own on the stream ipclb200_batch_shard returned
int ipclb200_batch_touch(const ipclb200_batch* b, int shard, void* stream) {
  if (!b || shard < 0 || shard >= (int)b->shards.size())
    return fail(IPCLB200_ERR_BAD_ARG, "batch_touch: bad argument");
  CUDA_TRY(cudaSetDevice(b->shards[shard].dev->id));
  return batch_touch({b}, (size_t)shard, (cudaStream_t)stream);
}

// scatter: a contiguous device buffer (count x words, on the device of shard 0)
// -> the shards; gather: the reverse.  NCCL grouped send/recv over NVLink
// (NCCL has no native scatter/gather); one device: a copy.
static int batch_exchange(ipclb200_batch* b, uint32_t* d_flat, bool scatter) {
  if (!b || !d_flat) return fail(IPCLB200_ERR_BAD_ARG, "batch scatter/gather: null pointer");
  DeviceGuard guard;
  const size_t W = (size_t)b->words;
  Dev* root = b->shards[0].dev;
  {
    Dev* owner = nullptr;
    TRY(dev_of_pointer(d_flat, &owner));
    if (owner != root)
      return fail(IPCLB200_ERR_BAD_ARG,
                  "batch scatter/gather: buffer is not on the first device");
  }
  std::vector<cudaStream_t> ss(b->shards.size());
  for (size_t i = 0; i < b->shards.size(); i++)
    TRY(batch_stream({b}, i, b->shards[i].dev, &ss[i]));
  CUDA_TRY(cudaSetDevice(root->id));
  if (scatter)
    CUDA_TRY(cudaMemcpyAsync(b->d[0], d_flat, b->shards[0].count * W * 4,
                             cudaMemcpyDeviceToDevice, ss[0]));
  else
    CUDA_TRY(cudaMemcpyAsync(d_flat, b->d[0], b->shards[0].count * W * 4,
                             cudaMemcpyDeviceToDevice, ss[0]));
  if (b->shards.size() > 1) {
    std::lock_guard<std::mutex> lk(g_nccl.mu);
    TRY(nccl_comms(b->shards));
    NCCL_TRY(g_nccl.GroupStart());
    for (size_t i = 1; i < b->shards.size(); i++) {
      const Shard& sh = b->shards[i];
      uint32_t* at_root = d_flat + sh.begin * W;
      if (scatter) {
        NCCL_TRY(g_nccl.Send(at_root, sh.count * W, kNcclUint32, (int)i, g_nccl.comms[0],
                             ss[0]));
        NCCL_TRY(g_nccl.Recv(b->d[i], sh.count * W, kNcclUint32, 0, g_nccl.comms[i], ss[i]));
      } else {
        NCCL_TRY(g_nccl.Send(b->d[i], sh.count * W, kNcclUint32, 0, g_nccl.comms[i], ss[i]));
        NCCL_TRY(g_nccl.Recv(at_root, sh.count * W, kNcclUint32, (int)i, g_nccl.comms[0],
                             ss[0]));
      }
    }
    NCCL_TRY(g_nccl.GroupEnd());
  }
  for (size_t i = 0; i < b->shards.size(); i++) {
    CUDA_TRY(cudaSetDevice(b->shards[i].dev->id));
    TRY(batch_touch({b}, i, ss[i]));
  }
  return 0;
}

int ipclb200_batch_scatter(ipclb200_batch* b, const uint32_t* d_src) {
  return batch_exchange(b, const_cast<uint32_t*>(d_src), true);
}
int ipclb200_batch_gather(const ipclb200_batch* b, uint32_t* d_dst) {
  return batch_exchange(const_cast<ipclb200_batch*>(b), d_dst, false);
}

int ipclb200_encrypt_batch(const ipclb200_pubkey* pk, const ipclb200_batch* pt,
                           const ipclb200_batch* r, int r_bits, int make_secure,
                           ipclb200_batch* ct) {
  if (!pk || !pt || !ct) return fail(IPCLB200_ERR_BAD_ARG, "encrypt_batch: null pointer");
  if (make_secure && !r) return fail(IPCLB200_ERR_BAD_ARG, "encrypt_batch: randoms missing");
  if (pk->L != 2 * pk->nl || ct->words != pk->L)
    return fail(IPCLB200_ERR_UNSUPPORTED,
                "encrypt_batch: ciphertext words must be 2*n_words, a size class");
  if (pt->words > pk->nl || (make_secure && r->words > 2 * pk->nl))
    return fail(IPCLB200_ERR_BAD_ARG, "encrypt_batch: operand too wide");
  if (!same_plan(pt, ct) || (make_secure && !same_plan(r, ct)))
    return fail(IPCLB200_ERR_BAD_ARG, "encrypt_batch: batches are sharded differently");
  DeviceGuard guard;
  for (size_t i = 0; i < ct->shards.size(); i++) {
    const Shard& sh = ct->shards[i];
    if (sh.count == 0) continue;
    Op op;
    cudaStream_t bs = nullptr;
    TRY(batch_stream({pt, make_secure ? r : nullptr, ct}, i, sh.dev, &bs));
    TRY(op.open_on(sh.dev, bs));
    TRY(encrypt_dev_impl(op, pk, pt->d[i], pt->words, make_secure ? r->d[i] : nullptr,
                         make_secure ? r->words : 0,
                         make_secure ? (r_bits > 0 ? r_bits : r->words * 32) : 0, sh.count,
                         make_secure, ct->d[i]));
    TRY(batch_touch({pt, make_secure ? r : nullptr, ct}, i, bs));
  }
  return 0;
}

int ipclb200_decrypt_batch(const ipclb200_privkey* sk, const ipclb200_batch* ct, int use_crt,
                           ipclb200_batch* pt) {
  if (!sk || !ct || !pt) return fail(IPCLB200_ERR_BAD_ARG, "decrypt_batch: null pointer");
  const int pl = sk->pl;
  if ((use_crt && sk->L != 2 * pl) || (!use_crt && sk->Lnsq != 4 * pl) ||
      ct->words != 4 * pl || pt->words != 2 * pl)
    return fail(IPCLB200_ERR_UNSUPPORTED,
                "decrypt_batch: key width is not a kernel size class");
  if (!same_plan(ct, pt))
    return fail(IPCLB200_ERR_BAD_ARG, "decrypt_batch: batches are sharded differently");
  DeviceGuard guard;
  for (size_t i = 0; i < ct->shards.size(); i++) {
    const Shard& sh = ct->shards[i];
    if (sh.count == 0) continue;
    Op op;
    cudaStream_t bs = nullptr;
    TRY(batch_stream({ct, pt}, i, sh.dev, &bs));
    TRY(op.open_on(sh.dev, bs));
    TRY(decrypt_dev_impl(op, sk, ct->d[i], sh.count, use_crt, pt->d[i]));
    TRY(batch_touch({ct, pt}, i, bs));
  }
  return 0;
}

int ipclb200_modmul_batch(const ipclb200_batch* a, const ipclb200_batch* b,
                          const uint32_t* h_b_shared, const uint32_t* h_mod, int mod_words,
                          ipclb200_batch* out) {
  if (!a || !out || !h_mod || (!b && !h_b_shared))
    return fail(IPCLB200_ERR_BAD_ARG, "modmul_batch: null pointer");
  if (class_words(mod_words) != mod_words || a->words != mod_words || out->words != mod_words ||
      (b && b->words != mod_words))
    return fail(IPCLB200_ERR_UNSUPPORTED,
                "modmul_batch: words must equal mod_words, a size class");
  if (!same_plan(a, out) || (b && !same_plan(b, out)))
    return fail(IPCLB200_ERR_BAD_ARG, "modmul_batch: batches are sharded differently");
  Limbs n;
  TRY(check_modulus(h_mod, mod_words, &n));
  DeviceGuard guard;
  for (size_t i = 0; i < out->shards.size(); i++) {
    const Shard& sh = out->shards[i];
    if (sh.count == 0) continue;
    Op op;
    cudaStream_t bs = nullptr;
    TRY(batch_stream({a, b, out}, i, sh.dev, &bs));
    TRY(op.open_on(sh.dev, bs));
    const uint32_t* d_b = b ? b->d[i] : nullptr;
    if (!b) {
      uint32_t* t;
      TRY(op.words(mod_words, &t));
      CUDA_TRY(cudaMemcpyAsync(t, h_b_shared, (size_t)mod_words * 4, cudaMemcpyHostToDevice,
                               op.s));
      d_b = t;
    }
    TRY(modmul_on(op, a->d[i], d_b, n, mod_words, sh.count, b ? 0u : IPCLB200_SHARED_B,
                  out->d[i]));
    TRY(batch_touch({a, b, out}, i, bs));
  }
  return 0;
}

int ipclb200_modexp_batch(const ipclb200_batch* base, const ipclb200_batch* exp,
                          const uint32_t* h_exp_shared, int exp_words, int exp_bits,
                          const uint32_t* h_mod, int mod_words, ipclb200_batch* out) {
  if (!base || !out || !h_mod || (!exp && !h_exp_shared) || exp_words <= 0)
    return fail(IPCLB200_ERR_BAD_ARG, "modexp_batch: bad argument");
  if (class_words(mod_words) != mod_words || base->words != mod_words ||
      out->words != mod_words || (exp && exp->words != exp_words))
    return fail(IPCLB200_ERR_UNSUPPORTED,
                "modexp_batch: words must equal mod_words, a size class");
  if (!same_plan(base, out) || (exp && !same_plan(exp, out)))
    return fail(IPCLB200_ERR_BAD_ARG, "modexp_batch: batches are sharded differently");
  Limbs n;
  TRY(check_modulus(h_mod, mod_words, &n));
  DeviceGuard guard;
  std::vector<uint8_t> sched;
  if (!exp) {
    const int eb = max_bits(h_exp_shared, exp_words, 1, exp_words);
    if (exp_bits <= 0) exp_bits = eb;
    const char* ns = getenv("IPCLB200_NO_SCHED");
    if (out->count >= 64 && eb > 64 && !(ns && ns[0] == '1'))
      sched = build_schedule(hbn::from_words(h_exp_shared, exp_words), kSchedWindow);
  }
  for (size_t i = 0; i < out->shards.size(); i++) {
    const Shard& sh = out->shards[i];
    if (sh.count == 0) continue;
    Op op;
    cudaStream_t bs = nullptr;
    TRY(batch_stream({base, exp, out}, i, sh.dev, &bs));
    TRY(op.open_on(sh.dev, bs));
    const uint32_t* d_exp = exp ? exp->d[i] : nullptr;
    const uint8_t* d_sched = nullptr;
    if (!exp) {
      uint32_t* t;
      TRY(op.words(exp_words, &t));
      CUDA_TRY(cudaMemcpyAsync(t, h_exp_shared, (size_t)exp_words * 4, cudaMemcpyHostToDevice,
                               op.s));
      d_exp = t;
      if (!sched.empty()) {
        uint32_t* sc;
        TRY(op.words((sched.size() + 3) / 4, &sc));
        CUDA_TRY(cudaMemcpyAsync(sc, sched.data(), sched.size(), cudaMemcpyHostToDevice, op.s));
        d_sched = reinterpret_cast<const uint8_t*>(sc);
      }
    }
    TRY(modexp_shared_dev(op, base->d[i], d_exp, n, mod_words, exp_words, exp_bits, sh.count,
                          IPCLB200_SHARED_MOD | (exp ? 0u : IPCLB200_SHARED_EXP), d_sched,
                          out->d[i]));
    TRY(batch_touch({base, exp, out}, i, bs));
    if (!exp) CUDA_TRY(cudaStreamSynchronize(op.s));  // pageable staging of sched / h_exp
  }
  return 0;
}

// ---- measurement ----------------------------------------------------------
int ipclb200_decrypt_layout(size_t count, int p_words, int sms) {
  return pick_hensel_spread(count, p_words, sms > 0 ? sms : 148);
}

int ipclb200_int_peak(double* mac32_per_s, double* sm_clock_mhz) {
  Dev* dev = nullptr;
  TRY(primary_device(&dev));
  Op op;
  TRY(op.open(dev));
  cudaStream_t s = op.s;
  const int threads = 256, blocks = dev->sms * 8;
  uint32_t* d_out;
  TRY(op.words((size_t)threads * blocks, &d_out));
  cudaEvent_t a, b;
  CUDA_TRY(cudaEventCreate(&a));
  CUDA_TRY(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 6; rep++) {
    CUDA_TRY(cudaEventRecord(a, s));
    int_peak_kernel<<<blocks, threads, 0, s>>>(d_out, 3u + rep, 5u);
    CUDA_TRY(cudaEventRecord(b, s));
    CUDA_TRY(cudaEventSynchronize(b));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0 && ms < best) best = ms;
  }
  g.launches += 6;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  double macs = (double)threads * blocks * kPeakIters * 16.0;
  if (mac32_per_s) *mac32_per_s = macs / (best * 1e-3);
  if (sm_clock_mhz) {
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev->id);
    *sm_clock_mhz = khz / 1000.0;
  }
  return 0;
}

int ipclb200_int_peak_sustained(double seconds, double* mac32_per_s) {
  if (!mac32_per_s || seconds <= 0 || seconds > 20)
    return fail(IPCLB200_ERR_BAD_ARG, "int_peak_sustained: bad argument");
  Dev* dev = nullptr;
  TRY(primary_device(&dev));
  Op op;
  TRY(op.open(dev));
  cudaStream_t s = op.s;
  const int threads = 256, blocks = dev->sms * 8;
  uint32_t* d_out;
  TRY(op.words((size_t)threads * blocks, &d_out));
  cudaEvent_t a, b;
  CUDA_TRY(cudaEventCreate(&a));
  CUDA_TRY(cudaEventCreate(&b));
  // one launch is ~1 ms: time a back-to-back train of them.
  // IPCLB200_PEAK_PATTERN=row: distinct multiplicand registers, as in a CIOS row
  const char* pat = getenv("IPCLB200_PEAK_PATTERN");
  const bool rowpat = pat && !strcmp(pat, "row");
  int_peak_kernel<<<blocks, threads, 0, s>>>(d_out, 3u, 5u);
  CUDA_TRY(cudaStreamSynchronize(s));
  const int launches = (int)(seconds * 1000.0) + 1;
  CUDA_TRY(cudaEventRecord(a, s));
  for (int i = 0; i < launches; i++) {
    if (rowpat)
      int_peak_row_kernel<<<blocks, threads, 0, s>>>(d_out, 3u + i, 5u);
    else
      int_peak_kernel<<<blocks, threads, 0, s>>>(d_out, 3u + i, 5u);
  }
  CUDA_TRY(cudaEventRecord(b, s));
  CUDA_TRY(cudaEventSynchronize(b));
  float ms = 0;
  CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
  g.launches += launches + 1;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  *mac32_per_s = (double)threads * blocks * kPeakIters * 16.0 * launches / (ms * 1e-3);
  return 0;
}

int ipclb200_debug_montsqr(const uint32_t* a, const uint32_t* mod, size_t count,
                           uint32_t* out_sqr, uint32_t* out_mul) {
#ifndef IPCLB200_EXPERIMENTS
  (void)a; (void)mod; (void)count; (void)out_sqr; (void)out_mul;
  return fail(IPCLB200_ERR_UNSUPPORTED, "debug_montsqr: built without -DIPCLB200_EXPERIMENTS");
#else
  return debug_montsqr_experiment(a, mod, count, out_sqr, out_mul);
#endif
}

int ipclb200_pipe_mix(int mode, double* ms_out) {
#ifndef IPCLB200_EXPERIMENTS
  (void)mode; (void)ms_out;
  return fail(IPCLB200_ERR_UNSUPPORTED, "pipe_mix: built without -DIPCLB200_EXPERIMENTS");
#else
  return pipe_mix_experiment(mode, ms_out);
#endif
}

}  // extern "C"
