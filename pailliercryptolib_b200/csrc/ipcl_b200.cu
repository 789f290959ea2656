// ipcl_b200.cu -- implementation of the C ABI in include/ipcl_b200.h.
//
// Host side of the boundary: argument checking, derivation of the per-modulus
// Montgomery constants (hostbn.hpp), staging of the flat limb buffers, launch
// configuration, and the error convention.  All arithmetic on batch elements
// happens in the kernels of kernels.cuh; nothing here falls back to the CPU.
#include <cuda_runtime.h>

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
#include "hostbn.hpp"
#include "kernels.cuh"
#ifdef IPCLB200_EXPERIMENTS
#include "experiments/kernels_exp.cuh"
#endif

using namespace ipclb200;
using hbn::Limbs;

namespace {

thread_local std::string t_err;

int fail(int code, const std::string& msg) {
  t_err = msg;
  return code;
}

#define CUDA_TRY(expr)                                                   \
  do {                                                                   \
    cudaError_t e_ = (expr);                                             \
    if (e_ != cudaSuccess)                                               \
      return fail(IPCLB200_ERR_CUDA,                                     \
                  std::string(#expr) + ": " + cudaGetErrorString(e_));   \
  } while (0)

#define TRY(expr)            \
  do {                       \
    int rc_ = (expr);        \
    if (rc_ != 0) return rc_; \
  } while (0)

// ---------------------------------------------------------------------------
// size classes: modulus words -> (limbs per lane K, lanes per integer T)
// ---------------------------------------------------------------------------
const int kClasses[] = {16, 32, 48, 64, 96, 128, 192, 256};
// largest number of jobs for which the wide (4 limbs per lane) layout is used
// (measured crossovers, profiles/r01_wide_layout.md)
constexpr size_t kWideMax32 = 6144, kWideMax64 = 3072, kWideMax128 = 1536;
// ... and for the middle layout (8 limbs per lane)
constexpr size_t kMidMax32 = 12288, kMidMax64 = 6144, kMidMax128 = 3072;

int class_words(int words) {
  for (int c : kClasses)
    if (words <= c) return c;
  return 0;
}

#define IPCLB200_DISPATCH(L, F)            \
  switch (L) {                             \
    case 16:  F(8, 2); break;              \
    case 32:  F(16, 2); break;             \
    case 48:  F(12, 4); break;             \
    case 64:  F(16, 4); break;             \
    case 96:  F(12, 8); break;             \
    case 128: F(16, 8); break;             \
    case 192: F(12, 16); break;            \
    case 256: F(16, 16); break;            \
    default: return fail(IPCLB200_ERR_UNSUPPORTED, "unsupported width"); \
  }

// Small batches: the same kernels with one integer spread over four times as
// many lanes (4 limbs per lane).  A batch that cannot fill the 148 SMs is
// latency bound -- a modexp is ~1200-2500 dependent Montgomery products -- and
// the wide layout shortens every product (8 multiplies per row and lane
// instead of 32) at the price of more shuffles per multiply.
#define IPCLB200_DISPATCH_WIDE(L, F)       \
  switch (L) {                             \
    case 32:  F(4, 8); break;              \
    case 64:  F(4, 16); break;             \
    case 128: F(4, 32); break;             \
    default: return fail(IPCLB200_ERR_UNSUPPORTED, "unsupported width"); \
  }

// in between: 8 limbs per lane, twice the default number of lanes
#define IPCLB200_DISPATCH_MID(L, F)        \
  switch (L) {                             \
    case 32:  F(8, 4); break;              \
    case 64:  F(8, 8); break;              \
    case 128: F(8, 16); break;             \
    default: return fail(IPCLB200_ERR_UNSUPPORTED, "unsupported width"); \
  }

// 0 = default layout, 1 = wide (4 limbs per lane), 2 = middle (8 limbs per
// lane); tasks = independent big-integer jobs of L words in the launch
int pick_layout(size_t tasks, int L);
bool use_wide(size_t tasks, int L) { return pick_layout(tasks, L) == 1; }

int pick_layout(size_t tasks, int L) {
  if (!(L == 32 || L == 64 || L == 128)) return 0;
  const char* e = getenv("IPCLB200_WIDE");
  if (e && e[0] == '0') return 0;
  if (e && e[0] == '1') return 1;
  if (e && e[0] == '2') return 2;
  {
    size_t wide_max = 0, mid_max = 0;
    switch (L) {
      case 32: wide_max = kWideMax32; mid_max = kMidMax32; break;
      case 64: wide_max = kWideMax64; mid_max = kMidMax64; break;
      default: wide_max = kWideMax128; mid_max = kMidMax128; break;
    }
    if (const char* m = getenv("IPCLB200_WIDE_MAX")) wide_max = strtoul(m, nullptr, 10);
    if (const char* m = getenv("IPCLB200_MID_MAX")) mid_max = strtoul(m, nullptr, 10);
    if (tasks <= wide_max) return 1;
    if (tasks <= mid_max) return 2;
    return 0;
  }
}

int lanes_for(int L) {
  switch (L) {
    case 16: case 32: return 2;
    case 48: case 64: return 4;
    case 96: case 128: return 8;
    default: return 16;
  }
}

// fixed-window width minimising (2^w - 2) + bits + bits/w multiplies
int pick_window(int ebits) {
  int best = 1;
  long best_cost = -1;
  for (int w = 1; w <= kMaxWindow; w++) {
    long cost = ((1L << w) - 2) + ebits + (ebits + w - 1) / w;
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = w;
    }
  }
  return best;
}

constexpr int kSchedWindow = 5;  // 16 odd powers per table

#ifdef IPCLB200_EXPERIMENTS
// byte code for decrypt_tile_kernel (opcodes: kernels.cuh) from a
// sliding-window schedule: reduce the ciphertext, enter Montgomery form, build
// the odd powers x, x^3, ... with x^2 parked in an extra slot, run the
// schedule, leave Montgomery form
std::vector<uint8_t> build_tile_program(const std::vector<uint8_t>& sched) {
  const int nodd = sched[0];
  std::vector<uint8_t> p = {0xc0, 0xc1, 0x40};
  if (nodd > 1) {
    p.push_back(0x00);
    p.push_back((uint8_t)(0x40 + nodd));
    p.push_back(0x80);
    for (int k = 1; k < nodd; k++) {
      p.push_back((uint8_t)(0x01 + nodd));
      p.push_back((uint8_t)(0x40 + k));
    }
  }
  p.push_back((uint8_t)(0x80 + sched[1]));
  for (size_t i = 2; sched[i] != 0xff; i++)
    p.push_back(sched[i] == 0 ? 0x00 : sched[i]);  // multiply by slot op-1
  p.push_back(0xc2);
  return p;
}

#endif

// left-to-right sliding-window schedule for a fixed exponent (format: see
// modexp_sched_core in kernels.cuh).  e > 0.
std::vector<uint8_t> build_schedule(const Limbs& e, int w) {
  std::vector<uint8_t> s;
  s.push_back((uint8_t)(1 << (w - 1)));
  auto bit = [&](int i) { return i >= 0 && ((e[(size_t)i / 32] >> (i % 32)) & 1u); };
  int i = hbn::bitlen(e) - 1;
  bool first = true;
  while (i >= 0) {
    if (!bit(i)) {
      s.push_back(0);
      i--;
      continue;
    }
    int l = i - w + 1;
    if (l < 0) l = 0;
    while (!bit(l)) l++;
    unsigned v = 0;
    for (int k = i; k >= l; k--) v = (v << 1) | (bit(k) ? 1u : 0u);
    if (first) {
      s.push_back((uint8_t)((v - 1) / 2));
      first = false;
    } else {
      for (int k = i; k >= l; k--) s.push_back(0);
      s.push_back((uint8_t)((v - 1) / 2 + 1));
    }
    i = l - 1;
  }
  s.push_back(0xff);
  return s;
}

// decrypt_hensel_kernel's schedule: the byte schedule of build_schedule as 32-bit
// words [nodd, first, (run << 8 | entry)..., (run << 8 | 0xff)]: `run` squarings,
// then a multiply by odd power `entry` (0xff: none, end)
std::vector<uint32_t> hensel_schedule(const std::vector<uint8_t>& sched) {
  std::vector<uint32_t> out = {sched[0], sched[1]};
  uint32_t run = 0;
  for (size_t i = 2; sched[i] != 0xff; i++) {
    if (sched[i] == 0) {
      run++;
    } else {
      out.push_back((run << 8) | (uint32_t)(sched[i] - 1));
      run = 0;
    }
  }
  out.push_back((run << 8) | 0xffu);
  return out;
}

// per-side constants of the two-digit decrypt (10*pl words):
//   p | pairs (k0_j, kw_j) of R^(j+1) mod p^2 in Montgomery form, j = 0..3 | -hp mod p
// A pair (x0, w) stands for x0 - w*p mod p^2 (mont_hensel.cuh).
void hensel_side_block(const Limbs& p, const Limbs& psq, const Limbs& hp, int pl,
                       uint32_t* out) {
  hbn::to_words(p, out, pl);
  const Limbs R = hbn::pow2(32u * (unsigned)pl);
  Limbs t = hbn::mod(R, psq);  // R^1
  for (int j = 0; j < 4; j++) {
    t = hbn::mod(hbn::mul(t, R), psq);  // R^(j+2) = R^(j+1) in Montgomery form
    Limbs hi, lo;
    hbn::divmod(t, p, &hi, &lo);
    Limbs wneg = hbn::mod(hi, p);
    Limbs w = hbn::is_zero(wneg) ? wneg : hbn::sub(p, wneg);
    hbn::to_words(lo, out + (size_t)(1 + 2 * j) * pl, pl);
    hbn::to_words(w, out + (size_t)(2 + 2 * j) * pl, pl);
  }
  Limbs nhp = hbn::is_zero(hp) ? hp : hbn::sub(p, hp);
  hbn::to_words(nhp, out + (size_t)9 * pl, pl);
}

#ifdef IPCLB200_EXPERIMENTS
// FP64 role (mont_fp64.cuh): the 64-word class (p^2 of a 2048-bit key) as 96
// limbs of 22 bits over 4 lanes
constexpr int kFpWords = 64;
constexpr int kFpK = 24, kFpT = 4;
constexpr int kFpLimbs = kFpK * kFpT;

void fp_limbs(const Limbs& x, double* out, int L) {
  for (int g = 0; g < L; g++) {
    const unsigned off = (unsigned)kFpW * (unsigned)g;
    const size_t idx = off / 32;
    const unsigned sh = off % 32;
    uint64_t v = 0;
    if (idx < x.size()) v = x[idx];
    if (idx + 1 < x.size()) v |= (uint64_t)x[idx + 1] << 32;
    out[g] = (double)((v >> sh) & kFpMask);
  }
}

#endif

// ---------------------------------------------------------------------------
// per-modulus constants
// ---------------------------------------------------------------------------
struct DevModulus {
  int L = 0;
  Limbs n;
  uint32_t* d = nullptr;  // [n | rr | r3 | one] each L words, then n0inv
  const uint32_t* d_n0inv = nullptr;
  ModConst mc{};
  ~DevModulus() {
    if (d) cudaFree(d);
  }
};

struct HostModConst {
  std::vector<uint32_t> n, rr, r3, one;
  uint32_t n0inv;
  uint32_t small_mod;
};

void host_mod_const(const Limbs& n, int L, HostModConst* h) {
  Limbs R = hbn::pow2(32u * (unsigned)L);
  Limbs one = hbn::mod(R, n);
  Limbs rr = hbn::mod(hbn::mul(one, one), n);
  Limbs r3 = hbn::mod(hbn::mul(rr, one), n);
  h->n.resize(L);
  h->rr.resize(L);
  h->r3.resize(L);
  h->one.resize(L);
  hbn::to_words(n, h->n.data(), L);
  hbn::to_words(rr, h->rr.data(), L);
  hbn::to_words(r3, h->r3.data(), L);
  hbn::to_words(one, h->one.data(), L);
  h->n0inv = hbn::neg_inv32(n[0]);
  h->small_mod = hbn::bitlen(n) <= 32 * L - 2 ? 1u : 0u;
}

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
struct Ctx {
  std::mutex mu;
  bool ready = false;
  int device = -1;
  int sms = 0;
  cudaStream_t stream = nullptr;
  // second stream + events for the two-kernel dual-pipe decrypt
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // grow-only scratch buffers for the host-pointer entry points
  static const int kSlots = 10;
  uint32_t* scratch[kSlots] = {};
  size_t scratch_words[kSlots] = {};
  // window-table workspace per stream
  std::map<void*, std::pair<uint32_t*, size_t>> table_ws;
  // small cache of shared moduli
  std::vector<std::shared_ptr<DevModulus>> mod_cache;
  std::atomic<uint64_t> launches{0};
};

Ctx g_ctx;
int g_requested_device = -1;  // set by ipclb200_init(device >= 0)

int ensure_init_locked() {
  if (g_ctx.ready) {
    CUDA_TRY(cudaSetDevice(g_ctx.device));
    return 0;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(IPCLB200_ERR_NO_DEVICE,
                std::string("no CUDA device: ") + cudaGetErrorString(e));
  int dev = 0;
  if (g_requested_device >= 0) {
    dev = g_requested_device;
  } else if (const char* lr = getenv("LOCAL_RANK")) {
    dev = atoi(lr) % ndev;  // one process per GPU under torchrun
  } else {
    cudaGetDevice(&dev);
  }
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return fail(IPCLB200_ERR_NO_DEVICE,
                std::string("device is sm_") + std::to_string(prop.major) +
                    std::to_string(prop.minor) +
                    ", this library holds sm_100a code only");
  CUDA_TRY(cudaSetDevice(dev));
  CUDA_TRY(cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&g_ctx.aux_stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&g_ctx.ev_fork, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&g_ctx.ev_join, cudaEventDisableTiming));
  {
    // device-resident batches come from the stream-ordered pool: keep freed
    // blocks cached instead of returning them to the OS at every synchronise
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = ~(uint64_t)0;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  g_ctx.device = dev;
  g_ctx.sms = prop.multiProcessorCount;
  g_ctx.ready = true;
  return 0;
}

int scratch_get(int slot, size_t words, uint32_t** out) {
  if (g_ctx.scratch_words[slot] < words) {
    if (g_ctx.scratch[slot]) {
      // the *_dev entry points enqueue on caller streams: wait for all of them
      CUDA_TRY(cudaDeviceSynchronize());
      CUDA_TRY(cudaFree(g_ctx.scratch[slot]));
      g_ctx.scratch[slot] = nullptr;
      g_ctx.scratch_words[slot] = 0;
    }
    size_t want = words + words / 4 + 1024;
    CUDA_TRY(cudaMalloc(&g_ctx.scratch[slot], want * sizeof(uint32_t)));
    g_ctx.scratch_words[slot] = want;
  }
  *out = g_ctx.scratch[slot];
  return 0;
}

int table_ws_get(void* stream, size_t words, uint32_t** out) {
  auto& slot = g_ctx.table_ws[stream];
  if (slot.second < words) {
    if (slot.first) {
      CUDA_TRY(cudaDeviceSynchronize());
      CUDA_TRY(cudaFree(slot.first));
      slot = {nullptr, 0};
    }
    uint32_t* p = nullptr;
    CUDA_TRY(cudaMalloc(&p, words * sizeof(uint32_t)));
    slot = {p, words};
  }
  *out = slot.first;
  return 0;
}

// table workspace of `words` words plus the zeroed work counter behind it
int table_ws_with_counter(cudaStream_t s, size_t words, uint32_t** ws,
                          unsigned int** counter, int key = 0) {
  words = (words + 3) & ~(size_t)3;
  TRY(table_ws_get((void*)((uintptr_t)s + (uintptr_t)key), words + 4, ws));
  *counter = reinterpret_cast<unsigned int*>(*ws + words);
  CUDA_TRY(cudaMemsetAsync(*counter, 0, 16, s));
  return 0;
}

int make_modulus(const Limbs& n, int L, std::shared_ptr<DevModulus>* out) {
  for (auto& m : g_ctx.mod_cache)
    if (m->L == L && m->n == n) {
      *out = m;
      return 0;
    }
  auto m = std::make_shared<DevModulus>();
  HostModConst h;
  host_mod_const(n, L, &h);
  m->L = L;
  m->n = n;
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
  if (g_ctx.mod_cache.size() >= 16) g_ctx.mod_cache.erase(g_ctx.mod_cache.begin());
  g_ctx.mod_cache.push_back(m);
  *out = m;
  return 0;
}

int check_modulus(const uint32_t* mod, int words, Limbs* out) {
  Limbs n = hbn::from_words(mod, words);
  if (n.empty()) return fail(IPCLB200_ERR_BAD_ARG, "modulus is zero");
  if (!(n[0] & 1u))
    return fail(IPCLB200_ERR_EVEN_MODULUS,
                "modulus is even (Montgomery arithmetic needs an odd modulus)");
  *out = n;
  return 0;
}

// grid for a persistent kernel: enough blocks for `groups` groups, capped at
// what is co-resident (a multiple of the SM count)
template <typename Kern>
int grid_for(Kern kern, size_t groups, int T, int* grid) {
  int per_sm = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern,
                                                         kBlockThreads, 0));
  if (per_sm < 1) return fail(IPCLB200_ERR_CUDA, "kernel does not fit an SM");
  size_t gpb = kBlockThreads / T;
  size_t need = (groups + gpb - 1) / gpb;
  size_t cap = (size_t)per_sm * g_ctx.sms;
  if (need < 1) need = 1;
  // persistent grid: all resident blocks (work is claimed dynamically), or
  // just enough blocks when the batch is smaller than one wave
  *grid = (int)(need < cap ? need : cap);
  return 0;
}

int max_bits(const uint32_t* v, int words, size_t count, size_t stride) {
  int best = 0;
  for (size_t i = 0; i < count; i++) {
    const uint32_t* e = v + i * stride;
    for (int w = words - 1; w >= 0; w--) {
      if (e[w]) {
        int b = w * 32 + 32 - __builtin_clz(e[w]);
        if (b > best) best = b;
        break;
      }
      if ((w + 1) * 32 <= best) break;
    }
  }
  return best;
}

// host (count x words) -> device (count x L), zero padded
int upload_padded(uint32_t* d, const uint32_t* h, int words, int L,
                  size_t count, cudaStream_t s) {
  if (words == L) {
    CUDA_TRY(cudaMemcpyAsync(d, h, count * (size_t)L * 4, cudaMemcpyHostToDevice, s));
  } else {
    CUDA_TRY(cudaMemsetAsync(d, 0, count * (size_t)L * 4, s));
    CUDA_TRY(cudaMemcpy2DAsync(d, (size_t)L * 4, h, (size_t)words * 4,
                               (size_t)words * 4, count, cudaMemcpyHostToDevice, s));
  }
  return 0;
}
int download_padded(uint32_t* h, const uint32_t* d, int words, int L,
                    size_t count, cudaStream_t s) {
  if (words == L) {
    CUDA_TRY(cudaMemcpyAsync(h, d, count * (size_t)L * 4, cudaMemcpyDeviceToHost, s));
  } else {
    CUDA_TRY(cudaMemcpy2DAsync(h, (size_t)words * 4, d, (size_t)L * 4,
                               (size_t)words * 4, count, cudaMemcpyDeviceToHost, s));
  }
  return 0;
}

// ---------------------------------------------------------------------------
// launch helpers (device pointers, any stream)
// ---------------------------------------------------------------------------
int launch_modexp(ModexpParams p, int L, cudaStream_t s) {
  // a schedule needs 2^(kSchedWindow-1) table entries per group
  p.window = p.sched ? kSchedWindow - 1 : pick_window(p.exp_bits);
  const int T = lanes_for(L);
  int grid = 0;
#define F(K_, T_)                                                         \
  {                                                                       \
    TRY(grid_for(modexp_kernel<K_, T_>, p.count, T_, &grid));             \
    size_t groups = (size_t)grid * (kBlockThreads / T_);                  \
    TRY(table_ws_with_counter(s, groups * ((size_t)L << p.window),        \
                              &p.table_ws, &p.work_counter));             \
    modexp_kernel<K_, T_><<<grid, kBlockThreads, 0, s>>>(p);              \
  }
  switch (pick_layout(p.count, L)) {
    case 1: IPCLB200_DISPATCH_WIDE(L, F) break;
    case 2: IPCLB200_DISPATCH_MID(L, F) break;
    default: IPCLB200_DISPATCH(L, F)
  }
#undef F
  (void)T;
  g_ctx.launches++;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int launch_modmul(ModmulParams p, int L, cudaStream_t s) {
  int grid = 0;
#define F(K_, T_)                                               \
  {                                                             \
    TRY(grid_for(modmul_kernel<K_, T_>, p.count, T_, &grid));   \
    modmul_kernel<K_, T_><<<grid, kBlockThreads, 0, s>>>(p);    \
  }
  IPCLB200_DISPATCH(L, F)
#undef F
  g_ctx.launches++;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

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
  std::shared_ptr<DevModulus> msq;
  uint32_t* d_const = nullptr;  // [nR (L) | hs_m (L) | n as exponent (nl)]
  // lazily built fixed-base comb table for hs
  mutable uint32_t* d_comb = nullptr;
  mutable int comb_w = 0;
  mutable int comb_windows = 0;
  mutable size_t enc_total = 0;  // elements encrypted with this key so far
  mutable bool comb_full = false;  // the table was built at the budget width
  uint8_t* d_sched_n = nullptr;  // sliding-window schedule of the exponent n
  ~ipclb200_pubkey() {
    if (d_const) cudaFree(d_const);
    if (d_comb) cudaFree(d_comb);
    if (d_sched_n) cudaFree(d_sched_n);
  }
};

struct ipclb200_privkey {
  int pl = 0;  // words of p (and q)
  int L = 0;   // class words of p^2 (== 2*pl required)
  Limbs p, q, n, nsq, lambda;
  std::shared_ptr<DevModulus> mp2, mq2, mnsq;
  uint32_t* d_const = nullptr;
  // offsets into d_const (words)
  const uint32_t *d_p = nullptr, *d_q = nullptr, *d_pm1 = nullptr,
                 *d_qm1 = nullptr, *d_hpR = nullptr, *d_hqR = nullptr,
                 *d_pinvR = nullptr, *d_n = nullptr, *d_muR = nullptr,
                 *d_lambda = nullptr;
  uint32_t p_inv32 = 0, q_inv32 = 0, p_n0inv = 0, q_n0inv = 0, n_inv32 = 0,
           n_n0inv = 0;
  int pm1_bits = 0, qm1_bits = 0, lambda_bits = 0;
  // sliding-window schedules of the shared exponents p-1, q-1
  uint8_t* d_sched = nullptr;
  const uint8_t *d_sched_p = nullptr, *d_sched_q = nullptr,
                *d_sched_lambda = nullptr;
  // two-digit (Hensel) decrypt, mont_hensel.cuh: per side p | K_0..K_3 | -hp
  // (10*pl words) and the run-length schedules; hensel_ok: p and q fill their
  // pl words and pl is a layout of decrypt_hensel_kernel
  uint32_t* d_hensel = nullptr;
  const uint32_t *d_hblk_p = nullptr, *d_hblk_q = nullptr, *d_hsched_p = nullptr,
                 *d_hsched_q = nullptr;
  bool hensel_ok = false;
#ifdef IPCLB200_EXPERIMENTS
  // byte-code programs of the thread-per-integer kernel and -N^-1 mod 2^256
  const uint8_t *d_prog_p = nullptr, *d_prog_q = nullptr;
  uint32_t ninv_p[8] = {}, ninv_q[8] = {};
  int tile_slots = 0;
  // FP64-pipe constants (mont_fp64.cuh), only for the 64-word class of p^2
  double* d_fp = nullptr;
  FpModConst fp0{}, fp1{};
  bool fp_ok = false;
#endif
  ~ipclb200_privkey() {
    if (d_const) cudaFree(d_const);
    if (d_sched) cudaFree(d_sched);
    if (d_hensel) cudaFree(d_hensel);
#ifdef IPCLB200_EXPERIMENTS
    if (d_fp) cudaFree(d_fp);
#endif
  }
};

namespace {

// single modexp on the device through the batch kernel (key-setup scalars:
// what the reference routes to ippSBModExp, ipcl/mod_exp.cpp:535-585)
int modexp_scalar(const Limbs& base, const Limbs& e, const Limbs& mod,
                  Limbs* out);

int modexp_host_impl(const uint32_t* base, const uint32_t* exp,
                     const uint32_t* mod, int mod_words, int exp_words,
                     size_t count, unsigned flags, uint32_t* out) {
  const int L = class_words(mod_words);
  if (!L) return fail(IPCLB200_ERR_UNSUPPORTED, "modulus wider than 8192 bits");
  cudaStream_t s = g_ctx.stream;
  const bool sh_mod = flags & IPCLB200_SHARED_MOD;
  const bool sh_base = flags & IPCLB200_SHARED_BASE;
  const bool sh_exp = flags & IPCLB200_SHARED_EXP;
  ModexpParams p{};
  p.count = count;
  p.exp_words = exp_words;
  p.exp_bits = max_bits(exp, exp_words, sh_exp ? 1 : count, exp_words);
  std::shared_ptr<DevModulus> dm;
  uint32_t* d_n0 = nullptr;
  if (sh_mod) {
    Limbs n;
    TRY(check_modulus(mod, mod_words, &n));
    TRY(make_modulus(n, L, &dm));
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
    uint32_t *dn, *drr, *done;
    TRY(scratch_get(3, count * (size_t)L, &dn));
    TRY(scratch_get(4, count * (size_t)L, &drr));
    TRY(scratch_get(6, count * (size_t)L, &done));
    TRY(scratch_get(5, count, &d_n0));
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
  TRY(scratch_get(0, (sh_base ? 1 : count) * (size_t)L, &d_base));
  TRY(scratch_get(1, (sh_exp ? 1 : count) * (size_t)exp_words, &d_exp));
  TRY(scratch_get(2, count * (size_t)L, &d_out));
  TRY(upload_padded(d_base, base, mod_words, L, sh_base ? 1 : count, s));
  CUDA_TRY(cudaMemcpyAsync(d_exp, exp, (sh_exp ? 1 : count) * (size_t)exp_words * 4,
                           cudaMemcpyHostToDevice, s));
  p.base = d_base;
  p.base_stride = sh_base ? 0 : L;
  p.exp = d_exp;
  p.exp_stride = sh_exp ? 0 : exp_words;
  p.out = d_out;
  if (sh_exp && count >= 64 && p.exp_bits > 64) {
    // one exponent for the whole batch (ct * scalar, ipcl/ciphertext.cpp:97-99;
    // Miller-Rabin rounds): sliding-window schedule instead of scanning
    const char* ns = getenv("IPCLB200_NO_SCHED");
    if (!(ns && ns[0] == '1')) {
      std::vector<uint8_t> sc = build_schedule(hbn::from_words(exp, exp_words), kSchedWindow);
      uint32_t* d_sc;
      TRY(scratch_get(8, (sc.size() + 3) / 4, &d_sc));
      CUDA_TRY(cudaMemcpyAsync(d_sc, sc.data(), sc.size(), cudaMemcpyHostToDevice, s));
      CUDA_TRY(cudaStreamSynchronize(s));  // sc dies at scope end
      p.sched = reinterpret_cast<const uint8_t*>(d_sc);
    }
  }
  TRY(launch_modexp(p, L, s));
  TRY(download_padded(out, d_out, mod_words, L, count, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return 0;
}

int modexp_scalar(const Limbs& base, const Limbs& e, const Limbs& mod,
                  Limbs* out) {
  int mw = (int)mod.size();
  int ew = e.empty() ? 1 : (int)e.size();
  std::vector<uint32_t> b(mw), x(ew), m(mw), r(mw);
  Limbs br = hbn::mod(base, mod);
  hbn::to_words(br, b.data(), mw);
  hbn::to_words(e, x.data(), ew);
  hbn::to_words(mod, m.data(), mw);
  TRY(modexp_host_impl(b.data(), x.data(), m.data(), mw, ew, 1,
                       IPCLB200_SHARED_MOD, r.data()));
  *out = hbn::from_words(r.data(), mw);
  return 0;
}

// small = true: the starter table (8-bit windows, 17 MB at a 2048-bit key);
// false: the widest table the budget allows
int build_comb(const ipclb200_pubkey* pk, int bits, bool small, cudaStream_t s) {
  const int L = pk->L;
  constexpr int kStarterWindow = 8;
  if (pk->d_comb && pk->comb_windows * pk->comb_w >= bits && (small || pk->comb_full))
    return 0;
  // Widest window (<= 16 bits) whose table fits the budget.  B200 has 180 GB of
  // HBM and the kernel needs one 4L-byte entry per window and element, so the
  // table can be large: 1024-bit r at a 2048-bit key, w = 16 -> 64 windows x
  // 65536 entries x 512 B = 2.1 GB and 64 multiplies per encryption.
  // Measured on B200, ms per 65536 encryptions: w = 8/9/10/11 -> 40.4/36.0/
  // 32.6/30.1 (static split); see DESIGN.md for the wider ones.
  size_t budget_mb = 4096;
  if (const char* e = getenv("IPCLB200_COMB_MAX_MB")) budget_mb = strtoul(e, nullptr, 10);
  int w = 16;
  auto table_words = [&](int ww) {
    return (size_t)((bits + ww - 1) / ww) * ((size_t)L << ww);
  };
  while (w > 4 && table_words(w) * 4 > (budget_mb << 20)) w--;
  if (small && w > kStarterWindow) w = kStarterWindow;
  if (const char* cw = getenv("IPCLB200_COMB_WINDOW")) {
    int v = atoi(cw);
    if (v >= 1 && v <= 16) w = v;
  }
  if (pk->d_comb) {
    // a wider exponent than the table covers: rebuild
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaFree(pk->d_comb));
    pk->d_comb = nullptr;
    pk->comb_windows = 0;
  }
  // back off to narrower windows if the allocation does not fit
  for (;; w--) {
    cudaError_t e = cudaMalloc(&pk->d_comb, table_words(w) * sizeof(uint32_t));
    if (e == cudaSuccess) break;
    cudaGetLastError();
    pk->d_comb = nullptr;
    if (w <= 4) return fail(IPCLB200_ERR_CUDA, "cannot allocate the fixed-base table");
  }
  int windows = (bits + w - 1) / w;
  if (windows < 1) windows = 1;
  CombParams cp{};
  cp.m = pk->msq->mc;
  cp.hs_m = pk->d_const + L;
  cp.comb = pk->d_comb;
  cp.w = w;
  cp.w_lo = w > 11 ? (w + 1) / 2 : w;  // two-level build for wide windows
  cp.windows = windows;
  const int nchains = cp.w_lo < w ? 2 * windows : windows;
#define F(K_, T_)                                                       \
  {                                                                     \
    comb_spine_kernel<K_, T_><<<1, 32, 0, s>>>(cp);                     \
    int gpb = kBlockThreads / T_;                                       \
    int grid = (nchains + gpb - 1) / gpb;                               \
    comb_fill_kernel<K_, T_><<<grid, kBlockThreads, 0, s>>>(cp);        \
    if (cp.w_lo < w) {                                                  \
      int eg = 0;                                                       \
      TRY(grid_for(comb_expand_kernel<K_, T_>,                          \
                   (size_t)windows << w, T_, &eg));                     \
      comb_expand_kernel<K_, T_><<<eg, kBlockThreads, 0, s>>>(cp);      \
      g_ctx.launches++;                                                 \
    }                                                                   \
  }
  IPCLB200_DISPATCH(L, F)
#undef F
  g_ctx.launches += 2;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(s));  // one-off; later users may be on other streams
  pk->comb_w = w;
  pk->comb_windows = windows;
  pk->comb_full = !small;
  return 0;
}

int encrypt_dev_impl(const ipclb200_pubkey* pk, const uint32_t* d_pt,
                     int pt_words, const uint32_t* d_r, int r_words,
                     int r_bits, size_t count, int make_secure, uint32_t* d_ct,
                     cudaStream_t s) {
  const int L = pk->L;
  EncryptParams p{};
  p.pt = d_pt;
  p.pt_words = pt_words;
  p.r = d_r;
  p.r_words = r_words;
  p.r_bits = r_bits;
  p.m = pk->msq->mc;
  p.nR = pk->d_const;
  p.hs_m = pk->d_const + L;
  p.n_exp = pk->d_const + 2 * L;
  p.n_exp_words = pk->nl;
  p.ct = d_ct;
  p.count = count;
  p.window = 1;
  if (!make_secure) {
    p.mode = 0;
  } else if (pk->djn) {
    const char* no_comb = getenv("IPCLB200_NO_COMB");
    // DJN: fixed-base comb (built on first use, ~20 ms and up to 160 MB per
    // key; IPCLB200_NO_COMB=1 keeps the generic windowed path instead)
    if (!(no_comb && no_comb[0] == '1')) {
      // a key starts with a small table (8-bit windows: 17 MB and ~130 products
      // per encryption at a 2048-bit key) and moves to the widest one the budget
      // allows (16-bit windows, 2.1 GB, 66 products) once it has encrypted
      // IPCLB200_COMB_UPGRADE (default 8192) elements: a key that is used a few
      // times does not pay for 2 GB of HBM
      size_t upgrade_at = 8192;
      if (const char* e = getenv("IPCLB200_COMB_UPGRADE")) upgrade_at = strtoul(e, nullptr, 10);
      pk->enc_total += count;
      TRY(build_comb(pk, r_bits > pk->rand_bits ? r_bits : pk->rand_bits,
                     pk->enc_total < upgrade_at, s));
      p.mode = 1;
      p.comb = pk->d_comb;
      p.comb_w = pk->comb_w;
      p.comb_windows = (r_bits + pk->comb_w - 1) / pk->comb_w;
      if (p.comb_windows < 1) p.comb_windows = 1;
    } else {
      p.mode = 2;
      p.window = pick_window(r_bits);
    }
  } else {
    p.mode = 3;
    const char* ns = getenv("IPCLB200_NO_SCHED");
    if (pk->d_sched_n && !(ns && ns[0] == '1')) {
      p.sched_n = pk->d_sched_n;
      p.window = kSchedWindow - 1;
    } else {
      p.window = pick_window(pk->nl * 32);
    }
  }
  int grid = 0;
#define F(K_, T_)                                                          \
  {                                                                        \
    TRY(grid_for(encrypt_kernel<K_, T_>, count, T_, &grid));               \
    size_t groups = (size_t)grid * (kBlockThreads / T_);                   \
    TRY(table_ws_with_counter(s, groups * ((size_t)L << p.window),         \
                              &p.table_ws, &p.work_counter));              \
    encrypt_kernel<K_, T_><<<grid, kBlockThreads, 0, s>>>(p);              \
  }
  switch (pick_layout(count, L)) {
    case 1: IPCLB200_DISPATCH_WIDE(L, F) break;
    case 2: IPCLB200_DISPATCH_MID(L, F) break;
    default: IPCLB200_DISPATCH(L, F)
  }
#undef F
  g_ctx.launches++;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// CRT decrypt in two-digit arithmetic (decrypt_hensel_kernel + crt_combine_kernel).
// d_x: count x 2*pl words (mp | mq per ciphertext).
int decrypt_hensel_impl(const ipclb200_privkey* sk, const uint32_t* d_ct,
                        size_t count, uint32_t* d_pt, uint32_t* d_x,
                        cudaStream_t s) {
  const int pl = sk->pl;
  DecryptHenselParams p{};
  p.ct = d_ct;
  p.s0.blk = sk->d_hblk_p;
  p.s0.sched = sk->d_hsched_p;
  p.s0.n0inv = sk->p_n0inv;
  p.s1.blk = sk->d_hblk_q;
  p.s1.sched = sk->d_hsched_q;
  p.s1.n0inv = sk->q_n0inv;
  p.mpq = d_x;
  p.count = count;
  p.table_entries = 1 << (kSchedWindow - 1);
  // warps per SM: IPCLB200_HENSEL_BLOCKS blocks of 128 threads (default 3)
  int want_blocks = 3;
  if (const char* e = getenv("IPCLB200_HENSEL_BLOCKS")) want_blocks = atoi(e);
  if (want_blocks < 1 || want_blocks > 4) want_blocks = 3;
#define FH(K_, T_, MINB_, ROWS_)                                                        \
  {                                                                                \
    auto kern = decrypt_hensel_kernel<K_, T_, MINB_, ROWS_>;                          \
    constexpr size_t smem = hensel_smem_bytes<K_, T_>(kBlockThreads);              \
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  (int)smem));                                     \
    int per_sm = 0;                                                                \
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern,          \
                                                           kBlockThreads, smem));  \
    if (per_sm < 1) return fail(IPCLB200_ERR_CUDA, "hensel kernel does not fit an SM"); \
    if (per_sm > want_blocks) per_sm = want_blocks;                                \
    const size_t gpb = kBlockThreads / T_;                                         \
    const size_t chunks = 2 * ((count + (32 / T_) - 1) / (32 / T_));               \
    const size_t need = (chunks + 3) / 4;                                          \
    const size_t cap = (size_t)per_sm * g_ctx.sms;                                 \
    const int grid = (int)(need < cap ? need : cap);                               \
    TRY(table_ws_with_counter(s, (size_t)grid * gpb * 2 * pl * p.table_entries,    \
                              &p.table_ws, &p.work_counter));                      \
    kern<<<grid, kBlockThreads, smem, s>>>(p);                                     \
  }
  int rows = 8;
  if (const char* e = getenv("IPCLB200_HENSEL_ROWS")) rows = atoi(e);
  switch (pl) {
    case 16: FH(8, 2, 3, 8) break;
    case 32:
      if (rows == 4) FH(16, 2, 3, 4) else if (rows == 16) FH(16, 2, 3, 16) else FH(16, 2, 3, 8)
      break;
    case 48: FH(24, 2, 2, 8) break;
    case 64: FH(16, 4, 3, 8) break;
    default: return fail(IPCLB200_ERR_UNSUPPORTED, "hensel: unsupported prime width");
  }
#undef FH
  g_ctx.launches++;
  CUDA_TRY(cudaGetLastError());
  CrtCombineParams f{};
  f.mpq = d_x;
  f.p = sk->d_p;
  f.q = sk->d_q;
  f.pinvR = sk->d_pinvR;
  f.q_n0inv = sk->q_n0inv;
  f.pl = pl;
  f.pt = d_pt;
  f.count = count;
  crt_combine_kernel<<<(unsigned)((count + 63) / 64), 64, 0, s>>>(f);
  g_ctx.launches++;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int decrypt_dev_impl(const ipclb200_privkey* sk, const uint32_t* d_ct,
                     size_t count, int use_crt, uint32_t* d_pt,
                     uint32_t* d_x /* count x 4*pl words scratch */,
                     cudaStream_t s) {
  const int pl = sk->pl;
  // the thread-per-integer kernel (mont_tile.cuh) does 15 % fewer multiplies
  // but measured slower on B200 (194 ms vs 165 ms per 65536 at a 2048-bit
  // key: only 8 warps/SM fit its shared-memory columns and it issues 2.7
  // instructions per multiply); it stays opt-in
  const char* force = getenv("IPCLB200_DECRYPT");
  // default: two-digit (Hensel) arithmetic, half the multiplies of the generic
  // kernel; IPCLB200_DECRYPT=int keeps the full-width kernel
  if (use_crt && sk->hensel_ok && (!force || !strcmp(force, "hensel")))
    return decrypt_hensel_impl(sk, d_ct, count, d_pt, d_x, s);
#ifdef IPCLB200_EXPERIMENTS
  const bool tile = force && !strcmp(force, "tile");
  if (use_crt && tile && (sk->L == 32 || sk->L == 48 || sk->L == 64)) {
    const int L = sk->L;
    DecryptTileParams p{};
    p.ct = d_ct;
    p.m0 = sk->mp2->mc;
    p.m1 = sk->mq2->mc;
    p.ninv0_lo = make_uint4(sk->ninv_p[0], sk->ninv_p[1], sk->ninv_p[2], sk->ninv_p[3]);
    p.ninv0_hi = make_uint4(sk->ninv_p[4], sk->ninv_p[5], sk->ninv_p[6], sk->ninv_p[7]);
    p.ninv1_lo = make_uint4(sk->ninv_q[0], sk->ninv_q[1], sk->ninv_q[2], sk->ninv_q[3]);
    p.ninv1_hi = make_uint4(sk->ninv_q[4], sk->ninv_q[5], sk->ninv_q[6], sk->ninv_q[7]);
    p.prog0 = sk->d_prog_p;
    p.prog1 = sk->d_prog_q;
    p.x = d_x;
    p.count = count;
    p.slots = sk->tile_slots;
    constexpr int NT = 128;
    const int V = L / 4;
    const size_t smem = (size_t)(2 * V * NT + 4 * V) * 16;
    int grid = 0;
#define FT(NB_)                                                                   \
  {                                                                               \
    auto kern = decrypt_tile_kernel<NB_, NT>;                                     \
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  (int)smem));                                    \
    int per_sm = 0;                                                               \
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem)); \
    if (per_sm < 1) return fail(IPCLB200_ERR_CUDA, "tile kernel does not fit an SM"); \
    size_t need = (2 * ((count + 31) / 32) + NT / 32 - 1) / (NT / 32);            \
    size_t cap = (size_t)per_sm * g_ctx.sms;                                      \
    grid = (int)(need < cap ? need : cap);                                        \
    size_t ws_words = (size_t)grid * (NT / 32) * p.slots * V * 32 * 4 + 4;        \
    uint32_t* ws = nullptr;                                                       \
    TRY(table_ws_get((void*)s, ws_words, &ws));                                   \
    p.table_ws = reinterpret_cast<uint4*>(ws);                                    \
    p.work_counter = ws + (ws_words - 4);                                         \
    CUDA_TRY(cudaMemsetAsync(p.work_counter, 0, 16, s));                          \
    kern<<<grid, NT, smem, s>>>(p);                                               \
  }
    if (L == 32) FT(4) else if (L == 48) FT(6) else FT(8)
#undef FT
    CrtFinishParams f{};
    f.x = d_x;
    f.p = sk->d_p;
    f.q = sk->d_q;
    f.hpR = sk->d_hpR;
    f.hqR = sk->d_hqR;
    f.pinvR = sk->d_pinvR;
    f.p_inv32 = sk->p_inv32;
    f.q_inv32 = sk->q_inv32;
    f.p_n0inv = sk->p_n0inv;
    f.q_n0inv = sk->q_n0inv;
    f.pl = pl;
    f.xl = L;
    f.pt = d_pt;
    f.count = count;
    crt_finish_kernel<<<(unsigned)((count + 63) / 64), 64, 0, s>>>(f);
    g_ctx.launches += 2;
    CUDA_TRY(cudaGetLastError());
  } else
#endif
  if (use_crt) {
    const int L = sk->L;
    DecryptCrtParams p{};
    p.ct = d_ct;
    p.m0 = sk->mp2->mc;
    p.m1 = sk->mq2->mc;
    p.sched0 = sk->d_sched_p;
    p.sched1 = sk->d_sched_q;
    p.x = d_x;
    p.count = count;
    p.table_entries = 1 << (kSchedWindow - 1);
    // which pipes: "int" = integer kernel only, "fp" = FP64 kernel only,
    // "dual" = both roles in one kernel, "dual2" = two kernels on two streams
    // sharing the work counter.  FP64 needs the 64-word class (2048-bit key).
#ifdef IPCLB200_EXPERIMENTS
    const char* mode = force ? force : "int";
    const bool want_fp = !strcmp(mode, "fp"), want_dual = !strcmp(mode, "dual"),
               want_dual2 = !strcmp(mode, "dual2"), want_sqr = !strcmp(mode, "sqr");
    if ((!strcmp(mode, "k32s") || !strcmp(mode, "k32s2")) && L == 64) {
      // 32 x 2 layout, multiplier rows from shared memory; 3 (or 2) blocks per SM
      const bool three = !strcmp(mode, "k32s");
      const size_t smem = (size_t)(kBlockThreads / 2) * kSbGroupWords * sizeof(uint32_t);
      int per_sm = 0;
      if (three)
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &per_sm, decrypt_crt_k32s_kernel<3>, kBlockThreads, smem));
      else
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &per_sm, decrypt_crt_k32s_kernel<2>, kBlockThreads, smem));
      if (per_sm < 1) return fail(IPCLB200_ERR_CUDA, "k32s kernel does not fit an SM");
      const size_t need = (2 * ((count + 15) / 16) + 3) / 4;
      const size_t capb = (size_t)per_sm * g_ctx.sms;
      const int grid = (int)(need < capb ? need : capb);
      TRY(table_ws_with_counter(s, (size_t)grid * 64 * L * p.table_entries, &p.table_ws,
                                &p.work_counter));
      if (three)
        decrypt_crt_k32s_kernel<3><<<grid, kBlockThreads, smem, s>>>(p);
      else
        decrypt_crt_k32s_kernel<2><<<grid, kBlockThreads, smem, s>>>(p);
      g_ctx.launches++;
      CUDA_TRY(cudaGetLastError());
    } else if (!strcmp(mode, "sqr2") && L == 64 && pick_layout(2 * count, L) == 0) {
      // symmetric squarings in the 32 x 2 layout (MontSqr2)
      constexpr size_t smem = sqr2_smem_bytes(kBlockThreads);
      static bool attr_set2 = false;
      if (!attr_set2) {
        CUDA_TRY(cudaFuncSetAttribute(decrypt_crt_sqr2_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set2 = true;
      }
      int per_sm = 0;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decrypt_crt_sqr2_kernel,
                                                             kBlockThreads, smem));
      if (per_sm < 1) return fail(IPCLB200_ERR_CUDA, "sqr2 kernel does not fit an SM");
      const size_t need = (2 * ((count + 15) / 16) + 3) / 4;
      const size_t capb = (size_t)per_sm * g_ctx.sms;
      const int grid = (int)(need < capb ? need : capb);
      TRY(table_ws_with_counter(s, (size_t)grid * 64 * L * p.table_entries, &p.table_ws,
                                &p.work_counter));
      decrypt_crt_sqr2_kernel<<<grid, kBlockThreads, smem, s>>>(p);
      g_ctx.launches++;
      CUDA_TRY(cudaGetLastError());
    } else if (!strcmp(mode, "k32") && L == 64) {
      int per_sm = 0;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decrypt_crt_k32_kernel,
                                                             kBlockThreads, 0));
      if (per_sm < 1) return fail(IPCLB200_ERR_CUDA, "k32 kernel does not fit an SM");
      const size_t need = (2 * ((count + 15) / 16) + 3) / 4;
      const size_t capb = (size_t)per_sm * g_ctx.sms;
      const int grid = (int)(need < capb ? need : capb);
      TRY(table_ws_with_counter(s, (size_t)grid * 64 * L * p.table_entries, &p.table_ws,
                                &p.work_counter));
      decrypt_crt_k32_kernel<<<grid, kBlockThreads, 0, s>>>(p);
      g_ctx.launches++;
      CUDA_TRY(cudaGetLastError());
    } else if (want_sqr && L == 64 && pick_layout(2 * count, L) == 0) {
      // symmetric squarings (mont_sqr.cuh), 64-word class
      constexpr size_t smem = sqr_smem_bytes(kBlockThreads, 4);
      static bool attr_set = false;
      if (!attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(decrypt_crt_sqr_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
      }
      int per_sm = 0;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decrypt_crt_sqr_kernel,
                                                             kBlockThreads, smem));
      if (per_sm < 1) return fail(IPCLB200_ERR_CUDA, "sqr kernel does not fit an SM");
      const size_t need = (2 * ((count + 7) / 8) + 3) / 4;
      const size_t capb = (size_t)per_sm * g_ctx.sms;
      const int grid = (int)(need < capb ? need : capb);
      TRY(table_ws_with_counter(s, (size_t)grid * 32 * L * p.table_entries, &p.table_ws,
                                &p.work_counter));
      decrypt_crt_sqr_kernel<<<grid, kBlockThreads, smem, s>>>(p);
      g_ctx.launches++;
      CUDA_TRY(cudaGetLastError());
    } else
    if ((want_fp || want_dual || want_dual2) && sk->fp_ok && L == kFpWords) {
      constexpr int FL = kFpLimbs;
      constexpr size_t smem = fp_role_smem(kFpK, kFpT);
      const int sms = g_ctx.sms;
      const size_t chunks = 2 * ((count + 7) / 8);          // warps of work
      const size_t need = (chunks + 3) / 4;                 // blocks of 4 warps
      auto env_int = [](const char* name, int dflt) {
        const char* e = getenv(name);
        return e ? atoi(e) : dflt;
      };
      DecryptFpParams f{};
      f.ct = d_ct;
      f.f0 = sk->fp0;
      f.f1 = sk->fp1;
      f.sched0 = sk->d_sched_p;
      f.sched1 = sk->d_sched_q;
      f.x = d_x;
      f.count = count;
      f.table_entries = p.table_entries;
      f.debug_stage = want_fp ? env_int("IPCLB200_FP_DEBUG_STAGE", 0) : 0;
      auto cap = [&](int per_sm) {
        size_t c = (size_t)per_sm * sms;
        return (int)(need < c ? need : c);
      };
      if (want_fp) {
        int blocks = env_int("IPCLB200_FP_BLOCKS", 2);
        if (blocks < 1 || blocks > 3) blocks = 2;
        const int grid = cap(blocks);
        TRY(table_ws_with_counter(s, (size_t)grid * 32 * FL * f.table_entries,
                                  &f.table_ws, &f.work_counter, 1));
        if (blocks == 3)
          decrypt_crt_fp_kernel<kFpK, kFpT, kFpWords, 3><<<grid, kBlockThreads, smem, s>>>(f);
        else
          decrypt_crt_fp_kernel<kFpK, kFpT, kFpWords, 2><<<grid, kBlockThreads, smem, s>>>(f);
        g_ctx.launches++;
      } else if (want_dual) {
        DecryptDualParams d{};
        const int grid = cap(3);
        TRY(table_ws_with_counter(s, (size_t)grid * 32 * L * p.table_entries,
                                  &p.table_ws, &p.work_counter));
        uint32_t* fws = nullptr;
        const size_t fwords = (size_t)grid * 32 * FL * f.table_entries;
        TRY(table_ws_get((void*)((uintptr_t)s + 1), fwords + 1024, &fws));
        f.table_ws = fws;
        f.work_counter = p.work_counter;
        d.i = p;
        d.f = f;
        d.sm_slots = fws + fwords;
        d.fp_mask = (unsigned)env_int("IPCLB200_FP_MASK", 4);
        d.slots_per_sm = 3;
        CUDA_TRY(cudaMemsetAsync(d.sm_slots, 0, 1024 * sizeof(uint32_t), s));
        decrypt_crt_dual_kernel<16, 4, kFpK, kFpT><<<grid, kBlockThreads, smem, s>>>(d);
        g_ctx.launches++;
      } else {
        int ib = 2, fb = 1;
        if (const char* e = getenv("IPCLB200_DUAL2")) sscanf(e, "%d,%d", &ib, &fb);
        if (ib < 0 || ib > 3) ib = 2;
        if (fb < 0 || fb > 2 || (ib == 0 && fb == 0)) fb = 1;
        const int igrid = cap(ib ? ib : 1), fgrid = cap(fb ? fb : 1);
        TRY(table_ws_with_counter(s, (size_t)igrid * 32 * L * p.table_entries,
                                  &p.table_ws, &p.work_counter));
        uint32_t* fws = nullptr;
        TRY(table_ws_get((void*)((uintptr_t)s + 1),
                         (size_t)fgrid * 32 * FL * f.table_entries, &fws));
        f.table_ws = fws;
        f.work_counter = p.work_counter;
        // kernels with different shared-memory carve-outs cannot share an SM:
        // ask for the same one for both
        static bool carve_set = false;
        if (!carve_set) {
          const int carve = env_int("IPCLB200_CARVEOUT", 100);
          if (carve >= 0) {
            CUDA_TRY(cudaFuncSetAttribute(decrypt_crt_kernel<16, 4>,
                                          cudaFuncAttributePreferredSharedMemoryCarveout, carve));
            CUDA_TRY(cudaFuncSetAttribute(decrypt_crt_fp224_kernel<kFpK, kFpT, kFpWords>,
                                          cudaFuncAttributePreferredSharedMemoryCarveout, carve));
            CUDA_TRY(cudaFuncSetAttribute(decrypt_crt_fp_kernel<kFpK, kFpT, kFpWords, 3>,
                                          cudaFuncAttributePreferredSharedMemoryCarveout, carve));
          }
          carve_set = true;
        }
        cudaStream_t s2 = g_ctx.aux_stream;
        CUDA_TRY(cudaEventRecord(g_ctx.ev_fork, s));
        CUDA_TRY(cudaStreamWaitEvent(s2, g_ctx.ev_fork, 0));
        if (ib) decrypt_crt_kernel<16, 4><<<igrid, kBlockThreads, 0, s>>>(p);
        if (fb == 1)
          decrypt_crt_fp224_kernel<kFpK, kFpT, kFpWords><<<fgrid, kBlockThreads, smem, s2>>>(f);
        else if (fb == 2)
          decrypt_crt_fp_kernel<kFpK, kFpT, kFpWords, 3><<<fgrid, kBlockThreads, smem, s2>>>(f);
        CUDA_TRY(cudaEventRecord(g_ctx.ev_join, s2));
        CUDA_TRY(cudaStreamWaitEvent(s, g_ctx.ev_join, 0));
        g_ctx.launches += 2;
      }
      CUDA_TRY(cudaGetLastError());
    } else
#endif
    {
      int grid = 0;
#define F(K_, T_)                                                          \
  {                                                                        \
    TRY(grid_for(decrypt_crt_kernel<K_, T_>, 2 * count, T_, &grid));       \
    size_t groups = (size_t)grid * (kBlockThreads / T_);                   \
    TRY(table_ws_with_counter(s, groups * ((size_t)L * p.table_entries),   \
                              &p.table_ws, &p.work_counter));              \
    decrypt_crt_kernel<K_, T_><<<grid, kBlockThreads, 0, s>>>(p);          \
  }
      switch (pick_layout(2 * count, L)) {
        case 1: IPCLB200_DISPATCH_WIDE(L, F) break;
        case 2: IPCLB200_DISPATCH_MID(L, F) break;
        default: IPCLB200_DISPATCH(L, F)
      }
#undef F
      g_ctx.launches++;
    }
    CrtFinishParams f{};
    f.x = d_x;
    f.p = sk->d_p;
    f.q = sk->d_q;
    f.hpR = sk->d_hpR;
    f.hqR = sk->d_hqR;
    f.pinvR = sk->d_pinvR;
    f.p_inv32 = sk->p_inv32;
    f.q_inv32 = sk->q_inv32;
    f.p_n0inv = sk->p_n0inv;
    f.q_n0inv = sk->q_n0inv;
    f.pl = pl;
    f.xl = L;
    f.pt = d_pt;
    f.count = count;
    crt_finish_kernel<<<(unsigned)((count + 63) / 64), 64, 0, s>>>(f);
    g_ctx.launches++;
    CUDA_TRY(cudaGetLastError());
  } else {
    const int L = sk->mnsq->L;
    ModexpParams p{};
    p.base = d_ct;
    p.base_stride = L;
    p.exp = sk->d_lambda;
    p.exp_stride = 0;
    p.exp_words = 2 * pl;
    p.exp_bits = sk->lambda_bits;
    p.n = sk->mnsq->mc.n;
    p.rr = sk->mnsq->mc.rr;
    p.one = sk->mnsq->mc.one;
    p.n0inv = sk->mnsq->d_n0inv;
    p.mod_stride = 0;
    p.n0_stride = 0;
    p.out = d_x;
    p.count = count;
    {
      const char* ns = getenv("IPCLB200_NO_SCHED");
      if (!(ns && ns[0] == '1')) p.sched = sk->d_sched_lambda;
    }
    TRY(launch_modexp(p, L, s));
    RawFinishParams f{};
    f.x = d_x;
    f.n = sk->d_n;
    f.muR = sk->d_muR;
    f.n_inv32 = sk->n_inv32;
    f.n_n0inv = sk->n_n0inv;
    f.nl = 2 * pl;
    f.xl = L;
    f.pt = d_pt;
    f.count = count;
    raw_finish_kernel<<<(unsigned)((count + 63) / 64), 64, 0, s>>>(f);
    g_ctx.launches++;
    CUDA_TRY(cudaGetLastError());
  }
  return 0;
}

}  // namespace

// ===========================================================================
// exported C ABI
// ===========================================================================
extern "C" {

const char* ipclb200_version(void) { return "ipcl_b200 0.1 (sm_100a)"; }
const char* ipclb200_last_error(void) { return t_err.c_str(); }

int ipclb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int ipclb200_init(int device) {
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  if (g_ctx.ready && (device < 0 || device == g_ctx.device)) return 0;
  if (g_ctx.ready)
    return fail(IPCLB200_ERR_BAD_ARG, "already initialised on another device");
  if (device >= 0) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev)
      return fail(IPCLB200_ERR_NO_DEVICE, "no such CUDA device");
    g_requested_device = device;
  }
  return ensure_init_locked();
}

void ipclb200_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  if (!g_ctx.ready) return;
  cudaSetDevice(g_ctx.device);
  cudaDeviceSynchronize();
  for (int i = 0; i < Ctx::kSlots; i++) {
    if (g_ctx.scratch[i]) cudaFree(g_ctx.scratch[i]);
    g_ctx.scratch[i] = nullptr;
    g_ctx.scratch_words[i] = 0;
  }
  for (auto& kv : g_ctx.table_ws)
    if (kv.second.first) cudaFree(kv.second.first);
  g_ctx.table_ws.clear();
  g_ctx.mod_cache.clear();
  if (g_ctx.stream) cudaStreamDestroy(g_ctx.stream);
  g_ctx.stream = nullptr;
  if (g_ctx.aux_stream) cudaStreamDestroy(g_ctx.aux_stream);
  g_ctx.aux_stream = nullptr;
  if (g_ctx.ev_fork) cudaEventDestroy(g_ctx.ev_fork);
  if (g_ctx.ev_join) cudaEventDestroy(g_ctx.ev_join);
  g_ctx.ev_fork = g_ctx.ev_join = nullptr;
  g_ctx.ready = false;
}

uint64_t ipclb200_launch_count(void) { return g_ctx.launches.load(); }

int ipclb200_modexp(const uint32_t* base, const uint32_t* exp,
                    const uint32_t* mod, int mod_words, int exp_words,
                    size_t count, unsigned flags, uint32_t* out) {
  if (!base || !exp || !mod || !out)
    return fail(IPCLB200_ERR_BAD_ARG, "modexp: null pointer");
  if (mod_words <= 0 || exp_words <= 0)
    return fail(IPCLB200_ERR_BAD_ARG, "modexp: non-positive width");
  if (mod_words > IPCLB200_MAX_MOD_WORDS)
    return fail(IPCLB200_ERR_UNSUPPORTED, "modexp: modulus wider than 8192 bits");
  if (count == 0) return 0;
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  TRY(ensure_init_locked());
  return modexp_host_impl(base, exp, mod, mod_words, exp_words, count, flags, out);
}

int ipclb200_modexp_dev(const uint32_t* d_base, const uint32_t* d_exp,
                        const uint32_t* h_mod, int mod_words, int exp_words,
                        int exp_bits, size_t count, unsigned flags,
                        uint32_t* d_out, void* stream) {
  if (!d_base || !d_exp || !h_mod || !d_out)
    return fail(IPCLB200_ERR_BAD_ARG, "modexp_dev: null pointer");
  if (!(flags & IPCLB200_SHARED_MOD))
    return fail(IPCLB200_ERR_UNSUPPORTED, "modexp_dev: needs IPCLB200_SHARED_MOD");
  if (mod_words <= 0 || exp_words <= 0 || class_words(mod_words) != mod_words)
    return fail(IPCLB200_ERR_UNSUPPORTED,
                "modexp_dev: mod_words must be one of 16,32,48,64,96,128,192,256");
  if (count == 0) return 0;
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  TRY(ensure_init_locked());
  Limbs n;
  TRY(check_modulus(h_mod, mod_words, &n));
  std::shared_ptr<DevModulus> dm;
  TRY(make_modulus(n, mod_words, &dm));
  cudaStream_t s = (cudaStream_t)stream;
  ModexpParams p{};
  p.base = d_base;
  p.base_stride = (flags & IPCLB200_SHARED_BASE) ? 0 : mod_words;
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
  return launch_modexp(p, mod_words, s);
}

static int modmul_common(const uint32_t* d_a, const uint32_t* d_b,
                         const Limbs& n, int L, size_t count, unsigned flags,
                         uint32_t* d_out, cudaStream_t s) {
  std::shared_ptr<DevModulus> dm;
  TRY(make_modulus(n, L, &dm));
  ModmulParams p{};
  p.a = d_a;
  p.b = d_b;
  p.b_stride = (flags & IPCLB200_SHARED_B) ? 0 : L;
  p.m = dm->mc;
  p.out = d_out;
  p.count = count;
  return launch_modmul(p, L, s);
}

int ipclb200_modmul(const uint32_t* a, const uint32_t* b, const uint32_t* mod,
                    int mod_words, size_t count, unsigned flags, uint32_t* out) {
  if (!a || !b || !mod || !out)
    return fail(IPCLB200_ERR_BAD_ARG, "modmul: null pointer");
  if (mod_words <= 0) return fail(IPCLB200_ERR_BAD_ARG, "modmul: non-positive width");
  const int L = class_words(mod_words);
  if (!L) return fail(IPCLB200_ERR_UNSUPPORTED, "modmul: modulus wider than 8192 bits");
  if (count == 0) return 0;
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  TRY(ensure_init_locked());
  Limbs n;
  TRY(check_modulus(mod, mod_words, &n));
  cudaStream_t s = g_ctx.stream;
  const bool sh_b = flags & IPCLB200_SHARED_B;
  uint32_t *d_a, *d_b, *d_out;
  TRY(scratch_get(0, count * (size_t)L, &d_a));
  TRY(scratch_get(1, (sh_b ? 1 : count) * (size_t)L, &d_b));
  TRY(scratch_get(2, count * (size_t)L, &d_out));
  TRY(upload_padded(d_a, a, mod_words, L, count, s));
  TRY(upload_padded(d_b, b, mod_words, L, sh_b ? 1 : count, s));
  TRY(modmul_common(d_a, d_b, n, L, count, flags, d_out, s));
  TRY(download_padded(out, d_out, mod_words, L, count, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return 0;
}

int ipclb200_modmul_dev(const uint32_t* d_a, const uint32_t* d_b,
                        const uint32_t* h_mod, int mod_words, size_t count,
                        unsigned flags, uint32_t* d_out, void* stream) {
  if (!d_a || !d_b || !h_mod || !d_out)
    return fail(IPCLB200_ERR_BAD_ARG, "modmul_dev: null pointer");
  if (mod_words <= 0 || class_words(mod_words) != mod_words)
    return fail(IPCLB200_ERR_UNSUPPORTED,
                "modmul_dev: mod_words must be one of 16,32,48,64,96,128,192,256");
  if (count == 0) return 0;
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  TRY(ensure_init_locked());
  Limbs n;
  TRY(check_modulus(h_mod, mod_words, &n));
  return modmul_common(d_a, d_b, n, mod_words, count, flags, d_out,
                       (cudaStream_t)stream);
}

// ---- public key -----------------------------------------------------------
int ipclb200_pubkey_create(const uint32_t* n, int n_words, const uint32_t* hs,
                           int rand_bits, ipclb200_pubkey** out) {
  if (!n || !out || n_words <= 0)
    return fail(IPCLB200_ERR_BAD_ARG, "pubkey_create: bad argument");
  if (2 * n_words > IPCLB200_MAX_MOD_WORDS)
    return fail(IPCLB200_ERR_UNSUPPORTED, "pubkey_create: key wider than 4096 bits");
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  TRY(ensure_init_locked());
  std::unique_ptr<ipclb200_pubkey> pk(new ipclb200_pubkey);
  pk->nl = n_words;
  pk->n = hbn::from_words(n, n_words);
  if (pk->n.empty()) return fail(IPCLB200_ERR_BAD_ARG, "pubkey_create: n is zero");
  if (!(pk->n[0] & 1u))
    return fail(IPCLB200_ERR_EVEN_MODULUS, "pubkey_create: n is even");
  pk->nsq = hbn::mul(pk->n, pk->n);
  pk->L = class_words(2 * n_words);
  const int L = pk->L;
  TRY(make_modulus(pk->nsq, L, &pk->msq));
  // n*R mod n^2, n (exponent), hs*R mod n^2
  Limbs R = hbn::pow2(32u * (unsigned)L);
  Limbs nR = hbn::mod(hbn::mul(pk->n, R), pk->nsq);
  std::vector<uint32_t> blk((size_t)L + n_words + L, 0u);
  hbn::to_words(nR, blk.data(), L);
  hbn::to_words(pk->n, blk.data() + 2 * L, n_words);
  if (hs) {
    pk->djn = true;
    pk->rand_bits = rand_bits > 0 ? rand_bits : n_words * 16;
    pk->hs = hbn::mod(hbn::from_words(hs, 2 * n_words), pk->nsq);
    Limbs hsm = hbn::mod(hbn::mul(pk->hs, R), pk->nsq);
    hbn::to_words(hsm, blk.data() + L, L);
  }
  CUDA_TRY(cudaMalloc(&pk->d_const, blk.size() * sizeof(uint32_t)));
  CUDA_TRY(cudaMemcpy(pk->d_const, blk.data(), blk.size() * sizeof(uint32_t),
                      cudaMemcpyHostToDevice));
  if (!hs) {
    // non-DJN obfuscator r^n: every element has the exponent n
    std::vector<uint8_t> sn = build_schedule(pk->n, kSchedWindow);
    CUDA_TRY(cudaMalloc(&pk->d_sched_n, sn.size()));
    CUDA_TRY(cudaMemcpy(pk->d_sched_n, sn.data(), sn.size(), cudaMemcpyHostToDevice));
  }
  *out = pk.release();
  return 0;
}

void ipclb200_pubkey_destroy(ipclb200_pubkey* pk) {
  if (!pk) return;
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  if (g_ctx.ready) {
    cudaSetDevice(g_ctx.device);
    cudaDeviceSynchronize();
  }
  delete pk;
}

int ipclb200_encrypt(const ipclb200_pubkey* pk, const uint32_t* pt,
                     int pt_words, const uint32_t* r, int r_words, size_t count,
                     int make_secure, uint32_t* ct) {
  if (!pk || !pt || !ct) return fail(IPCLB200_ERR_BAD_ARG, "encrypt: null pointer");
  if (make_secure && !r) return fail(IPCLB200_ERR_BAD_ARG, "encrypt: randoms missing");
  if (pt_words <= 0 || pt_words > pk->nl)
    return fail(IPCLB200_ERR_BAD_ARG, "encrypt: pt_words out of range");
  if (make_secure && (r_words <= 0 || r_words > 2 * pk->nl))
    return fail(IPCLB200_ERR_BAD_ARG, "encrypt: r_words out of range");
  if (count == 0) return 0;
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  TRY(ensure_init_locked());
  cudaStream_t s = g_ctx.stream;
  const int L = pk->L, CW = 2 * pk->nl;
  uint32_t *d_pt, *d_r = nullptr, *d_ct;
  TRY(scratch_get(0, count * (size_t)pt_words, &d_pt));
  TRY(scratch_get(2, count * (size_t)L, &d_ct));
  CUDA_TRY(cudaMemcpyAsync(d_pt, pt, count * (size_t)pt_words * 4,
                           cudaMemcpyHostToDevice, s));
  int r_bits = 0;
  if (make_secure) {
    TRY(scratch_get(1, count * (size_t)r_words, &d_r));
    CUDA_TRY(cudaMemcpyAsync(d_r, r, count * (size_t)r_words * 4,
                             cudaMemcpyHostToDevice, s));
    r_bits = max_bits(r, r_words, count, r_words);
  }
  TRY(encrypt_dev_impl(pk, d_pt, pt_words, d_r, r_words, r_bits, count,
                       make_secure, d_ct, s));
  TRY(download_padded(ct, d_ct, CW, L, count, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return 0;
}

int ipclb200_encrypt_dev(const ipclb200_pubkey* pk, const uint32_t* d_pt,
                         int pt_words, const uint32_t* d_r, int r_words,
                         size_t count, int make_secure, uint32_t* d_ct,
                         void* stream) {
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
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  TRY(ensure_init_locked());
  return encrypt_dev_impl(pk, d_pt, pt_words, d_r, r_words, r_words * 32, count,
                          make_secure, d_ct, (cudaStream_t)stream);
}

// ---- private key ----------------------------------------------------------
int ipclb200_privkey_create(const uint32_t* p_in, const uint32_t* q_in,
                            int p_words, ipclb200_privkey** out) {
  if (!p_in || !q_in || !out || p_words <= 0)
    return fail(IPCLB200_ERR_BAD_ARG, "privkey_create: bad argument");
  if (2 * p_words > kMaxPrimeWords)
    return fail(IPCLB200_ERR_UNSUPPORTED, "privkey_create: primes wider than 2048 bits");
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  TRY(ensure_init_locked());
  std::unique_ptr<ipclb200_privkey> sk(new ipclb200_privkey);
  const int pl = p_words;
  sk->pl = pl;
  Limbs p = hbn::from_words(p_in, pl), q = hbn::from_words(q_in, pl);
  if (hbn::cmp(q, p) < 0) p.swap(q);  // pri_key.cpp:19-22
  if (p.empty() || !(p[0] & 1u) || !(q[0] & 1u))
    return fail(IPCLB200_ERR_EVEN_MODULUS, "privkey_create: p and q must be odd");
  if (hbn::cmp(p, q) == 0)
    return fail(IPCLB200_ERR_BAD_ARG, "privkey_create: p and q are same");
  sk->p = p;
  sk->q = q;
  sk->n = hbn::mul(p, q);
  sk->nsq = hbn::mul(sk->n, sk->n);
  Limbs one = hbn::from_u64(1);
  Limbs psq = hbn::mul(p, p), qsq = hbn::mul(q, q);
  Limbs pm1 = hbn::sub(p, one), qm1 = hbn::sub(q, one);
  Limbs g = hbn::add(sk->n, one);
  sk->L = class_words(2 * pl);
  TRY(make_modulus(psq, sk->L, &sk->mp2));
  TRY(make_modulus(qsq, sk->L, &sk->mq2));
  TRY(make_modulus(sk->nsq, class_words(4 * pl), &sk->mnsq));
  // computeHfun (pri_key.cpp:159-167)
  Limbs hp, hq, pinv, tmp, lq;
  TRY(modexp_scalar(hbn::mod(g, psq), pm1, psq, &tmp));
  hbn::divmod(hbn::sub(tmp, one), p, &lq, nullptr);
  if (!hbn::modinv(lq, p, &hp))
    return fail(IPCLB200_ERR_BAD_ARG, "privkey_create: hp not invertible (p not prime?)");
  TRY(modexp_scalar(hbn::mod(g, qsq), qm1, qsq, &tmp));
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
  TRY(modexp_scalar(g, lam, sk->nsq, &tmp));
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
  std::vector<uint32_t> blk(7 * (size_t)pl + 4 * (size_t)pl + 4 + 2 * (size_t)pl, 0u);
  uint32_t* b = blk.data();
  hbn::to_words(p, b, pl);
  hbn::to_words(q, b + pl, pl);
  hbn::to_words(pm1, b + 2 * pl, pl);
  hbn::to_words(qm1, b + 3 * pl, pl);
  hbn::to_words(hpR, b + 4 * pl, pl);
  hbn::to_words(hqR, b + 5 * pl, pl);
  hbn::to_words(pinvR, b + 6 * pl, pl);
  hbn::to_words(sk->n, b + 7 * pl, 2 * pl);
  hbn::to_words(muR, b + 9 * pl, 2 * pl);
  b[11 * pl] = sk->mnsq->mc.n0inv;
  hbn::to_words(lam, b + 11 * pl + 4, 2 * pl);
  CUDA_TRY(cudaMalloc(&sk->d_const, blk.size() * sizeof(uint32_t)));
  CUDA_TRY(cudaMemcpy(sk->d_const, blk.data(), blk.size() * sizeof(uint32_t),
                      cudaMemcpyHostToDevice));
  const uint32_t* d = sk->d_const;
  sk->d_p = d;
  sk->d_q = d + pl;
  sk->d_pm1 = d + 2 * pl;
  sk->d_qm1 = d + 3 * pl;
  sk->d_hpR = d + 4 * pl;
  sk->d_hqR = d + 5 * pl;
  sk->d_pinvR = d + 6 * pl;
  sk->d_n = d + 7 * pl;
  sk->d_muR = d + 9 * pl;
  sk->d_lambda = d + 11 * pl + 4;
  sk->p_n0inv = hbn::neg_inv32(p[0]);
  sk->q_n0inv = hbn::neg_inv32(q[0]);
  sk->n_n0inv = hbn::neg_inv32(sk->n[0]);
  sk->p_inv32 = 0u - sk->p_n0inv;
  sk->q_inv32 = 0u - sk->q_n0inv;
  sk->n_inv32 = 0u - sk->n_n0inv;
  sk->pm1_bits = hbn::bitlen(pm1);
  sk->qm1_bits = hbn::bitlen(qm1);
  {
    std::vector<uint8_t> sp = build_schedule(pm1, kSchedWindow);
    std::vector<uint8_t> sq = build_schedule(qm1, kSchedWindow);
#ifdef IPCLB200_EXPERIMENTS
    std::vector<uint8_t> pp = build_tile_program(sp), pq = build_tile_program(sq);
#else
    std::vector<uint8_t> pp, pq;
#endif
    std::vector<uint8_t> sl = build_schedule(lam, kSchedWindow);
    std::vector<uint8_t> both(sp);
    both.insert(both.end(), sq.begin(), sq.end());
    both.insert(both.end(), pp.begin(), pp.end());
    both.insert(both.end(), pq.begin(), pq.end());
    both.insert(both.end(), sl.begin(), sl.end());
    CUDA_TRY(cudaMalloc(&sk->d_sched, both.size()));
    CUDA_TRY(cudaMemcpy(sk->d_sched, both.data(), both.size(), cudaMemcpyHostToDevice));
    sk->d_sched_p = sk->d_sched;
    sk->d_sched_q = sk->d_sched + sp.size();
    sk->d_sched_lambda = sk->d_sched_q + sq.size() + pp.size() + pq.size();
#ifdef IPCLB200_EXPERIMENTS
    sk->d_prog_p = sk->d_sched_q + sq.size();
    sk->d_prog_q = sk->d_prog_p + pp.size();
    sk->tile_slots = sp[0] + 1;
    Limbs two256 = hbn::pow2(256), inv;
    hbn::modinv(hbn::mod(psq, two256), two256, &inv);
    hbn::to_words(hbn::sub(two256, inv), sk->ninv_p, 8);
    hbn::modinv(hbn::mod(qsq, two256), two256, &inv);
    hbn::to_words(hbn::sub(two256, inv), sk->ninv_q, 8);
#endif
    // two-digit decrypt: needs primes that fill their words (so that a digit
    // < R is < 2p) and a prime width decrypt_hensel_kernel is instantiated for
    sk->hensel_ok = hbn::bitlen(p) == 32 * pl && hbn::bitlen(q) == 32 * pl &&
                    (pl == 16 || pl == 32 || pl == 48 || pl == 64);
    if (sk->hensel_ok) {
      std::vector<uint32_t> hs_p = hensel_schedule(sp), hs_q = hensel_schedule(sq);
      std::vector<uint32_t> blk(20 * (size_t)pl + hs_p.size() + hs_q.size(), 0u);
      hensel_side_block(p, psq, hp, pl, blk.data());
      hensel_side_block(q, qsq, hq, pl, blk.data() + 10 * (size_t)pl);
      std::copy(hs_p.begin(), hs_p.end(), blk.begin() + 20 * (size_t)pl);
      std::copy(hs_q.begin(), hs_q.end(), blk.begin() + 20 * (size_t)pl + hs_p.size());
      CUDA_TRY(cudaMalloc(&sk->d_hensel, blk.size() * sizeof(uint32_t)));
      CUDA_TRY(cudaMemcpy(sk->d_hensel, blk.data(), blk.size() * sizeof(uint32_t),
                          cudaMemcpyHostToDevice));
      sk->d_hblk_p = sk->d_hensel;
      sk->d_hblk_q = sk->d_hensel + 10 * (size_t)pl;
      sk->d_hsched_p = sk->d_hensel + 20 * (size_t)pl;
      sk->d_hsched_q = sk->d_hsched_p + hs_p.size();
    }
  }
  sk->lambda_bits = hbn::bitlen(lam);
#ifdef IPCLB200_EXPERIMENTS
  if (sk->L == kFpWords) {
    // radix-2^22 constants of the FP64 role: n and R^3 mod n, R = 2^(22*96)
    constexpr int FL = kFpLimbs;
    std::vector<double> blk(4 * (size_t)FL);
    Limbs Rf = hbn::pow2((unsigned)(kFpW * FL));
    const Limbs* mods[2] = {&psq, &qsq};
    for (int i = 0; i < 2; i++) {
      const Limbs& m = *mods[i];
      Limbs r1 = hbn::mod(Rf, m);
      Limbs r3 = hbn::mod(hbn::mul(hbn::mod(hbn::mul(r1, r1), m), r1), m);
      fp_limbs(m, blk.data() + (size_t)(2 * i) * FL, FL);
      fp_limbs(r3, blk.data() + (size_t)(2 * i + 1) * FL, FL);
    }
    CUDA_TRY(cudaMalloc(&sk->d_fp, blk.size() * sizeof(double)));
    CUDA_TRY(cudaMemcpy(sk->d_fp, blk.data(), blk.size() * sizeof(double),
                        cudaMemcpyHostToDevice));
    sk->fp0.n = sk->d_fp;
    sk->fp0.r3 = sk->d_fp + FL;
    sk->fp0.n32 = sk->mp2->mc.n;
    sk->fp0.n0inv = hbn::neg_inv32(psq[0]) & kFpMask;
    sk->fp1.n = sk->d_fp + 2 * FL;
    sk->fp1.r3 = sk->d_fp + 3 * FL;
    sk->fp1.n32 = sk->mq2->mc.n;
    sk->fp1.n0inv = hbn::neg_inv32(qsq[0]) & kFpMask;
    sk->fp_ok = true;
  }
#endif
  *out = sk.release();
  return 0;
}

void ipclb200_privkey_destroy(ipclb200_privkey* sk) {
  if (!sk) return;
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  if (g_ctx.ready) {
    cudaSetDevice(g_ctx.device);
    cudaDeviceSynchronize();
  }
  delete sk;
}

int ipclb200_decrypt(const ipclb200_privkey* sk, const uint32_t* ct,
                     size_t count, int use_crt, uint32_t* pt) {
  if (!sk || !ct || !pt) return fail(IPCLB200_ERR_BAD_ARG, "decrypt: null pointer");
  if (count == 0) return 0;
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  TRY(ensure_init_locked());
  cudaStream_t s = g_ctx.stream;
  const int pl = sk->pl;
  const int CW = 4 * pl;  // caller's ciphertext words
  uint32_t *d_ct, *d_x, *d_pt;
  // the kernels read a ciphertext as 2L (CRT, L = class of p^2) or L (RAW,
  // L = class of n^2) words; zero padding keeps the value
  const int ctw = use_crt ? 2 * sk->L : sk->mnsq->L;
  const int xw = use_crt ? 2 * sk->L : sk->mnsq->L;
  TRY(scratch_get(0, count * (size_t)ctw, &d_ct));
  TRY(scratch_get(1, count * (size_t)xw, &d_x));
  TRY(scratch_get(2, count * (size_t)(2 * pl), &d_pt));
  TRY(upload_padded(d_ct, ct, CW, ctw, count, s));
  TRY(decrypt_dev_impl(sk, d_ct, count, use_crt, d_pt, d_x, s));
  CUDA_TRY(cudaMemcpyAsync(pt, d_pt, count * (size_t)(2 * pl) * 4,
                           cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return 0;
}

int ipclb200_crt_residues(const ipclb200_privkey* sk, const uint32_t* ct,
                          size_t count, uint32_t* x, int* x_words) {
  if (!sk || !ct || !x) return fail(IPCLB200_ERR_BAD_ARG, "crt_residues: null pointer");
  if (x_words) *x_words = sk->L;
  if (count == 0) return 0;
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  TRY(ensure_init_locked());
  cudaStream_t s = g_ctx.stream;
  const int pl = sk->pl;
  uint32_t *d_ct, *d_x, *d_pt;
  TRY(scratch_get(0, count * (size_t)(2 * sk->L), &d_ct));
  TRY(scratch_get(1, count * (size_t)(2 * sk->L), &d_x));
  TRY(scratch_get(2, count * (size_t)(2 * pl), &d_pt));
  TRY(upload_padded(d_ct, ct, 4 * pl, 2 * sk->L, count, s));
  TRY(decrypt_dev_impl(sk, d_ct, count, 1, d_pt, d_x, s));
  CUDA_TRY(cudaMemcpyAsync(x, d_x, count * (size_t)(2 * sk->L) * 4,
                           cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return 0;
}

int ipclb200_decrypt_dev(const ipclb200_privkey* sk, const uint32_t* d_ct,
                         size_t count, int use_crt, uint32_t* d_pt,
                         void* stream) {
  if (!sk || !d_ct || !d_pt) return fail(IPCLB200_ERR_BAD_ARG, "decrypt_dev: null pointer");
  if (count == 0) return 0;
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  TRY(ensure_init_locked());
  const int pl = sk->pl;
  if ((use_crt && sk->L != 2 * pl) || (!use_crt && sk->mnsq->L != 4 * pl))
    return fail(IPCLB200_ERR_UNSUPPORTED, "decrypt_dev: key width is not a kernel size class");
  uint32_t* d_x;
  TRY(scratch_get(6, count * (size_t)(4 * pl), &d_x));
  return decrypt_dev_impl(sk, d_ct, count, use_crt, d_pt, d_x, (cudaStream_t)stream);
}

// ---- device-resident batches ------------------------------------------------
void* ipclb200_stream(void) {
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  if (ensure_init_locked() != 0) return nullptr;
  return (void*)g_ctx.stream;
}

int ipclb200_dev_alloc(size_t bytes, void** d_out) {
  if (!d_out) return fail(IPCLB200_ERR_BAD_ARG, "dev_alloc: null pointer");
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  TRY(ensure_init_locked());
  CUDA_TRY(cudaMallocAsync(d_out, bytes ? bytes : 4, g_ctx.stream));
  return 0;
}

int ipclb200_dev_free(void* d) {
  if (!d) return 0;
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  if (!g_ctx.ready) return 0;  // the context (and its pool) is already gone
  CUDA_TRY(cudaSetDevice(g_ctx.device));
  CUDA_TRY(cudaFreeAsync(d, g_ctx.stream));
  return 0;
}

int ipclb200_dev_upload(void* d, const void* h, size_t bytes) {
  if (!d || !h) return fail(IPCLB200_ERR_BAD_ARG, "dev_upload: null pointer");
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  TRY(ensure_init_locked());
  // pageable source: the call returns once the data is staged, so the caller
  // may reuse h; pinned source: wait for the copy
  CUDA_TRY(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, g_ctx.stream));
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, h) == cudaSuccess &&
      at.type == cudaMemoryTypeHost)
    CUDA_TRY(cudaStreamSynchronize(g_ctx.stream));
  cudaGetLastError();
  return 0;
}

int ipclb200_dev_download(void* h, const void* d, size_t bytes) {
  if (!d || !h) return fail(IPCLB200_ERR_BAD_ARG, "dev_download: null pointer");
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  TRY(ensure_init_locked());
  CUDA_TRY(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, g_ctx.stream));
  CUDA_TRY(cudaStreamSynchronize(g_ctx.stream));
  return 0;
}

int ipclb200_dev_copy(void* d_dst, const void* d_src, size_t bytes) {
  if (!d_dst || !d_src) return fail(IPCLB200_ERR_BAD_ARG, "dev_copy: null pointer");
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  TRY(ensure_init_locked());
  CUDA_TRY(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, g_ctx.stream));
  return 0;
}

int ipclb200_sync(void) {
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  TRY(ensure_init_locked());
  CUDA_TRY(cudaStreamSynchronize(g_ctx.stream));
  return 0;
}

int ipclb200_class_words(int words) { return words > 0 ? class_words(words) : 0; }

// ---- measurement ----------------------------------------------------------
int ipclb200_int_peak(double* mac32_per_s, double* sm_clock_mhz) {
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  TRY(ensure_init_locked());
  cudaStream_t s = g_ctx.stream;
  const int threads = 256, blocks = g_ctx.sms * 8;
  uint32_t* d_out;
  TRY(scratch_get(7, (size_t)threads * blocks, &d_out));
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
  g_ctx.launches += 6;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  double macs = (double)threads * blocks * kPeakIters * 16.0;
  if (mac32_per_s) *mac32_per_s = macs / (best * 1e-3);
  if (sm_clock_mhz) {
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, g_ctx.device);
    *sm_clock_mhz = khz / 1000.0;
  }
  return 0;
}

int ipclb200_int_peak_sustained(double seconds, double* mac32_per_s) {
  if (!mac32_per_s || seconds <= 0 || seconds > 20)
    return fail(IPCLB200_ERR_BAD_ARG, "int_peak_sustained: bad argument");
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  TRY(ensure_init_locked());
  cudaStream_t s = g_ctx.stream;
  const int threads = 256, blocks = g_ctx.sms * 8;
  uint32_t* d_out;
  TRY(scratch_get(7, (size_t)threads * blocks, &d_out));
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
  g_ctx.launches += launches + 1;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  *mac32_per_s = (double)threads * blocks * kPeakIters * 16.0 * launches / (ms * 1e-3);
  return 0;
}

int ipclb200_debug_montsqr(const uint32_t* a, const uint32_t* mod, size_t count,
                           uint32_t* out_sqr, uint32_t* out_mul) {
#ifndef IPCLB200_EXPERIMENTS
  return fail(IPCLB200_ERR_UNSUPPORTED, "debug_montsqr: built without -DIPCLB200_EXPERIMENTS");
#else
  if (!a || !mod || !out_sqr || !out_mul)
    return fail(IPCLB200_ERR_BAD_ARG, "debug_montsqr: null pointer");
  if (count == 0) return 0;
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  TRY(ensure_init_locked());
  cudaStream_t s = g_ctx.stream;
  const int L = 64;
  Limbs n;
  TRY(check_modulus(mod, L, &n));
  std::shared_ptr<DevModulus> dm;
  TRY(make_modulus(n, L, &dm));
  uint32_t *d_a, *d_s, *d_m;
  TRY(scratch_get(0, count * (size_t)L, &d_a));
  TRY(scratch_get(1, count * (size_t)L, &d_s));
  TRY(scratch_get(2, count * (size_t)L, &d_m));
  CUDA_TRY(cudaMemcpyAsync(d_a, a, count * (size_t)L * 4, cudaMemcpyHostToDevice, s));
  MontSqrTestParams p{};
  p.a = d_a;
  p.n = dm->mc.n;
  p.n0inv = dm->mc.n0inv;
  p.out_sqr = d_s;
  p.out_mul = d_m;
  p.count = count;
  const char* lay = getenv("IPCLB200_DEBUG_SQR_LAYOUT");
  if (lay && lay[0] == '2') {
    constexpr size_t smem = sqr2_smem_bytes(kBlockThreads);
    CUDA_TRY(cudaFuncSetAttribute(montsqr2_test_kernel,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)((count + 63) / 64);
    montsqr2_test_kernel<<<grid, kBlockThreads, smem, s>>>(p);
  } else {
    constexpr size_t smem = sqr_smem_bytes(kBlockThreads, 4);
    CUDA_TRY(cudaFuncSetAttribute(montsqr_test_kernel,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)((count + 31) / 32);
    montsqr_test_kernel<<<grid, kBlockThreads, smem, s>>>(p);
  }
  g_ctx.launches++;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(out_sqr, d_s, count * (size_t)L * 4, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(out_mul, d_m, count * (size_t)L * 4, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return 0;
#endif
}

int ipclb200_pipe_mix(int mode, double* ms_out) {
#ifndef IPCLB200_EXPERIMENTS
  return fail(IPCLB200_ERR_UNSUPPORTED, "pipe_mix: built without -DIPCLB200_EXPERIMENTS");
#else
  if (mode < 0 || mode > 8 || !ms_out)
    return fail(IPCLB200_ERR_BAD_ARG, "pipe_mix: bad argument");
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  TRY(ensure_init_locked());
  cudaStream_t s = g_ctx.stream;
  const int threads = 256, blocks = g_ctx.sms * 4, iters = 4096;
  uint32_t* d_out;
  TRY(scratch_get(7, (size_t)threads * blocks, &d_out));
  cudaEvent_t a, b;
  CUDA_TRY(cudaEventCreate(&a));
  CUDA_TRY(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CUDA_TRY(cudaEventRecord(a, s));
    pipe_mix_kernel<<<blocks, threads, 0, s>>>(d_out, mode, iters, 3u + rep, 1.5);
    CUDA_TRY(cudaEventRecord(b, s));
    CUDA_TRY(cudaEventSynchronize(b));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0 && ms < best) best = ms;
  }
  g_ctx.launches += 4;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  *ms_out = best;
  return 0;
#endif
}

}  // extern "C"
