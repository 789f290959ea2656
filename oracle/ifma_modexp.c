/*
 * ifma_modexp.c -- TEST/BENCH INFRASTRUCTURE ONLY (never linked into the
 * product).  A labelled RESTATEMENT of the CPU path the reference runs:
 * ipcl::ippModExp -> ippMBModExpWrapper -> mbx_exp_mb8 (ipcl/mod_exp.cpp:
 * 597-636, 446-533).  The genuine mbx_exp_mb8 lives in intel/ipp-crypto
 * (crypto_mb, tag ippcp_2021.6), which is not in /root/reference and cannot be
 * fetched here; this file restates its published algorithm so that bench.py
 * can time "the reference's AVX512-IFMA path" on the GPU box's host cores:
 *
 *   - 8 independent modexps per call, one per 64-bit lane of a 512-bit register
 *     ("multi-buffer"), operands transposed to digit-major layout;
 *   - radix 2^52 digits, products with vpmadd52luq / vpmadd52huq;
 *   - almost-Montgomery multiplication (inputs and outputs < 2n, R = 2^(52 len)
 *     >= 4n) with lazily carried 64-bit digit accumulators;
 *   - fixed 5-bit window exponentiation with a 32-entry table;
 *   - the caller splits a batch into chunks of 8 under `omp parallel for`,
 *     exactly like ippMBModExpWrapper.
 *
 * Numbers produced here are "restated", never Intel's.  Outputs are bit-exact
 * with paillier_oracle.c (tests/test_oracle.py::test_mb8_matches_scalar).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint32_t u32;
typedef uint64_t u64;

#ifdef ORC_HAVE_IFMA
#include <immintrin.h>

#define MB8 8
#define DIGIT_BITS 52
#define DIGIT_MASK ((1ull << DIGIT_BITS) - 1)
#define MAX_DIGITS 160 /* 8192-bit moduli */
#define WIN 5

typedef __m512i V;

/* provided by paillier_oracle.c */
extern int orc_radix52_constants(const u32* mod, int L, int len, u64* rr52,
                                 u64* k0);

static inline int digits_for(int L) { return (32 * L + 2 + DIGIT_BITS - 1) / DIGIT_BITS; }

/* little-endian u32 limbs -> 52-bit digits */
static void to_digits(u64* d, int len, const u32* x, int L) {
  for (int j = 0; j < len; j++) {
    int bit = j * DIGIT_BITS;
    int w = bit >> 5, sh = bit & 31;
    /* gather up to 96 bits */
    u64 lo = 0, hi = 0;
    if (w < L) lo = x[w];
    if (w + 1 < L) lo |= (u64)x[w + 1] << 32;
    if (w + 2 < L) hi = x[w + 2];
    u64 v = sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
    d[j] = v & DIGIT_MASK;
  }
}

static void from_digits(u32* x, int L, const u64* d, int len) {
  memset(x, 0, sizeof(u32) * (size_t)L);
  for (int j = 0; j < len; j++) {
    int bit = j * DIGIT_BITS;
    for (int k = 0; k < DIGIT_BITS; k += 1) {
      /* bit-serial is fine: runs once per modexp */
      int b = bit + k;
      if ((b >> 5) >= L) break;
      x[b >> 5] |= (u32)((d[j] >> k) & 1u) << (b & 31);
    }
  }
}

/* r = a*b*R^-1 (almost reduced, digits normalised).  a, b, n: len digit
 * vectors with every digit < 2^52; k0 = -n^-1 mod 2^52 per lane. */
static void amm8(V* r, const V* a, const V* b, const V* n, V k0, int len) {
  V acc[MAX_DIGITS + 1];
  const V zero = _mm512_setzero_si512();
  for (int j = 0; j <= len; j++) acc[j] = zero;
  for (int i = 0; i < len; i++) {
    const V bi = b[i];
    V t0 = _mm512_madd52lo_epu64(acc[0], a[0], bi);
    const V y = _mm512_madd52lo_epu64(zero, t0, k0);
    t0 = _mm512_madd52lo_epu64(t0, n[0], y);
    V carry = _mm512_srli_epi64(t0, DIGIT_BITS);
    for (int j = 1; j < len; j++) {
      V t = acc[j];
      t = _mm512_madd52lo_epu64(t, a[j], bi);
      t = _mm512_madd52lo_epu64(t, n[j], y);
      t = _mm512_madd52hi_epu64(t, a[j - 1], bi);
      t = _mm512_madd52hi_epu64(t, n[j - 1], y);
      acc[j - 1] = t;
    }
    acc[0] = _mm512_add_epi64(acc[0], carry);
    V t = acc[len];
    t = _mm512_madd52hi_epu64(t, a[len - 1], bi);
    t = _mm512_madd52hi_epu64(t, n[len - 1], y);
    acc[len - 1] = t;
    acc[len] = zero;
  }
  /* normalise the lazily carried digits */
  const V mask = _mm512_set1_epi64((long long)DIGIT_MASK);
  V c = zero;
  for (int j = 0; j < len; j++) {
    V t = _mm512_add_epi64(acc[j], c);
    r[j] = _mm512_and_si512(t, mask);
    c = _mm512_srli_epi64(t, DIGIT_BITS);
  }
}

static u32 window_of(const u32* e, int EL, int bitpos) {
  u32 v = 0;
  for (int k = 0; k < WIN; k++) {
    int b = bitpos + k;
    if (b < EL * 32) v |= ((e[b >> 5] >> (b & 31)) & 1u) << k;
  }
  return v;
}

/* 8 modexps with one shared modulus.  nlanes <= 8 real elements. */
static void modexp_mb8(u32* out, const u32* base, size_t base_stride,
                       const u32* exp, size_t exp_stride, int EL, int ebits,
                       const u32* mod, int L, const u64* rr52, u64 k0s,
                       int nlanes) {
  const int len = digits_for(L);
  u64 tmp[MB8][MAX_DIGITS];
  V n[MAX_DIGITS], rr[MAX_DIGITS], x[MAX_DIGITS], acc[MAX_DIGITS], one[MAX_DIGITS];
  u64 nd[MAX_DIGITS];
  to_digits(nd, len, mod, L);
  for (int j = 0; j < len; j++) {
    n[j] = _mm512_set1_epi64((long long)nd[j]);
    rr[j] = _mm512_set1_epi64((long long)rr52[j]);
    one[j] = _mm512_setzero_si512();
  }
  one[0] = _mm512_set1_epi64(1);
  const V k0 = _mm512_set1_epi64((long long)k0s);
  /* transpose the bases into digit-major vectors */
  memset(tmp, 0, sizeof(tmp));
  for (int l = 0; l < nlanes; l++) to_digits(tmp[l], len, base + l * base_stride, L);
  for (int j = 0; j < len; j++) {
    u64 lane[MB8];
    for (int l = 0; l < MB8; l++) lane[l] = tmp[l][j];
    x[j] = _mm512_loadu_si512((const void*)lane);
  }
  /* window table: tab[k] = base^k in Montgomery form */
  V(*tab)[MAX_DIGITS] = (V(*)[MAX_DIGITS])aligned_alloc(64, sizeof(V[MAX_DIGITS]) << WIN);
  amm8(tab[0], one, rr, n, k0, len); /* R mod n */
  amm8(tab[1], x, rr, n, k0, len);   /* base * R */
  for (int k = 2; k < (1 << WIN); k++) amm8(tab[k], tab[k - 1], tab[1], n, k0, len);
  int nwin = (ebits + WIN - 1) / WIN;
  if (nwin < 1) nwin = 1;
  for (int k = nwin - 1; k >= 0; k--) {
    /* per-lane table gather */
    u32 w[MB8];
    for (int l = 0; l < MB8; l++)
      w[l] = l < nlanes ? window_of(exp + l * exp_stride, EL, k * WIN) : 0;
    V sel[MAX_DIGITS];
    for (int j = 0; j < len; j++) {
      u64 lane[MB8];
      for (int l = 0; l < MB8; l++) lane[l] = ((const u64*)&tab[w[l]][j])[l];
      sel[j] = _mm512_loadu_si512((const void*)lane);
    }
    if (k == nwin - 1) {
      memcpy(acc, sel, sizeof(V) * (size_t)len);
    } else {
      for (int s = 0; s < WIN; s++) amm8(acc, acc, acc, n, k0, len);
      amm8(acc, acc, sel, n, k0, len);
    }
  }
  amm8(acc, acc, one, n, k0, len); /* leave Montgomery form: value <= n */
  for (int j = 0; j < len; j++) {
    u64 lane[MB8];
    _mm512_storeu_si512((void*)lane, acc[j]);
    for (int l = 0; l < MB8; l++) tmp[l][j] = lane[l];
  }
  for (int l = 0; l < nlanes; l++) {
    u32* o = out + (size_t)l * L;
    from_digits(o, L, tmp[l], len);
    /* canonical: subtract n once if the value equals n (it is <= n) */
    int ge = 1;
    for (int i = L - 1; i >= 0; i--)
      if (o[i] != mod[i]) {
        ge = o[i] > mod[i];
        break;
      }
    if (ge) {
      u64 br = 0;
      for (int i = 0; i < L; i++) {
        u64 t = (u64)o[i] - mod[i] - br;
        o[i] = (u32)t;
        br = (t >> 32) & 1;
      }
    }
  }
  free(tab);
}

int orc_have_ifma(void) { return __builtin_cpu_supports("avx512ifma") ? 1 : 0; }

/* Batched modexp, shared odd modulus: the shape of every ippModExp call on
 * the Paillier path.  base_stride/exp_stride in words (0 = one shared value).
 * Chunks of 8 under OpenMP (ippMBModExpWrapper, mod_exp.cpp:597-636). */
int orc_modexp_mb8(const u32* base, size_t base_stride, const u32* exp,
                   size_t exp_stride, int EL, const u32* mod, int L,
                   size_t count, u32* out) {
  if (L <= 0 || digits_for(L) > MAX_DIGITS) return -2;
  if (!(mod[0] & 1)) return -1;
  const int len = digits_for(L);
  u64 rr52[MAX_DIGITS], k0;
  int rc = orc_radix52_constants(mod, L, len, rr52, &k0);
  if (rc) return rc;
  /* exponent bit length over the batch (mod_exp.cpp:480-484) */
  int ebits = 0;
  size_t ne = exp_stride ? count : 1;
  for (size_t i = 0; i < ne; i++)
    for (int w = EL - 1; w >= 0; w--)
      if (exp[i * exp_stride + w]) {
        int b = w * 32 + 32 - __builtin_clz(exp[i * exp_stride + w]);
        if (b > ebits) ebits = b;
        break;
      }
  size_t chunks = (count + MB8 - 1) / MB8;
#pragma omp parallel for schedule(dynamic, 1)
  for (size_t c = 0; c < chunks; c++) {
    size_t off = c * MB8;
    int nl = (int)((count - off) < MB8 ? (count - off) : MB8);
    u32 bbuf[MB8 * 512], ebuf[MB8 * 512];
    /* materialise shared operands per lane so the kernel sees plain strides */
    for (int l = 0; l < nl; l++) {
      memcpy(bbuf + (size_t)l * L, base + (off + l) * base_stride, sizeof(u32) * (size_t)L);
      memcpy(ebuf + (size_t)l * EL, exp + (off + l) * exp_stride, sizeof(u32) * (size_t)EL);
    }
    modexp_mb8(out + off * (size_t)L, bbuf, (size_t)L, ebuf, (size_t)EL, EL, ebits,
               mod, L, rr52, k0, nl);
  }
  return 0;
}

#else /* !ORC_HAVE_IFMA */
int orc_have_ifma(void) { return 0; }
int orc_modexp_mb8(const u32* base, size_t base_stride, const u32* exp,
                   size_t exp_stride, int EL, const u32* mod, int L,
                   size_t count, u32* out) {
  (void)base; (void)base_stride; (void)exp; (void)exp_stride; (void)EL;
  (void)mod; (void)L; (void)count; (void)out;
  return -3;
}
#endif
