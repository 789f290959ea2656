/*
 * paillier_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement, in plain C, of the one hot path of intel/pailliercryptolib
 * (IPCL v2.0.0): batched modular exponentiation and the Paillier arithmetic
 * wrapped round it.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this file's shared object.
 *
 * Parity status: PINNED.  The arithmetic the reference calls (mbx_exp_mb8,
 * ippsMontExp, ipps*_BN) lives in intel/ipp-crypto tag ippcp_2021.6 (v11.4),
 * which is not vendored in /root/reference and cannot be fetched here
 * (cmake/ippcrypto.cmake:7-8,36).  Every function on the path returns the
 * canonical residue in [0, m), so any correct implementation is bit-identical;
 * this restatement is pinned against the reference's own known-answer test
 * ISO_IEC_18033_6_ComplianceTest (test/test_cryptography.cpp:99-241: c1, c2,
 * c1c2, m1m2) and against Python pow()/% on committed golden vectors
 * (tests/golden/, tests/test_oracle.py).
 *
 * Algorithm restated for modexp: the fixed-window Montgomery exponentiation
 * published for ipp-crypto's mbx_exp{1024,2048,3072,4096}_mb8 -- Montgomery
 * form via R^2 mod m, a 2^w-entry table of base powers, w squarings and one
 * table multiply per exponent window scanned from the top, leave Montgomery
 * form by multiplying with 1, canonical final subtraction -- here in radix
 * 2^32 CIOS instead of 8-lane radix 2^52 (same function, same outputs).
 *
 * Data layout everywhere: little-endian arrays of 32-bit words (the layout
 * ippsRef_BN exposes, ipcl/mod_exp.cpp:472-476), fixed stride per element.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint32_t u32;
typedef uint64_t u64;

#define ORC_MAX_WORDS 512 /* up to 16384-bit intermediates */

/* ------------------------------------------------------------------ */
/* multi-word helpers                                                  */
/* ------------------------------------------------------------------ */
static int bn_is_zero(const u32* a, int n) {
  for (int i = 0; i < n; i++)
    if (a[i]) return 0;
  return 1;
}
static int bn_cmp(const u32* a, const u32* b, int n) {
  for (int i = n - 1; i >= 0; i--) {
    if (a[i] > b[i]) return 1;
    if (a[i] < b[i]) return -1;
  }
  return 0;
}
static int bn_words(const u32* a, int n) {
  while (n > 0 && a[n - 1] == 0) n--;
  return n;
}
static u32 bn_add(u32* r, const u32* a, const u32* b, int n) {
  u64 c = 0;
  for (int i = 0; i < n; i++) {
    c += (u64)a[i] + b[i];
    r[i] = (u32)c;
    c >>= 32;
  }
  return (u32)c;
}
static u32 bn_sub(u32* r, const u32* a, const u32* b, int n) {
  u64 br = 0;
  for (int i = 0; i < n; i++) {
    u64 t = (u64)a[i] - b[i] - br;
    r[i] = (u32)t;
    br = (t >> 32) & 1;
  }
  return (u32)br;
}
static u32 bn_sub_word(u32* r, const u32* a, u32 w, int n) {
  u64 br = w;
  for (int i = 0; i < n; i++) {
    u64 t = (u64)a[i] - br;
    r[i] = (u32)t;
    br = (t >> 32) & 1;
  }
  return (u32)br;
}
/* r[na+nb] = a[na] * b[nb]; r must not alias a or b */
static void bn_mul(u32* r, const u32* a, int na, const u32* b, int nb) {
  memset(r, 0, sizeof(u32) * (size_t)(na + nb));
  for (int i = 0; i < na; i++) {
    u64 c = 0;
    u64 ai = a[i];
    for (int j = 0; j < nb; j++) {
      c += ai * b[j] + r[i + j];
      r[i + j] = (u32)c;
      c >>= 32;
    }
    r[i + nb] = (u32)c;
  }
}
/* Knuth algorithm D.  q[na] (may be NULL), r[nd] <- a[na] divmod d[nd]; d != 0 */
static void bn_divmod(u32* q, u32* r, const u32* a, int na, const u32* d,
                      int nd) {
  int n = bn_words(d, nd);
  int m = bn_words(a, na);
  if (q) memset(q, 0, sizeof(u32) * (size_t)na);
  memset(r, 0, sizeof(u32) * (size_t)nd);
  if (m < n) {
    memcpy(r, a, sizeof(u32) * (size_t)m);
    return;
  }
  if (n == 1) {
    u64 rem = 0;
    for (int i = m - 1; i >= 0; i--) {
      u64 cur = (rem << 32) | a[i];
      if (q) q[i] = (u32)(cur / d[0]);
      rem = cur % d[0];
    }
    r[0] = (u32)rem;
    return;
  }
  u32* un = (u32*)malloc(sizeof(u32) * (size_t)(m + 1));
  u32* vn = (u32*)malloc(sizeof(u32) * (size_t)n);
  int s = __builtin_clz(d[n - 1]);
  for (int i = n - 1; i > 0; i--)
    vn[i] = s ? (d[i] << s) | (d[i - 1] >> (32 - s)) : d[i];
  vn[0] = d[0] << s;
  un[m] = s ? a[m - 1] >> (32 - s) : 0;
  for (int i = m - 1; i > 0; i--)
    un[i] = s ? (a[i] << s) | (a[i - 1] >> (32 - s)) : a[i];
  un[0] = a[0] << s;
  for (int j = m - n; j >= 0; j--) {
    u64 num = ((u64)un[j + n] << 32) | un[j + n - 1];
    u64 qhat = num / vn[n - 1];
    u64 rhat = num % vn[n - 1];
    while (qhat >= (1ull << 32) ||
           qhat * vn[n - 2] > ((rhat << 32) | un[j + n - 2])) {
      qhat--;
      rhat += vn[n - 1];
      if (rhat >= (1ull << 32)) break;
    }
    int64_t borrow = 0;
    u64 carry = 0;
    for (int i = 0; i < n; i++) {
      u64 p = qhat * vn[i] + carry;
      carry = p >> 32;
      int64_t t = (int64_t)un[i + j] - borrow - (int64_t)(p & 0xffffffffu);
      un[i + j] = (u32)t;
      borrow = (t < 0) ? 1 : 0;
    }
    int64_t t = (int64_t)un[j + n] - borrow - (int64_t)carry;
    un[j + n] = (u32)t;
    if (t < 0) { /* add back */
      qhat--;
      u64 c = 0;
      for (int i = 0; i < n; i++) {
        c += (u64)un[i + j] + vn[i];
        un[i + j] = (u32)c;
        c >>= 32;
      }
      un[j + n] += (u32)c;
    }
    if (q) q[j] = (u32)qhat;
  }
  for (int i = 0; i < n - 1; i++)
    r[i] = s ? (un[i] >> s) | (un[i + 1] << (32 - s)) : un[i];
  r[n - 1] = un[n - 1] >> s;
  free(un);
  free(vn);
}
/* r[nd] = a[na] mod d[nd] */
static void bn_mod(u32* r, const u32* a, int na, const u32* d, int nd) {
  bn_divmod(NULL, r, a, na, d, nd);
}
/* r[L] = a[L]*b[L] mod m[L]  (plain product then division: BigNumber a*b%m,
 * ipcl/bignum.cpp:198-209,304-308) */
static void bn_modmul(u32* r, const u32* a, const u32* b, const u32* m, int L) {
  u32 t[2 * ORC_MAX_WORDS];
  bn_mul(t, a, L, b, L);
  bn_mod(r, t, 2 * L, m, L);
}

/* ------------------------------------------------------------------ */
/* Montgomery arithmetic, radix 2^32, CIOS                             */
/* ------------------------------------------------------------------ */
static u32 mont_n0inv(u32 n0) { /* -n0^{-1} mod 2^32, n0 odd */
  u32 x = n0;                   /* 3 correct bits */
  for (int i = 0; i < 5; i++) x *= 2 - n0 * x;
  return (u32)(0u - x);
}
/* r = a*b*R^{-1} mod n, fully reduced if a*b < n*R; r may alias a or b */
static void mont_mul(u32* r, const u32* a, const u32* b, const u32* n,
                     u32 n0inv, int L) {
  u32 t[ORC_MAX_WORDS + 2];
  memset(t, 0, sizeof(u32) * (size_t)(L + 2));
  for (int i = 0; i < L; i++) {
    u64 c = 0, bi = b[i];
    for (int j = 0; j < L; j++) {
      c += (u64)a[j] * bi + t[j];
      t[j] = (u32)c;
      c >>= 32;
    }
    c += t[L];
    t[L] = (u32)c;
    t[L + 1] = (u32)(c >> 32);
    u64 q = (u32)(t[0] * n0inv);
    c = (q * n[0] + t[0]) >> 32;
    for (int j = 1; j < L; j++) {
      c += q * n[j] + t[j];
      t[j - 1] = (u32)c;
      c >>= 32;
    }
    c += t[L];
    t[L - 1] = (u32)c;
    t[L] = t[L + 1] + (u32)(c >> 32);
  }
  if (t[L] || bn_cmp(t, n, L) >= 0) bn_sub(t, t, n, L);
  memcpy(r, t, sizeof(u32) * (size_t)L);
}

static u32 exp_window(const u32* e, int EL, int bitpos, int w) {
  /* w bits of e starting at bit `bitpos` (may run past the top: zeros) */
  u32 v = 0;
  for (int k = 0; k < w; k++) {
    int b = bitpos + k;
    if (b < EL * 32 && b >= 0) v |= ((e[b >> 5] >> (b & 31)) & 1u) << k;
  }
  return v;
}

/* One modexp: out[L] = base[L]^exp[EL] mod mod[L].  Follows ippMBModExp /
 * ippSBModExp (ipcl/mod_exp.cpp:446-585): operands zero-padded to the modulus
 * width, result 0 <= out < mod.  base may be >= mod (reduced first).
 * returns 0 ok, -1 mod zero/even (Montgomery needs an odd modulus; every
 * modulus on the path -- n^2, p^2, q^2, p, q, n -- is odd). */
static int modexp_one(u32* out, const u32* base, const u32* exp, int EL,
                      const u32* mod, int L) {
  if (bn_is_zero(mod, L) || !(mod[0] & 1)) return -1;
  const int w = 5;
  u32 n0inv = mont_n0inv(mod[0]);
  u32 rr[ORC_MAX_WORDS], one[ORC_MAX_WORDS], bm[ORC_MAX_WORDS];
  u32 acc[ORC_MAX_WORDS];
  /* RR = 2^(64L) mod m */
  {
    u32 t[2 * ORC_MAX_WORDS + 1];
    memset(t, 0, sizeof(u32) * (size_t)(2 * L + 1));
    t[2 * L] = 1;
    bn_mod(rr, t, 2 * L + 1, mod, L);
  }
  memset(one, 0, sizeof(u32) * (size_t)L);
  one[0] = 1;
  bn_mod(bm, base, L, mod, L);
  u32(*tab)[ORC_MAX_WORDS] = malloc(sizeof(u32[ORC_MAX_WORDS]) << w);
  mont_mul(tab[0], one, rr, mod, n0inv, L); /* R mod m  (Montgomery 1) */
  mont_mul(tab[1], bm, rr, mod, n0inv, L);  /* base in Montgomery form */
  for (int i = 2; i < (1 << w); i++)
    mont_mul(tab[i], tab[i - 1], tab[1], mod, n0inv, L);
  int ebits = EL * 32;
  int nwin = (ebits + w - 1) / w;
  memcpy(acc, tab[exp_window(exp, EL, (nwin - 1) * w, w)],
         sizeof(u32) * (size_t)L);
  for (int k = nwin - 2; k >= 0; k--) {
    for (int s = 0; s < w; s++) mont_mul(acc, acc, acc, mod, n0inv, L);
    mont_mul(acc, acc, tab[exp_window(exp, EL, k * w, w)], mod, n0inv, L);
  }
  mont_mul(out, acc, one, mod, n0inv, L);
  free(tab);
  return 0;
}

/* ------------------------------------------------------------------ */
/* exported batch API (mirrors include/ipcl_b200.h shapes)             */
/* ------------------------------------------------------------------ */

/* ipcl::ippModExp(vector,vector,vector) -- ipcl/mod_exp.cpp:655-678.
 * base: count x L words; exp: count x EL words; mod: count x L words, or a
 * single modulus when shared_mod != 0; out: count x L words. */
int orc_modexp(const u32* base, const u32* exp, const u32* mod, int L, int EL,
               size_t count, int shared_mod, int shared_base, int shared_exp,
               u32* out) {
  if (L <= 0 || L > ORC_MAX_WORDS || EL <= 0 || EL > ORC_MAX_WORDS) return -2;
  int rc = 0;
#pragma omp parallel for schedule(dynamic, 8)
  for (size_t i = 0; i < count; i++) {
    const u32* b = base + (shared_base ? 0 : i * (size_t)L);
    const u32* e = exp + (shared_exp ? 0 : i * (size_t)EL);
    const u32* m = mod + (shared_mod ? 0 : i * (size_t)L);
    int r = modexp_one(out + i * (size_t)L, b, e, EL, m, L);
    if (r) {
#pragma omp atomic write
      rc = r;
    }
  }
  return rc;
}

/* CipherText::raw_add, a*b % n^2 -- ipcl/ciphertext.cpp:135-141 (and the
 * loops :53-69; b_shared = the size-1 broadcast of :51-59). */
int orc_modmul(const u32* a, const u32* b, const u32* mod, int L, size_t count,
               int b_shared, u32* out) {
  if (L <= 0 || L > ORC_MAX_WORDS) return -2;
  if (bn_is_zero(mod, L)) return -1;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < count; i++)
    bn_modmul(out + i * (size_t)L, a + i * (size_t)L,
              b + (b_shared ? 0 : i * (size_t)L), mod, L);
  return 0;
}

/* PublicKey::raw_encrypt + applyObfuscator -- ipcl/pub_key.cpp:82-110.
 *   ct = ((n*pt + 1) mod n^2) * obf mod n^2
 *   DJN     (pub_key.cpp:51-64): obf = hs^r mod n^2   (same base for all)
 *   non-DJN (pub_key.cpp:66-80): obf = r^n  mod n^2   (same exponent for all)
 * n: NL words; nsq/hs/ct: 2*NL words; pt: count x NL words (pt < 2^(32 NL));
 * r: count x RL words.  make_secure == 0 skips the obfuscator
 * (pub_key.cpp:107), which is how ct+pt encodes its plaintext
 * (ciphertext.cpp:75-80). */
int orc_encrypt(const u32* n, int NL, const u32* hs /* NULL = non-DJN */,
                const u32* pt, const u32* r, int RL, size_t count,
                int make_secure, u32* ct) {
  int L = 2 * NL;
  if (NL <= 0 || L > ORC_MAX_WORDS || RL > L) return -2;
  u32 nsq[ORC_MAX_WORDS];
  bn_mul(nsq, n, NL, n, NL);
  int rc = 0;
#pragma omp parallel for schedule(dynamic, 8)
  for (size_t i = 0; i < count; i++) {
    u32 t[2 * ORC_MAX_WORDS + 1], c[ORC_MAX_WORDS], obf[ORC_MAX_WORDS];
    /* (n*pt + 1) % nsq */
    bn_mul(t, n, NL, pt + i * (size_t)NL, NL);
    u64 cy = 1;
    for (int j = 0; j < L && cy; j++) {
      cy += t[j];
      t[j] = (u32)cy;
      cy >>= 32;
    }
    t[L] = (u32)cy;
    bn_mod(c, t, L + 1, nsq, L);
    if (make_secure) {
      int r1;
      if (hs) {
        r1 = modexp_one(obf, hs, r + i * (size_t)RL, RL, nsq, L);
      } else {
        u32 rb[ORC_MAX_WORDS];
        memset(rb, 0, sizeof(u32) * (size_t)L);
        memcpy(rb, r + i * (size_t)RL, sizeof(u32) * (size_t)RL);
        r1 = modexp_one(obf, rb, n, NL, nsq, L);
      }
      if (r1) {
#pragma omp atomic write
        rc = r1;
      }
      bn_modmul(c, c, obf, nsq, L); /* sq.ModMul(ct, obf), pub_key.cpp:88-89 */
    }
    memcpy(ct + i * (size_t)L, c, sizeof(u32) * (size_t)L);
  }
  return rc;
}

/* x^(m-2) mod m for prime m (the value InverseMul returns, bignum.cpp:331-335) */
static void inv_mod_prime(u32* r, const u32* x, const u32* m, int L) {
  u32 e[ORC_MAX_WORDS], xr[ORC_MAX_WORDS];
  bn_sub_word(e, m, 2, L);
  bn_mod(xr, x, L, m, L);
  modexp_one(r, xr, e, L, m, L);
}

/* PrivateKey::computeLfun -- ipcl/pri_key.cpp:154-157: (a-1)/b.
 * a: LA words, b: LB words, result truncated/padded to LR words. */
static void l_fun(u32* r, int LR, const u32* a, int LA, const u32* b, int LB) {
  u32 am1[ORC_MAX_WORDS], q[ORC_MAX_WORDS], rem[ORC_MAX_WORDS];
  bn_sub_word(am1, a, 1, LA);
  bn_divmod(q, rem, am1, LA, b, LB);
  memset(r, 0, sizeof(u32) * (size_t)LR);
  memcpy(r, q, sizeof(u32) * (size_t)(LA < LR ? LA : LR));
}

/* Per-key CRT constants, PrivateKey ctor -- ipcl/pri_key.cpp:13-37,159-167.
 * p,q: PL words each (caller orders p < q as the ctor does, :19-22).
 * outputs: psq,qsq (2*PL), hp,hq,pinv (PL). g = n+1. */
int orc_crt_constants(const u32* p, const u32* q, int PL, u32* psq, u32* qsq,
                      u32* hp, u32* hq, u32* pinv) {
  int L2 = 2 * PL;
  if (PL <= 0 || 2 * L2 > ORC_MAX_WORDS) return -2;
  u32 n[ORC_MAX_WORDS], g[ORC_MAX_WORDS + 1];
  bn_mul(n, p, PL, q, PL);
  bn_mul(psq, p, PL, p, PL);
  bn_mul(qsq, q, PL, q, PL);
  memcpy(g, n, sizeof(u32) * (size_t)L2);
  u64 cy = 1;
  for (int j = 0; j < L2 && cy; j++) {
    cy += g[j];
    g[j] = (u32)cy;
    cy >>= 32;
  }
  const u32* pr[2] = {p, q};
  const u32* sq[2] = {psq, qsq};
  u32* h[2] = {hp, hq};
  for (int k = 0; k < 2; k++) {
    /* computeHfun(a=p, b=p^2): base = g % b; pm = base^(a-1) mod b;
     * lcrt = (pm-1)/a; return a.InverseMul(lcrt) */
    u32 base[ORC_MAX_WORDS], xm[ORC_MAX_WORDS], pm[ORC_MAX_WORDS];
    u32 lc[ORC_MAX_WORDS];
    bn_mod(base, g, L2, sq[k], L2);
    bn_sub_word(xm, pr[k], 1, PL);
    if (modexp_one(pm, base, xm, PL, sq[k], L2)) return -1;
    l_fun(lc, PL, pm, L2, pr[k], PL);
    inv_mod_prime(h[k], lc, pr[k], PL);
  }
  inv_mod_prime(pinv, p, q, PL); /* q.InverseMul(p), pri_key.cpp:27 */
  return 0;
}

/* PrivateKey::decryptCRT -- ipcl/pri_key.cpp:114-152.
 * ct: count x 2*NL words (NL = 2*PL, ct < n^2); pt out: count x NL words. */
int orc_decrypt_crt(const u32* p, const u32* q, int PL, const u32* ct,
                    size_t count, u32* pt) {
  int NL = 2 * PL, CL = 2 * NL;
  if (PL <= 0 || CL > ORC_MAX_WORDS) return -2;
  u32 psq[ORC_MAX_WORDS], qsq[ORC_MAX_WORDS], hp[ORC_MAX_WORDS];
  u32 hq[ORC_MAX_WORDS], pinv[ORC_MAX_WORDS], pm1[ORC_MAX_WORDS];
  u32 qm1[ORC_MAX_WORDS];
  int rc = orc_crt_constants(p, q, PL, psq, qsq, hp, hq, pinv);
  if (rc) return rc;
  bn_sub_word(pm1, p, 1, PL);
  bn_sub_word(qm1, q, 1, PL);
#pragma omp parallel for schedule(dynamic, 8)
  for (size_t i = 0; i < count; i++) {
    const u32* c = ct + i * (size_t)CL;
    u32 bp[ORC_MAX_WORDS], bq[ORC_MAX_WORDS], rp[ORC_MAX_WORDS];
    u32 rq[ORC_MAX_WORDS], lp[ORC_MAX_WORDS], lq[ORC_MAX_WORDS];
    u32 dp[ORC_MAX_WORDS], dq[ORC_MAX_WORDS], u[ORC_MAX_WORDS];
    u32 t[2 * ORC_MAX_WORDS];
    bn_mod(bp, c, CL, psq, NL); /* pri_key.cpp:127-130 */
    bn_mod(bq, c, CL, qsq, NL);
    modexp_one(rp, bp, pm1, PL, psq, NL); /* :133 */
    modexp_one(rq, bq, qm1, PL, qsq, NL); /* :134 */
    l_fun(lp, PL, rp, NL, p, PL);         /* :141-142 */
    l_fun(lq, PL, rq, NL, q, PL);
    bn_modmul(dp, lp, hp, p, PL);
    bn_modmul(dq, lq, hq, q, PL);
    /* computeCRT (:148-152): u = (mq - mp) * pinv mod q (non-negative
     * residue), pt = mp + u*p */
    u32 dpq[ORC_MAX_WORDS], diff[ORC_MAX_WORDS];
    bn_mod(dpq, dp, PL, q, PL);
    if (bn_cmp(dq, dpq, PL) >= 0) {
      bn_sub(diff, dq, dpq, PL);
    } else {
      bn_sub(diff, dq, dpq, PL);
      bn_add(diff, diff, q, PL);
    }
    bn_modmul(u, diff, pinv, q, PL);
    bn_mul(t, u, PL, p, PL);
    u32* o = pt + i * (size_t)NL;
    memset(lp, 0, sizeof(u32) * (size_t)NL);
    memcpy(lp, dp, sizeof(u32) * (size_t)PL);
    bn_add(o, t, lp, NL);
  }
  return 0;
}

/* PrivateKey::decryptRAW -- ipcl/pri_key.cpp:92-111, with lambda and
 * x = (L(g^lambda mod n^2))^{-1} mod n supplied by the caller (ctor :35-37). */
int orc_decrypt_raw(const u32* n, int NL, const u32* lambda, const u32* x,
                    const u32* ct, size_t count, u32* pt) {
  int L = 2 * NL;
  if (NL <= 0 || L > ORC_MAX_WORDS) return -2;
  u32 nsq[ORC_MAX_WORDS];
  bn_mul(nsq, n, NL, n, NL);
  int rc = 0;
#pragma omp parallel for schedule(dynamic, 8)
  for (size_t i = 0; i < count; i++) {
    u32 res[ORC_MAX_WORDS], l[ORC_MAX_WORDS];
    int r1 = modexp_one(res, ct + i * (size_t)L, lambda, NL, nsq, L);
    if (r1) {
#pragma omp atomic write
      rc = r1;
    }
    l_fun(l, NL, res, L, n, NL);
    bn_modmul(pt + i * (size_t)NL, l, x, n, NL);
  }
  return rc;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------ */
/* multi-buffer (8-lane AVX512-IFMA) variants: same functions, with the */
/* modexp stage going through ifma_modexp.c the way the reference goes  */
/* through mbx_exp_mb8.  Used as the CPU baseline when the host has IFMA.*/
/* ------------------------------------------------------------------ */
int orc_have_ifma(void);
int orc_modexp_mb8(const u32* base, size_t base_stride, const u32* exp,
                   size_t exp_stride, int EL, const u32* mod, int L,
                   size_t count, u32* out);

/* R = 2^(52 len): RR = R^2 mod m as 52-bit digits, k0 = -m^-1 mod 2^52 */
int orc_radix52_constants(const u32* mod, int L, int len, u64* rr52, u64* k0) {
  int bits = 2 * 52 * len;
  int tw = bits / 32 + 1;
  if (tw > 2 * ORC_MAX_WORDS + 1 || L > ORC_MAX_WORDS) return -2;
  u32* t = calloc((size_t)tw, sizeof(u32));
  u32 r[ORC_MAX_WORDS];
  t[bits >> 5] = 1u << (bits & 31);
  bn_mod(r, t, tw, mod, L);
  free(t);
  for (int j = 0; j < len; j++) {
    u64 v = 0;
    for (int k = 0; k < 52; k++) {
      int b = j * 52 + k;
      if ((b >> 5) < L) v |= (u64)((r[b >> 5] >> (b & 31)) & 1u) << k;
    }
    rr52[j] = v;
  }
  u64 n0 = mod[0] | (L > 1 ? (u64)mod[1] << 32 : 0);
  u64 x = n0;
  for (int i = 0; i < 6; i++) x *= 2 - n0 * x;
  *k0 = (0 - x) & ((1ull << 52) - 1);
  return 0;
}

int orc_encrypt_mb8(const u32* n, int NL, const u32* hs, const u32* pt,
                    const u32* r, int RL, size_t count, u32* ct) {
  int L = 2 * NL;
  if (NL <= 0 || L > ORC_MAX_WORDS || RL > L) return -2;
  u32 nsq[ORC_MAX_WORDS];
  bn_mul(nsq, n, NL, n, NL);
  u32* gm = malloc(sizeof(u32) * count * (size_t)L);
  u32* obf = malloc(sizeof(u32) * count * (size_t)L);
  /* the reference runs this loop serially (pub_key.cpp:105); OpenMP here only
   * flatters the baseline */
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < count; i++) {
    u32 t[2 * ORC_MAX_WORDS + 1];
    bn_mul(t, n, NL, pt + i * (size_t)NL, NL);
    u64 cy = 1;
    for (int j = 0; j < L && cy; j++) {
      cy += t[j];
      t[j] = (u32)cy;
      cy >>= 32;
    }
    t[L] = (u32)cy;
    bn_mod(gm + i * (size_t)L, t, L + 1, nsq, L);
  }
  int rc;
  if (hs) {
    rc = orc_modexp_mb8(hs, 0, r, (size_t)RL, RL, nsq, L, count, obf);
  } else {
    u32* rb = calloc(count * (size_t)L, sizeof(u32));
    for (size_t i = 0; i < count; i++)
      memcpy(rb + i * (size_t)L, r + i * (size_t)RL, sizeof(u32) * (size_t)RL);
    rc = orc_modexp_mb8(rb, (size_t)L, n, 0, NL, nsq, L, count, obf);
    free(rb);
  }
  if (!rc) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < count; i++)
      bn_modmul(ct + i * (size_t)L, gm + i * (size_t)L, obf + i * (size_t)L, nsq, L);
  }
  free(gm);
  free(obf);
  return rc;
}

int orc_decrypt_crt_mb8(const u32* p, const u32* q, int PL, const u32* ct,
                        size_t count, u32* pt) {
  int NL = 2 * PL, CL = 2 * NL;
  if (PL <= 0 || CL > ORC_MAX_WORDS) return -2;
  u32 psq[ORC_MAX_WORDS], qsq[ORC_MAX_WORDS], hp[ORC_MAX_WORDS];
  u32 hq[ORC_MAX_WORDS], pinv[ORC_MAX_WORDS], pm1[ORC_MAX_WORDS];
  u32 qm1[ORC_MAX_WORDS];
  int rc = orc_crt_constants(p, q, PL, psq, qsq, hp, hq, pinv);
  if (rc) return rc;
  bn_sub_word(pm1, p, 1, PL);
  bn_sub_word(qm1, q, 1, PL);
  u32* bp = malloc(sizeof(u32) * count * (size_t)NL);
  u32* bq = malloc(sizeof(u32) * count * (size_t)NL);
  u32* rp = malloc(sizeof(u32) * count * (size_t)NL);
  u32* rq = malloc(sizeof(u32) * count * (size_t)NL);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < count; i++) { /* pri_key.cpp:127-130 */
    bn_mod(bp + i * (size_t)NL, ct + i * (size_t)CL, CL, psq, NL);
    bn_mod(bq + i * (size_t)NL, ct + i * (size_t)CL, CL, qsq, NL);
  }
  rc = orc_modexp_mb8(bp, (size_t)NL, pm1, 0, PL, psq, NL, count, rp); /* :133 */
  if (!rc) rc = orc_modexp_mb8(bq, (size_t)NL, qm1, 0, PL, qsq, NL, count, rq);
  if (!rc) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < count; i++) { /* :141-145 */
      u32 lp[ORC_MAX_WORDS], lq[ORC_MAX_WORDS], dp[ORC_MAX_WORDS];
      u32 dq[ORC_MAX_WORDS], u[ORC_MAX_WORDS], dpq[ORC_MAX_WORDS];
      u32 diff[ORC_MAX_WORDS], t[2 * ORC_MAX_WORDS];
      l_fun(lp, PL, rp + i * (size_t)NL, NL, p, PL);
      l_fun(lq, PL, rq + i * (size_t)NL, NL, q, PL);
      bn_modmul(dp, lp, hp, p, PL);
      bn_modmul(dq, lq, hq, q, PL);
      bn_mod(dpq, dp, PL, q, PL);
      int neg = bn_cmp(dq, dpq, PL) < 0;
      bn_sub(diff, dq, dpq, PL);
      if (neg) bn_add(diff, diff, q, PL);
      bn_modmul(u, diff, pinv, q, PL);
      bn_mul(t, u, PL, p, PL);
      memset(lp, 0, sizeof(u32) * (size_t)NL);
      memcpy(lp, dp, sizeof(u32) * (size_t)PL);
      bn_add(pt + i * (size_t)NL, t, lp, NL);
    }
  }
  free(bp);
  free(bq);
  free(rp);
  free(rq);
  return rc;
}
