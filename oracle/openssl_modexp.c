/*
 * openssl_modexp.c -- TEST/BENCH INFRASTRUCTURE ONLY.
 * Independent second opinion for the oracle and a widely understood CPU line:
 * OpenSSL BN_mod_exp_mont_consttime under OpenMP.  The reference validates its
 * own accelerator path the same way (module/heqat/test/test_bnModExp.cpp:60,205
 * compares QAT results with BN_mod_exp).  Built only when the OpenSSL headers
 * are present (oracle/Makefile).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef ORC_HAVE_OPENSSL
#include <openssl/bn.h>

static BIGNUM* from_limbs(const uint32_t* x, int L) {
  unsigned char* be = malloc((size_t)L * 4);
  for (int i = 0; i < L; i++) {
    uint32_t w = x[L - 1 - i];
    be[4 * i] = (unsigned char)(w >> 24);
    be[4 * i + 1] = (unsigned char)(w >> 16);
    be[4 * i + 2] = (unsigned char)(w >> 8);
    be[4 * i + 3] = (unsigned char)w;
  }
  BIGNUM* r = BN_bin2bn(be, L * 4, NULL);
  free(be);
  return r;
}

static void to_limbs(uint32_t* out, int L, const BIGNUM* x) {
  unsigned char* be = calloc((size_t)L * 4, 1);
  BN_bn2binpad(x, be, L * 4);
  for (int i = 0; i < L; i++) {
    const unsigned char* p = be + 4 * (L - 1 - i);
    out[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
  }
  free(be);
}

int orc_have_openssl(void) { return 1; }

/* out[i] = base[i]^exp[i] mod mod (shared odd modulus) */
int orc_modexp_openssl(const uint32_t* base, size_t base_stride, const uint32_t* exp,
                       size_t exp_stride, int EL, const uint32_t* mod, int L,
                       size_t count, uint32_t* out) {
  int rc = 0;
#pragma omp parallel
  {
    BN_CTX* ctx = BN_CTX_new();
    BIGNUM* m = from_limbs(mod, L);
    BN_MONT_CTX* mont = BN_MONT_CTX_new();
    if (!BN_MONT_CTX_set(mont, m, ctx)) rc = -1;
    BIGNUM* r = BN_new();
#pragma omp for schedule(dynamic, 8)
    for (size_t i = 0; i < count; i++) {
      BIGNUM* b = from_limbs(base + i * base_stride, L);
      BIGNUM* e = from_limbs(exp + i * exp_stride, EL);
      BN_nnmod(b, b, m, ctx);
      if (!BN_mod_exp_mont_consttime(r, b, e, m, ctx, mont)) rc = -1;
      to_limbs(out + i * (size_t)L, L, r);
      BN_free(b);
      BN_free(e);
    }
    BN_free(r);
    BN_MONT_CTX_free(mont);
    BN_free(m);
    BN_CTX_free(ctx);
  }
  return rc;
}
#else
int orc_have_openssl(void) { return 0; }
int orc_modexp_openssl(const uint32_t* base, size_t base_stride, const uint32_t* exp,
                       size_t exp_stride, int EL, const uint32_t* mod, int L,
                       size_t count, uint32_t* out) {
  (void)base; (void)base_stride; (void)exp; (void)exp_stride; (void)EL;
  (void)mod; (void)L; (void)count; (void)out;
  return -3;
}
#endif
