/*
 * ipcl_b200.h -- C ABI of the B200 (sm_100a) back-end for the modexp hot path
 * of intel/pailliercryptolib (IPCL v2.0.0).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch
 * types.  It occupies the slot the reference's QAT offload occupies
 * (ipcl/mod_exp.cpp:68-184 calling HE_QAT_bnModExp_MT,
 * module/heqat/heqat/include/heqat/bnops.h:121-148): marshal a batch into flat
 * buffers, submit, wait, unmarshal.  The C++ `ipcl::` layer in
 * pailliercryptolib_b200/ipcl/ is the reference-side caller.
 *
 * Number format everywhere: little-endian arrays of 32-bit words -- the layout
 * ippsRef_BN exposes and ippMBModExp copies from (ipcl/mod_exp.cpp:472-476,
 * 503-506) -- with a fixed stride per element, element-major.
 *
 * Every function returns 0 on success or a negative IPCLB200_ERR_* code (the
 * convention of HE_QAT_STATUS, module/heqat/heqat/include/heqat/common/
 * types.h:35-43); ipclb200_last_error() gives the text for the calling thread.
 * There is no CPU fallback: without a CUDA device every compute entry point
 * fails with IPCLB200_ERR_NO_DEVICE.
 *
 * Host-pointer entry points copy in, compute and copy out synchronously.
 * *_dev entry points take device pointers on the current device plus a CUDA
 * stream (passed as void*) and only enqueue work.
 * All entry points are re-entrant and lock-free on the data path: every
 * host-pointer call runs on its own stream with stream-ordered temporaries,
 * every *_dev call allocates its intermediates on the caller's stream, so
 * concurrent callers overlap (the reference's encrypt/decrypt are called
 * concurrently on one key, test/test_cryptography.cpp:45-57).
 */
#ifndef IPCL_B200_H_
#define IPCL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IPCLB200_OK 0
#define IPCLB200_ERR_BAD_ARG (-1)      /* null pointer, zero/oversize width ... */
#define IPCLB200_ERR_EVEN_MODULUS (-2) /* Montgomery needs an odd modulus      */
#define IPCLB200_ERR_CUDA (-3)         /* a CUDA runtime call failed           */
#define IPCLB200_ERR_NO_DEVICE (-4)    /* no usable sm_100 device              */
#define IPCLB200_ERR_UNSUPPORTED (-5)  /* width not supported by this entry    */
#define IPCLB200_ERR_NCCL (-6)         /* NCCL missing or a NCCL call failed   */

/* operand broadcast flags: the operand is ONE value used by every element
 * (what std::vector<BigNumber>(sz, x) expresses at ipcl/pub_key.cpp:53-54,
 * 68-69 and ipcl/pri_key.cpp:119-120) */
#define IPCLB200_SHARED_BASE 1u
#define IPCLB200_SHARED_EXP 2u
#define IPCLB200_SHARED_MOD 4u
/* modmul: the second factor is one value (the size-1 broadcast of
 * CipherText::operator+, ipcl/ciphertext.cpp:51-59) */
#define IPCLB200_SHARED_B 8u

#define IPCLB200_MAX_MOD_WORDS 256 /* 8192-bit modulus */

/* ---- runtime ------------------------------------------------------------ */
/* replaces ipcl::initializeContext / terminateContext for this back-end
 * (ipcl/utils/context.cpp:40-86).  device < 0 selects the current device (or
 * LOCAL_RANK when set).  Idempotent. */
int ipclb200_init(int device);
/* One process, several GPUs: host-pointer entry points and ipclb200_batch_*
 * split every batch into contiguous blocks over devices 0..n_devices-1
 * (n_devices <= 0: all visible devices); per-key constants and fixed-base
 * tables are replicated per device on first use.  This is the slot of the
 * reference's static prefix/suffix split of a batch between the CPU and the
 * accelerator (ipcl/mod_exp.cpp:702-731).  Blocks smaller than 512 elements are
 * not split off (IPCLB200_MIN_SHARD).  Call before the first batch. */
int ipclb200_init_devices(int n_devices);
int ipclb200_active_devices(void);
void ipclb200_shutdown(void);
int ipclb200_device_count(void);
/* 1 if the library was built with -DIPCLB200_EXPERIMENTS (the kernels that
 * measured slower: FP64 pipe, symmetric squarings, thread-per-integer) */
int ipclb200_has_experiments(void);
const char* ipclb200_last_error(void);
const char* ipclb200_version(void);

/* ---- batched modular exponentiation --------------------------------------
 * out[i] = base[i] ^ exp[i] mod mod[i], 0 <= out[i] < mod[i].
 * Replaces ipcl::ippModExp(vector, vector, vector), ipcl/mod_exp.cpp:655-678
 * (and through it ippMBModExp :446-533 -> mbx_exp_mb8, ippSBModExp :535-585).
 *   base : count x mod_words   (1 x mod_words with IPCLB200_SHARED_BASE)
 *   exp  : count x exp_words   (1 x exp_words with IPCLB200_SHARED_EXP)
 *   mod  : count x mod_words   (1 x mod_words with IPCLB200_SHARED_MOD)
 *   out  : count x mod_words
 * base may be >= mod (it is reduced); mod must be odd and non-zero;
 * exp == 0 gives 1 mod m. */
int ipclb200_modexp(const uint32_t* base, const uint32_t* exp,
                    const uint32_t* mod, int mod_words, int exp_words,
                    size_t count, unsigned flags, uint32_t* out);

/* Device-resident variant.  Requires IPCLB200_SHARED_MOD (mod is a HOST
 * pointer: the per-modulus constants are derived on the host once) and
 * mod_words in {16,32,48,64,96,128,192,256}.  exp_bits: number of significant
 * exponent bits over the whole batch (0 = exp_words*32).  base/exp/out are
 * device pointers, strides as above. */
int ipclb200_modexp_dev(const uint32_t* d_base, const uint32_t* d_exp,
                        const uint32_t* h_mod, int mod_words, int exp_words,
                        int exp_bits, size_t count, unsigned flags,
                        uint32_t* d_out, void* stream);

/* ---- batched modular multiplication ---------------------------------------
 * out[i] = a[i] * b[i] mod mod, one modulus for the batch.
 * Replaces CipherText::raw_add (ct + ct), ipcl/ciphertext.cpp:135-141 and the
 * loops :53-69; also BigNumber::ModMul at ipcl/pub_key.cpp:88-89. */
int ipclb200_modmul(const uint32_t* a, const uint32_t* b, const uint32_t* mod,
                    int mod_words, size_t count, unsigned flags, uint32_t* out);
int ipclb200_modmul_dev(const uint32_t* d_a, const uint32_t* d_b,
                        const uint32_t* h_mod, int mod_words, size_t count,
                        unsigned flags, uint32_t* d_out, void* stream);

/* ---- Paillier public key: encrypt ------------------------------------------
 * Holds n, n^2 and (DJN) hs plus their device-side constants and the
 * fixed-base table for hs.  Replaces the state of ipcl::PublicKey that the
 * hot path reads (ipcl/pub_key.cpp:18-49). */
typedef struct ipclb200_pubkey ipclb200_pubkey;

/* n: n_words words (n_words*32 >= key bits).  hs: 2*n_words words or NULL for
 * a non-DJN key.  rand_bits: DJN obfuscator exponent width (bits/2,
 * ipcl/pub_key.cpp:46), ignored without hs. */
int ipclb200_pubkey_create(const uint32_t* n, int n_words, const uint32_t* hs,
                           int rand_bits, ipclb200_pubkey** out);
void ipclb200_pubkey_destroy(ipclb200_pubkey* pk);
/* Fixed-base table policy of a DJN key (hs^r is computed from a comb table of
 * hs, the same base for every element: ipcl/pub_key.cpp:53).  A key starts with
 * a 17 MB table (8-bit windows) and, once it has encrypted `upgrade_after`
 * elements (default 8192), gets the widest table below `max_table_mb` (default
 * 4096; windows of up to 18 bits: 17 bits = 3.8 GB at a 2048-bit key, 16 bits =
 * 2.1 GB with max_table_mb 2048..3903, 18 bits = 7.7 GB) built on a side stream --
 * encryptions keep using the small table until the wide one is ready.  Wide
 * tables of all keys on a device stay below IPCLB200_COMB_DEVICE_MB (64 GB).
 * A negative argument keeps the current value. */
int ipclb200_pubkey_set_table_policy(ipclb200_pubkey* pk, long max_table_mb,
                                     long upgrade_after);

/* ct[i] = ((n*pt[i] + 1) mod n^2) * obf[i] mod n^2,
 *   obf[i] = hs^r[i] mod n^2 (DJN) or r[i]^n mod n^2 (non-DJN).
 * Replaces PublicKey::raw_encrypt + applyObfuscator + get{DJN,Normal}
 * Obfuscator, ipcl/pub_key.cpp:51-110.  The randoms r are supplied by the
 * caller (host RNG stays in the ipcl:: layer, as setRandom injects them at
 * ipcl/pub_key.cpp:92-95).  make_secure == 0 skips the obfuscator
 * (ipcl/pub_key.cpp:107; r may be NULL).
 *   pt: count x pt_words (pt_words <= n_words), r: count x r_words
 *   (r_words <= 2*n_words), ct: count x 2*n_words. */
int ipclb200_encrypt(const ipclb200_pubkey* pk, const uint32_t* pt,
                     int pt_words, const uint32_t* r, int r_words, size_t count,
                     int make_secure, uint32_t* ct);
int ipclb200_encrypt_dev(const ipclb200_pubkey* pk, const uint32_t* d_pt,
                         int pt_words, const uint32_t* d_r, int r_words,
                         size_t count, int make_secure, uint32_t* d_ct,
                         void* stream);

/* ---- DJN randoms drawn on the device ------------------------------------------
 * The reference draws r = getRandomBN(randbits) per element on the host
 * (ipcl/pub_key.cpp:59-61 -> ipcl/utils/common.cpp:42-101: RDSEED / RDRAND /
 * ippsPRNGen).  Here the host supplies one fresh 256-bit key per call (OS
 * entropy, the ipcl:: layer) and the r of the whole batch are ChaCha20
 * keystream generated in HBM: the block function of RFC 8439 section 2.3 (20
 * rounds) under key[8] / nonce[3] (little-endian words); element e takes the
 * blocks with counters e*bpe .. e*bpe+bpe-1, bpe = ceil(words/16), truncated
 * to `bits` bits (higher words zero).  No r crosses PCIe.
 *   random_dev    : d_out = count x words on the current device; first_element
 *                   is the global index of element 0 (a shard of a batch)
 *   batch_random  : every shard of a batch (element indices are global)
 *   encrypt_drbg  : ipclb200_encrypt for a DJN key with r drawn this way
 *                   (randbits of the key), host plaintexts in, ciphertexts out */
int ipclb200_random_dev(uint32_t* d_out, size_t count, int words, int bits,
                        const uint32_t* key, const uint32_t* nonce,
                        uint64_t first_element, void* stream);
int ipclb200_encrypt_drbg(const ipclb200_pubkey* pk, const uint32_t* pt,
                          int pt_words, size_t count, const uint32_t* key,
                          const uint32_t* nonce, uint32_t* ct);

/* ---- Paillier private key: decrypt ------------------------------------------
 * Derives p^2, q^2, p^-1 mod q, hp, hq (and lambda, x for the non-CRT path)
 * exactly as the PrivateKey constructor does, ipcl/pri_key.cpp:13-37,159-167.
 * p and q: p_words words each; they are ordered p < q internally (:19-22). */
typedef struct ipclb200_privkey ipclb200_privkey;
int ipclb200_privkey_create(const uint32_t* p, const uint32_t* q, int p_words,
                            ipclb200_privkey** out);
void ipclb200_privkey_destroy(ipclb200_privkey* sk);
/* Schedule of the key's SECRET exponents (p-1, q-1 in decryptCRT, lambda in
 * decryptRAW).  constant_schedule != 0: fixed-window ladders whose sequence of
 * squarings and multiplies depends on the exponent's bit length only -- the
 * property of the reference's mbx_exp_mb8 -- at ~8 % more products;
 * 0: host-built sliding-window schedules (the default, fastest; the operation
 * sequence is the same for every ciphertext but is derived from the key).
 * Table indices are exponent-dependent addresses in both modes.  Keys whose
 * primes do not fill their words use the full-width CRT kernel, which only has
 * the sliding-window form.  Default for new keys: IPCLB200_CONSTANT_SCHEDULE. */
int ipclb200_privkey_set_schedule(ipclb200_privkey* sk, int constant_schedule);

/* pt[i] = Dec(ct[i]).  use_crt != 0: PrivateKey::decryptCRT,
 * ipcl/pri_key.cpp:114-152; use_crt == 0: decryptRAW, :92-111.
 *   ct: count x 4*p_words (values < n^2), pt: count x 2*p_words. */
int ipclb200_decrypt(const ipclb200_privkey* sk, const uint32_t* ct,
                     size_t count, int use_crt, uint32_t* pt);
int ipclb200_decrypt_dev(const ipclb200_privkey* sk, const uint32_t* d_ct,
                         size_t count, int use_crt, uint32_t* d_pt,
                         void* stream);

/* ---- device-resident batches ------------------------------------------------
 * Lets a host keep ciphertext batches in HBM between encrypt -> + -> * ->
 * decrypt instead of round-tripping through vector<BigNumber> (the copies at
 * ipcl/base_text.cpp:102 and ipcl/mod_exp.cpp:627-632); the ipcl:: layer uses
 * these for its device-resident CipherText.  The counterpart in the reference
 * is the buffer acquire/release of the QAT offload
 * (module/heqat/heqat/include/heqat/bnops.h:121-148).
 * All of them work on the library's own stream of the primary device,
 * ipclb200_stream(), which is also the stream to hand to the *_dev entry points
 * for work on these buffers (the *_dev calls run on the device that owns their
 * output pointer):
 * allocation, copies, kernels and frees are then ordered on that one stream.
 *   dev_alloc / dev_free : stream-ordered (cudaMallocAsync / cudaFreeAsync)
 *   dev_upload           : the host buffer may be reused on return
 *   dev_download         : waits until the data has arrived
 *   dev_copy             : device to device, `bytes` bytes
 *   class_words          : the kernel size class (16,32,48,64,96,128,192,256)
 *                          that holds `words` words, 0 if none; the *_dev
 *                          entry points need strides that are a size class */
void* ipclb200_stream(void);
int ipclb200_dev_alloc(size_t bytes, void** d_out);
int ipclb200_dev_free(void* d);
int ipclb200_dev_upload(void* d, const void* h, size_t bytes);
int ipclb200_dev_download(void* h, const void* d, size_t bytes);
int ipclb200_dev_copy(void* d_dst, const void* d_src, size_t bytes);
int ipclb200_sync(void);
int ipclb200_class_words(int words);
/* page-locked (pinned, portable) host memory: copies from/to it are DMA
 * transfers that overlap across devices and with kernels */
int ipclb200_host_alloc(size_t bytes, void** out);
int ipclb200_host_free(void* p);

/* ---- batches sharded over the active devices ---------------------------------
 * count x words limbs in HBM, split into contiguous blocks over the devices of
 * ipclb200_init_devices (one block on one device otherwise).  Every operation
 * on a batch is enqueued on the calling thread's own stream of each shard's
 * device, ordered by an event per shard behind whatever touched the operands
 * before -- threads working on different batches overlap -- and nothing waits
 * until batch_download / batch_sync.  Batches of equal count are sharded
 * identically, so element i of every operand lives on the same device.
 *   batch_upload / download : host buffer of count x h_words (h_words <= words,
 *                             zero padded / truncated per element)
 *   batch_scatter / gather  : from / to ONE contiguous device buffer on the
 *                             first device, NCCL grouped send/recv over NVLink
 *                             (the "trivial scatter/gather of independent
 *                             ciphertexts"; NCCL is loaded on first use)
 *   encrypt / decrypt / modmul / modexp _batch : the *_dev entry points per shard.
 *     modmul_batch: b == NULL -> one shared factor h_b_shared (host, mod_words);
 *     modexp_batch: exp == NULL -> one shared exponent h_exp_shared (host). */
typedef struct ipclb200_batch ipclb200_batch;
int ipclb200_batch_alloc(size_t count, int words, ipclb200_batch** out);
void ipclb200_batch_free(ipclb200_batch* b);
size_t ipclb200_batch_count(const ipclb200_batch* b);
int ipclb200_batch_words(const ipclb200_batch* b);
int ipclb200_batch_num_shards(const ipclb200_batch* b);
int ipclb200_batch_shard(const ipclb200_batch* b, int shard, int* device, void** d_ptr,
                         size_t* begin, size_t* count, void** stream);
int ipclb200_batch_upload(ipclb200_batch* b, const uint32_t* h, int h_words);
int ipclb200_batch_download(const ipclb200_batch* b, uint32_t* h, int h_words);
int ipclb200_batch_sync(const ipclb200_batch* b);
/* batch_shard hands out the calling thread's stream, already ordered behind the
 * work enqueued on the shard; a caller that enqueues kernels of its own there
 * calls batch_touch afterwards so that later library calls wait for them */
int ipclb200_batch_touch(const ipclb200_batch* b, int shard, void* stream);
int ipclb200_batch_random(ipclb200_batch* out, int bits, const uint32_t* key,
                          const uint32_t* nonce);
int ipclb200_batch_scatter(ipclb200_batch* b, const uint32_t* d_src);
int ipclb200_batch_gather(const ipclb200_batch* b, uint32_t* d_dst);
int ipclb200_encrypt_batch(const ipclb200_pubkey* pk, const ipclb200_batch* pt,
                           const ipclb200_batch* r, int r_bits, int make_secure,
                           ipclb200_batch* ct);
int ipclb200_decrypt_batch(const ipclb200_privkey* sk, const ipclb200_batch* ct,
                           int use_crt, ipclb200_batch* pt);
int ipclb200_modmul_batch(const ipclb200_batch* a, const ipclb200_batch* b,
                          const uint32_t* h_b_shared, const uint32_t* h_mod,
                          int mod_words, ipclb200_batch* out);
int ipclb200_modexp_batch(const ipclb200_batch* base, const ipclb200_batch* exp,
                          const uint32_t* h_exp_shared, int exp_words, int exp_bits,
                          const uint32_t* h_mod, int mod_words, ipclb200_batch* out);

/* Diagnostics: the lane layout the CRT decrypt picks for a batch of `count`
 * ciphertexts at p_words-word primes on a device with `sms` SMs (0 = 148):
 * 0 / 1 / 2 = one (ciphertext, side) task over 2, 4, 8 lanes of a warp (4, 8 at
 * 64-word primes), -2 = one task per thread.  Host-only (no device needed); the
 * cost model is fitted to profiles/r02_layout_sweep.jsonl. */
int ipclb200_decrypt_layout(size_t count, int p_words, int sms);

/* ---- measurement helpers --------------------------------------------------
 * Runs the integer-pipe microbenchmark (dependent-carry IMAD.WIDE.U32 chains on
 * every SM) and returns the measured 32x32->64 multiply-accumulate rate; this
 * is the roofline denominator SURVEY.md section 8(d) asks the builder to
 * measure.  Also returns the kernel-launch count since init (for bench.py's
 * gpu_launches). */
int ipclb200_int_peak(double* mac32_per_s, double* sm_clock_mhz);
/* the same microbenchmark as a back-to-back train of launches lasting about
 * `seconds`: the SUSTAINED integer rate, the denominator for a kernel that
 * runs for hundreds of milliseconds (int_peak is the burst figure) */
int ipclb200_int_peak_sustained(double seconds, double* mac32_per_s);
uint64_t ipclb200_launch_count(void);
/* Operands of host-pointer calls used IN PLACE so far.  ipclb200_encrypt (DJN
 * keys) on batches of >= 1024 elements per device does not stage page-locked
 * caller buffers (cudaHostAlloc / cudaHostRegister / ipclb200_host_alloc
 * memory): the kernel reads the plaintexts and writes the ciphertexts through
 * the mapped alias of the buffers, element by element as it claims work, so the
 * PCIe transfer overlaps the arithmetic (the role of the QAT path's in-place
 * request buffers, module/heqat/heqat/include/heqat/bnops.h:121-148).  Pageable
 * buffers are staged by copies as before.  IPCLB200_ZERO_COPY is a bit mask:
 * 1 = encrypt plaintexts, 2 = encrypt ciphertexts, 4 = ciphertexts of the CRT
 * decrypt (measured slower than the staged copy, hence default 3); 0 stages
 * everything. */
uint64_t ipclb200_zero_copy_count(void);

/* Pipe-overlap probe (profiles/r01_pipe_overlap.md): time of a fixed number of
 * IMAD.WIDE chains (mode 0), DFMA chains (1), both kinds on every SM
 * sub-partition at once (2), and each half of mode 2 alone (3: integer half,
 * 4: DFMA half).  Tells whether the FP64 pipe can work beside the integer
 * multiply pipe. */
int ipclb200_pipe_mix(int mode, double* ms_out);

/* Diagnostics for the symmetric-squaring kernel (mont_sqr.cuh): for `count`
 * 64-word integers a (< 2^2048) and one odd 64-word modulus, out_sqr[i] =
 * MontSqr::sqr(a[i]) and out_mul[i] = Mont::mul(a[i], a[i]); both are
 * a^2 * 2^-2048 mod n up to a multiple of n (values below 2^2048). */
int ipclb200_debug_montsqr(const uint32_t* a, const uint32_t* mod, size_t count,
                           uint32_t* out_sqr, uint32_t* out_mul);

/* Diagnostics: the two CRT residues of every ciphertext,
 *   x[i][0] = ct[i]^(p-1) mod p^2,  x[i][1] = ct[i]^(q-1) mod q^2
 * (the modexp results of ipcl/pri_key.cpp:128-134, before the L function), as
 * the decrypt kernel -- whichever pipe(s) IPCLB200_DECRYPT selects -- left
 * them.  x: count x 2 x *x_words words; *x_words is the kernel size class of
 * p^2 (>= 2*p_words).  Used by the parity tests to check the integer-pipe and
 * FP64-pipe kernels against the oracle independently of the CRT tail. */
int ipclb200_crt_residues(const ipclb200_privkey* sk, const uint32_t* ct,
                          size_t count, uint32_t* x, int* x_words);

#ifdef __cplusplus
}
#endif
#endif /* IPCL_B200_H_ */
