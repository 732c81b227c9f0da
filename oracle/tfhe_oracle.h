/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference's batched bootstrapping path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * The product (tfhe_gpu_b200/libtfhe_b200.so) never links or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function below bit-for-bit against the
 * UNMODIFIED reference CPU implementation compiled from /root/reference (oracle/_ref/libtfhe_ref.so) on keys
 * exported from the reference, and tests/golden/ holds reference-generated vectors (see make_golden.py).
 *
 * All arrays are flat little-endian uint64.  A ciphertext of dimension m is (m+1) words: a[0..m-1], b.
 */
#ifndef TFHE_ORACLE_H
#define TFHE_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { TFO_METHOD_AP = 1, TFO_METHOD_GINX = 2 };                        /* binfhe-constants.h:94-98  */
enum { TFO_OR, TFO_AND, TFO_NOR, TFO_NAND, TFO_XOR_FAST, TFO_XNOR_FAST, TFO_XOR, TFO_XNOR }; /* :101 */
enum { TFO_SET_TOY = 0, TFO_SET_STD128_AP = 2, TFO_SET_STD128 = 4 };    /* binfhe-constants.h:46-80  */

typedef struct tfo_params {
    uint32_t n, N;
    uint64_t q, Q, qKS;
    uint32_t baseKS, dKS;
    uint32_t baseG, digitsG, numDigitsToThrow;
    uint32_t baseR, digitsR;
    uint32_t method;
    uint32_t reserved;
    uint64_t psi;  /* minimal primitive 2N-th root of unity mod Q (nbtheory.cpp:284-345) */
    uint64_t beta; /* binfhecontext.h:348 (always 128) */
} tfo_params;

typedef struct tfo_ctx tfo_ctx;

/* ---- parameter derivation (binfhecontext.cpp:42-181, nbtheory.cpp:284-345,481-576) ---- */
int tfo_params_named(int set, int method, tfo_params* out);
int tfo_params_func(int set, int arbFunc, uint32_t logQ, uint64_t N, uint32_t baseG, uint32_t numDigitsToThrow,
                    tfo_params* out);
int tfo_params_custom(uint32_t n, uint32_t N, uint64_t q, uint64_t Q, uint32_t baseKS, uint32_t baseG, uint32_t baseR,
                      int method, tfo_params* out);

tfo_ctx* tfo_ctx_new(const tfo_params* p);
void tfo_ctx_free(tfo_ctx* c);
size_t tfo_bk_words(const tfo_params* p);
size_t tfo_ksk_words(const tfo_params* p);

/* ---- number theoretic transform (transformnat-impl.h:298-341,478-531,683-739) ---- */
void tfo_ntt_forward(const tfo_ctx* c, uint64_t* poly); /* COEFFICIENT -> EVALUATION (bit-reversed order) */
void tfo_ntt_inverse(const tfo_ctx* c, uint64_t* poly); /* EVALUATION -> COEFFICIENT */

/* ---- key generation / encryption with a deterministic PRNG (semantics of lwe-pke.cpp:48-118,218-295,
 *      rgsw-acc-cggi.cpp:44-75,213-240, rgsw-acc-dm.cpp:44-76,153-209; the random stream is our own) ---- */
void tfo_keygen(const tfo_ctx* c, uint64_t seed, uint64_t* sk /*n, mod qKS*/, uint64_t* bk, uint64_t* ksk);
void tfo_encrypt(const tfo_ctx* c, const uint64_t* sk, int64_t m, uint64_t p, uint64_t mod, uint64_t seed,
                 uint64_t* ct /*n+1*/);
int64_t tfo_decrypt(const tfo_ctx* c, const uint64_t* sk, const uint64_t* ct, uint64_t mod, uint64_t p);

/* ---- stages ---- */
void tfo_signed_digit_decompose(const tfo_ctx* c, const uint64_t* in /*[2][N]*/, uint64_t* out /*[d][N]*/);
/* contract of GPUFFTBootstrap::EvalAcc_CUDA (bootstrapping.cuh:111-124): acc [batch][2][N] COEFFICIENT in and out,
 * a-polynomial transposed on exit */
void tfo_eval_acc(const tfo_ctx* c, const uint64_t* bk, int batch, const uint64_t* a /*[batch][n]*/, uint64_t mod,
                  uint64_t* acc);
uint64_t tfo_round_qQ(uint64_t v, uint64_t q, uint64_t Q);              /* lwe-pke.cpp:41-46   */
void tfo_mod_switch(int batch, size_t len, const uint64_t* in, uint64_t from_mod, uint64_t to_mod, uint64_t* out);
void tfo_key_switch(const tfo_ctx* c, const uint64_t* ksk, int batch, const uint64_t* in /*[batch][N+1]*/,
                    uint64_t* out /*[batch][n+1]*/);                     /* lwe-pke.cpp:299-321 */
/* contract of GPUFFTBootstrap::MKMSwitch_CUDA (bootstrapping.cuh:126-136): MS(Q->qKS), KS, MS(qKS->fmod) */
void tfo_mkmswitch(const tfo_ctx* c, const uint64_t* ksk, int batch, const uint64_t* in /*[batch][N+1] mod Q*/,
                   uint64_t fmod, uint64_t* out /*[batch][n+1]*/);
/* accumulator initialisation (binfhe-base-scheme.cpp:1087-1145 gate, :1147-1192 function); writes [2][N] */
void tfo_init_acc_gate(const tfo_ctx* c, int gate, uint64_t b, uint64_t ctmod, uint64_t* acc);

/* ---- batched operations: semantics of BinFHEContext::{EvalBinGate,EvalFunc,EvalFloor,EvalSign,EvalDecomp,
 *      CiphertextMulMatrix} (binfhecontext.cpp:319-347) evaluated with the CPU (NTT) arithmetic ---- */
int tfo_eval_bin_gate(const tfo_ctx* c, const uint64_t* bk, const uint64_t* ksk, int gate, int batch,
                      const uint64_t* ct1, const uint64_t* ct2, uint64_t mod, uint64_t* out);
/* generic functional bootstrap: LUT-driven BootstrapFunc (binfhe-base-scheme.cpp:1194-1211):
 * table[x], x < ctmod, already holds f(x, ctmod, fmod); per-ct tables when per_ct != 0 ([batch][ctmod]) */
int tfo_bootstrap_func(const tfo_ctx* c, const uint64_t* bk, const uint64_t* ksk, int batch, const uint64_t* ct,
                       uint64_t ctmod, const uint64_t* table, int per_ct, uint64_t fmod, uint64_t* out);
int tfo_eval_func(const tfo_ctx* c, const uint64_t* bk, const uint64_t* ksk, int batch, const uint64_t* ct,
                  uint64_t mod, const uint64_t* lut, size_t lut_len, int per_ct, uint64_t* out);
int tfo_eval_floor(const tfo_ctx* c, const uint64_t* bk, const uint64_t* ksk, int batch, const uint64_t* ct,
                   uint64_t mod, uint32_t roundbits, uint64_t* out);
int tfo_eval_sign(const tfo_ctx* c, const uint64_t* bk, const uint64_t* ksk, int batch, const uint64_t* ct,
                  uint64_t mod, uint64_t* out);
/* out [batch][max_digits][n+1]; returns number of digits (or -1) and their moduli in out_mods */
int tfo_eval_decomp(const tfo_ctx* c, const uint64_t* bk, const uint64_t* ksk, int batch, const uint64_t* ct,
                    uint64_t mod, int max_digits, uint64_t* out, uint64_t* out_mods);
/* dynamic gadget base ("timeOptimization", binfhecontext.cpp:222-247): nk key sets, each with its own context
 * (same ring, different baseG / digitsG), BK and KSK; set 0 is the context's own.  With nk == 3 the scalar
 * EvalSign / EvalDecomp rule applies (binfhe-base-scheme.cpp:342-360, 411-428): after each floor + modulus switch,
 * base 2^27 once the modulus is <= 2^17, 2^18 once it is <= 2^26. */
int tfo_eval_sign_dyn(int nk, const tfo_ctx* const* cs, const uint64_t* const* bks, const uint64_t* const* ksks,
                      int batch, const uint64_t* ct, uint64_t mod, uint64_t* out);
int tfo_eval_decomp_dyn(int nk, const tfo_ctx* const* cs, const uint64_t* const* bks, const uint64_t* const* ksks,
                        int batch, const uint64_t* ct, uint64_t mod, int max_digits, uint64_t* out,
                        uint64_t* out_mods);
/* out[i] = sum_k ct[k] * M[k][i] mod modulus; ct [in][n+1], M [in][outc] int64 row-major, out [outc][n+1]
 * (lwe-operation.cu:50-141; exact integer semantics, Euclidean residue for negative entries) */
int tfo_mul_matrix(const tfo_ctx* c, int in, int outc, const uint64_t* ct, const int64_t* M, uint64_t modulus,
                   uint64_t* out);

int tfo_num_threads(void);
void tfo_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
