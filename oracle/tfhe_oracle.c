/* TEST INFRASTRUCTURE ONLY -- see tfhe_oracle.h.
 *
 * Plain-C restatement of the reference's CPU (NTT) bootstrapping path.  Every function cites the reference
 * file:line it follows (paths relative to /root/reference/src).  Parity status: PINNED against the compiled
 * reference (tests/test_oracle_vs_ref.py) and against reference-generated golden vectors (tests/golden/).
 */
#include "tfhe_oracle.h"

#include <math.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;
typedef int64_t i64;
typedef unsigned __int128 u128;

/* ------------------------------------------------------------------------------------------------------ */
/* modular arithmetic helpers                                                                              */
/* ------------------------------------------------------------------------------------------------------ */
static inline u64 mulmod(u64 a, u64 b, u64 Q) {
    return (u64)(((u128)a * b) % Q);
}
/* fast variant for Q < 2^62 using an extended-precision reciprocal; exact after the two fix-ups */
static inline u64 mulmod_f(u64 a, u64 b, u64 Q, long double Qinv) {
    u64 qh = (u64)((long double)a * (long double)b * Qinv);
    i64 r  = (i64)(a * b - qh * Q);
    if (r < 0)
        r += (i64)Q;
    else if (r >= (i64)Q)
        r -= (i64)Q;
    return (u64)r;
}
static inline u64 addmod(u64 a, u64 b, u64 Q) {
    u64 r = a + b;
    return r >= Q ? r - Q : r;
}
static inline u64 submod(u64 a, u64 b, u64 Q) {
    return a >= b ? a - b : a + Q - b;
}
static u64 powmod(u64 a, u64 e, u64 Q) {
    u64 r = 1 % Q;
    a %= Q;
    while (e) {
        if (e & 1)
            r = mulmod(r, a, Q);
        a = mulmod(a, a, Q);
        e >>= 1;
    }
    return r;
}
/* Shoup multiplication by a constant w with companion w' = floor(w * 2^64 / Q) */
static inline u64 shoup_pre(u64 w, u64 Q) {
    return (u64)((((u128)w) << 64) / Q);
}
static inline u64 mulmod_shoup(u64 x, u64 w, u64 wp, u64 Q) {
    u64 qh = (u64)(((u128)x * wp) >> 64);
    u64 r  = x * w - qh * Q;
    return r >= Q ? r - Q : r;
}

/* deterministic Miller-Rabin for 64-bit integers (stands in for nbtheory.cpp MillerRabinPrimalityTest, which
 * is probabilistic; both decide primality of the same candidates) */
static int is_prime(u64 n) {
    if (n < 2)
        return 0;
    static const u64 small[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (size_t i = 0; i < sizeof(small) / sizeof(small[0]); i++) {
        if (n % small[i] == 0)
            return n == small[i];
    }
    u64 d = n - 1;
    int s = 0;
    while (!(d & 1)) {
        d >>= 1;
        s++;
    }
    for (size_t i = 0; i < sizeof(small) / sizeof(small[0]); i++) {
        u64 x = powmod(small[i], d, n);
        if (x == 1 || x == n - 1)
            continue;
        int comp = 1;
        for (int r = 1; r < s; r++) {
            x = mulmod(x, x, n);
            if (x == n - 1) {
                comp = 0;
                break;
            }
        }
        if (comp)
            return 0;
    }
    return 1;
}

/* nbtheory.cpp:481-520 FirstPrime: first prime = 1 mod m that is > 2^nBits */
static u64 first_prime(u64 nBits, u64 m) {
    u64 r  = powmod(2, nBits, m);
    u64 q  = 1ULL << nBits;
    q      = (r > 0) ? q + (m - r) + 1 : q + 1;
    while (!is_prime(q))
        q += m;
    return q;
}
/* nbtheory.cpp:559-574 PreviousPrime */
static u64 previous_prime(u64 q, u64 m) {
    u64 c = q - m;
    while (!is_prime(c))
        c -= m;
    return c;
}
/* nbtheory.cpp:284-345 RootOfUnity: the MINIMAL primitive m-th root of unity (m a power of two here, so the
 * primitive roots are exactly the odd powers of any one of them) */
static u64 min_root_of_unity(u64 m, u64 Q) {
    u64 g = 0;
    for (u64 x = 2; x < Q; x++) {
        g = powmod(x, (Q - 1) / m, Q);
        if (powmod(g, m / 2, Q) == Q - 1)
            break;
    }
    u64 g2  = mulmod(g, g, Q);
    u64 cur = g, best = g;
    for (u64 i = 1; i < m / 2; i++) {
        cur = mulmod(cur, g2, Q);
        if (cur < best)
            best = cur;
    }
    return best;
}

static uint32_t ceil_log_ratio(double x, double base) {
    return (uint32_t)ceil(log(x) / log(base));
}

static void finish_params(tfo_params* p) {
    p->dKS     = ceil_log_ratio((double)p->qKS, (double)p->baseKS);            /* lwe-pke.cpp:305 */
    p->digitsG = ceil_log_ratio((double)p->Q, (double)p->baseG);               /* rgsw-cryptoparameters.h:86 */
    p->digitsR = (p->method == TFO_METHOD_AP) ? ceil_log_ratio((double)p->q, (double)p->baseR) : 0; /* :88-89 */
    p->psi     = min_root_of_unity(2ULL * p->N, p->Q);                         /* :79 */
    p->beta    = 128;                                                          /* binfhecontext.h:348 */
    p->reserved = 0;
}

/* binfhecontext.cpp:114-181 (rows of paramsMap needed by BASELINE.json's configs) */
int tfo_params_named(int set, int method, tfo_params* p) {
    memset(p, 0, sizeof(*p));
    uint32_t bits, cyc, n, mod, modKS, baseKS, baseG, baseR;
    switch (set) {
        case TFO_SET_TOY:       bits = 27; cyc = 1024; n = 64;  mod = 512;  modKS = 0;       baseKS = 25;  baseG = 1u << 9; baseR = 23; break;
        case TFO_SET_STD128_AP: bits = 27; cyc = 2048; n = 512; mod = 1024; modKS = 1u << 14; baseKS = 128; baseG = 1u << 9; baseR = 32; break;
        case TFO_SET_STD128:    bits = 27; cyc = 2048; n = 512; mod = 1024; modKS = 1u << 14; baseKS = 128; baseG = 1u << 7; baseR = 32; break;
        default: return -1;
    }
    p->Q      = previous_prime(first_prime(bits, cyc), cyc);
    p->N      = cyc / 2;
    p->n      = n;
    p->q      = mod;
    p->qKS    = modKS ? modKS : p->Q;
    p->baseKS = baseKS;
    p->baseG  = baseG;
    p->baseR  = baseR;
    p->method = (uint32_t)method;
    finish_params(p);
    return 0;
}

/* binfhecontext.cpp:51-112 */
int tfo_params_func(int set, int arbFunc, uint32_t logQ, uint64_t N, uint32_t baseG, uint32_t numDigitsToThrow,
                    tfo_params* p) {
    memset(p, 0, sizeof(*p));
    if ((set != TFO_SET_STD128 && set != TFO_SET_TOY) || logQ > 29 || logQ < 11)
        return -1;
    uint32_t logQprime = 54;
    if (baseG == 0) {
        if (logQ > 25)
            baseG = 1u << 14;
        else if (logQ > 16)
            baseG = 1u << 18;
        else if (logQ > 11)
            baseG = 1u << 27;
        else {
            baseG     = 1u << 5;
            logQprime = 27;
        }
    }
    /* StdLatticeParm::FindRingDim(HEStd_ternary, HEStd_128_classic, logQprime): 1024 for 27 bits, 2048 for 54
     * (core/lib/lattice/stdlatticeparms.cpp table: n=1024 -> 27 bits, n=2048 -> 54 bits) */
    uint64_t ringDim = (logQprime <= 27) ? 1024 : 2048;
    if (N >= ringDim)
        ringDim = N;
    p->Q      = previous_prime(first_prime(logQprime, 2 * ringDim), 2 * ringDim);
    p->N      = (uint32_t)ringDim;
    p->q      = arbFunc ? ringDim : 2 * ringDim;
    p->qKS    = 1ULL << 35;
    p->n      = (set == TFO_SET_TOY) ? 32 : 1305;
    p->baseKS = 32;
    p->baseG  = baseG;
    p->baseR  = 23;
    p->method = TFO_METHOD_GINX;
    p->numDigitsToThrow = numDigitsToThrow;
    finish_params(p);
    return 0;
}

/* binfhecontext.cpp:42-49 (qKS = Q) */
int tfo_params_custom(uint32_t n, uint32_t N, uint64_t q, uint64_t Q, uint32_t baseKS, uint32_t baseG, uint32_t baseR,
                      int method, tfo_params* p) {
    memset(p, 0, sizeof(*p));
    p->n = n; p->N = N; p->q = q; p->Q = Q; p->qKS = Q;
    p->baseKS = baseKS; p->baseG = baseG; p->baseR = baseR; p->method = (uint32_t)method;
    finish_params(p);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------ */
/* context: NTT tables                                                                                     */
/* ------------------------------------------------------------------------------------------------------ */
struct tfo_ctx {
    tfo_params p;
    uint32_t logN, d, gBits;
    u64 *w, *wp;     /* psi^bitrev(k), Shoup companion            (transformnat-impl.h:683-739 PreCompute) */
    u64 *wi, *wip;   /* psi^-bitrev(k), Shoup companion */
    u64 Ninv, Ninvp;
    long double Qinv;
};

static uint32_t bitrev(uint32_t x, uint32_t bits) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < bits; i++) {
        r = (r << 1) | (x & 1);
        x >>= 1;
    }
    return r;
}

tfo_ctx* tfo_ctx_new(const tfo_params* p) {
    tfo_ctx* c = (tfo_ctx*)calloc(1, sizeof(tfo_ctx));
    c->p       = *p;
    u64 N = p->N, Q = p->Q;
    c->logN = 0;
    while ((1ULL << c->logN) < N)
        c->logN++;
    c->d     = 2 * (p->digitsG - p->numDigitsToThrow);
    c->gBits = (uint32_t)log2((double)p->baseG);      /* rgsw-acc.cpp:70 */
    c->w   = (u64*)malloc(sizeof(u64) * N);
    c->wp  = (u64*)malloc(sizeof(u64) * N);
    c->wi  = (u64*)malloc(sizeof(u64) * N);
    c->wip = (u64*)malloc(sizeof(u64) * N);
    u64 psi = p->psi, psii = powmod(psi, Q - 2, Q);
    u64 x = 1, xi = 1;
    for (u64 k = 0; k < N; k++) {
        uint32_t r = bitrev((uint32_t)k, c->logN);
        c->w[r]    = x;
        c->wi[r]   = xi;
        x  = mulmod(x, psi, Q);
        xi = mulmod(xi, psii, Q);
    }
    for (u64 k = 0; k < N; k++) {
        c->wp[k]  = shoup_pre(c->w[k], Q);
        c->wip[k] = shoup_pre(c->wi[k], Q);
    }
    c->Ninv  = powmod(N, Q - 2, Q);
    c->Ninvp = shoup_pre(c->Ninv, Q);
    c->Qinv  = 1.0L / (long double)Q;
    return c;
}

void tfo_ctx_free(tfo_ctx* c) {
    if (!c)
        return;
    free(c->w); free(c->wp); free(c->wi); free(c->wip);
    free(c);
}

size_t tfo_bk_words(const tfo_params* p) {
    size_t N = p->N, n = p->n;
    if (p->method == TFO_METHOD_GINX)
        return 2 * n * (size_t)(2 * (p->digitsG - p->numDigitsToThrow)) * 2 * N;
    return n * (size_t)p->baseR * p->digitsR * (size_t)(2 * p->digitsG) * 2 * N;
}
size_t tfo_ksk_words(const tfo_params* p) {
    return (size_t)p->N * p->baseKS * p->dKS * (p->n + 1);
}

/* transformnat-impl.h:298-341 ForwardTransformToBitReverseInPlace (Cooley-Tukey, natural in, bit-reversed out) */
void tfo_ntt_forward(const tfo_ctx* c, u64* a) {
    u64 N = c->p.N, Q = c->p.Q;
    u64 t = N;
    for (u64 m = 1; m < N; m <<= 1) {
        t >>= 1;
        for (u64 i = 0; i < m; i++) {
            u64 j1 = 2 * i * t, S = c->w[m + i], Sp = c->wp[m + i];
            for (u64 j = j1; j < j1 + t; j++) {
                u64 U = a[j], V = mulmod_shoup(a[j + t], S, Sp, Q);
                a[j]     = addmod(U, V, Q);
                a[j + t] = submod(U, V, Q);
            }
        }
    }
}
/* transformnat-impl.h:478-531 InverseTransformFromBitReverseInPlace (Gentleman-Sande), then * N^-1 */
void tfo_ntt_inverse(const tfo_ctx* c, u64* a) {
    u64 N = c->p.N, Q = c->p.Q;
    u64 t = 1;
    for (u64 m = N; m > 1; m >>= 1) {
        u64 h = m >> 1, j1 = 0;
        for (u64 i = 0; i < h; i++) {
            u64 S = c->wi[h + i], Sp = c->wip[h + i];
            for (u64 j = j1; j < j1 + t; j++) {
                u64 U = a[j], V = a[j + t];
                a[j]     = addmod(U, V, Q);
                a[j + t] = mulmod_shoup(submod(U, V, Q), S, Sp, Q);
            }
            j1 += 2 * t;
        }
        t <<= 1;
    }
    for (u64 j = 0; j < N; j++)
        a[j] = mulmod_shoup(a[j], c->Ninv, c->Ninvp, Q);
}

/* ------------------------------------------------------------------------------------------------------ */
/* PRNG (our own stream; the reference uses a randomly seeded BLAKE2 generator, so streams can never match) */
/* ------------------------------------------------------------------------------------------------------ */
typedef struct { u64 s[4]; } rng_t;
static u64 splitmix(u64* x) {
    u64 z = (*x += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static void rng_seed(rng_t* r, u64 seed) {
    for (int i = 0; i < 4; i++)
        r->s[i] = splitmix(&seed);
}
static inline u64 rotl(u64 x, int k) {
    return (x << k) | (x >> (64 - k));
}
static u64 rng_next(rng_t* r) {
    u64* s = r->s;
    u64 res = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return res;
}
static u64 rng_uniform(rng_t* r, u64 m) {
    return (u64)(((u128)rng_next(r) * m) >> 64);
}
static i64 rng_gauss(rng_t* r, double std) {
    double u1 = ((double)(rng_next(r) >> 11) + 1.0) / 9007199254740993.0;
    double u2 = (double)(rng_next(r) >> 11) / 9007199254740992.0;
    return (i64)llround(std * sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2));
}
static u64 signed_to_mod(i64 v, u64 m) {
    i64 r = v % (i64)m;
    return (u64)(r < 0 ? r + (i64)m : r);
}

/* ------------------------------------------------------------------------------------------------------ */
/* key generation                                                                                          */
/* ------------------------------------------------------------------------------------------------------ */
static const double TFO_STD = 3.19;

/* one RGSW row pair list; rows [d][2][N] in EVALUATION format.
 * CGGI: rgsw-acc-cggi.cpp:213-240;  DM: rgsw-acc-dm.cpp:153-209 */
static void rgsw_encrypt(const tfo_ctx* c, rng_t* r, const u64* skN_ntt, int is_dm, i64 m, u64* rows) {
    const tfo_params* p = &c->p;
    u64 N = p->N, Q = p->Q;
    uint32_t d = is_dm ? 2 * p->digitsG : c->d;
    uint32_t thr = is_dm ? 0 : p->numDigitsToThrow;
    u64 Gpow[64];
    u64 g = 1;
    for (uint32_t i = 0; i < p->digitsG; i++) {
        Gpow[i] = g;
        g       = mulmod(g, p->baseG % Q, Q);
    }
    i64 mm = 0;
    int neg = 0;
    if (is_dm) {
        i64 q = (i64)p->q;
        mm    = (((m % q) + q) % q) * (i64)(2 * N / p->q);
        if (mm >= (i64)N) {
            mm -= (i64)N;
            neg = 1;
        }
    }
    u64* tmp = (u64*)malloc(sizeof(u64) * N);
    for (uint32_t i = 0; i < d; i++) {
        u64* A = rows + ((size_t)i * 2 + 0) * N;
        u64* B = rows + ((size_t)i * 2 + 1) * N;
        for (u64 k = 0; k < N; k++)
            A[k] = rng_uniform(r, Q);
        for (u64 k = 0; k < N; k++)
            B[k] = signed_to_mod(rng_gauss(r, TFO_STD), Q);
        memcpy(tmp, A, sizeof(u64) * N);
        if (is_dm) {
            u64 G = Gpow[i >> 1];
            u64* tgt = (i & 1) ? B : A;
            tgt[mm]  = neg ? submod(tgt[mm], G, Q) : addmod(tgt[mm], G, Q);
        }
        else if (m) {
            u64* tgt = (i & 1) ? B : A;
            tgt[0]   = addmod(tgt[0], Gpow[(i >> 1) + thr], Q);
        }
        tfo_ntt_forward(c, A);
        tfo_ntt_forward(c, B);
        tfo_ntt_forward(c, tmp);
        for (u64 k = 0; k < N; k++)
            B[k] = addmod(B[k], mulmod(tmp[k], skN_ntt[k], Q), Q);
    }
    free(tmp);
}

void tfo_keygen(const tfo_ctx* c, u64 seed, u64* sk, u64* bk, u64* ksk) {
    const tfo_params* p = &c->p;
    u64 N = p->N, n = p->n, Q = p->Q, qKS = p->qKS;
    rng_t r0;
    rng_seed(&r0, seed);
    /* lwe-pke.cpp:48-51 (ternary secret, stored mod qKS: binfhecontext.cpp:183-186) */
    i64* s  = (i64*)malloc(sizeof(i64) * n);
    i64* sN = (i64*)malloc(sizeof(i64) * N);
    for (u64 i = 0; i < n; i++) {
        s[i]  = (i64)rng_uniform(&r0, 3) - 1;
        sk[i] = signed_to_mod(s[i], qKS);
    }
    for (u64 i = 0; i < N; i++)
        sN[i] = (i64)rng_uniform(&r0, 3) - 1;
    u64 base_seed = rng_next(&r0);

    /* key switching key: lwe-pke.cpp:218-295 */
#pragma omp parallel for schedule(static)
    for (u64 i = 0; i < N; i++) {
        rng_t r;
        rng_seed(&r, base_seed ^ (0x1000000ULL + i));
        for (u64 j = 0; j < p->baseKS; j++) {
            u64 dig = 1;
            for (u64 k = 0; k < p->dKS; k++) {
                u64* row = ksk + (((i * p->baseKS + j) * p->dKS + k) * (n + 1));
                u64 b = signed_to_mod(rng_gauss(&r, TFO_STD), qKS);
                b     = addmod(b, mulmod(signed_to_mod(sN[i], qKS), mulmod(j % qKS, dig % qKS, qKS), qKS), qKS);
                for (u64 t = 0; t < n; t++) {
                    row[t] = rng_uniform(&r, qKS);
                    b      = addmod(b, mulmod(row[t], sk[t], qKS), qKS);
                }
                row[n] = b;
                dig *= p->baseKS;
            }
        }
    }

    /* bootstrapping key: binfhe-base-scheme.cpp:38-57 */
    u64* skN_ntt = (u64*)malloc(sizeof(u64) * N);
    for (u64 i = 0; i < N; i++)
        skN_ntt[i] = signed_to_mod(sN[i], Q);
    tfo_ntt_forward(c, skN_ntt);
    if (p->method == TFO_METHOD_GINX) {
        size_t stride = (size_t)c->d * 2 * N;
        /* rgsw-acc-cggi.cpp:44-75: s=0 -> {0,0}; 1 -> {1,0}; -1 -> {0,1} */
#pragma omp parallel for schedule(dynamic)
        for (u64 i = 0; i < n; i++) {
            rng_t r;
            rng_seed(&r, base_seed ^ (0x2000000ULL + i));
            rgsw_encrypt(c, &r, skN_ntt, 0, s[i] == 1, bk + ((0 * n + i) * stride));
            rgsw_encrypt(c, &r, skN_ntt, 0, s[i] == -1, bk + ((1 * n + i) * stride));
        }
    }
    else {
        /* rgsw-acc-dm.cpp:44-76 */
        size_t stride = (size_t)(2 * p->digitsG) * 2 * N;
        u64 bR = p->baseR, dR = p->digitsR;
#pragma omp parallel for schedule(dynamic)
        for (u64 i = 0; i < n; i++) {
            rng_t r;
            rng_seed(&r, base_seed ^ (0x3000000ULL + i));
            memset(bk + (i * bR * dR) * stride, 0, sizeof(u64) * dR * stride); /* a0 = 0 rows unused */
            for (u64 a0 = 1; a0 < bR; a0++) {
                i64 dig = 1;
                for (u64 k = 0; k < dR; k++) {
                    rgsw_encrypt(c, &r, skN_ntt, 1, s[i] * (i64)a0 * dig, bk + (((i * bR + a0) * dR + k) * stride));
                    dig *= (i64)bR;
                }
            }
        }
    }
    free(skN_ntt);
    free(s);
    free(sN);
}

/* lwe-pke.cpp:53-84 */
void tfo_encrypt(const tfo_ctx* c, const u64* sk, i64 m, u64 pt, u64 mod, u64 seed, u64* ct) {
    const tfo_params* p = &c->p;
    rng_t r;
    rng_seed(&r, seed);
    u64 n = p->n;
    u64 b = ((u64)(((m % (i64)pt) + (i64)pt) % (i64)pt)) * (mod / pt);
    b     = (b + signed_to_mod(rng_gauss(&r, TFO_STD), mod)) % mod;
    for (u64 i = 0; i < n; i++) {
        ct[i]  = rng_uniform(&r, mod);
        /* secret is ternary mod qKS; SwitchModulus maps qKS-1 -> mod-1 */
        u64 si = sk[i] == 0 ? 0 : (sk[i] == 1 ? 1 : mod - 1);
        b      = addmod(b, mulmod(ct[i], si, mod), mod);
    }
    ct[n] = b;
}
/* lwe-pke.cpp:86-130 */
i64 tfo_decrypt(const tfo_ctx* c, const u64* sk, const u64* ct, u64 mod, u64 pt) {
    u64 n = c->p.n, inner = 0;
    for (u64 i = 0; i < n; i++) {
        u64 si = sk[i] == 0 ? 0 : (sk[i] == 1 ? 1 : mod - 1);
        inner  = addmod(inner, mulmod(ct[i] % mod, si, mod), mod);
    }
    u64 r = submod(ct[n] % mod, inner, mod);
    r     = addmod(r, (mod / (pt * 2)) % mod, mod);
    return (i64)((pt * r) / mod);
}

/* ------------------------------------------------------------------------------------------------------ */
/* stages                                                                                                  */
/* ------------------------------------------------------------------------------------------------------ */
/* rgsw-acc.cpp:57-111 (VARIANT A) */
void tfo_signed_digit_decompose(const tfo_ctx* c, const u64* in, u64* out) {
    const tfo_params* p = &c->p;
    u64 N = p->N;
    i64 Q = (i64)p->Q;
    u64 QHalf = p->Q >> 1;
    uint32_t thr = p->numDigitsToThrow, dg = p->digitsG - thr;
    int gBits = (int)c->gBits, sh = 64 - gBits;
    for (int j = 0; j < 2; j++)
        for (u64 k = 0; k < N; k++) {
            u64 t = in[(size_t)j * N + k];
            i64 d = (t < QHalf) ? (i64)t : (i64)t - Q;
            i64 r;
            for (uint32_t i = 0; i < thr; i++) {
                r = (i64)((u64)d << sh) >> sh;
                d = (d - r) >> gBits;
            }
            for (uint32_t l = 0; l < dg; l++) {
                r = (i64)((u64)d << sh) >> sh;
                d -= r;
                d >>= gBits;
                if (r < 0)
                    r += Q;
                out[(size_t)(j + 2 * l) * N + k] = (u64)r;
            }
        }
}

/* monomial X^m - 1 (m < N) or -X^(m-N) - 1 (N <= m < 2N) in EVALUATION format: rgsw-cryptoparameters.h:141-159 */
static void monomial_ntt(const tfo_ctx* c, u64 m, u64* out) {
    u64 N = c->p.N, Q = c->p.Q;
    memset(out, 0, sizeof(u64) * N);
    if (m < N)
        out[m] = addmod(out[m], 1, Q);
    else
        out[m - N] = submod(out[m - N], 1, Q);
    out[0] = submod(out[0], 1, Q);
    tfo_ntt_forward(c, out);
}

typedef struct {
    u64 *ct, *dct, *mp, *mn, *t;
} acc_scratch;

/* rgsw-acc-cggi.cpp:246-307 */
static void add_to_acc_cggi(const tfo_ctx* c, const u64* ek1, const u64* ek2, u64 a, u64* acc, acc_scratch* s) {
    const tfo_params* p = &c->p;
    u64 N = p->N, Q = p->Q, M = 2 * N;
    uint32_t d = c->d;
    memcpy(s->ct, acc, sizeof(u64) * 2 * N);
    tfo_ntt_inverse(c, s->ct);
    tfo_ntt_inverse(c, s->ct + N);
    tfo_signed_digit_decompose(c, s->ct, s->dct);
    for (uint32_t l = 0; l < d; l++)
        tfo_ntt_forward(c, s->dct + (size_t)l * N);
    u64 ipos = a % M, ineg = (M - a % M) % M;
    monomial_ntt(c, ipos, s->mp);
    monomial_ntt(c, ineg, s->mn);
    for (int key = 0; key < 2; key++) {
        const u64* ek  = key ? ek2 : ek1;
        const u64* mon = key ? s->mn : s->mp;
        for (int j = 0; j < 2; j++) {
            for (u64 k = 0; k < N; k++) {
                u64 t = 0;
                for (uint32_t l = 0; l < d; l++)
                    t = addmod(t, mulmod_f(s->dct[(size_t)l * N + k], ek[((size_t)l * 2 + j) * N + k], Q, c->Qinv), Q);
                acc[(size_t)j * N + k] = addmod(acc[(size_t)j * N + k], mulmod_f(t, mon[k], Q, c->Qinv), Q);
            }
        }
    }
}

/* rgsw-acc-dm.cpp:306-359 -- NOTE: the sums start at l = 1 (dct[0] is dropped), exactly as the reference does */
static void add_to_acc_dm(const tfo_ctx* c, const u64* ek, u64* acc, acc_scratch* s) {
    const tfo_params* p = &c->p;
    u64 N = p->N, Q = p->Q;
    uint32_t d = 2 * p->digitsG;
    memcpy(s->ct, acc, sizeof(u64) * 2 * N);
    tfo_ntt_inverse(c, s->ct);
    tfo_ntt_inverse(c, s->ct + N);
    tfo_signed_digit_decompose(c, s->ct, s->dct);
    for (uint32_t l = 0; l < d; l++)
        tfo_ntt_forward(c, s->dct + (size_t)l * N);
    for (int j = 0; j < 2; j++)
        for (u64 k = 0; k < N; k++) {
            u64 t = 0;
            for (uint32_t l = 1; l < d; l++)
                t = addmod(t, mulmod_f(s->dct[(size_t)l * N + k], ek[((size_t)l * 2 + j) * N + k], Q, c->Qinv), Q);
            acc[(size_t)j * N + k] = t;
        }
}

/* one ciphertext: acc [2][N] COEFFICIENT in/out (a-poly transposed on exit) */
static void eval_acc_one(const tfo_ctx* c, const u64* bk, const u64* a, u64 mod, u64* acc, acc_scratch* s) {
    const tfo_params* p = &c->p;
    u64 N = p->N, n = p->n, Q = p->Q;
    tfo_ntt_forward(c, acc);
    tfo_ntt_forward(c, acc + N);
    if (p->method == TFO_METHOD_GINX) {
        /* rgsw-acc-cggi.cpp:143-155 */
        size_t stride = (size_t)c->d * 2 * N;
        u64 M = 2 * N;
        for (u64 i = 0; i < n; i++) {
            u64 ai = a[i] % mod;
            u64 e  = ((mod - ai) % mod) * (M / mod);
            if (e == 0)
                continue; /* monomial X^0 - 1 == 0: the update adds zero */
            add_to_acc_cggi(c, bk + (0 * n + i) * stride, bk + (1 * n + i) * stride, e, acc, s);
        }
    }
    else {
        /* rgsw-acc-dm.cpp:80-110 */
        size_t stride = (size_t)(2 * p->digitsG) * 2 * N;
        u64 q = p->q, bR = p->baseR, dR = p->digitsR;
        for (u64 i = 0; i < n; i++) {
            u64 aI = (q - a[i] % q) % q;
            for (u64 k = 0; k < dR; k++, aI /= bR) {
                u64 a0 = aI % bR;
                if (a0)
                    add_to_acc_dm(c, bk + (((i * bR + a0) * dR + k) * stride), acc, s);
            }
        }
    }
    tfo_ntt_inverse(c, acc);
    tfo_ntt_inverse(c, acc + N);
    /* Transpose of the a-polynomial (binfhe-base-scheme.cpp:93): a'(X) = a(X^-1) */
    u64* t = s->t;
    t[0]   = acc[0];
    for (u64 i = 1; i < N; i++)
        t[i] = acc[N - i] ? Q - acc[N - i] : 0;
    memcpy(acc, t, sizeof(u64) * N);
}

static void scratch_new(const tfo_ctx* c, acc_scratch* s) {
    u64 N = c->p.N;
    uint32_t d = c->p.method == TFO_METHOD_GINX ? c->d : 2 * c->p.digitsG;
    s->ct  = (u64*)malloc(sizeof(u64) * 2 * N);
    s->dct = (u64*)calloc((size_t)d * N, sizeof(u64));
    s->mp  = (u64*)malloc(sizeof(u64) * N);
    s->mn  = (u64*)malloc(sizeof(u64) * N);
    s->t   = (u64*)malloc(sizeof(u64) * N);
}
static void scratch_free(acc_scratch* s) {
    free(s->ct); free(s->dct); free(s->mp); free(s->mn); free(s->t);
}

void tfo_eval_acc(const tfo_ctx* c, const u64* bk, int batch, const u64* a, u64 mod, u64* acc) {
    u64 N = c->p.N, n = c->p.n;
#pragma omp parallel
    {
        acc_scratch s;
        scratch_new(c, &s);
#pragma omp for schedule(dynamic)
        for (int b = 0; b < batch; b++)
            eval_acc_one(c, bk, a + (size_t)b * n, mod, acc + (size_t)b * 2 * N, &s);
        scratch_free(&s);
    }
}

/* lwe-pke.cpp:41-46: three IEEE double operations, then floor, then mod q */
u64 tfo_round_qQ(u64 v, u64 q, u64 Q) {
    volatile double prod = (double)v * (double)q;
    volatile double quot = prod / (double)Q;
    volatile double sum  = 0.5 + quot;
    return ((u64)floor(sum)) % q;
}

/* lwe-pke.cpp:204-215 */
void tfo_mod_switch(int batch, size_t len, const u64* in, u64 from_mod, u64 to_mod, u64* out) {
    for (size_t i = 0; i < (size_t)batch * len; i++)
        out[i] = tfo_round_qQ(in[i], to_mod, from_mod);
}

/* lwe-pke.cpp:299-321 (note: digit value 0 rows are subtracted too) */
static void key_switch_one(const tfo_ctx* c, const u64* ksk, const u64* in, u64* out) {
    const tfo_params* p = &c->p;
    u64 N = p->N, n = p->n, qKS = p->qKS, bKS = p->baseKS, dKS = p->dKS;
    memset(out, 0, sizeof(u64) * n);
    out[n] = in[N];
    for (u64 i = 0; i < N; i++) {
        u64 atmp = in[i];
        for (u64 j = 0; j < dKS; j++, atmp /= bKS) {
            u64 a0 = atmp % bKS;
            const u64* row = ksk + (((i * bKS + a0) * dKS + j) * (n + 1));
            for (u64 k = 0; k <= n; k++)
                out[k] = submod(out[k], row[k], qKS);
        }
    }
}
void tfo_key_switch(const tfo_ctx* c, const u64* ksk, int batch, const u64* in, u64* out) {
    u64 N = c->p.N, n = c->p.n;
#pragma omp parallel for schedule(dynamic)
    for (int b = 0; b < batch; b++)
        key_switch_one(c, ksk, in + (size_t)b * (N + 1), out + (size_t)b * (n + 1));
}

/* binfhe-base-scheme.cpp:102-107 (scalar) == contract of MKMSwitch_CUDA (bootstrapping.cu:73-118) */
void tfo_mkmswitch(const tfo_ctx* c, const u64* ksk, int batch, const u64* in, u64 fmod, u64* out) {
    const tfo_params* p = &c->p;
    u64 N = p->N, n = p->n;
#pragma omp parallel
    {
        u64* ms = (u64*)malloc(sizeof(u64) * (N + 1));
        u64* ks = (u64*)malloc(sizeof(u64) * (n + 1));
#pragma omp for schedule(dynamic)
        for (int b = 0; b < batch; b++) {
            for (u64 i = 0; i <= N; i++)
                ms[i] = tfo_round_qQ(in[(size_t)b * (N + 1) + i], p->qKS, p->Q);
            key_switch_one(c, ksk, ms, ks);
            for (u64 i = 0; i <= n; i++)
                out[(size_t)b * (n + 1) + i] = tfo_round_qQ(ks[i], fmod, p->qKS);
        }
        free(ms);
        free(ks);
    }
}

/* binfhe-base-scheme.cpp:1087-1138; gate constants rgsw-cryptoparameters.h:130-137 */
void tfo_init_acc_gate(const tfo_ctx* c, int gate, u64 b, u64 q, u64* acc) {
    const tfo_params* p = &c->p;
    u64 N = p->N, Q = p->Q;
    static const u64 mult[6] = {5, 7, 1, 3, 5, 1}; /* OR AND NOR NAND XOR_FAST XNOR_FAST, times (params q >> 3) */
    u64 q1 = mult[gate] * (p->q >> 3);
    u64 qHalf = q >> 1;
    u64 q2 = addmod(q1, qHalf, q);
    u64 Q8 = Q / 8 + 1, Q8Neg = Q - Q8;
    u64 factor = 2 * N / q;
    memset(acc, 0, sizeof(u64) * 2 * N);
    u64* m = acc + N;
    for (u64 j = 0; j < qHalf; j++) {
        u64 temp = submod(b % q, j, q);
        if (q1 < q2)
            m[j * factor] = ((temp >= q1) && (temp < q2)) ? Q8Neg : Q8;
        else
            m[j * factor] = ((temp >= q2) && (temp < q1)) ? Q8 : Q8Neg;
    }
}

/* binfhe-base-scheme.cpp:1147-1185: m[j*factor] = (Q / fmod) * f((b - j) mod ctMod) */
static void init_acc_func(const tfo_ctx* c, const u64* table, u64 b, u64 ctmod, u64 fmod, u64* acc) {
    u64 N = c->p.N, Q = c->p.Q;
    u64 factor = 2 * N / ctmod, scale = Q / fmod;
    memset(acc, 0, sizeof(u64) * 2 * N);
    u64* m = acc + N;
    for (u64 j = 0; j < (ctmod >> 1); j++) {
        u64 temp      = submod(b % ctmod, j, ctmod);
        m[j * factor] = scale * table[temp];
    }
}

/* ------------------------------------------------------------------------------------------------------ */
/* batched operations                                                                                      */
/* ------------------------------------------------------------------------------------------------------ */
static void lwe_not(u64 n, u64 q, const u64* in, u64* out) { /* binfhe-base-scheme.cpp:145-158 */
    for (u64 i = 0; i < n; i++)
        out[i] = in[i] == 0 ? 0 : q - in[i];
    out[n] = submod(q >> 2, in[n], q);
}

/* binfhe-base-scheme.cpp:58-108 (scalar) / :598-677 (batched) */
int tfo_eval_bin_gate(const tfo_ctx* c, const u64* bk, const u64* ksk, int gate, int batch, const u64* ct1,
                      const u64* ct2, u64 q, u64* out) {
    const tfo_params* p = &c->p;
    u64 N = p->N, n = p->n, Q = p->Q;
    size_t W = n + 1;
    if (batch <= 0)
        return -1;
    if (gate == TFO_XOR || gate == TFO_XNOR) {
        size_t tot = (size_t)batch * W;
        u64* n1 = (u64*)malloc(sizeof(u64) * tot);
        u64* n2 = (u64*)malloc(sizeof(u64) * tot);
        u64* a1 = (u64*)malloc(sizeof(u64) * tot);
        u64* a2 = (u64*)malloc(sizeof(u64) * tot);
        for (int b = 0; b < batch; b++) {
            lwe_not(n, q, ct1 + b * W, n1 + b * W);
            lwe_not(n, q, ct2 + b * W, n2 + b * W);
        }
        tfo_eval_bin_gate(c, bk, ksk, TFO_AND, batch, ct1, n2, q, a1);
        tfo_eval_bin_gate(c, bk, ksk, TFO_AND, batch, n1, ct2, q, a2);
        tfo_eval_bin_gate(c, bk, ksk, TFO_OR, batch, a1, a2, q, out);
        if (gate == TFO_XNOR)
            for (int b = 0; b < batch; b++) {
                lwe_not(n, q, out + b * W, n1 + b * W);
                memcpy(out + b * W, n1 + b * W, sizeof(u64) * W);
            }
        free(n1); free(n2); free(a1); free(a2);
        return 0;
    }
    u64* prep = (u64*)malloc(sizeof(u64) * batch * W);
    u64* acc  = (u64*)malloc(sizeof(u64) * (size_t)batch * 2 * N);
    u64* ext  = (u64*)malloc(sizeof(u64) * (size_t)batch * (N + 1));
    u64* avec = (u64*)malloc(sizeof(u64) * (size_t)batch * n);
    for (int b = 0; b < batch; b++) {
        const u64 *x = ct1 + b * W, *y = ct2 + b * W;
        u64* z = prep + b * W;
        for (size_t i = 0; i < W; i++) {
            if (gate == TFO_XOR_FAST || gate == TFO_XNOR_FAST) {
                u64 t = submod(x[i], y[i], q);
                z[i]  = addmod(t, t, q);
            }
            else
                z[i] = addmod(x[i], y[i], q);
        }
        tfo_init_acc_gate(c, gate, z[n], q, acc + (size_t)b * 2 * N);
        memcpy(avec + (size_t)b * n, z, sizeof(u64) * n);
    }
    tfo_eval_acc(c, bk, batch, avec, q, acc);
    u64 Q8 = Q / 8 + 1;
    for (int b = 0; b < batch; b++) {
        memcpy(ext + (size_t)b * (N + 1), acc + (size_t)b * 2 * N, sizeof(u64) * N);
        ext[(size_t)b * (N + 1) + N] = addmod(Q8, acc[(size_t)b * 2 * N + N], Q);
    }
    tfo_mkmswitch(c, ksk, batch, ext, q, out);
    free(prep); free(acc); free(ext); free(avec);
    return 0;
}

/* binfhe-base-scheme.cpp:534-592 (scalar) / :1194-1211 (batched) */
int tfo_bootstrap_func(const tfo_ctx* c, const u64* bk, const u64* ksk, int batch, const u64* ct, u64 ctmod,
                       const u64* table, int per_ct, u64 fmod, u64* out) {
    const tfo_params* p = &c->p;
    u64 N = p->N, n = p->n;
    size_t W = n + 1;
    u64* acc  = (u64*)malloc(sizeof(u64) * (size_t)batch * 2 * N);
    u64* ext  = (u64*)malloc(sizeof(u64) * (size_t)batch * (N + 1));
    u64* avec = (u64*)malloc(sizeof(u64) * (size_t)batch * n);
    for (int b = 0; b < batch; b++) {
        init_acc_func(c, per_ct ? table + (size_t)b * ctmod : table, ct[b * W + n], ctmod, fmod,
                      acc + (size_t)b * 2 * N);
        memcpy(avec + (size_t)b * n, ct + b * W, sizeof(u64) * n);
    }
    tfo_eval_acc(c, bk, batch, avec, ctmod, acc);
    for (int b = 0; b < batch; b++) {
        memcpy(ext + (size_t)b * (N + 1), acc + (size_t)b * 2 * N, sizeof(u64) * N);
        ext[(size_t)b * (N + 1) + N] = acc[(size_t)b * 2 * N + N];
    }
    tfo_mkmswitch(c, ksk, batch, ext, fmod, out);
    free(acc); free(ext); free(avec);
    return 0;
}

/* binfhe-base-scheme.cpp:162-186 */
static int check_input_function(const u64* lut, size_t len, u64 mod) {
    int ret = 0;
    if (lut[0] == mod - lut[len / 2]) {
        for (size_t i = 1; i < len / 2; i++)
            if (lut[i] != mod - lut[len / 2 + i]) {
                ret = 2;
                break;
            }
    }
    else if (lut[0] == lut[len / 2]) {
        ret = 1;
        for (size_t i = 1; i < len / 2; i++)
            if (lut[i] != lut[len / 2 + i]) {
                ret = 2;
                break;
            }
    }
    else
        ret = 2;
    return ret;
}

/* binfhe-base-scheme.cpp:189-267 (scalar) / :679-924 (batched; the LUT_vec overload classifies with LUT_vec[0]) */
int tfo_eval_func(const tfo_ctx* c, const u64* bk, const u64* ksk, int batch, const u64* ct, u64 q, const u64* lut,
                  size_t lut_len, int per_ct, u64* out) {
    const tfo_params* p = &c->p;
    u64 n = p->n, beta = p->beta;
    size_t W = n + 1, tot = (size_t)batch * W;
    if (batch <= 0 || lut_len != q)
        return -1;
    int prop = check_input_function(lut, lut_len, q);
    size_t ntab = per_ct ? (size_t)batch : 1;
    u64* ct1 = (u64*)malloc(sizeof(u64) * tot);
    memcpy(ct1, ct, sizeof(u64) * tot);
    int rc = 0;
    if (prop == 0) {
        for (int b = 0; b < batch; b++)
            ct1[b * W + n] = addmod(ct1[b * W + n], beta, q);
        rc = tfo_bootstrap_func(c, bk, ksk, batch, ct1, q, lut, per_ct, q, out);
    }
    else if (prop == 2) {
        if (q > p->N) {
            free(ct1);
            return -2;
        }
        u64 dq = q << 1;
        u64* ct2 = (u64*)malloc(sizeof(u64) * tot);
        u64* ct3 = (u64*)malloc(sizeof(u64) * tot);
        u64* f0  = (u64*)malloc(sizeof(u64) * dq);
        u64* t2  = (u64*)malloc(sizeof(u64) * dq * ntab);
        memcpy(ct2, ct1, sizeof(u64) * tot);
        for (int b = 0; b < batch; b++)
            ct2[b * W + n] = addmod(ct2[b * W + n], beta, dq);
        for (u64 x = 0; x < dq; x++)
            f0[x] = (x < dq / 2) ? dq - dq / 4 : dq / 4;
        rc = tfo_bootstrap_func(c, bk, ksk, batch, ct2, dq, f0, 0, dq, ct3);
        for (int b = 0; b < batch; b++) {
            for (size_t i = 0; i < W; i++)
                ct3[b * W + i] = submod(ct1[b * W + i], ct3[b * W + i], dq);
            ct3[b * W + n] = addmod(ct3[b * W + n], beta, dq);
            ct3[b * W + n] = submod(ct3[b * W + n], q >> 1, dq);
        }
        for (size_t t = 0; t < ntab; t++)
            for (u64 x = 0; x < dq; x++) {
                const u64* L = lut + t * q; /* LUT2 = LUT || LUT */
                t2[t * dq + x] = (x < dq / 2) ? L[x % q] : dq - L[(x - dq / 2) % q];
            }
        rc |= tfo_bootstrap_func(c, bk, ksk, batch, ct3, dq, t2, per_ct, dq, out);
        for (size_t i = 0; i < tot; i++)
            out[i] %= q;
        free(ct2); free(ct3); free(f0); free(t2);
    }
    else {
        u64* ct2 = (u64*)malloc(sizeof(u64) * tot);
        u64* f0  = (u64*)malloc(sizeof(u64) * q);
        u64* t1  = (u64*)malloc(sizeof(u64) * q * ntab);
        for (int b = 0; b < batch; b++)
            ct1[b * W + n] = addmod(ct1[b * W + n], beta, q);
        for (u64 x = 0; x < q; x++)
            f0[x] = (x < q / 2) ? q - q / 4 : q / 4;
        rc = tfo_bootstrap_func(c, bk, ksk, batch, ct1, q, f0, 0, q, ct2);
        for (int b = 0; b < batch; b++) {
            for (size_t i = 0; i < W; i++)
                ct2[b * W + i] = submod(ct[b * W + i], ct2[b * W + i], q);
            ct2[b * W + n] = addmod(ct2[b * W + n], beta, q);
            ct2[b * W + n] = submod(ct2[b * W + n], q >> 2, q);
        }
        for (size_t t = 0; t < ntab; t++)
            for (u64 x = 0; x < q; x++) {
                const u64* L = lut + t * q;
                t1[t * q + x] = (x < q / 2) ? L[x] : q - L[x - q / 2];
            }
        rc |= tfo_bootstrap_func(c, bk, ksk, batch, ct2, q, t1, per_ct, q, out);
        free(ct2); free(f0); free(t1);
    }
    free(ct1);
    return rc;
}

/* binfhe-base-scheme.cpp:270-311 (scalar) / :926-987 (batched) */
int tfo_eval_floor(const tfo_ctx* c, const u64* bk, const u64* ksk, int batch, const u64* ct, u64 mod,
                   uint32_t roundbits, u64* out) {
    const tfo_params* p = &c->p;
    u64 n = p->n, beta = p->beta;
    size_t W = n + 1, tot = (size_t)batch * W;
    if (batch <= 0)
        return -1;
    u64 q = roundbits == 0 ? p->q : beta * 2 * (1ULL << roundbits);
    u64* ct1 = out;
    u64* cq  = (u64*)malloc(sizeof(u64) * tot);
    u64* r   = (u64*)malloc(sizeof(u64) * tot);
    u64* f   = (u64*)malloc(sizeof(u64) * q);
    memcpy(ct1, ct, sizeof(u64) * tot);
    for (int b = 0; b < batch; b++)
        ct1[b * W + n] = addmod(ct1[b * W + n], beta, mod);
    for (size_t i = 0; i < tot; i++)
        cq[i] = ct1[i] % q;
    for (u64 x = 0; x < q; x++)
        f[x] = (x < q / 2) ? mod - q / 4 : q / 4;
    int rc = tfo_bootstrap_func(c, bk, ksk, batch, cq, q, f, 0, mod, r);
    for (size_t i = 0; i < tot; i++)
        ct1[i] = submod(ct1[i], r[i], mod);
    for (size_t i = 0; i < tot; i++)
        cq[i] = ct1[i] % q;
    for (u64 x = 0; x < q; x++) {
        if (x < q / 4)
            f[x] = mod - q / 2 - x;
        else if (x < 3 * q / 4)
            f[x] = x;
        else
            f[x] = mod + q / 2 - x;
    }
    rc |= tfo_bootstrap_func(c, bk, ksk, batch, cq, q, f, 0, mod, r);
    for (size_t i = 0; i < tot; i++)
        ct1[i] = submod(ct1[i], r[i], mod);
    free(cq); free(r); free(f);
    return rc;
}

/* gadget base the scalar EvalSign / EvalDecomp switch to once the modulus has shrunk to `mod`
 * (binfhe-base-scheme.cpp:342-349, 411-418); 0 = keep the current base */
static uint32_t dynamic_base(u64 mod) {
    uint32_t binLog = (uint32_t)ceil(log2((double)mod));
    if (binLog <= 17)
        return 1u << 27;
    if (binLog <= 26)
        return 1u << 18;
    return 0;
}

/* key set for gadget base `base` among the nk loaded ones; -1 = "No key [..] found in the map" */
static int find_key(int nk, const tfo_ctx* const* cs, uint32_t base) {
    for (int k = 0; k < nk; k++)
        if (cs[k]->p.baseG == base)
            return k;
    return -1;
}

/* binfhe-base-scheme.cpp:314-372 (scalar EvalSign over the key map) / :989-1037 (batched, single key).
 * cs[0] / bks[0] / ksks[0] is the context's own key set (curBase); with nk == 3 the gadget base follows the
 * shrinking modulus ("if (EKs.size() == 3)"), otherwise the first key set is used throughout. */
int tfo_eval_sign_dyn(int nk, const tfo_ctx* const* cs, const u64* const* bks, const u64* const* ksks, int batch,
                      const u64* ct, u64 mod, u64* out) {
    const tfo_params* p = &cs[0]->p;
    u64 n = p->n, beta = p->beta, q = p->q;
    size_t W = n + 1, tot = (size_t)batch * W;
    if (batch <= 0 || nk <= 0)
        return -1;
    u64* cur = (u64*)malloc(sizeof(u64) * tot);
    u64* nxt = (u64*)malloc(sizeof(u64) * tot);
    memcpy(cur, ct, sizeof(u64) * tot);
    int rc = 0, k = 0;
    while (mod > q && rc == 0) {
        rc |= tfo_eval_floor(cs[k], bks[k], ksks[k], batch, cur, mod, 0, nxt);
        u64 newmod = mod / q * 2 * beta;
        tfo_mod_switch(batch, W, nxt, mod, newmod, cur);
        mod = newmod;
        if (nk == 3) {
            uint32_t base = dynamic_base(mod);
            if (base && (k = find_key(nk, cs, base)) < 0)
                rc = -1;
        }
    }
    if (rc == 0) {
        for (int b = 0; b < batch; b++)
            cur[b * W + n] = addmod(cur[b * W + n], beta, mod);
        u64* f3 = (u64*)malloc(sizeof(u64) * mod);
        for (u64 x = 0; x < mod; x++)
            f3[x] = (x < mod / 2) ? q / 4 : q - q / 4;
        rc |= tfo_bootstrap_func(cs[k], bks[k], ksks[k], batch, cur, mod, f3, 0, q, out);
        for (int b = 0; b < batch; b++)
            out[b * W + n] = submod(out[b * W + n], q >> 2, q);
        free(f3);
    }
    free(cur); free(nxt);
    return rc;
}

int tfo_eval_sign(const tfo_ctx* c, const u64* bk, const u64* ksk, int batch, const u64* ct, u64 mod, u64* out) {
    return tfo_eval_sign_dyn(1, &c, &bk, &ksk, batch, ct, mod, out);
}

/* binfhe-base-scheme.cpp:375-434 (scalar EvalDecomp over the key map) / :1039-1085 (batched, single key) */
int tfo_eval_decomp_dyn(int nk, const tfo_ctx* const* cs, const u64* const* bks, const u64* const* ksks, int batch,
                        const u64* ct, u64 mod, int max_digits, u64* out, u64* out_mods) {
    const tfo_params* p = &cs[0]->p;
    u64 n = p->n, beta = p->beta, q = p->q;
    size_t W = n + 1, tot = (size_t)batch * W;
    if (batch <= 0 || nk <= 0 || mod <= q)
        return -1;
    u64* cur = (u64*)malloc(sizeof(u64) * tot);
    u64* nxt = (u64*)malloc(sizeof(u64) * tot);
    memcpy(cur, ct, sizeof(u64) * tot);
    int nd = 0, rc = 0, k = 0;
    while (mod > q && rc == 0) {
        if (nd >= max_digits) { rc = -1; break; }
        for (int b = 0; b < batch; b++)
            for (size_t i = 0; i < W; i++)
                out[((size_t)b * max_digits + nd) * W + i] = cur[b * W + i] % q;
        out_mods[nd++] = q;
        rc |= tfo_eval_floor(cs[k], bks[k], ksks[k], batch, cur, mod, 0, nxt);
        u64 newmod = mod / q * 2 * beta;
        tfo_mod_switch(batch, W, nxt, mod, newmod, cur);
        mod = newmod;
        if (nk == 3) {
            uint32_t base = dynamic_base(mod);
            if (base && (k = find_key(nk, cs, base)) < 0)
                rc = -1;
        }
    }
    if (nd >= max_digits)
        rc = -1;
    if (rc == 0) {
        for (int b = 0; b < batch; b++)
            memcpy(out + ((size_t)b * max_digits + nd) * W, cur + b * W, sizeof(u64) * W);
        out_mods[nd++] = mod;
    }
    free(cur); free(nxt);
    return rc ? -1 : nd;
}

int tfo_eval_decomp(const tfo_ctx* c, const u64* bk, const u64* ksk, int batch, const u64* ct, u64 mod,
                    int max_digits, u64* out, u64* out_mods) {
    return tfo_eval_decomp_dyn(1, &c, &bk, &ksk, batch, ct, mod, max_digits, out, out_mods);
}

/* lwe-operation.cu:50-141, exact-integer restatement */
int tfo_mul_matrix(const tfo_ctx* c, int in, int outc, const u64* ct, const i64* M, u64 modulus, u64* out) {
    u64 n = c->p.n;
    size_t W = n + 1;
    if (in <= 0 || outc <= 0)
        return -1;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < outc; i++)
        for (size_t w = 0; w < W; w++) {
            u64 acc = 0;
            for (int k = 0; k < in; k++) {
                i64 mv = M[(size_t)k * outc + i] % (i64)modulus;
                u64 mu = (u64)(mv < 0 ? mv + (i64)modulus : mv);
                acc    = addmod(acc, mulmod(ct[(size_t)k * W + w] % modulus, mu, modulus), modulus);
            }
            out[(size_t)i * W + w] = acc;
        }
    return 0;
}

int tfo_num_threads(void) {
    return omp_get_max_threads();
}

/* bench.py sets the thread count explicitly: launchers such as torchrun export OMP_NUM_THREADS=1 */
void tfo_set_num_threads(int n) {
    if (n > 0)
        omp_set_num_threads(n);
}
