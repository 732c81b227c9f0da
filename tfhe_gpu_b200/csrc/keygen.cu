// evaluation-key generation on the GPU, straight into device memory in the reference's
// element order (so the result feeds tfhe_b200_setup(..., key_space = TFHE_B200_DEVICE) and never exists on the host).
// Follows BinFHEScheme::KeyGen (binfhe-base-scheme.cpp:38-57): key-switching key (lwe-pke.cpp:218-295), RingGSW
// bootstrapping key for CGGI (rgsw-acc-cggi.cpp:43-75, 213-240) and DM (rgsw-acc-dm.cpp:44-76, 153-209).
// Randomness: ChaCha20 in counter mode under a 256-bit key (from the caller or the operating system), separate derived
// keys for the public masks and the secret errors (the reference uses a BLAKE2-based PRNG seeded from the OS,
// core/include/math/distributiongenerator.h:86-130); uniform residues by Lemire's multiply-shift with rejection
// (unbiased), errors from an inversion table of the discrete Gaussian D_{Z, 3.19} (64-bit cumulative probabilities).  Key generation is randomised, so it cannot be bit-compared with the reference: the tests check
// decryption correctness of everything evaluated under these keys and the noise distribution of the key material.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "engine.cuh"

namespace tfhe_b200 {

namespace {

// ChaCha20 block function (RFC 8439 state layout: constants | 256-bit key | 128 bits of counter / nonce) used as a
// counter-mode PRF: every random element is a function of (key, stream, index, attempt), so the result does not depend
// on the launch shape, and with a 256-bit key from a CSPRNG the published masks reveal nothing about the error stream
// (the evaluation keys go to an untrusted evaluator: a generator that can be inverted or whose seed can be searched
// exposes every error term and with it the secret key).
struct PrfKey {
    u32 k[8];
};
__host__ __device__ __forceinline__ u32 rotl32(u32 x, int r) {
    return (x << r) | (x >> (32 - r));
}
#define CHACHA_QR(a, b, c, d)                    \
    a += b; d ^= a; d = rotl32(d, 16);           \
    c += d; b ^= c; b = rotl32(b, 12);           \
    a += b; d ^= a; d = rotl32(d, 8);            \
    c += d; b ^= c; b = rotl32(b, 7);
// first 128 bits of the ChaCha20 block with counter/nonce words (c0, c1, c2, c3)
__host__ __device__ __forceinline__ void chacha20_block(const PrfKey& K, u32 c0, u32 c1, u32 c2, u32 c3, u32 out[16]) {
    u32 x[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, K.k[0], K.k[1], K.k[2], K.k[3],
                 K.k[4], K.k[5], K.k[6], K.k[7], c0, c1, c2, c3};
    u32 w[16];
#pragma unroll
    for (int i = 0; i < 16; i++)
        w[i] = x[i];
#pragma unroll
    for (int r = 0; r < 10; r++) {
        CHACHA_QR(w[0], w[4], w[8], w[12]) CHACHA_QR(w[1], w[5], w[9], w[13])
        CHACHA_QR(w[2], w[6], w[10], w[14]) CHACHA_QR(w[3], w[7], w[11], w[15])
        CHACHA_QR(w[0], w[5], w[10], w[15]) CHACHA_QR(w[1], w[6], w[11], w[12])
        CHACHA_QR(w[2], w[7], w[8], w[13]) CHACHA_QR(w[3], w[4], w[9], w[14])
    }
#pragma unroll
    for (int i = 0; i < 16; i++)
        out[i] = w[i] + x[i];
}
// 128 random bits for (stream, index, attempt)
__device__ __forceinline__ uint4 rnd128(const PrfKey& key, u32 stream, u64 idx, u32 attempt) {
    u32 o[16];
    chacha20_block(key, (u32)idx, (u32)(idx >> 32), stream, attempt, o);
    return make_uint4(o[0], o[1], o[2], o[3]);
}
__device__ __forceinline__ u64 uniform_mod(const PrfKey& seed, u32 stream, u64 idx, u64 m) {
    for (u32 attempt = 0;; attempt++) {
        const uint4 r = rnd128(seed, stream, idx, attempt);
        const u64 x = ((u64)r.y << 32) | r.x;
        const u64 lo = x * m, hi = __umul64hi(x, m);
        if (lo < m) {
            const u64 thresh = (0 - m) % m;
            if (lo < thresh)
                continue;
        }
        return hi;
    }
}
// discrete Gaussian: cdt[k] = floor(2^64 * P(|X| <= k)), the last entry saturated
__device__ __forceinline__ int gauss(const PrfKey& seed, u32 stream, u64 idx, const u64* cdt, int len) {
    const uint4 r = rnd128(seed, stream, idx, 0);
    const u64 x = ((u64)r.y << 32) | r.x;
    int k = 0;
    while (k < len - 1 && x >= cdt[k])
        k++;
    return (r.z & 1) ? -k : k;
}
__device__ __forceinline__ u64 signed_mod(int v, u64 m) {
    if (v >= 0)
        return (u64)v % m;
    const u64 a = (u64)(-(long long)v) % m;
    return a ? m - a : 0;
}

enum { STREAM_KSK_A = 1, STREAM_KSK_E = 2, STREAM_BK_A = 3, STREAM_BK_E = 4 };

// in-place negacyclic forward NTT of x[N] in shared memory (Cooley-Tukey, bit-reversed output, twiddles W[m + i] =
// psi^bitrev(m + i) in Montgomery form): the transform of transformnat-impl.h:298-341
__device__ void ntt_forward_smem(u64* x, const u64* WM, u32 N, const ModCtx<u64>& M) {
    for (u32 m = 1, t = N >> 1; m < N; m <<= 1, t >>= 1) {
        __syncthreads();
        for (u32 b = threadIdx.x; b < N / 2; b += blockDim.x) {
            const u32 i = b / t, jj = b - i * t, j = 2 * i * t + jj;
            const u64 U = x[j], V = M.mont_mul(x[j + t], WM[m + i]);
            x[j] = M.add(U, V);
            x[j + t] = M.sub(U, V);
        }
    }
    __syncthreads();
}

__global__ void sk_ntt_kernel(u64* skM, const signed char* sk_ring, const u64* WM, u32 N, ModCtx<u64> M) {
    extern __shared__ u64 sx[];
    for (u32 k = threadIdx.x; k < N; k += blockDim.x)
        sx[k] = signed_mod(sk_ring[k], M.Q);
    ntt_forward_smem(sx, WM, N, M);
    for (u32 k = threadIdx.x; k < N; k += blockDim.x)
        skM[k] = M.mont_mul(sx[k], M.r2);   // Montgomery form
}

// one CTA per (RGSW ciphertext, row): comp 0 = a (+ message on even rows), comp 1 = NTT(e) + a * sk (+ message on odd
// rows); `a` is drawn directly in evaluation form (uniform either way)
struct BKArgs {
    u64* bk;
    const signed char* sk_lwe;   // [n]
    const u64* skM;              // NTT(sk_ring), Montgomery form
    const u64* WM;               // twiddles, Montgomery form
    const u64* psiM;             // psi^x, x < 2N, Montgomery form (DM monomials)
    const u64* cdt;
    PrfKey key_mask, key_err;    // independent PRF keys for the public masks and the secret errors
    u64 baseG;
    u32 n, N, d2, throwd, method, baseR, digitsR, q;
    int cdt_len;
    ModCtx<u64> M;
};
__global__ void bk_gen_kernel(BKArgs A) {
    extern __shared__ u64 se[];
    const u32 N = A.N, d2 = A.d2;
    const u64 Q = A.M.Q;
    const u64 rowid = blockIdx.x;
    const u32 r = (u32)(rowid % d2);
    const u64 keyidx = rowid / d2;
    u64* out = A.bk + rowid * 2 * N;
    // message of this RGSW ciphertext
    bool has_msg = false, neg = false;
    u32 mm = 0;
    if (A.method == TFHE_B200_METHOD_GINX) {
        const u32 key = (u32)(keyidx / A.n), i = (u32)(keyidx % A.n);
        const int s = A.sk_lwe[i];
        has_msg = key == 0 ? s == 1 : s == -1;                      // rgsw-acc-cggi.cpp:57-71
    }
    else {
        const u32 k = (u32)(keyidx % A.digitsR), a0 = (u32)((keyidx / A.digitsR) % A.baseR);
        const u32 i = (u32)(keyidx / ((u64)A.digitsR * A.baseR));
        if (a0 == 0) {                                               // never read (rgsw-acc-dm.cpp:60-70)
            for (u32 x = threadIdx.x; x < 2 * N; x += blockDim.x)
                out[x] = 0;
            return;
        }
        long long dig = 1;
        for (u32 x = 0; x < k; x++)
            dig *= A.baseR;
        long long m = (long long)A.sk_lwe[i] * (long long)a0 * dig;  // s_i * a0 * baseR^k
        const long long q = A.q;
        long long e = (((m % q) + q) % q) * (long long)(2 * N / A.q);  // X^e, e < 2N (rgsw-acc-dm.cpp:157-170)
        has_msg = true;
        if (e >= (long long)N) {
            e -= N;
            neg = true;
        }
        mm = (u32)e;
    }
    u64 G = 1;
    for (u32 x = 0; x < (r >> 1) + A.throwd; x++)
        G = (u64)(((unsigned __int128)G * (A.baseG % Q)) % Q);
    for (u32 k = threadIdx.x; k < N; k += blockDim.x)
        se[k] = signed_mod(gauss(A.key_err, STREAM_BK_E, rowid * N + k, A.cdt, A.cdt_len), Q);
    ntt_forward_smem(se, A.WM, N, A.M);
    const u32 logN = 31 - __clz(N);
    for (u32 k = threadIdx.x; k < N; k += blockDim.x) {
        const u64 a = uniform_mod(A.key_mask, STREAM_BK_A, rowid * N + k, Q);
        u64 msg = 0;
        if (has_msg) {
            if (A.method == TFHE_B200_METHOD_GINX)
                msg = G;                                             // constant polynomial: the same value in every slot
            else {
                const u32 br = __brev(k) >> (32 - logN);
                const u64 mono = A.M.mont_mul(A.psiM[((2 * br + 1) * mm) & (2 * N - 1)], 1);   // psi^((2 br + 1) mm), plain
                msg = (u64)(((unsigned __int128)mono * G) % Q);
                if (neg && msg)
                    msg = Q - msg;
            }
        }
        u64 c0 = a, c1 = A.M.add(se[k], A.M.mont_mul(a, A.skM[k]));
        if (r & 1)
            c1 = A.M.add(c1, msg);
        else
            c0 = A.M.add(c0, msg);
        out[k] = c0;
        out[N + k] = c1;
    }
}

// one CTA per key-switching row (i, j, k): a uniform mod qKS, b = <a, s> + e + s_ring[i] * j * baseKS^k
struct KSKArgs {
    u64* ksk;
    const signed char* sk_lwe;
    const signed char* sk_ring;
    const u64* cdt;
    PrfKey key_mask, key_err;
    u64 qKS;
    u32 n, N, baseKS, dKS;
    int cdt_len;
};
__global__ void ksk_gen_kernel(KSKArgs A) {
    __shared__ u64 red[128];
    const u64 rowid = blockIdx.x, qKS = A.qKS;
    const u32 k = (u32)(rowid % A.dKS), j = (u32)((rowid / A.dKS) % A.baseKS), i = (u32)(rowid / ((u64)A.dKS * A.baseKS));
    u64* row = A.ksk + rowid * (A.n + 1);
    u64 part = 0;
    for (u32 t = threadIdx.x; t < A.n; t += blockDim.x) {
        const u64 a = uniform_mod(A.key_mask, STREAM_KSK_A, rowid * A.n + t, qKS);
        row[t] = a;
        const int s = A.sk_lwe[t];
        if (s == 1)
            part += a;
        else if (s == -1)
            part += qKS - a;
        part %= qKS;
    }
    red[threadIdx.x] = part;
    __syncthreads();
    for (int off = 64; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off)
            red[threadIdx.x] = (red[threadIdx.x] + red[threadIdx.x + off]) % qKS;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        u64 dig = 1;
        for (u32 x = 0; x < k; x++)
            dig = (u64)(((unsigned __int128)dig * A.baseKS) % qKS);
        u64 msg = (u64)(((unsigned __int128)(j % qKS) * dig) % qKS);
        const int s = A.sk_ring[i];
        msg = s == 0 ? 0 : (s == 1 ? msg : (msg ? qKS - msg : 0));
        const u64 e = signed_mod(gauss(A.key_err, STREAM_KSK_E, rowid, A.cdt, A.cdt_len), qKS);
        row[A.n] = (red[0] + e + msg) % qKS;
    }
}

std::vector<u64> gaussian_cdt(double sigma) {
    const int len = (int)std::ceil(sigma * 13) + 2;
    std::vector<long double> pmf(len);
    long double S = 0;
    for (int k = 0; k < len; k++) {
        pmf[k] = expl(-(long double)k * k / (2.0L * sigma * sigma));
        S += k == 0 ? pmf[k] : 2 * pmf[k];
    }
    std::vector<u64> cdt(len);
    long double c = 0;
    for (int k = 0; k < len; k++) {
        c += (k == 0 ? pmf[k] : 2 * pmf[k]) / S;
        long double v = c * 18446744073709551616.0L;
        cdt[k] = v >= 18446744073709551615.0L ? ~0ULL : (u64)v;
    }
    cdt[len - 1] = ~0ULL;
    return cdt;
}

u32 bitrev_k(u32 x, u32 bits) {
    u32 r = 0;
    for (u32 i = 0; i < bits; i++) {
        r = (r << 1) | (x & 1);
        x >>= 1;
    }
    return r;
}

thread_local std::string g_kerr;

}  // namespace

const char* keygen_last_error() {
    return g_kerr.c_str();
}

#define KG_TRY(x)                                                                                       \
    do {                                                                                                \
        cudaError_t e__ = (x);                                                                          \
        if (e__ != cudaSuccess) {                                                                       \
            g_kerr = std::string(#x) + " failed: " + cudaGetErrorString(e__);                           \
            rc = TFHE_B200_ECUDA;                                                                       \
            goto done;                                                                                  \
        }                                                                                               \
    } while (0)

int keygen_device(const tfhe_b200_params& p, const signed char* sk_lwe, const signed char* sk_ring,
                  const unsigned char key[32], int device, u64* bk_dev, u64* ksk_dev) {
    int rc = 0;
    // two independent sub-keys (mask stream, error stream) derived from the caller's 256-bit key: one ChaCha20 block under
    // the master key with a derivation label as nonce yields 512 bits
    PrfKey master, key_mask, key_err;
    memcpy(master.k, key, 32);
    {
        u32 o[16];
        chacha20_block(master, 0x6b646600u /* "kdf" */, 0x62323030u /* "b200" */, 0x74666865u /* "tfhe" */, 0, o);
        memcpy(key_mask.k, o, 32);
        memcpy(key_err.k, o + 8, 32);
    }
    const u32 N = p.N, n = p.n;
    const ModCtx<u64> M = make_modctx<u64>(p.Q);
    const u32 logN = 31 - __builtin_clz(N);
    std::vector<u64> WM(N), psiM(2 * N);
    {
        const u64 psi = p.psi % p.Q;
        u64 x = 1;
        for (u32 k = 0; k < N; k++) {
            WM[bitrev_k(k, logN)] = to_mont<u64>(x, M);
            x = h_mulmod(x, psi, p.Q);
        }
        x = 1;
        for (u32 k = 0; k < 2 * N; k++) {
            psiM[k] = to_mont<u64>(x, M);
            x = h_mulmod(x, psi, p.Q);
        }
    }
    const std::vector<u64> cdt = gaussian_cdt(3.19);   // binfhecontext.cpp:134 STD_DEV
    signed char *d_s = nullptr, *d_sN = nullptr;
    u64 *d_W = nullptr, *d_psi = nullptr, *d_cdt = nullptr, *d_skM = nullptr;
    cudaStream_t st = nullptr;
    KG_TRY(cudaSetDevice(device));
    KG_TRY(cudaStreamCreate(&st));
    KG_TRY(cudaMalloc((void**)&d_s, n));
    KG_TRY(cudaMalloc((void**)&d_sN, N));
    KG_TRY(cudaMalloc((void**)&d_W, N * 8));
    KG_TRY(cudaMalloc((void**)&d_psi, 2 * N * 8));
    KG_TRY(cudaMalloc((void**)&d_cdt, cdt.size() * 8));
    KG_TRY(cudaMalloc((void**)&d_skM, N * 8));
    KG_TRY(cudaMemcpyAsync(d_s, sk_lwe, n, cudaMemcpyHostToDevice, st));
    KG_TRY(cudaMemcpyAsync(d_sN, sk_ring, N, cudaMemcpyHostToDevice, st));
    KG_TRY(cudaMemcpyAsync(d_W, WM.data(), N * 8, cudaMemcpyHostToDevice, st));
    KG_TRY(cudaMemcpyAsync(d_psi, psiM.data(), 2 * N * 8, cudaMemcpyHostToDevice, st));
    KG_TRY(cudaMemcpyAsync(d_cdt, cdt.data(), cdt.size() * 8, cudaMemcpyHostToDevice, st));
    sk_ntt_kernel<<<1, 256, N * 8, st>>>(d_skM, d_sN, d_W, N, M);
    KG_TRY(cudaGetLastError());
    {
        KSKArgs a;
        a.ksk = ksk_dev; a.sk_lwe = d_s; a.sk_ring = d_sN; a.cdt = d_cdt; a.key_mask = key_mask; a.key_err = key_err; a.qKS = p.qKS;
        a.n = n; a.N = N; a.baseKS = p.baseKS; a.dKS = p.dKS; a.cdt_len = (int)cdt.size();
        const u64 rows = (u64)N * p.baseKS * p.dKS;
        ksk_gen_kernel<<<(unsigned)rows, 128, 0, st>>>(a);
        KG_TRY(cudaGetLastError());
    }
    {
        BKArgs a;
        a.bk = bk_dev; a.sk_lwe = d_s; a.skM = d_skM; a.WM = d_W; a.psiM = d_psi; a.cdt = d_cdt; a.key_mask = key_mask; a.key_err = key_err;
        a.baseG = p.baseG; a.n = n; a.N = N; a.method = p.method; a.baseR = p.baseR; a.digitsR = p.digitsR;
        a.q = (u32)p.q; a.cdt_len = (int)cdt.size(); a.M = M;
        u64 rows;
        if (p.method == TFHE_B200_METHOD_GINX) {
            a.d2 = 2 * (p.digitsG - p.numDigitsToThrow);
            a.throwd = p.numDigitsToThrow;
            rows = (u64)2 * n * a.d2;
        }
        else {
            a.d2 = 2 * p.digitsG;
            a.throwd = 0;
            rows = (u64)n * p.baseR * p.digitsR * a.d2;
        }
        bk_gen_kernel<<<(unsigned)rows, 256, N * 8, st>>>(a);
        KG_TRY(cudaGetLastError());
    }
    KG_TRY(cudaStreamSynchronize(st));
done:
    cudaFree(d_s); cudaFree(d_sN); cudaFree(d_W); cudaFree(d_psi); cudaFree(d_cdt); cudaFree(d_skM);
    if (st)
        cudaStreamDestroy(st);
    return rc;
}

}  // namespace tfhe_b200
