// Modular arithmetic primitives for the blind-rotation kernels (sm_100a integer pipe).
//
//  * Montgomery multiplication for 32-bit (Q < 2^31) and 64-bit (Q < 2^62) prime moduli.  The bootstrapping
//    key, the NTT twiddles and the monomial table are stored in Montgomery form, so mont_mul(x, cM) = x*c mod Q
//    for a plain x -- one reduction per product, no conversion of the data path.
//  * Shoup (precomputed-quotient) multiplication for the 32-bit NTT butterflies: 1 IMAD.HI + 2 IMAD.
//
// The arithmetic is exact mod Q, which is all that bit-exactness with OpenFHE's NativeInteger path needs
// (SURVEY.md section 8 a': any exact negacyclic convolution reproduces the oracle).
#pragma once
#include <cstdint>

namespace tfhe_b200 {

typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;

template <typename T>
struct ModCtx;

template <>
struct ModCtx<u32> {
    u32 Q, qinv;  // qinv = Q^-1 mod 2^32
    u32 oneM;     // R mod Q  (Montgomery form of 1)
    u32 r2;       // R^2 mod Q
    __host__ __device__ static inline u32 mulhi(u32 a, u32 b) {
#ifdef __CUDA_ARCH__
        return __umulhi(a, b);
#else
        return (u32)(((u64)a * b) >> 32);
#endif
    }
    // a*b*R^-1 mod Q, requires a*b < Q*2^32; result in [0,Q)
    __host__ __device__ inline u32 mont_mul(u32 a, u32 b) const {
        u32 lo = a * b, hi = mulhi(a, b);
        u32 m = lo * qinv;
        u32 t = mulhi(m, Q);
        u32 r = hi - t;
        return hi < t ? r + Q : r;
    }
    // REDC of a 64-bit value x < Q*2^32; result in [0,Q)
    __host__ __device__ inline u32 redc(u64 x) const {
        u32 lo = (u32)x, hi = (u32)(x >> 32);
        u32 m = lo * qinv;
        u32 t = mulhi(m, Q);
        u32 r = hi - t;
        return hi < t ? r + Q : r;
    }
    __host__ __device__ inline u32 add(u32 a, u32 b) const {
        u32 r = a + b;
        return r >= Q ? r - Q : r;
    }
    __host__ __device__ inline u32 sub(u32 a, u32 b) const {
        return a >= b ? a - b : a + Q - b;
    }
};

template <>
struct ModCtx<u64> {
    u64 Q, qinv;  // qinv = Q^-1 mod 2^64
    u64 oneM;     // R mod Q
    u64 r2;       // R^2 mod Q
    __host__ __device__ static inline u64 mulhi(u64 a, u64 b) {
#ifdef __CUDA_ARCH__
        return __umul64hi(a, b);
#else
        return (u64)(((unsigned __int128)a * b) >> 64);
#endif
    }
    __host__ __device__ inline u64 mont_mul(u64 a, u64 b) const {
        u64 lo = a * b, hi = mulhi(a, b);
        u64 m = lo * qinv;
        u64 t = mulhi(m, Q);
        u64 r = hi - t;
        return hi < t ? r + Q : r;
    }
    __host__ __device__ inline u64 add(u64 a, u64 b) const {
        u64 r = a + b;
        return r >= Q ? r - Q : r;
    }
    __host__ __device__ inline u64 sub(u64 a, u64 b) const {
        return a >= b ? a - b : a + Q - b;
    }
};

// ---- host-side helpers to build the contexts and tables -------------------------------------------------
inline u64 h_mulmod(u64 a, u64 b, u64 Q) {
    return (u64)(((unsigned __int128)a * b) % Q);
}
inline u64 h_powmod(u64 a, u64 e, u64 Q) {
    u64 r = 1 % Q;
    a %= Q;
    while (e) {
        if (e & 1)
            r = h_mulmod(r, a, Q);
        a = h_mulmod(a, a, Q);
        e >>= 1;
    }
    return r;
}
template <typename T>
inline ModCtx<T> make_modctx(u64 Q) {
    ModCtx<T> m;
    m.Q = (T)Q;
    // Newton iteration for Q^-1 mod 2^w
    T inv = (T)Q;
    for (int i = 0; i < 6; i++)
        inv *= (T)2 - (T)Q * inv;
    m.qinv = inv;
    const int w = sizeof(T) * 8;
    u64 R = (w == 64) ? (u64)((((unsigned __int128)1) << 64) % Q) : (u64)((1ULL << 32) % Q);
    m.oneM = (T)R;
    m.r2 = (T)h_mulmod(R, R, Q);
    return m;
}
// Montgomery form of x: x*R mod Q
template <typename T>
inline T to_mont(u64 x, const ModCtx<T>& m) {
    return (T)h_mulmod(x % (u64)m.Q, (u64)m.oneM, (u64)m.Q);
}

// Exact restatement of LWEEncryptionScheme::RoundqQ (lwe-pke.cpp:41-46): three correctly rounded IEEE double
// operations (mul, div, add), floor, conversion, mod q.  Must NOT be replaced by exact integer rounding: for
// Q ~ 2^54, q = 2^35 the intermediate product exceeds 2^53 and the double rounding is part of the result.
__device__ __forceinline__ u64 round_qQ(u64 v, u64 q, double dq, double dQ) {
    double x = __dadd_rn(0.5, __ddiv_rn(__dmul_rn(__ull2double_rn(v), dq), dQ));
    return ((u64)floor(x)) % q;
}

}  // namespace tfhe_b200
