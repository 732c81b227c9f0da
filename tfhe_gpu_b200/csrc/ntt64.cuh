// Shared device helpers of the 64-bit register-resident NTT kernels (br_cggi64.cu, br_cggi64w.cu): explicit 32-bit
// multiply pieces, the approximate-quotient Shoup multiplication, 128-bit accumulation, the 27-bit limb pointwise
// multiply-accumulate.  See br_cggi64.cu for the design notes.
#pragma once
#include "engine.cuh"

namespace tfhe_b200 {

// runtime zero (never written): a third addend that keeps ptxas from turning 32-bit adds into IMAD.IADD on the
// multiplier pipe (7.8 % of its issue slots in the ncu profile of br_cggi64w before this)
static __constant__ u32 kZero32;

// explicit 32 x 32 (+ 64) -> 64: written as C, NVVM widens the operands to 64 bits and ptxas re-derives IMAD.WIDE with
// leftover adds of zero high halves
__device__ __forceinline__ u64 mulwide(u32 a, u32 b) {
    u64 r;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ u64 madwide(u32 a, u32 b, u64 c) {
    u64 r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
    return r;
}
__device__ __forceinline__ void unpack(u64 x, u32& lo, u32& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(x));
}
__device__ __forceinline__ u64 pack(u32 lo, u32 hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
// Shoup multiplication y * w mod Q with an APPROXIMATE quotient, in explicit 32-bit pieces (14 multiplier-pipe issue
// slots instead of the 16 + carry adds of __umul64hi + two 64-bit low products):
//   q' = y1 p1 + hi32(y1 p0) + hi32(y0 p1)  in [q - 2, q],  q = floor(y wp / 2^64), wp = floor(w 2^64 / Q)
//   t  = y w - q' Q (mod 2^64) = lo64(y0 w0 + q0 nQ0) + 2^32 (y0 w1 + y1 w0 + q0 nQ1 + q1 nQ0),   nQ = 2^64 - Q
// For y < 2^60: t < (1 + 2^-4) Q + 2 Q < 3.07 Q, so lazy values use an offset of 4Q (QO) where the exact form used 2Q.
// Z is a runtime 64-bit 0: it turns two-input adds into three-input IADD3 / IADD3.X (ALU pipe) where ptxas would
// otherwise pick IMAD.IADD / IMAD.X / IMAD.MOV on the multiplier pipe, which is the one this kernel saturates.
__device__ __forceinline__ u64 shoup64(u64 y, u64 w, u64 wp, u64 nQ, u64 Z) {
    u32 y0, y1, w0, w1, p0, p1, nq0, nq1, a0, a1, b0, b1, q0, q1, l0, l1, z0, z1;
    unpack(y, y0, y1);
    unpack(w, w0, w1);
    unpack(wp, p0, p1);
    unpack(nQ, nq0, nq1);
    unpack(Z, z0, z1);
    unpack(mulwide(y1, p0), a0, a1);
    unpack(mulwide(y0, p1), b0, b1);
    // the runtime-zero addend keeps ptxas from folding one of the 32-bit terms into the multiply-add as a 64-bit
    // register pair (two IMAD.MOVs); the two terms then go through one three-input IADD3 / IADD3.X pair
    const u64 q = madwide(y1, p1, Z) + (u64)a1 + (u64)b1;
    unpack(q, q0, q1);
    const u32 hy = y0 * w1 + y1 * w0;
    unpack(madwide(q0, nq0, mulwide(y0, w0)), l0, l1);
    const u32 h = q1 * nq0 + (q0 * nq1 + hy);
    return pack(l0, l1 + h + kZero32);
}
__device__ __forceinline__ u64 csub(u64 x, u64 m) {
    return x >= m ? x - m : x;
}
__device__ __forceinline__ void group_sync(int id) {
    asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

struct A128 {
    u64 lo, hi;
    __device__ __forceinline__ void mac(u64 x, u64 b) {
        u64 pl = x * b, ph = __umul64hi(x, b);
        lo += pl;
        hi += ph + (lo < pl);
    }
};
__device__ __forceinline__ u64 redc128(const A128& X, u64 Q, u64 qinv) {
    u64 m = X.lo * qinv;
    u64 t = __umul64hi(m, Q);
    u64 r = X.hi - t;
    return X.hi < t ? r + Q : r;
}

// 64 x 64 -> 128-bit multiply-accumulate in 27-bit limbs (Karatsuba): x = x0 + 2^27 x1 (x < 29 Q < 2^59: x1 < 2^32),
// b = b0 + 2^27 b1 (b < Q < 2^55: b1 < 2^28).  Three IMAD.WIDE per term and NO carry handling: with at most 8 terms the
// column sums s0 = sum x0 b0, s2 = sum x1 b1, kk = sum (x0 + x1)(b0 + b1) stay below 2^64 (checked on the host:
// cggi64_supported).  Key words arrive pre-split as b0 | b1 << 32 (bk_relayout_cggi64_kernel).
constexpr u32 M27 = (1u << 27) - 1;
struct Limb {
    u32 l0, l1, ls;
    __device__ __forceinline__ explicit Limb(u64 x) {
        u32 lo, hi;
        unpack(x, lo, hi);
        l0 = lo & M27;
        l1 = __funnelshift_r(lo, hi, 27);
        ls = l0 + l1 + kZero32;
    }
};
struct L3 {
    u64 s0, s2, kk;
    // first term (no zero-initialised accumulators)
    __device__ __forceinline__ L3(const Limb& x, u64 bpacked) {
        u32 b0, b1;
        unpack(bpacked, b0, b1);
        s0 = mulwide(x.l0, b0);
        s2 = mulwide(x.l1, b1);
        kk = mulwide(x.ls, b0 + b1 + kZero32);
    }
    __device__ __forceinline__ L3(const Limb& x, const Limb& b) {
        s0 = mulwide(x.l0, b.l0);
        s2 = mulwide(x.l1, b.l1);
        kk = mulwide(x.ls, b.ls);
    }
    __device__ __forceinline__ void mac(const Limb& x, u64 bpacked) {
        u32 b0, b1;
        unpack(bpacked, b0, b1);
        s0 = madwide(x.l0, b0, s0);
        s2 = madwide(x.l1, b1, s2);
        kk = madwide(x.ls, b0 + b1 + kZero32, kk);
    }
    __device__ __forceinline__ void mac(const Limb& x, const Limb& b) {
        s0 = madwide(x.l0, b.l0, s0);
        s2 = madwide(x.l1, b.l1, s2);
        kk = madwide(x.ls, b.ls, kk);
    }
    // s0 + 2^27 (kk - s0 - s2) + 2^54 s2 as a 128-bit value
    __device__ __forceinline__ A128 value() const {
        const u64 s1 = kk - s0 - s2;
        A128 r;
        r.lo = s0 + (s1 << 27);
        r.hi = (s1 >> 37) + (r.lo < s0);
        const u64 t = s2 << 54;
        r.lo += t;
        r.hi += (s2 >> 10) + (r.lo < t);
        return r;
    }
};


}  // namespace tfhe_b200
