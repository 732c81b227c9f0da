// Shared device helpers of the 32-bit register-resident NTT kernels (br_cggi32.cu, br_dm32.cu): lazy Shoup
// butterflies, the two-pass radix-32 x radix-N/32 transform, region addressing.  See br_cggi32.cu for the design.
#pragma once
#include "engine.cuh"

namespace tfhe_b200 {

__device__ __forceinline__ u32 mulhi_w(u32 a, u32 b) {
    // high half through IMAD.WIDE (same issue cost as IMAD.HI on B200, measured)
    u32 hi, lo;
    asm("{ .reg .u64 p; mul.wide.u32 p, %2, %3; mov.b64 {%1, %0}, p; }" : "=r"(hi), "=r"(lo) : "r"(a), "r"(b));
    (void)lo;
    return hi;
}
__device__ __forceinline__ u32 shoup_mul(u32 y, u32 w, u32 wp, u32 Q) {
    // y*w mod Q up to one extra Q: result in [0, 2Q) for any 32-bit y.
    // The quotient estimate is the high half of an IMAD.WIDE; on B200 IMAD.WIDE and IMAD.HI both issue at half the
    // IMAD rate (profiles/r01_imad_peak.json), so a butterfly costs 4 fma-heavy issue slots: 2 + 1 + 1.
    u32 q, lo;
    asm("{ .reg .u64 p; mul.wide.u32 p, %2, %3; mov.b64 {%1, %0}, p; }" : "=r"(q), "=r"(lo) : "r"(y), "r"(wp));
    (void)lo;
    return y * w - q * Q;
}
__device__ __forceinline__ u32 cond_sub(u32 x, u32 m) {
    // x < 2m  ->  x mod m
    return min(x, x - m);
}
__device__ __forceinline__ u32 pos_of(u32 idx) {
    return idx + ((idx >> 5) << 2);
}

template <int LOGN, int DK, int G>
struct KCfg {
    static constexpr int N = 1 << LOGN;
    static constexpr int TPN = N / 32;            // threads per NTT
    static constexpr bool XS = LOGN == 11;        // N = 2048: one cross-lane stage (stride 32) between the two passes
    static constexpr int PB = XS ? 5 : LOGN - 5;  // in-thread pass-B stages
    static constexpr int NTW = XS ? 32 : 32 - (32 >> PB);   // per-thread twiddles (N = 2048: slot 31 = the cross-stage twiddle)
    static constexpr int D = 2 * DK;              // digit polynomials
    static constexpr int RS = N + N / 8 + (LOGN == 9 ? 16 : 0);  // padded region stride (words)
    static constexpr int NT = G * 2 * TPN;        // threads per CTA
    static constexpr size_t smem_bytes(int n) {
        return (size_t)G * D * RS * 4 + (size_t)2 * N * 4 + (size_t)G * ((n + 1) / 2 * 2) * 2 + 64;
    }
    // TMA variant of br_cggi32: the key ring starts behind the rotation exponents, 128-byte aligned
    static constexpr size_t ring_offset(int n) {
        return ((size_t)G * D * RS * 4 + (size_t)2 * N * 4 + (size_t)G * (size_t)n * 2 + 127) & ~(size_t)127;
    }
};

// ---- register-resident NTT passes -----------------------------------------------------------------------------
// forward pass A: Cooley-Tukey stages with stride TPN*2^s, s = 4..0 (uniform twiddles from the parameter bank)
// `add0`: a constant still to be added to v[0..15] (the lower operands of the first stage): it rides in the third operand
// of that stage's additions (digit polynomials: the offset Q - B/2 of the first sixteen coefficients, br_cggi32.cu)
template <typename A>
__device__ __forceinline__ void fwd_passA(u32 (&v)[32], const A& args, u32 Q, u32 Q2, u32 add0) {
    const u32 Z = args.zero;
#pragma unroll
    for (int s = 4; s >= 0; s--) {
#pragma unroll
        for (int r = 0; r < 32; r++) {
            if (r & (1 << s))
                continue;
            const int ti = (16 >> s) + (r >> (s + 1));
            u32 t = shoup_mul(v[r + (1 << s)], args.twA_f[ti][0], args.twA_f[ti][1], Q);
            u32 x = v[r];
            v[r] = x + t + (s == 4 ? add0 : Z);
            v[r + (1 << s)] = x - t + (s == 4 ? Q2 + add0 : Q2);
        }
    }
}
template <typename A>
__device__ __forceinline__ void fwd_passA(u32 (&v)[32], const A& args, u32 Q, u32 Q2) {
    fwd_passA(v, args, Q, Q2, args.zero);
}
// Mid-transform sweep for 28-bit moduli: the lazy forward transform lets values grow by 2Q per stage, (2 + 2 s) Q after
// s stages, and 22 Q must fit 32 bits -- true for the 27-bit primes only.  Bringing the values back below 2Q between the
// two passes (three min-subtract steps per coefficient on the ALU pipe, none on the multiplier pipe) bounds both passes by
// 12 Q, which admits every Q < 2^28 (MEDIUM, SIGNED_MOD_TEST: binfhecontext.cpp:140,155).
__device__ __forceinline__ void sweep_below_2q(u32 (&v)[32], u32 Q2) {
#pragma unroll
    for (int r = 0; r < 32; r++) {
        u32 x = v[r];                // < 12 Q
        x = cond_sub(x, 4 * Q2);     // < 8 Q
        x = cond_sub(x, 2 * Q2);     // < 4 Q
        v[r] = cond_sub(x, Q2);      // < 2 Q
    }
}
// ---- N = 2048 (64 threads x 32 coefficients per polynomial) ----------------------------------------------------------
// 2048 = 32 x 2 x 32: five in-thread stages on either side of ONE stage (stride 32) that pairs the 32-blocks of two
// neighbouring lanes.  Each lane computes half of its butterflies: the lane holding the lower block (U) takes indices
// 16..31, its neighbour (V) takes 0..15; 16 shuffles bring the operands together, 16 more send the results home.
// FWD: Cooley-Tukey with the pair's forward twiddle; the lower block lives in the EVEN lane.  Inverse (mirrored blocks,
// see inv_passB): Gentleman-Sande with (V - U) * w for the same forward twiddle; the lower block lives in the ODD lane.
template <bool FWD>
__device__ __forceinline__ void cross_stage(u32 (&v)[32], u32 w, u32 wp, u32 Q, u32 Q2, u32 Z, bool odd_lane) {
    const bool holdsU = FWD ? !odd_lane : odd_lane;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const u32 send = holdsU ? v[i] : v[16 + i];
        const u32 recv = __shfl_xor_sync(0xffffffffu, send, 1);
        const u32 U = holdsU ? v[16 + i] : recv, V = holdsU ? recv : v[i];
        u32 nu, nv;
        if (FWD) {
            const u32 t = shoup_mul(V, w, wp, Q);
            nu = U + t + Z;
            nv = U - t + Q2;
        }
        else {
            nu = cond_sub(U + V + Z, Q2);
            nv = shoup_mul(V - U + Q2, w, wp, Q);
        }
        const u32 back = __shfl_xor_sync(0xffffffffu, holdsU ? nv : nu, 1);
        if (holdsU) {
            v[16 + i] = nu;
            v[i] = back;
        }
        else {
            v[i] = nv;
            v[16 + i] = back;
        }
    }
}
// values < 8 Q -> < 2 Q (29-bit moduli: 8 Q still fits 32 bits, three lazy stages between sweeps)
__device__ __forceinline__ void sweep_8q(u32 (&v)[32], u32 Q2) {
#pragma unroll
    for (int r = 0; r < 32; r++)
        v[r] = cond_sub(cond_sub(v[r], 2 * Q2), Q2);
}
// forward pass A with a sweep after its third stage (29-bit moduli)
template <typename A>
__device__ __forceinline__ void fwd_passA_sw(u32 (&v)[32], const A& args, u32 Q, u32 Q2) {
    const u32 Z = args.zero;
#pragma unroll
    for (int s = 4; s >= 0; s--) {
#pragma unroll
        for (int r = 0; r < 32; r++) {
            if (r & (1 << s))
                continue;
            const int ti = (16 >> s) + (r >> (s + 1));
            u32 t = shoup_mul(v[r + (1 << s)], args.twA_f[ti][0], args.twA_f[ti][1], Q);
            u32 x = v[r];
            v[r] = x + t + Z;
            v[r + (1 << s)] = x - t + Q2;
        }
        if (s == 2)
            sweep_8q(v, Q2);
    }
}
// forward pass B (five stages) with a sweep after its second stage (the cross stage came first) and after the last
__device__ __forceinline__ void fwd_passB_sw(u32 (&v)[32], const u32 (&tw)[32], const u32 (&twp)[32], u32 Q, u32 Q2,
                                             u32 Z) {
#pragma unroll
    for (int s = 4; s >= 0; s--) {
        const int off = (32 >> (s + 1)) - 1;
#pragma unroll
        for (int r = 0; r < 32; r++) {
            if (r & (1 << s))
                continue;
            const int ti = off + (r >> (s + 1));
            u32 t = shoup_mul(v[r + (1 << s)], tw[ti], twp[ti], Q);
            u32 x = v[r];
            v[r] = x + t + Z;
            v[r + (1 << s)] = x - t + Q2;
        }
        if (s == 3 || s == 0)
            sweep_8q(v, Q2);
    }
}
// the threads of one polynomial: a warp (or part of one) up to N = 1024, two warps (named barrier) for N = 2048
template <int TPN>
__device__ __forceinline__ void poly_sync(int bar_id) {
    if (TPN <= 32)
        __syncwarp();
    else
        asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(TPN) : "memory");
}

// forward pass B: strides 2^s, s = PB-1..0, per-thread twiddles
template <int PB>
__device__ __forceinline__ void fwd_passB(u32 (&v)[32], const u32 (&tw)[32], const u32 (&twp)[32], u32 Q, u32 Q2,
                                          u32 Z) {
#pragma unroll
    for (int s = PB - 1; s >= 0; s--) {
        const int off = (32 >> (s + 1)) - (32 >> PB);
#pragma unroll
        for (int r = 0; r < 32; r++) {
            if (r & (1 << s))
                continue;
            const int ti = off + (r >> (s + 1));
            u32 t = shoup_mul(v[r + (1 << s)], tw[ti], twp[ti], Q);
            u32 x = v[r];
            v[r] = x + t + Z;
            v[r + (1 << s)] = x - t + Q2;
        }
    }
}
// inverse pass B' on the MIRRORED block (virtual thread TPN-1-T): Gentleman-Sande stages 2^s, s = 0..PB-1, with
// (U - V) * psi^-x == (V - U) * psi^{mirror}; all values kept below 2Q
// CANON: the inputs are canonical (< Q), so the sums of the first stage are below 2Q without a correction
template <int PB, bool CANON = false>
__device__ __forceinline__ void inv_passB(u32 (&v)[32], const u32 (&tw)[32], const u32 (&twp)[32], u32 Q, u32 Q2,
                                          u32 Z) {
#pragma unroll
    for (int s = 0; s < PB; s++) {
        const int off = (32 >> (s + 1)) - (32 >> PB);
        const int cnt = 32 >> (s + 1);
#pragma unroll
        for (int r = 0; r < 32; r++) {
            if (r & (1 << s))
                continue;
            const int ti = off + (cnt - 1 - (r >> (s + 1)));
            u32 U = v[r], V = v[r + (1 << s)];
            v[r] = (CANON && s == 0) ? U + V + Z : cond_sub(U + V + Z, Q2);
            v[r + (1 << s)] = shoup_mul(V - U + Q2, tw[ti], twp[ti], Q);
        }
    }
}
// inverse pass A': strides TPN*2^s, s = 0..4, uniform inverse twiddles
template <typename A>
__device__ __forceinline__ void inv_passA(u32 (&v)[32], const A& args, u32 Q, u32 Q2) {
    const u32 Z = args.zero;
#pragma unroll
    for (int s = 0; s < 5; s++) {
#pragma unroll
        for (int r = 0; r < 32; r++) {
            if (r & (1 << s))
                continue;
            const int ti = (16 >> s) + (r >> (s + 1));
            u32 U = v[r], V = v[r + (1 << s)];
            v[r] = cond_sub(U + V + Z, Q2);
            v[r + (1 << s)] = shoup_mul(U - V + Q2, args.twA_i[ti][0], args.twA_i[ti][1], Q);
        }
    }
}


}  // namespace tfhe_b200
