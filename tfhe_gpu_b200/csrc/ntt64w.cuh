// Shared device helpers of the "wide" 64-bit register-resident NTT kernels (br_cggi64w.cu, br_dm64w.cu): N = 2048 as
// 16 x 8 x 16 with 128 threads x 16 coefficients per polynomial -- region addressing, the three transform passes in both
// directions, the CTA shape.  See br_cggi64w.cu for the design notes.
#pragma once
#include "ntt64.cuh"

namespace tfhe_b200 {
namespace w64 {

constexpr int LOGN = 11, N = 2048, TPN = 128, CPT = 16;

// position -> physical u64 index inside a region: 16-byte chunks of every 16-position block XOR-ed with the block index
__device__ __forceinline__ u32 posw(u32 p) {
    return (p & ~14u) | ((((p >> 1) ^ (p >> 4)) & 7u) << 1);
}
__device__ __forceinline__ void group_sync128(int id) {
    asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory");
}
// 16 consecutive positions of block b16
__device__ __forceinline__ void load_C(u64 (&v)[CPT], const u64* reg, int b16) {
    const ulonglong2* p = reinterpret_cast<const ulonglong2*>(reg) + 8 * b16;
#pragma unroll
    for (int x = 0; x < 8; x++) {
        ulonglong2 w = p[x ^ (b16 & 7)];
        v[2 * x] = w.x;
        v[2 * x + 1] = w.y;
    }
}
__device__ __forceinline__ void store_C(const u64 (&v)[CPT], u64* reg, int b16) {
    ulonglong2* p = reinterpret_cast<ulonglong2*>(reg) + 8 * b16;
#pragma unroll
    for (int x = 0; x < 8; x++)
        p[x ^ (b16 & 7)] = make_ulonglong2(v[2 * x], v[2 * x + 1]);
}
// layout B of 128-block blk: thread u of the octet holds positions blk*128 + 16 r + 2 u + c as v[2 r + c]
__device__ __forceinline__ void load_Bw(u64 (&v)[CPT], const u64* reg, int blk, int u) {
    const ulonglong2* p = reinterpret_cast<const ulonglong2*>(reg) + 64 * blk;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        ulonglong2 w = p[8 * r + (u ^ r)];
        v[2 * r] = w.x;
        v[2 * r + 1] = w.y;
    }
}
__device__ __forceinline__ void store_Bw(const u64 (&v)[CPT], u64* reg, int blk, int u) {
    ulonglong2* p = reinterpret_cast<ulonglong2*>(reg) + 64 * blk;
#pragma unroll
    for (int r = 0; r < 8; r++)
        p[8 * r + (u ^ r)] = make_ulonglong2(v[2 * r], v[2 * r + 1]);
}

// four Cooley-Tukey stages on v[16] (index bits 3..0); twiddle (cnt - 1 + x) of the table, x = index >> (s + 1)
__device__ __forceinline__ void fwd_pass4(u64 (&v)[CPT], const ulonglong2* __restrict__ tab, int stride, int lane_off,
                                          u64 nQ, u64 QO, u64 Z) {
#pragma unroll
    for (int s = 3; s >= 0; s--) {
        const int cnt = 8 >> s, off = cnt - 1;
        ulonglong2 w[8];
#pragma unroll
        for (int x = 0; x < cnt; x++)
            w[x] = tab[(off + x) * stride + lane_off];
#pragma unroll
        for (int r = 0; r < CPT; r++) {
            if (r & (1 << s))
                continue;
            const int ti = r >> (s + 1);
            u64 t = shoup64(v[r + (1 << s)], w[ti].x, w[ti].y, nQ, Z);
            u64 x = v[r];
            v[r] = x + t + Z;
            v[r + (1 << s)] = x - t + QO;
        }
    }
}
__device__ __forceinline__ void inv_pass4(u64 (&v)[CPT], const ulonglong2* __restrict__ tab, int stride, int lane_off,
                                          bool mirror, u64 nQ, u64 QO, u64 Z) {
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const int cnt = 8 >> s, off = cnt - 1;
        ulonglong2 w[8];
#pragma unroll
        for (int x = 0; x < cnt; x++)
            w[x] = tab[(off + (mirror ? cnt - 1 - x : x)) * stride + lane_off];
#pragma unroll
        for (int r = 0; r < CPT; r++) {
            if (r & (1 << s))
                continue;
            const int ti = r >> (s + 1);
            u64 U = v[r], V = v[r + (1 << s)];
            v[r] = csub(U + V + Z, QO);
            v[r + (1 << s)] = shoup64(V - U + QO, w[ti].x, w[ti].y, nQ, Z);
        }
    }
}
// three stages on layout B (index 2 r + c, r bits 2..0), twiddles (cnt - 1 + x) of the block's row
__device__ __forceinline__ void fwd_pass3(u64 (&v)[CPT], const ulonglong2* __restrict__ row, u64 nQ, u64 QO, u64 Z) {
#pragma unroll
    for (int s = 2; s >= 0; s--) {
        const int cnt = 4 >> s, off = cnt - 1;
        ulonglong2 w[4];
#pragma unroll
        for (int x = 0; x < cnt; x++)
            w[x] = row[off + x];
#pragma unroll
        for (int q = 0; q < CPT; q++) {
            const int r = q >> 1;
            if (r & (1 << s))
                continue;
            const int ti = r >> (s + 1), q2 = q + (2 << s);
            u64 t = shoup64(v[q2], w[ti].x, w[ti].y, nQ, Z);
            u64 x = v[q];
            v[q] = x + t + Z;
            v[q2] = x - t + QO;
        }
    }
}
__device__ __forceinline__ void inv_pass3(u64 (&v)[CPT], const ulonglong2* __restrict__ row, u64 nQ, u64 QO, u64 Z) {
#pragma unroll
    for (int s = 0; s < 3; s++) {
        const int cnt = 4 >> s, off = cnt - 1;
        ulonglong2 w[4];
#pragma unroll
        for (int x = 0; x < cnt; x++)
            w[x] = row[off + (cnt - 1 - x)];   // mirrored block
#pragma unroll
        for (int q = 0; q < CPT; q++) {
            const int r = q >> 1;
            if (r & (1 << s))
                continue;
            const int ti = r >> (s + 1), q2 = q + (2 << s);
            u64 U = v[q], V = v[q2];
            v[q] = csub(U + V + Z, QO);
            v[q2] = shoup64(V - U + QO, w[ti].x, w[ti].y, nQ, Z);
        }
    }
}

template <int DK, int G>
struct KW {
    static constexpr int D = 2 * DK;
    static constexpr int NT = G * 2 * TPN;
    static constexpr int WB = N / 32;
    static constexpr size_t smem = (size_t)G * D * N * 8 + (size_t)15 * TPN * 16 + 16 * 8 * 16 + 2 * 15 * 16 + 64 +
                                   2 * (size_t)G * 2 * (WB + 4) * 4;
};

}  // namespace w64
}  // namespace tfhe_b200
