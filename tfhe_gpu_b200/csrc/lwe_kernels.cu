// LWE-side kernels: fused ModSwitch -> KeySwitch -> ModSwitch, LWE affine glue between bootstraps, exact integer
// CiphertextMulMatrix, and the setup-time key re-encoding kernels.
//
// References (semantics): LWEEncryptionScheme::RoundqQ/ModSwitch/KeySwitch (lwe-pke.cpp:41-46,204-215,299-321),
// MKMSwitchKernel (bootstrapping.cu:73-118), EvalAddEq/EvalSubEq/EvalSubEq2/EvalAdd(Sub)ConstEq
// (lwe-pke.cpp:172-200), LWECiphertextImpl::SetModulus (lwe-ciphertext.h:120-124),
// CiphertextMulMatrix_CUDA + applyFmod (lwe-operation.cu:42-141).
#include <cstdlib>

#include "engine.cuh"

namespace tfhe_b200 {

// ---------------------------------------------------------------------------------------------------------
// MS -> KS -> MS.  One CTA per ciphertext.  The key-switching table is re-encoded at setup to the narrowest
// word that holds qKS (u16 for 2^14, u32 for the 27-bit prime, u64 for 2^35) with rows padded to 16 bytes, so
// every thread streams the gathered rows with 128-bit loads.  Threads are arranged as RG row groups x TC column
// vectors; partial column sums are kept in registers (sums of at most N*dKS < 2^15 values below 2^35 fit a u64
// without reduction), combined through shared memory, and the subtraction mod qKS happens once per column.
// ---------------------------------------------------------------------------------------------------------
template <typename TK, int VW>  // VW = entries per 16-byte vector
struct alignas(16) KVec {
    TK v[VW];
};

template <typename TK, int VW, int S, bool SPLIT>
__global__ void __launch_bounds__(512) mkmswitch_kernel(KSArgs A, int TC, int RG) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const u32 N = A.N, n = A.n, dKS = A.dKS;
    unsigned long long* colsum = reinterpret_cast<unsigned long long*>(smem_raw);  // [row_stride]
    unsigned short* dig = reinterpret_cast<unsigned short*>(colsum + A.row_stride);  // [rows]
    __shared__ u64 b_ms;

    const int ct = blockIdx.x;
    const u64* ext = A.ext + (size_t)ct * (N + 1);
    const double dQ = (double)A.Q, dqKS = (double)A.qKS;
    // this CTA's share of the mask entries (all of them unless the launch is split)
    // (SPLIT = false is the plain one-CTA-per-ciphertext kernel: the range is everything and folds away)
    const u32 ipc = SPLIT ? (N + gridDim.y - 1) / gridDim.y : N, i_lo = SPLIT ? blockIdx.y * ipc : 0;
    const u32 i_hi = SPLIT ? min(N, i_lo + ipc) : N;
    const u32 r_lo = i_lo * dKS, r_hi = i_hi * dKS;

    // ModSwitch Q -> qKS and base-baseKS digit extraction of the N mask entries
    for (u32 i = i_lo + threadIdx.x; i <= N; i += blockDim.x) {
        if (SPLIT && i >= i_hi && i != N)
            continue;
        u64 v = round_qQ(ext[i], A.qKS, dqKS, dQ);
        if (i == N)
            b_ms = v;
        else {
            for (u32 j = 0; j < dKS; j++) {
                dig[i * dKS + j] = (unsigned short)(v % A.baseKS);
                v /= A.baseKS;
            }
        }
    }
    for (u32 k = threadIdx.x; k < A.row_stride; k += blockDim.x)
        colsum[k] = 0;
    __syncthreads();

    const int tc = threadIdx.x % TC, rg = threadIdx.x / TC;
    const u32 CV = A.row_stride / VW;
    const KVec<TK, VW>* tab = reinterpret_cast<const KVec<TK, VW>*>(A.ksk);
    if (rg < RG) {
        u64 acc[S][VW];
#pragma unroll
        for (int s = 0; s < S; s++)
#pragma unroll
            for (int v = 0; v < VW; v++)
                acc[s][v] = 0;
        // up to four gathered rows per trip: all their loads are issued before the first add (the gather is latency / L2
        // bound; one row per trip left only S loads in flight per thread)
        auto row_ptr = [&](u32 r) {
            u32 i = r / dKS, j = r - i * dKS;
            u32 a0 = dig[r];
            size_t row = ((size_t)i * A.baseKS + a0) * dKS + j;
            return tab + row * CV;
        };
        // rows per trip: 4 measured 20.1 ms per 2048 ciphertexts of the 54-bit table (1 row: 29.1 ms; 8 rows with a
        // 255-register budget: 38.8 ms, occupancy lost)
        constexpr int U = S <= 3 ? 4 : (S == 4 ? 2 : 1);
        const u32 rows = r_hi;
        u32 r = r_lo + rg;
        for (; r + (U - 1) * RG < rows; r += U * RG) {
            const KVec<TK, VW>* rp[U];
#pragma unroll
            for (int q = 0; q < U; q++)
                rp[q] = row_ptr(r + q * RG);
            KVec<TK, VW> x[U][S];
#pragma unroll
            for (int q = 0; q < U; q++)
#pragma unroll
                for (int s = 0; s < S; s++) {
                    u32 cv = tc + s * TC;
                    if (cv < CV)
                        x[q][s] = rp[q][cv];
                }
#pragma unroll
            for (int q = 0; q < U; q++)
#pragma unroll
                for (int s = 0; s < S; s++) {
                    u32 cv = tc + s * TC;
                    if (cv < CV) {
#pragma unroll
                        for (int v = 0; v < VW; v++)
                            acc[s][v] += (u64)x[q][s].v[v];
                    }
                }
        }
        for (; r < rows; r += RG) {
            const KVec<TK, VW>* rp = row_ptr(r);
#pragma unroll
            for (int s = 0; s < S; s++) {
                u32 cv = tc + s * TC;
                if (cv < CV) {
                    KVec<TK, VW> x = rp[cv];
#pragma unroll
                    for (int v = 0; v < VW; v++)
                        acc[s][v] += (u64)x.v[v];
                }
            }
        }
#pragma unroll
        for (int s = 0; s < S; s++) {
            u32 cv = tc + s * TC;
            if (cv < CV) {
#pragma unroll
                for (int v = 0; v < VW; v++)
                    atomicAdd(&colsum[cv * VW + v], (unsigned long long)acc[s][v]);
            }
        }
    }
    __syncthreads();
    if (SPLIT) {   // partial column sums -> global accumulator; mkmswitch_finish_kernel completes the switch
        unsigned long long* part = reinterpret_cast<unsigned long long*>(A.partial) + (size_t)ct * A.row_stride;
        for (u32 k = threadIdx.x; k <= n; k += blockDim.x)
            atomicAdd(part + k, colsum[k]);
        return;
    }

    // a_out = 0 - sum, b_out = b - sum (mod qKS), then ModSwitch qKS -> fmod
    u64* out = A.out + (size_t)ct * (n + 1);
    const double dfmod = (double)A.fmod;
    for (u32 k = threadIdx.x; k <= n; k += blockDim.x) {
        u64 sum = colsum[k] % A.qKS;
        u64 base = (k == n) ? b_ms : 0;
        u64 v = base >= sum ? base - sum : base + A.qKS - sum;
        out[k] = round_qQ(v, A.fmod, dfmod, dqKS);
    }
}

template <typename TK, int VW>
static cudaError_t launch_ks_t(const KSArgs& a, cudaStream_t s) {
    const u32 CV = a.row_stride / VW;
    int TC = CV < 256 ? (int)CV : 256;
    int S = (int)((CV + TC - 1) / TC);
    int RG = 256 / TC;
    if (RG < 1)
        RG = 1;
    int threads = TC * RG;
    threads = (threads + 31) / 32 * 32;
    size_t smem = (size_t)a.row_stride * 8 + (size_t)a.N * a.dKS * 2 + 16;
#define KS_LAUNCH(SS)                                                                                            \
    {                                                                                                            \
        auto kern = a.splits > 1 ? mkmswitch_kernel<TK, VW, SS, true> : mkmswitch_kernel<TK, VW, SS, false>;     \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
        if (e != cudaSuccess)                                                                                    \
            return e;                                                                                            \
        kern<<<dim3(a.batch, a.splits), threads, smem, s>>>(a, TC, RG);                                          \
    }
    if (S == 1)
        KS_LAUNCH(1)
    else if (S == 2)
        KS_LAUNCH(2)
    else if (S == 3)
        KS_LAUNCH(3)
    else if (S == 4)
        KS_LAUNCH(4)
    else if (S <= 8)
        KS_LAUNCH(8)
    else
        return cudaErrorInvalidValue;
#undef KS_LAUNCH
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// Packed variant for u16 tables with qKS <= 2^14 (STD128: qKS = 2^14, 520-entry rows).  The ncu capture of the
// generic kernel shows it is instruction-issue bound on unpack-and-add (sm__throughput 71 %, DRAM 35 %), so here
// four gathered rows are first added as PACKED pairs of u16 (4 * (2^14 - 1) < 2^16: the 16-bit lanes cannot carry
// into each other) and only the 4-row partial sums are unpacked into 32-bit column accumulators: 0.75 instructions
// per table entry instead of ~3.  Row offsets (not digits) are precomputed in shared memory.
// ---------------------------------------------------------------------------------------------------------
// Gather-accumulate shared by the plain and the split packed kernels: rows r_lo + rg, r_lo + rg + RG, ... < rows of the
// ciphertext's row list, four gathered rows in flight per trip, packed u16 pair sums unpacked into 32-bit column
// accumulators, combined in `colsum` (shared memory).  Force-inlined: the plain kernel passes r_lo = 0 and compiles to
// the same SASS as when this body was written out in it (checked with cuobjdump; its load scheduling is what makes it
// fast).
template <int S>
__device__ __forceinline__ void ks16_gather(const KSArgs& A, int TC, int RG, const u32* rowoff, u32* colsum, u32 r_lo,
                                            u32 rows) {
    const int tc = threadIdx.x % TC, rg = threadIdx.x / TC;
    const u32 CV = A.row_stride / 8;
    const uint4* tab = reinterpret_cast<const uint4*>(A.ksk);
    if (rg < RG) {
        u32 lo[S][4], hi[S][4];
#pragma unroll
        for (int s = 0; s < S; s++)
#pragma unroll
            for (int v = 0; v < 4; v++)
                lo[s][v] = hi[s][v] = 0;
        u32 r = r_lo + rg;
        for (; r + 3 * RG < rows; r += 4 * RG) {
            const u32 o0 = rowoff[r], o1 = rowoff[r + RG], o2 = rowoff[r + 2 * RG], o3 = rowoff[r + 3 * RG];
#pragma unroll
            for (int s = 0; s < S; s++) {
                const u32 cv = tc + s * TC;
                if (cv < CV) {
                    const uint4 a = __ldg(tab + (size_t)o0 * CV + cv), b = __ldg(tab + (size_t)o1 * CV + cv);
                    const uint4 c = __ldg(tab + (size_t)o2 * CV + cv), d = __ldg(tab + (size_t)o3 * CV + cv);
                    const u32 p[4] = {a.x + b.x + c.x + d.x, a.y + b.y + c.y + d.y, a.z + b.z + c.z + d.z,
                                      a.w + b.w + c.w + d.w};
#pragma unroll
                    for (int v = 0; v < 4; v++) {
                        lo[s][v] += p[v] & 0xffffu;
                        hi[s][v] += p[v] >> 16;
                    }
                }
            }
        }
        for (; r < rows; r += RG) {
            const u32 o0 = rowoff[r];
#pragma unroll
            for (int s = 0; s < S; s++) {
                const u32 cv = tc + s * TC;
                if (cv < CV) {
                    const uint4 a = __ldg(tab + (size_t)o0 * CV + cv);
                    const u32 p[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                    for (int v = 0; v < 4; v++) {
                        lo[s][v] += p[v] & 0xffffu;
                        hi[s][v] += p[v] >> 16;
                    }
                }
            }
        }
#pragma unroll
        for (int s = 0; s < S; s++) {
            const u32 cv = tc + s * TC;
            if (cv < CV) {
#pragma unroll
                for (int v = 0; v < 4; v++) {
                    atomicAdd(&colsum[cv * 8 + 2 * v], lo[s][v]);
                    atomicAdd(&colsum[cv * 8 + 2 * v + 1], hi[s][v]);
                }
            }
        }
    }
}

template <int S>
__global__ void __launch_bounds__(288) mkmswitch_packed16_kernel(KSArgs A, int TC, int RG) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const u32 N = A.N, n = A.n, dKS = A.dKS;
    const u32 rows = N * dKS;
    u32* colsum = reinterpret_cast<u32*>(smem_raw);           // [row_stride]
    u32* rowoff = colsum + A.row_stride;                      // [rows] table row index of every gathered row
    __shared__ u64 b_ms;

    const int ct = blockIdx.x;
    const u64* ext = A.ext + (size_t)ct * (N + 1);
    const double dQ = (double)A.Q, dqKS = (double)A.qKS;
    for (u32 i = threadIdx.x; i <= N; i += blockDim.x) {
        u64 v = round_qQ(ext[i], A.qKS, dqKS, dQ);
        if (i == N)
            b_ms = v;
        else {
            for (u32 j = 0; j < dKS; j++) {
                u32 a0 = (u32)(v % A.baseKS);
                v /= A.baseKS;
                rowoff[i * dKS + j] = (i * A.baseKS + a0) * dKS + j;
            }
        }
    }
    for (u32 k = threadIdx.x; k < A.row_stride; k += blockDim.x)
        colsum[k] = 0;
    __syncthreads();

    ks16_gather<S>(A, TC, RG, rowoff, colsum, 0, rows);
    __syncthreads();
    u64* out = A.out + (size_t)ct * (n + 1);
    const double dfmod = (double)A.fmod;
    for (u32 k = threadIdx.x; k <= n; k += blockDim.x) {
        u64 sum = (u64)colsum[k] % A.qKS;
        u64 base = (k == n) ? b_ms : 0;
        u64 v = base >= sum ? base - sum : base + A.qKS - sum;
        out[k] = round_qQ(v, A.fmod, dfmod, dqKS);
    }
}

// Split variant (small batches): gridDim.y CTAs share the rows of a ciphertext.  A separate __global__ function around
// the shared gather body: merging the two kernels behind one template flag changed the plain kernel's load scheduling
// (3.09 ms instead of 2.22 ms per 16384 ciphertexts); with the body force-inlined into two kernels the plain one keeps
// its SASS.
template <int S>
__global__ void __launch_bounds__(288) mkmswitch_packed16_split_kernel(KSArgs A, int TC, int RG) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const u32 N = A.N, n = A.n, dKS = A.dKS;
    u32* colsum = reinterpret_cast<u32*>(smem_raw);           // [row_stride]
    u32* rowoff = colsum + A.row_stride;                      // [rows] table row index of every gathered row
    __shared__ u64 b_ms;

    const int ct = blockIdx.x;
    const u64* ext = A.ext + (size_t)ct * (N + 1);
    const double dQ = (double)A.Q, dqKS = (double)A.qKS;
    const u32 ipc = (N + gridDim.y - 1) / gridDim.y, i_lo = blockIdx.y * ipc;
    const u32 i_hi = min(N, i_lo + ipc);
    const u32 r_lo = i_lo * dKS, rows = i_hi * dKS;
    for (u32 i = i_lo + threadIdx.x; i < i_hi; i += blockDim.x) {
        u64 v = round_qQ(ext[i], A.qKS, dqKS, dQ);
        for (u32 j = 0; j < dKS; j++) {
            u32 a0 = (u32)(v % A.baseKS);
            v /= A.baseKS;
            rowoff[i * dKS + j] = (i * A.baseKS + a0) * dKS + j;
        }
    }
    if (threadIdx.x == blockDim.x - 1)   // a thread with the fewest loop trips
        b_ms = round_qQ(ext[N], A.qKS, dqKS, dQ);
    for (u32 k = threadIdx.x; k < A.row_stride; k += blockDim.x)
        colsum[k] = 0;
    __syncthreads();

    ks16_gather<S>(A, TC, RG, rowoff, colsum, r_lo, rows);
    __syncthreads();
    {
        unsigned long long* part = reinterpret_cast<unsigned long long*>(A.partial) + (size_t)ct * A.row_stride;
        for (u32 k = threadIdx.x; k <= n; k += blockDim.x)
            atomicAdd(part + k, (unsigned long long)colsum[k]);
    }
}

static cudaError_t launch_ks_packed16(const KSArgs& a, cudaStream_t s) {
    const u32 CV = a.row_stride / 8;
    int TC = CV < 256 ? (int)CV : 256;
    int S = (int)((CV + TC - 1) / TC);
    int RG = 256 / TC;
    if (RG < 1)
        RG = 1;
    int threads = (TC * RG + 31) / 32 * 32;
    size_t smem = (size_t)a.row_stride * 4 + (size_t)a.N * a.dKS * 4 + 16;
    if (S != 1 || smem > 200 * 1024)
        return cudaErrorNotSupported;
    auto kern = a.splits > 1 ? mkmswitch_packed16_split_kernel<1> : mkmswitch_packed16_kernel<1>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess)
        return e;
    kern<<<dim3(a.batch, a.splits), threads, smem, s>>>(a, TC, RG);
    return cudaGetLastError();
}

// second half of a split key switch: a_out = 0 - sum, b_out = MS(b) - sum (mod qKS), then ModSwitch qKS -> fmod
__global__ void mkmswitch_finish_kernel(KSArgs A) {
    const u32 n = A.n;
    const double dQ = (double)A.Q, dqKS = (double)A.qKS, dfmod = (double)A.fmod;
    const size_t total = (size_t)A.batch * (n + 1);
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t ct = idx / (n + 1);
        const u32 k = (u32)(idx - ct * (n + 1));
        const u64 sum = A.partial[ct * A.row_stride + k] % A.qKS;
        const u64 base = (k == n) ? round_qQ(A.ext[ct * (A.N + 1) + A.N], A.qKS, dqKS, dQ) : 0;
        const u64 v = base >= sum ? base - sum : base + A.qKS - sum;
        A.out[idx] = round_qQ(v, A.fmod, dfmod, dqKS);
    }
}

size_t mkmswitch_partial_bytes(u32 row_stride, int max_batch) {
    return (size_t)max_batch * row_stride * 8;
}

static cudaError_t launch_mkmswitch_main(const KSArgs& a, cudaStream_t s);

cudaError_t launch_mkmswitch(const KSArgs& a0, cudaStream_t s) {
    if (a0.batch <= 0)
        return cudaSuccess;
    KSArgs a = a0;
    // One CTA per ciphertext walks N*dKS dependent gather trips; a batch below two ciphertexts per SM is split so that
    // about four CTAs per SM share the rows of each ciphertext (the column sums meet in a global accumulator).
    a.splits = 1;
    if (a.partial && a.sm_count > 0 && a.batch < 2 * a.sm_count && !getenv("TFHE_B200_NO_KSSPLIT")) {
        int sp = (4 * a.sm_count + a.batch - 1) / a.batch;   // about four CTAs per SM in total
        if (sp > 16)
            sp = 16;
        if (sp > (int)a.N)
            sp = (int)a.N;
        a.splits = sp < 2 ? 1 : sp;
    }
    if (a.splits == 1) {
        a.partial = nullptr;
        return launch_mkmswitch_main(a, s);
    }
    cudaError_t e = cudaMemsetAsync(a.partial, 0, mkmswitch_partial_bytes(a.row_stride, a.batch), s);
    if (e != cudaSuccess)
        return e;
    e = launch_mkmswitch_main(a, s);
    if (e != cudaSuccess)
        return e;
    const size_t total = (size_t)a.batch * (a.n + 1);
    mkmswitch_finish_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(a);
    return cudaGetLastError();
}

static cudaError_t launch_mkmswitch_main(const KSArgs& a, cudaStream_t s) {
    // packed path: u16 entries, 4 rows cannot overflow a 16-bit lane, 32-bit column sums cannot overflow
    if (a.ksk_bytes == 2 && a.qKS <= (1u << 14) && (u64)a.N * a.dKS * a.qKS < (1ULL << 32) && !getenv("TFHE_B200_NO_KSPACK")) {
        cudaError_t e = launch_ks_packed16(a, s);
        if (e != cudaErrorNotSupported)
            return e;
    }
    switch (a.ksk_bytes) {
        case 2: return launch_ks_t<unsigned short, 8>(a, s);
        case 4: return launch_ks_t<u32, 4>(a, s);
        case 8: return launch_ks_t<u64, 2>(a, s);
    }
    return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------------------------
// LWE glue
// ---------------------------------------------------------------------------------------------------------
__global__ void lwe_affine_kernel(u64* out, const u64* x, const u64* y, int sx, int sy, int dbl, u64 cb, u64 m, u64 m2,
                                  size_t total, u32 words) {
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        u64 v = x[idx] % m;
        if (sx < 0)
            v = v ? m - v : 0;
        if (sy != 0) {
            u64 w = y[idx] % m;
            if (sy < 0)
                w = w ? m - w : 0;
            v += w;
            if (v >= m)
                v -= m;
        }
        if (dbl) {
            v += v;
            if (v >= m)
                v -= m;
        }
        if ((idx % words) == words - 1) {
            v += cb % m;
            if (v >= m)
                v -= m;
        }
        if (m2)
            v %= m2;
        out[idx] = v;
    }
}

cudaError_t launch_lwe_affine(u64* out, const u64* x, const u64* y, int sx, int sy, int dbl, u64 cb, u64 m, u64 m2,
                              int batch, u32 words, cudaStream_t s) {
    size_t total = (size_t)batch * words;
    if (!total)
        return cudaSuccess;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16)
        blocks = 148 * 16;
    lwe_affine_kernel<<<blocks, 256, 0, s>>>(out, x, y, sx, sy, dbl, cb, m, m2, total, words);
    return cudaGetLastError();
}

__global__ void mod_switch_kernel(u64* out, const u64* in, u64 from_mod, u64 to_mod, size_t count) {
    const double dq = (double)to_mod, dQ = (double)from_mod;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < count;
         idx += (size_t)gridDim.x * blockDim.x)
        out[idx] = round_qQ(in[idx], to_mod, dq, dQ);
}

cudaError_t launch_mod_switch(u64* out, const u64* in, u64 from_mod, u64 to_mod, size_t count, cudaStream_t s) {
    if (!count)
        return cudaSuccess;
    int blocks = (int)((count + 255) / 256);
    if (blocks > 148 * 16)
        blocks = 148 * 16;
    mod_switch_kernel<<<blocks, 256, 0, s>>>(out, in, from_mod, to_mod, count);
    return cudaGetLastError();
}

// Closed-form step tables of the functional operators, built on the device so that the per-GPU bodies of a sharded
// call never touch host memory between launches (the reference evaluates these lambdas per coefficient on the host:
// binfhe-base-scheme.cpp:716-720 f0, :934-952 f1/f2, :1004-1010 f3).
//   STEP_HALF : x < len/2 ? a : b                                   (f0 of EvalFunc, f1 of EvalFloor, f3 of EvalSign)
//   STEP_FLOOR2: x < len/4 ? a - len/2 - x : (x < 3 len/4 ? x : a + len/2 - x)        (f2 of EvalFloor, a = mod)
__global__ void step_table_kernel(u64* tab, int kind, u64 len, u64 a, u64 b) {
    for (u64 x = (u64)blockIdx.x * blockDim.x + threadIdx.x; x < len; x += (u64)gridDim.x * blockDim.x) {
        u64 v;
        if (kind == STEP_HALF)
            v = x < len / 2 ? a : b;
        else
            v = x < len / 4 ? a - len / 2 - x : (x < 3 * len / 4 ? x : a + len / 2 - x);
        tab[x] = v;
    }
}

cudaError_t launch_step_table(u64* tab, int kind, u64 len, u64 a, u64 b, cudaStream_t s) {
    if (!len)
        return cudaSuccess;
    int blocks = (int)((len + 255) / 256);
    if (blocks > 148 * 4)
        blocks = 148 * 4;
    step_table_kernel<<<blocks, 256, 0, s>>>(tab, kind, len, a, b);
    return cudaGetLastError();
}

__global__ void copy_mod_kernel(u64* out, size_t out_stride, const u64* in, size_t in_stride, u64 m, int batch,
                                u32 words) {
    size_t total = (size_t)batch * words;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        size_t b = idx / words, w = idx - b * words;
        u64 v = in[b * in_stride + w];
        out[b * out_stride + w] = m ? v % m : v;
    }
}

cudaError_t launch_copy_mod(u64* out, size_t out_stride, const u64* in, size_t in_stride, u64 m, int batch, u32 words,
                            cudaStream_t s) {
    size_t total = (size_t)batch * words;
    if (!total)
        return cudaSuccess;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16)
        blocks = 148 * 16;
    copy_mod_kernel<<<blocks, 256, 0, s>>>(out, out_stride, in, in_stride, m, batch, words);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// CiphertextMulMatrix: out[i][w] = sum_k ct[k][w] * M[k][i] mod modulus, exact.
// Tiled 32x32 through shared memory; products are reduced with 128-bit arithmetic so any modulus < 2^63 and any
// int64 matrix entry is handled exactly (the reference's FP64 GEMM is exact only while sums stay below 2^53).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 mulmod_u64(u64 a, u64 b, u64 m) {
    // a, b < m < 2^63: schoolbook via 128-bit product and division-free reduction using __umul64hi is overkill
    // here; this kernel is bandwidth/latency trivial compared with the bootstraps, so use the simple form.
    unsigned __int128 p = (unsigned __int128)a * b;
    return (u64)(p % m);
}

__global__ void mul_matrix_kernel(u64* out, const u64* ct, const i64* M, int in, int outc, u32 words, u64 modulus) {
    __shared__ u64 sC[32][33];  // ct tile   [k][w]
    __shared__ u64 sM[32][33];  // matrix tile [k][i]
    const int w = blockIdx.x * 32 + threadIdx.x;  // ciphertext word (column of ct)
    const int i0 = blockIdx.y * 32;
    u64 acc[4] = {0, 0, 0, 0};  // threadIdx.y in 0..7 handles i = i0 + threadIdx.y + 8*r
    for (int k0 = 0; k0 < in; k0 += 32) {
        for (int r = threadIdx.y; r < 32; r += 8) {
            int k = k0 + r;
            sC[r][threadIdx.x] = (k < in && w < (int)words) ? ct[(size_t)k * words + w] % modulus : 0;
            int i = i0 + threadIdx.x;
            u64 mv = 0;
            if (k < in && i < outc) {
                i64 x = M[(size_t)k * outc + i] % (i64)modulus;
                mv = (u64)(x < 0 ? x + (i64)modulus : x);
            }
            sM[r][threadIdx.x] = mv;
        }
        __syncthreads();
        for (int r = 0; r < 4; r++) {
            int il = threadIdx.y + 8 * r;
            u64 a = acc[r];
            for (int k = 0; k < 32; k++) {
                u64 p = mulmod_u64(sC[k][threadIdx.x], sM[k][il], modulus);
                a += p;
                if (a >= modulus)
                    a -= modulus;
            }
            acc[r] = a;
        }
        __syncthreads();
    }
    for (int r = 0; r < 4; r++) {
        int i = i0 + threadIdx.y + 8 * r;
        if (i < outc && w < (int)words)
            out[(size_t)i * words + w] = acc[r];
    }
}

// Fast path for moduli that fit 32 bits (every CiphertextMulMatrix the reference's own examples run: q = 2^10..2^17,
// GEMM.cpp:70-88): operands are reduced once into u32 planes, then a register-tiled integer GEMM accumulates
// 32 x 32 -> 64-bit products (IMAD.WIDE) in 64-bit registers.  A power-of-two modulus needs no intermediate reduction
// (wrap-around mod 2^64 is compatible); any other modulus is reduced every `chunk` terms, chunk * (m-1)^2 + m <= 2^64.
__global__ void mm_reduce_kernel(u32* dct, u32* dM, const u64* ct, const i64* M, size_t nct, size_t nM, u64 modulus) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < nct; x += stride)
        dct[x] = (u32)(ct[x] % modulus);
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < nM; x += stride) {
        i64 v = M[x] % (i64)modulus;
        dM[x] = (u32)(v < 0 ? v + (i64)modulus : v);
    }
}

constexpr int MM_TI = 64, MM_TW = 128, MM_TK = 16;   // CTA tile: 64 output rows x 128 ciphertext words, 16 terms a stage

__global__ void __launch_bounds__(256) mul_matrix32_kernel(u64* out, const u32* ct, const u32* M, int in, int outc,
                                                             u32 words, u64 modulus, int pow2, int chunk) {
    __shared__ __align__(16) u32 sM[MM_TK][MM_TI];
    __shared__ __align__(16) u32 sC[MM_TK][MM_TW];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int w0 = blockIdx.x * MM_TW, i0 = blockIdx.y * MM_TI;
    u64 acc[4][8];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 8; b++)
            acc[a][b] = 0;
    int since = 0;
    for (int k0 = 0; k0 < in; k0 += MM_TK) {
        // stage the two tiles (rows k0..k0+15): 16 x 64 matrix words, 16 x 128 ciphertext words
        for (int x = tid; x < MM_TK * MM_TI; x += 256) {
            const int k = k0 + x / MM_TI, i = i0 + x % MM_TI;
            sM[x / MM_TI][x % MM_TI] = (k < in && i < outc) ? M[(size_t)k * outc + i] : 0;
        }
        for (int x = tid; x < MM_TK * MM_TW; x += 256) {
            const int k = k0 + x / MM_TW, w = w0 + x % MM_TW;
            sC[x / MM_TW][x % MM_TW] = (k < in && w < (int)words) ? ct[(size_t)k * words + w] : 0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < MM_TK; k++) {
            const uint4 m4 = *reinterpret_cast<const uint4*>(&sM[k][ty * 4]);
            const uint4 c0 = *reinterpret_cast<const uint4*>(&sC[k][tx * 4]);
            const uint4 c1 = *reinterpret_cast<const uint4*>(&sC[k][64 + tx * 4]);
            const u32 mv[4] = {m4.x, m4.y, m4.z, m4.w};
            const u32 cv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 8; b++)
                    acc[a][b] += (u64)mv[a] * cv[b];
        }
        __syncthreads();
        since += MM_TK;
        if (!pow2 && since + MM_TK > chunk) {
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 8; b++)
                    acc[a][b] %= modulus;
            since = 2;   // the carried residue (< m) counts as at most two terms
        }
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {
        const int i = i0 + ty * 4 + a;
        if (i >= outc)
            continue;
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const int w = w0 + (b < 4 ? tx * 4 + b : 64 + tx * 4 + b - 4);
            if (w < (int)words)
                out[(size_t)i * words + w] = pow2 ? (acc[a][b] & (modulus - 1)) : (acc[a][b] % modulus);
        }
    }
}

size_t mul_matrix_scratch_bytes(int in, int outc, u32 words, u64 modulus) {
    if (modulus > (1ULL << 32))
        return 0;
    return ((size_t)in * words + (size_t)in * outc) * 4 + 64;
}

cudaError_t launch_mul_matrix(u64* out, const u64* ct, const i64* M, int in, int outc, u32 words, u64 modulus,
                              void* scratch, cudaStream_t s) {
    const bool pow2 = (modulus & (modulus - 1)) == 0;
    int chunk = 0;
    bool fast = scratch && modulus <= (1ULL << 32);
    if (fast && !pow2) {
        const unsigned __int128 sq = (unsigned __int128)(modulus - 1) * (modulus - 1);
        const unsigned __int128 c = ((((unsigned __int128)1) << 64) - modulus) / (sq ? sq : 1);
        chunk = c > (unsigned __int128)(1 << 30) ? (1 << 30) : (int)c;
        if (chunk < 2 * MM_TK + 1)
            fast = false;   // modulus above ~2^29.5: reductions would dominate; exact 128-bit kernel below
    }
    if (fast) {
        u32* dct = reinterpret_cast<u32*>(scratch);
        u32* dM = dct + (((size_t)in * words + 3) & ~(size_t)3);
        mm_reduce_kernel<<<148 * 4, 256, 0, s>>>(dct, dM, ct, M, (size_t)in * words, (size_t)in * outc, modulus);
        dim3 grid((words + MM_TW - 1) / MM_TW, (outc + MM_TI - 1) / MM_TI);
        mul_matrix32_kernel<<<grid, 256, 0, s>>>(out, dct, dM, in, outc, words, modulus, pow2 ? 1 : 0, chunk);
        return cudaGetLastError();
    }
    dim3 grid((words + 31) / 32, (outc + 31) / 32), block(32, 8);
    mul_matrix_kernel<<<grid, block, 0, s>>>(out, ct, M, in, outc, words, modulus);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// Setup-time key re-encoding (replaces the CPU loops of GPUSetup_core / KeyCopy_FFT, bootstrapping.cu:933-975)
// ---------------------------------------------------------------------------------------------------------
// generic layout: same element order as the reference; value -> Montgomery form times N^-1
template <typename T>
__global__ void bk_convert_generic_kernel(T* dst, const u64* src, size_t count, ModCtx<T> mod, T ninvM2) {
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < count;
         idx += (size_t)gridDim.x * blockDim.x)
        dst[idx] = mod.mont_mul((T)src[idx], ninvM2);
}

template <typename T>
cudaError_t launch_bk_convert_generic(T* dst, const u64* src, size_t count, ModCtx<T> mod, T ninvM2, cudaStream_t s) {
    if (!count)
        return cudaSuccess;
    bk_convert_generic_kernel<T><<<148 * 8, 256, 0, s>>>(dst, src, count, mod, ninvM2);
    return cudaGetLastError();
}
template cudaError_t launch_bk_convert_generic<u32>(u32*, const u64*, size_t, ModCtx<u32>, u32, cudaStream_t);
template cudaError_t launch_bk_convert_generic<u64>(u64*, const u64*, size_t, ModCtx<u64>, u64, cudaStream_t);

// per-ciphertext LUT expansion for EvalFunc(vector, LUT_vec) (binfhe-base-scheme.cpp:791-924)
//   mode 1 (periodic):  t[x] = x < q/2 ? L[x] : q - L[x - q/2]                       (tab_len = q)
//   mode 2 (arbitrary): t[x] = x < dq/2 ? L[x mod q] : dq - L[(x - dq/2) mod q]        (tab_len = dq = 2q)
__global__ void lut_expand_kernel(u64* out, const u64* lut, u64 q, u64 tab_len, int mode, size_t total) {
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        size_t b = idx / tab_len;
        u64 x = idx - b * tab_len;
        const u64* L = lut + b * q;
        u64 v;
        if (mode == 1)
            v = (x < q / 2) ? L[x] : q - L[x - q / 2];
        else
            v = (x < tab_len / 2) ? L[x % q] : tab_len - L[(x - tab_len / 2) % q];
        out[idx] = v;
    }
}

cudaError_t launch_lut_expand(u64* out, const u64* lut, u64 q, u64 tab_len, int mode, int batch, cudaStream_t s) {
    size_t total = (size_t)batch * tab_len;
    if (!total)
        return cudaSuccess;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16)
        blocks = 148 * 16;
    lut_expand_kernel<<<blocks, 256, 0, s>>>(out, lut, q, tab_len, mode, total);
    return cudaGetLastError();
}

// KSK: u64 rows of `words` entries -> narrow rows of row_stride entries (zero padded)
template <typename TK>
__global__ void ksk_convert_kernel(TK* dst, u32 row_stride, const u64* src, size_t rows, u32 words) {
    size_t total = rows * row_stride;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        size_t r = idx / row_stride;
        u32 w = (u32)(idx - r * row_stride);
        dst[idx] = w < words ? (TK)src[r * words + w] : (TK)0;
    }
}

cudaError_t launch_ksk_convert(void* dst, int bytes, u32 row_stride, const u64* src, size_t rows, u32 words,
                               cudaStream_t s) {
    if (!rows)
        return cudaSuccess;
    if (bytes == 2)
        ksk_convert_kernel<unsigned short><<<148 * 8, 256, 0, s>>>((unsigned short*)dst, row_stride, src, rows, words);
    else if (bytes == 4)
        ksk_convert_kernel<u32><<<148 * 8, 256, 0, s>>>((u32*)dst, row_stride, src, rows, words);
    else
        ksk_convert_kernel<u64><<<148 * 8, 256, 0, s>>>((u64*)dst, row_stride, src, rows, words);
    return cudaGetLastError();
}

}  // namespace tfhe_b200
