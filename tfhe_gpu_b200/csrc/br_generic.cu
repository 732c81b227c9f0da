// Generic blind-rotation kernel: one CTA per ciphertext, any power-of-two N <= 4096, 32- or 64-bit modulus,
// CGGI/GINX (ternary, two keys) and AP/DM accumulators.  It is the production path for the 54-bit functional
// parameter sets and for DM, and the cross-check for the specialised 32-bit CGGI kernel (br_cggi32.cu).
//
// Follows (semantics only) RingGSWAccumulatorCGGI::EvalAcc/AddToAccCGGI (rgsw-acc-cggi.cpp:143-155,246-307),
// RingGSWAccumulatorDM::EvalAcc/AddToAccDM (rgsw-acc-dm.cpp:80-110,306-359), SignedDigitDecompose
// (rgsw-acc.cpp:57-111), BootstrapGateCore/BootstrapFuncCore (binfhe-base-scheme.cpp:1087-1192) and the
// extraction in EvalBinGate/BootstrapFunc (:660-672, :1199-1205).
//
// Re-design relative to the reference (both its CPU and its FFT GPU path):
//   * the accumulator lives in COEFFICIENT form in shared memory for the whole rotation; each step computes
//     delta = INTT( sum_l NTT(digit_l) * BK_l * monomial ) and adds it (INTT is linear, so this equals the
//     oracle's "acc_eval += ...; INTT(acc_eval)" exactly) -- no accumulator traffic to global memory at all;
//   * N^-1 of the inverse transform and the Montgomery factor are folded into the bootstrapping key at setup;
//   * monomial factors w_k^e - 1 come from one 2N-entry table of psi powers instead of a 2N x N table;
//   * accumulator initialisation (gate / LUT), the LWE-mask modulus switch, sample extraction and the
//     a(X)->a(X^-1) transpose are fused into the same kernel.
#include "engine.cuh"

namespace tfhe_b200 {

template <typename T>
struct GenArgs {
    BRCommon c;
    BRTables<T> t;
};

// forward negacyclic NTT (Cooley-Tukey, natural -> bit-reversed) of `npoly` polynomials stored back to back
template <typename T>
__device__ __forceinline__ void ntt_fwd_smem(T* a, u32 npoly, u32 N, u32 logN, const T* __restrict__ tw,
                                             const ModCtx<T>& M) {
    const u32 half = N >> 1;
    u32 t = N;
    for (u32 m = 1; m < N; m <<= 1) {
        t >>= 1;
        const u32 tshift = __ffs(t) - 1;
        for (u32 b = threadIdx.x; b < npoly * half; b += blockDim.x) {
            u32 p = b / half, x = b - p * half;
            u32 i = x >> tshift, jj = x & (t - 1);
            u32 j = (i << (tshift + 1)) + jj;
            T* P = a + (size_t)p * N;
            T W = tw[m + i];
            T U = P[j], V = M.mont_mul(P[j + t], W);
            P[j] = M.add(U, V);
            P[j + t] = M.sub(U, V);
        }
        __syncthreads();
    }
}

// inverse negacyclic NTT (Gentleman-Sande, bit-reversed -> natural), WITHOUT the N^-1 scaling
template <typename T>
__device__ __forceinline__ void ntt_inv_smem(T* a, u32 npoly, u32 N, u32 logN, const T* __restrict__ twi,
                                             const ModCtx<T>& M) {
    const u32 half = N >> 1;
    u32 t = 1;
    for (u32 m = N; m > 1; m >>= 1) {
        const u32 h = m >> 1;
        const u32 tshift = __ffs(t) - 1;
        for (u32 b = threadIdx.x; b < npoly * half; b += blockDim.x) {
            u32 p = b / half, x = b - p * half;
            u32 i = x >> tshift, jj = x & (t - 1);
            u32 j = (i << (tshift + 1)) + jj;
            T* P = a + (size_t)p * N;
            T S = twi[h + i];
            T U = P[j], V = P[j + t];
            P[j] = M.add(U, V);
            P[j + t] = M.mont_mul(M.sub(U, V), S);
        }
        __syncthreads();
        t <<= 1;
    }
}

// ---- 64-bit lazy variants (Q < 2^59 / 26): Shoup butterflies, no per-stage corrections in the forward transform ------
__device__ __forceinline__ u64 shoup_mul64(u64 y, u64 w, u64 wp, u64 Q) {
    u64 q = __umul64hi(y, wp);
    return y * w - q * Q;   // in [0, 2Q) for any 64-bit y
}
__device__ __forceinline__ void ntt_fwd_lazy64(u64* a, u32 npoly, u32 N, const u64* __restrict__ sh, u64 Q) {
    // forward: values grow by 2Q per stage (digits < Q on entry, < (1 + 2 logN) Q <= 25 Q on exit)
    const u32 half = N >> 1;
    const u64 Q2 = 2 * Q;
    u32 t = N;
    for (u32 m = 1; m < N; m <<= 1) {
        t >>= 1;
        const u32 tshift = __ffs(t) - 1;
        for (u32 b = threadIdx.x; b < npoly * half; b += blockDim.x) {
            u32 p = b / half, x = b - p * half;
            u32 i = x >> tshift, jj = x & (t - 1);
            u32 j = (i << (tshift + 1)) + jj;
            u64* P = a + (size_t)p * N;
            u64 U = P[j], V = shoup_mul64(P[j + t], sh[m + i], sh[N + m + i], Q);
            P[j] = U + V;
            P[j + t] = U - V + Q2;
        }
        __syncthreads();
    }
}
__device__ __forceinline__ void ntt_inv_lazy64(u64* a, u32 npoly, u32 N, const u64* __restrict__ sh, u64 Q) {
    // inverse (no N^-1): all values kept below 2Q
    const u32 half = N >> 1;
    const u64 Q2 = 2 * Q;
    u32 t = 1;
    for (u32 m = N; m > 1; m >>= 1) {
        const u32 h = m >> 1;
        const u32 tshift = __ffs(t) - 1;
        for (u32 b = threadIdx.x; b < npoly * half; b += blockDim.x) {
            u32 p = b / half, x = b - p * half;
            u32 i = x >> tshift, jj = x & (t - 1);
            u32 j = (i << (tshift + 1)) + jj;
            u64* P = a + (size_t)p * N;
            u64 U = P[j], V = P[j + t];
            u64 sum = U + V;
            P[j] = sum >= Q2 ? sum - Q2 : sum;
            P[j + t] = shoup_mul64(U - V + Q2, sh[h + i], sh[N + h + i], Q);
        }
        __syncthreads();
        t <<= 1;
    }
}
struct Acc128 {
    u64 lo, hi;
    __device__ __forceinline__ void mac(u64 x, u64 b) {
        u64 pl = x * b, ph = __umul64hi(x, b);
        lo += pl;
        hi += ph + (lo < pl);
    }
};
// Montgomery reduction of a 128-bit value X < Q * 2^64 -> [0, Q)
__device__ __forceinline__ u64 redc128(const Acc128& X, const ModCtx<u64>& M) {
    u64 m = X.lo * M.qinv;
    u64 t = __umul64hi(m, M.Q);
    u64 r = X.hi - t;
    return X.hi < t ? r + M.Q : r;
}

template <typename T>
__device__ __forceinline__ void decompose_smem(const T* c, T* D, u32 N, u32 gBits, u32 numThrow, u32 kept,
                                               const ModCtx<T>& M) {
    // rgsw-acc.cpp:83-108: centred representative, sign-extended low gBits digits, no final carry fix-up
    const u64 Q = (u64)M.Q, QHalf = Q >> 1;
    const int sh = 64 - (int)gBits;
    for (u32 idx = threadIdx.x; idx < 2 * N; idx += blockDim.x) {
        u32 j = idx / N, k = idx - j * N;
        u64 tv = (u64)c[idx];
        i64 dv = (tv < QHalf) ? (i64)tv : (i64)tv - (i64)Q;
        for (u32 i = 0; i < numThrow; i++) {
            i64 r = (i64)((u64)dv << sh) >> sh;
            dv = (dv - r) >> gBits;
        }
        for (u32 l = 0; l < kept; l++) {
            i64 r = (i64)((u64)dv << sh) >> sh;
            dv = (dv - r) >> gBits;
            if (r < 0)
                r += (i64)Q;
            D[(size_t)(j + 2 * l) * N + k] = (T)r;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(1024, 1) br_generic_kernel(GenArgs<T> A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const BRCommon& C = A.c;
    const ModCtx<T> M = A.t.mod;
    const u32 N = C.N, n = C.n, d = C.d, logN = C.logN;
    T* c = reinterpret_cast<T*>(smem_raw);  // [2][N] coefficient-form accumulator
    T* D = c + 2 * N;                       // [d][N] digits / NTT scratch; D[0..1] doubles as delta
    const u64 Q = (u64)M.Q;
    const u32 twoN = 2 * N;

    for (int ct = blockIdx.x; ct < C.batch; ct += gridDim.x) {
        const u64* lwe = C.ct + (size_t)ct * (n + 1);
        // ------------------------------------------------------------------ accumulator initialisation
        if (C.acc_init == ACC_EXPLICIT) {
            const u64* src = C.acc_io + (size_t)ct * 2 * N;
            for (u32 idx = threadIdx.x; idx < 2 * N; idx += blockDim.x)
                c[idx] = (T)src[idx];
        }
        else {
            const u64 b = lwe[n], q = C.ct_mod;
            const u32 factor = (u32)(twoN / q);
            const u64 qHalf = q >> 1;
            for (u32 idx = threadIdx.x; idx < 2 * N; idx += blockDim.x)
                c[idx] = 0;
            __syncthreads();
            if (C.acc_init == ACC_GATE) {
                const u64 q1 = C.gate_q1;
                u64 q2 = q1 + qHalf;
                if (q2 >= q)
                    q2 -= q;
                const u64 Q8 = C.Q8, Q8Neg = Q - C.Q8;
                for (u64 j = threadIdx.x; j < qHalf; j += blockDim.x) {
                    u64 bb = b % q;
                    u64 temp = bb >= j ? bb - j : bb + q - j;
                    u64 v;
                    if (q1 < q2)
                        v = ((temp >= q1) && (temp < q2)) ? Q8Neg : Q8;
                    else
                        v = ((temp >= q2) && (temp < q1)) ? Q8 : Q8Neg;
                    c[N + j * factor] = (T)v;
                }
            }
            else {
                const u64* tab = C.table + (C.acc_init == ACC_TABLE_PER ? (size_t)ct * q : 0);
                for (u64 j = threadIdx.x; j < qHalf; j += blockDim.x) {
                    u64 bb = b % q;
                    u64 temp = bb >= j ? bb - j : bb + q - j;
                    c[N + j * factor] = (T)(C.scale * tab[temp]);
                }
            }
        }
        __syncthreads();

        // ------------------------------------------------------------------ blind rotation
        const bool dm = (C.method == TFHE_B200_METHOD_AP);
        const u32 steps = dm ? n * C.digitsR : n;
        const size_t rgsw_stride = (size_t)d * 2 * N;
        u64 aI = 0;
        for (u32 step = 0; step < steps; step++) {
            const T* ek0;
            const T* ek1 = nullptr;
            u32 e = 0;
            if (!dm) {
                // rgsw-acc-cggi.cpp:146-153: e = ((mod - a_i) mod mod) * (2N / mod)
                u64 ai = lwe[step] % C.ct_mod;
                e = (u32)(((C.ct_mod - ai) % C.ct_mod) * (twoN / C.ct_mod));
                if (e == 0)
                    continue;  // monomial X^0 - 1 = 0
                // same element order as the reference export: [key][i][l][j][N]
                ek0 = A.t.bk + ((size_t)0 * n + step) * rgsw_stride;
                ek1 = A.t.bk + ((size_t)1 * n + step) * rgsw_stride;
            }
            else {
                // rgsw-acc-dm.cpp:102-109
                u32 i = step / C.digitsR, k = step - i * C.digitsR;
                if (k == 0) {
                    u64 q = C.q_lwe;
                    aI = (q - lwe[i] % q) % q;
                }
                u32 a0 = (u32)(aI % C.baseR);
                aI /= C.baseR;
                if (a0 == 0)
                    continue;
                ek0 = A.t.bk + (((size_t)i * C.baseR + a0) * C.digitsR + k) * rgsw_stride;
            }

            decompose_smem<T>(c, D, N, C.gBits, C.numThrow, C.digitsKept, M);
            __syncthreads();
            if constexpr (sizeof(T) == 8)
                ntt_fwd_lazy64(reinterpret_cast<u64*>(D), d, N, reinterpret_cast<const u64*>(A.t.sh_fwd), Q);
            else
                ntt_fwd_smem<T>(D, d, N, logN, A.t.tw_fwd, M);

            // pointwise multiply-accumulate against the RGSW key(s); delta written over D[0], D[1]
            if constexpr (sizeof(T) == 8) {
                // 64-bit path: lazy digits (< 25 Q) times key words (< Q) accumulate in 128 bits, ONE Montgomery
                // reduction per output (sum < d * 25 Q^2 < Q 2^64 because Q < 2^55 and d <= 12: checked at setup)
                const ModCtx<u64>& M8 = reinterpret_cast<const ModCtx<u64>&>(M);
                const u64* D8 = reinterpret_cast<const u64*>(D);
                const u64* e0 = reinterpret_cast<const u64*>(ek0);
                const u64* e1 = reinterpret_cast<const u64*>(ek1);
                for (u32 k = threadIdx.x; k < N; k += blockDim.x) {
                    Acc128 a00{0, 0}, a01{0, 0}, a10{0, 0}, a11{0, 0};
                    const u32 l0 = dm ? 1 : 0;
                    for (u32 l = l0; l < d; l++) {
                        u64 x = D8[(size_t)l * N + k];
                        a00.mac(x, e0[((size_t)l * 2 + 0) * N + k]);
                        a01.mac(x, e0[((size_t)l * 2 + 1) * N + k]);
                        if (!dm) {
                            a10.mac(x, e1[((size_t)l * 2 + 0) * N + k]);
                            a11.mac(x, e1[((size_t)l * 2 + 1) * N + k]);
                        }
                    }
                    u64 s00 = redc128(a00, M8), s01 = redc128(a01, M8);
                    if (!dm) {
                        u64 s10 = redc128(a10, M8), s11 = redc128(a11, M8);
                        u32 br = __brev(k) >> (32 - logN);
                        u32 x = ((2 * br + 1) * e) & (twoN - 1);
                        const u64* pp = reinterpret_cast<const u64*>(A.t.psi_pow);
                        u64 m1 = M8.sub(pp[x], M8.oneM);
                        u64 m2 = M8.sub(pp[(twoN - x) & (twoN - 1)], M8.oneM);
                        Acc128 t0{0, 0}, t1{0, 0};
                        t0.mac(s00, m1); t0.mac(s10, m2);
                        t1.mac(s01, m1); t1.mac(s11, m2);
                        s00 = redc128(t0, M8);
                        s01 = redc128(t1, M8);
                    }
                    reinterpret_cast<u64*>(D)[k] = s00;
                    reinterpret_cast<u64*>(D)[N + k] = s01;
                }
            }
            else
            for (u32 k = threadIdx.x; k < N; k += blockDim.x) {
                T s00 = 0, s01 = 0, s10 = 0, s11 = 0;
                const u32 l0 = dm ? 1 : 0;  // rgsw-acc-dm.cpp:353,357: the DM sums start at l = 1
                for (u32 l = l0; l < d; l++) {
                    T x = D[(size_t)l * N + k];
                    s00 = M.add(s00, M.mont_mul(x, ek0[((size_t)l * 2 + 0) * N + k]));
                    s01 = M.add(s01, M.mont_mul(x, ek0[((size_t)l * 2 + 1) * N + k]));
                    if (!dm) {
                        s10 = M.add(s10, M.mont_mul(x, ek1[((size_t)l * 2 + 0) * N + k]));
                        s11 = M.add(s11, M.mont_mul(x, ek1[((size_t)l * 2 + 1) * N + k]));
                    }
                }
                if (!dm) {
                    // evaluation slot k holds the evaluation at psi^(2*bitrev(k)+1); monomial X^e there is
                    // psi^((2*bitrev(k)+1)*e mod 2N)   (rgsw-cryptoparameters.h:141-159)
                    u32 br = __brev(k) >> (32 - logN);
                    u32 x = ((2 * br + 1) * e) & (twoN - 1);
                    T m1 = M.sub(A.t.psi_pow[x], M.oneM);
                    T m2 = M.sub(A.t.psi_pow[(twoN - x) & (twoN - 1)], M.oneM);
                    s00 = M.add(M.mont_mul(s00, m1), M.mont_mul(s10, m2));
                    s01 = M.add(M.mont_mul(s01, m1), M.mont_mul(s11, m2));
                }
                D[k] = s00;
                D[N + k] = s01;
            }
            __syncthreads();
            if constexpr (sizeof(T) == 8) {
                ntt_inv_lazy64(reinterpret_cast<u64*>(D), 2, N, reinterpret_cast<const u64*>(A.t.sh_inv), Q);
                for (u32 idx = threadIdx.x; idx < 2 * N; idx += blockDim.x) {   // D < 2Q
                    u64 v = (u64)D[idx];
                    v = v >= Q ? v - Q : v;
                    c[idx] = dm ? (T)v : M.add(c[idx], (T)v);
                }
            }
            else {
                ntt_inv_smem<T>(D, 2, N, logN, A.t.tw_inv, M);
                if (dm) {
                    for (u32 idx = threadIdx.x; idx < 2 * N; idx += blockDim.x)
                        c[idx] = D[idx];
                }
                else {
                    for (u32 idx = threadIdx.x; idx < 2 * N; idx += blockDim.x)
                        c[idx] = M.add(c[idx], D[idx]);
                }
            }
            __syncthreads();
        }

        // ------------------------------------------------------------------ extraction (+ transpose of a)
        // a'(X) = a(X^-1): a'[0] = a[0], a'[i] = -a[N-i]   (binfhe-base-scheme.cpp:93, bootstrapping.cu:675-685)
        if (C.write_acc) {
            u64* dst = C.acc_io + (size_t)ct * 2 * N;
            for (u32 idx = threadIdx.x; idx < N; idx += blockDim.x) {
                u64 v = idx == 0 ? (u64)c[0] : (u64)c[N - idx];
                dst[idx] = (idx == 0 || v == 0) ? v : Q - v;
                dst[N + idx] = (u64)c[N + idx];
            }
        }
        if (C.ext) {
            u64* dst = C.ext + (size_t)ct * (N + 1);
            for (u32 idx = threadIdx.x; idx <= N; idx += blockDim.x) {
                u64 v;
                if (idx == N) {
                    v = (u64)c[N] + C.ext_add_b;
                    if (v >= Q)
                        v -= Q;
                }
                else if (idx == 0)
                    v = (u64)c[0];
                else {
                    v = (u64)c[N - idx];
                    v = v ? Q - v : 0;
                }
                dst[idx] = v;
            }
        }
        __syncthreads();
    }
}

template <typename T>
cudaError_t launch_br_generic(const BRCommon& c, const BRTables<T>& t, cudaStream_t s, int sm_count) {
    GenArgs<T> A;
    A.c = c;
    A.t = t;
    size_t smem = (size_t)(2 + c.d) * c.N * sizeof(T);
    u32 threads = c.N / 2;
    if (threads > 1024)
        threads = 1024;
    if (threads < 128)
        threads = 128;
    static bool attr_set[2] = {false, false};
    (void)attr_set;
    cudaError_t e = cudaFuncSetAttribute(br_generic_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess)
        return e;
    int grid = c.batch;
    (void)sm_count;
    br_generic_kernel<T><<<grid, threads, smem, s>>>(A);
    return cudaGetLastError();
}

template cudaError_t launch_br_generic<u32>(const BRCommon&, const BRTables<u32>&, cudaStream_t, int);
template cudaError_t launch_br_generic<u64>(const BRCommon&, const BRTables<u64>&, cudaStream_t, int);

}  // namespace tfhe_b200
