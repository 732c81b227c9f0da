// AP/DM blind rotation for the N = 2048 rings (any modulus below 2^54, in 64-bit words): the DM accumulator
// (rgsw-acc-dm.cpp:80-110, 306-359) on the "wide" register-resident transform of br_cggi64w.cu (ntt64w.cuh: 128 threads x
// 16 coefficients per polynomial, 16 warps per SM with two ciphertexts per CTA).  Covers every N = 2048 set of the
// reference's paramsMap under method AP: STD192 / STD192_OPT / STD192Q / STD192Q_OPT (three digits) and STD256 /
// STD256_OPT / STD256Q / STD256Q_OPT (four digits) with top-digit elimination (their top signed digit is exact:
// cggi32_skip_top_ok), STD128Q / STD128Q_OPT (two digits, the top one can wrap) on the plain path.
//
//   for i < n, for k < digitsR:  a0 = k-th base-baseR digit of (q - a_i) mod q;  if a0 == 0 skip
//       acc[j] = sum_{l'=1}^{d-1} NTT(digit_l') * BK[i][a0][k][l'][j]            (REPLACE; row l' = 0 is dropped)
//
// Relative to br_cggi64w.cu (same relation as br_dm32.cu to br_cggi32.cu):
//   * the key row is selected by the ciphertext's own refresh digit, so the ciphertexts of a CTA share no key words: the
//     pointwise stage is ciphertext-major (the 256 threads of a ciphertext cover its 2048 slots, every thread loads the
//     key words of its own slots), ciphertexts are synchronised by their own named barrier and a ciphertext whose digit
//     is zero skips the step outright;
//   * the accumulator is REPLACED: the evaluation-domain accumulator that top-digit elimination needs is simply the
//     previous pointwise result, 2 (DK - 1) forward + 2 inverse transforms and 2 D multiply-accumulates per slot per
//     active step, no monomial factors;
//   * the dropped row l' = 0 and the eliminated top digit are absorbed by the key transform (bk_relayout_dm64_kernel);
//   * the refresh digit of a step is recomputed from the ciphertext (one broadcast load) instead of being tabulated:
//     the digit regions leave no shared memory for a step table.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ntt64w.cuh"

namespace tfhe_b200 {

namespace {

using namespace w64;

struct DM64WArgs {
    BRCommon c;
    ModCtx<u64> mod;
    const u64* bk;       // [row = (i*baseR + a0)*digitsR + k][x(D planes)][slot][2] 27-bit limb pairs: plane l', word j
    const u64* twC;      // tables of cggi64w_build_tables
    const u64* twB;
    const u64* twU;
    u64 Q2, dig_off, dig_add, ninvM, zero64;
};

// PLAIN = true: no top-digit elimination (two digits whose top digit can wrap: STD128Q / STD128Q_OPT): all DK digits of
// both components are transformed, the key keeps its own rows with row l' = 0 zeroed, the result lands in rows 0 / 1.
template <int DK, int G, bool PLAIN = false>
__global__ void __launch_bounds__(KW<DK, G>::NT, 1) br_dm64w_kernel(const __grid_constant__ DM64WArgs A) {
    using K = KW<DK, G>;
    constexpr int D = K::D, NT = K::NT, NF = PLAIN ? DK : DK - 1;
    constexpr int RES = PLAIN ? 0 : 2 * (DK - 1);   // rows that receive the pointwise result (a, b)
    constexpr int CT_THREADS = 2 * TPN;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* Dsm = reinterpret_cast<u64*>(smem_raw);                                  // [G][D][N]
    ulonglong2* twC = reinterpret_cast<ulonglong2*>(Dsm + (size_t)G * D * N);    // [15][128]
    ulonglong2* twB = twC + 15 * TPN;                                             // [16][8]
    ulonglong2* twUf = twB + 16 * 8;                                              // [15] uniform forward
    ulonglong2* twUi = twUf + 15;                                                 // [15] uniform inverse, negated

    const BRCommon& C = A.c;
    const u64 Q = A.mod.Q, Q2 = A.Q2, QO = 2 * A.Q2, nQ = 0 - A.mod.Q, qinv = A.mod.qinv;
    const u64 Z = A.zero64;
    const u32 n = C.n;
    const int tid = threadIdx.x;
    const int g = tid / CT_THREADS, j = (tid / TPN) & 1, T = tid % TPN, lt = tid % CT_THREADS;
    const int bar_id = 1 + g * 2 + j;          // the 128 threads of one (ciphertext, component)
    const int ct_bar = 1 + 2 * G + g;          // the 256 threads of one ciphertext
    const int ct = blockIdx.x * G + g;
    const bool live = ct < C.batch;
    const u64* lwe = C.ct + (size_t)(live ? ct : 0) * (n + 1);
    const int blk = T >> 3, u8 = T & 7;
    auto ct_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(ct_bar), "r"(CT_THREADS) : "memory"); };

    for (int x = tid; x < 15 * TPN; x += NT)
        twC[x] = reinterpret_cast<const ulonglong2*>(A.twC)[x];
    for (int x = tid; x < 16 * 8; x += NT)
        twB[x] = reinterpret_cast<const ulonglong2*>(A.twB)[x];
    for (int x = tid; x < 2 * 15; x += NT)
        twUf[x] = reinterpret_cast<const ulonglong2*>(A.twU)[x];

    // ---- accumulator initialisation in A layout (coefficient idx = T + 128 r) -----------------------------------
    u64 c[CPT];
    if (C.acc_init == ACC_EXPLICIT) {
        const u64* src = C.acc_io + ((size_t)(live ? ct : 0) * 2 + j) * N;
#pragma unroll
        for (int r = 0; r < CPT; r++)
            c[r] = live ? src[T + TPN * r] : 0;
    }
    else {
        const u32 q = (u32)C.ct_mod, b = (u32)(lwe[n] % q);
        const u32 factor = (2 * N) / q, fshift = __ffs(factor) - 1;
        const u32 q1 = (u32)C.gate_q1;
        u32 q2 = q1 + (q >> 1);
        if (q2 >= q)
            q2 -= q;
        const u64* tab = C.table + (C.acc_init == ACC_TABLE_PER ? (size_t)(live ? ct : 0) * q : 0);
#pragma unroll
        for (int r = 0; r < CPT; r++) {
            const u32 idx = T + TPN * r;
            u64 val = 0;
            if (j == 1 && live && (idx & (factor - 1)) == 0) {
                u32 jj = idx >> fshift;
                u32 temp = b >= jj ? b - jj : b + q - jj;
                if (C.acc_init == ACC_GATE) {
                    bool in = (q1 < q2) ? ((temp >= q1) && (temp < q2)) : !((temp >= q2) && (temp < q1));
                    val = in ? Q - C.Q8 : C.Q8;
                }
                else
                    val = C.scale * tab[temp];
            }
            c[r] = val;
        }
    }
    __syncthreads();

    u64* myD = Dsm + (size_t)g * D * N;
    u64* top0 = myD + (size_t)RES * N;                 // evaluation-domain accumulator rows (a, b)
    u64* top = top0 + (size_t)j * N;
    const u64 QHalf = Q >> 1;
    const u32 gBits = C.gBits;
    const u64 gmask = ((u64)1 << gBits) - 1;

    // forward transform of v (A layout in) through region `reg`; result in registers in C layout of block T, < 29 Q
    auto forward = [&](u64 (&v)[CPT], u64* reg) {
        fwd_pass4(v, twUf, 1, 0, nQ, QO, Z);
#pragma unroll
        for (int r = 0; r < CPT; r++)
            reg[posw(T + TPN * r)] = v[r];
        group_sync128(bar_id);
        load_Bw(v, reg, blk, u8);
        fwd_pass3(v, twB + 8 * blk, nQ, QO, Z);
        store_Bw(v, reg, blk, u8);
        __syncwarp();
        load_C(v, reg, T);
        fwd_pass4(v, twC, TPN, T, nQ, QO, Z);
        const u64 Q16 = 4 * QO;
#pragma unroll
        for (int r = 0; r < CPT; r++)
            v[r] = csub(v[r], Q16);   // 11 lazy stages: < 45 Q -> < 29 Q (27-bit limb split of the pointwise stage)
    };

    if (!PLAIN) {
        // evaluation-domain accumulator (scaled by N^-1, see br_cggi32.cu) of the initial accumulator
        u64 v[CPT];
#pragma unroll
        for (int r = 0; r < CPT; r++)
            v[r] = c[r];
        forward(v, top);
#pragma unroll
        for (int r = 0; r < CPT; r++)
            v[r] = A.mod.mont_mul(v[r], A.ninvM);
        __syncwarp();
        store_C(v, top, T);
    }
    __syncthreads();

    // rgsw-acc-dm.cpp:102-109: aI = (q - a_i) mod q with the scheme's q; digit k of aI in base baseR
    const u32 qs = (u32)C.q_lwe, baseR = C.baseR, digitsR = C.digitsR;
    for (u32 i = 0; i < n; i++) {
        u32 aI = live ? (qs - (u32)(lwe[i] % qs)) % qs : 0;
        for (u32 k = 0; k < digitsR; k++, aI /= baseR) {
            const u32 a0 = aI % baseR;       // uniform over the ciphertext's threads
            if (a0 == 0)
                continue;
            const size_t row = ((size_t)i * baseR + a0) * digitsR + k;
            // ---- phase 1: digits 0..DK-2 of component j -> forward transforms -------------------------------------
            if (NF > 1) {
                u64* park = myD + (size_t)j * N;
#pragma unroll
                for (int r = 0; r < CPT; r++)
                    park[posw(T + TPN * r)] = c[r];
            }
#pragma unroll 1
            for (int li = 0; li < NF; li++) {
                const int l = NF - 1 - li;
                if (NF > 1) {
                    const u64* park = myD + (size_t)j * N;
#pragma unroll
                    for (int r = 0; r < CPT; r++)
                        c[r] = park[posw(T + TPN * r)];
                }
                u64 v[CPT];
                const u32 sh = gBits * l;
#pragma unroll
                for (int r = 0; r < CPT; r++) {
                    i64 dv = (c[r] < QHalf) ? (i64)c[r] : (i64)c[r] - (i64)Q;
                    u64 Dv = (u64)(dv + (i64)A.dig_off);
                    v[r] = ((u64)((i64)Dv >> sh) & gmask) + A.dig_add;
                }
                u64* reg = myD + (size_t)(j + 2 * l) * N;
                forward(v, reg);
                __syncwarp();
                store_C(v, reg, T);
            }
            ct_sync();

            // ---- phase 2: pointwise product with the ciphertext's own key row; the result REPLACES the accumulator --
            {
                constexpr int ITERS = N / CT_THREADS;
                const ulonglong2* bkr = reinterpret_cast<const ulonglong2*>(A.bk) + row * (size_t)D * N;
#pragma unroll 1
                for (int it = 0; it < ITERS; it++) {
                    const int k2 = lt + it * CT_THREADS;
                    ulonglong2 kw[D];
#pragma unroll
                    for (int x = 0; x < D; x++)
                        kw[x] = __ldg(bkr + (size_t)x * N + k2);
                    const u32 pk = posw(k2);
                    u64* dreg = myD + pk;
                    u64 xd[D];
#pragma unroll
                    for (int l = 0; l < D; l++)
                        xd[l] = dreg[(size_t)l * N];
                    const Limb x0(xd[0]);
                    L3 a0s(x0, kw[0].x), a1s(x0, kw[0].y);
#pragma unroll
                    for (int l = 1; l < D; l++) {
                        const Limb x(xd[l]);
                        a0s.mac(x, kw[l].x);
                        a1s.mac(x, kw[l].y);
                    }
                    dreg[(size_t)RES * N] = redc128(a0s.value(), Q, qinv);
                    dreg[(size_t)(RES + 1) * N] = redc128(a1s.value(), Q, qinv);
                }
            }
            ct_sync();

            // ---- phase 3: c = INTT(evaluation-domain accumulator), mirrored blocks, scratch = row j ------------------
            {
                u64 v[CPT];
                u64* reg = myD + (size_t)j * N;
                const int Tv = TPN - 1 - T;                 // mirrored 16-block; its 128-block is 15 - blk
                load_C(v, top, Tv);
                inv_pass4(v, twC, TPN, T, true, nQ, QO, Z);
                store_C(v, reg, Tv);
                __syncwarp();
                load_Bw(v, reg, 15 - blk, u8);
                inv_pass3(v, twB + 8 * blk, nQ, QO, Z);
                store_Bw(v, reg, 15 - blk, u8);
                group_sync128(bar_id);
#pragma unroll
                for (int r = 0; r < CPT; r++)
                    v[r] = reg[posw(T + TPN * r)];
                group_sync128(bar_id);                      // the next phase 1 overwrites row j (parking / digit 0)
                inv_pass4(v, twUi, 1, 0, false, nQ, QO, Z);
#pragma unroll
                for (int r = 0; r < CPT; r++)
                    c[r] = csub(csub(v[r], Q2), Q);         // v < 4Q
            }
        }
    }

    if (live) {
        if (C.write_acc) {
            u64* dst = C.acc_io + (size_t)ct * 2 * N;
#pragma unroll
            for (int r = 0; r < CPT; r++) {
                const u32 idx = T + TPN * r;
                if (j == 0) {
                    u64 val = c[r];
                    dst[idx == 0 ? 0 : N - idx] = (idx == 0 || val == 0) ? val : Q - val;
                }
                else
                    dst[N + idx] = c[r];
            }
        }
        if (C.ext) {
            u64* dst = C.ext + (size_t)ct * (N + 1);
#pragma unroll
            for (int r = 0; r < CPT; r++) {
                const u32 idx = T + TPN * r;
                if (j == 0) {
                    u64 val = c[r];
                    dst[idx == 0 ? 0 : N - idx] = (idx == 0 || val == 0) ? val : Q - val;
                }
                else if (idx == 0) {
                    u64 val = c[r] + C.ext_add_b;
                    dst[N] = val >= Q ? val - Q : val;
                }
            }
        }
    }
}

template <int DK, int G, bool PLAIN = false>
cudaError_t launch_dm_w(const DM64WArgs& a, cudaStream_t s) {
    using K = KW<DK, G>;
    if (K::smem > 227 * 1024)
        return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(br_dm64w_kernel<DK, G, PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)K::smem);
    if (e != cudaSuccess)
        return e;
    br_dm64w_kernel<DK, G, PLAIN><<<(a.c.batch + G - 1) / G, K::NT, K::smem, s>>>(a);
    return cudaGetLastError();
}

}  // namespace

// N = 2048, method AP, any Q < 2^54: three or four digits with an exact top digit (elimination without wrap repair), or
// two digits on the plain path (STD128Q: the top digit can wrap)
bool dm64w_supported(const tfhe_b200_params& p) {
    if (p.method != TFHE_B200_METHOD_AP || p.N != 2048 || p.numDigitsToThrow != 0)
        return false;
    if (p.Q >= (1ULL << 54) || p.q == 0 || p.q > 4096 || p.baseR < 2 || p.digitsR == 0)
        return false;
    if (p.digitsG == 2)
        return true;   // plain or eliminated, as cggi32_skip_top_ok decides
    if (p.digitsG != 3 && p.digitsG != 4)
        return false;
    return cggi32_skip_top_ok(p);
}

cudaError_t launch_br_dm64w(const BRCommon& c, const CGGI64WTables& t, cudaStream_t s, int sm_count, int group) {
    DM64WArgs a;
    a.c = c;
    a.mod = t.mod;
    a.bk = t.bk;
    a.twC = t.twC;
    a.twB = t.twB;
    a.twU = t.twU;
    a.Q2 = 2 * t.mod.Q;
    const u64 B = 1ULL << c.gBits;
    unsigned __int128 off = 0, pw = 1;
    for (u32 i = 0; i < c.digitsKept; i++) {
        off += (B / 2) * pw;
        pw *= B;
    }
    a.dig_off = (u64)off;
    a.dig_add = t.mod.Q - B / 2;
    a.zero64 = 0;
    a.ninvM = to_mont<u64>(h_powmod((u64)w64::N, t.mod.Q - 2, t.mod.Q), t.mod);
    const bool one = group == 1 || (group == 0 && sm_count > 0 && c.batch <= sm_count);
    if (c.digitsKept == 2) {
        if (!t.plain)
            return one ? launch_dm_w<2, 1>(a, s) : launch_dm_w<2, 2>(a, s);
        return one ? launch_dm_w<2, 1, true>(a, s) : launch_dm_w<2, 2, true>(a, s);
    }
    if (t.plain)
        return cudaErrorInvalidConfiguration;
    if (c.digitsKept == 3)
        return one ? launch_dm_w<3, 1>(a, s) : launch_dm_w<3, 2>(a, s);
    if (c.digitsKept == 4)
        return launch_dm_w<4, 1>(a, s);
    return cudaErrorInvalidConfiguration;
}

}  // namespace tfhe_b200
