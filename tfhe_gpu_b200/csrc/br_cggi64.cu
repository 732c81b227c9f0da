// Specialised CGGI/GINX blind rotation for the 54-bit functional-bootstrapping parameter sets (N = 2048,
// 2^31 <= Q < 2^55): EvalFunc / EvalFloor / EvalSign / EvalDecomp (BASELINE.json configs[3], [4]).
//
// Same design as br_cggi32.cu, in 64-bit words:
//   * a CTA owns G ciphertexts that walk the n steps in lock-step and share the RGSW key words of the step;
//   * each accumulator polynomial lives in coefficient form in registers: 64 threads x 32 coefficients (u64);
//   * register-resident three-part NTT: pass A (strides 1024..64, twiddles identical in every thread -> kernel-parameter
//     constant bank), one swizzled shared-memory transpose, the stride-32 stage done between neighbouring lanes with
//     warp shuffles (each lane takes half of the butterflies), pass B (strides 16..1, per-thread twiddles read from a
//     shared-memory table);
//   * lazy Harvey/Shoup butterflies on 64-bit words with an approximate quotient (shoup64; values < 45 Q < 2^60, one
//     conditional subtraction at the end of the forward transform), inverse transform through the mirrored
//     block so the forward twiddle table serves both directions;
//   * pointwise stage accumulates in 128 bits with one Montgomery reduction per output;
//   * closed-form signed digits; top-digit elimination (logQ = 12: 2 forward + 2 inverse transforms per step instead of
//     4 + 2; logQ = 17: 4 + 2 instead of 6 + 2).  For these rings the reference's top digit CAN wrap (B^d/2 - B/2 < Q/2),
//     so the kernel records the wrapped coefficients of a step and corrects the evaluation-domain accumulator for them
//     (cggi_skip_top_wrapfix_ok, wrap_fix below) -- still bit-exact;
//   * fused accumulator init / extraction / transpose.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ntt64.cuh"

namespace tfhe_b200 {

namespace {

constexpr int LOGN = 11, N = 2048, TPN = 64, NTW = 31;

struct CGGI64Args {
    BRCommon c;
    ModCtx<u64> mod;
    const u64* bk;       // [i][x(2D planes)][slot][2]: word w = (key*D + l')*2 + j lives in plane w/2, lane w%2;
                         // each word split into 27-bit limbs b0 | b1 << 32 (Montgomery form)
    const u64* psi_pow;  // [2N] Montgomery form (global memory)
    const u64* twB;      // [NTW][TPN][2] per-thread pass-B twiddles (value, Shoup companion), already [x][T] order
    const u64* tw32;     // [32][2] stride-32 stage twiddles: psi^bitrev(32 + u) and companion
    const u64* twU;      // [2][31][2] uniform pass-A twiddles: forward W[e+1], then NEGATED inverse Q - WI[e+1]
    u64 Q2, dig_off, dig_add, ninvM;
    u64 zero64;          // runtime 0 (see shoup64)
    u64 kfix;            // B^d / N mod Q (plain): weight of a wrapped top digit in the evaluation-domain accumulator
    u32 zero;
};

// u64 position -> physical u64 index inside a region: XOR the 16-byte chunk index with the owner block's low bits
__device__ __forceinline__ u32 pos64(u32 p) {
    return p ^ (((p >> 5) & 7) << 1);
}
__device__ __forceinline__ u64 shfl_xor64(u64 v) {
    u32 lo = (u32)v, hi = (u32)(v >> 32);
    lo = __shfl_xor_sync(0xffffffffu, lo, 1);
    hi = __shfl_xor_sync(0xffffffffu, hi, 1);
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ void load_B(u64 (&v)[32], const u64* reg, int blk) {
    const ulonglong2* p = reinterpret_cast<const ulonglong2*>(reg) + 16 * blk;
#pragma unroll
    for (int x = 0; x < 16; x++) {
        ulonglong2 w = p[x ^ (blk & 7)];
        v[2 * x] = w.x;
        v[2 * x + 1] = w.y;
    }
}
__device__ __forceinline__ void store_B(const u64 (&v)[32], u64* reg, int blk) {
    ulonglong2* p = reinterpret_cast<ulonglong2*>(reg) + 16 * blk;
#pragma unroll
    for (int x = 0; x < 16; x++)
        p[x ^ (blk & 7)] = make_ulonglong2(v[2 * x], v[2 * x + 1]);
}

// ---- five in-register stages on v[32] with twiddles (value, companion) read from a shared-memory table ------------
// One body serves pass A (uniform twiddles: stride 0 table twU) and pass B (per-thread twiddles: twS[x][T]); it is
// executed from a non-unrolled loop so the code exists once (the fully unrolled 64-bit transforms did not fit the
// instruction cache: ncu showed `no_instruction` as the top stall).  Table entry e of a pass = twiddle (off(s) + x).
__device__ __forceinline__ void fwd_pass5(u64 (&v)[32], const ulonglong2* __restrict__ tab, int stride, int lane_off,
                                          u64 nQ, u64 QO, u64 Z) {
#pragma unroll
    for (int s = 4; s >= 0; s--) {
        const int off = (32 >> (s + 1)) - 1, cnt = 32 >> (s + 1);
        ulonglong2 w[16];
#pragma unroll
        for (int x = 0; x < cnt; x++)
            w[x] = tab[(off + x) * stride + lane_off];
#pragma unroll
        for (int r = 0; r < 32; r++) {
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
// Gentleman-Sande stages 2^s, s = 0..4, in the form v' = (V - U) * w: pass B' reads the FORWARD per-thread table
// mirrored (psi^-bitrev(m+i) = -psi^bitrev(m + m-1-i)), pass A' reads a table of NEGATED inverse uniform twiddles.
__device__ __forceinline__ void inv_pass5(u64 (&v)[32], const ulonglong2* __restrict__ tab, int stride, int lane_off,
                                          bool mirror, u64 nQ, u64 QO, u64 Z) {
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const int off = (32 >> (s + 1)) - 1, cnt = 32 >> (s + 1);
        ulonglong2 w[16];
#pragma unroll
        for (int x = 0; x < cnt; x++)
            w[x] = tab[(off + (mirror ? cnt - 1 - x : x)) * stride + lane_off];
#pragma unroll
        for (int r = 0; r < 32; r++) {
            if (r & (1 << s))
                continue;
            const int ti = r >> (s + 1);
            u64 U = v[r], V = v[r + (1 << s)];
            v[r] = csub(U + V + Z, QO);
            v[r + (1 << s)] = shoup64(V - U + QO, w[ti].x, w[ti].y, nQ, Z);
        }
    }
}
// stride 32: positions p (block 2u) and p + 32 (block 2u + 1) are held by neighbouring lanes.  Each lane computes half
// of the 32 butterflies: the even lane those of its local slots 0..15, the odd lane those of its local slots 16..31.
template <bool INV>
__device__ __forceinline__ void stage32(u64 (&v)[32], bool odd_blk, u64 w, u64 wp, u64 nQ, u64 QO, u64 Z) {
    // odd_blk: this lane holds the upper block (the "y" side) of the pair
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const u64 lo = v[k], hi = v[16 + k];
        const u64 send = odd_blk ? lo : hi;
        const u64 own = odd_blk ? hi : lo;
        const u64 recv = shfl_xor64(send);
        const u64 x = odd_blk ? recv : own;
        const u64 y = odd_blk ? own : recv;
        u64 nx, ny;
        if (!INV) {
            u64 t = shoup64(y, w, wp, nQ, Z);
            nx = x + t + Z;
            ny = x - t + QO;
        }
        else {
            // Gentleman-Sande with the mirrored (negated) forward twiddle: y' = (x - y) * (-w) = (y - x) * w
            nx = csub(x + y + Z, QO);
            ny = shoup64(y - x + QO, w, wp, nQ, Z);
        }
        const u64 keep = odd_blk ? ny : nx;
        const u64 ret = odd_blk ? nx : ny;
        const u64 recv2 = shfl_xor64(ret);
        v[k] = odd_blk ? recv2 : keep;
        v[16 + k] = odd_blk ? keep : recv2;
    }
}
template <int DK, int G>
struct K64 {
    static constexpr int D = 2 * DK;
    static constexpr int NT = G * 2 * TPN;
    static constexpr int WB = N / 32;   // words of a wrapped-coefficient bitmap
    static constexpr size_t smem = (size_t)G * D * N * 8 + (size_t)NTW * TPN * 16 + 2 * NTW * 16 + 64 +
                                   2 * (size_t)G * 2 * (WB + 2) * 4;
};

template <int DK, int G, bool SKIP>
__global__ void __launch_bounds__(K64<DK, G>::NT, 1) br_cggi64_kernel(const __grid_constant__ CGGI64Args A) {
    using K = K64<DK, G>;
    constexpr int D = K::D, NT = K::NT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* Dsm = reinterpret_cast<u64*>(smem_raw);                                  // [G][D][N]
    ulonglong2* twS = reinterpret_cast<ulonglong2*>(Dsm + (size_t)G * D * N);    // [NTW][TPN]
    ulonglong2* twUf = twS + NTW * TPN;                                           // [NTW] uniform forward
    ulonglong2* twUi = twUf + NTW;                                                // [NTW] uniform inverse, negated
    // wrapped-top-digit bookkeeping (SKIP only), double-buffered by step parity: bitmaps [2][G][2][WB], warp flags
    // [2][G][2][2]
    u32* wbits = reinterpret_cast<u32*>(twUi + NTW);
    u32* wany = wbits + 2 * G * 2 * K::WB;

    const BRCommon& C = A.c;
    const u64 Q = A.mod.Q, Q2 = A.Q2, QO = 2 * A.Q2, nQ = 0 - A.mod.Q, qinv = A.mod.qinv, oneM = A.mod.oneM;
    const u64 Z = A.zero64;
    const u32 n = C.n;
    const int tid = threadIdx.x;
    const int g = tid / (2 * TPN), j = (tid / TPN) & 1, T = tid % TPN;
    const int bar_id = 1 + g * 2 + j;
    const int ct = blockIdx.x * G + g;
    const bool live = ct < C.batch;
    const u64* lwe = C.ct + (size_t)(live ? ct : 0) * (n + 1);

    for (int x = tid; x < NTW * TPN; x += NT)
        twS[x] = reinterpret_cast<const ulonglong2*>(A.twB)[x];
    for (int x = tid; x < 2 * NTW; x += NT)
        twUf[x] = reinterpret_cast<const ulonglong2*>(A.twU)[x];
    const u64 w32 = A.tw32[(T >> 1) * 2], w32p = A.tw32[(T >> 1) * 2 + 1];

    // ---- accumulator initialisation in A layout (coefficient idx = T + 64 r) ------------------------------------
    u64 c[32];
    if (C.acc_init == ACC_EXPLICIT) {
        const u64* src = C.acc_io + ((size_t)(live ? ct : 0) * 2 + j) * N;
#pragma unroll
        for (int r = 0; r < 32; r++)
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
        for (int r = 0; r < 32; r++) {
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
    const u64 QHalf = Q >> 1;
    const u32 gBits = C.gBits;
    const u64 gmask = ((u64)1 << gBits) - 1;
    const bool odd_blk_f = T & 1;          // forward: lane holds block T
    const bool odd_blk_i = !(T & 1);       // inverse: lane holds mirrored block 63 - T

    // forward transform of v (A layout in) through region `reg`, result left in registers in B layout of block T
    auto forward = [&](u64 (&v)[32], u64* reg) {
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
            fwd_pass5(v, pass ? twS : twUf, pass ? TPN : 1, pass ? T : 0, nQ, QO, Z);
            if (pass == 0) {
#pragma unroll
                for (int r = 0; r < 32; r++)
                    reg[pos64(T + TPN * r)] = v[r];
                group_sync(bar_id);
                load_B(v, reg, T);
                group_sync(bar_id);
                stage32<false>(v, odd_blk_f, w32, w32p, nQ, QO, Z);
            }
        }
        // 11 lazy stages leave values below Q + 11 * 4Q = 45 Q; the pointwise stage's 27-bit limb split needs < 31 Q
        const u64 Q16 = 4 * QO;
#pragma unroll
        for (int r = 0; r < 32; r++)
            v[r] = csub(v[r], Q16);   // < max(16 Q, 29 Q)
    };

    if (SKIP) {
        // evaluation-domain accumulator (scaled by N^-1), see br_cggi32.cu
        u64 v[32];
#pragma unroll
        for (int r = 0; r < 32; r++)
            v[r] = c[r];
        u64* reg = myD + (size_t)(j + 2 * (DK - 1)) * N;
        forward(v, reg);
#pragma unroll
        for (int r = 0; r < 32; r++)
            v[r] = A.mod.mont_mul(v[r], A.ninvM);   // lazy (< 29 Q) * (N^-1 R) * R^-1 -> canonical
        store_B(v, reg, T);
        __syncthreads();
    }

    for (u32 i = 0; i < n; i++) {
        // ---- phase 1 ---------------------------------------------------------------------------------------------
        if (SKIP) {
            // Top-digit elimination assumes c = sum_l d_l B^l.  The reference's top digit is truncated to its gBits
            // window, so for the few centred values above B^d/2 - B/2 it comes out B too small and
            // c = sum_l d_l B^l + B^d.  Those coefficients are recorded here (bit DK*gBits of the offset value) and the
            // pointwise stage corrects the evaluation-domain accumulator for them (wrap_fix below).
            u32 wm = 0;
            const u32 wsh = gBits * DK;
#pragma unroll
            for (int r = 0; r < 32; r++) {
                i64 dv = (c[r] < QHalf) ? (i64)c[r] : (i64)c[r] - (i64)Q;
                u64 Dv = (u64)(dv + (i64)A.dig_off);
                wm |= (u32)((Dv >> wsh) & 1) << r;
            }
            const bool anyw = __any_sync(0xffffffffu, wm != 0);
            const int par = i & 1;
            if (anyw) {
                // coefficient index T + 64 r = 32 * (2 r + T / 32) + T % 32: a warp's ballot IS the bitmap word
                u32* wb = wbits + ((size_t)(par * G + g) * 2 + j) * K::WB + (T >> 5);
#pragma unroll
                for (int r = 0; r < 32; r++) {
                    u32 word = __ballot_sync(0xffffffffu, (wm >> r) & 1);
                    if ((T & 31) == 0)
                        wb[2 * r] = word;
                }
            }
            if ((T & 31) == 0)
                wany[((par * G + g) * 2 + j) * 2 + (T >> 5)] = anyw;
        }
        // With top-digit elimination the accumulator exists in evaluation form (top rows), and phase 3 inverse-transforms
        // THAT (not the step's delta), so the coefficient-form copy c[] is only needed here, for the digits: it does not
        // stay in registers through the transforms.  With more than one digit it is parked in row j (the scratch of the
        // LAST digit processed, l = 0) and re-read per digit; every thread reads back only what it wrote itself.
        constexpr int NF = SKIP ? DK - 1 : DK;
        if (SKIP && NF > 1) {
            u64* park = myD + (size_t)j * N;
#pragma unroll
            for (int r = 0; r < 32; r++)
                park[pos64(T + TPN * r)] = c[r];
        }
#pragma unroll 1
        for (int li = 0; li < NF; li++) {
            const int l = SKIP ? NF - 1 - li : li;
            if (SKIP && NF > 1) {
                const u64* park = myD + (size_t)j * N;
#pragma unroll
                for (int r = 0; r < 32; r++)
                    c[r] = park[pos64(T + TPN * r)];
            }
            u64 v[32];
            const u32 sh = gBits * (l + C.numThrow);
#pragma unroll
            for (int r = 0; r < 32; r++) {
                i64 dv = (c[r] < QHalf) ? (i64)c[r] : (i64)c[r] - (i64)Q;
                u64 Dv = (u64)(dv + (i64)A.dig_off);
                v[r] = ((u64)((i64)Dv >> sh) & gmask) + A.dig_add;
            }
            u64* reg = myD + (size_t)(j + 2 * l) * N;
            forward(v, reg);
            store_B(v, reg, T);
        }
        __syncthreads();

        // ---- phase 2: pointwise stage -------------------------------------------------------------------------------
        // rare path (about 1 ciphertext-step in 130 for logQ = 17, 1 in 65000 for logQ = 12): for every wrapped
        // coefficient k0 of ciphertext gg / component jj, NTT(X^k0)[slot] = psi^((2 bitrev(slot) + 1) k0), so the
        // top-row operand becomes acc_eval - (B^d / N) * sum_k0 psi^(...).  Applied in place before the stage and
        // undone after it (the stage adds delta on top), all CTA-uniform.
        bool anyflag = false;
        if (SKIP) {
            const u32* fl = wany + (size_t)(i & 1) * G * 4;
            u32 f = 0;
#pragma unroll
            for (int x = 0; x < G * 4; x++)
                f |= fl[x];
            anyflag = f != 0;
        }
        auto wrap_fix = [&](bool undo) {
            const u32* fl = wany + (size_t)(i & 1) * G * 4;
#pragma unroll 1
            for (int gj = 0; gj < G * 2; gj++) {
                if (!(fl[gj * 2] | fl[gj * 2 + 1]))
                    continue;
                const u32* wb = wbits + ((size_t)(i & 1) * G * 2 + gj) * K::WB;
                u64* reg = Dsm + (size_t)(gj >> 1) * D * N + (size_t)(2 * (DK - 1) + (gj & 1)) * N;
#pragma unroll 1
                for (int k = tid; k < N; k += NT) {
                    const u32 br = __brev((u32)k) >> (32 - LOGN);
                    u64 sum = 0;
#pragma unroll 1
                    for (int wd = 0; wd < K::WB; wd++) {
                        if (!fl[gj * 2 + (wd & 1)])
                            continue;
                        u32 bits = wb[wd];
                        while (bits) {
                            const u32 k0 = 32 * wd + (__ffs(bits) - 1);
                            bits &= bits - 1;
                            sum = csub(sum + __ldg(A.psi_pow + (((2 * br + 1) * k0) & (2 * N - 1))), Q);
                        }
                    }
                    const u64 term = A.mod.mont_mul(sum, A.kfix);
                    const u64 x = reg[pos64(k)];
                    reg[pos64(k)] = undo ? csub(x + term, Q) : (x >= term ? x - term : x + Q - term);
                }
            }
        };
        if (SKIP && anyflag) {
            wrap_fix(false);
            __syncthreads();
        }
        {
            constexpr int ITERS = N / NT;
            static_assert(N % NT == 0, "unsupported CTA shape");
            constexpr int PL = 2 * D;   // uint4-sized planes per slot (4D u64 words)
            const ulonglong2* bki = reinterpret_cast<const ulonglong2*>(A.bk) + (size_t)i * PL * N;
            u64 ee[G];
#pragma unroll
            for (int gg = 0; gg < G; gg++) {
                // rgsw-acc-cggi.cpp:146-153: e_i = ((mod - a_i) mod mod) * (2N / mod); 0 for dead slots
                const int cg = blockIdx.x * G + gg;
                u64 e = 0;
                if (cg < C.batch) {
                    u64 ai = C.ct[(size_t)cg * (n + 1) + i] % C.ct_mod;
                    e = ((C.ct_mod - ai) % C.ct_mod) * ((2 * N) / C.ct_mod);
                }
                ee[gg] = e;
            }
#pragma unroll 1
            for (int it = 0; it < ITERS; it++) {
                const int k = tid + it * NT;
                u64 bkv[4 * D];
#pragma unroll
                for (int x = 0; x < PL; x++) {
                    ulonglong2 w = bki[(size_t)x * N + k];
                    bkv[2 * x] = w.x;
                    bkv[2 * x + 1] = w.y;
                }
                const u32 pk = pos64(k);
                const u32 br = __brev((u32)k) >> (32 - LOGN);
                // all shared-memory loads of the CTA's ciphertexts first, then the arithmetic (independent streams), then
                // the stores: written per ciphertext, the store of one and the loads of the next cannot be reordered
                u64 xd[G][D], m1[G], m2[G], dl0[G], dl1[G];
#pragma unroll
                for (int gg = 0; gg < G; gg++) {
                    const u64* dreg = Dsm + (size_t)gg * D * N + pk;
#pragma unroll
                    for (int l = 0; l < D; l++)
                        xd[gg][l] = dreg[(size_t)l * N];
                    const u32 xx = (u32)(((2 * br + 1) * ee[gg]) & (2 * N - 1));
                    m1[gg] = __ldg(A.psi_pow + xx);
                    m2[gg] = __ldg(A.psi_pow + ((2 * N - xx) & (2 * N - 1)));
                }
#pragma unroll
                for (int gg = 0; gg < G; gg++) {
                    const Limb x0(xd[gg][0]);
                    L3 a00(x0, bkv[(0 * D) * 2 + 0]), a01(x0, bkv[(0 * D) * 2 + 1]);
                    L3 a10(x0, bkv[(1 * D) * 2 + 0]), a11(x0, bkv[(1 * D) * 2 + 1]);
#pragma unroll
                    for (int l = 1; l < D; l++) {
                        const Limb x(xd[gg][l]);
                        a00.mac(x, bkv[(0 * D + l) * 2 + 0]);
                        a01.mac(x, bkv[(0 * D + l) * 2 + 1]);
                        a10.mac(x, bkv[(1 * D + l) * 2 + 0]);
                        a11.mac(x, bkv[(1 * D + l) * 2 + 1]);
                    }
                    const Limb s00(redc128(a00.value(), Q, qinv)), s01(redc128(a01.value(), Q, qinv));
                    const Limb s10(redc128(a10.value(), Q, qinv)), s11(redc128(a11.value(), Q, qinv));
                    u64 f1 = m1[gg], f2 = m2[gg];
                    f1 = f1 >= oneM ? f1 - oneM : f1 + Q - oneM;
                    f2 = f2 >= oneM ? f2 - oneM : f2 + Q - oneM;
                    const Limb F1(f1), F2(f2);
                    L3 t0(s00, F1), t1(s01, F1);
                    t0.mac(s10, F2);
                    t1.mac(s11, F2);
                    dl0[gg] = redc128(t0.value(), Q, qinv);
                    dl1[gg] = redc128(t1.value(), Q, qinv);
                    if (SKIP) {
                        m1[gg] = csub(xd[gg][2 * (DK - 1)] + dl0[gg], Q);
                        m2[gg] = csub(xd[gg][2 * (DK - 1) + 1] + dl1[gg], Q);
                    }
                }
#pragma unroll
                for (int gg = 0; gg < G; gg++) {
                    u64* dreg = Dsm + (size_t)gg * D * N + pk;
                    if (!SKIP) {
                        dreg[0] = dl0[gg];
                        dreg[N] = dl1[gg];
                    }
                    else {
                        dreg[(size_t)(2 * (DK - 1)) * N] = m1[gg];
                        dreg[(size_t)(2 * (DK - 1) + 1) * N] = m2[gg];
                    }
                }
            }
        }
        __syncthreads();
        if (SKIP && anyflag) {
            wrap_fix(true);
            __syncthreads();   // phase 3 reads the top rows
        }

        // ---- phase 3: inverse transform of delta_j through the mirrored block, accumulate -------------------------
        // (SKIP: inverse transform of the whole evaluation-domain accumulator, read from the top row, through row j)
        {
            u64 v[32];
            u64* reg = myD + (size_t)j * N;
            const int Tv = TPN - 1 - T;
            load_B(v, SKIP ? myD + (size_t)(j + 2 * (DK - 1)) * N : reg, Tv);
#pragma unroll 1
            for (int pass = 0; pass < 2; pass++) {
                inv_pass5(v, pass ? twUi : twS, pass ? 1 : TPN, pass ? 0 : T, pass == 0, nQ, QO, Z);
                if (pass == 0) {
                    stage32<true>(v, odd_blk_i, w32, w32p, nQ, QO, Z);
                    group_sync(bar_id);
                    store_B(v, reg, Tv);
                    group_sync(bar_id);
#pragma unroll
                    for (int r = 0; r < 32; r++)
                        v[r] = reg[pos64(T + TPN * r)];
                    group_sync(bar_id);
                }
            }
#pragma unroll
            for (int r = 0; r < 32; r++)
                c[r] = SKIP ? csub(csub(v[r], Q2), Q) : csub(csub(csub(c[r] + v[r], QO), Q2), Q);   // v < 4Q
        }
    }

    if (live) {
        if (C.write_acc) {
            u64* dst = C.acc_io + (size_t)ct * 2 * N;
#pragma unroll
            for (int r = 0; r < 32; r++) {
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
            for (int r = 0; r < 32; r++) {
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

u32 bitrev_h(u32 x, u32 bits) {
    u32 r = 0;
    for (u32 i = 0; i < bits; i++) {
        r = (r << 1) | (x & 1);
        x >>= 1;
    }
    return r;
}
u64 shoup_h(u64 w, u64 Q) {
    return (u64)((((unsigned __int128)w) << 64) / Q);
}

template <int DK, int G, bool SKIP>
cudaError_t launch_t(const CGGI64Args& a, cudaStream_t s) {
    using K = K64<DK, G>;
    if (K::smem > 227 * 1024)
        return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(br_cggi64_kernel<DK, G, SKIP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)K::smem);
    if (e != cudaSuccess)
        return e;
    const int grid = (a.c.batch + G - 1) / G;
    br_cggi64_kernel<DK, G, SKIP><<<grid, K::NT, K::smem, s>>>(a);
    return cudaGetLastError();
}

}  // namespace

bool cggi64_supported(const tfhe_b200_params& p) {
    if (p.method != TFHE_B200_METHOD_GINX || p.N != 2048)
        return false;
    // Q < 2^54: lazy transform outputs (< 29 Q) must split into 27-bit limbs with x1 + x0 < 2^32 (see L3).  There is no
    // lower bound: the 27- and 29-bit moduli of the N = 2048 sets (STD256 / STD256Q and their _OPT variants,
    // binfhecontext.cpp:147-148,153-154) run here as well, in 64-bit words (tfhe_b200_setup selects the 64-bit engine
    // for them), which beats the generic 32-bit kernel they would otherwise fall back to.
    if (p.Q >= (1ULL << 54))
        return false;
    const u32 dk = p.digitsG - p.numDigitsToThrow;
    return dk >= 1 && dk <= 4;
}

// twB: [NTW][TPN][2] (x-major, so the kernel copies it straight to shared memory); tw32: [32][2]; twA: [fwd|inv][32][2]
void cggi64_build_tables(const tfhe_b200_params& p, std::vector<u64>& twA, std::vector<u64>& twB, std::vector<u64>& tw32) {
    const u64 Q = p.Q;
    std::vector<u64> W(N), WI(N);
    u64 psi = p.psi % Q, psii = h_powmod(psi, Q - 2, Q), x = 1, xi = 1;
    for (u32 k = 0; k < (u32)N; k++) {
        u32 r = bitrev_h(k, LOGN);
        W[r] = x;
        WI[r] = xi;
        x = h_mulmod(x, psi, Q);
        xi = h_mulmod(xi, psii, Q);
    }
    // uniform tables, entry e <-> twiddle index e + 1 = (16 >> s) + x: forward W, then NEGATED inverse (pass A' uses
    // the (V - U) * w form)
    twA.assign(2 * NTW * 2, 0);
    for (u32 e = 0; e < (u32)NTW; e++) {
        twA[(0 * NTW + e) * 2 + 0] = W[e + 1];
        twA[(0 * NTW + e) * 2 + 1] = shoup_h(W[e + 1], Q);
        u64 neg = (Q - WI[e + 1]) % Q;
        twA[(1 * NTW + e) * 2 + 0] = neg;
        twA[(1 * NTW + e) * 2 + 1] = shoup_h(neg, Q);
    }
    tw32.assign(32 * 2, 0);
    for (u32 u = 0; u < 32; u++) {
        tw32[u * 2] = W[32 + u];
        tw32[u * 2 + 1] = shoup_h(W[32 + u], Q);
    }
    twB.assign((size_t)NTW * TPN * 2, 0);
    for (int s = 4; s >= 0; s--) {
        const u32 cnt = 32 >> (s + 1), off = cnt - 1;
        for (u32 T = 0; T < (u32)TPN; T++)
            for (u32 xx = 0; xx < cnt; xx++) {
                u64 w = W[(N >> (s + 1)) + T * cnt + xx];
                twB[((size_t)(off + xx) * TPN + T) * 2 + 0] = w;
                twB[((size_t)(off + xx) * TPN + T) * 2 + 1] = shoup_h(w, Q);
            }
    }
}

cudaError_t launch_br_cggi64(const BRCommon& c, const CGGI64Tables& t, cudaStream_t s, int group) {
    CGGI64Args a;
    a.c = c;
    a.mod = t.mod;
    a.bk = t.bk;
    a.psi_pow = t.psi_pow;
    a.twB = t.twB;
    a.tw32 = t.tw32;
    a.twU = t.twA;
    a.Q2 = 2 * t.mod.Q;
    const u64 B = 1ULL << c.gBits;
    const u32 total_digits = c.digitsKept + c.numThrow;
    unsigned __int128 off = 0, pw = 1;
    for (u32 i = 0; i < total_digits; i++) {
        off += (B / 2) * pw;
        pw *= B;
    }
    a.dig_off = (u64)off;
    a.dig_add = t.mod.Q - B / 2;
    a.zero = 0;
    a.zero64 = 0;
    a.ninvM = to_mont<u64>(h_powmod((u64)N, t.mod.Q - 2, t.mod.Q), t.mod);
    a.kfix = h_mulmod((u64)(pw % t.mod.Q), h_powmod((u64)N, t.mod.Q - 2, t.mod.Q), t.mod.Q);
    const int dk = (int)c.digitsKept;
    (void)group;
    if (dk == 1)
        return launch_t<1, 2, false>(a, s);
    if (dk == 2)
        return t.skip_top ? launch_t<2, 2, true>(a, s) : launch_t<2, 2, false>(a, s);
    if (dk == 3)
        return t.skip_top ? launch_t<3, 2, true>(a, s) : launch_t<3, 2, false>(a, s);
    if (dk == 4)
        return t.skip_top ? launch_t<4, 1, true>(a, s) : launch_t<4, 1, false>(a, s);
    return cudaErrorInvalidConfiguration;
}

}  // namespace tfhe_b200
