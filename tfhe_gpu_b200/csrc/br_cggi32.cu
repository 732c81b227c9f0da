// Specialised CGGI/GINX blind rotation for 32-bit moduli (Q < 2^32/22, i.e. the 27-bit primes of TOY / STD128 /
// logQ=11), N in {512, 1024}.  This is the kernel the headline metric (STD128 CGGI NAND gates/s) runs on.
//
// Design (B200-first, nothing like the reference's one-CTA-per-ciphertext FP64 FFT kernel):
//   * a CTA owns a GROUP of G ciphertexts and walks the n rotation steps in lock-step, so the RGSW key of step i
//     (128 KB for STD128) is read from L2 once per CTA and reused from registers by all G ciphertexts;
//   * the accumulator never leaves the register file: each (ciphertext, component) polynomial is held in
//     coefficient form by N/32 threads, 32 coefficients per thread ("A layout": thread T owns T + (N/32) r);
//   * every NTT is a register-resident radix-32 x radix-(N/32) two-pass transform done by N/32 threads (one warp
//     for N=1024): pass A (strides >= N/32) uses 31 twiddles that are the same for every thread and are read as
//     constant-bank operands straight from the kernel parameters; one shared-memory transpose; pass B (strides
//     < 32) uses 31 per-thread twiddles that stay in registers for the whole kernel.  Butterflies are lazy
//     (Harvey) with Shoup multiplication: 1 IMAD.HI + 2 IMAD + 2 IADD3, no conditional corrections in the forward
//     transform (values stay < 22 Q < 2^32);
//   * the inverse transform re-uses the forward per-thread twiddles through the identity
//     psi^-bitrev(m+i) = -psi^bitrev(m + (m-1-i)): thread T processes the mirrored block N/32-1-T;
//   * signed digit decomposition is computed in closed form per digit ((d + offset) >> g*l) & (B-1)) - B/2, which
//     is bit-identical to the reference's sequential carry loop (rgsw-acc.cpp:86-108);
//   * the pointwise stage accumulates the d products per output in 64 bits (IMAD.WIDE) and performs ONE Montgomery
//     reduction per output, then applies the monomial factors (psi-power table in shared memory);
//   * N^-1 and the Montgomery factor are folded into the key; accumulator init, the LWE-mask modulus switch,
//     sample extraction and the a(X^-1) transpose are fused in.
#include <cstring>

#include "ntt32.cuh"

// Words per entry of the monomial-factor table in shared memory (phase 2):
//   1: psi^x (Montgomery form); the two factors psi^x - 1 and psi^-x - 1 are two lookups plus a modular subtraction each
//      (kept for the TMA variant, whose shared-memory ring is laid out behind the 8 KB table)
//   2: (psi^x - 1, psi^-x - 1): one index computation, one 64-bit lookup, no subtraction -- 160.8 -> 156.0 ms per 16384
//      STD128 bootstraps.  (A four-word entry (a R, a) that folds the factors into the 64-bit sums as hi(s) (a R) + lo(s) a
//      with one reduction per output saves four more multiplier slots per slot but measured 157.2 ms: not kept.)
#ifndef CGGI32_TABW
#define CGGI32_TABW 2
#endif

namespace tfhe_b200 {

struct CGGI32Args {
    BRCommon c;
    ModCtx<u32> mod;
    const u32* bk;       // [i][x][k][4]: word w = (key*D + l')*2 + jout of slot k lives in plane x = w/4, lane w%4
    const u32* psi_pow;  // [2N] Montgomery form
    const u32* twB;      // [TPN][NTW][2] per-thread pass-B twiddles (value, Shoup companion)
    u32 twA_f[32][2];    // uniform pass-A twiddles, forward: index (16>>s) + (r>>(s+1))
    u32 twA_i[32][2];    // inverse
    u32 Q2;              // 2Q
    u32 dig_off;         // closed-form decomposition offset  sum_i (B/2) B^i
    u32 dig_add;         // Q - B/2 (digits are fed to the lazy NTT as r + Q)
    u32 ninvM;           // N^-1 in Montgomery form (SKIP: the evaluation-domain accumulator is kept scaled by N^-1)
    u32 zero;            // always 0: third IADD3 operand that keeps ptxas from turning adds into IMAD.IADD (fma-heavy pipe)
    // persistent variant (PERS): hand-over slots of the groups that are split between two CTAs, see the kernel header
    u32* pers_state;     // [slot][2][G][2][N] accumulator image: coefficient registers, evaluation-domain rows (SKIP path)
    u32* pers_flags;     // [slot] launch epoch once the slot's image is complete
    u32 pers_epoch;
    u32 pers_groups;     // ceil(batch / G)
    u32* pers_ticket;    // running count of persistent CTAs started on this device (never reset)
    u32 pers_ticket_base;   // ... its value when this launch starts
};

// ---- TMA bulk copies + mbarriers (key streaming of the TMA variant) ----------------------------------------------
__device__ __forceinline__ u32 smem_u32(const void* p) {
    return (u32)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u32 bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(u32 dst, const void* src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// TMA = true (opt-in variant, see launch_t): the RGSW key words of a pointwise iteration (D planes x NT slots x 16 B =
// 32 KB for the headline shape) are streamed by TMA bulk copies into a two-stage shared-memory ring, two iterations
// ahead (full / empty mbarriers, one producer thread), instead of register-staged __ldg prefetches.
//
// LAT = true (latency layout, one ciphertext per CTA, top digit eliminated): 2*DK warps; warps 0 .. 2*(DK-1)-1 own one
// DIGIT polynomial each instead of one accumulator component.  The DK-1 warps of a component each run the inverse
// transform redundantly (through their own digit region), extract their own digit and do ONE forward transform; the
// last pair of warps only helps in the pointwise stage.  The critical path of a rotation step is 1 inverse + 1 forward
// transform + N / (64 DK) pointwise iterations instead of 1 + (DK-1) + N/64.  Chosen for batches of at most one
// ciphertext per SM.
//
// SWEEP = 1 (28-bit moduli): one reduction sweep between the two passes of every forward transform, see
// sweep_below_2q in ntt32.cuh; everything else is unchanged (the pointwise bounds hold for Q < 2^28: lazy rows < 12 Q,
// eight of them times a key word < Q stay below 2^63, and the two reduced sums times the monomial factors below Q 2^32).
// SWEEP = 2 (29-bit moduli, N = 2048 only): 8 Q is all that fits 32 bits, so a sweep follows every third lazy stage and
// the transform ends below 2 Q (rows < 2 Q times key words < Q: eight of them stay below 2^63).
//
// LOGN = 11 (N = 2048: the STD256 family, binfhecontext.cpp:147-148,153-154): 64 threads x 32 coefficients per
// polynomial, i.e. two warps per (ciphertext, component); the transposes are fenced by a 64-thread named barrier and one
// cross-lane stage sits between the two in-thread passes (cross_stage in ntt32.cuh).
//
// PERS = true (persistent variant, no wave quantisation): CTAs walk the n steps of a group in lock-step, so a plain
// launch costs ceil(groups / SMs) wave times whatever the remainder.  Here the grid is ONE CTA per SM and the
// groups * n rotation steps of the launch are cut into gridDim.x equal contiguous ranges (McNaughton's wrap-around rule
// for preemptive scheduling): a CTA's range is the last steps of group gA, whole groups, and the first sB steps of group
// gB.  It runs the head of gB FIRST (and leaves the accumulator image in its hand-over slot), then its whole groups,
// then the rest of gA LAST, whose head the previous CTA ran at the very start of the launch -- a range is at least n
// steps long, so that image has been waiting since long before it is needed (the flag wait is a formality) and the two
// parts of a group never overlap in time.  Every CTA finishes at the same moment: groups * n / SMs step times instead
// of ceil(groups / SMs) * n.  The image is the coefficient registers plus, on the SKIP path, the evaluation-domain rows:
// the very words the next step would have read, so the results are bit-exact by construction.
template <int LOGN, int DK, int G, bool SKIP, bool TMA = false, bool LAT = false, int SWEEP = 0, bool PERS = false>
__global__ void __launch_bounds__(LAT ? 2 * DK * ((1 << LOGN) / 32) : KCfg<LOGN, DK, G>::NT, 1)
    br_cggi32_kernel(const __grid_constant__ CGGI32Args A) {
    using K = KCfg<LOGN, DK, G>;
    constexpr int N = K::N, TPN = K::TPN, PB = K::PB, NTW = K::NTW, D = K::D, RS = K::RS;
    constexpr int NT = LAT ? 2 * DK * TPN : K::NT;
    static_assert(!LAT || (G == 1 && SKIP && !TMA && DK >= 2), "latency layout: one ciphertext per CTA, skip-top path");
    static_assert(!PERS || (!LAT && !TMA), "persistent variant: throughput shapes with register-staged key loads");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u32* Dsm = reinterpret_cast<u32*>(smem_raw);                      // [G][D][RS]
    constexpr int TABW = TMA ? 1 : CGGI32_TABW;                       // words per monomial-factor entry (see the top of the file)
    u32* psiM = Dsm + (size_t)G * D * RS;                             // [2N][TABW]
    unsigned short* es = reinterpret_cast<unsigned short*>(psiM + 2 * N * TABW);  // [G][n] rotation exponents
    // TMA variant: key ring [2][D][NT] uint4 and 4 mbarriers (full[2], empty[2]) behind the exponents, 128-byte aligned
    uint4* ring = reinterpret_cast<uint4*>(smem_raw + K::ring_offset((int)A.c.n));
    const u32 bar0 = smem_u32(ring + 2 * D * NT);
    constexpr u32 STAGE_BYTES = D * NT * 16;
    constexpr int MIT = (N + NT - 1) / NT;
    const u32 total_fills = A.c.n * MIT;
    auto issue_fill = [&](u32 f) {   // chunk f = (step f / MIT, iteration f % MIT) -> stage f & 1
        const u32 st = f & 1, fb = bar0 + 8 * st;
        const uint4* src = reinterpret_cast<const uint4*>(A.bk) + (size_t)(f / MIT) * D * N + (f % MIT) * NT;
        mbar_expect_tx(fb, STAGE_BYTES);
#pragma unroll
        for (int x = 0; x < D; x++)
            tma_bulk_g2s(smem_u32(ring + (st * D + x) * NT), src + (size_t)x * N, NT * 16, fb);
    };

    const BRCommon& C = A.c;
    const u32 Q = A.mod.Q, Q2 = A.Q2, qinv = A.mod.qinv, oneM = A.mod.oneM;
    const u32 n = C.n;
    const int tid = threadIdx.x;
    const int g = LAT ? 0 : tid / (2 * TPN);   // ciphertext slot within the CTA
    const int j = (tid / TPN) & 1;         // accumulator component (0 = a, 1 = b)
    const int lw = LAT ? tid / (2 * TPN) : 0;  // latency layout: the digit polynomial this warp transforms
    const bool helper = LAT && lw == DK - 1;   // latency layout: pointwise-only warps
    const int T = tid % TPN;               // thread index within the NTT
    const int pbar = 1 + 2 * g + j;        // named barrier of this polynomial's threads (N = 2048 only)
    const bool odd_lane = tid & 1;

    // ---- persistent variant: this CTA's range of the launch's groups * n rotation steps ----------------------------
    // The ranges are handed out in the order the CTAs START (a ticket, as in decoupled look-back scans), not by
    // blockIdx: range k only ever waits for range k - 1, whose CTA is then running or done whatever order the hardware
    // dispatches blocks in.
    u32 bid = blockIdx.x;
    if (PERS) {
        __shared__ u32 s_bid;
        if (tid == 0)
            s_bid = atomicAdd(A.pers_ticket, 1u) - A.pers_ticket_base;
        __syncthreads();
        bid = s_bid;
    }
    PersRange range;
    range.n_items = 1;
    if (PERS)
        range = PersRange::of(A.pers_groups, n, gridDim.x, bid);

    // ---- one-time loads: psi-power table, per-thread twiddles ------------------------------------------------
    // bit-rotated index: the distinct exponents a warp touches differ in their top bits only, which become the low
    // index bits so that they fall into distinct banks (16 x 4 B per warp, 16 x 8 B per half-warp)
    auto tab_index = [&](u32 x) -> u32 { return ((x & (2 * N / 16 - 1)) << 4) | (x >> (LOGN + 1 - 4)); };
    for (int x = tid; x < 2 * N; x += NT) {
        if (TABW == 1)
            psiM[tab_index(x)] = A.psi_pow[x];
        else {
            u32 a1 = A.psi_pow[x], a2 = A.psi_pow[(2 * N - x) & (2 * N - 1)];
            a1 = a1 >= oneM ? a1 - oneM : a1 + Q - oneM;
            a2 = a2 >= oneM ? a2 - oneM : a2 + Q - oneM;
            reinterpret_cast<uint2*>(psiM)[tab_index(x)] = make_uint2(a1, a2);
        }
    }
    u32 tw[32], twp[32];
    {
        const uint2* src = reinterpret_cast<const uint2*>(A.twB) + (size_t)T * NTW;
#pragma unroll
        for (int x = 0; x < NTW; x++) {
            uint2 w = src[x];
            tw[x] = w.x;
            twp[x] = w.y;
        }
    }
    u32* const myD = Dsm + (size_t)g * D * RS;
    const u32 QHalf = Q >> 1;
    const u32 gBits = C.gBits, gmask = (1u << gBits) - 1;
    u32 c[32];
    u32 bk_pre[TMA ? 1 : 4 * D];   // key slice of (step, slot tid), requested one step ahead (register-staged variant)

    for (int item = 0; item < range.n_items; item++) {
    // this item: rotation steps [sb, se) of group grp (the whole rotation unless PERS)
    u32 grp = bid, sb = 0, se = n;
    if (PERS) {
        range.item(item, n, grp, sb, se);
        __syncthreads();   // the previous item is done with the exponents and the digit regions
    }
    const int ct = (int)grp * G + g;
    const bool live = ct < C.batch;
    const u64* lwe = C.ct + (size_t)(live ? ct : 0) * (n + 1);
    {
        // rgsw-acc-cggi.cpp:146-153: e_i = ((mod - a_i) mod mod) * (2N / mod); 0 for dead slots
        const u32 mod = (u32)C.ct_mod, fac = (2 * N) / mod;
        const int lt = tid % (2 * TPN);
        for (u32 i = lt; i < n; i += 2 * TPN) {
            u32 ai = (u32)(lwe[i] % mod);
            u32 e = ((mod - ai) % mod) * fac;
            es[g * n + i] = live ? (unsigned short)e : 0;
        }
    }

    // ---- accumulator initialisation in A layout (coefficient idx = T + TPN*r) --------------------------------
    if (PERS && sb > 0) {
        // resumed group: the accumulator comes from the hand-over slot (below)
    }
    else if (C.acc_init == ACC_EXPLICIT) {
        const u64* src = C.acc_io + ((size_t)(live ? ct : 0) * 2 + j) * N;
#pragma unroll
        for (int r = 0; r < 32; r++)
            c[r] = live ? (u32)src[T + TPN * r] : 0;
    }
    else {
        const u64 q = C.ct_mod, b = lwe[n] % q;
        const u32 factor = (u32)((2 * N) / q);
        const u64 q1 = C.gate_q1;
        u64 q2 = q1 + (q >> 1);
        if (q2 >= q)
            q2 -= q;
        const u64* tab = C.table + (C.acc_init == ACC_TABLE_PER ? (size_t)(live ? ct : 0) * q : 0);
#pragma unroll
        for (int r = 0; r < 32; r++) {
            const u32 idx = T + TPN * r;
            u32 val = 0;
            if (j == 1 && live && (idx % factor) == 0) {
                u64 jj = idx / factor;
                u64 temp = b >= jj ? b - jj : b + q - jj;
                if (C.acc_init == ACC_GATE) {
                    bool in;
                    if (q1 < q2)
                        in = (temp >= q1) && (temp < q2);
                    else
                        in = !((temp >= q2) && (temp < q1));
                    val = in ? (u32)(Q - C.Q8) : (u32)C.Q8;
                }
                else
                    val = (u32)(C.scale * tab[temp]);
            }
            c[r] = val;
        }
    }
    if (TMA && tid == 0) {
        mbar_init(bar0, 1);              // full[0], full[1]: one arrive (the producer's expect_tx) + the bytes
        mbar_init(bar0 + 8, 1);
        mbar_init(bar0 + 16, NT / 32);   // empty[0], empty[1]: one arrive per warp
        mbar_init(bar0 + 24, NT / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (!TMA) {
        const uint4* p4 = reinterpret_cast<const uint4*>(A.bk) + (PERS ? (size_t)sb * D * N : (size_t)0) + tid;
#pragma unroll
        for (int x = 0; x < D; x++) {
            uint4 w = __ldg(p4 + (size_t)x * N);
            bk_pre[4 * x] = w.x; bk_pre[4 * x + 1] = w.y; bk_pre[4 * x + 2] = w.z; bk_pre[4 * x + 3] = w.w;
        }
    }
    else if (tid == 0) {
        issue_fill(0);
        if (total_fills > 1)
            issue_fill(1);
    }

    if (PERS && sb > 0) {
        // the head of this group was run by the previous CTA at the start of the launch: wait for its image
        const u32 slot = bid - 1;
        if (tid == 0) {
            u32 f, spins = 0;
            for (;;) {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(A.pers_flags + slot) : "memory");
                if (f == A.pers_epoch)
                    break;
                __nanosleep(200);
                if (++spins > (1u << 26))   // ~20 s: cannot happen (CTA k-1 is dispatched before CTA k and fills the slot
                    __trap();               // first thing); fail the launch loudly rather than hang or return garbage
            }
        }
        __syncthreads();
        // image: [slot][2][G][2][N] words -- the coefficient registers, then (SKIP path) the evaluation-domain rows
        const uint4* src4 = reinterpret_cast<const uint4*>(A.pers_state) + (((size_t)slot * 2 * G + g) * 2 + j) * (N / 4) + T * 8;
#pragma unroll
        for (int x = 0; x < 8; x++) {
            const uint4 w = __ldcg(src4 + x);
            c[4 * x] = w.x; c[4 * x + 1] = w.y; c[4 * x + 2] = w.z; c[4 * x + 3] = w.w;
        }
        if (SKIP) {
            uint4* p4 = reinterpret_cast<uint4*>(myD + (size_t)(2 * (DK - 1) + j) * RS + 36 * T);
#pragma unroll
            for (int x = 0; x < 8; x++)
                p4[x] = __ldcg(src4 + (size_t)G * 2 * (N / 4) + x);
            __syncthreads();
        }
    }
    else if (SKIP) {
        // Top-digit elimination.  With no digits thrown and B^(DK-1) * (B/2 - 1) > Q/2 the signed digits satisfy
        // c = sum_l d_l B^l EXACTLY (the top digit never wraps), hence NTT(d_top) = B^-(DK-1) (NTT(c) - sum_{l<top}
        // B^l NTT(d_l)).  Substituting into sum_l NTT(d_l) BK_l gives sum_{l<top} NTT(d_l) BK'_l + NTT(c) BK'_top
        // with BK'_l = BK_l - B^(l-top) BK_top and BK'_top = B^-top BK_top (done once at setup).  NTT(c) is simply the
        // evaluation-domain accumulator, maintained as acc_eval += delta in the pointwise stage and kept in the
        // shared-memory region the top digit would have used: 2 of the 2*DK forward transforms per step disappear.
        if (!LAT || lw == 0) {
        u32 v[32];
#pragma unroll
        for (int r = 0; r < 32; r++)
            v[r] = c[r];
        if (SWEEP == 2)
            fwd_passA_sw(v, A, Q, Q2);
        else
            fwd_passA(v, A, Q, Q2);
        if (SWEEP == 1)
            sweep_below_2q(v, Q2);
        if (SWEEP == 2)
            sweep_8q(v, Q2);
        u32* reg = Dsm + (size_t)g * D * RS + (size_t)(j + 2 * (DK - 1)) * RS;
#pragma unroll
        for (int r = 0; r < 32; r++)
            reg[pos_of(T + TPN * r)] = v[r];
        poly_sync<TPN>(pbar);
        {
            const uint4* p4 = reinterpret_cast<const uint4*>(reg + 36 * T);
#pragma unroll
            for (int x = 0; x < 8; x++) {
                uint4 w = p4[x];
                v[4 * x] = w.x; v[4 * x + 1] = w.y; v[4 * x + 2] = w.z; v[4 * x + 3] = w.w;
            }
        }
        poly_sync<TPN>(pbar);
        if (K::XS)
            cross_stage<true>(v, tw[31], twp[31], Q, Q2, A.zero, odd_lane);
        if (SWEEP == 2)
            fwd_passB_sw(v, tw, twp, Q, Q2, A.zero);
        else
            fwd_passB<PB>(v, tw, twp, Q, Q2, A.zero);
#pragma unroll
        for (int r = 0; r < 32; r++) {   // < 24Q (< 2Q with SWEEP = 2) -> canonical
            u32 x = v[r];
            if (SWEEP != 2) {
                x = cond_sub(x, 16 * Q); x = cond_sub(x, 8 * Q); x = cond_sub(x, 4 * Q); x = cond_sub(x, Q2);
            }
            x = cond_sub(x, Q);
            // the key carries N^-1 (unscaled inverse transform), so delta = true_delta / N: keep acc_eval / N as well
            // (the transformed top row carries the compensating factor N, see bk_relayout_cggi32_kernel)
            v[r] = A.mod.mont_mul(x, A.ninvM);
        }
        {
            uint4* p4 = reinterpret_cast<uint4*>(reg + 36 * T);
#pragma unroll
            for (int x = 0; x < 8; x++)
                p4[x] = make_uint4(v[4 * x], v[4 * x + 1], v[4 * x + 2], v[4 * x + 3]);
        }
        }
        __syncthreads();
    }

    // =========================================================================================================
    for (u32 i = sb; i < se; i++) {
        // ---- phase 1: decompose + forward NTT of the DK digit polynomials of component j ---------------------
#ifdef CGGI32_UNROLL_L
#pragma unroll
#else
#pragma unroll 1
#endif
        for (int l = (LAT ? lw : 0); l < (LAT ? (helper ? lw : lw + 1) : (SKIP ? DK - 1 : DK)); l++) {
            u32 v[32];
            const u32 sh = gBits * (l + C.numThrow);
#pragma unroll
            for (int r = 0; r < 32; r++) {
                // centred representative, closed-form signed digit, fed to the lazy NTT as digit + Q
                int dv = (c[r] < QHalf) ? (int)c[r] : (int)c[r] - (int)Q;
                u32 Dv = (u32)(dv + (int)A.dig_off);
                // (the offset of the first sixteen coefficients is added inside the first butterfly stage)
                v[r] = ((u32)((int)Dv >> sh) & gmask) + ((SWEEP == 2 || r >= 16) ? A.dig_add : 0u);
            }
            if (SWEEP == 2)
                fwd_passA_sw(v, A, Q, Q2);
            else
                fwd_passA(v, A, Q, Q2, A.dig_add);
            if (SWEEP == 1)
                sweep_below_2q(v, Q2);
            if (SWEEP == 2)
                sweep_8q(v, Q2);
            u32* reg = myD + (size_t)(j + 2 * l) * RS;
            // transpose A layout -> B layout through the (padded) region
#pragma unroll
            for (int r = 0; r < 32; r++)
                reg[pos_of(T + TPN * r)] = v[r];
            poly_sync<TPN>(pbar);
            {
                const uint4* p4 = reinterpret_cast<const uint4*>(reg + 36 * T);
#pragma unroll
                for (int x = 0; x < 8; x++) {
                    uint4 w = p4[x];
                    v[4 * x] = w.x; v[4 * x + 1] = w.y; v[4 * x + 2] = w.z; v[4 * x + 3] = w.w;
                }
            }
            poly_sync<TPN>(pbar);
            if (K::XS)
                cross_stage<true>(v, tw[31], twp[31], Q, Q2, A.zero, odd_lane);
            if (SWEEP == 2)
                fwd_passB_sw(v, tw, twp, Q, Q2, A.zero);
            else
                fwd_passB<PB>(v, tw, twp, Q, Q2, A.zero);
            {
                uint4* p4 = reinterpret_cast<uint4*>(reg + 36 * T);
#pragma unroll
                for (int x = 0; x < 8; x++)
                    p4[x] = make_uint4(v[4 * x], v[4 * x + 1], v[4 * x + 2], v[4 * x + 3]);
            }
        }
        __syncthreads();

        // ---- phase 2: pointwise MAC against the RGSW keys of step i, monomial factors, delta -> regions 0,1 ---
        // The key slice of evaluation slot k is 4*D words, stored as D "planes" of uint4 ([i][x][k][4]) so a warp
        // reads 512 contiguous bytes per load.  Loads are double-buffered in registers; the first slot of the NEXT
        // step is requested before the inverse transform so its latency hides behind phases 3 and 1.
        {
            constexpr int ITERS = (N + NT - 1) / NT;
            u32 bkv[4 * D];
            if (!TMA) {
#pragma unroll
                for (int x = 0; x < 4 * D; x++)
                    bkv[x] = bk_pre[x];
            }
#pragma unroll
            for (int it = 0; it < ITERS; it++) {
                const int k = tid + it * NT;
                u32 bkn[TMA ? 1 : 4 * D];
                if (TMA) {
                    const u32 f = i * ITERS + it, st = f & 1, par = (f >> 1) & 1;
                    mbar_wait(bar0 + 8 * st, par);                 // the bytes of chunk f have landed
#pragma unroll
                    for (int x = 0; x < D; x++) {
                        const uint4 w = ring[(st * D + x) * NT + tid];
                        bkv[4 * x] = w.x; bkv[4 * x + 1] = w.y; bkv[4 * x + 2] = w.z; bkv[4 * x + 3] = w.w;
                    }
                    __syncwarp();
                    if ((tid & 31) == 0)
                        mbar_arrive(bar0 + 16 + 8 * st);           // this warp no longer needs the stage
                    if (tid == 0 && f + 2 < total_fills) {
                        mbar_wait(bar0 + 16 + 8 * st, par);        // ... nor does any other warp: refill it
                        issue_fill(f + 2);
                    }
                }
                if (!TMA && it + 1 < ITERS && (!LAT || k + NT < N)) {
                    const uint4* p4 = reinterpret_cast<const uint4*>(A.bk) + (size_t)i * D * N + (k + NT);
#pragma unroll
                    for (int x = 0; x < D; x++) {
                        uint4 w = __ldg(p4 + (size_t)x * N);
                        bkn[4 * x] = w.x; bkn[4 * x + 1] = w.y; bkn[4 * x + 2] = w.z; bkn[4 * x + 3] = w.w;
                    }
                }
                const u32 pk = pos_of(k);
                const u32 br = __brev((u32)k) >> (32 - LOGN);
                // Ciphertexts are processed GB at a time: all shared-memory loads of the group first, then the
                // arithmetic (independent streams the scheduler can interleave), then the stores.  Written per ciphertext
                // the store of one and the loads of the next cannot be reordered (possible aliasing), which serialises the
                // long multiply-accumulate / reduction chains (ncu: `wait` was half of this phase's samples).
                #ifndef CGGI32_GB
#define CGGI32_GB 2
#endif
                constexpr int GB = (G % CGGI32_GB == 0) ? CGGI32_GB : ((G % 2 == 0) ? 2 : 1);   // 4 at a time measured slower
                auto redc_lazy = [&](u64 x) -> u32 {   // x < 2^63 -> < 2^31 + Q, congruent x R^-1 (mod Q)
                    u32 lo = (u32)x, hi = (u32)(x >> 32);
                    u32 t = mulhi_w(lo * qinv, Q);
                    return hi - t + Q;
                };
                auto redc_full = [&](u64 x) -> u32 {   // x < Q * 2^32 -> [0, Q)
                    u32 lo = (u32)x, hi = (u32)(x >> 32);
                    u32 t = mulhi_w(lo * qinv, Q);
                    u32 r = hi - t;
                    return hi < t ? r + Q : r;
                };
                if (!LAT || k < N)
#pragma unroll
                for (int g0 = 0; g0 < G; g0 += GB) {
                    u32 xd[GB][D], m1[GB], m2[GB], dl0[GB], dl1[GB];
#pragma unroll
                    for (int b = 0; b < GB; b++) {
                        const u32* dreg = Dsm + (size_t)(g0 + b) * D * RS + pk;
#pragma unroll
                        for (int l = 0; l < D; l++)
                            xd[b][l] = dreg[(size_t)l * RS];
                        const u32 e = es[(g0 + b) * n + i];
                        const u32 xx = ((2 * br + 1) * e) & (2 * N - 1);
                        // the table is stored bit-rotated (see tab_index) so that the distinct exponents a warp touches
                        // (they differ by multiples of 2N/16) fall into distinct banks
                        if (TABW == 1) {
                            const u32 x2 = (2 * N - xx) & (2 * N - 1);
                            m1[b] = psiM[tab_index(xx)];
                            m2[b] = psiM[tab_index(x2)];
                        }
                        else {
                            const uint2 w = reinterpret_cast<const uint2*>(psiM)[tab_index(xx)];
                            m1[b] = w.x; m2[b] = w.y;
                        }
                    }
#pragma unroll
                    for (int b = 0; b < GB; b++) {
                        u64 s00 = 0, s01 = 0, s10 = 0, s11 = 0;
#pragma unroll
                        for (int l = 0; l < D; l++) {
                            const u32 x = xd[b][l];
                            s00 += (u64)x * bkv[(0 * D + l) * 2 + 0];
                            s01 += (u64)x * bkv[(0 * D + l) * 2 + 1];
                            s10 += (u64)x * bkv[(1 * D + l) * 2 + 0];
                            s11 += (u64)x * bkv[(1 * D + l) * 2 + 1];
                        }
                        {
                        const u32 r00 = redc_lazy(s00), r01 = redc_lazy(s01), r10 = redc_lazy(s10), r11 = redc_lazy(s11);
                        u32 a1 = m1[b], a2 = m2[b];
                        if (TABW == 1) {
                            a1 = a1 >= oneM ? a1 - oneM : a1 + Q - oneM;
                            a2 = a2 >= oneM ? a2 - oneM : a2 + Q - oneM;
                        }
                        if (SKIP) {
                            // acc_eval += delta in one go: y = top + hi - t lies in (-Q, 2Q) (mod 2^32), and the canonical
                            // residue is whichever of y, y + Q, y - Q is below Q -- the unsigned minimum of the three
                            // (two VIADDMNMX instead of compare / select / add / add / min: 153.7 -> 152.5 ms per 16384)
                            auto redc_acc = [&](u64 x, u32 top) -> u32 {
                                u32 lo = (u32)x, hi = (u32)(x >> 32);
                                u32 t = mulhi_w(lo * qinv, Q);
                                u32 y = hi - t + top;
                                return min(min(y, y + Q), y - Q);
                            };
                            m1[b] = redc_acc((u64)r00 * a1 + (u64)r10 * a2, xd[b][2 * (DK - 1)]);
                            m2[b] = redc_acc((u64)r01 * a1 + (u64)r11 * a2, xd[b][2 * (DK - 1) + 1]);
                        }
                        else
                        {
                        dl0[b] = redc_full((u64)r00 * a1 + (u64)r10 * a2);
                        dl1[b] = redc_full((u64)r01 * a1 + (u64)r11 * a2);
                        }
                        }
                    }
#pragma unroll
                    for (int b = 0; b < GB; b++) {
                        u32* wreg = Dsm + (size_t)(g0 + b) * D * RS + pk;
                        if (SKIP) {   // phase 3 transforms the updated acc_eval itself; delta is not needed on its own
                            wreg[(size_t)(2 * (DK - 1)) * RS] = m1[b];
                            wreg[(size_t)(2 * (DK - 1) + 1) * RS] = m2[b];
                        }
                        else {
                            wreg[0] = dl0[b];
                            wreg[RS] = dl1[b];
                        }
                    }
                }
                if (!TMA && it + 1 < ITERS) {
#pragma unroll
                    for (int x = 0; x < 4 * D; x++)
                        bkv[x] = bkn[x];
                }
            }
            // request the first slot of the next step now
            if (!TMA && i + 1 < se) {
                const uint4* p4 = reinterpret_cast<const uint4*>(A.bk) + (size_t)(i + 1) * D * N + tid;
#pragma unroll
                for (int x = 0; x < D; x++) {
                    uint4 w = __ldg(p4 + (size_t)x * N);
                    bk_pre[4 * x] = w.x; bk_pre[4 * x + 1] = w.y; bk_pre[4 * x + 2] = w.z; bk_pre[4 * x + 3] = w.w;
                }
            }
        }
        __syncthreads();

        // ---- phase 3: inverse NTT (mirrored block) ------------------------------------------------------------
        // plain path: of delta_j, accumulated into c.  SKIP path: of the evaluation-domain accumulator itself,
        // c = INTT(acc_eval) (identical mod Q because the transform is linear), read from its own region and
        // transposed through the free digit region j, so acc_eval survives for the next step and c is dead between
        // the digit extraction and this point (32 registers less through phases 1 and 2).
        // (the last step of a head part runs it as well: the hand-over image holds both forms of the accumulator)
        if (!helper) {
            u32 v[32];
            u32* reg = myD + (size_t)(LAT ? j + 2 * lw : j) * RS;   // scratch: this warp's own digit region
            const int Tv = TPN - 1 - T;
            {
                const uint4* p4 = reinterpret_cast<const uint4*>((SKIP ? myD + (size_t)(2 * (DK - 1) + j) * RS : reg) + 36 * Tv);
#pragma unroll
                for (int x = 0; x < 8; x++) {
                    uint4 w = p4[x];
                    v[4 * x] = w.x; v[4 * x + 1] = w.y; v[4 * x + 2] = w.z; v[4 * x + 3] = w.w;
                }
            }
            inv_passB<PB, true>(v, tw, twp, Q, Q2, A.zero);   // both paths transform canonical rows
            if (K::XS)
                cross_stage<false>(v, tw[31], twp[31], Q, Q2, A.zero, odd_lane);
            poly_sync<TPN>(pbar);
            {
                uint4* p4 = reinterpret_cast<uint4*>(reg + 36 * Tv);
#pragma unroll
                for (int x = 0; x < 8; x++)
                    p4[x] = make_uint4(v[4 * x], v[4 * x + 1], v[4 * x + 2], v[4 * x + 3]);
            }
            poly_sync<TPN>(pbar);
#pragma unroll
            for (int r = 0; r < 32; r++)
                v[r] = reg[pos_of(T + TPN * r)];
            poly_sync<TPN>(pbar);
            inv_passA(v, A, Q, Q2);
#pragma unroll
            for (int r = 0; r < 32; r++)
                c[r] = SKIP ? cond_sub(v[r], Q) : cond_sub(cond_sub(c[r] + v[r], Q2), Q);   // c < Q, v < 2Q
        }
        // no CTA barrier needed here: phase 1 of the next step writes regions that phase 2 finished reading
        // (barrier above) and region j, which only this thread group read in phase 3 (__syncwarp above).
    }

    if (PERS && se < n) {
        // head part of a split group: leave the accumulator image for the next CTA
        const u32 slot = bid;
        uint4* dst4 = reinterpret_cast<uint4*>(A.pers_state) + (((size_t)slot * 2 * G + g) * 2 + j) * (N / 4) + T * 8;
#pragma unroll
        for (int x = 0; x < 8; x++)
            __stcg(dst4 + x, make_uint4(c[4 * x], c[4 * x + 1], c[4 * x + 2], c[4 * x + 3]));
        if (SKIP) {
            // phase 3 only read the evaluation-domain rows: they are as phase 2 left them
            const uint4* p4 = reinterpret_cast<const uint4*>(myD + (size_t)(2 * (DK - 1) + j) * RS + 36 * T);
#pragma unroll
            for (int x = 0; x < 8; x++)
                __stcg(dst4 + (size_t)G * 2 * (N / 4) + x, p4[x]);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0)
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(A.pers_flags + slot), "r"(A.pers_epoch) : "memory");
        continue;
    }
    // ---- extraction: a'(X) = a(X^-1), b = acc_b[0] (+ Q8 for gates) ---------------------------------------------
    if (live && (!LAT || lw == 0)) {
        if (C.write_acc) {
            u64* dst = C.acc_io + (size_t)ct * 2 * N;
#pragma unroll
            for (int r = 0; r < 32; r++) {
                const u32 idx = T + TPN * r;
                if (j == 0) {
                    u32 val = c[r];
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
                    u32 val = c[r];
                    dst[idx == 0 ? 0 : N - idx] = (idx == 0 || val == 0) ? val : Q - val;
                }
                else if (idx == 0) {
                    u64 val = (u64)c[r] + C.ext_add_b;
                    dst[N] = val >= Q ? val - Q : val;
                }
            }
        }
    }
    }   // items
}

// ---------------------------------------------------------------------------------------------------------------
// Mirrors the CASE table of launch_br_cggi32 exactly (instantiated shapes, shared-memory footprint of the default
// group, 32-bit digit extraction): a set that passes here can be launched, anything else stays on the generic kernel.
bool cggi32_supported(const tfhe_b200_params& p) {
    if (p.method != TFHE_B200_METHOD_GINX)
        return false;
    if (p.N != 512 && p.N != 1024 && p.N != 2048)
        return false;
    if (p.digitsG <= p.numDigitsToThrow)
        return false;
    const u32 dk = p.digitsG - p.numDigitsToThrow;
    bool inst = p.N == 1024 ? (dk >= 2 && dk <= 6) : (dk == 2 || dk == 3 || dk == 4 || dk == 6);
    if (p.N == 2048) {
        // the STD256 family: four digits, top digit exact (elimination), two ciphertexts per CTA; 27-bit moduli on the
        // plain lazy transform (24 Q < 2^32), 29-bit moduli with a sweep every third stage (8 Q < 2^32)
        if (dk != 4 || p.numDigitsToThrow != 0 || !cggi32_skip_top_ok(p))
            return false;
        if (p.Q >= (1ULL << 32) / 24 && p.Q >= (1ULL << 32) / 8)
            return false;
        inst = true;
    }
    else if (cggi32_needs_sweep(p.Q)) {
        // lazy forward NTT bound: values < 22 Q must fit 32 bits; 28-bit moduli run the variant with a mid-transform
        // sweep (12 Q per pass), instantiated for N = 1024 with three or four kept digits (MEDIUM, SIGNED_MOD_TEST)
        if (p.Q >= (1ULL << 28))
            return false;
        inst = p.N == 1024 && (dk == 3 || dk == 4);
    }
    if (!inst)
        return false;
    // digits are extracted from the low 32 bits of (centred value + offset): every digit window must lie inside them
    u32 gbits = 0;
    while ((1ULL << gbits) < p.baseG)
        gbits++;
    if ((1ULL << gbits) != p.baseG || gbits * p.digitsG > 32)
        return false;
    // shared memory of the throughput shape: G * D digit regions + psi table + G * n rotation exponents
    const u32 G = p.N == 2048 ? 2 : (p.N == 1024 ? (dk <= 4 ? 4 : 2) : (dk <= 4 ? 8 : 4));
    const size_t RS = p.N + p.N / 8 + (p.N == 512 ? 16 : 0);
    const size_t smem = (size_t)G * 2 * dk * RS * 4 + (size_t)2 * p.N * 4 * CGGI32_TABW + (size_t)G * ((p.n + 1) / 2 * 2) * 2 + 64;
    if (smem > 227 * 1024)
        return false;
    return true;
}

// top-digit elimination is exact iff the top signed digit can never wrap (see the kernel prologue)
bool cggi32_skip_top_ok(const tfhe_b200_params& p) {
    if (p.numDigitsToThrow != 0 || p.digitsG < 2)
        return false;
    // Exact check: the digit extraction d -> floor((d + B/2) / B) is monotone, so the extreme top digits come from
    // the extreme centred values d_max = QHalf - 1 and d_min = QHalf - Q (rgsw-acc.cpp:83).  The top digit is exact
    // iff it lies in [-B/2, B/2 - 1] BEFORE the sign-extension truncates it.
    const __int128 B = p.baseG, QH = p.Q >> 1;
    __int128 ext[2] = {QH - 1, QH - (__int128)p.Q};
    for (int e = 0; e < 2; e++) {
        __int128 d = ext[e];
        for (u32 i = 1; i < p.digitsG; i++) {
            __int128 t = d + B / 2;                       // floor division for negative values
            d = (t >= 0) ? t / B : -((-t + B - 1) / B);
        }
        if (d < -(B / 2) || d > B / 2 - 1)
            return false;
    }
    return true;
}

// Weaker condition used by the 64-bit kernel, which repairs wrapped top digits on the fly: the offset value
// D = d + sum_i (B/2) B^i must stay in [0, 2 B^digits), so that the masked digits reproduce D mod B^digits and the only
// possible discrepancy is c - sum_l d_l B^l = B^digits, flagged by bit digits*gBits of D.
bool cggi_skip_top_wrapfix_ok(const tfhe_b200_params& p) {
    if (p.numDigitsToThrow != 0 || p.digitsG < 2)
        return false;
    u32 gbits = 0;
    while ((1ULL << gbits) < p.baseG)
        gbits++;
    if ((1ULL << gbits) != p.baseG || gbits * p.digitsG > 62)
        return false;
    const __int128 B = p.baseG, QH = p.Q >> 1;
    __int128 off = 0, pw = 1;
    for (u32 i = 0; i < p.digitsG; i++) {
        off += (B / 2) * pw;
        pw *= B;
    }
    const __int128 dmax = QH - 1, dmin = QH - (__int128)p.Q;
    return dmin + off >= 0 && dmax + off < 2 * pw;
}

static u32 bitrev_h(u32 x, u32 bits) {
    u32 r = 0;
    for (u32 i = 0; i < bits; i++) {
        r = (r << 1) | (x & 1);
        x >>= 1;
    }
    return r;
}
static u32 shoup_h(u64 w, u64 Q) {
    return (u32)((w << 32) / Q);
}

// twA: [fwd|inv][32][2] ; twB: [TPN][NTW][2]   (plain residues + Shoup companions, NOT Montgomery form)
void cggi32_build_tables(const tfhe_b200_params& p, std::vector<u32>& twA, std::vector<u32>& twB) {
    const u64 Q = p.Q, N = p.N;
    const u32 logN = N == 512 ? 9 : (N == 1024 ? 10 : 11);
    const bool xs = logN == 11;   // N = 2048: five in-thread pass-B stages + the cross-lane stage (twiddle in slot 31)
    const u32 TPN = N / 32, PB = xs ? 5 : logN - 5, NTW = xs ? 32 : 32 - (32 >> PB);
    std::vector<u64> W(N), WI(N);
    u64 psi = p.psi % Q, psii = h_powmod(psi, Q - 2, Q), x = 1, xi = 1;
    for (u64 k = 0; k < N; k++) {
        u32 r = bitrev_h((u32)k, logN);
        W[r] = x;
        WI[r] = xi;
        x = h_mulmod(x, psi, Q);
        xi = h_mulmod(xi, psii, Q);
    }
    twA.assign(2 * 32 * 2, 0);
    for (u32 k = 1; k < 32; k++) {
        twA[(0 * 32 + k) * 2 + 0] = (u32)W[k];
        twA[(0 * 32 + k) * 2 + 1] = shoup_h(W[k], Q);
        twA[(1 * 32 + k) * 2 + 0] = (u32)WI[k];
        twA[(1 * 32 + k) * 2 + 1] = shoup_h(WI[k], Q);
    }
    twB.assign((size_t)TPN * NTW * 2, 0);
    for (u32 T = 0; T < TPN; T++)
        for (int s = (int)PB - 1; s >= 0; s--) {
            const u32 cnt = 32 >> (s + 1), off = cnt - (32 >> PB);
            for (u32 xx = 0; xx < cnt; xx++) {
                u64 w = W[(N >> (s + 1)) + T * cnt + xx];
                twB[((size_t)T * NTW + off + xx) * 2 + 0] = (u32)w;
                twB[((size_t)T * NTW + off + xx) * 2 + 1] = shoup_h(w, Q);
            }
        }
    if (xs)
        for (u32 T = 0; T < TPN; T++) {   // stage with 32 groups of 64: the pair of lanes (2b, 2b+1) owns group b
            u64 w = W[32 + (T >> 1)];
            twB[((size_t)T * NTW + 31) * 2 + 0] = (u32)w;
            twB[((size_t)T * NTW + 31) * 2 + 1] = shoup_h(w, Q);
        }
}

// Persistent variant: one CTA per SM, the launch's groups * n steps cut into equal ranges (see the kernel header).
// Instantiated for the shapes the named sets run at large batches.
template <int LOGN, int DK, int G, bool SKIP>
static cudaError_t launch_pers(CGGI32Args a, cudaStream_t s, int ctas) {
    using K = KCfg<LOGN, DK, G>;
    const size_t smem = (K::smem_bytes((int)a.c.n) + (size_t)(CGGI32_TABW - 1) * 2 * K::N * 4);
    if (smem > 227 * 1024)
        return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(br_cggi32_kernel<LOGN, DK, G, SKIP, false, false, 0, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess)
        return e;
    a.pers_groups = (a.c.batch + G - 1) / G;
    br_cggi32_kernel<LOGN, DK, G, SKIP, false, false, 0, true><<<ctas, K::NT, smem, s>>>(a);
    return cudaGetLastError();
}

// Does the throughput shape of this parameter set have a persistent instantiation?  (group = ciphertexts per CTA)
bool cggi32_pers_shape(u32 logN, u32 dk, bool skip_top, u64 Q, int* group) {
    if (cggi32_needs_sweep(Q))
        return false;
    if (logN == 10 && dk == 4 && skip_top) { *group = 4; return true; }
    if (logN == 9 && dk == 3 && !skip_top) { *group = 8; return true; }
    return false;
}

template <int LOGN, int DK, int G, bool SKIP, bool TMA>
static cudaError_t launch_t2(const CGGI32Args& a, cudaStream_t s) {
    using K = KCfg<LOGN, DK, G>;
    const size_t smem = TMA ? K::ring_offset((int)a.c.n) + (size_t)2 * K::D * K::NT * 16 + 64 : (K::smem_bytes((int)a.c.n) + (size_t)(CGGI32_TABW - 1) * 2 * K::N * 4);
    if (smem > 227 * 1024)
        return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(br_cggi32_kernel<LOGN, DK, G, SKIP, TMA>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess)
        return e;
    const int grid = (a.c.batch + G - 1) / G;
    br_cggi32_kernel<LOGN, DK, G, SKIP, TMA><<<grid, K::NT, smem, s>>>(a);
    return cudaGetLastError();
}

// 28-bit moduli: throughput shape only (N = 1024, four ciphertexts per CTA)
template <int DK, bool SKIP>
static cudaError_t launch_sweep(const CGGI32Args& a, cudaStream_t s) {
    using K = KCfg<10, DK, 4>;
    const size_t smem = (K::smem_bytes((int)a.c.n) + (size_t)(CGGI32_TABW - 1) * 2 * K::N * 4);
    if (smem > 227 * 1024)
        return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(br_cggi32_kernel<10, DK, 4, SKIP, false, false, 1>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess)
        return e;
    br_cggi32_kernel<10, DK, 4, SKIP, false, false, 1><<<(a.c.batch + 3) / 4, K::NT, smem, s>>>(a);
    return cudaGetLastError();
}

// N = 2048 (STD256 family): four digits with top-digit elimination, two ciphertexts per CTA (8 warps)
template <int SW>
static cudaError_t launch_n2048(const CGGI32Args& a, cudaStream_t s) {
    using K = KCfg<11, 4, 2>;
    const size_t smem = (K::smem_bytes((int)a.c.n) + (size_t)(CGGI32_TABW - 1) * 2 * K::N * 4);
    if (smem > 227 * 1024)
        return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(br_cggi32_kernel<11, 4, 2, true, false, false, SW>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess)
        return e;
    br_cggi32_kernel<11, 4, 2, true, false, false, SW><<<(a.c.batch + 1) / 2, K::NT, smem, s>>>(a);
    return cudaGetLastError();
}

// latency layout: one ciphertext per CTA, 2*DK warps
template <int LOGN, int DK>
static cudaError_t launch_lat(const CGGI32Args& a, cudaStream_t s) {
    using K = KCfg<LOGN, DK, 1>;
    const size_t smem = (K::smem_bytes((int)a.c.n) + (size_t)(CGGI32_TABW - 1) * 2 * K::N * 4);
    cudaError_t e = cudaFuncSetAttribute(br_cggi32_kernel<LOGN, DK, 1, true, false, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess)
        return e;
    br_cggi32_kernel<LOGN, DK, 1, true, false, true><<<a.c.batch, 2 * DK * K::TPN, smem, s>>>(a);
    return cudaGetLastError();
}

template <int LOGN, int DK, int G>
static cudaError_t launch_t(const CGGI32Args& a, cudaStream_t s, bool skip) {
    // TMA key streaming exists for the headline shape (N = 1024, four digits, four ciphertexts per CTA, top digit
    // eliminated) and is OPT-IN (TFHE_B200_TMA=1): measured 173.5 ms per 16384 bootstraps against 162.8 ms for the
    // register-staged loads on the same box.  The key is L2-resident and the register prefetch already hides its
    // latency a whole phase ahead; the ring adds 8 LDS.128 per thread-iteration, the mbarrier waits and a bubble at the
    // head of every iteration (the multiply-accumulates cannot start before the ring reads return).
    if (LOGN == 10 && DK == 4 && G == 4 && skip && getenv("TFHE_B200_TMA"))
        return launch_t2<LOGN, DK, G, true, (LOGN == 10 && DK == 4 && G == 4)>(a, s);
    return skip ? launch_t2<LOGN, DK, G, true, false>(a, s) : launch_t2<LOGN, DK, G, false, false>(a, s);
}

cudaError_t launch_br_cggi32(const BRCommon& c, const CGGI32Tables& t, cudaStream_t s, int sm_count, int group) {
    CGGI32Args a;
    a.c = c;
    a.mod = t.mod;
    a.bk = t.bk;
    a.psi_pow = t.psi_pow;
    a.twB = t.twB;
    memcpy(a.twA_f, t.twA, sizeof(a.twA_f));
    memcpy(a.twA_i, t.twA + 64, sizeof(a.twA_i));
    a.Q2 = 2 * t.mod.Q;
    const u32 B = 1u << c.gBits, total_digits = c.digitsKept + c.numThrow;
    u64 off = 0, pw = 1;
    for (u32 i = 0; i < total_digits; i++) {
        off += (B / 2) * pw;
        pw *= B;
    }
    a.dig_off = (u32)off;
    a.dig_add = t.mod.Q - B / 2;
    a.zero = 0;
    a.ninvM = to_mont<u32>(h_powmod((u64)1 << c.logN, t.mod.Q - 2, t.mod.Q), t.mod);
    const int dk = (int)c.digitsKept;
    a.pers_state = nullptr; a.pers_flags = nullptr; a.pers_epoch = 0; a.pers_groups = 0;
    a.pers_ticket = nullptr; a.pers_ticket_base = 0;
    {
        // Persistent variant: whenever the plain launch would end on a partial wave (t.pers_ctas > 0 forces a CTA count,
        // tests).  The caller owns the hand-over slots and bumps the epoch per launch.
        int pg = 0;
        if (group == 0 && t.pers_state && t.pers_mode != 0 && cggi32_pers_shape(c.logN, dk, t.skip_top, t.mod.Q, &pg)) {
            const int groups = (c.batch + pg - 1) / pg;
            int ctas = std::min(sm_count, groups);
            bool use = groups > sm_count && groups % sm_count != 0;
            if (t.pers_ctas > 0) {
                ctas = std::min(std::min(t.pers_ctas, groups), sm_count);
                use = true;
            }
            if (use && ctas <= t.pers_slots) {
                a.pers_state = t.pers_state; a.pers_flags = t.pers_flags; a.pers_epoch = t.pers_epoch;
                a.pers_ticket = t.pers_ticket; a.pers_ticket_base = t.pers_ticket_base;
                if (t.pers_launched)
                    *t.pers_launched = ctas;
                if (c.logN == 10)
                    return launch_pers<10, 4, 4, true>(a, s, ctas);
                return launch_pers<9, 3, 8, false>(a, s, ctas);
            }
        }
    }
    if (c.logN == 11) {
        if (dk != 4 || !t.skip_top)
            return cudaErrorInvalidConfiguration;
        return t.mod.Q < (1ULL << 32) / 24 ? launch_n2048<0>(a, s) : launch_n2048<2>(a, s);
    }
    if (cggi32_needs_sweep(t.mod.Q)) {   // 28-bit modulus: see cggi32_supported
        if (c.logN == 10 && dk == 3)
            return t.skip_top ? launch_sweep<3, true>(a, s) : launch_sweep<3, false>(a, s);
        if (c.logN == 10 && dk == 4)
            return t.skip_top ? launch_sweep<4, true>(a, s) : launch_sweep<4, false>(a, s);
        return cudaErrorInvalidConfiguration;
    }
#define CASE(LOGN, DK, GG) \
    if (c.logN == LOGN && dk == DK && group == GG) return launch_t<LOGN, DK, GG>(a, s, t.skip_top);
    if (c.logN == 10) {
        // Throughput shape: 4 ciphertexts per CTA share every key word.  A batch that leaves SMs idle that way is
        // latency-bound instead and runs as CTAs of 2 ciphertexts (4 warps): a rotation step finishes sooner because
        // fewer warps compete for the SM's multiplier pipe.  Measured per bootstrap at batch <= 296 on one B200: 5.77 ms
        // (4 per CTA), 4.47 ms (2 per CTA), 5.1 ms (1 per CTA: one warp per scheduler, latency-bound -- not instantiated).
        // At most one ciphertext per SM: the latency layout (one warp per digit polynomial), see the kernel header.
        if (dk == 4 && t.skip_top && ((group == 0 && c.batch <= sm_count) || group == 1))
            return launch_lat<10, 4>(a, s);
        if (group == 0 && dk == 4 && c.batch <= 2 * sm_count)
            group = 2;
        if (group == 0) group = (dk <= 4) ? 4 : 2;
        CASE(10, 4, 4) CASE(10, 4, 2)
        CASE(10, 3, 4)
        CASE(10, 2, 4)
        CASE(10, 6, 2)
        CASE(10, 5, 2)   // logQ = 11 with one thrown digit (the reference's EvalFloor timing set, time-estimate.cpp:96-123)
    }
    else if (c.logN == 9) {
        if (group == 0) group = (dk <= 4) ? 8 : 4;
        CASE(9, 3, 8)
        CASE(9, 2, 8)
        CASE(9, 4, 8)
        CASE(9, 6, 4)
    }
#undef CASE
    return cudaErrorInvalidConfiguration;
}

}  // namespace tfhe_b200
