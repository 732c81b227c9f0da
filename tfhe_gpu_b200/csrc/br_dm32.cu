// Specialised AP/DM blind rotation for 32-bit moduli (STD128 with method AP: N = 1024, Q 27-bit, baseG = 2^7,
// baseR = 32, digitsR = 2).  Same register-resident machinery as br_cggi32.cu (ntt32.cuh), adapted to the DM
// accumulator (rgsw-acc-dm.cpp:80-110, 306-359):
//
//   for i < n, for k < digitsR:  a0 = k-th base-baseR digit of (q - a_i) mod q;  if a0 == 0 skip
//       acc[j] = sum_{l'=1}^{d-1} NTT(digit_l') * BK[i][a0][k][l'][j]            (REPLACE; row l' = 0 is dropped)
//
// Differences to the CGGI kernel:
//   * the key row is selected by the ciphertext's own digit a0, so the G ciphertexts of a CTA do NOT share key words:
//     every (ciphertext, slot) reads its own 16 words (64 KB per ciphertext-step out of a 2.1 GB table, straight from
//     HBM); all G loads of a slot are issued before the first multiply;
//   * the G ciphertexts still walk the n*digitsR steps in lock-step (barriers), a ciphertext whose digit is zero sits
//     the step out (warp-uniform predicate: a warp is one (ciphertext, component));
//   * the accumulator is REPLACED, so the evaluation-domain accumulator needed by the top-digit elimination is simply
//     the pointwise result of the previous active step -- no accumulation, no extra transform;
//   * top-digit elimination is mandatory here (cggi32_skip_top_ok): 6 forward + 2 inverse transforms per active step;
//     the dropped row l' = 0 is handled by the key transform (BK'_0 = -B^-top BK_top for component a).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ntt32.cuh"

namespace tfhe_b200 {

struct DM32Args {
    BRCommon c;
    ModCtx<u32> mod;
    const u32* bk;       // [keyidx = (i*baseR + a0)*digitsR + k][x(D/2)][slot(N)][4]: word w = l'*2 + j
    const u32* twB;
    u32 twA_f[32][2];
    u32 twA_i[32][2];
    u32 Q2, dig_off, dig_add, ninvM, zero;
};

// LAT = true: latency layout for batches of at most one ciphertext per SM (same idea as br_cggi32.cu): G = 1, 2*DK
// warps; warps 0 .. 2*(DK-1)-1 own one digit polynomial each (the DK-1 warps of a component run the inverse transform
// redundantly, extract their own digit, do ONE forward transform), the last pair only helps in the pointwise stage.
//
// PLAIN = true: no top-digit elimination, for gadgets whose top digit can wrap (baseG = 2^9 with a 27-bit modulus: the
// named STD128_AP / STD128_APOPT sets and TOY): all DK digits of both components are transformed (2*DK forward
// transforms per active step), the key keeps its own rows with row l' = 0 zeroed (rgsw-acc-dm.cpp:353 starts at 1), and
// the pointwise result lands in rows 0 / 1, from where the inverse transform picks it up.
// SWEEP = 1: 28-bit moduli (MEDIUM, SIGNED_MOD_TEST), see sweep_below_2q in ntt32.cuh; SWEEP = 2: 29-bit moduli at N = 2048.
// LOGN = 11 (N = 2048: the STD256 family under AP): 64 threads x 32 coefficients per polynomial with the cross-lane stage
// and the 64-thread named barriers of br_cggi32.cu.
template <int LOGN, int DK, int G, bool LAT = false, bool PLAIN = false, int SWEEP = 0>
__global__ void __launch_bounds__(LAT ? 2 * DK * ((1 << LOGN) / 32) : KCfg<LOGN, DK, G>::NT, 1)
    br_dm32_kernel(const __grid_constant__ DM32Args A) {
    using K = KCfg<LOGN, DK, G>;
    static_assert(!(LAT && (PLAIN || SWEEP)), "latency layout exists for the top-digit-elimination path only");
    constexpr int N = K::N, TPN = K::TPN, PB = K::PB, NTW = K::NTW, D = K::D, RS = K::RS;
    constexpr int NT = LAT ? 2 * DK * TPN : K::NT;
    constexpr int CT_THREADS = LAT ? NT : 2 * TPN;   // threads that serve one ciphertext
    static_assert(!LAT || G == 1, "latency layout: one ciphertext per CTA");
    constexpr int P = D / 2;   // uint4 planes per slot (2*D words)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u32* Dsm = reinterpret_cast<u32*>(smem_raw);              // [G][D][RS]
    int* kidx = reinterpret_cast<int*>(Dsm + (size_t)G * D * RS);  // [G][steps] key row of the step, -1 = sit out

    const BRCommon& C = A.c;
    const u32 Q = A.mod.Q, Q2 = A.Q2, qinv = A.mod.qinv;
    const u32 n = C.n;
    const u32 steps = n * C.digitsR;
    const int tid = threadIdx.x;
    const int g = LAT ? 0 : tid / (2 * TPN), j = (tid / TPN) & 1, T = tid % TPN;
    const int lw = LAT ? tid / (2 * TPN) : 0;     // latency layout: digit polynomial of this warp
    const int pbar = 1 + G + 2 * g + j;           // named barrier of this polynomial's threads (N = 2048 only)
    const bool odd_lane = tid & 1;
    const bool helper = LAT && lw == DK - 1;      // latency layout: pointwise-only warps
    const int ct = blockIdx.x * G + g;
    const bool live = ct < C.batch;
    const u64* lwe = C.ct + (size_t)(live ? ct : 0) * (n + 1);

    {
        // rgsw-acc-dm.cpp:102-109: aI = (q - a_i) mod q with q = the scheme's q (NOT the ciphertext modulus)
        const u32 q = (u32)C.q_lwe;
        const int lt = tid % CT_THREADS;
        for (u32 i = lt; i < n; i += CT_THREADS) {
            u32 aI = (q - (u32)(lwe[i] % q)) % q;
            for (u32 k = 0; k < C.digitsR; k++, aI /= C.baseR) {
                u32 a0 = aI % C.baseR;
                kidx[g * steps + i * C.digitsR + k] = (live && a0) ? (int)((i * C.baseR + a0) * C.digitsR + k) : -1;
            }
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

    // ---- accumulator initialisation in A layout -------------------------------------------------------------------
    u32 c[32];
    if (C.acc_init == ACC_EXPLICIT) {
        const u64* src = C.acc_io + ((size_t)(live ? ct : 0) * 2 + j) * N;
#pragma unroll
        for (int r = 0; r < 32; r++)
            c[r] = live ? (u32)src[T + TPN * r] : 0;
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
            u32 val = 0;
            if (j == 1 && live && (idx & (factor - 1)) == 0) {
                u32 jj = idx >> fshift;
                u32 temp = b >= jj ? b - jj : b + q - jj;
                if (C.acc_init == ACC_GATE) {
                    bool in = (q1 < q2) ? ((temp >= q1) && (temp < q2)) : !((temp >= q2) && (temp < q1));
                    val = in ? (u32)(Q - C.Q8) : (u32)C.Q8;
                }
                else
                    val = (u32)(C.scale * tab[temp]);
            }
            c[r] = val;
        }
    }

    u32* myD = Dsm + (size_t)g * D * RS;
    const u32 QHalf = Q >> 1;
    const u32 gBits = C.gBits, gmask = (1u << gBits) - 1;

    auto load_B = [&](u32 (&v)[32], const u32* reg, int tt) {
        const uint4* p4 = reinterpret_cast<const uint4*>(reg + 36 * tt);
#pragma unroll
        for (int x = 0; x < 8; x++) {
            uint4 w = p4[x];
            v[4 * x] = w.x; v[4 * x + 1] = w.y; v[4 * x + 2] = w.z; v[4 * x + 3] = w.w;
        }
    };
    auto store_B = [&](const u32 (&v)[32], u32* reg, int tt) {
        uint4* p4 = reinterpret_cast<uint4*>(reg + 36 * tt);
#pragma unroll
        for (int x = 0; x < 8; x++)
            p4[x] = make_uint4(v[4 * x], v[4 * x + 1], v[4 * x + 2], v[4 * x + 3]);
    };

    // evaluation-domain accumulator (scaled by N^-1, see br_cggi32.cu) of the initial accumulator -> top-digit region
    if (!PLAIN && (!LAT || lw == 0)) {
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
        u32* reg = myD + (size_t)(j + 2 * (DK - 1)) * RS;
#pragma unroll
        for (int r = 0; r < 32; r++)
            reg[pos_of(T + TPN * r)] = v[r];
        poly_sync<TPN>(pbar);
        load_B(v, reg, T);
        poly_sync<TPN>(pbar);
        if (K::XS)
            cross_stage<true>(v, tw[31], twp[31], Q, Q2, A.zero, odd_lane);
        if (SWEEP == 2)
            fwd_passB_sw(v, tw, twp, Q, Q2, A.zero);
        else
            fwd_passB<PB>(v, tw, twp, Q, Q2, A.zero);
#pragma unroll
        for (int r = 0; r < 32; r++) {
            u32 x = v[r];
            if (SWEEP != 2) {
                x = cond_sub(x, 16 * Q); x = cond_sub(x, 8 * Q); x = cond_sub(x, 4 * Q); x = cond_sub(x, Q2);
            }
            x = cond_sub(x, Q);
            v[r] = A.mod.mont_mul(x, A.ninvM);
        }
        store_B(v, reg, T);
    }
    __syncthreads();

    // The ciphertexts of a CTA share nothing but the tables: from here on every (ciphertext) pair of warps runs on its
    // own, synchronised by a named barrier, and a ciphertext whose refresh digit is zero skips the step entirely.
    const int lt = tid % CT_THREADS;                // thread within the ciphertext
    const int bar_id = 1 + g;                       // ciphertext barriers 1..G, polynomial barriers (N = 2048) above them
    auto ct_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(CT_THREADS) : "memory"); };
    constexpr int MIT = N / CT_THREADS;             // pointwise iterations per thread (slots lt + CT_THREADS*it)
    // evaluation-domain accumulator rows (a, b): the top-digit rows -- or rows 0 / 1 when every row is a digit row
    u32* top0 = myD + (size_t)(PLAIN ? 0 : 2 * (DK - 1)) * RS;

    for (u32 s = 0; s < steps; s++) {
        const int row = kidx[g * steps + s];        // uniform over the ciphertext's threads
        if (row < 0)
            continue;
        // key words of the first pointwise iteration: requested now, consumed after the forward transforms
        const uint4* kp = reinterpret_cast<const uint4*>(A.bk) + (size_t)row * P * N + lt;
        uint4 cur[P];
#pragma unroll
        for (int x = 0; x < P; x++)
            cur[x] = __ldg(kp + (size_t)x * N);
        // ---- phase 1: digits 0..DK-2 of component j -> forward NTT -------------------------------------------------
#pragma unroll 1
        for (int l = (LAT ? lw : 0); l < (LAT ? (helper ? lw : lw + 1) : (PLAIN ? DK : DK - 1)); l++) {
            u32 v[32];
            const u32 sh = gBits * l;
#pragma unroll
            for (int r = 0; r < 32; r++) {
                int dv = (c[r] < QHalf) ? (int)c[r] : (int)c[r] - (int)Q;
                u32 Dv = (u32)(dv + (int)A.dig_off);
                v[r] = ((u32)((int)Dv >> sh) & gmask) + A.dig_add;
            }
            if (SWEEP == 2)
                fwd_passA_sw(v, A, Q, Q2);
            else
                fwd_passA(v, A, Q, Q2);
            if (SWEEP == 1)
                sweep_below_2q(v, Q2);
            if (SWEEP == 2)
                sweep_8q(v, Q2);
            u32* reg = myD + (size_t)(j + 2 * l) * RS;
#pragma unroll
            for (int r = 0; r < 32; r++)
                reg[pos_of(T + TPN * r)] = v[r];
            poly_sync<TPN>(pbar);
            load_B(v, reg, T);
            poly_sync<TPN>(pbar);
            if (K::XS)
                cross_stage<true>(v, tw[31], twp[31], Q, Q2, A.zero, odd_lane);
            if (SWEEP == 2)
                fwd_passB_sw(v, tw, twp, Q, Q2, A.zero);
            else
                fwd_passB<PB>(v, tw, twp, Q, Q2, A.zero);
            store_B(v, reg, T);
        }
        ct_sync();

        // ---- phase 2: pointwise product with the ciphertext's own key row; the result REPLACES the accumulator -----
        // (ciphertext-major: the two warps of the ciphertext cover its N slots, next iteration's key words in flight)
#pragma unroll 1
        for (int it = 0; it < MIT; it++) {
            uint4 nxt[P];
            if (it + 1 < MIT) {
#pragma unroll
                for (int x = 0; x < P; x++)
                    nxt[x] = __ldg(kp + (size_t)x * N + (size_t)(it + 1) * CT_THREADS);
            }
            const u32 pk = pos_of(lt + it * CT_THREADS);
            u32 xd[D];
#pragma unroll
            for (int l = 0; l < D; l++)
                xd[l] = myD[(size_t)l * RS + pk];
            u64 s0 = 0, s1 = 0;
#pragma unroll
            for (int x = 0; x < P; x++) {
                const u32 kw[4] = {cur[x].x, cur[x].y, cur[x].z, cur[x].w};   // words (2x, 0), (2x, 1), (2x+1, 0), (2x+1, 1)
                s0 += (u64)xd[2 * x] * kw[0];
                s1 += (u64)xd[2 * x] * kw[1];
                s0 += (u64)xd[2 * x + 1] * kw[2];
                s1 += (u64)xd[2 * x + 1] * kw[3];
            }
            // sums < D * 22Q * Q < 2^62: reduce in two steps (lazy, then canonical)
            auto redc2 = [&](u64 x) -> u32 {
                u32 lo = (u32)x, hi = (u32)(x >> 32);
                u32 t = mulhi_w(lo * qinv, Q);
                u32 r = hi - t + Q;                       // < 2^31 + Q, == x R^-1 (mod Q)
                r = cond_sub(r, 8 * Q); r = cond_sub(r, 4 * Q); r = cond_sub(r, Q2); r = cond_sub(r, Q);
                return r;
            };
            top0[pk] = redc2(s0);        // the new evaluation-domain accumulator; phase 3 transforms it back
            top0[RS + pk] = redc2(s1);
            if (it + 1 < MIT) {
#pragma unroll
                for (int x = 0; x < P; x++)
                    cur[x] = nxt[x];
            }
        }
        ct_sync();

        // ---- phase 3: c = INTT(result), read from the accumulator row, through row j as scratch ---------------------
        if (!helper) {
            u32 v[32];
            u32* reg = myD + (size_t)(LAT ? j + 2 * lw : j) * RS;   // scratch: this warp's own digit region
            const int Tv = TPN - 1 - T;
            load_B(v, top0 + (size_t)j * RS, Tv);
            inv_passB<PB>(v, tw, twp, Q, Q2, A.zero);
            if (K::XS)
                cross_stage<false>(v, tw[31], twp[31], Q, Q2, A.zero, odd_lane);
            poly_sync<TPN>(pbar);
            store_B(v, reg, Tv);
            poly_sync<TPN>(pbar);
#pragma unroll
            for (int r = 0; r < 32; r++)
                v[r] = reg[pos_of(T + TPN * r)];
            poly_sync<TPN>(pbar);
            inv_passA(v, A, Q, Q2);
#pragma unroll
            for (int r = 0; r < 32; r++)
                c[r] = cond_sub(v[r], Q);   // v < 2Q
        }
    }

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
}

// Shapes instantiated below: N = 1024 with four digits and top-digit elimination (STD128 / STD128_OPT with the DM
// accumulator: the headline AP shape, with latency layouts), three digits plain (the named STD128_AP sets), three digits
// with elimination + sweep (MEDIUM), four digits plain + sweep (SIGNED_MOD_TEST); N = 512 with three digits plain (TOY).
bool dm32_supported(const tfhe_b200_params& p) {
    if (p.method != TFHE_B200_METHOD_AP || p.numDigitsToThrow != 0)
        return false;
    if (p.Q >= (1ULL << 28) && p.N != 2048)
        return false;
    if ((u64)p.n * p.digitsR > 8192)
        return false;
    u32 gbits = 0;
    while ((1ULL << gbits) < p.baseG)
        gbits++;
    if ((1ULL << gbits) != p.baseG || gbits * p.digitsG > 32)
        return false;
    const bool skip = cggi32_skip_top_ok(p), sweep = cggi32_needs_sweep(p.Q);
    if (p.N == 2048)   // the STD256 family: four digits, exact top digit; 27-bit moduli plain lazy, up to 29 bits with sweeps
        return p.digitsG == 4 && skip && p.Q < (1ULL << 32) / 8 && (u64)p.n * p.digitsR <= 4096;
    if (p.N == 1024 && p.digitsG == 4 && skip && !sweep)
        return true;
    if (p.N == 1024 && p.digitsG == 3 && !skip && !sweep)
        return true;
    if (p.N == 1024 && p.digitsG == 3 && skip && sweep)
        return true;
    if (p.N == 1024 && p.digitsG == 4 && !skip && sweep)
        return true;
    if (p.N == 512 && p.digitsG == 3 && !skip && !sweep)
        return true;
    return false;
}

template <int LOGN, int DK, int G, bool PLAIN, int SWEEP>
static cudaError_t launch_dm_t(const DM32Args& a, cudaStream_t s) {
    using K = KCfg<LOGN, DK, G>;
    const size_t smem = (size_t)G * K::D * K::RS * 4 + (size_t)G * a.c.n * a.c.digitsR * 4 + 64;
    if (smem > 227 * 1024)
        return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(br_dm32_kernel<LOGN, DK, G, false, PLAIN, SWEEP>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess)
        return e;
    br_dm32_kernel<LOGN, DK, G, false, PLAIN, SWEEP><<<(a.c.batch + G - 1) / G, K::NT, smem, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_br_dm32(const BRCommon& c, const CGGI32Tables& t, cudaStream_t s, int sm_count, int group) {
    DM32Args a;
    a.c = c;
    a.mod = t.mod;
    a.bk = t.bk;
    a.twB = t.twB;
    memcpy(a.twA_f, t.twA, sizeof(a.twA_f));
    memcpy(a.twA_i, t.twA + 64, sizeof(a.twA_i));
    a.Q2 = 2 * t.mod.Q;
    const u32 B = 1u << c.gBits;
    u64 off = 0, pw = 1;
    for (u32 i = 0; i < c.digitsKept; i++) {
        off += (B / 2) * pw;
        pw *= B;
    }
    a.dig_off = (u32)off;
    a.dig_add = t.mod.Q - B / 2;
    a.zero = 0;
    a.ninvM = to_mont<u32>(h_powmod((u64)1 << c.logN, t.mod.Q - 2, t.mod.Q), t.mod);
    {
        const bool sweep = cggi32_needs_sweep(t.mod.Q), plain = !t.skip_top;
        const int dk = (int)c.digitsKept;
        if (c.logN == 11) {
            if (dk != 4 || plain)
                return cudaErrorInvalidConfiguration;
            return t.mod.Q < (1ULL << 32) / 24 ? launch_dm_t<11, 4, 2, false, 0>(a, s) : launch_dm_t<11, 4, 2, false, 2>(a, s);
        }
        if (c.logN == 9 && dk == 3 && plain && !sweep)
            return launch_dm_t<9, 3, 8, true, 0>(a, s);
        if (c.logN == 10 && dk == 3 && plain && !sweep)
            return (sm_count > 0 && c.batch <= 2 * sm_count) ? launch_dm_t<10, 3, 2, true, 0>(a, s)
                                                             : launch_dm_t<10, 3, 4, true, 0>(a, s);
        if (c.logN == 10 && dk == 3 && !plain && sweep)
            return launch_dm_t<10, 3, 4, false, 1>(a, s);
        if (c.logN == 10 && dk == 4 && plain && sweep)
            return launch_dm_t<10, 4, 4, true, 1>(a, s);
        if (!(c.logN == 10 && dk == 4 && !plain && !sweep))
            return cudaErrorInvalidConfiguration;
    }
    constexpr int G = 4;   // 5 ciphertexts per CTA (168 registers, small spills) measured 43.7 k gates/s against 52.8 k
    using K = KCfg<10, 4, G>;
    const size_t smem = (size_t)G * K::D * K::RS * 4 + (size_t)G * c.n * c.digitsR * 4 + 64;
    if (smem > 227 * 1024)
        return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(br_dm32_kernel<10, 4, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess)
        return e;
    if (group == 1 || (group == 0 && sm_count > 0 && c.batch <= sm_count)) {
        // at most one ciphertext per SM: latency layout (one ciphertext per CTA, 8 warps)
        using K1 = KCfg<10, 4, 1>;
        const size_t smem1 = (size_t)K1::D * K1::RS * 4 + (size_t)c.n * c.digitsR * 4 + 64;
        e = cudaFuncSetAttribute(br_dm32_kernel<10, 4, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
        if (e != cudaSuccess)
            return e;
        br_dm32_kernel<10, 4, 1, true><<<c.batch, 2 * 4 * K1::TPN, smem1, s>>>(a);
        return cudaGetLastError();
    }
    if (group == 2 || (group == 0 && sm_count > 0 && c.batch <= 2 * sm_count)) {
        // at most two ciphertexts per SM: CTAs of two (4 warps per SM instead of 8 competing for the multiplier pipe)
        using K2 = KCfg<10, 4, 2>;
        const size_t smem2 = (size_t)2 * K2::D * K2::RS * 4 + (size_t)2 * c.n * c.digitsR * 4 + 64;
        e = cudaFuncSetAttribute(br_dm32_kernel<10, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e != cudaSuccess)
            return e;
        br_dm32_kernel<10, 4, 2><<<(c.batch + 1) / 2, K2::NT, smem2, s>>>(a);
        return cudaGetLastError();
    }
    const int grid = (c.batch + G - 1) / G;
    br_dm32_kernel<10, 4, G><<<grid, K::NT, smem, s>>>(a);
    return cudaGetLastError();
}

}  // namespace tfhe_b200
