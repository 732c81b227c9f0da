// "Wide" variant of the 54-bit CGGI blind rotation (br_cggi64.cu): 128 threads x 16 coefficients per polynomial instead
// of 64 x 32, i.e. 16 warps per SM instead of 8 at the same shared-memory footprint (two ciphertexts per CTA) and 128
// registers per thread.  ncu on br_cggi64 showed 2 warps per scheduler leaving the multiplier pipe 60 % busy with `wait`
// (fixed-latency dependencies) as the top stall, and one ciphertext per SM running at 0.62x the throughput of two: the
// kernel is starved for warps, not for work per thread.  Same arithmetic (ntt64.cuh), same key layout, same results.
//
// Transform (N = 2048 = 16 x 8 x 16): pass A, strides 1024..128, in registers with coefficient T + 128 r (uniform
// twiddles); transpose through the region (one 128-thread named barrier); pass B, strides 64..16, on the 128-block
// shared by 8 neighbouring threads (7 twiddles per block); warp-local exchange through the region; pass C, strides 8..1,
// on 16 consecutive positions per thread (15 per-thread twiddles from a shared-memory table).  The inverse runs the same
// three passes backwards on the MIRRORED blocks so that the forward tables serve both directions
// (psi^-bitrev(m+i) = -psi^bitrev(m + m-1-i), the sign absorbed by computing (V - U) w).
// Only the top-digit-elimination path exists here (with the wrap repair of br_cggi64.cu); everything else stays on
// br_cggi64.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ntt64w.cuh"

namespace tfhe_b200 {

namespace {

using namespace w64;

struct CGGI64WArgs {
    BRCommon c;
    ModCtx<u64> mod;
    const u64* bk;       // [i][x(2D planes)][slot][2] 27-bit limb pairs (bk_relayout_cggi64_kernel)
    const u64* psi_pow;  // [2N] Montgomery form
    const u64* twC;      // [15][128][2] pass-C twiddles (value, Shoup companion): entry (cnt-1+x, t) = W[128 cnt + cnt t + x]
    const u64* twB;      // [16][8][2]   pass-B twiddles: entry (blk, cnt-1+x) = W[16 cnt + cnt blk + x] (slot 7 unused)
    const u64* twU;      // [2][15][2]   uniform pass-A twiddles: forward W[e+1], then NEGATED inverse
    u64 Q2, dig_off, dig_add, ninvM, zero64, kfix;
    // persistent variant (PERS, see br_cggi32.cu): hand-over slots of the groups split between two CTAs
    u64* pers_state;     // [slot][2][G][2][N]: coefficient registers, then the evaluation-domain accumulator rows
    u32* pers_flags;     // [slot] launch epoch once the slot's image is complete
    u32 pers_epoch;
    u32 pers_groups;     // ceil(batch / G)
    u32* pers_ticket;    // running count of persistent CTAs started on this device (never reset)
    u32 pers_ticket_base;   // ... its value when this launch starts
};

// PLAIN = true: no top-digit elimination (thrown digits, or a top digit that is neither exact nor repairable).  The
// region layout stays the same -- DK - 1 digit rows per component plus the evaluation-domain accumulator rows, which
// are still maintained (acc_eval += delta) and inverse-transformed, only no longer multiplied by a key row: DK - 1 is
// then the number of KEPT digits, the key has 2 (DK - 1) rows per secret-key half, and there is no wrap repair.
//
// PERS = true: persistent variant without wave quantisation -- one CTA per SM, the launch's groups x n rotation steps cut
// into equal ranges, split groups handed over through an accumulator image; the scheme and its argument are in the
// header of br_cggi32_kernel.  The wrap bitmaps are rebuilt from the coefficients in every phase 1, so the image is the
// coefficient registers plus the evaluation-domain rows here as well.
template <int DK, int G, bool PLAIN = false, bool PERS = false>
__global__ void __launch_bounds__(KW<DK, G>::NT, 1) br_cggi64w_kernel(const __grid_constant__ CGGI64WArgs A) {
    using K = KW<DK, G>;
    constexpr int D = K::D, NT = K::NT, NF = DK - 1;
    constexpr int DKEY = PLAIN ? D - 2 : D;   // key rows per secret-key half
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* Dsm = reinterpret_cast<u64*>(smem_raw);                                  // [G][D][N]
    ulonglong2* twC = reinterpret_cast<ulonglong2*>(Dsm + (size_t)G * D * N);    // [15][128]
    ulonglong2* twB = twC + 15 * TPN;                                             // [16][8]
    ulonglong2* twUf = twB + 16 * 8;                                              // [15] uniform forward
    ulonglong2* twUi = twUf + 15;                                                 // [15] uniform inverse, negated
    u32* wbits = reinterpret_cast<u32*>(twUi + 15 + 2);                           // [2][G][2][WB]
    u32* wany = wbits + 2 * G * 2 * K::WB;                                        // [2][G][2][4]

    const BRCommon& C = A.c;
    const u64 Q = A.mod.Q, Q2 = A.Q2, QO = 2 * A.Q2, nQ = 0 - A.mod.Q, qinv = A.mod.qinv, oneM = A.mod.oneM;
    const u64 Z = A.zero64;
    const u32 n = C.n;
    const int tid = threadIdx.x;
    const int g = tid / (2 * TPN), j = (tid / TPN) & 1, T = tid % TPN;
    const int bar_id = 1 + g * 2 + j;
    const int blk = T >> 3, u8 = T & 7;

    // persistent variant: this CTA's range of the launch's groups * n rotation steps (br_cggi32.cu)
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

    for (int x = tid; x < 15 * TPN; x += NT)
        twC[x] = reinterpret_cast<const ulonglong2*>(A.twC)[x];
    for (int x = tid; x < 16 * 8; x += NT)
        twB[x] = reinterpret_cast<const ulonglong2*>(A.twB)[x];
    for (int x = tid; x < 2 * 15; x += NT)
        twUf[x] = reinterpret_cast<const ulonglong2*>(A.twU)[x];

    u64 c[CPT];
    u64* const myD = Dsm + (size_t)g * D * N;
    u64* const top = myD + (size_t)(j + 2 * (DK - 1)) * N;   // evaluation-domain accumulator row of this component
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

    for (int item = 0; item < range.n_items; item++) {
    // this item: rotation steps [sb, se) of group grp (the whole rotation unless PERS)
    u32 grp = bid, sb = 0, se = n;
    if (PERS) {
        range.item(item, n, grp, sb, se);
        __syncthreads();   // the previous item is done with the digit regions and the wrap bitmaps
    }
    const int ct = (int)grp * G + g;
    const bool live = ct < C.batch;
    const u64* lwe = C.ct + (size_t)(live ? ct : 0) * (n + 1);

    // ---- accumulator initialisation in A layout (coefficient idx = T + 128 r) -----------------------------------
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
                if (++spins > (1u << 26))   // cannot happen: fail the launch loudly rather than hang (br_cggi32.cu)
                    __trap();
            }
        }
        __syncthreads();
        const u64* src = A.pers_state + (((size_t)slot * 2 * G + g) * 2 + j) * N;
#pragma unroll
        for (int r = 0; r < CPT; r++)
            c[r] = __ldcg(src + T + TPN * r);
#pragma unroll
        for (int r = 0; r < CPT; r++)
            top[T + TPN * r] = __ldcg(src + (size_t)G * 2 * N + T + TPN * r);
    }
    else if (C.acc_init == ACC_EXPLICIT) {
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

    if (!PERS || sb == 0) {
        // evaluation-domain accumulator (scaled by N^-1), see br_cggi32.cu
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

    for (u32 i = sb; i < se; i++) {
        // ---- phase 1: wrapped-top-digit detection, digits 0..DK-2 -> forward transforms ----------------------------
        if (!PLAIN) {
            u32 wm = 0;
            const u32 wsh = gBits * DK;
#pragma unroll
            for (int r = 0; r < CPT; r++) {
                i64 dv = (c[r] < QHalf) ? (i64)c[r] : (i64)c[r] - (i64)Q;
                u64 Dv = (u64)(dv + (i64)A.dig_off);
                wm |= (u32)((Dv >> wsh) & 1) << r;
            }
            const bool anyw = __any_sync(0xffffffffu, wm != 0);
            const int par = i & 1;
            if (anyw) {
                // coefficient index T + 128 r = 32 * (4 r + T / 32) + T % 32: a warp's ballot IS the bitmap word
                u32* wb = wbits + ((size_t)(par * G + g) * 2 + j) * K::WB + (T >> 5);
#pragma unroll
                for (int r = 0; r < CPT; r++) {
                    u32 word = __ballot_sync(0xffffffffu, (wm >> r) & 1);
                    if ((T & 31) == 0)
                        wb[4 * r] = word;
                }
            }
            if ((T & 31) == 0)
                wany[((par * G + g) * 2 + j) * 4 + (T >> 5)] = anyw;
        }
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
            const u32 sh = gBits * (l + (PLAIN ? C.numThrow : 0));
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
        __syncthreads();

        // ---- phase 2: pointwise stage (wrap repair first, see br_cggi64.cu) ------------------------------------------
        bool anyflag = false;
        if (!PLAIN) {
            const u32* fl = wany + (size_t)(i & 1) * G * 8;
            u32 f = 0;
#pragma unroll
            for (int x = 0; x < G * 8; x++)
                f |= fl[x];
            anyflag = f != 0;
        }
        auto wrap_fix = [&](bool undo) {
            const u32* fl = wany + (size_t)(i & 1) * G * 8;
#pragma unroll 1
            for (int gj = 0; gj < G * 2; gj++) {
                if (!(fl[gj * 4] | fl[gj * 4 + 1] | fl[gj * 4 + 2] | fl[gj * 4 + 3]))
                    continue;
                const u32* wb = wbits + ((size_t)(i & 1) * G * 2 + gj) * K::WB;
                u64* reg = Dsm + (size_t)(gj >> 1) * D * N + (size_t)(2 * (DK - 1) + (gj & 1)) * N;
#pragma unroll 1
                for (int k = tid; k < N; k += NT) {
                    const u32 br = __brev((u32)k) >> (32 - LOGN);
                    u64 sum = 0;
#pragma unroll 1
                    for (int wd = 0; wd < K::WB; wd++) {
                        if (!fl[gj * 4 + (wd & 3)])
                            continue;
                        u32 bits = wb[wd];
                        while (bits) {
                            const u32 k0 = 32 * wd + (__ffs(bits) - 1);
                            bits &= bits - 1;
                            sum = csub(sum + __ldg(A.psi_pow + (((2 * br + 1) * k0) & (2 * N - 1))), Q);
                        }
                    }
                    const u64 term = A.mod.mont_mul(sum, A.kfix);
                    const u64 x = reg[posw(k)];
                    reg[posw(k)] = undo ? csub(x + term, Q) : (x >= term ? x - term : x + Q - term);
                }
            }
        };
        if (anyflag) {
            wrap_fix(false);
            __syncthreads();
        }
        {
            constexpr int ITERS = N / NT;
            static_assert(N % NT == 0, "unsupported CTA shape");
            constexpr int PL = 2 * DKEY;
            const ulonglong2* bki = reinterpret_cast<const ulonglong2*>(A.bk) + (size_t)i * PL * N;
            u32 ee[G];
#pragma unroll
            for (int gg = 0; gg < G; gg++) {
                // rgsw-acc-cggi.cpp:146-153: e_i = ((mod - a_i) mod mod) * (2N / mod); 0 for dead slots
                const int cg = (int)grp * G + gg;
                u64 e = 0;
                if (cg < C.batch) {
                    u64 ai = C.ct[(size_t)cg * (n + 1) + i] % C.ct_mod;
                    e = ((C.ct_mod - ai) % C.ct_mod) * ((2 * N) / C.ct_mod);
                }
                ee[gg] = (u32)e;
            }
#pragma unroll 1
            for (int it = 0; it < ITERS; it++) {
                const int k = tid + it * NT;
                u64 bkv[4 * DKEY];
#pragma unroll
                for (int x = 0; x < PL; x++) {
                    ulonglong2 w = bki[(size_t)x * N + k];
                    bkv[2 * x] = w.x;
                    bkv[2 * x + 1] = w.y;
                }
                const u32 pk = posw(k);
                const u32 br = __brev((u32)k) >> (32 - LOGN);
#pragma unroll
                for (int gg = 0; gg < G; gg++) {
                    const u32 e = ee[gg];
                    u64* dreg = Dsm + (size_t)gg * D * N + pk;
                    u64 xd[D];
#pragma unroll
                    for (int l = 0; l < D; l++)
                        xd[l] = dreg[(size_t)l * N];
                    const u32 xx = (u32)(((2 * br + 1) * e) & (2 * N - 1));
                    u64 f1 = __ldg(A.psi_pow + xx);
                    u64 f2 = __ldg(A.psi_pow + ((2 * N - xx) & (2 * N - 1)));
                    const Limb x0(xd[0]);
                    L3 a00(x0, bkv[(0 * DKEY) * 2 + 0]), a01(x0, bkv[(0 * DKEY) * 2 + 1]);
                    L3 a10(x0, bkv[(1 * DKEY) * 2 + 0]), a11(x0, bkv[(1 * DKEY) * 2 + 1]);
#pragma unroll
                    for (int l = 1; l < DKEY; l++) {
                        const Limb x(xd[l]);
                        a00.mac(x, bkv[(0 * DKEY + l) * 2 + 0]);
                        a01.mac(x, bkv[(0 * DKEY + l) * 2 + 1]);
                        a10.mac(x, bkv[(1 * DKEY + l) * 2 + 0]);
                        a11.mac(x, bkv[(1 * DKEY + l) * 2 + 1]);
                    }
                    const Limb s00(redc128(a00.value(), Q, qinv)), s01(redc128(a01.value(), Q, qinv));
                    const Limb s10(redc128(a10.value(), Q, qinv)), s11(redc128(a11.value(), Q, qinv));
                    f1 = f1 >= oneM ? f1 - oneM : f1 + Q - oneM;
                    f2 = f2 >= oneM ? f2 - oneM : f2 + Q - oneM;
                    const Limb F1(f1), F2(f2);
                    L3 t0(s00, F1), t1(s01, F1);
                    t0.mac(s10, F2);
                    t1.mac(s11, F2);
                    const u64 dl0 = redc128(t0.value(), Q, qinv);
                    const u64 dl1 = redc128(t1.value(), Q, qinv);
                    dreg[(size_t)(2 * (DK - 1)) * N] = csub(xd[2 * (DK - 1)] + dl0, Q);
                    dreg[(size_t)(2 * (DK - 1) + 1) * N] = csub(xd[2 * (DK - 1) + 1] + dl1, Q);
                }
            }
        }
        __syncthreads();
        if (anyflag) {
            wrap_fix(true);
            __syncthreads();
        }

        // ---- phase 3: c = INTT(evaluation-domain accumulator), mirrored blocks, scratch = row j -----------------------
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

    if (PERS && se < n) {
        // head part of a split group: leave the accumulator image for the next CTA (phase 3 only read the top rows)
        const u32 slot = bid;
        u64* dst = A.pers_state + (((size_t)slot * 2 * G + g) * 2 + j) * N;
#pragma unroll
        for (int r = 0; r < CPT; r++)
            __stcg(dst + T + TPN * r, c[r]);
#pragma unroll
        for (int r = 0; r < CPT; r++)
            __stcg(dst + (size_t)G * 2 * N + T + TPN * r, top[T + TPN * r]);
        __threadfence();
        __syncthreads();
        if (tid == 0)
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(A.pers_flags + slot), "r"(A.pers_epoch) : "memory");
        continue;
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
    }   // items
}

u32 bitrev_w(u32 x, u32 bits) {
    u32 r = 0;
    for (u32 i = 0; i < bits; i++) {
        r = (r << 1) | (x & 1);
        x >>= 1;
    }
    return r;
}
u64 shoup_w(u64 w, u64 Q) {
    return (u64)((((unsigned __int128)w) << 64) / Q);
}

template <int DK, int G, bool PLAIN = false>
cudaError_t launch_w_pers(CGGI64WArgs a, cudaStream_t s, int ctas) {
    using K = KW<DK, G>;
    cudaError_t e = cudaFuncSetAttribute(br_cggi64w_kernel<DK, G, PLAIN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)K::smem);
    if (e != cudaSuccess)
        return e;
    a.pers_groups = (a.c.batch + G - 1) / G;
    br_cggi64w_kernel<DK, G, PLAIN, true><<<ctas, K::NT, K::smem, s>>>(a);
    return cudaGetLastError();
}

template <int DK, int G, bool PLAIN = false>
cudaError_t launch_w(const CGGI64WArgs& a, cudaStream_t s) {
    using K = KW<DK, G>;
    if (K::smem > 227 * 1024)
        return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(br_cggi64w_kernel<DK, G, PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)K::smem);
    if (e != cudaSuccess)
        return e;
    const int grid = (a.c.batch + G - 1) / G;
    br_cggi64w_kernel<DK, G, PLAIN><<<grid, K::NT, K::smem, s>>>(a);
    return cudaGetLastError();
}

}  // namespace

// plain path (no top-digit elimination): one or two kept digits
bool cggi64w_plain_supported(const tfhe_b200_params& p) {
    if (!cggi64_supported(p))
        return false;
    const u32 kept = p.digitsG - p.numDigitsToThrow;
    return kept == 1 || kept == 2;
}

bool cggi64w_supported(const tfhe_b200_params& p) {
    if (!cggi64_supported(p) || p.numDigitsToThrow != 0)
        return false;
    // 4 digits: 8 digit regions of 16 KB leave room for ONE ciphertext per CTA (8 warps) -- still twice the warps of the
    // 64 x 32 layout, which is limited to one ciphertext per CTA as well at this size
    return p.digitsG == 2 || p.digitsG == 3 || p.digitsG == 4;
}

// twC: [15][128][2]; twB: [16][8][2]; twU: [fwd | negated inv][15][2]
void cggi64w_build_tables(const tfhe_b200_params& p, std::vector<u64>& twU, std::vector<u64>& twB, std::vector<u64>& twC) {
    const u64 Q = p.Q;
    std::vector<u64> W(N), WI(N);
    u64 psi = p.psi % Q, psii = h_powmod(psi, Q - 2, Q), x = 1, xi = 1;
    for (u32 k = 0; k < (u32)N; k++) {
        u32 r = bitrev_w(k, LOGN);
        W[r] = x;
        WI[r] = xi;
        x = h_mulmod(x, psi, Q);
        xi = h_mulmod(xi, psii, Q);
    }
    twU.assign(2 * 15 * 2, 0);
    for (u32 e = 0; e < 15; e++) {
        twU[(0 * 15 + e) * 2 + 0] = W[e + 1];
        twU[(0 * 15 + e) * 2 + 1] = shoup_w(W[e + 1], Q);
        u64 neg = (Q - WI[e + 1]) % Q;
        twU[(1 * 15 + e) * 2 + 0] = neg;
        twU[(1 * 15 + e) * 2 + 1] = shoup_w(neg, Q);
    }
    twB.assign(16 * 8 * 2, 0);
    for (u32 blk = 0; blk < 16; blk++)
        for (u32 cnt = 1; cnt <= 4; cnt *= 2)
            for (u32 xx = 0; xx < cnt; xx++) {
                u64 w = W[16 * cnt + cnt * blk + xx];
                twB[(blk * 8 + cnt - 1 + xx) * 2 + 0] = w;
                twB[(blk * 8 + cnt - 1 + xx) * 2 + 1] = shoup_w(w, Q);
            }
    twC.assign((size_t)15 * TPN * 2, 0);
    for (u32 cnt = 1; cnt <= 8; cnt *= 2)
        for (u32 t = 0; t < (u32)TPN; t++)
            for (u32 xx = 0; xx < cnt; xx++) {
                u64 w = W[128 * cnt + cnt * t + xx];
                twC[((size_t)(cnt - 1 + xx) * TPN + t) * 2 + 0] = w;
                twC[((size_t)(cnt - 1 + xx) * TPN + t) * 2 + 1] = shoup_w(w, Q);
            }
}

cudaError_t launch_br_cggi64w(const BRCommon& c, const CGGI64WTables& t, cudaStream_t s, int sm_count, int group) {
    CGGI64WArgs a;
    a.c = c;
    a.mod = t.mod;
    a.bk = t.bk;
    a.psi_pow = t.psi_pow;
    a.twC = t.twC;
    a.twB = t.twB;
    a.twU = t.twU;
    a.Q2 = 2 * t.mod.Q;
    const u64 B = 1ULL << c.gBits;
    unsigned __int128 off = 0, pw = 1;
    for (u32 i = 0; i < c.digitsKept + (t.plain ? c.numThrow : 0); i++) {
        off += (B / 2) * pw;
        pw *= B;
    }
    a.dig_off = (u64)off;
    a.dig_add = t.mod.Q - B / 2;
    a.zero64 = 0;
    const u64 ninv = h_powmod((u64)N, t.mod.Q - 2, t.mod.Q);
    a.ninvM = to_mont<u64>(ninv, t.mod);
    a.kfix = h_mulmod((u64)(pw % t.mod.Q), ninv, t.mod.Q);
    // a batch of at most one ciphertext per SM is latency-bound: one ciphertext per CTA (8 warps instead of 16 competing
    // for the SM) finishes a rotation step sooner; `group` = 1 / 2 forces a shape (tests, measurements)
    const bool one = group == 1 || (group == 0 && sm_count > 0 && c.batch <= sm_count);
    a.pers_state = nullptr; a.pers_flags = nullptr; a.pers_epoch = 0; a.pers_groups = 0;
    a.pers_ticket = nullptr; a.pers_ticket_base = 0;
    if (!t.plain && group == 0 && t.pers_state && t.pers_mode != 0 && (c.digitsKept == 2 || c.digitsKept == 3)) {
        // persistent variant of the two-ciphertext shapes: whenever the plain launch would end on a partial wave
        // (t.pers_ctas > 0 forces a CTA count, tests)
        const int groups = (c.batch + 1) / 2;
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
            return c.digitsKept == 2 ? launch_w_pers<2, 2>(a, s, ctas) : launch_w_pers<3, 2>(a, s, ctas);
        }
    }
    if (t.plain) {   // DK template = kept digits + 1 (the accumulator rows take the place of the eliminated digit)
        if (c.digitsKept == 1)
            return one ? launch_w<2, 1, true>(a, s) : launch_w<2, 2, true>(a, s);
        if (c.digitsKept == 2)
            return one ? launch_w<3, 1, true>(a, s) : launch_w<3, 2, true>(a, s);
        return cudaErrorInvalidConfiguration;
    }
    if (c.digitsKept == 2)
        return one ? launch_w<2, 1>(a, s) : launch_w<2, 2>(a, s);
    if (c.digitsKept == 3)
        return one ? launch_w<3, 1>(a, s) : launch_w<3, 2>(a, s);
    if (c.digitsKept == 4)
        return launch_w<4, 1>(a, s);
    return cudaErrorInvalidConfiguration;
}

}  // namespace tfhe_b200
