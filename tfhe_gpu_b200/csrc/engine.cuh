// Internal declarations shared by the translation units of libtfhe_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/tfhe_b200.h"
#include "modarith.cuh"

namespace tfhe_b200 {

// ---- accumulator initialisation descriptor (binfhe-base-scheme.cpp:1087-1138 gate, :1147-1185 function) ----
enum AccInit : int {
    ACC_GATE = 0,      // m[j*factor] = in_range((b-j) mod q) ? -Q8 : +Q8
    ACC_TABLE = 1,     // m[j*factor] = (Q/fmod) * table[(b-j) mod ctmod]          (shared table)
    ACC_TABLE_PER = 2, // same, table + ct*ctmod
    ACC_EXPLICIT = 3,  // acc given by the caller ([batch][2][N] u64, COEFFICIENT)  (EvalAcc_CUDA contract)
};

struct BRCommon {
    // ring / gadget
    u32 N, logN, n, d;          // d = number of digit polynomials (2 * kept digits)
    u32 gBits, numThrow, digitsKept;
    u32 method;                 // TFHE_B200_METHOD_*
    u32 baseR, digitsR;
    u64 q_lwe;                  // params q (DM uses it for the refresh digits, gates for the constants)
    // per call
    int batch;
    const u64* ct;              // [batch][n+1] prepared ciphertext (a, b), modulus ct_mod
    u64 ct_mod;
    int acc_init;
    u64 gate_q1;                // ACC_GATE: q1 = gateConst[gate]
    u64 Q8;                     // Q/8 + 1
    u64 scale;                  // ACC_TABLE*: Q / fmod
    const u64* table;           // ACC_TABLE*: f values, ct_mod entries (per ct: [batch][ct_mod])
    u64* acc_io;                // ACC_EXPLICIT input and/or optional output [batch][2][N] (may be null)
    int write_acc;              // write the final accumulator (a transposed) to acc_io
    u64* ext;                   // [batch][N+1] extracted LWE mod Q (b += Q8 for gates); may be null
    u64 ext_add_b;              // constant added to b on extraction (Q8 for gates, 0 for functions)
};

// Persistent blind rotation (br_cggi32.cu, br_cggi64w.cu): the groups * n rotation steps of a launch are cut into
// `ctas` equal contiguous ranges (McNaughton's wrap-around rule).  Range k is the last steps of group gA (from step sA),
// n_full whole groups starting at first_full, and the first sB steps of group gB; it is worked off as items in the order
// head of gB, whole groups, tail of gA, so that the two parts of a split group never run at the same time.
struct PersRange {
    u32 gA, sA, gB, sB, first_full;
    int n_full, n_items;
    __host__ __device__ static inline PersRange of(u32 groups, u32 n, u32 ctas, u32 k) {
        PersRange r;
        const u64 Wt = (u64)groups * n;
        const u64 lo = Wt * k / ctas, hi = Wt * (k + 1) / ctas;
        r.gA = (u32)(lo / n); r.sA = (u32)(lo % n); r.gB = (u32)(hi / n); r.sB = (u32)(hi % n);
        r.first_full = r.gA + (r.sA ? 1 : 0);
        r.n_full = (int)r.gB - (int)r.first_full;
        r.n_items = (r.sB ? 1 : 0) + r.n_full + (r.sA ? 1 : 0);
        return r;
    }
    // item t of the range: rotation steps [sb, se) of group grp
    __host__ __device__ inline void item(int t, u32 n, u32& grp, u32& sb, u32& se) const {
        const int u = t - (sB ? 1 : 0);
        sb = 0; se = n;
        if (u < 0) { grp = gB; se = sB; }
        else if (u < n_full) grp = first_full + (u32)u;
        else { grp = gA; sb = sA; }
    }
};

template <typename T>
struct BRTables {
    ModCtx<T> mod;
    const T* tw_fwd;   // [N] psi^bitrev(k) in Montgomery form
    const T* tw_inv;   // [N] psi^-bitrev(k) in Montgomery form
    const T* psi_pow;  // [2N] psi^x in Montgomery form (monomial factors)
    const T* sh_fwd;   // [2][N] plain psi^bitrev(k) and its Shoup companion floor(w 2^w / Q)  (64-bit lazy NTT path)
    const T* sh_inv;   // [2][N] same for psi^-bitrev(k)
    const T* bk;       // generic layout, Montgomery form, pre-multiplied by N^-1
};

// generic kernels (br_generic.cu)
template <typename T>
cudaError_t launch_br_generic(const BRCommon& c, const BRTables<T>& t, cudaStream_t s, int sm_count);

// optimised CGGI kernel for Q < 2^31, N in {512, 1024} (br_cggi32.cu)
struct CGGI32Tables {
    ModCtx<u32> mod;
    const u32* bk;        // [i][k][key][l][j]  Montgomery form * N^-1
    const u32* psi_pow;   // [2N] Montgomery form
    const u32* twB;       // per-thread pass-B twiddles + Shoup companions, forward
    bool skip_top;        // keys were transformed for top-digit elimination (see br_cggi32.cu)
    const u32* twA;       // HOST pointer: uniform pass-A twiddles + companions [fwd|inv][32][2] (kernel params)
    // persistent variant (br_cggi32.cu): hand-over slots of split groups, owned by the device record
    u32* pers_state = nullptr;   // [pers_slots][PERS_SLOT_WORDS]
    u32* pers_flags = nullptr;   // [pers_slots], zero at allocation
    u32 pers_epoch = 0;          // unique per launch on this device (never 0)
    u32* pers_ticket = nullptr;  // device counter of persistent CTAs started so far (ranges are handed out in start order)
    u32 pers_ticket_base = 0;    // its value at the start of this launch (host shadow)
    int* pers_launched = nullptr;   // HOST out: persistent CTAs launched by this call (0 = a plain launch)
    int pers_slots = 0;
    int pers_mode = 1;           // 0 = never, 1 = automatic
    int pers_ctas = 0;           // > 0: force the persistent variant with this many CTAs (tests)
};
constexpr size_t PERS_SLOT_WORDS = 16384;   // G * 2 * N words of the largest persistent shape (4 x 2 x 1024, 8 x 2 x 512)
bool cggi32_pers_shape(u32 logN, u32 dk, bool skip_top, u64 Q, int* group);
bool cggi32_supported(const tfhe_b200_params& p);
// moduli between 2^32/22 and 2^28 run the cggi32 variant with a mid-transform reduction sweep (ntt32.cuh)
inline bool cggi32_needs_sweep(u64 Q) { return Q >= (1ULL << 32) / 22; }
bool cggi32_skip_top_ok(const tfhe_b200_params& p);
bool cggi_skip_top_wrapfix_ok(const tfhe_b200_params& p);   // top digit may wrap but the 64-bit kernel can repair it
cudaError_t launch_br_cggi32(const BRCommon& c, const CGGI32Tables& t, cudaStream_t s, int sm_count, int group);
bool dm32_supported(const tfhe_b200_params& p);
cudaError_t launch_br_dm32(const BRCommon& c, const CGGI32Tables& t, cudaStream_t s, int sm_count = 0, int group = 0);
void cggi32_build_tables(const tfhe_b200_params& p, std::vector<u32>& twA, std::vector<u32>& twB);

// optimised CGGI kernel for the 54-bit sets, N = 2048 (br_cggi64.cu)
struct CGGI64Tables {
    ModCtx<u64> mod;
    const u64* bk;        // [i][x(2D)][slot][2]
    const u64* psi_pow;   // [2N] Montgomery form
    const u64* twB;       // device [31][64][2]
    const u64* tw32;      // device [32][2]
    const u64* twA;       // device [fwd | negated inv][31][2] uniform pass-A twiddles
    bool skip_top;
};
bool cggi64_supported(const tfhe_b200_params& p);
void cggi64_build_tables(const tfhe_b200_params& p, std::vector<u64>& twA, std::vector<u64>& twB, std::vector<u64>& tw32);
cudaError_t launch_br_cggi64(const BRCommon& c, const CGGI64Tables& t, cudaStream_t s, int group);

// wide variant (128 threads x 16 coefficients per polynomial, br_cggi64w.cu): top-digit-elimination path only
struct CGGI64WTables {
    ModCtx<u64> mod;
    const u64* bk;        // same re-laid-out key as CGGI64Tables (skip-top transform applied)
    const u64* psi_pow;
    const u64* twC;       // device [15][128][2]
    const u64* twB;       // device [16][8][2]
    const u64* twU;       // device [2][15][2]
    bool plain = false;   // no top-digit elimination (untransformed key, thrown digits honoured)
    // persistent variant (see CGGI32Tables): slots of PERS_SLOT_WORDS64 u64
    u64* pers_state = nullptr;
    u32* pers_flags = nullptr;
    u32 pers_epoch = 0;
    u32* pers_ticket = nullptr;
    u32 pers_ticket_base = 0;
    int* pers_launched = nullptr;
    int pers_slots = 0, pers_mode = 1, pers_ctas = 0;
};
constexpr size_t PERS_SLOT_WORDS64 = 16384;   // 2 images x 2 ciphertexts x 2 components x 2048 coefficients
bool cggi64w_supported(const tfhe_b200_params& p);
bool cggi64w_plain_supported(const tfhe_b200_params& p);
void cggi64w_build_tables(const tfhe_b200_params& p, std::vector<u64>& twU, std::vector<u64>& twB, std::vector<u64>& twC);
cudaError_t launch_br_cggi64w(const BRCommon& c, const CGGI64WTables& t, cudaStream_t s, int sm_count = 0, int group = 0);

// Serialized-key reader (serial_reader.cu): an index over OpenFHE's cereal portable-binary stream of a RingGSWACCKey /
// LWESwitchingKey -- where every polynomial / key-switching row lives in the byte stream -- plus gather functions that
// produce chunks of the flat element order tfhe_b200_setup takes.  No OpenFHE object is built.
struct SerializedAccKey {
    const unsigned char* base = nullptr;
    u64 dim[3] = {0, 0, 0};          // m_key dimensions ([1][2][n] for CGGI, [n][baseR][digitsR] for DM)
    u64 rows = 0;                    // RGSW rows per evaluation key (d)
    u64 N = 0, Q = 0, psi = 0;
    std::vector<size_t> poly_off;    // byte offset of the N raw coefficients of polynomial p ((size_t)-1: null entry)
    void gather(u64* dst, size_t off_words, size_t cnt_words) const;
};
struct SerializedSwitchKey {
    const unsigned char* base = nullptr;
    u64 N = 0, baseKS = 0, dKS = 0, n = 0, qKS = 0;
    std::vector<size_t> rowA_off;    // [N * baseKS * dKS] byte offset of the n mask words
    std::vector<size_t> rowB_off;    // [N * baseKS]       byte offset of the dKS b words
    void gather_rows(u64* dst, size_t row0, size_t nrows) const;
};
int index_serialized_acc_key(const void* data, size_t bytes, SerializedAccKey* out, std::string* err);
int index_serialized_switch_key(const void* data, size_t bytes, SerializedSwitchKey* out, std::string* err);

// AP/DM on the N = 2048 rings (br_dm64w.cu): same tables as the wide CGGI kernel, key from bk_relayout_dm64_kernel
bool dm64w_supported(const tfhe_b200_params& p);
cudaError_t launch_br_dm64w(const BRCommon& c, const CGGI64WTables& t, cudaStream_t s, int sm_count = 0, int group = 0);

// GPU key generation (keygen.cu)
int keygen_device(const tfhe_b200_params& p, const signed char* sk_lwe, const signed char* sk_ring,
                  const unsigned char key[32], int device, u64* bk_dev, u64* ksk_dev);
const char* keygen_last_error();

// LWE-side kernels (lwe_kernels.cu)
struct KSArgs {
    u32 N, n, baseKS, dKS, row_stride;  // row_stride in entries (padded to 16 B)
    u64 Q, qKS, fmod;
    int batch;
    const u64* ext;    // [batch][N+1] mod Q
    u64* out;          // [batch][n+1] mod fmod
    const void* ksk;   // [N][baseKS][dKS][row_stride] entries of ksk_bytes each
    int ksk_bytes;     // 2, 4 or 8
    // small batches: the N*dKS gathered rows of a ciphertext are split over `splits` CTAs that add their column sums
    // into `partial` ([batch][row_stride] u64, zeroed by the launcher); a second tiny kernel finishes (b - sum, ModSwitch)
    u64* partial = nullptr;
    int splits = 1;
    int sm_count = 0;
};
// bytes of KSArgs::partial scratch needed for split launches of up to max_batch ciphertexts
size_t mkmswitch_partial_bytes(u32 row_stride, int max_batch);
cudaError_t launch_mkmswitch(const KSArgs& a, cudaStream_t s);

// out = ((sx*x + sy*y) mod m, b += cb) then optionally reduced mod m2 (SetModulus); words = n+1
cudaError_t launch_lwe_affine(u64* out, const u64* x, const u64* y, int sx, int sy, int dbl, u64 cb, u64 m, u64 m2,
                              int batch, u32 words, cudaStream_t s);
// ModSwitch (lwe-pke.cpp:204-215) elementwise
cudaError_t launch_mod_switch(u64* out, const u64* in, u64 from_mod, u64 to_mod, size_t count, cudaStream_t s);
// strided copy / reduce helpers
cudaError_t launch_copy_mod(u64* out, size_t out_stride, const u64* in, size_t in_stride, u64 m, int batch, u32 words,
                            cudaStream_t s);
// closed-form step tables of EvalFunc / EvalFloor / EvalSign, generated on the device (lwe_kernels.cu)
enum StepTable : int { STEP_HALF = 0, STEP_FLOOR2 = 1 };
cudaError_t launch_step_table(u64* tab, int kind, u64 len, u64 a, u64 b, cudaStream_t s);
cudaError_t launch_lut_expand(u64* out, const u64* lut, u64 q, u64 tab_len, int mode, int batch, cudaStream_t s);
size_t mul_matrix_scratch_bytes(int in, int outc, u32 words, u64 modulus);
cudaError_t launch_mul_matrix(u64* out, const u64* ct, const i64* M, int in, int outc, u32 words, u64 modulus,
                              void* scratch, cudaStream_t s);

// key re-encoding kernels (lwe_kernels.cu)
// dst[perm(idx)] = to_mont(src[idx]) * Ninv ; layouts described in capi.cu
template <typename T>
cudaError_t launch_bk_convert_generic(T* dst, const u64* src, size_t count, ModCtx<T> mod, T ninvM2, cudaStream_t s);
cudaError_t launch_ksk_convert(void* dst, int bytes, u32 row_stride, const u64* src, size_t rows, u32 words,
                               cudaStream_t s);

}  // namespace tfhe_b200
