/* tfhe_b200.h -- C ABI of the B200-native batched TFHE/FHEW bootstrapping engine (libtfhe_b200.so).
 *
 * This is the drop-in boundary for the hot path of eric070021/TFHE-GPU (paths below are relative to
 * /root/reference/src/binfhe).  Every entry point replaces one reference operator; the reference-side
 * binding (a ~150-line C++ shim defining lbcrypto::GPUFFTBootstrap::* / GPULWEOperation::*) is
 * tfhe_gpu_b200/adapter/binfhe_b200_shim.cpp and is described in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success or a negative tfhe_b200_status; tfhe_b200_last_error() gives the
 *     message (the reference prints and std::exit()s instead: lib/bootstrapping.cu:28-37)
 *   - all integers are little-endian uint64_t, exactly the memory image of OpenFHE's NativeInteger /
 *     NativeVector (&v[0]), so the shim can pass OpenFHE buffers without conversion
 *   - an LWE ciphertext of dimension m is (m+1) words: a[0..m-1], b; batches are dense row-major
 *   - `space` says whether ciphertext / LUT / matrix pointers are host memory (TFHE_B200_HOST; copied through
 *     pinned staging inside the call) or device memory on the handle's first GPU (TFHE_B200_DEVICE)
 *   - calls are synchronous (like the reference: lib/bootstrapping.cu:1642-1646) and re-entrant per handle
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with TFHE_B200_ENODEV
 */
#ifndef TFHE_B200_H
#define TFHE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum tfhe_b200_status {
    TFHE_B200_OK      = 0,
    TFHE_B200_EINVAL  = -1, /* bad argument (OPENFHE_THROW(openfhe_error / config_error) in the reference)      */
    TFHE_B200_ENODEV  = -2, /* no usable CUDA device                                                             */
    TFHE_B200_ECUDA   = -3, /* CUDA runtime error, message has the details                                       */
    TFHE_B200_ENOTSUP = -4, /* parameter combination the engine has no kernel for (reference: exit(1))           */
    TFHE_B200_ENOMEM  = -5
} tfhe_b200_status;

enum { TFHE_B200_HOST = 0, TFHE_B200_DEVICE = 1 };
enum { TFHE_B200_METHOD_AP = 1, TFHE_B200_METHOD_GINX = 2 };            /* include/binfhe-constants.h:94-98 */
/* tfhe_b200_params.flags.  KEEP_GENERIC: keep the generic-layout copy of the bootstrapping key in device memory beside
 * the specialised layout, so that set_option("force_generic") can run the cross-check kernel (tests); by default it is
 * released at the end of GPUSetup (saves one copy of the key per GPU: 2.1 GB for STD128 AP). */
enum { TFHE_B200_FLAG_KEEP_GENERIC = 1 };
/* include/binfhe-constants.h:101 */
enum { TFHE_B200_OR = 0, TFHE_B200_AND, TFHE_B200_NOR, TFHE_B200_NAND, TFHE_B200_XOR_FAST, TFHE_B200_XNOR_FAST,
       TFHE_B200_XOR, TFHE_B200_XNOR };

/* Flattened BinFHECryptoParams (replaces the reference's params_CUDA[10], lib/bootstrapping.cu:917-929). */
typedef struct tfhe_b200_params {
    uint32_t n, N;                 /* LWE dimension, ring dimension                                            */
    uint64_t q, Q, qKS;            /* LWE modulus, RLWE/RGSW prime modulus, key-switch modulus                 */
    uint32_t baseKS, dKS;          /* key-switch base and digit count ceil(log qKS / log baseKS)               */
    uint32_t baseG, digitsG, numDigitsToThrow;
    uint32_t baseR, digitsR;       /* AP/DM refresh base and digit count (0 for GINX)                          */
    uint32_t method;               /* TFHE_B200_METHOD_*                                                       */
    uint32_t flags;                /* TFHE_B200_FLAG_* (0 = defaults)                                          */
    uint64_t psi;                  /* primitive 2N-th root of unity the BK polynomials were transformed with
                                      (ILNativeParams::GetRootOfUnity(); the engine uses the same bit-reversed
                                      evaluation order as core/include/math/hal/intnat/transformnat-impl.h)   */
    uint64_t beta;                 /* BinFHEContext::GetBeta(), include/binfhecontext.h:348                    */
} tfhe_b200_params;

typedef struct tfhe_b200_handle tfhe_b200_handle;

/* Per-call phase timings in milliseconds (CUDA events), filled when stats != NULL.  Replaces the commented-out
 * std::chrono blocks of the reference (lib/bootstrapping.cu:1610,1678-1680,1829-1831,1918-1920). */
typedef struct tfhe_b200_stats {
    float h2d_ms, prep_ms, blind_rotate_ms, keyswitch_ms, d2h_ms, total_ms;
    uint32_t bootstraps;           /* number of blind rotations per input performed by the call                */
    uint32_t kernel_launches;      /* kernels launched by the call on all GPUs                                 */
} tfhe_b200_stats;

/* --------------------------------------------------------------------------------------------------------
 * Lifetime.  Replaces GPUFFTBootstrap::GPUSetup/GPUClean (include/bootstrapping.cuh:104-109,
 * lib/bootstrapping.cu:725-1110) and GPULWEOperation::GPUSetup/GPUClean (include/lwe-operation.cuh:57-62).
 *
 *   bk : bootstrapping key in the reference's own element order, EVALUATION format, one u64 per coefficient
 *        GINX: [key(2)][i(n)][l(d)][j(2)][N], d = 2*(digitsG-numDigitsToThrow)   == (*BSkey)[0][key][i]->[l][j]
 *        AP  : [i(n)][a0(baseR)][k(digitsR)][l(d)][j(2)][N], d = 2*digitsG       == (*BSkey)[i][a0][k]->[l][j]
 *   ksk: [i(N)][a0(baseKS)][j(dKS)][n+1]  (A row, then B)  == the flattening of lib/bootstrapping.cu:961-975
 *   key_space: TFHE_B200_HOST or TFHE_B200_DEVICE (device = already broadcast to this process's GPU, e.g. by NCCL)
 *   num_gpus : GPUs driven by THIS process (0 = all visible), devices first_device .. first_device+num_gpus-1;
 *              keys are uploaded once and replicated peer-to-peer, the batch is split evenly, no collectives
 * -------------------------------------------------------------------------------------------------------- */
int tfhe_b200_setup(const tfhe_b200_params* params, const uint64_t* bk, size_t bk_words, const uint64_t* ksk,
                    size_t ksk_words, int key_space, int first_device, int num_gpus, tfhe_b200_handle** out);
/* SURVEY.md section 8(f) rank 2 -- GPUSetup straight from OpenFHE's serialized keys.  bk_stream / ksk_stream are the
 * byte streams Serial::SerializeToFile(path, cc.GetRefreshKey() / cc.GetSwitchKey(), SerType::BINARY) writes (cereal
 * portable binary; examples/boolean-serial-binary.cpp:76-88, core/include/utils/serial.h:99-115), e.g. the mmap'ed files.
 * The reference deserialises them into a tree of shared_ptr<RingGSWEvalKeyImpl> / NativePoly objects
 * (boolean-serial-binary.cpp:115-131) that its GPUSetup then copies coefficient by coefficient
 * (lib/bootstrapping.cu:933-975); here the streams are indexed once and the raw coefficients are gathered from them
 * directly into the pinned upload chunks -- no OpenFHE object is built and no flat copy of the keys exists on the host.
 * `params` must describe the same key set (checked: dimensions, N, Q, psi, n, qKS, baseKS, dKS).  The resulting handle
 * is indistinguishable from one made by tfhe_b200_setup on the flattened keys. */
int tfhe_b200_setup_from_serialized(const tfhe_b200_params* params, const void* bk_stream, size_t bk_bytes,
                                    const void* ksk_stream, size_t ksk_bytes, int first_device, int num_gpus,
                                    tfhe_b200_handle** out);
/* What the two streams hold (host only, no GPU needed): m_key dimensions ([1][2][n] for CGGI, [n][baseR][digitsR] for
 * DM), RGSW rows per evaluation key, ring dimension, moduli, the root of unity, and the sizes of the flat arrays. */
typedef struct tfhe_b200_serialized_info_t {
    uint64_t bk_dim[3], bk_rows, N, Q, psi;
    uint64_t ks_N, baseKS, dKS, n, qKS;
    size_t bk_words, ksk_words;
} tfhe_b200_serialized_info_t;
int tfhe_b200_serialized_info(const void* bk_stream, size_t bk_bytes, const void* ksk_stream, size_t ksk_bytes,
                              tfhe_b200_serialized_info_t* info);
/* The streams flattened into the element order of tfhe_b200_setup (host only; sizes from tfhe_b200_serialized_info). */
int tfhe_b200_flatten_serialized(const void* bk_stream, size_t bk_bytes, const void* ksk_stream, size_t ksk_bytes,
                                 uint64_t* bk_out, size_t bk_words, uint64_t* ksk_out, size_t ksk_words);
int tfhe_b200_clean(tfhe_b200_handle* h);
const char* tfhe_b200_last_error(void);
int tfhe_b200_num_gpus(const tfhe_b200_handle* h);
/* expected key sizes for a parameter set (words) */
size_t tfhe_b200_bk_words(const tfhe_b200_params* params);
size_t tfhe_b200_ksk_words(const tfhe_b200_params* params);
/* name of the blind-rotation kernel variant the handle dispatches to ("cggi_u32_ntt32", "generic_u64", ...) */
const char* tfhe_b200_kernel_variant(const tfhe_b200_handle* h);
/* tuning knobs (tests and measurements; results never depend on them):
 *   "force_generic" = 1  run the generic kernel instead of the specialised one (cross-check of the two implementations)
 *   "group"              ciphertexts per CTA of the specialised CGGI kernels: 0 = automatic (throughput shape for large
 *                        batches, latency shapes for batches that cannot fill the SMs), 4 / 2 = fixed, 1 = one
 *                        ciphertext per CTA (32-bit rings: the latency layout, one warp per digit polynomial)
 *   "persistent"         1 (default) = launches that would end on a partial wave of CTAs run the persistent variant of the
 *                        blind rotation (one CTA per SM, the launch's rotation steps cut into equal ranges); 0 = never
 *   "persistent_ctas"    > 0 forces the persistent variant with this many CTAs (tests: splits at arbitrary steps) */
int tfhe_b200_set_option(tfhe_b200_handle* h, const char* key, int64_t value);

/* Persistent blind rotation, host view of the schedule (no GPU needed; tests and capacity planning).  A launch of
 * `groups` CTA groups x `n` rotation steps on `ctas` persistent CTAs cuts the groups * n steps into `ctas` equal ranges;
 * range `cta` is worked off as items in the order the kernel runs them (head of the group shared with the next range,
 * whole groups, tail of the group shared with the previous range).  Returns the number of items of range `cta` (or a
 * negative status); if item >= 0, out[0..2] = (group, first step, end step) of that item.  The same code runs on the
 * device (csrc/engine.cuh: PersRange). */
int tfhe_b200_persistent_plan(uint32_t groups, uint32_t n, uint32_t ctas, uint32_t cta, int item, uint32_t out[3]);

/* --------------------------------------------------------------------------------------------------------
 * SURVEY.md section 8(f) rank 4 -- dynamic gadget base ("timeOptimization") for EvalSign / EvalDecomp.
 * BinFHEContext::BTKeyGen with timeOptimization fills m_BTKey_map with one RingGSWBTKey (BK and KSK) per gadget base
 * in {2^14, 2^18, 2^27} (lib/binfhecontext.cpp:222-247, include/rgsw-cryptoparameters.h:105-120); the scalar EvalSign /
 * EvalDecomp start with the context's own base and switch to 2^18 once the ciphertext modulus is <= 2^26 and to 2^27
 * once it is <= 2^17 (lib/binfhe-base-scheme.cpp:342-360, 411-428).  The reference's GPU path refuses such contexts
 * (lib/binfhecontext.cpp:350-353); here the extra key sets are loaded beside the one given to tfhe_b200_setup and
 * tfhe_b200_eval_sign / tfhe_b200_eval_decomp follow the same switching rule -- exactly when, like the reference
 * (`EKs.size() == 3`), three key sets are loaded.  Every other entry point keeps using the setup key.
 *   baseG: gadget base of this key set; digitsG = ceil(log Q / log baseG) as Change_BaseG computes it
 *   bk / ksk: same element order as tfhe_b200_setup, sized for (baseG, digitsG) */
int tfhe_b200_add_key_set(tfhe_b200_handle* h, uint32_t baseG, const uint64_t* bk, size_t bk_words,
                          const uint64_t* ksk, size_t ksk_words, int key_space);
/* number of key sets loaded (1 after tfhe_b200_setup) */
int tfhe_b200_num_key_sets(const tfhe_b200_handle* h);

/* --------------------------------------------------------------------------------------------------------
 * SURVEY.md section 8(f) rank 3 -- evaluation-key generation on the GPU (reference: BinFHEContext::BTKeyGen ->
 * BinFHEScheme::KeyGen, lib/binfhe-base-scheme.cpp:38-57; lib/lwe-pke.cpp:218-295; lib/rgsw-acc-cggi.cpp:43-75,213-240;
 * lib/rgsw-acc-dm.cpp:44-76,153-209).  The keys are written to DEVICE memory in the element order tfhe_b200_setup
 * expects (pass them on with key_space = TFHE_B200_DEVICE), so a 4.8 GB key-switching key never exists on the host.
 *   sk_lwe : the n ternary LWE secret coefficients in {-1, 0, 1};  sk_ring : the N ternary ring secret coefficients
 *   bk_dev : tfhe_b200_bk_words(params) u64 on `device`;  ksk_dev : tfhe_b200_ksk_words(params) u64 on `device`
 * Randomness: ChaCha20 in counter mode under a 256-bit key, independent derived keys for the public masks and the secret
 * errors, unbiased uniform residues, discrete Gaussian errors (sigma 3.19).
 *   key32  : 32 bytes of secret key material for the generator, from a cryptographically secure source (the reference
 *            seeds a BLAKE2-based PRNG from the OS, core/include/math/distributiongenerator.h:86-130); NULL = the engine
 *            draws them from the operating system (getrandom).  The same key reproduces the same evaluation keys.
 * Key generation is randomised: results are not comparable bit for bit with the reference (tests check decryption
 * correctness under these keys and the noise distribution of the key material). */
int tfhe_b200_keygen(const tfhe_b200_params* params, const int8_t* sk_lwe, const int8_t* sk_ring, const uint8_t* key32,
                     int device, uint64_t* bk_dev, uint64_t* ksk_dev);
/* TEST ONLY -- deterministic generation from a 64-bit seed (fixtures, statistics tests).  A 64-bit seed can be searched
 * exhaustively and every error term recomputed from it: keys made this way offer at most 64 bits of security whatever
 * the parameter set claims.  Never use it for keys that leave the machine. */
int tfhe_b200_keygen_test_seed(const tfhe_b200_params* params, const int8_t* sk_lwe, const int8_t* sk_ring,
                               uint64_t seed, int device, uint64_t* bk_dev, uint64_t* ksk_dev);

/* --------------------------------------------------------------------------------------------------------
 * Operator-level entry points (what the reference's host code calls).
 * -------------------------------------------------------------------------------------------------------- */
/* GPUFFTBootstrap::EvalAcc_CUDA (include/bootstrapping.cuh:111-124, lib/bootstrapping.cu:1139-1853).
 * a: [batch][n] mask, modulus ct_mod; acc: [batch][2][N] in COEFFICIENT format on entry and exit; on exit the
 * a-polynomial is already transposed (a(X) -> a(X^-1)), exactly as the reference kernel leaves it. */
int tfhe_b200_eval_acc(tfhe_b200_handle* h, int batch, const uint64_t* a, uint64_t ct_mod, uint64_t* acc, int space,
                       tfhe_b200_stats* stats);
/* GPUFFTBootstrap::MKMSwitch_CUDA (include/bootstrapping.cuh:126-136, lib/bootstrapping.cu:73-118,1855-1935):
 * ModSwitch Q->qKS, KeySwitch, ModSwitch qKS->fmod.  in: [batch][N+1] mod Q, out: [batch][n+1] mod fmod. */
int tfhe_b200_mkmswitch(tfhe_b200_handle* h, int batch, const uint64_t* in, uint64_t fmod, uint64_t* out, int space,
                        tfhe_b200_stats* stats);
/* GPULWEOperation::CiphertextMulMatrix_CUDA (include/lwe-operation.cuh:49-50, lib/lwe-operation.cu:50-141):
 * out[i] = sum_k ct[k] * matrix[k][i] mod modulus.  ct: [in][n+1], matrix: [in][out_cols] row-major int64,
 * out: [out_cols][n+1].  Exact integer arithmetic (the reference uses FP64 and is exact only below 2^53). */
int tfhe_b200_mul_matrix(tfhe_b200_handle* h, int in, int out_cols, const uint64_t* ct, const int64_t* matrix,
                         uint64_t modulus, uint64_t* out, int space, tfhe_b200_stats* stats);

/* --------------------------------------------------------------------------------------------------------
 * Fused batched API: one call == one BinFHEContext batched method (lib/binfhecontext.cpp:319-347 ->
 * lib/binfhe-base-scheme.cpp:598-1277).  Ciphertexts stay on the device between the bootstraps of a call.
 * -------------------------------------------------------------------------------------------------------- */
/* BinFHEContext::EvalBinGate(gate, vector, vector); ct modulus = ct_mod (normally q). */
int tfhe_b200_eval_bin_gate(tfhe_b200_handle* h, int gate, int batch, const uint64_t* ct1, const uint64_t* ct2,
                            uint64_t ct_mod, uint64_t* out, int space, tfhe_b200_stats* stats);

/* The same for callers that hold every ciphertext as a separate object -- std::vector<LWECiphertext>, the reference's
 * own argument type (lib/binfhecontext.cpp:323-325): a1[i] / a2[i] / a_out[i] point to the n mask words of ciphertext i
 * (&ct->GetA()[0]; a NativeVector is a flat u64 array), b1 / b2 / b_out are dense arrays of the b words; the output
 * objects are allocated by the caller.  Host memory only.  The gather of chunk k+1 into the handle's pinned staging and
 * the scatter of chunk k-1 back into the output objects run inside the chunk pipeline while the GPU works on chunk k,
 * so the flattening the reference does up front (lib/bootstrapping.cu:1562-1600) costs no wall time. */
int tfhe_b200_eval_bin_gate_v(tfhe_b200_handle* h, int gate, int batch, const uint64_t* const* a1, const uint64_t* b1,
                              const uint64_t* const* a2, const uint64_t* b2, uint64_t ct_mod, uint64_t* const* a_out,
                              uint64_t* b_out, tfhe_b200_stats* stats);

/* --------------------------------------------------------------------------------------------------------
 * SURVEY.md section 8(f) rank 1 -- gate-graph submission: a whole netlist of binary gates over a batch in ONE call,
 * every intermediate ciphertext device-resident.  Replaces the host loop a reference user writes around
 * BinFHEContext::EvalBinGate / EvalNOT (one std::vector<LWECiphertext> round trip per gate; the glue of
 * lib/binfhe-base-scheme.cpp:598-677 and the AND/AND/OR expansion of XOR at :617-640 run on the host there).
 * Semantics: identical, bit for bit, to evaluating the nodes one by one in order with EvalBinGate / EvalNOT.
 *   wires 0 .. n_inputs-1 are the inputs, wire n_inputs + g is the output of node g; nodes are in topological
 *   order (in0, in1 < n_inputs + g); gate is a TFHE_B200_* binary gate or TFHE_B200_NOT (in1 ignored).
 *   inputs: [n_inputs][batch][n+1]; out: [n_outputs][batch][n+1] = the wires listed in output_wires.
 * Independent nodes of the same kind and depth are bootstrapped by ONE launch over (nodes x batch) ciphertexts, so a
 * small batch still fills the GPU.  Like every batched call the batch is sharded over the handle's GPUs. */
enum { TFHE_B200_NOT = 8 };
typedef struct tfhe_b200_gate_node {
    int32_t gate;
    int32_t in0;
    int32_t in1;
} tfhe_b200_gate_node;
int tfhe_b200_eval_circuit(tfhe_b200_handle* h, int batch, int n_inputs, const uint64_t* inputs, uint64_t ct_mod,
                           int n_nodes, const tfhe_b200_gate_node* nodes, int n_outputs, const int32_t* output_wires,
                           uint64_t* out, int space, tfhe_b200_stats* stats);
/* BinFHEScheme::BootstrapFunc(vector) (lib/binfhe-base-scheme.cpp:1194-1211, 1260-1277): table[x] = f(x, ct_mod,
 * fmod) for x < ct_mod; per_ct != 0 => one table per ciphertext ([batch][ct_mod]).  Output modulus fmod. */
int tfhe_b200_bootstrap_func(tfhe_b200_handle* h, int batch, const uint64_t* ct, uint64_t ct_mod,
                             const uint64_t* table, int per_ct, uint64_t fmod, uint64_t* out, int space,
                             tfhe_b200_stats* stats);
/* BinFHEContext::EvalFunc(vector, LUT) / (vector, LUT_vec): lut has ct_mod entries (per_ct: [batch][ct_mod]). */
int tfhe_b200_eval_func(tfhe_b200_handle* h, int batch, const uint64_t* ct, uint64_t ct_mod, const uint64_t* lut,
                        size_t lut_len, int per_ct, uint64_t* out, int space, tfhe_b200_stats* stats);
/* BinFHEContext::EvalFloor(vector, roundbits) */
int tfhe_b200_eval_floor(tfhe_b200_handle* h, int batch, const uint64_t* ct, uint64_t ct_mod, uint32_t roundbits,
                         uint64_t* out, int space, tfhe_b200_stats* stats);
/* BinFHEContext::EvalSign(vector): output modulus q */
int tfhe_b200_eval_sign(tfhe_b200_handle* h, int batch, const uint64_t* ct, uint64_t ct_mod, uint64_t* out,
                        int space, tfhe_b200_stats* stats);
/* BinFHEContext::EvalDecomp(vector): out [batch][max_digits][n+1]; returns the digit count (>0) or a negative
 * status; out_mods[k] (host memory) receives the modulus of digit k. */
int tfhe_b200_eval_decomp(tfhe_b200_handle* h, int batch, const uint64_t* ct, uint64_t ct_mod, int max_digits,
                          uint64_t* out, uint64_t* out_mods, int space, tfhe_b200_stats* stats);

#ifdef __cplusplus
}
#endif
#endif /* TFHE_B200_H */
