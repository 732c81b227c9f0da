// TEST INFRASTRUCTURE ONLY.
//
// Flat C wrapper (for ctypes / plain C++ harnesses) around the UNMODIFIED reference CPU implementation
// (OpenFHE 1.0.4 + TFHE-GPU host code, compiled in place from /root/reference by oracle/Makefile into
// oracle/_ref/libtfhe_ref.so).  It exposes:
//   * context creation for the parameter sets BASELINE.json names (binfhecontext.cpp:42-181),
//   * key generation + export of BK / KSK / secret key as flat little-endian u64 arrays in the SAME element
//     order the reference's own GPUSetup flattens them (bootstrapping.cu:933-975),
//   * encrypt / decrypt,
//   * the reference's SCALAR CPU evaluation API (binfhecontext.cpp:248-289 -> binfhe-base-scheme.cpp:58-592),
//     looped over a batch with OpenMP -- this is the bit-exact oracle and the `cpu_baseline` ("reference"),
//   * stage-level entry points (NTT, blind rotation, mod switch, key switch) for stage KATs.
//   * the reference's BATCHED API (binfhecontext.cpp:319-365); in libtfhe_ref.so these throw (GPU stubs), in
//     libtfhe_ref_dropin.so they run on our engine through tfhe_gpu_b200/adapter/binfhe_b200_shim.cpp.
//
// Ciphertext wire format everywhere: (n+1) u64 per ciphertext = a[0..n-1], b.  The modulus is passed beside it.
#include "binfhecontext-ser.h"   // binfhecontext.h + cereal registration of the key types (serialization fixtures)
#ifdef TFHE_B200_DROPIN
#include "binfhe_b200.hpp"   // fused C++ adapter (only in the drop-in build, which links libtfhe_b200.so)
#endif

#include <chrono>
#include <omp.h>
#include <cstring>
#include <string>

using namespace lbcrypto;
typedef uint64_t u64;

struct RefCtx {
    BinFHEContext cc;
    LWEPrivateKey sk;
    bool has_keys = false;
    std::string err;
};

static thread_local std::string g_err;

#define REF_TRY try {
#define REF_CATCH(ret)                 \
    }                                  \
    catch (const std::exception& e) {  \
        g_err = e.what();              \
        return ret;                    \
    }

static LWECiphertext make_ct(const u64* p, uint32_t n, u64 mod) {
    NativeVector a(n, NativeInteger(mod));
    for (uint32_t i = 0; i < n; i++)
        a[i] = NativeInteger(p[i]);
    return std::make_shared<LWECiphertextImpl>(std::move(a), NativeInteger(p[n]));
}
static void put_ct(ConstLWECiphertext ct, u64* p) {
    uint32_t n = ct->GetLength();
    for (uint32_t i = 0; i < n; i++)
        p[i] = ct->GetA(i).ConvertToInt();
    p[n] = ct->GetB().ConvertToInt();
}

extern "C" {

const char* ref_last_error() {
    return g_err.c_str();
}

// kind 0: named set + method            (binfhecontext.cpp:114-181)   a0=BINFHE_PARAMSET a1=BINFHE_METHOD
// kind 1: functional bootstrapping set  (binfhecontext.cpp:51-112)    a0=set a1=arbFunc a2=logQ a3=N a4=baseG a5=numDigitsToThrow
// kind 2: custom                        (binfhecontext.cpp:42-49)     a0=n a1=N a2=q a3=Q a4=baseKS a5=baseG a6=baseR a7=method
void* ref_ctx_create(int kind, const u64* a) {
    REF_TRY
    auto* c = new RefCtx();
    if (kind == 0)
        c->cc.GenerateBinFHEContext((BINFHE_PARAMSET)a[0], (BINFHE_METHOD)a[1]);
    else if (kind == 1 || kind == 3)  // kind 3: timeOptimization = true (three-key map, dynamic gadget base)
        c->cc.GenerateBinFHEContext((BINFHE_PARAMSET)a[0], a[1] != 0, (uint32_t)a[2], (int64_t)a[3], GINX, kind == 3,
                                    (uint32_t)a[4], (uint32_t)a[5]);
    else
        c->cc.GenerateBinFHEContext((uint32_t)a[0], (uint32_t)a[1], NativeInteger(a[2]), NativeInteger(a[3]), 3.19,
                                    (uint32_t)a[4], (uint32_t)a[5], (uint32_t)a[6], (BINFHE_METHOD)a[7]);
    return c;
    REF_CATCH(nullptr)
}

void ref_ctx_destroy(void* h) {
    delete (RefCtx*)h;
}

// out[16]: n N q Q qKS baseKS dKS baseG digitsG numDigitsToThrow baseR digitsR method(1=AP,2=GINX) psi beta reserved
int ref_ctx_params(void* h, u64* out) {
    REF_TRY
    auto* c  = (RefCtx*)h;
    auto P   = c->cc.GetParams();
    auto L   = P->GetLWEParams();
    auto R   = P->GetRingGSWParams();
    u64 qKS  = L->GetqKS().ConvertToInt();
    out[0]   = L->Getn();
    out[1]   = L->GetN();
    out[2]   = L->Getq().ConvertToInt();
    out[3]   = L->GetQ().ConvertToInt();
    out[4]   = qKS;
    out[5]   = L->GetBaseKS();
    out[6]   = (u64)std::ceil(log((double)qKS) / log((double)L->GetBaseKS()));  // lwe-pke.cpp:305
    out[7]   = R->GetBaseG();
    out[8]   = R->GetDigitsG();
    out[9]   = R->GetNumDigitsToThrow();
    out[10]  = R->GetBaseR();
    out[11]  = R->GetDigitsR().size();
    out[12]  = (u64)R->GetMethod();
    out[13]  = R->GetPolyParams()->GetRootOfUnity().ConvertToInt();
    out[14]  = c->cc.GetBeta().ConvertToInt();
    out[15]  = 0;
    return 0;
    REF_CATCH(-1)
}

int ref_keygen(void* h) {
    REF_TRY
    auto* c = (RefCtx*)h;
    c->sk   = c->cc.KeyGen();
    c->cc.BTKeyGen(c->sk);
    c->has_keys = true;
    return 0;
    REF_CATCH(-1)
}

// timeOptimization contexts: gadget bases of m_BTKey_map (binfhecontext.cpp:222-247); returns their count
int ref_key_map_bases(void* h, u64* out, int max) {
    REF_TRY
    auto* c  = (RefCtx*)h;
    auto map = c->cc.GetBTKeyMap();
    int k    = 0;
    for (auto& kv : *map)
        if (k < max)
            out[k++] = kv.first;
    return (int)map->size();
    REF_CATCH(-1)
}
// make the key set of `baseG` the context's current one (BTKeyLoad + Change_BaseG), so that ref_ctx_params /
// ref_export_bk / ref_export_ksk describe it; select the original base again before evaluating
int ref_select_key(void* h, u64 baseG) {
    REF_TRY
    auto* c  = (RefCtx*)h;
    auto map = c->cc.GetBTKeyMap();
    auto it  = map->find((uint32_t)baseG);
    if (it == map->end())
        throw std::runtime_error("no such key in the map");
    c->cc.GetParams()->GetRingGSWParams()->Change_BaseG((uint32_t)baseG);
    c->cc.BTKeyLoad(it->second);
    return 0;
    REF_CATCH(-1)
}

// number of u64 words of the flattened bootstrapping key
//   GINX: [key(2)][i(n)][l(d)][j(2)][N]                 d = 2*(digitsG - numDigitsToThrow)   (rgsw-acc-cggi.cpp:143-155)
//   AP  : [i(n)][a0(baseR)][k(digitsR)][l(d)][j(2)][N]  (a0 = 0 rows are all-zero, never read)  (rgsw-acc-dm.cpp:102-110)
// polynomials are exported as stored: EVALUATION format (bit-reversed NTT order, transformnat-impl.h).
u64 ref_bk_words(void* h) {
    auto* c = (RefCtx*)h;
    auto R  = c->cc.GetParams()->GetRingGSWParams();
    auto L  = c->cc.GetParams()->GetLWEParams();
    u64 N = L->GetN(), n = L->Getn();
    if (R->GetMethod() == GINX) {
        u64 d = 2 * (R->GetDigitsG() - R->GetNumDigitsToThrow());
        return 2 * n * d * 2 * N;
    }
    u64 d = 2 * R->GetDigitsG();
    return n * R->GetBaseR() * R->GetDigitsR().size() * d * 2 * N;
}

int ref_export_bk(void* h, u64* out) {
    REF_TRY
    auto* c = (RefCtx*)h;
    auto R  = c->cc.GetParams()->GetRingGSWParams();
    auto L  = c->cc.GetParams()->GetLWEParams();
    u64 N = L->GetN(), n = L->Getn();
    auto BK = c->cc.GetRefreshKey();
    if (R->GetMethod() == GINX) {
        u64 d = 2 * (R->GetDigitsG() - R->GetNumDigitsToThrow());
#pragma omp parallel for collapse(2)
        for (u64 key = 0; key < 2; key++)
            for (u64 i = 0; i < n; i++) {
                const auto& ev = (*BK)[0][key][i]->GetElements();
                for (u64 l = 0; l < d; l++)
                    for (u64 j = 0; j < 2; j++) {
                        u64* dst = out + ((((key * n + i) * d + l) * 2 + j) * N);
                        const NativePoly& p = ev[l][j];
                        for (u64 k = 0; k < N; k++)
                            dst[k] = p[k].ConvertToInt();
                    }
            }
    }
    else {
        u64 d = 2 * R->GetDigitsG(), bR = R->GetBaseR(), dR = R->GetDigitsR().size();
        memset(out, 0, sizeof(u64) * n * bR * dR * d * 2 * N);
#pragma omp parallel for
        for (u64 i = 0; i < n; i++)
            for (u64 a0 = 1; a0 < bR; a0++)
                for (u64 k = 0; k < dR; k++) {
                    const auto& ev = (*BK)[i][a0][k]->GetElements();
                    for (u64 l = 0; l < d; l++)
                        for (u64 j = 0; j < 2; j++) {
                            u64* dst = out + ((((((i * bR + a0) * dR + k) * d + l) * 2 + j)) * N);
                            const NativePoly& p = ev[l][j];
                            for (u64 x = 0; x < N; x++)
                                dst[x] = p[x].ConvertToInt();
                        }
                }
    }
    return 0;
    REF_CATCH(-1)
}

// KSK flattened as [i(N)][a0(baseKS)][j(dKS)][n+1]  (A row then B), same order as bootstrapping.cu:961-975
u64 ref_ksk_words(void* h) {
    auto* c = (RefCtx*)h;
    auto L  = c->cc.GetParams()->GetLWEParams();
    u64 qKS = L->GetqKS().ConvertToInt();
    u64 dKS = (u64)std::ceil(log((double)qKS) / log((double)L->GetBaseKS()));
    return (u64)L->GetN() * L->GetBaseKS() * dKS * (L->Getn() + 1);
}

int ref_export_ksk(void* h, u64* out) {
    REF_TRY
    auto* c = (RefCtx*)h;
    auto L  = c->cc.GetParams()->GetLWEParams();
    u64 N = L->GetN(), n = L->Getn(), bKS = L->GetBaseKS();
    u64 qKS = L->GetqKS().ConvertToInt();
    u64 dKS = (u64)std::ceil(log((double)qKS) / log((double)bKS));
    auto KS = c->cc.GetSwitchKey();
    const auto& A = KS->GetElementsA();
    const auto& B = KS->GetElementsB();
#pragma omp parallel for
    for (u64 i = 0; i < N; i++)
        for (u64 a0 = 0; a0 < bKS; a0++)
            for (u64 j = 0; j < dKS; j++) {
                u64* dst = out + (((i * bKS + a0) * dKS + j) * (n + 1));
                for (u64 k = 0; k < n; k++)
                    dst[k] = A[i][a0][j][k].ConvertToInt();
                dst[n] = B[i][a0][j].ConvertToInt();
            }
    return 0;
    REF_CATCH(-1)
}

// secret key as residues mod qKS (ternary: 0, 1, qKS-1)
int ref_export_sk(void* h, u64* out) {
    REF_TRY
    auto* c = (RefCtx*)h;
    const auto& s = c->sk->GetElement();
    for (uint32_t i = 0; i < s.GetLength(); i++)
        out[i] = s[i].ConvertToInt();
    return 0;
    REF_CATCH(-1)
}

int ref_encrypt(void* h, int64_t m, u64 p, u64 mod, u64* ct_out) {
    REF_TRY
    auto* c = (RefCtx*)h;
    auto ct = c->cc.Encrypt(c->sk, m, FRESH, p, NativeInteger(mod));
    put_ct(ct, ct_out);
    return 0;
    REF_CATCH(-1)
}

int64_t ref_decrypt(void* h, const u64* ct, u64 mod, u64 p) {
    REF_TRY
    auto* c  = (RefCtx*)h;
    uint32_t n = c->cc.GetParams()->GetLWEParams()->Getn();
    LWEPlaintext r;
    c->cc.Decrypt(c->sk, make_ct(ct, n, mod), &r, p);
    return r;
    REF_CATCH(-1)
}

// ---------------------------------------------------------------------------------------------------------
// Scalar CPU API (the oracle), looped over the batch with OpenMP.
// ---------------------------------------------------------------------------------------------------------
int ref_eval_bin_gate(void* h, int gate, int batch, const u64* ct1, const u64* ct2, u64 mod, u64* out) {
    REF_TRY
    auto* c    = (RefCtx*)h;
    uint32_t n = c->cc.GetParams()->GetLWEParams()->Getn();
    int bad    = 0;
#pragma omp parallel for schedule(dynamic)
    for (int s = 0; s < batch; s++) {
        try {
            auto r = c->cc.EvalBinGate((BINGATE)gate, make_ct(ct1 + (size_t)s * (n + 1), n, mod),
                                       make_ct(ct2 + (size_t)s * (n + 1), n, mod));
            put_ct(r, out + (size_t)s * (n + 1));
        }
        catch (...) {
#pragma omp atomic
            bad++;
        }
    }
    return bad ? -1 : 0;
    REF_CATCH(-1)
}

int ref_eval_func(void* h, int batch, const u64* ct, u64 mod, const u64* lut, u64 lut_len, u64* out) {
    REF_TRY
    auto* c    = (RefCtx*)h;
    uint32_t n = c->cc.GetParams()->GetLWEParams()->Getn();
    std::vector<NativeInteger> LUT(lut_len);
    for (u64 i = 0; i < lut_len; i++)
        LUT[i] = NativeInteger(lut[i]);
    int bad = 0;
#pragma omp parallel for schedule(dynamic)
    for (int s = 0; s < batch; s++) {
        try {
            auto r = c->cc.EvalFunc(make_ct(ct + (size_t)s * (n + 1), n, mod), LUT);
            put_ct(r, out + (size_t)s * (n + 1));
        }
        catch (...) {
#pragma omp atomic
            bad++;
        }
    }
    return bad ? -1 : 0;
    REF_CATCH(-1)
}

int ref_eval_floor(void* h, int batch, const u64* ct, u64 mod, uint32_t roundbits, u64* out) {
    REF_TRY
    auto* c    = (RefCtx*)h;
    uint32_t n = c->cc.GetParams()->GetLWEParams()->Getn();
    int bad    = 0;
#pragma omp parallel for schedule(dynamic)
    for (int s = 0; s < batch; s++) {
        try {
            auto r = c->cc.EvalFloor(make_ct(ct + (size_t)s * (n + 1), n, mod), roundbits);
            put_ct(r, out + (size_t)s * (n + 1));
        }
        catch (...) {
#pragma omp atomic
            bad++;
        }
    }
    return bad ? -1 : 0;
    REF_CATCH(-1)
}

// EvalSign / EvalDecomp temporarily mutate the shared RGSW params (Change_BaseG, binfhe-base-scheme.cpp:331,369),
// which is a no-op when timeOptimization == false (single-key map), so the OpenMP loop is safe.
int ref_eval_sign(void* h, int batch, const u64* ct, u64 mod, u64* out) {
    REF_TRY
    auto* c    = (RefCtx*)h;
    uint32_t n = c->cc.GetParams()->GetLWEParams()->Getn();
    int bad    = 0;
    // with the three-key map EvalSign calls Change_BaseG on the shared RingGSWCryptoParams: not thread-safe
    const bool par = c->cc.GetBTKeyMap()->size() < 3;
#pragma omp parallel for schedule(dynamic) if (par)
    for (int s = 0; s < batch; s++) {
        try {
            auto r = c->cc.EvalSign(make_ct(ct + (size_t)s * (n + 1), n, mod));
            put_ct(r, out + (size_t)s * (n + 1));
        }
        catch (...) {
#pragma omp atomic
            bad++;
        }
    }
    return bad ? -1 : 0;
    REF_CATCH(-1)
}

// out: [batch][max_digits][n+1]; out_mods: [max_digits] moduli of the digits; returns #digits or -1
int ref_eval_decomp(void* h, int batch, const u64* ct, u64 mod, int max_digits, u64* out, u64* out_mods) {
    REF_TRY
    auto* c    = (RefCtx*)h;
    uint32_t n = c->cc.GetParams()->GetLWEParams()->Getn();
    int bad = 0, nd = 0;
    const bool par = c->cc.GetBTKeyMap()->size() < 3;
#pragma omp parallel for schedule(dynamic) if (par)
    for (int s = 0; s < batch; s++) {
        try {
            auto r = c->cc.EvalDecomp(make_ct(ct + (size_t)s * (n + 1), n, mod));
            if ((int)r.size() > max_digits)
                throw std::runtime_error("too many digits");
            for (size_t k = 0; k < r.size(); k++) {
                put_ct(r[k], out + ((size_t)s * max_digits + k) * (n + 1));
                if (s == 0)
                    out_mods[k] = r[k]->GetModulus().ConvertToInt();
            }
            if (s == 0)
                nd = (int)r.size();
        }
        catch (...) {
#pragma omp atomic
            bad++;
        }
    }
    return bad ? -1 : nd;
    REF_CATCH(-1)
}

// ---------------------------------------------------------------------------------------------------------
// Stage-level entry points (for KATs)
// ---------------------------------------------------------------------------------------------------------
// forward / inverse negacyclic NTT of one polynomial with the context's (N, Q, psi)   transformnat-impl.h
int ref_ntt(void* h, int inverse, u64* poly) {
    REF_TRY
    auto* c = (RefCtx*)h;
    auto R  = c->cc.GetParams()->GetRingGSWParams();
    u64 N   = R->GetN();
    NativePoly p(R->GetPolyParams(), inverse ? Format::EVALUATION : Format::COEFFICIENT, true);
    for (u64 i = 0; i < N; i++)
        p[i] = NativeInteger(poly[i]);
    p.SetFormat(inverse ? Format::COEFFICIENT : Format::EVALUATION);
    for (u64 i = 0; i < N; i++)
        poly[i] = p[i].ConvertToInt();
    return 0;
    REF_CATCH(-1)
}

// signed digit decomposition (rgsw-acc.cpp:57-111): in [2][N] coefficient form, out [d][N]
int ref_signed_digit_decompose(void* h, const u64* in, u64* out) {
    REF_TRY
    auto* c = (RefCtx*)h;
    auto R  = c->cc.GetParams()->GetRingGSWParams();
    u64 N = R->GetN(), d = 2 * (R->GetDigitsG() - R->GetNumDigitsToThrow());
    std::vector<NativePoly> ct(2), dct(d);
    for (int j = 0; j < 2; j++) {
        ct[j] = NativePoly(R->GetPolyParams(), Format::COEFFICIENT, true);
        for (u64 k = 0; k < N; k++)
            ct[j][k] = NativeInteger(in[j * N + k]);
    }
    for (u64 l = 0; l < d; l++)
        dct[l] = NativePoly(R->GetPolyParams(), Format::COEFFICIENT, true);
    struct Acc : public RingGSWAccumulator {
        void EvalAcc(const std::shared_ptr<RingGSWCryptoParams>, const RingGSWACCKey, RLWECiphertext&,
                     const NativeVector&, std::string, uint64_t) const override {}
        RingGSWACCKey KeyGenAcc(const std::shared_ptr<RingGSWCryptoParams>, const NativePoly&,
                                ConstLWEPrivateKey) const override {
            return nullptr;
        }
        using RingGSWAccumulator::SignedDigitDecompose;
    } acc;
    acc.SignedDigitDecompose(R, ct, dct);
    for (u64 l = 0; l < d; l++)
        for (u64 k = 0; k < N; k++)
            out[l * N + k] = dct[l][k].ConvertToInt();
    return 0;
    REF_CATCH(-1)
}

// Blind rotation with the contract of GPUFFTBootstrap::EvalAcc_CUDA (bootstrapping.cuh:111-124): acc is
// [batch][2][N] in COEFFICIENT format on entry and exit, and the a-polynomial is already transposed on exit.
// Computed with the reference CPU accumulators (rgsw-acc-cggi.cpp:143-155 / rgsw-acc-dm.cpp:80-110, mode "NTT").
int ref_eval_acc(void* h, int batch, const u64* a, u64 mod, u64* acc) {
    REF_TRY
    auto* c    = (RefCtx*)h;
    auto R     = c->cc.GetParams()->GetRingGSWParams();
    uint32_t n = c->cc.GetParams()->GetLWEParams()->Getn();
    u64 N      = R->GetN();
    auto BK    = c->cc.GetRefreshKey();
    auto method = R->GetMethod();
    int bad = 0;
#pragma omp parallel for schedule(dynamic)
    for (int s = 0; s < batch; s++) {
        try {
            std::vector<NativePoly> res(2);
            for (int j = 0; j < 2; j++) {
                res[j] = NativePoly(R->GetPolyParams(), Format::COEFFICIENT, true);
                for (u64 k = 0; k < N; k++)
                    res[j][k] = NativeInteger(acc[((size_t)s * 2 + j) * N + k]);
                res[j].SetFormat(Format::EVALUATION);
            }
            RLWECiphertext A = std::make_shared<RLWECiphertextImpl>(std::move(res));
            NativeVector av(n, NativeInteger(mod));
            for (uint32_t i = 0; i < n; i++)
                av[i] = NativeInteger(a[(size_t)s * n + i]);
            if (method == GINX)
                RingGSWAccumulatorCGGI().EvalAcc(R, BK, A, av, "NTT", 0);
            else
                RingGSWAccumulatorDM().EvalAcc(R, BK, A, av, "NTT", 0);
            auto& el = A->GetElements();
            el[0]    = el[0].Transpose();
            el[0].SetFormat(Format::COEFFICIENT);
            el[1].SetFormat(Format::COEFFICIENT);
            for (int j = 0; j < 2; j++)
                for (u64 k = 0; k < N; k++)
                    acc[((size_t)s * 2 + j) * N + k] = el[j][k].ConvertToInt();
        }
        catch (...) {
#pragma omp atomic
            bad++;
        }
    }
    return bad ? -1 : 0;
    REF_CATCH(-1)
}

// LWEEncryptionScheme::ModSwitch (lwe-pke.cpp:204-215) on len-word ciphertexts (len = dim + 1)
int ref_mod_switch(void* h, int batch, u64 len, const u64* in, u64 from_mod, u64 to_mod, u64* out) {
    REF_TRY
    (void)h;
    LWEEncryptionScheme S;
    for (int s = 0; s < batch; s++) {
        auto r = S.ModSwitch(NativeInteger(to_mod), make_ct(in + (size_t)s * len, (uint32_t)len - 1, from_mod));
        put_ct(r, out + (size_t)s * len);
    }
    return 0;
    REF_CATCH(-1)
}

// LWEEncryptionScheme::KeySwitch (lwe-pke.cpp:299-321): in [batch][N+1] mod qKS, out [batch][n+1] mod qKS
int ref_key_switch(void* h, int batch, const u64* in, u64* out) {
    REF_TRY
    auto* c = (RefCtx*)h;
    auto L  = c->cc.GetParams()->GetLWEParams();
    u64 N = L->GetN(), n = L->Getn(), qKS = L->GetqKS().ConvertToInt();
    LWEEncryptionScheme S;
    auto KS = c->cc.GetSwitchKey();
#pragma omp parallel for
    for (int s = 0; s < batch; s++) {
        auto r = S.KeySwitch(L, KS, make_ct(in + (size_t)s * (N + 1), (uint32_t)N, qKS));
        put_ct(r, out + (size_t)s * (n + 1));
    }
    return 0;
    REF_CATCH(-1)
}

// ---------------------------------------------------------------------------------------------------------
// Batched API of the reference (binfhecontext.cpp:319-365).  Throws in libtfhe_ref.so (GPU stubs); runs on our
// engine in libtfhe_ref_dropin.so.
// ---------------------------------------------------------------------------------------------------------
int ref_gpu_setup(void* h, int num_gpus) {
    REF_TRY
    ((RefCtx*)h)->cc.GPUSetup(num_gpus);
    return 0;
    REF_CATCH(-1)
}
int ref_gpu_clean(void* h) {
    REF_TRY
    ((RefCtx*)h)->cc.GPUClean();
    return 0;
    REF_CATCH(-1)
}

static std::vector<LWECiphertext> make_vec(const u64* p, int batch, uint32_t n, u64 mod) {
    std::vector<LWECiphertext> v(batch);
#pragma omp parallel for if (batch > 512)
    for (int s = 0; s < batch; s++)
        v[s] = make_ct(p + (size_t)s * (n + 1), n, mod);
    return v;
}
static void put_vec(const std::vector<LWECiphertext>& v, u64* out, uint32_t n) {
#pragma omp parallel for if (v.size() > 512)
    for (size_t s = 0; s < v.size(); s++)
        put_ct(v[s], out + s * (n + 1));
}

int ref_batched_eval_bin_gate(void* h, int gate, int batch, const u64* ct1, const u64* ct2, u64 mod, u64* out) {
    REF_TRY
    auto* c    = (RefCtx*)h;
    uint32_t n = c->cc.GetParams()->GetLWEParams()->Getn();
    auto r     = c->cc.EvalBinGate((BINGATE)gate, make_vec(ct1, batch, n, mod), make_vec(ct2, batch, n, mod));
    put_vec(r, out, n);
    return 0;
    REF_CATCH(-1)
}
int ref_batched_eval_func(void* h, int batch, const u64* ct, u64 mod, const u64* lut, u64 lut_len, u64* out) {
    REF_TRY
    auto* c    = (RefCtx*)h;
    uint32_t n = c->cc.GetParams()->GetLWEParams()->Getn();
    std::vector<NativeInteger> LUT(lut_len);
    for (u64 i = 0; i < lut_len; i++)
        LUT[i] = NativeInteger(lut[i]);
    auto r = c->cc.EvalFunc(make_vec(ct, batch, n, mod), LUT);
    put_vec(r, out, n);
    return 0;
    REF_CATCH(-1)
}
int ref_batched_eval_floor(void* h, int batch, const u64* ct, u64 mod, uint32_t roundbits, u64* out) {
    REF_TRY
    auto* c    = (RefCtx*)h;
    uint32_t n = c->cc.GetParams()->GetLWEParams()->Getn();
    auto r     = c->cc.EvalFloor(make_vec(ct, batch, n, mod), roundbits);
    put_vec(r, out, n);
    return 0;
    REF_CATCH(-1)
}
int ref_batched_eval_sign(void* h, int batch, const u64* ct, u64 mod, u64* out) {
    REF_TRY
    auto* c    = (RefCtx*)h;
    uint32_t n = c->cc.GetParams()->GetLWEParams()->Getn();
    auto r     = c->cc.EvalSign(make_vec(ct, batch, n, mod));
    put_vec(r, out, n);
    return 0;
    REF_CATCH(-1)
}
int ref_batched_eval_decomp(void* h, int batch, const u64* ct, u64 mod, int max_digits, u64* out, u64* out_mods) {
    REF_TRY
    auto* c    = (RefCtx*)h;
    uint32_t n = c->cc.GetParams()->GetLWEParams()->Getn();
    auto r     = c->cc.EvalDecomp(make_vec(ct, batch, n, mod));
    int nd     = r.empty() ? 0 : (int)r[0].size();
    if (nd > max_digits)
        throw std::runtime_error("too many digits");
    for (size_t s = 0; s < r.size(); s++)
        for (int k = 0; k < nd; k++) {
            put_ct(r[s][k], out + (s * max_digits + k) * (n + 1));
            if (s == 0)
                out_mods[k] = r[s][k]->GetModulus().ConvertToInt();
        }
    return nd;
    REF_CATCH(-1)
}
// matrix: [in][out_cols] row-major int64; ct: [in][n+1]; out: [out_cols][n+1]
int ref_batched_mul_matrix(void* h, int in, int out_cols, const u64* ct, u64 mod, const int64_t* matrix, u64 modulus,
                           u64* out) {
    REF_TRY
    auto* c    = (RefCtx*)h;
    uint32_t n = c->cc.GetParams()->GetLWEParams()->Getn();
    std::vector<std::vector<int64_t>> M(in, std::vector<int64_t>(out_cols));
    for (int k = 0; k < in; k++)
        for (int i = 0; i < out_cols; i++)
            M[k][i] = matrix[(size_t)k * out_cols + i];
    auto r = c->cc.CiphertextMulMatrix(make_vec(ct, in, n, mod), M, modulus);
    put_vec(r, out, n);
    return 0;
    REF_CATCH(-1)
}

#ifdef TFHE_B200_DROPIN
// ---------------------------------------------------------------------------------------------------------
// Fused C++ adapter (tfhe_gpu_b200/adapter/binfhe_b200.hpp) driven with the reference's own objects: keys come from
// cc.GetRefreshKey()/GetSwitchKey(), inputs/outputs are std::vector<LWECiphertext>.
// ---------------------------------------------------------------------------------------------------------
void* fused_create(void* h, int num_gpus) {
    REF_TRY
    return new tfhe_b200::BatchedBinFHE(((RefCtx*)h)->cc, num_gpus);
    REF_CATCH(nullptr)
}
void fused_destroy(void* f) {
    delete (tfhe_b200::BatchedBinFHE*)f;
}
int fused_eval_bin_gate(void* h, void* f, int gate, int batch, const u64* ct1, const u64* ct2, u64 mod, u64* out) {
    REF_TRY
    uint32_t n = ((RefCtx*)h)->cc.GetParams()->GetLWEParams()->Getn();
    auto r = ((tfhe_b200::BatchedBinFHE*)f)->EvalBinGate((BINGATE)gate, make_vec(ct1, batch, n, mod), make_vec(ct2, batch, n, mod));
    put_vec(r, out, n);
    return 0;
    REF_CATCH(-1)
}
// Micro-benchmark of the drop-in C++ surface (measurement infrastructure): `reps` calls of
// BatchedBinFHE::EvalBinGate(gate, std::vector<LWECiphertext>, std::vector<LWECiphertext>) on ciphertext OBJECTS built once
// from the flat arrays; seconds[r] = wall time of call r, vectors in, vector out, everything a caller pays.  `out`
// receives the result of the last call (for the parity check of the caller).
int fused_bench_eval_bin_gate(void* h, void* f, int gate, int batch, const u64* ct1, const u64* ct2, u64 mod, int reps,
                              double* seconds, u64* out) {
    REF_TRY
    uint32_t n = ((RefCtx*)h)->cc.GetParams()->GetLWEParams()->Getn();
    auto v1 = make_vec(ct1, batch, n, mod), v2 = make_vec(ct2, batch, n, mod);
    auto* gpu = (tfhe_b200::BatchedBinFHE*)f;
    std::vector<LWECiphertext> r;
    for (int k = 0; k < reps; k++) {
        auto t0 = std::chrono::steady_clock::now();
        r = gpu->EvalBinGate((BINGATE)gate, v1, v2);
        seconds[k] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    put_vec(r, out, n);
    return 0;
    REF_CATCH(-1)
}
int fused_eval_func(void* h, void* f, int batch, const u64* ct, u64 mod, const u64* lut, u64 lut_len, u64* out) {
    REF_TRY
    uint32_t n = ((RefCtx*)h)->cc.GetParams()->GetLWEParams()->Getn();
    std::vector<NativeInteger> LUT(lut_len);
    for (u64 i = 0; i < lut_len; i++)
        LUT[i] = NativeInteger(lut[i]);
    auto r = ((tfhe_b200::BatchedBinFHE*)f)->EvalFunc(make_vec(ct, batch, n, mod), LUT);
    put_vec(r, out, n);
    return 0;
    REF_CATCH(-1)
}
int fused_eval_sign(void* h, void* f, int batch, const u64* ct, u64 mod, u64* out) {
    REF_TRY
    uint32_t n = ((RefCtx*)h)->cc.GetParams()->GetLWEParams()->Getn();
    auto r = ((tfhe_b200::BatchedBinFHE*)f)->EvalSign(make_vec(ct, batch, n, mod));
    put_vec(r, out, n);
    return 0;
    REF_CATCH(-1)
}
int fused_eval_decomp(void* h, void* f, int batch, const u64* ct, u64 mod, int max_digits, u64* out, u64* out_mods) {
    REF_TRY
    uint32_t n = ((RefCtx*)h)->cc.GetParams()->GetLWEParams()->Getn();
    auto r = ((tfhe_b200::BatchedBinFHE*)f)->EvalDecomp(make_vec(ct, batch, n, mod));
    int nd = r.empty() ? 0 : (int)r[0].size();
    if (nd > max_digits)
        throw std::runtime_error("too many digits");
    for (size_t s = 0; s < r.size(); s++)
        for (int k = 0; k < nd; k++) {
            put_ct(r[s][k], out + (s * max_digits + k) * (n + 1));
            if (s == 0)
                out_mods[k] = r[s][k]->GetModulus().ConvertToInt();
        }
    return nd;
    REF_CATCH(-1)
}
#endif

// Serialises the refreshing key and the key-switching key exactly as the reference's own example does
// (examples/boolean-serial-binary.cpp:76-88): Serial::SerializeToFile(..., SerType::BINARY), i.e. cereal's portable
// binary archive of RingGSWACCKey / LWESwitchingKey.  Fixtures for the serialized-key reader of the engine.
int ref_serialize_keys(void* h, const char* bk_path, const char* ksk_path) {
    REF_TRY
    auto* c = (RefCtx*)h;
    if (!Serial::SerializeToFile(bk_path, c->cc.GetRefreshKey(), SerType::BINARY))
        throw std::runtime_error("cannot write the refreshing key");
    if (!Serial::SerializeToFile(ksk_path, c->cc.GetSwitchKey(), SerType::BINARY))
        throw std::runtime_error("cannot write the switching key");
    return 0;
    REF_CATCH(-1)
}

int ref_num_threads() {
    return omp_get_max_threads();
}

// bench.py sets the thread count explicitly: launchers such as torchrun export OMP_NUM_THREADS=1
void ref_set_num_threads(int n) {
    if (n > 0)
        omp_set_num_threads(n);
}

}  // extern "C"
