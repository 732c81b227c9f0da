// Drop-in shim: defines the reference's GPU operator entry points on top of the C ABI of libtfhe_b200.so.
//
// The reference's host code (binfhecontext.cpp:349-365, binfhe-base-scheme.cpp:598-1277, rgsw-acc-cggi.cpp:196-205)
// calls six static functions that live in its two .cu files:
//
//   GPUFFTBootstrap::GPUSetup / GPUClean / EvalAcc_CUDA / MKMSwitch_CUDA      (bootstrapping.cuh:104-136)
//   GPULWEOperation::GPUSetup / GPUClean / CiphertextMulMatrix_CUDA           (lwe-operation.cuh:49-62)
//
// Linking the UNMODIFIED reference host sources against this file (instead of bootstrapping.cu + lwe-operation.cu)
// makes `cc.GPUSetup(); cc.EvalBinGate(NAND, ct1_vec, ct2_vec); ...` run on the B200-native engine.  Unlike the
// reference's FFT kernels the results are bit-identical to the scalar CPU API.  Errors become openfhe_error
// instead of std::exit (bootstrapping.cu:28-37).
//
// Built by oracle/Makefile target `dropin` (needs the reference headers, hence not part of libtfhe_b200.so).
#include "binfhecontext.h"
#include "tfhe_b200.h"

#include <cmath>
#include <mutex>

namespace lbcrypto {

namespace {
tfhe_b200_handle* g_handle = nullptr;
std::mutex g_mu;

void check(int rc, const char* what) {
    if (rc < 0)
        OPENFHE_THROW(openfhe_error, std::string(what) + ": " + tfhe_b200_last_error());
}

tfhe_b200_params flatten_params(const std::shared_ptr<BinFHECryptoParams>& params) {
    auto L = params->GetLWEParams();
    auto R = params->GetRingGSWParams();
    tfhe_b200_params p{};
    p.n = L->Getn();
    p.N = L->GetN();
    p.q = L->Getq().ConvertToInt();
    p.Q = L->GetQ().ConvertToInt();
    p.qKS = L->GetqKS().ConvertToInt();
    p.baseKS = L->GetBaseKS();
    p.dKS = (uint32_t)std::ceil(log((double)p.qKS) / log((double)p.baseKS));   // lwe-pke.cpp:305
    p.baseG = R->GetBaseG();
    p.digitsG = R->GetDigitsG();
    p.numDigitsToThrow = R->GetNumDigitsToThrow();
    p.baseR = R->GetBaseR();
    p.digitsR = (uint32_t)R->GetDigitsR().size();
    p.method = (uint32_t)R->GetMethod();
    p.psi = R->GetPolyParams()->GetRootOfUnity().ConvertToInt();
    p.beta = 128;   // binfhecontext.h:348
    return p;
}
}  // namespace

// replaces bootstrapping.cu:725-1083 (no FFT re-encoding on the host: keys go up as stored, in EVALUATION format)
void GPUFFTBootstrap::GPUSetup(const std::shared_ptr<BinFHECryptoParams> params, RingGSWACCKey BSkey,
                               LWESwitchingKey KSkey, int numGPUs) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_handle) {   // idempotent (the reference double-counts GPUs on a second call, bootstrapping.cu:762)
        tfhe_b200_clean(g_handle);
        g_handle = nullptr;
    }
    tfhe_b200_params p = flatten_params(params);
    const uint64_t N = p.N, n = p.n;
    std::vector<uint64_t> bk(tfhe_b200_bk_words(&p)), ksk(tfhe_b200_ksk_words(&p));
    if (p.method == TFHE_B200_METHOD_GINX) {
        const uint64_t d = 2 * (p.digitsG - p.numDigitsToThrow);
#pragma omp parallel for collapse(2)
        for (uint64_t key = 0; key < 2; key++)
            for (uint64_t i = 0; i < n; i++) {
                const auto& ev = (*BSkey)[0][key][i]->GetElements();
                for (uint64_t l = 0; l < d; l++)
                    for (uint64_t j = 0; j < 2; j++) {
                        uint64_t* dst = bk.data() + ((((key * n + i) * d + l) * 2 + j) * N);
                        const NativePoly& poly = ev[l][j];
                        for (uint64_t k = 0; k < N; k++)
                            dst[k] = poly[k].ConvertToInt();
                    }
            }
    }
    else {
        const uint64_t d = 2 * p.digitsG, bR = p.baseR, dR = p.digitsR;
#pragma omp parallel for
        for (uint64_t i = 0; i < n; i++)
            for (uint64_t a0 = 1; a0 < bR; a0++)
                for (uint64_t k = 0; k < dR; k++) {
                    const auto& ev = (*BSkey)[i][a0][k]->GetElements();
                    for (uint64_t l = 0; l < d; l++)
                        for (uint64_t j = 0; j < 2; j++) {
                            uint64_t* dst = bk.data() + ((((((i * bR + a0) * dR + k) * d + l) * 2 + j)) * N);
                            const NativePoly& poly = ev[l][j];
                            for (uint64_t x = 0; x < N; x++)
                                dst[x] = poly[x].ConvertToInt();
                        }
                }
    }
    {
        const auto& A = KSkey->GetElementsA();
        const auto& B = KSkey->GetElementsB();
        const uint64_t bKS = p.baseKS, dKS = p.dKS;
#pragma omp parallel for
        for (uint64_t i = 0; i < N; i++)
            for (uint64_t a0 = 0; a0 < bKS; a0++)
                for (uint64_t j = 0; j < dKS; j++) {
                    uint64_t* dst = ksk.data() + (((i * bKS + a0) * dKS + j) * (n + 1));
                    for (uint64_t k = 0; k < n; k++)
                        dst[k] = A[i][a0][j][k].ConvertToInt();
                    dst[n] = B[i][a0][j].ConvertToInt();
                }
    }
    check(tfhe_b200_setup(&p, bk.data(), bk.size(), ksk.data(), ksk.size(), TFHE_B200_HOST, 0, numGPUs, &g_handle),
          "GPUSetup");
}

void GPUFFTBootstrap::GPUClean() {
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_handle)
        tfhe_b200_clean(g_handle);
    g_handle = nullptr;
}

// replaces bootstrapping.cu:1139-1853.  acc: COEFFICIENT format in and out, a-polynomial transposed on exit.
void GPUFFTBootstrap::EvalAcc_CUDA(const std::shared_ptr<RingGSWCryptoParams> params, const std::vector<NativeVector>& a,
                                   std::shared_ptr<std::vector<RLWECiphertext>> acc, uint64_t fmod) {
    (void)fmod;
    if (!g_handle)
        OPENFHE_THROW(openfhe_error, "EvalAcc_CUDA: GPUSetup has not been called");
    const size_t batch = a.size();
    if (batch == 0 || acc->size() != batch)
        OPENFHE_THROW(openfhe_error, "EvalAcc_CUDA: empty or mismatched batch");
    const uint64_t N = params->GetN(), Q = params->GetQ().ConvertToInt();
    const size_t n = a[0].GetLength();
    const uint64_t mod = a[0].GetModulus().ConvertToInt();
    std::vector<uint64_t> fa(batch * n), facc(batch * 2 * N);
#pragma omp parallel for if (batch > 512)
    for (size_t s = 0; s < batch; s++) {
        for (size_t i = 0; i < n; i++)
            fa[s * n + i] = a[s][i].ConvertToInt();
        const auto& el = (*acc)[s]->GetElements();
        for (int j = 0; j < 2; j++)
            for (uint64_t k = 0; k < N; k++)
                facc[(s * 2 + j) * N + k] = el[j][k].ConvertToInt();
    }
    check(tfhe_b200_eval_acc(g_handle, (int)batch, fa.data(), mod, facc.data(), TFHE_B200_HOST, nullptr),
          "EvalAcc_CUDA");
    auto polyParams = params->GetPolyParams();
#pragma omp parallel for if (batch > 512)
    for (size_t s = 0; s < batch; s++) {
        std::vector<NativePoly> res(2);
        for (int j = 0; j < 2; j++) {
            NativeVector v(N, NativeInteger(Q));
            for (uint64_t k = 0; k < N; k++)
                v[k] = NativeInteger(facc[(s * 2 + j) * N + k]);
            res[j] = NativePoly(polyParams, Format::COEFFICIENT, false);
            res[j].SetValues(std::move(v), Format::COEFFICIENT);
        }
        (*acc)[s] = std::make_shared<RLWECiphertextImpl>(std::move(res));
    }
}

// replaces bootstrapping.cu:73-118,1855-1935
void GPUFFTBootstrap::MKMSwitch_CUDA(const std::shared_ptr<LWECryptoParams> params,
                                     std::shared_ptr<std::vector<LWECiphertext>> ctExt, NativeInteger fmod) {
    if (!g_handle)
        OPENFHE_THROW(openfhe_error, "MKMSwitch_CUDA: GPUSetup has not been called");
    const size_t batch = ctExt->size();
    if (batch == 0)
        OPENFHE_THROW(openfhe_error, "MKMSwitch_CUDA: empty batch");
    const uint64_t N = params->GetN(), n = params->Getn();
    std::vector<uint64_t> in(batch * (N + 1)), out(batch * (n + 1));
#pragma omp parallel for if (batch > 512)
    for (size_t s = 0; s < batch; s++) {
        const auto& ct = (*ctExt)[s];
        for (uint64_t i = 0; i < N; i++)
            in[s * (N + 1) + i] = ct->GetA(i).ConvertToInt();
        in[s * (N + 1) + N] = ct->GetB().ConvertToInt();
    }
    check(tfhe_b200_mkmswitch(g_handle, (int)batch, in.data(), fmod.ConvertToInt(), out.data(), TFHE_B200_HOST, nullptr),
          "MKMSwitch_CUDA");
#pragma omp parallel for if (batch > 512)
    for (size_t s = 0; s < batch; s++) {
        NativeVector a(n, fmod);
        for (uint64_t k = 0; k < n; k++)
            a[k] = NativeInteger(out[s * (n + 1) + k]);
        (*ctExt)[s] = std::make_shared<LWECiphertextImpl>(std::move(a), NativeInteger(out[s * (n + 1) + n]));
    }
}

void GPULWEOperation::GPUSetup(int) {}   // the engine needs no cuBLAS handle (lwe-operation.cu:143-146)
void GPULWEOperation::GPUClean() {}

// replaces lwe-operation.cu:50-141 (exact integers instead of FP64 cublasDgemm + fmod)
std::shared_ptr<std::vector<LWECiphertext>> GPULWEOperation::CiphertextMulMatrix_CUDA(
    const std::shared_ptr<BinFHECryptoParams> params, const std::vector<LWECiphertext>& ct,
    const std::vector<std::vector<int64_t>>& matrix, uint64_t modulus) {
    if (!g_handle)
        OPENFHE_THROW(openfhe_error, "CiphertextMulMatrix_CUDA: GPUSetup has not been called");
    if (ct.empty())
        OPENFHE_THROW(openfhe_error, "Input ciphertexts are empty.");
    if (matrix.empty() || matrix[0].empty())
        OPENFHE_THROW(openfhe_error, "Input matrix is empty.");
    if (ct.size() != matrix.size())
        OPENFHE_THROW(openfhe_error,
                      "The number of rows of the matrix must be equal to the number of input ciphertexts.");
    const size_t K = ct.size(), outc = matrix[0].size();
    const uint64_t n = params->GetLWEParams()->Getn();
    std::vector<uint64_t> fct(K * (n + 1)), fout(outc * (n + 1));
    std::vector<int64_t> fm(K * outc);
    for (size_t k = 0; k < K; k++) {
        for (uint64_t i = 0; i < n; i++)
            fct[k * (n + 1) + i] = ct[k]->GetA(i).ConvertToInt();
        fct[k * (n + 1) + n] = ct[k]->GetB().ConvertToInt();
        for (size_t i = 0; i < outc; i++)
            fm[k * outc + i] = matrix[k][i];
    }
    check(tfhe_b200_mul_matrix(g_handle, (int)K, (int)outc, fct.data(), fm.data(), modulus, fout.data(),
                               TFHE_B200_HOST, nullptr),
          "CiphertextMulMatrix_CUDA");
    auto res = std::make_shared<std::vector<LWECiphertext>>(outc);
    for (size_t i = 0; i < outc; i++) {
        NativeVector a(n, NativeInteger(modulus));
        for (uint64_t k = 0; k < n; k++)
            a[k] = NativeInteger(fout[i * (n + 1) + k]);
        (*res)[i] = std::make_shared<LWECiphertextImpl>(std::move(a), NativeInteger(fout[i * (n + 1) + n]));
    }
    return res;
}

}  // namespace lbcrypto
