// Header-only C++ adapter: the batched BinFHEContext surface of eric070021/TFHE-GPU (binfhecontext.h:352-421) on
// std::vector<LWECiphertext>, implemented with the FUSED entry points of the C ABI (include/tfhe_b200.h).
//
// Compared with the link-time shim (binfhe_b200_shim.cpp, which keeps the reference's host control flow and its two
// host round trips per bootstrap) this route hands a whole batched method to the engine in one call: ciphertexts are
// flattened once, stay on the device between the bootstraps of the call, and come back once.
//
//   lbcrypto::BinFHEContext cc;  cc.GenerateBinFHEContext(STD128, GINX);  auto sk = cc.KeyGen();  cc.BTKeyGen(sk);
//   tfhe_b200::BatchedBinFHE gpu(cc, /*numGPUs=*/0);                 // == cc.GPUSetup()
//   auto out = gpu.EvalBinGate(NAND, ct1_vec, ct2_vec);              // == cc.EvalBinGate(NAND, ct1_vec, ct2_vec)
//   // ~BatchedBinFHE()                                              // == cc.GPUClean()
//
// Error behaviour mirrors the reference: API misuse throws lbcrypto::openfhe_error with the reference's messages.
#pragma once
#include <cmath>
#include <string>
#include <vector>

#include "binfhecontext.h"
#include "tfhe_b200.h"

static_assert(sizeof(lbcrypto::NativeInteger) == sizeof(uint64_t),
              "the adapter passes NativeVector storage to the C ABI as flat uint64_t arrays");

namespace tfhe_b200 {

class BatchedBinFHE {
public:
    using CT = lbcrypto::LWECiphertext;

    explicit BatchedBinFHE(lbcrypto::BinFHEContext& cc, int numGPUs = 0) {
        using namespace lbcrypto;
        auto params = cc.GetParams();
        auto L = params->GetLWEParams();
        auto R = params->GetRingGSWParams();
        auto BSkey = cc.GetRefreshKey();
        auto KSkey = cc.GetSwitchKey();
        if (BSkey == nullptr || KSkey == nullptr)
            OPENFHE_THROW(openfhe_error, "ERROR: Need to call BTKeyGen before calling GPUSetup");
        tfhe_b200_params& p = m_p;
        p = tfhe_b200_params{};
        p.n = L->Getn();
        p.N = L->GetN();
        p.q = L->Getq().ConvertToInt();
        p.Q = L->GetQ().ConvertToInt();
        p.qKS = L->GetqKS().ConvertToInt();
        p.baseKS = L->GetBaseKS();
        p.dKS = (uint32_t)std::ceil(log((double)p.qKS) / log((double)p.baseKS));
        p.baseG = R->GetBaseG();
        p.digitsG = R->GetDigitsG();
        p.numDigitsToThrow = R->GetNumDigitsToThrow();
        p.baseR = R->GetBaseR();
        p.digitsR = (uint32_t)R->GetDigitsR().size();
        p.method = (uint32_t)R->GetMethod();
        p.psi = R->GetPolyParams()->GetRootOfUnity().ConvertToInt();
        p.beta = cc.GetBeta().ConvertToInt();
        std::vector<uint64_t> bk, ksk;
        flatten_keys(p, BSkey, KSkey, bk, ksk);
        check(tfhe_b200_setup(&p, bk.data(), bk.size(), ksk.data(), ksk.size(), TFHE_B200_HOST, 0, numGPUs, &m_h),
              "GPUSetup");
        // timeOptimization contexts (BTKeyGen filled m_BTKey_map with the 2^14 / 2^18 / 2^27 key sets,
        // binfhecontext.cpp:222-247): the reference's GPUSetup throws here (binfhecontext.cpp:350-353); we load the other
        // sets so that EvalSign / EvalDecomp switch base as the scalar CPU path does (binfhe-base-scheme.cpp:342-360)
        auto keyMap = cc.GetBTKeyMap();
        if (keyMap->size() == 3 && p.method == TFHE_B200_METHOD_GINX) {
            for (const auto& kv : *keyMap) {
                if (kv.first == p.baseG)
                    continue;
                tfhe_b200_params p2 = p;
                p2.baseG = kv.first;
                p2.digitsG = (uint32_t)std::ceil(log((double)p.Q) / log((double)kv.first));
                flatten_keys(p2, kv.second.BSkey, kv.second.KSkey, bk, ksk);
                check(tfhe_b200_add_key_set(m_h, p2.baseG, bk.data(), bk.size(), ksk.data(), ksk.size(), TFHE_B200_HOST),
                      "GPUSetup");
            }
        }
    }
    ~BatchedBinFHE() {
        if (m_h)
            tfhe_b200_clean(m_h);
    }
    BatchedBinFHE(const BatchedBinFHE&) = delete;
    BatchedBinFHE& operator=(const BatchedBinFHE&) = delete;

    // binfhecontext.cpp:323-325 -> binfhe-base-scheme.cpp:598-677
    std::vector<CT> EvalBinGate(lbcrypto::BINGATE gate, const std::vector<CT>& ct1, const std::vector<CT>& ct2) const {
        using namespace lbcrypto;
        if (ct1.empty() || ct2.empty())
            OPENFHE_THROW(openfhe_error, "ERROR: EvalBinGate: input vector is empty");
        if (ct1.size() != ct2.size())
            OPENFHE_THROW(openfhe_error, "ERROR: EvalBinGate: input ciphertexts size unmatched");
        if (&ct1 == &ct2 || ct1 == ct2)
            OPENFHE_THROW(config_error, "Input ciphertexts should be independant");
        const uint64_t mod = ct1[0]->GetModulus().ConvertToInt();
        // no flattening: the engine gathers straight from the ciphertext objects into its pinned staging, chunk by chunk
        // under the running bootstraps, and scatters the results into the output objects the same way
        // (tfhe_b200_eval_bin_gate_v); &GetA()[0] is a flat u64 array
        const size_t batch = ct1.size(), n = m_p.n;
        std::vector<const uint64_t*> a1(batch), a2(batch);
        std::vector<uint64_t> b1(batch), b2(batch), bo(batch);
        std::vector<uint64_t*> ao(batch);
        std::vector<NativeVector> outA(batch);
#pragma omp parallel for if (batch > 512)
        for (size_t s = 0; s < batch; s++) {
            if (ct1[s]->GetLength() != n || ct2[s]->GetLength() != n)
                continue;   // reported below (no exceptions out of a parallel region)
            a1[s] = reinterpret_cast<const uint64_t*>(&ct1[s]->GetA()[0]);
            a2[s] = reinterpret_cast<const uint64_t*>(&ct2[s]->GetA()[0]);
            b1[s] = ct1[s]->GetB().ConvertToInt();
            b2[s] = ct2[s]->GetB().ConvertToInt();
            outA[s] = NativeVector(n, NativeInteger(mod));
            ao[s] = reinterpret_cast<uint64_t*>(&outA[s][0]);
        }
        for (size_t s = 0; s < batch; s++)
            if (!a1[s])
                OPENFHE_THROW(openfhe_error, "ERROR: EvalBinGate: ciphertext dimension does not match the key");
        check(tfhe_b200_eval_bin_gate_v(m_h, (int)gate, (int)batch, a1.data(), b1.data(), a2.data(), b2.data(), mod,
                                        ao.data(), bo.data(), nullptr),
              "EvalBinGate");
        std::vector<CT> ret(batch);
#pragma omp parallel for if (batch > 512)
        for (size_t s = 0; s < batch; s++)
            ret[s] = std::make_shared<LWECiphertextImpl>(std::move(outA[s]), NativeInteger(bo[s]));
        return ret;
    }
    // binfhecontext.cpp:327-330
    std::vector<CT> EvalFunc(const std::vector<CT>& ct, const std::vector<lbcrypto::NativeInteger>& LUT) const {
        need(ct, "EvalFunc");
        const uint64_t mod = ct[0]->GetModulus().ConvertToInt();
        std::vector<uint64_t> lut(LUT.size());
        for (size_t i = 0; i < LUT.size(); i++)
            lut[i] = LUT[i].ConvertToInt();
        auto a = flatten(ct);
        std::vector<uint64_t> o(a.size());
        check(tfhe_b200_eval_func(m_h, (int)ct.size(), a.data(), mod, lut.data(), lut.size(), 0, o.data(), TFHE_B200_HOST,
                                  nullptr),
              "EvalFunc");
        return unflatten(o, ct.size(), mod);
    }
    // binfhecontext.cpp:332-335
    std::vector<CT> EvalFunc(const std::vector<CT>& ct,
                             const std::vector<std::vector<lbcrypto::NativeInteger>>& LUT_vec) const {
        using namespace lbcrypto;
        need(ct, "EvalFunc");
        if (ct.size() != LUT_vec.size())
            OPENFHE_THROW(openfhe_error, "ERROR: EvalFunc: input ciphertexts size unmatched with LUT size");
        const uint64_t mod = ct[0]->GetModulus().ConvertToInt();
        const size_t len = LUT_vec[0].size();
        std::vector<uint64_t> lut(ct.size() * len);
        for (size_t s = 0; s < ct.size(); s++)
            for (size_t i = 0; i < len; i++)
                lut[s * len + i] = LUT_vec[s][i].ConvertToInt();
        auto a = flatten(ct);
        std::vector<uint64_t> o(a.size());
        check(tfhe_b200_eval_func(m_h, (int)ct.size(), a.data(), mod, lut.data(), len, 1, o.data(), TFHE_B200_HOST, nullptr),
              "EvalFunc");
        return unflatten(o, ct.size(), mod);
    }
    // binfhecontext.cpp:337-339
    std::vector<CT> EvalFloor(const std::vector<CT>& ct, uint32_t roundbits = 0) const {
        need(ct, "EvalFunc");
        const uint64_t mod = ct[0]->GetModulus().ConvertToInt();
        auto a = flatten(ct);
        std::vector<uint64_t> o(a.size());
        check(tfhe_b200_eval_floor(m_h, (int)ct.size(), a.data(), mod, roundbits, o.data(), TFHE_B200_HOST, nullptr),
              "EvalFloor");
        return unflatten(o, ct.size(), mod);
    }
    // binfhecontext.cpp:341-343
    std::vector<CT> EvalSign(const std::vector<CT>& ct) const {
        need(ct, "EvalFunc");
        const uint64_t mod = ct[0]->GetModulus().ConvertToInt();
        auto a = flatten(ct);
        std::vector<uint64_t> o(a.size());
        check(tfhe_b200_eval_sign(m_h, (int)ct.size(), a.data(), mod, o.data(), TFHE_B200_HOST, nullptr), "EvalSign");
        return unflatten(o, ct.size(), m_p.q);
    }
    // binfhecontext.cpp:345-347
    std::vector<std::vector<CT>> EvalDecomp(const std::vector<CT>& ct) const {
        need(ct, "EvalFunc");
        const uint64_t mod = ct[0]->GetModulus().ConvertToInt();
        const int maxd = 16;
        const size_t W = m_p.n + 1;
        auto a = flatten(ct);
        std::vector<uint64_t> o(ct.size() * maxd * W), mods(maxd);
        int nd = tfhe_b200_eval_decomp(m_h, (int)ct.size(), a.data(), mod, maxd, o.data(), mods.data(), TFHE_B200_HOST,
                                       nullptr);
        check(nd, "EvalDecomp");
        std::vector<std::vector<CT>> ret(ct.size());
        for (size_t s = 0; s < ct.size(); s++)
            for (int k = 0; k < nd; k++)
                ret[s].push_back(make(o.data() + (s * maxd + k) * W, mods[k]));
        return ret;
    }
    // SURVEY section 8(f) rank 1 (no reference counterpart: replaces a host loop of EvalBinGate / EvalNOT calls).
    // inputs[w] = the batch on input wire w; nodes in topological order, wire ids = inputs first, then node outputs.
    struct GateNode {
        int gate;   // lbcrypto::BINGATE value (OR, AND, NOR, NAND, XOR_FAST, XNOR_FAST, XOR, XNOR) or kNot
        int in0, in1;
    };
    static constexpr int kNot = TFHE_B200_NOT;
    std::vector<std::vector<CT>> EvalCircuit(const std::vector<std::vector<CT>>& inputs, const std::vector<GateNode>& nodes,
                                             const std::vector<int>& outputs) const {
        using namespace lbcrypto;
        if (inputs.empty() || inputs[0].empty())
            OPENFHE_THROW(openfhe_error, "ERROR: EvalCircuit: input vector is empty");
        const size_t batch = inputs[0].size(), W = m_p.n + 1;
        std::vector<uint64_t> in(inputs.size() * batch * W);
        for (size_t w = 0; w < inputs.size(); w++) {
            if (inputs[w].size() != batch)
                OPENFHE_THROW(openfhe_error, "ERROR: EvalCircuit: input ciphertexts size unmatched");
            auto a = flatten(inputs[w]);
            std::copy(a.begin(), a.end(), in.begin() + w * batch * W);
        }
        std::vector<tfhe_b200_gate_node> nd(nodes.size());
        for (size_t g = 0; g < nodes.size(); g++)
            nd[g] = tfhe_b200_gate_node{nodes[g].gate, nodes[g].in0, nodes[g].in1};
        std::vector<int32_t> ow(outputs.begin(), outputs.end());
        std::vector<uint64_t> o(ow.size() * batch * W);
        const uint64_t mod = inputs[0][0]->GetModulus().ConvertToInt();
        check(tfhe_b200_eval_circuit(m_h, (int)batch, (int)inputs.size(), in.data(), mod, (int)nd.size(), nd.data(),
                                     (int)ow.size(), ow.data(), o.data(), TFHE_B200_HOST, nullptr),
              "EvalCircuit");
        std::vector<std::vector<CT>> ret(ow.size());
        for (size_t k = 0; k < ow.size(); k++) {
            std::vector<uint64_t> slice(o.begin() + k * batch * W, o.begin() + (k + 1) * batch * W);
            ret[k] = unflatten(slice, batch, mod);
        }
        return ret;
    }
    // binfhecontext.cpp:319-321
    std::vector<CT> CiphertextMulMatrix(const std::vector<CT>& ct, const std::vector<std::vector<int64_t>>& matrix,
                                        uint64_t modulus) const {
        using namespace lbcrypto;
        if (ct.empty())
            OPENFHE_THROW(openfhe_error, "Input ciphertexts are empty.");
        if (matrix.empty() || matrix[0].empty())
            OPENFHE_THROW(openfhe_error, "Input matrix is empty.");
        if (ct.size() != matrix.size())
            OPENFHE_THROW(openfhe_error,
                          "The number of rows of the matrix must be equal to the number of input ciphertexts.");
        const size_t outc = matrix[0].size(), W = m_p.n + 1;
        std::vector<int64_t> m(ct.size() * outc);
        for (size_t k = 0; k < ct.size(); k++)
            for (size_t i = 0; i < outc; i++)
                m[k * outc + i] = matrix[k][i];
        auto a = flatten(ct);
        std::vector<uint64_t> o(outc * W);
        check(tfhe_b200_mul_matrix(m_h, (int)ct.size(), (int)outc, a.data(), m.data(), modulus, o.data(), TFHE_B200_HOST,
                                   nullptr),
              "CiphertextMulMatrix");
        return unflatten(o, outc, modulus);
    }

    const tfhe_b200_params& params() const { return m_p; }
    tfhe_b200_handle* handle() const { return m_h; }

private:
    // BK / KSK in the element order tfhe_b200_setup documents (the flattening of bootstrapping.cu:933-975)
    static void flatten_keys(const tfhe_b200_params& p, const lbcrypto::RingGSWACCKey& BSkey,
                             const lbcrypto::LWESwitchingKey& KSkey, std::vector<uint64_t>& bk,
                             std::vector<uint64_t>& ksk) {
        const uint64_t N = p.N, n = p.n;
        bk.assign(tfhe_b200_bk_words(&p), 0);
        ksk.assign(tfhe_b200_ksk_words(&p), 0);
        if (p.method == TFHE_B200_METHOD_GINX) {
            const uint64_t d = 2 * (p.digitsG - p.numDigitsToThrow);
#pragma omp parallel for collapse(2)
            for (uint64_t key = 0; key < 2; key++)
                for (uint64_t i = 0; i < n; i++) {
                    const auto& ev = (*BSkey)[0][key][i]->GetElements();
                    for (uint64_t l = 0; l < d; l++)
                        for (uint64_t j = 0; j < 2; j++) {
                            uint64_t* dst = bk.data() + ((((key * n + i) * d + l) * 2 + j) * N);
                            for (uint64_t k = 0; k < N; k++)
                                dst[k] = ev[l][j][k].ConvertToInt();
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
                                for (uint64_t x = 0; x < N; x++)
                                    dst[x] = ev[l][j][x].ConvertToInt();
                            }
                    }
        }
        const auto& A = KSkey->GetElementsA();
        const auto& B = KSkey->GetElementsB();
#pragma omp parallel for
        for (uint64_t i = 0; i < N; i++)
            for (uint64_t a0 = 0; a0 < p.baseKS; a0++)
                for (uint64_t j = 0; j < p.dKS; j++) {
                    uint64_t* dst = ksk.data() + (((i * p.baseKS + a0) * p.dKS + j) * (n + 1));
                    for (uint64_t k = 0; k < n; k++)
                        dst[k] = A[i][a0][j][k].ConvertToInt();
                    dst[n] = B[i][a0][j].ConvertToInt();
                }
    }
    tfhe_b200_handle* m_h = nullptr;
    tfhe_b200_params m_p;

    static void check(int rc, const char* what) {
        if (rc < 0)
            OPENFHE_THROW(lbcrypto::openfhe_error, std::string(what) + ": " + tfhe_b200_last_error());
    }
    static void need(const std::vector<CT>& ct, const char* what) {
        if (ct.empty())
            OPENFHE_THROW(lbcrypto::openfhe_error, std::string("ERROR: ") + what + ": input vector is empty");
    }
    std::vector<uint64_t> flatten(const std::vector<CT>& v) const {
        const size_t n = m_p.n, W = n + 1;
        std::vector<uint64_t> f(v.size() * W);
#pragma omp parallel for if (v.size() > 512)
        for (size_t s = 0; s < v.size(); s++) {
            for (size_t i = 0; i < n; i++)
                f[s * W + i] = v[s]->GetA(i).ConvertToInt();
            f[s * W + n] = v[s]->GetB().ConvertToInt();
        }
        return f;
    }
    CT make(const uint64_t* p, uint64_t mod) const {
        const size_t n = m_p.n;
        lbcrypto::NativeVector a(n, lbcrypto::NativeInteger(mod));
        for (size_t i = 0; i < n; i++)
            a[i] = lbcrypto::NativeInteger(p[i]);
        return std::make_shared<lbcrypto::LWECiphertextImpl>(std::move(a), lbcrypto::NativeInteger(p[n]));
    }
    std::vector<CT> unflatten(const std::vector<uint64_t>& f, size_t count, uint64_t mod) const {
        const size_t W = m_p.n + 1;
        std::vector<CT> v(count);
#pragma omp parallel for if (count > 512)
        for (size_t s = 0; s < count; s++)
            v[s] = make(f.data() + s * W, mod);
        return v;
    }
};

}  // namespace tfhe_b200
