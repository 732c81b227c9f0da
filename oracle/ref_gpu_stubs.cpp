// TEST INFRASTRUCTURE ONLY (oracle build of the reference CPU path).
//
// The reference's host sources (binfhe-base-scheme.cpp, binfhecontext.cpp, rgsw-acc-cggi.cpp) reference six
// static GPU entry points that live in its two .cu files (bootstrapping.cu, lwe-operation.cu).  The oracle
// build does not compile those files (cuFFTDx FFT path, approximate, no sm_100 dispatch); the entry points
// are defined here so the library links, and they throw: the oracle is the *scalar CPU* API only.
//
// Replaces: bootstrapping.cuh:111-136, lwe-operation.cuh:49-62 (declarations only are used).
#include "binfhecontext.h"

namespace lbcrypto {

static void no_gpu(const char* what) {
    OPENFHE_THROW(not_available_error, std::string(what) + ": GPU path is not part of the CPU oracle build");
}

void GPUFFTBootstrap::GPUSetup(const std::shared_ptr<BinFHECryptoParams>, RingGSWACCKey, LWESwitchingKey, int) {
    no_gpu("GPUFFTBootstrap::GPUSetup");
}
void GPUFFTBootstrap::GPUClean() {
    no_gpu("GPUFFTBootstrap::GPUClean");
}
void GPUFFTBootstrap::EvalAcc_CUDA(const std::shared_ptr<RingGSWCryptoParams>, const std::vector<NativeVector>&,
                                   std::shared_ptr<std::vector<RLWECiphertext>>, uint64_t) {
    no_gpu("GPUFFTBootstrap::EvalAcc_CUDA");
}
void GPUFFTBootstrap::MKMSwitch_CUDA(const std::shared_ptr<LWECryptoParams>,
                                     std::shared_ptr<std::vector<LWECiphertext>>, NativeInteger) {
    no_gpu("GPUFFTBootstrap::MKMSwitch_CUDA");
}
void GPULWEOperation::GPUSetup(int) {
    no_gpu("GPULWEOperation::GPUSetup");
}
void GPULWEOperation::GPUClean() {
    no_gpu("GPULWEOperation::GPUClean");
}
std::shared_ptr<std::vector<LWECiphertext>> GPULWEOperation::CiphertextMulMatrix_CUDA(
    const std::shared_ptr<BinFHECryptoParams>, const std::vector<LWECiphertext>&,
    const std::vector<std::vector<int64_t>>&, uint64_t) {
    no_gpu("GPULWEOperation::CiphertextMulMatrix_CUDA");
    return nullptr;
}

}  // namespace lbcrypto
