// placeholder: specialised kernel not built yet
#include "engine.cuh"
namespace tfhe_b200 {
bool cggi32_supported(const tfhe_b200_params&) { return false; }
cudaError_t launch_br_cggi32(const BRCommon&, const CGGI32Tables&, cudaStream_t, int, int) { return cudaErrorNotSupported; }
size_t cggi32_twB_words(u32) { return 0; }
size_t cggi32_twA_words() { return 0; }
void cggi32_build_tables(const tfhe_b200_params&, std::vector<u32>&, std::vector<u32>&) {}
}
