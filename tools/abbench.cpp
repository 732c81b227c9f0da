// Developer A/B harness (not the contract bench): times tfhe_b200_eval_bin_gate of several builds of libtfhe_b200.so on
// the same device-resident inputs and keys, round-robin, and compares the outputs bit for bit (FNV-1a of the result).
// No Python, no torch: start-up is a second, so a GPU call spends its time on the kernels.
//
//   g++ -O2 -o build/abbench tools/abbench.cpp -I include -I /usr/local/cuda/include -L /usr/local/cuda/lib64 -lcudart -ldl
//   abbench [--set std128|ap|toy|medium|std128_ap_set|std256|std256q|func12|sign17] [--batch B[,B..]] [--reps R] SPEC [SPEC ...]
//   SPEC = path/to/libtfhe_b200.so[:option=value[:option=value]]     (options of tfhe_b200_set_option)
//
// Keys come from tfhe_b200_keygen_test_seed of the FIRST library (deterministic), inputs are uniform random words.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "tfhe_b200.h"

struct Lib {
    std::string spec, path;
    std::vector<std::pair<std::string, long long>> opts;
    void* dl = nullptr;
    tfhe_b200_handle* h = nullptr;
    decltype(&tfhe_b200_setup) setup;
    decltype(&tfhe_b200_clean) clean;
    decltype(&tfhe_b200_last_error) last_error;
    decltype(&tfhe_b200_set_option) set_option;
    decltype(&tfhe_b200_eval_bin_gate) eval_bin_gate;
    decltype(&tfhe_b200_bootstrap_func) bootstrap_func;
    decltype(&tfhe_b200_keygen_test_seed) keygen;
    decltype(&tfhe_b200_bk_words) bk_words;
    decltype(&tfhe_b200_ksk_words) ksk_words;
    decltype(&tfhe_b200_kernel_variant) variant;
};

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                            \
        }                                                                                       \
    } while (0)

static unsigned long long fnv(const void* p, size_t bytes) {
    const unsigned long long* w = (const unsigned long long*)p;
    unsigned long long h = 1469598103934665603ULL;
    for (size_t i = 0; i < bytes / 8; i++) {
        h ^= w[i];
        h *= 1099511628211ULL;
    }
    return h;
}

int main(int argc, char** argv) {
    std::string set = "std128";
    std::vector<int> batches = {16384};
    int reps = 5, gate = TFHE_B200_NAND;
    bool func = false;   // time ONE BootstrapFunc (blind rotation + MS/KS/MS) instead of a gate
    std::vector<Lib> libs;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--set") set = argv[++i];
        else if (a == "--reps") reps = atoi(argv[++i]);
        else if (a == "--gate") gate = atoi(argv[++i]);
        else if (a == "--batch") {
            batches.clear();
            char* s = argv[++i];
            for (char* t = strtok(s, ","); t; t = strtok(nullptr, ",")) batches.push_back(atoi(t));
        }
        else {
            Lib L;
            L.spec = a;
            size_t p = a.find(':');
            L.path = a.substr(0, p);
            while (p != std::string::npos) {
                size_t q = a.find(':', p + 1);
                std::string kv = a.substr(p + 1, q == std::string::npos ? std::string::npos : q - p - 1);
                size_t e = kv.find('=');
                L.opts.push_back({kv.substr(0, e), atoll(kv.substr(e + 1).c_str())});
                p = q;
            }
            libs.push_back(L);
        }
    }
    if (libs.empty()) {
        fprintf(stderr, "usage: abbench [--set std128|ap|toy|medium|std128_ap_set|std256|std256q|func12|sign17] [--batch B,..] [--reps R] LIB[:opt=val] ...\n");
        return 1;
    }
    tfhe_b200_params P;
    memset(&P, 0, sizeof(P));
    if (set == "std128" || set == "ap") {   // binfhecontext.cpp:141 (STD128), GINX / AP
        P.n = 512; P.N = 1024; P.q = 1024; P.Q = 134215681ULL; P.qKS = 16384; P.baseKS = 128; P.dKS = 2;
        P.baseG = 128; P.digitsG = 4; P.baseR = 32; P.psi = 282116; P.beta = 128;
        P.method = set == "ap" ? TFHE_B200_METHOD_AP : TFHE_B200_METHOD_GINX;
        P.digitsR = set == "ap" ? 2 : 0;
    }
    else if (set == "toy") {            // binfhecontext.cpp:139
        P.n = 64; P.N = 512; P.q = 512; P.Q = 134215681ULL; P.qKS = 134215681ULL; P.baseKS = 25; P.dKS = 6;
        P.baseG = 512; P.digitsG = 3; P.baseR = 23; P.psi = 78074; P.beta = 128; P.method = TFHE_B200_METHOD_GINX;
    }
    else if (set == "medium") {         // binfhecontext.cpp:140 (28-bit modulus: reduction sweep)
        P.n = 422; P.N = 1024; P.q = 1024; P.Q = 268369921ULL; P.qKS = 16384; P.baseKS = 128; P.dKS = 2;
        P.baseG = 1024; P.digitsG = 3; P.baseR = 32; P.psi = 326097; P.beta = 128; P.method = TFHE_B200_METHOD_GINX;
    }
    else if (set == "std128_ap_set") {  // binfhecontext.cpp:141 under GINX (three digits, top digit may wrap: plain path)
        P.n = 512; P.N = 1024; P.q = 1024; P.Q = 134215681ULL; P.qKS = 16384; P.baseKS = 128; P.dKS = 2;
        P.baseG = 512; P.digitsG = 3; P.baseR = 32; P.psi = 282116; P.beta = 128; P.method = TFHE_B200_METHOD_GINX;
    }
    else if (set == "std256") {         // binfhecontext.cpp:147 (N = 2048, 29-bit modulus)
        P.n = 1024; P.N = 2048; P.q = 2048; P.Q = 536813569ULL; P.qKS = 16384; P.baseKS = 128; P.dKS = 2;
        P.baseG = 256; P.digitsG = 4; P.baseR = 46; P.psi = 145054; P.beta = 128; P.method = TFHE_B200_METHOD_GINX;
    }
    else if (set == "std256q") {        // binfhecontext.cpp:153 (N = 2048, 27-bit modulus)
        P.n = 2048; P.N = 2048; P.q = 2048; P.Q = 134176769ULL; P.qKS = 65536; P.baseKS = 16; P.dKS = 4;
        P.baseG = 128; P.digitsG = 4; P.baseR = 46; P.psi = 100530; P.beta = 128; P.method = TFHE_B200_METHOD_GINX;
    }
    else if (set == "func12" || set == "sign17") {   // the STD128 functional sets (logQ = 12 arbitrary LUT, logQ = 17)
        P.n = 1305; P.N = 2048; P.q = set == "func12" ? 2048 : 4096; P.Q = 18014398509404161ULL; P.qKS = 34359738368ULL;
        P.baseKS = 32; P.dKS = 7; P.baseG = set == "func12" ? 134217728u : 262144u; P.digitsG = set == "func12" ? 2 : 3;
        P.baseR = 23; P.psi = 2604308523238ULL; P.beta = 128; P.method = TFHE_B200_METHOD_GINX;
        func = true;
    }
    else {
        fprintf(stderr, "unknown set %s\n", set.c_str());
        return 1;
    }
    for (auto& L : libs) {
        L.dl = dlopen(L.path.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!L.dl) {
            fprintf(stderr, "dlopen %s: %s\n", L.path.c_str(), dlerror());
            return 2;
        }
#define SYM(f, name) L.f = (decltype(L.f))dlsym(L.dl, name); if (!L.f) { fprintf(stderr, "missing %s\n", name); return 2; }
        SYM(setup, "tfhe_b200_setup") SYM(clean, "tfhe_b200_clean") SYM(last_error, "tfhe_b200_last_error")
        SYM(set_option, "tfhe_b200_set_option") SYM(eval_bin_gate, "tfhe_b200_eval_bin_gate")
        SYM(bootstrap_func, "tfhe_b200_bootstrap_func")
        SYM(keygen, "tfhe_b200_keygen_test_seed") SYM(bk_words, "tfhe_b200_bk_words") SYM(ksk_words, "tfhe_b200_ksk_words")
        SYM(variant, "tfhe_b200_kernel_variant")
    }
    CK(cudaSetDevice(0));
    const size_t bkw = libs[0].bk_words(&P), ksw = libs[0].ksk_words(&P);
    uint64_t *bk, *ksk;
    CK(cudaMalloc(&bk, bkw * 8));
    CK(cudaMalloc(&ksk, ksw * 8));
    {
        std::vector<int8_t> s1(P.n), s2(P.N);
        unsigned long long x = 88172645463325252ULL;
        auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
        for (auto& v : s1) v = (int8_t)(rnd() % 3) - 1;
        for (auto& v : s2) v = (int8_t)(rnd() % 3) - 1;
        if (libs[0].keygen(&P, s1.data(), s2.data(), 1, 0, bk, ksk)) {
            fprintf(stderr, "keygen: %s\n", libs[0].last_error());
            return 2;
        }
    }
    for (auto& L : libs) {
        if (L.setup(&P, bk, bkw, ksk, ksw, TFHE_B200_DEVICE, 0, 1, &L.h)) {
            fprintf(stderr, "setup %s: %s\n", L.spec.c_str(), L.last_error());
            return 2;
        }
        for (auto& o : L.opts)
            if (L.set_option(L.h, o.first.c_str(), o.second)) {
                fprintf(stderr, "set_option %s: %s\n", L.spec.c_str(), L.last_error());
                return 2;
            }
    }
    CK(cudaFree(bk));
    CK(cudaFree(ksk));
    const size_t W = P.n + 1;
    for (int batch : batches) {
        std::vector<uint64_t> hin(2 * (size_t)batch * W), hout((size_t)batch * W), htab(P.q);
        for (size_t i = 0; i < htab.size(); i++)
            htab[i] = (i * 2654435761ULL) % P.q;
        uint64_t* tab;
        CK(cudaMalloc(&tab, P.q * 8));
        CK(cudaMemcpy(tab, htab.data(), P.q * 8, cudaMemcpyHostToDevice));
        unsigned long long x = 0x9E3779B97F4A7C15ULL + batch;
        for (auto& v : hin) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            v = x % P.q;
        }
        uint64_t *c1, *c2, *out;
        CK(cudaMalloc(&c1, (size_t)batch * W * 8));
        CK(cudaMalloc(&c2, (size_t)batch * W * 8));
        CK(cudaMalloc(&out, (size_t)batch * W * 8));
        CK(cudaMemcpy(c1, hin.data(), (size_t)batch * W * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c2, hin.data() + (size_t)batch * W, (size_t)batch * W * 8, cudaMemcpyHostToDevice));
        std::vector<std::vector<float>> br(libs.size()), tot(libs.size());
        std::vector<unsigned long long> sum(libs.size());
        for (int r = -2; r < reps; r++)
            for (size_t k = 0; k < libs.size(); k++) {
                Lib& L = libs[k];
                tfhe_b200_stats st;
                CK(cudaMemset(out, 0, (size_t)batch * W * 8));
                if (func ? L.bootstrap_func(L.h, batch, c1, P.q, tab, 0, P.q, out, TFHE_B200_DEVICE, &st)
                         : L.eval_bin_gate(L.h, gate, batch, c1, c2, P.q, out, TFHE_B200_DEVICE, &st)) {
                    fprintf(stderr, "eval %s: %s\n", L.spec.c_str(), L.last_error());
                    return 2;
                }
                CK(cudaDeviceSynchronize());
                if (r >= 0) {
                    br[k].push_back(st.blind_rotate_ms);
                    tot[k].push_back(st.total_ms);
                }
                if (r == reps - 1) {
                    CK(cudaMemcpy(hout.data(), out, (size_t)batch * W * 8, cudaMemcpyDeviceToHost));
                    sum[k] = fnv(hout.data(), (size_t)batch * W * 8);
                }
            }
        for (size_t k = 0; k < libs.size(); k++) {
            std::sort(br[k].begin(), br[k].end());
            std::sort(tot[k].begin(), tot[k].end());
            printf("{\"spec\": \"%s\", \"set\": \"%s\", \"batch\": %d, \"variant\": \"%s\", \"br_ms_min\": %.3f, \"br_ms_med\": %.3f, "
                   "\"total_ms_min\": %.3f, \"total_ms_med\": %.3f, \"gates_per_s\": %.0f, \"fnv\": \"%016llx\", \"same_as_first\": %s}\n",
                   libs[k].spec.c_str(), set.c_str(), batch, libs[k].variant(libs[k].h), br[k].front(), br[k][br[k].size() / 2],
                   tot[k].front(), tot[k][tot[k].size() / 2], batch / (tot[k][tot[k].size() / 2] * 1e-3), sum[k],
                   sum[k] == sum[0] ? "true" : "false");
            fflush(stdout);
        }
        CK(cudaFree(tab));
        CK(cudaFree(c1));
        CK(cudaFree(c2));
        CK(cudaFree(out));
    }
    for (auto& L : libs)
        L.clean(L.h);
    return 0;
}
