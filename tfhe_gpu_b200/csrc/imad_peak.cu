// Integer-pipe microbenchmark: measures the peak issue rate of the 32-bit integer multiply-add family on this GPU.
// The blind-rotation roofline ("fraction of the IMAD peak") uses the number measured here,
// because MEASURED_PEAKS.json only carries HBM and bf16 figures.
//
//   imad_peak [out.json]
//
// Variants: IMAD (mad.lo.u32), IMAD.HI.U32 (mul.hi.u32), IMAD.WIDE.U32 (mad.wide.u32), IADD3 (alu pipe), a 1:1
// IMAD+IADD mix (dual issue across the fma and alu pipes) and the exact instruction mix of one lazy Shoup
// butterfly (1 IMAD.HI + 2 IMAD + 2 IADD3).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CHECK(x)                                                                          \
    do {                                                                                  \
        cudaError_t e = (x);                                                              \
        if (e != cudaSuccess) {                                                           \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e));     \
            exit(1);                                                                      \
        }                                                                                 \
    } while (0)

constexpr int ILP = 8;

template <int VARIANT>
__global__ void __launch_bounds__(256) k(unsigned* out, int iters, unsigned a, unsigned b, long long* cyc) {
    unsigned x[ILP];
    unsigned long long w[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) {
        x[i] = threadIdx.x * 7 + i + a;
        w[i] = x[i];
    }
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (VARIANT == 0)
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
            else if (VARIANT == 1)
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
            else if (VARIANT == 2)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((unsigned)w[i]), "r"(a));
            else if (VARIANT == 3)
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
            else if (VARIANT == 4) {
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x[(i + 4) % ILP]) : "r"(b));
            }
            else if (VARIANT == 5) {
                // lazy Shoup butterfly on (x[i], x[i^1]) -- executed for even i only
                if ((i & 1) == 0) {
                    unsigned q, t;
                    asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(q) : "r"(x[i + 1]), "r"(b));
                    asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(t) : "r"(x[i + 1]), "r"(a));
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(t) : "r"(q), "r"(0x7ffe001u));
                    unsigned u = x[i];
                    asm volatile("add.u32 %0, %1, %2;" : "=r"(x[i]) : "r"(u), "r"(t));
                    asm volatile("sub.u32 %0, %1, %2;" : "=r"(x[i + 1]) : "r"(u), "r"(t));
                }
            }
        }
    }
    long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++)
        s += x[i] + (unsigned)w[i] + (unsigned)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0)
        cyc[blockIdx.x] = t1 - t0;
}

struct Res {
    const char* name;
    double gops, per_sm_clk, ms;
    double ops_per_iter;
};

template <int V>
Res run(const char* name, double ops_per_unrolled_iter, int sms, int ctas_per_sm) {
    int grid = sms * ctas_per_sm, block = 256, iters = 20000;
    unsigned* out;
    long long* cyc;
    CHECK(cudaMalloc(&out, (size_t)grid * block * 4));
    CHECK(cudaMalloc(&cyc, grid * 8));
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; w++)
        k<V><<<grid, block>>>(out, iters, 3, 5, cyc);
    CHECK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        CHECK(cudaEventRecord(e0));
        k<V><<<grid, block>>>(out, iters, 3, 5, cyc);
        CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1));
        float ms;
        CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best)
            best = ms;
    }
    std::vector<long long> hc(grid);
    CHECK(cudaMemcpy(hc.data(), cyc, grid * 8, cudaMemcpyDeviceToHost));
    double avg_cyc = 0;
    for (auto c : hc)
        avg_cyc += (double)c;
    avg_cyc /= grid;
    double total_ops = (double)grid * block * iters * ops_per_unrolled_iter;
    Res r;
    r.name = name;
    r.ms = best;
    r.gops = total_ops / (best * 1e-3) / 1e9;
    // ops issued by one SM (ctas_per_sm resident CTAs run concurrently) per SM clock
    r.per_sm_clk = (double)ctas_per_sm * block * iters * ops_per_unrolled_iter / avg_cyc;
    r.ops_per_iter = ops_per_unrolled_iter;
    CHECK(cudaFree(out));
    CHECK(cudaFree(cyc));
    return r;
}

int main(int argc, char** argv) {
    cudaDeviceProp p;
    CHECK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    std::vector<Res> rs;
    rs.push_back(run<0>("imad_lo", ILP, sms, 8));
    rs.push_back(run<1>("imad_hi_u32", ILP, sms, 8));
    rs.push_back(run<2>("imad_wide_u32", ILP, sms, 8));
    rs.push_back(run<3>("iadd3", ILP, sms, 8));
    rs.push_back(run<4>("imad_lo+iadd_1to1", 2 * ILP, sms, 8));
    rs.push_back(run<5>("shoup_butterfly_5instr", 5 * (ILP / 2), sms, 8));
    FILE* f = argc > 1 ? fopen(argv[1], "w") : stdout;
    fprintf(f, "{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, \"variants\": {", p.name, sms, p.clockRate);
    for (size_t i = 0; i < rs.size(); i++)
        fprintf(f, "%s\"%s\": {\"gops\": %.1f, \"per_sm_per_clk\": %.2f, \"ms\": %.3f}", i ? ", " : "", rs[i].name,
                rs[i].gops, rs[i].per_sm_clk, rs[i].ms);
    fprintf(f, "}}\n");
    if (f != stdout)
        fclose(f);
    return 0;
}
