// Integer-pipe microbenchmark: operand-form dependence of IMAD / IMAD.HI / IMAD.WIDE throughput and the
// co-issue behaviour with the ALU pipe.  Clocks are warmed for ~1 s first.  Prints JSON.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)
constexpr int ILP = 8;

template <int V>
__global__ void __launch_bounds__(256) k(unsigned* out, const unsigned* in, int iters, unsigned ua, unsigned ub, long long* cyc) {
    unsigned x[ILP], y[ILP], z[ILP];
    unsigned long long w[ILP];
    double dx[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) {
        x[i] = in[threadIdx.x + 256 * i];
        y[i] = in[threadIdx.x + 256 * (i + 8)] | 1;
        z[i] = in[threadIdx.x + 256 * (i + 16)];
        w[i] = x[i];
        dx[i] = (double)x[i];
    }
    const double dy = (double)y[0] * 1e-9, dz = (double)z[0];
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (V == 0)       asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(z[i]));          // R,R,R (distinct)
            else if (V == 1)  asm volatile("mul.lo.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));                           // R,R,RZ
            else if (V == 2)  asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(ua), "r"(z[i]));             // R,UR,R
            else if (V == 3)  asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(ub));             // R,R,UR
            else if (V == 4)  asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(ua), "r"(ub));               // R,UR,UR
            else if (V == 5)  asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));                           // HI R,R
            else if (V == 6)  asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(ua));                             // HI R,UR
            else if (V == 7)  asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((unsigned)w[i]), "r"(y[i])); // WIDE acc
            else if (V == 8)  asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"((unsigned)w[i]), "r"(y[i]));    // WIDE no acc
            else if (V == 9)  asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));                              // IADD3 R,R
            else if (V == 10) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(z[i]));        // LOP3
            else if (V == 11) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dx[i]) : "d"(dy), "d"(dz));               // DFMA
            else if (V == 12) { // shoup butterfly, register twiddles (pass B form)
                if ((i & 1) == 0) {
                    unsigned q, t;
                    asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(q) : "r"(x[i + 1]), "r"(z[i]));
                    asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(t) : "r"(x[i + 1]), "r"(y[i]));
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(t) : "r"(q), "r"(ua));
                    unsigned u = x[i];
                    asm volatile("add.u32 %0, %1, %2;" : "=r"(x[i]) : "r"(u), "r"(t));
                    asm volatile("{ .reg .u32 tt; sub.u32 tt, %1, %2; add.u32 %0, tt, %3; }" : "=r"(x[i + 1]) : "r"(u), "r"(t), "r"(ub));
                }
            }
            else if (V == 13) { // shoup butterfly, uniform twiddles (pass A form)
                if ((i & 1) == 0) {
                    unsigned q, t;
                    asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(q) : "r"(x[i + 1]), "r"(ub));
                    asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(t) : "r"(x[i + 1]), "r"(ua));
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(t) : "r"(q), "r"(ua));
                    unsigned u = x[i];
                    asm volatile("add.u32 %0, %1, %2;" : "=r"(x[i]) : "r"(u), "r"(t));
                    asm volatile("{ .reg .u32 tt; sub.u32 tt, %1, %2; add.u32 %0, tt, %3; }" : "=r"(x[i + 1]) : "r"(u), "r"(t), "r"(ub));
                }
            }
            else if (V == 15) { // shoup butterfly with the quotient taken from a full-rate IMAD.WIDE (no addend)
                if ((i & 1) == 0) {
                    unsigned long long qq; unsigned q, t;
                    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(qq) : "r"(x[i + 1]), "r"(z[i]));
                    q = (unsigned)(qq >> 32);
                    asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(t) : "r"(x[i + 1]), "r"(y[i]));
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(t) : "r"(q), "r"(ua));
                    unsigned u = x[i];
                    asm volatile("add.u32 %0, %1, %2;" : "=r"(x[i]) : "r"(u), "r"(t));
                    asm volatile("{ .reg .u32 tt; sub.u32 tt, %1, %2; add.u32 %0, tt, %3; }" : "=r"(x[i + 1]) : "r"(u), "r"(t), "r"(ub));
                }
            }
            else if (V == 16) { // MAC: full-rate IMAD.WIDE (no addend) + 64-bit add on the ALU pipe
                unsigned long long pp;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(pp) : "r"(x[i]), "r"(y[i]));
                asm volatile("add.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(pp));
                x[i] += (unsigned)w[i];
            }
            else if (V == 17) { // MAC: IMAD.WIDE with 64-bit addend
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x[i]), "r"(y[i]));
                x[i] += (unsigned)w[i];
            }
            else if (V == 18 || V == 19 || V == 20) {
                // pipe-overlap probe: V18 = even warps integer butterflies / odd warps DFMA; V19 = butterflies only in
                // even warps (odd idle); V20 = DFMA only in odd warps (even idle)
                const bool int_warp = ((threadIdx.x >> 5) & 1) == 0;
                if (int_warp && V != 20) {
                    if ((i & 1) == 0) {
                        unsigned q, t;
                        asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(q) : "r"(x[i + 1]), "r"(z[i]));
                        asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(t) : "r"(x[i + 1]), "r"(y[i]));
                        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(t) : "r"(q), "r"(ua));
                        unsigned u = x[i];
                        asm volatile("add.u32 %0, %1, %2;" : "=r"(x[i]) : "r"(u), "r"(t));
                        asm volatile("{ .reg .u32 tt; sub.u32 tt, %1, %2; add.u32 %0, tt, %3; }" : "=r"(x[i + 1]) : "r"(u), "r"(t), "r"(ub));
                    }
                }
                else if (!int_warp && V != 19) {
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dx[i]) : "d"(dy), "d"(dz));
                }
            }
            else if (V == 14) { // 1 IMAD (R,R,R) + 1 IADD (R,R) interleaved
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(z[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(z[(i + 3) % ILP]) : "r"(y[i]));
            }
        }
    }
    long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i] + y[i] + z[i] + (unsigned)w[i] + (unsigned)(w[i] >> 32) + (unsigned)dx[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

struct Res { const char* name; double gops, per_sm_clk, per_sm_clk_wall, ms; int resident; };
static unsigned *g_out, *g_in; static long long* g_cyc; static int g_sms; static double g_clock_hz;

template <int V>
Res run(const char* name, double ops, int ctas_per_sm) {
    int grid = g_sms * ctas_per_sm, block = 256, iters = 20000;
    cudaEvent_t e0, e1; CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    k<V><<<grid, block>>>(g_out, g_in, iters, 3, 0x7ffe001u, g_cyc);
    CHECK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        CHECK(cudaEventRecord(e0));
        k<V><<<grid, block>>>(g_out, g_in, iters, 3, 0x7ffe001u, g_cyc);
        CHECK(cudaEventRecord(e1)); CHECK(cudaEventSynchronize(e1));
        float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    std::vector<long long> hc(grid);
    CHECK(cudaMemcpy(hc.data(), g_cyc, grid * 8, cudaMemcpyDeviceToHost));
    double avg = 0; for (auto c : hc) avg += (double)c; avg /= grid;
    Res r; r.name = name; r.ms = best;
    r.gops = (double)grid * block * iters * ops / (best * 1e-3) / 1e9;
    // CTAs actually co-resident on an SM (the ~60-register probes fit 4 of 256 threads, not the 8 launched per SM): the
    // launched CTAs run in ceil(ctas_per_sm / resident) rounds of `avg` cycles each
    int resident = 0;
    CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k<V>, block, 0));
    if (resident > ctas_per_sm) resident = ctas_per_sm;
    if (resident < 1) resident = 1;
    const int rounds = (ctas_per_sm + resident - 1) / resident;
    r.resident = resident;
    r.per_sm_clk = (double)ctas_per_sm * block * iters * ops / (avg * rounds);
    r.per_sm_clk_wall = r.gops * 1e9 / g_sms / g_clock_hz;   // from the wall clock and the maximum SM clock
    return r;
}

int main(int argc, char** argv) {
    cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0)); g_sms = p.multiProcessorCount;
    { int khz = 0; CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0)); g_clock_hz = khz * 1e3; }
    int maxgrid = g_sms * 8;
    CHECK(cudaMalloc(&g_out, (size_t)maxgrid * 256 * 4)); CHECK(cudaMalloc(&g_cyc, maxgrid * 8));
    std::vector<unsigned> hin(256 * 24); for (size_t i = 0; i < hin.size(); i++) hin[i] = (unsigned)(i * 2654435761u + 12345u);
    CHECK(cudaMalloc(&g_in, hin.size() * 4)); CHECK(cudaMemcpy(g_in, hin.data(), hin.size() * 4, cudaMemcpyHostToDevice));
    // warm the clocks for ~1 s
    for (int i = 0; i < 150; i++) k<0><<<maxgrid, 256>>>(g_out, g_in, 20000, 3, 5, g_cyc);
    CHECK(cudaDeviceSynchronize());
    std::vector<Res> rs;
    for (int occ : {8, 2}) {
        rs.push_back(run<0>(occ == 8 ? "imad_rrr" : "imad_rrr_occ2", ILP, occ));
        rs.push_back(run<5>(occ == 8 ? "imadhi_rr" : "imadhi_rr_occ2", ILP, occ));
        rs.push_back(run<12>(occ == 8 ? "butterfly_regtw" : "butterfly_regtw_occ2", 5 * (ILP / 2), occ));
        rs.push_back(run<13>(occ == 8 ? "butterfly_unitw" : "butterfly_unitw_occ2", 5 * (ILP / 2), occ));
    }
    rs.push_back(run<1>("imul_rr", ILP, 8));
    rs.push_back(run<2>("imad_r_ur_r", ILP, 8));
    rs.push_back(run<3>("imad_r_r_ur", ILP, 8));
    rs.push_back(run<4>("imad_r_ur_ur", ILP, 8));
    rs.push_back(run<6>("imadhi_r_ur", ILP, 8));
    rs.push_back(run<7>("imadwide_acc", ILP, 8));
    rs.push_back(run<8>("imulwide", ILP, 8));
    rs.push_back(run<9>("iadd_rr", ILP, 8));
    rs.push_back(run<10>("lop3_rrr", ILP, 8));
    rs.push_back(run<11>("dfma", ILP, 8));
    rs.push_back(run<14>("imad_rrr+iadd", 2 * ILP, 8));
    rs.push_back(run<18>("overlap_int_even_dfma_odd", 1, 8));
    rs.push_back(run<19>("overlap_int_even_only", 1, 8));
    rs.push_back(run<20>("overlap_dfma_odd_only", 1, 8));
    rs.push_back(run<15>("butterfly_widehi", 5 * (ILP / 2), 8));
    rs.push_back(run<15>("butterfly_widehi_occ2", 5 * (ILP / 2), 2));
    rs.push_back(run<16>("mac_mulwide_add64", ILP, 8));
    rs.push_back(run<17>("mac_madwide", ILP, 8));
    rs.push_back(run<16>("mac_mulwide_add64_occ2", ILP, 2));
    rs.push_back(run<17>("mac_madwide_occ2", ILP, 2));
    FILE* f = argc > 1 ? fopen(argv[1], "w") : stdout;
    fprintf(f, "{\"gpu\": \"%s\", \"sms\": %d, \"variants\": {", p.name, g_sms);
    for (size_t i = 0; i < rs.size(); i++)
        fprintf(f, "%s\"%s\": {\"gops\": %.1f, \"per_sm_per_clk\": %.2f, \"per_sm_per_clk_wall\": %.2f, \"resident_ctas\": %d, \"ms\": %.3f}", i ? ", " : "", rs[i].name, rs[i].gops, rs[i].per_sm_clk, rs[i].per_sm_clk_wall, rs[i].resident, rs[i].ms);
    fprintf(f, "}}\n"); if (f != stdout) fclose(f);
    return 0;
}
