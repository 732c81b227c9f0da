// C ABI of libtfhe_b200.so (see include/tfhe_b200.h) and the host-side orchestration of the batched operations.
//
// The control flow of every batched method re-expresses the reference's host code
// (binfhe-base-scheme.cpp:598-1277) as a sequence of device kernels over device-resident ciphertext batches:
// nothing returns to the host between the bootstraps of one call (the reference round-trips through
// std::vector<LWECiphertext> twice per bootstrap, bootstrapping.cu:1616-1667,1877-1905).
#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <errno.h>
#include <nvtx3/nvToolsExt.h>
#include <sys/random.h>

#include "engine.cuh"

using namespace tfhe_b200;

static_assert(sizeof(tfhe_b200_params) == 88, "tfhe_b200_params layout is part of the ABI");
static_assert(sizeof(tfhe_b200_stats) == 32, "tfhe_b200_stats layout is part of the ABI");

static thread_local std::string g_err;

// NVTX range around a phase of a call (host side; shows up in Nsight Systems timelines, a no-op without a profiler)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

#define CUDA_TRY(x)                                                                                        \
    do {                                                                                                   \
        cudaError_t e__ = (x);                                                                             \
        if (e__ != cudaSuccess) {                                                                          \
            char buf__[512];                                                                               \
            snprintf(buf__, sizeof(buf__), "%s:%d: %s failed: %s", __FILE__, __LINE__, #x,                 \
                     cudaGetErrorString(e__));                                                             \
            g_err = buf__;                                                                                 \
            return TFHE_B200_ECUDA;                                                                        \
        }                                                                                                  \
    } while (0)

#define FAIL(code, msg)  \
    do {                 \
        g_err = (msg);   \
        return (code);   \
    } while (0)

namespace {

struct Arena {
    unsigned char* base = nullptr;
    size_t cap = 0, off = 0;
};

// Handle-owned pinned staging for pageable host buffers (the reference keeps acc_host / ctExt_host pinned for 65536
// ciphertexts, bootstrapping.cu:904-905): grow-only, carved per call like the device arena.
struct Pinned {
    unsigned char* base = nullptr;
    size_t cap = 0, off = 0;
};
struct PendingOut {   // staged device->host copy: finished by a host memcpy once the stream has drained
    void* dst;
    const void* src;
    size_t bytes;
};

// One host worker thread per GPU of a multi-GPU handle: the per-device body of a sharded call (launches, staging
// memcpys, the final stream synchronisation) runs on it, so GPU k+1 is fed while GPU k is still being fed / drained.
// The reference drives all GPUs from a single host thread (bootstrapping.cu:1616-1667).
class Worker {
  public:
    Worker() : th_([this] { loop(); }) {}
    ~Worker() {
        {
            std::lock_guard<std::mutex> l(m_);
            quit_ = true;
        }
        cv_.notify_all();
        th_.join();
    }
    void submit(std::function<int()> f) {
        {
            std::lock_guard<std::mutex> l(m_);
            job_ = std::move(f);
            has_job_ = true;
            done_ = false;
        }
        cv_.notify_all();
    }
    int wait(std::string* err) {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [this] { return done_; });
        if (rc_ && err)
            *err = err_;
        return rc_;
    }

  private:
    void loop();
    std::mutex m_;
    std::condition_variable cv_;
    std::function<int()> job_;
    bool has_job_ = false, done_ = true, quit_ = false;
    int rc_ = 0;
    std::string err_;
    std::thread th_;
};

constexpr int MAX_CHUNKS = 10;   // chunks of the pipelined host-buffer path

struct Dev {
    int id = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t xfer_in = nullptr, xfer_out = nullptr;   // host<->device copies of the pipelined host-buffer path
    cudaEvent_t ev[6] = {};
    unsigned ev_mask = 0;                                 // which of ev[] were recorded by the current call
    cudaEvent_t pev[3 * MAX_CHUNKS + 1] = {};             // per-chunk hand-over events (no timing)
    // key material / tables
    void* bk_generic = nullptr;
    u32* bk_cggi32 = nullptr;
    void* tw_fwd = nullptr;
    void* tw_inv = nullptr;
    void* psi_pow = nullptr;
    void* sh_fwd = nullptr;
    void* sh_inv = nullptr;
    u32* twB = nullptr;
    u64* bk_cggi64 = nullptr;
    u64* twB64 = nullptr;
    u64* tw32_64 = nullptr;
    u64* twU64 = nullptr;
    u64* twCw = nullptr;   // wide 64-bit kernel tables
    u64* twBw = nullptr;
    u64* twUw = nullptr;
    void* ksk = nullptr;
    u64* ks_partial = nullptr;   // column-sum accumulator of split key switches (batches below one ciphertext per SM)
    void* pers_state = nullptr;  // persistent blind rotation: hand-over slots of split groups, one per SM (br_cggi32.cu)
    size_t pers_slot_bytes = 0;
    u32* pers_flags = nullptr;   // ... and their completion flags (hold the epoch of the launch that filled the slot);
                                 // word [sm_count] counts the persistent CTAs started so far (range tickets)
    u32 pers_epoch = 0;
    u32 pers_tickets = 0;        // host shadow of that counter: its value when the next launch starts
    Arena ws;
    Pinned pin;                         // pinned host staging (inputs and outputs of the current call)
    std::vector<PendingOut> pending;    // staged outputs to hand to the caller after the stream has drained
    std::unique_ptr<Worker> worker;     // multi-GPU handles only
};

void Worker::loop() {
    std::unique_lock<std::mutex> l(m_);
    for (;;) {
        cv_.wait(l, [this] { return has_job_ || quit_; });
        if (quit_)
            return;
        std::function<int()> f = std::move(job_);
        has_job_ = false;
        l.unlock();
        g_err.clear();
        int rc = f();
        l.lock();
        rc_ = rc;
        err_ = g_err;   // g_err is thread-local: hand the message to the calling thread
        done_ = true;
        cv_.notify_all();
    }
}

}  // namespace

struct tfhe_b200_handle {
    tfhe_b200_params p;
    bool is64 = false;
    bool have_cggi32 = false;
    bool skip_top = false;
    bool have_dm32 = false;
    bool have_dm64w = false;     // AP/DM on the N = 2048 rings (br_dm64w.cu), 64-bit words
    bool have_cggi64 = false;
    bool have_cggi64w = false;   // wide variant usable (skip-top path of a supported ring)
    std::vector<u64> twA64_host;
    int force_generic = 0;
    bool keep_generic = false;   // generic key layout retained beside the specialised one (cross-check kernel)
    int group = 0;  // ciphertexts per CTA of the cggi32 kernel (0 = default)
    int pers_mode = 1;   // persistent blind rotation: 0 = never, 1 = whenever a plain launch would end on a partial wave
    int pers_ctas = 0;   // > 0: force the persistent variant with this many CTAs (tests: splits at arbitrary steps)
    u32 logN = 0, d = 0, gBits = 0;
    ModCtx<u32> m32;
    ModCtx<u64> m64;
    int ksk_bytes = 8;
    u32 row_stride = 0;
    size_t bk_words = 0, ksk_words = 0;
    std::vector<Dev> devs;
    std::vector<u32> twA_host;
    std::string variant;
    std::mutex mu;
    // additional bootstrapping-key sets (own BK + KSK) for other gadget bases, keyed by baseG;
    // complete child handles on the same devices that borrow this handle's streams (binfhecontext.h:437 m_BTKey_map)
    std::map<u32, tfhe_b200_handle*> key_map;
    bool borrowed_streams = false;
};

// ---------------------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------------------
static size_t bk_words_of(const tfhe_b200_params* p) {
    size_t N = p->N, n = p->n;
    if (p->method == TFHE_B200_METHOD_GINX)
        return 2 * n * (size_t)(2 * (p->digitsG - p->numDigitsToThrow)) * 2 * N;
    return n * (size_t)p->baseR * p->digitsR * (size_t)(2 * p->digitsG) * 2 * N;
}
static size_t ksk_words_of(const tfhe_b200_params* p) {
    return (size_t)p->N * p->baseKS * p->dKS * (p->n + 1);
}

static int arena_reserve(Dev& d, size_t bytes) {
    if (bytes <= d.ws.cap) {
        d.ws.off = 0;
        return 0;
    }
    CUDA_TRY(cudaSetDevice(d.id));
    CUDA_TRY(cudaStreamSynchronize(d.stream));
    if (d.ws.base)
        CUDA_TRY(cudaFree(d.ws.base));
    d.ws.base = nullptr;
    d.ws.cap = 0;
    size_t cap = bytes + (bytes >> 3) + (1 << 20);
    if (cudaMalloc((void**)&d.ws.base, cap) != cudaSuccess) {
        cudaGetLastError();
        d.ws.base = nullptr;
        FAIL(TFHE_B200_ENOMEM, "workspace: cudaMalloc of " + std::to_string(cap) + " bytes failed on device " +
                                   std::to_string(d.id));
    }
    d.ws.cap = cap;
    d.ws.off = 0;
    return 0;
}
// nullptr when the reservation was computed too small (a programming error in this file): the caller reports
// TFHE_B200_ENOMEM instead of handing a kernel a bad pointer -- the library never exits the process
template <typename T>
static T* arena_take(Dev& d, size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
    if (d.ws.off + bytes > d.ws.cap) {
        char buf[160];
        snprintf(buf, sizeof(buf), "workspace arena overrun (%zu + %zu > %zu)", d.ws.off, bytes, d.ws.cap);
        g_err = buf;
        return nullptr;
    }
    T* p = reinterpret_cast<T*>(d.ws.base + d.ws.off);
    d.ws.off += bytes;
    return p;
}
#define TAKE(var, T, d, count)              \
    T* var = arena_take<T>(d, count);       \
    if (!var)                               \
        return TFHE_B200_ENOMEM

// ---- pinned staging --------------------------------------------------------------------------------------------
static const size_t PIN_LIMIT = (size_t)3 << 30;   // larger calls copy straight from / to the caller's pageable memory

// true when the CUDA driver can DMA from / to this host pointer directly (cudaHostAlloc / cudaHostRegister memory)
static bool host_is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}
// reserve `bytes` of pinned staging for the current call (grow-only); failure is not an error: the copies then go
// through the driver's own pageable path
static void pin_reserve(Dev& d, size_t bytes) {
    d.pin.off = 0;
    if (bytes <= d.pin.cap || bytes > PIN_LIMIT)
        return;
    cudaSetDevice(d.id);
    cudaStreamSynchronize(d.stream);
    if (d.pin.base)
        cudaFreeHost(d.pin.base);
    d.pin.base = nullptr;
    d.pin.cap = 0;
    size_t cap = bytes + (bytes >> 3) + (1 << 20);
    if (cudaHostAlloc((void**)&d.pin.base, cap, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        d.pin.base = nullptr;
        return;
    }
    d.pin.cap = cap;
}
static unsigned char* pin_take(Dev& d, size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (!d.pin.base || d.pin.off + bytes > d.pin.cap)
        return nullptr;
    unsigned char* p = d.pin.base + d.pin.off;
    d.pin.off += bytes;
    return p;
}
// host memcpy on a few threads (a single core moves ~10 GB/s, less than one GPU's PCIe link)
static void par_memcpy(void* dst, const void* src, size_t bytes) {
    const size_t blk = (size_t)2 << 20;
    if (bytes < 2 * blk) {
        memcpy(dst, src, bytes);
        return;
    }
    const long nblk = (long)((bytes + blk - 1) / blk);
#pragma omp parallel for num_threads(4) schedule(static)
    for (long b = 0; b < nblk; b++) {
        const size_t o = (size_t)b * blk;
        memcpy((char*)dst + o, (const char*)src + o, std::min(blk, bytes - o));
    }
}

static u32 bitrev32(u32 x, u32 bits) {
    u32 r = 0;
    for (u32 i = 0; i < bits; i++) {
        r = (r << 1) | (x & 1);
        x >>= 1;
    }
    return r;
}

// shard [start, start+count) of the batch owned by device k of nd
static void shard_of(int batch, int nd, int k, int* start, int* count) {
    int base = batch / nd, rem = batch % nd;
    *start = k * base + (k < rem ? k : rem);
    *count = base + (k < rem ? 1 : 0);
}

// Host-space copies of pageable memory go through the handle's pinned staging when the call reserved room for them
// (pin_reserve): a host memcpy + a true asynchronous DMA instead of the driver's blocking pageable path.
static int copy_in(Dev& d, Dev& d0, void* dst, const void* src, size_t bytes, int space) {
    if (!bytes)
        return 0;
    if (space == TFHE_B200_HOST) {
        unsigned char* st = (bytes >= (64 << 10) && !host_is_pinned(src)) ? pin_take(d, bytes) : nullptr;
        if (st) {
            par_memcpy(st, src, bytes);
            src = st;
        }
        CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, d.stream));
    }
    else if (d.id == d0.id)
        CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, d.stream));
    else
        CUDA_TRY(cudaMemcpyPeerAsync(dst, d.id, src, d0.id, bytes, d.stream));
    return 0;
}
static int copy_out(Dev& d, Dev& d0, void* dst, const void* src, size_t bytes, int space) {
    if (!bytes)
        return 0;
    if (space == TFHE_B200_HOST) {
        unsigned char* st = (bytes >= (64 << 10) && !host_is_pinned(dst)) ? pin_take(d, bytes) : nullptr;
        if (st) {
            d.pending.push_back(PendingOut{dst, st, bytes});   // finished by run_sharded after the stream has drained
            dst = st;
        }
        CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, d.stream));
    }
    else if (d.id == d0.id)
        CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, d.stream));
    else
        CUDA_TRY(cudaMemcpyPeerAsync(dst, d0.id, src, d.id, bytes, d.stream));
    return 0;
}
// phase marker i of the current call (timing events of tfhe_b200_stats)
static cudaError_t rec_ev(Dev& d, int i) {
    d.ev_mask |= 1u << i;
    return cudaEventRecord(d.ev[i], d.stream);
}

// ---------------------------------------------------------------------------------------------------------
// setup
// ---------------------------------------------------------------------------------------------------------
template <typename T>
static int build_tables(tfhe_b200_handle* h, Dev& d, const ModCtx<T>& M) {
    const tfhe_b200_params& p = h->p;
    const u64 Q = p.Q, N = p.N;
    std::vector<T> fw(N), iv(N), pp(2 * N);
    u64 psi = p.psi % Q, psii = h_powmod(psi, Q - 2, Q);
    u64 x = 1, xi = 1;
    for (u64 k = 0; k < N; k++) {
        u32 r = bitrev32((u32)k, h->logN);
        fw[r] = to_mont<T>(x, M);
        iv[r] = to_mont<T>(xi, M);
        x = h_mulmod(x, psi, Q);
        xi = h_mulmod(xi, psii, Q);
    }
    x = 1;
    for (u64 k = 0; k < 2 * N; k++) {
        pp[k] = to_mont<T>(x, M);
        x = h_mulmod(x, psi, Q);
    }
    CUDA_TRY(cudaMalloc(&d.tw_fwd, N * sizeof(T)));
    CUDA_TRY(cudaMalloc(&d.tw_inv, N * sizeof(T)));
    CUDA_TRY(cudaMalloc(&d.psi_pow, 2 * N * sizeof(T)));
    CUDA_TRY(cudaMemcpy(d.tw_fwd, fw.data(), N * sizeof(T), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d.tw_inv, iv.data(), N * sizeof(T), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d.psi_pow, pp.data(), 2 * N * sizeof(T), cudaMemcpyHostToDevice));
    // plain twiddles + Shoup companions for the lazy 64-bit transform
    {
        std::vector<T> sf(2 * N), si(2 * N);
        const int w = sizeof(T) * 8;
        x = 1; xi = 1;
        for (u64 k = 0; k < N; k++) {
            u32 r = bitrev32((u32)k, h->logN);
            sf[r] = (T)x;
            sf[N + r] = (T)((((unsigned __int128)x) << w) / Q);
            si[r] = (T)xi;
            si[N + r] = (T)((((unsigned __int128)xi) << w) / Q);
            x = h_mulmod(x, psi, Q);
            xi = h_mulmod(xi, psii, Q);
        }
        CUDA_TRY(cudaMalloc(&d.sh_fwd, 2 * N * sizeof(T)));
        CUDA_TRY(cudaMalloc(&d.sh_inv, 2 * N * sizeof(T)));
        CUDA_TRY(cudaMemcpy(d.sh_fwd, sf.data(), 2 * N * sizeof(T), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(d.sh_inv, si.data(), 2 * N * sizeof(T), cudaMemcpyHostToDevice));
    }
    return 0;
}

// Host-side producer of the two flat key arrays (element order of tfhe_b200_setup), chunk by chunk: either the caller's
// flat arrays, or an index over OpenFHE's serialized streams (serial_reader.cu) -- the staging loop below does not care.
namespace {
struct KeySource {
    virtual ~KeySource() {}
    virtual void bk_words(u64* dst, size_t off, size_t cnt) const = 0;
    virtual void ksk_rows(u64* dst, size_t row0, size_t nrows, u32 words) const = 0;
};
struct FlatKeys : KeySource {
    const u64 *bk, *ksk;
    FlatKeys(const u64* b, const u64* k) : bk(b), ksk(k) {}
    void bk_words(u64* dst, size_t off, size_t cnt) const override { par_memcpy(dst, bk + off, cnt * 8); }
    void ksk_rows(u64* dst, size_t row0, size_t nrows, u32 words) const override {
        par_memcpy(dst, ksk + row0 * words, nrows * words * 8);
    }
};
struct SerializedKeys : KeySource {
    SerializedAccKey acc;
    SerializedSwitchKey sw;
    void bk_words(u64* dst, size_t off, size_t cnt) const override {
        const size_t blk = (size_t)acc.N * 16;   // a few polynomials per task
        const long nblk = (long)((cnt + blk - 1) / blk);
#pragma omp parallel for num_threads(4) schedule(static) if (nblk > 8)
        for (long b = 0; b < nblk; b++) {
            const size_t o = (size_t)b * blk;
            acc.gather(dst + o, off + o, std::min(blk, cnt - o));
        }
    }
    void ksk_rows(u64* dst, size_t row0, size_t nrows, u32 words) const override {
        const size_t blk = 256;
        const long nblk = (long)((nrows + blk - 1) / blk);
#pragma omp parallel for num_threads(4) schedule(static) if (nblk > 8)
        for (long b = 0; b < nblk; b++) {
            const size_t o = (size_t)b * blk;
            sw.gather_rows(dst + o * words, row0 + o, std::min(blk, nrows - o));
        }
    }
};
}  // namespace

// upload + re-encode the keys on device 0.  Device-resident keys (key_space = DEVICE: bk_dev / ksk_dev) are converted in
// place; host keys stream through two pinned 32 MiB chunks: the host fills chunk c+1 (memcpy from the caller's arrays,
// or a gather from the serialized stream) while the GPU copies and converts chunk c.
template <typename T>
static int encode_keys(tfhe_b200_handle* h, Dev& d, const ModCtx<T>& M, const KeySource* src, const u64* bk, const u64* ksk,
                       int key_space) {
    const tfhe_b200_params& p = h->p;
    const u64 Q = p.Q;
    u64 ninv = h_powmod(p.N, Q - 2, Q);
    u64 R = (u64)M.oneM;
    T ninvM2 = (T)h_mulmod(h_mulmod(ninv, R, Q), R, Q);  // N^-1 * R^2: mont_mul(x, .) = x * N^-1 * R
    const size_t stage_words = (size_t)4 << 20;           // 32 MiB staging chunks for host-resident keys
    u64* stage[2] = {nullptr, nullptr};                   // device
    u64* hstage[2] = {nullptr, nullptr};                  // pinned host
    cudaEvent_t ev[2] = {nullptr, nullptr};
    if (key_space == TFHE_B200_HOST)
        for (int k = 0; k < 2; k++) {
            CUDA_TRY(cudaMalloc((void**)&stage[k], stage_words * sizeof(u64)));
            CUDA_TRY(cudaHostAlloc((void**)&hstage[k], stage_words * sizeof(u64), cudaHostAllocDefault));
            CUDA_TRY(cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming));
        }
    int chunk = 0;
    // fills pinned slot, uploads it and returns the device staging pointer; the caller launches the conversion and then
    // calls done(slot)
    auto upload = [&](int slot, size_t words) -> int {
        CUDA_TRY(cudaMemcpyAsync(stage[slot], hstage[slot], words * sizeof(u64), cudaMemcpyHostToDevice, d.stream));
        return 0;
    };

    // ---- bootstrapping key, generic layout (same element order as the source) ----
    CUDA_TRY(cudaMalloc(&d.bk_generic, h->bk_words * sizeof(T)));
    if (key_space == TFHE_B200_DEVICE)
        CUDA_TRY(launch_bk_convert_generic<T>((T*)d.bk_generic, bk, h->bk_words, M, ninvM2, d.stream));
    else {
        for (size_t off = 0; off < h->bk_words; off += stage_words, chunk++) {
            const int slot = chunk & 1;
            const size_t cnt = std::min(h->bk_words - off, stage_words);
            CUDA_TRY(cudaEventSynchronize(ev[slot]));          // the previous user of this slot has been consumed
            src->bk_words(hstage[slot], off, cnt);
            int r = upload(slot, cnt);
            if (r) return r;
            CUDA_TRY(launch_bk_convert_generic<T>((T*)d.bk_generic + off, stage[slot], cnt, M, ninvM2, d.stream));
            CUDA_TRY(cudaEventRecord(ev[slot], d.stream));
        }
    }
    // ---- key switching key: narrowest word, 16-byte padded rows ----
    {
        const size_t rows = (size_t)p.N * p.baseKS * p.dKS;
        const u32 words = p.n + 1;
        CUDA_TRY(cudaMalloc(&d.ksk, rows * h->row_stride * h->ksk_bytes));
        if (key_space == TFHE_B200_DEVICE)
            CUDA_TRY(launch_ksk_convert(d.ksk, h->ksk_bytes, h->row_stride, ksk, rows, words, d.stream));
        else {
            const size_t rows_per = stage_words / words;
            for (size_t r0 = 0; r0 < rows; r0 += rows_per, chunk++) {
                const int slot = chunk & 1;
                const size_t cnt = std::min(rows - r0, rows_per);
                CUDA_TRY(cudaEventSynchronize(ev[slot]));
                src->ksk_rows(hstage[slot], r0, cnt, words);
                int r = upload(slot, cnt * words);
                if (r) return r;
                CUDA_TRY(launch_ksk_convert((unsigned char*)d.ksk + r0 * h->row_stride * h->ksk_bytes, h->ksk_bytes,
                                            h->row_stride, stage[slot], cnt, words, d.stream));
                CUDA_TRY(cudaEventRecord(ev[slot], d.stream));
            }
        }
    }
    CUDA_TRY(cudaStreamSynchronize(d.stream));
    for (int k = 0; k < 2; k++) {
        if (stage[k])
            CUDA_TRY(cudaFree(stage[k]));
        if (hstage[k])
            CUDA_TRY(cudaFreeHost(hstage[k]));
        if (ev[k])
            CUDA_TRY(cudaEventDestroy(ev[k]));
    }
    return 0;
}

// the specialised 32-bit CGGI layout is derived on the device from the (already Montgomery-encoded) generic copy
__global__ void bk_relayout_cggi32_kernel(u32* dst, const u32* src, u32 n, u32 d, u32 N, ModCtx<u32> M, int skip,
                                          const u32* cM /* [d/2] Montgomery constants, see below */) {
    // destination [i][x][k][c]: word w = 4x + c = (key*d + l')*2 + j of evaluation slot k  (see br_cggi32.cu)
    // With top-digit elimination (skip != 0) the rows are transformed (row l' = jin + 2l, top = d/2 - 1):
    //   l < top : BK'_l = BK_l - B^(l-top) * BK_top        (cM[l]   = B^(l-top) in Montgomery form)
    //   l = top : BK'_top = N * B^-top * BK_top            (cM[top] = N * B^-top in Montgomery form; the kernel
    //             multiplies this row by acc_eval / N because every stored key word already carries N^-1)
    const size_t total = (size_t)n * N * 2 * d * 2;
    const u32 top = d / 2 - 1;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        size_t r = idx;
        u32 c = r % 4; r /= 4;
        u32 k = r % N; r /= N;
        u32 x = r % d; r /= d;     // 4d words per slot = d planes
        u32 i = (u32)r;
        u32 w = 4 * x + c;
        u32 j = w % 2, lp = (w / 2) % d, key = w / (2 * d);
        const size_t base = (((size_t)key * n + i) * d) * 2 * N;
        u32 val = src[base + ((size_t)lp * 2 + j) * N + k];
        if (skip) {
            const u32 jin = lp & 1, l = lp >> 1;
            const u32 vt = src[base + ((size_t)(jin + 2 * top) * 2 + j) * N + k];
            const u32 t = M.mont_mul(vt, cM[l]);
            val = (l == top) ? t : M.sub(val, t);
        }
        dst[idx] = val;
    }
}

// 64-bit CGGI layout: destination [i][x(2d)][slot][2], word w = 2x + c = (key*d + l')*2 + j; optional top-digit transform
__global__ void bk_relayout_cggi64_kernel(u64* dst, const u64* src, u32 n, u32 d, u32 N, ModCtx<u64> M, int skip,
                                          const u64* cM) {
    const size_t total = (size_t)n * N * 2 * d * 2;
    const u32 top = d / 2 - 1;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        size_t r = idx;
        u32 c = r % 2; r /= 2;
        u32 k = r % N; r /= N;
        u32 x = r % (2 * d); r /= (2 * d);
        u32 i = (u32)r;
        u32 w = 2 * x + c;
        u32 j = w % 2, lp = (w / 2) % d, key = w / (2 * d);
        const size_t base = (((size_t)key * n + i) * d) * 2 * N;
        u64 val = src[base + ((size_t)lp * 2 + j) * N + k];
        if (skip) {
            const u32 jin = lp & 1, l = lp >> 1;
            const u64 vt = src[base + ((size_t)(jin + 2 * top) * 2 + j) * N + k];
            const u64 t = M.mont_mul(vt, cM[l]);
            val = (l == top) ? t : M.sub(val, t);
        }
        dst[idx] = (val & ((1ULL << 27) - 1)) | ((val >> 27) << 32);   // 27-bit limbs for the kernel's Karatsuba MAC
    }
}

// AP/DM variant of the re-layout: source [row][l'(d)][j(2)][N] (row = (i*baseR + a0)*digitsR + k), destination
// [row][x(d/2)][slot][4] with word w = l'*2 + j.  The DM accumulator drops row l' = 0 (rgsw-acc-dm.cpp:353,357) and the
// kernel eliminates the top digit, so:  l < top: BK' = [l' >= 1] BK_l' - B^(l-top) BK_top(jin);  l = top: N B^-top BK_top.
__global__ void bk_relayout_dm32_kernel(u32* dst, const u32* src, size_t rows, u32 d, u32 N, ModCtx<u32> M,
                                        const u32* cM, int skip) {
    const size_t per_row = (size_t)d * 2 * N;
    const size_t total = rows * per_row;
    const u32 top = d / 2 - 1;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        size_t r = idx;
        u32 c = r % 4; r /= 4;
        u32 k = r % N; r /= N;
        u32 x = r % (d / 2); r /= (d / 2);
        size_t row = r;
        u32 w = 4 * x + c;
        u32 j = w % 2, lp = w / 2;
        const u32 jin = lp & 1, l = lp >> 1;
        const size_t base = row * per_row;
        if (!skip) {   // plain path: the key's own rows, row l' = 0 never enters the sum (rgsw-acc-dm.cpp:353,357)
            dst[idx] = lp >= 1 ? src[base + ((size_t)lp * 2 + j) * N + k] : 0;
            continue;
        }
        const u32 vt = src[base + ((size_t)(jin + 2 * top) * 2 + j) * N + k];
        const u32 t = M.mont_mul(vt, cM[l]);
        u32 val;
        if (l == top)
            val = t;
        else {
            u32 own = lp >= 1 ? src[base + ((size_t)lp * 2 + j) * N + k] : 0;
            val = M.sub(own, t);
        }
        dst[idx] = val;
    }
}

// 64-bit AP/DM layout (br_dm64w.cu): source [row][l'(d)][j(2)][N], destination [row][x(d)][slot][2] with plane x = l' and
// word c = j, packed as 27-bit limb pairs; top-digit elimination as in bk_relayout_dm32_kernel.
__global__ void bk_relayout_dm64_kernel(u64* dst, const u64* src, size_t rows, u32 d, u32 N, ModCtx<u64> M,
                                        const u64* cM, int skip) {
    const size_t per_row = (size_t)d * 2 * N;
    const size_t total = rows * per_row;
    const u32 top = d / 2 - 1;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        size_t r = idx;
        const u32 j = r % 2; r /= 2;
        const u32 k = r % N; r /= N;
        const u32 lp = r % d; r /= d;
        const size_t base = r * per_row;
        const u32 jin = lp & 1, l = lp >> 1;
        u64 val;
        if (!skip)   // plain path: the key's own rows, row l' = 0 never enters the sum (rgsw-acc-dm.cpp:353,357)
            val = lp >= 1 ? src[base + ((size_t)lp * 2 + j) * N + k] : 0;
        else {
            const u64 vt = src[base + ((size_t)(jin + 2 * top) * 2 + j) * N + k];
            const u64 t = M.mont_mul(vt, cM[l]);
            if (l == top)
                val = t;
            else {
                const u64 own = lp >= 1 ? src[base + ((size_t)lp * 2 + j) * N + k] : 0;
                val = M.sub(own, t);
            }
        }
        dst[idx] = (val & ((1ULL << 27) - 1)) | ((val >> 27) << 32);
    }
}

extern "C" size_t tfhe_b200_bk_words(const tfhe_b200_params* p) {
    return p ? bk_words_of(p) : 0;
}
extern "C" size_t tfhe_b200_ksk_words(const tfhe_b200_params* p) {
    return p ? ksk_words_of(p) : 0;
}
extern "C" const char* tfhe_b200_last_error(void) {
    return g_err.c_str();
}
extern "C" int tfhe_b200_num_gpus(const tfhe_b200_handle* h) {
    return h ? (int)h->devs.size() : 0;
}
extern "C" const char* tfhe_b200_kernel_variant(const tfhe_b200_handle* h) {
    if (!h)
        return "";
    if (h->have_cggi32 && !h->force_generic)
        return h->skip_top ? "cggi_u32_ntt32_skiptop" : "cggi_u32_ntt32";
    if (h->have_dm32 && !h->force_generic)
        return h->skip_top ? "dm_u32_ntt32_skiptop" : "dm_u32_ntt32";
    if (h->have_dm64w && !h->force_generic)
        return h->skip_top ? "dm_u64_ntt16x128_skiptop" : "dm_u64_ntt16x128";
    if (h->have_cggi64 && !h->force_generic)
        return h->have_cggi64w ? (h->skip_top ? "cggi_u64_ntt16x128_skiptop" : "cggi_u64_ntt16x128")
                               : (h->skip_top ? "cggi_u64_ntt32x64_skiptop" : "cggi_u64_ntt32x64");
    return h->variant.c_str();
}
extern "C" int tfhe_b200_set_option(tfhe_b200_handle* h, const char* key, int64_t value) {
    if (!h || !key)
        FAIL(TFHE_B200_EINVAL, "set_option: null argument");
    std::string k(key);
    if (k == "force_generic") {
        if (value && !h->devs.empty() && !h->devs[0].bk_generic)
            FAIL(TFHE_B200_EINVAL, "set_option: force_generic needs the generic key layout, which this handle released "
                                   "(set TFHE_B200_FLAG_KEEP_GENERIC in params.flags or TFHE_B200_KEEP_GENERIC=1 "
                                   "before GPUSetup)");
        h->force_generic = (int)value;
    }
    else if (k == "group")
        h->group = (int)value;
    else if (k == "persistent")
        h->pers_mode = (int)value;
    else if (k == "persistent_ctas")
        h->pers_ctas = (int)value;
    else
        FAIL(TFHE_B200_EINVAL, "set_option: unknown key " + k);
    for (auto& kv : h->key_map)
        tfhe_b200_set_option(kv.second, key, value);
    return 0;
}

static int free_dev(Dev& d, bool borrowed_streams = false) {
    d.worker.reset();   // joins the worker thread (idle between calls)
    cudaSetDevice(d.id);
    if (borrowed_streams)
        d.stream = d.xfer_in = d.xfer_out = nullptr;
    if (d.stream)
        cudaStreamSynchronize(d.stream);
    void* ptrs[] = {d.bk_generic, d.bk_cggi32, d.tw_fwd, d.tw_inv, d.psi_pow, d.sh_fwd, d.sh_inv, d.bk_cggi64, d.twB64, d.tw32_64, d.twU64, d.twCw, d.twBw, d.twUw, d.twB, d.ksk, d.ks_partial, d.pers_state, d.pers_flags, d.ws.base};
    for (void* p : ptrs)
        if (p)
            cudaFree(p);
    if (d.pin.base)
        cudaFreeHost(d.pin.base);
    for (auto& e : d.ev)
        if (e)
            cudaEventDestroy(e);
    for (auto& e : d.pev)
        if (e)
            cudaEventDestroy(e);
    if (d.stream)
        cudaStreamDestroy(d.stream);
    if (d.xfer_in)
        cudaStreamDestroy(d.xfer_in);
    if (d.xfer_out)
        cudaStreamDestroy(d.xfer_out);
    d = Dev();
    return 0;
}

extern "C" int tfhe_b200_clean(tfhe_b200_handle* h) {
    if (!h)
        return 0;  // GPUClean without GPUSetup is a no-op
    for (auto& d : h->devs)
        if (d.stream && !h->borrowed_streams) {
            cudaSetDevice(d.id);
            cudaStreamSynchronize(d.stream);
        }
    for (auto& kv : h->key_map)
        tfhe_b200_clean(kv.second);
    h->key_map.clear();
    for (auto& d : h->devs)
        free_dev(d, h->borrowed_streams);
    delete h;
    return 0;
}

// `src` produces host-resident keys chunk by chunk (key_space = HOST); bk / ksk are the device arrays for key_space = DEVICE
static int setup_impl(const tfhe_b200_params* params, const KeySource* src, const uint64_t* bk, size_t bk_words,
                      const uint64_t* ksk, size_t ksk_words, int key_space, int first_device, int num_gpus,
                      tfhe_b200_handle** out) {
    const tfhe_b200_params& p = *params;
    if (p.N < 16 || (p.N & (p.N - 1)) || p.N > 4096)
        FAIL(TFHE_B200_ENOTSUP, "setup: ring dimension N must be a power of two in [16, 4096]");
    if (p.method != TFHE_B200_METHOD_GINX && p.method != TFHE_B200_METHOD_AP)
        FAIL(TFHE_B200_EINVAL, "setup: method must be AP (1) or GINX (2)");
    if (p.Q >= (1ULL << 55) || !(p.Q & 1))
        FAIL(TFHE_B200_ENOTSUP, "setup: Q must be an odd prime below 2^55 (lazy 128-bit accumulation bound)");
    {
        // the generic 64-bit kernel accumulates d products of lazily reduced transform outputs (< (1 + 2 log N) Q) in
        // 128 bits and reduces once: d (1 + 2 log N) Q must stay below 2^64 (br_generic.cu, pointwise stage)
        u32 lg = 0;
        while ((1u << lg) < p.N)
            lg++;
        const unsigned __int128 bound = (unsigned __int128)(2 * p.digitsG) * (1 + 2 * lg) * p.Q;
        if (p.Q >= (1ULL << 31) && bound >= ((unsigned __int128)1 << 64))
            FAIL(TFHE_B200_ENOTSUP, "setup: 2*digitsG*(1 + 2 log N)*Q must stay below 2^64 (lazy accumulation bound)");
    }
    if ((p.Q - 1) % (2ULL * p.N) != 0 || h_powmod(p.psi, p.N, p.Q) != p.Q - 1)
        FAIL(TFHE_B200_EINVAL, "setup: psi is not a primitive 2N-th root of unity mod Q");
    if (p.baseG == 0 || (p.baseG & (p.baseG - 1)) || p.digitsG <= p.numDigitsToThrow)
        FAIL(TFHE_B200_EINVAL, "setup: gadget base must be a power of two and leave at least one digit");
    if (p.baseKS < 2 || p.dKS == 0 || p.qKS == 0 || p.n == 0)
        FAIL(TFHE_B200_EINVAL, "setup: bad key-switch parameters");
    if (p.method == TFHE_B200_METHOD_AP && (p.baseR < 2 || p.digitsR == 0))
        FAIL(TFHE_B200_EINVAL, "setup: bad refresh base for AP");
    if (bk_words != bk_words_of(&p) || ksk_words != ksk_words_of(&p))
        FAIL(TFHE_B200_EINVAL, "setup: key sizes do not match the parameter set");

    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        FAIL(TFHE_B200_ENODEV, std::string("setup: no CUDA device available (") + cudaGetErrorString(ce) +
                                   "); this engine has no CPU fallback");
    if (first_device < 0 || first_device >= ndev)
        FAIL(TFHE_B200_EINVAL, "setup: first_device out of range");
    if (num_gpus <= 0 || first_device + num_gpus > ndev)
        num_gpus = ndev - first_device;

    auto* h = new tfhe_b200_handle();
    h->p = p;
    // 64-bit words for Q >= 2^31 -- and for the small-modulus N = 2048 rings, whose only specialised kernel is the 64-bit one
    // (N = 2048 under AP: br_dm64w; under GINX with a modulus the 32-bit kernel cannot hold: br_cggi64 / br_cggi64w)
    const bool c32 = p.Q < (1ULL << 31) && cggi32_supported(p) && !getenv("TFHE_B200_NO_CGGI32");
    const bool d32 = p.Q < (1ULL << 31) && dm32_supported(p) && !getenv("TFHE_B200_NO_DM32");
    h->is64 = p.Q >= (1ULL << 31) || (!c32 && cggi64_supported(p) && !getenv("TFHE_B200_NO_CGGI64")) ||
              (!d32 && dm64w_supported(p) && !getenv("TFHE_B200_NO_DM64"));
    while ((1u << h->logN) < p.N)
        h->logN++;
    h->d = (p.method == TFHE_B200_METHOD_GINX) ? 2 * (p.digitsG - p.numDigitsToThrow) : 2 * p.digitsG;
    h->gBits = (u32)std::log2((double)p.baseG);
    h->bk_words = bk_words;
    h->ksk_words = ksk_words;
    h->ksk_bytes = p.qKS <= (1ULL << 16) ? 2 : (p.qKS <= (1ULL << 32) ? 4 : 8);
    {
        u32 vw = 16 / h->ksk_bytes;
        h->row_stride = (p.n + 1 + vw - 1) / vw * vw;
    }
    if (h->is64)
        h->m64 = make_modctx<u64>(p.Q);
    else
        h->m32 = make_modctx<u32>(p.Q);
    h->have_cggi32 = !h->is64 && cggi32_supported(p) && !getenv("TFHE_B200_NO_CGGI32");
    // TFHE_B200_NO_SKIPTOP=1 (debug) keeps the untransformed keys and the full set of forward transforms
    h->have_cggi64 = h->is64 && cggi64_supported(p) && !getenv("TFHE_B200_NO_CGGI64");
    h->have_dm32 = !h->is64 && dm32_supported(p) && !getenv("TFHE_B200_NO_DM32");
    h->have_dm64w = h->is64 && dm64w_supported(p) && !getenv("TFHE_B200_NO_DM64");
    h->skip_top = ((h->have_cggi32 && cggi32_skip_top_ok(p)) || (h->have_dm32 && cggi32_skip_top_ok(p)) || (h->have_dm64w && cggi32_skip_top_ok(p)) ||
                   (h->have_cggi64 && (cggi32_skip_top_ok(p) || cggi_skip_top_wrapfix_ok(p)))) &&
                  !getenv("TFHE_B200_NO_SKIPTOP");
    h->have_cggi64w = h->have_cggi64 && (h->skip_top ? cggi64w_supported(p) : cggi64w_plain_supported(p)) &&
                      !getenv("TFHE_B200_NO_C64W");
    h->variant = h->is64 ? "generic_u64" : "generic_u32";
    if (p.method == TFHE_B200_METHOD_AP)
        h->variant += "_dm";
    h->keep_generic = (p.flags & TFHE_B200_FLAG_KEEP_GENERIC) || getenv("TFHE_B200_KEEP_GENERIC");

    h->devs.resize(num_gpus);
    int rc = 0;
    for (int k = 0; k < num_gpus && rc == 0; k++) {
        Dev& d = h->devs[k];
        d.id = first_device + k;
        auto init = [&]() -> int {
            CUDA_TRY(cudaSetDevice(d.id));
            cudaDeviceProp prop;
            CUDA_TRY(cudaGetDeviceProperties(&prop, d.id));
            d.sm_count = prop.multiProcessorCount;
            CUDA_TRY(cudaMalloc((void**)&d.ks_partial, mkmswitch_partial_bytes(h->row_stride, 2 * d.sm_count)));
            CUDA_TRY(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
            CUDA_TRY(cudaStreamCreateWithFlags(&d.xfer_in, cudaStreamNonBlocking));
            CUDA_TRY(cudaStreamCreateWithFlags(&d.xfer_out, cudaStreamNonBlocking));
            for (auto& e : d.ev)
                CUDA_TRY(cudaEventCreate(&e));
            for (auto& e : d.pev)
                CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            if (num_gpus > 1)
                d.worker.reset(new Worker());
            int r = h->is64 ? build_tables<u64>(h, d, h->m64) : build_tables<u32>(h, d, h->m32);
            if (r)
                return r;
            if (h->have_cggi32 || h->have_dm32) {
                std::vector<u32> twB;
                cggi32_build_tables(p, h->twA_host, twB);
                CUDA_TRY(cudaMalloc((void**)&d.twB, twB.size() * 4));
                CUDA_TRY(cudaMemcpy(d.twB, twB.data(), twB.size() * 4, cudaMemcpyHostToDevice));
            }
            if (h->have_cggi64) {
                std::vector<u64> twB, tw32;
                cggi64_build_tables(p, h->twA64_host, twB, tw32);
                CUDA_TRY(cudaMalloc((void**)&d.twB64, twB.size() * 8));
                CUDA_TRY(cudaMalloc((void**)&d.tw32_64, tw32.size() * 8));
                CUDA_TRY(cudaMemcpy(d.twB64, twB.data(), twB.size() * 8, cudaMemcpyHostToDevice));
                CUDA_TRY(cudaMemcpy(d.tw32_64, tw32.data(), tw32.size() * 8, cudaMemcpyHostToDevice));
                CUDA_TRY(cudaMalloc((void**)&d.twU64, h->twA64_host.size() * 8));
                CUDA_TRY(cudaMemcpy(d.twU64, h->twA64_host.data(), h->twA64_host.size() * 8, cudaMemcpyHostToDevice));
            }
            if (h->have_cggi64w || h->have_dm64w) {
                std::vector<u64> tU, tB, tC;
                cggi64w_build_tables(p, tU, tB, tC);
                CUDA_TRY(cudaMalloc((void**)&d.twUw, tU.size() * 8));
                CUDA_TRY(cudaMalloc((void**)&d.twBw, tB.size() * 8));
                CUDA_TRY(cudaMalloc((void**)&d.twCw, tC.size() * 8));
                CUDA_TRY(cudaMemcpy(d.twUw, tU.data(), tU.size() * 8, cudaMemcpyHostToDevice));
                CUDA_TRY(cudaMemcpy(d.twBw, tB.data(), tB.size() * 8, cudaMemcpyHostToDevice));
                CUDA_TRY(cudaMemcpy(d.twCw, tC.data(), tC.size() * 8, cudaMemcpyHostToDevice));
            }
            return 0;
        };
        rc = init();
    }
    // keys: upload + encode on device 0, then replicate peer-to-peer (the reference re-copies from the host to
    // every GPU, bootstrapping.cu:1007-1069)
    if (rc == 0) {
        Dev& d0 = h->devs[0];
        auto enc = [&]() -> int {
            CUDA_TRY(cudaSetDevice(d0.id));
            int r = h->is64 ? encode_keys<u64>(h, d0, h->m64, src, bk, ksk, key_space)
                            : encode_keys<u32>(h, d0, h->m32, src, bk, ksk, key_space);
            if (r)
                return r;
            // constants of the top-digit elimination: B^(l-top) (l < top) and N * B^-top, Montgomery form (the kernels
            // keep acc_eval scaled by N^-1: compensated in the row it multiplies); see the re-layout kernels above
            const u32 dk = h->d / 2, top = dk - 1;
            auto upload_cM = [&](void** dcM) -> int {
                std::vector<u64> cM(dk);
                const u64 Binv = h_powmod(p.baseG % p.Q, p.Q - 2, p.Q);
                for (u32 l = 0; l < dk; l++) {
                    u64 cst = h_powmod(Binv, l == top ? top : top - l, p.Q);
                    if (l == top)
                        cst = h_mulmod(cst, p.N % p.Q, p.Q);
                    cM[l] = h->is64 ? to_mont<u64>(cst, h->m64) : (u64)to_mont<u32>(cst, h->m32);
                }
                const size_t tsz = h->is64 ? 8 : 4;
                std::vector<u32> cM32(cM.begin(), cM.end());
                CUDA_TRY(cudaMalloc(dcM, dk * tsz));
                CUDA_TRY(cudaMemcpy(*dcM, h->is64 ? (const void*)cM.data() : (const void*)cM32.data(), dk * tsz,
                                    cudaMemcpyHostToDevice));
                return 0;
            };
            void* dcM = nullptr;
            if (h->have_cggi32 || h->have_cggi64 || h->have_dm32 || h->have_dm64w) {
                r = upload_cM(&dcM);
                if (r)
                    return r;
            }
            if (h->have_cggi32) {
                CUDA_TRY(cudaMalloc((void**)&d0.bk_cggi32, h->bk_words * 4));
                bk_relayout_cggi32_kernel<<<148 * 8, 256, 0, d0.stream>>>(d0.bk_cggi32, (const u32*)d0.bk_generic, p.n,
                                                                          h->d, p.N, h->m32, h->skip_top ? 1 : 0,
                                                                          (const u32*)dcM);
                CUDA_TRY(cudaGetLastError());
            }
            if (h->have_cggi64) {
                CUDA_TRY(cudaMalloc((void**)&d0.bk_cggi64, h->bk_words * 8));
                bk_relayout_cggi64_kernel<<<148 * 8, 256, 0, d0.stream>>>(d0.bk_cggi64, (const u64*)d0.bk_generic, p.n,
                                                                          h->d, p.N, h->m64, h->skip_top ? 1 : 0,
                                                                          (const u64*)dcM);
                CUDA_TRY(cudaGetLastError());
            }
            if (h->have_dm32) {
                CUDA_TRY(cudaMalloc((void**)&d0.bk_cggi32, h->bk_words * 4));
                bk_relayout_dm32_kernel<<<148 * 8, 256, 0, d0.stream>>>(d0.bk_cggi32, (const u32*)d0.bk_generic,
                                                                        (size_t)p.n * p.baseR * p.digitsR, h->d, p.N,
                                                                        h->m32, (const u32*)dcM, h->skip_top ? 1 : 0);
                CUDA_TRY(cudaGetLastError());
            }
            if (h->have_dm64w) {
                CUDA_TRY(cudaMalloc((void**)&d0.bk_cggi64, h->bk_words * 8));
                bk_relayout_dm64_kernel<<<148 * 8, 256, 0, d0.stream>>>(d0.bk_cggi64, (const u64*)d0.bk_generic,
                                                                        (size_t)p.n * p.baseR * p.digitsR, h->d, p.N,
                                                                        h->m64, (const u64*)dcM, h->skip_top ? 1 : 0);
                CUDA_TRY(cudaGetLastError());
            }
            CUDA_TRY(cudaStreamSynchronize(d0.stream));
            if (dcM)
                CUDA_TRY(cudaFree(dcM));
            // The generic layout is only the source of the specialised ones: release it once they exist, unless the
            // caller asked to keep the cross-check kernel available (params.flags bit 0 or TFHE_B200_KEEP_GENERIC=1;
            // set_option("force_generic") needs it).  STD128-AP: 2.1 GB, logQ = 17: 513 MB saved per GPU.
            if (!h->keep_generic && (h->have_cggi32 || h->have_cggi64 || h->have_dm32 || h->have_dm64w)) {
                CUDA_TRY(cudaFree(d0.bk_generic));
                d0.bk_generic = nullptr;
            }
            const size_t tsz = h->is64 ? 8 : 4;
            const size_t ksk_bytes_total = (size_t)p.N * p.baseKS * p.dKS * h->row_stride * h->ksk_bytes;
            for (size_t k = 1; k < h->devs.size(); k++) {
                Dev& d = h->devs[k];
                CUDA_TRY(cudaSetDevice(d.id));
                int can = 0;
                cudaDeviceCanAccessPeer(&can, d.id, d0.id);
                if (can) {
                    cudaError_t pe = cudaDeviceEnablePeerAccess(d0.id, 0);
                    if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled)
                        CUDA_TRY(pe);
                    cudaGetLastError();
                }
                if (d0.bk_generic) {
                    CUDA_TRY(cudaMalloc(&d.bk_generic, h->bk_words * tsz));
                    CUDA_TRY(cudaMemcpyPeerAsync(d.bk_generic, d.id, d0.bk_generic, d0.id, h->bk_words * tsz, d.stream));
                }
                if (h->have_cggi32 || h->have_dm32) {
                    CUDA_TRY(cudaMalloc((void**)&d.bk_cggi32, h->bk_words * 4));
                    CUDA_TRY(cudaMemcpyPeerAsync(d.bk_cggi32, d.id, d0.bk_cggi32, d0.id, h->bk_words * 4, d.stream));
                }
                if (h->have_cggi64 || h->have_dm64w) {
                    CUDA_TRY(cudaMalloc((void**)&d.bk_cggi64, h->bk_words * 8));
                    CUDA_TRY(cudaMemcpyPeerAsync(d.bk_cggi64, d.id, d0.bk_cggi64, d0.id, h->bk_words * 8, d.stream));
                }
                CUDA_TRY(cudaMalloc(&d.ksk, ksk_bytes_total));
                CUDA_TRY(cudaMemcpyPeerAsync(d.ksk, d.id, d0.ksk, d0.id, ksk_bytes_total, d.stream));
            }
            for (auto& d : h->devs) {
                CUDA_TRY(cudaSetDevice(d.id));
                CUDA_TRY(cudaStreamSynchronize(d.stream));
            }
            return 0;
        };
        rc = enc();
    }
    if (rc) {
        std::string keep = g_err;
        for (auto& d : h->devs)
            free_dev(d);
        delete h;
        g_err = keep;
        return rc;
    }
    *out = h;
    return 0;
}

extern "C" int tfhe_b200_setup(const tfhe_b200_params* params, const uint64_t* bk, size_t bk_words, const uint64_t* ksk,
                               size_t ksk_words, int key_space, int first_device, int num_gpus,
                               tfhe_b200_handle** out) {
    if (!params || !bk || !ksk || !out)
        FAIL(TFHE_B200_EINVAL, "setup: null argument (the reference throws 'Need to call BTKeyGen before calling GPUSetup')");
    FlatKeys src(bk, ksk);
    return setup_impl(params, &src, bk, bk_words, ksk, ksk_words, key_space, first_device, num_gpus, out);
}

// ---- keys straight from OpenFHE's serialized streams ------------------------------------------------
static int index_streams(const void* bk_stream, size_t bk_bytes, const void* ksk_stream, size_t ksk_bytes,
                         SerializedKeys* sk) {
    if (!bk_stream || !ksk_stream || !bk_bytes || !ksk_bytes)
        FAIL(TFHE_B200_EINVAL, "serialized keys: null or empty stream");
    std::string err;
    int rc = index_serialized_acc_key(bk_stream, bk_bytes, &sk->acc, &err);
    if (rc == 0)
        rc = index_serialized_switch_key(ksk_stream, ksk_bytes, &sk->sw, &err);
    if (rc)
        g_err = "serialized keys: " + err;
    return rc;
}

extern "C" int tfhe_b200_serialized_info(const void* bk_stream, size_t bk_bytes, const void* ksk_stream, size_t ksk_bytes,
                                         tfhe_b200_serialized_info_t* info) {
    if (!info)
        FAIL(TFHE_B200_EINVAL, "serialized keys: null info");
    SerializedKeys sk;
    int rc = index_streams(bk_stream, bk_bytes, ksk_stream, ksk_bytes, &sk);
    if (rc)
        return rc;
    memset(info, 0, sizeof(*info));
    for (int k = 0; k < 3; k++)
        info->bk_dim[k] = sk.acc.dim[k];
    info->bk_rows = sk.acc.rows; info->N = sk.acc.N; info->Q = sk.acc.Q; info->psi = sk.acc.psi;
    info->ks_N = sk.sw.N; info->baseKS = sk.sw.baseKS; info->dKS = sk.sw.dKS; info->n = sk.sw.n; info->qKS = sk.sw.qKS;
    info->bk_words = sk.acc.poly_off.size() * sk.acc.N;
    info->ksk_words = sk.sw.rowA_off.size() * (sk.sw.n + 1);
    return 0;
}

extern "C" int tfhe_b200_flatten_serialized(const void* bk_stream, size_t bk_bytes, const void* ksk_stream,
                                            size_t ksk_bytes, uint64_t* bk_out, size_t bk_words, uint64_t* ksk_out,
                                            size_t ksk_words) {
    if (!bk_out || !ksk_out)
        FAIL(TFHE_B200_EINVAL, "serialized keys: null output");
    SerializedKeys sk;
    int rc = index_streams(bk_stream, bk_bytes, ksk_stream, ksk_bytes, &sk);
    if (rc)
        return rc;
    if (bk_words != sk.acc.poly_off.size() * sk.acc.N || ksk_words != sk.sw.rowA_off.size() * (sk.sw.n + 1))
        FAIL(TFHE_B200_EINVAL, "serialized keys: output sizes do not match the streams (see tfhe_b200_serialized_info)");
    sk.bk_words(bk_out, 0, bk_words);
    sk.ksk_rows(ksk_out, 0, sk.sw.rowA_off.size(), (u32)sk.sw.n + 1);
    return 0;
}

extern "C" int tfhe_b200_setup_from_serialized(const tfhe_b200_params* params, const void* bk_stream, size_t bk_bytes,
                                               const void* ksk_stream, size_t ksk_bytes, int first_device, int num_gpus,
                                               tfhe_b200_handle** out) {
    if (!params || !out)
        FAIL(TFHE_B200_EINVAL, "setup: null argument");
    SerializedKeys sk;
    int rc = index_streams(bk_stream, bk_bytes, ksk_stream, ksk_bytes, &sk);
    if (rc)
        return rc;
    // the streams must describe exactly the key set the parameters name
    const tfhe_b200_params& p = *params;
    const bool ginx = p.method == TFHE_B200_METHOD_GINX;
    const u64 want_rows = ginx ? 2 * (u64)(p.digitsG - p.numDigitsToThrow) : 2 * (u64)p.digitsG;
    const bool dims_ok = ginx ? (sk.acc.dim[0] == 1 && sk.acc.dim[1] == 2 && sk.acc.dim[2] == p.n)
                              : (sk.acc.dim[0] == p.n && sk.acc.dim[1] == p.baseR && sk.acc.dim[2] == p.digitsR);
    if (!dims_ok || sk.acc.rows != want_rows || sk.acc.N != p.N || sk.acc.Q != p.Q)
        FAIL(TFHE_B200_EINVAL, "setup: the serialized refreshing key does not match the parameter set (dimensions, ring "
                               "dimension or modulus)");
    if (sk.acc.psi && sk.acc.psi != p.psi)
        FAIL(TFHE_B200_EINVAL, "setup: the serialized refreshing key was transformed with another root of unity than "
                               "params.psi");
    if (sk.sw.N != p.N || sk.sw.baseKS != p.baseKS || sk.sw.dKS != p.dKS || sk.sw.n != p.n || sk.sw.qKS != p.qKS)
        FAIL(TFHE_B200_EINVAL, "setup: the serialized switching key does not match the parameter set");
    return setup_impl(params, &sk, nullptr, bk_words_of(&p), nullptr, ksk_words_of(&p), TFHE_B200_HOST, first_device,
                      num_gpus, out);
}

// SURVEY 8(f) rank 4 -- BinFHEContext::BTKeyGen with timeOptimization fills m_BTKey_map with one RingGSWBTKey (BK and
// KSK) per gadget base in {2^14, 2^18, 2^27} (binfhecontext.cpp:222-247, rgsw-cryptoparameters.h:105-120); the scalar
// EvalSign / EvalDecomp switch between them as the ciphertext modulus shrinks (binfhe-base-scheme.cpp:342-360,411-428).
extern "C" int tfhe_b200_add_key_set(tfhe_b200_handle* h, uint32_t baseG, const uint64_t* bk, size_t bk_words,
                                     const uint64_t* ksk, size_t ksk_words, int key_space) {
    if (!h)
        FAIL(TFHE_B200_EINVAL, "add_key_set: GPUSetup has not been called");
    if (h->borrowed_streams)
        FAIL(TFHE_B200_EINVAL, "add_key_set: not a top-level handle");
    if (h->p.method != TFHE_B200_METHOD_GINX)
        FAIL(TFHE_B200_ENOTSUP, "add_key_set: the key map exists for CGGI/GINX functional parameter sets only");
    if (baseG < 2 || (baseG & (baseG - 1)))
        FAIL(TFHE_B200_EINVAL, "add_key_set: gadget base must be a power of two");
    tfhe_b200_params p = h->p;
    p.baseG = baseG;
    // RingGSWCryptoParams::Change_BaseG (rgsw-cryptoparameters.h:276-282)
    p.digitsG = (uint32_t)std::ceil(std::log((double)p.Q) / std::log((double)baseG));
    std::lock_guard<std::mutex> lock(h->mu);
    if (baseG == h->p.baseG || h->key_map.count(baseG))
        FAIL(TFHE_B200_EINVAL, "add_key_set: a key set for this gadget base is already loaded");
    tfhe_b200_handle* c = nullptr;
    int rc = tfhe_b200_setup(&p, bk, bk_words, ksk, ksk_words, key_space, h->devs[0].id, (int)h->devs.size(), &c);
    if (rc)
        return rc;
    // the child only ever runs inside the parent's calls: same streams, so stream order covers every dependency
    for (size_t k = 0; k < c->devs.size(); k++) {
        Dev& d = c->devs[k];
        d.worker.reset();   // the child runs on the parent's workers
        cudaSetDevice(d.id);
        cudaStreamDestroy(d.stream);
        cudaStreamDestroy(d.xfer_in);
        cudaStreamDestroy(d.xfer_out);
        d.stream = h->devs[k].stream;
        d.xfer_in = h->devs[k].xfer_in;
        d.xfer_out = h->devs[k].xfer_out;
    }
    c->borrowed_streams = true;
    c->force_generic = h->force_generic;
    h->key_map[baseG] = c;
    return 0;
}
extern "C" int tfhe_b200_persistent_plan(uint32_t groups, uint32_t n, uint32_t ctas, uint32_t cta, int item,
                                         uint32_t out[3]) {
    if (groups == 0 || n == 0 || ctas == 0 || cta >= ctas)
        FAIL(TFHE_B200_EINVAL, "persistent_plan: empty launch or range index out of bounds");
    if (ctas > groups)
        FAIL(TFHE_B200_EINVAL, "persistent_plan: more CTAs than groups (a range must hold at least n steps)");
    const PersRange r = PersRange::of(groups, n, ctas, cta);
    if (item >= 0) {
        if (item >= r.n_items || !out)
            FAIL(TFHE_B200_EINVAL, "persistent_plan: item out of range");
        r.item(item, n, out[0], out[1], out[2]);
    }
    return r.n_items;
}

extern "C" int tfhe_b200_num_key_sets(const tfhe_b200_handle* h) {
    return h ? 1 + (int)h->key_map.size() : 0;
}

// ---------------------------------------------------------------------------------------------------------
// device-level building blocks (all asynchronous on d.stream)
// ---------------------------------------------------------------------------------------------------------
struct AccDesc {
    int mode = ACC_GATE;
    u64 gate_q1 = 0;
    const u64* table = nullptr;  // device
    u64 fmod = 0;
    u64* acc_io = nullptr;
    int write_acc = 0;
    u64 ext_add_b = 0;
};

// Hand-over slots of the persistent blind rotation (one per SM, allocated on first use) and a fresh epoch: the flags
// hold the epoch of the launch that filled them, so they never need clearing between launches.
static int pers_prepare(Dev& d, size_t slot_bytes) {
    if (!d.pers_flags) {
        CUDA_TRY(cudaMalloc((void**)&d.pers_flags, (size_t)(d.sm_count + 1) * 4));
        CUDA_TRY(cudaMemsetAsync(d.pers_flags, 0, (size_t)(d.sm_count + 1) * 4, d.stream));
        d.pers_tickets = 0;
    }
    if (d.pers_slot_bytes < slot_bytes) {
        if (d.pers_state) {
            CUDA_TRY(cudaStreamSynchronize(d.stream));
            CUDA_TRY(cudaFree(d.pers_state));
            d.pers_state = nullptr;
            d.pers_slot_bytes = 0;
        }
        CUDA_TRY(cudaMalloc(&d.pers_state, (size_t)d.sm_count * slot_bytes));
        d.pers_slot_bytes = slot_bytes;
    }
    if (++d.pers_epoch == 0) {   // wrapped: start over
        CUDA_TRY(cudaMemsetAsync(d.pers_flags, 0, (size_t)d.sm_count * 4, d.stream));
        d.pers_epoch = 1;
    }
    return 0;
}

static int blind_rotate_launch(tfhe_b200_handle* h, Dev& d, const BRCommon& c, int* launches, bool allow_pers = true) {
    if (c.batch <= 0)
        return 0;
    NvtxRange nvtx("tfhe_b200:blind_rotate");
    int pers_launched = 0;
    if (h->have_cggi32 && !h->force_generic) {
        CGGI32Tables t;
        t.mod = h->m32; t.bk = d.bk_cggi32; t.psi_pow = (const u32*)d.psi_pow; t.twA = h->twA_host.data(); t.twB = d.twB; t.skip_top = h->skip_top;
        t.pers_mode = (getenv("TFHE_B200_NO_PERSISTENT") || !allow_pers) ? 0 : h->pers_mode; t.pers_ctas = h->pers_ctas;
        if (t.pers_mode) {
            int rc = pers_prepare(d, PERS_SLOT_WORDS * 4);
            if (rc) return rc;
            t.pers_state = (u32*)d.pers_state; t.pers_flags = d.pers_flags; t.pers_epoch = d.pers_epoch; t.pers_slots = d.sm_count;
            t.pers_ticket = d.pers_flags + d.sm_count; t.pers_ticket_base = d.pers_tickets; t.pers_launched = &pers_launched;
        }
        CUDA_TRY(launch_br_cggi32(c, t, d.stream, d.sm_count, h->group));
        d.pers_tickets += (u32)pers_launched;
    }
    else if (h->have_dm32 && !h->force_generic) {
        CGGI32Tables t;
        t.mod = h->m32; t.bk = d.bk_cggi32; t.psi_pow = (const u32*)d.psi_pow; t.twA = h->twA_host.data(); t.twB = d.twB;
        t.skip_top = h->skip_top;
        CUDA_TRY(launch_br_dm32(c, t, d.stream, d.sm_count, h->group));
    }
    else if (h->have_dm64w && !h->force_generic) {
        CGGI64WTables t;
        t.mod = h->m64; t.bk = d.bk_cggi64; t.psi_pow = (const u64*)d.psi_pow; t.twC = d.twCw; t.twB = d.twBw; t.twU = d.twUw;
        t.plain = !h->skip_top;
        CUDA_TRY(launch_br_dm64w(c, t, d.stream, d.sm_count, h->group));
    }
    else if (h->have_cggi64w && !h->force_generic && !getenv("TFHE_B200_C64_NARROW")) {
        CGGI64WTables t;
        t.mod = h->m64; t.bk = d.bk_cggi64; t.psi_pow = (const u64*)d.psi_pow; t.twC = d.twCw; t.twB = d.twBw; t.twU = d.twUw;
        t.plain = !h->skip_top;
        t.pers_mode = (getenv("TFHE_B200_NO_PERSISTENT") || !allow_pers) ? 0 : h->pers_mode; t.pers_ctas = h->pers_ctas;
        if (t.pers_mode && !t.plain) {
            int rc = pers_prepare(d, PERS_SLOT_WORDS64 * 8);
            if (rc) return rc;
            t.pers_state = (u64*)d.pers_state; t.pers_flags = d.pers_flags; t.pers_epoch = d.pers_epoch; t.pers_slots = d.sm_count;
            t.pers_ticket = d.pers_flags + d.sm_count; t.pers_ticket_base = d.pers_tickets; t.pers_launched = &pers_launched;
        }
        CUDA_TRY(launch_br_cggi64w(c, t, d.stream, d.sm_count, h->group));
        d.pers_tickets += (u32)pers_launched;
    }
    else if (h->have_cggi64 && !h->force_generic) {
        CGGI64Tables t;
        t.mod = h->m64; t.bk = d.bk_cggi64; t.psi_pow = (const u64*)d.psi_pow; t.twB = d.twB64; t.tw32 = d.tw32_64;
        t.twA = d.twU64; t.skip_top = h->skip_top;
        CUDA_TRY(launch_br_cggi64(c, t, d.stream, h->group));
    }
    else if (!d.bk_generic)
        FAIL(TFHE_B200_EINVAL, "blind rotation: the generic key layout was released at GPUSetup");
    else if (h->is64) {
        BRTables<u64> t;
        t.mod = h->m64; t.tw_fwd = (const u64*)d.tw_fwd; t.tw_inv = (const u64*)d.tw_inv;
        t.psi_pow = (const u64*)d.psi_pow; t.bk = (const u64*)d.bk_generic;
        t.sh_fwd = (const u64*)d.sh_fwd; t.sh_inv = (const u64*)d.sh_inv;
        CUDA_TRY(launch_br_generic<u64>(c, t, d.stream, d.sm_count));
    }
    else {
        BRTables<u32> t;
        t.mod = h->m32; t.tw_fwd = (const u32*)d.tw_fwd; t.tw_inv = (const u32*)d.tw_inv;
        t.psi_pow = (const u32*)d.psi_pow; t.bk = (const u32*)d.bk_generic;
        t.sh_fwd = (const u32*)d.sh_fwd; t.sh_inv = (const u32*)d.sh_inv;
        CUDA_TRY(launch_br_generic<u32>(c, t, d.stream, d.sm_count));
    }
    if (launches)
        (*launches)++;
    return 0;
}

// Ciphertexts per CTA of the throughput shape of the active kernel (one CTA per SM: a wave is sm_count x this many).
static int throughput_group(const tfhe_b200_handle* h) {
    const u32 dk = h->d / 2;
    if (h->force_generic)
        return 1;
    if (h->group > 0)
        return h->group;
    if (h->have_cggi32)
        return h->logN == 11 ? 2 : (h->logN == 10 ? (dk <= 4 ? 4 : 2) : (dk <= 4 ? 8 : 4));
    if (h->have_dm32)
        return h->logN == 9 ? 8 : (h->logN == 11 ? 2 : 4);
    if (h->have_dm64w)
        return dk <= 3 ? 2 : 1;
    if (h->have_cggi64)
        return dk <= 3 ? 2 : 1;
    return 1;
}
// ... and the largest remainder (in ciphertexts per SM) for which the kernel has a latency shape; 0 = none.
// Does the active kernel have a persistent variant (no wave quantisation: any batch above one wave costs
// groups / SMs wave times)?  Then neither the tail launch nor wave-aligned chunks are needed.
static bool persistent_shape(const tfhe_b200_handle* h) {
    int g = 0;
    if (h->force_generic || h->group != 0 || h->pers_mode == 0 || getenv("TFHE_B200_NO_PERSISTENT"))
        return false;
    if (h->have_cggi32)
        return cggi32_pers_shape(h->logN, h->d / 2, h->skip_top, h->p.Q, &g);
    if (h->have_dm32 || h->have_dm64w)
        return false;
    if (h->have_cggi64w && !getenv("TFHE_B200_C64_NARROW"))   // the two-ciphertext shapes with top-digit elimination
        return h->skip_top && (h->d / 2 == 2 || h->d / 2 == 3);
    return false;
}
static void tail_shapes(const tfhe_b200_handle* h, int* per_cta, int* tail_per_sm) {
    *per_cta = *tail_per_sm = 0;
    if (h->force_generic || h->group != 0 || persistent_shape(h))
        return;
    const u32 dk = h->d / 2;
    if (h->have_cggi32) {
        if (h->logN == 10 && dk == 4 && !cggi32_needs_sweep(h->p.Q)) {   // CTAs of 4; CTAs of 2 (<= 2 per SM) and the latency layout (<= 1 per SM)
            *per_cta = 4;
            *tail_per_sm = 2;
        }
    }
    else if (h->have_dm32) {
        if (h->logN == 10 && dk == 4 && h->skip_top && !cggi32_needs_sweep(h->p.Q)) {   // the shape with latency layouts
            *per_cta = 4;
            *tail_per_sm = 2;
        }
    }
    else if ((h->have_dm64w || (h->have_cggi64w && !getenv("TFHE_B200_C64_NARROW"))) && dk <= 3) {
        *per_cta = 2;                      // CTAs of 2 (16 warps); one ciphertext per CTA (<= 1 per SM)
        *tail_per_sm = 1;
    }
}

static int blind_rotate(tfhe_b200_handle* h, Dev& d, int batch, const u64* ct, u64 ct_mod, const AccDesc& a, u64* ext,
                        int* launches) {
    const tfhe_b200_params& p = h->p;
    BRCommon c;
    memset(&c, 0, sizeof(c));
    c.N = p.N; c.logN = h->logN; c.n = p.n; c.d = h->d;
    c.gBits = h->gBits; c.numThrow = (p.method == TFHE_B200_METHOD_GINX) ? p.numDigitsToThrow : 0;
    c.digitsKept = h->d / 2;
    c.method = p.method; c.baseR = p.baseR; c.digitsR = p.digitsR; c.q_lwe = p.q;
    c.batch = batch; c.ct = ct; c.ct_mod = ct_mod;
    c.acc_init = a.mode; c.gate_q1 = a.gate_q1; c.Q8 = p.Q / 8 + 1;
    c.scale = a.fmod ? p.Q / a.fmod : 0;
    c.table = a.table; c.acc_io = a.acc_io; c.write_acc = a.write_acc;
    c.ext = ext; c.ext_add_b = a.ext_add_b;
    if (batch <= 0)
        return 0;
    // Wave quantisation: CTAs walk the n rotation steps in lock-step, so a launch costs ceil(CTAs / SMs) wave times and
    // a last wave that fills only part of the SMs still costs a whole one.  When that remainder is small enough for a
    // latency shape (fewer ciphertexts per CTA: a shorter step), it runs as a second launch of that shape instead --
    // STD128 at 2048 ciphertexts per GPU (16384 over 8 GPUs): 3 waves of 5.7 ms + one of 4.4 ms instead of 4 x 5.7 ms.
    int per_cta, tail_per_sm;
    tail_shapes(h, &per_cta, &tail_per_sm);
    int wave = per_cta * d.sm_count;
    int rem = wave ? batch % wave : 0;
    bool split = wave && batch > wave && rem > 0 && rem <= tail_per_sm * d.sm_count && !getenv("TFHE_B200_NO_TAIL");
    int head = batch - rem;
    bool head_pers = true;
    if (persistent_shape(h) && h->have_cggi64w && !h->have_cggi32 && h->pers_ctas == 0) {
        // The persistent variant staggers its CTAs over the rotation steps.  The 32-bit kernels do not care (their key
        // lives in L2); the 54-bit keys (342 / 513 MB) are streamed from HBM once per wave by CTAs in lock-step, and once
        // per CTA when staggered: measured 2-3 % per step.  So only the end of a launch -- the last whole wave plus the
        // remainder -- runs persistently; everything before it runs as plain lock-step waves.
        wave = throughput_group(h) * d.sm_count;
        rem = batch % wave;
        head = (batch / wave - 1) * wave;
        split = rem > 0 && head > 0;
        head_pers = false;
    }
    if (split) {
        c.batch = head;
        int r = blind_rotate_launch(h, d, c, launches, head_pers);
        if (r)
            return r;
        rem = batch - head;
        c.batch = rem;
        c.ct = ct + (size_t)head * (p.n + 1);
        if (c.table && a.mode == ACC_TABLE_PER)
            c.table = a.table + (size_t)head * ct_mod;
        if (c.acc_io)
            c.acc_io = a.acc_io + (size_t)head * 2 * p.N;
        if (c.ext)
            c.ext = ext + (size_t)head * (p.N + 1);
        return blind_rotate_launch(h, d, c, launches);
    }
    return blind_rotate_launch(h, d, c, launches);
}

static int mkmswitch_dev(tfhe_b200_handle* h, Dev& d, int batch, const u64* ext, u64 fmod, u64* out, int* launches) {
    const tfhe_b200_params& p = h->p;
    KSArgs a;
    a.N = p.N; a.n = p.n; a.baseKS = p.baseKS; a.dKS = p.dKS; a.row_stride = h->row_stride;
    a.Q = p.Q; a.qKS = p.qKS; a.fmod = fmod; a.batch = batch; a.ext = ext; a.out = out;
    a.ksk = d.ksk; a.ksk_bytes = h->ksk_bytes;
    a.partial = d.ks_partial; a.sm_count = d.sm_count;
    NvtxRange nvtx("tfhe_b200:mkmswitch");
    CUDA_TRY(launch_mkmswitch(a, d.stream));
    if (launches)
        (*launches)++;
    return 0;
}

// One bootstrap = blind rotation + extraction + MS/KS/MS.  `ext` is scratch of batch*(N+1) words.
static int bootstrap_dev(tfhe_b200_handle* h, Dev& d, int batch, const u64* ct, u64 ct_mod, const AccDesc& a, u64 fmod,
                         u64* ext, u64* out, int* launches, int ev_mid = -1) {
    int rc = blind_rotate(h, d, batch, ct, ct_mod, a, ext, launches);
    if (rc)
        return rc;
    if (ev_mid >= 0)
        CUDA_TRY(rec_ev(d, ev_mid));
    return mkmswitch_dev(h, d, batch, ext, fmod, out, launches);
}

#define AFFINE(out, x, y, sx, sy, dbl, cb, m, m2)                                                         \
    do {                                                                                                  \
        CUDA_TRY(launch_lwe_affine(out, x, y, sx, sy, dbl, cb, m, m2, batch, W, d.stream));               \
        if (launches)                                                                                     \
            (*launches)++;                                                                                \
    } while (0)

// binfhe-base-scheme.cpp:598-677
static int gate_dev(tfhe_b200_handle* h, Dev& d, int gate, int batch, const u64* c1, const u64* c2, u64 q, u64* out,
                    u64* tmp /* 4*batch*W */, u64* ext, int* launches, int* nboot) {
    const tfhe_b200_params& p = h->p;
    const u32 W = p.n + 1;
    const size_t S = (size_t)batch * W;
    static const u64 mult[6] = {5, 7, 1, 3, 5, 1};  // rgsw-cryptoparameters.h:130-137
    if (gate == TFHE_B200_XOR || gate == TFHE_B200_XNOR) {
        u64 *n1 = tmp, *n2 = tmp + S, *a1 = tmp + 2 * S, *a2 = tmp + 3 * S;
        AFFINE(n1, c1, nullptr, -1, 0, 0, q >> 2, q, 0);  // EvalNOT
        AFFINE(n2, c2, nullptr, -1, 0, 0, q >> 2, q, 0);
        // the recursive calls need their own prepared-ct scratch: reuse `out` for it
        int rc = gate_dev(h, d, TFHE_B200_AND, batch, c1, n2, q, a1, out, ext, launches, nboot);
        if (rc) return rc;
        rc = gate_dev(h, d, TFHE_B200_AND, batch, n1, c2, q, a2, out, ext, launches, nboot);
        if (rc) return rc;
        // OR(a1, a2): prepared ct goes to n1 (free now)
        AFFINE(n1, a1, a2, 1, 1, 0, 0, q, 0);
        AccDesc a;
        a.mode = ACC_GATE; a.gate_q1 = mult[TFHE_B200_OR] * (p.q >> 3); a.ext_add_b = p.Q / 8 + 1;
        rc = bootstrap_dev(h, d, batch, n1, q, a, q, ext, out, launches);
        if (rc) return rc;
        (*nboot)++;
        if (gate == TFHE_B200_XNOR) {
            AFFINE(n1, out, nullptr, -1, 0, 0, q >> 2, q, 0);
            CUDA_TRY(cudaMemcpyAsync(out, n1, S * 8, cudaMemcpyDeviceToDevice, d.stream));
        }
        return 0;
    }
    u64* prep = tmp;
    if (gate == TFHE_B200_XOR_FAST || gate == TFHE_B200_XNOR_FAST)
        AFFINE(prep, c1, c2, 1, -1, 1, 0, q, 0);  // 2*(ct1 - ct2)
    else
        AFFINE(prep, c1, c2, 1, 1, 0, 0, q, 0);   // ct1 + ct2
    AccDesc a;
    a.mode = ACC_GATE; a.gate_q1 = mult[gate] * (p.q >> 3); a.ext_add_b = p.Q / 8 + 1;
    int rc = bootstrap_dev(h, d, batch, prep, q, a, q, ext, out, launches, 2);
    (*nboot)++;
    return rc;
}

// binfhe-base-scheme.cpp:926-987.  in/out may alias.  tmp: 2*batch*W words.
static int floor_dev(tfhe_b200_handle* h, Dev& d, int batch, const u64* in, u64 mod, u32 roundbits, u64* out, u64* tmp,
                     u64* ext, u64* tab /* device, >= q words */, int* launches, int* nboot) {
    const tfhe_b200_params& p = h->p;
    const u32 W = p.n + 1;
    const size_t S = (size_t)batch * W;
    const u64 beta = p.beta;
    const u64 q = roundbits == 0 ? p.q : beta * 2 * (1ULL << roundbits);
    u64 *cq = tmp, *r = tmp + S;
    AFFINE(out, in, nullptr, 1, 0, 0, beta, mod, 0);   // ct1 = ct + beta
    AFFINE(cq, out, nullptr, 1, 0, 0, 0, mod, q);      // ct1Modq
    // f1(x) = x < q/2 ? Q - q/4 : q/4  (binfhe-base-scheme.cpp:934-939); tables are generated on the device in stream
    // order, so the body never waits for the GPU
    CUDA_TRY(launch_step_table(tab, STEP_HALF, q, mod - q / 4, q / 4, d.stream));
    (*launches)++;
    AccDesc a;
    a.mode = ACC_TABLE; a.table = tab; a.fmod = mod;
    int rc = bootstrap_dev(h, d, batch, cq, q, a, mod, ext, r, launches);
    if (rc) return rc;
    (*nboot)++;
    AFFINE(out, out, r, 1, -1, 0, 0, mod, 0);          // ct1 -= ct2
    AFFINE(cq, out, nullptr, 1, 0, 0, 0, mod, q);      // ct2Modq
    // f2 (binfhe-base-scheme.cpp:941-952)
    CUDA_TRY(launch_step_table(tab, STEP_FLOOR2, q, mod, 0, d.stream));
    (*launches)++;
    rc = bootstrap_dev(h, d, batch, cq, q, a, mod, ext, r, launches);
    if (rc) return rc;
    (*nboot)++;
    AFFINE(out, out, r, 1, -1, 0, 0, mod, 0);          // ct1 -= ct3
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// sharded execution driver
// ---------------------------------------------------------------------------------------------------------
struct CallCtx {
    tfhe_b200_handle* h;
    tfhe_b200_stats* stats;
    int launches = 0;
    int nboot = 0;
};

// Runs `body` once per GPU on that GPU's shard.  A single-GPU handle runs it inline; a multi-GPU handle hands every
// shard to that GPU's worker thread, so all GPUs are fed, drained and synchronised concurrently (the reference walks its
// GPUs from one host thread: bootstrapping.cu:1616-1667, :1642-1646).
template <typename F>
static int run_sharded(tfhe_b200_handle* h, int batch, tfhe_b200_stats* stats, F body) {
    std::lock_guard<std::mutex> lock(h->mu);
    const int nd = (int)h->devs.size();
    std::vector<int> launches(nd, 0), nboot(nd, 0);
    auto per_dev = [&](int k) -> int {
        Dev& d = h->devs[k];
        int start, count;
        shard_of(batch, nd, k, &start, &count);
        d.ev_mask = 0;
        d.pending.clear();
        d.pin.off = 0;
        auto run = [&]() -> int {
            NvtxRange nvtx("tfhe_b200:shard");
            CUDA_TRY(cudaSetDevice(d.id));
            CUDA_TRY(rec_ev(d, 0));
            int r = count > 0 ? body(d, start, count, &launches[k], &nboot[k]) : 0;
            if (r)
                return r;
            CUDA_TRY(rec_ev(d, 5));
            return 0;
        };
        int rc = run();
        cudaError_t e = cudaStreamSynchronize(d.stream);   // also on failure: nothing of this call may stay in flight
        if (e != cudaSuccess && rc == 0) {
            g_err = std::string("stream sync failed: ") + cudaGetErrorString(e);
            rc = TFHE_B200_ECUDA;
        }
        if (rc == 0)
            for (const PendingOut& po : d.pending)
                par_memcpy(po.dst, po.src, po.bytes);
        d.pending.clear();
        return rc;
    };
    int rc = 0;
    if (nd == 1)
        rc = per_dev(0);
    else {
        for (int k = 0; k < nd; k++)
            h->devs[k].worker->submit([&per_dev, k] { return per_dev(k); });
        for (int k = 0; k < nd; k++) {
            std::string err;
            int r = h->devs[k].worker->wait(&err);
            if (r && rc == 0) {
                rc = r;
                g_err = err;
            }
        }
    }
    if (rc == 0 && stats) {
        memset(stats, 0, sizeof(*stats));
        // phases from the first GPU (fields whose markers this call did not record stay 0), total = slowest GPU
        Dev& d0 = h->devs[0];
        auto span = [&](Dev& d, int a, int b, float* out) {
            float t = 0;
            if ((d.ev_mask >> a & 1) && (d.ev_mask >> b & 1) && cudaEventElapsedTime(&t, d.ev[a], d.ev[b]) == cudaSuccess)
                *out = t;
        };
        for (Dev& d : h->devs) {
            cudaSetDevice(d.id);
            float t = 0;
            span(d, 0, 5, &t);
            stats->total_ms = std::max(stats->total_ms, t);
        }
        cudaSetDevice(d0.id);
        span(d0, 0, 1, &stats->h2d_ms);
        span(d0, 1, 2, &stats->blind_rotate_ms);
        span(d0, 2, 3, &stats->keyswitch_ms);
        span(d0, 4, 5, &stats->d2h_ms);
        cudaGetLastError();
        stats->bootstraps = (uint32_t)nboot[0];
        for (int k = 0; k < nd; k++)
            stats->kernel_launches += (uint32_t)launches[k];
    }
    return rc;
}

static int check_call(tfhe_b200_handle* h, int batch, const void* a, const void* b, const char* what) {
    if (!h)
        FAIL(TFHE_B200_EINVAL, std::string(what) + ": GPUSetup has not been called");
    if (batch <= 0)
        FAIL(TFHE_B200_EINVAL, std::string("ERROR: ") + what + ": input vector is empty");
    if (!a || !b)
        FAIL(TFHE_B200_EINVAL, std::string(what) + ": null pointer");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// operator-level entry points
// ---------------------------------------------------------------------------------------------------------
extern "C" int tfhe_b200_keygen(const tfhe_b200_params* params, const int8_t* sk_lwe, const int8_t* sk_ring,
                                const uint8_t* key32, int device, uint64_t* bk_dev, uint64_t* ksk_dev) {
    if (!params || !sk_lwe || !sk_ring || !bk_dev || !ksk_dev)
        FAIL(TFHE_B200_EINVAL, "KeyGen: null argument");
    const tfhe_b200_params& p = *params;
    if (p.n == 0 || p.N == 0 || (p.N & (p.N - 1)) || p.Q < 3 || p.Q >= (1ULL << 62) || p.qKS < 2 || p.baseKS < 2 ||
        p.baseG < 2 || p.digitsG == 0 || p.dKS == 0)
        FAIL(TFHE_B200_EINVAL, "KeyGen: inconsistent parameters");
    if (p.method != TFHE_B200_METHOD_GINX && p.method != TFHE_B200_METHOD_AP)
        FAIL(TFHE_B200_ENOTSUP, "KeyGen: unknown bootstrapping method");
    for (uint32_t i = 0; i < p.n; i++)
        if (sk_lwe[i] < -1 || sk_lwe[i] > 1)
            FAIL(TFHE_B200_EINVAL, "ERROR: only ternary secret key distributions are supported.");
    for (uint32_t i = 0; i < p.N; i++)
        if (sk_ring[i] < -1 || sk_ring[i] > 1)
            FAIL(TFHE_B200_EINVAL, "ERROR: only ternary secret key distributions are supported.");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        FAIL(TFHE_B200_ENODEV, "KeyGen: no CUDA device (this engine has no CPU fallback)");
    if (device < 0 || device >= ndev)
        FAIL(TFHE_B200_EINVAL, "KeyGen: bad device index");
    unsigned char key[32];
    if (key32)
        memcpy(key, key32, 32);
    else {   // 256 bits from the operating system's CSPRNG
        size_t got = 0;
        while (got < sizeof(key)) {
            ssize_t r = getrandom(key + got, sizeof(key) - got, 0);
            if (r < 0) {
                if (errno == EINTR)
                    continue;
                FAIL(TFHE_B200_EINVAL, "KeyGen: getrandom() failed; pass 32 bytes of key material explicitly");
            }
            got += (size_t)r;
        }
    }
    int rc = keygen_device(p, (const signed char*)sk_lwe, (const signed char*)sk_ring, key, device, bk_dev, ksk_dev);
    volatile unsigned char* wipe = key;
    for (size_t i = 0; i < sizeof(key); i++)
        wipe[i] = 0;
    if (rc)
        g_err = keygen_last_error();
    return rc;
}

// TEST ONLY: deterministic key generation from a 64-bit seed (reproducible fixtures, noise-statistics tests).  A 64-bit
// seed can be searched exhaustively, so keys generated this way offer at most 64 bits of security whatever the
// parameter set claims -- production callers use tfhe_b200_keygen with 32 bytes from a CSPRNG (or NULL).
extern "C" int tfhe_b200_keygen_test_seed(const tfhe_b200_params* params, const int8_t* sk_lwe, const int8_t* sk_ring,
                                          uint64_t seed, int device, uint64_t* bk_dev, uint64_t* ksk_dev) {
    uint8_t key[32] = {};
    memcpy(key, &seed, 8);
    memcpy(key + 8, "tfhe_b200 TEST-ONLY seed", 24);
    return tfhe_b200_keygen(params, sk_lwe, sk_ring, key, device, bk_dev, ksk_dev);
}

extern "C" int tfhe_b200_eval_acc(tfhe_b200_handle* h, int batch, const uint64_t* a, uint64_t ct_mod, uint64_t* acc,
                                  int space, tfhe_b200_stats* stats) {
    int rc = check_call(h, batch, a, acc, "EvalAcc");
    if (rc) return rc;
    const tfhe_b200_params& p = h->p;
    if (ct_mod == 0 || (2ULL * p.N) % ct_mod)
        FAIL(TFHE_B200_EINVAL, "EvalAcc: ciphertext modulus must divide 2N");
    Dev& d0 = h->devs[0];
    return run_sharded(h, batch, stats, [&](Dev& d, int start, int count, int* launches, int* nboot) -> int {
        const u32 n = p.n, W = n + 1, N = p.N;
        size_t need = (size_t)count * (W + 2 * N) * 8 + 4096;
        int r = arena_reserve(d, need);
        if (r) return r;
        TAKE(ct, u64, d, (size_t)count * W);
        TAKE(ac, u64, d, (size_t)count * 2 * N);
        if (space == TFHE_B200_HOST)
            pin_reserve(d, 2 * (size_t)count * 2 * N * 8 + 4096);
        // the operator passes only the mask; b is unused because the accumulator is explicit
        CUDA_TRY(cudaMemsetAsync(ct, 0, (size_t)count * W * 8, d.stream));
        if (space == TFHE_B200_HOST)
            CUDA_TRY(cudaMemcpy2DAsync(ct, W * 8, a + (size_t)start * n, n * 8, n * 8, count, cudaMemcpyHostToDevice,
                                       d.stream));
        else if (d.id == d0.id)
            CUDA_TRY(cudaMemcpy2DAsync(ct, W * 8, a + (size_t)start * n, n * 8, n * 8, count,
                                       cudaMemcpyDeviceToDevice, d.stream));
        else {
            TAKE(tmpa, u64, d, (size_t)count * n);
            CUDA_TRY(cudaMemcpyPeerAsync(tmpa, d.id, a + (size_t)start * n, d0.id, (size_t)count * n * 8, d.stream));
            CUDA_TRY(cudaMemcpy2DAsync(ct, W * 8, tmpa, n * 8, n * 8, count, cudaMemcpyDeviceToDevice, d.stream));
        }
        r = copy_in(d, d0, ac, acc + (size_t)start * 2 * N, (size_t)count * 2 * N * 8, space);
        if (r) return r;
        CUDA_TRY(rec_ev(d, 1));
        AccDesc ad;
        ad.mode = ACC_EXPLICIT; ad.acc_io = ac; ad.write_acc = 1;
        r = blind_rotate(h, d, count, ct, ct_mod, ad, nullptr, launches);
        if (r) return r;
        (*nboot) = 1;
        {
            CUDA_TRY(rec_ev(d, 2));
            CUDA_TRY(rec_ev(d, 3));
            CUDA_TRY(rec_ev(d, 4));
        }
        return copy_out(d, d0, acc + (size_t)start * 2 * N, ac, (size_t)count * 2 * N * 8, space);
    });
}

extern "C" int tfhe_b200_mkmswitch(tfhe_b200_handle* h, int batch, const uint64_t* in, uint64_t fmod, uint64_t* out,
                                   int space, tfhe_b200_stats* stats) {
    int rc = check_call(h, batch, in, out, "MKMSwitch");
    if (rc) return rc;
    if (fmod == 0)
        FAIL(TFHE_B200_EINVAL, "MKMSwitch: fmod must be non-zero");
    const tfhe_b200_params& p = h->p;
    Dev& d0 = h->devs[0];
    return run_sharded(h, batch, stats, [&](Dev& d, int start, int count, int* launches, int* nboot) -> int {
        const u32 W = p.n + 1, N = p.N;
        int r = arena_reserve(d, (size_t)count * (W + N + 1) * 8 + 4096);
        if (r) return r;
        TAKE(ext, u64, d, (size_t)count * (N + 1));
        TAKE(o, u64, d, (size_t)count * W);
        if (space == TFHE_B200_HOST)
            pin_reserve(d, (size_t)count * (N + 1 + W) * 8 + 4096);
        r = copy_in(d, d0, ext, in + (size_t)start * (N + 1), (size_t)count * (N + 1) * 8, space);
        if (r) return r;
        {
            CUDA_TRY(rec_ev(d, 1));
            CUDA_TRY(rec_ev(d, 2));
        }
        r = mkmswitch_dev(h, d, count, ext, fmod, o, launches);
        if (r) return r;
        (void)nboot;
        {
            CUDA_TRY(rec_ev(d, 3));
            CUDA_TRY(rec_ev(d, 4));
        }
        return copy_out(d, d0, out + (size_t)start * W, o, (size_t)count * W * 8, space);
    });
}

extern "C" int tfhe_b200_mul_matrix(tfhe_b200_handle* h, int in, int out_cols, const uint64_t* ct, const int64_t* matrix,
                                    uint64_t modulus, uint64_t* out, int space, tfhe_b200_stats* stats) {
    if (!h)
        FAIL(TFHE_B200_EINVAL, "CiphertextMulMatrix: GPUSetup has not been called");
    if (in <= 0 || !ct)
        FAIL(TFHE_B200_EINVAL, "Input ciphertexts are empty.");
    if (out_cols <= 0 || !matrix)
        FAIL(TFHE_B200_EINVAL, "Input matrix is empty.");
    if (!out || modulus == 0 || modulus >= (1ULL << 63))
        FAIL(TFHE_B200_EINVAL, "CiphertextMulMatrix: bad output pointer or modulus");
    const tfhe_b200_params& p = h->p;
    // like the reference (lwe-operation.cu:91) this runs on the first GPU only: the operand is tiny
    std::lock_guard<std::mutex> lock(h->mu);
    Dev& d = h->devs[0];
    const u32 W = p.n + 1;
    auto run = [&]() -> int {
        CUDA_TRY(cudaSetDevice(d.id));
        const size_t scratch_bytes = mul_matrix_scratch_bytes(in, out_cols, W, modulus);
        int r = arena_reserve(d, ((size_t)in * W + (size_t)in * out_cols + (size_t)out_cols * W) * 8 + scratch_bytes + 4096);
        if (r) return r;
        TAKE(dct, u64, d, (size_t)in * W);
        TAKE(dM, i64, d, (size_t)in * out_cols);
        TAKE(dout, u64, d, (size_t)out_cols * W);
        void* scratch = scratch_bytes ? (void*)arena_take<u32>(d, scratch_bytes / 4) : nullptr;
        if (scratch_bytes && !scratch)
            return TFHE_B200_ENOMEM;
        d.ev_mask = 0;
        d.pending.clear();
        d.pin.off = 0;
        if (space == TFHE_B200_HOST)
            pin_reserve(d, ((size_t)in * W + (size_t)in * out_cols + (size_t)out_cols * W) * 8 + 4096);
        CUDA_TRY(rec_ev(d, 0));
        r = copy_in(d, d, dct, ct, (size_t)in * W * 8, space);
        if (r) return r;
        r = copy_in(d, d, dM, matrix, (size_t)in * out_cols * 8, space);
        if (r) return r;
        CUDA_TRY(launch_mul_matrix(dout, dct, dM, in, out_cols, W, modulus, scratch, d.stream));
        r = copy_out(d, d, out, dout, (size_t)out_cols * W * 8, space);
        if (r) return r;
        CUDA_TRY(rec_ev(d, 5));
        CUDA_TRY(cudaStreamSynchronize(d.stream));
        for (const PendingOut& po : d.pending)
            par_memcpy(po.dst, po.src, po.bytes);
        d.pending.clear();
        if (stats) {
            memset(stats, 0, sizeof(*stats));
            cudaEventElapsedTime(&stats->total_ms, d.ev[0], d.ev[5]);
            stats->kernel_launches = scratch ? 2 : 1;
        }
        return 0;
    };
    return run();
}

// ---------------------------------------------------------------------------------------------------------
// fused batched API
// ---------------------------------------------------------------------------------------------------------
// Host-side view of the operands of one EvalBinGate call: dense [batch][n+1] arrays (flat != nullptr), or one pointer
// per ciphertext for callers that hold every ciphertext as a separate object (std::vector<LWECiphertext>).
namespace {
struct GateIO {
    const u64 *ct1 = nullptr, *ct2 = nullptr;        // dense
    u64* out = nullptr;
    const u64 *const *a1 = nullptr, *const *a2 = nullptr;   // scattered: a1[i] -> n mask words, b1[i]
    const u64 *b1 = nullptr, *b2 = nullptr;
    u64* const* a_out = nullptr;
    u64* b_out = nullptr;
    bool scattered() const { return a1 != nullptr; }
};
// gather ciphertexts [first, first+cnt) of a scattered operand into a dense staging block
void gather_cts(u64* dst, const u64* const* a, const u64* b, size_t first, size_t cnt, u32 n) {
#pragma omp parallel for num_threads(4) schedule(static) if (cnt >= 256)
    for (long i = 0; i < (long)cnt; i++) {
        u64* row = dst + (size_t)i * (n + 1);
        memcpy(row, a[first + i], (size_t)n * 8);
        row[n] = b[first + i];
    }
}
void scatter_cts(u64* const* a, u64* b, const u64* src, size_t first, size_t cnt, u32 n) {
#pragma omp parallel for num_threads(4) schedule(static) if (cnt >= 256)
    for (long i = 0; i < (long)cnt; i++) {
        const u64* row = src + (size_t)i * (n + 1);
        memcpy(a[first + i], row, (size_t)n * 8);
        b[first + i] = row[n];
    }
}
}  // namespace

static int eval_bin_gate_impl(tfhe_b200_handle* h, int gate, int batch, const GateIO& io, uint64_t ct_mod, int space,
                              tfhe_b200_stats* stats) {
    const tfhe_b200_params& p = h->p;
    if (gate < 0 || gate > TFHE_B200_XNOR)
        FAIL(TFHE_B200_EINVAL, "EvalBinGate: unknown gate");
    if (ct_mod == 0 || (2ULL * p.N) % ct_mod)
        FAIL(TFHE_B200_EINVAL, "EvalBinGate: ciphertext modulus must divide 2N");
    Dev& d0 = h->devs[0];
    const bool sc = io.scattered();
    return run_sharded(h, batch, stats, [&](Dev& d, int start, int count, int* launches, int* nboot) -> int {
        const u32 W = p.n + 1, N = p.N, n = p.n;
        const size_t S = (size_t)count * W;
        int r = arena_reserve(d, (7 * S + (size_t)count * (N + 1)) * 8 + 8192);
        if (r) return r;
        TAKE(c1, u64, d, S);
        TAKE(c2, u64, d, S);
        TAKE(o, u64, d, S);
        TAKE(tmp, u64, d, 4 * S);
        TAKE(ext, u64, d, (size_t)count * (N + 1));
        const int unit = d.sm_count * throughput_group(h);   // one wave of the throughput shape
        const bool pin_in = !sc && space == TFHE_B200_HOST && host_is_pinned(io.ct1) && host_is_pinned(io.ct2);
        const bool pin_out = !sc && space == TFHE_B200_HOST && host_is_pinned(io.out);
        if (space == TFHE_B200_HOST)
            pin_reserve(d, (pin_in ? 0 : 2 * (S * 8 + 256)) + (pin_out ? 0 : S * 8 + 256) + 4096);
        unsigned char *sin1 = nullptr, *sin2 = nullptr, *sout = nullptr;
        if (space == TFHE_B200_HOST) {
            if (!pin_in) {
                sin1 = pin_take(d, S * 8);
                sin2 = pin_take(d, S * 8);
            }
            if (!pin_out)
                sout = pin_take(d, S * 8);
            if (sc && !(sin1 && sin2 && sout))
                FAIL(TFHE_B200_ENOMEM, "EvalBinGate: pinned staging for the gathered ciphertexts could not be allocated");
        }
        // stage rows [off, off+cnt) of the shard's inputs (host memcpy / gather) and return the DMA sources
        auto stage_in = [&](int off, int cnt, const u64** s1, const u64** s2) {
            const size_t o8 = (size_t)off * W, b8 = (size_t)cnt * W * 8;
            if (sc) {
                gather_cts(reinterpret_cast<u64*>(sin1) + o8, io.a1, io.b1, (size_t)start + off, cnt, n);
                gather_cts(reinterpret_cast<u64*>(sin2) + o8, io.a2, io.b2, (size_t)start + off, cnt, n);
            }
            else if (sin1 && sin2) {
                par_memcpy(sin1 + o8 * 8, io.ct1 + (size_t)(start + off) * W, b8);
                par_memcpy(sin2 + o8 * 8, io.ct2 + (size_t)(start + off) * W, b8);
            }
            else {
                *s1 = io.ct1 + (size_t)(start + off) * W;
                *s2 = io.ct2 + (size_t)(start + off) * W;
                return;
            }
            *s1 = reinterpret_cast<const u64*>(sin1) + o8;
            *s2 = reinterpret_cast<const u64*>(sin2) + o8;
        };
        auto hand_out = [&](int off, int cnt) {   // staged results of rows [off, off+cnt) to the caller
            const size_t o8 = (size_t)off * W;
            if (sc)
                scatter_cts(io.a_out, io.b_out, reinterpret_cast<const u64*>(sout) + o8, (size_t)start + off, cnt, n);
            else
                par_memcpy(io.out + (size_t)(start + off) * W, sout + o8 * 8, (size_t)cnt * W * 8);
        };
        if (space == TFHE_B200_HOST && count >= 3 * unit && !getenv("TFHE_B200_NO_PIPELINE")) {
            // Host buffers: the shard goes through in chunks so that the upload of chunk k+1 and the download of chunk
            // k-1 ride under the bootstraps of chunk k (three streams, hand-over by events).  The reference copies
            // everything in, computes, copies everything out (bootstrapping.cu:1562-1853).  Chunk boundaries lie on
            // whole waves of CTAs (sm_count x ciphertexts per CTA), or the partial last wave of every chunk would cost more
            // than the overlap gains; the first and the last chunk are ONE wave, so that only one wave's worth of input
            // is exposed before the first launch and one wave's worth of output after the last.  Pageable buffers (and
            // per-object ciphertexts) are staged through the handle's pinned memory: the host memcpy / gather of chunk
            // k+1 and the hand-back of chunk k-1 run while the GPU works on chunk k.
            // (with a persistent blind rotation the remainder costs only its share of a wave: it joins the last chunk)
            const int waves = persistent_shape(h) ? count / unit : (count + unit - 1) / unit;
            int wsz[MAX_CHUNKS], nch = 0;
            {
                const int mid = waves - 2, nmid = std::min(MAX_CHUNKS - 2, (mid + 5) / 6);   // middle chunks of <= ~6 waves
                wsz[nch++] = 1;
                for (int k = 0; k < nmid; k++)
                    wsz[nch++] = mid / nmid + (k < mid % nmid ? 1 : 0);
                wsz[nch++] = 1;
            }
            int coff[MAX_CHUNKS + 1];
            coff[0] = 0;
            for (int k = 0; k < nch; k++)
                coff[k + 1] = std::min(count, coff[k] + wsz[k] * unit);
            coff[nch] = count;
            if (persistent_shape(h) && nch > 2) {
                // A persistent blind rotation costs groups / SMs wave times whatever the chunk size, and its staggered
                // CTAs run a step 1.7 % faster than lock-step waves (all SMs asking L2 for the same key lines at the same
                // moment), so the middle chunks share everything between the first and the last wave EVENLY instead of
                // in whole waves: 16384 gates = 1 + 5 x 5.13 + 1 waves.
                const int g = throughput_group(h), nmid = nch - 2;
                const int each = (count - 2 * unit) / nmid / g * g;
                for (int k = 1; k < nch - 1; k++)
                    coff[k + 1] = coff[k] + each;   // the last chunk (one wave + the rounding) ends at count
            }
            cudaEvent_t* ev_in = d.pev;
            cudaEvent_t* ev_done = d.pev + MAX_CHUNKS;
            cudaEvent_t* ev_out = d.pev + 2 * MAX_CHUNKS;
            CUDA_TRY(cudaStreamWaitEvent(d.xfer_in, d.ev[0], 0));
            CUDA_TRY(rec_ev(d, 1));
            auto upload = [&](int k) -> int {
                const int off = coff[k], cnt = coff[k + 1] - off;
                const size_t o8 = (size_t)off * W, b8 = (size_t)cnt * W * 8;
                const u64 *s1, *s2;
                stage_in(off, cnt, &s1, &s2);
                CUDA_TRY(cudaMemcpyAsync(c1 + o8, s1, b8, cudaMemcpyHostToDevice, d.xfer_in));
                CUDA_TRY(cudaMemcpyAsync(c2 + o8, s2, b8, cudaMemcpyHostToDevice, d.xfer_in));
                CUDA_TRY(cudaEventRecord(ev_in[k], d.xfer_in));
                return 0;
            };
            r = upload(0);
            if (r) return r;
            int handed = 0;   // chunks already handed back to the caller (staged outputs)
            for (int k = 0; k < nch; k++) {
                const int off = coff[k], cnt = coff[k + 1] - off;
                const size_t o8 = (size_t)off * W;
                int nb = 0;
                CUDA_TRY(cudaStreamWaitEvent(d.stream, ev_in[k], 0));
                r = gate_dev(h, d, gate, cnt, c1 + o8, c2 + o8, ct_mod, o + o8, tmp, ext, launches, &nb);
                if (r) return r;
                *nboot = nb;
                CUDA_TRY(cudaEventRecord(ev_done[k], d.stream));
                CUDA_TRY(cudaStreamWaitEvent(d.xfer_out, ev_done[k], 0));
                u64* dst = sout ? reinterpret_cast<u64*>(sout) + o8 : io.out + (size_t)(start + off) * W;
                CUDA_TRY(cudaMemcpyAsync(dst, o + o8, (size_t)cnt * W * 8, cudaMemcpyDeviceToHost, d.xfer_out));
                CUDA_TRY(cudaEventRecord(ev_out[k], d.xfer_out));
                if (k + 1 < nch) {   // staged while the GPU works on chunk k
                    r = upload(k + 1);
                    if (r) return r;
                }
                // results that have already landed go back to the caller while the GPU keeps working
                while (sout && handed < k && cudaEventQuery(ev_out[handed]) == cudaSuccess) {
                    hand_out(coff[handed], coff[handed + 1] - coff[handed]);
                    handed++;
                }
                cudaGetLastError();   // a cudaErrorNotReady from the query is not an error
            }
            {
                if (gate == TFHE_B200_XOR || gate == TFHE_B200_XNOR)
                    CUDA_TRY(rec_ev(d, 2));
                CUDA_TRY(rec_ev(d, 3));
                CUDA_TRY(rec_ev(d, 4));
            }
            CUDA_TRY(cudaEventRecord(d.pev[3 * MAX_CHUNKS], d.xfer_out));
            CUDA_TRY(cudaStreamWaitEvent(d.stream, d.pev[3 * MAX_CHUNKS], 0));   // the final synchronisation covers the downloads
            if (sout)
                for (; handed < nch; handed++) {
                    CUDA_TRY(cudaEventSynchronize(ev_out[handed]));
                    hand_out(coff[handed], coff[handed + 1] - coff[handed]);
                }
            return 0;
        }
        if (space == TFHE_B200_HOST) {
            const u64 *s1, *s2;
            stage_in(0, count, &s1, &s2);
            CUDA_TRY(cudaMemcpyAsync(c1, s1, S * 8, cudaMemcpyHostToDevice, d.stream));
            CUDA_TRY(cudaMemcpyAsync(c2, s2, S * 8, cudaMemcpyHostToDevice, d.stream));
        }
        else {
            r = copy_in(d, d0, c1, io.ct1 + (size_t)start * W, S * 8, space);
            if (r) return r;
            r = copy_in(d, d0, c2, io.ct2 + (size_t)start * W, S * 8, space);
            if (r) return r;
        }
        CUDA_TRY(rec_ev(d, 1));
        r = gate_dev(h, d, gate, count, c1, c2, ct_mod, o, tmp, ext, launches, nboot);
        if (r) return r;
        {
            if (gate == TFHE_B200_XOR || gate == TFHE_B200_XNOR)
                CUDA_TRY(rec_ev(d, 2));
            CUDA_TRY(rec_ev(d, 3));
            CUDA_TRY(rec_ev(d, 4));
        }
        if (space == TFHE_B200_HOST && sout) {
            CUDA_TRY(cudaMemcpyAsync(sout, o, S * 8, cudaMemcpyDeviceToHost, d.stream));
            CUDA_TRY(cudaStreamSynchronize(d.stream));
            hand_out(0, count);
            return 0;
        }
        return copy_out(d, d0, io.out + (size_t)start * W, o, S * 8, space);
    });
}

extern "C" int tfhe_b200_eval_bin_gate(tfhe_b200_handle* h, int gate, int batch, const uint64_t* ct1,
                                       const uint64_t* ct2, uint64_t ct_mod, uint64_t* out, int space,
                                       tfhe_b200_stats* stats) {
    int rc = check_call(h, batch, ct1, ct2, "EvalBinGate");
    if (rc) return rc;
    if (!out)
        FAIL(TFHE_B200_EINVAL, "EvalBinGate: null output");
    if (ct1 == ct2)
        FAIL(TFHE_B200_EINVAL, "Input ciphertexts should be independant");
    GateIO io;
    io.ct1 = ct1; io.ct2 = ct2; io.out = out;
    return eval_bin_gate_impl(h, gate, batch, io, ct_mod, space, stats);
}

extern "C" int tfhe_b200_eval_bin_gate_v(tfhe_b200_handle* h, int gate, int batch, const uint64_t* const* a1,
                                         const uint64_t* b1, const uint64_t* const* a2, const uint64_t* b2,
                                         uint64_t ct_mod, uint64_t* const* a_out, uint64_t* b_out,
                                         tfhe_b200_stats* stats) {
    int rc = check_call(h, batch, a1, a2, "EvalBinGate");
    if (rc) return rc;
    if (!b1 || !b2 || !a_out || !b_out)
        FAIL(TFHE_B200_EINVAL, "EvalBinGate: null pointer");
    if (a1 == a2)
        FAIL(TFHE_B200_EINVAL, "Input ciphertexts should be independant");
    GateIO io;
    io.a1 = a1; io.b1 = b1; io.a2 = a2; io.b2 = b2; io.a_out = a_out; io.b_out = b_out;
    return eval_bin_gate_impl(h, gate, batch, io, ct_mod, TFHE_B200_HOST, stats);
}

// ---------------------------------------------------------------------------------------------------------
// Gate-graph submission.  The netlist is lowered on the host to primitive nodes
// (XOR / XNOR -> NOT, NOT, AND, AND, OR [, NOT], exactly the expansion gate_dev performs, binfhe-base-scheme.cpp:617-640),
// levelised, and every (level, gate kind) group is bootstrapped by one launch over group x shard ciphertexts.
// ---------------------------------------------------------------------------------------------------------
namespace {
struct PrimNode {
    int gate;      // TFHE_B200_OR .. TFHE_B200_XNOR_FAST, or TFHE_B200_NOT
    int in0, in1;  // wire ids (primitive numbering)
    int level;     // bootstrap depth
};
}  // namespace

extern "C" int tfhe_b200_eval_circuit(tfhe_b200_handle* h, int batch, int n_inputs, const uint64_t* inputs,
                                      uint64_t ct_mod, int n_nodes, const tfhe_b200_gate_node* nodes, int n_outputs,
                                      const int32_t* output_wires, uint64_t* out, int space, tfhe_b200_stats* stats) {
    if (!h)
        FAIL(TFHE_B200_EINVAL, "EvalCircuit: GPUSetup has not been called");
    if (batch <= 0 || n_inputs <= 0 || !inputs)
        FAIL(TFHE_B200_EINVAL, "ERROR: EvalCircuit: input vector is empty");
    if (n_nodes < 0 || (n_nodes > 0 && !nodes) || n_outputs <= 0 || !output_wires || !out)
        FAIL(TFHE_B200_EINVAL, "EvalCircuit: bad netlist or output description");
    const tfhe_b200_params& p = h->p;
    if (ct_mod == 0 || (2ULL * p.N) % ct_mod)
        FAIL(TFHE_B200_EINVAL, "EvalCircuit: ciphertext modulus must divide 2N");
    // ---- lower to primitive nodes; wire_of[user wire] = primitive wire --------------------------------------------
    std::vector<PrimNode> prim;
    std::vector<int> wire_of(n_inputs + n_nodes), level_of_wire;
    level_of_wire.assign(n_inputs, 0);
    for (int i = 0; i < n_inputs; i++)
        wire_of[i] = i;
    auto add = [&](int gate, int a, int b) -> int {
        PrimNode nd;
        nd.gate = gate; nd.in0 = a; nd.in1 = b;
        nd.level = gate == TFHE_B200_NOT ? level_of_wire[a] : 1 + std::max(level_of_wire[a], level_of_wire[b]);
        prim.push_back(nd);
        level_of_wire.push_back(nd.level);
        return n_inputs + (int)prim.size() - 1;
    };
    for (int g = 0; g < n_nodes; g++) {
        const tfhe_b200_gate_node& nd = nodes[g];
        const int lim = n_inputs + g;
        if (nd.in0 < 0 || nd.in0 >= lim)
            FAIL(TFHE_B200_EINVAL, "EvalCircuit: node input is not an earlier wire");
        const int a = wire_of[nd.in0];
        if (nd.gate == TFHE_B200_NOT) {
            wire_of[lim] = add(TFHE_B200_NOT, a, a);
            continue;
        }
        if (nd.gate < 0 || nd.gate > TFHE_B200_XNOR)
            FAIL(TFHE_B200_EINVAL, "EvalCircuit: unknown gate");
        if (nd.in1 < 0 || nd.in1 >= lim)
            FAIL(TFHE_B200_EINVAL, "EvalCircuit: node input is not an earlier wire");
        if (nd.in0 == nd.in1)
            FAIL(TFHE_B200_EINVAL, "Input ciphertexts should be independant");
        const int b = wire_of[nd.in1];
        if (nd.gate == TFHE_B200_XOR || nd.gate == TFHE_B200_XNOR) {
            const int n1 = add(TFHE_B200_NOT, a, a), n2 = add(TFHE_B200_NOT, b, b);
            const int a1 = add(TFHE_B200_AND, a, n2), a2 = add(TFHE_B200_AND, n1, b);
            int o = add(TFHE_B200_OR, a1, a2);
            if (nd.gate == TFHE_B200_XNOR)
                o = add(TFHE_B200_NOT, o, o);
            wire_of[lim] = o;
        }
        else
            wire_of[lim] = add(nd.gate, a, b);
    }
    for (int o = 0; o < n_outputs; o++)
        if (output_wires[o] < 0 || output_wires[o] >= n_inputs + n_nodes)
            FAIL(TFHE_B200_EINVAL, "EvalCircuit: output wire out of range");
    const int n_prim = (int)prim.size();
    int max_level = 0, n_boot_nodes = 0;
    for (const PrimNode& nd : prim) {
        max_level = std::max(max_level, nd.level);
        n_boot_nodes += nd.gate != TFHE_B200_NOT;
    }
    // execution order: per level, the bootstrapped nodes grouped by gate kind, then that level's NOT nodes (index order)
    struct Group { int gate; std::vector<int> ids; };
    std::vector<std::vector<Group>> plan(max_level + 1);
    std::vector<std::vector<int>> nots(max_level + 1);
    size_t max_group = 1;
    for (int id = 0; id < n_prim; id++) {
        const PrimNode& nd = prim[id];
        if (nd.gate == TFHE_B200_NOT) {
            nots[nd.level].push_back(id);
            continue;
        }
        auto& lv = plan[nd.level];
        Group* grp = nullptr;
        for (Group& g : lv)
            if (g.gate == nd.gate)
                grp = &g;
        if (!grp) {
            lv.push_back(Group{nd.gate, {}});
            grp = &lv.back();
        }
        grp->ids.push_back(id);
    }
    Dev& d0 = h->devs[0];
    const int nd_gpus = (int)h->devs.size();
    const int shard_max = (batch + nd_gpus - 1) / nd_gpus;
    // at most ~32768 ciphertexts per launch keeps the scratch bounded
    const size_t group_cap = std::max<size_t>(1, 32768 / (size_t)std::max(1, shard_max));
    for (auto& lv : plan)
        for (Group& g : lv)
            max_group = std::max(max_group, std::min(g.ids.size(), group_cap));
    static const u64 mult[6] = {5, 7, 1, 3, 5, 1};  // rgsw-cryptoparameters.h:130-137
    return run_sharded(h, batch, stats, [&](Dev& d, int start, int count, int* launches, int* nboot) -> int {
        const u32 W = p.n + 1, N = p.N;
        const size_t S = (size_t)count * W;
        const size_t n_wires = (size_t)n_inputs + n_prim;
        int r = arena_reserve(d, ((n_wires + max_group) * S + max_group * (size_t)count * (N + 1)) * 8 + 16384);
        if (r) return r;
        TAKE(wires, u64, d, n_wires * S);
        TAKE(prep, u64, d, max_group * S);
        TAKE(ext, u64, d, max_group * (size_t)count * (N + 1));
        if (space == TFHE_B200_HOST)
            pin_reserve(d, ((size_t)n_inputs + n_outputs) * (S * 8 + 256) + 4096);
        for (int i = 0; i < n_inputs; i++) {
            r = copy_in(d, d0, wires + (size_t)i * S, inputs + ((size_t)i * batch + start) * W, S * 8, space);
            if (r) return r;
        }
        CUDA_TRY(rec_ev(d, 1));
        const int batch_ = count;   // AFFINE uses `batch`
        int depth = 0;
        for (int lvl = 0; lvl <= max_level; lvl++) {
            for (const Group& g : plan[lvl]) {
                for (size_t off = 0; off < g.ids.size(); off += group_cap) {
                    const size_t k = std::min(group_cap, g.ids.size() - off);
                    // prepared ciphertexts of the chunk's k nodes side by side -> one blind rotation + one key switch
                    for (size_t x = 0; x < k; x++) {
                        const PrimNode& nd = prim[g.ids[off + x]];
                        const u64* c1 = wires + (size_t)nd.in0 * S;
                        const u64* c2 = wires + (size_t)nd.in1 * S;
                        const int batch = batch_;
                        if (g.gate == TFHE_B200_XOR_FAST || g.gate == TFHE_B200_XNOR_FAST)
                            AFFINE(prep + x * S, c1, c2, 1, -1, 1, 0, ct_mod, 0);
                        else
                            AFFINE(prep + x * S, c1, c2, 1, 1, 0, 0, ct_mod, 0);
                    }
                    AccDesc a;
                    a.mode = ACC_GATE; a.gate_q1 = mult[g.gate] * (p.q >> 3); a.ext_add_b = p.Q / 8 + 1;
                    // the key switch writes [k*count][W]; reuse `prep` as the landing block once the rotation has
                    // consumed it (blind_rotate reads prep, mkmswitch reads ext and writes prep)
                    r = bootstrap_dev(h, d, (int)(k * count), prep, ct_mod, a, ct_mod, ext, prep, launches);
                    if (r) return r;
                    for (size_t x = 0; x < k; x++)
                        CUDA_TRY(cudaMemcpyAsync(wires + ((size_t)n_inputs + g.ids[off + x]) * S, prep + x * S, S * 8,
                                                 cudaMemcpyDeviceToDevice, d.stream));
                }
                depth = lvl;
            }
            for (int id : nots[lvl]) {
                const PrimNode& nd = prim[id];
                const int batch = batch_;
                AFFINE(wires + ((size_t)n_inputs + id) * S, wires + (size_t)nd.in0 * S, nullptr, -1, 0, 0, ct_mod >> 2,
                       ct_mod, 0);   // EvalNOT (binfhe-base-scheme.cpp:741-745)
            }
        }
        (void)depth;
        *nboot = n_boot_nodes;   // bootstraps per batch element
        {
            CUDA_TRY(rec_ev(d, 2));
            CUDA_TRY(rec_ev(d, 3));
            CUDA_TRY(rec_ev(d, 4));
        }
        for (int o = 0; o < n_outputs; o++) {
            r = copy_out(d, d0, out + ((size_t)o * batch + start) * W, wires + (size_t)wire_of[output_wires[o]] * S,
                         S * 8, space);
            if (r) return r;
        }
        return 0;
    });
}

extern "C" int tfhe_b200_bootstrap_func(tfhe_b200_handle* h, int batch, const uint64_t* ct, uint64_t ct_mod,
                                        const uint64_t* table, int per_ct, uint64_t fmod, uint64_t* out, int space,
                                        tfhe_b200_stats* stats) {
    int rc = check_call(h, batch, ct, table, "BootstrapFunc");
    if (rc) return rc;
    const tfhe_b200_params& p = h->p;
    if (!out || fmod == 0 || ct_mod == 0 || (2ULL * p.N) % ct_mod)
        FAIL(TFHE_B200_EINVAL, "BootstrapFunc: bad modulus or null output");
    Dev& d0 = h->devs[0];
    return run_sharded(h, batch, stats, [&](Dev& d, int start, int count, int* launches, int* nboot) -> int {
        const u32 W = p.n + 1, N = p.N;
        const size_t S = (size_t)count * W;
        size_t tabw = per_ct ? (size_t)count * ct_mod : ct_mod;
        int r = arena_reserve(d, (2 * S + (size_t)count * (N + 1) + tabw) * 8 + 8192);
        if (r) return r;
        TAKE(c, u64, d, S);
        TAKE(o, u64, d, S);
        TAKE(ext, u64, d, (size_t)count * (N + 1));
        TAKE(tab, u64, d, tabw);
        if (space == TFHE_B200_HOST)
            pin_reserve(d, (2 * S + tabw) * 8 + 4096);
        r = copy_in(d, d0, c, ct + (size_t)start * W, S * 8, space);
        if (r) return r;
        r = copy_in(d, d0, tab, table + (per_ct ? (size_t)start * ct_mod : 0), tabw * 8, space);
        if (r) return r;
        CUDA_TRY(rec_ev(d, 1));
        AccDesc a;
        a.mode = per_ct ? ACC_TABLE_PER : ACC_TABLE; a.table = tab; a.fmod = fmod;
        r = bootstrap_dev(h, d, count, c, ct_mod, a, fmod, ext, o, launches, 2);
        if (r) return r;
        (*nboot) = 1;
        {
            CUDA_TRY(rec_ev(d, 3));
            CUDA_TRY(rec_ev(d, 4));
        }
        return copy_out(d, d0, out + (size_t)start * W, o, S * 8, space);
    });
}

// binfhe-base-scheme.cpp:162-186
static int check_input_function(const std::vector<u64>& lut, u64 mod) {
    size_t len = lut.size();
    int ret = 0;
    if (lut[0] == mod - lut[len / 2]) {
        for (size_t i = 1; i < len / 2; i++)
            if (lut[i] != mod - lut[len / 2 + i]) { ret = 2; break; }
    }
    else if (lut[0] == lut[len / 2]) {
        ret = 1;
        for (size_t i = 1; i < len / 2; i++)
            if (lut[i] != lut[len / 2 + i]) { ret = 2; break; }
    }
    else
        ret = 2;
    return ret;
}

extern "C" int tfhe_b200_eval_func(tfhe_b200_handle* h, int batch, const uint64_t* ct, uint64_t ct_mod,
                                   const uint64_t* lut, size_t lut_len, int per_ct, uint64_t* out, int space,
                                   tfhe_b200_stats* stats) {
    int rc = check_call(h, batch, ct, lut, "EvalFunc");
    if (rc) return rc;
    const tfhe_b200_params& p = h->p;
    const u64 q = ct_mod;
    if (!out || q == 0 || lut_len != q || (2ULL * p.N) % q)
        FAIL(TFHE_B200_EINVAL, "EvalFunc: LUT length must equal the ciphertext modulus, which must divide 2N");
    // classification uses the first LUT only, as the reference does (binfhe-base-scheme.cpp:816)
    std::vector<u64> lut0(lut_len);
    if (space == TFHE_B200_HOST)
        memcpy(lut0.data(), lut, lut_len * 8);
    else {
        CUDA_TRY(cudaSetDevice(h->devs[0].id));
        CUDA_TRY(cudaMemcpy(lut0.data(), lut, lut_len * 8, cudaMemcpyDeviceToHost));
    }
    const int prop = check_input_function(lut0, q);
    if (prop == 2 && q > p.N)
        FAIL(TFHE_B200_ENOTSUP,
             "ERROR: ciphertext modulus q needs to be <= ring dimension for arbitrary function evaluation");
    const u64 beta = p.beta;
    Dev& d0 = h->devs[0];
    return run_sharded(h, batch, stats, [&](Dev& d, int start, int count, int* launches, int* nboot) -> int {
        const int batch = count;  // for AFFINE
        const u32 W = p.n + 1, N = p.N;
        const size_t S = (size_t)count * W;
        const u64 dq = q << 1;
        const size_t ntab = per_ct ? (size_t)count : 1;
        int r = arena_reserve(d, (4 * S + (size_t)count * (N + 1) + ntab * (q + dq) + 2 * dq) * 8 + 16384);
        if (r) return r;
        TAKE(c0, u64, d, S);
        TAKE(c1, u64, d, S);
        TAKE(c2, u64, d, S);
        TAKE(o, u64, d, S);
        TAKE(ext, u64, d, (size_t)count * (N + 1));
        TAKE(dl, u64, d, ntab * q);    // raw LUT(s)
        TAKE(dt, u64, d, ntab * dq);   // expanded table(s)
        TAKE(df0, u64, d, dq);
        if (space == TFHE_B200_HOST)
            pin_reserve(d, (2 * S + ntab * q) * 8 + 4096);
        r = copy_in(d, d0, c0, ct + (size_t)start * W, S * 8, space);
        if (r) return r;
        r = copy_in(d, d0, dl, lut + (per_ct ? (size_t)start * q : 0), ntab * q * 8, space);
        if (r) return r;
        CUDA_TRY(rec_ev(d, 1));
        AccDesc a;
        if (prop == 0) {
            AFFINE(c1, c0, nullptr, 1, 0, 0, beta, q, 0);
            a.mode = per_ct ? ACC_TABLE_PER : ACC_TABLE; a.table = dl; a.fmod = q;
            r = bootstrap_dev(h, d, count, c1, q, a, q, ext, o, launches);
            if (r) return r;
            (*nboot) = 1;
        }
        else if (prop == 2) {
            CUDA_TRY(launch_step_table(df0, STEP_HALF, dq, dq - dq / 4, dq / 4, d.stream));   // f0, :716-720
            (*launches)++;
            AFFINE(c2, c0, nullptr, 1, 0, 0, beta, dq, 0);          // ct2 = ct1 + beta (mod 2q)
            a.mode = ACC_TABLE; a.table = df0; a.fmod = dq;
            r = bootstrap_dev(h, d, count, c2, dq, a, dq, ext, c1, launches);   // ct3
            if (r) return r;
            AFFINE(c1, c0, c1, 1, -1, 0, beta, dq, 0);              // ct3 = ct1 - ct3 + beta
            AFFINE(c1, c1, nullptr, 1, 0, 0, dq - (q >> 1), dq, 0); // ct3 -= q/2
            CUDA_TRY(launch_lut_expand(dt, dl, q, dq, 2, (int)ntab, d.stream));
            (*launches)++;
            a.mode = per_ct ? ACC_TABLE_PER : ACC_TABLE; a.table = dt; a.fmod = dq;
            r = bootstrap_dev(h, d, count, c1, dq, a, dq, ext, c2, launches);   // ct4
            if (r) return r;
            AFFINE(o, c2, nullptr, 1, 0, 0, 0, dq, q);              // SetModulus(q)
            (*nboot) = 2;
        }
        else {
            CUDA_TRY(launch_step_table(df0, STEP_HALF, q, q - q / 4, q / 4, d.stream));
            (*launches)++;
            AFFINE(c1, c0, nullptr, 1, 0, 0, beta, q, 0);
            a.mode = ACC_TABLE; a.table = df0; a.fmod = q;
            r = bootstrap_dev(h, d, count, c1, q, a, q, ext, c2, launches);     // ct2
            if (r) return r;
            AFFINE(c2, c0, c2, 1, -1, 0, beta, q, 0);               // ct2 = ct - ct2 + beta
            AFFINE(c2, c2, nullptr, 1, 0, 0, q - (q >> 2), q, 0);   // ct2 -= q/4
            CUDA_TRY(launch_lut_expand(dt, dl, q, q, 1, (int)ntab, d.stream));
            (*launches)++;
            a.mode = per_ct ? ACC_TABLE_PER : ACC_TABLE; a.table = dt; a.fmod = q;
            r = bootstrap_dev(h, d, count, c2, q, a, q, ext, o, launches);
            if (r) return r;
            (*nboot) = 2;
        }
        {
            CUDA_TRY(rec_ev(d, 2));
            CUDA_TRY(rec_ev(d, 3));
            CUDA_TRY(rec_ev(d, 4));
        }
        return copy_out(d, d0, out + (size_t)start * W, o, S * 8, space);
    });
}

extern "C" int tfhe_b200_eval_floor(tfhe_b200_handle* h, int batch, const uint64_t* ct, uint64_t ct_mod,
                                    uint32_t roundbits, uint64_t* out, int space, tfhe_b200_stats* stats) {
    int rc = check_call(h, batch, ct, out, "EvalFloor");
    if (rc) return rc;
    const tfhe_b200_params& p = h->p;
    const u64 qf = roundbits == 0 ? p.q : p.beta * 2 * (1ULL << roundbits);
    if (ct_mod == 0 || (2ULL * p.N) % qf)
        FAIL(TFHE_B200_EINVAL, "EvalFloor: bad modulus");
    Dev& d0 = h->devs[0];
    return run_sharded(h, batch, stats, [&](Dev& d, int start, int count, int* launches, int* nboot) -> int {
        const u32 W = p.n + 1, N = p.N;
        const size_t S = (size_t)count * W;
        int r = arena_reserve(d, (4 * S + (size_t)count * (N + 1) + qf) * 8 + 8192);
        if (r) return r;
        TAKE(c, u64, d, S);
        TAKE(o, u64, d, S);
        TAKE(tmp, u64, d, 2 * S);
        TAKE(ext, u64, d, (size_t)count * (N + 1));
        TAKE(tab, u64, d, qf);
        if (space == TFHE_B200_HOST)
            pin_reserve(d, 2 * S * 8 + 4096);
        r = copy_in(d, d0, c, ct + (size_t)start * W, S * 8, space);
        if (r) return r;
        CUDA_TRY(rec_ev(d, 1));
        r = floor_dev(h, d, count, c, ct_mod, roundbits, o, tmp, ext, tab, launches, nboot);
        if (r) return r;
        {
            CUDA_TRY(rec_ev(d, 2));
            CUDA_TRY(rec_ev(d, 3));
            CUDA_TRY(rec_ev(d, 4));
        }
        return copy_out(d, d0, out + (size_t)start * W, o, S * 8, space);
    });
}

// shared body of EvalSign (binfhe-base-scheme.cpp:989-1037) and EvalDecomp (:1039-1085)
// gadget base the reference selects for a ciphertext modulus (0 = keep the current one)
static u32 dynamic_base(u64 mod) {
    const u32 binLog = (u32)std::ceil(std::log2((double)mod));
    if (binLog <= 17)
        return 1u << 27;
    if (binLog <= 26)
        return 1u << 18;
    return 0;
}

static int sign_decomp(tfhe_b200_handle* h, int batch, const uint64_t* ct, uint64_t ct_mod, bool decomp, int max_digits,
                       uint64_t* out, uint64_t* out_mods, int space, tfhe_b200_stats* stats, int* ndigits) {
    const tfhe_b200_params& p = h->p;
    const u64 q = p.q, beta = p.beta;
    if (decomp && ct_mod <= q)
        FAIL(TFHE_B200_ENOTSUP,
             "ERROR: EvalDecomp is only for large precision. For small precision, please use bootstrapping directly");
    // modulus chain
    std::vector<u64> mods;
    {
        u64 m = ct_mod;
        while (m > q) {
            mods.push_back(m);
            m = m / q * 2 * beta;
            if (mods.size() > 64)
                FAIL(TFHE_B200_EINVAL, "EvalSign/EvalDecomp: modulus chain does not terminate");
        }
        mods.push_back(m);  // final modulus (<= q)
    }
    const int nd = (int)mods.size();
    if (decomp && nd > max_digits)
        FAIL(TFHE_B200_EINVAL, "EvalDecomp: max_digits too small");
    const u64 mfin = mods.back();
    if (!decomp && (mfin == 0 || (2ULL * p.N) % mfin))
        FAIL(TFHE_B200_EINVAL, "EvalSign: final modulus must divide 2N");
    if (ndigits)
        *ndigits = nd;
    if (decomp && out_mods) {
        for (int k = 0; k < nd; k++)
            out_mods[k] = k < nd - 1 ? q : mfin;
    }
    // "if (EKs.size() == 3)": the reference switches gadget base only when the full three-key map is present; every base
    // the modulus chain asks for must then be loaded ("ERROR: No key [..] found in the map")
    const bool dynamic = h->key_map.size() == 2;
    if (dynamic) {
        for (size_t k = 1; k < mods.size(); k++) {
            const u32 base = dynamic_base(mods[k]);
            if (base && base != p.baseG && !h->key_map.count(base))
                FAIL(TFHE_B200_EINVAL, "ERROR: No key [" + std::to_string(base) + "] found in the map");
        }
    }
    Dev& d0 = h->devs[0];
    return run_sharded(h, batch, stats, [&](Dev& d, int start, int count, int* launches, int* nboot) -> int {
        const int batch = count;
        const u32 W = p.n + 1, N = p.N;
        const size_t S = (size_t)count * W;
        const size_t outw = decomp ? S * max_digits : S;
        int r = arena_reserve(d, (5 * S + outw + (size_t)count * (N + 1) + 2 * q) * 8 + 16384);
        if (r) return r;
        TAKE(cur, u64, d, S);
        TAKE(nxt, u64, d, S);
        TAKE(tmp, u64, d, 2 * S);
        TAKE(o, u64, d, outw);
        TAKE(ext, u64, d, (size_t)count * (N + 1));
        TAKE(tab, u64, d, 2 * q);
        if (space == TFHE_B200_HOST)
            pin_reserve(d, (S + outw) * 8 + 4096);
        r = copy_in(d, d0, cur, ct + (size_t)start * W, S * 8, space);
        if (r) return r;
        CUDA_TRY(rec_ev(d, 1));
        u64 mod = ct_mod;
        int digit = 0;
        // current key set (curEK of binfhe-base-scheme.cpp:327-333); the child handle's Dev of the same index carries
        // that set's key tables and borrows this Dev's streams
        const size_t dev_index = (size_t)(&d - &h->devs[0]);
        tfhe_b200_handle* kh = h;
        while (mod > q) {
            if (decomp) {
                CUDA_TRY(launch_copy_mod(o + (size_t)digit * W, (size_t)max_digits * W, cur, W, q, count, W, d.stream));
                (*launches)++;
                digit++;
            }
            r = floor_dev(kh, kh->devs[dev_index], count, cur, mod, 0, nxt, tmp, ext, tab, launches, nboot);
            if (r) return r;
            u64 newmod = mod / q * 2 * beta;
            CUDA_TRY(launch_mod_switch(cur, nxt, mod, newmod, S, d.stream));   // lwe-pke.cpp:204-215
            (*launches)++;
            mod = newmod;
            if (dynamic) {  // binfhe-base-scheme.cpp:342-360 / :411-428
                const u32 base = dynamic_base(mod);
                if (base)
                    kh = base == h->p.baseG ? h : h->key_map.find(base)->second;
            }
        }
        if (decomp) {
            CUDA_TRY(launch_copy_mod(o + (size_t)digit * W, (size_t)max_digits * W, cur, W, 0, count, W, d.stream));
            (*launches)++;
        }
        else {
            AFFINE(cur, cur, nullptr, 1, 0, 0, beta, mod, 0);
            CUDA_TRY(launch_step_table(tab, STEP_HALF, mod, q / 4, q - q / 4, d.stream));   // f3, :1004-1010
            (*launches)++;
            AccDesc a;
            a.mode = ACC_TABLE; a.table = tab; a.fmod = q;
            r = bootstrap_dev(kh, kh->devs[dev_index], count, cur, mod, a, q, ext, nxt, launches);
            if (r) return r;
            (*nboot)++;
            AFFINE(o, nxt, nullptr, 1, 0, 0, q - (q >> 2), q, 0);
        }
        {
            CUDA_TRY(rec_ev(d, 2));
            CUDA_TRY(rec_ev(d, 3));
            CUDA_TRY(rec_ev(d, 4));
        }
        const size_t ow = decomp ? (size_t)max_digits * W : W;
        return copy_out(d, d0, out + (size_t)start * ow, o, (size_t)count * ow * 8, space);
    });
}

extern "C" int tfhe_b200_eval_sign(tfhe_b200_handle* h, int batch, const uint64_t* ct, uint64_t ct_mod, uint64_t* out,
                                   int space, tfhe_b200_stats* stats) {
    int rc = check_call(h, batch, ct, out, "EvalSign");
    if (rc) return rc;
    return sign_decomp(h, batch, ct, ct_mod, false, 1, out, nullptr, space, stats, nullptr);
}

extern "C" int tfhe_b200_eval_decomp(tfhe_b200_handle* h, int batch, const uint64_t* ct, uint64_t ct_mod,
                                     int max_digits, uint64_t* out, uint64_t* out_mods, int space,
                                     tfhe_b200_stats* stats) {
    int rc = check_call(h, batch, ct, out, "EvalDecomp");
    if (rc) return rc;
    if (max_digits <= 0 || !out_mods)
        FAIL(TFHE_B200_EINVAL, "EvalDecomp: bad max_digits / out_mods");
    int nd = 0;
    rc = sign_decomp(h, batch, ct, ct_mod, true, max_digits, out, out_mods, space, stats, &nd);
    return rc ? rc : nd;
}
