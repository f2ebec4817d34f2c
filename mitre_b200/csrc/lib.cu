// lib.cu -- library plumbing: version / errors / device info, row packing kernels, and the
// host-pointer compatibility entry with the reference's exact argument list.
#include <immintrin.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace mitre {

static int g_sm_count = 0;
static size_t g_smem_optin = 0;

static void query_device()
{
    if (g_sm_count) return;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return;
    g_sm_count = prop.multiProcessorCount;
    g_smem_optin = prop.sharedMemPerBlockOptin;
}

int sm_count()
{
    query_device();
    return g_sm_count > 0 ? g_sm_count : 148;
}

size_t smem_optin()
{
    query_device();
    return g_smem_optin ? g_smem_optin : 227 * 1024;
}

// ---- packing -----------------------------------------------------------------------------------
// One warp per (row, 32-subject half word): lanes read 32 consecutive subjects (coalesced),
// ballot, bit-reverse so that subject i lands at bit 63-(i&63).
template <typename T>
__global__ void __launch_bounds__(256)
pack_rows_kernel(const T *__restrict__ truth, int64_t P, int N, uint64_t *__restrict__ packed)
{
    const int W64 = words_for(N);
    const int halves = 2 * W64;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    uint32_t *packed32 = reinterpret_cast<uint32_t *>(packed);
    for (int64_t item = warp0; item < P * halves; item += nwarps) {
        const int64_t row = item / halves;
        const int h = (int)(item % halves);
        const int i = h * 32 + lane;
        const bool t = (i < N) && (truth[row * N + i] != (T)0);
        const uint32_t bits = __brev(__ballot_sync(0xffffffffu, t));
        // little-endian: the high 32 bits of word w are packed32[2w+1]
        if (lane == 0) packed32[(row * W64 + (h >> 1)) * 2 + ((h & 1) ? 0 : 1)] = bits;
    }
}

__global__ void __launch_bounds__(256)
unpack_rows_kernel(const uint64_t *__restrict__ packed, int64_t P, int N, uint8_t *__restrict__ truth)
{
    const int W64 = words_for(N);
    const int64_t total = P * N;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t row = e / N;
        const int i = (int)(e % N);
        truth[e] = (packed[row * W64 + (i >> 6)] & subject_bit(i)) ? 1 : 0;
    }
}

// fp64 pipe probe: 8 independent DFMA chains per thread (the denominator for the scoring kernel's
// fp64 roofline; MEASURED_PEAKS.json carries no fp64 figure).
__global__ void __launch_bounds__(256)
fp64_probe_kernel(int iters, double seed, double *__restrict__ sink)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 123.456) sink[0] = r;      // never true: keeps the chains alive
}

static unsigned grid_for(int64_t work_items, int threads)
{
    int64_t blocks = ceil_div64(work_items, threads);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace mitre

using namespace mitre;

extern "C" const char *mitre_version(void) { return "mitre_b200 0.1.0 (sm_100a)"; }

extern "C" const char *mitre_error_string(int code)
{
    switch (code) {
        case MITRE_OK: return "ok";
        case MITRE_ERR_INVALID: return "invalid argument";
        case MITRE_ERR_UNSUPPORTED: return "unsupported shape";
        case MITRE_ERR_WORKSPACE: return "workspace too small";
        case MITRE_ERR_NOT_PD: return "matrix is not positive definite";
        case MITRE_ERR_NO_DEVICE: return "no usable CUDA device";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown error";
}

extern "C" int mitre_device_info(int *sm, int *cc_major, int *cc_minor, size_t *smem)
{
    int dev = 0;
    MITRE_CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    MITRE_CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (sm) *sm = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (smem) *smem = prop.sharedMemPerBlockOptin;
    return MITRE_OK;
}

extern "C" int mitre_pack_rows_f64(void *stream, const double *truth, int64_t P, int N, uint64_t *packed)
{
    if (P < 0 || N <= 0 || (P > 0 && (!truth || !packed))) return MITRE_ERR_INVALID;
    if (P == 0) return MITRE_OK;
    pack_rows_kernel<double><<<grid_for(P * 2 * words_for(N) * 32, 256), 256, 0, as_stream(stream)>>>(
        truth, P, N, packed);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

extern "C" int mitre_pack_rows_u8(void *stream, const uint8_t *truth, int64_t P, int N, uint64_t *packed)
{
    if (P < 0 || N <= 0 || (P > 0 && (!truth || !packed))) return MITRE_ERR_INVALID;
    if (P == 0) return MITRE_OK;
    pack_rows_kernel<uint8_t><<<grid_for(P * 2 * words_for(N) * 32, 256), 256, 0, as_stream(stream)>>>(
        truth, P, N, packed);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

extern "C" int mitre_unpack_rows_u8(void *stream, const uint64_t *packed, int64_t P, int N, uint8_t *truth)
{
    if (P < 0 || N <= 0 || (P > 0 && (!truth || !packed))) return MITRE_ERR_INVALID;
    if (P == 0) return MITRE_OK;
    unpack_rows_kernel<<<grid_for(P * N, 256), 256, 0, as_stream(stream)>>>(packed, P, N, truth);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

// ---- host-pointer compatibility entry -------------------------------------------------------------
// Same arguments as efficient_likelihoods.likelihoods_of_all_options (pyx:15-21): the caller hands over
// the float64[P, N] truth matrix the reference builds on every sweep (logit_rules.py:214-217), 8 N bytes
// per candidate, in pageable host memory.  Sending that over PCIe (15 GB/s from pageable memory) made
// this entry no faster than the reference itself; the rows are therefore packed to bits ON THE HOST
// (SSE2 compares, a few threads over row blocks, bound by reading the matrix from host DRAM) and only
// 8 W64 bytes per candidate cross the bus.  Two pinned staging sets alternate: packing of chunk c + 1
// overlaps the copies and the kernel of chunk c.  Staging resources live for the life of the process.
namespace {

struct HostStage {
    std::mutex lock;
    int device = -1;
    int64_t chunk_rows = 0;
    int W64 = 0;
    size_t ws_bytes = 0, small_elems = 0;
    cudaStream_t st[2] = {nullptr, nullptr};
    cudaEvent_t inputs_ready = nullptr;
    uint64_t *h_packed[2] = {nullptr, nullptr}, *d_packed[2] = {nullptr, nullptr};
    double *h_out[2] = {nullptr, nullptr}, *d_out[2] = {nullptr, nullptr};
    void *d_ws[2] = {nullptr, nullptr};
    double *h_small = nullptr, *d_small = nullptr;    // X, omega, kappa
    int32_t *d_status = nullptr, *h_status = nullptr;

    void release()
    {
        for (int b = 0; b < 2; ++b) {
            if (st[b]) cudaStreamDestroy(st[b]);
            cudaFree(d_packed[b]); cudaFree(d_out[b]); cudaFree(d_ws[b]);
            if (h_packed[b]) cudaFreeHost(h_packed[b]);
            if (h_out[b]) cudaFreeHost(h_out[b]);
            st[b] = nullptr; d_packed[b] = h_packed[b] = nullptr; d_out[b] = h_out[b] = nullptr; d_ws[b] = nullptr;
        }
        if (inputs_ready) cudaEventDestroy(inputs_ready);
        inputs_ready = nullptr;
        cudaFree(d_small); cudaFree(d_status);
        if (h_small) cudaFreeHost(h_small);
        if (h_status) cudaFreeHost(h_status);
        d_small = h_small = nullptr; d_status = h_status = nullptr;
        chunk_rows = 0; W64 = 0; ws_bytes = 0; small_elems = 0; device = -1;
    }

    cudaError_t ensure(int64_t rows, int w64, size_t ws, size_t small)
    {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev == device && rows <= chunk_rows && w64 <= W64 && ws <= ws_bytes && small <= small_elems)
            return cudaSuccess;
        release();
        device = dev;
        for (int b = 0; b < 2; ++b) {
            if ((e = cudaStreamCreateWithFlags(&st[b], cudaStreamNonBlocking)) != cudaSuccess) return e;
            if ((e = cudaMalloc(&d_packed[b], sizeof(uint64_t) * rows * w64)) != cudaSuccess) return e;
            if ((e = cudaMalloc(&d_out[b], sizeof(double) * rows)) != cudaSuccess) return e;
            if ((e = cudaMalloc(&d_ws[b], ws)) != cudaSuccess) return e;
            if ((e = cudaMallocHost(&h_packed[b], sizeof(uint64_t) * rows * w64)) != cudaSuccess) return e;
            if ((e = cudaMallocHost(&h_out[b], sizeof(double) * rows)) != cudaSuccess) return e;
        }
        if ((e = cudaEventCreateWithFlags(&inputs_ready, cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&d_small, sizeof(double) * small)) != cudaSuccess) return e;
        if ((e = cudaMallocHost(&h_small, sizeof(double) * small)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&d_status, sizeof(int32_t))) != cudaSuccess) return e;
        if ((e = cudaMallocHost(&h_status, sizeof(int32_t))) != cudaSuccess) return e;
        chunk_rows = rows; W64 = w64; ws_bytes = ws; small_elems = small;
        return cudaSuccess;
    }
};

HostStage g_host_stage;
int64_t g_host_chunk_bytes = 0;   // test hook: packed bytes per staging chunk (0 = default)

unsigned host_threads()
{
    unsigned n = std::thread::hardware_concurrency();
    if (n > 32) n = 32;
    return n < 1 ? 1 : n;
}

template <class F>
void parallel_rows(int64_t rows, int64_t min_rows_per_thread, F f)
{
    unsigned n = host_threads();
    if ((int64_t)n * min_rows_per_thread > rows) n = (unsigned)(rows / min_rows_per_thread);
    if (n < 2) {
        f((int64_t)0, rows);
        return;
    }
    const int64_t piece = (rows + n - 1) / n;
    std::vector<std::thread> workers;
    for (unsigned t = 0; t < n; ++t) {
        const int64_t lo = (int64_t)t * piece, hi = (lo + piece < rows) ? lo + piece : rows;
        if (lo >= hi) break;
        workers.emplace_back([=] { f(lo, hi); });
    }
    for (auto &w : workers) w.join();
}

// float64[rows, N] (non-zero = True) -> packed uint64[rows, W64]; subject i at bit 63 - i % 64 of word i / 64
inline void pack_row_sse2(const double *x, int N, int W64, uint64_t *out)
{
    const __m128d zero = _mm_setzero_pd();
    for (int w = 0; w < W64; ++w) {
        const int i0 = w * 64, n = (N - i0 < 64) ? (N - i0) : 64;
        uint64_t word = 0;
        int i = 0;
        for (; i + 2 <= n; i += 2) {
            const unsigned m = (unsigned)_mm_movemask_pd(_mm_cmpneq_pd(_mm_loadu_pd(x + i0 + i), zero));
            // bit 0 of m = element i, bit 1 = element i + 1; element i goes to bit 63 - i
            word |= ((uint64_t)(m & 1u) << (63 - i)) | ((uint64_t)(m >> 1) << (62 - i));
        }
        if (i < n && x[i0 + i] != 0.0) word |= 1ull << (63 - i);
        out[w] = word;
    }
}

// AVX2: eight doubles per step -> one byte of the word, already in subject order after a bit reversal
// (compiled for the target inside this function only; taken when the CPU reports AVX2)
__attribute__((target("avx2"))) void pack_rows_avx2(const double *truth, int64_t lo, int64_t hi, int N, int W64,
                                                     uint64_t *packed)
{
    static const unsigned char rev4[16] = {0, 8, 4, 12, 2, 10, 6, 14, 1, 9, 5, 13, 3, 11, 7, 15};
    const __m256d zero = _mm256_setzero_pd();
    for (int64_t r = lo; r < hi; ++r) {
        const double *x = truth + r * N;
        uint64_t *out = packed + r * W64;
        for (int w = 0; w < W64; ++w) {
            const int i0 = w * 64, n = (N - i0 < 64) ? (N - i0) : 64;
            uint64_t word = 0;
            int i = 0;
            for (; i + 4 <= n; i += 4) {
                const unsigned m = (unsigned)_mm256_movemask_pd(_mm256_cmp_pd(_mm256_loadu_pd(x + i0 + i), zero, _CMP_NEQ_UQ));
                // bit j of m = element i + j, which goes to bit 63 - (i + j): reverse the nibble, place its top at 63 - i
                word |= (uint64_t)rev4[m] << (60 - i);
            }
            for (; i < n; ++i)
                if (x[i0 + i] != 0.0) word |= 1ull << (63 - i);
            out[w] = word;
        }
    }
}

void pack_rows_host(const double *truth, int64_t rows, int N, int W64, uint64_t *packed)
{
    static const bool avx2 = __builtin_cpu_supports("avx2");
    parallel_rows(rows, 4096, [=](int64_t lo, int64_t hi) {
        if (avx2) {
            pack_rows_avx2(truth, lo, hi, N, W64, packed);
            return;
        }
        for (int64_t r = lo; r < hi; ++r) pack_row_sse2(truth + r * N, N, W64, packed + r * W64);
    });
}

void copy_rows_host(double *dst, const double *src, int64_t count)
{
    parallel_rows(count, 1 << 18, [=](int64_t lo, int64_t hi) { memcpy(dst + lo, src + lo, sizeof(double) * (hi - lo)); });
}

}  // namespace

extern "C" int mitre_set_host_chunk_bytes(int64_t bytes)
{
    std::lock_guard<std::mutex> guard(g_host_stage.lock);
    g_host_chunk_bytes = bytes > 0 ? bytes : 0;
    g_host_stage.release();
    return MITRE_OK;
}

extern "C" int mitre_score_all_host(const double *X_host, const double *omega_host,
                                    const double *response_host, const double *truth_host, double sigma2,
                                    int N, int p, int64_t P, double *out_host)
{
    if (N <= 0 || p <= 0 || P < 0 || !X_host || !omega_host || !response_host) return MITRE_ERR_INVALID;
    if (P > 0 && (!truth_host || !out_host)) return MITRE_ERR_INVALID;
    if (p > MITRE_MAX_P) return MITRE_ERR_UNSUPPORTED;
    const int W64 = words_for(N);
    const size_t ws_bytes = mitre_score_workspace_bytes(N, p);
    // chunk: 8 Mi rows (64 MiB of packed rows at N <= 64), fewer for wide rows.  (512 Ki-row chunks hide the
    // copies behind the packing of the next chunk but measured slower, 33.9 vs 27.9 ms for 8 Mi rows: the
    // packing threads are started per chunk.)
    int64_t chunk = (g_host_chunk_bytes > 0 ? g_host_chunk_bytes : (64ll << 20)) / ((int64_t)W64 * 8);
    if (chunk < 1024) chunk = 1024;
    if (chunk > P) chunk = P > 0 ? P : 1;

    std::lock_guard<std::mutex> guard(g_host_stage.lock);
    HostStage &hs = g_host_stage;
    int rc = MITRE_OK;
#define TRY(expr)                                   \
    do {                                            \
        cudaError_t _e = (expr);                    \
        if (_e != cudaSuccess) { hs.release(); return (int)_e; } \
    } while (0)
#define TRYRC(expr)                                 \
    do {                                            \
        rc = (expr);                                \
        if (rc != MITRE_OK) { cudaDeviceSynchronize(); return rc; } \
    } while (0)
    const size_t small = (size_t)N * p + 2 * (size_t)N;
    TRY(hs.ensure(chunk, W64, ws_bytes, small));
    double *d_X = hs.d_small, *d_omega = d_X + (size_t)N * p, *d_kappa = d_omega + N;
    // sweep inputs and the cleared status word: stream-ordered on st[0] from pinned memory; st[1] waits for
    // them through an event (the streams are non-blocking: nothing orders them against the default stream)
    memcpy(hs.h_small, X_host, sizeof(double) * N * p);
    memcpy(hs.h_small + (size_t)N * p, omega_host, sizeof(double) * N);
    for (int i = 0; i < N; ++i) hs.h_small[(size_t)N * p + N + i] = response_host[i] * omega_host[i];  // z = kappa/omega
    TRY(cudaMemcpyAsync(hs.d_small, hs.h_small, sizeof(double) * small, cudaMemcpyHostToDevice, hs.st[0]));
    TRY(cudaMemsetAsync(hs.d_status, 0, sizeof(int32_t), hs.st[0]));
    TRY(cudaEventRecord(hs.inputs_ready, hs.st[0]));
    TRY(cudaStreamWaitEvent(hs.st[1], hs.inputs_ready, 0));
    {
        int64_t start_of[2] = {0, 0}, rows_of[2] = {0, 0};     // chunk whose result sits in h_out[b]
        int64_t done = 0;
        int b = 0;
        while (done < P) {
            const int64_t n = (P - done < chunk) ? (P - done) : chunk;
            TRY(cudaStreamSynchronize(hs.st[b]));              // the chunk that used this staging set is complete
            if (rows_of[b]) copy_rows_host(out_host + start_of[b], hs.h_out[b], rows_of[b]);
            pack_rows_host(truth_host + done * N, n, N, W64, hs.h_packed[b]);   // overlaps the other set's GPU work
            TRY(cudaMemcpyAsync(hs.d_packed[b], hs.h_packed[b], sizeof(uint64_t) * n * W64, cudaMemcpyHostToDevice,
                                hs.st[b]));
            TRYRC(mitre_score_all(hs.st[b], hs.d_packed[b], n, W64, nullptr, d_X, d_omega, d_kappa, N, p, sigma2,
                                  hs.d_ws[b], ws_bytes, hs.d_out[b], hs.d_status));
            TRY(cudaMemcpyAsync(hs.h_out[b], hs.d_out[b], sizeof(double) * n, cudaMemcpyDeviceToHost, hs.st[b]));
            start_of[b] = done;
            rows_of[b] = n;
            done += n;
            b ^= 1;
        }
        for (int k = 0; k < 2; ++k, b ^= 1) {
            TRY(cudaStreamSynchronize(hs.st[b]));
            if (rows_of[b]) copy_rows_host(out_host + start_of[b], hs.h_out[b], rows_of[b]);
            rows_of[b] = 0;
        }
        if (P == 0) {
            TRYRC(mitre_score_all(hs.st[0], nullptr, 0, W64, nullptr, d_X, d_omega, d_kappa, N, p, sigma2, hs.d_ws[0],
                                  ws_bytes, nullptr, hs.d_status));
        }
        TRY(cudaMemcpyAsync(hs.h_status, hs.d_status, sizeof(int32_t), cudaMemcpyDeviceToHost, hs.st[0]));
        TRY(cudaStreamSynchronize(hs.st[0]));
        if (*hs.h_status & (MITRE_STATUS_BASE_NOT_PD | MITRE_STATUS_BAD_SCHUR)) rc = MITRE_ERR_NOT_PD;
    }
#undef TRY
#undef TRYRC
    return rc;
}

// Launches the DFMA probe; *flops_launched = 2 * 8 * iters * threads.  Time it with stream events.
extern "C" int mitre_probe_fp64(void *stream, int iters, double *sink, double *flops_launched)
{
    if (iters <= 0 || !sink) return MITRE_ERR_INVALID;
    const int blocks = sm_count() * 8;
    fp64_probe_kernel<<<blocks, 256, 0, as_stream(stream)>>>(iters, 1.0, sink);
    MITRE_LAUNCH_CHECK();
    if (flops_launched) *flops_launched = 2.0 * 8.0 * (double)iters * 256.0 * (double)blocks;
    return MITRE_OK;
}
