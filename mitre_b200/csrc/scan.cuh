// scan.cuh -- exclusive scan of (mult * int32 input) into int64 offsets; out has n+1 entries,
// out[n] = total.  Three launches: per-tile sums, scan of the tile sums (one CTA), per-tile
// scan + base.  Used for row offsets (mult = 2: two rows per threshold) and for the radix
// sort's digit histograms.
#pragma once
#include "common.cuh"

namespace mitre {

constexpr int kScanTile = 2048;  // elements per CTA (256 threads x 8)

static __global__ void __launch_bounds__(256)
scan_tile_sums_kernel(const int32_t *__restrict__ in, int64_t n, int mult, int64_t *__restrict__ tile_sums)
{
    __shared__ long long sred[256];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    long long acc = 0;
    for (int i = threadIdx.x; i < kScanTile; i += 256) {
        const int64_t g = base + i;
        if (g < n) acc += (long long)mult * in[g];
    }
    sred[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sred[threadIdx.x] += sred[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = sred[0];
}

static __global__ void __launch_bounds__(1024)
scan_tile_bases_kernel(int64_t *__restrict__ tile_sums, int64_t ntiles, int64_t *__restrict__ total_out)
{
    __shared__ long long s[1024];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t c0 = 0; c0 < ntiles; c0 += 1024) {
        const int64_t i = c0 + threadIdx.x;
        const long long v = i < ntiles ? tile_sums[i] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            const long long t = threadIdx.x >= off ? s[threadIdx.x - off] : 0;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < ntiles) tile_sums[i] = carry + s[threadIdx.x] - v;  // exclusive
        __syncthreads();
        if (threadIdx.x == 0) carry += s[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}

static __global__ void __launch_bounds__(256)
scan_apply_kernel(const int32_t *__restrict__ in, int64_t n, int mult, const int64_t *__restrict__ tile_bases,
                  int64_t *__restrict__ offsets)
{
    __shared__ long long swarp[8];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t g0 = base + (int64_t)threadIdx.x * 8;  // 8 consecutive elements per thread
    long long v[8], tsum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        v[j] = (g0 + j < n) ? (long long)mult * in[g0 + j] : 0;
        tsum += v[j];
    }
    long long inc = tsum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
    }
    if (lane == 31) swarp[wid] = inc;
    __syncthreads();
    long long wbase = 0;
    for (int w = 0; w < wid; ++w) wbase += swarp[w];
    long long run = tile_bases[blockIdx.x] + wbase + inc - tsum;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (g0 + j < n) offsets[g0 + j] = run;
        run += v[j];
    }
}

static inline size_t scan_workspace_bytes(int64_t n)
{
    if (n < 0) return 0;
    return sizeof(int64_t) * (size_t)(ceil_div64(n > 0 ? n : 1, kScanTile) + 2);
}

// out[0..n) = exclusive prefix sums of mult*in, out[n] = total.
static inline int exclusive_scan_i32(cudaStream_t st, const int32_t *in, int64_t n, int mult, int64_t *out,
                              void *workspace)
{
    if (n == 0) {
        MITRE_CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(int64_t), st));
        return MITRE_OK;
    }
    const int64_t ntiles = ceil_div64(n, kScanTile);
    if (ntiles > 0x7fffffffll) return MITRE_ERR_UNSUPPORTED;
    int64_t *tiles = static_cast<int64_t *>(workspace);
    scan_tile_sums_kernel<<<(unsigned)ntiles, 256, 0, st>>>(in, n, mult, tiles);
    MITRE_LAUNCH_CHECK();
    scan_tile_bases_kernel<<<1, 1024, 0, st>>>(tiles, ntiles, out + n);
    MITRE_LAUNCH_CHECK();
    scan_apply_kernel<<<(unsigned)ntiles, 256, 0, st>>>(in, n, mult, tiles, out);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

}  // namespace mitre
