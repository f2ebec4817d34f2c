// draw.cu -- proposal draw over all candidates (replaces the host-side weight assembly and
// choose_from_relative_probabilities, mitre/rules.py:1261-1264 as used at
// logit_rules.py:861-863, 958-959, 1168-1170):
//
//     p = cumsum(exp(w));  marker = u * p[-1];  index = #{ p < marker }
//
// mitre_weight_stats gives (max w, sum exp(w - max)) -- the 16-byte pair ranks exchange for the
// two-stage multi-GPU draw and the log-sum-exp the add/remove moves need
// (logit_rules.py:959, 1035-1036, 1064).
//
// The cumulative sum is hierarchical (8 consecutive elements per thread, warp scan, tile of
// 2048, sequential over tiles): exact for the selection except when `marker` falls within
// rounding of a partial sum, where a sequential np.cumsum could differ by one index.
#include <math.h>

#include "common.cuh"

namespace mitre {

constexpr int kDrawTile = 2048;

__device__ __forceinline__ double weight_at(const double *__restrict__ a, const double *__restrict__ b,
                                            int64_t i)
{
    return b ? a[i] + b[i] : a[i];
}

// (max, sum exp(x - max)) merge; empty = (-inf, 0)
struct MaxSum {
    double m, s;
};
__device__ __forceinline__ MaxSum ms_merge(MaxSum x, MaxSum y)
{
    if (y.s == 0.0 && y.m == -INFINITY) return x;
    if (x.s == 0.0 && x.m == -INFINITY) return y;
    MaxSum r;
    if (x.m >= y.m) { r.m = x.m; r.s = x.s + y.s * exp(y.m - x.m); }
    else { r.m = y.m; r.s = y.s + x.s * exp(x.m - y.m); }
    return r;
}

__global__ void __launch_bounds__(256)
stats_partial_kernel(const double *__restrict__ a, const double *__restrict__ b, int64_t P,
                     double *__restrict__ part /*[grid][2]*/)
{
    __shared__ double sm[256], ss[256];
    MaxSum acc = {-INFINITY, 0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += stride) {
        const double w = weight_at(a, b, i);
        if (w == -INFINITY) continue;
        if (w > acc.m) { acc.s = acc.s * exp(acc.m - w) + 1.0; acc.m = w; }
        else acc.s += exp(w - acc.m);
    }
    sm[threadIdx.x] = acc.m; ss[threadIdx.x] = acc.s;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            MaxSum x = {sm[threadIdx.x], ss[threadIdx.x]}, y = {sm[threadIdx.x + s], ss[threadIdx.x + s]};
            x = ms_merge(x, y);
            sm[threadIdx.x] = x.m; ss[threadIdx.x] = x.s;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = sm[0]; part[2 * blockIdx.x + 1] = ss[0]; }
}

__global__ void __launch_bounds__(256)
stats_final_kernel(const double *__restrict__ part, int nparts, double *__restrict__ stats)
{
    __shared__ double sm[256], ss[256];
    MaxSum acc = {-INFINITY, 0.0};
    for (int i = threadIdx.x; i < nparts; i += 256) {
        MaxSum y = {part[2 * i], part[2 * i + 1]};
        acc = ms_merge(acc, y);
    }
    sm[threadIdx.x] = acc.m; ss[threadIdx.x] = acc.s;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            MaxSum x = {sm[threadIdx.x], ss[threadIdx.x]}, y = {sm[threadIdx.x + s], ss[threadIdx.x + s]};
            x = ms_merge(x, y);
            sm[threadIdx.x] = x.m; ss[threadIdx.x] = x.s;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { stats[0] = sm[0]; stats[1] = ss[0]; }
}

// Inclusive scan of one 2048-element tile of exp(w - shift); returns this thread's 8 inclusive
// values in inc[] and the tile total through smem.
__device__ __forceinline__ void tile_scan(const double *__restrict__ a, const double *__restrict__ b, int64_t P,
                                          int64_t tile, double shift, double (&inc)[8], double *swarp,
                                          double *tile_total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t i0 = tile * kDrawTile + (int64_t)threadIdx.x * 8;
    double run = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double e = (i0 + j < P) ? exp(weight_at(a, b, i0 + j) - shift) : 0.0;
        run += e;
        inc[j] = run;
    }
    double winc = run;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, winc, off);
        if (lane >= off) winc += t;
    }
    if (lane == 31) swarp[wid] = winc;
    __syncthreads();
    double base = winc - run;
    double tot = 0.0;
    for (int w = 0; w < 8; ++w) {
        if (w < wid) base += swarp[w];
        tot += swarp[w];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) inc[j] += base;
    if (threadIdx.x == 0) *tile_total = tot;
    __syncthreads();
}

__global__ void __launch_bounds__(256)
draw_tile_sums_kernel(const double *__restrict__ a, const double *__restrict__ b, int64_t P,
                      const double *__restrict__ stats, int stabilise, double *__restrict__ tile_sums)
{
    __shared__ double swarp[8];
    __shared__ double tot;
    double inc[8];
    const double shift = stabilise ? stats[0] : 0.0;
    tile_scan(a, b, P, blockIdx.x, shift, inc, swarp, &tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// single CTA: locate the tile the marker falls in by narrowing [lo, hi) 256-fold per round (each
// thread sums one contiguous run of tile sums, thread 0 walks the 256 run sums left to right), then
// rescan that tile and count.  A single thread walking all P/2048 tile sums, as the first version
// did, costs 14 ms at P = 2.9e8.
__global__ void __launch_bounds__(256)
draw_select_kernel(const double *__restrict__ a, const double *__restrict__ b, int64_t P,
                   const double *__restrict__ stats, int stabilise, double *__restrict__ tile_sums,
                   int64_t ntiles, double u, int64_t *__restrict__ index_out)
{
    __shared__ double swarp[8];
    __shared__ double tot;
    __shared__ long long s_tile;
    __shared__ double s_before, s_marker;
    __shared__ double s_run[256];
    {
        int64_t lo = 0, hi = ntiles;
        double before = 0.0;
        bool have_marker = false;
        while (hi - lo > 1 || !have_marker) {
            const int64_t n = hi - lo;
            const int64_t c = (n + 255) / 256;
            const int64_t i0 = lo + (int64_t)threadIdx.x * c;
            const int64_t i1 = (i0 + c < hi) ? i0 + c : hi;
            double sum = 0.0;
            for (int64_t i = i0; i < i1; ++i) sum += tile_sums[i];
            s_run[threadIdx.x] = sum;
            __syncthreads();
            if (threadIdx.x == 0) {
                if (!have_marker) {
                    double total = 0.0;
                    for (int k = 0; k < 256; ++k) total += s_run[k];
                    s_marker = u * total;
                }
                const double marker = s_marker;
                const int used = (int)((n + c - 1) / c);      // runs that hold at least one tile
                double run = before;
                int sel = used - 1;
                for (int k = 0; k < used; ++k) {
                    const double nxt = run + s_run[k];
                    if (!(nxt < marker) || k == used - 1) { sel = k; break; }
                    run = nxt;
                }
                s_tile = sel;
                s_before = run;
            }
            __syncthreads();
            have_marker = true;
            const int64_t nlo = lo + (int64_t)s_tile * c;
            hi = (nlo + c < hi) ? nlo + c : hi;
            lo = nlo;
            before = s_before;
            __syncthreads();
        }
        if (threadIdx.x == 0) s_tile = lo;
    }
    __syncthreads();
    const long long sel = s_tile;
    double inc[8];
    const double shift = stabilise ? stats[0] : 0.0;
    tile_scan(a, b, P, sel, shift, inc, swarp, &tot);
    int c = 0;
    const int64_t i0 = sel * kDrawTile + (int64_t)threadIdx.x * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (i0 + j < P && (s_before + inc[j]) < s_marker) ++c;
    // block sum of c
    __shared__ int scount[256];
    scount[threadIdx.x] = c;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) scount[threadIdx.x] += scount[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *index_out = sel * kDrawTile + scount[0];
}

// ---- primitive prior on the device (flat_prior, logit_rules.py:424-457) ---------------------------
// row -> group by binary search of the row's enumeration index in the group row offsets
__global__ void __launch_bounds__(256)
row_groups_kernel(const int64_t *__restrict__ perm, int64_t P, const int64_t *__restrict__ row_offsets,
                  int64_t G, int32_t *__restrict__ row_group)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += stride) {
        const int64_t r = perm[i];
        int64_t lo = 0, hi = G;          // last g with row_offsets[g] <= r
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (row_offsets[mid] <= r) lo = mid;
            else hi = mid;
        }
        row_group[i] = (int32_t)lo;
    }
}

// out[i] = by_variable[gvar[g]] + by_window[gwin[g]] - norm,  g = row_group[i]
__global__ void __launch_bounds__(256)
row_prior_kernel(const int32_t *__restrict__ row_group, int64_t P, const int32_t *__restrict__ gvar,
                 const int32_t *__restrict__ gwin, const double *__restrict__ by_variable,
                 const double *__restrict__ by_window, double norm, double *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += stride) {
        const int g = row_group[i];
        out[i] = (__ldg(by_variable + __ldg(gvar + g)) + __ldg(by_window + __ldg(gwin + g))) - norm;
    }
}

constexpr int kStatsBlocks = 592;   // 148 SMs x 4

}  // namespace mitre

using namespace mitre;

extern "C" size_t mitre_draw_workspace_bytes(int64_t P)
{
    if (P < 0) return 0;
    const size_t ntiles = (size_t)ceil_div64(P > 0 ? P : 1, kDrawTile);
    return sizeof(double) * (2 * (size_t)kStatsBlocks + 2 + ntiles);
}

extern "C" int mitre_weight_stats(void *stream, const double *a, const double *b, int64_t P, double *stats,
                                  void *workspace, size_t workspace_bytes)
{
    if (P <= 0 || !a || !stats || !workspace) return MITRE_ERR_INVALID;
    if (workspace_bytes < mitre_draw_workspace_bytes(P)) return MITRE_ERR_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    double *part = static_cast<double *>(workspace);
    int64_t blocks = ceil_div64(P, 256);
    if (blocks > kStatsBlocks) blocks = kStatsBlocks;
    stats_partial_kernel<<<(unsigned)blocks, 256, 0, st>>>(a, b, P, part);
    MITRE_LAUNCH_CHECK();
    stats_final_kernel<<<1, 256, 0, st>>>(part, (int)blocks, stats);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

extern "C" int mitre_draw_index(void *stream, const double *a, const double *b, int64_t P, double u,
                                int stabilise, double *stats, int64_t *index_out, void *workspace,
                                size_t workspace_bytes)
{
    if (P <= 0 || !a || !index_out || !workspace) return MITRE_ERR_INVALID;
    if (stabilise && !stats) return MITRE_ERR_INVALID;
    if (workspace_bytes < mitre_draw_workspace_bytes(P)) return MITRE_ERR_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    if (stabilise == 1) {   // 2: stats[0] already holds max(a + b), e.g. from mitre_score_all_stats
        int rc = mitre_weight_stats(stream, a, b, P, stats, workspace, workspace_bytes);
        if (rc != MITRE_OK) return rc;
    }
    const int64_t ntiles = ceil_div64(P, kDrawTile);
    if (ntiles > 0x7fffffffll) return MITRE_ERR_UNSUPPORTED;
    double *tile_sums = static_cast<double *>(workspace) + 2 * kStatsBlocks + 2;
    draw_tile_sums_kernel<<<(unsigned)ntiles, 256, 0, st>>>(a, b, P, stats, stabilise, tile_sums);
    MITRE_LAUNCH_CHECK();
    draw_select_kernel<<<1, 256, 0, st>>>(a, b, P, stats, stabilise, tile_sums, ntiles, u, index_out);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

extern "C" int mitre_row_groups(void *stream, const int64_t *perm, int64_t P, const int64_t *row_offsets,
                                int64_t G, int32_t *row_group)
{
    if (P < 0 || G <= 0) return MITRE_ERR_INVALID;
    if (P == 0) return MITRE_OK;
    if (!perm || !row_offsets || !row_group) return MITRE_ERR_INVALID;
    int64_t blocks = ceil_div64(P, 256);
    if (blocks > 16 * sm_count()) blocks = 16 * sm_count();
    row_groups_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(perm, P, row_offsets, G, row_group);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

extern "C" int mitre_row_prior(void *stream, const int32_t *row_group, int64_t P, const int32_t *group_variable,
                               const int32_t *group_window, const double *by_variable, const double *by_window,
                               double norm, double *out)
{
    if (P < 0) return MITRE_ERR_INVALID;
    if (P == 0) return MITRE_OK;
    if (!row_group || !group_variable || !group_window || !by_variable || !by_window || !out)
        return MITRE_ERR_INVALID;
    int64_t blocks = ceil_div64(P, 256);
    if (blocks > 16 * sm_count()) blocks = 16 * sm_count();
    row_prior_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(row_group, P, group_variable, group_window,
                                                                      by_variable, by_window, norm, out);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}
