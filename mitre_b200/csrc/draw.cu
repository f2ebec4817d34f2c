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

// ---- per-tile pairs ---------------------------------------------------------------------------------
// pairs[t] = (max, sum exp(w - max)) over the kDrawTileRows candidates of tile t.  The large-table
// scoring kernel writes these itself (scoring.cu, score_w1d_kernel); every other producer of weights
// gets them from this pass: one warp per tile, four consecutive candidates per lane.
__global__ void __launch_bounds__(256)
tile_pairs_kernel(const double *__restrict__ a, const double *__restrict__ b, int64_t P,
                  double *__restrict__ pairs)
{
    const int lane = threadIdx.x & 31;
    const int64_t ntiles = draw_ntiles(P);
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < ntiles; t += nwarps) {
        const int64_t k0 = t * kDrawTileRows + 4 * lane;
        double w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = (k0 + i < P) ? weight_at(a, b, k0 + i) : -INFINITY;
        double m = fmax(fmax(w[0], w[1]), fmax(w[2], w[3]));
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
        double sum = 0.0;
        if (m > -INFINITY && m < INFINITY) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double d = w[i] - m;
                sum += (d == d) ? exp(d) : 0.0;       // NaN weights (flagged rows) carry no mass
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        if (lane == 0) reinterpret_cast<double2 *>(pairs)[t] = make_double2(m, sum);
    }
}

// mass of tile t in units of exp(shift)
__device__ __forceinline__ double tile_term(const double *__restrict__ pairs, int64_t t, double shift)
{
    const double2 ms = reinterpret_cast<const double2 *>(pairs)[t];
    return ms.y > 0.0 ? ms.y * exp(ms.x - shift) : 0.0;
}

// chunk_sums[c] = mass of the kDrawChunkTiles tiles of chunk c, fixed summation order
__global__ void __launch_bounds__(kDrawChunkTiles)
draw_chunk_sums_kernel(const double *__restrict__ pairs, int64_t ntiles, const double *__restrict__ stats,
                       int stabilise, double *__restrict__ chunk_sums)
{
    __shared__ double sred[kDrawChunkTiles];
    const double shift = stabilise ? stats[0] : 0.0;
    const int64_t t = (int64_t)blockIdx.x * kDrawChunkTiles + threadIdx.x;
    sred[threadIdx.x] = t < ntiles ? tile_term(pairs, t, shift) : 0.0;
    __syncthreads();
    for (int s = kDrawChunkTiles / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sred[threadIdx.x] += sred[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) chunk_sums[blockIdx.x] = sred[0];
}

// Single CTA.  (1) locate the chunk the marker falls in by narrowing [lo, hi) 256-fold per round
// (each thread sums one contiguous run of chunk sums, thread 0 walks the 256 run sums left to
// right); (2) walk the chunk's 256 tile masses left to right; (3) resolve the 128 candidates of
// that tile with a strict left-to-right cumulative sum, the reference's own order
// (rules.py:1261-1264: np.cumsum, #{p < marker}).
__global__ void __launch_bounds__(256)
draw_select_kernel(const double *__restrict__ a, const double *__restrict__ b, int64_t P,
                   const double *__restrict__ stats, int stabilise, const double *__restrict__ pairs,
                   const double *__restrict__ chunk_sums, double u, int64_t *__restrict__ index_out)
{
    __shared__ long long s_sel;
    __shared__ double s_before, s_marker;
    __shared__ double s_run[256];
    const double shift = stabilise ? stats[0] : 0.0;
    const int64_t ntiles = draw_ntiles(P), nchunks = draw_nchunks(P);
    {
        int64_t lo = 0, hi = nchunks;
        double before = 0.0;
        bool have_marker = false;
        while (hi - lo > 1 || !have_marker) {
            const int64_t n = hi - lo;
            const int64_t c = (n + 255) / 256;
            const int64_t i0 = lo + (int64_t)threadIdx.x * c;
            const int64_t i1 = (i0 + c < hi) ? i0 + c : hi;
            double sum = 0.0;
            for (int64_t i = i0; i < i1; ++i) sum += chunk_sums[i];
            s_run[threadIdx.x] = sum;
            __syncthreads();
            if (threadIdx.x == 0) {
                if (!have_marker) {
                    double total = 0.0;
                    for (int k = 0; k < 256; ++k) total += s_run[k];
                    s_marker = u * total;
                }
                const double marker = s_marker;
                const int used = (int)((n + c - 1) / c);      // runs that hold at least one chunk
                double run = before;
                int sel = used - 1;
                for (int k = 0; k < used; ++k) {
                    const double nxt = run + s_run[k];
                    if (!(nxt < marker) || k == used - 1) { sel = k; break; }
                    run = nxt;
                }
                s_sel = sel;
                s_before = run;
            }
            __syncthreads();
            have_marker = true;
            const int64_t nlo = lo + (int64_t)s_sel * c;
            hi = (nlo + c < hi) ? nlo + c : hi;
            lo = nlo;
            before = s_before;
            __syncthreads();
        }
        // the chunk's tiles
        const int64_t t0 = lo * kDrawChunkTiles;
        const int64_t t = t0 + threadIdx.x;
        s_run[threadIdx.x] = t < ntiles ? tile_term(pairs, t, shift) : 0.0;
        __syncthreads();
        if (threadIdx.x == 0) {
            const int used = (int)((ntiles - t0 < kDrawChunkTiles) ? (ntiles - t0) : kDrawChunkTiles);
            const double marker = s_marker;
            double run = before;
            int sel = used - 1;
            for (int k = 0; k < used; ++k) {
                const double nxt = run + s_run[k];
                if (!(nxt < marker) || k == used - 1) { sel = k; break; }
                run = nxt;
            }
            s_sel = t0 + sel;
            s_before = run;
        }
        __syncthreads();
    }
    const int64_t tile = s_sel;
    const int64_t k0 = tile * kDrawTileRows;
    __syncthreads();
    if (threadIdx.x < kDrawTileRows) {
        const int64_t k = k0 + threadIdx.x;
        double e = 0.0;
        if (k < P) {
            const double d = weight_at(a, b, k) - shift;
            e = (d == d) ? exp(d) : 0.0;
        }
        s_run[threadIdx.x] = e;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int n = (int)((P - k0 < kDrawTileRows) ? (P - k0) : kDrawTileRows);
        const double marker = s_marker;
        double cum = s_before;
        int count = 0;
        for (int i = 0; i < n; ++i) {
            cum += s_run[i];
            count += (cum < marker) ? 1 : 0;
        }
        int64_t idx = k0 + count;
        if (idx > P - 1) idx = P - 1;
        *index_out = idx;
    }
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

int launch_tile_pairs(cudaStream_t st, const double *a, const double *b, int64_t P, double *pairs)
{
    int64_t blocks = ceil_div64(draw_ntiles(P), 8);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    tile_pairs_kernel<<<(unsigned)blocks, 256, 0, st>>>(a, b, P, pairs);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

}  // namespace mitre

using namespace mitre;

extern "C" size_t mitre_draw_workspace_bytes(int64_t P)
{
    if (P < 0) return 0;
    return sizeof(double) * draw_ws_doubles(P > 0 ? P : 1);
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

static int draw_select(cudaStream_t st, const double *a, const double *b, int64_t P, double u, int stabilise,
                       const double *stats, int64_t *index_out, double *ws)
{
    const int64_t nchunks = draw_nchunks(P);
    if (nchunks > 0x7fffffffll) return MITRE_ERR_UNSUPPORTED;
    const double *pairs = ws + draw_off_pairs();
    double *chunk_sums = ws + draw_off_chunks(P);
    draw_chunk_sums_kernel<<<(unsigned)nchunks, kDrawChunkTiles, 0, st>>>(pairs, draw_ntiles(P), stats, stabilise,
                                                                         chunk_sums);
    MITRE_LAUNCH_CHECK();
    draw_select_kernel<<<1, 256, 0, st>>>(a, b, P, stats, stabilise, pairs, chunk_sums, u, index_out);
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
    double *ws = static_cast<double *>(workspace);
    int rc = launch_tile_pairs(st, a, b, P, ws + draw_off_pairs());
    if (rc != MITRE_OK) return rc;
    return draw_select(st, a, b, P, u, stabilise, stats, index_out, ws);
}

extern "C" int mitre_draw_from_tiles(void *stream, const double *a, const double *b, int64_t P, double u,
                                     int stabilise, const double *stats, int64_t *index_out, void *workspace,
                                     size_t workspace_bytes)
{
    if (P <= 0 || !a || !index_out || !workspace) return MITRE_ERR_INVALID;
    if (stabilise && !stats) return MITRE_ERR_INVALID;
    if (workspace_bytes < mitre_draw_workspace_bytes(P)) return MITRE_ERR_WORKSPACE;
    return draw_select(as_stream(stream), a, b, P, u, stabilise, stats, index_out, static_cast<double *>(workspace));
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
