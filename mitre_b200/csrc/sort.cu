// sort.cu -- stable lexicographic sort of the packed truth rows (replaces
// DiscreteRulePopulation.sort_base_rules, mitre/rules.py:439-462: np.lexsort(np.rot90(truth)),
// subject 0 the primary key, False < True, ties in enumeration order).
//
// With subject i at bit 63-(i%64) of word i/64, the reference order is plain unsigned
// lexicographic order of the words, so an LSD radix sort over 8-subject digits (least
// significant digit = the last 8 subjects) with a stable scatter reproduces np.lexsort exactly.
//
//   W64 == 1 : (key, index) pairs move together: coalesced key reads, no gathers.
//   W64  > 1 : only the permutation moves; each pass gathers its digit through the permutation.
//
// Per pass: digit histogram per 2048-row tile -> exclusive scan over [digit][tile] -> stable
// scatter (warp match_any ranking inside the tile, tile/warp order preserved).
#include "common.cuh"
#include "scan.cuh"

namespace mitre {

constexpr int kSortTile = 2048;   // rows per CTA: 8 warps x 8 rounds x 32 lanes
constexpr int kRounds = 8;

// Digit d (0 = most significant = subjects 0..7) of a row given by its words.
__device__ __forceinline__ int row_digit(const uint64_t *__restrict__ rows, int64_t row, int W64, int d)
{
    const uint64_t w = rows[row * W64 + (d >> 3)];
    return (int)((w >> (56 - 8 * (d & 7))) & 0xff);
}

// KV = true: keys_in holds one uint64 key per position (W64 == 1) and moves with the index.
template <bool KV>
__global__ void __launch_bounds__(256)
sort_hist_kernel(const uint64_t *__restrict__ rows, const uint64_t *__restrict__ keys_in,
                 const int64_t *__restrict__ perm_in, int64_t P, int W64, int d, int64_t ntiles,
                 int32_t *__restrict__ hist /*[256][ntiles]*/)
{
    __shared__ int cnt[256];
    cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
    for (int i = threadIdx.x; i < kSortTile; i += 256) {
        const int64_t pos = base + i;
        if (pos < P) {
            int dig;
            if (KV) dig = (int)((keys_in[pos] >> (56 - 8 * d)) & 0xff);
            else dig = row_digit(rows, perm_in ? perm_in[pos] : pos, W64, d);
            atomicAdd(&cnt[dig], 1);
        }
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * ntiles + blockIdx.x] = cnt[threadIdx.x];
}

template <bool KV>
__global__ void __launch_bounds__(256)
sort_scatter_kernel(const uint64_t *__restrict__ rows, const uint64_t *__restrict__ keys_in,
                    const int64_t *__restrict__ perm_in, int64_t P, int W64, int d, int64_t ntiles,
                    const int64_t *__restrict__ offsets /*[256][ntiles]*/, uint64_t *__restrict__ keys_out,
                    int64_t *__restrict__ perm_out)
{
    __shared__ int cnt[8][256];            // per-warp digit counts, then running bases
    __shared__ long long gbase[256];       // global offset of (digit, this tile)
    __shared__ int dbase[256];             // offset of the digit's run inside the tile
    __shared__ int wsum[8];
    __shared__ uint64_t skey[KV ? kSortTile : 1];
    __shared__ int64_t sidx[kSortTile];
    __shared__ uint8_t sdig[kSortTile];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * 256; i += 256) (&cnt[0][0])[i] = 0;
    gbase[threadIdx.x] = offsets[(int64_t)threadIdx.x * ntiles + blockIdx.x];
    __syncthreads();

    const int64_t tile_base = (int64_t)blockIdx.x * kSortTile;
    const int64_t wbase = tile_base + (int64_t)wid * (kRounds * 32);
    int dig[kRounds];
    uint64_t key[kRounds];
    int64_t idx[kRounds];
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const int64_t pos = wbase + r * 32 + lane;
        dig[r] = -1;
        key[r] = 0;
        idx[r] = 0;
        if (pos < P) {
            idx[r] = perm_in ? perm_in[pos] : pos;
            if (KV) {
                key[r] = keys_in[pos];
                dig[r] = (int)((key[r] >> (56 - 8 * d)) & 0xff);
            } else {
                dig[r] = row_digit(rows, idx[r], W64, d);
            }
        }
        const unsigned peers = __match_any_sync(0xffffffffu, dig[r]);
        if (dig[r] >= 0 && (peers & lt) == 0) cnt[wid][dig[r]] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // digit = threadIdx.x: exclusive prefix over warps, then exclusive prefix over digits of the totals
    {
        int run = 0;
        for (int w = 0; w < 8; ++w) {
            const int c = cnt[w][threadIdx.x];
            cnt[w][threadIdx.x] = run;
            run += c;
        }
        int inc = run;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += y;
        }
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        int before = 0;
        for (int w = 0; w < wid; ++w) before += wsum[w];
        dbase[threadIdx.x] = before + inc - run;
    }
    __syncthreads();
    // stable position inside the tile, ordered by digit: stage there, then write digit runs out
    // with consecutive threads on consecutive addresses (a direct scatter writes 8-byte pieces of
    // 256 different runs from every warp)
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const unsigned peers = __match_any_sync(0xffffffffu, dig[r]);
        int base = 0;
        if (dig[r] >= 0) base = cnt[wid][dig[r]];
        __syncwarp();
        if (dig[r] >= 0) {
            const int local = dbase[dig[r]] + base + __popc(peers & lt);
            sidx[local] = idx[r];
            sdig[local] = (uint8_t)dig[r];
            if (KV) skey[local] = key[r];
            if ((peers & lt) == 0) cnt[wid][dig[r]] = base + __popc(peers);
        }
        __syncwarp();
    }
    __syncthreads();
    const int n_valid = (int)((P - tile_base < kSortTile) ? (P - tile_base) : kSortTile);
    for (int i = threadIdx.x; i < n_valid; i += 256) {
        const int dg = sdig[i];
        const int64_t dst = gbase[dg] + (i - dbase[dg]);
        perm_out[dst] = sidx[i];
        if (KV) keys_out[dst] = skey[i];
    }
}

__global__ void __launch_bounds__(256)
iota_kernel(int64_t *__restrict__ p, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = i;
}

__global__ void __launch_bounds__(256)
gather_rows_kernel(const uint64_t *__restrict__ rows, const int64_t *__restrict__ perm, int64_t P, int W64,
                   uint64_t *__restrict__ out)
{
    const int64_t total = P * W64;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t i = e / W64;
        const int w = (int)(e % W64);
        out[e] = rows[perm[i] * W64 + w];
    }
}

struct SortWs {
    int64_t *perm_tmp;   // [P]
    uint64_t *keys_a;    // [P]   (W64 == 1 only)
    uint64_t *keys_b;    // [P]   (W64 == 1 only)
    int32_t *hist;       // [256 * ntiles]
    int64_t *offsets;    // [256 * ntiles + 1]
    void *scan_ws;
};

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t sort_ws_layout(int64_t P, int W64, void *base, SortWs *ws)
{
    const int64_t ntiles = ceil_div64(P > 0 ? P : 1, kSortTile);
    size_t off = 0;
    char *b = static_cast<char *>(base);
    auto take = [&](size_t bytes) {
        void *p = b ? b + off : nullptr;
        off += align256(bytes);
        return p;
    };
    void *perm_tmp = take(sizeof(int64_t) * (size_t)P);
    void *keys_a = take(W64 == 1 ? sizeof(uint64_t) * (size_t)P : 0);
    void *keys_b = take(W64 == 1 ? sizeof(uint64_t) * (size_t)P : 0);
    void *hist = take(sizeof(int32_t) * 256 * (size_t)ntiles);
    void *offsets = take(sizeof(int64_t) * (256 * (size_t)ntiles + 1));
    void *scan_ws = take(scan_workspace_bytes(256 * ntiles));
    if (ws) {
        ws->perm_tmp = static_cast<int64_t *>(perm_tmp);
        ws->keys_a = static_cast<uint64_t *>(keys_a);
        ws->keys_b = static_cast<uint64_t *>(keys_b);
        ws->hist = static_cast<int32_t *>(hist);
        ws->offsets = static_cast<int64_t *>(offsets);
        ws->scan_ws = scan_ws;
    }
    return off;
}

// sort_packed.cu: one-sweep sort of (row | index) words for tables with N + bits(P) <= 64
size_t lexsort_packed_workspace_bytes(int64_t P);
bool lexsort_packed_eligible(int64_t P, int N);
int lexsort_packed(cudaStream_t st, const uint64_t *rows, int64_t P, int N, int padded, int64_t P_valid,
                   const int64_t *row_offsets, int L, int64_t *perm, uint64_t *sorted_rows, void *workspace,
                   size_t workspace_bytes);
int lexsort_packed_failed(cudaStream_t st, void *workspace, int64_t P);

static int g_packed_sort = 1;

}  // namespace mitre

using namespace mitre;

extern "C" int mitre_set_packed_sort(int enable)
{
    const int old = g_packed_sort;
    g_packed_sort = enable ? 1 : 0;
    return old;
}

extern "C" size_t mitre_lexsort_workspace_bytes(int64_t P, int W64)
{
    if (P < 0 || W64 <= 0) return 0;
    size_t general = sort_ws_layout(P, W64, nullptr, nullptr);
    if (W64 == 1) {
        const size_t packed = lexsort_packed_workspace_bytes(P);
        if (packed > general) general = packed;
    }
    return general;
}

extern "C" int mitre_lexsort_padded(void *stream, const uint64_t *rows_padded, int64_t P_slots, int N,
                                    int64_t P_valid, const int64_t *row_offsets, int64_t *perm,
                                    uint64_t *sorted_rows, void *workspace, size_t workspace_bytes)
{
    if (P_slots < 0 || N < 2 || P_valid < 0 || P_valid > P_slots) return MITRE_ERR_INVALID;
    if (P_slots == 0 || P_valid == 0) return MITRE_OK;
    if (!rows_padded || !perm || !row_offsets || !workspace) return MITRE_ERR_INVALID;
    if (!g_packed_sort || !lexsort_packed_eligible(P_slots, N)) return MITRE_ERR_UNSUPPORTED;
    return lexsort_packed(as_stream(stream), rows_padded, P_slots, N, 1, P_valid, row_offsets, 2 * (N - 1), perm,
                          sorted_rows, workspace, workspace_bytes);
}

extern "C" int mitre_lexsort_check(void *stream, void *workspace, int64_t P, int N)
{
    if (!workspace || P <= 0 || !g_packed_sort || !lexsort_packed_eligible(P, N)) return MITRE_OK;
    return lexsort_packed_failed(as_stream(stream), workspace, P) ? MITRE_ERR_UNSUPPORTED : MITRE_OK;
}

extern "C" int mitre_lexsort_rows(void *stream, const uint64_t *rows, int64_t P, int W64, int N,
                                  int64_t *perm, uint64_t *sorted_rows, void *workspace,
                                  size_t workspace_bytes)
{
    if (P < 0 || N <= 0 || W64 != words_for(N)) return MITRE_ERR_INVALID;
    if (P == 0) return MITRE_OK;
    if (!rows || !perm || !workspace) return MITRE_ERR_INVALID;
    if (workspace_bytes < mitre_lexsort_workspace_bytes(P, W64)) return MITRE_ERR_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    if (W64 == 1 && g_packed_sort && lexsort_packed_eligible(P, N))
        return lexsort_packed(st, rows, P, N, 0, P, nullptr, 0, perm, sorted_rows, workspace, workspace_bytes);
    SortWs ws;
    sort_ws_layout(P, W64, workspace, &ws);
    const int64_t ntiles = ceil_div64(P, kSortTile);
    if (ntiles > 0x7fffffffll) return MITRE_ERR_UNSUPPORTED;
    const int ndigits = (N + 7) / 8;
    // ping-pong so that the final pass lands in `perm`
    int64_t *pin = nullptr;   // identity on the first pass
    int64_t *pout = (ndigits % 2 == 1) ? perm : ws.perm_tmp;
    const uint64_t *kin = rows;
    uint64_t *kout = (W64 == 1) ? ws.keys_a : nullptr;
    for (int d = ndigits - 1; d >= 0; --d) {
        if (W64 == 1) {
            sort_hist_kernel<true><<<(unsigned)ntiles, 256, 0, st>>>(rows, kin, pin, P, W64, d, ntiles, ws.hist);
        } else {
            sort_hist_kernel<false><<<(unsigned)ntiles, 256, 0, st>>>(rows, nullptr, pin, P, W64, d, ntiles,
                                                                       ws.hist);
        }
        MITRE_LAUNCH_CHECK();
        int rc = exclusive_scan_i32(st, ws.hist, 256 * ntiles, 1, ws.offsets, ws.scan_ws);
        if (rc != MITRE_OK) return rc;
        if (W64 == 1) {
            sort_scatter_kernel<true><<<(unsigned)ntiles, 256, 0, st>>>(rows, kin, pin, P, W64, d, ntiles,
                                                                         ws.offsets, kout, pout);
        } else {
            sort_scatter_kernel<false><<<(unsigned)ntiles, 256, 0, st>>>(rows, nullptr, pin, P, W64, d, ntiles,
                                                                          ws.offsets, nullptr, pout);
        }
        MITRE_LAUNCH_CHECK();
        pin = pout;
        pout = (pout == perm) ? ws.perm_tmp : perm;
        if (W64 == 1) {
            kin = kout;
            kout = (kout == ws.keys_a) ? ws.keys_b : ws.keys_a;
        }
    }
    if (sorted_rows) {
        if (W64 == 1) {
            MITRE_CUDA_TRY(cudaMemcpyAsync(sorted_rows, kin, sizeof(uint64_t) * (size_t)P,
                                           cudaMemcpyDeviceToDevice, st));
        } else {
            int64_t blocks = ceil_div64(P * W64, 256);
            const int64_t cap = (int64_t)sm_count() * 16;
            if (blocks > cap) blocks = cap;
            gather_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(rows, perm, P, W64, sorted_rows);
            MITRE_LAUNCH_CHECK();
        }
    }
    return MITRE_OK;
}
