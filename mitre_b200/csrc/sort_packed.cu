// sort_packed.cu -- the lexsort of sort.cu for tables whose row AND enumeration index fit one 64-bit word
// (N + bits(P) <= 64; S2 and S5 are N = 35 with P < 2^29), as a one-sweep LSD radix sort.
//
// Reference: DiscreteRulePopulation.sort_base_rules, mitre/rules.py:439-462 (np.lexsort(np.rot90(truth)),
// subject 0 the primary key, ties in enumeration order).
//
// A packed row keeps subject i at bit 63 - i, so its low 64 - N bits are zero: word = row | index is a
// single sort key whose unsigned order is exactly (row lexicographic, then enumeration index) -- the
// reference's stable order.  The sort then moves 8 bytes per row per pass instead of 16 (key + int64
// index), runs ceil(N / 9) passes of <= 9-bit digits over the N row bits only (4 at N = 35 instead of 5
// passes of 8 bits over N + 1), and every pass reads its input ONCE:
//   * the digit histograms of all passes come from one read of the table (os_hist_kernel);
//   * a pass is one kernel (os_pass_kernel): a CTA takes the next 4096-row tile by ticket, ranks its rows
//     by digit (warp match_any ranking, warp / round order preserved => stable), publishes the tile's
//     per-digit counts, obtains the counts of all earlier tiles by decoupled look-back over the published
//     (flag | count) words of its predecessors, stages the tile in shared memory ordered by digit and
//     writes each digit's run with consecutive threads on consecutive addresses;
//   * the FIRST pass reads the caller's rows directly and drops the ~0 sentinel slots of the padded
//     detector output (fused_detectors.cuh), so sentinels are never sorted;
//   * the LAST pass writes the two outputs themselves -- the sorted rows and perm, with the padded slot
//     index already mapped to the dense enumeration index (what densify_perm_kernel did in a pass of its
//     own).
// Tiles are taken by ticket, so every tile a CTA waits for belongs to a CTA that is already running;
// every wait is bounded all the same (an expired wait sets an error word that mitre_lexsort_check reads back).
#include "common.cuh"

namespace mitre {

constexpr int kOsThreads = 512;
constexpr int kOsWarps = kOsThreads / 32;
#ifndef MITRE_OS_ITEMS
#define MITRE_OS_ITEMS 8
#endif
#ifndef MITRE_OS_MINB
#define MITRE_OS_MINB 2
#endif
constexpr int kOsItems = MITRE_OS_ITEMS;
constexpr int kOsTile = kOsThreads * kOsItems;   // 4096 rows per CTA
constexpr int kOsMaxBits = 9;
constexpr int kOsBins = 1 << kOsMaxBits;
constexpr int kOsMaxPasses = 8;                  // ceil(64 / 9)
constexpr uint32_t kOsFlagAgg = 1u << 30;        // status word = flag | count (count < 2^30)
constexpr uint32_t kOsFlagIncl = 2u << 30;
constexpr uint32_t kOsCountMask = (1u << 30) - 1u;
constexpr int kOsLook = 8;                       // predecessors read per look-back round

struct OsPlan {
    int npass;
    int shift[kOsMaxPasses];   // pass 0 = least significant digit
    int bits[kOsMaxPasses];
};

static OsPlan make_plan(int N)
{
    OsPlan pl{};
    pl.npass = (N + kOsMaxBits - 1) / kOsMaxBits;
    const int base = N / pl.npass, extra = N % pl.npass;
    int shift = 64 - N;
    for (int p = 0; p < pl.npass; ++p) {
        pl.bits[p] = base + (p < extra ? 1 : 0);
        pl.shift[p] = shift;
        shift += pl.bits[p];
    }
    return pl;
}

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t *p, uint32_t v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Digit histograms of every pass from one read of the table.  Neighbouring slots of the enumeration order
// are the nested threshold sets of one group -- (above_k, below_k) pairs, the sets shrinking slowly with k --
// so the lanes of a warp mostly hit a handful of bins.  A thread therefore owns 16 consecutive rows (one
// 128-byte line) and keeps a one-entry run (digit, count) per pass and per parity (the even slots are the
// `above` rows, the odd ones their complements), touching shared memory only when the digit changes.
constexpr int kOsHistRows = 16;

__global__ void __launch_bounds__(kOsThreads)
os_hist_kernel(const uint64_t *__restrict__ rows, int64_t P, int padded, OsPlan plan,
               uint32_t *__restrict__ ghist /*[npass][kOsBins]*/)
{
    __shared__ uint32_t h[kOsMaxPasses][kOsBins];
    for (int i = threadIdx.x; i < kOsMaxPasses * kOsBins; i += kOsThreads) (&h[0][0])[i] = 0;
    __syncthreads();
    const int64_t nchunks = ceil_div64(P, kOsHistRows);
    const int64_t stride = (int64_t)gridDim.x * kOsThreads;
    for (int64_t c = (int64_t)blockIdx.x * kOsThreads + threadIdx.x; c < nchunks; c += stride) {
        const int64_t i0 = c * kOsHistRows;
        uint64_t k[kOsHistRows];
        if (i0 + kOsHistRows <= P && (reinterpret_cast<uintptr_t>(rows) & 15) == 0) {
            const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(rows + i0);
#pragma unroll
            for (int j = 0; j < kOsHistRows / 2; ++j) {
                const ulonglong2 v = __ldg(src + j);
                k[2 * j] = v.x;
                k[2 * j + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < kOsHistRows; ++j) k[j] = (i0 + j < P) ? __ldg(rows + i0 + j) : ~0ull;
        }
        const bool tail = i0 + kOsHistRows > P;
        for (int p = 0; p < plan.npass; ++p) {
            const int sh = plan.shift[p];
            const uint32_t mask = (1u << plan.bits[p]) - 1u;
#pragma unroll
            for (int par = 0; par < 2; ++par) {
                uint32_t cur = 0, cnt = 0;
#pragma unroll
                for (int j = par; j < kOsHistRows; j += 2) {
                    // past the end the slot reads ~0; without padding ~0 is not a possible row for N <= 63
                    const bool valid = !((padded || tail) && k[j] == ~0ull);
                    const uint32_t d = (uint32_t)(k[j] >> sh) & mask;
                    if (valid) {
                        if (cnt && d != cur) {
                            atomicAdd(&h[p][cur], cnt);
                            cnt = 0;
                        }
                        cur = d;
                        ++cnt;
                    }
                }
                if (cnt) atomicAdd(&h[p][cur], cnt);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plan.npass * kOsBins; i += kOsThreads) {
        const uint32_t v = (&h[0][0])[i];
        if (v) atomicAdd(&ghist[i], v);
    }
}

// exclusive scan over the bins of every pass, in place (one CTA per pass)
__global__ void __launch_bounds__(kOsBins)
os_offsets_kernel(uint32_t *__restrict__ ghist)
{
    __shared__ uint32_t wsum[kOsBins / 32];
    uint32_t *h = ghist + (size_t)blockIdx.x * kOsBins;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t v = h[threadIdx.x];
    uint32_t inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += y;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    uint32_t before = 0;
    for (int w = 0; w < wid; ++w) before += wsum[w];
    h[threadIdx.x] = before + inc - v;
}

struct OsShared {
    uint64_t key[kOsTile];               // the tile, ordered by digit
    uint32_t wcnt[kOsWarps][kOsBins];    // per-warp digit counts, then the warps' exclusive prefixes
    uint32_t dbase[kOsBins];             // offset of the digit's run inside the tile
    uint32_t gbase[kOsBins];             // global offset of (digit, this tile)
    uint32_t run[kOsBins];               // the tile's count of the digit
    uint32_t wsum[kOsWarps];
    uint32_t tile;
    uint32_t n_valid;
};

// One pass.  FIRST: `in` is the caller's table (P slots; ~0 = dropped when `padded`), the index is OR-ed in.
// LAST: the outputs are perm / rows_out instead of packed words.
template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(kOsThreads, MITRE_OS_MINB)
os_pass_kernel(const uint64_t *__restrict__ in, int64_t P, int padded, int shift, int bits, int idx_bits,
               const uint32_t *__restrict__ gofs /*[kOsBins]*/, uint32_t *__restrict__ status /*[ntiles][kOsBins]*/,
               uint32_t *__restrict__ ticket, uint64_t *__restrict__ out_packed, int64_t *__restrict__ perm_out,
               uint64_t *__restrict__ rows_out, const int64_t *__restrict__ row_offsets, uint32_t L,
               int32_t *__restrict__ err)
{
    extern __shared__ __align__(16) unsigned char os_smem[];
    OsShared &S = *reinterpret_cast<OsShared *>(os_smem);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t mask = (1u << bits) - 1u;
    const unsigned lt = (1u << lane) - 1u;
    for (int i = tid; i < kOsWarps * kOsBins; i += kOsThreads) (&S.wcnt[0][0])[i] = 0;
    if (tid == 0) S.tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = S.tile;
    const int64_t wbase = (int64_t)tile * kOsTile + (int64_t)wid * (kOsItems * 32);

    uint64_t key[kOsItems];
    uint32_t dig[kOsItems];
    uint32_t rank[kOsItems];
#pragma unroll
    for (int r = 0; r < kOsItems; ++r) {
        const int64_t pos = wbase + r * 32 + lane;
        key[r] = 0;
        dig[r] = 0xffffffffu;
        if (pos < P) {
            uint64_t k = ld_stream_u64(in + pos);
            bool valid = true;
            if (FIRST) {
                valid = !(padded && k == ~0ull);
                k |= (uint64_t)pos;
            }
            key[r] = k;
            if (valid) dig[r] = (uint32_t)(k >> shift) & mask;
        }
    }
#pragma unroll
    for (int r = 0; r < kOsItems; ++r) {
        const unsigned peers = __match_any_sync(0xffffffffu, dig[r]);
        const bool valid = dig[r] != 0xffffffffu;
        uint32_t prev = 0;
        if (valid) prev = S.wcnt[wid][dig[r]];
        __syncwarp();
        rank[r] = prev + __popc(peers & lt);
        if (valid && (peers & lt) == 0) S.wcnt[wid][dig[r]] = prev + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // thread = bin: exclusive prefix over the warps, the tile's count of the bin, publish, look back
    {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kOsWarps; ++w) {
            const uint32_t c = S.wcnt[w][tid];
            S.wcnt[w][tid] = run;
            run += c;
        }
        // bins beyond 2^bits stay empty: nothing to publish or to wait for
        if (tid <= (int)mask && tile > 0) st_relaxed_u32(status + (size_t)tile * kOsBins + tid, kOsFlagAgg | run);
        S.run[tid] = run;
        uint32_t inc = run;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += y;
        }
        if (lane == 31) S.wsum[wid] = inc;
        __syncthreads();
        uint32_t before = 0;
        for (int w = 0; w < wid; ++w) before += S.wsum[w];
        S.dbase[tid] = before + inc - run;
        if (tid == kOsBins - 1) S.n_valid = before + inc;
    }
    __syncthreads();
    // the tile goes to shared memory ordered by digit while the predecessors make progress
#pragma unroll
    for (int r = 0; r < kOsItems; ++r) {
        if (dig[r] != 0xffffffffu) S.key[S.dbase[dig[r]] + S.wcnt[wid][dig[r]] + rank[r]] = key[r];
    }
    // decoupled look-back, thread = bin: counts of this bin in all earlier tiles.  kOsLook predecessors are
    // read at once (independent loads in flight) and consumed nearest first up to the first inclusive one.
    {
        const bool active = tid <= (int)mask;
        uint32_t *mine = status + (size_t)tile * kOsBins + tid;
        uint32_t excl = 0;
        bool failed = false, done = !active || tile == 0;
        for (int64_t t = (int64_t)tile - 1; !done; t -= kOsLook) {
            uint32_t sv[kOsLook];
#pragma unroll
            for (int j = 0; j < kOsLook; ++j)
                sv[j] = (t - j >= 0) ? ld_relaxed_u32(status + (size_t)(t - j) * kOsBins + tid) : kOsFlagIncl;
#pragma unroll
            for (int j = 0; j < kOsLook; ++j) {
                if (done) break;
                uint32_t v = sv[j];
                uint32_t spins = 0;
                while ((v & ~kOsCountMask) == 0) {
                    if (++spins > (1u << 22)) {
                        failed = true;
                        v = kOsFlagIncl;
                        break;
                    }
                    if (spins > 16) __nanosleep(64);
                    v = ld_relaxed_u32(status + (size_t)(t - j) * kOsBins + tid);
                }
                excl += v & kOsCountMask;
                if (v & kOsFlagIncl) done = true;
            }
        }
        if (failed) atomicExch(err, 1);
        if (active) st_relaxed_u32(mine, kOsFlagIncl | ((excl + S.run[tid]) & kOsCountMask));
        S.gbase[tid] = gofs[tid] + excl;
    }
    __syncthreads();
    const uint32_t n_valid = S.n_valid;
    const uint64_t idx_mask = (idx_bits >= 64) ? ~0ull : ((1ull << idx_bits) - 1ull);
    for (uint32_t i = tid; i < n_valid; i += kOsThreads) {
        const uint64_t k = S.key[i];
        const uint32_t dg = (uint32_t)(k >> shift) & mask;
        const int64_t dst = (int64_t)S.gbase[dg] + (i - S.dbase[dg]);
        if (LAST) {
            const uint64_t slot = k & idx_mask;
            int64_t e = (int64_t)slot;
            if (row_offsets) {
                const uint32_t g = (uint32_t)slot / L;
                e = row_offsets[g] + (int64_t)((uint32_t)slot - g * L);
            }
            perm_out[dst] = e;
            if (rows_out) rows_out[dst] = k & ~idx_mask;
        } else {
            out_packed[dst] = k;
        }
    }
}

struct OsWs {
    uint64_t *buf_a;      // [P] packed words
    uint64_t *buf_b;      // [P]
    uint32_t *status;     // [ntiles][kOsBins]
    uint32_t *ghist;      // [kOsMaxPasses][kOsBins]
    uint32_t *ticket;     // [kOsMaxPasses]
    int32_t *err;         // [1]
};

static size_t os_align(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t os_layout(int64_t P, void *base, OsWs *ws)
{
    const int64_t ntiles = ceil_div64(P > 0 ? P : 1, kOsTile);
    size_t off = 0;
    char *b = static_cast<char *>(base);
    auto take = [&](size_t bytes) {
        void *p = b ? b + off : nullptr;
        off += os_align(bytes);
        return p;
    };
    void *a = take(sizeof(uint64_t) * (size_t)P);
    void *bb = take(sizeof(uint64_t) * (size_t)P);
    void *status = take(sizeof(uint32_t) * (size_t)ntiles * kOsBins);
    void *ghist = take(sizeof(uint32_t) * kOsMaxPasses * kOsBins);
    void *ticket = take(sizeof(uint32_t) * kOsMaxPasses);
    void *err = take(sizeof(int32_t));
    if (ws) {
        ws->buf_a = static_cast<uint64_t *>(a);
        ws->buf_b = static_cast<uint64_t *>(bb);
        ws->status = static_cast<uint32_t *>(status);
        ws->ghist = static_cast<uint32_t *>(ghist);
        ws->ticket = static_cast<uint32_t *>(ticket);
        ws->err = static_cast<int32_t *>(err);
    }
    return off;
}

size_t lexsort_packed_workspace_bytes(int64_t P) { return os_layout(P, nullptr, nullptr); }

bool lexsort_packed_eligible(int64_t P, int N)
{
    if (N <= 0 || N > 63 || P <= 0 || P >= (1ll << 30)) return false;
    int idx_bits = 1;
    while ((1ll << idx_bits) < P) ++idx_bits;
    return N + idx_bits <= 64;
}

template <bool FIRST, bool LAST>
static int launch_pass(cudaStream_t st, int64_t ntiles, const uint64_t *in, int64_t P, int padded, int shift, int bits,
                       int idx_bits, const uint32_t *gofs, uint32_t *status, uint32_t *ticket, uint64_t *out_packed,
                       int64_t *perm_out, uint64_t *rows_out, const int64_t *row_offsets, uint32_t L, int32_t *err)
{
    // per launch, not once per process: the attribute belongs to the current device's copy of the kernel
    MITRE_CUDA_TRY(cudaFuncSetAttribute(os_pass_kernel<FIRST, LAST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(OsShared)));
    os_pass_kernel<FIRST, LAST><<<(unsigned)ntiles, kOsThreads, sizeof(OsShared), st>>>(
        in, P, padded, shift, bits, idx_bits, gofs, status, ticket, out_packed, perm_out, rows_out, row_offsets, L, err);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

// rows uint64[P] (W64 == 1).  padded: slots equal to ~0 are dropped and only the P_valid others are sorted;
// row_offsets (optional, with L = slots per group): slot index -> dense enumeration index in perm.
// Returns MITRE_ERR_UNSUPPORTED when the table is not eligible (the caller takes the general path).
int lexsort_packed(cudaStream_t st, const uint64_t *rows, int64_t P, int N, int padded, int64_t P_valid,
                   const int64_t *row_offsets, int L, int64_t *perm, uint64_t *sorted_rows, void *workspace,
                   size_t workspace_bytes)
{
    if (!lexsort_packed_eligible(P, N)) return MITRE_ERR_UNSUPPORTED;
    if (workspace_bytes < lexsort_packed_workspace_bytes(P)) return MITRE_ERR_WORKSPACE;
    OsWs ws;
    os_layout(P, workspace, &ws);
    int idx_bits = 1;
    while ((1ll << idx_bits) < P) ++idx_bits;
    const OsPlan plan = make_plan(N);
    MITRE_CUDA_TRY(cudaMemsetAsync(ws.ghist, 0,
                                   sizeof(uint32_t) * kOsMaxPasses * kOsBins + os_align(sizeof(uint32_t) * kOsMaxPasses) +
                                       sizeof(int32_t),
                                   st));
    {
        int64_t blocks = ceil_div64(ceil_div64(P, kOsHistRows), kOsThreads);
        const int64_t cap = (int64_t)sm_count() * 4;
        if (blocks > cap) blocks = cap;
        os_hist_kernel<<<(unsigned)blocks, kOsThreads, 0, st>>>(rows, P, padded, plan, ws.ghist);
        MITRE_LAUNCH_CHECK();
        os_offsets_kernel<<<plan.npass, kOsBins, 0, st>>>(ws.ghist);
        MITRE_LAUNCH_CHECK();
    }
    const uint64_t *in = rows;
    uint64_t *out = ws.buf_a;
    int64_t Pin = P;
    for (int p = 0; p < plan.npass; ++p) {
        const int64_t ntiles = ceil_div64(Pin, kOsTile);
        MITRE_CUDA_TRY(cudaMemsetAsync(ws.status, 0, sizeof(uint32_t) * (size_t)ntiles * kOsBins, st));
        const bool first = p == 0, last = p == plan.npass - 1;
        const uint32_t *gofs = ws.ghist + (size_t)p * kOsBins;
        int rc;
#define OS_ARGS st, ntiles, in, Pin, padded, plan.shift[p], plan.bits[p], idx_bits, gofs, ws.status, ws.ticket + p, out, \
                perm, sorted_rows, row_offsets, (uint32_t)(L > 0 ? L : 1), ws.err
        if (first && last) rc = launch_pass<true, true>(OS_ARGS);
        else if (first) rc = launch_pass<true, false>(OS_ARGS);
        else if (last) rc = launch_pass<false, true>(OS_ARGS);
        else rc = launch_pass<false, false>(OS_ARGS);
#undef OS_ARGS
        if (rc != MITRE_OK) return rc;
        in = out;
        out = (out == ws.buf_a) ? ws.buf_b : ws.buf_a;
        Pin = padded ? P_valid : P;
        if (Pin <= 0) break;
    }
    return MITRE_OK;
}

// the look-back's bounded wait expired (never observed; a lost CTA would otherwise hang the device)
int lexsort_packed_failed(cudaStream_t st, void *workspace, int64_t P)
{
    OsWs ws;
    os_layout(P, workspace, &ws);
    int32_t e = 0;
    if (cudaMemcpyAsync(&e, ws.err, sizeof(e), cudaMemcpyDeviceToHost, st) != cudaSuccess) return 1;
    if (cudaStreamSynchronize(st) != cudaSuccess) return 1;
    return e;
}

}  // namespace mitre
