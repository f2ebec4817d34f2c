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
#include <string.h>

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

// chunk_sums[c] = mass of the kDrawChunkTiles tiles of chunk c, fixed summation order; shift = *shift_ptr
// (NULL: 0, the reference's raw exp)
__global__ void __launch_bounds__(kDrawChunkTiles)
draw_chunk_sums_kernel(const double *__restrict__ pairs, int64_t ntiles, const double *__restrict__ shift_ptr,
                       double *__restrict__ chunk_sums)
{
    __shared__ double sred[kDrawChunkTiles];
    const double shift = shift_ptr ? shift_ptr[0] : 0.0;
    const int64_t t = (int64_t)blockIdx.x * kDrawChunkTiles + threadIdx.x;
    sred[threadIdx.x] = t < ntiles ? tile_term(pairs, t, shift) : 0.0;
    __syncthreads();
    for (int s = kDrawChunkTiles / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sred[threadIdx.x] += sred[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) chunk_sums[blockIdx.x] = sred[0];
}

// Shared scratch of select_in_table (one CTA of 256 threads).
struct SelectSmem {
    long long sel;
    double before, marker, total;
    double run[256];
};

// CTA-wide (256 threads): over the `used` leading values of sm.run, the first k whose inclusive prefix
// before + run[0] + ... + run[k] is NOT < marker (the last one if none is), and the prefix preceding it.
// Fixed-shape parallel scan (warp shuffles, then the 8 warp totals): deterministic, and the prefix it
// returns is the one the decision was made with.  total_out (optional): before + everything.
__device__ void select_run(SelectSmem &sm, int used, double before, double marker, double *total_out)
{
    __shared__ double wtot[8];
    __shared__ int s_first[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const double x = (int)threadIdx.x < used ? sm.run[threadIdx.x] : 0.0;
    double inc = x;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += y;
    }
    if (lane == 31) wtot[wid] = inc;
    __syncthreads();
    double base = before;
    for (int w = 0; w < wid; ++w) base += wtot[w];
    const double incl = base + inc;                     // inclusive prefix at this thread
    const double prev = __shfl_up_sync(0xffffffffu, incl, 1);
    const double excl = lane ? prev : base;             // the previous thread's inclusive prefix, bit for bit
    if (total_out && threadIdx.x == 255) *total_out = incl;
    const bool hit = (int)threadIdx.x < used && !(incl < marker);
    const unsigned ballot = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) s_first[wid] = ballot ? (wid * 32 + __ffs(ballot) - 1) : -1;
    __syncthreads();
    int sel = -1;
    for (int w = 0; w < 8 && sel < 0; ++w) sel = s_first[w];
    if (sel < 0) sel = used - 1;
    if ((int)threadIdx.x == sel) {
        sm.sel = sel;
        sm.before = excl;
    }
    __syncthreads();
}

// CTA-wide selection (256 threads, all must call).  The table's candidates carry masses
// scale * exp(w - shift_local) (chunk sums and tile pairs are relative to shift_local; scale rescales
// them to the units of `marker`; in-tile terms are exp(w - shift_rows) with shift_rows = shift_local -
// log(scale), passed explicitly so that no rounding enters).  `before0` = mass preceding this table.
// marker < 0: marker = u * (before0 + this table's mass), the single-table case.
//   (1) narrow [lo, hi) over the chunk sums 256-fold per round (each thread sums one contiguous run,
//       thread 0 walks the 256 run sums left to right);
//   (2) walk the chunk's 256 tile masses left to right;
//   (3) resolve the 128 candidates of that tile with a strict left-to-right cumulative sum, the
//       reference's own order (rules.py:1261-1264: np.cumsum, #{p < marker}).
// Returns the index (valid in every thread).
__device__ int64_t select_in_table(const double *__restrict__ a, const double *__restrict__ b, int64_t P,
                                   double shift_local, double scale, double shift_rows,
                                   const double *__restrict__ pairs, const double *__restrict__ chunk_sums,
                                   double u, double marker_in, double before0, SelectSmem &sm)
{
    const int64_t ntiles = draw_ntiles(P), nchunks = draw_nchunks(P);
    int64_t lo = 0, hi = nchunks;
    double before = before0;
    bool have_marker = false;
    __syncthreads();
    if (threadIdx.x == 0) sm.marker = marker_in;
    while (hi - lo > 1 || !have_marker) {
        const int64_t n = hi - lo;
        const int64_t c = (n + 255) / 256;
        const int64_t i0 = lo + (int64_t)threadIdx.x * c;
        const int64_t i1 = (i0 + c < hi) ? i0 + c : hi;
        double sum = 0.0;
#pragma unroll 4
        for (int64_t i = i0; i < i1; ++i) sum += __ldg(chunk_sums + i);
        sm.run[threadIdx.x] = sum * scale;
        __syncthreads();
        const int used = (int)((n + c - 1) / c);      // runs that hold at least one chunk
        if (!have_marker && marker_in < 0.0) {
            // single table: marker = u * total, total by the same scan (marker = +inf selects nothing)
            select_run(sm, used, before, INFINITY, &sm.total);
            if (threadIdx.x == 0) sm.marker = u * sm.total;
            __syncthreads();
        }
        select_run(sm, used, before, sm.marker, nullptr);
        have_marker = true;
        const int64_t nlo = lo + (int64_t)sm.sel * c;
        hi = (nlo + c < hi) ? nlo + c : hi;
        lo = nlo;
        before = sm.before;
        __syncthreads();
    }
    // the chunk's tiles
    const int64_t t0 = lo * kDrawChunkTiles;
    {
        const int64_t t = t0 + threadIdx.x;
        sm.run[threadIdx.x] = t < ntiles ? tile_term(pairs, t, shift_local) * scale : 0.0;
    }
    __syncthreads();
    {
        const int used = (int)((ntiles - t0 < kDrawChunkTiles) ? (ntiles - t0) : kDrawChunkTiles);
        select_run(sm, used, before, sm.marker, nullptr);
        if (threadIdx.x == 0) sm.sel += t0;
        __syncthreads();
    }
    const int64_t k0 = (int64_t)sm.sel * kDrawTileRows;
    const double tile_before = sm.before;
    __syncthreads();
    if (threadIdx.x < kDrawTileRows) {
        const int64_t k = k0 + threadIdx.x;
        double e = 0.0;
        if (k < P) {
            const double d = weight_at(a, b, k) - shift_rows;
            e = (d == d) ? exp(d) : 0.0;
        }
        sm.run[threadIdx.x] = e;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int n = (int)((P - k0 < kDrawTileRows) ? (P - k0) : kDrawTileRows);
        const double marker = sm.marker;
        double cum = tile_before;
        int count = 0;
        for (int i = 0; i < n; ++i) {
            cum += sm.run[i];
            count += (cum < marker) ? 1 : 0;
        }
        int64_t idx = k0 + count;
        if (idx > P - 1) idx = P - 1;
        sm.sel = idx;
    }
    __syncthreads();
    return sm.sel;
}

__global__ void __launch_bounds__(256)
draw_select_kernel(const double *__restrict__ a, const double *__restrict__ b, int64_t P,
                   const double *__restrict__ shift_ptr, const double *__restrict__ pairs,
                   const double *__restrict__ chunk_sums, double u, int64_t *__restrict__ index_out)
{
    __shared__ SelectSmem sm;
    const double shift = shift_ptr ? shift_ptr[0] : 0.0;
    const int64_t idx = select_in_table(a, b, P, shift, 1.0, shift, pairs, chunk_sums, u, -1.0, 0.0, sm);
    if (threadIdx.x == 0) *index_out = idx;
}

// ---- sharded draw over peer memory (candidate scoring, the scalable exchange) -----------
//
// One problem's lexsorted candidates are cut into contiguous slices, one per GPU (the reference's own
// precedent: _likelihood_worker's [start, stop) slices, logit_rules.py:21-60).  After a rank has
// scored its slice it holds (M_r, S_r) = (max, sum exp(w - max)) of its log-weights.  The draw of ONE
// global index (choose_from_relative_probabilities over all P weights, rules.py:1261-1264) needs
//   (1) every rank's pair on every rank  -> total mass, marker = u * total, the owner of the marker;
//   (2) the owner's selection inside its slice -> the index, on every rank.
// Both exchanges are a few bytes.  Instead of two NCCL collectives (two launches and ~10 us of
// protocol latency each, which is what bounds a 0.25 ms step) the kernel below does them itself over
// NVLink peer memory: every rank STORES its pair straight into every peer's slot buffer (mapped with
// CUDA IPC), spins on its own local slots until all pairs of this step's sequence number have
// arrived, computes the same plan as everyone else, and the owner stores the drawn index into every
// peer's result slot.  One launch per rank, no host involvement, ~2 NVLink store latencies.
// Every wait is bounded (kCommTimeoutNs): a missing peer sets MITRE_STATUS_COMM_TIMEOUT instead of
// hanging the GPU.
struct CommSlot {          // 32 bytes; seq is written last (release), read first (acquire)
    double m, s;
    unsigned long long seq;
    unsigned long long pad;
};
struct CommResult {
    long long index;
    double log_total;
    unsigned long long seq;
    unsigned long long pad;
};
constexpr int kCommMaxWorld = 16;
constexpr unsigned long long kCommTimeoutNs = 2000000000ull;   // 2 s

// The per-rank buffer other ranks write into.  Two copies, used by even / odd sequence numbers: a fast
// rank may publish step seq+1 while a slow one is still reading step seq, but nobody can reach seq+2
// before everyone has published seq+1, i.e. has finished reading seq.
struct CommLayout {
    CommSlot slots[2][kCommMaxWorld];
    CommResult result[2];
};
struct CommView {          // passed by value to the kernel
    int world, rank;
    CommLayout *peer[kCommMaxWorld];    // peer[r] = rank r's buffer in this process's address space
};

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_f64(double *p, double v)
{
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// The plan every rank derives from the gathered pairs (mirrors parallel.owner_of_marker):
// M = max_r M_r, mass_r = S_r exp(M_r - M), marker = u * sum_r mass_r, owner = first rank whose
// cumulative mass is NOT < marker, before = mass of the ranks preceding the owner.
struct DrawPlan {
    double M, total, marker, before;
    int owner;
};
__device__ DrawPlan make_plan(const double *m, const double *s, int world, double u)
{
    DrawPlan p;
    p.M = -INFINITY;
    for (int r = 0; r < world; ++r)
        if (s[r] > 0.0 && m[r] > p.M) p.M = m[r];
    double cum = 0.0;
    p.total = 0.0;
    for (int r = 0; r < world; ++r) p.total += (s[r] > 0.0) ? s[r] * exp(m[r] - p.M) : 0.0;
    p.marker = u * p.total;
    p.owner = -1;
    p.before = 0.0;
    int last_live = -1;
    for (int r = 0; r < world; ++r) {
        const double mass = (s[r] > 0.0) ? s[r] * exp(m[r] - p.M) : 0.0;
        if (mass > 0.0) last_live = r;
        if (p.owner < 0 && mass > 0.0 && !(cum + mass < p.marker)) {
            p.owner = r;
            p.before = cum;
        }
        if (p.owner < 0) cum += mass;
    }
    if (p.owner < 0) {                 // rounding left the marker past the end, or no mass at all
        p.owner = last_live >= 0 ? last_live : 0;
        p.before = 0.0;
        for (int r = 0; r < p.owner; ++r) p.before += (s[r] > 0.0) ? s[r] * exp(m[r] - p.M) : 0.0;
    }
    return p;
}

// out[0] = global index (as double bits of an int64 would not fit: out_index is separate),
// out_stats[0] = M, out_stats[1] = total (log-sum-exp over all ranks = M + log total)
__global__ void __launch_bounds__(256)
sharded_draw_kernel(CommView cv, const double *__restrict__ u_ptr, double *__restrict__ seq_ptr,
                    const double *__restrict__ local_stats,
                    const double *__restrict__ a, const double *__restrict__ b, int64_t P_local,
                    int64_t start, const double *__restrict__ pairs, const double *__restrict__ chunk_sums,
                    int64_t *__restrict__ index_out, double *__restrict__ stats_out, int32_t *status)
{
    __shared__ SelectSmem sm;
    __shared__ double s_m[kCommMaxWorld], s_s[kCommMaxWorld];
    __shared__ DrawPlan plan;
    __shared__ int s_fail;
    const int world = cv.world, rank = cv.rank;
    CommLayout *mine = cv.peer[rank];
    const double u = u_ptr[0];
    const unsigned long long seq = (unsigned long long)seq_ptr[0];
    const int par = (int)(seq & 1);
    if (threadIdx.x == 0) s_fail = 0;
    __syncthreads();
    // (1) publish this rank's pair into every rank's slot[rank], then collect everyone's
    if (threadIdx.x < world) {
        CommSlot *dst = &cv.peer[threadIdx.x]->slots[par][rank];
        st_relaxed_sys_f64(&dst->m, local_stats[0]);
        st_relaxed_sys_f64(&dst->s, local_stats[1]);
        st_release_sys_u64(&dst->seq, seq);
        const CommSlot *src = &mine->slots[par][threadIdx.x];
        const unsigned long long t0 = globaltimer_ns();
        while (ld_acquire_sys_u64(&src->seq) != seq) {
            if (globaltimer_ns() - t0 > kCommTimeoutNs) { s_fail = 1; break; }
        }
        s_m[threadIdx.x] = *reinterpret_cast<const volatile double *>(&src->m);
        s_s[threadIdx.x] = *reinterpret_cast<const volatile double *>(&src->s);
    }
    __syncthreads();
    if (s_fail) {
        if (threadIdx.x == 0) {
            if (status) atomicOr(reinterpret_cast<unsigned int *>(status), MITRE_STATUS_COMM_TIMEOUT);
            *index_out = -1;
        }
        return;
    }
    if (threadIdx.x == 0) plan = make_plan(s_m, s_s, world, u);
    __syncthreads();
    // (2) the owner selects inside its slice and publishes the index
    if (plan.owner == rank) {
        const double Mr = s_m[rank];
        const double scale = exp(Mr - plan.M);
        const int64_t local = P_local > 0 ? select_in_table(a, b, P_local, Mr, scale, plan.M, pairs, chunk_sums, u,
                                                            plan.marker, plan.before, sm)
                                          : 0;
        if (threadIdx.x < world) {
            CommResult *dst = &cv.peer[threadIdx.x]->result[par];
            st_relaxed_sys_u64(reinterpret_cast<unsigned long long *>(&dst->index), (unsigned long long)(start + local));
            st_relaxed_sys_f64(&dst->log_total, plan.M + log(plan.total));
            st_release_sys_u64(&dst->seq, seq);
        }
    }
    if (threadIdx.x == 0) {
        const unsigned long long t0 = globaltimer_ns();
        bool ok = true;
        while (ld_acquire_sys_u64(&mine->result[par].seq) != seq) {
            if (globaltimer_ns() - t0 > kCommTimeoutNs) { ok = false; break; }
        }
        if (ok) {
            *index_out = *reinterpret_cast<const volatile long long *>(&mine->result[par].index);
            stats_out[0] = plan.M;
            stats_out[1] = plan.total;
            seq_ptr[0] = (double)(seq + 1);      // the next launch is the next step, on every rank
        } else {
            if (status) atomicOr(reinterpret_cast<unsigned int *>(status), MITRE_STATUS_COMM_TIMEOUT);
            *index_out = -1;
        }
    }
}

// The same step with the exchanges done by the caller (NCCL all-gather of the pairs before, all-reduce
// MAX of index_out after): gathered = float64[world, 2].  Non-owners write index -1.
__global__ void __launch_bounds__(256)
sharded_select_kernel(int world, int rank, const double *__restrict__ u_ptr, const double *__restrict__ gathered,
                      const double *__restrict__ a, const double *__restrict__ b, int64_t P_local, int64_t start,
                      const double *__restrict__ pairs, const double *__restrict__ chunk_sums,
                      int64_t *__restrict__ index_out, double *__restrict__ stats_out)
{
    __shared__ SelectSmem sm;
    __shared__ double s_m[kCommMaxWorld], s_s[kCommMaxWorld];
    __shared__ DrawPlan plan;
    const double u = u_ptr[0];
    if (threadIdx.x < world) {
        s_m[threadIdx.x] = gathered[2 * threadIdx.x];
        s_s[threadIdx.x] = gathered[2 * threadIdx.x + 1];
    }
    __syncthreads();
    if (threadIdx.x == 0) plan = make_plan(s_m, s_s, world, u);
    __syncthreads();
    int64_t idx = -1;
    if (plan.owner == rank) {
        const double Mr = s_m[rank];
        const double scale = exp(Mr - plan.M);
        idx = start + (P_local > 0 ? select_in_table(a, b, P_local, Mr, scale, plan.M, pairs, chunk_sums, u,
                                                     plan.marker, plan.before, sm)
                                   : 0);
    }
    if (threadIdx.x == 0) {
        *index_out = idx;
        stats_out[0] = plan.M;
        stats_out[1] = plan.total;
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
    const double *shift_ptr = stabilise ? stats : nullptr;
    draw_chunk_sums_kernel<<<(unsigned)nchunks, kDrawChunkTiles, 0, st>>>(pairs, draw_ntiles(P), shift_ptr, chunk_sums);
    MITRE_LAUNCH_CHECK();
    draw_select_kernel<<<1, 256, 0, st>>>(a, b, P, shift_ptr, pairs, chunk_sums, u, index_out);
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

// ---- sharded draw: host side ----------------------------------------------------------------------------
namespace {
struct Comm {
    int world = 0, rank = 0, device = 0;
    CommLayout *local = nullptr;                  // cudaMalloc'ed here, written by the peers
    CommLayout *peer[kCommMaxWorld] = {nullptr};  // peer[rank] == local
    bool opened[kCommMaxWorld] = {false};
    unsigned long long seq = 0;
};
}  // namespace

extern "C" int mitre_comm_create(int world, int rank, void **comm_out, void *handle_out)
{
    if (world < 1 || world > kCommMaxWorld || rank < 0 || rank >= world || !comm_out || !handle_out)
        return MITRE_ERR_INVALID;
    Comm *c = new Comm();
    c->world = world;
    c->rank = rank;
    MITRE_CUDA_TRY(cudaGetDevice(&c->device));
    cudaError_t e = cudaMalloc(&c->local, sizeof(CommLayout));
    if (e != cudaSuccess) { delete c; return (int)e; }
    e = cudaMemset(c->local, 0, sizeof(CommLayout));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    memset(&h, 0, sizeof(h));
    if (e == cudaSuccess && world > 1) e = cudaIpcGetMemHandle(&h, c->local);
    if (e != cudaSuccess) { cudaFree(c->local); delete c; return (int)e; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    memcpy(handle_out, &h, 64);
    c->peer[rank] = c->local;
    *comm_out = c;
    return MITRE_OK;
}

extern "C" int mitre_comm_connect(void *comm, const void *all_handles)
{
    Comm *c = static_cast<Comm *>(comm);
    if (!c || (c->world > 1 && !all_handles)) return MITRE_ERR_INVALID;
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(all_handles) + 64 * (size_t)r, 64);
        void *p = nullptr;
        MITRE_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer[r] = static_cast<CommLayout *>(p);
        c->opened[r] = true;
    }
    return MITRE_OK;
}

extern "C" int mitre_comm_destroy(void *comm)
{
    Comm *c = static_cast<Comm *>(comm);
    if (!c) return MITRE_OK;
    for (int r = 0; r < c->world; ++r)
        if (c->opened[r]) cudaIpcCloseMemHandle(c->peer[r]);
    cudaFree(c->local);
    delete c;
    return MITRE_OK;
}

// chunk sums of the slice's tile pairs relative to the slice's own max (local_stats[0]); the first half of
// a sharded draw, independent of the other ranks
extern "C" int mitre_sharded_chunk_sums(void *stream, int64_t P_local, const double *local_stats, void *workspace,
                                        size_t workspace_bytes)
{
    if (P_local < 0 || !local_stats || !workspace) return MITRE_ERR_INVALID;
    if (workspace_bytes < mitre_draw_workspace_bytes(P_local)) return MITRE_ERR_WORKSPACE;
    if (P_local == 0) return MITRE_OK;
    double *ws = static_cast<double *>(workspace);
    draw_chunk_sums_kernel<<<(unsigned)draw_nchunks(P_local), kDrawChunkTiles, 0, as_stream(stream)>>>(
        ws + draw_off_pairs(), draw_ntiles(P_local), local_stats, ws + draw_off_chunks(P_local));
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

extern "C" int mitre_sharded_draw(void *stream, void *comm, const double *u, double *seq, const double *local_stats,
                                  const double *a, const double *b, int64_t P_local, int64_t start,
                                  void *workspace, size_t workspace_bytes, int64_t *index_out, double *stats_out,
                                  int32_t *status)
{
    Comm *c = static_cast<Comm *>(comm);
    if (!c || !u || !seq || !local_stats || !index_out || !stats_out || P_local < 0) return MITRE_ERR_INVALID;
    if (P_local > 0 && (!a || !workspace || workspace_bytes < mitre_draw_workspace_bytes(P_local)))
        return MITRE_ERR_WORKSPACE;
    CommView cv;
    cv.world = c->world;
    cv.rank = c->rank;
    for (int r = 0; r < kCommMaxWorld; ++r) cv.peer[r] = r < c->world ? c->peer[r] : nullptr;
    const double *ws = static_cast<const double *>(workspace);
    sharded_draw_kernel<<<1, 256, 0, as_stream(stream)>>>(cv, u, seq, local_stats, a, b, P_local, start,
                                                          ws ? ws + draw_off_pairs() : nullptr,
                                                          ws ? ws + draw_off_chunks(P_local) : nullptr, index_out,
                                                          stats_out, status);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

extern "C" int mitre_sharded_select(void *stream, int world, int rank, const double *u, const double *gathered,
                                    const double *a, const double *b, int64_t P_local, int64_t start,
                                    void *workspace, size_t workspace_bytes, int64_t *index_out, double *stats_out)
{
    if (world < 1 || world > kCommMaxWorld || rank < 0 || rank >= world || !u || !gathered || !index_out || !stats_out ||
        P_local < 0)
        return MITRE_ERR_INVALID;
    if (P_local > 0 && (!a || !workspace || workspace_bytes < mitre_draw_workspace_bytes(P_local)))
        return MITRE_ERR_WORKSPACE;
    const double *ws = static_cast<const double *>(workspace);
    sharded_select_kernel<<<1, 256, 0, as_stream(stream)>>>(world, rank, u, gathered, a, b, P_local, start,
                                                            ws ? ws + draw_off_pairs() : nullptr,
                                                            ws ? ws + draw_off_chunks(P_local) : nullptr, index_out,
                                                            stats_out);
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
