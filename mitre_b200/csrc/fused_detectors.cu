// fused_detectors.cu -- C ABI entry points of the fused detector path (kernel: fused_detectors.cuh,
// instantiated per sorting-network size in fused_inst.cu).
#define MITRE_FUSED_MAIN 1
#include <vector>

#include "fused_detectors.cuh"

namespace mitre {
int g_detector_split = 1;
int g_sort_emit_rolled = 0;
SlopeGuardArgs g_slope_guard = {nullptr, nullptr, nullptr};
int g_emit_blocks_per_sm = 0;
int launch_slope_guard(cudaStream_t st, const double *values, int64_t G, int N, const uint8_t *group_type,
                       uint8_t *group_flags, unsigned long long *flagged);   // detectors.cu
int launch_count_flags(cudaStream_t st, const uint8_t *group_flags, int64_t G, unsigned long long *flagged);
int launch_fused_nb4(cudaStream_t st, const double *series_t, int64_t vstride, int N, int V,
                     const int32_t *wlo, const int32_t *wcnt, const int32_t *wstart,
                     const double *winv_n, const double *winv_sxx, const int32_t *widx,
                     const double *wdt, int64_t ld, const int64_t *group_base,
                     const uint8_t *slope_ok, const int32_t *task_window, const int8_t *task_type,
                     int n_wt, int64_t G, double *values, double *thresholds, int32_t *K, uint64_t *rows);
int launch_fused_nb8(cudaStream_t st, const double *series_t, int64_t vstride, int N, int V,
                     const int32_t *wlo, const int32_t *wcnt, const int32_t *wstart,
                     const double *winv_n, const double *winv_sxx, const int32_t *widx,
                     const double *wdt, int64_t ld, const int64_t *group_base,
                     const uint8_t *slope_ok, const int32_t *task_window, const int8_t *task_type,
                     int n_wt, int64_t G, double *values, double *thresholds, int32_t *K, uint64_t *rows);
int launch_fused_nb12(cudaStream_t st, const double *series_t, int64_t vstride, int N, int V,
                     const int32_t *wlo, const int32_t *wcnt, const int32_t *wstart,
                     const double *winv_n, const double *winv_sxx, const int32_t *widx,
                     const double *wdt, int64_t ld, const int64_t *group_base,
                     const uint8_t *slope_ok, const int32_t *task_window, const int8_t *task_type,
                     int n_wt, int64_t G, double *values, double *thresholds, int32_t *K, uint64_t *rows);
int launch_fused_nb16(cudaStream_t st, const double *series_t, int64_t vstride, int N, int V,
                     const int32_t *wlo, const int32_t *wcnt, const int32_t *wstart,
                     const double *winv_n, const double *winv_sxx, const int32_t *widx,
                     const double *wdt, int64_t ld, const int64_t *group_base,
                     const uint8_t *slope_ok, const int32_t *task_window, const int8_t *task_type,
                     int n_wt, int64_t G, double *values, double *thresholds, int32_t *K, uint64_t *rows);
int launch_fused_nb20(cudaStream_t st, const double *series_t, int64_t vstride, int N, int V,
                     const int32_t *wlo, const int32_t *wcnt, const int32_t *wstart,
                     const double *winv_n, const double *winv_sxx, const int32_t *widx,
                     const double *wdt, int64_t ld, const int64_t *group_base,
                     const uint8_t *slope_ok, const int32_t *task_window, const int8_t *task_type,
                     int n_wt, int64_t G, double *values, double *thresholds, int32_t *K, uint64_t *rows);
int launch_fused_nb24(cudaStream_t st, const double *series_t, int64_t vstride, int N, int V,
                     const int32_t *wlo, const int32_t *wcnt, const int32_t *wstart,
                     const double *winv_n, const double *winv_sxx, const int32_t *widx,
                     const double *wdt, int64_t ld, const int64_t *group_base,
                     const uint8_t *slope_ok, const int32_t *task_window, const int8_t *task_type,
                     int n_wt, int64_t G, double *values, double *thresholds, int32_t *K, uint64_t *rows);
int launch_fused_nb28(cudaStream_t st, const double *series_t, int64_t vstride, int N, int V,
                     const int32_t *wlo, const int32_t *wcnt, const int32_t *wstart,
                     const double *winv_n, const double *winv_sxx, const int32_t *widx,
                     const double *wdt, int64_t ld, const int64_t *group_base,
                     const uint8_t *slope_ok, const int32_t *task_window, const int8_t *task_type,
                     int n_wt, int64_t G, double *values, double *thresholds, int32_t *K, uint64_t *rows);
int launch_fused_nb32(cudaStream_t st, const double *series_t, int64_t vstride, int N, int V,
                     const int32_t *wlo, const int32_t *wcnt, const int32_t *wstart,
                     const double *winv_n, const double *winv_sxx, const int32_t *widx,
                     const double *wdt, int64_t ld, const int64_t *group_base,
                     const uint8_t *slope_ok, const int32_t *task_window, const int8_t *task_type,
                     int n_wt, int64_t G, double *values, double *thresholds, int32_t *K, uint64_t *rows);
int launch_fused_nb36(cudaStream_t st, const double *series_t, int64_t vstride, int N, int V,
                     const int32_t *wlo, const int32_t *wcnt, const int32_t *wstart,
                     const double *winv_n, const double *winv_sxx, const int32_t *widx,
                     const double *wdt, int64_t ld, const int64_t *group_base,
                     const uint8_t *slope_ok, const int32_t *task_window, const int8_t *task_type,
                     int n_wt, int64_t G, double *values, double *thresholds, int32_t *K, uint64_t *rows);
int launch_fused_nb40(cudaStream_t st, const double *series_t, int64_t vstride, int N, int V,
                     const int32_t *wlo, const int32_t *wcnt, const int32_t *wstart,
                     const double *winv_n, const double *winv_sxx, const int32_t *widx,
                     const double *wdt, int64_t ld, const int64_t *group_base,
                     const uint8_t *slope_ok, const int32_t *task_window, const int8_t *task_type,
                     int n_wt, int64_t G, double *values, double *thresholds, int32_t *K, uint64_t *rows);
int launch_fused_nb44(cudaStream_t st, const double *series_t, int64_t vstride, int N, int V,
                     const int32_t *wlo, const int32_t *wcnt, const int32_t *wstart,
                     const double *winv_n, const double *winv_sxx, const int32_t *widx,
                     const double *wdt, int64_t ld, const int64_t *group_base,
                     const uint8_t *slope_ok, const int32_t *task_window, const int8_t *task_type,
                     int n_wt, int64_t G, double *values, double *thresholds, int32_t *K, uint64_t *rows);
int launch_fused_nb48(cudaStream_t st, const double *series_t, int64_t vstride, int N, int V,
                     const int32_t *wlo, const int32_t *wcnt, const int32_t *wstart,
                     const double *winv_n, const double *winv_sxx, const int32_t *widx,
                     const double *wdt, int64_t ld, const int64_t *group_base,
                     const uint8_t *slope_ok, const int32_t *task_window, const int8_t *task_type,
                     int n_wt, int64_t G, double *values, double *thresholds, int32_t *K, uint64_t *rows);
int launch_fused_nb52(cudaStream_t st, const double *series_t, int64_t vstride, int N, int V,
                     const int32_t *wlo, const int32_t *wcnt, const int32_t *wstart,
                     const double *winv_n, const double *winv_sxx, const int32_t *widx,
                     const double *wdt, int64_t ld, const int64_t *group_base,
                     const uint8_t *slope_ok, const int32_t *task_window, const int8_t *task_type,
                     int n_wt, int64_t G, double *values, double *thresholds, int32_t *K, uint64_t *rows);
int launch_fused_nb56(cudaStream_t st, const double *series_t, int64_t vstride, int N, int V,
                     const int32_t *wlo, const int32_t *wcnt, const int32_t *wstart,
                     const double *winv_n, const double *winv_sxx, const int32_t *widx,
                     const double *wdt, int64_t ld, const int64_t *group_base,
                     const uint8_t *slope_ok, const int32_t *task_window, const int8_t *task_type,
                     int n_wt, int64_t G, double *values, double *thresholds, int32_t *K, uint64_t *rows);
int launch_fused_nb60(cudaStream_t st, const double *series_t, int64_t vstride, int N, int V,
                     const int32_t *wlo, const int32_t *wcnt, const int32_t *wstart,
                     const double *winv_n, const double *winv_sxx, const int32_t *widx,
                     const double *wdt, int64_t ld, const int64_t *group_base,
                     const uint8_t *slope_ok, const int32_t *task_window, const int8_t *task_type,
                     int n_wt, int64_t G, double *values, double *thresholds, int32_t *K, uint64_t *rows);
int launch_fused_nb64(cudaStream_t st, const double *series_t, int64_t vstride, int N, int V,
                     const int32_t *wlo, const int32_t *wcnt, const int32_t *wstart,
                     const double *winv_n, const double *winv_sxx, const int32_t *widx,
                     const double *wdt, int64_t ld, const int64_t *group_base,
                     const uint8_t *slope_ok, const int32_t *task_window, const int8_t *task_type,
                     int n_wt, int64_t G, double *values, double *thresholds, int32_t *K, uint64_t *rows);

// padded slot index -> dense enumeration index (the reference's _reordering_cache values)
__global__ void __launch_bounds__(256)
densify_perm_kernel(const int64_t *__restrict__ perm_pad, int64_t P, int L,
                    const int64_t *__restrict__ row_offsets, int64_t *__restrict__ perm)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += stride) {
        const int64_t r = perm_pad[i];
        const int64_t g = r / L;
        perm[i] = row_offsets[g] + (r - g * L);
    }
}
}  // namespace mitre

using namespace mitre;

extern "C" int mitre_window_lists(void *stream, const double *times, const int32_t *t_offsets, int N,
                                  const double *windows, int W, int64_t ld, int64_t vstride, int32_t *lo,
                                  int32_t *cnt, int32_t *start, double *inv_n, double *inv_sxx, int32_t *idx,
                                  double *dt)
{
    if (N <= 0 || W < 0 || ld <= 0 || vstride <= 0 || !times || !t_offsets) return MITRE_ERR_INVALID;
    if (ld * vstride >= 0x7fffffffll) return MITRE_ERR_UNSUPPORTED;   // int32 element offsets
    if (W == 0) return MITRE_OK;
    if (!windows || !lo || !cnt || !start || !inv_n || !inv_sxx || !idx || !dt) return MITRE_ERR_INVALID;
    window_lists_kernel<<<(W + 127) / 128, 128, 0, as_stream(stream)>>>(times, t_offsets, N, windows, W, ld,
                                                                        vstride, lo, cnt, start, inv_n, inv_sxx,
                                                                        idx, dt);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

static int detectors_fused_launch(void *stream, const double *series_t, int64_t vstride, int N, int V,
                                  const int32_t *wlo, const int32_t *wcnt, const int32_t *wstart,
                                  const double *winv_n, const double *winv_sxx, const int32_t *widx,
                                  const double *wdt, int64_t ld, const int64_t *group_base,
                                  const uint8_t *slope_ok, const int32_t *task_window,
                                  const int8_t *task_type, int n_tasks_wt, int n_windows, int64_t G,
                                  double *values, double *thresholds, int32_t *K, uint64_t *rows_padded);

// sorting-network size -> translation unit
static int nb_dispatch(cudaStream_t st, const double *series_t, int64_t vstride, int N, int V, const int32_t *wlo,
                       const int32_t *wcnt, const int32_t *wstart, const double *winv_n, const double *winv_sxx,
                       const int32_t *widx, const double *wdt, int64_t ld, const int64_t *group_base,
                       const uint8_t *slope_ok, const int32_t *task_window, const int8_t *task_type, int n_tasks_wt,
                       int64_t G, double *values, double *thresholds, int32_t *K, uint64_t *rows_padded)
{
    if (N <= 4) return launch_fused_nb4(st, series_t, vstride, N, V, wlo, wcnt, wstart, winv_n, winv_sxx, widx, wdt, ld, group_base, slope_ok, task_window, task_type, n_tasks_wt, G, values, thresholds, K, rows_padded);
    if (N <= 8) return launch_fused_nb8(st, series_t, vstride, N, V, wlo, wcnt, wstart, winv_n, winv_sxx, widx, wdt, ld, group_base, slope_ok, task_window, task_type, n_tasks_wt, G, values, thresholds, K, rows_padded);
    if (N <= 12) return launch_fused_nb12(st, series_t, vstride, N, V, wlo, wcnt, wstart, winv_n, winv_sxx, widx, wdt, ld, group_base, slope_ok, task_window, task_type, n_tasks_wt, G, values, thresholds, K, rows_padded);
    if (N <= 16) return launch_fused_nb16(st, series_t, vstride, N, V, wlo, wcnt, wstart, winv_n, winv_sxx, widx, wdt, ld, group_base, slope_ok, task_window, task_type, n_tasks_wt, G, values, thresholds, K, rows_padded);
    if (N <= 20) return launch_fused_nb20(st, series_t, vstride, N, V, wlo, wcnt, wstart, winv_n, winv_sxx, widx, wdt, ld, group_base, slope_ok, task_window, task_type, n_tasks_wt, G, values, thresholds, K, rows_padded);
    if (N <= 24) return launch_fused_nb24(st, series_t, vstride, N, V, wlo, wcnt, wstart, winv_n, winv_sxx, widx, wdt, ld, group_base, slope_ok, task_window, task_type, n_tasks_wt, G, values, thresholds, K, rows_padded);
    if (N <= 28) return launch_fused_nb28(st, series_t, vstride, N, V, wlo, wcnt, wstart, winv_n, winv_sxx, widx, wdt, ld, group_base, slope_ok, task_window, task_type, n_tasks_wt, G, values, thresholds, K, rows_padded);
    if (N <= 32) return launch_fused_nb32(st, series_t, vstride, N, V, wlo, wcnt, wstart, winv_n, winv_sxx, widx, wdt, ld, group_base, slope_ok, task_window, task_type, n_tasks_wt, G, values, thresholds, K, rows_padded);
    if (N <= 36) return launch_fused_nb36(st, series_t, vstride, N, V, wlo, wcnt, wstart, winv_n, winv_sxx, widx, wdt, ld, group_base, slope_ok, task_window, task_type, n_tasks_wt, G, values, thresholds, K, rows_padded);
    if (N <= 40) return launch_fused_nb40(st, series_t, vstride, N, V, wlo, wcnt, wstart, winv_n, winv_sxx, widx, wdt, ld, group_base, slope_ok, task_window, task_type, n_tasks_wt, G, values, thresholds, K, rows_padded);
    if (N <= 44) return launch_fused_nb44(st, series_t, vstride, N, V, wlo, wcnt, wstart, winv_n, winv_sxx, widx, wdt, ld, group_base, slope_ok, task_window, task_type, n_tasks_wt, G, values, thresholds, K, rows_padded);
    if (N <= 48) return launch_fused_nb48(st, series_t, vstride, N, V, wlo, wcnt, wstart, winv_n, winv_sxx, widx, wdt, ld, group_base, slope_ok, task_window, task_type, n_tasks_wt, G, values, thresholds, K, rows_padded);
    if (N <= 52) return launch_fused_nb52(st, series_t, vstride, N, V, wlo, wcnt, wstart, winv_n, winv_sxx, widx, wdt, ld, group_base, slope_ok, task_window, task_type, n_tasks_wt, G, values, thresholds, K, rows_padded);
    if (N <= 56) return launch_fused_nb56(st, series_t, vstride, N, V, wlo, wcnt, wstart, winv_n, winv_sxx, widx, wdt, ld, group_base, slope_ok, task_window, task_type, n_tasks_wt, G, values, thresholds, K, rows_padded);
    if (N <= 60) return launch_fused_nb60(st, series_t, vstride, N, V, wlo, wcnt, wstart, winv_n, winv_sxx, widx, wdt, ld, group_base, slope_ok, task_window, task_type, n_tasks_wt, G, values, thresholds, K, rows_padded);
    if (N <= 64) return launch_fused_nb64(st, series_t, vstride, N, V, wlo, wcnt, wstart, winv_n, winv_sxx, widx, wdt, ld, group_base, slope_ok, task_window, task_type, n_tasks_wt, G, values, thresholds, K, rows_padded);
    return MITRE_ERR_UNSUPPORTED;
}


extern "C" int mitre_detectors_fused(void *stream, const double *series_t, int64_t vstride, int N, int V,
                                     const int32_t *wlo, const int32_t *wcnt, const int32_t *wstart,
                                     const double *winv_n, const double *winv_sxx, const int32_t *widx,
                                     const double *wdt, int64_t ld, const int64_t *group_base,
                                     const uint8_t *slope_ok, const int32_t *task_window,
                                     const int8_t *task_type, int n_tasks_wt, int n_windows, int64_t G,
                                     double *values, double *thresholds, int32_t *K, uint64_t *rows_padded,
                                     const uint8_t *group_type, uint8_t *group_flags, int64_t *flagged)
{
    if (N < 2 || N > 63) return MITRE_ERR_UNSUPPORTED;
    if ((group_flags || flagged) && !group_type) return MITRE_ERR_INVALID;
    if (flagged && !group_flags) return MITRE_ERR_INVALID;
    // Slope guard band.  Default variants: the values kernel sees a subject's mean and slope together and
    // flags constant non-zero series, the staged sort/emit kernel checks near-ties on the sorted values it
    // already holds -- no extra pass.  The cross-check variants (single kernel, flat values, direct
    // sort/emit) get the separate pass.
    const bool want_guard = group_type && group_flags && G > 0;
    const bool in_kernel = want_guard && (g_detector_split == 1 || g_detector_split == 3);
    if (want_guard) MITRE_CUDA_TRY(cudaMemsetAsync(group_flags, 0, (size_t)G, as_stream(stream)));
    if (flagged) MITRE_CUDA_TRY(cudaMemsetAsync(flagged, 0, sizeof(int64_t), as_stream(stream)));
    struct GuardScope {
        GuardScope(bool on, const uint8_t *t, uint8_t *f)
        {
            g_slope_guard = on ? SlopeGuardArgs{t, f, nullptr} : SlopeGuardArgs{nullptr, nullptr, nullptr};
        }
        ~GuardScope() { g_slope_guard = SlopeGuardArgs{nullptr, nullptr, nullptr}; }
    } guard_scope(in_kernel, group_type, group_flags);
    int rc = detectors_fused_launch(stream, series_t, vstride, N, V, wlo, wcnt, wstart, winv_n, winv_sxx, widx, wdt, ld,
                                    group_base, slope_ok, task_window, task_type, n_tasks_wt, n_windows, G, values,
                                    thresholds, K, rows_padded);
    if (rc != MITRE_OK) return rc;
    if (want_guard && g_detector_split != 5) {
        if (!in_kernel)
            return launch_slope_guard(as_stream(stream), values, G, N, group_type, group_flags,
                                      reinterpret_cast<unsigned long long *>(flagged));
        if (flagged) return launch_count_flags(as_stream(stream), group_flags, G, reinterpret_cast<unsigned long long *>(flagged));
    }
    return MITRE_OK;
}

static int detectors_fused_launch(void *stream, const double *series_t, int64_t vstride, int N, int V,
                                  const int32_t *wlo, const int32_t *wcnt, const int32_t *wstart,
                                  const double *winv_n, const double *winv_sxx, const int32_t *widx,
                                  const double *wdt, int64_t ld, const int64_t *group_base,
                                  const uint8_t *slope_ok, const int32_t *task_window,
                                  const int8_t *task_type, int n_tasks_wt, int n_windows, int64_t G,
                                  double *values, double *thresholds, int32_t *K, uint64_t *rows_padded)
{
    if (V <= 0 || n_tasks_wt < 0 || vstride < V || ld <= 0) return MITRE_ERR_INVALID;
    if (n_tasks_wt == 0) return MITRE_OK;
    if (!series_t || !wlo || !wcnt || !wstart || !winv_n || !winv_sxx || !widx || !wdt || !group_base ||
        !slope_ok || !task_window || !task_type || !values || !thresholds || !K || !rows_padded)
        return MITRE_ERR_INVALID;
    if (g_detector_split) {
        if (G <= 0) return MITRE_ERR_INVALID;
        if (n_windows <= 0) return MITRE_ERR_INVALID;
        if (g_detector_split == 2) {
            const int64_t bx = (int64_t)n_tasks_wt * N;
            if (bx > 0x7fffffffll) return MITRE_ERR_UNSUPPORTED;
            int by = (V + 2047) / 2048;   // up to 8 variables per thread along y
            if (by > 65535) by = 65535;
            values_flat_kernel<<<dim3((unsigned)bx, (unsigned)by), 256, 0, as_stream(stream)>>>(
                series_t, vstride, N, V, wlo, wcnt, wstart, winv_n, winv_sxx, wdt, ld, group_base, slope_ok,
                task_window, task_type, values);
        } else if (g_detector_split != 3 &&
                   (size_t)ld * 256 + sizeof(double) * 128 * (size_t)N + 2 * resident_meta_bytes(N, ld) <=
                       smem_optin()) {
            // the whole sample axis of a 32-variable tile fits in shared memory
            const size_t smem = (size_t)ld * 256 + sizeof(double) * 128 * (size_t)N +
                                2 * resident_meta_bytes(N, ld);
            const int n_warps = N < 16 ? N : 16;                // warps take subjects dynamically
            MITRE_CUDA_TRY(cudaFuncSetAttribute(values_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)smem));
            // split the windows over blockIdx.y so that the grid fills whole waves of one block per SM
            const int tiles = (V + 31) / 32, sms = sm_count();
            int chunks = 1;
            double best = 0.0;
            for (int k = 1; k <= 16 && n_windows / k >= 8; ++k) {
                const int64_t blocks = (int64_t)tiles * k;
                const double fill = (double)blocks / (double)(ceil_div64(blocks, sms) * sms) - 0.004 * k;
                if (fill > best) { best = fill; chunks = k; }
            }
            const int per_block = (n_windows + chunks - 1) / chunks;
            chunks = (n_windows + per_block - 1) / per_block;
            values_resident_kernel<<<dim3((unsigned)tiles, (unsigned)chunks), 32 * n_warps, smem, as_stream(stream)>>>(
                series_t, vstride, N, V, n_windows, wlo, wcnt, wstart, winv_n, winv_sxx, wdt, ld, group_base,
                slope_ok, per_block, values, g_slope_guard.group_flags);
        } else {
            const int n_vtiles = (V + 31) / 32;
            const int64_t blocks = (int64_t)n_windows * n_vtiles;
            if (blocks > 0x7fffffffll) return MITRE_ERR_UNSUPPORTED;
            const int rounds = (N + 7) / 8;                     // subjects per warp
            const int n_warps = (N + rounds - 1) / rounds;      // <= 8, so that the warps finish together
            const size_t smem = sizeof(double) * 64 * (size_t)(N | 1);
            values_tile_kernel<<<(unsigned)blocks, 32 * n_warps, smem, as_stream(stream)>>>(
                series_t, vstride, N, V, n_vtiles, wlo, wcnt, wstart, winv_n, winv_sxx, wdt, ld, group_base,
                slope_ok, values, g_slope_guard.group_flags);
        }
        MITRE_LAUNCH_CHECK();
    }
    return nb_dispatch(as_stream(stream), series_t, vstride, N, V, wlo, wcnt, wstart, winv_n, winv_sxx, widx, wdt, ld,
                       group_base, slope_ok, task_window, task_type, n_tasks_wt, G, values, thresholds, K, rows_padded);
}

// ---- walk-forward values + stand-alone sort/emit (the default device path from round 2) --------------------
extern "C" int mitre_window_tbar(void *stream, const double *times, int N, const double *windows, int W,
                                 const int32_t *wlo, const int32_t *wcnt, double *tbar)
{
    if (N <= 0 || W < 0 || !times) return MITRE_ERR_INVALID;
    if (W == 0) return MITRE_OK;
    if (!windows || !wlo || !wcnt || !tbar) return MITRE_ERR_INVALID;
    window_tbar_kernel<<<(W * N + 127) / 128, 128, 0, as_stream(stream)>>>(times, N, windows, W, wlo, wcnt, tbar);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

// The block's start groups are chosen so that the (window, subject) tables of ALL its windows fit in shared
// memory next to the series tile: sg_first is a host-visible copy is not available here, so the launcher
// takes the window count of the fullest chunk from the caller (max_block_windows) for a given split.
static int g_walk_vpt = 1;   // variables per thread of values_walk_kernel
static int g_walk_persistent = 0;    // 1: values_walk_kernel CTAs loop over variable tiles, tables loaded once (measured slower)
static int g_walk_vb_cap = 16;       // variables per walk CTA (upper bound)
static int g_pipe_emit_per_sm = 0;   // pipelined call: sort/emit blocks per SM while a values kernel runs beside it (0 = all)
static int g_pipe_priority = 0;      // pipelined call: sort/emit stream at high priority

extern "C" int mitre_set_walk_tuning(int vb_cap, int emit_blocks_per_sm, int emit_priority)
{
    g_walk_vb_cap = vb_cap < 1 ? 1 : (vb_cap > 16 ? 16 : vb_cap);
    g_pipe_emit_per_sm = emit_blocks_per_sm < 0 ? 0 : emit_blocks_per_sm;
    g_pipe_priority = emit_priority & 1;
    g_walk_persistent = (emit_priority & 2) ? 1 : 0;      // bit 1: CTAs persistent over the variable tiles
    return MITRE_OK;
}

extern "C" int mitre_walk_vars_per_block(int N) { return N > 0 ? walk_vars_per_block(N, g_walk_vb_cap) : 0; }

extern "C" int mitre_set_walk_vpt(int vpt)
{
    const int old = g_walk_vpt;
    g_walk_vpt = vpt == 1 ? 1 : 2;
    return old;
}

static int launch_values_walk(cudaStream_t st, dim3 grid, size_t smem, const double *series, int64_t series_stride,
                              const double *times, const int32_t *t_offsets, int S, int N, int V, const double *windows,
                              const int32_t *wlo, const int32_t *wcnt, const double *winv_sxx, const double *wtbar,
                              const int64_t *group_base, const uint8_t *slope_ok, const int32_t *sg_first, int sg_base,
                              int n_sg, int sg_per_block, int max_block_windows, double *values, uint8_t *group_flags)
{
    const int VB = walk_vars_per_block(N, g_walk_vb_cap);
    const int vpt = (g_walk_vpt == 2 && VB >= 2) ? 2 : 1;
    const int threads = ((((VB + vpt - 1) / vpt) * N + 31) / 32) * 32;
    if (vpt == 2) {
        MITRE_CUDA_TRY(cudaFuncSetAttribute(values_walkn_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        values_walkn_kernel<2><<<grid, threads, smem, st>>>(series, series_stride, times, t_offsets, S, N, V, windows, wlo,
                                                           wcnt, winv_sxx, wtbar, group_base, slope_ok, sg_first, sg_base,
                                                           n_sg, sg_per_block, max_block_windows, values, group_flags, VB);
    } else {
        MITRE_CUDA_TRY(cudaFuncSetAttribute(values_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // persistent over the variable tiles: one CTA per SM (shared memory), the SMs dealt to the chunks
        if (g_walk_persistent && grid.y > 0) {
            unsigned per_chunk = (unsigned)sm_count() / grid.y;
            if (per_chunk < 1) per_chunk = 1;
            if (per_chunk < grid.x) grid.x = per_chunk;
        }
        values_walk_kernel<<<grid, threads, smem, st>>>(series, series_stride, times, t_offsets, S, N, V, windows, wlo,
                                                           wcnt, winv_sxx, wtbar, group_base, slope_ok, sg_first, sg_base,
                                                           n_sg, sg_per_block, max_block_windows, values, group_flags, VB);
    }
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

extern "C" size_t mitre_values_walk_smem_bytes(int N, int S, int max_block_windows)
{
    if (N <= 0 || S <= 0 || max_block_windows <= 0) return 0;
    return sizeof(double) * walk_smem_doubles(N, S, max_block_windows, walk_vars_per_block(N, g_walk_vb_cap));
}

extern "C" int mitre_detector_values_walk(void *stream, const double *series, int64_t series_stride,
                                          const double *times, const int32_t *t_offsets, int S, int N, int V,
                                          const double *windows, int W, const int32_t *wlo, const int32_t *wcnt,
                                          const double *winv_sxx, const double *wtbar, const int64_t *group_base,
                                          const uint8_t *slope_ok, const int32_t *sg_first, int n_sg,
                                          int sg_per_block, int max_block_windows, double *values,
                                          uint8_t *group_flags)
{
    if (N < 1 || N > kWalkMaxThreads || V <= 0 || S <= 0 || W < 0 || n_sg < 0 || series_stride < S ||
        sg_per_block < 1 || max_block_windows < 1)
        return MITRE_ERR_INVALID;
    if (W == 0 || n_sg == 0) return MITRE_OK;
    if (!series || !times || !t_offsets || !windows || !wlo || !wcnt || !winv_sxx || !wtbar || !group_base ||
        !slope_ok || !sg_first || !values)
        return MITRE_ERR_INVALID;
    const size_t smem = mitre_values_walk_smem_bytes(N, S, max_block_windows);
    if (smem > smem_optin()) return MITRE_ERR_UNSUPPORTED;
    const int VB = walk_vars_per_block(N, g_walk_vb_cap);
    const int tiles = (V + VB - 1) / VB;
    const int chunks = (n_sg + sg_per_block - 1) / sg_per_block;
    return launch_values_walk(as_stream(stream), dim3((unsigned)tiles, (unsigned)chunks), smem, series, series_stride,
                              times, t_offsets, S, N, V, windows, wlo, wcnt, winv_sxx, wtbar, group_base, slope_ok,
                              sg_first, 0, n_sg, sg_per_block, max_block_windows, values, group_flags);
}

// ---- the two halves pipelined over runs of start groups ---------------------------------------------------
// The values kernel is bound by instruction issue and dependent fp64 chains at ~28 % occupancy, the sort/emit
// kernel by HBM stores and its own issue rate at ~18 %: alone, each leaves most of an SM idle.  Here the
// start groups are cut into `n_stages` runs; run c's values are computed on one internal stream while run
// c - 1's values are sorted and emitted on another (event-ordered), so the two kernels share the SMs, and a
// run's values (tens of MB) are still in the 126 MB L2 when the sort/emit kernel reads them back.
namespace {
struct WalkPipe {
    cudaStream_t st[2] = {nullptr, nullptr};
    cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> done;
    int device = -1;
    cudaError_t ensure(int n)
    {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev != device) {
            int lo = 0, hi = 0;
            if ((e = cudaDeviceGetStreamPriorityRange(&lo, &hi)) != cudaSuccess) return e;
            for (int i = 0; i < 2; ++i) {
                // st[1] (sort/emit) at the highest priority: its pending blocks are placed before pending
                // blocks of the values kernel whenever an SM frees resources
                if ((e = cudaStreamCreateWithPriority(&st[i], cudaStreamNonBlocking, (i == 1 && g_pipe_priority) ? hi : lo)) != cudaSuccess) return e;
                if ((e = cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming)) != cudaSuccess) return e;
            }
            if ((e = cudaEventCreateWithFlags(&fork, cudaEventDisableTiming)) != cudaSuccess) return e;
            done.clear();
            device = dev;
        }
        while ((int)done.size() < n) {
            cudaEvent_t ev;
            if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return e;
            done.push_back(ev);
        }
        return cudaSuccess;
    }
};
WalkPipe g_walk_pipe;
}  // namespace

extern "C" int mitre_detectors_walk(void *stream, const double *series, int64_t series_stride, const double *times,
                                    const int32_t *t_offsets, int S, int N, int V, const double *windows, int W,
                                    const int32_t *wlo, const int32_t *wcnt, const double *winv_sxx,
                                    const double *wtbar, const int64_t *group_base, const uint8_t *slope_ok,
                                    const int32_t *sg_first, int n_sg, int sg_per_block, int max_block_windows,
                                    const int32_t *stage_sg_host, const int64_t *stage_group_host, int n_stages,
                                    int64_t G, double *values, double *thresholds, int32_t *K, uint64_t *rows_padded,
                                    const uint8_t *group_type, uint8_t *group_flags, int64_t *flagged)
{
    if (N < 2 || N > 63) return MITRE_ERR_UNSUPPORTED;
    if (V <= 0 || S <= 0 || W <= 0 || n_sg <= 0 || series_stride < S || sg_per_block < 1 || max_block_windows < 1 ||
        n_stages < 1 || G <= 0)
        return MITRE_ERR_INVALID;
    if (!series || !times || !t_offsets || !windows || !wlo || !wcnt || !winv_sxx || !wtbar || !group_base ||
        !slope_ok || !sg_first || !stage_sg_host || !stage_group_host || !values || !thresholds || !K || !rows_padded)
        return MITRE_ERR_INVALID;
    if ((group_flags || flagged) && !group_type) return MITRE_ERR_INVALID;
    if (flagged && !group_flags) return MITRE_ERR_INVALID;
    const size_t smem = mitre_values_walk_smem_bytes(N, S, max_block_windows);
    if (smem > smem_optin()) return MITRE_ERR_UNSUPPORTED;
    cudaStream_t user = as_stream(stream);
    WalkPipe &wp = g_walk_pipe;
    MITRE_CUDA_TRY(wp.ensure(n_stages));
    if (group_flags) MITRE_CUDA_TRY(cudaMemsetAsync(group_flags, 0, (size_t)G, user));
    if (flagged) MITRE_CUDA_TRY(cudaMemsetAsync(flagged, 0, sizeof(int64_t), user));
    MITRE_CUDA_TRY(cudaEventRecord(wp.fork, user));
    MITRE_CUDA_TRY(cudaStreamWaitEvent(wp.st[0], wp.fork, 0));
    MITRE_CUDA_TRY(cudaStreamWaitEvent(wp.st[1], wp.fork, 0));
    const int VB = walk_vars_per_block(N, g_walk_vb_cap);
    const int tiles = (V + VB - 1) / VB;
    const int M = N - 1;
    const int saved = g_detector_split;
    int rc = MITRE_OK;
    for (int c = 0; c < n_stages && rc == MITRE_OK; ++c) {
        const int sg0 = stage_sg_host[c], sg1 = stage_sg_host[c + 1];
        const int64_t g0 = stage_group_host[c], g1 = stage_group_host[c + 1];
        if (sg1 <= sg0 || g1 <= g0) continue;
        const int chunks = (sg1 - sg0 + sg_per_block - 1) / sg_per_block;
        rc = launch_values_walk(wp.st[0], dim3((unsigned)tiles, (unsigned)chunks), smem, series, series_stride, times,
                                t_offsets, S, N, V, windows, wlo, wcnt, winv_sxx, wtbar, group_base, slope_ok,
                                sg_first, sg0, sg1, sg_per_block, max_block_windows, values, group_flags);
        if (rc != MITRE_OK) break;
        MITRE_CUDA_TRY(cudaEventRecord(wp.done[c], wp.st[0]));
        MITRE_CUDA_TRY(cudaStreamWaitEvent(wp.st[1], wp.done[c], 0));
        g_detector_split = 1;
        g_emit_blocks_per_sm = (c + 1 < n_stages) ? g_pipe_emit_per_sm : 0;   // the last run has the SMs to itself
        g_slope_guard = (group_type && group_flags) ? SlopeGuardArgs{group_type + g0, group_flags + g0, nullptr}
                                                     : SlopeGuardArgs{nullptr, nullptr, nullptr};
        rc = nb_dispatch(wp.st[1], nullptr, 0, N, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0,
                         nullptr, nullptr, nullptr, nullptr, 0, g1 - g0, values + g0 * N, thresholds + g0 * M, K + g0,
                         rows_padded + g0 * 2 * M);
        g_slope_guard = SlopeGuardArgs{nullptr, nullptr, nullptr};
        g_emit_blocks_per_sm = 0;
        g_detector_split = saved;
    }
    MITRE_CUDA_TRY(cudaEventRecord(wp.join[0], wp.st[0]));
    MITRE_CUDA_TRY(cudaEventRecord(wp.join[1], wp.st[1]));
    MITRE_CUDA_TRY(cudaStreamWaitEvent(user, wp.join[0], 0));
    MITRE_CUDA_TRY(cudaStreamWaitEvent(user, wp.join[1], 0));
    if (rc != MITRE_OK) return rc;
    if (flagged) return launch_count_flags(user, group_flags, G, reinterpret_cast<unsigned long long *>(flagged));
    return MITRE_OK;
}

extern "C" int mitre_detector_sort_emit(void *stream, const double *values, int64_t G, int N, double *thresholds,
                                        int32_t *K, uint64_t *rows_padded, const uint8_t *group_type,
                                        uint8_t *group_flags, int64_t *flagged)
{
    if (N < 2 || N > 63) return MITRE_ERR_UNSUPPORTED;
    if (G < 0) return MITRE_ERR_INVALID;
    if (G == 0) return MITRE_OK;
    if (!values || !thresholds || !K || !rows_padded) return MITRE_ERR_INVALID;
    if ((group_flags || flagged) && !group_type) return MITRE_ERR_INVALID;
    if (flagged && !group_flags) return MITRE_ERR_INVALID;
    // group_flags is NOT cleared here: the values kernel may already have set constant-series flags
    if (flagged) MITRE_CUDA_TRY(cudaMemsetAsync(flagged, 0, sizeof(int64_t), as_stream(stream)));
    const int saved = g_detector_split;
    g_detector_split = 1;
    g_slope_guard = (group_type && group_flags) ? SlopeGuardArgs{group_type, group_flags, nullptr}
                                                 : SlopeGuardArgs{nullptr, nullptr, nullptr};
    const int rc = nb_dispatch(as_stream(stream), nullptr, 0, N, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                               nullptr, 0, nullptr, nullptr, nullptr, nullptr, 0, G, const_cast<double *>(values),
                               thresholds, K, rows_padded);
    g_slope_guard = SlopeGuardArgs{nullptr, nullptr, nullptr};
    g_detector_split = saved;
    if (rc != MITRE_OK) return rc;
    if (flagged) return launch_count_flags(as_stream(stream), group_flags, G, reinterpret_cast<unsigned long long *>(flagged));
    return MITRE_OK;
}

extern "C" int mitre_densify_perm(void *stream, const int64_t *perm_padded, int64_t P, int N,
                                  const int64_t *row_offsets, int64_t *perm)
{
    if (P < 0 || N < 2) return MITRE_ERR_INVALID;
    if (P == 0) return MITRE_OK;
    if (!perm_padded || !row_offsets || !perm) return MITRE_ERR_INVALID;
    int64_t blocks = ceil_div64(P, 256);
    if (blocks > 16 * sm_count()) blocks = 16 * sm_count();
    densify_perm_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(perm_padded, P, 2 * (N - 1), row_offsets,
                                                                         perm);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}

extern "C" int mitre_set_sort_emit_rolled(int enable)
{
    g_sort_emit_rolled = enable ? 1 : 0;
    return MITRE_OK;
}

// Test / tuning hook: 1 = values kernel + sort/emit kernel (default), 0 = the single fused kernel.
extern "C" int mitre_set_detector_split(int enable)
{
    g_detector_split = (enable >= 0 && enable <= 5) ? enable : 1;
    return MITRE_OK;
}


// Test hook: the exact small-integer division used for window means vs the IEEE divide.
extern "C" int mitre_selftest_divide(void *stream, const double *a, const int32_t *n, int64_t count,
                                     double *fast, double *exact)
{
    if (count < 0 || (count > 0 && (!a || !n || !fast || !exact))) return MITRE_ERR_INVALID;
    if (count == 0) return MITRE_OK;
    selftest_divide_kernel<<<(unsigned)ceil_div64(count, 256), 256, 0, as_stream(stream)>>>(a, n, count, fast,
                                                                                          exact);
    MITRE_LAUNCH_CHECK();
    return MITRE_OK;
}
