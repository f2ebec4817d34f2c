// fused_inst.cu -- one instantiation of the fused detector kernel per translation unit
// (compiled once per -DMITRE_FUSED_NB=<4|8|...|64> so the fully unrolled sorting networks build
// in parallel).
#include "fused_detectors.cuh"

#ifndef MITRE_FUSED_NB
#error "compile with -DMITRE_FUSED_NB=<NB>"
#endif
#define MITRE_CAT2(a, b) a##b
#define MITRE_CAT(a, b) MITRE_CAT2(a, b)

namespace mitre {
int MITRE_CAT(launch_fused_nb, MITRE_FUSED_NB)(cudaStream_t st, const double *series_t, int64_t vstride, int N, int V,
                     const int32_t *wlo, const int32_t *wcnt, const int32_t *wstart,
                     const double *winv_n, const double *winv_sxx, const int32_t *widx,
                     const double *wdt, int64_t ld, const int64_t *group_base,
                     const uint8_t *slope_ok, const int32_t *task_window, const int8_t *task_type,
                     int n_wt, int64_t G, double *values, double *thresholds, int32_t *K, uint64_t *rows)
{
    if (g_detector_split) return launch_split_impl<MITRE_FUSED_NB>(st, series_t, vstride, N, V, wlo, wcnt, wstart, winv_n, winv_sxx, widx, wdt, ld, group_base, slope_ok, task_window, task_type, n_wt, G, values, thresholds, K, rows);
    return launch_fused_impl<MITRE_FUSED_NB>(st, series_t, vstride, N, V, wlo, wcnt, wstart, winv_n, winv_sxx, widx, wdt, ld, group_base, slope_ok, task_window, task_type, n_wt, values, thresholds, K, rows);
}
}  // namespace mitre
