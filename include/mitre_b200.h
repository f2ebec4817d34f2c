/*
 * mitre_b200.h -- C ABI of libmitre_b200.so: the B200 (sm_100a) replacement for MITRE's MCMC
 * likelihood hot path.  Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * gerberlab/mitre tree).  Conventions:
 *
 *   * Every pointer is a DEVICE pointer unless the parameter name ends in `_host` or the
 *     function name ends in `_host`.  `stream` is a cudaStream_t passed as void* (NULL = the
 *     legacy default stream).  Device entry points are asynchronous and stream-ordered; they
 *     allocate nothing -- the caller owns inputs, outputs and workspaces.
 *   * Return value: 0 = MITRE_OK; < 0 = argument / capability error (MITRE_ERR_*);
 *     > 0 = a cudaError_t from the runtime.  Nothing is thrown across the ABI.
 *   * Numerical failures that the reference reports as numpy LinAlgError
 *     (efficient_likelihoods.pyx:42, mvnpdf_cholesky.py:66) are reported through the caller's
 *     `status` word on the device (bit flags MITRE_STATUS_*), OR-ed in by the kernels.
 *
 * PACKED TRUTH ROWS.  The reference keeps `flat_truth` as bool[P, N] (rules.py:510) and converts
 * it to float64[P, N] on every sweep (logit_rules.py:214-217).  Here a row is W64 = ceil(N/64)
 * uint64 words; subject i lives in word i/64 at bit (63 - i%64).  With that bit order, comparing
 * rows word by word as unsigned integers is exactly the reference's lexicographic row order
 * (subject 0 most significant, False < True; rules.py:447-450).  Padding bits are zero.
 */
#ifndef MITRE_B200_H
#define MITRE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MITRE_OK 0
#define MITRE_ERR_INVALID (-1)      /* null pointer, negative size, inconsistent shape      */
#define MITRE_ERR_UNSUPPORTED (-2)  /* shape outside what the kernels are built for          */
#define MITRE_ERR_WORKSPACE (-3)    /* workspace smaller than *_workspace_bytes() asks for    */
#define MITRE_ERR_NOT_PD (-4)       /* host-synchronous entries only: Cholesky failed         */
#define MITRE_ERR_NO_DEVICE (-5)    /* no sm_100 device / kernel image not loadable           */

#define MITRE_STATUS_BASE_NOT_PD 1u   /* base system I/sigma^2 + X'OmegaX not positive definite */
#define MITRE_STATUS_BAD_SCHUR 2u     /* a candidate's Schur complement was <= 0 or non-finite  */
#define MITRE_STATUS_COV_NOT_PD 4u    /* dense covariance not positive definite                 */
#define MITRE_STATUS_COMM_TIMEOUT 8u  /* sharded draw: a peer's pair / the owner's index never arrived */

#define MITRE_MAX_P 24                /* columns of the base design matrix X the kernels accept */

/* ---- library ------------------------------------------------------------------------------ */

const char *mitre_version(void);
const char *mitre_error_string(int code);
/* sm count, compute capability and opt-in shared memory of the current device. */
int mitre_device_info(int *sm_count, int *cc_major, int *cc_minor, size_t *smem_optin);

/* Measurement aid: launches a pure DFMA kernel (8 independent chains per thread); the caller times
 * it with stream events.  *flops_launched_host receives the flop count of the launch.  Gives the
 * fp64-pipe denominator for the scoring kernel (MEASURED_PEAKS.json has no fp64 entry). */
int mitre_probe_fp64(void *stream, int iters, double *sink, double *flops_launched_host);

/* ---- packing (layout conversion at the boundary) ------------------------------------------ */

/* float64[P, N] (non-zero = True; the array the reference passes as `list_of_truth_vectors`,
 * efficient_likelihoods.pyx:19) -> packed uint64[P, W64]. */
int mitre_pack_rows_f64(void *stream, const double *truth, int64_t P, int N, uint64_t *packed);
/* bool/uint8[P, N] (the reference's flat_truth, rules.py:510) -> packed. */
int mitre_pack_rows_u8(void *stream, const uint8_t *truth, int64_t P, int N, uint64_t *packed);
/* packed -> uint8[P, N] of 0/1. */
int mitre_unpack_rows_u8(void *stream, const uint64_t *packed, int64_t P, int N, uint8_t *truth);

/* ---- scoring: replaces efficient_likelihoods.likelihoods_of_all_options -------------------- */
/*
 * Reference: mitre/efficient_likelihoods.pyx:15-169, called from
 * LogisticRuleModel.likelihoods_of_all_options (logit_rules.py:186-229).
 *
 *   out[k] = log N( z ; 0, diag(1/omega) + sigma2 * [X | v_k][X | v_k]^T ),  z = kappa/omega,
 *   v_k = packed_truth[k] AND mask           (the AND is logit_rules.py:214-216)
 *
 * for all k in [0, P) in one launch.  X is float64[N, p] row-major (any real entries: rule
 * columns, the constant column, categorical / continuous covariates, rules.py:1152-1179),
 * kappa = y - 0.5 (logit_rules.py:222).  `mask` is packed uint64[W64] (all subjects set when
 * the rule has no other primitive).  `workspace` must hold mitre_score_workspace_bytes(N, p)
 * bytes and stay untouched until the launch has run.  `status` (device int32, caller-zeroed)
 * receives MITRE_STATUS_* flags.
 * Kernel selection (N <= 64): tables of subset sums -- 8-subject tables in shared memory for small
 * candidate sets; for P >= 4 Mi rows (mitre_set_large_table_min_rows moves that switch point; a
 * tuning / test hook) 16-subject tables in L2, the leading ones shared by the four consecutive rows
 * a thread owns.  mitre_set_split_tables(1) selects a variant that keeps the table of the last 10-11
 * subjects in shared memory instead (same speed on B200, see scoring.cu; off by default).
 */
size_t mitre_score_workspace_bytes(int N, int p);
int mitre_set_large_table_min_rows(int64_t rows);
int mitre_set_split_tables(int enable);
int mitre_set_wide_tables(int enable);   /* N > 64: 1 (default) byte tables in L2, 0 one fetch per set bit */
int mitre_set_dedupe_table_bits(int bits); /* ... its subset-sum tables cover `bits` subjects each (11..16; default 11:
                                            the last table is 2048 entries and stays in L1) */
int mitre_set_dedupe(int mode);          /* large tables, N <= 64: 1 (default) score each run of equal rows once;
                                            2 the same with four lanes per 64-byte table entry; 0 off */
int mitre_score_all(void *stream, const uint64_t *packed_truth, int64_t P, int W64,
                    const uint64_t *mask, const double *X, const double *omega,
                    const double *kappa, int N, int p, double sigma2, void *workspace,
                    size_t workspace_bytes, double *out, int32_t *status);

/*
 * Same sweep, additionally reducing (max, sum exp(w - max)) of w = out (+ log_prior, float64[P] or
 * NULL) into stats[2] in the scoring kernel's epilogue (N <= 64; for N > 64 a second pass over
 * `out` using stats_workspace of mitre_draw_workspace_bytes(P) bytes).  When stats_workspace is given
 * it also receives the per-128-candidate (max, sum-exp) pairs mitre_draw_from_tiles consumes, so the
 * draw that follows needs no second pass over the P weights.  log-sum-exp =
 * stats[0] + log(stats[1]): what the add/remove moves need (logit_rules.py:959, 1064) and the
 * 16-byte pair ranks exchange in the two-stage multi-GPU draw.
 */
int mitre_score_all_stats(void *stream, const uint64_t *packed_truth, int64_t P, int W64,
                          const uint64_t *mask, const double *X, const double *omega,
                          const double *kappa, int N, int p, double sigma2, void *workspace,
                          size_t workspace_bytes, double *out, int32_t *status,
                          const double *log_prior, double *stats, void *stats_workspace,
                          size_t stats_workspace_bytes);

/* Group scoring: the same log-likelihoods for ALL candidates of G (window, variable, type) groups,
 * in the reference's ENUMERATION order (rules.py:465-511: group, threshold ascending, above then
 * below), computed from the detector values without truth rows.  Within a group the candidates are
 * nested sets along the sorted order of its N values, so their masked sums are suffix sums: O(N p)
 * per group instead of per candidate (the equivalent algebra noted next to
 * efficient_likelihoods.pyx:66-167's shared-prefix reuse).  values float64[G, N], thresholds
 * float64[G, N-1], K int32[G], row_offsets int64[G+1] as produced by mitre_detector_values /
 * _thresholds / _row_offsets (or mitre_detectors_fused); out float64[row_offsets[G]].
 * mask / X / omega / kappa / workspace / status as for mitre_score_all.  N <= 65535.
 * group_prior float64[G] or NULL: added to every candidate of the group (a primitive's log-prior
 * depends only on its window and variable, logit_rules.py:424-457), so that `out` holds the
 * log-weights directly and no per-candidate prior vector is needed.
 * N <= 64: one warp per group (64-key bitonic sort over shuffles); larger N: one CTA per group
 * (mitre_set_groups_warp(0) forces the CTA kernel; test hook). */
int mitre_set_groups_warp(int enable);
int mitre_score_groups(void *stream, const double *values, const double *thresholds, const int32_t *K,
                       const int64_t *row_offsets, int64_t G, const uint64_t *mask, const double *X,
                       const double *omega, const double *kappa, int N, int p, double sigma2,
                       void *workspace, size_t workspace_bytes, const double *group_prior, double *out,
                       int32_t *status);

/*
 * Drop-in compatibility form with the reference's exact arguments, all HOST pointers
 * (efficient_likelihoods.pyx:15-21): X[N,p], omega[N], response[N] (= kappa/omega),
 * truth float64[P,N], prior_coefficient_variance; writes out_host[P].  Synchronous: stages the
 * rows through the device in chunks, packs, scores and copies back.
 * Returns MITRE_ERR_NOT_PD where the reference raises LinAlgError.
 */
int mitre_score_all_host(const double *X_host, const double *omega_host,
                         const double *response_host, const double *truth_host,
                         double sigma2, int N, int p, int64_t P, double *out_host);
int mitre_set_host_chunk_bytes(int64_t bytes); /* test hook: packed bytes per staging chunk of
                                                  mitre_score_all_host (0 = default, 64 MiB) */

/* ---- dense single-rule-list likelihood: replaces mc_logpdf_np ------------------------------ */
/*
 * Reference: LogisticRuleModel.likelihood_z_given_X (logit_rules.py:171-177) ->
 * mvnpdf_cholesky.mc_logpdf_np (mvnpdf_cholesky.py:53-74).  Batched so that move_primitive's
 * 2(2R'+1) sequential calls (logit_rules.py:1375-1401) are one launch.
 *
 *   out[b] = log N( kappa/omega ; 0, diag(1/omega) + sigma2 * X_b X_b^T )
 *
 * Xs is float64[B, N, pmax] row-major; design matrix b uses its first p_b[b] columns.
 */
int mitre_likelihood_z_given_X_batch(void *stream, const double *Xs, const int32_t *p_b,
                                     int pmax, const double *omega, const double *kappa,
                                     int N, int B, double sigma2, double *out,
                                     int32_t *status);
/*
 * mc_logpdf_np(x, cov) itself (mvnpdf_cholesky.py:53-74) for an arbitrary dense covariance:
 * x[N], cov[N,N] row-major (read only), scratch[N*N] doubles of device scratch, out[1].
 */
int mitre_mc_logpdf(void *stream, const double *x, const double *cov, int N, double *scratch,
                    double *out, int32_t *status);

/* ---- detector evaluation: replaces DiscreteRulePopulation.mine_rules ----------------------- */
/*
 * Reference: mitre/rules.py:325-436 (mine_rules) = window admission (391-421),
 * calculate_values (578-600) over PrimitiveRule.evaluate (80-141), update_base_rules
 * (465-511) over thresholds_from_values (513-576), sort_base_rules (439-462).
 *
 * Series layout in HBM: the N subjects' samples are concatenated along one axis of length
 * S = sum_s T_s; t_offsets int32[N+1] delimits each subject; times float64[S] must be
 * non-decreasing within a subject; series float64[V, series_stride] holds variable v's samples
 * for all subjects at series + v*series_stride (series_stride >= S, in elements).
 */

/* For every (window w, subject s): first sample index (global, into [0,S)) and the number of
 * samples with t0 <= t <= t1 (closed both ends, rules.py:100, 406-409).  lo, cnt: int32[W, N].
 * The caller admits a window iff min_s cnt >= 1, and its slope iff min_s cnt >= 2. */
int mitre_window_ranges(void *stream, const double *times, const int32_t *t_offsets, int N,
                        const double *windows, int W, int32_t *lo, int32_t *cnt);

/* calculate_values: values float64[G, N].  Groups are enumerated window-major, then variable,
 * then type (slope, average) (rules.py:474-480): group index of (w, v, type) is
 * group_base[w] + v*(1+slope_ok[w]) + (slope_ok[w] ? type : 0), type 0 = slope, 1 = average.
 * 'average' is bit-identical to np.mean of the masked copy (numpy pairwise summation order);
 * 'slope' is the closed-form OLS slope of x on (t - midpoint) (rules.py:110-114 reaches the
 * same quantity through LAPACK gelsd; values agree to ~1e-11 relative).
 * max_T = the largest per-subject sample count (sizes the shared-memory staging).  For the
 * TMA staging path `series` and `times` must be 16-byte aligned, series_stride even, and both
 * arrays readable up to the next even element past S (pad by one element if S is odd). */
int mitre_detector_values(void *stream, const double *series, int64_t series_stride,
                          const double *times, const int32_t *t_offsets, int N, int V,
                          const double *windows, const int32_t *lo, const int32_t *cnt,
                          const int64_t *group_base, const uint8_t *slope_ok, int W, int max_T,
                          double *values);

/* thresholds_from_values for every group (max_thresholds = None): sort, midpoints of
 * neighbours, unique.  thresholds float64[G, N-1] (first K[g] entries of each row valid),
 * K int32[G]. */
int mitre_detector_thresholds(void *stream, const double *values, int64_t G, int N,
                              double *thresholds, int32_t *K);

/* Exclusive scan of 2*K[g] -> row_offsets int64[G+1]; row_offsets[G] = P.
 * workspace: mitre_scan_workspace_bytes(G) bytes. */
size_t mitre_scan_workspace_bytes(int64_t n);
int mitre_detector_row_offsets(void *stream, const int32_t *K, int64_t G, int64_t *row_offsets,
                               void *workspace, size_t workspace_bytes);

/* update_base_rules: for group g, threshold k (ascending), direction d (0 above: v > th,
 * 1 below: v <= th; rules.py:494-497) write packed row row_offsets[g] + 2k + d. */
int mitre_detector_rows(void *stream, const double *values, const double *thresholds,
                        const int32_t *K, const int64_t *row_offsets, int64_t G, int N,
                        uint64_t *packed_truth);

/* sort_base_rules: stable lexicographic sort of the packed rows (np.lexsort(np.rot90(truth)),
 * rules.py:450).  perm int64[P]: perm[i] = enumeration index of the row at sorted position i
 * (the reference's _reordering_cache); sorted_rows uint64[P, W64] may be NULL. */
size_t mitre_lexsort_workspace_bytes(int64_t P, int W64);
int mitre_lexsort_rows(void *stream, const uint64_t *rows, int64_t P, int W64, int N,
                       int64_t *perm, uint64_t *sorted_rows, void *workspace,
                       size_t workspace_bytes);
/* Tables whose row and enumeration index fit one 64-bit word (W64 == 1, N + bits(P) <= 64, P < 2^30: S2, S5)
 * take a one-sweep LSD radix sort of (row | index) words -- ceil(N / 9) passes, each reading its input once
 * (decoupled look-back between tiles), same permutation.  mitre_set_packed_sort(0) keeps the general path
 * (returns the previous setting).
 *   mitre_lexsort_padded  the same sort straight from the padded detector output rows_padded
 *       uint64[P_slots] (mitre_detectors_fused: 2(N-1) slots per group, unused slots = ~0): sentinel slots are
 *       dropped in the first pass, the last pass writes sorted_rows uint64[P_valid] (may be NULL) and
 *       perm int64[P_valid] already mapped from slot indices to dense enumeration indices through
 *       row_offsets int64[G+1] (what mitre_densify_perm does).  MITRE_ERR_UNSUPPORTED when the table is not
 *       eligible: use mitre_lexsort_rows(..., N + 1, ...) + mitre_densify_perm then.
 *   mitre_lexsort_check   synchronises the stream and reports whether a look-back wait of the last packed
 *       sort in `workspace` expired (never observed; the waits are bounded so that a lost CTA cannot hang the
 *       device): MITRE_OK or MITRE_ERR_UNSUPPORTED. */
int mitre_lexsort_padded(void *stream, const uint64_t *rows_padded, int64_t P_slots, int N,
                         int64_t P_valid, const int64_t *row_offsets, int64_t *perm,
                         uint64_t *sorted_rows, void *workspace, size_t workspace_bytes);
int mitre_lexsort_check(void *stream, void *workspace, int64_t P, int N);
int mitre_set_packed_sort(int enable);

/* ---- fused detector evaluation (N <= 63): values + thresholds + truth rows, no host round trip --
 *
 * Same outputs as mitre_detector_values / _thresholds / _rows.  Default: two launches.
 *   values      block = 32 variables x a run of windows; the tile's samples sit in shared memory for
 *               the whole block (series_t is read from HBM once), both detector types come out of
 *               one pass over a window's samples, finished tiles leave as contiguous ranges.  When
 *               the sample axis does not fit in shared memory: block = window x 32 variables,
 *               samples read through L2.
 *   sort/emit   one thread owns one (window, variable, type) group: register-resident sorting
 *               network, then thresholds and both truth rows of every threshold; the block's slices
 *               of values / thresholds / rows move between HBM and shared memory as TMA bulk copies.
 * mitre_set_detector_split selects other variants for tests and timing: 0 = the earlier single
 * kernel that does all of it per thread; 1 = default; 2 = values with one thread per value;
 * 3 = values through L2 even when the resident kernel fits; 4 = sort/emit with direct global
 * accesses instead of staged tiles; 5 = timing probe: sort/emit moves its tiles but skips the
 * sort and the emission (outputs undefined) -- the kernel's data-movement floor.
 *
 * mitre_window_lists precomputes, once per window set, everything that does not depend on the
 * variable (ld = row stride of the per-sample lists, >= S):
 *   lo, cnt, start int32[W, N]; inv_n, inv_sxx float64[W, N]; idx int32[W, ld]; dt float64[W, ld]
 *   -- first sample / sample count / position in the window's flat sample list; 1/cnt and
 *   1/sum((t - mid) - tbar)^2; the flat list of element offsets (sample index * vstride, must fit
 *   int32) into series_t and of centred times.
 *
 * mitre_detectors_fused:
 *   series_t     float64[S_pad, vstride]: the series TRANSPOSED (sample-major, variable fastest,
 *                vstride >= V), so the 32 lanes of a warp read 32 consecutive variables of one sample
 *   task_window / task_type  int32 / int8 [n_tasks_wt]: the (window, type) pairs to evaluate
 *                (type 0 = slope, only for windows with slope_ok; 1 = average)
 *   n_windows    W, the number of windows behind the lists; G = number of groups (rows of values)
 *   rows_padded  uint64[G, 2(N-1)]: group g's first 2*K[g] slots hold its (above, below) row pairs
 *                in threshold order; unused slots hold the sentinel ~0.  A stable lexsort of this
 *                padded table with N+1 significant bits (mitre_lexsort_rows(..., N + 1, ...)) leaves
 *                the P = 2*sum K valid rows first, in the reference order; mitre_densify_perm maps
 *                the padded slot indices of that permutation to the reference's dense enumeration
 *                indices (row_offsets[g] + slot).
 * mitre_selftest_divide is a test hook: the kernel's exact small-integer division (window means)
 *   next to the IEEE divide, elementwise.
 */
int mitre_window_lists(void *stream, const double *times, const int32_t *t_offsets, int N,
                       const double *windows, int W, int64_t ld, int64_t vstride, int32_t *lo,
                       int32_t *cnt, int32_t *start, double *inv_n, double *inv_sxx, int32_t *idx,
                       double *dt);
int mitre_detectors_fused(void *stream, const double *series_t, int64_t vstride, int N, int V,
                          const int32_t *wlo, const int32_t *wcnt, const int32_t *wstart,
                          const double *winv_n, const double *winv_sxx, const int32_t *widx,
                          const double *wdt, int64_t ld, const int64_t *group_base,
                          const uint8_t *slope_ok, const int32_t *task_window,
                          const int8_t *task_type, int n_tasks_wt, int n_windows, int64_t G,
                          double *values, double *thresholds, int32_t *K, uint64_t *rows_padded,
                          const uint8_t *group_type, uint8_t *group_flags, int64_t *flagged);
/* ---- walk-forward detector values (default device path, N <= 63) + stand-alone sort/emit ------------------
 *
 * mitre_detector_values_walk computes the same values float64[G, N] as mitre_detector_values / the values
 * half of mitre_detectors_fused, an order of magnitude cheaper: windows are pairs (t_i, t_j) of one grid
 * (rules.py:391-402) enumerated start-major, so the admitted windows sharing a start ("start group":
 * windows sg_first[k] .. sg_first[k+1]-1, int32[n_sg + 1]) are prefixes of one another, and a thread per
 * (variable, subject) walks the samples once per start group, emitting the mean (numpy's pairwise order is
 * prefix-stable: bit-identical to np.mean) and the slope (running sums of x - x_first against t - t_i) at
 * every window's sample count.  series float64[V, series_stride] variable-major (Dataset.device_series
 * 'series'), times float64[S], wlo / wcnt int32[W, N] (mitre_window_ranges or mitre_window_lists), winv_sxx
 * float64[W, N] (mitre_window_lists), wtbar float64[W, N] (mitre_window_tbar: mean of t - t0 over the
 * window's samples), group_base / slope_ok as for mitre_detector_values.  group_flags (optional, uint8[G],
 * caller-zeroed): constant non-zero series are flagged (slope guard band, below).  A block takes
 * sg_per_block consecutive start groups of a tile of variables and keeps the (window, subject) tables of all
 * their windows in shared memory: max_block_windows = the largest window count of such a run; the caller
 * picks sg_per_block so that mitre_values_walk_smem_bytes(N, S, max_block_windows) fits the device's opt-in
 * shared memory (else MITRE_ERR_UNSUPPORTED: use mitre_detectors_fused).
 * mitre_detector_sort_emit is the second half of mitre_detectors_fused on its own: values -> thresholds, K,
 * padded truth rows (+ near-tie flags into group_flags, which it does NOT clear, and the count). */
int mitre_window_tbar(void *stream, const double *times, int N, const double *windows, int W,
                      const int32_t *wlo, const int32_t *wcnt, double *tbar);
size_t mitre_values_walk_smem_bytes(int N, int S, int max_block_windows);
int mitre_detector_values_walk(void *stream, const double *series, int64_t series_stride,
                               const double *times, const int32_t *t_offsets, int S, int N, int V,
                               const double *windows, int W, const int32_t *wlo, const int32_t *wcnt,
                               const double *winv_sxx, const double *wtbar, const int64_t *group_base,
                               const uint8_t *slope_ok, const int32_t *sg_first, int n_sg,
                               int sg_per_block, int max_block_windows, double *values, uint8_t *group_flags);
int mitre_detector_sort_emit(void *stream, const double *values, int64_t G, int N, double *thresholds,
                             int32_t *K, uint64_t *rows_padded, const uint8_t *group_type,
                             uint8_t *group_flags, int64_t *flagged);
/* Both halves pipelined: the start groups are cut into n_stages runs (stage_sg_host int32[n_stages + 1]: start
 * group bounds; stage_group_host int64[n_stages + 1]: the matching group bounds, group_base of each run's first
 * window and G at the end; both HOST arrays).  Run c's values are computed on one internal stream while run
 * c - 1 is sorted and emitted on another, ordered by events; the call returns with the caller's stream waiting
 * on both.  Neither kernel fills an SM alone (issue-bound at low occupancy vs store-bound), so they overlap,
 * and a run's values are re-read from L2 instead of HBM.  Outputs and flags as for the two calls above
 * (group_flags is cleared here). */
int mitre_detectors_walk(void *stream, const double *series, int64_t series_stride, const double *times,
                         const int32_t *t_offsets, int S, int N, int V, const double *windows, int W,
                         const int32_t *wlo, const int32_t *wcnt, const double *winv_sxx, const double *wtbar,
                         const int64_t *group_base, const uint8_t *slope_ok, const int32_t *sg_first, int n_sg,
                         int sg_per_block, int max_block_windows, const int32_t *stage_sg_host,
                         const int64_t *stage_group_host, int n_stages, int64_t G, double *values,
                         double *thresholds, int32_t *K, uint64_t *rows_padded, const uint8_t *group_type,
                         uint8_t *group_flags, int64_t *flagged);

/* Slope guard band.  'slope' detector values come from a closed form here and from LAPACK gelsd in the
 * reference (rules.py:110-114); the truth rows of a group depend only on the order and tie structure of
 * its N values, so they are the reference's unless (a) two neighbours of the sorted values are unequal but
 * closer than 1e-10 x the group's largest |value|, or (b) a subject's slope is exactly 0 while its window
 * mean is not (a constant non-zero series: the reference returns +-1e-20 rounding noise there, not 0).
 * Such groups are FLAGGED: group_type uint8[G] (0 = slope, 1 = average; a slope group is followed by its
 * window's average group), group_flags uint8[G] (optional) <- 0/1, *flagged (device int64, optional) <-
 * the number of flagged groups (needs group_flags).  mitre_detectors_fused does the check inside its own
 * kernels when group_type is given (last three arguments; all NULL = off): the values kernel sees a
 * subject's mean and slope together (b), the sort/emit kernel holds the sorted values (a);
 * mitre_slope_guard is the same check as a separate pass over values float64[G, N], any N (the
 * phase-by-phase path). */
int mitre_slope_guard(void *stream, const double *values, int64_t G, int N, const uint8_t *group_type,
                      uint8_t *group_flags, int64_t *flagged);
int mitre_densify_perm(void *stream, const int64_t *perm_padded, int64_t P, int N,
                       const int64_t *row_offsets, int64_t *perm);
int mitre_set_detector_split(int enable); /* tuning/test hook, see above */
int mitre_set_walk_tuning(int vb_cap, int emit_blocks_per_sm, int emit_priority); /* variables per walk CTA (1..16,
                                             default 16); pipelined call (mitre_detectors_walk): resident sort/emit
                                             blocks per SM while a values kernel runs beside them (0 = all) and
                                             whether the sort/emit stream runs at high priority (takes effect when
                                             the internal streams are created) */
int mitre_walk_vars_per_block(int N);     /* the tile width the plan of mitre_detector_values_walk must assume */
int mitre_set_walk_vpt(int vpt);          /* variables per thread of the walk-forward values kernel: 1 (default) or 2 (
                                             shared control flow, two interleaved fp64 chains; measured slower: 1.12 vs 0.82 ms at S5); same values.
                                             Returns the previous setting. */
int mitre_set_sort_emit_rolled(int enable); /* sort/emit with rolled emission loops over shared-memory copies of the
                                              sorted values (small code, ~70 registers) instead of the unrolled
                                              register-resident variant */
int mitre_selftest_divide(void *stream, const double *a, const int32_t *n, int64_t count,
                          double *fast, double *exact);

/* ---- proposal draw: replaces the host-side weight assembly + choose_from_relative_probabilities
 *
 * Reference: weights = likelihoods + primitive_priors + multinomial_priors + log(relative_rates)
 * (logit_rules.py:861-862) then choose_from_relative_probabilities(np.exp(weights))
 * (rules.py:1261-1264: cumsum, marker = u * total, index = #{cum < marker}).
 *
 * mitre_weight_stats: (max, sum exp(w - max)) of w = a + b (b may be NULL) -- stats[0] = max,
 *   stats[1] = sum; log-sum-exp = stats[0] + log(stats[1]).  This 16-byte pair is what ranks
 *   all-gather for the two-stage multi-GPU draw.
 * mitre_draw_index: index_out[0] = #{ i : cumsum(exp(w - shift))[i] < u * total }.
 *   stabilise = 0: shift = 0, i.e. the reference's raw np.exp(weights) including its underflow
 *   behaviour; stabilise = 1: shift = max w (stats is then computed first and must be non-NULL);
 *   stabilise = 2: shift = stats[0] as the caller left it (mitre_score_all_stats / mitre_weight_stats
 *   over the same weights), saving the extra pass.
 *   Two-level: candidates are cut into tiles of 128 whose masses (max, sum exp(w - max)) locate the
 *   tile holding the marker; inside that tile the cumulative sum is the reference's strict
 *   left-to-right order.  The drawn index can differ from np.cumsum's only when u*total lies within
 *   rounding (~1e-16 relative) of a partial sum.
 */
size_t mitre_draw_workspace_bytes(int64_t P);
int mitre_weight_stats(void *stream, const double *a, const double *b, int64_t P, double *stats,
                       void *workspace, size_t workspace_bytes);
int mitre_draw_index(void *stream, const double *a, const double *b, int64_t P, double u,
                     int stabilise, double *stats, int64_t *index_out, void *workspace,
                     size_t workspace_bytes);
/* Same draw without the pass over the weights: `workspace` still holds the per-tile pairs that
 * mitre_score_all_stats(..., stats_workspace = workspace, ...) left there for exactly these weights
 * (a = out, b = log_prior).  stabilise 0 / non-zero as for mitre_draw_index (non-zero: shift = stats[0]). */
int mitre_draw_from_tiles(void *stream, const double *a, const double *b, int64_t P, double u,
                          int stabilise, const double *stats, int64_t *index_out, void *workspace,
                          size_t workspace_bytes);

/* ---- sharded draw: one problem's candidates in contiguous slices over the GPUs of one box ---------
 *
 * Reference precedent: _likelihood_worker's [start, stop) slices of the candidate table
 * (logit_rules.py:21-60); the consumer is the same choose_from_relative_probabilities draw over ALL
 * candidates (logit_rules.py:861-863, rules.py:1261-1264).  Every rank scores its slice with
 * mitre_score_all_stats (stats_workspace = its draw workspace, so the per-tile pairs are there), calls
 * mitre_sharded_chunk_sums, and then either
 *   mitre_sharded_draw    does both exchanges itself over NVLink peer memory inside one kernel: every
 *                         rank stores its (max, sum-exp) pair into every peer's slot buffer, waits for
 *                         all pairs of this step's sequence number, derives the common plan (total mass,
 *                         marker, owner);
 *                         the owner selects inside its slice and stores the global index into every
 *                         peer's result slot.  index_out[0] = global index, stats_out = (M, total) with
 *                         log-sum-exp over all ranks' weights = M + log(total).  Waits are bounded
 *                         (MITRE_STATUS_COMM_TIMEOUT in *status, index -1).  ALL ranks must launch it,
 *                         each on its own GPU.
 *   mitre_sharded_select  the same selection with the exchanges left to the caller (NCCL): `gathered`
 *                         float64[world, 2] = all ranks' pairs; non-owners write index -1 (all-reduce
 *                         MAX of index_out afterwards).
 * `u` (the uniform draw) and `seq` are DEVICE float64[1]: seq is the step's sequence number, an integer
 * >= 1 that the caller initialises to the same value on all ranks and the kernel itself advances by one
 * per launch -- device-resident so that the whole step can be replayed as a CUDA graph.
 * mitre_comm_create allocates the slot buffer other ranks write into and returns its 64-byte CUDA IPC
 * handle in handle_out_host; mitre_comm_connect takes all ranks' handles (world x 64 bytes, rank order).
 * world <= 16.  world == 1 works without IPC (a single-GPU run of the same code path).
 */
int mitre_comm_create(int world, int rank, void **comm_out_host, void *handle_out_host);
int mitre_comm_connect(void *comm, const void *all_handles_host);
int mitre_comm_destroy(void *comm);
int mitre_sharded_chunk_sums(void *stream, int64_t P_local, const double *local_stats, void *workspace,
                             size_t workspace_bytes);
int mitre_sharded_draw(void *stream, void *comm, const double *u, double *seq, const double *local_stats,
                       const double *a, const double *b, int64_t P_local, int64_t start, void *workspace,
                       size_t workspace_bytes, int64_t *index_out, double *stats_out, int32_t *status);
int mitre_sharded_select(void *stream, int world, int rank, const double *u, const double *gathered,
                         const double *a, const double *b, int64_t P_local, int64_t start, void *workspace,
                         size_t workspace_bytes, int64_t *index_out, double *stats_out);

/* ---- primitive prior on the device: replaces the O(P) host evaluation of flat_prior ------------
 *
 * Reference: LogisticRuleModel.flat_prior (logit_rules.py:424-457): log-prior of every primitive =
 * Normal logpdf of its variable's log-weight + Beta logpdf of its window's relative duration,
 * minus the log-sum-exp over all primitives.  Both terms depend only on the primitive's
 * (variable, window), i.e. on its group, so the host evaluates V + W scalars and
 *   mitre_row_groups  maps every lexsorted row to its group once per population
 *                     (binary search of perm[i] in row_offsets int64[G+1]);
 *   mitre_row_prior   writes out[i] = by_variable[group_variable[g]] + by_window[group_window[g]] - norm.
 */
int mitre_row_groups(void *stream, const int64_t *perm, int64_t P, const int64_t *row_offsets,
                     int64_t G, int32_t *row_group);
int mitre_row_prior(void *stream, const int32_t *row_group, int64_t P, const int32_t *group_variable,
                    const int32_t *group_window, const double *by_variable, const double *by_window,
                    double norm, double *out);

/* ---- omega / beta Gibbs inner loop on the device ----------------------------------------------------
 *
 * Reference: LogisticRuleSampler.update_omega (logit_rules.py:753-760, pypolyagamma.pgdrawv) and
 * update_beta (762-787), run 10x per step (567-573).  One single-CTA launch:
 *   mode 0: n_iters x { omega_i ~ PG(1, (X beta)_i);  beta ~ N(V X'kappa, V), V = (I/v + X' Omega X)^-1 }
 *   mode 1: n_iters independent beta draws for the omega passed in `omega` (input)
 * X float64[N, p]; beta_io float64[p] (start value in, last draw out); omega float64[N];
 * betas_out float64[n_iters, p].  Philox4x32-10 keyed by `seed`.  Distributionally (not
 * stream-for-stream) equivalent to the host samplers.  mitre_pg_draw: out[i] ~ PG(1, z[i]).
 */
int mitre_gibbs_omega_beta(void *stream, const double *X, const double *kappa, int N, int p,
                           double prior_variance, int n_iters, int mode, uint64_t seed,
                           double *beta_io, double *omega, double *betas_out, int32_t *status);
int mitre_pg_draw(void *stream, const double *z, int64_t n, uint64_t seed, double *out);

#ifdef __cplusplus
}
#endif
#endif /* MITRE_B200_H */
