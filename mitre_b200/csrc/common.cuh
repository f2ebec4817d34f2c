// common.cuh -- shared device/host helpers for libmitre_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mitre_b200.h"

#define MITRE_CUDA_TRY(expr)                        \
    do {                                            \
        cudaError_t _e = (expr);                    \
        if (_e != cudaSuccess) return (int)_e;      \
    } while (0)

#define MITRE_LAUNCH_CHECK()                        \
    do {                                            \
        cudaError_t _e = cudaGetLastError();        \
        if (_e != cudaSuccess) return (int)_e;      \
    } while (0)

namespace mitre {

constexpr int kWarp = 32;

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// Number of SMs of the current device (148 on B200), cached.
int sm_count();
// Opt-in shared memory per block of the current device (227 KB on B200), cached.
size_t smem_optin();

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int words_for(int N) { return (N + 63) >> 6; }

// ---- packed-row bit helpers: subject i -> word i>>6, bit 63-(i&63) ----------------------------
__host__ __device__ inline uint64_t subject_bit(int i) { return 1ull << (63 - (i & 63)); }

// ---- shared-memory address / mbarrier / bulk-copy (TMA) PTX ------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// 1-D bulk async copy global -> shared through the TMA unit (SASS: UBLKCP).  dst, src and
// bytes must be multiples of 16.  Completion is signalled on `bar` as transaction bytes.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// 1-D bulk async copy shared -> global (coalesced full-line stores issued by the TMA unit).
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0()
{
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// Make generic-proxy smem writes visible to the async proxy before a bulk store reads them.
__device__ __forceinline__ void fence_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Streaming (evict-first) global accesses for data touched exactly once.
__device__ __forceinline__ uint64_t ld_stream_u64(const uint64_t *p)
{
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream_f64(double *p, double v)
{
    asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ ulonglong2 ld_stream_u64x2(const ulonglong2 *p)
{
    ulonglong2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ld_stream_f64x2(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream_f64x2(double2 *p, double a, double b)
{
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void st_stream_u64(uint64_t *p, uint64_t v)
{
    asm volatile("st.global.L1::no_allocate.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- draw workspace layout (doubles), shared by scoring.cu (writer of the per-tile pairs) and
// draw.cu (reader):  [partials 2*kStatsBlocks][misc 2][pairs 2*ntiles][chunk sums nchunks]
//   pairs[t] = (max, sum exp(w - max)) of the kDrawTileRows candidates [t*128, t*128+128)
//   chunk c  = kDrawChunkTiles consecutive tiles
constexpr int kStatsBlocks = 592;       // 148 SMs x 4
constexpr int kDrawTileRows = 128;      // one warp pass of the scoring kernel (4 rows per lane)
constexpr int kDrawChunkTiles = 256;
__host__ __device__ inline int64_t draw_ntiles(int64_t P) { return (P + kDrawTileRows - 1) / kDrawTileRows; }
__host__ __device__ inline int64_t draw_nchunks(int64_t P)
{
    return (draw_ntiles(P) + kDrawChunkTiles - 1) / kDrawChunkTiles;
}
__host__ __device__ inline size_t draw_off_pairs() { return 2 * (size_t)kStatsBlocks + 2; }
__host__ __device__ inline size_t draw_off_chunks(int64_t P) { return draw_off_pairs() + 2 * (size_t)draw_ntiles(P); }
__host__ __device__ inline size_t draw_ws_doubles(int64_t P)
{
    return draw_off_chunks(P) + (size_t)draw_nchunks(P);
}

// Order-preserving map double -> uint64 (for finite values and infinities; -0.0 < +0.0 is
// normalised by the caller where it matters).
__device__ __forceinline__ uint64_t f64_key(double v)
{
    uint64_t b = static_cast<uint64_t>(__double_as_longlong(v));
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

}  // namespace mitre
