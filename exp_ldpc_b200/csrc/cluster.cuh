// cluster.cuh — thread-block-cluster helpers shared by the cluster BP kernels (bpc_kernel.cuh, the cluster
// mode of bp7_kernel.cuh): distributed-shared-memory accesses by 32-bit cluster-window address and the
// cluster barrier.
#pragma once

#include <stdint.h>

namespace qd {

__device__ __forceinline__ uint32_t cl_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

// address of `shared_addr` (a shared::cta address of THIS CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t cl_map(uint32_t shared_addr, uint32_t rank)
{
    uint32_t out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(shared_addr), "r"(rank));
    return out;
}
__device__ __forceinline__ double cl_ld(uint32_t a, double) { double v; asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ float cl_ld(uint32_t a, float) { float v; asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void cl_st(uint32_t a, double v) { asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void cl_st(uint32_t a, float v) { asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void cl_st_u32(uint32_t a, uint32_t v) { asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void cl_st_s64(uint32_t a, long long v) { asm volatile("st.shared::cluster.s64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ void cl_red_or_u32(uint32_t a, uint32_t v) { asm volatile("red.relaxed.cluster.shared::cluster.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// Cluster-wide barrier after which every CTA sees the distributed-shared-memory stores and reductions that ANY thread of
// any CTA issued before it: each thread releases its own remote accesses.  (Rounds 1-2 used one `fence.acq_rel.cluster` by
// thread 0 behind the CTA barrier and relaxed arrives; measured again at the end of round 2 the per-thread release costs
// nothing measurable in bpc_kernel and 1.5 % in the cluster mode of bp7_kernel, and it does not lean on one thread's fence
// covering stores that other threads have in flight to another SM.)
__device__ __forceinline__ void cluster_sync_release_all()
{
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

}  // namespace qd
