// cluster.cuh — thread-block-cluster helpers shared by the cluster BP kernels (bpc_kernel.cuh, the cluster
// mode of bp7_kernel.cuh): distributed-shared-memory accesses by 32-bit cluster-window address and a
// cluster barrier whose release fence is paid once per CTA.
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

// Cluster-wide barrier with release/acquire semantics for the shared-memory traffic of the whole CTA, paid once:
// cooperative_groups' cluster.sync() makes every thread execute a cluster-scope release fence (measured: most of
// the kernel's time with 1024-thread CTAs); here the CTA barrier orders the CTA's writes before thread 0's single
// cumulative fence, and all threads then arrive relaxed and wait with acquire.
__device__ __forceinline__ void cluster_sync_once_fenced()
{
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("fence.acq_rel.cluster;" ::: "memory");
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Cluster-wide barrier after which every CTA sees the distributed-shared-memory stores and reductions that ANY thread of
// any CTA issued before it: each thread releases its own remote accesses (a single thread's fence, as above, orders what
// that thread has observed; stores other threads sent into another SM's shared memory may still be in flight).
__device__ __forceinline__ void cluster_sync_release_all()
{
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

}  // namespace qd
