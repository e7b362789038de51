// syndrome_kernel.cuh — bit-sliced H e mod 2 over 128-shot tiles (the HBM-bound utility kernel of the path).
//
// Computes what the reference does per shot with `H @ x % 2` (python/qldpc/misc/_experiment.py:47,58,124,
// python/qldpc/spacetime_code.py:109) on bit-packed, shot-major vectors; CPU restatement: oracle/qldpc_oracle.c
// orc_syndrome.  HBM traffic is the algorithmic minimum, 4 * (ceil(n/32) + ceil(m/32)) bytes per shot
// (SURVEY.md 8d); everything else is arranged so that the instruction count per shot stays near what
// the two bit transposes cost:
//
//   0. the packed rows of a 128-shot tile are ONE contiguous piece of global memory: a single cp.async.bulk
//      (TMA) brings them into a staging buffer, armed on an mbarrier and issued a whole phase ahead of its
//      use, so HBM latency never sits on the critical path and no lane spends instructions on addresses;
//   A. transpose in.  A lane owns a 32 x 32 bit block = (word w of the packed rows) x (32 shots of group g);
//      the 32 lanes of a warp are 8 consecutive words x the 4 shot groups of the tile, so a store
//      instruction hits 32 distinct banks.
//      The block is transposed in registers (two byte-permute stages + three shift/select stages: 256
//      operations instead of the 480 of the xor-swap form) and stored as word g of the 16-byte slot of
//      each of its 32 columns: X[column] = the column's bit for the 128 shots of the tile.
//   B. gather.  A thread owns a check and XORs the slots of its columns with 128-bit loads — one load
//      and four XORs per edge and 128 shots.  Slot offsets come from a host-built table (4 edges per
//      uint4, coalesced over the checks; missing edges point at an all-zero slot, so nothing is predicated).
//   C. transpose out, the mirror image of A: Y[check] slots -> 32 x 32 blocks -> packed syndrome rows.
//
// Slots are laid out with a pitch of 33 per 32 columns, which makes the stores of A and the loads of C
// conflict free; the gathers of B hit random banks (a quarter warp per wavefront).
#pragma once

#include "common.cuh"

namespace qd {

struct SynParams {
    int n, m, nwn, nwm;
    int xs, ys;                 // 16-byte slots of X (33 * nwn + 1, the last one is the zero slot) and of Y (33 * nwm)
    int Q;                      // uint4 table entries per check = ceil(max row degree / 4)
    const uint4 *tbl;           // [Q][m]: byte offsets (slot * 16) of the check's columns inside X, 4 edges per entry
    const uint32_t *vec; int vstride;
    long long shots;
    uint32_t *synd; int sstride;
};

// 32 x 32 bit transpose in registers (bit b of a[s] <-> bit s of a[b]): the 16- and 8-bit stages are byte
// permutes, the 4/2/1-bit stages a shift and ONE three-input logic operation per output word (the mask sits in
// a register: with an immediate mask the compiler needs two)
__device__ __forceinline__ uint32_t bitsel(uint32_t x, uint32_t y, uint32_t m)       // (x & m) | (y & ~m)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(x), "r"(y), "r"(m));
    return d;
}

__device__ __forceinline__ void transpose32_prmt(uint32_t (&a)[32])
{
    uint32_t m4 = 0x0F0F0F0Fu, m2 = 0x33333333u, m1 = 0x55555555u;
    asm volatile("" : "+r"(m4), "+r"(m2), "+r"(m1));
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const uint32_t x = a[k], y = a[k + 16];
        a[k] = __byte_perm(x, y, 0x5410);
        a[k + 16] = __byte_perm(x, y, 0x7632);
    }
#pragma unroll
    for (int k = 0; k < 32; k = (k + 9) & ~8) {
        const uint32_t x = a[k], y = a[k + 8];
        a[k] = __byte_perm(x, y, 0x6240);
        a[k + 8] = __byte_perm(x, y, 0x7351);
    }
#pragma unroll
    for (int k = 0; k < 32; k = (k + 5) & ~4) {
        const uint32_t x = a[k], y = a[k + 4];
        a[k] = bitsel(x, y << 4, m4);
        a[k + 4] = bitsel(x >> 4, y, m4);
    }
#pragma unroll
    for (int k = 0; k < 32; k = (k + 3) & ~2) {
        const uint32_t x = a[k], y = a[k + 2];
        a[k] = bitsel(x, y << 2, m2);
        a[k + 2] = bitsel(x >> 2, y, m2);
    }
#pragma unroll
    for (int k = 0; k < 32; k += 2) {
        const uint32_t x = a[k], y = a[k + 1];
        a[k] = bitsel(x, y << 1, m1);
        a[k + 1] = bitsel(x >> 1, y, m1);
    }
}

// ---- TMA bulk copy of a tile's packed rows (one contiguous piece of global memory) + mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // earlier generic reads of dst are through
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}

// TMA: the packed rows of a tile arrive through one cp.async.bulk into shared memory, issued a whole phase
//      ahead of their use (possible when the rows are 16-byte aligned and the staging buffer fits), else the
//      lanes load them from global memory themselves.
// QT:  compile-time number of uint4 table entries per check (1..3), 0 = run time.
// Shot (4 s + g) of a tile is bit s of group g (interleaved groups spread the staging reads over the banks).
template <bool TMA, int QT>
__global__ void __launch_bounds__(320, 3) syndrome128_kernel(const SynParams p)
{
    extern __shared__ uint4 syn_smem[];
    uint4 *X = syn_smem, *Y = syn_smem + p.xs;
    const uint32_t *RAW = reinterpret_cast<const uint32_t *>(Y + p.ys);   // [128][vstride] (TMA only)
    __shared__ __align__(8) unsigned long long sh_bar;
    const int tid = (int)threadIdx.x, NT = (int)blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
    const int wl = lane >> 2, g = lane & 3;                 // lane = (word within an 8-word strip, shot group)
    const unsigned char *Xb = reinterpret_cast<const unsigned char *>(X);
    const uint32_t bar = smem_u32(&sh_bar), raw_s = smem_u32(RAW);
    const int Q = QT > 0 ? QT : p.Q;
    // once per CTA: the slots no phase ever writes (zero slot of X, checks beyond m in the last word of Y)
    if (tid == 0) X[p.xs - 1] = make_uint4(0u, 0u, 0u, 0u);
    for (int c = p.m + tid; c < 32 * p.nwm; c += NT) Y[c + (c >> 5)] = make_uint4(0u, 0u, 0u, 0u);
    if (TMA && tid == 0) mbar_init(bar, 1);
    __syncthreads();

    const long long ntiles = (p.shots + 127) >> 7;
    const int nA = (p.nwn + 7) >> 3, nC = (p.sstride + 7) >> 3;       // 8-word strips of the packed rows in / out
    auto rows_of = [&](long long tile) -> int { return (int)min(128LL, p.shots - (tile << 7)); };
    auto fetch = [&](long long tile) {                                 // thread 0: start the tile's bulk load
        bulk_load(raw_s, p.vec + (tile << 7) * (long long)p.vstride, (uint32_t)rows_of(tile) * (uint32_t)p.vstride * 4u, bar);
    };
    auto slot = [&](uint32_t off) -> uint4 { return *reinterpret_cast<const uint4 *>(Xb + off); };
    auto fold4 = [&](uint4 &acc, const uint4 &v0, const uint4 &v1, const uint4 &v2, const uint4 &v3) {
        acc.x ^= v0.x ^ v1.x ^ v2.x ^ v3.x; acc.y ^= v0.y ^ v1.y ^ v2.y ^ v3.y;
        acc.z ^= v0.z ^ v1.z ^ v2.z ^ v3.z; acc.w ^= v0.w ^ v1.w ^ v2.w ^ v3.w;
    };

    // in-strips start behind the out-strips in the warps' round robin, except for the CTA's first tile
    // (nothing to write out yet)
    const int c_shift = nC % nwarps;
    uint32_t parity = 0u;
    if (TMA && tid == 0 && (long long)blockIdx.x < ntiles) fetch(blockIdx.x);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long s0 = tile << 7;
        const int rows = rows_of(tile);
        const int c_strips = tile == (long long)blockIdx.x ? 0 : c_shift;
        const int ns = (rows - g + 3) >> 2;                           // shots of this lane's group: 4 s + g < rows
        // ---- A. packed error rows -> X.  The strips (8 words) of both transposes are dealt out together: the out-strips of
        // the previous tile come first in the list, so a CTA with as many warps as strips gives every warp one strip
        if (TMA) { mbar_wait(bar, parity); parity ^= 1u; }
        for (int t = warp - c_strips; t < nA; t += nwarps) {
            if (t < 0) continue;
            const int w = t * 8 + wl;
            if (w >= p.nwn) continue;
            uint32_t a[32];
            if (TMA) {
                const uint32_t *src = RAW + g * p.vstride + w;
                const int step = 4 * p.vstride;
                if (rows == 128) {
#pragma unroll
                    for (int s = 0; s < 32; ++s) { a[s] = *src; src += step; }
                } else {
#pragma unroll
                    for (int s = 0; s < 32; ++s) { a[s] = s < ns ? *src : 0u; src += step; }
                }
            } else {
                const uint32_t *src = p.vec + (s0 + g) * (long long)p.vstride + w;
                const long long step = 4LL * p.vstride;
#pragma unroll
                for (int s = 0; s < 32; ++s) { a[s] = s < ns ? __ldg(src) : 0u; src += step; }
            }
            transpose32_prmt(a);
            uint32_t *dst = reinterpret_cast<uint32_t *>(X + w * 33) + g;
#pragma unroll
            for (int b = 0; b < 32; ++b) dst[b * 4] = a[b];
        }
        __syncthreads();
        // the staging buffer is free: the CTA's next tile starts to arrive while this one is worked on
        if (TMA && tid == 0 && tile + gridDim.x < ntiles) fetch(tile + gridDim.x);
        // ---- B. Y[check] = XOR of the slots of its columns
        for (int c = tid; c < p.m; c += NT) {
            uint4 acc = make_uint4(0u, 0u, 0u, 0u);
            if (QT > 0) {
                uint4 t[QT > 0 ? QT : 1];
#pragma unroll
                for (int q = 0; q < QT; ++q) t[q] = __ldg(p.tbl + (size_t)q * p.m + c);
#pragma unroll
                for (int q = 0; q < QT; ++q) fold4(acc, slot(t[q].x), slot(t[q].y), slot(t[q].z), slot(t[q].w));
            } else {
                for (int q = 0; q < Q; ++q) {
                    const uint4 t = __ldg(p.tbl + (size_t)q * p.m + c);
                    fold4(acc, slot(t.x), slot(t.y), slot(t.z), slot(t.w));
                }
            }
            Y[c + (c >> 5)] = acc;
        }
        __syncthreads();
        // ---- C. Y -> packed syndrome rows (stride padding words are written as zeros)
        for (int t = warp; t < nC; t += nwarps) {
            const int k = t * 8 + wl;
            if (k >= p.sstride) continue;
            uint32_t o[32];
            if (k < p.nwm) {
                const uint32_t *src = reinterpret_cast<const uint32_t *>(Y + k * 33) + g;
#pragma unroll
                for (int b = 0; b < 32; ++b) o[b] = src[b * 4];
                transpose32_prmt(o);
            } else {
#pragma unroll
                for (int b = 0; b < 32; ++b) o[b] = 0u;
            }
            uint32_t *dst = p.synd + (s0 + g) * (long long)p.sstride + k;
            const long long step = 4LL * p.sstride;
            if (rows == 128) {
#pragma unroll
                for (int s = 0; s < 32; ++s) { *dst = o[s]; dst += step; }
            } else {
#pragma unroll
                for (int s = 0; s < 32; ++s) { if (s < ns) *dst = o[s]; dst += step; }
            }
        }
        // no barrier: the next tile's phase A writes X (free since B), its phase B is behind a barrier of its own
    }
}

}  // namespace qd
