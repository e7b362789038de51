// syndrome_kernel.cuh — bit-sliced H e mod 2 over 128-shot tiles (the HBM-bound utility kernel of the path).
//
// Computes what the reference does per shot with `H @ x % 2` (python/qldpc/misc/_experiment.py:47,58,124,
// python/qldpc/spacetime_code.py:109) on bit-packed, shot-major vectors; CPU restatement: oracle/qldpc_oracle.c
// orc_syndrome.  HBM traffic is the algorithmic minimum, 4 * (ceil(n/32) + ceil(m/32)) bytes per shot
// (SURVEY.md 8d); everything else is arranged so that the instruction count per shot stays near what
// the two bit transposes cost:
//
//   A. transpose in.  A lane owns a 32 x 32 bit block = (word w of the packed rows) x (32 shots of group g);
//      the 32 lanes of a warp are 8 consecutive words x the 4 shot groups of the tile, so a load
//      instruction reads 4 rows x 32 contiguous bytes and a store instruction hits 32 distinct banks.
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
    const uint4 *tbl;           // [Q][m]: byte offsets (slot * 16) of the check's columns inside X
    const uint32_t *vec; int vstride;
    long long shots;
    uint32_t *synd; int sstride;
};

// 32 x 32 bit transpose in registers (bit b of a[s] <-> bit s of a[b]): the 16- and 8-bit stages are byte
// permutes, the 4/2/1-bit stages a shift and a bit select per output word
__device__ __forceinline__ uint32_t bitsel(uint32_t x, uint32_t y, uint32_t m) { return (x & m) | (y & ~m); }

__device__ __forceinline__ void transpose32_prmt(uint32_t (&a)[32])
{
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
        a[k] = bitsel(x, y << 4, 0x0F0F0F0Fu);
        a[k + 4] = bitsel(x >> 4, y, 0x0F0F0F0Fu);
    }
#pragma unroll
    for (int k = 0; k < 32; k = (k + 3) & ~2) {
        const uint32_t x = a[k], y = a[k + 2];
        a[k] = bitsel(x, y << 2, 0x33333333u);
        a[k + 2] = bitsel(x >> 2, y, 0x33333333u);
    }
#pragma unroll
    for (int k = 0; k < 32; k += 2) {
        const uint32_t x = a[k], y = a[k + 1];
        a[k] = bitsel(x, y << 1, 0x55555555u);
        a[k + 1] = bitsel(x >> 1, y, 0x55555555u);
    }
}

__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// QT: compile-time Q (1..4), 0 = run time
template <int QT>
__global__ void __launch_bounds__(256, 4) syndrome128_kernel(const SynParams p)
{
    extern __shared__ uint4 syn_smem[];
    uint4 *X = syn_smem, *Y = syn_smem + p.xs;
    const int tid = (int)threadIdx.x, NT = (int)blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
    const int wl = lane >> 2, g = lane & 3;                 // lane = (word within an 8-word strip, shot group)
    const uint32_t xbase = smem_u32(X);
    const int Q = QT > 0 ? QT : p.Q;
    // slots no phase ever writes: the zero slot of X; the checks beyond m in the last word of Y
    if (tid == 0) X[p.xs - 1] = make_uint4(0u, 0u, 0u, 0u);
    for (int c = p.m + tid; c < 32 * p.nwm; c += NT) Y[c + (c >> 5)] = make_uint4(0u, 0u, 0u, 0u);

    const long long ntiles = (p.shots + 127) >> 7;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long s0 = tile << 7;
        const int ns = (int)min(128LL, p.shots - s0) - 32 * g;      // shots of this lane's group in the tile (may be <= 0)
        // ---- A. packed error rows -> X
        for (int w0 = warp * 8; w0 < p.nwn; w0 += nwarps * 8) {
            const int w = w0 + wl;
            if (w < p.nwn) {
                uint32_t a[32];
                const uint32_t *src = p.vec + (s0 + 32 * g) * (long long)p.vstride + w;
#pragma unroll
                for (int s = 0; s < 32; ++s) a[s] = s < ns ? __ldg(src + (long long)s * p.vstride) : 0u;
                transpose32_prmt(a);
                uint32_t *dst = reinterpret_cast<uint32_t *>(X + w * 33) + g;
#pragma unroll
                for (int b = 0; b < 32; ++b) dst[b * 4] = a[b];
            }
        }
        __syncthreads();
        // ---- B. Y[check] = XOR of the slots of its columns
        for (int c = tid; c < p.m; c += NT) {
            uint4 acc = make_uint4(0u, 0u, 0u, 0u);
            if (QT > 0) {
                uint4 t[QT > 0 ? QT : 1];
#pragma unroll
                for (int q = 0; q < QT; ++q) t[q] = __ldg(p.tbl + (size_t)q * p.m + c);
#pragma unroll
                for (int q = 0; q < QT; ++q) {
                    const uint4 v0 = lds128(xbase + t[q].x), v1 = lds128(xbase + t[q].y);
                    const uint4 v2 = lds128(xbase + t[q].z), v3 = lds128(xbase + t[q].w);
                    acc.x ^= v0.x ^ v1.x ^ v2.x ^ v3.x; acc.y ^= v0.y ^ v1.y ^ v2.y ^ v3.y;
                    acc.z ^= v0.z ^ v1.z ^ v2.z ^ v3.z; acc.w ^= v0.w ^ v1.w ^ v2.w ^ v3.w;
                }
            } else {
                for (int q = 0; q < Q; ++q) {
                    const uint4 t = __ldg(p.tbl + (size_t)q * p.m + c);
                    const uint4 v0 = lds128(xbase + t.x), v1 = lds128(xbase + t.y);
                    const uint4 v2 = lds128(xbase + t.z), v3 = lds128(xbase + t.w);
                    acc.x ^= v0.x ^ v1.x ^ v2.x ^ v3.x; acc.y ^= v0.y ^ v1.y ^ v2.y ^ v3.y;
                    acc.z ^= v0.z ^ v1.z ^ v2.z ^ v3.z; acc.w ^= v0.w ^ v1.w ^ v2.w ^ v3.w;
                }
            }
            Y[c + (c >> 5)] = acc;
        }
        __syncthreads();
        // ---- C. Y -> packed syndrome rows (stride padding words are written as zeros)
        for (int k0 = warp * 8; k0 < p.sstride; k0 += nwarps * 8) {
            const int k = k0 + wl;
            if (k < p.sstride) {
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
                uint32_t *dst = p.synd + (s0 + 32 * g) * (long long)p.sstride + k;
#pragma unroll
                for (int s = 0; s < 32; ++s)
                    if (s < ns) dst[(long long)s * p.sstride] = o[s];
            }
        }
        __syncthreads();
    }
}

}  // namespace qd
