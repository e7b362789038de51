// ssf_kernel.cuh — small-set-flip as an integer kernel (specification: SURVEY.md Appendix C;
// CPU restatement: oracle/qldpc_oracle.c orc_ssf_decode; the reference itself has no SSF).
//
// One CTA per shot in flight (persistent, shots pulled from a global atomic queue).  Per round:
//   * the best candidate of every generator is CACHED in shared memory; a flip only invalidates the
//     generators whose check window contains a check that changed (host-built check -> generators
//     lists), so after the first round only a few hundred generators are rescanned instead of all;
//   * the stale generators are compacted into a list; the syndrome restricted to a generator's check
//     window is gathered into one 64-bit word (one thread per stale generator);
//   * the rescans are spread evenly over the CTA as (generator, chunk) work items: a chunk fixes the
//     subset's bits above the lowest five and walks the 32 settings of the low five in a compile-time
//     Gray-code order (toggle masks in registers, no indexing), one XOR + two popcounts per candidate:
//     delta = 2 popc(local & T) - popc(T); chunk winners merge with a shared-memory atomicMax;
//   * candidates are ranked by an exact integer key  (delta * 27720 / |F|, lowest generator,
//     lowest subset mask)  — 27720 = lcm(1..12) makes the ratio an integer, so the arg-max is a
//     plain 64-bit max and ties break exactly like the serial oracle, whatever the evaluation order;
//   * warp shuffles + one shared-memory stage pick the CTA-wide winner, which is flipped.
// Integer arithmetic only => results are bit-identical to the oracle.
#pragma once

#include "common.cuh"

namespace qd {

constexpr int SSF_MAX_W = 12;

struct SsfDev {
    int m, n, nwm, nwn, mg, max_w;
    const int32_t *gen_ptr, *gen_idx;        // generator supports (rows of Hx)
    const int32_t *win_ptr, *win_chk;        // per generator: its window of adjacent checks (<= 64)
    const unsigned long long *qmask;         // per support entry: toggle mask in window coordinates
    const int32_t *chk_ptr, *chk_gen;        // per check: the generators whose window contains it
};

struct SsfParams {
    SsfDev s;
    int max_rounds;
    const uint32_t *synd; int synd_stride;
    long long shots;
    uint32_t *corr; int corr_stride;
    uint8_t *conv; int32_t *rounds;
    const uint32_t *err; int err_stride;
    const uint32_t *lzT; int k;
    uint8_t *fail;
    long long *counters;
    unsigned long long *next_shot;
};

__host__ __device__ inline size_t ssf_smem_bytes(int nwn, int nwm, int mg, int threads)
{
    return (size_t)8 * (threads / 32 + 2) + (size_t)16 * mg + (size_t)4 * (nwn + nwm) + (size_t)((mg + 15) & ~15) +
           (size_t)((2 * mg + 15) & ~15) + 32;
}

// 27720 / size for size in 1..12
__constant__ unsigned inv_size[13] = {0, 27720, 13860, 9240, 6930, 5544, 4620, 3960, 3465, 3080, 2772, 2520, 2310};

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other > v ? other : v;
    }
    return v;
}

__device__ __forceinline__ int popc_w(unsigned long long v) { return __popcll(v); }
__device__ __forceinline__ int popc_w(uint32_t v) { return __popc(v); }

// WT: the window word — uint32_t when every generator's check window has at most 32 checks ((3,4) HGP: 28),
// which halves the popcounts of the scan; unsigned long long otherwise
template <typename WT>
__global__ void __launch_bounds__(256) ssf_kernel(const SsfParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SsfDev &s = p.s;
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
    unsigned long long *red = reinterpret_cast<unsigned long long *>(smem_raw);   // [nwarps + 2]
    unsigned long long *gkeys = red + nwarps + 2;                                  // [mg] cached best key per generator
    unsigned long long *glocal = gkeys + s.mg;                                     // [mg] window syndrome of the d-th stale generator
    uint32_t *sigma = reinterpret_cast<uint32_t *>(glocal + s.mg);                 // [nwm]
    uint32_t *dec = sigma + s.nwm;                                                // [nwn]
    uint8_t *dirty = reinterpret_cast<uint8_t *>(dec + s.nwn);                    // [mg] cached key is stale
    uint16_t *dlist = reinterpret_cast<uint16_t *>(dirty + ((s.mg + 15) & ~15));  // [mg] the stale generators
    __shared__ long long sh_shot;
    __shared__ int sh_weight, sh_ndirty;
    constexpr int LOWB = 5;                                                       // subset bits walked inside a work item

    long long c_fail = 0, c_shots = 0, c_nc = 0, c_it = 0;

    for (;;) {
        if (tid == 0) {
            const unsigned long long id = atomicAdd(p.next_shot, 1ULL);
            sh_shot = id < (unsigned long long)p.shots ? (long long)id : -1;
            sh_weight = 0;
            sh_ndirty = 0;
        }
        __syncthreads();
        const long long shot = sh_shot;
        if (shot < 0) break;
        {
            const uint32_t *src = p.synd + shot * (long long)p.synd_stride;
            int w8 = 0;
            for (int w = tid; w < s.nwm; w += NT) { const uint32_t v = src[w]; sigma[w] = v; w8 += __popc(v); }
            w8 = warp_sum(w8);
            if (lane == 0 && w8) atomicAdd(&sh_weight, w8);
            for (int w = tid; w < s.nwn; w += NT) dec[w] = 0u;
            for (int g = tid; g < s.mg; g += NT) dirty[g] = 1;
        }
        __syncthreads();
        int weight = sh_weight;
        int rounds = 0;
        while (weight > 0 && (p.max_rounds == 0 || rounds < p.max_rounds)) {
            // ---- stale generators -> list (order irrelevant: keys carry the generator index)
            for (int g = tid; g < s.mg; g += NT)
                if (dirty[g]) { dirty[g] = 0; gkeys[g] = 0ULL; dlist[atomicAdd(&sh_ndirty, 1)] = (uint16_t)g; }
            __syncthreads();
            const int ndirty = sh_ndirty;
            // ---- their window syndromes
            for (int d = tid; d < ndirty; d += NT) {
                const int g = dlist[d];
                const int wlo = s.win_ptr[g], nw = s.win_ptr[g + 1] - wlo;
                unsigned long long local = 0ULL;
                for (int z = 0; z < nw; ++z) {
                    const int c = s.win_chk[wlo + z];
                    local |= (unsigned long long)((sigma[c >> 5] >> (c & 31)) & 1u) << z;
                }
                glocal[d] = local;
            }
            __syncthreads();
            if (tid == 0) sh_ndirty = 0;                        // next round's list (read again only after two barriers)
            // ---- rescans: work item = (stale generator d, chunk of its subsets); generators of different size
            // have different chunk counts, so items are numbered over the largest and the surplus skipped
            const int maxch = s.max_w > LOWB ? 1 << (s.max_w - LOWB) : 1;
            for (int x = tid; x < ndirty * maxch; x += NT) {
                const int d = x / maxch, chunk = x - d * maxch;
                const WT local = (WT)glocal[d];
                if (!local) continue;
                const int g = dlist[d];
                const int lo = s.gen_ptr[g], w = s.gen_ptr[g + 1] - lo;
                if (chunk >= (w > LOWB ? 1 << (w - LOWB) : 1)) continue;
                WT q0 = 0, q1 = 0, q2 = 0, q3 = 0, q4 = 0, T = 0;
                if (w > 0) q0 = (WT)s.qmask[lo];
                if (w > 1) q1 = (WT)s.qmask[lo + 1];
                if (w > 2) q2 = (WT)s.qmask[lo + 2];
                if (w > 3) q3 = (WT)s.qmask[lo + 3];
                if (w > 4) q4 = (WT)s.qmask[lo + 4];
                for (int t = LOWB; t < w; ++t)
                    if ((chunk >> (t - LOWB)) & 1) T ^= (WT)s.qmask[lo + t];
                const unsigned hi_mask = (unsigned)chunk << LOWB;
                const int hi_size = __popc((unsigned)chunk);
                const unsigned low_lim = w < LOWB ? 1u << w : 1u << LOWB;      // low settings that stay inside the support
                const unsigned long long gkey = (unsigned long long)(0xFFFFFFu - (unsigned)g) << 16;
                // inside a work item the generator and the high subset bits are fixed, so candidates compare by
                // (delta * 27720 / |F|, lowest low setting): a 32-bit key  score * 32 + (31 - gray), three
                // instructions per candidate; the 64-bit key is formed once for the winner.  31 = "nothing positive".
                int iv[LOWB + 1];
#pragma unroll
                for (int t = 0; t <= LOWB; ++t) iv[t] = (int)inv_size[min(hi_size + t, 12)] * 32;
                int best32 = 31;
                unsigned gray = 0u;
#pragma unroll
                for (unsigned i = 0; i < (1u << LOWB); ++i) {
                    if (i > 0) {
                        const int b = (i & 1u) ? 0 : (i & 2u) ? 1 : (i & 4u) ? 2 : (i & 8u) ? 3 : 4;   // compile time
                        gray ^= 1u << b;
                        T ^= b == 0 ? q0 : b == 1 ? q1 : b == 2 ? q2 : b == 3 ? q3 : q4;
                    }
                    const unsigned mask = hi_mask | gray;
                    if (mask == 0u || gray >= low_lim) continue;
                    const int delta = 2 * popc_w((WT)(local & T)) - popc_w(T);
                    best32 = max(best32, delta * iv[__popc(gray)] + (31 - (int)gray));
                }
                unsigned long long gbest = 0ULL;
                if (best32 > 31) {
                    const unsigned wg = 31u - ((unsigned)best32 & 31u);
                    gbest = ((unsigned long long)((unsigned)best32 >> 5) << 40) | gkey | (unsigned long long)(0xFFFFu - (hi_mask | wg));
                }
                if (gbest) atomicMax(&gkeys[g], gbest);
            }
            __syncthreads();
            unsigned long long best = 0ULL;
            for (int g = tid; g < s.mg; g += NT) { const unsigned long long k = gkeys[g]; best = k > best ? k : best; }
            best = warp_max_u64(best);
            if (lane == 0) red[warp] = best;
            __syncthreads();
            if (warp == 0) {
                unsigned long long v = lane < nwarps ? red[lane] : 0ULL;
                v = warp_max_u64(v);
                if (lane == 0) red[nwarps] = v;
            }
            __syncthreads();
            const unsigned long long win = red[nwarps];
            if (win == 0ULL) break;                              // no subset lowers the syndrome weight
            const int g = (int)(0xFFFFFFu - (unsigned)((win >> 16) & 0xFFFFFFu));
            const unsigned mask = 0xFFFFu - (unsigned)(win & 0xFFFFu);
            const int size = __popc(mask);
            const int delta = (int)((unsigned)(win >> 40) / inv_size[size]);
            const int lo = s.gen_ptr[g], wlo = s.win_ptr[g], nw = s.win_ptr[g + 1] - wlo;
            if (tid < SSF_MAX_W && ((mask >> tid) & 1u)) {
                const int q = s.gen_idx[lo + tid];
                atomicXor(&dec[q >> 5], 1u << (q & 31));
            }
            unsigned long long T = 0ULL;
            for (int t = 0; t < SSF_MAX_W; ++t)
                if ((mask >> t) & 1u) T ^= s.qmask[lo + t];
            if (tid >= 32 && tid - 32 < nw && ((T >> (tid - 32)) & 1ULL)) {
                const int c = s.win_chk[wlo + tid - 32];
                atomicXor(&sigma[c >> 5], 1u << (c & 31));
                for (int z = s.chk_ptr[c]; z < s.chk_ptr[c + 1]; ++z) dirty[s.chk_gen[z]] = 1;
            }
            weight -= delta;
            ++rounds;
            __syncthreads();
        }

        // ---- retire the shot
        uint32_t *dst = p.corr + shot * (long long)p.corr_stride;
        for (int w = tid; w < p.corr_stride; w += NT) dst[w] = w < s.nwn ? dec[w] : 0u;
        int f = 0;
        if (p.err) {
            const uint32_t *e = p.err + shot * (long long)p.err_stride;
            __syncthreads();
            for (int w = tid; w < s.nwn; w += NT) dec[w] ^= e[w];
            __syncthreads();
            for (int i = tid; i < p.k; i += NT) {
                uint32_t acc = 0u;
                for (int w = 0; w < s.nwn; ++w) acc ^= p.lzT[(size_t)w * p.k + i] & dec[w];
                f |= __popc(acc) & 1;
            }
            f = __syncthreads_or(f);
        }
        if (tid == 0) {
            const int conv = weight == 0;
            if (p.conv) p.conv[shot] = (uint8_t)conv;
            if (p.rounds) p.rounds[shot] = rounds;
            if (p.fail) p.fail[shot] = (uint8_t)f;
            c_fail += f; c_shots += 1; c_nc += !conv; c_it += rounds;
        }
        __syncthreads();
    }
    if (tid == 0 && p.counters && c_shots) {
        atomicAdd((unsigned long long *)&p.counters[0], (unsigned long long)c_fail);
        atomicAdd((unsigned long long *)&p.counters[1], (unsigned long long)c_shots);
        atomicAdd((unsigned long long *)&p.counters[2], (unsigned long long)c_nc);
        atomicAdd((unsigned long long *)&p.counters[3], (unsigned long long)c_it);
    }
}

}  // namespace qd
