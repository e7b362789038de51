// bp_kernel.cuh — flooding belief propagation, messages resident in shared memory.
//
// What it computes is the syndrome-BP of ldpc.bp_decoder as called by the reference at
// python/qldpc/misc/_experiment.py:110-117,137-149 (semantics: SURVEY.md Appendix B;
// CPU restatement: oracle/qldpc_oracle.c).  How it computes it is B200-specific:
//
//   * persistent CTAs, each keeping S shots ("slots") in flight; a finished slot (converged or
//     out of iterations) is written out and immediately refilled from a global atomic queue, so
//     lanes never idle on the wide spread of per-shot iteration counts;
//   * ONE in-place message array per slot in shared memory: the check pass reads bit->check
//     messages of its row and overwrites them with check->bit messages; the bit pass reads the
//     check->bit messages of its column and overwrites them with bit->check messages.  The same
//     thread owns a whole row (resp. column), so no second array and no intra-phase hazards;
//   * row-major padded layout (odd row pitch) => the check pass is bank-conflict free;
//   * hard decisions are ballot-packed one word per warp; the residual syndrome s ^ H e_hat is
//     maintained incrementally (only bits whose decision flipped touch it) together with a count
//     of unsatisfied checks, so the convergence test of every iteration is one shared-memory read;
//   * fp64 path: every operation is an explicitly rounded __d*_rn intrinsic issued in exactly the
//     order of the serial reference (prefix pass then suffix pass), so hard decisions, iteration
//     counts and soft outputs are bit-identical to the CPU oracle.  fp32 path: same structure,
//     half the shared-memory traffic, not bit-reproducible.
#pragma once

#include "common.cuh"

namespace qd {

struct GraphDev {
    int m, n, nnz;
    int nwm, nwn;                 // packed words per syndrome / per error vector
    int npos;                     // message positions per slot (rows padded to `pitch`)
    int max_row_deg, max_col_deg;
    const uint32_t *row_desc;     // [m]   position of the row's first message | degree << 24
    const uint32_t *col_desc;     // [n]   offset into csc_* | degree << 24
    const uint32_t *csc_pos;      // [nnz] message position of the t-th edge of a column (rows ascending)
    const uint32_t *csc_row;      // [nnz] check index of that edge
    const int32_t *row_ptr;       // CSR (syndrome kernel)
    const int32_t *col_idx;
};

struct BpParams {
    GraphDev g;
    const void *prior;            // real[n]: log((1-p)/p) for min-sum, p/(1-p) for product-sum
    const void *alpha;            // real[max_iter + 1]: scaling factor of iteration it (min-sum)
    int max_iter;
    int slots;
    const uint32_t *synd; int synd_stride;
    long long shots;
    uint32_t *corr; int corr_stride;
    uint8_t *conv; int32_t *iters; double *llr;
    const uint32_t *err; int err_stride;
    const uint32_t *lzT; int k;   // logical operators, packed and transposed: [nwn][k]
    uint8_t *fail;
    long long *counters;          // {failures, shots, non_converged, iterations}
    unsigned long long *next_shot;
};

template <typename T> struct Ar;
template <> struct Ar<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double big() { return 1e308; }
    static __device__ __forceinline__ double lg(double a) { return log(a); }
};
template <> struct Ar<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float big() { return 3.0e38f; }
    static __device__ __forceinline__ float lg(float a) { return logf(a); }
};

struct BpCtl {              // per-slot control words, shared memory
    long long shot;         // global shot index, -1 = slot idle
    int it;                 // iterations completed
    int unsat;              // weight of the residual syndrome s ^ H e_hat
    int fresh;              // slot holds a newly fetched shot that is not loaded yet
    int pad;
};

constexpr int QD_BP_MIN_SUM_K = 3;

// shared-memory bytes for a launch
template <typename real>
__host__ __device__ inline size_t bp_smem_bytes(int npos, int nwn, int nwm, int slots)
{
    size_t msg = (size_t)slots * npos * sizeof(real);
    msg = (msg + 15) & ~(size_t)15;
    size_t bits = ((size_t)slots * (nwn + 2 * nwm) * 4 + 7) & ~(size_t)7;
    return msg + bits + (size_t)slots * sizeof(BpCtl);
}

template <typename real, int METHOD, int VMAX, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) bp_kernel(const BpParams p)
{
    using A = Ar<real>;
    constexpr int DMAX = (VMAX == 4) ? 8 : 16;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GraphDev &g = p.g;
    const int S = p.slots;
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31;

    real *msg = reinterpret_cast<real *>(smem_raw);
    size_t off = ((size_t)S * g.npos * sizeof(real) + 15) & ~(size_t)15;
    uint32_t *dec = reinterpret_cast<uint32_t *>(smem_raw + off);        // [S][nwn]
    uint32_t *syn = dec + (size_t)S * g.nwn;                            // [S][nwm] original syndrome
    uint32_t *res = syn + (size_t)S * g.nwm;                            // [S][nwm] s ^ H e_hat
    BpCtl *ctl = reinterpret_cast<BpCtl *>(
        smem_raw + off + (((size_t)S * (g.nwn + 2 * g.nwm) * 4 + 7) & ~(size_t)7));

    const real *prior = static_cast<const real *>(p.prior);
    const real *alpha_tab = static_cast<const real *>(p.alpha);

    long long c_fail = 0, c_shots = 0, c_nc = 0, c_it = 0;   // meaningful on thread 0 only

    if (tid < S) {
        const unsigned long long id = atomicAdd(p.next_shot, 1ULL);
        ctl[tid].shot = id < (unsigned long long)p.shots ? (long long)id : -1;
        ctl[tid].it = 0;
        ctl[tid].unsat = 0;
        ctl[tid].fresh = 1;
    }
    __syncthreads();

    const int n32 = (g.n + 31) & ~31;

    for (;;) {
        // ------------------------------------------------------------------ load fresh slots
        bool any_active = false;
        for (int s = 0; s < S; ++s) any_active |= ctl[s].shot >= 0;
        if (!any_active) break;
        for (int s = 0; s < S; ++s) {
            if (!ctl[s].fresh || ctl[s].shot < 0) continue;
            const uint32_t *src = p.synd + ctl[s].shot * (long long)p.synd_stride;
            int w8 = 0;
            for (int w = tid; w < g.nwm; w += NT) {
                const uint32_t v = src[w];
                syn[s * g.nwm + w] = v;
                res[s * g.nwm + w] = v;
                w8 += __popc(v);
            }
            if (tid < ((g.nwm + 31) & ~31)) {     // warps that may hold syndrome words
                w8 = warp_sum(w8);
                if (lane == 0 && w8) atomicAdd(&ctl[s].unsat, w8);
            }
            for (int w = tid; w < g.nwn; w += NT) dec[s * g.nwn + w] = 0u;
            real *M = msg + (size_t)s * g.npos;
            for (int j = tid; j < g.n; j += NT) {
                const uint32_t d = g.col_desc[j];
                const int o = d & 0xFFFFFF, deg = d >> 24;
                const real pj = prior[j];
                for (int t = 0; t < deg; ++t) M[g.csc_pos[o + t]] = pj;
            }
        }
        __syncthreads();
        if (tid < S) ctl[tid].fresh = 0;

        // ------------------------------------------------------------------ iterate
        for (;;) {
            // ---- check pass: row messages bit->check  ==>  check->bit, in place
            for (int c = tid; c < g.m; c += NT) {
                const uint32_t d = g.row_desc[c];
                const int pos0 = d & 0xFFFFFF, deg = d >> 24;
                for (int s = 0; s < S; ++s) {
                    if (ctl[s].shot < 0) continue;
                    real *M = msg + (size_t)s * g.npos + pos0;
                    const unsigned sbit = (syn[s * g.nwm + (c >> 5)] >> (c & 31)) & 1u;
                    if (METHOD == QD_BP_MIN_SUM_K) {
                        real min1 = A::big(), min2 = A::big();
                        int arg = -1;
                        unsigned long long neg = 0ULL;
                        unsigned par = sbit;
                        for (int k = 0; k < deg; ++k) {
                            const real v = M[k];
                            const real a = fabs(v);
                            const unsigned isneg = v <= (real)0 ? 1u : 0u;
                            neg |= (unsigned long long)isneg << (k & 63);
                            par ^= isneg;
                            if (a < min1) { min2 = min1; min1 = a; arg = k; }
                            else if (a < min2) { min2 = a; }
                        }
                        const real al = alpha_tab[ctl[s].it + 1];
                        const real nal = -al;
                        if (deg <= 64) {
                            for (int k = 0; k < deg; ++k) {
                                const real mn = (k == arg) ? min2 : min1;
                                const unsigned sg = par ^ (unsigned)((neg >> k) & 1ULL);
                                M[k] = A::mul(mn, sg ? nal : al);
                            }
                        } else {
                            for (int k = 0; k < deg; ++k) {
                                const real v = M[k];
                                const real mn = (k == arg) ? min2 : min1;
                                const unsigned sg = par ^ (v <= (real)0 ? 1u : 0u);
                                M[k] = A::mul(mn, sg ? nal : al);
                            }
                        }
                    } else {
                        // product-sum, ratio domain: prefix products left in place, then suffix pass
                        real t[DMAX];
                        real temp = sbit ? (real)-1 : (real)1;
#pragma unroll
                        for (int k = 0; k < DMAX; ++k) {
                            if (k < deg) {
                                const real v = M[k];
                                t[k] = A::sub(A::div((real)2, A::add((real)1, v)), (real)1);
                                M[k] = temp;
                                temp = A::mul(temp, t[k]);
                            }
                        }
                        temp = (real)1;
#pragma unroll
                        for (int k = DMAX - 1; k >= 0; --k) {
                            if (k < deg) {
                                const real x = A::mul(M[k], temp);
                                M[k] = A::div(A::sub((real)1, x), A::add((real)1, x));
                                temp = A::mul(temp, t[k]);
                            }
                        }
                    }
                }
            }
            __syncthreads();

            // ---- bit pass: column messages check->bit  ==>  bit->check, in place; decisions
            for (int j = tid; j < n32; j += NT) {
                const bool valid = j < g.n;
                int deg = 0, o = 0;
                real pj = (real)0;
                int pos[VMAX];
                if (valid) {
                    const uint32_t d = g.col_desc[j];
                    o = d & 0xFFFFFF; deg = d >> 24;
                    pj = prior[j];
#pragma unroll
                    for (int t = 0; t < VMAX; ++t) pos[t] = t < deg ? (int)g.csc_pos[o + t] : 0;
                }
                for (int s = 0; s < S; ++s) {
                    if (ctl[s].shot < 0) continue;
                    real *M = msg + (size_t)s * g.npos;
                    unsigned dbit = 0u;
                    if (valid) {
                        real cv[VMAX], pre[VMAX];
                        real temp = pj;
                        if (METHOD == QD_BP_MIN_SUM_K) {
#pragma unroll
                            for (int t = 0; t < VMAX; ++t)
                                if (t < deg) { cv[t] = M[pos[t]]; pre[t] = temp; temp = A::add(temp, cv[t]); }
                            dbit = temp <= (real)0 ? 1u : 0u;
                            if (p.llr) p.llr[ctl[s].shot * (long long)g.n + j] = (double)temp;
                            real suf = (real)0;
#pragma unroll
                            for (int t = VMAX - 1; t >= 0; --t)
                                if (t < deg) { M[pos[t]] = A::add(pre[t], suf); suf = A::add(suf, cv[t]); }
                        } else {
#pragma unroll
                            for (int t = 0; t < VMAX; ++t)
                                if (t < deg) {
                                    cv[t] = M[pos[t]]; pre[t] = temp; temp = A::mul(temp, cv[t]);
                                    if (temp != temp) temp = (real)1;
                                }
                            dbit = temp >= (real)1 ? 1u : 0u;
                            if (p.llr) p.llr[ctl[s].shot * (long long)g.n + j] = (double)A::lg(A::div((real)1, temp));
                            real suf = (real)1;
#pragma unroll
                            for (int t = VMAX - 1; t >= 0; --t)
                                if (t < deg) {
                                    M[pos[t]] = A::mul(pre[t], suf); suf = A::mul(suf, cv[t]);
                                    if (suf != suf) suf = (real)1;
                                }
                        }
                    }
                    // one decision word per warp (j is warp-aligned): flips drive the residual syndrome
                    const uint32_t neww = __ballot_sync(0xffffffffu, dbit);
                    uint32_t *dw = dec + s * g.nwn + (j >> 5);
                    const uint32_t oldw = *dw;   // j < n32 => word exists; warp-wide broadcast
                    const uint32_t flips = oldw ^ neww;
                    if (flips) {                                  // warp-uniform
                        int delta = 0;
                        if ((flips >> lane) & 1u) {
                            for (int t = 0; t < deg; ++t) {
                                const uint32_t r = g.csc_row[o + t];
                                const uint32_t bit = 1u << (r & 31);
                                const uint32_t old = atomicXor(&res[s * g.nwm + (r >> 5)], bit);
                                delta += (old & bit) ? -1 : 1;
                            }
                        }
                        delta = warp_sum(delta);
                        __syncwarp();
                        if (lane == 0) {
                            *dw = neww;
                            if (delta) atomicAdd(&ctl[s].unsat, delta);
                        }
                    }
                }
            }
            if (tid < S && ctl[tid].shot >= 0) ctl[tid].it += 1;
            __syncthreads();

            bool any_done = false;
            for (int s = 0; s < S; ++s)
                any_done |= ctl[s].shot >= 0 && (ctl[s].unsat == 0 || ctl[s].it >= p.max_iter);
            if (any_done) break;
        }

        // ------------------------------------------------------------------ retire finished slots
        for (int s = 0; s < S; ++s) {
            const long long shot = ctl[s].shot;
            if (shot < 0) continue;
            const int conv = ctl[s].unsat == 0;
            if (!conv && ctl[s].it < p.max_iter) continue;
            uint32_t *D = dec + s * g.nwn;
            uint32_t *dst = p.corr + shot * (long long)p.corr_stride;
            for (int w = tid; w < p.corr_stride; w += NT) dst[w] = w < g.nwn ? D[w] : 0u;
            int f = 0;
            if (p.err) {
                const uint32_t *e = p.err + shot * (long long)p.err_stride;
                __syncthreads();
                for (int w = tid; w < g.nwn; w += NT) D[w] ^= e[w];
                __syncthreads();
                for (int i = tid; i < p.k; i += NT) {
                    uint32_t acc = 0u;
                    for (int w = 0; w < g.nwn; ++w) acc ^= p.lzT[(size_t)w * p.k + i] & D[w];
                    f |= __popc(acc) & 1;
                }
                f = __syncthreads_or(f);
            }
            if (tid == 0) {
                if (p.conv) p.conv[shot] = (uint8_t)conv;
                if (p.iters) p.iters[shot] = ctl[s].it;
                if (p.fail) p.fail[shot] = (uint8_t)f;
                c_fail += f; c_shots += 1; c_nc += !conv; c_it += ctl[s].it;
            }
        }
        __syncthreads();
        if (tid < S) {
            const bool done = ctl[tid].shot >= 0 && (ctl[tid].unsat == 0 || ctl[tid].it >= p.max_iter);
            if (done) {
                const unsigned long long id = atomicAdd(p.next_shot, 1ULL);
                ctl[tid].shot = id < (unsigned long long)p.shots ? (long long)id : -1;
                ctl[tid].it = 0;
                ctl[tid].unsat = 0;
                ctl[tid].fresh = 1;
            }
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
