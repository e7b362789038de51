// bpg_kernel.cuh — belief propagation with the messages in global memory (L2-resident scratch).
//
// Fallback for Tanner graphs that the shared-memory kernels cannot take: more messages than one SM's
// shared memory holds (spacetime matrices of the large codes, circuit-noise detector error models)
// or degrees beyond their limits.  Same decoder (ldpc.bp_decoder semantics, SURVEY.md Appendix B;
// reference call sites python/qldpc/misc/_experiment.py:110-117,137-149), same operation order as the
// serial restatement (oracle/qldpc_oracle.c: bp_min_sum_log / bp_prob_ratios), so fp64 results are
// bit-identical to it; slower than the shared-memory kernels.
//
// One CTA per shot in flight (persistent, atomic shot queue).  Two message arrays per CTA in a global
// scratch buffer, CSR edge order: b2c (bit -> check) and c2b (check -> bit).
//   * Min-sum check pass: one WARP per row, lanes strided over the row's (contiguous) edges, all loads
//     issued before the first use.  The serial prefix/suffix sweeps give edge k the minimum of the other
//     edges' magnitudes and the parity of the other edges' signs; both are exact and order-free, so the
//     warp computes smallest / second-smallest magnitude (+ position of the smallest) and the sign
//     parity with shuffles and every lane writes its edges — detector-error-model rows have hundreds
//     of edges, a thread per row spent the whole iteration on serial L2 round trips.
//   * Product-sum check pass: one thread per row, the two literal sweeps (ordered products).
//   * Bit pass: one thread per column; columns
//     of degree <= 8 load all their messages first and keep the prefix sums in registers (one gather + one
//     scatter; tiles of 4 / 8), longer ones run the two literal sweeps.
//   * Hard decisions (bytes) and the syndrome live in shared memory; convergence is tested exactly after
//     every iteration, one warp per row.
#pragma once

#include "bp_kernel.cuh"

namespace qd {

struct BpgParams {
    int m, n, nnz, nwm, nwn;
    const int32_t *row_ptr, *col_idx;       // CSR
    const int32_t *col_ptr, *csc_edge;      // CSC offsets; CSR edge id of every CSC entry (rows ascending)
    const void *prior;                      // real[n], decoder domain
    double alpha_fixed;
    int max_iter;
    void *scratch;                          // gridDim.x * 2 * nnz reals
    const uint32_t *synd; int synd_stride;
    long long shots;
    uint32_t *corr; int corr_stride;
    uint8_t *conv; int32_t *iters; double *llr; int llr_mode;
    const uint32_t *err; int err_stride;
    const uint32_t *lzT; int k;
    uint8_t *fail;
    long long *counters;
    unsigned long long *next_shot;
};

constexpr int BPG_RT = 8;      // rows up to 32 * BPG_RT edges keep their messages in registers during the check update

// merge (smallest, second smallest, position of the smallest) over the warp; equal magnitudes may pick either
// position: the second smallest then equals the smallest, so every edge receives the same value either way
template <typename real>
__device__ __forceinline__ void warp_two_min(real &m1, real &m2, int &arg)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const real om1 = __shfl_xor_sync(0xffffffffu, m1, o), om2 = __shfl_xor_sync(0xffffffffu, m2, o);
        const int oarg = __shfl_xor_sync(0xffffffffu, arg, o);
        const bool take = om1 < m1 || (om1 == m1 && (unsigned)oarg < (unsigned)arg);
        const real hi1 = take ? m1 : om1;                         // the larger of the two smallest
        const real lo2 = om2 < m2 ? om2 : m2;
        m2 = hi1 < lo2 ? hi1 : lo2;
        if (take) { m1 = om1; arg = oarg; }
    }
}

// Column update with the column's messages in registers: one gather of check->bit messages, the prefix products /
// sums in the serial order, one scatter of bit->check messages.  Returns the soft output (total LLR / posterior ratio).
template <typename real, int METHOD, int CT>
__device__ __forceinline__ real bpg_column_in_registers(const int32_t *__restrict__ csc_edge, int lo, int hi, const real *c2b,
                                                        real *b2c, real prior)
{
    using A = Ar<real>;
    int ee[CT];
    real cv[CT], pre[CT];
#pragma unroll
    for (int k = 0; k < CT; ++k) if (lo + k < hi) ee[k] = csc_edge[lo + k];
#pragma unroll
    for (int k = 0; k < CT; ++k) if (lo + k < hi) cv[k] = c2b[ee[k]];
    real temp = prior;
#pragma unroll
    for (int k = 0; k < CT; ++k)
        if (lo + k < hi) {
            pre[k] = temp;
            if (METHOD == QD_BP_MIN_SUM_K) temp = A::add(temp, cv[k]);
            else { temp = A::mul(temp, cv[k]); if (temp != temp) temp = (real)1; }
        }
    const real soft = temp;
    temp = METHOD == QD_BP_MIN_SUM_K ? (real)0 : (real)1;
#pragma unroll
    for (int k = CT - 1; k >= 0; --k)
        if (lo + k < hi) {
            if (METHOD == QD_BP_MIN_SUM_K) { b2c[ee[k]] = A::add(pre[k], temp); temp = A::add(temp, cv[k]); }
            else { b2c[ee[k]] = A::mul(pre[k], temp); temp = A::mul(temp, cv[k]); if (temp != temp) temp = (real)1; }
        }
    return soft;
}

template <typename real, int METHOD>
__global__ void __launch_bounds__(512, 2) bpg_kernel(const BpgParams p)
{
    using A = Ar<real>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint8_t *dec = smem_raw;                                   // [n]
    uint8_t *syn = dec + ((p.n + 15) & ~15);                   // [m]
    uint32_t *bits = reinterpret_cast<uint32_t *>(syn + ((p.m + 15) & ~15));   // [nwn]
    __shared__ long long sh_shot;
    const int tid = (int)threadIdx.x, NT = (int)blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
    real *b2c = static_cast<real *>(p.scratch) + (size_t)blockIdx.x * 2 * p.nnz;
    real *c2b = b2c + p.nnz;
    const real *prior = static_cast<const real *>(p.prior);
    const real alpha_fixed = (real)p.alpha_fixed;
    long long c_fail = 0, c_shots = 0, c_nc = 0, c_it = 0;

    for (;;) {
        if (tid == 0) {
            const unsigned long long id = atomicAdd(p.next_shot, 1ULL);
            sh_shot = id < (unsigned long long)p.shots ? (long long)id : -1;
        }
        __syncthreads();
        const long long shot = sh_shot;
        if (shot < 0) break;
        const uint32_t *sw = p.synd + shot * (long long)p.synd_stride;
        for (int r = tid; r < p.m; r += NT) syn[r] = (uint8_t)((sw[r >> 5] >> (r & 31)) & 1u);
        for (int j = tid; j < p.n; j += NT) {
            dec[j] = 0;
            const real pj = prior[j];
            for (int t = p.col_ptr[j]; t < p.col_ptr[j + 1]; ++t) b2c[p.csc_edge[t]] = pj;
        }
        __syncthreads();
        int it = 0, converged = 0;
        while (it < p.max_iter) {
            ++it;
            // ---- check pass
            if (METHOD == QD_BP_MIN_SUM_K) {
                const real al = alpha_fixed != (real)0 ? alpha_fixed : A::sub((real)1, pow2_neg<real>(it));
                for (int r = warp; r < p.m; r += nwarps) {
                    const int lo = p.row_ptr[r], hi = p.row_ptr[r + 1];
                    real m1 = A::big(), m2 = A::big();
                    int arg = -1;
                    unsigned tot = 0u;
                    if (hi - lo <= 32 * BPG_RT) {
                        real v[BPG_RT];
#pragma unroll
                        for (int k = 0; k < BPG_RT; ++k) { const int e = lo + lane + 32 * k; v[k] = e < hi ? b2c[e] : A::big(); }
#pragma unroll
                        for (int k = 0; k < BPG_RT; ++k) {
                            const int e = lo + lane + 32 * k;
                            if (e < hi) {
                                const real a = abs_bits(v[k]);
                                tot ^= v[k] <= (real)0 ? 1u : 0u;   // the serial rule: a zero message counts as negative
                                if (a < m1) { m2 = m1; m1 = a; arg = e; }
                                else if (a < m2) m2 = a;
                            }
                        }
                        warp_two_min(m1, m2, arg);
                        tot = (__reduce_xor_sync(0xffffffffu, tot) ^ syn[r]) & 1u;
                        const real o1 = A::mul(m1, al), o2 = A::mul(m2, al);
#pragma unroll
                        for (int k = 0; k < BPG_RT; ++k) {
                            const int e = lo + lane + 32 * k;
                            if (e < hi) c2b[e] = A::flip(e == arg ? o2 : o1, tot ^ (v[k] <= (real)0 ? 1u : 0u));
                        }
                    } else {
                        for (int e = lo + lane; e < hi; e += 32) {
                            const real x = b2c[e], a = abs_bits(x);
                            tot ^= x <= (real)0 ? 1u : 0u;
                            if (a < m1) { m2 = m1; m1 = a; arg = e; }
                            else if (a < m2) m2 = a;
                        }
                        warp_two_min(m1, m2, arg);
                        tot = (__reduce_xor_sync(0xffffffffu, tot) ^ syn[r]) & 1u;
                        const real o1 = A::mul(m1, al), o2 = A::mul(m2, al);
                        for (int e = lo + lane; e < hi; e += 32)
                            c2b[e] = A::flip(e == arg ? o2 : o1, tot ^ (b2c[e] <= (real)0 ? 1u : 0u));
                    }
                }
            } else {
                for (int r = tid; r < p.m; r += NT) {
                    const int lo = p.row_ptr[r], hi = p.row_ptr[r + 1];
                    real temp = syn[r] ? (real)-1 : (real)1;
                    for (int e = lo; e < hi; ++e) {
                        c2b[e] = temp;
                        temp = A::mul(temp, A::sub(A::two_over(A::add((real)1, b2c[e])), (real)1));
                    }
                    temp = (real)1;
                    for (int e = hi - 1; e >= lo; --e) {
                        const real x = A::mul(c2b[e], temp);
                        c2b[e] = A::div(A::sub((real)1, x), A::add((real)1, x));
                        temp = A::mul(temp, A::sub(A::two_over(A::add((real)1, b2c[e])), (real)1));
                    }
                }
            }
            __syncthreads();
            // ---- bit pass
            const bool want_llr = p.llr && (p.llr_mode == 0 || it == p.max_iter);
            for (int j = tid; j < p.n; j += NT) {
                const int lo = p.col_ptr[j], hi = p.col_ptr[j + 1];
                real temp = prior[j];
                if (hi - lo <= 8) {
                    const real soft = hi - lo <= 4 ? bpg_column_in_registers<real, METHOD, 4>(p.csc_edge, lo, hi, c2b, b2c, temp)
                                                   : bpg_column_in_registers<real, METHOD, 8>(p.csc_edge, lo, hi, c2b, b2c, temp);
                    if (METHOD == QD_BP_MIN_SUM_K) {
                        dec[j] = soft <= (real)0 ? 1 : 0;
                        if (want_llr) p.llr[shot * (long long)p.n + j] = (double)soft;
                    } else {
                        dec[j] = soft >= (real)1 ? 1 : 0;
                        if (want_llr) p.llr[shot * (long long)p.n + j] = p.llr_mode ? (double)soft : (double)A::lg(A::div((real)1, soft));
                    }
                } else if (METHOD == QD_BP_MIN_SUM_K) {
                    for (int t = lo; t < hi; ++t) { const int e = p.csc_edge[t]; b2c[e] = temp; temp = A::add(temp, c2b[e]); }
                    dec[j] = temp <= (real)0 ? 1 : 0;
                    if (want_llr) p.llr[shot * (long long)p.n + j] = (double)temp;
                    temp = (real)0;
                    for (int t = hi - 1; t >= lo; --t) { const int e = p.csc_edge[t]; b2c[e] = A::add(b2c[e], temp); temp = A::add(temp, c2b[e]); }
                } else {
                    for (int t = lo; t < hi; ++t) {
                        const int e = p.csc_edge[t];
                        b2c[e] = temp;
                        temp = A::mul(temp, c2b[e]);
                        if (temp != temp) temp = (real)1;
                    }
                    dec[j] = temp >= (real)1 ? 1 : 0;
                    if (want_llr) p.llr[shot * (long long)p.n + j] = p.llr_mode ? (double)temp : (double)A::lg(A::div((real)1, temp));
                    temp = (real)1;
                    for (int t = hi - 1; t >= lo; --t) {
                        const int e = p.csc_edge[t];
                        b2c[e] = A::mul(b2c[e], temp);
                        temp = A::mul(temp, c2b[e]);
                        if (temp != temp) temp = (real)1;
                    }
                }
            }
            __syncthreads();
            // ---- H dec == s ?
            int unsat = 0;
            for (int r = warp; r < p.m; r += nwarps) {
                unsigned par = 0u;
                for (int e = p.row_ptr[r] + lane; e < p.row_ptr[r + 1]; e += 32) par ^= dec[p.col_idx[e]];
                par = (__reduce_xor_sync(0xffffffffu, par) ^ syn[r]) & 1u;
                unsat |= (int)par;
            }
            if (!__syncthreads_or(unsat)) { converged = 1; break; }
        }
        // ---- outputs
        for (int w = tid; w < p.nwn; w += NT) {
            uint32_t v = 0u;
            for (int b = 0; b < 32 && w * 32 + b < p.n; ++b) v |= (uint32_t)dec[w * 32 + b] << b;
            bits[w] = v;
        }
        __syncthreads();
        uint32_t *dst = p.corr + shot * (long long)p.corr_stride;
        for (int w = tid; w < p.corr_stride; w += NT) dst[w] = w < p.nwn ? bits[w] : 0u;
        int f = 0;
        if (p.err) {
            const uint32_t *e = p.err + shot * (long long)p.err_stride;
            __syncthreads();
            for (int w = tid; w < p.nwn; w += NT) bits[w] ^= e[w];
            __syncthreads();
            for (int i = tid; i < p.k; i += NT) {
                uint32_t acc = 0u;
                for (int w = 0; w < p.nwn; ++w) acc ^= p.lzT[(size_t)w * p.k + i] & bits[w];
                f |= __popc(acc) & 1;
            }
            f = __syncthreads_or(f);
        }
        if (tid == 0) {
            if (p.conv) p.conv[shot] = (uint8_t)converged;
            if (p.iters) p.iters[shot] = it;
            if (p.fail) p.fail[shot] = (uint8_t)f;
            c_fail += f; c_shots += 1; c_nc += !converged; c_it += it;
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
