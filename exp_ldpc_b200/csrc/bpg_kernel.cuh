// bpg_kernel.cuh — belief propagation with the messages in global memory (L2-resident scratch).
//
// Fallback for Tanner graphs that the shared-memory kernels cannot take: more messages than one SM's
// shared memory holds (spacetime matrices of the large codes, circuit-noise detector error models)
// or degrees beyond their limits.  Same decoder (ldpc.bp_decoder semantics, SURVEY.md Appendix B;
// reference call sites python/qldpc/misc/_experiment.py:110-117,137-149), same operation order as the
// serial restatement (oracle/qldpc_oracle.c: bp_min_sum_log / bp_prob_ratios), so fp64 results are
// bit-identical to it; an order of magnitude slower than the shared-memory kernels.
//
// One CTA per shot in flight (persistent, atomic shot queue).  Two message arrays per CTA in a global
// scratch buffer, CSR edge order: b2c (bit -> check) and c2b (check -> bit).  Check pass: one thread
// per row, forward sweep writes the exclusive prefix (min + sign parity packed in the sign bit, or
// the prefix product), backward sweep folds in the suffix.  Bit pass: one thread per column over its
// edges in row order (prefix sums, then suffix sums).  Hard decisions (bytes) and the syndrome live
// in shared memory; convergence is tested exactly after every iteration.
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

template <typename real, int METHOD>
__global__ void __launch_bounds__(512) bpg_kernel(const BpgParams p)
{
    using A = Ar<real>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint8_t *dec = smem_raw;                                   // [n]
    uint8_t *syn = dec + ((p.n + 15) & ~15);                   // [m]
    uint32_t *bits = reinterpret_cast<uint32_t *>(syn + ((p.m + 15) & ~15));   // [nwn]
    __shared__ long long sh_shot;
    const int tid = (int)threadIdx.x, NT = (int)blockDim.x;
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
            for (int r = tid; r < p.m; r += NT) {
                const int lo = p.row_ptr[r], hi = p.row_ptr[r + 1];
                if (METHOD == QD_BP_MIN_SUM_K) {
                    const real al = alpha_fixed != (real)0 ? alpha_fixed : A::sub((real)1, pow2_neg<real>(it));
                    real temp = A::big();
                    unsigned sgn = syn[r];
                    for (int e = lo; e < hi; ++e) {             // exclusive prefix: min, sign parity in the sign bit
                        const real v = b2c[e];
                        c2b[e] = A::flip(temp, sgn & 1u);
                        const real a = abs_bits(v);
                        if (a < temp) temp = a;
                        sgn += v <= (real)0 ? 1u : 0u;         // the serial rule: a zero message counts as negative
                    }
                    temp = A::big();
                    sgn = 0u;
                    for (int e = hi - 1; e >= lo; --e) {
                        const real pre = c2b[e], v = b2c[e];
                        real mag = abs_bits(pre);
                        if (temp < mag) mag = temp;
                        const unsigned par = ((hi_bits(pre) >> 31) + sgn) & 1u;
                        c2b[e] = A::flip(A::mul(mag, al), par);
                        const real a = abs_bits(v);
                        if (a < temp) temp = a;
                        sgn += v <= (real)0 ? 1u : 0u;
                    }
                } else {
                    real temp = syn[r] ? (real)-1 : (real)1;
                    for (int e = lo; e < hi; ++e) {
                        c2b[e] = temp;
                        temp = A::mul(temp, A::sub(A::div((real)2, A::add((real)1, b2c[e])), (real)1));
                    }
                    temp = (real)1;
                    for (int e = hi - 1; e >= lo; --e) {
                        const real x = A::mul(c2b[e], temp);
                        c2b[e] = A::div(A::sub((real)1, x), A::add((real)1, x));
                        temp = A::mul(temp, A::sub(A::div((real)2, A::add((real)1, b2c[e])), (real)1));
                    }
                }
            }
            __syncthreads();
            // ---- bit pass
            const bool want_llr = p.llr && (p.llr_mode == 0 || it == p.max_iter);
            for (int j = tid; j < p.n; j += NT) {
                const int lo = p.col_ptr[j], hi = p.col_ptr[j + 1];
                real temp = prior[j];
                if (METHOD == QD_BP_MIN_SUM_K) {
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
            for (int r = tid; r < p.m; r += NT) {
                unsigned par = syn[r];
                for (int e = p.row_ptr[r]; e < p.row_ptr[r + 1]; ++e) par ^= dec[p.col_idx[e]];
                unsat |= (int)(par & 1u);
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
