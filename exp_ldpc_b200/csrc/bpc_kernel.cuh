// bpc_kernel.cuh — belief propagation for Tanner graphs larger than one SM's shared memory: the
// messages of one shot are spread over the shared memories of a thread-block CLUSTER (2, 4 or 8 CTAs
// on neighbouring SMs) and exchanged through distributed shared memory.
//
// Same decoder as the other BP kernels (ldpc.bp_decoder semantics; reference call
// sites python/qldpc/misc/_experiment.py:110-117,137-149) and the same operation order as the serial
// restatement (oracle/qldpc_oracle.c: bp_min_sum_log / bp_prob_ratios), so fp64 results are bit-identical
// to it.  Used for the spacetime matrices of the large codes (spacetime_code.py:46-75) and for
// circuit-noise detector error models (:122-183), which the single-SM kernels cannot hold and which
// otherwise fall to the global-memory kernel (bpg_kernel.cuh).
//
// Layout.  Rows are split into C contiguous ranges of about nnz / C edges; CTA c of the cluster owns
// range c and keeps the two message arrays (bit->check, check->bit) of those rows in ITS shared memory
// in CSR edge order, so the check pass is purely local.  A column belongs to the CTA that holds most
// of its edges; its owner runs the column's serial sweeps and reaches the few messages that live in
// another CTA through the cluster address window (map_shared_rank) — for the block-structured matrices
// this kernel exists for, those are only the edges that cross a range boundary.  Hard decisions are
// ballot-packed per warp in "position" order (owner, index within the owner) and the packed word is
// stored into every CTA of the cluster, so the parity test of the next step reads local memory only.
// Per iteration: check pass | cluster barrier | bit pass + decision broadcast | cluster barrier |
// parity test, flags to every CTA | cluster barrier.  One cluster per shot in flight, persistent,
// shots handed out through a global atomic queue by CTA 0.
#pragma once

#include <cooperative_groups.h>

#include "bp_kernel.cuh"

namespace qd {

struct BpcParams {
    int m, n, nnz, nwn, C;
    int max_edges;                          // message-array length per CTA
    int max_rows;                           // rows per CTA (largest range)
    int nwpos;                              // words of the replicated decision array
    const int32_t *row_lo;                  // [C+1] first row of every range
    const int32_t *row_ptr;                 // CSR offsets
    const int32_t *edge_pos;                // [nnz] decision position (word << 5 | bit) of the edge's column
    const int32_t *lc_lo;                   // [C+1] first local column (flat index) of every CTA
    const int32_t *lc_col;                  // [ntot] original column of a local column
    const int32_t *lc_ptr;                  // [ntot+1] offsets into desc
    const uint32_t *desc;                   // [nnz] (owner rank << 24) | message index inside the owner, rows ascending
    const int32_t *word_lo;                 // [C+1] first decision word of every CTA
    const int32_t *col_pos;                 // [n] decision position of column j
    const void *prior;                      // real[ntot], decoder domain, local-column order
    double alpha_fixed;
    int max_iter;
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

struct BpcSmemLayout { size_t b2c, c2b, decw, syn, bits, total; };

__host__ __device__ inline BpcSmemLayout bpc_smem_layout(int max_edges, int max_rows, int nwpos, int nwn, int real_bytes)
{
    BpcSmemLayout L;
    size_t o = 0;
    L.b2c = o;  o += ((size_t)max_edges * real_bytes + 15) & ~(size_t)15;
    L.c2b = o;  o += ((size_t)max_edges * real_bytes + 15) & ~(size_t)15;
    L.decw = o; o += ((size_t)nwpos * 4 + 15) & ~(size_t)15;
    L.syn = o;  o += ((size_t)max_rows + 15) & ~(size_t)15;
    L.bits = o; o += ((size_t)nwn * 4 + 15) & ~(size_t)15;
    L.total = o;
    return L;
}

template <typename real, int METHOD>
__global__ void __launch_bounds__(1024, 1) bpc_kernel(const BpcParams p)
{
    namespace cg = cooperative_groups;
    using A = Ar<real>;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const BpcSmemLayout L = bpc_smem_layout(p.max_edges, p.max_rows, p.nwpos, p.nwn, (int)sizeof(real));
    real *b2c = reinterpret_cast<real *>(smem_raw + L.b2c);
    real *c2b = reinterpret_cast<real *>(smem_raw + L.c2b);
    uint32_t *decw = reinterpret_cast<uint32_t *>(smem_raw + L.decw);
    uint8_t *syn = smem_raw + L.syn;
    uint32_t *bits = reinterpret_cast<uint32_t *>(smem_raw + L.bits);
    __shared__ long long sh_shot;
    __shared__ int sh_unsat[2][8];
    const int tid = (int)threadIdx.x, NT = (int)blockDim.x, lane = tid & 31;
    const int C = p.C, rank = (int)cluster.block_rank();
    const int row0 = p.row_lo[rank], row1 = p.row_lo[rank + 1], e0 = p.row_ptr[row0];
    const int lc0 = p.lc_lo[rank], ncols = p.lc_lo[rank + 1] - lc0, ncols32 = (ncols + 31) & ~31;
    const int w0 = p.word_lo[rank];
    const real *prior = static_cast<const real *>(p.prior) + lc0;
    const int32_t *lc_ptr = p.lc_ptr + lc0, *lc_col = p.lc_col + lc0;
    const real alpha_fixed = (real)p.alpha_fixed;
    long long c_fail = 0, c_shots = 0, c_nc = 0, c_it = 0;
    // a message by its descriptor (owner rank << 24 | index inside the owner): a pointer into the owner's shared memory
    auto at_b2c = [&](uint32_t d) -> real * { return cluster.map_shared_rank(b2c, d >> 24) + (d & 0xffffffu); };
    auto at_c2b = [&](uint32_t d) -> real * { return cluster.map_shared_rank(c2b, d >> 24) + (d & 0xffffffu); };

    for (;;) {
        if (rank == 0 && tid == 0) {
            const unsigned long long id = atomicAdd(p.next_shot, 1ULL);
            const long long s = id < (unsigned long long)p.shots ? (long long)id : -1;
            for (int r = 0; r < C; ++r) *cluster.map_shared_rank(&sh_shot, r) = s;
        }
        cluster.sync();
        const long long shot = sh_shot;
        if (shot < 0) break;
        const uint32_t *sw = p.synd + shot * (long long)p.synd_stride;
        for (int r = row0 + tid; r < row1; r += NT) syn[r - row0] = (uint8_t)((sw[r >> 5] >> (r & 31)) & 1u);
        for (int i = tid; i < ncols; i += NT) {
            const real pj = prior[i];
            for (int t = lc_ptr[i]; t < lc_ptr[i + 1]; ++t) *at_b2c(p.desc[t]) = pj;
        }
        cluster.sync();
        int it = 0, converged = 0;
        while (it < p.max_iter) {
            ++it;
            // ---- check pass: the rows of this CTA, messages local
            for (int r = row0 + tid; r < row1; r += NT) {
                const int lo = p.row_ptr[r] - e0, hi = p.row_ptr[r + 1] - e0;
                if (METHOD == QD_BP_MIN_SUM_K) {
                    const real al = alpha_fixed != (real)0 ? alpha_fixed : A::sub((real)1, pow2_neg<real>(it));
                    real temp = A::big();
                    unsigned sgn = syn[r - row0];
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
                    real temp = syn[r - row0] ? (real)-1 : (real)1;
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
            cluster.sync();
            // ---- bit pass: the columns of this CTA; a message may live in another CTA of the cluster
            const bool want_llr = p.llr && (p.llr_mode == 0 || it == p.max_iter);
            for (int i = tid; i < ncols32; i += NT) {
                int d1 = 0;
                if (i < ncols) {
                    const int lo = lc_ptr[i], hi = lc_ptr[i + 1];
                    real temp = prior[i];
                    if (METHOD == QD_BP_MIN_SUM_K) {
                        for (int t = lo; t < hi; ++t) { const uint32_t d = p.desc[t]; *at_b2c(d) = temp; temp = A::add(temp, *at_c2b(d)); }
                        d1 = temp <= (real)0 ? 1 : 0;
                        if (want_llr) p.llr[shot * (long long)p.n + lc_col[i]] = (double)temp;
                        temp = (real)0;
                        for (int t = hi - 1; t >= lo; --t) {
                            const uint32_t d = p.desc[t];
                            real *q = at_b2c(d);
                            *q = A::add(*q, temp);
                            temp = A::add(temp, *at_c2b(d));
                        }
                    } else {
                        for (int t = lo; t < hi; ++t) {
                            const uint32_t d = p.desc[t];
                            *at_b2c(d) = temp;
                            temp = A::mul(temp, *at_c2b(d));
                            if (temp != temp) temp = (real)1;
                        }
                        d1 = temp >= (real)1 ? 1 : 0;
                        if (want_llr) p.llr[shot * (long long)p.n + lc_col[i]] = p.llr_mode ? (double)temp : (double)A::lg(A::div((real)1, temp));
                        temp = (real)1;
                        for (int t = hi - 1; t >= lo; --t) {
                            const uint32_t d = p.desc[t];
                            real *q = at_b2c(d);
                            *q = A::mul(*q, temp);
                            temp = A::mul(temp, *at_c2b(d));
                            if (temp != temp) temp = (real)1;
                        }
                    }
                }
                const unsigned word = __ballot_sync(0xffffffffu, d1);
                if (lane < C) cluster.map_shared_rank(decw, lane)[w0 + (i >> 5)] = word;
            }
            cluster.sync();
            // ---- H dec == s ?  (decisions are replicated: local reads)
            int unsat = 0;
            for (int r = row0 + tid; r < row1; r += NT) {
                unsigned par = syn[r - row0];
                for (int e = p.row_ptr[r]; e < p.row_ptr[r + 1]; ++e) {
                    const int ps = p.edge_pos[e];
                    par ^= decw[ps >> 5] >> (ps & 31);
                }
                unsat |= (int)(par & 1u);
            }
            unsat = __syncthreads_or(unsat);
            if (tid < C) cluster.map_shared_rank(&sh_unsat[it & 1][0], tid)[rank] = unsat;
            cluster.sync();
            int any = 0;
            for (int r = 0; r < C; ++r) any |= sh_unsat[it & 1][r];
            if (!any) { converged = 1; break; }
        }
        // ---- outputs (CTA 0 holds every decision)
        if (rank == 0) {
            for (int w = tid; w < p.nwn; w += NT) {
                uint32_t v = 0u;
                for (int b = 0; b < 32 && w * 32 + b < p.n; ++b) {
                    const int ps = p.col_pos[w * 32 + b];
                    v |= ((decw[ps >> 5] >> (ps & 31)) & 1u) << b;
                }
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
        }
        // the next shot's setup overwrites sh_shot / messages / decisions of other CTAs: everyone must be done
        cluster.sync();
    }
    if (rank == 0 && tid == 0 && p.counters && c_shots) {
        atomicAdd((unsigned long long *)&p.counters[0], (unsigned long long)c_fail);
        atomicAdd((unsigned long long *)&p.counters[1], (unsigned long long)c_shots);
        atomicAdd((unsigned long long *)&p.counters[2], (unsigned long long)c_nc);
        atomicAdd((unsigned long long *)&p.counters[3], (unsigned long long)c_it);
    }
}

}  // namespace qd
