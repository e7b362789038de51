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
// in CSR edge order (every row padded to an odd number of elements, which keeps the row-per-thread sweeps of
// equal-degree rows free of bank conflicts), so the check pass is purely local.  A column belongs to the CTA that holds most
// of its edges; its owner runs the column's serial sweeps and reaches the few messages that live in
// another CTA through the cluster address window (map_shared_rank) — for the block-structured matrices
// this kernel exists for, those are only the edges that cross a range boundary.  Hard decisions are
// ballot-packed per warp in "position" order (owner, index within the owner) and the packed word is
// stored into every CTA of the cluster, so the parity test of the next step reads local memory only.
// Per iteration: check pass (fused with the parity test of the previous iteration's decisions, flags to every
// CTA) | cluster barrier | bit pass + decision broadcast | cluster barrier.  One cluster per shot in flight, persistent,
// shots handed out through a global atomic queue by CTA 0.
#pragma once

#include <cooperative_groups.h>

#include "bp_kernel.cuh"
#include "cluster.cuh"

namespace qd {

struct BpcParams {
    int m, n, nnz, nwn, C;
    int max_edges;                          // message-array length per CTA
    int max_rows;                           // rows per CTA (largest range)
    int max_cols;                           // columns per CTA
    int max_cedges;                         // edges of the columns of one CTA
    int nwpos;                              // words of the replicated decision array
    const int32_t *row_lo;                  // [C+1] first row of every range
    const int32_t *row_ptr;                 // CSR offsets
    const int32_t *row_start;               // [m] first message of the row inside its CTA's arrays (rows padded to odd pitch)
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

// identical in every CTA of the cluster (sized by the largest share), so that an offset means the same everywhere
constexpr int BPC_CT = 4;              // columns up to this degree are updated in registers

struct BpcSmemLayout { size_t b2c, c2b, decw, syn, bits, desc, lcptr, prior, rowptr, epos, total; };

__host__ __device__ inline BpcSmemLayout bpc_smem_layout(int max_edges, int max_rows, int max_cols, int max_cedges, int nwpos,
                                                         int nwn, int real_bytes)
{
    BpcSmemLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o += (bytes + 15) & ~(size_t)15; return at; };
    L.b2c = take((size_t)max_edges * real_bytes);
    L.c2b = take((size_t)max_edges * real_bytes);
    L.decw = take((size_t)nwpos * 4);
    L.syn = take((size_t)max_rows);
    L.bits = take((size_t)nwn * 4);
    L.desc = take((size_t)max_cedges * 4);         // (owner rank << 24 | message index) of every column edge
    L.lcptr = take((size_t)(max_cols + 1) * 4);
    L.prior = take((size_t)max_cols * real_bytes);
    L.rowptr = take((size_t)max_rows * 8);         // first message, one past the last
    L.epos = take((size_t)max_edges * 4);
    L.total = o;
    return L;
}

template <typename real, int METHOD>
__global__ void __launch_bounds__(800, 1) bpc_kernel(const BpcParams p)
{
    namespace cg = cooperative_groups;
    using A = Ar<real>;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const BpcSmemLayout L = bpc_smem_layout(p.max_edges, p.max_rows, p.max_cols, p.max_cedges, p.nwpos, p.nwn, (int)sizeof(real));
    real *b2c = reinterpret_cast<real *>(smem_raw + L.b2c);
    real *c2b = reinterpret_cast<real *>(smem_raw + L.c2b);
    uint32_t *decw = reinterpret_cast<uint32_t *>(smem_raw + L.decw);
    uint8_t *syn = smem_raw + L.syn;
    uint32_t *bits = reinterpret_cast<uint32_t *>(smem_raw + L.bits);
    uint32_t *s_desc = reinterpret_cast<uint32_t *>(smem_raw + L.desc);
    int32_t *s_lcptr = reinterpret_cast<int32_t *>(smem_raw + L.lcptr);
    real *s_prior = reinterpret_cast<real *>(smem_raw + L.prior);
    int32_t *s_rowptr = reinterpret_cast<int32_t *>(smem_raw + L.rowptr);
    int32_t *s_epos = reinterpret_cast<int32_t *>(smem_raw + L.epos);
    __shared__ long long sh_shot;
    __shared__ int sh_unsat[2][8];
    const int tid = (int)threadIdx.x, NT = (int)blockDim.x, lane = tid & 31;
    const int C = p.C, rank = (int)cluster.block_rank();
    const int row0 = p.row_lo[rank], nrows = p.row_lo[rank + 1] - row0;
    const int lc0 = p.lc_lo[rank], ncols = p.lc_lo[rank + 1] - lc0, ncols32 = (ncols + 31) & ~31;
    const int w0 = p.word_lo[rank];
    const int32_t *lc_col = p.lc_col + lc0;
    const real alpha_fixed = (real)p.alpha_fixed;
    const uint32_t b2c_sa = (uint32_t)__cvta_generic_to_shared(b2c);
    const uint32_t c2b_sa = (uint32_t)__cvta_generic_to_shared(c2b);
    const uint32_t decw_sa = (uint32_t)__cvta_generic_to_shared(decw);
    long long c_fail = 0, c_shots = 0, c_nc = 0, c_it = 0;
#ifdef QD_BPC_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
#define QD_TICK(k) do { const long long now_ = clock64(); tacc[k] += now_ - tlast; tlast = now_; } while (0)
#else
#define QD_TICK(k) do { } while (0)
#endif
    // A message by descriptor.  Messages of this CTA's own rows (all but the edges that cross a range boundary) are
    // plain shared-memory accesses; only the others go through the cluster window, whose scattered 8-byte
    // accesses are an order of magnitude slower.
    const uint32_t myrank24 = (uint32_t)rank << 24;
    auto ld_msg = [&](real *base, uint32_t base_sa, uint32_t d) -> real {
        const uint32_t ix = d & 0xffffffu;
        if ((d ^ myrank24) < (1u << 24)) return base[ix];
        return cl_ld(cl_map(base_sa + ix * (uint32_t)sizeof(real), d >> 24), real());
    };
    auto st_msg = [&](real *base, uint32_t base_sa, uint32_t d, real v) {
        const uint32_t ix = d & 0xffffffu;
        if ((d ^ myrank24) < (1u << 24)) base[ix] = v;
        else cl_st(cl_map(base_sa + ix * (uint32_t)sizeof(real), d >> 24), v);
    };

    // ---- once per launch: this CTA's share of the graph tables into shared memory
    {
        const int d0 = p.lc_ptr[lc0];
        for (int i = tid; i <= ncols; i += NT) s_lcptr[i] = p.lc_ptr[lc0 + i] - d0;
        for (int i = tid; i < ncols; i += NT) s_prior[i] = static_cast<const real *>(p.prior)[lc0 + i];
        const int nd = p.lc_ptr[lc0 + ncols] - d0;
        for (int t = tid; t < nd; t += NT) s_desc[t] = p.desc[d0 + t];
        for (int r = tid; r < nrows; r += NT) {
            const int lo = p.row_start[row0 + r], e = p.row_ptr[row0 + r], deg = p.row_ptr[row0 + r + 1] - e;
            s_rowptr[2 * r] = lo; s_rowptr[2 * r + 1] = lo + deg;
            for (int k = 0; k < deg; ++k) s_epos[lo + k] = p.edge_pos[e + k];
        }
    }
    cluster_sync_release_all();                 // every CTA of the cluster is running before its shared memory is written remotely

    for (;;) {
        if (rank == 0 && tid == 0) {
            const unsigned long long id = atomicAdd(p.next_shot, 1ULL);
            const long long s = id < (unsigned long long)p.shots ? (long long)id : -1;
            const uint32_t sa = (uint32_t)__cvta_generic_to_shared(&sh_shot);
            for (int r = 0; r < C; ++r) cl_st_s64(cl_map(sa, (uint32_t)r), s);
        }
        cluster_sync_release_all();
        const long long shot = sh_shot;
        if (shot < 0) break;
        const uint32_t *sw = p.synd + shot * (long long)p.synd_stride;
        for (int r = tid; r < nrows; r += NT) syn[r] = (uint8_t)((sw[(row0 + r) >> 5] >> ((row0 + r) & 31)) & 1u);
        for (int i = tid; i < ncols; i += NT) {
            const real pj = s_prior[i];
            for (int t = s_lcptr[i]; t < s_lcptr[i + 1]; ++t) st_msg(b2c, b2c_sa, s_desc[t], pj);
        }
        cluster_sync_release_all();
        // Iteration it + 1's check pass also tests iteration it's decisions (H dec == s ?, decisions are replicated:
        // local reads), so an iteration costs two cluster barriers; the check pass that discovers convergence is
        // wasted, its messages are never used.  it = iterations completed.
        int it = 0, converged = 0;
        QD_TICK(0);
        for (;;) {
            const bool test = it > 0, last = it >= p.max_iter;
            int unsat = 0;
            // ---- check pass: the rows of this CTA, messages local
            for (int r = tid; r < nrows; r += NT) {
                const int lo = s_rowptr[2 * r], hi = s_rowptr[2 * r + 1];
                if (test) {
                    unsigned par = syn[r];
                    for (int e = lo; e < hi; ++e) {
                        const int ps = s_epos[e];
                        par ^= decw[ps >> 5] >> (ps & 31);
                    }
                    unsat |= (int)(par & 1u);
                }
                if (last) continue;
                if (METHOD == QD_BP_MIN_SUM_K) {
                    // The serial prefix/suffix sweeps give edge k the minimum of the other edges' magnitudes and the
                    // parity of the other edges' signs (+ the syndrome bit).  Minimum and parity are exact and
                    // order-free, so: smallest and second-smallest magnitude with the first position of the
                    // smallest, and the sign bits in a mask — one load and one store per edge.
                    const real al = alpha_fixed != (real)0 ? alpha_fixed : A::sub((real)1, pow2_neg<real>(it + 1));
                    real m1 = A::big(), m2 = A::big();
                    int arg = -1;
                    unsigned neg = 0u, tot = syn[r];
                    for (int e = lo; e < hi; ++e) {
                        const real v = b2c[e];
                        const real a = abs_bits(v);
                        const unsigned ng = v <= (real)0 ? 1u : 0u;      // the serial rule: a zero message counts as negative
                        tot ^= ng;
                        if (e - lo < 32) neg |= ng << (e - lo);
                        if (a < m1) { m2 = m1; m1 = a; arg = e; }
                        else if (a < m2) m2 = a;
                    }
                    const real o1 = A::mul(m1, al), o2 = A::mul(m2, al);
                    for (int e = lo; e < hi; ++e) {
                        const unsigned ng = e - lo < 32 ? (neg >> (e - lo)) & 1u : (b2c[e] <= (real)0 ? 1u : 0u);
                        c2b[e] = A::flip(e == arg ? o2 : o1, (tot ^ ng) & 1u);
                    }
                } else {
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
            QD_TICK(1);
            if (test) {
                unsat = __syncthreads_or(unsat);
                if (tid < C) cl_st_u32(cl_map((uint32_t)__cvta_generic_to_shared(&sh_unsat[it & 1][rank]), (uint32_t)tid), (uint32_t)unsat);
            }
            cluster_sync_release_all();
            QD_TICK(2);
            if (test) {
                int any = 0;
                for (int r = 0; r < C; ++r) any |= sh_unsat[it & 1][r];
                if (!any) { converged = 1; break; }
            }
            if (last) break;
            ++it;
            // ---- bit pass: the columns of this CTA; a message may live in another CTA of the cluster
            const bool want_llr = p.llr && (p.llr_mode == 0 || it == p.max_iter);
            for (int i = tid; i < ncols32; i += NT) {
                int d1 = 0;
                if (i < ncols) {
                    const int lo = s_lcptr[i], hi = s_lcptr[i + 1];
                    real temp = s_prior[i];
                    if (hi - lo <= BPC_CT) {
                        // short column: its messages in registers, one load and one store per edge, the same
                        // prefix-then-suffix operation order
                        uint32_t dd[BPC_CT];
                        real cv[BPC_CT], pre[BPC_CT];
#pragma unroll
                        for (int k = 0; k < BPC_CT; ++k)
                            if (lo + k < hi) { dd[k] = s_desc[lo + k]; cv[k] = ld_msg(c2b, c2b_sa, dd[k]); }
#pragma unroll
                        for (int k = 0; k < BPC_CT; ++k)
                            if (lo + k < hi) {
                                pre[k] = temp;
                                if (METHOD == QD_BP_MIN_SUM_K) temp = A::add(temp, cv[k]);
                                else { temp = A::mul(temp, cv[k]); if (temp != temp) temp = (real)1; }
                            }
                        if (METHOD == QD_BP_MIN_SUM_K) {
                            d1 = temp <= (real)0 ? 1 : 0;
                            if (want_llr) p.llr[shot * (long long)p.n + lc_col[i]] = (double)temp;
                            temp = (real)0;
                        } else {
                            d1 = temp >= (real)1 ? 1 : 0;
                            if (want_llr) p.llr[shot * (long long)p.n + lc_col[i]] = p.llr_mode ? (double)temp : (double)A::lg(A::div((real)1, temp));
                            temp = (real)1;
                        }
#pragma unroll
                        for (int k = BPC_CT - 1; k >= 0; --k)
                            if (lo + k < hi) {
                                if (METHOD == QD_BP_MIN_SUM_K) { st_msg(b2c, b2c_sa, dd[k], A::add(pre[k], temp)); temp = A::add(temp, cv[k]); }
                                else { st_msg(b2c, b2c_sa, dd[k], A::mul(pre[k], temp)); temp = A::mul(temp, cv[k]); if (temp != temp) temp = (real)1; }
                            }
                    } else if (METHOD == QD_BP_MIN_SUM_K) {
                        for (int t = lo; t < hi; ++t) { const uint32_t d = s_desc[t]; st_msg(b2c, b2c_sa, d, temp); temp = A::add(temp, ld_msg(c2b, c2b_sa, d)); }
                        d1 = temp <= (real)0 ? 1 : 0;
                        if (want_llr) p.llr[shot * (long long)p.n + lc_col[i]] = (double)temp;
                        temp = (real)0;
                        for (int t = hi - 1; t >= lo; --t) {
                            const uint32_t d = s_desc[t];
                            st_msg(b2c, b2c_sa, d, A::add(ld_msg(b2c, b2c_sa, d), temp));
                            temp = A::add(temp, ld_msg(c2b, c2b_sa, d));
                        }
                    } else {
                        for (int t = lo; t < hi; ++t) {
                            const uint32_t d = s_desc[t];
                            st_msg(b2c, b2c_sa, d, temp);
                            temp = A::mul(temp, ld_msg(c2b, c2b_sa, d));
                            if (temp != temp) temp = (real)1;
                        }
                        d1 = temp >= (real)1 ? 1 : 0;
                        if (want_llr) p.llr[shot * (long long)p.n + lc_col[i]] = p.llr_mode ? (double)temp : (double)A::lg(A::div((real)1, temp));
                        temp = (real)1;
                        for (int t = hi - 1; t >= lo; --t) {
                            const uint32_t d = s_desc[t];
                            st_msg(b2c, b2c_sa, d, A::mul(ld_msg(b2c, b2c_sa, d), temp));
                            temp = A::mul(temp, ld_msg(c2b, c2b_sa, d));
                            if (temp != temp) temp = (real)1;
                        }
                    }
                }
                const unsigned word = __ballot_sync(0xffffffffu, d1);
                if (lane < C) cl_st_u32(cl_map(decw_sa + (uint32_t)(w0 + (i >> 5)) * 4u, (uint32_t)lane), word);
            }
            QD_TICK(3);
            cluster_sync_release_all();
            QD_TICK(4);
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
                if ((size_t)p.k * 4 <= (size_t)p.max_edges * sizeof(real)) {
                    // (logical, 8-word chunk) pairs over the whole CTA; the messages are dead, c2b is scratch
                    uint32_t *lpar = reinterpret_cast<uint32_t *>(c2b);
                    for (int i = tid; i < p.k; i += NT) lpar[i] = 0u;
                    __syncthreads();
                    const int nch = (p.nwn + 7) / 8;
                    for (int x = tid; x < p.k * nch; x += NT) {
                        const int i = x % p.k, wlo = (x / p.k) * 8, whi = min(wlo + 8, p.nwn);
                        uint32_t acc = 0u;
                        for (int w = wlo; w < whi; ++w) acc ^= p.lzT[(size_t)w * p.k + i] & bits[w];
                        if (acc) atomicXor(&lpar[i], acc);
                    }
                    __syncthreads();
                    for (int i = tid; i < p.k; i += NT) f |= __popc(lpar[i]) & 1;
                } else {
                    for (int i = tid; i < p.k; i += NT) {
                        uint32_t acc = 0u;
                        for (int w = 0; w < p.nwn; ++w) acc ^= p.lzT[(size_t)w * p.k + i] & bits[w];
                        f |= __popc(acc) & 1;
                    }
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
        cluster_sync_release_all();
    }
#ifdef QD_BPC_TIMING
    if (blockIdx.x < (unsigned)C && (tid == 0 || tid == NT - 1))
        printf("bpc timing rank %d tid %d: setup/out %lld check+test %lld bar1 %lld bit %lld bar2 %lld  (iters %lld)\n", rank, tid,
               tacc[0] + tacc[4], tacc[1], tacc[2], tacc[3], 0LL, c_it);
#endif
    if (rank == 0 && tid == 0 && p.counters && c_shots) {
        atomicAdd((unsigned long long *)&p.counters[0], (unsigned long long)c_fail);
        atomicAdd((unsigned long long *)&p.counters[1], (unsigned long long)c_shots);
        atomicAdd((unsigned long long *)&p.counters[2], (unsigned long long)c_nc);
        atomicAdd((unsigned long long *)&p.counters[3], (unsigned long long)c_it);
    }
}

}  // namespace qd
