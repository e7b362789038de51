/*
 * qldpc_oracle.c — CPU ORACLE (test infrastructure; see qldpc_oracle.h for the rules and the
 * parity status).  Plain C, double precision, serial flooding BP, compiled with
 * -ffp-contract=off so every floating-point operation is an individually rounded IEEE-754
 * operation in exactly the order written here (the CUDA fp64 path reproduces that order).
 */
#include "qldpc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ============================================================================================
 * Philox4x32-10 (Salmon et al., SC'11).  Shared definition with csrc/philox.cuh.
 * ========================================================================================== */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

uint32_t orc_bernoulli_threshold(double q)
{
    if (!(q > 0.0)) return 0u;
    double t = floor(q * 4294967296.0);
    if (t >= 4294967295.0) return 4294967295u;
    return (uint32_t)t;
}

void orc_sample_bits(uint64_t seed, uint32_t stream, int64_t shot0, int64_t shots, int n,
                     const uint32_t *thr, uint8_t *bits)
{
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    for (int64_t s = 0; s < shots; ++s) {
        uint64_t g = (uint64_t)(shot0 + s);
        uint8_t *row = bits + s * (int64_t)n;
        for (int b = 0; b * 4 < n; ++b) {
            uint32_t ctr[4] = { (uint32_t)g, (uint32_t)(g >> 32), (uint32_t)b, stream };
            uint32_t r[4];
            orc_philox4x32_10(ctr, key, r);
            for (int t = 0; t < 4 && b * 4 + t < n; ++t)
                row[b * 4 + t] = (uint8_t)(r[t] < thr[b * 4 + t]);
        }
    }
}

/* ============================================================================================
 * H @ v mod 2 and the logical-failure flag
 *   reference: (self._checks @ data_readout) % 2      _experiment.py:47,58,124
 *              np.any(GF2(Lz) @ GF2(readout) != 0)    _experiment.py:208-209
 * ========================================================================================== */
void orc_syndrome(int m, const int32_t *row_ptr, const int32_t *col_idx, const uint8_t *vec,
                  uint8_t *synd)
{
    for (int i = 0; i < m; ++i) {
        unsigned acc = 0;
        for (int32_t e = row_ptr[i]; e < row_ptr[i + 1]; ++e) acc ^= vec[col_idx[e]];
        synd[i] = (uint8_t)(acc & 1u);
    }
}

int orc_logical_fail(int k, int n, const uint8_t *L, const uint8_t *vec)
{
    for (int i = 0; i < k; ++i) {
        unsigned acc = 0;
        const uint8_t *row = L + (int64_t)i * n;
        for (int j = 0; j < n; ++j) acc ^= (unsigned)(row[j] & vec[j]);
        if (acc & 1u) return 1;
    }
    return 0;
}

/* ============================================================================================
 * BP — restatement of ldpc.bp_decoder.
 * Edge storage: CSR order (row i, increasing column) is "next_in_row"; the CSC permutation
 * csc_edge[] lists, for column j, its edges in increasing row = "next_in_col".
 * ========================================================================================== */
struct orc_bp {
    int m, n;
    int32_t nnz;
    int32_t *row_ptr, *col_idx;
    int32_t *col_ptr, *csc_edge;
    double *probs;
    double *b2c, *c2b;     /* per-edge messages, CSR order */
    int *esgn;             /* per-edge sign counter of the min-sum pass */
    uint8_t *dec_synd;
    double *soft;          /* raw soft output of the last iteration: total LLR (min-sum), ratio R (product-sum) */
    int max_iter, method;
    double alpha;
};

orc_bp *orc_bp_create(int m, int n, const int32_t *row_ptr, const int32_t *col_idx,
                      const double *probs, int max_iter, int method, double alpha)
{
    orc_bp *d = (orc_bp *)calloc(1, sizeof(orc_bp));
    int32_t nnz = row_ptr[m];
    d->m = m; d->n = n; d->nnz = nnz;
    d->row_ptr = (int32_t *)malloc(sizeof(int32_t) * (m + 1));
    d->col_idx = (int32_t *)malloc(sizeof(int32_t) * (nnz ? nnz : 1));
    memcpy(d->row_ptr, row_ptr, sizeof(int32_t) * (m + 1));
    memcpy(d->col_idx, col_idx, sizeof(int32_t) * nnz);
    d->col_ptr = (int32_t *)calloc(n + 1, sizeof(int32_t));
    d->csc_edge = (int32_t *)malloc(sizeof(int32_t) * (nnz ? nnz : 1));
    for (int32_t e = 0; e < nnz; ++e) d->col_ptr[col_idx[e] + 1]++;
    for (int j = 0; j < n; ++j) d->col_ptr[j + 1] += d->col_ptr[j];
    int32_t *fill = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));
    memcpy(fill, d->col_ptr, sizeof(int32_t) * n);
    for (int i = 0; i < m; ++i)              /* rows ascending => each column list ascending */
        for (int32_t e = row_ptr[i]; e < row_ptr[i + 1]; ++e) d->csc_edge[fill[col_idx[e]]++] = e;
    free(fill);
    d->probs = (double *)malloc(sizeof(double) * (n ? n : 1));
    memcpy(d->probs, probs, sizeof(double) * n);
    d->b2c = (double *)malloc(sizeof(double) * (nnz ? nnz : 1));
    d->c2b = (double *)malloc(sizeof(double) * (nnz ? nnz : 1));
    d->esgn = (int *)malloc(sizeof(int) * (nnz ? nnz : 1));
    d->dec_synd = (uint8_t *)malloc(m ? m : 1);
    d->soft = (double *)calloc(n ? n : 1, sizeof(double));
    d->max_iter = max_iter > 0 ? max_iter : n;   /* ldpc: max_iter == 0 -> n */
    d->method = method;
    d->alpha = alpha;
    return d;
}

void orc_bp_destroy(orc_bp *d)
{
    if (!d) return;
    free(d->row_ptr); free(d->col_idx); free(d->col_ptr); free(d->csc_edge); free(d->probs);
    free(d->b2c); free(d->c2b); free(d->esgn); free(d->dec_synd); free(d->soft); free(d);
}

static int bp_converged(const orc_bp *d, const uint8_t *synd, const uint8_t *decoding)
{
    orc_syndrome(d->m, d->row_ptr, d->col_idx, decoding, d->dec_synd);
    return memcmp(d->dec_synd, synd, (size_t)d->m) == 0;
}

/* product-sum in the probability-ratio domain (bp_method "ps") */
static int bp_prob_ratios(orc_bp *d, const uint8_t *synd, uint8_t *decoding, double *llr, int *iters)
{
    const int m = d->m, n = d->n;
    double *b2c = d->b2c, *c2b = d->c2b;
    for (int j = 0; j < n; ++j) {
        double r = d->probs[j] / (1.0 - d->probs[j]);
        for (int32_t t = d->col_ptr[j]; t < d->col_ptr[j + 1]; ++t) b2c[d->csc_edge[t]] = r;
    }
    *iters = 0;
    for (int it = 1; it <= d->max_iter; ++it) {
        *iters = it;
        for (int i = 0; i < m; ++i) {
            const int32_t lo = d->row_ptr[i], hi = d->row_ptr[i + 1];
            double temp = synd[i] ? -1.0 : 1.0;
            for (int32_t e = lo; e < hi; ++e) {
                c2b[e] = temp;
                temp *= 2.0 / (1.0 + b2c[e]) - 1.0;
            }
            temp = 1.0;
            for (int32_t e = hi - 1; e >= lo; --e) {
                c2b[e] *= temp;
                c2b[e] = (1.0 - c2b[e]) / (1.0 + c2b[e]);
                temp *= 2.0 / (1.0 + b2c[e]) - 1.0;
            }
        }
        for (int j = 0; j < n; ++j) {
            const int32_t lo = d->col_ptr[j], hi = d->col_ptr[j + 1];
            double temp = d->probs[j] / (1.0 - d->probs[j]);
            for (int32_t t = lo; t < hi; ++t) {
                const int32_t e = d->csc_edge[t];
                b2c[e] = temp;
                temp *= c2b[e];
                if (isnan(temp)) temp = 1.0;
            }
            if (llr) llr[j] = log(1.0 / temp);
            d->soft[j] = temp;
            decoding[j] = (uint8_t)(temp >= 1.0);
            temp = 1.0;
            for (int32_t t = hi - 1; t >= lo; --t) {
                const int32_t e = d->csc_edge[t];
                b2c[e] *= temp;
                temp *= c2b[e];
                if (isnan(temp)) temp = 1.0;
            }
        }
        if (bp_converged(d, synd, decoding)) return 1;
    }
    return 0;
}

/* min-sum in the log-likelihood domain (bp_method "ms" / "msl") */
static int bp_min_sum_log(orc_bp *d, const uint8_t *synd, uint8_t *decoding, double *llr, int *iters)
{
    const int m = d->m, n = d->n;
    double *b2c = d->b2c, *c2b = d->c2b;
    int *esgn = d->esgn;
    for (int j = 0; j < n; ++j) {
        double l0 = log((1.0 - d->probs[j]) / d->probs[j]);
        for (int32_t t = d->col_ptr[j]; t < d->col_ptr[j + 1]; ++t) b2c[d->csc_edge[t]] = l0;
    }
    *iters = 0;
    for (int it = 1; it <= d->max_iter; ++it) {
        *iters = it;
        const double alpha = d->alpha == 0.0 ? 1.0 - pow(2.0, -1.0 * it / 1.0) : d->alpha;
        for (int i = 0; i < m; ++i) {
            const int32_t lo = d->row_ptr[i], hi = d->row_ptr[i + 1];
            double temp = 1e308;
            int sgn = synd[i] ? 1 : 0;
            for (int32_t e = lo; e < hi; ++e) {
                c2b[e] = temp;
                esgn[e] = sgn;
                if (fabs(b2c[e]) < temp) temp = fabs(b2c[e]);
                if (b2c[e] <= 0) sgn += 1;
            }
            temp = 1e308;
            sgn = 0;
            for (int32_t e = hi - 1; e >= lo; --e) {
                if (temp < c2b[e]) c2b[e] = temp;
                esgn[e] += sgn;
                c2b[e] *= ((esgn[e] & 1) ? -1.0 : 1.0) * alpha;
                if (fabs(b2c[e]) < temp) temp = fabs(b2c[e]);
                if (b2c[e] <= 0) sgn += 1;
            }
        }
        for (int j = 0; j < n; ++j) {
            const int32_t lo = d->col_ptr[j], hi = d->col_ptr[j + 1];
            double temp = log((1.0 - d->probs[j]) / d->probs[j]);
            for (int32_t t = lo; t < hi; ++t) {
                const int32_t e = d->csc_edge[t];
                b2c[e] = temp;
                temp += c2b[e];
            }
            if (llr) llr[j] = temp;
            d->soft[j] = temp;
            decoding[j] = (uint8_t)(temp <= 0);
            temp = 0.0;
            for (int32_t t = hi - 1; t >= lo; --t) {
                const int32_t e = d->csc_edge[t];
                b2c[e] += temp;
                temp += c2b[e];
            }
        }
        if (bp_converged(d, synd, decoding)) return 1;
    }
    return 0;
}

int orc_bp_decode(orc_bp *d, const uint8_t *synd, uint8_t *decoding, double *llr, int *iters)
{
    int it = 0, conv;
    memset(decoding, 0, (size_t)d->n);
    if (d->method == 0) conv = bp_prob_ratios(d, synd, decoding, llr, &it);
    else conv = bp_min_sum_log(d, synd, decoding, llr, &it);
    if (iters) *iters = it;
    return conv;
}

/* ============================================================================================
 * OSD-0 — restatement of ldpc.bposd_decoder with osd_method "osd0" (reference call sites
 * python/qldpc/misc/_experiment.py:23-27,51,59,77,82,96-100,125).  PARITY UNPINNED (see header).
 * Columns are ranked most-likely-in-error first by the BP soft output of the last iteration:
 * ascending total LLR for min-sum; for product-sum descending posterior ratio R (the order of
 * ascending log(1/R) without a trip through libm); ties by column index.  The first rank(H)
 * linearly independent columns in that order are the information set S; the decoding is the
 * unique solution of H_S x_S = s, zero elsewhere.  Plain Gauss-Jordan on the dense matrix.
 * ========================================================================================== */
typedef struct { double key; int idx; } orc_rank_t;
static int orc_rank_cmp(const void *a, const void *b)
{
    const orc_rank_t *x = (const orc_rank_t *)a, *y = (const orc_rank_t *)b;
    if (x->key < y->key) return -1;
    if (x->key > y->key) return 1;
    return x->idx - y->idx;
}

/* method 0: OSD-0.  method 1 ("osd_e", exhaustive): all 2^order - 1 non-zero patterns on the first `order`
 * non-information-set columns (in ranking order; pattern c sets column b iff bit b of c), c ascending.
 * method 2 ("osd_cs", combination sweep): every single non-information-set column (ranking order), then
 * every pair among the first `order` of them (i < j, lexicographic).  Each pattern t fixes the
 * information-set part as x_S = x_S(OSD-0) + R_T t; the candidate of smallest Hamming weight wins, the
 * OSD-0 solution first and ties to the earlier candidate (strict improvement only). */
int orc_osd(int m, int n, const int32_t *row_ptr, const int32_t *col_idx, const uint8_t *synd,
            const double *key_ascending, int method, int order, uint8_t *decoding)
{
    const int W = (n + 1 + 63) / 64;
    uint64_t *R = (uint64_t *)calloc((size_t)(m ? m : 1) * W, sizeof(uint64_t));
    orc_rank_t *ord = (orc_rank_t *)malloc(sizeof(orc_rank_t) * (n ? n : 1));
    int *pcol = (int *)malloc(sizeof(int) * (m ? m : 1));
    uint8_t *is_piv = (uint8_t *)calloc((size_t)(n ? n : 1), 1);
    for (int i = 0; i < m; ++i) {
        for (int32_t e = row_ptr[i]; e < row_ptr[i + 1]; ++e) R[(size_t)i * W + (col_idx[e] >> 6)] |= (uint64_t)1 << (col_idx[e] & 63);
        if (synd[i]) R[(size_t)i * W + (n >> 6)] |= (uint64_t)1 << (n & 63);
        pcol[i] = -1;
    }
    for (int j = 0; j < n; ++j) { ord[j].key = key_ascending[j]; ord[j].idx = j; }
    qsort(ord, (size_t)n, sizeof(orc_rank_t), orc_rank_cmp);
    int rank = 0;
    for (int t = 0; t < n && rank < m; ++t) {
        const int c = ord[t].idx;
        int piv = -1;
        for (int i = 0; i < m; ++i)
            if (pcol[i] < 0 && ((R[(size_t)i * W + (c >> 6)] >> (c & 63)) & 1)) { piv = i; break; }
        if (piv < 0) continue;
        pcol[piv] = c;
        is_piv[c] = 1;
        ++rank;
        for (int i = 0; i < m; ++i)
            if (i != piv && ((R[(size_t)i * W + (c >> 6)] >> (c & 63)) & 1))
                for (int w = 0; w < W; ++w) R[(size_t)i * W + w] ^= R[(size_t)piv * W + w];
    }
    memset(decoding, 0, (size_t)n);
    int consistent = 1;
    for (int i = 0; i < m; ++i) {
        const int sbit = (int)((R[(size_t)i * W + (n >> 6)] >> (n & 63)) & 1);
        if (pcol[i] >= 0) decoding[pcol[i]] = (uint8_t)sbit;
        else if (sbit) consistent = 0;      /* syndrome outside the column space: cannot happen for s = H e */
    }
    if (method != 0 && order > 0 && consistent) {
        /* T: the columns outside the information set, in ranking order */
        int k = 0;
        int *T = (int *)malloc(sizeof(int) * (n ? n : 1));
        for (int t = 0; t < n; ++t) if (!is_piv[ord[t].idx]) T[k++] = ord[t].idx;
        const int ord_k = order < k ? order : k;
        uint8_t *cand = (uint8_t *)malloc((size_t)(n ? n : 1)), *best = (uint8_t *)malloc((size_t)(n ? n : 1));
        int best_w = 0;
        for (int j = 0; j < n; ++j) best_w += decoding[j];
        memcpy(best, decoding, (size_t)n);
        #define ORC_GETBIT(i, c) ((int)((R[(size_t)(i) * W + ((c) >> 6)] >> ((c) & 63)) & 1))
        /* evaluate the pattern that sets T[a] (and T[b] if b >= 0, ... generic list) */
        int pat[32];
        long ncand = method == 1 ? ((1L << ord_k) - 1) : (long)k + (long)ord_k * (ord_k - 1) / 2;
        long ci = 0;
        int i2 = 0, j2 = 1;                                   /* pair cursor for osd_cs */
        for (; ci < ncand; ++ci) {
            int np = 0;
            if (method == 1) { for (int bpos = 0; bpos < ord_k; ++bpos) if (((ci + 1) >> bpos) & 1) pat[np++] = T[bpos]; }
            else if (ci < k) pat[np++] = T[ci];
            else { pat[np++] = T[i2]; pat[np++] = T[j2]; if (++j2 >= ord_k) { ++i2; j2 = i2 + 1; } }
            memset(cand, 0, (size_t)n);
            int w = np;
            for (int q = 0; q < np; ++q) cand[pat[q]] = 1;
            for (int i = 0; i < m; ++i) {
                if (pcol[i] < 0) continue;
                int bit = ORC_GETBIT(i, n);
                for (int q = 0; q < np; ++q) bit ^= ORC_GETBIT(i, pat[q]);
                cand[pcol[i]] = (uint8_t)bit;
                w += bit;
            }
            if (w < best_w) { best_w = w; memcpy(best, cand, (size_t)n); }
        }
        #undef ORC_GETBIT
        memcpy(decoding, best, (size_t)n);
        free(T); free(cand); free(best);
    }
    free(R); free(ord); free(pcol); free(is_piv);
    return consistent;
}

int orc_osd0(int m, int n, const int32_t *row_ptr, const int32_t *col_idx, const uint8_t *synd,
             const double *key_ascending, uint8_t *decoding)
{
    return orc_osd(m, n, row_ptr, col_idx, synd, key_ascending, 0, 0, decoding);
}

int orc_bposd_decode(orc_bp *d, const uint8_t *synd, int method, int order, uint8_t *decoding, int *iters)
{
    const int conv = orc_bp_decode(d, synd, decoding, NULL, iters);
    if (conv) return 1;
    double *key = (double *)malloc(sizeof(double) * (d->n ? d->n : 1));
    for (int j = 0; j < d->n; ++j) key[j] = d->method == 0 ? -d->soft[j] : d->soft[j];
    orc_osd(d->m, d->n, d->row_ptr, d->col_idx, synd, key, method, order, decoding);
    free(key);
    return 0;
}

/* BP, then OSD-0 on the shots BP did not converge on.  Returns the BP convergence flag. */
int orc_bposd0_decode(orc_bp *d, const uint8_t *synd, uint8_t *decoding, int *iters)
{
    const int conv = orc_bp_decode(d, synd, decoding, NULL, iters);
    if (conv) return 1;
    double *key = (double *)malloc(sizeof(double) * (d->n ? d->n : 1));
    for (int j = 0; j < d->n; ++j) key[j] = d->method == 0 ? -d->soft[j] : d->soft[j];
    orc_osd0(d->m, d->n, d->row_ptr, d->col_idx, synd, key, decoding);
    free(key);
    return 0;
}

const double *orc_bp_soft(const orc_bp *d) { return d->soft; }

/* ============================================================================================
 * Pauli-frame circuit sampler — restatement of what stim's compiled sampler returns for the circuits
 * of build_storage_simulation (python/qldpc/storage_sim.py:38-75,134-185; reference call sites
 * python/qldpc/misc/_experiment.py:171-197), relative to the noiseless reference sample.
 * Program rows {opcode | threshold << 8, q0, q1, site}; opcodes 0 reset, 1 CX, 2 CZ, 3 MZ, 4 MX,
 * 5 MRZ, 6 MRX, 7 DEPOLARIZE1, 8 DEPOLARIZE2, 9 X_ERROR, 10 Z_ERROR.  Noise site k of shot g draws
 * philox(key = seed, ctr = (g_lo, g_hi, k, 0x51)).  PARITY: stim is absent; what is pinned is the
 * circuit text (reference-generated fixtures) and the textbook frame propagation rules.
 * ========================================================================================== */
static void orc_circuit_run(int nops, const int32_t *ops, const uint32_t *thr, int nq, int nmeas, uint64_t seed,
                            int64_t shot0, int64_t shots, const int32_t *forced, uint8_t *records);

void orc_circuit_sample(int nops, const int32_t *ops, const uint32_t *thr, int nq, int nmeas, uint64_t seed,
                        int64_t shot0, int64_t shots, uint8_t *records)
{
    orc_circuit_run(nops, ops, thr, nq, nmeas, seed, shot0, shots, NULL, records);
}

/* record f = the noiseless circuit with exactly one fault {site, Pauli code} = forced[2f], forced[2f+1] */
void orc_circuit_propagate_faults(int nops, const int32_t *ops, int nq, int nmeas, int64_t nfaults, const int32_t *forced,
                                  uint8_t *records)
{
    orc_circuit_run(nops, ops, NULL, nq, nmeas, 0, 0, nfaults, forced, records);
}

static void orc_circuit_run(int nops, const int32_t *ops, const uint32_t *thr, int nq, int nmeas, uint64_t seed,
                            int64_t shot0, int64_t shots, const int32_t *forced, uint8_t *records)
{
    uint8_t *fx = (uint8_t *)malloc((size_t)(nq ? nq : 1)), *fz = (uint8_t *)malloc((size_t)(nq ? nq : 1));
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    for (int64_t s = 0; s < shots; ++s) {
        const uint64_t g = (uint64_t)(shot0 + s);
        uint8_t *rec = records + s * (int64_t)nmeas;
        int k = 0;
        memset(fx, 0, (size_t)nq); memset(fz, 0, (size_t)nq);
        for (int i = 0; i < nops; ++i) {
            const int code = ops[4 * i] & 0xff, ti = ops[4 * i] >> 8, q0 = ops[4 * i + 1], q1 = ops[4 * i + 2];
            uint32_t r[4] = {0, 0, 0, 0};
            int hit = 0;
            uint32_t fcode = 0;
            if (ti && forced) {
                hit = forced[2 * s] == ops[4 * i + 3];
                fcode = (uint32_t)forced[2 * s + 1];
            } else if (ti) {
                const uint32_t ctr[4] = { (uint32_t)g, (uint32_t)(g >> 32), (uint32_t)ops[4 * i + 3], 0x51u };
                orc_philox4x32_10(ctr, key, r);
                hit = r[0] < thr[ti];
            }
            switch (code) {
            case 0: fx[q0] = fz[q0] = 0; break;
            case 1: fx[q1] ^= fx[q0]; fz[q0] ^= fz[q1]; break;
            case 2: { const uint8_t a = fx[q0], b = fx[q1]; fz[q0] ^= b; fz[q1] ^= a; } break;
            case 3: rec[k++] = (uint8_t)(fx[q0] ^ hit); break;
            case 4: rec[k++] = (uint8_t)(fz[q0] ^ hit); break;
            case 5: rec[k++] = (uint8_t)(fx[q0] ^ hit); fx[q0] = fz[q0] = 0; break;
            case 6: rec[k++] = (uint8_t)(fz[q0] ^ hit); fx[q0] = fz[q0] = 0; break;
            case 7:
                if (hit) {
                    const uint32_t c = forced ? fcode : 1u + (uint32_t)(((uint64_t)r[1] * 3u) >> 32);
                    fx[q0] ^= (uint8_t)(c != 3u); fz[q0] ^= (uint8_t)(c != 1u);
                }
                break;
            case 8:
                if (hit) {
                    const uint32_t c = forced ? fcode : 1u + (uint32_t)(((uint64_t)r[1] * 15u) >> 32), pa = c & 3u, pb = c >> 2;
                    fx[q0] ^= (uint8_t)(pa == 1u || pa == 2u); fz[q0] ^= (uint8_t)(pa >= 2u);
                    fx[q1] ^= (uint8_t)(pb == 1u || pb == 2u); fz[q1] ^= (uint8_t)(pb >= 2u);
                }
                break;
            case 9: if (hit) fx[q0] ^= 1; break;
            case 10: if (hit) fz[q0] ^= 1; break;
            default: break;
            }
        }
    }
    free(fx); free(fz);
}

/* ============================================================================================
 * Small-set-flip.  Candidate sets: non-empty subsets F of the support of
 * a generator g (row of Hx).  delta(F) = |sigma| - |sigma ^ Hz 1_F|.  Every round flips the F
 * maximising delta/|F| (exact integer cross-multiplication) among those with delta > 0; ties go
 * to the lowest generator index, then the lowest subset mask (bit i of the mask = i-th qubit of
 * the generator's ascending support).  One flip per round.
 * ========================================================================================== */
#define SSF_MAX_W 12
struct orc_ssf {
    int m, n, mg;
    int32_t *hx_ptr, *hx_idx;       /* generator supports */
    int32_t *win_ptr, *win_chk;     /* per generator: distinct adjacent Hz rows (the window) */
    uint64_t *qmask;                /* per (generator, support position): toggle mask in window bits */
    uint64_t *tmask;                /* scratch: 2^SSF_MAX_W subset toggle masks */
};

orc_ssf *orc_ssf_create(int m, int n, const int32_t *hz_row_ptr, const int32_t *hz_col_idx,
                        int mg, const int32_t *hx_row_ptr, const int32_t *hx_col_idx)
{
    /* CSC of Hz: checks adjacent to each qubit, ascending */
    int32_t nnz = hz_row_ptr[m];
    int32_t *cptr = (int32_t *)calloc(n + 1, sizeof(int32_t));
    int32_t *crow = (int32_t *)malloc(sizeof(int32_t) * (nnz ? nnz : 1));
    for (int32_t e = 0; e < nnz; ++e) cptr[hz_col_idx[e] + 1]++;
    for (int j = 0; j < n; ++j) cptr[j + 1] += cptr[j];
    int32_t *fill = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));
    memcpy(fill, cptr, sizeof(int32_t) * n);
    for (int i = 0; i < m; ++i)
        for (int32_t e = hz_row_ptr[i]; e < hz_row_ptr[i + 1]; ++e) crow[fill[hz_col_idx[e]]++] = i;
    free(fill);

    orc_ssf *d = (orc_ssf *)calloc(1, sizeof(orc_ssf));
    d->m = m; d->n = n; d->mg = mg;
    int32_t gnnz = hx_row_ptr[mg];
    d->hx_ptr = (int32_t *)malloc(sizeof(int32_t) * (mg + 1));
    d->hx_idx = (int32_t *)malloc(sizeof(int32_t) * (gnnz ? gnnz : 1));
    memcpy(d->hx_ptr, hx_row_ptr, sizeof(int32_t) * (mg + 1));
    memcpy(d->hx_idx, hx_col_idx, sizeof(int32_t) * gnnz);
    d->win_ptr = (int32_t *)calloc(mg + 1, sizeof(int32_t));
    d->win_chk = (int32_t *)malloc(sizeof(int32_t) * 64 * (size_t)(mg ? mg : 1));
    d->qmask = (uint64_t *)calloc((size_t)(gnnz ? gnnz : 1), sizeof(uint64_t));
    d->tmask = (uint64_t *)malloc(sizeof(uint64_t) << SSF_MAX_W);
    int ok = 1;
    for (int g = 0; g < mg && ok; ++g) {
        int w = hx_row_ptr[g + 1] - hx_row_ptr[g];
        if (w > SSF_MAX_W) { ok = 0; break; }
        int32_t *win = d->win_chk + d->win_ptr[g];
        int nw = 0;
        for (int32_t t = hx_row_ptr[g]; t < hx_row_ptr[g + 1] && ok; ++t) {
            int q = hx_col_idx[t];
            for (int32_t u = cptr[q]; u < cptr[q + 1]; ++u) {
                int c = crow[u], pos = -1;
                for (int z = 0; z < nw; ++z) if (win[z] == c) { pos = z; break; }
                if (pos < 0) {
                    if (nw == 64) { ok = 0; break; }
                    pos = nw; win[nw++] = c;
                }
                d->qmask[t] ^= (uint64_t)1 << pos;
            }
        }
        d->win_ptr[g + 1] = d->win_ptr[g] + nw;
    }
    free(cptr); free(crow);
    if (!ok) { orc_ssf_destroy(d); return NULL; }
    return d;
}

void orc_ssf_destroy(orc_ssf *d)
{
    if (!d) return;
    free(d->hx_ptr); free(d->hx_idx); free(d->win_ptr); free(d->win_chk); free(d->qmask);
    free(d->tmask); free(d);
}

int orc_ssf_decode(orc_ssf *d, const uint8_t *synd, int max_rounds, uint8_t *decoding,
                   int *rounds, int *final_weight)
{
    uint8_t *sigma = (uint8_t *)malloc(d->m ? d->m : 1);
    memcpy(sigma, synd, (size_t)d->m);
    memset(decoding, 0, (size_t)d->n);
    int weight = 0, r = 0;
    for (int i = 0; i < d->m; ++i) weight += sigma[i];
    while (weight > 0 && (max_rounds == 0 || r < max_rounds)) {
        int best_delta = 0, best_size = 1, best_g = -1;
        unsigned best_mask = 0;
        for (int g = 0; g < d->mg; ++g) {
            const int32_t lo = d->hx_ptr[g];
            const int w = d->hx_ptr[g + 1] - lo;
            const int32_t *win = d->win_chk + d->win_ptr[g];
            const int nw = d->win_ptr[g + 1] - d->win_ptr[g];
            uint64_t local = 0;
            for (int z = 0; z < nw; ++z) local |= (uint64_t)sigma[win[z]] << z;
            if (!local) continue;
            d->tmask[0] = 0;
            for (unsigned mask = 1; mask < (1u << w); ++mask) {
                const unsigned low = mask & (~mask + 1u);
                const uint64_t t = d->tmask[mask ^ low] ^ d->qmask[lo + __builtin_ctz(low)];
                d->tmask[mask] = t;
                const int delta = 2 * __builtin_popcountll(local & t) - __builtin_popcountll(t);
                const int size = __builtin_popcount(mask);
                if (delta > 0 && delta * best_size > best_delta * size) {
                    best_delta = delta; best_size = size; best_g = g; best_mask = mask;
                }
            }
        }
        if (best_g < 0) break;
        const int32_t lo = d->hx_ptr[best_g];
        const int32_t *win = d->win_chk + d->win_ptr[best_g];
        uint64_t t = 0;
        for (int b = 0; best_mask >> b; ++b)
            if ((best_mask >> b) & 1u) { decoding[d->hx_idx[lo + b]] ^= 1; t ^= d->qmask[lo + b]; }
        for (int z = 0; t >> z; ++z)
            if ((t >> z) & 1u) sigma[win[z]] ^= 1;
        weight -= best_delta;
        ++r;
    }
    free(sigma);
    if (rounds) *rounds = r;
    if (final_weight) *final_weight = weight;
    return weight == 0;
}

/* ============================================================================================
 * Batched code-capacity experiment: sample -> syndrome -> decode -> residual -> logical flag.
 * Mirrors run_simulation's per-shot loop (_experiment.py:200-209) with rounds = 0 and the
 * Pool(cpu_count()) sharding of misc/p_sweep.py:18-32 (pthreads pulling disjoint shot chunks).
 * ========================================================================================== */
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

int orc_num_threads(void)
{
    long c = sysconf(_SC_NPROCESSORS_ONLN);
    return c > 0 ? (int)c : 1;
}

typedef struct {
    const orc_experiment *x;
    int decoder;
    uint64_t seed;
    int64_t shot0, shots;
    atomic_llong *next;
    uint8_t *decodings, *converged, *fail;
    int32_t *iters;
    int64_t c_fail, c_nc, c_it;
    int status;
} orc_worker;

#define ORC_CHUNK 64

static void *orc_worker_main(void *arg)
{
    orc_worker *w = (orc_worker *)arg;
    const orc_experiment *x = w->x;
    orc_bp *bp = NULL;
    orc_ssf *ssf = NULL;
    if (w->decoder == 0 || w->decoder >= 2)
        bp = orc_bp_create(x->m, x->n, x->row_ptr, x->col_idx, x->probs, x->max_iter, x->method, x->alpha);
    else
        ssf = orc_ssf_create(x->m, x->n, x->row_ptr, x->col_idx, x->mg, x->hx_row_ptr, x->hx_col_idx);
    if (!bp && !ssf) { w->status = -1; return NULL; }
    uint8_t *err = (uint8_t *)malloc(x->n ? x->n : 1), *dec = (uint8_t *)malloc(x->n ? x->n : 1);
    uint8_t *syn = (uint8_t *)malloc(x->m ? x->m : 1);
    for (;;) {
        int64_t lo = atomic_fetch_add(w->next, ORC_CHUNK);
        if (lo >= w->shots) break;
        int64_t hi = lo + ORC_CHUNK < w->shots ? lo + ORC_CHUNK : w->shots;
        for (int64_t s = lo; s < hi; ++s) {
            int it = 0, conv;
            orc_sample_bits(w->seed, 0u, w->shot0 + s, 1, x->n, x->thr, err);
            orc_syndrome(x->m, x->row_ptr, x->col_idx, err, syn);
            if (bp && w->decoder == 2) conv = orc_bposd0_decode(bp, syn, dec, &it);
            else if (bp && (w->decoder == 3 || w->decoder == 4)) conv = orc_bposd_decode(bp, syn, w->decoder - 2, x->osd_order, dec, &it);
            else if (bp) conv = orc_bp_decode(bp, syn, dec, NULL, &it);
            else conv = orc_ssf_decode(ssf, syn, x->max_iter, dec, &it, NULL);
            if (w->decodings) memcpy(w->decodings + s * (int64_t)x->n, dec, (size_t)x->n);
            for (int j = 0; j < x->n; ++j) err[j] ^= dec[j];
            int f = orc_logical_fail(x->k, x->n, x->Lz, err);
            if (w->converged) w->converged[s] = (uint8_t)conv;
            if (w->iters) w->iters[s] = it;
            if (w->fail) w->fail[s] = (uint8_t)f;
            w->c_fail += f; w->c_nc += !conv; w->c_it += it;
        }
    }
    free(err); free(dec); free(syn);
    orc_bp_destroy(bp); orc_ssf_destroy(ssf);
    return NULL;
}

int orc_run_code_capacity(const orc_experiment *x, int decoder, uint64_t seed, int64_t shot0,
                          int64_t shots, int threads, int64_t counters[4], uint8_t *decodings,
                          uint8_t *converged, int32_t *iters, uint8_t *fail)
{
    if (threads <= 0) threads = orc_num_threads();
    if (threads > 1024) threads = 1024;
    atomic_llong next = 0;
    orc_worker *w = (orc_worker *)calloc((size_t)threads, sizeof(orc_worker));
    pthread_t *tid = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    for (int t = 0; t < threads; ++t) {
        w[t].x = x; w[t].decoder = decoder; w[t].seed = seed; w[t].shot0 = shot0; w[t].shots = shots;
        w[t].next = &next; w[t].decodings = decodings; w[t].converged = converged;
        w[t].iters = iters; w[t].fail = fail;
        if (t > 0) pthread_create(&tid[t], NULL, orc_worker_main, &w[t]);
    }
    orc_worker_main(&w[0]);
    int status = w[0].status;
    counters[0] = w[0].c_fail; counters[1] = shots; counters[2] = w[0].c_nc; counters[3] = w[0].c_it;
    for (int t = 1; t < threads; ++t) {
        pthread_join(tid[t], NULL);
        counters[0] += w[t].c_fail; counters[2] += w[t].c_nc; counters[3] += w[t].c_it;
        if (w[t].status) status = w[t].status;
    }
    free(w); free(tid);
    return status;
}
