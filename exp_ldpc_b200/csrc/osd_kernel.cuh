// osd_kernel.cuh — ordered-statistics post-processing (OSD-0, osd_e, osd_cs) for the shots belief
// propagation did not converge on.
//
// Stands in for the OSD stage of ldpc.bposd_decoder as the reference calls it at
// python/qldpc/misc/_experiment.py:23-27,51,59,77,82,96-100,125 (defaults osd_cs order 7, :218-219);
// CPU restatement: oracle/qldpc_oracle.c orc_osd0 / orc_osd (parity unpinned, see oracle/qldpc_oracle.h).
//
// Definition: rank the columns most-likely-in-error first by the soft output of the last BP
// iteration (ascending total LLR for min-sum, descending posterior ratio for product-sum, ties by
// column index); the first rank(H) linearly independent columns in that order are the information
// set S; the OSD-0 decoding is the unique solution of H_S x_S = s, zero elsewhere.  The answer does
// not depend on how the elimination is carried out, only on the ranking, so the kernel is free to
// pick its own pivots:
//   * one CTA per shot; ranking = bitonic sort of (order-preserving 64-bit image of the key, column);
//   * the augmented matrix [H | s] is built bit-packed in shared memory WITH ITS COLUMNS IN RANKED
//     ORDER (m rows of W 64-bit words), so a block of 32 consecutive ranked columns is one aligned
//     32-bit word of every row and, for OSD-0, everything left of the current block is dead;
//   * Gauss-Jordan in lazy blocks of 32 columns, one thread per row: pivots are chosen on
//     register-resident column bits (one barrier per column), the pending row operations are
//     applied once per block through 4-bit lookup tables of the block's pivot rows
//     (method of the four Russians): one 128-bit table load + XOR per 4 pivots and 16 bytes of row;
//   * stops after rank(H) pivots (computed once on the host);
//   * matrices beyond one SM's shared memory (m * W * 8 bytes: the n = 6400 HGP code needs 2.5 MB) stay in a
//     per-CTA global scratch (L2-resident for a bounded number of CTAs); ranking, (B, Y) pairs and lookup tables
//     remain in shared memory.  Same code, slower.
// Integer only => bit-identical to the oracle.
#pragma once

#include "common.cuh"

namespace qd {

struct OsdParams {
    int m, n, nwm, nwn;
    int W;                      // 64-bit words per row of [H | s]
    int P;                      // sort size: power of two >= n
    int rank;                   // rank of H
    int negate_key;             // product-sum: rank by descending ratio
    int method, order;          // 0 OSD-0; 1 osd_e, 2 osd_cs with search depth `order`
    const int32_t *row_ptr, *col_idx;
    const uint32_t *synd; int synd_stride;
    const double *soft;         // [shots][n] raw soft output of the last BP iteration
    const uint8_t *conv;        // BP convergence flags (1: shot is skipped)
    long long shots;
    uint32_t *corr; int corr_stride;
    const uint32_t *err; int err_stride;
    const uint32_t *lzT; int k;
    uint8_t *fail;              // in: flag of the BP decoding; out: flag of the OSD decoding
    long long *counters;        // counters[0] (failures) is corrected by the change of the flag
    unsigned long long *next_shot;
    unsigned long long *rows_global;   // non-null: [gridDim.x][m * W] — the matrix does not fit shared memory and lives in L2 / HBM
};

struct OsdSmemLayout { size_t rows, keys, idx, pcol, dec, tables, nz, prow, total; };

constexpr int OSD_BLOCK = 32;     // ranked columns handled per lazy elimination block

__host__ __device__ inline OsdSmemLayout osd_smem_layout(int m, int W, int P, int nwm, int nwn, bool rows_in_smem)
{
    OsdSmemLayout L;
    size_t o = 0;
    L.rows = o;    o += rows_in_smem ? (size_t)m * W * 8 : 0;
    L.keys = o;    o += (size_t)P * 8;                             // sort keys; then inverse ranking; then (B, Y) pairs; then search scratch
    L.idx = o;     o += ((size_t)P * 2 + 7) & ~(size_t)7;
    L.pcol = o;    o += ((size_t)m * 2 + 7) & ~(size_t)7;
    L.dec = o;     o += ((size_t)nwn * 4 + 7) & ~(size_t)7;
    o = (o + 15) & ~(size_t)15;
    L.tables = o;  o += (size_t)(W / 2) * (OSD_BLOCK / 4) * 16 * 16;  // per 16-byte chunk: 8 nibble tables of 16 entries
    L.nz = o;      o += ((size_t)(W / 2) * 4 + 7) & ~(size_t)7;    // per chunk: which of the block's pivot rows are non-zero there
    L.prow = o;    o += (size_t)OSD_BLOCK * 2;
    L.total = o;
    return L;
}

template <int RPT>
__global__ void __launch_bounds__(1024, 1) osd0_kernel(const OsdParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = (int)threadIdx.x, NT = (int)blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
    const OsdSmemLayout L = osd_smem_layout(p.m, p.W, p.P, p.nwm, p.nwn, p.rows_global == nullptr);
    unsigned long long *R = p.rows_global ? p.rows_global + (size_t)blockIdx.x * p.m * p.W
                                          : reinterpret_cast<unsigned long long *>(smem_raw + L.rows);
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw + L.keys);
    uint16_t *idx = reinterpret_cast<uint16_t *>(smem_raw + L.idx);
    uint16_t *pcol = reinterpret_cast<uint16_t *>(smem_raw + L.pcol);
    uint32_t *dec = reinterpret_cast<uint32_t *>(smem_raw + L.dec);
    __shared__ long long sh_shot;
    __shared__ unsigned candrow[3][32];                       // pivot round r, slot r % 3: lowest unused row whose lowest set bit is column c
    __shared__ int sh_k;
    __shared__ int red32[32];
    __shared__ unsigned long long red64[32];
    const int W = p.W, W2 = p.W / 2, m = p.m, n = p.n;
    long long c_delta = 0;                                    // thread 0: change of the failure count
    int rbuf = 0;                                             // slot of candrow[] the next pivot round nominates into

    for (int i = tid; i < 96; i += NT) candrow[i >> 5][i & 31] = 0xFFFFFFFFu;
    for (int x = tid; x < W2 * (OSD_BLOCK / 4); x += NT)      // entry 0 of every nibble table: the empty XOR
        reinterpret_cast<ulonglong2 *>(smem_raw + L.tables)[x * 16] = make_ulonglong2(0ULL, 0ULL);

    for (;;) {
        if (tid == 0) {
            long long pick = -1;
            for (;;) {
                const unsigned long long id = atomicAdd(p.next_shot, 1ULL);
                if (id >= (unsigned long long)p.shots) break;
                if (!p.conv[id]) { pick = (long long)id; break; }
            }
            sh_shot = pick;
        }
        __syncthreads();
        const long long shot = sh_shot;
        if (shot < 0) break;

        // ---- 1. ranking: idx[t] = column ranked t-th
        for (int i = tid; i < p.P; i += NT) {
            unsigned long long key = ~0ULL;
            if (i < n) {
                double d = p.soft[shot * (long long)n + i];
                if (p.negate_key) d = -d;
                if (d == 0.0) d = 0.0;                              // -0.0 ranks like +0.0 (the oracle compares values)
                const unsigned long long b = (unsigned long long)__double_as_longlong(d);
                key = b ^ ((b >> 63) ? ~0ULL : 0x8000000000000000ULL);
            }
            keys[i] = key;
            idx[i] = (uint16_t)i;
        }
        __syncthreads();
        // bitonic network over (key, column).  Steps with a partner distance below 32 stay inside an aligned run of 32
        // elements: a warp takes such a run into registers (lane = position) and does ALL consecutive short steps with
        // shuffles, so only the steps with distance >= 32 go through shared memory and cost a barrier (21 of 66 at P = 2048).
        auto run_steps = [&](int k2, int jtop) {                  // steps j = jtop, jtop/2, ..., 1 of stage k2 on every run of 32
            for (int r = warp; r < p.P / 32; r += nwarps) {
                const int i = r * 32 + lane;
                unsigned long long key = keys[i];
                uint32_t ix = idx[i];
                for (int kk = k2, j0 = jtop; ; ) {
                    for (int j = j0; j > 0; j >>= 1) {
                        const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, j);
                        const uint32_t oi = __shfl_xor_sync(0xffffffffu, ix, j);
                        const bool mine_gt = key > ok || (key == ok && ix > oi);
                        const bool up = (i & kk) == 0, low = (lane & j) == 0;
                        if (mine_gt == (up == low)) { key = ok; ix = oi; }   // the lower position keeps the smaller one when ascending
                    }
                    if (kk >= 32 || kk >= p.P) break;             // stages 2..32 run back to back inside the registers
                    kk <<= 1; j0 = kk >> 1;
                }
                keys[i] = key;
                idx[i] = (uint16_t)ix;
            }
        };
        run_steps(2, 1);                                          // stages k2 = 2 .. 32 completely
        __syncthreads();
        for (int k2 = 64; k2 <= p.P; k2 <<= 1) {
            for (int j = k2 >> 1; j >= 32; j >>= 1) {
                for (int pr = tid; pr < p.P / 2; pr += NT) {           // one compare-exchange per thread and step: pair (i, i | j)
                    const int i = ((pr & ~(j - 1)) << 1) | (pr & (j - 1)), l = i | j;
                    const unsigned long long ka = keys[i], kb = keys[l];
                    const uint16_t ia = idx[i], ib = idx[l];
                    const bool a_gt_b = ka > kb || (ka == kb && ia > ib);
                    const bool up = (i & k2) == 0;
                    if (a_gt_b == up) { keys[i] = kb; keys[l] = ka; idx[i] = ib; idx[l] = ia; }
                }
                __syncthreads();
            }
            run_steps(k2, 16);
            __syncthreads();
        }

        // ---- 2. [H | s] with the columns in ranked order: column idx[t] sits at bit position t, s at position n
        {
            uint16_t *pos = reinterpret_cast<uint16_t *>(keys);      // the sort keys are dead
            for (int i = tid; i < m * W; i += NT) R[i] = 0ULL;
            for (int t = tid; t < n; t += NT) pos[idx[t]] = (uint16_t)t;
            __syncthreads();
            const uint32_t *s = p.synd + shot * (long long)p.synd_stride;
            for (int r = tid; r < m; r += NT) {
                unsigned long long *row = R + (size_t)r * W;
                for (int e = p.row_ptr[r]; e < p.row_ptr[r + 1]; ++e) {
                    const int q = pos[p.col_idx[e]];
                    row[q >> 6] |= 1ULL << (q & 63);
                }
                if ((s[r >> 5] >> (r & 31)) & 1u) row[n >> 6] |= 1ULL << (n & 63);
            }
            __syncthreads();
        }

        {
            // ---- 3. Gauss-Jordan (the host guarantees NT * RPT >= m), one thread per row (or two), in lazy blocks of 32 ranked columns.
            // A thread pulls its row's 32 bits of the block into a register (B).  Choosing pivots only needs
            // those bits: when row r absorbs pivot i, B_r ^= B_piv and the set Y_r of block-start pivot rows
            // making up the row's pending update becomes Y_r ^ Y_piv ^ {i}.  (B, Y) pairs are published in
            // shared memory (everyone reads the pivot's pair; its owner does not write it in that step), so a
            // column costs one barrier and a few register operations.  After the block,
            //     row_r ^= XOR_{i in Y_r} (pivot row i as it was at the start of the block)
            // is applied through nibble tables: T[chunk][g][e] = XOR of the pivot rows 4g+b, b in e, restricted
            // to the 16-byte chunk.  For OSD-0 the chunks left of the block are never read again and are skipped.
            uint2 *BY = reinterpret_cast<uint2 *>(keys);
            ulonglong2 *T = reinterpret_cast<ulonglong2 *>(smem_raw + L.tables);
            uint32_t *nz = reinterpret_cast<uint32_t *>(smem_raw + L.nz);
            uint16_t *prow = reinterpret_cast<uint16_t *>(smem_raw + L.prow);
            // Thread tid owns rows tid + q * NT, q < RPT (RPT = 2 halves the warps that take part in the
            // latency-bound pivot rounds).
            unsigned long long *myrow[RPT];
            bool mine[RPT], isused[RPT];
            uint32_t B[RPT], Y[RPT], avail[RPT];
            int mypcol[RPT];
#pragma unroll
            for (int q = 0; q < RPT; ++q) {
                mine[q] = tid + q * NT < m;
                myrow[q] = R + (size_t)(mine[q] ? tid + q * NT : 0) * W;
                isused[q] = false;
                mypcol[q] = 0xFFFF;
            }
            int rank = 0;
            for (int t0 = 0; t0 < n && rank < p.rank; t0 += OSD_BLOCK) {
#pragma unroll
                for (int q = 0; q < RPT; ++q) {
                    B[q] = 0u; Y[q] = 0u;
                    if (mine[q]) {
                        B[q] = (uint32_t)(myrow[q][t0 >> 6] >> (t0 & 32));
                        if (n - t0 < 32) B[q] &= (1u << (n - t0)) - 1u;       // position n is s, never a pivot column
                        BY[tid + q * NT] = make_uint2(B[q], 0u);
                    }
                    avail[q] = isused[q] ? 0u : B[q];
                }
                // One barrier per BATCH of pivots.  Any unused row whose block word is non-zero can serve as a pivot for the
                // column of its lowest set bit, as long as it is reduced against the pivots already chosen — the set of pivot
                // columns (leading positions of the span of the unused rows) does not depend on the order.  So a round nominates,
                // for EVERY block column at once, the lowest unused row whose lowest set bit is that column (one shared
                // atomicMin per warp and column value); after the barrier lane c of every warp holds the candidate of column c
                // and its (B, Y) pair, and the candidates are applied from the highest column down: a row with lowest bit c has
                // no bits below c, so applying it never disturbs the lowest bit of a candidate still to come, while that
                // candidate's bits at the higher pivot columns are cleared before its turn (the lanes reduce their candidates
                // in step, lane = column, so nothing is computed twice).  Rows with independent random words have
                // geometrically distributed lowest bits: a round yields ~log2(#unused rows) pivots instead of one.
                int npiv = 0;
                auto pivot_round = [&]() -> bool {              // false: the block is finished
                    const int buf = rbuf, nxt = buf == 2 ? 0 : buf + 1;
                    rbuf = nxt;
#pragma unroll
                    for (int q = 0; q < RPT; ++q) {
                        const int c = avail[q] ? __ffs((int)avail[q]) - 1 : 32;
                        const unsigned grp = __match_any_sync(0xffffffffu, c);
                        if (c < 32 && (grp & ((1u << lane) - 1u)) == 0u) atomicMin(&candrow[buf][c], (unsigned)(tid + q * NT));
                    }
                    if (tid < 32) candrow[nxt][tid] = 0xFFFFFFFFu;   // the slots of the next round (last read two barriers ago)
                    __syncthreads();
                    const unsigned prw = candrow[buf][lane];
                    unsigned todo = __ballot_sync(0xffffffffu, prw != 0xFFFFFFFFu);
                    if (!todo) return false;                    // the rest of the block depends on earlier columns
                    uint32_t pb = 0u, py = 0u;
                    if (prw != 0xFFFFFFFFu) { const uint2 pv = BY[prw]; pb = pv.x; py = pv.y; }
                    // a row knows by itself whether it was chosen: it is the candidate of the column it nominated itself for
                    int mycol[RPT];
                    bool changed[RPT];
#pragma unroll
                    for (int q = 0; q < RPT; ++q) {
                        const int c = avail[q] ? __ffs((int)avail[q]) - 1 : 0;
                        mycol[q] = (avail[q] && candrow[buf][c] == (unsigned)(tid + q * NT)) ? c : -1;
                        changed[q] = false;
                    }
                    // pivot i of the block = the i-th candidate from the top: warp 0 files the rows for the table pass
                    if (warp == 0 && prw != 0xFFFFFFFFu) prow[npiv + __popc(todo >> lane) - 1] = (uint16_t)prw;
                    uint32_t nbit = 1u << npiv;
                    const int k = __popc(todo);
                    while (todo) {
                        const int c = 31 - __clz((int)todo);
                        todo ^= 1u << c;
                        const uint32_t vb = __shfl_sync(0xffffffffu, pb, c);
                        const uint32_t vy = __shfl_sync(0xffffffffu, py, c) ^ nbit;
                        nbit <<= 1;
                        const uint32_t lm = (lane < c && ((pb >> c) & 1u)) ? 0xFFFFFFFFu : 0u;   // candidates of lower columns absorb it
                        pb ^= vb & lm; py ^= vy & lm;
#pragma unroll
                        for (int q = 0; q < RPT; ++q) {
                            const uint32_t hm = (mycol[q] != c && ((B[q] >> c) & 1u)) ? 0xFFFFFFFFu : 0u;
                            B[q] ^= vb & hm; Y[q] ^= vy & hm;
                            changed[q] |= hm != 0u;
                        }
                    }
                    npiv += k; rank += k;
#pragma unroll
                    for (int q = 0; q < RPT; ++q) {
                        // (a row chosen in this round keeps its published pair: others may still be reading it, and a used
                        // row is never nominated again)
                        if (mycol[q] >= 0) { isused[q] = true; mypcol[q] = t0 + mycol[q]; }
                        else if (changed[q]) BY[tid + q * NT] = make_uint2(B[q], Y[q]);
                        avail[q] = isused[q] ? 0u : B[q];
                    }
                    return rank < p.rank;
                };
                while (pivot_round()) { }
                __syncthreads();
                if (npiv == 0) continue;
                const int jc0 = p.method == 0 ? min((t0 + OSD_BLOCK) >> 7, n >> 7) : 0;
                const int ng = (npiv + 3) >> 2;
                for (int x = tid; x < (W2 - jc0) * ng * 16; x += NT) {
                    const int e = x & 15, gq = x >> 4, g = gq % ng, jc = jc0 + gq / ng;
                    ulonglong2 acc = make_ulonglong2(0ULL, 0ULL);
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (((e >> b) & 1) && 4 * g + b < npiv) {
                            const ulonglong2 v = reinterpret_cast<const ulonglong2 *>(R + (size_t)prow[4 * g + b] * W)[jc];
                            acc.x ^= v.x; acc.y ^= v.y;
                        }
                    T[(jc * (OSD_BLOCK / 4) + g) * 16 + e] = acc;
                }
                for (int jc = jc0 + warp; jc < W2; jc += nwarps) {
                    bool nzb = false;
                    if (lane < npiv) {
                        const ulonglong2 v = reinterpret_cast<const ulonglong2 *>(R + (size_t)prow[lane] * W)[jc];
                        nzb = (v.x | v.y) != 0ULL;
                    }
                    const unsigned b = __ballot_sync(0xffffffffu, nzb);
                    if (lane == 0) nz[jc] = b;
                }
                __syncthreads();
#pragma unroll
                for (int q = 0; q < RPT; ++q) {
                    if (!Y[q]) continue;
                    ulonglong2 *row = reinterpret_cast<ulonglong2 *>(myrow[q]);
                    for (int jc = jc0; jc < W2; ++jc) {
                        const uint32_t y = Y[q] & nz[jc];
                        if (!y) continue;
                        const ulonglong2 *Tj = T + jc * (OSD_BLOCK / 4) * 16;
                        // all 8 lookups are issued unconditionally: nibbles beyond the block's pivots are 0 and
                        // entry 0 of every table is zero (set once per launch, only ever rewritten with zero)
                        ulonglong2 tv[OSD_BLOCK / 4];
#pragma unroll
                        for (int g = 0; g < OSD_BLOCK / 4; ++g) tv[g] = Tj[g * 16 + ((y >> (4 * g)) & 15u)];
                        ulonglong2 acc = row[jc];
#pragma unroll
                        for (int g = 0; g < OSD_BLOCK / 4; g += 2) {
                            acc.x ^= tv[g].x ^ tv[g + 1].x; acc.y ^= tv[g].y ^ tv[g + 1].y;
                        }
                        row[jc] = acc;
                    }
                }
                __syncthreads();
            }
#pragma unroll
            for (int q = 0; q < RPT; ++q)
                if (mine[q]) pcol[tid + q * NT] = (uint16_t)mypcol[q];   // ranked position of the row's pivot column, 0xFFFF: none
            __syncthreads();
        }

        // ---- 4. higher-order search (osd_e / osd_cs): patterns on the columns outside the information set
        // T (ranking order).  A pattern t changes the information-set part to x_S = x_S(OSD-0) + R_T t, so
        // its weight is |t| + sum over pivot rows of (s_r ^ parity of the row's bits on t).  Smallest
        // weight wins; OSD-0 first, ties to the earlier candidate (oracle: orc_osd).  Columns are addressed by
        // their ranked position here; idx[] maps back to column numbers in step 5.
        long long best_ci = -1;                                  // -1: keep the OSD-0 solution
        uint32_t *Mrow = reinterpret_cast<uint32_t *>(keys);     // the (B, Y) pairs are dead: reuse them
        uint16_t *tlist = reinterpret_cast<uint16_t *>(Mrow + m);
        int k = 0, ord_k = 0;
        if (p.method != 0 && p.order > 0) {
            uint32_t *colpiv = dec;                               // bit per ranked position: in the information set
            for (int w = tid; w < p.nwn; w += NT) colpiv[w] = 0u;
            __syncthreads();
            for (int r = tid; r < m; r += NT) {
                const unsigned pc = pcol[r];
                if (pc != 0xFFFFu) atomicOr(&colpiv[pc >> 5], 1u << (pc & 31));
            }
            __syncthreads();
            if (warp == 0) {                                      // compaction of T in ranking order
                int count = 0;
                for (int base = 0; base < n; base += 32) {
                    const int t = base + lane;
                    const bool isT = t < n && !((colpiv[t >> 5] >> (t & 31)) & 1u);
                    const unsigned bal = __ballot_sync(0xffffffffu, isT);
                    if (isT) tlist[count + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)t;
                    count += __popc(bal);
                }
                if (lane == 0) sh_k = count;
            }
            __syncthreads();
            k = sh_k;
            ord_k = min(p.order, k);
            // per row: bit b = the row's bit on T[b] (b < ord_k <= 24), bit 30 = pivot row, bit 31 = OSD-0 bit
            int w0 = 0;
            for (int r = tid; r < m; r += NT) {
                uint32_t v = 0u;
                if (pcol[r] != 0xFFFFu) {
                    const unsigned long long *row = R + (size_t)r * W;
                    const uint32_t sb = (uint32_t)((row[n >> 6] >> (n & 63)) & 1ULL);
                    v = (1u << 30) | (sb << 31);
                    for (int b = 0; b < ord_k; ++b) { const int c = tlist[b]; v |= (uint32_t)((row[c >> 6] >> (c & 63)) & 1ULL) << b; }
                    w0 += (int)sb;
                }
                Mrow[r] = v;
            }
            {                                                     // block sum of the OSD-0 weight
                int v = __reduce_add_sync(0xffffffffu, w0);
                if (lane == 0) red32[warp] = v;
                __syncthreads();
                w0 = __reduce_add_sync(0xffffffffu, lane < nwarps ? red32[lane] : 0);
            }
            const long long ncand = p.method == 1 ? ((1LL << ord_k) - 1) : (long long)k + (long long)ord_k * (ord_k - 1) / 2;
            unsigned long long mybest = ~0ULL;                   // (weight << 32) | candidate index
            for (long long ci = tid; ci < ncand; ci += NT) {
                int wt;
                if (p.method == 2 && ci < k) {                   // one column of T
                    const int c = tlist[ci];
                    const int cw = c >> 6, cb = c & 63;
                    wt = 1;
                    for (int r = 0; r < m; ++r) {
                        const uint32_t v = Mrow[r];
                        wt += (int)((v >> 30) & ((v >> 31) ^ (uint32_t)((R[(size_t)r * W + cw] >> cb) & 1ULL)) & 1u);
                    }
                } else {
                    uint32_t mask;
                    if (p.method == 1) mask = (uint32_t)(ci + 1);
                    else {                                       // pair (i, j), i < j < ord_k, lexicographic
                        long long q = ci - k;
                        int i = 0;
                        while (q >= ord_k - 1 - i) { q -= ord_k - 1 - i; ++i; }
                        mask = (1u << i) | (1u << (i + 1 + (int)q));
                    }
                    wt = __popc(mask);
                    for (int r = 0; r < m; ++r) {
                        const uint32_t v = Mrow[r];
                        wt += (int)((v >> 30) & ((v >> 31) ^ (uint32_t)__popc(v & mask)) & 1u);
                    }
                }
                const unsigned long long key = ((unsigned long long)(unsigned)wt << 32) | (unsigned long long)ci;
                mybest = key < mybest ? key : mybest;
            }
            {                                                     // block minimum
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { const unsigned long long o2 = __shfl_xor_sync(0xffffffffu, mybest, o); mybest = o2 < mybest ? o2 : mybest; }
                if (lane == 0) red64[warp] = mybest;
                __syncthreads();
                mybest = lane < nwarps ? red64[lane] : ~0ULL;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { const unsigned long long o2 = __shfl_xor_sync(0xffffffffu, mybest, o); mybest = o2 < mybest ? o2 : mybest; }
            }
            if (mybest != ~0ULL && (int)(mybest >> 32) < w0) best_ci = (long long)(mybest & 0xffffffffULL);
            __syncthreads();
        }

        // ---- 5. the chosen solution, back in column numbers
        for (int w = tid; w < p.nwn; w += NT) dec[w] = 0u;
        __syncthreads();
        {
            uint32_t mask = 0u;
            int single = -1;
            if (best_ci >= 0) {
                if (p.method == 2 && best_ci < k) single = tlist[best_ci];
                else if (p.method == 1) mask = (uint32_t)(best_ci + 1);
                else {
                    long long q = best_ci - k;
                    int i = 0;
                    while (q >= ord_k - 1 - i) { q -= ord_k - 1 - i; ++i; }
                    mask = (1u << i) | (1u << (i + 1 + (int)q));
                }
                if (tid == 0 && single >= 0) { const int c = idx[single]; atomicOr(&dec[c >> 5], 1u << (c & 31)); }
                if (tid < 32 && ((mask >> tid) & 1u)) { const int c = idx[tlist[tid]]; atomicOr(&dec[c >> 5], 1u << (c & 31)); }
            }
            for (int r = tid; r < m; r += NT) {
                const unsigned pc = pcol[r];
                if (pc == 0xFFFFu) continue;
                const unsigned long long *row = R + (size_t)r * W;
                uint32_t bit = (uint32_t)((row[n >> 6] >> (n & 63)) & 1ULL);
                if (single >= 0) bit ^= (uint32_t)((row[single >> 6] >> (single & 63)) & 1ULL);
                if (mask) bit ^= (uint32_t)__popc(Mrow[r] & mask) & 1u;
                if (bit) { const int c = idx[pc]; atomicOr(&dec[c >> 5], 1u << (c & 31)); }
            }
        }
        __syncthreads();
        uint32_t *dst = p.corr + shot * (long long)p.corr_stride;
        for (int w = tid; w < p.corr_stride; w += NT) dst[w] = w < p.nwn ? dec[w] : 0u;
        if (p.err) {
            const uint32_t *e = p.err + shot * (long long)p.err_stride;
            __syncthreads();
            for (int w = tid; w < p.nwn; w += NT) dec[w] ^= e[w];
            __syncthreads();
            int f = 0;
            for (int i = tid; i < p.k; i += NT) {
                uint32_t acc = 0u;
                for (int w = 0; w < p.nwn; ++w) acc ^= p.lzT[(size_t)w * p.k + i] & dec[w];
                f |= __popc(acc) & 1;
            }
            f = __syncthreads_or(f);
            if (tid == 0) {
                const int old = p.fail ? p.fail[shot] : 0;
                if (p.fail) p.fail[shot] = (uint8_t)f;
                c_delta += f - old;
            }
        }
        __syncthreads();
    }
    if (tid == 0 && p.counters && c_delta)
        atomicAdd((unsigned long long *)&p.counters[0], (unsigned long long)c_delta);   // two's complement: may be negative
}

}  // namespace qd
