// bp7_cluster_layout.h — host-side plan for the cluster mode of the thread-owned-state BP kernel (bp7_kernel.cuh, CL):
// how a Tanner graph that does not fit one SM is spread over the C CTAs of a thread-block cluster.
//
//   * rows: C contiguous shares of equal size (for the spacetime matrices [blockdiag(H x (T+1)) | M] of
//     spacetime_code.py:46-75 with (T+1) a multiple of C these are whole rounds, and the only columns cut by the split
//     are the weight-2 measurement columns between two rounds);
//   * columns: a column whose edges all lie in one share is a record of that CTA only.  A cut column has a replica record
//     in every CTA holding one of its edges; a replica addresses its local edges in the CTA's message array and its
//     remote edges in MAILBOX elements.  All replicas run the same serial sums, so they agree bit for bit; the CTA holding
//     the column's first edge reports its decision;
//   * placement inside a CTA: local messages by the Koenig colouring of bp7_layout.h over the CTA's own accesses; the
//     mailbox elements of one bit-pass access (16 consecutive records, edge t) share ONE line and take the banks the local
//     messages of that access leave free, so the access stays conflict free and the exporter fills a line with
//     consecutive 8-byte stores;
//   * exports: for every local edge of a cut column and every other CTA holding a replica, one entry (local element,
//     destination rank, destination mailbox element), sorted by destination so that consecutive lanes write consecutive
//     elements through the cluster window.
//
// Positions are (line, colour) pairs: the element index of slot 0 is line * NB * S + colour once the slot count S is known.
#pragma once

#include <algorithm>
#include <cstdint>
#include <vector>

#include "bp7_layout.h"

namespace qd {

struct Bp7Pos { int32_t line = -1; int32_t colour = 0; };

struct Bp7ClusterRank {
    int row0 = 0, rows = 0;
    int nrec = 0;                              // records incl. class padding (multiple of 32)
    std::vector<int32_t> rec_col;              // [nrec] global column (-1: padding)
    std::vector<uint8_t> rec_owner;            // [nrec] this CTA reports the column's decision
    std::vector<uint8_t> chunk_deg;            // [nrec / 32]
    std::vector<Bp7Pos> rec_pos;               // [nrec * V] message position of edge t (line < 0: none)
    std::vector<int32_t> rec_flag_row;         // [nrec * V] local row of edge t (-1: remote edge or none)
    std::vector<uint8_t> rec_flag_k;           // [nrec * V] index of the edge within that row
    std::vector<Bp7Pos> row_pos;               // [rows * RD] message position of the k-th edge of a local row (line < 0: none)
    int lines = 0;                             // local lines + mailbox lines
    int local_edges = 0, mailbox_elems = 0;
    struct Export { Bp7Pos src; int32_t rank; Bp7Pos dst; };
    std::vector<Export> exports;
};

struct Bp7ClusterPlan {
    int C = 0, nb = 16, vmax = 4, rd = 9;
    int m_max = 0, m32 = 0, nrec = 0, lines = 0, nexp_max = 0;
    long local_edges = 0, mailbox_elems = 0;
    std::vector<Bp7ClusterRank> rank;
    bool ok = false;
};

// CSR (row_ptr, col_idx sorted) + CSC (col_ptr, csc_row ascending, csc_k = index of the edge within its row)
struct Bp7ClusterProblem {
    int m = 0, n = 0;
    const int32_t *row_ptr = nullptr, *col_idx = nullptr;
    const int32_t *col_ptr = nullptr, *csc_row = nullptr;
    const uint8_t *csc_k = nullptr;
    int vmax = 4, nb = 16, rd = 9;
};

inline Bp7ClusterPlan bp7_cluster_plan(const Bp7ClusterProblem &P, int C)
{
    Bp7ClusterPlan plan;
    plan.C = C; plan.nb = P.nb; plan.vmax = P.vmax; plan.rd = P.rd;
    const int m = P.m, n = P.n, NB = P.nb, V = P.vmax, RD = P.rd;
    if (C < 2 || m < C) return plan;
    for (int r = 0; r < m; ++r) if (P.row_ptr[r + 1] - P.row_ptr[r] > RD) return plan;
    for (int j = 0; j < n; ++j) if (P.col_ptr[j + 1] - P.col_ptr[j] > V) return plan;
    plan.rank.resize(C);
    // share c = rows [ceil(c m / C), ceil((c + 1) m / C))
    for (int c = 0; c < C; ++c) {
        const int lo = (int)(((long)c * m + C - 1) / C), hi = (int)(((long)(c + 1) * m + C - 1) / C);
        plan.rank[c].row0 = lo; plan.rank[c].rows = hi - lo;
    }
    // record index of column j inside rank c (-1: no replica there)
    std::vector<std::vector<int32_t>> rec_of(C, std::vector<int32_t>(n, -1));
    for (int c = 0; c < C; ++c) {
        Bp7ClusterRank &R = plan.rank[c];
        const int lo = R.row0, hi = R.row0 + R.rows;
        // columns with an edge in the share, by full degree, then by the pattern of remote positions
        std::vector<std::vector<std::pair<unsigned, int>>> cls(V + 1);
        {
            std::vector<uint8_t> seen(n, 0);
            for (int r = lo; r < hi; ++r)
                for (int e = P.row_ptr[r]; e < P.row_ptr[r + 1]; ++e) {
                    const int j = P.col_idx[e];
                    if (seen[j]) continue;
                    seen[j] = 1;
                    unsigned remote = 0u;
                    const int dg = P.col_ptr[j + 1] - P.col_ptr[j];
                    for (int t = 0; t < dg; ++t) {
                        const int row = P.csc_row[P.col_ptr[j] + t];
                        if (row < lo || row >= hi) remote |= 1u << t;
                    }
                    cls[dg].push_back({remote, j});
                }
        }
        int nrec = 0;
        for (int dg = 1; dg <= V; ++dg) {
            std::sort(cls[dg].begin(), cls[dg].end());
            nrec += ((int)cls[dg].size() + 31) / 32 * 32;
        }
        R.nrec = nrec;
        R.rec_col.assign(std::max(nrec, 1), -1);
        R.rec_owner.assign(std::max(nrec, 1), 0);
        R.chunk_deg.assign(std::max(nrec / 32, 1), 0);
        R.rec_pos.assign((size_t)std::max(nrec, 1) * V, Bp7Pos());
        R.rec_flag_row.assign((size_t)std::max(nrec, 1) * V, -1);
        R.rec_flag_k.assign((size_t)std::max(nrec, 1) * V, 0);
        R.row_pos.assign((size_t)std::max(R.rows, 1) * RD, Bp7Pos());
        int r0 = 0;
        for (int dg = 1; dg <= V; ++dg) {
            const int cnt = (int)cls[dg].size(), padded = (cnt + 31) / 32 * 32;
            for (int i = 0; i < padded; ++i) {
                R.chunk_deg[(r0 + i) >> 5] = (uint8_t)dg;
                if (i >= cnt) continue;
                const int j = cls[dg][i].second;
                R.rec_col[r0 + i] = j;
                rec_of[c][j] = r0 + i;
                const int first_row = P.csc_row[P.col_ptr[j]];
                R.rec_owner[r0 + i] = first_row >= lo && first_row < hi;
            }
            r0 += padded;
        }
        // local edges: Koenig colouring over this CTA's accesses
        std::vector<int> lv, rv, erow, ek, erec, et;
        for (int r = lo; r < hi; ++r)
            for (int e = P.row_ptr[r]; e < P.row_ptr[r + 1]; ++e) {
                const int j = P.col_idx[e], k = e - P.row_ptr[r], rec = rec_of[c][j];
                int t = 0;
                while (P.csc_row[P.col_ptr[j] + t] != r || P.csc_k[P.col_ptr[j] + t] != k) ++t;
                lv.push_back(((r - lo) / NB) * RD + k);
                rv.push_back((rec / NB) * V + t);
                erow.push_back(r - lo); ek.push_back(k); erec.push_back(rec); et.push_back(t);
            }
        const int nloc = (int)lv.size();
        R.local_edges = nloc;
        const int nleft = ((R.rows + NB - 1) / NB) * RD, nright = ((nrec + NB - 1) / NB + 1) * V;
        const std::vector<int> col = bp7_colour_bipartite(nloc, nleft, nright, lv, rv, NB);
        std::vector<int> used(NB, 0);
        for (int e = 0; e < nloc; ++e) {
            Bp7Pos ps; ps.colour = col[e]; ps.line = used[col[e]]++;
            R.row_pos[(size_t)erow[e] * RD + ek[e]] = ps;
            R.rec_pos[(size_t)erec[e] * V + et[e]] = ps;
            R.rec_flag_row[(size_t)erec[e] * V + et[e]] = erow[e];
            R.rec_flag_k[(size_t)erec[e] * V + et[e]] = (uint8_t)ek[e];
        }
        int lines = nloc ? *std::max_element(used.begin(), used.end()) : 0;
        // mailbox: one line per (group of NB records, edge t) that has a remote edge; its elements take the banks that
        // the local messages of the same access leave free
        for (int g = 0; g * NB < nrec; ++g)
            for (int t = 0; t < V; ++t) {
                unsigned taken = 0u;
                bool any = false;
                for (int i = g * NB; i < std::min(nrec, (g + 1) * NB); ++i) {
                    const Bp7Pos ps = R.rec_pos[(size_t)i * V + t];
                    if (ps.line >= 0) taken |= 1u << ps.colour;
                }
                for (int i = g * NB; i < std::min(nrec, (g + 1) * NB); ++i) {
                    const int j = R.rec_col[i];
                    if (j < 0 || t >= P.col_ptr[j + 1] - P.col_ptr[j]) continue;
                    const int row = P.csc_row[P.col_ptr[j] + t];
                    if (row >= lo && row < hi) continue;
                    int c = 0;
                    while (taken & (1u << c)) ++c;
                    taken |= 1u << c;
                    Bp7Pos ps; ps.line = lines; ps.colour = c;
                    R.rec_pos[(size_t)i * V + t] = ps;
                    ++R.mailbox_elems;
                    any = true;
                }
                if (any) ++lines;
            }
        R.lines = lines;
    }
    // exports
    for (int c = 0; c < C; ++c) {
        Bp7ClusterRank &R = plan.rank[c];
        const int lo = R.row0, hi = R.row0 + R.rows;
        for (int r = lo; r < hi; ++r)
            for (int e = P.row_ptr[r]; e < P.row_ptr[r + 1]; ++e) {
                const int j = P.col_idx[e], k = e - P.row_ptr[r];
                const int dg = P.col_ptr[j + 1] - P.col_ptr[j];
                int t = 0;
                while (P.csc_row[P.col_ptr[j] + t] != r || P.csc_k[P.col_ptr[j] + t] != k) ++t;
                for (int q = 0; q < C; ++q) {
                    if (q == c || rec_of[q][j] < 0) continue;
                    Bp7ClusterRank::Export x;
                    x.src = R.row_pos[(size_t)(r - lo) * RD + k];
                    x.rank = q;
                    x.dst = plan.rank[q].rec_pos[(size_t)rec_of[q][j] * V + t];
                    R.exports.push_back(x);
                }
                (void)dg;
            }
        std::sort(R.exports.begin(), R.exports.end(), [](const Bp7ClusterRank::Export &a, const Bp7ClusterRank::Export &b) {
            if (a.rank != b.rank) return a.rank < b.rank;
            if (a.dst.line != b.dst.line) return a.dst.line < b.dst.line;
            return a.dst.colour < b.dst.colour;
        });
    }
    for (int c = 0; c < C; ++c) {
        const Bp7ClusterRank &R = plan.rank[c];
        plan.m_max = std::max(plan.m_max, R.rows);
        plan.nrec = std::max(plan.nrec, R.nrec);
        plan.lines = std::max(plan.lines, R.lines);
        plan.nexp_max = std::max(plan.nexp_max, (int)R.exports.size());
        plan.local_edges += R.local_edges;
        plan.mailbox_elems += R.mailbox_elems;
    }
    plan.m32 = (plan.m_max + 31) / 32 * 32;
    plan.ok = true;
    return plan;
}

// self-check of a plan: every edge of the graph is a local edge of exactly one CTA, every remote position of a replica
// is fed by exactly one export, no two elements of one access share a bank, no element is used twice.  0 = consistent.
inline long bp7_cluster_plan_errors(const Bp7ClusterProblem &P, const Bp7ClusterPlan &plan)
{
    if (!plan.ok) return -1;
    const int NB = plan.nb, V = plan.vmax, RD = plan.rd, C = plan.C;
    long bad = 0, edges = 0;
    for (int c = 0; c < C; ++c) {
        const Bp7ClusterRank &R = plan.rank[c];
        std::vector<int> owner_of((size_t)std::max(R.lines, 1) * NB, 0);      // users of an element
        std::vector<int> fed((size_t)std::max(R.lines, 1) * NB, 0);           // exports arriving at an element
        // check-pass accesses
        for (int g = 0; g * NB < R.rows; ++g)
            for (int k = 0; k < RD; ++k) {
                unsigned banks = 0u;
                for (int r = g * NB; r < std::min(R.rows, (g + 1) * NB); ++r) {
                    const Bp7Pos ps = R.row_pos[(size_t)r * RD + k];
                    const int deg = P.row_ptr[R.row0 + r + 1] - P.row_ptr[R.row0 + r];
                    if ((ps.line >= 0) != (k < deg)) ++bad;
                    if (ps.line < 0) continue;
                    if (banks & (1u << ps.colour)) ++bad;
                    banks |= 1u << ps.colour;
                    ++edges;
                }
            }
        // bit-pass accesses
        for (int g = 0; g * NB < R.nrec; ++g)
            for (int t = 0; t < V; ++t) {
                unsigned banks = 0u;
                for (int i = g * NB; i < std::min(R.nrec, (g + 1) * NB); ++i) {
                    const int j = R.rec_col[i];
                    const Bp7Pos ps = R.rec_pos[(size_t)i * V + t];
                    const int dg = j >= 0 ? P.col_ptr[j + 1] - P.col_ptr[j] : 0;
                    if ((ps.line >= 0) != (t < dg)) ++bad;
                    if (ps.line < 0) continue;
                    if (ps.line >= R.lines) { ++bad; continue; }
                    if (banks & (1u << ps.colour)) ++bad;
                    banks |= 1u << ps.colour;
                    ++owner_of[(size_t)ps.line * NB + ps.colour];
                    if (R.rec_flag_row[(size_t)i * V + t] < 0) --fed[(size_t)ps.line * NB + ps.colour];   // expects one export
                }
            }
        for (int v : owner_of) if (v > 1) ++bad;
        for (int q = 0; q < C; ++q)
            for (const auto &x : plan.rank[q].exports)
                if (x.rank == c) {
                    if (x.dst.line < 0 || x.dst.line >= R.lines) { ++bad; continue; }
                    ++fed[(size_t)x.dst.line * NB + x.dst.colour];
                }
        for (int v : fed) if (v != 0) ++bad;
    }
    const long nnz = P.m > 0 ? P.row_ptr[P.m] : 0;
    if (edges != nnz) ++bad;
    return bad;
}

}  // namespace qd
