// bp7_layout.h — host-side message placement for the thread-owned-state BP kernel (bp7_kernel.cuh).
//
// Both passes of the kernel walk the same message array: in the check pass a group of NB lanes
// (NB = 16 for 8-byte messages, 32 for 4-byte messages: the lanes that share one 128-byte
// shared-memory wavefront) owns NB consecutive rows and touches the k-th edge of each; in the bit
// pass a group of NB lanes owns NB consecutive column records and touches the t-th edge of each.
// An access is bank-conflict free iff its NB messages lie in NB different banks.
//
// Build the bipartite multigraph whose left vertices are the check-pass accesses (row group, k),
// whose right vertices are the bit-pass accesses (record group, t), and whose edges are the
// messages (each message belongs to exactly one access of either kind).  Every vertex has degree
// <= NB, so by Koenig's edge-colouring theorem the messages can be coloured with NB colours such
// that no two messages of one access share a colour.  Colour = bank: message e is stored at
// element index  NB * i(e) + colour(e)  (i = running index within its colour), and BOTH passes are
// conflict free for ANY Tanner graph.  The placement only decides where a message lives, never
// the order of any arithmetic, so decoder outputs do not depend on it.
#pragma once

#include <algorithm>
#include <cstdint>
#include <vector>

namespace qd {

struct Bp7Layout {
    int nb = 16;
    std::vector<std::vector<int>> cls;         // per degree (index = degree): columns in record order
    std::vector<int> rec_of_col;               // column -> record index (-1: no edges)
    int nrec = 0;                              // records incl. padding (classes padded to 32)
    std::vector<uint8_t> colour;               // per CSR edge: bank
    std::vector<int32_t> index;                // per CSR edge: running index within its colour
    int lines = 0;                             // max index + 1: the array holds `lines` x NB messages per slot
};

struct Bp7LayoutProblem {
    int m = 0, n = 0;
    const int32_t *row_ptr = nullptr, *col_idx = nullptr;     // CSR, sorted columns
    const int32_t *col_ptr = nullptr;                          // CSC offsets
    int vmax = 4, nb = 16;
};

// Kempe-chain edge colouring of a bipartite multigraph with maximum degree <= NB: edge e joins left vertex lv[e] and
// right vertex rv[e]; returns a colour in [0, NB) per edge, no colour twice at a vertex (Koenig).
inline std::vector<int> bp7_colour_bipartite(int nnz, int nleft, int nright, const std::vector<int> &lv, const std::vector<int> &rv,
                                             int NB)
{
    std::vector<int> at_l((size_t)std::max(nleft, 1) * NB, -1), at_r((size_t)std::max(nright, 1) * NB, -1);   // edge using colour c at a vertex
    std::vector<int> col(std::max(nnz, 1), -1);
    auto free_colour = [&](const std::vector<int> &at, int v) {
        for (int c = 0; c < NB; ++c) if (at[(size_t)v * NB + c] < 0) return c;
        return -1;
    };
    std::vector<int> path;
    for (int e = 0; e < nnz; ++e) {
        const int u = lv[e], v = rv[e];
        const int a = free_colour(at_l, u), b = free_colour(at_r, v);
        if (a != b && at_r[(size_t)v * NB + a] >= 0) {
            // colour a is taken at v: swap a <-> b along the alternating path that starts at v with
            // its a-edge (it cannot reach u: u has no a-edge and the graph is bipartite)
            path.clear();
            int side = 1, w = v, c = a;                            // side 1: w is a right vertex
            for (;;) {
                const int f = side ? at_r[(size_t)w * NB + c] : at_l[(size_t)w * NB + c];
                if (f < 0) break;
                path.push_back(f);
                w = side ? lv[f] : rv[f];
                side ^= 1;
                c = c == a ? b : a;
            }
            for (int f : path) { at_l[(size_t)lv[f] * NB + col[f]] = -1; at_r[(size_t)rv[f] * NB + col[f]] = -1; }
            for (int f : path) {
                col[f] = col[f] == a ? b : a;
                at_l[(size_t)lv[f] * NB + col[f]] = f;
                at_r[(size_t)rv[f] * NB + col[f]] = f;
            }
        }
        col[e] = a;
        at_l[(size_t)u * NB + a] = e;
        at_r[(size_t)v * NB + a] = e;
    }
    return col;
}

inline Bp7Layout bp7_colour_layout(const Bp7LayoutProblem &P)
{
    Bp7Layout L;
    const int m = P.m, n = P.n, NB = P.nb, V = P.vmax;
    const int nnz = m > 0 ? P.row_ptr[m] : 0;
    L.nb = NB;
    L.cls.assign(V + 1, {});
    L.rec_of_col.assign(n, -1);
    for (int j = 0; j < n; ++j) {
        const int dg = P.col_ptr[j + 1] - P.col_ptr[j];
        if (dg >= 1 && dg <= V) L.cls[dg].push_back(j);
    }
    int nrec = 0;
    for (int dg = 1; dg <= V; ++dg) {
        for (size_t i = 0; i < L.cls[dg].size(); ++i) L.rec_of_col[L.cls[dg][i]] = nrec + (int)i;
        nrec += ((int)L.cls[dg].size() + 31) / 32 * 32;
    }
    L.nrec = nrec;
    L.colour.assign(std::max(nnz, 1), 0);
    L.index.assign(std::max(nnz, 1), 0);
    if (nnz == 0) { L.lines = 0; return L; }

    // vertices: left (row group, k), right (record group, t)
    int max_rd = 0;
    for (int r = 0; r < m; ++r) max_rd = std::max(max_rd, P.row_ptr[r + 1] - P.row_ptr[r]);
    const int nleft = ((m + NB - 1) / NB) * max_rd;
    const int nright = ((nrec + NB - 1) / NB) * V;
    std::vector<int> lv(nnz), rv(nnz);
    {
        std::vector<int> fill(n, 0);                               // rank of the row within its column
        for (int r = 0; r < m; ++r)
            for (int e = P.row_ptr[r]; e < P.row_ptr[r + 1]; ++e) {
                const int j = P.col_idx[e], t = fill[j]++;
                lv[e] = (r / NB) * max_rd + (e - P.row_ptr[r]);
                rv[e] = (L.rec_of_col[j] / NB) * V + t;
            }
    }
    const std::vector<int> col = bp7_colour_bipartite(nnz, nleft, nright, lv, rv, NB);
    std::vector<int> used(NB, 0);
    for (int e = 0; e < nnz; ++e) {
        L.colour[e] = (uint8_t)col[e];
        L.index[e] = used[col[e]]++;
    }
    L.lines = *std::max_element(used.begin(), used.end());
    return L;
}

// number of (access, bank) collisions of a placement: 0 for a proper colouring (self-check)
inline long bp7_layout_conflicts(const Bp7LayoutProblem &P, const Bp7Layout &L)
{
    const int m = P.m, NB = P.nb, V = P.vmax;
    int max_rd = 0;
    for (int r = 0; r < m; ++r) max_rd = std::max(max_rd, P.row_ptr[r + 1] - P.row_ptr[r]);
    std::vector<uint8_t> seen_l((size_t)((m + NB - 1) / NB) * std::max(max_rd, 1) * NB, 0);
    std::vector<uint8_t> seen_r((size_t)((L.nrec + NB - 1) / NB + 1) * V * NB, 0);
    std::vector<int> fill(P.n, 0);
    long bad = 0;
    for (int r = 0; r < m; ++r)
        for (int e = P.row_ptr[r]; e < P.row_ptr[r + 1]; ++e) {
            const int j = P.col_idx[e], t = fill[j]++;
            uint8_t &a = seen_l[((size_t)(r / NB) * max_rd + (e - P.row_ptr[r])) * NB + L.colour[e]];
            uint8_t &b = seen_r[((size_t)(L.rec_of_col[j] / NB) * V + t) * NB + L.colour[e]];
            bad += a + b;
            a = b = 1;
        }
    return bad;
}

}  // namespace qd
