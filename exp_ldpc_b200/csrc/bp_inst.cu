// bp_inst.cu — explicit instantiations of the BP kernels, one translation unit per
// (precision, method): compiled several times with -DQD_INST_REAL / -DQD_INST_METHOD /
// -DQD_INST_TAG so that the objects build in parallel.  Each object exports two pickers.
#include "bp7_kernel.cuh"
#include "bp_pick.h"

#ifndef QD_INST_REAL
#error "compile with -DQD_INST_REAL=double|float -DQD_INST_METHOD=0|3 -DQD_INST_TAG=f64_ms|..."
#endif

namespace qd {

#define QD_CAT_(a, b) a##b
#define QD_CAT(a, b) QD_CAT_(a, b)

// ---- generic kernel (any row degrees): column tile 8 or 16, 32-bit message positions ----------
bp_kernel_fn QD_CAT(pick_bp6_, QD_INST_TAG)(int vmax, int threads)
{
    typedef QD_INST_REAL real;
    constexpr int M = QD_INST_METHOD;
    if (threads <= 512)
        return vmax <= 8 ? bp_kernel<real, M, 0, 8, uint32_t, 512> : bp_kernel<real, M, 0, 16, uint32_t, 512>;
    return vmax <= 8 ? bp_kernel<real, M, 0, 8, uint32_t, 1024> : bp_kernel<real, M, 0, 16, uint32_t, 1024>;
}

// ---- thread-owned-state kernel: (max row degree, column tile) x slots; IRREG = rows may be shorter
template <typename real, int M, int RD, int VMAX, bool IRREG, int MAXT = BP7_MAXT>
static bp7_kernel_fn pick7_s(int slots)
{
    switch (slots) {
    case 1: return bp7_kernel<real, M, RD, VMAX, 1, IRREG, MAXT, 1>;
    case 2: return bp7_kernel<real, M, RD, VMAX, 2, IRREG, MAXT, 1>;
#ifndef QD_FAST_BUILD
    case 3: return bp7_kernel<real, M, RD, VMAX, 3, IRREG, MAXT, 1>;
    case 6: return bp7_kernel<real, M, RD, VMAX, 6, IRREG, MAXT, 1>;
#endif
    case 4: return bp7_kernel<real, M, RD, VMAX, 4, IRREG, MAXT, 1>;
    case 8: return bp7_kernel<real, M, RD, VMAX, 8, IRREG, MAXT, 1>;
    default: return nullptr;
    }
}

// Column tile 8 (three buffers of up to 8 messages in the bit pass) does not fit the 72 registers an 800-thread CTA
// leaves: the same kernels are also compiled for CTAs of at most 640 threads (96 registers), and the launch takes the
// tightest bound that covers its thread count (a 704-thread bound, 80 registers, was measured to change nothing).
template <typename real, int M, int RD>
static bp7_kernel_fn pick7_wide(int slots, int threads)
{
    if (threads <= 640) return pick7_s<real, M, RD, 8, false, 640>(slots);
    return pick7_s<real, M, RD, 8, false>(slots);
}

// cluster mode (rows split over the CTAs of a thread-block cluster): the padded row degree 9 covers the spacetime
// matrices it exists for, and any graph with rows up to 9 and columns up to 4 for the tests
template <typename real, int M>
static bp7_kernel_fn pick7_cluster(int rd, int vmax, int slots)
{
    if (rd != 9 || vmax != 4) return nullptr;
    switch (slots) {
    case 1: return bp7_kernel<real, M, 9, 4, 1, true, BP7_MAXT, 1, true>;
    case 2: return bp7_kernel<real, M, 9, 4, 2, true, BP7_MAXT, 1, true>;
    case 3: return bp7_kernel<real, M, 9, 4, 3, true, BP7_MAXT, 1, true>;
    case 4: return bp7_kernel<real, M, 9, 4, 4, true, BP7_MAXT, 1, true>;
    default: return nullptr;
    }
}

bp7_kernel_fn QD_CAT(pick_bp7_, QD_INST_TAG)(int rd, int vmax, int slots, int irregular, int cluster, int threads)
{
    typedef QD_INST_REAL real;
    constexpr int M = QD_INST_METHOD;
    if (cluster) return pick7_cluster<real, M>(rd, vmax, slots);
    switch ((irregular ? 10000 : 0) + rd * 100 + vmax) {
    case 704: return pick7_s<real, M, 7, 4, false>(slots);
    case 10904: return pick7_s<real, M, 9, 4, true>(slots);      // spacetime matrices of (3,4) HGP codes
#ifndef QD_FAST_BUILD
    case 604: return pick7_s<real, M, 6, 4, false>(slots);
    case 808: return pick7_wide<real, M, 8>(slots, threads);
    case 908: return pick7_wide<real, M, 9>(slots, threads);
    case 10804: return pick7_s<real, M, 8, 4, true>(slots);      // single-shot [H | I], spacetime of weight-6 rows
#endif
    default: return nullptr;
    }
}

}  // namespace qd
