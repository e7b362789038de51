// bp_kernel.cuh — flooding belief propagation, messages resident in shared memory.
//
// What it computes is the syndrome-BP of ldpc.bp_decoder as called by the reference at
// python/qldpc/misc/_experiment.py:110-117,137-149 (semantics: SURVEY.md Appendix B; the CPU
// restatement used by the tests lives under oracle/).  How it computes it is B200-specific:
//
//   * persistent CTAs, each keeping S shots ("slots") in flight; a finished slot (converged or
//     out of iterations) is written out and immediately refilled from a global atomic queue, so
//     lanes never idle on the wide spread of per-shot iteration counts.  A thread owns the same
//     rows / columns in every slot, so graph indices and addresses are formed once per pass;
//   * ONE in-place message array per slot in shared memory: the check pass reads the bit->check
//     messages of its row and overwrites them with check->bit messages; the bit pass reads the
//     check->bit messages of its column and overwrites them with bit->check messages.  The same
//     thread owns a whole row (resp. column), so no second array and no intra-phase hazards;
//   * rows padded to an odd pitch => the check pass is bank-conflict free; regular row degree
//     is a template parameter so a row lives in registers between the two sweeps of the pass;
//   * no atomics on the message path: hard decisions are ballot-packed one word per warp; a bit
//     whose decision flipped raises one flag byte per incident edge, the next check pass folds
//     the flags of its row into the parity of that check (one 64-bit load per row) and the
//     unsatisfied-check count is one atomicAdd per warp.  The convergence test therefore trails
//     the decisions by half an iteration; the decisions themselves are final, so outputs are
//     unchanged;
//   * fp64 path: every operation is an explicitly rounded __d*_rn intrinsic (or an exact
//     min/max/negate) issued in exactly the order of the serial reference (prefix sweep then
//     suffix sweep), so hard decisions, iteration counts and soft outputs are bit-identical to
//     the CPU restatement.  fp32 path: same structure, half the shared-memory traffic, not
//     bit-reproducible.
#pragma once

#include <type_traits>

#include "common.cuh"

namespace qd {

struct GraphDev {
    int m, n, nnz;
    int nwm, nwn;                 // packed words per syndrome / per error vector
    int npos;                     // message positions per slot (rows padded to `pitch`)
    int pitch;                    // row pitch of the message array (0 = CSR offsets, irregular)
    int regular_deg;              // common row degree, 0 if rows differ
    int fb;                       // flag bytes per row (multiple of 8, >= max row degree)
    int max_row_deg, max_col_deg;
    const uint32_t *row_desc;     // [m]   position of the row's first message | degree << 24
    const uint32_t *col_desc;     // [n]   offset into csc_* | degree << 24
    const uint32_t *csc_pos;      // [nnz] message position of the t-th edge of a column (rows ascending)
    const uint32_t *csc_row;      // [nnz] check index of that edge
    const uint32_t *fpos_of_pos;  // [npos] flag-byte index (row * fb + k) of a message position
    const int32_t *row_ptr;       // CSR (syndrome kernel)
    const int32_t *col_idx;
    const int32_t *ell;           // [32][ell_deg][nwm]: column of the e-th edge of check 32 k + b at (b * ell_deg + e) * nwm + k, -1 = none
    int ell_deg;                  // max row degree if the table exists (<= 64), else 0
};

struct BpParams {
    GraphDev g;
    const void *prior;            // real[n]: log((1-p)/p) for min-sum, p/(1-p) for product-sum
    const void *alpha;            // real[max_iter + 1]: scaling factor of iteration it (min-sum)
    const void *vidx;             // IT[n][VMAX]: message positions of each column, ~0 = none
    const uint8_t *chunk_deg;     // [ceil(n/32)] common degree of the 32 columns of a chunk, 0 = mixed
    int max_iter;
    int slots;                    // shots in flight per CTA (all teams)
    int groups;                   // warp teams per CTA; slots % groups == 0, blockDim % (32 * groups) == 0
    int stage_idx;                // copy vidx + prior into shared memory
    const uint32_t *synd; int synd_stride;
    long long shots;
    uint32_t *corr; int corr_stride;
    uint8_t *conv; int32_t *iters; double *llr;
    int llr_mode;                 // 0: log-ratio after every iteration; 1: raw soft output of the last allowed iteration
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
    // 2 / b: doubling is exact, so twice the correctly rounded reciprocal IS the correctly rounded quotient
    // (b = 1 + ratio >= 1 here: no overflow, no subnormal result) — a reciprocal is cheaper than a division
    static __device__ __forceinline__ double two_over(double b) { return __dmul_rn(2.0, __drcp_rn(b)); }
    static __device__ __forceinline__ double big() { return 1e308; }
    static __device__ __forceinline__ double lg(double a) { return log(a); }
    // no NaNs on the min-sum path, so plain compare-select (fmin/fmax cost ~8 instructions each)
    static __device__ __forceinline__ double mn(double a, double b) { return a < b ? a : b; }
    static __device__ __forceinline__ double mx(double a, double b) { return a < b ? b : a; }
    static __device__ __forceinline__ double flip(double a, unsigned s)   // s ? -a : a, exact
    {
        return __hiloint2double(__double2hiint(a) ^ (int)(s << 31), __double2loint(a));
    }
};
template <> struct Ar<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float two_over(float b) { return __fmul_rn(2.0f, __frcp_rn(b)); }
    static __device__ __forceinline__ float big() { return 3.0e38f; }
    static __device__ __forceinline__ float lg(float a) { return logf(a); }
    static __device__ __forceinline__ float mn(float a, float b) { return fminf(a, b); }
    static __device__ __forceinline__ float mx(float a, float b) { return fmaxf(a, b); }
    static __device__ __forceinline__ float flip(float a, unsigned s)
    {
        return __int_as_float(__float_as_int(a) ^ (int)(s << 31));
    }
};

struct BpCtl {              // per-slot control block, shared memory (cold fields)
    long long shot;         // global shot index, -1 = slot idle
    int unsat[2];           // unsatisfied checks seen by the check pass of CTA iteration T (buffer T & 1)
    int born;               // hot word read by every work item: the CTA iteration T in which the slot was
                            // (re)filled, BP_IDLE if it holds no shot.  Written only between barriers.
                            //   fresh  <=> born == T   (filled after T's check pass: T's bit pass skips it)
                            //   iterations completed before the check pass of T = T - born - 1
    int pad;
};
constexpr int BP_IDLE = 0x7fffffff;

constexpr int QD_BP_MIN_SUM_K = 3;

struct BpSmemLayout {
    size_t msg, flags, vidx, prior, bits, ctl, total;
};

// shared-memory carve-up, identical on host and device
__host__ __device__ inline BpSmemLayout bp_smem_layout(int npos, int m, int n, int fb, int nwn, int nwm, int slots,
                                                       int real_bytes, int idx_bytes_per_col, bool stage_idx)
{
    BpSmemLayout L;
    size_t o = 0;
    L.msg = o;   o += ((size_t)slots * npos * real_bytes + 15) & ~(size_t)15;
    L.flags = o; o += ((size_t)slots * m * fb + 15) & ~(size_t)15;
    L.vidx = o;  o += stage_idx ? (((size_t)n * idx_bytes_per_col + 15) & ~(size_t)15) : 0;
    L.prior = o; o += stage_idx ? (((size_t)n * real_bytes + 15) & ~(size_t)15) : 0;
    L.bits = o;  o += ((size_t)slots * (nwn + 2 * nwm) * 4 + 15) & ~(size_t)15;
    L.ctl = o;   o += (size_t)slots * sizeof(BpCtl);
    L.total = o;
    return L;
}

// ---- explicit 32-bit shared-memory addressing: keeps address arithmetic to one IMAD/LEA per access
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double lds_real(uint32_t a, double) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ float lds_real(uint32_t a, float) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_real(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_real(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned long long lds_u64(uint32_t a) { unsigned long long v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u64(uint32_t a, unsigned long long v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint2 lds_v2(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds_v4(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }

template <typename real> __device__ __forceinline__ real pow2_neg(int it);   // 2^-it, exact
template <> __device__ __forceinline__ double pow2_neg<double>(int it) { return it < 1023 ? __hiloint2double((1023 - it) << 20, 0) : 0.0; }
template <> __device__ __forceinline__ float pow2_neg<float>(int it) { return it < 127 ? __int_as_float((127 - it) << 23) : 0.0f; }


// ---- bit-node update of one column with a compile-time edge count D (no predication), split into
// load / compute / store so that two slots can be interleaved by the instruction scheduler.
template <typename real, int D>
__device__ __forceinline__ void bit_load(const uint32_t (&qa)[D], real (&cv)[D])
{
#pragma unroll
    for (int t = 0; t < D; ++t) cv[t] = lds_real(qa[t], real());
}
// cv: check->bit messages in; bit->check messages out.  Returns the hard decision.
template <typename real, int METHOD, int D>
__device__ __forceinline__ unsigned bit_compute(real (&cv)[D], real pj, real &soft)
{
    using A = Ar<real>;
    real pre[D];
    real temp = pj;
    unsigned dbit;
    if (METHOD == QD_BP_MIN_SUM_K) {
#pragma unroll
        for (int t = 0; t < D; ++t) { pre[t] = temp; temp = A::add(temp, cv[t]); }
        dbit = temp <= (real)0 ? 1u : 0u;
        soft = temp;
        real suf = (real)0;
#pragma unroll
        for (int t = D - 1; t >= 0; --t) { const real c = cv[t]; cv[t] = A::add(pre[t], suf); suf = A::add(suf, c); }
    } else {
#pragma unroll
        for (int t = 0; t < D; ++t) {
            pre[t] = temp; temp = A::mul(temp, cv[t]);
            if (temp != temp) temp = (real)1;
        }
        dbit = temp >= (real)1 ? 1u : 0u;
        soft = temp;
        real suf = (real)1;
#pragma unroll
        for (int t = D - 1; t >= 0; --t) {
            const real c = cv[t];
            cv[t] = A::mul(pre[t], suf); suf = A::mul(suf, c);
            if (suf != suf) suf = (real)1;
        }
    }
    return dbit;
}
template <typename real, int D>
__device__ __forceinline__ void bit_store(const uint32_t (&qa)[D], const real (&cv)[D])
{
#pragma unroll
    for (int t = 0; t < D; ++t) sts_real(qa[t], cv[t]);
}

// |x| and sign bit through the integer pipe (keeps the fp64 pipe for compares)
__device__ __forceinline__ double abs_bits(double v) { return __hiloint2double(__double2hiint(v) & 0x7fffffff, __double2loint(v)); }
__device__ __forceinline__ float abs_bits(float v) { return __int_as_float(__float_as_int(v) & 0x7fffffff); }
__device__ __forceinline__ unsigned hi_bits(double v) { return (unsigned)__double2hiint(v); }
__device__ __forceinline__ unsigned hi_bits(float v) { return (unsigned)__float_as_int(v); }
__device__ __forceinline__ double xor_hi(double v, unsigned m) { return __hiloint2double(__double2hiint(v) ^ (int)m, __double2loint(v)); }
__device__ __forceinline__ float xor_hi(float v, unsigned m) { return __int_as_float(__float_as_int(v) ^ (int)m); }


// ---- min-sum check-node update of a row of compile-time degree RD, register resident.
// v: bit->check messages in, check->bit messages out.  sign0 = syndrome bit << 31.
// Sign handling uses the sign BIT instead of (v <= 0): the two differ only for v = +0.0, and then
// min1 = 0, so every message whose sign would change is a zero, whose sign no later comparison,
// sum or output can observe.
template <typename real, int RD>
__device__ __forceinline__ void check_ms_compute(real (&v)[RD], unsigned sign0, real al)
{
    using A = Ar<real>;
    unsigned sx = sign0;
    unsigned hk[RD];
#pragma unroll
    for (int k = 0; k < RD; ++k) { hk[k] = hi_bits(v[k]); sx ^= hk[k]; v[k] = abs_bits(v[k]); }
    real m1, m2;
    if (RD >= 2) {
        const bool lt = v[1] < v[0];
        m1 = lt ? v[1] : v[0];
        m2 = lt ? v[0] : v[1];
    } else { m1 = v[0]; m2 = A::big(); }
#pragma unroll
    for (int k = 2; k < RD; ++k) {
        const real a = v[k];
        const bool lt1 = a < m1, lt2 = a < m2;
        m2 = lt1 ? m1 : (lt2 ? a : m2);
        m1 = lt1 ? a : m1;
    }
    const real m1a = A::mul(m1, al), m2a = A::mul(m2, al);
#pragma unroll
    for (int k = 0; k < RD; ++k) {
        const real out = v[k] == m1 ? m2a : m1a;
        v[k] = xor_hi(out, (sx ^ hk[k]) & 0x80000000u);
    }
}

template <typename real, int METHOD, int RD, int VMAX, typename IT, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) bp_kernel(const BpParams p)
{
    using A = Ar<real>;
    constexpr int DMAX = RD > 0 ? RD : 16;           // register tile of a row (generic rows: <= 16 for product-sum)
    constexpr IT NONE = (IT)~(IT)0;
    constexpr int RB = (int)sizeof(real);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GraphDev &g = p.g;
    const int S_all = p.slots;
    const int lane = threadIdx.x & 31;
    const int P = RD > 0 ? (RD | 1) : g.pitch;
    const int FB = g.fb;

    const BpSmemLayout L = bp_smem_layout(g.npos, g.m, g.n, FB, g.nwn, g.nwm, S_all, RB, VMAX * (int)sizeof(IT),
                                          p.stage_idx != 0);
    const real *prior = static_cast<const real *>(p.prior);
    const IT *vidx = static_cast<const IT *>(p.vidx);
    if (p.stage_idx) {                                       // whole CTA: stage the static tables once
        IT *sv = reinterpret_cast<IT *>(smem_raw + L.vidx);
        real *sp = reinterpret_cast<real *>(smem_raw + L.prior);
        for (int i = threadIdx.x; i < g.n * VMAX; i += blockDim.x) sv[i] = vidx[i];
        for (int i = threadIdx.x; i < g.n; i += blockDim.x) sp[i] = prior[i];
    }
    __syncthreads();                                         // the only CTA-wide barrier

    // ---- warp groups: the CTA's warps are split into `groups` teams, each owning S_all/groups slots
    // and synchronising on its own named barrier.  Teams drift out of phase, so one team's
    // shared-memory bursts overlap another team's arithmetic.
    const int G = p.groups;
    const int NT = (int)blockDim.x / G;                      // threads of a team (multiple of 32)
    const int gid = (int)threadIdx.x / NT;
    const int tid = (int)threadIdx.x - gid * NT;
    const int S = S_all / G;                                 // slots of a team
    const int s_lo = gid * S;
    auto gsync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(gid + 1), "r"(NT) : "memory"); };
    auto gsync_or = [&](int f) -> int {
        int r;
        asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.s32 q, %1, 0;\n\tbar.red.or.pred p, %2, %3, q;\n\tselp.s32 %0, 1, 0, p;\n\t}"
                     : "=r"(r) : "r"(f), "r"(gid + 1), "r"(NT) : "memory");
        return r;
    };
    // team-local views of the per-slot arrays
    real *msg = reinterpret_cast<real *>(smem_raw + L.msg) + (size_t)s_lo * g.npos;
    unsigned char *flags = smem_raw + L.flags + (size_t)s_lo * g.m * FB;      // [S][m][FB]
    uint32_t *dec_all = reinterpret_cast<uint32_t *>(smem_raw + L.bits);        // [S_all][nwn]
    uint32_t *syn_all = dec_all + (size_t)S_all * g.nwn;                       // [S_all][nwm] original syndrome
    uint32_t *res_all = syn_all + (size_t)S_all * g.nwm;                       // [S_all][nwm] s ^ H e_hat
    uint32_t *dec = dec_all + (size_t)s_lo * g.nwn, *syn = syn_all + (size_t)s_lo * g.nwm,
             *res = res_all + (size_t)s_lo * g.nwm;
    BpCtl *ctl = reinterpret_cast<BpCtl *>(smem_raw + L.ctl) + s_lo;
    if (p.stage_idx) {
        vidx = reinterpret_cast<const IT *>(smem_raw + L.vidx);
        prior = reinterpret_cast<const real *>(smem_raw + L.prior);
    }
    // 32-bit shared addresses of the hot arrays
    const uint32_t a_msg = smem_u32(msg), a_flags = smem_u32(flags), a_dec = smem_u32(dec), a_syn = smem_u32(syn),
                   a_res = smem_u32(res), a_born = smem_u32(&ctl[0].born),
                   a_vidx = smem_u32(smem_raw + L.vidx), a_prior = smem_u32(smem_raw + L.prior);
    const uint32_t msg_stride = (uint32_t)g.npos * RB, flag_stride = (uint32_t)g.m * FB;
    const real alpha_fixed = static_cast<const real *>(p.alpha)[0];      // 0 => alpha_it = 1 - 2^-it
    const bool staged = p.stage_idx != 0;

    const int m32 = (g.m + 31) & ~31, n32 = (g.n + 31) & ~31;
    const int flag_words = (int)(((size_t)g.m * FB) >> 3);              // 8-byte flag words per slot

    long long c_fail = 0, c_shots = 0, c_nc = 0, c_it = 0;   // meaningful on thread 0 of a team only

    if (tid < S) {
        const unsigned long long id = atomicAdd(p.next_shot, 1ULL);
        ctl[tid].shot = id < (unsigned long long)p.shots ? (long long)id : -1;
        ctl[tid].unsat[0] = ctl[tid].unsat[1] = 0;
        ctl[tid].born = id < (unsigned long long)p.shots ? -1 : BP_IDLE;   // -1: part of iteration 0 in full
    }
    gsync();

    // ---- (re)load every slot born in `stamp`: syndrome, cleared decisions/flags, bit->check messages = priors
    auto load_slots = [&](int stamp) {
        for (int s = 0; s < S; ++s) {
            if (ctl[s].born != stamp) continue;
            const uint32_t *src = p.synd + ctl[s].shot * (long long)p.synd_stride;
            for (int w = tid; w < g.nwm; w += NT) {
                const uint32_t v = src[w];
                syn[s * g.nwm + w] = v;
                res[s * g.nwm + w] = v;
            }
            for (int w = tid; w < g.nwn; w += NT) dec[s * g.nwn + w] = 0u;
            unsigned long long *fw = reinterpret_cast<unsigned long long *>(flags + (size_t)s * g.m * FB);
            for (int w = tid; w < flag_words; w += NT) fw[w] = 0ULL;
            real *M = msg + (size_t)s * g.npos;
            for (int j = tid; j < g.n; j += NT) {
                const real pj = prior[j];
#pragma unroll
                for (int t = 0; t < VMAX; ++t) {
                    const IT q = vidx[j * VMAX + t];
                    if (q != NONE) M[q] = pj;
                }
            }
        }
    };
    load_slots(-1);
    gsync();
    bool any_shot = false;
    for (int s = 0; s < S; ++s) any_shot |= ctl[s].born != BP_IDLE;

    for (int T = 0; any_shot; ++T) {
        const int ub = T & 1;
        // ================================================================== check pass
        // row messages bit->check ==> check->bit in place; fold decision flips into the parity
        {
            for (int c = tid; c < m32; c += NT) {                 // a thread owns its rows in every slot
              int pos0 = 0, deg = 0;
              if (c < g.m) {
                  if (RD > 0) { pos0 = c * P; deg = RD; }
                  else { const uint32_t d = g.row_desc[c]; pos0 = d & 0xFFFFFF; deg = d >> 24; }
              }
              if (METHOD == QD_BP_MIN_SUM_K && RD > 0 && RD <= 8) {
                constexpr int RDN = RD > 0 ? RD : 1;
                // fast path: two slots at a time, loads / arithmetic / stores of both in one basic block
                const bool valid = c < g.m;
                const uint32_t rowoff = (uint32_t)(valid ? pos0 : 0) * RB, floff = (uint32_t)(valid ? c : 0) * FB;
                const uint32_t woff = (uint32_t)(c >> 5) * 4u;
                auto alpha_of = [&](int born) -> real {
                    return alpha_fixed != (real)0 ? alpha_fixed : A::sub((real)1, pow2_neg<real>(T - born));
                };
                auto finish = [&](int s, uint32_t wofs, uint32_t resw, unsigned newbit) {
                    const uint32_t neww = __ballot_sync(0xffffffffu, valid ? newbit : 0u);
                    if (lane == 0) {
                        if (neww != resw) sts_u32(a_res + wofs, neww);
                        if (neww) atomicAdd(&ctl[s].unsat[ub], __popc(neww));
                    }
                };
                for (int s = 0; s < S; s += 2) {
                    const int born0 = (int)lds_u32(a_born + s * (uint32_t)sizeof(BpCtl));
                    const int born1 = s + 1 < S ? (int)lds_u32(a_born + (s + 1) * (uint32_t)sizeof(BpCtl)) : BP_IDLE;
                    const bool act0 = born0 != BP_IDLE, act1 = born1 != BP_IDLE;      // warp-uniform
                    const uint32_t row0 = a_msg + s * msg_stride + rowoff, row1 = row0 + msg_stride;
                    const uint32_t fa0 = a_flags + s * flag_stride + floff, fa1 = fa0 + flag_stride;
                    const uint32_t w0 = (uint32_t)(s * g.nwm) * 4u + woff, w1 = w0 + (uint32_t)g.nwm * 4u;
                    if (act0 && act1) {
                        real v0[RDN], v1[RDN];
                        const uint32_t resw0 = lds_u32(a_res + w0), resw1 = lds_u32(a_res + w1);
                        const uint32_t synw0 = lds_u32(a_syn + w0), synw1 = lds_u32(a_syn + w1);
                        const unsigned long long f0 = lds_u64(fa0), f1 = lds_u64(fa1);
#pragma unroll
                        for (int k = 0; k < RDN; ++k) v0[k] = lds_real(row0 + k * RB, real());
#pragma unroll
                        for (int k = 0; k < RDN; ++k) v1[k] = lds_real(row1 + k * RB, real());
                        const unsigned nb0 = ((resw0 >> lane) ^ (unsigned)__popcll(f0)) & 1u;
                        const unsigned nb1 = ((resw1 >> lane) ^ (unsigned)__popcll(f1)) & 1u;
                        check_ms_compute<real, RDN>(v0, ((synw0 >> lane) & 1u) << 31, alpha_of(born0));
                        check_ms_compute<real, RDN>(v1, ((synw1 >> lane) & 1u) << 31, alpha_of(born1));
                        if (valid) {
#pragma unroll
                            for (int k = 0; k < RDN; ++k) sts_real(row0 + k * RB, v0[k]);
#pragma unroll
                            for (int k = 0; k < RDN; ++k) sts_real(row1 + k * RB, v1[k]);
                            if (f0) sts_u64(fa0, 0ULL);
                            if (f1) sts_u64(fa1, 0ULL);
                        }
                        finish(s, w0, resw0, nb0);
                        finish(s + 1, w1, resw1, nb1);
                    } else if (act0 || act1) {
                        const int sa = act0 ? s : s + 1;
                        const uint32_t row = act0 ? row0 : row1, fa = act0 ? fa0 : fa1, w = act0 ? w0 : w1;
                        real v[RDN];
                        const uint32_t resw = lds_u32(a_res + w), synw = lds_u32(a_syn + w);
                        const unsigned long long f = lds_u64(fa);
#pragma unroll
                        for (int k = 0; k < RDN; ++k) v[k] = lds_real(row + k * RB, real());
                        const unsigned nb = ((resw >> lane) ^ (unsigned)__popcll(f)) & 1u;
                        check_ms_compute<real, RDN>(v, ((synw >> lane) & 1u) << 31, alpha_of(act0 ? born0 : born1));
                        if (valid) {
#pragma unroll
                            for (int k = 0; k < RDN; ++k) sts_real(row + k * RB, v[k]);
                            if (f) sts_u64(fa, 0ULL);
                        }
                        finish(sa, w, resw, nb);
                    }
                }
              } else
              for (int s = 0; s < S; ++s) {
                const int born = (int)lds_u32(a_born + s * (uint32_t)sizeof(BpCtl));
                if (born != BP_IDLE) {                                       // warp-uniform
                    const uint32_t wofs = (uint32_t)(s * g.nwm + (c >> 5)) * 4u;
                    const uint32_t resw = lds_u32(a_res + wofs);
                    unsigned newbit = 0u;
                    if (c < g.m) {
                        const uint32_t synw = lds_u32(a_syn + wofs);
                        const uint32_t row = a_msg + s * msg_stride + (uint32_t)pos0 * RB;
                        // ---- parity of this check under the current hard decisions
                        const uint32_t fa = a_flags + s * flag_stride + (uint32_t)c * FB;
                        unsigned flipped = 0u;
                        if (RD > 0 && RD <= 8) {
                            const unsigned long long f = lds_u64(fa);
                            if (f) { flipped = (unsigned)__popcll(f); sts_u64(fa, 0ULL); }
                        } else {
#pragma unroll 1
                            for (int w = 0; w < (FB >> 3); ++w) {
                                const unsigned long long f = lds_u64(fa + 8 * w);
                                if (f) { flipped ^= (unsigned)__popcll(f); sts_u64(fa + 8 * w, 0ULL); }
                            }
                        }
                        newbit = ((resw >> lane) ^ flipped) & 1u;
                        const unsigned sbit = (synw >> lane) & 1u;
                        if (METHOD == QD_BP_MIN_SUM_K) {
                            const real al = alpha_fixed != (real)0 ? alpha_fixed
                                                                   : A::sub((real)1, pow2_neg<real>(T - born));
                            if (RD > 0) {
                                real v[DMAX];
#pragma unroll
                                for (int k = 0; k < RD; ++k) v[k] = lds_real(row + k * RB, real());
                                check_ms_compute<real, DMAX>(v, sbit << 31, al);
#pragma unroll
                                for (int k = 0; k < RD; ++k) sts_real(row + k * RB, v[k]);
                            } else {
                                real m1 = A::big(), m2 = A::big();
                                unsigned par = sbit;
                                for (int k = 0; k < deg; ++k) {
                                    const real v = lds_real(row + k * RB, real());
                                    par ^= v <= (real)0 ? 1u : 0u;
                                    const real a = fabs(v);
                                    const bool lt1 = a < m1, lt2 = a < m2;
                                    m2 = lt1 ? m1 : (lt2 ? a : m2);
                                    m1 = lt1 ? a : m1;
                                }
                                const real m1a = A::mul(m1, al), m2a = A::mul(m2, al);
                                for (int k = 0; k < deg; ++k) {
                                    const real v = lds_real(row + k * RB, real());
                                    const real out = fabs(v) == m1 ? m2a : m1a;
                                    sts_real(row + k * RB, A::flip(out, par ^ (v <= (real)0 ? 1u : 0u)));
                                }
                            }
                        } else {
                            // product-sum, ratio domain: prefix products left in place, then the suffix sweep
                            real t[DMAX];
                            real temp = sbit ? (real)-1 : (real)1;
#pragma unroll
                            for (int k = 0; k < DMAX; ++k) {
                                if (k < deg) {
                                    const real v = lds_real(row + k * RB, real());
                                    t[k] = A::sub(A::two_over(A::add((real)1, v)), (real)1);
                                    sts_real(row + k * RB, temp);
                                    temp = A::mul(temp, t[k]);
                                }
                            }
                            temp = (real)1;
#pragma unroll
                            for (int k = DMAX - 1; k >= 0; --k) {
                                if (k < deg) {
                                    const real x = A::mul(lds_real(row + k * RB, real()), temp);
                                    sts_real(row + k * RB, A::div(A::sub((real)1, x), A::add((real)1, x)));
                                    temp = A::mul(temp, t[k]);
                                }
                            }
                        }
                    }
                    const uint32_t neww = __ballot_sync(0xffffffffu, newbit);
                    if (lane == 0) {
                        if (neww != resw) sts_u32(a_res + wofs, neww);
                        if (neww) atomicAdd(&ctl[s].unsat[ub], __popc(neww));
                    }
                }
              }
            }
        }
        gsync();

        // ================================================================== retire / refill
        bool any_done = false;
        for (int s = 0; s < S; ++s) {
            const int born = ctl[s].born, it = T - born - 1;
            any_done |= born != BP_IDLE && it >= 1 && (ctl[s].unsat[ub] == 0 || it >= p.max_iter);
        }
        if (any_done) {
            for (int s = 0; s < S; ++s) {
                const int born = ctl[s].born, it = T - born - 1;
                if (born == BP_IDLE || it < 1) continue;
                const int conv = ctl[s].unsat[ub] == 0;
                if (!conv && it < p.max_iter) continue;
                const long long shot = ctl[s].shot;
                uint32_t *D = dec + s * g.nwn;
                uint32_t *dst = p.corr + shot * (long long)p.corr_stride;
                for (int w = tid; w < p.corr_stride; w += NT) dst[w] = w < g.nwn ? D[w] : 0u;
                int f = 0;
                if (p.err) {
                    const uint32_t *e = p.err + shot * (long long)p.err_stride;
                    gsync();
                    for (int w = tid; w < g.nwn; w += NT) D[w] ^= e[w];
                    gsync();
                    for (int i = tid; i < p.k; i += NT) {
                        uint32_t acc = 0u;
                        for (int w = 0; w < g.nwn; ++w) acc ^= p.lzT[(size_t)w * p.k + i] & D[w];
                        f |= __popc(acc) & 1;
                    }
                    f = gsync_or(f);
                }
                if (tid == 0) {
                    if (p.conv) p.conv[shot] = (uint8_t)conv;
                    if (p.iters) p.iters[shot] = it;
                    if (p.fail) p.fail[shot] = (uint8_t)f;
                    c_fail += f; c_shots += 1; c_nc += !conv; c_it += it;
                }
            }
            gsync();
            if (tid < S) {
                const int born = ctl[tid].born, it = T - born - 1;
                const bool done = born != BP_IDLE && it >= 1 && (ctl[tid].unsat[ub] == 0 || it >= p.max_iter);
                if (done) {
                    const unsigned long long id = atomicAdd(p.next_shot, 1ULL);
                    const bool more = id < (unsigned long long)p.shots;
                    ctl[tid].shot = more ? (long long)id : -1;
                    ctl[tid].unsat[0] = ctl[tid].unsat[1] = 0;
                    ctl[tid].born = more ? T : BP_IDLE;
                }
            }
            gsync();
            bool any_active = false;
            for (int s = 0; s < S; ++s) any_active |= ctl[s].born != BP_IDLE;
            if (!any_active) break;
            load_slots(T);
            // no barrier needed here: the bit pass below skips slots born in T, and the barrier that ends
            // the bit pass orders these stores before the next check pass
        }

        // ================================================================== bit pass
        // column messages check->bit ==> bit->check in place; hard decisions; raise flip flags
        {
            for (int j = tid; j < n32; j += NT) {                 // a thread owns its columns in every slot
              alignas(16) IT pos[VMAX];
              uint32_t qo[VMAX];                                    // byte offsets of the column's messages
              real pj = (real)0;
              int cd = 0;
#pragma unroll
              for (int t = 0; t < VMAX; ++t) pos[t] = NONE;
              if (j < g.n) {
                  if (staged) {
                      if (VMAX * sizeof(IT) == 8) *reinterpret_cast<uint2 *>(pos) = lds_v2(a_vidx + (uint32_t)j * 8u);
                      else if (VMAX * sizeof(IT) == 16) *reinterpret_cast<uint4 *>(pos) = lds_v4(a_vidx + (uint32_t)j * 16u);
                      else {
#pragma unroll
                          for (int t = 0; t < VMAX; t += 16 / (int)sizeof(IT))
                              *reinterpret_cast<uint4 *>(pos + t) = lds_v4(a_vidx + ((uint32_t)j * VMAX + t) * (uint32_t)sizeof(IT));
                      }
                      pj = lds_real(a_prior + (uint32_t)j * RB, real());
                  } else {
#pragma unroll
                      for (int t = 0; t < VMAX; ++t) pos[t] = vidx[(size_t)j * VMAX + t];
                      pj = prior[j];
                  }
                  cd = __ldg(p.chunk_deg + (j >> 5));
              }
#pragma unroll
              for (int t = 0; t < VMAX; ++t) qo[t] = (uint32_t)pos[t] * RB;
              // ---- decision word of the warp + flip flags for one slot
              auto publish = [&](int s, uint32_t mbase, unsigned dbit, int deg) {
                  const uint32_t neww = __ballot_sync(0xffffffffu, dbit);
                  const uint32_t da = a_dec + (uint32_t)(s * g.nwn + (j >> 5)) * 4u;
                  const uint32_t flips = lds_u32(da) ^ neww;
                  if (flips) {                                                // warp-uniform
                      if ((flips >> lane) & 1u) {
                          const uint32_t fbase = a_flags + s * flag_stride;
#pragma unroll
                          for (int t = 0; t < VMAX; ++t)
                              if (t < deg) {
                                  const uint32_t q = (uint32_t)pos[t];           // message position
                                  uint32_t fp;
                                  if (RD > 0) {                                // row * FB + k  =  q + (FB - P) * row
                                      constexpr int Pc = RD | 1, FBc = RD <= 8 ? 8 : 16;
                                      fp = q + (uint32_t)(FBc - Pc) * (q / (uint32_t)Pc);
                                  } else fp = g.fpos_of_pos[q];
                                  sts_u8(fbase + fp, 1u);
                              }
                      }
                      __syncwarp();
                      if (lane == 0) sts_u32(da, neww);
                  }
              };
              auto soft_out = [&](int s, real soft) {
                  if (p.llr && j < g.n && (p.llr_mode == 0 || T - ctl[s].born == p.max_iter))
                      p.llr[ctl[s].shot * (long long)g.n + j] =
                          (METHOD == QD_BP_MIN_SUM_K || p.llr_mode) ? (double)soft : (double)A::lg(A::div((real)1, soft));
              };
              // ---- uniform-degree chunk: unpredicated update, two slots interleaved
              auto run_uniform = [&](auto dtag) {
                  constexpr int D = decltype(dtag)::value;
                  const bool valid = j < g.n;
                  uint32_t qb[D];
#pragma unroll
                  for (int t = 0; t < D; ++t) qb[t] = valid ? qo[t < VMAX ? t : 0] : 0u;
                  for (int s = 0; s < S; s += 2) {
                      const int born0 = (int)lds_u32(a_born + s * (uint32_t)sizeof(BpCtl));
                      const int born1 = s + 1 < S ? (int)lds_u32(a_born + (s + 1) * (uint32_t)sizeof(BpCtl)) : BP_IDLE;
                      const bool act0 = born0 != BP_IDLE && born0 != T, act1 = born1 != BP_IDLE && born1 != T;
                      const uint32_t mb0 = a_msg + s * msg_stride, mb1 = mb0 + msg_stride;
                      if (act0 && act1) {
                          uint32_t q0[D], q1[D];
#pragma unroll
                          for (int t = 0; t < D; ++t) { q0[t] = mb0 + qb[t]; q1[t] = mb1 + qb[t]; }
                          real c0[D], c1[D], soft0, soft1;
                          bit_load<real, D>(q0, c0);
                          bit_load<real, D>(q1, c1);
                          const unsigned d0 = bit_compute<real, METHOD, D>(c0, pj, soft0);
                          const unsigned d1 = bit_compute<real, METHOD, D>(c1, pj, soft1);
                          if (valid) { bit_store<real, D>(q0, c0); bit_store<real, D>(q1, c1); }
                          soft_out(s, soft0);
                          soft_out(s + 1, soft1);
                          publish(s, mb0, valid ? d0 : 0u, D);
                          publish(s + 1, mb1, valid ? d1 : 0u, D);
                      } else if (act0 || act1) {
                          const int sa = act0 ? s : s + 1;
                          const uint32_t mb = act0 ? mb0 : mb1;
                          uint32_t q[D];
#pragma unroll
                          for (int t = 0; t < D; ++t) q[t] = mb + qb[t];
                          real c[D], soft;
                          bit_load<real, D>(q, c);
                          const unsigned d = bit_compute<real, METHOD, D>(c, pj, soft);
                          if (valid) bit_store<real, D>(q, c);
                          soft_out(sa, soft);
                          publish(sa, mb, valid ? d : 0u, D);
                      }
                  }
              };
              // cd: common degree of the 32 columns of this chunk (0 = mixed); warp-uniform
              const int cdw = __shfl_sync(0xffffffffu, cd, 0);
              if (cdw == 3 && VMAX >= 3) run_uniform(std::integral_constant<int, 3>());
              else if (cdw == 4 && VMAX >= 4) run_uniform(std::integral_constant<int, 4>());
              else if (cdw == 5 && VMAX >= 5) run_uniform(std::integral_constant<int, 5>());
              else if (cdw == 6 && VMAX >= 6) run_uniform(std::integral_constant<int, 6>());
              else
              for (int s = 0; s < S; ++s) {
                const int born = (int)lds_u32(a_born + s * (uint32_t)sizeof(BpCtl));
                if (born != BP_IDLE && born != T) {                          // warp-uniform
                    const uint32_t mbase = a_msg + s * msg_stride;
                    unsigned dbit = 0u;
                    int deg = 0;
                    real soft = (real)0;
                    if (j < g.n) {
                        // mixed-degree chunk: predicated sweep over the column tile
                        uint32_t qa[VMAX];
#pragma unroll
                        for (int t = 0; t < VMAX; ++t) qa[t] = mbase + qo[t];
                        real cv[VMAX], pre[VMAX];
                        real temp = pj;
#pragma unroll
                        for (int t = 0; t < VMAX; ++t) {
                            if (pos[t] != NONE) { cv[t] = lds_real(qa[t], real()); deg = t + 1; }
                        }
                        if (METHOD == QD_BP_MIN_SUM_K) {
#pragma unroll
                            for (int t = 0; t < VMAX; ++t)
                                if (t < deg) { pre[t] = temp; temp = A::add(temp, cv[t]); }
                            dbit = temp <= (real)0 ? 1u : 0u;
                            real suf = (real)0;
#pragma unroll
                            for (int t = VMAX - 1; t >= 0; --t)
                                if (t < deg) { sts_real(qa[t], A::add(pre[t], suf)); suf = A::add(suf, cv[t]); }
                        } else {
#pragma unroll
                            for (int t = 0; t < VMAX; ++t)
                                if (t < deg) {
                                    pre[t] = temp; temp = A::mul(temp, cv[t]);
                                    if (temp != temp) temp = (real)1;
                                }
                            dbit = temp >= (real)1 ? 1u : 0u;
                            real suf = (real)1;
#pragma unroll
                            for (int t = VMAX - 1; t >= 0; --t)
                                if (t < deg) {
                                    sts_real(qa[t], A::mul(pre[t], suf)); suf = A::mul(suf, cv[t]);
                                    if (suf != suf) suf = (real)1;
                                }
                        }
                        soft = temp;
                    }
                    soft_out(s, soft);
                    publish(s, mbase, dbit, deg);
                }
              }
            }
        }
        if (tid < S) ctl[tid].unsat[ub ^ 1] = 0;          // buffer of the next check pass; nobody reads it now
        gsync();
    }

    if (tid == 0 && p.counters && c_shots) {
        atomicAdd((unsigned long long *)&p.counters[0], (unsigned long long)c_fail);
        atomicAdd((unsigned long long *)&p.counters[1], (unsigned long long)c_shots);
        atomicAdd((unsigned long long *)&p.counters[2], (unsigned long long)c_nc);
        atomicAdd((unsigned long long *)&p.counters[3], (unsigned long long)c_it);
    }
}

}  // namespace qd
