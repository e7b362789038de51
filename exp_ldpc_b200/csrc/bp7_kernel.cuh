// bp7_kernel.cuh — flooding BP for row-regular Tanner graphs, "thread-owned state" variant.
//
// Same computation as bp_kernel.cuh (syndrome BP of ldpc.bp_decoder as the reference calls it at
// python/qldpc/misc/_experiment.py:110-117,137-149; CPU restatement: oracle/qldpc_oracle.c), same
// persistent-CTA / slot-refill scheme, but laid out so that the hot loops are almost pure
// load-compute-store:
//
//   * S (compile time) shots per CTA share one message array organised in 128-byte lines (one
//     line = one sweep over all shared-memory banks): message e of slot s lives in line
//     index(e) * S + s at bank colour(e).  The (colour, index) placement comes from a Koenig edge
//     colouring computed on the host (bp7_layout.h) and makes EVERY access of BOTH passes bank-
//     conflict free for any Tanner graph; the slot only adds the immediate s * 128 to an address,
//     so every shared-memory access is [edge base register + immediate];
//   * a thread owns the same rows and the same column records in every iteration, so the per-row
//     and per-column bits of state live in REGISTERS: syndrome bit, running residual-syndrome bit
//     (one bit per owned row x slot) and the previous hard decision (one bit per owned column x
//     slot).  No ballots, no decision words and no syndrome words in shared memory on the hot path;
//   * a column whose decision flipped raises one flag byte per incident edge (addresses come
//     precomputed with the column record); the owning row thread folds the 8 flag bytes of its
//     (row, slot) into the residual bit with one 64-bit load + popcount.  Unsatisfied slots are
//     OR-reduced per warp (one REDUX + at most one shared atomic per warp and pass);
//   * column records are sorted by degree and every degree class is padded to a multiple of 32
//     with records that point at a scratch row, so a warp always runs an unpredicated,
//     compile-time-degree column update; rows beyond m use the same scratch row;
//   * min-sum check update = exclusive prefix/suffix minima, exactly the two sweeps of the serial
//     decoder (3(RD-2) minima per row instead of min/second-min bookkeeping);
//   * fp64: every operation is an explicitly rounded intrinsic issued in the serial order, so
//     decisions, iteration counts and soft outputs are bit-identical to the CPU restatement;
//   * a finished shot leaves its slot at once: the slot's next shot was drawn from the queue ahead of
//     time and its syndrome staged in shared memory by a service warp, and the finished shot's
//     outputs (decisions back in column order, correction, residual, logical parities, counters)
//     drain over the next two iterations beside the decoding — no global round trip and a single
//     CTA barrier on the critical path of a retire;
//   * cluster mode (template flag CL): graphs that exceed one SM but are block structured (the
//     multi-round matrices of spacetime_code.py:46-75) are split by rows over a thread-block
//     cluster; see the comment at the kernel and bp7_cluster_layout.h.
#pragma once

#include "bp_kernel.cuh"
#include "cluster.cuh"

namespace qd {

constexpr int BP7_SMAX = 8;

struct Bp7Params {
    int m, n, nwm, nwn;
    int m32;                    // rows rounded up to a multiple of 32
    int lines;                  // 128-byte message lines per slot
    int off_flags, off_table, off_out, off_ctl, off_syn, off_cb;   // shared-memory carve-up (bp7_smem_layout), bytes
    int nrec;                   // column records (degree classes padded to multiples of 32)
    int max_iter;
    int staged;                 // record table copied to shared memory
    int uniform_prior;
    int llr_mode;               // 0: posterior log-ratio after every iteration; 1: raw soft output (total LLR for
                                // min-sum, posterior ratio for product-sum) of the LAST allowed iteration only
    double prior0;              // the common prior (decoder domain) when uniform_prior
    double alpha_fixed;         // 0 => alpha_it = 1 - 2^-it
    const uint4 *table;         // VMAX/4 uint4 per record: u16 message element indices, then u16 flag offsets
    const uint4 *row_table;     // bp7_row_quads uint4 per row (m32 rows): u16 element indices of the row's messages (slot 0),
                                // then (rows of different degrees) a u16 mask of the missing edges
    const uint8_t *chunk_deg;   // [nrec / 32]
    const void *prior_rec;      // real[nrec], decoder domain
    const uint16_t *col_id;     // [nrec]; 0xFFFF = padding record
    const uint32_t *const_bits; // [nwn] decisions of columns without edges
    const uint32_t *synd; int synd_stride;
    long long shots;
    uint32_t *corr; int corr_stride;
    uint8_t *conv; int32_t *iters; double *llr;
    const uint32_t *err; int err_stride;
    const uint32_t *lzT; int k;
    uint8_t *fail;
    long long *counters;        // {failures, shots, non_converged, iterations}
    unsigned long long *next_shot;
    // cluster mode (template flag CL): a thread-block cluster of cl_C CTAs decodes S shots together.  CTA `rank` owns a
    // contiguous share of the rows and keeps their messages; every per-CTA table above is an array over the ranks
    // (stride = the common, padded size), m / m32 / nrec / lines are those common sizes.
    int cl_C;
    const int32_t *cl_row0;     // [C] first (global) row of the share
    const int32_t *cl_rows;     // [C] rows of the share
    const int32_t *cl_nexp;     // [C] export entries
    const uint2 *cl_export;     // [C][cl_exp_stride]: x = local message element | mailbox element << 16, y = destination rank
    int cl_exp_stride;
};

struct Bp7Ctl {
    long long shot[BP7_SMAX];
    int born[BP7_SMAX];         // see BpCtl::born
    unsigned unsat[2];          // bit s: slot s has an unsatisfied check (buffer T & 1)
    int pad[2];                 // pad[0] bit s: the residual of the shot draining from slot s anticommutes with a logical
    double alpha[BP7_SMAX];     // min-sum scaling factor of slot s in the coming check pass
    long long resv[BP7_SMAX];   // the slot's next shot, drawn from the queue ahead of time (-1: none)
    long long drain_shot[BP7_SMAX];   // the finished shot whose outputs are draining from slot s
    int drain_it[BP7_SMAX], drain_conv[BP7_SMAX];
    unsigned need_resv, fresh_resv;   // bit s: slot s wants a reservation / got one in this iteration (syndrome not staged yet)
    int pad2[2];
};

struct Bp7SmemLayout {
    size_t msg, flags, table, out, ctl, syn, cb, total;
    int msg_elems;              // elements of the message array (incl. scratch row)
};

// uint4 per row-table entry: RD 16-bit message positions, one more 16-bit word for the missing-edge mask of irregular rows
__host__ __device__ constexpr int bp7_row_quads(int rd, bool irregular) { return (rd + (irregular ? 1 : 0) + 7) / 8; }

// flag bytes of a row: S records of 8 (row degree <= 8) or 16 bytes; the pitch is an odd number of
// 8-byte words, which keeps the check pass conflict free
// (rows of 9 edges: the ninth edge's flag byte lives in a second array behind the records, one byte per (row, slot), rows
// interleaved in blocks of 8 so that the slots of a row stay 8 bytes apart like everywhere else)
__host__ __device__ inline int bp7_flag_record(int rd) { (void)rd; return 8; }
__host__ __device__ inline int bp7_flag_pitch(int s, int rd) { return ((s * bp7_flag_record(rd) / 8) | 1) * 8; }
// byte offset (from the flag area) of the flag of edge k of `row`, slot 0; rows = rows incl. the scratch row (m + 1)
__host__ __device__ inline int bp7_flag_offset(int row, int k, int rows, int s, int rd)
{
    if (k < 8) return row * bp7_flag_pitch(s, rd) + k;
    return rows * bp7_flag_pitch(s, rd) + (row >> 3) * 8 * s + (row & 7);
}
__host__ __device__ inline int bp7_flag_bytes(int rows, int s, int rd)
{
    return rows * bp7_flag_pitch(s, rd) + (rd > 8 ? ((rows + 7) / 8) * 8 * s : 0);
}

// lines: 128-byte lines per slot (from the layout); one more line group is scratch for padding lanes
__host__ __device__ inline Bp7SmemLayout bp7_smem_layout(int m, int nrec, int nwn, int nwm, int lines, int rd, int vmax, int s,
                                                         int real_bytes, bool staged)
{
    Bp7SmemLayout L;
    L.msg_elems = (lines + 1) * s * (128 / real_bytes);
    size_t o = 0;
    L.msg = o;   o += (size_t)(lines + 1) * s * 128;
    L.flags = o; o += (size_t)bp7_flag_bytes(m + 1, s, rd);
    o = (o + 15) & ~(size_t)15;
    L.table = o; o += staged ? (size_t)nrec * vmax * 4 : 0;
    L.out = o;   o += ((size_t)s * nwn * 4 + 15) & ~(size_t)15;
    L.syn = o;   o += ((size_t)s * nwm * 4 + 15) & ~(size_t)15;     // staged syndromes of the slots' next shots
    L.cb = o;    o += ((size_t)nwn * 4 + 15) & ~(size_t)15;         // decisions of columns without edges
    L.ctl = o;   o += sizeof(Bp7Ctl);
    L.total = o;
    return L;
}

// ---- [register + immediate] shared-memory accesses
template <int OFF> __device__ __forceinline__ double lds_off(uint32_t a, double)
{ double v; asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF)); return v; }
template <int OFF> __device__ __forceinline__ float lds_off(uint32_t a, float)
{ float v; asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF)); return v; }
template <int OFF> __device__ __forceinline__ void sts_off(uint32_t a, double v)
{ asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(a), "n"(OFF), "d"(v) : "memory"); }
template <int OFF> __device__ __forceinline__ void sts_off(uint32_t a, float v)
{ asm volatile("st.shared.f32 [%0+%1], %2;" ::"r"(a), "n"(OFF), "f"(v) : "memory"); }
template <int OFF> __device__ __forceinline__ unsigned long long lds_u64_off(uint32_t a)
{ unsigned long long v; asm volatile("ld.shared.u64 %0, [%1+%2];" : "=l"(v) : "r"(a), "n"(OFF)); return v; }
template <int OFF> __device__ __forceinline__ void sts_u64_off(uint32_t a, unsigned long long v)
{ asm volatile("st.shared.u64 [%0+%1], %2;" ::"r"(a), "n"(OFF), "l"(v) : "memory"); }
template <int OFF> __device__ __forceinline__ void sts_u8_off(uint32_t a, uint32_t v)
{ asm volatile("st.shared.u8 [%0+%1], %2;" ::"r"(a), "n"(OFF), "r"(v) : "memory"); }

// ---- min-sum check update, exclusive prefix / suffix minima (the two sweeps of the serial decoder)
// v: bit->check messages in, check->bit messages out.  sign0 = syndrome bit << 31.  Signs travel
// as sign BITS (see check_ms_compute in bp_kernel.cuh for why that equals the serial `<= 0` rule).
// value of smaller magnitude (ties keep a): the compare and the final multiply take |x| through
// operand modifiers, so no instruction is spent on absolute values
__device__ __forceinline__ double mn_abs(double a, double b) { return fabs(b) < fabs(a) ? b : a; }
__device__ __forceinline__ float mn_abs(float a, float b) { return fabsf(b) < fabsf(a) ? b : a; }

template <typename real, int RD>
__device__ __forceinline__ void check_ms7(real (&v)[RD], unsigned sign0, real al)
{
    using A = Ar<real>;
    unsigned sx = sign0;
#pragma unroll
    for (int k = 0; k < RD; ++k) sx ^= hi_bits(v[k]);
    real pm[RD];                                   // pm[k]: smallest-magnitude value among v[0..k-1], k >= 1
    pm[0] = A::big();
    if (RD > 1) pm[1] = v[0];
#pragma unroll
    for (int k = 2; k < RD; ++k) pm[k] = mn_abs(pm[k - 1], v[k - 1]);
    real sm = A::big();                            // same among v[k+1..]
    // the row's sign product rides on the multiplier: |pick| * (-al) = -(|pick| * al) bit for bit, so the
    // per-edge sign costs one LOP3 (own sign bit) instead of two
    const real als = xor_hi(al, sx & 0x80000000u);
#pragma unroll
    for (int k = RD - 1; k >= 0; --k) {
        const real pick = k == RD - 1 ? pm[k] : (k == 0 ? sm : mn_abs(pm[k], sm));
        sm = k == RD - 1 ? v[k] : mn_abs(sm, v[k]);
        const real mag = A::mul(fabs(pick), als);
        v[k] = xor_hi(mag, hi_bits(v[k]) & 0x80000000u);
    }
}

// Long rows (RD >= 8: spacetime matrices): smallest and second-smallest magnitude plus the position of the first
// smallest.  The value edge k receives is the minimum over the OTHER edges either way (second smallest at that position,
// smallest elsewhere; equal magnitudes are equal values), so the outputs are bit-identical to the two sweeps above; this
// form keeps RD + 2 values live instead of 2 RD + 1 and multiplies twice per row instead of once per edge.
template <typename real, int RD>
__device__ __forceinline__ void check_ms_m12(real (&v)[RD], unsigned sign0, real al)
{
    using A = Ar<real>;
    unsigned sx = sign0;
#pragma unroll
    for (int k = 0; k < RD; ++k) sx ^= hi_bits(v[k]);
    real m1 = v[0], m2 = A::big();
    int arg = 0;
#pragma unroll
    for (int k = 1; k < RD; ++k) {
        const bool lt1 = fabs(v[k]) < fabs(m1);
        const bool lt2 = fabs(v[k]) < fabs(m2);
        m2 = lt1 ? m1 : (lt2 ? v[k] : m2);
        m1 = lt1 ? v[k] : m1;
        arg = lt1 ? k : arg;
    }
    const real als = xor_hi(al, sx & 0x80000000u);
    const real o1 = A::mul(fabs(m1), als), o2 = A::mul(fabs(m2), als);
#pragma unroll
    for (int k = 0; k < RD; ++k) v[k] = xor_hi(k == arg ? o2 : o1, hi_bits(v[k]) & 0x80000000u);
}

// ---- product-sum check update in the probability-ratio domain, register resident
template <typename real, int RD>
__device__ __forceinline__ void check_ps7(real (&v)[RD], unsigned sbit)
{
    using A = Ar<real>;
    real t[RD], pre[RD];
    real temp = sbit ? (real)-1 : (real)1;
#pragma unroll
    for (int k = 0; k < RD; ++k) {
        t[k] = A::sub(A::two_over(A::add((real)1, v[k])), (real)1);
        pre[k] = temp;
        temp = A::mul(temp, t[k]);
    }
    temp = (real)1;
#pragma unroll
    for (int k = RD - 1; k >= 0; --k) {
        const real x = A::mul(pre[k], temp);
        v[k] = A::div(A::sub((real)1, x), A::add((real)1, x));
        temp = A::mul(temp, t[k]);
    }
}

// ---- one slot of one row: load, and compute + store.  The shared-memory accesses are volatile asm
// and keep their program order, so the callers issue the loads of the next slots by hand before the
// arithmetic of the current one (software pipeline over the slots).
template <int OFF> __device__ __forceinline__ uint32_t lds_u8_off(uint32_t a)
{ uint32_t v; asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF)); return v; }

// fb: the row's flag record, f9: its ninth-edge flag byte (RD > 8), both for slot 0
template <typename real, int METHOD, int RD, int s>
__device__ __forceinline__ void row_ld(const uint32_t (&ra)[RD], uint32_t fb, uint32_t f9, real (&v)[RD], unsigned long long &f,
                                       unsigned dummy)
{
    constexpr int FB = 8;
    f = lds_u64_off<s * FB>(fb);
    if (RD > 8) {                       // ninth flag: fold it in, keep "any flag set" in bit 63
        const unsigned g = lds_u8_off<s * FB>(f9);
        f = ((unsigned long long)((__popcll(f) + (int)g) & 1)) | ((f | g) ? 0x8000000000000002ULL : 0ULL);
    }
#pragma unroll
    for (int k = 0; k < RD; ++k) v[k] = lds_off<s * 128>(ra[k], real());
    if (dummy) {                        // rows shorter than RD: the missing edges read as the neutral message
#pragma unroll
        for (int k = 0; k < RD; ++k)
            if ((dummy >> k) & 1u) v[k] = METHOD == QD_BP_MIN_SUM_K ? Ar<real>::big() : (real)0;
    }
}

template <typename real, int METHOD, int RD, int S, int s, bool LONGROWS>
__device__ __forceinline__ void row_cs(const uint32_t (&ra)[RD], uint32_t fb, uint32_t f9, real (&v)[RD], unsigned long long f, int i,
                                       bool valid, unsigned synbits, unsigned &resbits, unsigned &un, real al)
{
    constexpr int FB = 8;
    const int bi = i * S + s;
    resbits ^= ((unsigned)__popcll(f) & 1u) << bi;      // RD > 8: f was folded by row_ld (even popcount pad)
    if (f) {
        sts_u64_off<s * FB>(fb, 0ULL);
        if (RD > 8) sts_u8_off<s * FB>(f9, 0u);
    }
    un |= (valid ? (resbits >> bi) & 1u : 0u) << s;
    const unsigned sbit = (synbits >> bi) & 1u;
    if (METHOD == QD_BP_MIN_SUM_K) {
        if (LONGROWS) check_ms_m12<real, RD>(v, sbit << 31, al);
        else check_ms7<real, RD>(v, sbit << 31, al);
    }
    else check_ps7<real, RD>(v, sbit);
#pragma unroll
    for (int k = 0; k < RD; ++k) sts_off<s * 128>(ra[k], v[k]);
}

template <typename real, int D, int VMAX, int s>
__device__ __forceinline__ void col_ld(const uint32_t (&ea)[VMAX], real (&cv)[D])
{
#pragma unroll
    for (int t = 0; t < D; ++t) cv[t] = lds_off<s * 128>(ea[t], real());
}

// Soft output of one (column record, slot): rare (single-shot calls, the first stage of bposd), so it lives out of line
// and re-derives everything it needs from the kernel parameters instead of holding registers through the bit pass.
struct Bp7LlrOut { unsigned want; const Bp7Params *p; const long long *shot; uint32_t rank; };

template <typename real, int METHOD>
__device__ __forceinline__ void bp7_write_llr(const Bp7Params *p, const long long *shots, uint32_t rank, int r, int s, real soft)
{
    using A = Ar<real>;
    const unsigned col = p->col_id[(size_t)rank * p->nrec + r];            // (rank = 0 outside cluster mode)
    if (col != 0xFFFFu)
        p->llr[shots[s] * (long long)p->n + col] =
            (METHOD == QD_BP_MIN_SUM_K || p->llr_mode) ? (double)soft : (double)A::lg(A::div((real)1, soft));
}

template <typename real, int METHOD, int D, int VMAX, int S, int FB, int s>
__device__ __forceinline__ void col_cs(const uint32_t (&ea)[VMAX], const uint32_t (&fa)[VMAX], real (&cv)[D], real pj, int i,
                                       unsigned &decbits, const Bp7LlrOut &lo, int r)
{
    using A = Ar<real>;
    const int bi = i * S + s;
    real soft;
    const unsigned dbit = bit_compute<real, METHOD, D>(cv, pj, soft);
#pragma unroll
    for (int t = 0; t < D; ++t) sts_off<s * 128>(ea[t], cv[t]);
    const unsigned flip = ((decbits >> bi) & 1u) ^ dbit;
    if (flip) {
#pragma unroll
        for (int t = 0; t < D; ++t) sts_u8_off<s * FB>(fa[t], 1u);
    }
    decbits ^= flip << bi;
    if ((lo.want >> s) & 1u) bp7_write_llr<real, METHOD>(lo.p, lo.shot, lo.rank, r, s, soft);   // rare
}

// all slots of one column record of compile-time degree D
template <typename real, int METHOD, int D, int VMAX, int S, int FB>
__device__ __forceinline__ void col_update(const uint32_t (&ea)[VMAX], const uint32_t (&fa)[VMAX], real pj, int i, unsigned run,
                                           unsigned &decbits, const Bp7LlrOut &lo, int r)
{
    constexpr unsigned SMASK = (1u << S) - 1u;
    real ca[D], cb[D], cc[D];
#define QD_LD(n, buf) col_ld<real, D, VMAX, (n) % S>(ea, buf)
#define QD_CS(n, buf) col_cs<real, METHOD, D, VMAX, S, FB, (n) % S>(ea, fa, buf, pj, i, decbits, lo, r)
    if (run == SMASK) {                                  // three buffers: two slots of loads in flight
        QD_LD(0, ca);
        if (S > 1) QD_LD(1, cb);
        if (S > 2) QD_LD(2, cc);
        QD_CS(0, ca);
        if (S > 3) QD_LD(3, ca);
        if (S > 1) QD_CS(1, cb);
        if (S > 4) QD_LD(4, cb);
        if (S > 2) QD_CS(2, cc);
        if (S > 5) QD_LD(5, cc);
        if (S > 3) QD_CS(3, ca);
        if (S > 6) QD_LD(6, ca);
        if (S > 4) QD_CS(4, cb);
        if (S > 7) QD_LD(7, cb);
        if (S > 5) QD_CS(5, cc);
        if (S > 6) QD_CS(6, ca);
        if (S > 7) QD_CS(7, cb);
    } else {
        if (S > 0 && (run & 1u)) { QD_LD(0, ca); QD_CS(0, ca); }
        if (S > 1 && (run & 2u)) { QD_LD(1, ca); QD_CS(1, ca); }
        if (S > 2 && (run & 4u)) { QD_LD(2, ca); QD_CS(2, ca); }
        if (S > 3 && (run & 8u)) { QD_LD(3, ca); QD_CS(3, ca); }
        if (S > 4 && (run & 16u)) { QD_LD(4, ca); QD_CS(4, ca); }
        if (S > 5 && (run & 32u)) { QD_LD(5, ca); QD_CS(5, ca); }
        if (S > 6 && (run & 64u)) { QD_LD(6, ca); QD_CS(6, ca); }
        if (S > 7 && (run & 128u)) { QD_LD(7, ca); QD_CS(7, ca); }
    }
#undef QD_LD
#undef QD_CS
}

// CL (cluster mode): the Tanner graph is too large for one SM.  The rows are split over the CTAs of a cluster; a column
// whose edges all lie in one share is updated there, a column cut by the split has a replica record in every CTA that
// holds one of its edges.  A replica reads the check->bit messages of its remote edges from a MAILBOX (message elements
// nobody's rows own) that the row owners fill through distributed shared memory after the check pass ("export"), runs
// the same serial sums as every other replica of the column, and keeps the outputs of its local edges.  Everything else
// — in-place message array, conflict-free placement, thread-owned state, slot refill — is the single-CTA kernel; the
// slot control block is replicated and advanced identically in every CTA (CTA 0 draws from the shot queue and assembles
// the outputs).  Arithmetic and its order are unchanged, so results are bit-identical to the single-CTA kernels.
template <typename real, int METHOD, int RD, int VMAX, int S, bool IRREG, int MAXT, int MINB, bool CL = false>
__global__ void __launch_bounds__(MAXT, MINB) bp7_kernel(const __grid_constant__ Bp7Params p)
{
    using A = Ar<real>;
    static_assert(VMAX == 4 || VMAX == 8, "column tile");
    static_assert(S >= 1 && S <= BP7_SMAX, "slots");
    constexpr int RB = (int)sizeof(real);
    constexpr int LINE = 128;                            // bytes between the slots of one message
    static_assert(RD <= 9, "flag layout: eight flags in the row's record, one more in the second array");
    constexpr int FB = 8;                                // flag bytes per (row, slot)
    constexpr int FP = ((S * FB / 8) | 1) * 8;           // flag bytes per row (bp7_flag_pitch)
    constexpr int RQ = bp7_row_quads(RD, IRREG);         // uint4 per row-table entry
    constexpr unsigned SMASK = (1u << S) - 1u;
    constexpr bool PREFETCH = S == 1;                    // table entries a turn ahead (see the check pass; no gain in cluster mode)
    constexpr bool LONGROWS = IRREG && RD >= 8;          // spacetime matrices: min / second-min check update, one row buffer
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = (int)threadIdx.x, NT = (int)blockDim.x, lane = tid & 31;

    // shared-memory carve-up comes from the host (bp7_smem_layout); the three hot base addresses are made
    // opaque so that they live in registers instead of being re-derived inside the loops
    uint32_t a_msg = smem_u32(smem_raw), a_flags = a_msg + (uint32_t)p.off_flags, a_tab = a_msg + (uint32_t)p.off_table;
    asm volatile("" : "+r"(a_msg), "+r"(a_flags), "+r"(a_tab));
    uint32_t *outbits = reinterpret_cast<uint32_t *>(smem_raw + p.off_out);      // [S][nwn]
    Bp7Ctl *ctl = reinterpret_cast<Bp7Ctl *>(smem_raw + p.off_ctl);
    const int max_iter = p.max_iter;
    const real alpha_fixed = (real)p.alpha_fixed;
    const bool staged = p.staged != 0;
    // this CTA's tables and share of the rows (cluster mode: indexed by the rank within the cluster)
    const uint32_t rank = CL ? cl_rank() : 0u;
    const uint32_t NC = CL ? (uint32_t)p.cl_C : 1u;
    const uint4 *const tbl = p.table + (CL ? (size_t)rank * p.nrec * (VMAX / 4) : 0);
    const uint4 *const rtbl = p.row_table + (CL ? (size_t)rank * p.m32 * RQ : 0);
    const uint8_t *const cdeg = p.chunk_deg + (CL ? (size_t)rank * (p.nrec / 32) : 0);
    const real *const prior_rec = static_cast<const real *>(p.prior_rec) + (CL ? (size_t)rank * p.nrec : 0);
    const uint16_t *const colid = p.col_id + (CL ? (size_t)rank * p.nrec : 0);
    const int my_m = CL ? p.cl_rows[rank] : p.m;
    const int row0 = CL ? p.cl_row0[rank] : 0;
    // slot control: one copy per CTA, written identically everywhere (cluster mode: by CTA 0 through the cluster window)
    auto set_slot = [&](int s, long long shot, int born) {
        if (!CL) { ctl->shot[s] = shot; ctl->born[s] = born; }
        else
            for (uint32_t r = 0; r < NC; ++r) {
                cl_st_s64(cl_map(smem_u32(&ctl->shot[s]), r), shot);
                cl_st_u32(cl_map(smem_u32(&ctl->born[s]), r), (uint32_t)born);
            }
    };

    uint32_t *synbuf = reinterpret_cast<uint32_t *>(smem_raw + p.off_syn);       // [S][nwm] syndrome of the slot's NEXT shot
    uint32_t *cbits = reinterpret_cast<uint32_t *>(smem_raw + p.off_cb);         // [nwn] decisions of columns without edges
    const bool lead = !CL || rank == 0;                                          // the CTA that talks to the queue and writes outputs
    const int warp = tid >> 5, svc_warp = (NT >> 5) - 1;                         // service warp: the one with the fewest rows

    if (staged) {
        uint4 *st = reinterpret_cast<uint4 *>(smem_raw + p.off_table);
        for (int i = tid; i < p.nrec * (VMAX / 4); i += NT) st[i] = tbl[i];
    }
    for (int w = tid; w < p.nwn; w += NT) cbits[w] = p.const_bits[w];
    for (int w = tid; w < S * p.nwn; w += NT) outbits[w] = 0u;
    if (tid < BP7_SMAX) { ctl->resv[tid] = -1; ctl->drain_shot[tid] = -1; }
    if (tid == 0) { ctl->unsat[0] = ctl->unsat[1] = 0u; ctl->pad[0] = 0; ctl->need_resv = 0u; ctl->fresh_resv = 0u; }
    if (CL) cluster_sync_release_all(); else __syncthreads();      // (cluster: every CTA runs before its shared memory is written)
    if (lead && tid < S) {
        const unsigned long long id = atomicAdd(p.next_shot, 1ULL);
        const bool more = id < (unsigned long long)p.shots;
        set_slot(tid, more ? (long long)id : -1, more ? -1 : BP_IDLE);     // -1: takes part in iteration 0 in full
    }
    if (CL) cluster_sync_release_all(); else __syncthreads();
    if (tid == 0) {                                                // slots that hold a shot reserve their next one
        unsigned need = 0u;
#pragma unroll
        for (int s = 0; s < S; ++s) need |= (unsigned)(ctl->born[s] != BP_IDLE) << s;
        ctl->need_resv = need;
    }

    // the slots' birth stamps, one or two 128-bit broadcast loads (the control block is read by every thread)
    auto read_born = [&](int (&born)[S]) {
        int b[BP7_SMAX];
        const uint4 lo4 = lds_v4(smem_u32(&ctl->born[0]));
        b[0] = (int)lo4.x; b[1] = (int)lo4.y; b[2] = (int)lo4.z; b[3] = (int)lo4.w;
        if (S > 4) {
            const uint4 hi4 = lds_v4(smem_u32(&ctl->born[4]));
            b[4] = (int)hi4.x; b[5] = (int)hi4.y; b[6] = (int)hi4.z; b[7] = (int)hi4.w;
        }
#pragma unroll
        for (int s = 0; s < S; ++s) born[s] = b[s];
    };
    auto active_mask = [&]() -> unsigned {
        int born[S];
        read_born(born);
        unsigned act = 0u;
#pragma unroll
        for (int s = 0; s < S; ++s) act |= (unsigned)(born[s] != BP_IDLE) << s;
        return act;
    };

    // ---- thread-owned state, bit (i * S + s): i-th owned row / column record, slot s
    unsigned synbits = 0u, resbits = 0u, decbits = 0u;
    long long c_fail = 0, c_shots = 0, c_nc = 0, c_it = 0;      // thread 0 only

    // record -> shared addresses of its edges (slot 0) and of their flip-flag bytes (slot 0)
    auto fetch_record = [&](int r, uint32_t (&w)[VMAX]) {       // VMAX/2 words of mq, then VMAX/2 of fq
#pragma unroll
        for (int q = 0; q < VMAX / 4; ++q) {
            const uint4 x = staged ? lds_v4(a_tab + (uint32_t)(r * (VMAX / 4) + q) * 16u)
                                   : __ldg(tbl + (size_t)r * (VMAX / 4) + q);
            w[4 * q + 0] = x.x; w[4 * q + 1] = x.y; w[4 * q + 2] = x.z; w[4 * q + 3] = x.w;
        }
    };
    auto decode_record = [&](const uint32_t (&w)[VMAX], uint32_t (&ea)[VMAX], uint32_t (&fa)[VMAX]) {
#pragma unroll
        for (int t = 0; t < VMAX; ++t) {
            const uint32_t mw = w[t >> 1], fw = w[VMAX / 2 + (t >> 1)];
            ea[t] = a_msg + ((t & 1) ? (mw >> 16) : (mw & 0xffffu)) * RB;
            fa[t] = a_flags + ((t & 1) ? (fw >> 16) : (fw & 0xffffu));
        }
    };
    auto load_record = [&](int r, uint32_t (&ea)[VMAX], uint32_t (&fa)[VMAX]) {
        uint32_t w[VMAX];
        fetch_record(r, w);
        decode_record(w, ea, fa);
    };
    // row -> shared addresses of its messages (slot 0), in column order
    auto fetch_row = [&](int c, uint32_t (&w)[4 * RQ]) {
#pragma unroll
        for (int q = 0; q < RQ; ++q) {
            const uint4 x = __ldg(rtbl + (size_t)c * RQ + q);
            w[4 * q + 0] = x.x; w[4 * q + 1] = x.y; w[4 * q + 2] = x.z; w[4 * q + 3] = x.w;
        }
    };
    auto decode_row = [&](const uint32_t (&w)[4 * RQ], uint32_t (&ra)[RD], unsigned &dummy) {
#pragma unroll
        for (int k = 0; k < RD; ++k) ra[k] = a_msg + ((k & 1) ? (w[k >> 1] >> 16) : (w[k >> 1] & 0xffffu)) * RB;
        // the 16-bit entry after the last edge: bit k set = no k-th edge in this row
        dummy = IRREG ? ((RD & 1) ? (w[RD >> 1] >> 16) : (w[RD >> 1] & 0xffffu)) : 0u;
    };
    auto prior_of = [&](int r) -> real { return p.uniform_prior ? (real)p.prior0 : prior_rec[r]; };

    // ---- (re)load every slot born in `stamp`.  The first shots read their syndromes from global memory; later ones find
    // them in shared memory (synbuf, staged while the slot's previous shot was running).  The previous shot's decisions
    // (decbits) stay: the bit pass of the same iteration files them and clears the bits.
    auto load_slots = [&](int stamp, bool from_global) {
        unsigned fresh = 0u;
#pragma unroll
        for (int s = 0; s < S; ++s) fresh |= (ctl->born[s] == stamp ? 1u : 0u) << s;
        if (!fresh) return;
        int i = 0;
        for (int c = tid; c < p.m32; c += NT, ++i) {
            const bool valid = c < my_m;
            const uint32_t fb = a_flags + (uint32_t)(valid ? c : p.m) * FP;
            const uint32_t f9 = a_flags + (uint32_t)bp7_flag_offset(valid ? c : p.m, 8, p.m + 1, S, RD);
            const int gc = row0 + c;                                  // the row's index in the syndrome
#pragma unroll
            for (int s = 0; s < S; ++s)
                if ((fresh >> s) & 1u) {
                    const int bi = i * S + s;
                    unsigned b = 0u;
                    if (valid) {
                        const uint32_t word = from_global ? p.synd[ctl->shot[s] * (long long)p.synd_stride + (gc >> 5)]
                                                          : synbuf[s * p.nwm + (gc >> 5)];
                        b = (word >> (gc & 31)) & 1u;
                    }
                    synbits = (synbits & ~(1u << bi)) | (b << bi);
                    resbits = (resbits & ~(1u << bi)) | (b << bi);
                    sts_u64(fb + s * FB, 0ULL);
                    if (RD > 8) sts_u8(f9 + s * FB, 0u);
                }
        }
        for (int r = tid; r < p.nrec; r += NT) {
            uint32_t ea[VMAX], fa[VMAX];
            load_record(r, ea, fa);
            const int deg = cdeg[r >> 5];
            const real pj = prior_of(r);
#pragma unroll
            for (int s = 0; s < S; ++s)
                if ((fresh >> s) & 1u) {
#pragma unroll
                    for (int t = 0; t < VMAX; ++t)
                        if (t < deg) sts_real(ea[t] + s * LINE, pj);
                }
        }
    };
    load_slots(-1, true);
    // scaling factor of slot s in check pass T (a slot born in b runs its iteration T - b there): one thread per
    // slot writes it after the bit pass, the row owners read it instead of re-deriving it per row and slot
    auto publish_alpha = [&](int T) {
        if (METHOD == QD_BP_MIN_SUM_K && tid < S)
            ctl->alpha[tid] = (double)(alpha_fixed != (real)0 ? alpha_fixed
                                                              : A::sub((real)1, pow2_neg<real>(T - ctl->born[tid])));
    };
    publish_alpha(0);
    __syncthreads();
    // warp-uniform copies of the control state every thread would otherwise re-read each pass: the slots holding a shot,
    // and the slots retired one / two iterations ago (their outputs are draining)
    unsigned act = active_mask(), d1 = 0u, d2 = 0u;
    const bool any_shot = act != 0u;

#ifdef QD_BP7_TIMING
    long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64(), n_iter = 0;
#define QD_TICK7(k) do { const long long now_ = clock64(); tacc[k] += now_ - tlast; tlast = now_; } while (0)
#else
#define QD_TICK7(k) do { } while (0)
#endif
    // A finished shot leaves its slot at once (the slot's next shot and its syndrome are already in shared memory), and
    // its outputs drain over the following two iterations beside the decoding of the other slots:
    //   iteration T   (shot found finished after the check pass)  bit pass: the column owners file its decisions into
    //                                                              the output words; its error words are requested
    //   iteration T+1 check pass: correction written, residual formed;   bit pass: logical parities of the residual
    //   iteration T+2 check pass: per-shot outputs and counters, output words cleared
    // Meanwhile the service warp draws the slot's next-but-one shot from the queue and stages its syndrome.
    uint32_t err_pf = 0u;                                      // error word (slot, word) = tid of a shot retired in the last iteration
    // corrections of the slots in `mask` out, their residuals in place (each word by one thread)
    auto drain_corrections = [&](unsigned mask) {
        if (!lead || !mask) return;
        for (int x = tid; x < S * p.corr_stride; x += NT) {
            const int s = x / p.corr_stride, w = x - s * p.corr_stride;
            if (!((mask >> s) & 1u)) continue;
            const long long shot = ctl->drain_shot[s];
            const uint32_t v = w < p.nwn ? (outbits[s * p.nwn + w] | cbits[w]) : 0u;
            p.corr[shot * (long long)p.corr_stride + w] = v;
            if (p.err && w < p.nwn)
                outbits[s * p.nwn + w] = v ^ (x == tid ? err_pf : p.err[shot * (long long)p.err_stride + w]);
        }
    };
    bool b2_pending = false;                                   // cluster mode: arrived at the end-of-iteration barrier, not yet waited
    for (int T = 0; any_shot; ++T) {
        const int ub = T & 1;
        QD_TICK7(5);
        // ================================================================== check pass
        {
            // ---- drain steps due now (lead CTA): d1 = slots retired in T - 1, d2 = in T - 2
            if (!CL) drain_corrections(d1);                   // (cluster mode: after the wait below, the decisions come from all CTAs)
            if (lead && d2) {
                for (int x = tid; x < S * p.nwn; x += NT)
                    if ((d2 >> (x / p.nwn)) & 1u) outbits[x] = 0u;
                if (tid == 0) {
                    const unsigned failing = (unsigned)ctl->pad[0];
                    for (int s = 0; s < S; ++s) {
                        if (!((d2 >> s) & 1u)) continue;
                        const int it = ctl->drain_it[s], conv = ctl->drain_conv[s];
                        const int f = (int)((failing >> s) & 1u);
                        const long long shot = ctl->drain_shot[s];
                        if (p.conv) p.conv[shot] = (uint8_t)conv;
                        if (p.iters) p.iters[shot] = it;
                        if (p.fail) p.fail[shot] = (uint8_t)f;
                        c_fail += f; c_shots += 1; c_nc += !conv; c_it += it;
                    }
                    ctl->pad[0] = (int)(failing & ~d2);
                }
            }
            // ---- queue service: the syndromes requested in the last bit pass have landed; draw the next shots
            long long resv_new = -1;
            bool resv_have = false;
            if (warp == svc_warp) {
                asm volatile("cp.async.wait_all;" ::: "memory");
                const unsigned need = ctl->need_resv;
                if (need) {
                    if (lead && lane < S && ((need >> lane) & 1u)) {
                        const unsigned long long id = atomicAdd(p.next_shot, 1ULL);
                        resv_new = id < (unsigned long long)p.shots ? (long long)id : -1;
                        resv_have = true;
                        if (!CL) ctl->resv[lane] = resv_new;  // (cluster mode: sent to every CTA after the wait below)
                    }
                    __syncwarp();
                    if (lane == 0) { ctl->fresh_resv = need; ctl->need_resv = 0u; }
                }
            }
            unsigned un = 0u;
            int i = 0;
            // single-slot geometries (graphs that fill an SM: tables in global memory, several rows and records per thread,
            // no second slot to overlap with): the table entry of the thread's NEXT row / record is requested a turn ahead
            uint32_t wrow[4 * RQ];
            if (PREFETCH && tid < p.m32) fetch_row(tid, wrow);
            for (int c = tid; c < p.m32; c += NT, ++i) {
                const bool valid = c < my_m;
                const uint32_t fb = a_flags + (uint32_t)(valid ? c : p.m) * FP;
                const uint32_t f9 = RD > 8 ? a_flags + (uint32_t)bp7_flag_offset(valid ? c : p.m, 8, p.m + 1, S, RD) : 0u;
                uint32_t ra[RD];
                unsigned dummy;
                if (!PREFETCH) fetch_row(c, wrow);
                decode_row(wrow, ra, dummy);
                if (PREFETCH && c + NT < p.m32) fetch_row(c + NT, wrow);
                real va[RD], vb[RD];
                unsigned long long fa = 0ULL, fbb = 0ULL;
                auto alpha_of = [&](int s) -> real { return METHOD == QD_BP_MIN_SUM_K ? (real)ctl->alpha[s] : (real)0; };
#define QD_LD(n, buf, f) row_ld<real, METHOD, RD, (n) % S>(ra, fb, f9, buf, f, dummy)
#define QD_CS(n, buf, f) row_cs<real, METHOD, RD, S, (n) % S, LONGROWS>(ra, fb, f9, buf, f, i, valid, synbits, resbits, un, alpha_of((n) % S))
                if (act == SMASK && !LONGROWS) {                // steady state: straight line, two buffers (long rows: one
                                                                // buffer, two would not fit the register file)
                    QD_LD(0, va, fa);
                    if (S > 1) QD_LD(1, vb, fbb);
                    QD_CS(0, va, fa);
                    if (S > 2) QD_LD(2, va, fa);
                    if (S > 1) QD_CS(1, vb, fbb);
                    if (S > 3) QD_LD(3, vb, fbb);
                    if (S > 2) QD_CS(2, va, fa);
                    if (S > 4) QD_LD(4, va, fa);
                    if (S > 3) QD_CS(3, vb, fbb);
                    if (S > 5) QD_LD(5, vb, fbb);
                    if (S > 4) QD_CS(4, va, fa);
                    if (S > 6) QD_LD(6, va, fa);
                    if (S > 5) QD_CS(5, vb, fbb);
                    if (S > 7) QD_LD(7, vb, fbb);
                    if (S > 6) QD_CS(6, va, fa);
                    if (S > 7) QD_CS(7, vb, fbb);
                } else {
                    if (S > 0 && (act & 1u)) { QD_LD(0, va, fa); QD_CS(0, va, fa); }
                    if (S > 1 && (act & 2u)) { QD_LD(1, va, fa); QD_CS(1, va, fa); }
                    if (S > 2 && (act & 4u)) { QD_LD(2, va, fa); QD_CS(2, va, fa); }
                    if (S > 3 && (act & 8u)) { QD_LD(3, va, fa); QD_CS(3, va, fa); }
                    if (S > 4 && (act & 16u)) { QD_LD(4, va, fa); QD_CS(4, va, fa); }
                    if (S > 5 && (act & 32u)) { QD_LD(5, va, fa); QD_CS(5, va, fa); }
                    if (S > 6 && (act & 64u)) { QD_LD(6, va, fa); QD_CS(6, va, fa); }
                    if (S > 7 && (act & 128u)) { QD_LD(7, va, fa); QD_CS(7, va, fa); }
                }
#undef QD_LD
#undef QD_CS
            }
            un = __reduce_or_sync(0xffffffffu, un);
            if (!CL) {
                if (lane == 0 && un) atomicOr(&ctl->unsat[ub], un);
            } else {
                // Cluster mode: everything that touches another CTA waits for the end-of-iteration barrier of the previous
                // iteration, which this CTA only ARRIVED at: the check pass above ran in its shadow.
                __syncthreads();
                if (b2_pending) { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); b2_pending = false; }
                if (lane == 0 && un)
                    for (uint32_t r = 0; r < NC; ++r) cl_red_or_u32(cl_map(smem_u32(&ctl->unsat[ub]), r), un);
                if (resv_have)
                    for (uint32_t r = 0; r < NC; ++r) cl_st_s64(cl_map(smem_u32(&ctl->resv[lane]), r), resv_new);
                drain_corrections(d1);
            }
        }
        if (!CL) __syncthreads();
        QD_TICK7(0);
        if (CL) {
            // ============================================================== export: check->bit messages of cut columns
            // into the mailboxes of the other CTAs that hold a replica of the column (consecutive entries fill
            // consecutive elements of a destination line)
            const int nexp = p.cl_nexp[rank];
            const uint2 *ex = p.cl_export + (size_t)rank * p.cl_exp_stride;
            for (int x = tid; x < nexp; x += NT) {
                const uint2 e = __ldg(ex + x);
                const uint32_t src = a_msg + (e.x & 0xffffu) * RB;
                const uint32_t dst = cl_map(a_msg + (e.x >> 16) * RB, e.y);
#pragma unroll
                for (int s = 0; s < S; ++s)
                    if ((act >> s) & 1u) cl_st(dst + s * LINE, lds_real(src + s * LINE, real()));
            }
            cluster_sync_release_all();
        }
        QD_TICK7(1);

        // ================================================================== retire / refill: nothing here touches global memory
        unsigned done = 0u;
        {
            const unsigned unsat = ctl->unsat[ub];
            int born[S];
            read_born(born);
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const int b = born[s], it = T - b - 1;
                const unsigned fin = (unsigned)(b != BP_IDLE) & (unsigned)(it >= 1) &
                                     ((~(unsat >> s) & 1u) | (unsigned)(it >= max_iter));
                done |= fin << s;
            }
            if (done) {
                // every CTA of a cluster holds the same control block and advances it in the same way
                if (tid < S && ((done >> tid) & 1u)) {
                    const int s = tid;
                    ctl->drain_shot[s] = ctl->shot[s];
                    ctl->drain_it[s] = T - ctl->born[s] - 1;
                    ctl->drain_conv[s] = (int)(~(unsat >> s) & 1u);
                    const long long nx = ctl->resv[s];
                    ctl->resv[s] = -1;
                    ctl->shot[s] = nx;
                    ctl->born[s] = nx >= 0 ? T : BP_IDLE;
                    if (nx >= 0) atomicOr(&ctl->need_resv, 1u << s);      // (an empty queue stays empty)
                }
                __syncthreads();
                act = active_mask();
                load_slots(T, false);
                // no barrier: the bit pass below skips slots born in T and ends with a barrier
            }
        }

        QD_TICK7(2);
        // ================================================================== bit pass
        {
            // ---- drain: logical parities of the residuals formed in this iteration's check pass (8 lanes per logical)
            if (lead && p.err) {
                if (d1) {
                    const int total = S * p.k * 8;
                    for (int base = warp * 32; base < total; base += NT) {
                        const int item = base + lane, g = item >> 3, j = item & 7;
                        const int s = g / p.k, i = g - s * p.k;
                        uint32_t acc = 0u;
                        if (item < total && ((d1 >> s) & 1u)) {
                            const uint32_t *D = outbits + s * p.nwn;
                            for (int w = j; w < p.nwn; w += 8) acc ^= __ldg(p.lzT + (size_t)w * p.k + i) & D[w];
                        }
                        acc ^= __shfl_xor_sync(0xffffffffu, acc, 1);
                        acc ^= __shfl_xor_sync(0xffffffffu, acc, 2);
                        acc ^= __shfl_xor_sync(0xffffffffu, acc, 4);
                        if (j == 0 && (__popc(acc) & 1)) atomicOr(&ctl->pad[0], 1 << s);
                    }
                }
            }
            // ---- queue service: request the syndromes of the shots drawn in this iteration's check pass
            if (warp == svc_warp) {
                const unsigned fr = ctl->fresh_resv;
                if (fr) {
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        const long long shot = ((fr >> s) & 1u) ? ctl->resv[s] : -1;
                        if (shot < 0) continue;
                        const uint32_t *src = p.synd + shot * (long long)p.synd_stride;
                        for (int w = lane; w < p.nwm; w += 32)
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(synbuf + s * p.nwm + w)), "l"(src + w) : "memory");
                    }
                    __syncwarp();
                    if (lane == 0) ctl->fresh_resv = 0u;
                }
            }
            // ---- drain: decisions of the shots retired in this iteration, back in column order
            if (done) {
                int i = 0;
                for (int r = tid; r < p.nrec; r += NT, ++i) {
                    const unsigned bits = (decbits >> (i * S)) & done;
                    if (bits) {
                        const unsigned col = colid[r];
                        if (col != 0xFFFFu) {
#pragma unroll
                            for (int s = 0; s < S; ++s)
                                if ((bits >> s) & 1u) {
                                    if (!CL) atomicOr(&outbits[s * p.nwn + (col >> 5)], 1u << (col & 31));
                                    else cl_red_or_u32(cl_map(smem_u32(&outbits[s * p.nwn + (col >> 5)]), 0u), 1u << (col & 31));
                                }
                        }
                        decbits &= ~(done << (i * S));
                    }
                }
            }
            // slots to update: the ones holding a shot that was not loaded in this iteration; slots whose soft output is
            // written in this pass (mode 1: only the iteration that hits the cap)
            const unsigned run = act & ~done;
            unsigned want = 0u;
            if (p.llr) {
                int born[S];
                read_born(born);
#pragma unroll
                for (int s = 0; s < S; ++s)
                    want |= (((run >> s) & 1u) & ((unsigned)(p.llr_mode == 0) | (unsigned)(T - born[s] == max_iter))) << s;
            }
            int i = 0;
            uint32_t wrec[VMAX];
            int degn = 0;
            if (PREFETCH && tid < p.nrec) { fetch_record(tid, wrec); degn = (int)__ldg(cdeg + (tid >> 5)); }
            for (int r = tid; r < p.nrec; r += NT, ++i) {
                uint32_t ea[VMAX], fa[VMAX];
                if (!PREFETCH) { fetch_record(r, wrec); degn = (int)__ldg(cdeg + (r >> 5)); }
                decode_record(wrec, ea, fa);
                const real pj = prior_of(r);
                const int deg = __shfl_sync(0xffffffffu, degn, 0);
                if (PREFETCH && r + NT < p.nrec) { fetch_record(r + NT, wrec); degn = (int)__ldg(cdeg + ((r + NT) >> 5)); }
                const Bp7LlrOut lo{want, &p, ctl->shot, rank};
#define QD_COL(D_) col_update<real, METHOD, (D_), VMAX, S, FB>(ea, fa, pj, i, run, decbits, lo, r)
                switch (deg) {
                case 1: QD_COL(1); break;
                case 2: QD_COL(2); break;
                case 3: QD_COL(3); break;
                case 4: QD_COL(4); break;
                case 5: if (VMAX > 4) QD_COL(VMAX > 4 ? 5 : 1); break;
                case 6: if (VMAX > 4) QD_COL(VMAX > 4 ? 6 : 1); break;
                case 7: if (VMAX > 4) QD_COL(VMAX > 4 ? 7 : 1); break;
                case 8: if (VMAX > 4) QD_COL(VMAX > 4 ? 8 : 1); break;
                default: break;
                }
#undef QD_COL
            }
            // ---- drain: ask for the error words of the shots retired in this iteration (used at the top of the next one)
            if (lead && done && p.err && tid < S * p.corr_stride) {
                const int s = tid / p.corr_stride, w = tid - s * p.corr_stride;
                if (((done >> s) & 1u) && w < p.nwn) err_pf = p.err[ctl->drain_shot[s] * (long long)p.err_stride + w];
            }
        }
        QD_TICK7(3);
        if (tid == 0) ctl->unsat[ub ^ 1] = 0u;              // buffer of the next check pass
        publish_alpha(T + 1);
        // (cluster mode: nobody exports into a mailbox that is still being read, nobody sets a flag that is being reset —
        // but this CTA only arrives here and waits after its next check pass, which touches no other CTA)
        __syncthreads();
        if (CL) {
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");   // every thread releases what it sent to CTA 0
            b2_pending = true;
#ifdef QD_BP7_NO_SPLIT_BARRIER                                  // (experiments: wait at once)
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
            b2_pending = false;
#endif
        }
        QD_TICK7(4);
#ifdef QD_BP7_TIMING
        ++n_iter;
#endif
        // go on while a slot decodes or drains
        d2 = d1; d1 = done;
        if (!(act | d1 | d2)) break;
    }
#ifdef QD_BP7_TIMING
    if (blockIdx.x < 8 && (tid == 0 || tid == NT - 1))
        printf("bp7 timing cta %d tid %d: check %lld export+bar %lld retire %lld bit %lld endbar %lld (iters %lld)\n", (int)blockIdx.x, tid,
               tacc[0], tacc[1], tacc[2], tacc[3], tacc[4], n_iter);
#endif
    if (CL) {                                               // no CTA leaves while its shared memory may still be addressed
        if (b2_pending) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        cluster_sync_release_all();
    }

    if (tid == 0 && p.counters && c_shots) {
        atomicAdd((unsigned long long *)&p.counters[0], (unsigned long long)c_fail);
        atomicAdd((unsigned long long *)&p.counters[1], (unsigned long long)c_shots);
        atomicAdd((unsigned long long *)&p.counters[2], (unsigned long long)c_nc);
        atomicAdd((unsigned long long *)&p.counters[3], (unsigned long long)c_it);
    }
}

}  // namespace qd
