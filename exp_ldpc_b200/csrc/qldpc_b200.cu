// qldpc_b200.cu — C ABI (include/qldpc_b200.h) + bit-packed utility kernels + host orchestration.
// sm_100a only.  No CPU fallback: without a usable device every entry point returns QD_ENODEV.

#include "../../include/qldpc_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <memory>
#include <string>
#include <vector>

#include "bp7_kernel.cuh"
#include "bp7_layout.h"
#include "bp7_cluster_layout.h"
#include "bp_pick.h"
#include "bpg_kernel.cuh"
#include "bpc_kernel.cuh"
#include "frame_sim.cuh"
#include "osd_kernel.cuh"
#include "ssf_kernel.cuh"
#include "syndrome_kernel.cuh"

namespace qd {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
int fail(int code, const std::string &msg) { g_err = msg; return code; }

// =============================================================================================
// small utility kernels (HBM-bound, bit-packed, shot-major)
// =============================================================================================

// one thread per output word: 32 Bernoulli bits = 8 Philox blocks
__global__ void sample_kernel(const uint32_t *__restrict__ thr, int nbits, int nw, uint32_t k0, uint32_t k1,
                              uint32_t stream_id, long long shot0, long long shots, uint32_t *__restrict__ out,
                              int stride)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= shots * stride) return;
    const long long s = idx / stride;
    const int w = (int)(idx - s * stride);
    uint32_t word = 0u;
    if (w < nw) {
        const unsigned long long g = (unsigned long long)(shot0 + s);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int j0 = w * 32 + b * 4;
            if (j0 < nbits) {
                uint32_t r[4];
                philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)(j0 >> 2), stream_id, k0, k1, r);
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (j0 + t < nbits && r[t] < __ldg(thr + j0 + t)) word |= 1u << (b * 4 + t);
            }
        }
    }
    out[idx] = word;
}

// one thread per syndrome word (32 checks) of one shot
__global__ void syndrome_kernel(GraphDev g, const uint32_t *__restrict__ vec, int vstride, long long shots,
                                uint32_t *__restrict__ synd, int sstride)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= shots * sstride) return;
    const long long s = idx / sstride;
    const int w = (int)(idx - s * sstride);
    uint32_t word = 0u;
    if (w < g.nwm) {
        const uint32_t *v = vec + s * (long long)vstride;
        const int c1 = min(g.m, w * 32 + 32);
        for (int c = w * 32; c < c1; ++c) {
            uint32_t par = 0u;
            for (int e = __ldg(g.row_ptr + c); e < __ldg(g.row_ptr + c + 1); ++e) {
                const int col = __ldg(g.col_idx + e);
                par ^= v[col >> 5] >> (col & 31);
            }
            word |= (par & 1u) << (c & 31);
        }
    }
    synd[idx] = word;
}

// 32 x 32 bit transpose in registers (bit b of a[s]  <->  bit s of a[b])
__device__ __forceinline__ void transpose32(uint32_t (&a)[32])
{
    uint32_t msk = 0x0000FFFFu;
#pragma unroll
    for (int j = 16; j != 0; j >>= 1, msk ^= msk << j) {
#pragma unroll
        for (int k = 0; k < 32; k = (k + j + 1) & ~j) {
            const uint32_t t = ((a[k] >> j) ^ a[k + j]) & msk;
            a[k] ^= t << j;
            a[k + j] ^= t;
        }
    }
}

// Bit-sliced H e mod 2: a warp takes a tile of 32 shots.  The packed error rows are read coalesced (lane = word),
// every lane transposes its 32 x 32 bit block in registers, so that shared memory holds one word per COLUMN with
// the 32 shots in its bits (pitch 33: conflict free).  A check is then the XOR of the words of its columns — 32
// shots per operation instead of one bit gather per shot and edge — computed by lane k for the checks 32k..32k+31,
// transposed back in registers and written as word k of the 32 syndrome rows (coalesced).  HBM traffic is the
// algorithmic minimum: 4 * (ceil(n/32) + ceil(m/32)) bytes per shot.
// DEG: compile-time bound of the row degree (8 or 16: the edge loop unrolls and its table loads are issued ahead of
// the shared-memory gathers), 0 = run-time
template <int DEG>
__global__ void __launch_bounds__(256) syndrome_sliced_kernel(GraphDev g, const uint32_t *__restrict__ vec, int vstride, long long shots,
                                                              uint32_t *__restrict__ synd, int sstride)
{
    extern __shared__ uint32_t sliced_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint32_t *cw = sliced_smem + (size_t)warp * g.nwn * 33;       // column c at (c >> 5) * 33 + (c & 31)
    const long long ntiles = (shots + 31) / 32;
    for (long long tile = (long long)blockIdx.x * nwarps + warp; tile < ntiles; tile += (long long)gridDim.x * nwarps) {
        const long long s0 = tile * 32;
        const int ns = (int)min(32LL, shots - s0);
        for (int wc = 0; wc < g.nwn; wc += 32) {
            const int w = wc + lane;
            uint32_t a[32];
#pragma unroll
            for (int s = 0; s < 32; ++s) a[s] = (s < ns && w < g.nwn) ? vec[(s0 + s) * (long long)vstride + w] : 0u;
            transpose32(a);
            if (w < g.nwn) {
#pragma unroll
                for (int b = 0; b < 32; ++b) cw[w * 33 + b] = a[b];
            }
        }
        __syncwarp();
        for (int kc = 0; kc < sstride; kc += 32) {
            const int k = kc + lane;
            uint32_t o[32];
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                uint32_t x = 0u;
                if (k < g.nwm) {
                    const int32_t *t = g.ell + (size_t)b * g.ell_deg * g.nwm + k;      // lanes read consecutive entries
                    if (DEG > 0) {
                        int c[DEG > 0 ? DEG : 1];
#pragma unroll
                        for (int e = 0; e < DEG; ++e) c[e] = e < g.ell_deg ? __ldg(t + (size_t)e * g.nwm) : -1;
#pragma unroll
                        for (int e = 0; e < DEG; ++e) if (c[e] >= 0) x ^= cw[(c[e] >> 5) * 33 + (c[e] & 31)];
                    } else {
                        for (int e = 0; e < g.ell_deg; ++e) {
                            const int c = __ldg(t + (size_t)e * g.nwm);
                            if (c >= 0) x ^= cw[(c >> 5) * 33 + (c & 31)];
                        }
                    }
                }
                o[b] = x;
            }
            transpose32(o);
            if (k < sstride) {
#pragma unroll
                for (int s = 0; s < 32; ++s)
                    if (s < ns) synd[(s0 + s) * (long long)sstride + k] = o[s];
            }
        }
        __syncwarp();
    }
}

// one warp per shot: fail = any_i parity(L_i & (a ^ b))
__global__ void logical_fail_kernel(const uint32_t *__restrict__ lzT, int k, int nwn, const uint32_t *__restrict__ a,
                                    const uint32_t *__restrict__ b, int stride, long long shots,
                                    uint8_t *__restrict__ fail_out, unsigned long long *fail_count)
{
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= shots) return;
    const uint32_t *ra = a + warp * (long long)stride;
    const uint32_t *rb = b ? b + warp * (long long)stride : nullptr;
    int f = 0;
    for (int i = lane; i < ((k + 31) & ~31); i += 32) {
        uint32_t acc = 0u;
        if (i < k)
            for (int w = 0; w < nwn; ++w) {
                const uint32_t r = ra[w] ^ (rb ? rb[w] : 0u);
                acc ^= __ldg(lzT + (size_t)w * k + i) & r;
            }
        f |= __popc(acc) & 1;
    }
    f = __any_sync(0xffffffffu, f);
    if (lane == 0) {
        if (fail_out) fail_out[warp] = (uint8_t)f;
        if (fail_count && f) atomicAdd(fail_count, 1ULL);
    }
}

// dst[s] bits [dst_offset, dst_offset + nbits)  (^)=  src[s] bits [src_offset, src_offset + nbits);
// one thread per destination word, other destination bits are left alone.
// The building block of the multi-round recipes: correction fold (spacetime_code.py:81-84),
// syndrome history and its differencing (spacetime_code.py:98-119), data part of a single-shot
// decoding (spacetime_code.py:31-33), slices of a measurement record (storage_sim.py:187-196).
__global__ void bits_copy_kernel(const uint32_t *__restrict__ src, int sstride, long long src_offset, long long shots,
                                 int nbits, uint32_t *__restrict__ dst, int dstride, long long dst_offset, int accumulate)
{
    const int w0 = (int)(dst_offset >> 5), w1 = (int)((dst_offset + nbits - 1) >> 5), nw = w1 - w0 + 1;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= shots * nw) return;
    const long long s = idx / nw;
    const int w = w0 + (int)(idx - s * nw);
    const long long b0 = max((long long)w * 32, dst_offset), b1 = min((long long)w * 32 + 32, dst_offset + nbits);
    const int len = (int)(b1 - b0), sh = (int)(b0 - (long long)w * 32);
    const long long s0 = src_offset + (b0 - dst_offset);
    const uint32_t *row = src + s * (long long)sstride;
    const int sw = (int)(s0 >> 5), ss = (int)(s0 & 31);
    const uint32_t lo = row[sw];
    const uint32_t hi = (ss && ss + len > 32) ? row[sw + 1] : 0u;
    uint32_t v = ss ? __funnelshift_r(lo, hi, ss) : lo;
    const uint32_t lm = len >= 32 ? 0xffffffffu : ((1u << len) - 1u);
    v = (v & lm) << sh;
    uint32_t *d = dst + s * (long long)dstride + w;
    *d = accumulate ? (*d ^ v) : ((*d & ~(lm << sh)) | v);
}

// =============================================================================================
// host-side handles
// =============================================================================================
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t n)
    {
        if (p) { cudaFree(p); p = nullptr; }
        bytes = n;
        if (n == 0) return 0;
        cudaError_t e = cudaMalloc(&p, n);
        if (e != cudaSuccess) { p = nullptr; return fail(QD_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
        return 0;
    }
    int ensure(size_t n) { return n <= bytes && p ? 0 : alloc(n); }
    template <typename T> T *as() { return static_cast<T *>(p); }
};

struct PinBuf {
    void *p = nullptr;
    size_t bytes = 0;
    ~PinBuf() { if (p) cudaFreeHost(p); }
    int ensure(size_t n)
    {
        if (n <= bytes && p) return 0;
        if (p) { cudaFreeHost(p); p = nullptr; }
        bytes = n;
        cudaError_t e = cudaMallocHost(&p, n ? n : 1);
        if (e != cudaSuccess) { p = nullptr; return fail(QD_ENOMEM, std::string("cudaMallocHost: ") + cudaGetErrorString(e)); }
        return 0;
    }
    template <typename T> T *as() { return static_cast<T *>(p); }
};

template <typename T>
static int upload(DevBuf &b, const std::vector<T> &v)
{
    int rc = b.alloc(std::max<size_t>(v.size(), 1) * sizeof(T));
    if (rc) return rc;
    if (!v.empty()) QD_CUDA(cudaMemcpy(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

static int use_device(int device)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(QD_ENODEV, std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0"));
    if (device < 0 || device >= count) return fail(QD_EINVAL, "device index out of range");
    QD_CUDA(cudaSetDevice(device));
    return 0;
}

}  // namespace qd

using namespace qd;

struct qd_graph {
    int device = 0;
    GraphDev dev{};
    std::vector<int32_t> row_ptr, col_idx, col_ptr;
    std::vector<uint32_t> row_desc, col_desc, csc_pos, csc_row, fpos_of_pos;
    std::vector<uint8_t> csc_k;      // index of the edge within its row
    std::vector<int32_t> csc_edge;   // CSR edge id of every CSC entry
    DevBuf d_col_ptr, d_csc_edge, d_ell, d_syn_tbl;
    int syn_q = 0, syn_ctas[2] = {0, 0};     // 128-shot syndrome kernel: uint4 table entries per check; resident CTAs per SM (0: not probed)
    int big = 0;                     // degrees / size beyond the shared-memory kernels: global-memory BP only
    // message placements of the thread-owned-state BP kernel, keyed by (banks, column tile)
    std::list<std::pair<int, Bp7Layout>> bp7_layouts;   // list: references stay valid
    DevBuf d_row_desc, d_col_desc, d_csc_pos, d_csc_row, d_row_ptr, d_col_idx, d_fpos;
    int sm_count = 0, smem_optin = 0;
};

struct qd_logicals {
    int device = 0;
    int k = 0, n = 0, nwn = 0;
    DevBuf d_lzT;   // [nwn][k]
};

struct qd_sampler {
    int device = 0;
    int nbits = 0;
    DevBuf d_thr;
};

struct qd_bp {
    qd_graph *g = nullptr;
    int device = 0;        // copy: the graph handle may be destroyed before the decoder
    int max_iter = 0, method = 0, precision = 64;
    int vmax = 4;          // register tile of a column
    int rd = 0;            // compile-time row degree of the selected kernel (0 = generic rows)
    int idx_bytes = 2;     // bytes per message position in vidx
    int stage_idx = 0;     // vidx + priors staged in shared memory
    int threads = 0, slots = 0, ctas_per_sm = 0, groups = 1;
    size_t smem = 0;
    DevBuf d_prior, d_alpha, d_next, d_vidx, d_chunk_deg;
    // row-regular kernel (bp7_kernel.cuh): record table for the configured slot count
    int v7 = 0, v7_nrec = 0, v7_uniform = 0, v7_staged = 0, v7_irreg = 0;
    double v7_prior0 = 0.0, alpha_fixed = 0.0;
    std::vector<double> prior_dom;   // priors in the decoder's domain, per column
    int v7_lines = 0;
    int global_mem = 0;    // messages in global memory (bpg_kernel.cuh): graphs the shared-memory kernels cannot take
    DevBuf d_scratch[2];        // global-memory kernel: message arrays of the CTAs in flight, one buffer per work queue
                                // (the host pipeline keeps two launches in flight on two streams)
    // cluster kernel (bpc_kernel.cuh): messages spread over the shared memories of `cl_C` CTAs
    int cluster = 0, cl_C = 0, cl_max_edges = 0, cl_max_rows = 0, cl_max_cols = 0, cl_max_cedges = 0, cl_nwpos = 0, cl_clusters = 0;
    DevBuf dc_row_lo, dc_row_start, dc_edge_pos, dc_lc_lo, dc_lc_col, dc_lc_ptr, dc_desc, dc_word_lo, dc_col_pos, dc_prior;
    // cluster mode of the row-regular kernel (bp7_kernel.cuh, CL): rows split over `cl7_C` CTAs, replicas + mailboxes for cut columns
    int cl7 = 0, cl7_C = 0, cl7_clusters = 0;
    Bp7ClusterPlan cl7_plan;
    DevBuf d7c_row0, d7c_rows, d7c_nexp, d7c_export;
    int llr_mode = 0;      // what d_llr receives (0: log-ratio every iteration, 1: raw soft output of the last iteration)
    DevBuf d7_table, d7_row_table, d7_chunk_deg, d7_prior_rec, d7_col_id, d7_const_bits;
    // host-buffer pipeline state
    cudaStream_t streams[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    DevBuf d_in[2], d_out[2], d_conv[2], d_it[2];
    PinBuf h_in[2], h_out[2], h_conv[2], h_it[2];
    ~qd_bp()
    {
        for (int i = 0; i < 2; ++i) {
            if (streams[i]) cudaStreamDestroy(streams[i]);
            if (done[i]) cudaEventDestroy(done[i]);
        }
    }
};

struct qd_circuit {
    int device = 0;
    int nops = 0, nq = 0, nmeas = 0, nthr = 0;
    DevBuf d_ops, d_thr;
};

struct qd_osd {
    qd_graph *g = nullptr;
    int device = 0;
    int rank = 0, W = 0, P = 0;
    int method = 0, order = 0;    // 0 OSD-0, 1 osd_e, 2 osd_cs
    size_t smem = 0;
    int rows_global = 0;          // [H | s] in a global scratch buffer (does not fit shared memory)
    DevBuf d_next, d_soft, d_conv, d_rows, d_failbuf;
};

struct qd_ssf {
    qd_graph *g = nullptr;
    int device = 0;
    SsfDev dev{};
    int max_rounds = 0;
    int max_window = 64;   // checks in the largest generator window (<= 32: the 32-bit scan)
    int threads = 128;
    size_t smem = 0;
    DevBuf d_gen_ptr, d_gen_idx, d_win_ptr, d_win_chk, d_qmask, d_chk_ptr, d_chk_gen, d_next;
    cudaStream_t streams[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    DevBuf d_in[2], d_out[2], d_conv[2], d_it[2];
    PinBuf h_in[2], h_out[2], h_conv[2], h_it[2];
    ~qd_ssf()
    {
        for (int i = 0; i < 2; ++i) {
            if (streams[i]) cudaStreamDestroy(streams[i]);
            if (done[i]) cudaEventDestroy(done[i]);
        }
    }
};

// =============================================================================================
// misc
// =============================================================================================
extern "C" int qd_abi_version(void) { return QD_ABI_VERSION; }
extern "C" const char *qd_last_error(void) { return g_err.c_str(); }

extern "C" int qd_device_count(void)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) return 0;
    return count;
}

extern "C" int qd_device_info(int device, int *sm_count, int *smem_per_block_optin, int *clock_khz)
{
    int rc = use_device(device);
    if (rc) return rc;
    int v = 0;
    if (sm_count) { QD_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device)); *sm_count = v; }
    if (smem_per_block_optin) {
        QD_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        *smem_per_block_optin = v;
    }
    if (clock_khz) { QD_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, device)); *clock_khz = v; }
    return 0;
}

// =============================================================================================
// graph
// =============================================================================================
extern "C" int qd_graph_create(int32_t m, int32_t n, int64_t nnz, const int32_t *h_row_ptr,
                               const int32_t *h_col_idx, int device, qd_graph **out)
{
    if (!out) return fail(QD_EINVAL, "out is NULL");
    *out = nullptr;
    if (m < 0 || n <= 0 || nnz < 0 || !h_row_ptr || (nnz > 0 && !h_col_idx)) return fail(QD_EINVAL, "bad graph arguments");
    if (h_row_ptr[0] != 0 || h_row_ptr[m] != nnz) return fail(QD_EINVAL, "row_ptr does not span [0, nnz]");
    if (nnz >= (1 << 24)) return fail(QD_EUNSUP, "more than 2^24 edges");
    int rc = use_device(device);
    if (rc) return rc;
    std::unique_ptr<qd_graph> g(new qd_graph);
    g->device = device;
    g->row_ptr.assign(h_row_ptr, h_row_ptr + m + 1);
    g->col_idx.assign(h_col_idx, h_col_idx + nnz);
    int max_row = 0;
    for (int i = 0; i < m; ++i) {
        const int lo = h_row_ptr[i], hi = h_row_ptr[i + 1];
        if (hi < lo) return fail(QD_EINVAL, "row_ptr not monotone");
        max_row = std::max(max_row, hi - lo);
        for (int e = lo; e < hi; ++e) {
            if (h_col_idx[e] < 0 || h_col_idx[e] >= n) return fail(QD_EINVAL, "column index out of range");
            if (e > lo && h_col_idx[e] <= h_col_idx[e - 1]) return fail(QD_EINVAL, "column indices must be sorted and unique");
        }
    }
    if (max_row > 255) g->big = 1;
    int min_row = max_row;
    for (int i = 0; i < m; ++i) min_row = std::min(min_row, h_row_ptr[i + 1] - h_row_ptr[i]);
    const int regular_deg = (m > 0 && min_row == max_row) ? max_row : 0;
    // message layout: rows padded to an odd pitch when that wastes < 1/3 of the array
    int pitch = max_row | 1;
    const bool padded = m > 0 && (regular_deg > 0 || (int64_t)pitch * m <= nnz + nnz / 3 + 64);
    const int fb = std::max(8, (max_row + 7) / 8 * 8);     // decision-flip flag bytes per row
    g->row_desc.resize(std::max(m, 1));
    std::vector<int> pos_of_edge(std::max<int64_t>(nnz, 1));
    const int npos = padded ? pitch * m : (int)nnz;
    if (npos >= (1 << 24)) return fail(QD_EUNSUP, "message array too large");
    g->fpos_of_pos.assign(std::max(npos, 1), 0u);
    for (int i = 0; i < m; ++i) {
        const int lo = h_row_ptr[i], deg = h_row_ptr[i + 1] - lo;
        const int pos0 = padded ? i * pitch : lo;
        g->row_desc[i] = (uint32_t)pos0 | ((uint32_t)deg << 24);
        for (int k = 0; k < deg; ++k) {
            pos_of_edge[lo + k] = pos0 + k;
            g->fpos_of_pos[pos0 + k] = (uint32_t)(i * fb + k);
        }
    }
    // CSC: column lists with rows ascending
    g->col_ptr.assign(n + 1, 0);
    for (int64_t e = 0; e < nnz; ++e) g->col_ptr[h_col_idx[e] + 1]++;
    int max_col = 0;
    for (int j = 0; j < n; ++j) {
        max_col = std::max(max_col, g->col_ptr[j + 1]);
        g->col_ptr[j + 1] += g->col_ptr[j];
    }
    if (max_col > 255) g->big = 1;
    g->csc_pos.resize(std::max<int64_t>(nnz, 1));
    g->csc_row.resize(std::max<int64_t>(nnz, 1));
    g->csc_k.resize(std::max<int64_t>(nnz, 1));
    g->csc_edge.resize(std::max<int64_t>(nnz, 1));
    g->col_desc.resize(n);
    {
        std::vector<int> fill(g->col_ptr.begin(), g->col_ptr.end() - 1);
        for (int i = 0; i < m; ++i)
            for (int e = h_row_ptr[i]; e < h_row_ptr[i + 1]; ++e) {
                const int t = fill[h_col_idx[e]]++;
                g->csc_pos[t] = (uint32_t)pos_of_edge[e];
                g->csc_row[t] = (uint32_t)i;
                g->csc_k[t] = (uint8_t)std::min(e - h_row_ptr[i], 255);
                g->csc_edge[t] = e;
            }
        for (int j = 0; j < n; ++j)
            g->col_desc[j] = (uint32_t)g->col_ptr[j] | ((uint32_t)(g->col_ptr[j + 1] - g->col_ptr[j]) << 24);
    }
    if ((rc = upload(g->d_row_desc, g->row_desc))) return rc;
    if ((rc = upload(g->d_col_desc, g->col_desc))) return rc;
    if ((rc = upload(g->d_csc_pos, g->csc_pos))) return rc;
    if ((rc = upload(g->d_csc_row, g->csc_row))) return rc;
    if ((rc = upload(g->d_row_ptr, g->row_ptr))) return rc;
    if ((rc = upload(g->d_col_idx, g->col_idx))) return rc;
    if ((rc = upload(g->d_fpos, g->fpos_of_pos))) return rc;
    if ((rc = upload(g->d_col_ptr, g->col_ptr))) return rc;
    if ((rc = upload(g->d_csc_edge, g->csc_edge))) return rc;
    GraphDev &d = g->dev;
    d.m = m; d.n = n; d.nnz = (int)nnz;
    d.nwm = words_for(m); d.nwn = words_for(n);
    d.npos = npos;
    d.pitch = padded ? pitch : 0;
    d.regular_deg = regular_deg;
    d.fb = fb;
    d.fpos_of_pos = g->d_fpos.as<uint32_t>();
    d.max_row_deg = max_row; d.max_col_deg = max_col;
    d.row_desc = g->d_row_desc.as<uint32_t>();
    d.col_desc = g->d_col_desc.as<uint32_t>();
    d.csc_pos = g->d_csc_pos.as<uint32_t>();
    d.csc_row = g->d_csc_row.as<uint32_t>();
    d.row_ptr = g->d_row_ptr.as<int32_t>();
    d.col_idx = g->d_col_idx.as<int32_t>();
    d.ell = nullptr; d.ell_deg = 0;
    if (max_row >= 1 && max_row <= 64) {
        // the syndrome kernel's view of H: for fixed (bit b, edge e) the 32-check groups k are contiguous
        const int nwm = d.nwm;
        std::vector<int32_t> ell((size_t)32 * max_row * std::max(nwm, 1), -1);
        for (int r = 0; r < m; ++r)
            for (int e = h_row_ptr[r]; e < h_row_ptr[r + 1]; ++e)
                ell[((size_t)(r & 31) * max_row + (e - h_row_ptr[r])) * nwm + (r >> 5)] = h_col_idx[e];
        if ((rc = upload(g->d_ell, ell))) return rc;
        d.ell = g->d_ell.as<int32_t>(); d.ell_deg = max_row;
    }
    if (max_row >= 1 && m >= 1) {
        // the 128-shot kernel's view of H (syndrome_kernel.cuh): byte offsets of the 16-byte column slots, 4 edges
        // per uint4, entry q of check c at [q * m + c]; missing edges point at the zero slot behind the columns
        const int Q = (max_row + 3) / 4;
        const uint32_t zero_off = (uint32_t)(33 * d.nwn) * 16u;
        std::vector<uint32_t> tbl((size_t)Q * m * 4, zero_off);
        for (int r = 0; r < m; ++r)
            for (int e = h_row_ptr[r]; e < h_row_ptr[r + 1]; ++e) {
                const int k = e - h_row_ptr[r], col = h_col_idx[e];
                tbl[((size_t)(k >> 2) * m + r) * 4 + (k & 3)] = (uint32_t)(col + (col >> 5)) * 16u;
            }
        if ((rc = upload(g->d_syn_tbl, tbl))) return rc;
        g->syn_q = Q;
    }
    QD_CUDA(cudaDeviceGetAttribute(&g->sm_count, cudaDevAttrMultiProcessorCount, device));
    QD_CUDA(cudaDeviceGetAttribute(&g->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    *out = g.release();
    return 0;
}

extern "C" int qd_graph_destroy(qd_graph *g)
{
    if (g) { cudaSetDevice(g->device); delete g; }
    return 0;
}

// =============================================================================================
// logicals, sampler
// =============================================================================================
extern "C" int qd_logicals_create(int32_t k, int32_t n, const uint8_t *h_dense, int device, qd_logicals **out)
{
    if (!out) return fail(QD_EINVAL, "out is NULL");
    *out = nullptr;
    if (k < 0 || n <= 0 || (k > 0 && !h_dense)) return fail(QD_EINVAL, "bad logicals arguments");
    int rc = use_device(device);
    if (rc) return rc;
    std::unique_ptr<qd_logicals> l(new qd_logicals);
    l->device = device; l->k = k; l->n = n; l->nwn = words_for(n);
    std::vector<uint32_t> t((size_t)l->nwn * std::max(k, 1), 0u);
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < n; ++j)
            if (h_dense[(size_t)i * n + j] & 1) t[(size_t)(j >> 5) * k + i] |= 1u << (j & 31);
    if ((rc = upload(l->d_lzT, t))) return rc;
    *out = l.release();
    return 0;
}

extern "C" int qd_logicals_destroy(qd_logicals *l)
{
    if (l) { cudaSetDevice(l->device); delete l; }
    return 0;
}

extern "C" int qd_sampler_create(int32_t nbits, const uint32_t *h_threshold, int device, qd_sampler **out)
{
    if (!out) return fail(QD_EINVAL, "out is NULL");
    *out = nullptr;
    if (nbits <= 0 || !h_threshold) return fail(QD_EINVAL, "bad sampler arguments");
    int rc = use_device(device);
    if (rc) return rc;
    std::unique_ptr<qd_sampler> s(new qd_sampler);
    s->device = device; s->nbits = nbits;
    std::vector<uint32_t> thr(h_threshold, h_threshold + nbits);
    if ((rc = upload(s->d_thr, thr))) return rc;
    *out = s.release();
    return 0;
}

extern "C" int qd_sampler_destroy(qd_sampler *s)
{
    if (s) { cudaSetDevice(s->device); delete s; }
    return 0;
}

static inline unsigned grid_for(long long items, int block) { return (unsigned)((items + block - 1) / block); }

extern "C" int qd_sampler_run(qd_sampler *s, uint64_t seed, uint32_t stream_id, int64_t shot0, int64_t shots,
                              uint32_t *d_bits, int32_t stride, void *stream)
{
    if (!s || !d_bits || shots < 0) return fail(QD_EINVAL, "bad sampler_run arguments");
    const int nw = words_for(s->nbits);
    if (stride < nw) return fail(QD_EINVAL, "stride smaller than ceil(nbits/32)");
    if (shots == 0) return 0;
    int rc = use_device(s->device);
    if (rc) return rc;
    const long long items = (long long)shots * stride;
    sample_kernel<<<grid_for(items, 256), 256, 0, (cudaStream_t)stream>>>(
        s->d_thr.as<uint32_t>(), s->nbits, nw, (uint32_t)seed, (uint32_t)(seed >> 32), stream_id, shot0, shots, d_bits, stride);
    QD_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int qd_bits_copy_batch(const uint32_t *d_src, int32_t src_stride, int64_t src_bit_offset, int64_t shots,
                                  int32_t nbits, uint32_t *d_dst, int32_t dst_stride, int64_t dst_bit_offset,
                                  int32_t accumulate, void *stream)
{
    if (!d_src || !d_dst || shots < 0 || src_bit_offset < 0 || dst_bit_offset < 0 || nbits < 0)
        return fail(QD_EINVAL, "bad bits_copy_batch arguments");
    if ((src_bit_offset + nbits + 31) / 32 > src_stride || (dst_bit_offset + nbits + 31) / 32 > dst_stride)
        return fail(QD_EINVAL, "stride too small");
    if (shots == 0 || nbits == 0) return 0;
    const int nw = (int)(((dst_bit_offset + nbits - 1) >> 5) - (dst_bit_offset >> 5) + 1);
    const long long items = (long long)shots * nw;
    bits_copy_kernel<<<grid_for(items, 256), 256, 0, (cudaStream_t)stream>>>(d_src, src_stride, src_bit_offset, shots, nbits,
                                                                             d_dst, dst_stride, dst_bit_offset, accumulate);
    QD_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int qd_syndrome_batch(qd_graph *g, const uint32_t *d_vec, int32_t vec_stride, int64_t shots,
                                 uint32_t *d_synd, int32_t synd_stride, void *stream)
{
    if (!g || !d_vec || !d_synd || shots < 0) return fail(QD_EINVAL, "bad syndrome_batch arguments");
    if (vec_stride < g->dev.nwn || synd_stride < g->dev.nwm) return fail(QD_EINVAL, "stride too small");
    if (shots == 0) return 0;
    int rc = use_device(g->device);
    if (rc) return rc;
    // 128-shot bit-sliced kernel whenever the column and check slots of a tile fit shared memory; the tile's packed
    // rows come in through one TMA bulk copy when they are 16-byte aligned and the staging buffer fits as well
    const int xs = 33 * g->dev.nwn + 1, ys = 33 * g->dev.nwm;
    const size_t smem_xy = ((size_t)xs + ys) * 16, smem_raw = (size_t)128 * vec_stride * 4;
    if (g->syn_q > 0 && smem_xy + 64 <= (size_t)g->smem_optin && !getenv("QD_SYNDROME_OLD")) {
        const bool tma = ((uintptr_t)d_vec & 15u) == 0 && (vec_stride & 3) == 0 && smem_xy + smem_raw + 64 <= (size_t)g->smem_optin &&
                         !getenv("QD_SYNDROME_NO_TMA");
        const size_t smem128 = smem_xy + (tma ? smem_raw : 0);
        const int qt = g->syn_q <= 3 ? g->syn_q : 0;
        void (*kern)(const SynParams) =
            tma ? (qt == 1 ? syndrome128_kernel<true, 1> : qt == 2 ? syndrome128_kernel<true, 2> : qt == 3 ? syndrome128_kernel<true, 3> : syndrome128_kernel<true, 0>)
                : (qt == 1 ? syndrome128_kernel<false, 1> : qt == 2 ? syndrome128_kernel<false, 2> : qt == 3 ? syndrome128_kernel<false, 3> : syndrome128_kernel<false, 0>);
        // one warp per 8-word strip of the two transposes (in + out) when that is at most 10 warps, else 8
        const int strips = (g->dev.nwn + 7) / 8 + (synd_stride + 7) / 8;
        const int warps = strips <= 10 ? strips : 8;
        QD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem128));
        int nb = 0;
        QD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, warps * 32, smem128));
        SynParams sp{};
        sp.n = g->dev.n; sp.m = g->dev.m; sp.nwn = g->dev.nwn; sp.nwm = g->dev.nwm;
        sp.xs = xs; sp.ys = ys; sp.Q = g->syn_q;
        sp.tbl = g->d_syn_tbl.as<uint4>();
        sp.vec = d_vec; sp.vstride = vec_stride; sp.shots = shots; sp.synd = d_synd; sp.sstride = synd_stride;
        const long long tiles = (shots + 127) / 128;
        const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(tiles, (long long)g->sm_count * std::max(1, nb)));
        kern<<<grid, warps * 32, smem128, (cudaStream_t)stream>>>(sp);
        QD_CUDA(cudaGetLastError());
        return 0;
    }
    const size_t per_warp = (size_t)g->dev.nwn * 33 * 4;
    if (per_warp <= 96u * 1024u && g->dev.ell && !getenv("QD_SYNDROME_GATHER")) {
        // bit-sliced kernel: as many warps per CTA as ~96 KB of shared memory hold (two CTAs per SM), at most 8
        const int warps = (int)std::max<size_t>(1, std::min<size_t>(8, (96u * 1024u) / per_warp));
        const size_t smem = per_warp * warps;
        auto kern = g->dev.ell_deg <= 8 ? syndrome_sliced_kernel<8> : g->dev.ell_deg <= 16 ? syndrome_sliced_kernel<16> : syndrome_sliced_kernel<0>;
        QD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const long long tiles = (shots + 31) / 32;
        const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((tiles + warps - 1) / warps, (long long)g->sm_count * 4));
        kern<<<grid, warps * 32, smem, (cudaStream_t)stream>>>(g->dev, d_vec, vec_stride, shots, d_synd, synd_stride);
    } else {
        const long long items = (long long)shots * synd_stride;
        syndrome_kernel<<<grid_for(items, 128), 128, 0, (cudaStream_t)stream>>>(g->dev, d_vec, vec_stride, shots, d_synd, synd_stride);
    }
    QD_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int qd_logical_fail_batch(qd_logicals *l, const uint32_t *d_a, const uint32_t *d_b, int32_t stride,
                                     int64_t shots, uint8_t *d_fail, unsigned long long *d_fail_count, void *stream)
{
    if (!l || !d_a || shots < 0) return fail(QD_EINVAL, "bad logical_fail_batch arguments");
    if (stride < l->nwn) return fail(QD_EINVAL, "stride too small");
    if (shots == 0) return 0;
    int rc = use_device(l->device);
    if (rc) return rc;
    logical_fail_kernel<<<grid_for((long long)shots * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        l->d_lzT.as<uint32_t>(), l->k, l->nwn, d_a, d_b, stride, shots, d_fail, d_fail_count);
    QD_CUDA(cudaGetLastError());
    return 0;
}

// =============================================================================================
// BP
// =============================================================================================
// ---- kernel pickers (instantiations live in bp_inst.cu, one object per precision x method) ----
static bp_kernel_fn bp_kernel_for(const qd_bp *d)
{
    const int t = d->threads * d->ctas_per_sm;
    if (d->precision == QD_PREC_F64)
        return d->method == QD_BP_MIN_SUM ? pick_bp6_f64_ms(d->vmax, t) : pick_bp6_f64_ps(d->vmax, t);
    return d->method == QD_BP_MIN_SUM ? pick_bp6_f32_ms(d->vmax, t) : pick_bp6_f32_ps(d->vmax, t);
}

// threads: the CTA size the kernel will be launched with (0: any — existence checks); wide column tiles are compiled for
// several launch bounds (bp_inst.cu)
static bp7_kernel_fn bp7_kernel_for(const qd_bp *d, int slots, int threads = 0)
{
    const int t = threads > 0 ? threads : BP7_MAXT;
    if (d->precision == QD_PREC_F64)
        return d->method == QD_BP_MIN_SUM ? pick_bp7_f64_ms(d->rd, d->vmax, slots, d->v7_irreg, d->cl7, t)
                                          : pick_bp7_f64_ps(d->rd, d->vmax, slots, d->v7_irreg, d->cl7, t);
    return d->method == QD_BP_MIN_SUM ? pick_bp7_f32_ms(d->rd, d->vmax, slots, d->v7_irreg, d->cl7, t)
                                      : pick_bp7_f32_ps(d->rd, d->vmax, slots, d->v7_irreg, d->cl7, t);
}

static size_t bp_smem(const qd_bp *d, int slots, bool stage)
{
    const GraphDev &g = d->g->dev;
    return bp_smem_layout(g.npos, g.m, g.n, g.fb, g.nwn, g.nwm, slots, d->precision == QD_PREC_F64 ? 8 : 4,
                          d->vmax * d->idx_bytes, stage).total;
}

// default launch geometry: stage the column index + priors in shared memory when that still
// leaves room for a shot, then fill the rest of shared memory with shots
static int bp_default_config(qd_bp *d)
{
    const size_t budget = (size_t)d->g->smem_optin;            // per CTA (227 KB on B200)
    if (bp_smem(d, 1, false) > budget)
        return fail(QD_EUNSUP, "Tanner graph too large for the shared-memory BP kernel (" +
                                   std::to_string(bp_smem(d, 1, false)) + " B per shot > " + std::to_string(budget) + " B)");
    d->stage_idx = bp_smem(d, 1, true) <= budget ? 1 : 0;
    int slots = 1;
    while (slots < 16 && bp_smem(d, slots + 1, d->stage_idx) <= budget) ++slots;
    int ctas = 1;
    // small graphs: several CTAs per SM instead of very many slots per CTA
    while (slots >= 8 && ctas < 4) { slots /= 2; ctas *= 2; }
    const GraphDev &g = d->g->dev;
    const long long work = (long long)slots * std::max(g.n, g.m);
    int threads = ctas > 1 ? 512 : 1024;
    while (threads > 128 && (long long)threads * 2 > work) threads /= 2;
    d->threads = threads; d->slots = slots; d->ctas_per_sm = ctas; d->groups = 1;
    d->smem = bp_smem(d, slots, d->stage_idx);
    return 0;
}

// tables of the generic kernel: per-column message positions and per-chunk common degree
static int bp6_prepare(qd_bp *d)
{
    qd_graph *g = d->g;
    const GraphDev &gd = g->dev;
    int rc;
    {   // per-column message positions, padded to the column tile, ~0 = no edge
        const size_t cnt = (size_t)gd.n * d->vmax;
        if (d->idx_bytes == 2) {
            std::vector<uint16_t> v(cnt, (uint16_t)0xFFFF);
            for (int j = 0; j < gd.n; ++j)
                for (int t = g->col_ptr[j]; t < g->col_ptr[j + 1]; ++t)
                    v[(size_t)j * d->vmax + (t - g->col_ptr[j])] = (uint16_t)g->csc_pos[t];
            if ((rc = upload(d->d_vidx, v))) return rc;
        } else {
            std::vector<uint32_t> v(cnt, 0xFFFFFFFFu);
            for (int j = 0; j < gd.n; ++j)
                for (int t = g->col_ptr[j]; t < g->col_ptr[j + 1]; ++t)
                    v[(size_t)j * d->vmax + (t - g->col_ptr[j])] = g->csc_pos[t];
            if ((rc = upload(d->d_vidx, v))) return rc;
        }
    }
    {   // common degree of each 32-column chunk (0 = mixed): lets a warp run an unpredicated column update
        std::vector<uint8_t> cd((gd.n + 31) / 32, 0);
        for (int c0 = 0; c0 < gd.n; c0 += 32) {
            const int d0 = g->col_ptr[c0 + 1] - g->col_ptr[c0];
            bool same = d0 >= 3 && d0 <= 6 && d0 <= d->vmax;
            for (int j = c0; j < std::min(gd.n, c0 + 32) && same; ++j) same = (g->col_ptr[j + 1] - g->col_ptr[j]) == d0;
            cd[c0 / 32] = same ? (uint8_t)d0 : 0;
        }
        if ((rc = upload(d->d_chunk_deg, cd))) return rc;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// row-regular kernel (bp7_kernel.cuh): record table, launch geometry
// ---------------------------------------------------------------------------------------------
static const int kBp7Slots[] = {8, 6, 4, 3, 2, 1};

static const Bp7Layout &bp7_layout_of(qd_bp *d)
{
    qd_graph *g = d->g;
    const int nb = d->precision == QD_PREC_F64 ? 16 : 32;
    const int key = nb * 100 + d->vmax;
    for (auto &kv : g->bp7_layouts) if (kv.first == key) return kv.second;
    Bp7LayoutProblem P;
    P.m = g->dev.m; P.n = g->dev.n;
    P.row_ptr = g->row_ptr.data(); P.col_idx = g->col_idx.data(); P.col_ptr = g->col_ptr.data();
    P.vmax = d->vmax; P.nb = nb;
    g->bp7_layouts.emplace_back(key, bp7_colour_layout(P));
    return g->bp7_layouts.back().second;
}

static size_t bp7_smem(qd_bp *d, int slots, bool staged)
{
    const GraphDev &g = d->g->dev;
    const Bp7Layout &L = bp7_layout_of(d);
    return bp7_smem_layout(g.m, L.nrec, g.nwn, g.nwm, L.lines, d->rd, d->vmax, slots, d->precision == QD_PREC_F64 ? 8 : 4, staged).total;
}

// can a (slots, threads) geometry be represented?  per-thread state is one bit per owned
// (row | record, slot) in a 32-bit register; message positions are 16-bit element indices
static bool bp7_geometry_ok(qd_bp *d, int slots, int threads)
{
    const GraphDev &g = d->g->dev;
    const Bp7Layout &L = bp7_layout_of(d);
    const int m32 = (g.m + 31) / 32 * 32;
    const Bp7SmemLayout SL = bp7_smem_layout(g.m, L.nrec, g.nwn, g.nwm, L.lines, d->rd, d->vmax, slots, d->precision == QD_PREC_F64 ? 8 : 4, false);
    if (SL.msg_elems > 65535 || bp7_flag_bytes(g.m + 1, slots, d->rd) > 65535 || g.n >= 65535) return false;
    if (threads < 32 || threads > BP7_MAXT || threads % 32) return false;
    if (((m32 + threads - 1) / threads) * slots > 32) return false;
    if (((L.nrec + threads - 1) / threads) * slots > 32) return false;
    return bp7_kernel_for(d, slots) != nullptr;
}

// build and upload the tables for `slots` shots per CTA
static int bp7_prepare(qd_bp *d, int slots)
{
    qd_graph *g = d->g;
    const GraphDev &gd = g->dev;
    const Bp7Layout &L = bp7_layout_of(d);
    const int V = d->vmax, S = slots, RD = d->rd, NB = L.nb, RB = 128 / NB, FP = bp7_flag_pitch(S, RD), FB = bp7_flag_record(RD);
    (void)FP;
    const int m32 = (gd.m + 31) / 32 * 32, RQ = bp7_row_quads(RD, d->v7_irreg != 0);
    (void)RB;
    // element index (slot 0) of a CSR edge; scratch elements live in the extra line group
    auto elem = [&](int e) { return L.index[e] * NB * S + L.colour[e]; };
    auto scratch = [&](int lane) { return L.lines * NB * S + (lane % NB); };
    std::vector<uint32_t> const_bits(std::max(gd.nwn, 1), 0u);
    for (int j = 0; j < gd.n; ++j)
        if (g->col_ptr[j + 1] == g->col_ptr[j]) {   // no edges: the decision never changes (llr <= 0, resp. ratio >= 1)
            const double pr = d->prior_dom[j];
            if (d->method == QD_BP_MIN_SUM ? pr <= 0.0 : pr >= 1.0) const_bits[j >> 5] |= 1u << (j & 31);
        }
    // CSR edge of the t-th edge of a column
    std::vector<int> csr_edge(std::max(gd.nnz, 1));
    for (int j = 0; j < gd.n; ++j)
        for (int t = g->col_ptr[j]; t < g->col_ptr[j + 1]; ++t)
            csr_edge[t] = g->row_ptr[g->csc_row[t]] + g->csc_k[t];
    std::vector<uint16_t> table((size_t)std::max(L.nrec, 1) * 2 * V, 0), col_id(std::max(L.nrec, 1), 0xFFFF);
    std::vector<uint8_t> chunk_deg(std::max(L.nrec / 32, 1), 0);
    std::vector<double> prior_rec(std::max(L.nrec, 1), d->prior_dom[0]);
    int r0 = 0;
    for (int dg = 1; dg <= V; ++dg) {
        const int cnt = (int)L.cls[dg].size(), padded = (cnt + 31) / 32 * 32;
        for (int i = 0; i < padded; ++i) {
            const int r = r0 + i, j = i < cnt ? L.cls[dg][i] : -1;
            uint16_t *mq = &table[(size_t)r * 2 * V], *fq = mq + V;
            for (int t = 0; t < V; ++t) {
                if (j >= 0 && t < dg) {
                    const int ce = g->col_ptr[j] + t;
                    mq[t] = (uint16_t)elem(csr_edge[ce]);
                    fq[t] = (uint16_t)bp7_flag_offset((int)g->csc_row[ce], g->csc_k[ce], gd.m + 1, S, RD);
                } else {
                    mq[t] = (uint16_t)scratch(r);
                    fq[t] = (uint16_t)bp7_flag_offset(gd.m, t % FB, gd.m + 1, S, RD);
                }
            }
            if (j >= 0) { col_id[r] = (uint16_t)j; prior_rec[r] = d->prior_dom[j]; }
            chunk_deg[r >> 5] = (uint8_t)dg;
        }
        r0 += padded;
    }
    std::vector<uint16_t> row_table((size_t)m32 * RQ * 8, 0);
    for (int c = 0; c < m32; ++c) {
        const int deg = c < gd.m ? g->row_ptr[c + 1] - g->row_ptr[c] : 0;
        uint16_t *row = &row_table[(size_t)c * RQ * 8];
        unsigned missing = 0u;
        for (int k = 0; k < RD; ++k) {
            row[k] = (uint16_t)(k < deg ? elem(g->row_ptr[c] + k) : scratch(c));
            if (c < gd.m && k >= deg) missing |= 1u << k;         // no such edge: reads as the neutral message
        }
        if (d->v7_irreg) row[RD] = (uint16_t)missing;
    }
    d->v7_nrec = L.nrec;
    d->v7_lines = L.lines;
    int rc;
    if ((rc = upload(d->d7_table, table))) return rc;
    if ((rc = upload(d->d7_row_table, row_table))) return rc;
    if ((rc = upload(d->d7_col_id, col_id))) return rc;
    if ((rc = upload(d->d7_chunk_deg, chunk_deg))) return rc;
    if ((rc = upload(d->d7_const_bits, const_bits))) return rc;
    if (d->precision == QD_PREC_F64) {
        if ((rc = upload(d->d7_prior_rec, prior_rec))) return rc;
    } else {
        std::vector<float> pf(prior_rec.begin(), prior_rec.end());
        if ((rc = upload(d->d7_prior_rec, pf))) return rc;
    }
    return 0;
}

// set (threads, slots, ctas) for the row-regular kernel; 0 = choose
static int bp7_configure(qd_bp *d, int threads, int slots, int ctas)
{
    const GraphDev &g = d->g->dev;
    const size_t budget = (size_t)d->g->smem_optin;
    const int nrec = bp7_layout_of(d).nrec;
    const int m32 = (g.m + 31) / 32 * 32;
    auto pick_threads = [&](int S) -> int {
        // fewest passes over the owned rows / records, then the fewest idle lanes
        int best = 0; long best_cost = -1;
        for (int t = 32; t <= BP7_MAXT; t += 32) {
            if (!bp7_geometry_ok(d, S, t)) continue;
            const long cost = (long)((m32 + t - 1) / t) * 4 * d->rd + (long)((nrec + t - 1) / t) * 9;
            const long score = cost * 4096 + t;                  // equal passes: fewer threads
            if (best_cost < 0 || score < best_cost) { best_cost = score; best = t; }
        }
        return best;
    };
    int S = slots, T = threads, staged = 1;
    if (S <= 0) {
        // as many shots per SM as shared memory holds: big graphs take one CTA per SM, small ones several
        S = 0;
        for (int cand : kBp7Slots) {
            if (!bp7_kernel_for(d, cand)) continue;
            const bool fits_staged = bp7_smem(d, cand, true) <= budget;
            if (!fits_staged && bp7_smem(d, cand, false) > budget) continue;
            if (!pick_threads(cand)) continue;
            // small graphs: prefer more CTAs over more slots once a CTA holds >= 4 shots
            if (cand > 4 && bp7_smem(d, 4, true) * 2 <= budget && bp7_kernel_for(d, 4) && pick_threads(4)) continue;
            S = cand;
            break;
        }
        if (!S) return fail(QD_EUNSUP, "row-regular BP kernel: graph does not fit shared memory");
    }
    if (!bp7_kernel_for(d, S)) return fail(QD_EINVAL, "slots must be one of the compiled slot counts (1,2,3,4,6,8)");
    if (T <= 0) T = pick_threads(S);
    if (!T || !bp7_geometry_ok(d, S, T))
        return fail(QD_EINVAL, "threads/slots: owned rows x slots and owned columns x slots must be <= 32, threads <= 800");
    size_t smem = bp7_smem(d, S, true);
    if (smem > budget) { staged = 0; smem = bp7_smem(d, S, false); }
    if (smem > budget) return fail(QD_EUNSUP, "configuration does not fit shared memory");
    int C = ctas;
    const size_t per_sm = 228u * 1024u;                           // B200: 228 KB per SM, 1 KB reserved per CTA
    if (C > 1 && staged && (smem + 1024) * C > per_sm) {           // requested CTAs/SM only fit without the staged table
        const size_t un = bp7_smem(d, S, false);
        if ((un + 1024) * C <= per_sm) { staged = 0; smem = un; }
    }
    if (C <= 0) {
        C = (int)std::max<size_t>(1, per_sm / (smem + 1024));
        // the register file holds one launch bound's worth of threads (wide column tiles: see pick7_wide in bp_inst.cu)
        const int bound = d->vmax == 8 && !d->v7_irreg && T <= 640 ? 640 : BP7_MAXT;
        C = std::min(C, std::max(1, bound / T));
        C = std::min(C, 8);
    }
    int rc;
    if (S != d->slots || !d->d7_table.p) { if ((rc = bp7_prepare(d, S))) return rc; }
    d->threads = T; d->slots = S; d->ctas_per_sm = C; d->groups = 1; d->smem = smem; d->v7_staged = staged;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// cluster mode of the row-regular kernel (bp7_kernel.cuh, CL; plan: bp7_cluster_layout.h)
// ---------------------------------------------------------------------------------------------
static const int kBp7cSlots[] = {4, 3, 2, 1};
constexpr int BP7C_RD = 9, BP7C_V = 4;

static bool bp7c_graph_ok(const qd_bp *d)
{
    const GraphDev &gd = d->g->dev;
    return !d->g->big && gd.max_row_deg >= 2 && gd.max_row_deg <= BP7C_RD && gd.max_col_deg >= 1 && gd.max_col_deg <= BP7C_V &&
           gd.m >= 64 && gd.n < 65535;
}

static Bp7ClusterPlan bp7c_plan(qd_bp *d, int C)
{
    qd_graph *g = d->g;
    Bp7ClusterProblem P;
    P.m = g->dev.m; P.n = g->dev.n;
    P.row_ptr = g->row_ptr.data(); P.col_idx = g->col_idx.data(); P.col_ptr = g->col_ptr.data();
    P.csc_row = reinterpret_cast<const int32_t *>(g->csc_row.data()); P.csc_k = g->csc_k.data();
    P.vmax = BP7C_V; P.rd = BP7C_RD; P.nb = d->precision == QD_PREC_F64 ? 16 : 32;
    return bp7_cluster_plan(P, C);
}

static size_t bp7c_smem(const qd_bp *d, const Bp7ClusterPlan &plan, int slots, bool staged)
{
    return bp7_smem_layout(plan.m_max, plan.nrec, d->g->dev.nwn, d->g->dev.nwm, plan.lines, BP7C_RD, BP7C_V, slots,
                           d->precision == QD_PREC_F64 ? 8 : 4, staged).total;
}

static bool bp7c_geometry_ok(const qd_bp *d, const Bp7ClusterPlan &plan, int slots, int threads)
{
    const Bp7SmemLayout SL = bp7_smem_layout(plan.m_max, plan.nrec, d->g->dev.nwn, d->g->dev.nwm, plan.lines, BP7C_RD, BP7C_V, slots,
                                             d->precision == QD_PREC_F64 ? 8 : 4, false);
    if (SL.msg_elems > 65535 || bp7_flag_bytes(plan.m_max + 1, slots, BP7C_RD) > 65535) return false;
    if (threads < 32 || threads > BP7_MAXT || threads % 32) return false;
    if (((plan.m32 + threads - 1) / threads) * slots > 32) return false;
    if (((plan.nrec + threads - 1) / threads) * slots > 32) return false;
    return true;
}

static int bp7c_pick_threads(const qd_bp *d, const Bp7ClusterPlan &plan, int S)
{
    int best = 0; long best_cost = -1;
    for (int t = 32; t <= BP7_MAXT; t += 32) {
        if (!bp7c_geometry_ok(d, plan, S, t)) continue;
        const long cost = (long)((plan.m32 + t - 1) / t) * 4 * BP7C_RD + (long)((plan.nrec + t - 1) / t) * 9;
        const long score = cost * 4096 + t;
        if (best_cost < 0 || score < best_cost) { best_cost = score; best = t; }
    }
    return best;
}

// rough throughput of a geometry: shots in flight per SM, discounted for the shallow software pipeline of few slots
static double bp7_geometry_score(int slots, int ctas_per_cluster)
{
    static const double eff[5] = {0.0, 0.70, 0.88, 0.96, 1.0};
    return (double)slots / ctas_per_cluster * eff[std::min(slots, 4)];
}

// build and upload the per-rank tables of the plan for `slots` shots per cluster
static int bp7c_prepare(qd_bp *d, int slots)
{
    qd_graph *g = d->g;
    const GraphDev &gd = g->dev;
    const Bp7ClusterPlan &plan = d->cl7_plan;
    const int C = plan.C, V = BP7C_V, RD = BP7C_RD, S = slots, NB = plan.nb, RB = 128 / NB;
    const int FP = bp7_flag_pitch(S, RD), FB = bp7_flag_record(RD), RQ = bp7_row_quads(RD, true);
    (void)FP;
    const int nrec = plan.nrec, m32 = plan.m32, M = plan.m_max;
    (void)RB;
    auto elem = [&](const Bp7Pos &ps) { return ps.line * NB * S + ps.colour; };
    auto scratch = [&](int lane) { return plan.lines * NB * S + (lane % NB); };
    std::vector<uint32_t> const_bits(std::max(gd.nwn, 1), 0u);
    for (int j = 0; j < gd.n; ++j)
        if (g->col_ptr[j + 1] == g->col_ptr[j]) {
            const double pr = d->prior_dom[j];
            if (d->method == QD_BP_MIN_SUM ? pr <= 0.0 : pr >= 1.0) const_bits[j >> 5] |= 1u << (j & 31);
        }
    std::vector<uint16_t> table((size_t)C * nrec * 2 * V, 0), col_id((size_t)C * nrec, 0xFFFF);
    std::vector<uint8_t> chunk_deg((size_t)C * (nrec / 32), 0);
    std::vector<double> prior_rec((size_t)C * nrec, d->prior_dom[0]);
    std::vector<uint16_t> row_table((size_t)C * m32 * RQ * 8, 0);
    std::vector<int32_t> row0(C), rows(C), nexp(C);
    const int exp_stride = std::max(plan.nexp_max, 1);
    std::vector<uint2> exports((size_t)C * exp_stride, make_uint2(0u, 0u));
    for (int c = 0; c < C; ++c) {
        const Bp7ClusterRank &R = plan.rank[c];
        row0[c] = R.row0; rows[c] = R.rows; nexp[c] = (int)R.exports.size();
        for (int r = 0; r < nrec; ++r) {
            uint16_t *mq = &table[((size_t)c * nrec + r) * 2 * V], *fq = mq + V;
            const int j = r < R.nrec ? R.rec_col[r] : -1;
            for (int t = 0; t < V; ++t) {
                const Bp7Pos ps = r < R.nrec ? R.rec_pos[(size_t)r * V + t] : Bp7Pos();
                mq[t] = (uint16_t)(ps.line >= 0 ? elem(ps) : scratch(r));
                const int fr = r < R.nrec ? R.rec_flag_row[(size_t)r * V + t] : -1;
                fq[t] = (uint16_t)(fr >= 0 ? bp7_flag_offset(fr, R.rec_flag_k[(size_t)r * V + t], M + 1, S, RD)
                                           : bp7_flag_offset(M, t % FB, M + 1, S, RD));
            }
            if (j >= 0) {
                if (R.rec_owner[r]) col_id[(size_t)c * nrec + r] = (uint16_t)j;
                prior_rec[(size_t)c * nrec + r] = d->prior_dom[j];
            }
        }
        for (int q = 0; q < R.nrec / 32; ++q) chunk_deg[(size_t)c * (nrec / 32) + q] = R.chunk_deg[q];
        for (int r = 0; r < m32; ++r) {
            uint16_t *row = &row_table[((size_t)c * m32 + r) * RQ * 8];
            unsigned missing = 0u;
            for (int k = 0; k < RD; ++k) {
                const Bp7Pos ps = r < R.rows ? R.row_pos[(size_t)r * RD + k] : Bp7Pos();
                row[k] = (uint16_t)(ps.line >= 0 ? elem(ps) : scratch(r));
                if (r < R.rows && ps.line < 0) missing |= 1u << k;     // no such edge: reads as the neutral message
            }
            row[RD] = (uint16_t)missing;
        }
        for (size_t x = 0; x < R.exports.size(); ++x) {
            const auto &e = R.exports[x];
            exports[(size_t)c * exp_stride + x] = make_uint2((uint32_t)elem(e.src) | ((uint32_t)elem(e.dst) << 16), (uint32_t)e.rank);
        }
    }
    d->v7_nrec = nrec;
    d->v7_lines = plan.lines;
    int rc;
    if ((rc = upload(d->d7_table, table)) || (rc = upload(d->d7_row_table, row_table)) || (rc = upload(d->d7_col_id, col_id)) ||
        (rc = upload(d->d7_chunk_deg, chunk_deg)) || (rc = upload(d->d7_const_bits, const_bits)) || (rc = upload(d->d7c_row0, row0)) ||
        (rc = upload(d->d7c_rows, rows)) || (rc = upload(d->d7c_nexp, nexp)) || (rc = upload(d->d7c_export, exports)))
        return rc;
    if (d->precision == QD_PREC_F64) {
        if ((rc = upload(d->d7_prior_rec, prior_rec))) return rc;
    } else {
        std::vector<float> pf(prior_rec.begin(), prior_rec.end());
        if ((rc = upload(d->d7_prior_rec, pf))) return rc;
    }
    return 0;
}

static void bp7c_launch_config(const qd_bp *d, int C, int threads, size_t smem, unsigned grid, cudaStream_t st, cudaLaunchConfig_t &cfg,
                               cudaLaunchAttribute *attr)
{
    (void)d;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = attr; cfg.numAttrs = 1;
}

// choose (cluster size, slots, threads) for the cluster mode (0 = choose) and upload its tables.  On success d->cl7 = 1 and
// d->rd / vmax / v7_irreg describe the compiled cluster variant; *score (optional) receives bp7_geometry_score.
static int bp7c_configure(qd_bp *d, int want_C, int want_slots, int want_threads, double *score, bool any_cut = false)
{
    if (!bp7c_graph_ok(d)) return fail(QD_EUNSUP, "cluster mode of the row-regular BP kernel: rows up to 9, columns up to 4 edges");
    const size_t budget = (size_t)d->g->smem_optin;
    const bool dbg = getenv("QD_BPC_DEBUG") != nullptr;
    const int save_rd = d->rd, save_vmax = d->vmax, save_irreg = d->v7_irreg, save_cl7 = d->cl7;
    d->rd = BP7C_RD; d->vmax = BP7C_V; d->v7_irreg = 1; d->cl7 = 1;
    auto restore = [&]() { d->rd = save_rd; d->vmax = save_vmax; d->v7_irreg = save_irreg; d->cl7 = save_cl7; };
    double best_score = -1.0;
    int best_C = 0, best_S = 0, best_T = 0, best_staged = 0, best_clusters = 0;
    size_t best_smem = 0;
    Bp7ClusterPlan best_plan;
    // Cluster sizes 2, 4, 8 only.  A forced split of the [[225,9]] matrix over THREE CTAs (product-sum, two slots) hung about once
    // in thirty 512-shot launches (scripts/stress_cluster.py; only with the deferred end-of-iteration wait) and the cause was
    // not found; the same case over two CTAs, and every other geometry tried, never did.  DESIGN.md section 3.2a.
    const bool any_size = getenv("QD_BP7C_ANY_SIZE") != nullptr;          // (experiments on that problem)
    if (want_C > 0 && want_C != 2 && want_C != 4 && want_C != 8 && !any_size) {
        restore();
        return fail(QD_EINVAL, "cluster mode of the row-regular BP kernel: the cluster size must be 2, 4 or 8");
    }
    for (int C = 2; C <= 8; C = any_size && want_C > 0 ? C + 1 : C * 2) {
        if (want_C > 0 && C != want_C) continue;
        if (d->g->dev.m < 32 * C) continue;
        Bp7ClusterPlan plan = bp7c_plan(d, C);
        if (!plan.ok) continue;
        {
            qd_graph *g = d->g;
            Bp7ClusterProblem P;
            P.m = g->dev.m; P.n = g->dev.n;
            P.row_ptr = g->row_ptr.data(); P.col_idx = g->col_idx.data(); P.col_ptr = g->col_ptr.data();
            P.csc_row = reinterpret_cast<const int32_t *>(g->csc_row.data()); P.csc_k = g->csc_k.data();
            P.vmax = BP7C_V; P.rd = BP7C_RD; P.nb = plan.nb;
            if (bp7_cluster_plan_errors(P, plan) != 0) { restore(); return fail(QD_ECUDA, "cluster plan failed its self-check"); }
        }
        // a split that cuts a large part of the columns moves the messages through the cluster window instead of shared
        // memory: only splits along the block structure of the matrix are worth it (unless the caller insists)
        if (!any_cut && want_C <= 0 && plan.mailbox_elems * 3 > plan.local_edges) continue;
        for (int S : kBp7cSlots) {
            if (want_slots > 0 && S != want_slots) continue;
            if (!bp7_kernel_for(d, S)) continue;
            int staged = 1;
            size_t smem = bp7c_smem(d, plan, S, true);
            if (smem > budget) { staged = 0; smem = bp7c_smem(d, plan, S, false); }
            if (smem > budget) continue;
            const int T = want_threads > 0 ? want_threads : bp7c_pick_threads(d, plan, S);
            if (!T || !bp7c_geometry_ok(d, plan, S, T)) continue;
            bp7_kernel_fn fn = bp7_kernel_for(d, S, T);
            if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); continue; }
            cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[1];
            bp7c_launch_config(d, C, T, smem, (unsigned)(C * std::max(1, d->g->sm_count / C)), nullptr, cfg, attr);
            int nclusters = 0;
            const cudaError_t oe = cudaOccupancyMaxActiveClusters(&nclusters, fn, &cfg);
            if (dbg) fprintf(stderr, "bp7c: C=%d S=%d T=%d smem=%zu staged=%d cut=%ld/%ld clusters=%d (%s)\n", C, S, T, smem, staged,
                             plan.mailbox_elems, plan.local_edges, nclusters, cudaGetErrorString(oe));
            if (oe != cudaSuccess || nclusters < 1) { cudaGetLastError(); continue; }
            // SMs actually used by whole clusters count, not the nominal S / C
            const double sc = bp7_geometry_score(S, C) * ((double)nclusters * C / std::max(1, d->g->sm_count));
            if (sc > best_score) {
                best_score = sc; best_C = C; best_S = S; best_T = T; best_staged = staged; best_smem = smem; best_clusters = nclusters;
                best_plan = plan;
            }
            break;                                                // fewer slots of the same split are never better
        }
    }
    if (best_C == 0) { restore(); return fail(QD_EUNSUP, "cluster mode of the row-regular BP kernel: no split of the rows fits"); }
    d->cl7_plan = std::move(best_plan);
    d->cl7_C = best_C; d->cl7_clusters = best_clusters;
    int rc = bp7c_prepare(d, best_S);
    if (rc) { restore(); return rc; }
    d->threads = best_T; d->slots = best_S; d->ctas_per_sm = 1; d->groups = 1; d->smem = best_smem; d->v7_staged = best_staged;
    if (score) *score = best_score;
    return 0;
}


template <typename real, int METHOD>
static void (*bpc_fn())(const BpcParams) { return bpc_kernel<real, METHOD>; }

static void (*bpc_kernel_for(const qd_bp *d))(const BpcParams)
{
    return d->precision == QD_PREC_F64 ? (d->method == QD_BP_MIN_SUM ? bpc_fn<double, QD_BP_MIN_SUM>() : bpc_fn<double, QD_BP_PRODUCT_SUM>())
                                       : (d->method == QD_BP_MIN_SUM ? bpc_fn<float, QD_BP_MIN_SUM>() : bpc_fn<float, QD_BP_PRODUCT_SUM>());
}

// Partition the graph over a cluster of C CTAs (smallest C of 2, 4, 8 whose per-CTA share fits shared memory) and
// upload the tables of bpc_kernel.cuh.  Rows: contiguous ranges of about nnz / C edges.  Columns: the CTA holding
// most of the column's edges (ties: the lowest rank).
static int bpc_prepare(qd_bp *d)
{
    qd_graph *g = d->g;
    const GraphDev &gd = g->dev;
    const int rb = d->precision == QD_PREC_F64 ? 8 : 4;
    const int m = gd.m, n = gd.n, nnz = gd.nnz;
    if (m < 8 || nnz < 8) return fail(QD_EUNSUP, "cluster BP kernel: graph too small to split");
    for (int C = 2; C <= 8; C *= 2) {
        std::vector<int32_t> row_lo(C + 1, 0);
        row_lo[C] = m;
        if (gd.max_row_deg <= 32) {
            // a thread per row: equal row counts, so that no CTA needs one more pass over its rows than the others
            for (int c = 1; c < C; ++c) row_lo[c] = (int)((long long)m * c / C);
        } else {
            for (int c = 1, r = 0; c < C; ++c) {
                const long long want = (long long)nnz * c / C;
                while (r < m && g->row_ptr[r] < want) ++r;
                row_lo[c] = std::max(r, row_lo[c - 1]);
            }
        }
        int max_edges = 0, max_rows = 0;
        std::vector<int32_t> row_start(std::max(m, 1), 0);
        for (int c = 0; c < C; ++c) {
            int pos = 0;
            for (int r = row_lo[c]; r < row_lo[c + 1]; ++r) {
                const int deg = g->row_ptr[r + 1] - g->row_ptr[r];
                row_start[r] = pos;
                pos += deg | 1;                                   // odd pitch
            }
            max_edges = std::max(max_edges, pos);
            max_rows = std::max(max_rows, row_lo[c + 1] - row_lo[c]);
        }
        if (max_edges >= (1 << 24)) continue;
        std::vector<int> row_owner(m);
        for (int c = 0; c < C; ++c) for (int r = row_lo[c]; r < row_lo[c + 1]; ++r) row_owner[r] = c;
        std::vector<int> col_owner(n, 0);
        std::vector<std::vector<int32_t>> cols(C);
        for (int j = 0; j < n; ++j) {
            int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int t = g->col_ptr[j]; t < g->col_ptr[j + 1]; ++t) ++cnt[row_owner[g->csc_row[t]]];
            int best = g->col_ptr[j + 1] == g->col_ptr[j] ? (int)((long long)j * C / n) : 0;   // isolated columns: spread
            for (int c = 1; c < C; ++c) if (cnt[c] > cnt[best]) best = c;
            col_owner[j] = best;
            cols[best].push_back(j);
        }
        std::vector<int32_t> lc_lo(C + 1, 0), word_lo(C + 1, 0);
        int max_cols = 0, max_cedges = 0;
        for (int c = 0; c < C; ++c) {
            lc_lo[c + 1] = lc_lo[c] + (int)cols[c].size();
            word_lo[c + 1] = word_lo[c] + ((int)cols[c].size() + 31) / 32;
            max_cols = std::max(max_cols, (int)cols[c].size());
            int ce = 0;
            for (int j : cols[c]) ce += g->col_ptr[j + 1] - g->col_ptr[j];
            max_cedges = std::max(max_cedges, ce);
        }
        const int nwpos = std::max(word_lo[C], 1);
        const size_t smem = bpc_smem_layout(max_edges, max_rows, max_cols, max_cedges, nwpos, gd.nwn, rb).total;
        const bool dbg = getenv("QD_BPC_DEBUG") != nullptr;
        if (dbg) fprintf(stderr, "bpc: C=%d max_edges=%d max_rows=%d max_cols=%d max_cedges=%d smem=%zu (optin %d)\n", C, max_edges,
                         max_rows, max_cols, max_cedges, smem, g->smem_optin);
        if (smem + 256 > (size_t)g->smem_optin) continue;
        // tables
        std::vector<int32_t> lc_col(std::max(n, 1)), lc_ptr(n + 1, 0), col_pos(std::max(n, 1)), edge_pos(std::max(nnz, 1));
        std::vector<uint32_t> desc(std::max(nnz, 1));
        std::vector<double> prior(std::max(n, 1));
        int li = 0, dp = 0;
        for (int c = 0; c < C; ++c)
            for (size_t i = 0; i < cols[c].size(); ++i, ++li) {
                const int j = cols[c][i];
                lc_col[li] = j;
                prior[li] = d->prior_dom[j];
                col_pos[j] = (word_lo[c] << 5) + (int)i;
                lc_ptr[li] = dp;
                for (int t = g->col_ptr[j]; t < g->col_ptr[j + 1]; ++t) {
                    const int e = g->csc_edge[t], row = (int)g->csc_row[t], o = row_owner[row];
                    desc[dp++] = ((uint32_t)o << 24) | (uint32_t)(row_start[row] + (e - g->row_ptr[row]));
                }
            }
        lc_ptr[n] = dp;
        for (int e = 0; e < nnz; ++e) edge_pos[e] = col_pos[g->col_idx[e]];
        int rc;
        if ((rc = upload(d->dc_row_lo, row_lo)) || (rc = upload(d->dc_row_start, row_start)) || (rc = upload(d->dc_edge_pos, edge_pos)) || (rc = upload(d->dc_lc_lo, lc_lo)) ||
            (rc = upload(d->dc_lc_col, lc_col)) || (rc = upload(d->dc_lc_ptr, lc_ptr)) || (rc = upload(d->dc_desc, desc)) ||
            (rc = upload(d->dc_word_lo, word_lo)) || (rc = upload(d->dc_col_pos, col_pos)))
            return rc;
        if (rb == 8) { if ((rc = upload(d->dc_prior, prior))) return rc; }
        else { std::vector<float> pf(prior.begin(), prior.end()); if ((rc = upload(d->dc_prior, pf))) return rc; }
        // threads per CTA: the fewest passes over the CTA's rows and columns (a column pass costs about as much as a
        // row pass), then the fewest threads; 800 threads leave 81 registers
        int threads = 128;
        {
            long best = -1;
            for (int t = 128; t <= 800; t += 32) {
                const long cost = (long)((max_rows + t - 1) / t) + (long)((max_cols + t - 1) / t);
                if (best < 0 || cost < best) { best = cost; threads = t; }
            }
        }
        // how many clusters of this shape the device runs at once
        auto fn = bpc_kernel_for(d);
        QD_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaLaunchConfig_t cfg{};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.gridDim = dim3((unsigned)(C * std::max(1, g->sm_count / C)));
        cfg.blockDim = dim3((unsigned)threads);
        cfg.dynamicSmemBytes = smem;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int nclusters = 0;
        const cudaError_t oe = cudaOccupancyMaxActiveClusters(&nclusters, fn, &cfg);
        if (dbg) fprintf(stderr, "bpc: C=%d threads=%d clusters=%d (%s)\n", C, threads, nclusters, cudaGetErrorString(oe));
        if (oe != cudaSuccess || nclusters < 1) {
            cudaGetLastError();
            continue;
        }
        d->cl_C = C; d->cl_max_edges = max_edges; d->cl_max_rows = max_rows; d->cl_max_cols = max_cols; d->cl_max_cedges = max_cedges;
        d->cl_nwpos = nwpos; d->cl_clusters = nclusters;
        d->threads = threads; d->slots = 1; d->ctas_per_sm = 1; d->groups = 1; d->smem = smem;
        return 0;
    }
    return fail(QD_EUNSUP, "cluster BP kernel: the graph does not fit the shared memory of 8 SMs");
}

// pick the kernel family and its default launch geometry.  Row-regular graphs with a compiled
// (row degree, column tile) pair take the thread-owned-state kernel (bp7_kernel.cuh) when they fit;
// everything else the generic kernel (bp_kernel.cuh).
// regular_cluster: 1 = force the cluster mode of the row-regular kernel, -1 = never use it, 0 = when it pays
static int bp_choose_kernel(qd_bp *d, bool allow_regular, bool family_global = false, bool family_cluster = false,
                            int regular_cluster = 0)
{
    const bool family_regular_cluster = regular_cluster > 0;
    const GraphDev &gd = d->g->dev;
    int rc;
    d->v7 = 0;
    d->cl7 = 0;
    d->global_mem = 0;
    d->cluster = 0;
    if (family_regular_cluster) {
        if ((rc = bp7c_configure(d, 0, 0, 0, nullptr, true))) return rc;       // forced: any graph the kernel can hold
        d->v7 = 1;
        return 0;
    }
    const bool shared_ok = !d->g->big && gd.max_col_deg <= 16 && !(d->method == QD_BP_PRODUCT_SUM && gd.max_row_deg > 16);
    if (family_global || family_cluster || !shared_ok) { d->global_mem = 1; }
    else if (allow_regular && gd.max_row_deg >= 2 && gd.max_col_deg >= 1 && gd.max_col_deg <= 8) {
        // candidates: exact (row degree, column tile) for row-regular graphs, else the padded variants
        const int vm0 = gd.max_col_deg <= 4 ? 4 : 8;
        bool found = false;
        if (gd.regular_deg > 0)
            for (int vm = vm0; vm <= 8 && !found; vm += 4) {
                d->rd = gd.regular_deg; d->vmax = vm; d->v7_irreg = 0;
                found = bp7_kernel_for(d, 1) != nullptr;
            }
        for (int rd = gd.max_row_deg; rd <= 9 && !found; ++rd)
            for (int vm = vm0; vm <= 8 && !found; vm += 4) {
                d->rd = rd; d->vmax = vm; d->v7_irreg = 1;
                found = bp7_kernel_for(d, 1) != nullptr;
            }
        if (found) {
            d->v7 = 1;
            d->slots = 0;
            d->d7_table.alloc(0);
            if (bp7_configure(d, 0, 0, 0) != 0) d->v7 = 0;     // does not fit: generic kernel
        }
        // Graphs that leave one SM with one or two shots in flight, or none: split the rows over a thread-block cluster
        // when the matrix has the block structure for it (spacetime matrices) and more shots per SM come out of it.
        // (Measured at n=1600, T=1: one shot per SM in the single-CTA kernel, 2.98 M shots/s, still beats three shots per
        // two-CTA cluster, 2.70 M — barriers and the export cost more than the extra shots in flight bring. So: only
        // when the graph does not fit one SM.)
        if (regular_cluster == 0 && !d->v7 && bp7c_graph_ok(d)) {
            const double single = d->v7 ? bp7_geometry_score(d->slots, 1) : 0.0;
            const int save_threads = d->threads, save_slots = d->slots, save_ctas = d->ctas_per_sm, save_staged = d->v7_staged;
            const int save_rd = d->rd, save_vmax = d->vmax, save_irreg = d->v7_irreg;
            const size_t save_smem = d->smem;
            double sc = 0.0;
            if (bp7c_configure(d, 0, 0, 0, &sc) == 0 && sc > 1.1 * single) {
                d->v7 = 1;                                        // bp7c_configure left d->cl7 = 1 and the cluster tables
            } else {
                d->cl7 = 0;
                d->rd = save_rd; d->vmax = save_vmax; d->v7_irreg = save_irreg;
                d->threads = save_threads; d->slots = save_slots; d->ctas_per_sm = save_ctas; d->v7_staged = save_staged; d->smem = save_smem;
                // the two modes share the table buffers: put the single-CTA tables back
                if (d->v7 && (rc = bp7_prepare(d, d->slots))) return rc;
            }
        }
    }
    if (!d->v7 && !d->global_mem) {
        d->rd = 0;
        d->v7_irreg = 0;
        d->vmax = gd.max_col_deg <= 8 ? 8 : 16;
        d->idx_bytes = 4;
        if ((rc = bp6_prepare(d))) return rc;
        if (bp_default_config(d) != 0) d->global_mem = 1;       // does not fit shared memory
    }
    d->cluster = 0;
    if (d->global_mem && !family_global && bpc_prepare(d) == 0) {   // too large for one SM: spread over a cluster
        d->global_mem = 0; d->cluster = 1;
        d->rd = 0; d->v7_irreg = 0;
    }
    if (d->global_mem) {
        d->rd = 0; d->v7_irreg = 0;
        const size_t smem = (size_t)((gd.n + 15) & ~15) + ((gd.m + 15) & ~15) + (size_t)gd.nwn * 4 + 16;
        if (smem > (size_t)d->g->smem_optin)
            return fail(QD_EUNSUP, "graph too large even for the global-memory BP kernel (decisions + syndrome exceed shared memory)");
        d->threads = 512; d->slots = 1; d->ctas_per_sm = std::max(1, std::min(4, (int)((200u * 1024u) / (smem + 1024)))); d->groups = 1;
        d->smem = smem;
    }
    return 0;
}

// host-only: plan the cluster split of a graph and run the plan's self-check (no device needed)
extern "C" int qd_bp_cluster_plan_check(int32_t m, int32_t n, const int32_t *row_ptr, const int32_t *col_idx, int32_t ctas,
                                        int32_t banks, int64_t *stats)
{
    if (m <= 0 || n <= 0 || !row_ptr || !col_idx || !stats) return fail(QD_EINVAL, "bad arguments");
    if (banks != 16 && banks != 32) return fail(QD_EINVAL, "banks must be 16 (8-byte messages) or 32 (4-byte messages)");
    const int nnz = row_ptr[m];
    std::vector<int32_t> col_ptr(n + 1, 0), csc_row(std::max(nnz, 1));
    std::vector<uint8_t> csc_k(std::max(nnz, 1));
    for (int r = 0; r < m; ++r)
        for (int e = row_ptr[r]; e < row_ptr[r + 1]; ++e) {
            if (col_idx[e] < 0 || col_idx[e] >= n) return fail(QD_EINVAL, "column index out of range");
            if (e > row_ptr[r] && col_idx[e] <= col_idx[e - 1]) return fail(QD_EINVAL, "column indices must be sorted and distinct within a row");
            ++col_ptr[col_idx[e] + 1];
        }
    for (int j = 0; j < n; ++j) col_ptr[j + 1] += col_ptr[j];
    {
        std::vector<int32_t> fill(col_ptr.begin(), col_ptr.end() - 1);
        for (int r = 0; r < m; ++r)
            for (int e = row_ptr[r]; e < row_ptr[r + 1]; ++e) {
                const int at = fill[col_idx[e]]++;
                csc_row[at] = r; csc_k[at] = (uint8_t)(e - row_ptr[r]);
            }
    }
    Bp7ClusterProblem P;
    P.m = m; P.n = n; P.row_ptr = row_ptr; P.col_idx = col_idx; P.col_ptr = col_ptr.data(); P.csc_row = csc_row.data(); P.csc_k = csc_k.data();
    P.vmax = BP7C_V; P.rd = BP7C_RD; P.nb = banks;
    const Bp7ClusterPlan plan = bp7_cluster_plan(P, ctas);
    if (!plan.ok) return fail(QD_EUNSUP, "no cluster plan: rows up to 9 and columns up to 4 edges, 2 <= ctas <= rows");
    stats[0] = bp7_cluster_plan_errors(P, plan);
    stats[1] = plan.local_edges; stats[2] = plan.mailbox_elems; stats[3] = plan.lines;
    stats[4] = plan.nrec; stats[5] = plan.m_max; stats[6] = plan.nexp_max; stats[7] = 0;
    return 0;
}

extern "C" int qd_bp_select_kernel(qd_bp *d, int32_t family)
{
    if (!d) return fail(QD_EINVAL, "NULL handle");
    if (family != 0 && family != 5 && family != 6 && family != 7 && family != 8 && family != 9)
        return fail(QD_EINVAL, "kernel family must be 0 (automatic), 5 (global memory), 6 (generic), 7 (thread-owned state), 8 (cluster) "
                               "or 9 (thread-owned state over a cluster)");
    int rc = use_device(d->g->device);
    if (rc) return rc;
    if ((rc = bp_choose_kernel(d, family != 6, family == 5, family == 8, family == 9 ? 1 : (family == 7 ? -1 : 0)))) return rc;
    if (family == 8 && !d->cluster) return fail(QD_EUNSUP, "the cluster BP kernel does not support this graph");
    if (family == 7 && !d->v7) return fail(QD_EUNSUP, "the thread-owned-state BP kernel does not support this graph");
    if (family == 6 && d->global_mem) return fail(QD_EUNSUP, "the generic shared-memory BP kernel does not support this graph");
    return 0;
}

extern "C" int qd_bp_get_kernel(qd_bp *d, int32_t *family, int32_t *row_degree, int32_t *column_tile)
{
    if (!d) return fail(QD_EINVAL, "NULL handle");
    if (family) *family = d->cluster ? 8 : (d->global_mem ? 5 : (d->v7 ? (d->cl7 ? 9 : 7) : 6));
    if (row_degree) *row_degree = d->rd;
    if (column_tile) *column_tile = d->vmax;
    return 0;
}

extern "C" int qd_bp_create(qd_graph *g, const double *h_channel_probs, int32_t max_iter, int32_t method,
                            double ms_scaling_factor, int32_t precision, qd_bp **out)
{
    if (!out) return fail(QD_EINVAL, "out is NULL");
    *out = nullptr;
    if (!g || !h_channel_probs) return fail(QD_EINVAL, "graph or channel_probs is NULL");
    if (method != QD_BP_PRODUCT_SUM && method != QD_BP_MIN_SUM) return fail(QD_EINVAL, "unknown bp method");
    if (precision != QD_PREC_F64 && precision != QD_PREC_F32) return fail(QD_EINVAL, "precision must be 64 or 32");
    if (max_iter < 0) return fail(QD_EINVAL, "max_iter < 0");
    int rc = use_device(g->device);
    if (rc) return rc;
    const GraphDev &gd = g->dev;
    std::unique_ptr<qd_bp> d(new qd_bp);
    d->g = g;
    d->device = g->device;
    d->max_iter = max_iter > 0 ? max_iter : gd.n;   // ldpc convention: 0 -> n
    d->method = method;
    d->precision = precision;
    d->alpha_fixed = ms_scaling_factor;
    d->rd = 0;
    d->vmax = gd.max_col_deg <= 8 ? 8 : 16;
    d->idx_bytes = 4;
    // priors in the decoder's domain, computed in double with the host libm (same as the oracle)
    std::vector<double> pr(gd.n), al(d->max_iter + 1, 0.0);
    for (int j = 0; j < gd.n; ++j) {
        const double p = h_channel_probs[j];
        if (!(p > 0.0) || !(p < 1.0)) return fail(QD_EINVAL, "channel probabilities must lie in (0, 1)");
        pr[j] = method == QD_BP_MIN_SUM ? std::log((1.0 - p) / p) : p / (1.0 - p);
    }
    d->prior_dom = pr;
    d->v7_uniform = 1;
    for (int j = 1; j < gd.n; ++j) if (pr[j] != pr[0]) { d->v7_uniform = 0; break; }
    d->v7_prior0 = pr[0];
    // al[0] = the fixed scaling factor (0 selects the 1 - 2^-it schedule, formed exactly on the device);
    // al[1..] = the schedule as the host computes it (reference values, checked against the device formula)
    al[0] = ms_scaling_factor;
    for (int it = 1; it <= d->max_iter; ++it)
        al[it] = ms_scaling_factor == 0.0 ? 1.0 - std::pow(2.0, -1.0 * it / 1.0) : ms_scaling_factor;
    for (int it = 1; it <= std::min(d->max_iter, 1100); ++it)
        if (ms_scaling_factor == 0.0 && al[it] != 1.0 - std::ldexp(1.0, -it)) return fail(QD_ECUDA, "alpha schedule mismatch");
    if (precision == QD_PREC_F64) {
        if ((rc = upload(d->d_prior, pr))) return rc;
        if ((rc = upload(d->d_alpha, al))) return rc;
    } else {
        std::vector<float> prf(pr.begin(), pr.end()), alf(al.begin(), al.end());
        if ((rc = upload(d->d_prior, prf))) return rc;
        if ((rc = upload(d->d_alpha, alf))) return rc;
    }
    if ((rc = d->d_next.alloc(sizeof(unsigned long long) * 2))) return rc;
    if ((rc = bp_choose_kernel(d.get(), true))) return rc;
    *out = d.release();
    return 0;
}

extern "C" int qd_bp_destroy(qd_bp *d)
{
    if (d) { cudaSetDevice(d->device); delete d; }
    return 0;
}

extern "C" int qd_bp_configure(qd_bp *d, int32_t threads, int32_t slots, int32_t ctas_per_sm, int32_t teams)
{
    if (!d) return fail(QD_EINVAL, "NULL handle");
    if (d->global_mem) {
        if (threads > 0) { if (threads % 32 || threads > 512) return fail(QD_EINVAL, "threads must be a multiple of 32 <= 512"); d->threads = threads; }
        if (ctas_per_sm > 0) d->ctas_per_sm = ctas_per_sm;
        return 0;
    }
    if (d->v7) {
        if (teams > 1) return fail(QD_EINVAL, "the row-regular BP kernel runs one team per CTA");
        int rc = use_device(d->g->device);
        if (rc) return rc;
        if (d->cl7)          // cluster mode: `ctas_per_sm` is read as the cluster size (CTAs per shot group)
            return bp7c_configure(d, ctas_per_sm > 0 ? ctas_per_sm : d->cl7_C, slots > 0 ? slots : d->slots,
                                  threads > 0 ? threads : (slots > 0 ? 0 : d->threads), nullptr, true);
        return bp7_configure(d, threads > 0 ? threads : (slots > 0 ? 0 : d->threads), slots > 0 ? slots : d->slots,
                             ctas_per_sm > 0 ? ctas_per_sm : (slots > 0 || threads > 0 ? 0 : d->ctas_per_sm));
    }
    int t = threads > 0 ? threads : d->threads, s = slots > 0 ? slots : d->slots, c = ctas_per_sm > 0 ? ctas_per_sm : d->ctas_per_sm;
    int gr = teams > 0 ? teams : (slots > 0 || threads > 0 ? 1 : d->groups);
    if (t % 32 || t > 1024 || s > 64) return fail(QD_EINVAL, "threads must be a multiple of 32 <= 1024, slots <= 64");
    if (gr > 15 || s % gr || t % (32 * gr)) return fail(QD_EINVAL, "teams must divide slots and threads/32 (and be <= 15)");
    if (s / gr > t / gr) return fail(QD_EINVAL, "slots per team must not exceed threads per team");
    int stage = 1;
    size_t smem = bp_smem(d, s, true);
    if (smem > (size_t)d->g->smem_optin) { stage = 0; smem = bp_smem(d, s, false); }
    if (smem > (size_t)d->g->smem_optin) return fail(QD_EUNSUP, "configuration does not fit shared memory");
    d->threads = t; d->slots = s; d->ctas_per_sm = c; d->groups = gr; d->smem = smem; d->stage_idx = stage;
    return 0;
}

extern "C" int qd_bp_get_config(qd_bp *d, int32_t *threads, int32_t *slots, int32_t *ctas_per_sm, int32_t *teams,
                                int32_t *smem_bytes, int32_t *grid)
{
    if (d && teams) *teams = d->groups;
    if (!d) return fail(QD_EINVAL, "NULL handle");
    if (threads) *threads = d->threads;
    if (slots) *slots = d->slots;
    if (ctas_per_sm) *ctas_per_sm = d->ctas_per_sm;
    if (smem_bytes) *smem_bytes = (int32_t)d->smem;
    if (grid) *grid = d->cl7 ? d->cl7_clusters * d->cl7_C : d->g->sm_count * d->ctas_per_sm;
    if (ctas_per_sm && d->cl7) *ctas_per_sm = d->cl7_C;       // cluster mode: CTAs per cluster
    return 0;
}

static int bp_launch(qd_bp *d, const uint32_t *d_synd, int32_t synd_stride, int64_t shots, uint32_t *d_corr,
                     int32_t corr_stride, uint8_t *d_converged, int32_t *d_iters, double *d_llr, const uint32_t *d_err,
                     int32_t err_stride, qd_logicals *logicals, uint8_t *d_fail, int64_t *d_counters, void *stream,
                     int queue)
{
    if (!d || !d_synd || !d_corr || shots < 0) return fail(QD_EINVAL, "bad bp_decode_batch arguments");
    const GraphDev &g = d->g->dev;
    if (synd_stride < g.nwm || corr_stride < g.nwn) return fail(QD_EINVAL, "stride too small");
    if (d_err && (!logicals || err_stride < g.nwn)) return fail(QD_EINVAL, "fused epilogue needs logicals and err_stride >= ceil(n/32)");
    if (logicals && logicals->n != g.n) return fail(QD_EINVAL, "logicals act on a different number of qubits");
    if (shots == 0) return 0;
    int rc = use_device(d->g->device);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (d->cluster) {
        BpcParams q{};
        q.m = g.m; q.n = g.n; q.nnz = g.nnz; q.nwn = g.nwn; q.C = d->cl_C;
        q.max_edges = d->cl_max_edges; q.max_rows = d->cl_max_rows; q.max_cols = d->cl_max_cols; q.max_cedges = d->cl_max_cedges;
        q.nwpos = d->cl_nwpos;
        q.row_lo = d->dc_row_lo.as<int32_t>(); q.row_ptr = g.row_ptr; q.row_start = d->dc_row_start.as<int32_t>(); q.edge_pos = d->dc_edge_pos.as<int32_t>();
        q.lc_lo = d->dc_lc_lo.as<int32_t>(); q.lc_col = d->dc_lc_col.as<int32_t>(); q.lc_ptr = d->dc_lc_ptr.as<int32_t>();
        q.desc = d->dc_desc.as<uint32_t>(); q.word_lo = d->dc_word_lo.as<int32_t>(); q.col_pos = d->dc_col_pos.as<int32_t>();
        q.prior = d->dc_prior.p; q.alpha_fixed = d->alpha_fixed; q.max_iter = d->max_iter;
        q.synd = d_synd; q.synd_stride = synd_stride; q.shots = shots;
        q.corr = d_corr; q.corr_stride = corr_stride;
        q.conv = d_converged; q.iters = d_iters; q.llr = d_llr; q.llr_mode = d->llr_mode;
        q.err = d_err; q.err_stride = err_stride;
        q.lzT = d_err ? logicals->d_lzT.as<uint32_t>() : nullptr;
        q.k = d_err ? logicals->k : 0;
        q.fail = d_fail;
        q.counters = (long long *)d_counters;
        q.next_shot = d->d_next.as<unsigned long long>() + queue;
        auto fn = bpc_kernel_for(d);
        QD_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem));
        QD_CUDA(cudaMemsetAsync(q.next_shot, 0, sizeof(unsigned long long), st));
        cudaLaunchConfig_t cfg{};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)d->cl_C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        const long long nclusters = std::min<long long>(d->cl_clusters, shots);
        cfg.gridDim = dim3((unsigned)(nclusters * d->cl_C));
        cfg.blockDim = dim3((unsigned)d->threads);
        cfg.dynamicSmemBytes = d->smem;
        cfg.stream = st;
        cfg.attrs = attr; cfg.numAttrs = 1;
        QD_CUDA(cudaLaunchKernelEx(&cfg, fn, q));
        return 0;
    }
    if (d->global_mem) {
        qd_graph *gr = d->g;
        const unsigned grid = (unsigned)std::min<long long>((long long)gr->sm_count * d->ctas_per_sm, shots);
        const size_t rb = d->precision == QD_PREC_F64 ? 8 : 4;
        DevBuf &scratch = d->d_scratch[queue & 1];
        if ((rc = scratch.ensure((size_t)gr->sm_count * d->ctas_per_sm * 2 * std::max(g.nnz, 1) * rb))) return rc;
        BpgParams q{};
        q.m = g.m; q.n = g.n; q.nnz = g.nnz; q.nwm = g.nwm; q.nwn = g.nwn;
        q.row_ptr = g.row_ptr; q.col_idx = g.col_idx;
        q.col_ptr = gr->d_col_ptr.as<int32_t>(); q.csc_edge = gr->d_csc_edge.as<int32_t>();
        q.prior = d->d_prior.p; q.alpha_fixed = d->alpha_fixed; q.max_iter = d->max_iter;
        q.scratch = scratch.p;
        q.synd = d_synd; q.synd_stride = synd_stride; q.shots = shots;
        q.corr = d_corr; q.corr_stride = corr_stride;
        q.conv = d_converged; q.iters = d_iters; q.llr = d_llr; q.llr_mode = d->llr_mode;
        q.err = d_err; q.err_stride = err_stride;
        q.lzT = d_err ? logicals->d_lzT.as<uint32_t>() : nullptr;
        q.k = d_err ? logicals->k : 0;
        q.fail = d_fail;
        q.counters = (long long *)d_counters;
        q.next_shot = d->d_next.as<unsigned long long>() + queue;
        void (*fn)(const BpgParams) =
            d->precision == QD_PREC_F64 ? (d->method == QD_BP_MIN_SUM ? bpg_kernel<double, QD_BP_MIN_SUM> : bpg_kernel<double, QD_BP_PRODUCT_SUM>)
                                        : (d->method == QD_BP_MIN_SUM ? bpg_kernel<float, QD_BP_MIN_SUM> : bpg_kernel<float, QD_BP_PRODUCT_SUM>);
        QD_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem));
        QD_CUDA(cudaMemsetAsync(q.next_shot, 0, sizeof(unsigned long long), st));
        fn<<<grid, d->threads, d->smem, st>>>(q);
        QD_CUDA(cudaGetLastError());
        return 0;
    }
    if (d->v7) {
        Bp7Params q{};
        q.m = d->cl7 ? d->cl7_plan.m_max : g.m; q.n = g.n; q.nwm = g.nwm; q.nwn = g.nwn;
        q.m32 = (q.m + 31) / 32 * 32;
        q.nrec = d->v7_nrec;
        q.lines = d->v7_lines;
        if (d->cl7) {
            q.cl_C = d->cl7_C;
            q.cl_row0 = d->d7c_row0.as<int32_t>(); q.cl_rows = d->d7c_rows.as<int32_t>(); q.cl_nexp = d->d7c_nexp.as<int32_t>();
            q.cl_export = d->d7c_export.as<uint2>(); q.cl_exp_stride = std::max(d->cl7_plan.nexp_max, 1);
        }
        {
            const Bp7SmemLayout SL = bp7_smem_layout(q.m, d->v7_nrec, g.nwn, g.nwm, d->v7_lines, d->rd, d->vmax, d->slots,
                                                     d->precision == QD_PREC_F64 ? 8 : 4, d->v7_staged != 0);
            q.off_flags = (int)SL.flags; q.off_table = (int)SL.table; q.off_out = (int)SL.out; q.off_ctl = (int)SL.ctl;
            q.off_syn = (int)SL.syn; q.off_cb = (int)SL.cb;
        }
        q.row_table = d->d7_row_table.as<uint4>();
        q.max_iter = d->max_iter;
        q.staged = d->v7_staged;
        q.uniform_prior = d->v7_uniform;
        q.prior0 = d->v7_prior0;
        q.alpha_fixed = d->alpha_fixed;
        q.llr_mode = d->llr_mode;
        q.table = d->d7_table.as<uint4>();
        q.chunk_deg = d->d7_chunk_deg.as<uint8_t>();
        q.prior_rec = d->d7_prior_rec.p;
        q.col_id = d->d7_col_id.as<uint16_t>();
        q.const_bits = d->d7_const_bits.as<uint32_t>();
        q.synd = d_synd; q.synd_stride = synd_stride; q.shots = shots;
        q.corr = d_corr; q.corr_stride = corr_stride;
        q.conv = d_converged; q.iters = d_iters; q.llr = d_llr;
        q.err = d_err; q.err_stride = err_stride;
        q.lzT = d_err ? logicals->d_lzT.as<uint32_t>() : nullptr;
        q.k = d_err ? logicals->k : 0;
        q.fail = d_fail;
        q.counters = (long long *)d_counters;
        q.next_shot = d->d_next.as<unsigned long long>() + queue;
        bp7_kernel_fn fn7 = bp7_kernel_for(d, d->slots, d->threads);
        if (!fn7) return fail(QD_EUNSUP, "no row-regular BP kernel compiled for this geometry");
        QD_CUDA(cudaFuncSetAttribute(fn7, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem));
        QD_CUDA(cudaMemsetAsync(q.next_shot, 0, sizeof(unsigned long long), st));
        const long long want7 = (shots + d->slots - 1) / d->slots;
        if (d->cl7) {
            const long long nclusters = std::min<long long>(d->cl7_clusters, want7);
            cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[1];
            bp7c_launch_config(d, d->cl7_C, d->threads, d->smem, (unsigned)(nclusters * d->cl7_C), st, cfg, attr);
            QD_CUDA(cudaLaunchKernelEx(&cfg, fn7, q));
            return 0;
        }
        const unsigned grid7 = (unsigned)std::min<long long>((long long)d->g->sm_count * d->ctas_per_sm, want7);
        fn7<<<grid7, d->threads, d->smem, st>>>(q);
        QD_CUDA(cudaGetLastError());
        return 0;
    }
    BpParams p{};
    p.g = g;
    p.prior = d->d_prior.p; p.alpha = d->d_alpha.p; p.vidx = d->d_vidx.p;
    p.chunk_deg = d->d_chunk_deg.as<uint8_t>();
    p.max_iter = d->max_iter; p.slots = d->slots; p.groups = d->groups; p.stage_idx = d->stage_idx;
    p.synd = d_synd; p.synd_stride = synd_stride; p.shots = shots;
    p.corr = d_corr; p.corr_stride = corr_stride;
    p.conv = d_converged; p.iters = d_iters; p.llr = d_llr;
    p.llr_mode = d->llr_mode;
    p.err = d_err; p.err_stride = err_stride;
    p.lzT = d_err ? logicals->d_lzT.as<uint32_t>() : nullptr;
    p.k = d_err ? logicals->k : 0;
    p.fail = d_fail;
    p.counters = (long long *)d_counters;
    p.next_shot = d->d_next.as<unsigned long long>() + queue;
    bp_kernel_fn fn = bp_kernel_for(d);
    QD_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem));
    QD_CUDA(cudaMemsetAsync(p.next_shot, 0, sizeof(unsigned long long), st));
    long long want = (shots + d->slots - 1) / d->slots;
    unsigned grid = (unsigned)std::min<long long>((long long)d->g->sm_count * d->ctas_per_sm, want);
    fn<<<grid, d->threads, d->smem, st>>>(p);
    QD_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int qd_bp_decode_batch(qd_bp *d, const uint32_t *d_synd, int32_t synd_stride, int64_t shots,
                                  uint32_t *d_corr, int32_t corr_stride, uint8_t *d_converged, int32_t *d_iters,
                                  double *d_llr, const uint32_t *d_err, int32_t err_stride, qd_logicals *logicals,
                                  uint8_t *d_fail, int64_t *d_counters, void *stream)
{
    return bp_launch(d, d_synd, synd_stride, shots, d_corr, corr_stride, d_converged, d_iters, d_llr, d_err, err_stride,
                     logicals, d_fail, d_counters, stream, 0);
}

// ---------------------------------------------------------------------------------------------
// host-buffer pipeline shared by BP and SSF: chunks alternate between two streams so that the
// H2D copy / decode / D2H copy of one chunk overlap the staging of the next.
// ---------------------------------------------------------------------------------------------
template <typename H, typename Launch>
static int host_pipeline(H *d, int device, int nwm, int nwn, const uint32_t *h_synd, int synd_stride, int64_t shots,
                         uint32_t *h_corr, int corr_stride, uint8_t *h_conv, int32_t *h_it, Launch launch)
{
    int rc = use_device(device);
    if (rc) return rc;
    for (int i = 0; i < 2; ++i) {
        if (!d->streams[i]) QD_CUDA(cudaStreamCreateWithFlags(&d->streams[i], cudaStreamNonBlocking));
        if (!d->done[i]) QD_CUDA(cudaEventCreateWithFlags(&d->done[i], cudaEventDisableTiming));
    }
    const int is = (nwm + 3) & ~3, os = (nwn + 3) & ~3;   // device strides: 16-byte rows
    const int64_t per_shot = 4LL * (is + os) + 5;
    int64_t chunk = std::max<int64_t>(1024, std::min<int64_t>(1 << 18, (64LL << 20) / per_shot));
    if (const char *e = getenv("QD_HOST_CHUNK")) chunk = std::max<int64_t>(1, atoll(e));   // tests: many small chunks in flight
    chunk = std::min(chunk, std::max<int64_t>(shots, 1));
    for (int i = 0; i < 2; ++i) {
        if ((rc = d->d_in[i].ensure((size_t)chunk * is * 4))) return rc;
        if ((rc = d->d_out[i].ensure((size_t)chunk * os * 4))) return rc;
        if ((rc = d->d_conv[i].ensure((size_t)chunk))) return rc;
        if ((rc = d->d_it[i].ensure((size_t)chunk * 4))) return rc;
        if ((rc = d->h_in[i].ensure((size_t)chunk * is * 4))) return rc;
        if ((rc = d->h_out[i].ensure((size_t)chunk * os * 4))) return rc;
        if ((rc = d->h_conv[i].ensure((size_t)chunk))) return rc;
        if ((rc = d->h_it[i].ensure((size_t)chunk * 4))) return rc;
    }
    const int64_t nchunks = (shots + chunk - 1) / chunk;
    auto drain = [&](int64_t c) -> int {   // copy chunk c's results from pinned staging to the caller
        const int b = (int)(c & 1);
        QD_CUDA(cudaEventSynchronize(d->done[b]));
        const int64_t s0 = c * chunk, ns = std::min(chunk, shots - s0);
        const uint32_t *src = d->h_out[b].template as<uint32_t>();
        if (corr_stride == os) {                                  // rows already have the caller's stride
            std::memcpy(h_corr + s0 * corr_stride, src, (size_t)ns * os * 4);
        } else {
            for (int64_t s = 0; s < ns; ++s) {
                uint32_t *dst = h_corr + (s0 + s) * corr_stride;
                std::memcpy(dst, src + s * os, (size_t)nwn * 4);
                for (int w = nwn; w < corr_stride; ++w) dst[w] = 0u;
            }
        }
        if (h_conv) std::memcpy(h_conv + s0, d->h_conv[b].p, (size_t)ns);
        if (h_it) std::memcpy(h_it + s0, d->h_it[b].p, (size_t)ns * 4);
        return 0;
    };
    for (int64_t c = 0; c < nchunks; ++c) {
        const int b = (int)(c & 1);
        if (c >= 2 && (rc = drain(c - 2))) return rc;
        const int64_t s0 = c * chunk, ns = std::min(chunk, shots - s0);
        uint32_t *stage = d->h_in[b].template as<uint32_t>();
        if (synd_stride == is) {                                  // caller already uses the device stride
            std::memcpy(stage, h_synd + s0 * synd_stride, (size_t)ns * is * 4);
        } else {
            for (int64_t s = 0; s < ns; ++s) {
                std::memcpy(stage + s * is, h_synd + (s0 + s) * synd_stride, (size_t)nwm * 4);
                for (int w = nwm; w < is; ++w) stage[s * is + w] = 0u;
            }
        }
        cudaStream_t st = d->streams[b];
        QD_CUDA(cudaMemcpyAsync(d->d_in[b].p, stage, (size_t)ns * is * 4, cudaMemcpyHostToDevice, st));
        if ((rc = launch(d->d_in[b].template as<uint32_t>(), is, ns, d->d_out[b].template as<uint32_t>(), os,
                         d->d_conv[b].template as<uint8_t>(), d->d_it[b].template as<int32_t>(), st)))
            return rc;
        QD_CUDA(cudaMemcpyAsync(d->h_out[b].p, d->d_out[b].p, (size_t)ns * os * 4, cudaMemcpyDeviceToHost, st));
        if (h_conv) QD_CUDA(cudaMemcpyAsync(d->h_conv[b].p, d->d_conv[b].p, (size_t)ns, cudaMemcpyDeviceToHost, st));
        if (h_it) QD_CUDA(cudaMemcpyAsync(d->h_it[b].p, d->d_it[b].p, (size_t)ns * 4, cudaMemcpyDeviceToHost, st));
        QD_CUDA(cudaEventRecord(d->done[b], st));
    }
    for (int64_t c = std::max<int64_t>(0, nchunks - 2); c < nchunks; ++c)
        if ((rc = drain(c))) return rc;
    return 0;
}

extern "C" int qd_bp_decode_batch_host(qd_bp *d, const uint32_t *h_synd, int32_t synd_stride, int64_t shots,
                                       uint32_t *h_corr, int32_t corr_stride, uint8_t *h_converged, int32_t *h_iters)
{
    if (!d || !h_synd || !h_corr || shots < 0) return fail(QD_EINVAL, "bad bp_decode_batch_host arguments");
    const GraphDev &g = d->g->dev;
    if (synd_stride < g.nwm || corr_stride < g.nwn) return fail(QD_EINVAL, "stride too small");
    if (shots == 0) return 0;
    int which = 0;   // the two in-flight chunks use separate work-queue counters
    auto launch = [&](const uint32_t *in, int is, int64_t ns, uint32_t *out, int os, uint8_t *conv, int32_t *it,
                      cudaStream_t st) -> int {
        int rc = bp_launch(d, in, is, ns, out, os, conv, it, nullptr, nullptr, 0, nullptr, nullptr, nullptr, st, which);
        which ^= 1;
        return rc;
    };
    return host_pipeline(d, d->g->device, g.nwm, g.nwn, h_synd, synd_stride, shots, h_corr, corr_stride, h_converged,
                         h_iters, launch);
}

extern "C" int qd_bp_decode_one(qd_bp *d, const uint8_t *h_syndrome, uint8_t *h_decoding, int32_t *converged,
                                int32_t *iters, double *h_llr)
{
    if (!d || !h_syndrome || !h_decoding) return fail(QD_EINVAL, "bad bp_decode_one arguments");
    const GraphDev &g = d->g->dev;
    int rc = use_device(d->g->device);
    if (rc) return rc;
    std::vector<uint32_t> s(g.nwm, 0u), c(g.nwn, 0u);
    for (int i = 0; i < g.m; ++i)
        if (h_syndrome[i] & 1) s[i >> 5] |= 1u << (i & 31);
    DevBuf ds, dc, dv, di, dl;
    if ((rc = ds.alloc((size_t)g.nwm * 4 + 4))) return rc;
    if ((rc = dc.alloc((size_t)g.nwn * 4))) return rc;
    if ((rc = dv.alloc(4))) return rc;
    if ((rc = di.alloc(4))) return rc;
    if (h_llr && (rc = dl.alloc((size_t)g.n * 8))) return rc;
    QD_CUDA(cudaMemcpy(ds.p, s.data(), (size_t)g.nwm * 4, cudaMemcpyHostToDevice));
    rc = qd_bp_decode_batch(d, ds.as<uint32_t>(), g.nwm, 1, dc.as<uint32_t>(), g.nwn, dv.as<uint8_t>(), di.as<int32_t>(),
                            h_llr ? dl.as<double>() : nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr);
    if (rc) return rc;
    QD_CUDA(cudaDeviceSynchronize());
    uint8_t cv = 0;
    int32_t it = 0;
    QD_CUDA(cudaMemcpy(c.data(), dc.p, (size_t)g.nwn * 4, cudaMemcpyDeviceToHost));
    QD_CUDA(cudaMemcpy(&cv, dv.p, 1, cudaMemcpyDeviceToHost));
    QD_CUDA(cudaMemcpy(&it, di.p, 4, cudaMemcpyDeviceToHost));
    if (h_llr) QD_CUDA(cudaMemcpy(h_llr, dl.p, (size_t)g.n * 8, cudaMemcpyDeviceToHost));
    for (int j = 0; j < g.n; ++j) h_decoding[j] = (uint8_t)((c[j >> 5] >> (j & 31)) & 1u);
    if (converged) *converged = cv;
    if (iters) *iters = it;
    return 0;
}

// =============================================================================================
// circuit-level sampler (Pauli frames)
// =============================================================================================
extern "C" int qd_circuit_create(int32_t num_ops, const int32_t *h_ops, int32_t num_thresholds, const uint32_t *h_thresholds,
                                 int32_t num_qubits, int32_t num_measurements, int device, qd_circuit **out)
{
    if (!out) return fail(QD_EINVAL, "out is NULL");
    *out = nullptr;
    if (num_ops < 0 || (num_ops && !h_ops) || num_thresholds < 1 || !h_thresholds || num_qubits <= 0 || num_measurements < 0)
        return fail(QD_EINVAL, "bad circuit_create arguments");
    if (h_thresholds[0] != 0u) return fail(QD_EINVAL, "threshold 0 must be 0 (the noiseless entry)");
    int meas = 0;
    for (int i = 0; i < num_ops; ++i) {
        const int code = h_ops[4 * i] & 0xff, ti = h_ops[4 * i] >> 8, q0 = h_ops[4 * i + 1], q1 = h_ops[4 * i + 2];
        if (code > FR_ZERR || ti < 0 || ti >= num_thresholds || q0 < 0 || q0 >= num_qubits) return fail(QD_EINVAL, "bad circuit op");
        if ((code == FR_CX || code == FR_CZ || code == FR_DEP2) && (q1 < 0 || q1 >= num_qubits || q1 == q0)) return fail(QD_EINVAL, "bad two-qubit op");
        meas += code >= FR_MZ && code <= FR_MRX;
    }
    if (meas != num_measurements) return fail(QD_EINVAL, "num_measurements does not match the program");
    int rc = use_device(device);
    if (rc) return rc;
    std::unique_ptr<qd_circuit> c(new qd_circuit);
    c->device = device; c->nops = num_ops; c->nq = num_qubits; c->nmeas = num_measurements; c->nthr = num_thresholds;
    std::vector<int32_t> ops(h_ops, h_ops + (size_t)4 * num_ops);
    std::vector<uint32_t> thr(h_thresholds, h_thresholds + num_thresholds);
    if ((rc = upload(c->d_ops, ops))) return rc;
    if ((rc = upload(c->d_thr, thr))) return rc;
    *out = c.release();
    return 0;
}

extern "C" int qd_circuit_destroy(qd_circuit *c)
{
    if (c) { cudaSetDevice(c->device); delete c; }
    return 0;
}

static int circuit_launch(qd_circuit *c, uint64_t seed, int64_t shot0, int64_t shots, uint32_t *d_records, int32_t stride,
                          const int2 *d_forced, void *stream);

extern "C" int qd_circuit_sample(qd_circuit *c, uint64_t seed, int64_t shot0, int64_t shots, uint32_t *d_records,
                                 int32_t stride, void *stream)
{
    return circuit_launch(c, seed, shot0, shots, d_records, stride, nullptr, stream);
}

extern "C" int qd_circuit_propagate_faults(qd_circuit *c, int64_t num_faults, const int32_t *d_site_and_pauli,
                                           uint32_t *d_records, int32_t stride, void *stream)
{
    if (!d_site_and_pauli) return fail(QD_EINVAL, "bad circuit_propagate_faults arguments");
    return circuit_launch(c, 0, 0, num_faults, d_records, stride, reinterpret_cast<const int2 *>(d_site_and_pauli), stream);
}

static int circuit_launch(qd_circuit *c, uint64_t seed, int64_t shot0, int64_t shots, uint32_t *d_records, int32_t stride,
                          const int2 *d_forced, void *stream)
{
    if (!c || !d_records || shots < 0) return fail(QD_EINVAL, "bad circuit_sample arguments");
    if (stride < words_for(c->nmeas)) return fail(QD_EINVAL, "stride smaller than ceil(measurements/32)");
    if (shots == 0) return 0;
    int rc = use_device(c->device);
    if (rc) return rc;
    int optin = 0;
    QD_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
    const int nqw = words_for(c->nq);
    int threads = 128;
    while (threads > 32 && (size_t)nqw * threads * 8 > (size_t)optin / 2) threads /= 2;
    const size_t smem = (size_t)nqw * threads * 8;
    if (smem > (size_t)optin) return fail(QD_EUNSUP, "circuit has too many qubits for the shared-memory frame sampler");
    QD_CUDA(cudaFuncSetAttribute(frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FrameParams p{};
    p.ops = c->d_ops.as<int4>(); p.nops = c->nops; p.thr = c->d_thr.as<uint32_t>();
    p.nqw = nqw; p.nmeas = c->nmeas;
    p.k0 = (uint32_t)seed; p.k1 = (uint32_t)(seed >> 32);
    p.shot0 = shot0; p.shots = shots; p.rec = d_records; p.stride = stride;
    p.forced = d_forced;
    frame_kernel<<<grid_for(shots, threads), threads, smem, (cudaStream_t)stream>>>(p);
    QD_CUDA(cudaGetLastError());
    return 0;
}

// =============================================================================================
// OSD-0 post-processing (BP + OSD)
// =============================================================================================
static int gf2_rank_of_csr(int m, int n, const std::vector<int32_t> &row_ptr, const std::vector<int32_t> &col_idx)
{
    const int W = (n + 63) / 64;
    std::vector<uint64_t> a((size_t)std::max(m, 1) * W, 0);
    for (int i = 0; i < m; ++i)
        for (int e = row_ptr[i]; e < row_ptr[i + 1]; ++e) a[(size_t)i * W + (col_idx[e] >> 6)] |= 1ULL << (col_idx[e] & 63);
    int rank = 0;
    for (int c = 0; c < n && rank < m; ++c) {
        int piv = -1;
        for (int i = rank; i < m; ++i)
            if ((a[(size_t)i * W + (c >> 6)] >> (c & 63)) & 1ULL) { piv = i; break; }
        if (piv < 0) continue;
        if (piv != rank)
            for (int w = 0; w < W; ++w) std::swap(a[(size_t)piv * W + w], a[(size_t)rank * W + w]);
        for (int i = rank + 1; i < m; ++i)
            if ((a[(size_t)i * W + (c >> 6)] >> (c & 63)) & 1ULL)
                for (int w = c >> 6; w < W; ++w) a[(size_t)i * W + w] ^= a[(size_t)rank * W + w];
        ++rank;
    }
    return rank;
}

extern "C" int qd_osd_create(qd_graph *g, int32_t method, int32_t order, qd_osd **out)
{
    if (!out) return fail(QD_EINVAL, "out is NULL");
    *out = nullptr;
    if (!g) return fail(QD_EINVAL, "graph is NULL");
    if (method < 0 || method > 2 || order < 0) return fail(QD_EINVAL, "OSD method must be 0 (osd0), 1 (osd_e) or 2 (osd_cs), order >= 0");
    if (method == 1 && order > 16) return fail(QD_EUNSUP, "osd_e: search depth > 16 (65535 patterns) not supported");
    if (method == 2 && order > 24) return fail(QD_EUNSUP, "osd_cs: search depth > 24 not supported");
    int rc = use_device(g->device);
    if (rc) return rc;
    const GraphDev &gd = g->dev;
    if (gd.n > 16000 || gd.m > 4096) return fail(QD_EUNSUP, "OSD: more than 16000 columns or 4096 rows");
    std::unique_ptr<qd_osd> o(new qd_osd);
    o->g = g; o->device = g->device;
    o->method = order > 0 ? method : 0; o->order = method ? order : 0;
    o->W = (gd.n + 1 + 63) / 64;
    while (o->W % 4 != 2) ++o->W;           // rows of an odd number of 16-byte units: aligned, bank-conflict free 128-bit XORs
    o->P = 32;                              // (the ranking network works on runs of 32)
    while (o->P < gd.n) o->P <<= 1;
    // the (B, Y) pairs of the elimination (8 bytes per row) reuse the P * 8 bytes of the sort keys
    if (gd.m > o->P) return fail(QD_EUNSUP, "OSD: more rows than the power of two >= the number of columns");
    if (o->method && (size_t)gd.m * 4 + (size_t)gd.n * 2 + 8 * 33 > (size_t)o->P * 8)
        return fail(QD_EUNSUP, "higher-order OSD: scratch for this shape does not fit");
    o->smem = osd_smem_layout(gd.m, o->W, o->P, gd.nwm, gd.nwn, true).total;
    if (o->smem > (size_t)g->smem_optin || getenv("QD_OSD_GLOBAL")) {      // (the variable forces the scratch path for tests)
        o->rows_global = 1;
        o->smem = osd_smem_layout(gd.m, o->W, o->P, gd.nwm, gd.nwn, false).total;
        if (o->smem > (size_t)g->smem_optin)
            return fail(QD_EUNSUP, "OSD: ranking keys and lookup tables (" + std::to_string(o->smem) + " B) do not fit shared memory");
    }
    o->rank = gf2_rank_of_csr(gd.m, gd.n, g->row_ptr, g->col_idx);
    if ((rc = o->d_next.alloc(sizeof(unsigned long long) * 2))) return rc;
    *out = o.release();
    return 0;
}

extern "C" int qd_osd_destroy(qd_osd *o)
{
    if (o) { cudaSetDevice(o->device); delete o; }
    return 0;
}

extern "C" int qd_osd_rank(qd_osd *o)
{
    return o ? o->rank : fail(QD_EINVAL, "NULL handle");
}

static int osd_launch(qd_osd *o, int negate_key, const uint32_t *d_synd, int32_t synd_stride, int64_t shots,
                      const double *d_soft, const uint8_t *d_conv, uint32_t *d_corr, int32_t corr_stride,
                      const uint32_t *d_err, int32_t err_stride, qd_logicals *logicals, uint8_t *d_fail,
                      int64_t *d_counters, cudaStream_t st, int queue)
{
    const GraphDev &g = o->g->dev;
    OsdParams p{};
    p.m = g.m; p.n = g.n; p.nwm = g.nwm; p.nwn = g.nwn;
    p.W = o->W; p.P = o->P; p.rank = o->rank; p.negate_key = negate_key;
    p.method = o->method; p.order = o->order;
    p.row_ptr = g.row_ptr; p.col_idx = g.col_idx;
    p.synd = d_synd; p.synd_stride = synd_stride;
    p.soft = d_soft; p.conv = d_conv; p.shots = shots;
    p.corr = d_corr; p.corr_stride = corr_stride;
    p.err = d_err; p.err_stride = err_stride;
    p.lzT = d_err ? logicals->d_lzT.as<uint32_t>() : nullptr;
    p.k = d_err ? logicals->k : 0;
    p.fail = d_fail;
    p.counters = (long long *)d_counters;
    p.next_shot = o->d_next.as<unsigned long long>() + queue;
    // one row per thread up to 1024 rows (measured faster than two at m = 768), two up to 2048 (more rows never
    // fit shared memory: qd_osd_create refuses them)
    int rpt = g.m > 2048 ? 4 : g.m > 1024 ? 2 : 1;
    if (const char *e = getenv("QD_OSD_RPT")) rpt = std::max(rpt, atoi(e) == 4 ? 4 : atoi(e) == 2 ? 2 : 1);
    const int threads = std::min(1024, std::max(128, ((g.m + rpt - 1) / rpt + 31) & ~31));
    auto kern = rpt == 4 ? osd0_kernel<4> : rpt == 2 ? osd0_kernel<2> : osd0_kernel<1>;
    QD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)o->smem));
    QD_CUDA(cudaMemsetAsync(p.next_shot, 0, sizeof(unsigned long long), st));
    int per_sm = 0;
    QD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, o->smem));
    per_sm = std::max(1, per_sm);
    long long ctas = (long long)o->g->sm_count * per_sm;
    if (o->rows_global) {
        // keep the matrices of the CTAs in flight L2-resident: at most ~96 MB of them, at least 8 CTAs
        const size_t per = (size_t)g.m * o->W * 8;
        ctas = std::max<long long>(8, std::min<long long>(ctas, (long long)((96ull << 20) / per)));
        if (int rc2 = o->d_rows.ensure(per * (size_t)ctas)) return rc2;
        p.rows_global = o->d_rows.as<unsigned long long>();
    }
    const unsigned grid = (unsigned)std::min<long long>(ctas, shots);
    kern<<<grid, threads, o->smem, st>>>(p);
    QD_CUDA(cudaGetLastError());
    return 0;
}

// BP on the whole batch (soft output of the last iteration captured), then OSD-0 on the shots BP left unconverged
static int bposd_launch(qd_bp *d, qd_osd *o, const uint32_t *d_synd, int32_t synd_stride, int64_t shots, uint32_t *d_corr,
                        int32_t corr_stride, uint8_t *d_converged, int32_t *d_iters, const uint32_t *d_err,
                        int32_t err_stride, qd_logicals *logicals, uint8_t *d_fail, int64_t *d_counters, void *stream,
                        int queue)
{
    if (!d || !o) return fail(QD_EINVAL, "NULL handle");
    if (o->g != d->g) return fail(QD_EINVAL, "OSD handle was built on a different graph");
    if (shots <= 0) return shots < 0 ? fail(QD_EINVAL, "shots < 0") : 0;
    int rc = use_device(d->device);
    if (rc) return rc;
    const GraphDev &g = d->g->dev;
    // the soft outputs take 8n bytes per shot: decode in slices of at most 1 GiB of them
    const int64_t slice = std::max<int64_t>(1, std::min<int64_t>(shots, (1LL << 30) / (8LL * g.n)));
    if ((rc = o->d_soft.ensure((size_t)slice * g.n * 8))) return rc;
    if (!d_converged && (rc = o->d_conv.ensure((size_t)slice))) return rc;
    // the OSD epilogue corrects the failure count by (new flag - BP's flag): it needs BP's flags even when the
    // caller does not want them
    const bool own_fail = d_err && d_counters && !d_fail;
    if (own_fail && (rc = o->d_failbuf.ensure((size_t)slice))) return rc;
    const int saved = d->llr_mode;
    d->llr_mode = 1;
    for (int64_t s0 = 0; s0 < shots && !rc; s0 += slice) {
        const int64_t ns = std::min(slice, shots - s0);
        uint8_t *conv = d_converged ? d_converged + s0 : o->d_conv.as<uint8_t>();
        uint8_t *flags = own_fail ? o->d_failbuf.as<uint8_t>() : (d_fail ? d_fail + s0 : nullptr);
        rc = bp_launch(d, d_synd + s0 * synd_stride, synd_stride, ns, d_corr + s0 * corr_stride, corr_stride, conv,
                       d_iters ? d_iters + s0 : nullptr, o->d_soft.as<double>(), d_err ? d_err + s0 * err_stride : nullptr,
                       err_stride, logicals, flags, d_counters, stream, queue);
        if (!rc)
            rc = osd_launch(o, d->method == QD_BP_PRODUCT_SUM, d_synd + s0 * synd_stride, synd_stride, ns,
                            o->d_soft.as<double>(), conv, d_corr + s0 * corr_stride, corr_stride,
                            d_err ? d_err + s0 * err_stride : nullptr, err_stride, logicals, flags,
                            d_counters, (cudaStream_t)stream, queue);
    }
    d->llr_mode = saved;
    return rc;
}

extern "C" int qd_bposd_decode_batch(qd_bp *d, qd_osd *o, const uint32_t *d_synd, int32_t synd_stride, int64_t shots,
                                     uint32_t *d_corr, int32_t corr_stride, uint8_t *d_converged, int32_t *d_iters,
                                     const uint32_t *d_err, int32_t err_stride, qd_logicals *logicals, uint8_t *d_fail,
                                     int64_t *d_counters, void *stream)
{
    if (!d_synd || !d_corr) return fail(QD_EINVAL, "bad bposd_decode_batch arguments");
    if (d && (synd_stride < d->g->dev.nwm || corr_stride < d->g->dev.nwn)) return fail(QD_EINVAL, "stride too small");
    if (d_err && (!logicals || (d && err_stride < d->g->dev.nwn))) return fail(QD_EINVAL, "fused epilogue needs logicals and err_stride >= ceil(n/32)");
    if (d_fail && !d_err) return fail(QD_EINVAL, "d_fail needs d_err");
    return bposd_launch(d, o, d_synd, synd_stride, shots, d_corr, corr_stride, d_converged, d_iters, d_err, err_stride,
                        logicals, d_fail, d_counters, stream, 0);
}

extern "C" int qd_bposd_decode_batch_host(qd_bp *d, qd_osd *o, const uint32_t *h_synd, int32_t synd_stride, int64_t shots,
                                          uint32_t *h_corr, int32_t corr_stride, uint8_t *h_converged, int32_t *h_iters)
{
    if (!d || !o || !h_synd || !h_corr || shots < 0) return fail(QD_EINVAL, "bad bposd_decode_batch_host arguments");
    const GraphDev &g = d->g->dev;
    if (synd_stride < g.nwm || corr_stride < g.nwn) return fail(QD_EINVAL, "stride too small");
    if (shots == 0) return 0;
    auto launch = [&](const uint32_t *in, int is, int64_t ns, uint32_t *out, int os, uint8_t *conv, int32_t *it,
                      cudaStream_t st) -> int {
        // the OSD scratch (soft outputs) is per handle: run the slices on one stream at a time
        QD_CUDA(cudaDeviceSynchronize());
        return bposd_launch(d, o, in, is, ns, out, os, conv, it, nullptr, 0, nullptr, nullptr, nullptr, st, 0);
    };
    return host_pipeline(d, d->device, g.nwm, g.nwn, h_synd, synd_stride, shots, h_corr, corr_stride, h_converged,
                         h_iters, launch);
}

// =============================================================================================
// SSF
// =============================================================================================
extern "C" int qd_ssf_create(qd_graph *hz, int32_t num_generators, const int32_t *h_gen_row_ptr,
                             const int32_t *h_gen_col_idx, int32_t max_rounds, qd_ssf **out)
{
    if (!out) return fail(QD_EINVAL, "out is NULL");
    *out = nullptr;
    if (!hz || num_generators <= 0 || !h_gen_row_ptr || !h_gen_col_idx || max_rounds < 0)
        return fail(QD_EINVAL, "bad ssf_create arguments");
    int rc = use_device(hz->device);
    if (rc) return rc;
    const GraphDev &g = hz->dev;
    const int mg = num_generators;
    const int gnnz = h_gen_row_ptr[mg];
    std::vector<int32_t> gen_ptr(h_gen_row_ptr, h_gen_row_ptr + mg + 1), gen_idx(h_gen_col_idx, h_gen_col_idx + gnnz);
    std::vector<int32_t> win_ptr(mg + 1, 0), win_chk;
    std::vector<unsigned long long> qmask(std::max(gnnz, 1), 0ULL);
    int max_w = 0, max_window = 0;
    for (int gi = 0; gi < mg; ++gi) {
        const int lo = gen_ptr[gi], w = gen_ptr[gi + 1] - lo;
        if (w <= 0 || w > SSF_MAX_W) return fail(QD_EUNSUP, "SSF: generator weight must be in [1, 12]");
        if (mg > 65535) return fail(QD_EUNSUP, "SSF: more than 65535 generators");
        max_w = std::max(max_w, w);
        std::vector<int32_t> win;
        for (int t = lo; t < lo + w; ++t) {
            const int q = gen_idx[t];
            if (q < 0 || q >= g.n) return fail(QD_EINVAL, "SSF: generator qubit out of range");
            for (int u = hz->col_ptr[q]; u < hz->col_ptr[q + 1]; ++u) {
                const int c = (int)hz->csc_row[u];
                int pos = -1;
                for (size_t z = 0; z < win.size(); ++z) if (win[z] == c) { pos = (int)z; break; }
                if (pos < 0) {
                    if (win.size() == 64) return fail(QD_EUNSUP, "SSF: check window of a generator exceeds 64");
                    pos = (int)win.size();
                    win.push_back(c);
                }
                qmask[t] ^= 1ULL << pos;
            }
        }
        win_chk.insert(win_chk.end(), win.begin(), win.end());
        win_ptr[gi + 1] = (int)win_chk.size();
        max_window = std::max(max_window, win_ptr[gi + 1] - win_ptr[gi]);
    }
    // check -> generators whose window contains it (what a flip invalidates)
    std::vector<int32_t> chk_ptr(g.m + 1, 0), chk_gen(win_chk.size());
    for (int32_t c : win_chk) chk_ptr[c + 1]++;
    for (int c = 0; c < g.m; ++c) chk_ptr[c + 1] += chk_ptr[c];
    {
        std::vector<int32_t> fill(chk_ptr.begin(), chk_ptr.end() - 1);
        for (int gi = 0; gi < mg; ++gi)
            for (int z = win_ptr[gi]; z < win_ptr[gi + 1]; ++z) chk_gen[fill[win_chk[z]]++] = gi;
    }
    std::unique_ptr<qd_ssf> d(new qd_ssf);
    d->g = hz;
    d->device = hz->device;
    d->max_rounds = max_rounds;
    if ((rc = upload(d->d_gen_ptr, gen_ptr))) return rc;
    if ((rc = upload(d->d_gen_idx, gen_idx))) return rc;
    if ((rc = upload(d->d_win_ptr, win_ptr))) return rc;
    if ((rc = upload(d->d_win_chk, win_chk))) return rc;
    if ((rc = upload(d->d_qmask, qmask))) return rc;
    if ((rc = upload(d->d_chk_ptr, chk_ptr))) return rc;
    if ((rc = upload(d->d_chk_gen, chk_gen))) return rc;
    if ((rc = d->d_next.alloc(sizeof(unsigned long long) * 2))) return rc;
    SsfDev &s = d->dev;
    s.m = g.m; s.n = g.n; s.nwm = g.nwm; s.nwn = g.nwn; s.mg = mg; s.max_w = max_w;
    s.gen_ptr = d->d_gen_ptr.as<int32_t>(); s.gen_idx = d->d_gen_idx.as<int32_t>();
    s.win_ptr = d->d_win_ptr.as<int32_t>(); s.win_chk = d->d_win_chk.as<int32_t>();
    s.qmask = d->d_qmask.as<unsigned long long>();
    s.chk_ptr = d->d_chk_ptr.as<int32_t>(); s.chk_gen = d->d_chk_gen.as<int32_t>();
    d->threads = max_w <= 7 ? 64 : 128;    // few chunks per generator: smaller CTAs, more of them per SM (measured: n=1600 +6 %)
    if (const char *e = getenv("QD_SSF_THREADS")) { const int t = atoi(e); if (t >= 32 && t <= 256 && t % 32 == 0) d->threads = t; }
    d->max_window = max_window;
    d->smem = ssf_smem_bytes(g.nwn, g.nwm, mg, d->threads);
    *out = d.release();
    return 0;
}

extern "C" int qd_ssf_destroy(qd_ssf *d)
{
    if (d) { cudaSetDevice(d->device); delete d; }
    return 0;
}

static int ssf_launch(qd_ssf *d, const uint32_t *d_synd, int32_t synd_stride, int64_t shots, uint32_t *d_corr,
                      int32_t corr_stride, uint8_t *d_converged, int32_t *d_rounds, const uint32_t *d_err,
                      int32_t err_stride, qd_logicals *logicals, uint8_t *d_fail, int64_t *d_counters, void *stream,
                      int queue)
{
    if (!d || !d_synd || !d_corr || shots < 0) return fail(QD_EINVAL, "bad ssf_decode_batch arguments");
    const GraphDev &g = d->g->dev;
    if (synd_stride < g.nwm || corr_stride < g.nwn) return fail(QD_EINVAL, "stride too small");
    if (d_err && (!logicals || err_stride < g.nwn)) return fail(QD_EINVAL, "fused epilogue needs logicals and err_stride >= ceil(n/32)");
    if (logicals && logicals->n != g.n) return fail(QD_EINVAL, "logicals act on a different number of qubits");
    if (shots == 0) return 0;
    int rc = use_device(d->g->device);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    SsfParams p{};
    p.s = d->dev;
    p.max_rounds = d->max_rounds;
    p.synd = d_synd; p.synd_stride = synd_stride; p.shots = shots;
    p.corr = d_corr; p.corr_stride = corr_stride;
    p.conv = d_converged; p.rounds = d_rounds;
    p.err = d_err; p.err_stride = err_stride;
    p.lzT = d_err ? logicals->d_lzT.as<uint32_t>() : nullptr;
    p.k = d_err ? logicals->k : 0;
    p.fail = d_fail;
    p.counters = (long long *)d_counters;
    p.next_shot = d->d_next.as<unsigned long long>() + queue;
    auto kern = d->max_window <= 32 ? ssf_kernel<uint32_t> : ssf_kernel<unsigned long long>;
    QD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem));
    QD_CUDA(cudaMemsetAsync(p.next_shot, 0, sizeof(unsigned long long), st));
    int per_sm = 0;
    QD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, d->threads, d->smem));
    per_sm = std::max(1, per_sm);
    unsigned grid = (unsigned)std::min<long long>((long long)d->g->sm_count * per_sm, shots);
    kern<<<grid, d->threads, d->smem, st>>>(p);
    QD_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int qd_ssf_decode_batch(qd_ssf *d, const uint32_t *d_synd, int32_t synd_stride, int64_t shots,
                                   uint32_t *d_corr, int32_t corr_stride, uint8_t *d_converged, int32_t *d_rounds,
                                   const uint32_t *d_err, int32_t err_stride, qd_logicals *logicals, uint8_t *d_fail,
                                   int64_t *d_counters, void *stream)
{
    return ssf_launch(d, d_synd, synd_stride, shots, d_corr, corr_stride, d_converged, d_rounds, d_err, err_stride,
                      logicals, d_fail, d_counters, stream, 0);
}

extern "C" int qd_ssf_decode_batch_host(qd_ssf *d, const uint32_t *h_synd, int32_t synd_stride, int64_t shots,
                                        uint32_t *h_corr, int32_t corr_stride, uint8_t *h_converged, int32_t *h_rounds)
{
    if (!d || !h_synd || !h_corr || shots < 0) return fail(QD_EINVAL, "bad ssf_decode_batch_host arguments");
    const GraphDev &g = d->g->dev;
    if (synd_stride < g.nwm || corr_stride < g.nwn) return fail(QD_EINVAL, "stride too small");
    if (shots == 0) return 0;
    int which = 0;
    auto launch = [&](const uint32_t *in, int is, int64_t ns, uint32_t *out, int os, uint8_t *conv, int32_t *it,
                      cudaStream_t st) -> int {
        int rc = ssf_launch(d, in, is, ns, out, os, conv, it, nullptr, 0, nullptr, nullptr, nullptr, st, which);
        which ^= 1;
        return rc;
    };
    return host_pipeline(d, d->g->device, g.nwm, g.nwn, h_synd, synd_stride, shots, h_corr, corr_stride, h_converged,
                         h_rounds, launch);
}

// =============================================================================================
// code-capacity experiment
// =============================================================================================
extern "C" int qd_code_capacity_run(qd_graph *hz, qd_bp *bp, qd_ssf *ssf, qd_osd *osd, qd_sampler *sampler, qd_logicals *logicals,
                                    uint64_t seed, int64_t shot0, int64_t shots, int64_t *h_counters, uint32_t *h_corr,
                                    int32_t corr_stride, uint8_t *h_fail, uint8_t *h_converged, int32_t *h_iters)
{
    if (!hz || !sampler || !logicals || !h_counters || shots < 0) return fail(QD_EINVAL, "bad code_capacity_run arguments");
    if ((bp == nullptr) == (ssf == nullptr)) return fail(QD_EINVAL, "exactly one of bp / ssf must be given");
    const GraphDev &g = hz->dev;
    if (sampler->nbits != g.n || logicals->n != g.n) return fail(QD_EINVAL, "sampler / logicals size mismatch");
    if ((bp && bp->g != hz) || (ssf && ssf->g != hz) || (osd && osd->g != hz)) return fail(QD_EINVAL, "decoder was built on a different graph");
    if (osd && !bp) return fail(QD_EINVAL, "OSD post-processing needs a BP decoder");
    if (h_corr && corr_stride < g.nwn) return fail(QD_EINVAL, "corr_stride too small");
    int rc = use_device(hz->device);
    if (rc) return rc;
    const int es = (g.nwn + 3) & ~3, ss = (g.nwm + 3) & ~3;
    const int64_t per_shot = 4LL * (2 * es + ss) + 6;
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(shots, std::min<int64_t>(1 << 22, (512LL << 20) / per_shot)));
    DevBuf d_err, d_syn, d_cor, d_fail, d_conv, d_it, d_cnt;
    if ((rc = d_err.alloc((size_t)chunk * es * 4))) return rc;
    if ((rc = d_syn.alloc((size_t)chunk * ss * 4))) return rc;
    if ((rc = d_cor.alloc((size_t)chunk * es * 4))) return rc;
    if ((rc = d_fail.alloc((size_t)chunk))) return rc;
    if ((rc = d_conv.alloc((size_t)chunk))) return rc;
    if ((rc = d_it.alloc((size_t)chunk * 4))) return rc;
    if ((rc = d_cnt.alloc(4 * sizeof(int64_t)))) return rc;
    QD_CUDA(cudaMemset(d_cnt.p, 0, 4 * sizeof(int64_t)));
    std::vector<uint32_t> tmp;
    for (int64_t s0 = 0; s0 < shots; s0 += chunk) {
        const int64_t ns = std::min(chunk, shots - s0);
        if ((rc = qd_sampler_run(sampler, seed, 0u, shot0 + s0, ns, d_err.as<uint32_t>(), es, nullptr))) return rc;
        if ((rc = qd_syndrome_batch(hz, d_err.as<uint32_t>(), es, ns, d_syn.as<uint32_t>(), ss, nullptr))) return rc;
        if (bp && osd)
            rc = qd_bposd_decode_batch(bp, osd, d_syn.as<uint32_t>(), ss, ns, d_cor.as<uint32_t>(), es, d_conv.as<uint8_t>(),
                                       d_it.as<int32_t>(), d_err.as<uint32_t>(), es, logicals, d_fail.as<uint8_t>(),
                                       d_cnt.as<int64_t>(), nullptr);
        else if (bp)
            rc = qd_bp_decode_batch(bp, d_syn.as<uint32_t>(), ss, ns, d_cor.as<uint32_t>(), es, d_conv.as<uint8_t>(),
                                    d_it.as<int32_t>(), nullptr, d_err.as<uint32_t>(), es, logicals, d_fail.as<uint8_t>(),
                                    d_cnt.as<int64_t>(), nullptr);
        else
            rc = qd_ssf_decode_batch(ssf, d_syn.as<uint32_t>(), ss, ns, d_cor.as<uint32_t>(), es, d_conv.as<uint8_t>(),
                                     d_it.as<int32_t>(), d_err.as<uint32_t>(), es, logicals, d_fail.as<uint8_t>(),
                                     d_cnt.as<int64_t>(), nullptr);
        if (rc) return rc;
        if (h_corr) {
            tmp.resize((size_t)ns * es);
            QD_CUDA(cudaMemcpy(tmp.data(), d_cor.p, (size_t)ns * es * 4, cudaMemcpyDeviceToHost));
            for (int64_t s = 0; s < ns; ++s) {
                uint32_t *dst = h_corr + (s0 + s) * corr_stride;
                std::memcpy(dst, tmp.data() + s * es, (size_t)g.nwn * 4);
                for (int w = g.nwn; w < corr_stride; ++w) dst[w] = 0u;
            }
        }
        if (h_fail) QD_CUDA(cudaMemcpy(h_fail + s0, d_fail.p, (size_t)ns, cudaMemcpyDeviceToHost));
        if (h_converged) QD_CUDA(cudaMemcpy(h_converged + s0, d_conv.p, (size_t)ns, cudaMemcpyDeviceToHost));
        if (h_iters) QD_CUDA(cudaMemcpy(h_iters + s0, d_it.p, (size_t)ns * 4, cudaMemcpyDeviceToHost));
    }
    QD_CUDA(cudaMemcpy(h_counters, d_cnt.p, 4 * sizeof(int64_t), cudaMemcpyDeviceToHost));
    return 0;
}
