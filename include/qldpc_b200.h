/*
 * qldpc_b200.h — C ABI of libqldpc_b200.so, the B200 (sm_100a) batched qLDPC decoding engine.
 *
 * This is the drop-in boundary for the decode hot path of qldpc/exp_ldpc.  The reference has no
 * FFI of its own for this path: its seam is the Python duck type of the third-party `ldpc`
 * decoder objects plus `run_simulation`.  Every entry point below names the reference interface
 * it stands in for (paths relative to the reference tree):
 *
 *   qd_bp_create          <- ldpc.bp_decoder(H, channel_probs=|error_rate=, max_iter=, bp_method=,
 *                            ms_scaling_factor=)            python/qldpc/misc/_experiment.py:110-113,137-140
 *   qd_bp_decode_one      <- bp_decoder.decode(syndrome)     python/qldpc/misc/_experiment.py:117,149
 *   qd_bp_decode_batch*   <- the `for i in range(samples)` loop around .decode()
 *                                                            python/qldpc/misc/_experiment.py:200-209
 *   qd_syndrome_batch     <- (checks @ readout) % 2          python/qldpc/misc/_experiment.py:47,58,124
 *                                                            python/qldpc/spacetime_code.py:109
 *   qd_logical_fail_batch <- np.any(GF2(Lz) @ GF2(readout ^ corr) != 0)
 *                                                            python/qldpc/misc/_experiment.py:208-209
 *   qd_sampler_*          <- stim sampler.sample(samples) restricted to the code-capacity circuit
 *                            (rounds = 0)                    python/qldpc/misc/_experiment.py:193-197
 *   qd_circuit_*          <- stim.Circuit(text).compile_sampler().sample(samples) for the circuits of
 *                            build_storage_simulation        python/qldpc/misc/_experiment.py:171-197
 *                                                            python/qldpc/storage_sim.py:38-75,134-185
 *   qd_code_capacity_run  <- run_simulation(samples, ...) for rounds = 0, returning the counts that
 *                            p_sweep sums                    python/qldpc/misc/_experiment.py:154-210
 *                                                            python/qldpc/misc/p_sweep.py:28-32
 *   qd_osd_*, qd_bposd_*  <- ldpc.bposd_decoder(..., osd_method=, osd_order=).decode(syndrome): BP, then
 *                            OSD (osd0 / osd_e / osd_cs) on the shots BP left unconverged
 *                                                            python/qldpc/misc/_experiment.py:23-27,51,59,82,96-100,125
 *   qd_ssf_*              <- (absent from the reference; SURVEY.md Appendix C is the specification)
 *
 * Conventions
 *   - All functions return 0 on success and a negative QD_E* code on failure; the message is
 *     available from qd_last_error() (thread-local).
 *   - Bit vectors are packed shot-major: shot s starts at word s*stride; bit j of a shot is bit
 *     (j & 31) of word (j >> 5).  `stride` is in 32-bit words and must be >= ceil(bits/32);
 *     padding bits must be zero on input and are written as zero on output.
 *   - Pointers named d_* are device pointers on the handle's device; h_* are host pointers.
 *     `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - Handles are not thread-safe; use one handle per host thread / per stream.
 *   - There is no CPU fallback: every entry point fails with QD_ENODEV when no CUDA device is
 *     usable.
 */
#ifndef QLDPC_B200_H
#define QLDPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QD_ABI_VERSION 2

#define QD_EINVAL   (-1)   /* bad argument                       */
#define QD_ENODEV   (-2)   /* no usable CUDA device              */
#define QD_ECUDA    (-3)   /* CUDA runtime error                 */
#define QD_ENOMEM   (-4)   /* allocation failure                 */
#define QD_EUNSUP   (-5)   /* graph geometry not supported       */

#define QD_BP_PRODUCT_SUM 0   /* bp_method "ps"         */
#define QD_BP_MIN_SUM     3   /* bp_method "ms" / "msl" */

#define QD_PREC_F64 64       /* reproduces the fp64 operation order of the reference decoder  */
#define QD_PREC_F32 32       /* faster, not bit-reproducible                                  */

typedef struct qd_graph    qd_graph;
typedef struct qd_logicals qd_logicals;
typedef struct qd_sampler  qd_sampler;
typedef struct qd_bp       qd_bp;
typedef struct qd_ssf      qd_ssf;
typedef struct qd_osd      qd_osd;
typedef struct qd_circuit  qd_circuit;

int         qd_abi_version(void);
const char *qd_last_error(void);
int         qd_device_count(void);
/* multiprocessor count, max opt-in shared memory per block (bytes), clock (kHz) of a device */
int         qd_device_info(int device, int *sm_count, int *smem_per_block_optin, int *clock_khz);

/* ---- Tanner graph: H (m x n) in CSR with sorted column indices, host arrays ---------------- */
int qd_graph_create(int32_t m, int32_t n, int64_t nnz, const int32_t *h_row_ptr,
                    const int32_t *h_col_idx, int device, qd_graph **out);
int qd_graph_destroy(qd_graph *g);

/* ---- logical operators: dense k x n 0/1 bytes (host) -------------------------------------- */
int qd_logicals_create(int32_t k, int32_t n, const uint8_t *h_dense, int device, qd_logicals **out);
int qd_logicals_destroy(qd_logicals *l);

/* ---- Philox4x32-10 Bernoulli sampler: bit j of shot g = philox(key=seed,
 *      ctr=(g_lo, g_hi, j/4, stream_id))[j%4] < h_threshold[j] ------------------------------ */
int qd_sampler_create(int32_t nbits, const uint32_t *h_threshold, int device, qd_sampler **out);
int qd_sampler_destroy(qd_sampler *s);
int qd_sampler_run(qd_sampler *s, uint64_t seed, uint32_t stream_id, int64_t shot0, int64_t shots,
                   uint32_t *d_bits, int32_t stride, void *stream);

/* ---- circuit-level sampler (Pauli frames).  Program: num_ops rows {opcode | threshold << 8, q0, q1, site}
 *      with opcodes 0 reset, 1 CX, 2 CZ, 3 MZ, 4 MX, 5 MRZ, 6 MRX, 7 DEPOLARIZE1, 8 DEPOLARIZE2,
 *      9 X_ERROR, 10 Z_ERROR; h_thresholds[t] = floor(p * 2^32), entry 0 = 0; `site` numbers the noisy
 *      operations (Philox counter).  Records are bit-packed like stim's bit_packed samples:
 *      bit k of shot s = k-th measurement, relative to the noiseless reference sample. ---------- */
int qd_circuit_create(int32_t num_ops, const int32_t *h_ops, int32_t num_thresholds,
                      const uint32_t *h_thresholds, int32_t num_qubits, int32_t num_measurements,
                      int device, qd_circuit **out);
int qd_circuit_destroy(qd_circuit *c);
int qd_circuit_sample(qd_circuit *c, uint64_t seed, int64_t shot0, int64_t shots,
                      uint32_t *d_records, int32_t stride, void *stream);
/* Fault propagation for detector error models (what circuit.detector_error_model() needs,
 * python/qldpc/misc/_experiment.py:174, python/qldpc/spacetime_code.py:122-183): record f = the
 * NOISELESS circuit with exactly one fault, d_site_and_pauli[2f] = noise site, [2f+1] = Pauli code
 * (DEPOLARIZE1: 1 X, 2 Y, 3 Z; DEPOLARIZE2: 1..15 = pa + 4 pb with 0 I, 1 X, 2 Y, 3 Z; ignored
 * for measurement flips and X/Z_ERROR). */
int qd_circuit_propagate_faults(qd_circuit *c, int64_t num_faults, const int32_t *d_site_and_pauli,
                                uint32_t *d_records, int32_t stride, void *stream);

/* ---- H @ v mod 2, batched and bit-packed (python/qldpc/misc/_experiment.py:47,58,124;
 *      python/qldpc/spacetime_code.py:109).  All words of d_synd up to synd_stride are written (padding
 *      words as zeros); d_vec is only read.  Any strides >= the packed row lengths are accepted; rows that
 *      are 16-byte aligned (d_vec 16-byte aligned, vec_stride a multiple of 4 words) take the TMA path. */
int qd_syndrome_batch(qd_graph *g, const uint32_t *d_vec, int32_t vec_stride, int64_t shots,
                      uint32_t *d_synd, int32_t synd_stride, void *stream);

/* ---- d_dst[s] bits [dst_bit_offset, +nbits) (^)= d_src[s] bits [src_bit_offset, +nbits); other bits
 *      of d_dst are left alone (accumulate != 0: XOR into them).  Building block of the multi-round
 *      recipes: correction fold, syndrome history and its differencing, data part of a single-shot
 *      decoding, slices of a measurement record (python/qldpc/spacetime_code.py:31-33,81-84,98-119;
 *      python/qldpc/storage_sim.py:187-196). ----------------------------------------------------- */
int qd_bits_copy_batch(const uint32_t *d_src, int32_t src_stride, int64_t src_bit_offset, int64_t shots,
                       int32_t nbits, uint32_t *d_dst, int32_t dst_stride, int64_t dst_bit_offset,
                       int32_t accumulate, void *stream);

/* ---- fail[s] = any(L @ (a[s] ^ b[s]) mod 2); d_b may be NULL; d_fail_count (nullable) is
 *      incremented by the number of failing shots ------------------------------------------- */
int qd_logical_fail_batch(qd_logicals *l, const uint32_t *d_a, const uint32_t *d_b, int32_t stride,
                          int64_t shots, uint8_t *d_fail, unsigned long long *d_fail_count,
                          void *stream);

/* ---- belief propagation -------------------------------------------------------------------- */
/* h_channel_probs: n priors.  max_iter == 0 -> n.  ms_scaling_factor == 0 -> 1 - 2^-iteration. */
int qd_bp_create(qd_graph *g, const double *h_channel_probs, int32_t max_iter, int32_t method,
                 double ms_scaling_factor, int32_t precision, qd_bp **out);
int qd_bp_destroy(qd_bp *d);
/* launch geometry override for tuning (threads per CTA, shots in flight per CTA, CTAs per SM,
 * warp teams per CTA: the CTA's warps are split into `teams` groups with their own barrier and
 * their own share of the slots); 0 keeps the current value.  Fails if the configuration does
 * not fit shared memory. */
int qd_bp_configure(qd_bp *d, int32_t threads, int32_t slots, int32_t ctas_per_sm, int32_t teams);
int qd_bp_get_config(qd_bp *d, int32_t *threads, int32_t *slots, int32_t *ctas_per_sm,
                     int32_t *teams, int32_t *smem_bytes, int32_t *grid);

/* Kernel family (testing / tuning).  qd_bp_create picks automatically: 7 = thread-owned-state
 * shared-memory kernel (a variant is compiled for the graph's degrees), 6 = generic shared-memory
 * kernel, 8 = messages spread over the shared memories of a thread-block cluster of 2 / 4 / 8 SMs
 * (graphs too large or too dense for one SM), 5 = messages in global memory (graphs beyond that; slower),
 * 9 = the thread-owned-state kernel over a thread-block cluster: the rows are split over 2, 4 or 8 CTAs, columns cut by the
 * split are updated redundantly in every CTA that holds one of their edges and their check->bit messages travel through
 * distributed shared memory (block-structured matrices that leave one SM with at most two shots in flight: the
 * spacetime matrices of spacetime_code.py:46-75; rows up to 9, columns up to 4 edges).  In that family qd_bp_configure /
 * qd_bp_get_config read `ctas_per_sm` as the cluster size.
 * select: 0 = automatic, 5 / 6 / 7 / 8 / 9 = force; resets the launch geometry. */
int qd_bp_select_kernel(qd_bp *d, int32_t family);
/* Host-only self-check of family 9's plan (no device needed): split the rows of the CSR graph over `ctas` CTAs, place the
 * messages, and verify the plan (every edge local to exactly one CTA, every remote edge of a replica fed by exactly one
 * export, every access bank-conflict free).  banks = 16 (8-byte messages) or 32 (4-byte).
 * stats[8] = {errors (0 = consistent), local edges, mailbox elements, message lines per CTA, records per CTA,
 *             rows per CTA, exports per CTA (largest), 0}. */
int qd_bp_cluster_plan_check(int32_t m, int32_t n, const int32_t *row_ptr, const int32_t *col_idx, int32_t ctas,
                             int32_t banks, int64_t *stats);
int qd_bp_get_kernel(qd_bp *d, int32_t *family, int32_t *row_degree, int32_t *column_tile);

/* Decode `shots` packed syndromes.  Required: d_synd, d_corr.  Optional (NULL to skip):
 *   d_converged[shots], d_iters[shots], d_llr[shots*n] (posterior log-ratio of the last
 *   iteration), and the fused failure epilogue: d_err (the sampled error / readout the syndrome
 *   came from) + logicals -> d_fail[shots] = any(L (err ^ corr)), and d_counters[4] +=
 *   {failures, shots, non_converged, iterations}.                                             */
int qd_bp_decode_batch(qd_bp *d, const uint32_t *d_synd, int32_t synd_stride, int64_t shots,
                       uint32_t *d_corr, int32_t corr_stride, uint8_t *d_converged,
                       int32_t *d_iters, double *d_llr, const uint32_t *d_err, int32_t err_stride,
                       qd_logicals *logicals, uint8_t *d_fail, int64_t *d_counters, void *stream);
/* Same with HOST buffers: chunked, double-buffered H2D -> decode -> D2H; returns when done. */
int qd_bp_decode_batch_host(qd_bp *d, const uint32_t *h_synd, int32_t synd_stride, int64_t shots,
                            uint32_t *h_corr, int32_t corr_stride, uint8_t *h_converged,
                            int32_t *h_iters);
/* One shot, byte-per-bit host I/O: the literal bp_decoder.decode() replacement. */
int qd_bp_decode_one(qd_bp *d, const uint8_t *h_syndrome, uint8_t *h_decoding, int32_t *converged,
                     int32_t *iters, double *h_llr);

/* ---- BP + OSD ------------------------------------------------------------------------------- */
/* OSD-0: rank the columns by the soft output of the last BP iteration (most likely in error
 * first, ties by index), take the first rank(H) independent ones as information set, solve.
 * method 1 ("osd_e"): then try all non-zero patterns on the first `order` columns outside the
 * information set; method 2 ("osd_cs"): every single such column, then every pair among the
 * first `order`; the lightest candidate wins (OSD-0 first, ties to the earlier candidate).
 * [H | s] is eliminated bit-packed in shared memory when its m * ceil((n+1)/64) * 8 bytes fit, else in a
 * global scratch buffer (slower); QD_EUNSUP beyond 4096 checks / 16000 columns. */
int qd_osd_create(qd_graph *g, int32_t method, int32_t order, qd_osd **out);
int qd_osd_destroy(qd_osd *o);
int qd_osd_rank(qd_osd *o);                       /* rank of H over GF(2) */
/* As qd_bp_decode_batch; shots with converged == 0 get the OSD-0 decoding (d_converged keeps the
 * BP flag; d_fail / d_counters[0] describe the final decoding). */
int qd_bposd_decode_batch(qd_bp *d, qd_osd *o, const uint32_t *d_synd, int32_t synd_stride,
                          int64_t shots, uint32_t *d_corr, int32_t corr_stride, uint8_t *d_converged,
                          int32_t *d_iters, const uint32_t *d_err, int32_t err_stride,
                          qd_logicals *logicals, uint8_t *d_fail, int64_t *d_counters, void *stream);
int qd_bposd_decode_batch_host(qd_bp *d, qd_osd *o, const uint32_t *h_synd, int32_t synd_stride,
                               int64_t shots, uint32_t *h_corr, int32_t corr_stride,
                               uint8_t *h_converged, int32_t *h_iters);

/* ---- small-set-flip ------------------------------------------------------------------------ */
/* hz: graph of the checks whose syndrome is decoded; generators: rows of Hx (host CSR).        */
int qd_ssf_create(qd_graph *hz, int32_t num_generators, const int32_t *h_gen_row_ptr,
                  const int32_t *h_gen_col_idx, int32_t max_rounds, qd_ssf **out);
int qd_ssf_destroy(qd_ssf *d);
int qd_ssf_decode_batch(qd_ssf *d, const uint32_t *d_synd, int32_t synd_stride, int64_t shots,
                        uint32_t *d_corr, int32_t corr_stride, uint8_t *d_converged,
                        int32_t *d_rounds, const uint32_t *d_err, int32_t err_stride,
                        qd_logicals *logicals, uint8_t *d_fail, int64_t *d_counters, void *stream);
int qd_ssf_decode_batch_host(qd_ssf *d, const uint32_t *h_synd, int32_t synd_stride, int64_t shots,
                             uint32_t *h_corr, int32_t corr_stride, uint8_t *h_converged,
                             int32_t *h_rounds);

/* ---- code-capacity experiment: sample -> syndrome -> decode -> residual -> logical flag ---- */
/* Exactly one of bp / ssf is non-NULL; osd (nullable, with bp) adds OSD-0 post-processing.
 * Decodes global shots [shot0, shot0+shots) drawn from
 * `sampler` with key `seed`; h_counters[4] = {failures, shots, non_converged, iterations}.
 * Optional per-shot host outputs (NULL to skip): h_corr (packed, stride words), h_fail,
 * h_converged, h_iters. */
int qd_code_capacity_run(qd_graph *hz, qd_bp *bp, qd_ssf *ssf, qd_osd *osd, qd_sampler *sampler,
                         qd_logicals *logicals, uint64_t seed, int64_t shot0, int64_t shots,
                         int64_t *h_counters, uint32_t *h_corr, int32_t corr_stride,
                         uint8_t *h_fail, uint8_t *h_converged, int32_t *h_iters);

#ifdef __cplusplus
}
#endif
#endif /* QLDPC_B200_H */
