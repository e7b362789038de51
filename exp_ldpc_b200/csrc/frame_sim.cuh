// frame_sim.cuh — Pauli-frame sampler for the storage-simulation circuits.
//
// Stands in for `stim.Circuit(text).compile_sampler().sample(shots)` as the reference uses it at
// python/qldpc/misc/_experiment.py:171-197, restricted to the gate set its circuit builder emits
// (python/qldpc/storage_sim.py:38-75,134-185; noise python/qldpc/noise_model.py:117-151):
// R/RZ/RX, CX, CZ, M/MZ/MX/MRZ/MRX with flip probability, DEPOLARIZE1/2, X_ERROR/Z_ERROR; TICK,
// DETECTOR, OBSERVABLE_INCLUDE, SHIFT_COORDS carry no sampling semantics; REPEAT is unrolled by the
// host compiler (exp_ldpc_b200/circuits.py).  CPU restatement: oracle/qldpc_oracle.c
// orc_circuit_sample (same Philox stream => bit-identical records).
//
// One thread per shot.  The frame (one X bit and one Z bit per qubit: the Pauli by which the noisy
// run differs from the noiseless one) lives in shared memory, word-major ([word][thread]) so that a
// warp touching the same qubit touches 32 different banks.  A measurement record bit is the frame
// component that anticommutes with the measured operator, XOR the readout flip: outcomes are
// reported relative to the noiseless reference sample (all deterministic outcomes of these
// circuits — Z checks on |0..0>, the data readout — are 0 there).  Noise site k of shot g draws
// philox(key = seed, ctr = (g_lo, g_hi, k, 0x51)): word 0 against the site's threshold, word 1
// picks the Pauli (1 of 3, or 1 of 15 for two-qubit depolarisation).
#pragma once

#include "common.cuh"

namespace qd {

enum FrameOp : int { FR_RESET = 0, FR_CX = 1, FR_CZ = 2, FR_MZ = 3, FR_MX = 4, FR_MRZ = 5, FR_MRX = 6, FR_DEP1 = 7, FR_DEP2 = 8,
                     FR_XERR = 9, FR_ZERR = 10 };

struct FrameParams {
    const int4 *ops;            // {opcode | threshold index << 8, q0, q1, noise site}
    int nops;
    const uint32_t *thr;        // thresholds: event iff u32 < thr[i]; thr[0] = 0 (noiseless)
    int nqw;                    // frame words per Pauli component: ceil(num_qubits / 32)
    int nmeas;
    uint32_t k0, k1;            // Philox key = seed
    long long shot0, shots;
    uint32_t *rec; int stride;  // [shots][stride] packed measurement records
    const int2 *forced;         // fault-propagation mode (detector error models): "shot" g runs the NOISELESS circuit
                                // with exactly one fault, {noise site, Pauli code} = forced[g]; no random numbers
};

__global__ void frame_kernel(const FrameParams p)
{
    extern __shared__ uint32_t fsm[];
    const int NT = (int)blockDim.x, tid = (int)threadIdx.x;
    uint32_t *fx = fsm, *fz = fsm + (size_t)p.nqw * NT;
    const long long s = (long long)blockIdx.x * NT + tid;
    const bool live = s < p.shots;
    const unsigned long long g = (unsigned long long)(p.shot0 + (live ? s : 0));
    for (int w = 0; w < p.nqw; ++w) { fx[w * NT + tid] = 0u; fz[w * NT + tid] = 0u; }
    uint32_t *out = p.rec + (live ? s : 0) * (long long)p.stride;
    uint32_t cur = 0u;
    int nbit = 0, wout = 0;
    auto getb = [&](const uint32_t *f, int q) -> uint32_t { return (f[(q >> 5) * NT + tid] >> (q & 31)) & 1u; };
    auto xorb = [&](uint32_t *f, int q, uint32_t v) { f[(q >> 5) * NT + tid] ^= v << (q & 31); };
    auto clrb = [&](uint32_t *f, int q) { f[(q >> 5) * NT + tid] &= ~(1u << (q & 31)); };
    auto emit = [&](uint32_t bit) {
        cur |= bit << nbit;
        if (++nbit == 32) { if (live) out[wout] = cur; ++wout; cur = 0u; nbit = 0; }
    };
    for (int i = 0; i < p.nops; ++i) {
        const int4 op = __ldg(p.ops + i);
        const int code = op.x & 0xff, ti = op.x >> 8, q0 = op.y, q1 = op.z;
        uint32_t r[4] = {0u, 0u, 0u, 0u};
        bool hit = false;
        int forced_code = 0;
        if (ti) {
            if (p.forced) {
                const int2 f = __ldg(p.forced + g);
                hit = f.x == op.w;
                forced_code = f.y;
            } else {
                philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)op.w, 0x51u, p.k0, p.k1, r);
                hit = r[0] < __ldg(p.thr + ti);
            }
        }
        switch (code) {
        case FR_RESET: clrb(fx, q0); clrb(fz, q0); break;
        case FR_CX: xorb(fx, q1, getb(fx, q0)); xorb(fz, q0, getb(fz, q1)); break;
        case FR_CZ: { const uint32_t a = getb(fx, q0), b = getb(fx, q1); xorb(fz, q0, b); xorb(fz, q1, a); } break;
        case FR_MZ: emit(getb(fx, q0) ^ (hit ? 1u : 0u)); break;
        case FR_MX: emit(getb(fz, q0) ^ (hit ? 1u : 0u)); break;
        case FR_MRZ: emit(getb(fx, q0) ^ (hit ? 1u : 0u)); clrb(fx, q0); clrb(fz, q0); break;
        case FR_MRX: emit(getb(fz, q0) ^ (hit ? 1u : 0u)); clrb(fx, q0); clrb(fz, q0); break;
        case FR_DEP1:
            if (hit) {
                const uint32_t c = p.forced ? (uint32_t)forced_code
                                            : 1u + (uint32_t)(((unsigned long long)r[1] * 3ull) >> 32);      // 1 X, 2 Y, 3 Z
                xorb(fx, q0, c != 3u ? 1u : 0u); xorb(fz, q0, c != 1u ? 1u : 0u);
            }
            break;
        case FR_DEP2:
            if (hit) {
                const uint32_t c = p.forced ? (uint32_t)forced_code
                                            : 1u + (uint32_t)(((unsigned long long)r[1] * 15ull) >> 32);     // (pa, pb) != (I, I)
                const uint32_t pa = c & 3u, pb = c >> 2;                                            // 0 I, 1 X, 2 Y, 3 Z
                xorb(fx, q0, (pa == 1u || pa == 2u) ? 1u : 0u); xorb(fz, q0, pa >= 2u ? 1u : 0u);
                xorb(fx, q1, (pb == 1u || pb == 2u) ? 1u : 0u); xorb(fz, q1, pb >= 2u ? 1u : 0u);
            }
            break;
        case FR_XERR: if (hit) xorb(fx, q0, 1u); break;
        case FR_ZERR: if (hit) xorb(fz, q0, 1u); break;
        default: break;
        }
    }
    if (live) {
        if (nbit) out[wout++] = cur;
        for (; wout < p.stride; ++wout) out[wout] = 0u;
    }
}

}  // namespace qd
