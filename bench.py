#!/usr/bin/env python
"""bench.py — decoded shots/s of the BP decode hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one pass of the hot path over one batch of synthetic code-capacity syndromes
(``--shots`` per GPU, default 2^20) of the workload BASELINE.json quotes the metric on:
configs[1], the ~1.6k-qubit random (3,4) hypergraph-product code, BP (min-sum, alpha schedule,
iteration cap 32, early exit) at p = 0.05, prior 2p/3 (scripts/p_sweep.py:4-8 of the reference).

Printed JSON (one line, rank 0):
  value      whole-job decoded shots/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the public host API (bp_decoder.decode_batch: pinned host
             syndromes -> H2D -> decode -> D2H corrections/flags), copies inside the timed region
  roofline   HBM roofline of the dominant kernel (bp_kernel) from algorithmic bytes / event time;
             roofline_smem adds the shared-memory roofline that actually bounds it
  cpu_baseline  the CPU oracle (port of the reference decoder) timed on this host's cores
``--impl reference`` times that CPU port alone (the reference's own decoder, the third-party
`ldpc` package, is not installable here: see DESIGN.md).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

_BP = dict(max_iter=32, bp_method="ms", ms_scaling_factor=0.0)
WORKLOADS = {
    # name: (qecc file | code_builders function, p, bp options, description)
    "hgp_34_n1600": ("hgp_34_n1600.qecc", 0.05, _BP, "(3,4) HGP"),                       # BASELINE configs[1]
    "hgp_34_n6400": ("hgp_34_n6400.qecc", 0.05, _BP, "(3,4) HGP"),                       # BASELINE configs[1]
    "hgp_34_n225": ("hgp_34_n225.qecc", 0.05, _BP, "(3,4) HGP"),                         # BASELINE configs[0]
    "lp_882": ("builder:pk19_b1_882", 0.05, _BP, "QC lifted product [[882,24]]"),        # BASELINE configs[2]
    "qclp_1054": ("builder:pk20_qc_1054", 0.05, _BP, "QC lifted product [[1054,140]]"),
    "a5lp_2700": ("builder:a5_lift_2700", 0.05, _BP, "PSL(2,5) lifted product n=2700"),  # BASELINE configs[3]
    # BASELINE configs[3] as a Cayley-graph Tanner code: PSL(2,4) ~ A5, Morgenstern B generators, [6,3,3] local codes
    "cayley_lp_4320": ("builder:psl2_tanner_lp_4320", 0.05, _BP, "PSL(2,4) Cayley-graph Tanner lifted product [[4320,28]]"),
    # BASELINE configs[4]: spacetime matrix of the [[225,9]] code, T rounds, phenomenological faults drawn from
    # the decoder's priors (data and measurement 2p/3); Stim circuit-level sampling is a "next" row
    "st_hgp225_T3": ("spacetime:hgp_34_n225.qecc:3", 0.02, _BP, "spacetime (T=3) of the (3,4) HGP [[225,9]] code"),
    "st_hgp1600_T3": ("spacetime:hgp_34_n1600.qecc:3", 0.02, _BP, "spacetime (T=3) of the (3,4) HGP n=1600 code"),
    # the reference's default number of rounds (misc/p_sweep.py:62) on the BASELINE configs[1] code
    "st_hgp1600_T1": ("spacetime:hgp_34_n1600.qecc:1", 0.02, _BP, "spacetime (T=1) of the (3,4) HGP n=1600 code"),
}
SSF_DEFAULT_P = 0.01


def decoding_problem(name, code, p):
    """(check matrix, logical operators, per-column flip probability, per-column prior) the decoder sees."""
    src = WORKLOADS[name][0]
    q, prior = probabilities(p)
    if src.startswith("spacetime:"):
        from exp_ldpc_b200.spacetime import SpacetimeCode
        T = int(src.split(":")[2])
        m, n = code.checks.z.shape
        h_st = SpacetimeCode(code.checks.z, T).spacetime_check_matrix
        lz = np.asarray(code.logicals.z) & 1
        l_st = np.concatenate([lz] * (T + 1) + [np.zeros((lz.shape[0], T * m), dtype=lz.dtype)], axis=1)
        pr = np.full(h_st.shape[1], prior)
        return h_st, l_st, pr, pr
    n = code.checks.z.shape[1]
    return code.checks.z, np.asarray(code.logicals.z) & 1, np.full(n, q), np.full(n, prior)


def load_workload(name):
    src = WORKLOADS[name][0]
    if src.startswith("spacetime:"):
        from exp_ldpc_b200.codes import load_code
        return load_code(os.path.join(ROOT, "codes", src.split(":")[1]))
    if src.startswith("builder:"):
        from exp_ldpc_b200 import cayley_codes, code_builders
        name = src.split(":", 1)[1]
        return getattr(code_builders if hasattr(code_builders, name) else cayley_codes, name)()
    from exp_ldpc_b200.codes import load_code
    return load_code(os.path.join(ROOT, "codes", src))


def probabilities(p):
    """depolarizing_noise(p, pm=p) with rounds=0 (noise_model.py:117-123) and the 2p/3 prior."""
    return 2.0 * p / 3.0 + p - 4.0 / 3.0 * p * p, 2.0 * p / 3.0


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()          # exact PID we started
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_port_throughput(code, p, opts, target_seconds=12.0, threads=0, decoder="bp", problem=None):
    """Time the CPU oracle (the port of the reference's decoder) on a bounded sample of the same
    workload; returns (shots/s, cores, sample description, mean iterations)."""
    from oracle import pyoracle as po
    hz, lz, q, prior = problem
    cores = po.num_threads() if threads <= 0 else threads
    osd_dec = {"osd0": "bposd0", "osd_e": "bposd_e", "osd_cs": "bposd_cs"}[opts.get("osd_method", "osd0")]
    kw = dict(decoder={"bposd": osd_dec}.get(decoder, decoder), hx=code.checks.x, osd_order=opts.get("osd_order", 0),
              max_iter=0 if decoder == "ssf" else opts["max_iter"],        # SSF: no round cap, like ssf_decoder's default
              bp_method=opts["bp_method"],
              ms_scaling_factor=opts["ms_scaling_factor"], seed=0, threads=cores)
    probe = 64 * cores
    t0 = time.perf_counter()
    po.run_code_capacity(hz, lz, q, prior, shots=probe, **kw)
    dt = max(time.perf_counter() - t0, 1e-4)
    shots = int(max(probe, min(5_000_000, probe * target_seconds / dt)))
    t0 = time.perf_counter()
    r = po.run_code_capacity(hz, lz, q, prior, shots=shots, shot0=probe, **kw)
    dt = time.perf_counter() - t0
    return shots / dt, cores, f"{shots} shots of the same workload, {dt:.1f} s wall on {cores} threads", r["iters_total"] / shots


def single_core_baseline(code, p, opts, args, problem):
    """the port on ONE core beside the all-core figure (a short bounded sample), and what the figure means."""
    v1, _, sample1, _ = cpu_port_throughput(code, p, opts, target_seconds=min(3.0, max(0.2, (args.cpu_seconds or 12.0) / 4.0)),
                                            threads=1, decoder=args.decoder, problem=problem)
    note = ("C port of the decoder alone: the reference adds a Python call, GF(2) array construction and slicing per shot "
            "(_experiment.py:200-209), so this is an upper bound on its throughput")
    if args.decoder == "ssf":
        note = ("the reference has no small-set-flip; this is the plain restatement the parity tests check against: every round "
                "rescans all generators with a non-zero local syndrome (no candidate cache, unlike the CUDA kernel)")
    return {"single_core": {"value": v1, "unit": "shots/s", "sample": sample1}, "note": note}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="hgp_34_n1600", choices=sorted(WORKLOADS))
    ap.add_argument("--shots", type=int, default=1 << 20, help="shots per GPU per step")
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--osd-method", default="osd0", choices=["osd0", "osd_e", "osd_cs"])
    ap.add_argument("--osd-order", type=int, default=0)
    ap.add_argument("--decoder", default="bp", choices=["bp", "ssf", "bposd"],
                    help="bposd = BP, then OSD-0 on the shots BP left unconverged")
    ap.add_argument("--p", type=float, default=None)
    ap.add_argument("--max-iter", type=int, default=None)
    ap.add_argument("--bp-method", default=None)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--slots", type=int, default=0)
    ap.add_argument("--ctas", type=int, default=0)
    ap.add_argument("--teams", type=int, default=0)
    ap.add_argument("--kernel", default="auto", choices=["auto", "regular", "regular_cluster", "generic", "cluster", "global"],
                    help="force a BP kernel family (tuning / comparison); the default is the library's own choice")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=None,
                    help="wall seconds of CPU work per reference-arm step / for the cpu_baseline sample (default: bounded automatically)")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    _, p, opts, family = WORKLOADS[args.workload]
    opts = dict(opts)
    ssf = args.decoder == "ssf"
    if ssf:
        p = SSF_DEFAULT_P
    if args.p is not None:
        p = args.p
    if args.max_iter is not None:
        opts["max_iter"] = args.max_iter
    if args.bp_method is not None:
        opts["bp_method"] = args.bp_method
    if args.decoder == "bposd":
        opts["osd_method"], opts["osd_order"] = args.osd_method, args.osd_order
    code = load_workload(args.workload)
    problem = decoding_problem(args.workload, code, p)
    hz_dec, lz_dec, q_vec, prior_vec = problem
    m, n = hz_dec.shape
    nnz = int(hz_dec.nnz)
    q, prior = float(q_vec[0]), float(prior_vec[0])
    osd = args.decoder == "bposd"
    how = ("small-set-flip (candidate sets = subsets of the X-generator supports)" if ssf else
           f"BP{' + ' + args.osd_method + ' order ' + str(args.osd_order) if osd else ''} {opts['bp_method']} alpha={'1-2^-it' if opts['ms_scaling_factor'] == 0 else opts['ms_scaling_factor']} "
           f"max_iter={opts['max_iter']} early-exit")
    metric = ("decoded shots/sec (SSF, fixed p)" if ssf else
              "decoded shots/sec (BP + OSD, fixed p and iteration cap)" if osd else
              "decoded shots/sec (BP, fixed p and iteration cap)")
    dtype = "u32" if ssf else args.precision
    config = {"workload": f"{args.workload}: {family} n={n} m={m} nnz={nnz} k={code.num_logicals}, "
                          f"{'phenomenological faults' if WORKLOADS[args.workload][0].startswith('spacetime:') else 'code capacity'}, "
                          f"{how}, p={p} (flip q={q:.5f}, prior 2p/3)",
              "shots_per_step_per_gpu": args.shots, "p": p, "decoder": args.decoder, "max_iter": 0 if ssf else opts["max_iter"],
              "bp_method": opts["bp_method"], "precision": dtype, "parallelism": f"shots sharded over {world} rank(s), one all-reduce of int64[4] per step",
              "l2": "inputs larger than L2 (no flush needed)"}

    # ------------------------------------------------------------------ reference arm (CPU port)
    if args.impl == "reference":
        if rank != 0:
            return
        rates, step_ms = [], []
        cores = 0
        sample = ""
        per_step = args.cpu_seconds or max(2.0, min(20.0, 120.0 / (args.steps + args.warmup)))
        for i in range(args.warmup + args.steps):
            r, cores, sample, _ = cpu_port_throughput(code, p, opts, target_seconds=per_step if i >= args.warmup else min(1.0, per_step),
                                                      decoder=args.decoder, problem=problem)
            if i >= args.warmup:
                rates.append(r)
                step_ms.append(1e3 * float(sample.split(" shots")[0]) / r)     # wall time of this step's bounded sample
        v = statistics.mean(rates)
        line = {"metric": metric, "value": v, "unit": "shots/s",
                "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": statistics.mean(step_ms), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u32" if ssf else "f64",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "shots/s", "cores": cores, "kind": "port",
                                 "sample": sample + " per step (oracle/qldpc_oracle.c; the reference's own decoder `ldpc` is not installable here)",
                                 **single_core_baseline(code, p, opts, args, problem)},
                "e2e": {"value": v, "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from exp_ldpc_b200.decoders import Logicals, Sampler, TannerGraph, bp_decoder, bposd_decoder, row_words, ssf_decoder
    graph = TannerGraph(hz_dec, device=local_rank)
    logicals = Logicals(lz_dec, device=local_rank)
    sampler = Sampler(q_vec, n, device=local_rank)
    if ssf:
        dec = ssf_decoder(graph, code.checks.x)
        cfg = {"kernel": "ssf"}
    else:
        dec = (bposd_decoder if osd else bp_decoder)(graph, channel_probs=prior_vec, precision=args.precision, kernel=args.kernel,
                                                     **opts)
        if args.threads or args.slots or args.ctas or args.teams:
            dec.configure(args.threads, args.slots, args.ctas, args.teams)
        cfg = dec.config()
    config["launch"] = cfg

    def launch(cnt_row):
        if ssf:
            dec.decode_batch_device(d_syn, ss, shots, d_cor, es, d_conv, d_it, d_err, es, logicals, d_fail, cnt_row, stream)
        else:
            dec.decode_batch_device(d_syn, ss, shots, d_cor, es, d_conv, d_it, None, d_err, es, logicals, d_fail, cnt_row, stream)

    shots = args.shots
    es, ss = row_words(n), row_words(m)
    stream = torch.cuda.current_stream().cuda_stream
    d_err = torch.empty((shots, es), dtype=torch.int32, device=dev)
    d_syn = torch.empty((shots, ss), dtype=torch.int32, device=dev)
    d_cor = torch.empty((shots, es), dtype=torch.int32, device=dev)
    d_fail = torch.empty(shots, dtype=torch.uint8, device=dev)
    d_conv = torch.empty(shots, dtype=torch.uint8, device=dev)
    d_it = torch.empty(shots, dtype=torch.int32, device=dev)
    d_cnt = torch.zeros((args.steps + 1, 4), dtype=torch.int64, device=dev)   # one counter row per step
    shot0 = rank * shots * (args.steps + args.warmup + 1)

    def make_inputs(step):
        sampler.run_device(0, 0, shot0 + step * shots, shots, d_err, es, stream)
        graph.syndrome_batch_device(d_err, es, shots, d_syn, ss, stream)

    def decode_step(row):
        cnt_row = d_cnt[row]
        launch(cnt_row)
        if world > 1:
            dist.all_reduce(cnt_row)   # the one collective of the path: sum of the 4 counters

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    make_inputs(0)
    for _ in range(args.warmup):
        decode_step(args.steps)
    barrier()
    d_cnt.zero_()
    clocks = ClockSampler(local_rank)
    clocks.start()
    time.sleep(0.25)
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(args.steps):
        cnt_row = d_cnt[i]
        k_ev[i][0].record()
        launch(cnt_row)
        k_ev[i][1].record()
        if world > 1:
            dist.all_reduce(cnt_row)
    stop.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clk = clocks.stop()
    ms_total = start.elapsed_time(stop)
    kernel_ms = statistics.mean(a.elapsed_time(b) for a, b in k_ev)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    cnt = d_cnt[:args.steps].sum(dim=0).cpu().numpy().astype(np.int64)
    total_shots = shots * world * args.steps
    value = total_shots / (ms_total * 1e-3)
    mean_iters = float(cnt[3]) / max(1, int(cnt[1]))

    # ------------------------------------------------------------------ e2e through the host API
    e2e = None
    if not args.no_e2e:
        h_syn = torch.empty((shots, ss), dtype=torch.int32, pin_memory=True)
        h_syn.copy_(d_syn)
        torch.cuda.synchronize()
        h_np = h_syn.numpy().view(np.uint32)
        for _ in range(2):          # full-size warm-up: the pinned staging buffers of the library are sized here
            dec.decode_batch(h_np, packed=True, return_packed=True)
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 3))
        for _ in range(e2e_steps):
            corr, conv, its = dec.decode_batch(h_np, packed=True, return_packed=True)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": shots * world * e2e_steps / float(tt.item()), "unit": "shots/s",
               "h2d_bytes_per_step": int(shots * ss * 4), "d2h_bytes_per_step": int(shots * (es * 4 + 1 + 4)),
               "api": f"{args.decoder}_decoder.decode_batch(packed host syndromes) -> qd_{args.decoder}_decode_batch_host",
               "steps": e2e_steps}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ rooflines (rank 0's kernel)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    bytes_per_shot = 4 * ((m + 31) // 32) + 4 * ((n + 31) // 32) + 1          # SURVEY.md §8(d) B_dec
    achieved = bytes_per_shot * shots / (kernel_ms * 1e-3) / 1e9
    msg_bytes = 8 if args.precision == "f64" else 4
    smem_bytes_per_shot = mean_iters * nnz * 4 * msg_bytes                      # SURVEY.md §8(d) B_smem
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    sm_mhz = clk.get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    smem_peak = sm_count * 128 * sm_mhz * 1e6 / 1e9
    smem_achieved = smem_bytes_per_shot * shots / (kernel_ms * 1e-3) / 1e9
    kernel_name = {"regular": "bp7_kernel", "regular_cluster": "bp7_kernel (cluster mode)", "ssf": "ssf_kernel", "cluster": "bpc_kernel",
                   "global": "bpg_kernel"}.get(cfg.get("kernel"), "bp_kernel")
    if osd:
        kernel_name += " + osd0_kernel"
    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed `ncu --set full`
    # capture of this workload (per launch of the captured shot count, scaled to this launch)
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"{args.workload}:{dtype}")
        if tr and tr.get("kernel") == kernel_name:
            traffic = tr["dram_bytes_per_shot"] * shots
    except (OSError, ValueError):
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "kernel": kernel_name, "kernel_ms": kernel_ms,
                "algorithmic_bytes_per_launch": bytes_per_shot * shots,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_shot": bytes_per_shot,
                "note": ("HBM is touched for I/O only; the kernel is bound by integer issue rate" if ssf else
                         "HBM is touched for I/O only; the kernel is bound by shared memory / issue, see roofline_smem")}
    roofline_smem = {"bound": "shared-memory", "achieved": smem_achieved, "peak": smem_peak, "unit": "GB/s",
                     "frac": smem_achieved / smem_peak, "algorithmic_bytes_per_shot": smem_bytes_per_shot,
                     "mean_iterations": mean_iters, "edge_updates_per_s": 2.0 * nnz * mean_iters * shots / (kernel_ms * 1e-3),
                     "peak_source": f"{sm_count} SMs x 128 B/clk x {sm_mhz:.0f} MHz (median SM clock during the timed region)"}

    # the HBM-bound kernel of the path (H e mod 2 on the step's packed errors), timed alone: SURVEY.md §8(d) B_syn
    # without the logical bits = 4 * (ceil(n/32) + ceil(m/32)) bytes per shot against the measured copy bandwidth
    for _ in range(3):
        graph.syndrome_batch_device(d_err, es, shots, d_syn, ss, stream)
    s_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s_ev[0].record()
    for _ in range(10):
        graph.syndrome_batch_device(d_err, es, shots, d_syn, ss, stream)
    s_ev[1].record()
    torch.cuda.synchronize()
    syn_ms = s_ev[0].elapsed_time(s_ev[1]) / 10
    syn_bytes = 4 * ((n + 31) // 32 + (m + 31) // 32)
    syn_achieved = syn_bytes * shots / (syn_ms * 1e-3) / 1e9
    roofline_syndrome = {"bound": "hbm", "achieved": syn_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": syn_achieved / hbm_peak,
                         "kernel": "syndrome128_kernel", "kernel_ms": syn_ms, "algorithmic_bytes_per_shot": syn_bytes,
                         "note": "inputs of one step (larger than L2 at the default shot count); not part of `value`"}

    cpu = None
    if not args.no_cpu_baseline:
        v, cores, sample, cpu_iters = cpu_port_throughput(code, p, opts, target_seconds=args.cpu_seconds or 12.0, decoder=args.decoder,
                                                          problem=problem)
        cpu = {"value": v, "unit": "shots/s", "cores": cores, "kind": "port", "sample": sample,
               "mean_iterations": cpu_iters}
        cpu.update(single_core_baseline(code, p, opts, args, problem))

    if ssf:
        roofline_smem = None
    line = {"metric": metric, "value": value, "unit": "shots/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": config, "e2e": e2e, "gpu_launches": args.steps * (2 if osd else 1), "clocks": clk, "roofline": roofline,
            "roofline_smem": roofline_smem, "roofline_syndrome": roofline_syndrome, "cpu_baseline": cpu,
            "counters": {"failures": int(cnt[0]), "shots": int(cnt[1]), "non_converged": int(cnt[2]),
                         "iterations": int(cnt[3])},
            "wall_s_timed_region": t_wall}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
