#!/usr/bin/env python
"""Sweep the BP launch geometry (threads, slots, CTAs/SM) on the GPU for one workload and print
decoded shots/s for each (device-resident inputs, CUDA-event timing).  Development tool."""
import argparse
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import WORKLOADS, load_workload, probabilities  # noqa: E402
from exp_ldpc_b200.codes import load_code  # noqa: E402
from exp_ldpc_b200.decoders import Logicals, Sampler, TannerGraph, bp_decoder, row_words  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="hgp_34_n1600")
ap.add_argument("--shots", type=int, default=1 << 18)
ap.add_argument("--precision", default="f64")
ap.add_argument("--p", type=float, default=None)
ap.add_argument("--bp-method", default=None)
ap.add_argument("--max-iter", type=int, default=None)
ap.add_argument("--threads", default="256,512,768,1024")
ap.add_argument("--slots", default="1,2,3,4,5,6,8")
ap.add_argument("--ctas", default="1,2,3,4")
ap.add_argument("--teams", default="1")
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()

_, p, opts, _family = WORKLOADS[a.workload]
opts = dict(opts)
if a.p is not None:
    p = a.p
if a.bp_method:
    opts["bp_method"] = a.bp_method
if a.max_iter is not None:
    opts["max_iter"] = a.max_iter
code = load_workload(a.workload)
m, n = code.checks.z.shape
q, prior = probabilities(p)
dev = torch.device("cuda:0")
graph, logicals, sampler = TannerGraph(code.checks.z), Logicals(code.logicals.z), Sampler(q, n)
dec = bp_decoder(graph, error_rate=prior, precision=a.precision, **opts)
es, ss = row_words(n), row_words(m)
shots = a.shots
d_err = torch.empty((shots, es), dtype=torch.int32, device=dev)
d_syn = torch.empty((shots, ss), dtype=torch.int32, device=dev)
d_cor = torch.empty((shots, es), dtype=torch.int32, device=dev)
d_fail = torch.empty(shots, dtype=torch.uint8, device=dev)
d_cnt = torch.zeros(4, dtype=torch.int64, device=dev)
sampler.run_device(0, 0, 0, shots, d_err, es)
graph.syndrome_batch_device(d_err, es, shots, d_syn, ss)
print("default config", dec.config(), flush=True)
results = []
combos = [dec.config()] + [dict(threads=t, slots=s, ctas_per_sm=c, teams=g) for t, s, c, g in itertools.product(
    [int(x) for x in a.threads.split(",")], [int(x) for x in a.slots.split(",")], [int(x) for x in a.ctas.split(",")],
    [int(x) for x in a.teams.split(",")])]
for cfg in combos:
    t, s, c, g = cfg["threads"], cfg["slots"], cfg["ctas_per_sm"], cfg["teams"]
    if t * c > 2048:
        continue
    try:
        got = dec.configure(t, s, c, g)
    except Exception:
        continue
    if got["smem_bytes"] * c > 232448 - 1024 * c:
        continue
    best = 1e9
    for r in range(a.reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d_cnt.zero_()
        e0.record()
        dec.decode_batch_device(d_syn, ss, shots, d_cor, es, None, None, None, d_err, es, logicals, d_fail, d_cnt)
        e1.record()
        torch.cuda.synchronize()
        if r:
            best = min(best, e0.elapsed_time(e1))
    cnt = d_cnt.tolist()
    rate = shots / (best * 1e-3)
    results.append((rate, t, s, c, g, got["smem_bytes"]))
    print(f"threads={t:5d} slots={s:2d} ctas={c} teams={g} smem={got['smem_bytes']:7d}  {rate / 1e6:8.3f} Mshots/s  mean_iters={cnt[3] / max(1, cnt[1]):.2f}", flush=True)
results.sort(reverse=True)
print("BEST", results[:5])
