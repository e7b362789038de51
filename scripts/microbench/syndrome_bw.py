#!/usr/bin/env python
"""HBM roofline of the bit-packed syndrome kernel (qd_syndrome_batch): algorithmic bytes
4 * (ceil(n/32) + ceil(m/32)) per shot over the CUDA-event time, against MEASURED_PEAKS.json.
QD_SYNDROME_OLD=1 selects the 32-shot-tile kernel of round 1, QD_SYNDROME_OLD=1 QD_SYNDROME_GATHER=1 the
one-thread-per-syndrome-word gather kernel, for comparison.  Code names may be given as arguments."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from exp_ldpc_b200.codes import load_code
from exp_ldpc_b200.decoders import Sampler, TannerGraph, row_words

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
dev = torch.device("cuda:0")
for name in (sys.argv[1:] or ("hgp_34_n225", "hgp_34_n1600", "hgp_34_n6400")):
    code = load_code(os.path.join(ROOT, "codes", name + ".qecc"))
    hz = code.checks.z
    m, n = hz.shape
    shots = 1 << 22 if n < 2000 else 1 << 20
    g = TannerGraph(hz)
    smp = Sampler(0.05, n)
    es, ss = row_words(n), row_words(m)
    err = torch.empty((shots, es), dtype=torch.int32, device=dev)
    syn = torch.empty((shots, ss), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    smp.run_device(0, 0, 0, shots, err, es, st)
    for _ in range(3):
        g.syndrome_batch_device(err, es, shots, syn, ss, st)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    reps = 10
    for _ in range(reps):
        g.syndrome_batch_device(err, es, shots, syn, ss, st)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    alg = 4 * ((n + 31) // 32 + (m + 31) // 32) * shots
    moved = 4 * (es + ss) * shots
    print(f"{name}: {shots} shots {ms:.3f} ms  algorithmic {alg / ms / 1e6:.0f} GB/s = {alg / ms / 1e6 / peak:.3f} of {peak:.0f} GB/s"
          f"  (bytes moved incl. stride padding: {moved / ms / 1e6:.0f} GB/s)")
