#!/usr/bin/env python
"""GPU probe: G1 MSM sweep 2^11 .. 2^20 (BASELINE configs[2]) -> gpurun_out/probe_msm.json.
Variable-base bucket method over the synthetic SRS tau^i * G, three scalar distributions; fixed-base commit beside it for n <= 6145."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dot_ring_b200 import _native  # noqa: E402
from tests import msm_cases  # noqa: E402

ctx = _native.Context(0)
out = {"imad_peak": ctx.microbench("imad", 20000)[0], "sweep": []}
NAMES = {0: "uniform", 1: "ones", 2: "bits"}
for k in range(11, 21):
    n = 1 << k
    for dist in (0, 1, 2):
        iters = 5 if k <= 16 else 2
        ms, c, res = ctx.g1_msm_bench(n, iters, 7, dist, msm_cases.TAU)
        ok = res == msm_cases.expected_synthetic(n, 7, dist)
        # canonical Pippenger work: min_c ceil(255/c) (10 n + 14 2^c) Fq mul, 600 IMAD each
        canon = min(-(-255 // cc) * (n * 10 + (1 << cc) * 14) for cc in range(2, 21)) * 600
        row = {"log2_n": k, "distribution": NAMES[dist], "window_bits": c, "ms": ms, "points_per_s": n / (ms * 1e-3), "parity": ok,
               "frac_of_imad_peak_canonical": canon / (ms * 1e-3) / out["imad_peak"]}
        out["sweep"].append(row)
        print(row, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_msm.json", "w"), indent=1)
