#!/usr/bin/env python
"""Integer / FP64 pipe ceilings and single-chain latencies (dr_microbench) -> JSON on stdout."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dot_ring_b200 import _native
ctx = _native.Context(0)
info = ctx.device_info()
out = {"device": info["name"], "sm_count": info["sm_count"]}
for kind, iters in (("imad", 20000), ("imad_wide", 20000), ("dfma", 20000), ("imad_dfma", 20000), ("fq_mul", 2000), ("fq_sqr", 2000), ("fr_mul", 4000), ("fr_sqr", 4000), ("g1_madd", 400),
                    ("fr_chain", 4000), ("fq_chain", 2000), ("fr_inv_chain", 40), ("fq_inv_chain", 20)):
    ops, ms = ctx.microbench(kind, iters)
    out[kind] = {"ops_per_s": ops, "ms": ms}
    if kind.endswith("chain"):
        out[kind]["ns_per_op_one_chain"] = 1e9 / (ops / (32 * info["sm_count"]))
print(json.dumps(out, indent=1))
