#!/usr/bin/env python
"""One-shot GPU probe: integer-pipe ceilings, SRS table build time, commit throughput.
Writes gpurun_out/probe.json (run under gpurun)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dot_ring_b200 import _native  # noqa: E402
from dot_ring_b200.srs import read_srs_file  # noqa: E402

out = {}
if os.environ.get("DR_LIB"):  # A/B builds of the same sources (tools only)
    _native.set_default_library(_native.Library(os.environ["DR_LIB"]))
ctx = _native.Context(0)
ctx.set_commit_mode(int(os.environ.get("COMMIT_MODE", "0")))
out["device"] = ctx.device_info()
for kind, iters in (("imad", 20000), ("imad_wide", 20000), ("fq_mul", 2000), ("fr_mul", 4000), ("g1_madd", 300)):
    ops, ms = ctx.microbench(kind, iters)
    out[f"mb_{kind}"] = {"ops_per_s": ops, "ms": ms}
    print(kind, f"{ops:.4g} ops/s", f"{ms:.2f} ms", flush=True)
raw = read_srs_file()
cbits = int(os.environ.get("DR_WINDOW_BITS", "12"))
t0 = time.time()
srs = _native.NativeSrs(ctx, raw.g1_be96, raw.g2_be192, cbits)
out["table"] = {"window_bits": cbits, "bytes": srs.table_bytes, "build_s": time.time() - t0}
print("table", out["table"], flush=True)
res = []
for n, batch in ((2048, 16), (2048, 256), (2048, 4096), (6145, 1024), (2048, 1), (6145, 1)):
    ms, first = srs.commit_bench(n, batch, 3, seed=n)
    pts = n * batch / (ms * 1e-3)
    res.append({"n": n, "batch": batch, "ms": ms, "points_per_s": pts, "first": first.hex()[:32]})
    print(res[-1], flush=True)
out["commit"] = res
out["launches"] = ctx.library.launch_count()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(os.environ.get("PROBE_OUT", "gpurun_out/probe.json"), "w"), indent=1)
