#!/usr/bin/env python
"""A/B of build variants of the dense commit kernel: python tools/gpu_commit_ab.py default build/var/lib_m3.so ...
Each library builds the bench's table (GLV, 16-bit windows, full SRS) and times dr_kzg_commit_bench at the prover's shapes."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dot_ring_b200 import _native
from dot_ring_b200.srs import read_srs_file

raw = read_srs_file(None, None)
out = {}
for name in sys.argv[1:] or ["default"]:
    lib = _native.default_library() if name == "default" else _native.Library(os.path.abspath(name))
    ctx = _native.Context(0, lib)
    t0 = time.perf_counter()
    srs = _native.NativeSrs(ctx, raw.g1_be96, raw.g2_be192, 16, 0, True)
    ctx.sync()
    res = {"table_s": round(time.perf_counter() - t0, 2)}
    for n, batch in ((6145, 1024), (6145, 4096), (2047, 4096), (6145, 512)):
        ms, first = srs.commit_bench(n, batch, 4)
        res[f"n{n}_b{batch}"] = {"ms": round(ms, 3), "madd_per_s": round(n * batch * 16 / (ms * 1e-3) / 1e9, 4), "first": first.hex()[:16]}
    out[name] = res
    print(name, json.dumps(res), flush=True)
    srs.close()
    ctx.trim()
    ctx.close()
