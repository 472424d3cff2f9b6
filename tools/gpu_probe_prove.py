#!/usr/bin/env python
"""GPU probe: Ring VRF prove throughput at ring 1023 / N=2048 for several batch sizes (gpurun)."""
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dot_ring_b200 import _native  # noqa: E402
from oracle import fr, ring_proof as rp  # noqa: E402
from tests.helpers import bench_ring_keys, le64  # noqa: E402
from tests.ring_fixtures import native_ring, native_srs  # noqa: E402

cbits = int(os.environ.get("DR_WINDOW_BITS", "12"))
if os.environ.get("DR_LIB"):  # A/B builds of the same sources (tools only)
    _native.set_default_library(_native.Library(os.environ["DR_LIB"]))
ctx = _native.Context(0)
t0 = time.time()
srs = native_srs(ctx, None, cbits)
print("srs+table s", time.time() - t0, "table GB", srs.table_bytes / 1e9, flush=True)
pk, sk, keys = bench_ring_keys(1023)
params = rp.Params.from_ring_size(1023)
t0 = time.time()
ring = native_ring(srs, keys, params)
print("ring create s", time.time() - t0, flush=True)
out = {"window_bits": cbits, "runs": []}
rng = random.Random(0)
for n in [int(x) for x in os.environ.get("PROVE_NS", "16,256,1024,4096").split(",")]:
    zk = [rng.randrange(fr.R) for _ in range(12 * n)]
    alphas = [b"bench-batch-input" + le64(j) for j in range(n)]
    ads = [b"bench-batch-ad" + le64(j) for j in range(n)]
    ring.prove_batch(alphas[:8], ads[:8], [sk] * 8, [3] * 8, zk_rows=zk[:96])
    t0 = time.time()
    proofs, status = ring.prove_batch(alphas, ads, [sk] * n, [3] * n, zk_rows=zk)
    dt = time.time() - t0
    ph = ring.prove_phase_ms()
    assert not any(status)
    out["runs"].append({"n": n, "wall_s": dt, "proofs_per_s": n / dt, "phase_ms": ph, "device_ms": sum(ph)})
    print(out["runs"][-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(os.environ.get("PROBE_OUT", "gpurun_out/probe_prove.json"), "w"), indent=1)
