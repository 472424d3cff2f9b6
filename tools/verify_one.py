#!/usr/bin/env python
"""One batch of Ring VRF verifications (per-item and aggregated) for profiling: python tools/verify_one.py [n]."""
import os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dot_ring_b200 import _native
from oracle import fr, ring_proof as rp
from tests import verify_cases as cases
from tests.helpers import bench_ring_keys, le64
from tests.ring_fixtures import native_ring, native_srs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
if os.environ.get("DR_LIB"):
    _native.set_default_library(_native.Library(os.environ["DR_LIB"]))
ctx = _native.Context(0)
srs = native_srs(ctx, None, 10)
pk, sk, keys = bench_ring_keys(1023)
ring = native_ring(srs, keys, rp.Params.from_ring_size(1023))
rng = random.Random(0)
al = [b"bench-batch-input" + le64(j) for j in range(n)]
ad = [b"bench-batch-ad" + le64(j) for j in range(n)]
proofs, status = ring.prove_batch(al, ad, [sk] * n, [3] * n, zk_rows=[rng.randrange(fr.R) for _ in range(12 * n)])
for agg in (False, True):
    co = cases.coeffs_for(n, 1, independent=not agg)
    times = []
    for rep in range(5):
        t0 = time.perf_counter()
        v, ok = ring.verify_batch(al, ad, proofs, co, aggregate=agg)
        times.append(round((time.perf_counter() - t0) * 1e3, 1))
    assert ok
    dt = min(times) * 1e-3
    print("aggregate" if agg else "per-item", n, times, "best %.0f verifies/s" % (n / dt), flush=True)
