#!/usr/bin/env python
"""Where does a prove call spend device time outside its kernels?  Runs bench-like calls with DOT_RING_B200_DEBUG=1 (the library prints
copy-in / copy-out / side-stream spans) and the event-timed call duration next to the phase sum.  python tools/gpu_probe_timeline.py [window_bits]"""
import os, random, sys, time
os.environ["DOT_RING_B200_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dot_ring_b200 import _native
from oracle import fr, ring_proof as rp
from tests.helpers import bench_ring_keys, le64
from tests.ring_fixtures import native_ring, native_srs

wb = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ctx = _native.Context(0)
srs = native_srs(ctx, None, wb)
pk, sk, keys = bench_ring_keys(1023)
ring = native_ring(srs, keys, rp.Params.from_ring_size(1023))
ring.time_calls = True
rng = random.Random(0)
for n in (1, 512, 1024, 2048, 4096, 512, 4096):
    al = [b"bench-batch-input" + le64(j) for j in range(n)]
    ad = [b"bench-batch-ad" + le64(j) for j in range(n)]
    zk = b"".join(rng.randrange(fr.R).to_bytes(32, "little") for _ in range(12 * n))
    for rep in range(3):
        t0 = time.perf_counter()
        proofs, status = ring.prove_batch(al, ad, [sk] * n, [3] * n, zk_rows=zk)
        wall = (time.perf_counter() - t0) * 1e3
        ph = ring.prove_phase_ms()
        print(f"n={n} rep={rep}: wall {wall:.2f} ms, event pair {ring.last_call_ms:.2f} ms, phase sum {sum(ph):.2f} ms, phases {[round(x, 2) for x in ph]}", flush=True)
