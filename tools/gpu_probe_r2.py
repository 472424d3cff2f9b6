#!/usr/bin/env python
"""Round-2 probe for launch lists and ncu captures: one ring-1023 proving pass and Tiny / Pedersen / Ring verification batches.

    python tools/gpu_probe_r2.py [proofs] [verify_items] [window_bits]

Prints wall times per call; the kernels of interest for `ncu -k regex:`: PedersenProve, Witness(Body|Coop), WitnessLde,
Constraint, IetfVerify, PedersenVerify, RingVerify*."""
import os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dot_ring_b200 import _native
from oracle import fr, ring_proof as rp
from tests import verify_cases as cases
from tests.helpers import bench_ring_keys, le64
from tests.ring_fixtures import native_ring, native_srs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
wb = int(sys.argv[3]) if len(sys.argv) > 3 else 10
ctx = _native.Context(0)
srs = native_srs(ctx, None, wb)
pk, sk, keys = bench_ring_keys(1023)
ring = native_ring(srs, keys, rp.Params.from_ring_size(1023))
rng = random.Random(0)
al = [b"bench-batch-input" + le64(j) for j in range(n)]
ad = [b"bench-batch-ad" + le64(j) for j in range(n)]
zk = [rng.randrange(fr.R) for _ in range(12 * n)]
for rep in range(3):
    t0 = time.perf_counter()
    proofs, status = ring.prove_batch(al, ad, [sk] * n, [3] * n, zk_rows=zk)
    dt = time.perf_counter() - t0
    print(f"prove {n}: {dt * 1e3:.1f} ms, phases {[round(x, 2) for x in ring.prove_phase_ms()]}", flush=True)
assert status == [0] * n
one = ring.prove_batch(al[:1], ad[:1], [sk], [3], zk_rows=zk[:12])
for rep in range(3):
    t0 = time.perf_counter()
    ring.prove_batch(al[:1], ad[:1], [sk], [3], zk_rows=zk[:12])
    print(f"prove 1: {(time.perf_counter() - t0) * 1e3:.2f} ms, phases {[round(x, 2) for x in ring.prove_phase_ms()]}", flush=True)
m = min(n, 1024)
for agg in (False, True):
    co = cases.coeffs_for(m, 1, independent=not agg)
    for rep in range(3):
        t0 = time.perf_counter()
        v, ok = ring.verify_batch(al[:m], ad[:m], proofs[:m], co, aggregate=agg)
        dt = time.perf_counter() - t0
    assert ok
    print(f"ring verify {'aggregate' if agg else 'per-item'} {m}: {dt * 1e3:.1f} ms", flush=True)
for rep in range(3):
    t0 = time.perf_counter()
    v, ok = ring.verify_batch(al[:1], ad[:1], proofs[:1], cases.coeffs_for(1, 1))
    print(f"ring verify 1: {(time.perf_counter() - t0) * 1e3:.2f} ms", flush=True)
su = cases.suite_struct()
sks = [(1 + 7919 * i).to_bytes(32, "little") for i in range(nv)]
ins = [b"bench-ietf-input" + le64(i) for i in range(nv)]
ads = [b"bench-ietf-ad" + le64(i) for i in range(nv)]
pks = [p for p in ctx.te_mul([__import__("oracle.bandersnatch", fromlist=["x"]).point_to_string(__import__("oracle.bandersnatch", fromlist=["x"]).GENERATOR)], [int.from_bytes(s, "little") for s in sks])]
tiny = ctx.vrf_prove("tiny", su, ins, ads, sks)
ped = ctx.vrf_prove("pedersen", su, ins, ads, sks)
thin = ctx.vrf_prove("thin", su, ins, ads, sks)
for name, fn in (("tiny", lambda: ctx.tiny_verify(su, pks, ins, ads, tiny)), ("thin", lambda: ctx.thin_verify(su, pks, ins, ads, thin)), ("pedersen", lambda: ctx.pedersen_verify(su, ins, ads, ped))):
    for rep in range(3):
        t0 = time.perf_counter()
        v = fn()
        dt = time.perf_counter() - t0
    assert v == [1] * nv, name
    print(f"{name} verify {nv}: {dt * 1e3:.1f} ms = {nv / dt:.0f} /s", flush=True)
    one_t = []
