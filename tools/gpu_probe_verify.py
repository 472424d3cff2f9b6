#!/usr/bin/env python
"""GPU probe: verification throughput (BASELINE configs[3] and the verify half of configs[1]); writes gpurun_out/probe_verify.json."""
import hashlib
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dot_ring_b200 import _native  # noqa: E402
from oracle import fr, ring_proof as rp  # noqa: E402
from tests import verify_cases as cases  # noqa: E402
from tests.helpers import bench_ring_keys, le64  # noqa: E402
from tests.ring_fixtures import native_ring, native_srs  # noqa: E402

out = {}
if os.environ.get("DR_LIB"):
    _native.set_default_library(_native.Library(os.environ["DR_LIB"]))
ctx = _native.Context(0)
su = cases.suite_struct()
N_VRF = int(os.environ.get("N_VRF", "100000"))
sks = [hashlib.sha256(b"ietf-signer" + le64(i)).digest()[:31] + b"\x00" for i in range(N_VRF)]
alphas = [b"bench-ietf-input" + le64(i) for i in range(N_VRF)]
ads = [b"bench-ietf-ad" + le64(i) for i in range(N_VRF)]
gen = cases.bs.point_to_string(cases.bs.GENERATOR)
t0 = time.time()
pks = ctx.te_mul([gen], [int.from_bytes(k, "little") for k in sks])
out["keygen_s"] = time.time() - t0
for kind in ("tiny", "pedersen"):
    t0 = time.time()
    proofs = ctx.vrf_prove(kind, su, alphas, ads, sks)
    t_prove = time.time() - t0
    bad = [bytearray(p) for p in proofs]
    for i in range(0, N_VRF, 100):
        bad[i][-1 - 8] ^= 1
    bad = [bytes(b) for b in bad]
    for rep in range(2):
        t0 = time.time()
        ctx.timer_start()
        if kind == "tiny":
            v = ctx.tiny_verify(su, pks, alphas, ads, bad)
        else:
            v = ctx.pedersen_verify(su, alphas, ads, bad)
        dev_ms = ctx.timer_stop()
        dt = time.time() - t0
    n_bad = sum(1 for x in v if x != 1)
    assert n_bad == len(range(0, N_VRF, 100)), n_bad
    assert all((v[i] != 1) == (i % 100 == 0) for i in range(N_VRF))
    out[kind] = {"n": N_VRF, "prove_wall_s": t_prove, "verify_wall_s": dt, "verify_device_ms": dev_ms, "verifies_per_s": N_VRF / dt, "verifies_per_s_device": N_VRF / (dev_ms * 1e-3), "proves_per_s": N_VRF / t_prove, "rejected": n_bad}
    print(kind, out[kind], flush=True)

srs = native_srs(ctx, None, int(os.environ.get("DR_WINDOW_BITS", "12")))
pk, sk, keys = bench_ring_keys(1023)
params = rp.Params.from_ring_size(1023)
ring = native_ring(srs, keys, params)
rng = random.Random(0)
for n in (1, 64, 4096):
    zk = [rng.randrange(fr.R) for _ in range(12 * n)]
    al = [b"bench-batch-input" + le64(j) for j in range(n)]
    ad = [b"bench-batch-ad" + le64(j) for j in range(n)]
    proofs, status = ring.prove_batch(al, ad, [sk] * n, [3] * n, zk_rows=zk)
    assert not any(status)
    res = {"n": n}
    for agg in (False, True):
        co = cases.coeffs_for(n, 1, independent=not agg)
        for rep in range(2):
            t0 = time.time()
            ctx.timer_start()
            v, ok = ring.verify_batch(al, ad, proofs, co, aggregate=agg)
            dev_ms = ctx.timer_stop()
            dt = time.time() - t0
        assert ok and v == [1] * n
        res["aggregate" if agg else "per_item"] = {"wall_s": dt, "device_ms": dev_ms, "verifies_per_s": n / dt}
    out[f"ring_verify_{n}"] = res
    print(res, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_verify.json", "w"), indent=1)
