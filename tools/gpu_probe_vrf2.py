#!/usr/bin/env python
"""Repeatability probe for the Tiny / Pedersen batch verifiers (wall time of the C-ABI call only, inputs pre-packed)."""
import ctypes, hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dot_ring_b200 import _native
from tests import verify_cases as cases
from tests.helpers import le64
if os.environ.get("DR_LIB"):
    _native.set_default_library(_native.Library(os.environ["DR_LIB"]))
ctx = _native.Context(0)
su = cases.suite_struct()
L = ctx.library.lib
for n in (20000, 100000):
    sks = [hashlib.sha256(b"ietf-signer" + le64(i)).digest()[:31] + b"\x00" for i in range(n)]
    alphas = [b"bench-ietf-input" + le64(i) for i in range(n)]
    ads = [b"bench-ietf-ad" + le64(i) for i in range(n)]
    gen = cases.bs.point_to_string(cases.bs.GENERATOR)
    pks = ctx.te_mul([gen], [int.from_bytes(k, "little") for k in sks])
    blob, a, b, c, d = _native.pack_items(alphas, ads)
    out = ctypes.create_string_buffer(n)
    for kind in ("tiny", "pedersen"):
        proofs = b"".join(ctx.vrf_prove(kind, su, alphas, ads, sks))
        times = []
        for rep in range(4):
            t0 = time.perf_counter()
            if kind == "tiny":
                rc = L.dr_tiny_verify_batch(ctx.handle, ctypes.byref(su), n, blob, a, b, c, d, b"".join(pks), proofs, out)
            else:
                rc = L.dr_pedersen_verify_batch(ctx.handle, ctypes.byref(su), n, blob, a, b, c, d, proofs, out)
            times.append(round((time.perf_counter() - t0) * 1e3, 1))
            assert rc == 0 and out.raw == b"\x01" * n
        print(os.environ.get("DR_LIB", "default")[-14:], kind, n, "ms:", times, "best/s:", int(n / (min(times) * 1e-3)), flush=True)
