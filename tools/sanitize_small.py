#!/usr/bin/env python
"""Small end-to-end pass for compute-sanitizer memcheck: SRS table (4-bit windows), ring 8 prove + verify, VRF verify, MSM, NTT."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dot_ring_b200 import _native
from oracle import fr, ring_proof as rp
from tests import msm_cases, verify_cases as cases
from tests.helpers import hx, load, ring_proof_bytes, split_keys
from tests.ring_fixtures import native_ring, native_srs
lib = _native.Library(os.environ["DR_EMUL_LIB"], require_cuda=False) if os.environ.get("DR_EMUL_LIB") else None  # ASan build of the emulation (CPU)
ctx = _native.Context(0, lib)
srs = native_srs(ctx, 1537, 4)
v = load("bandersnatch_sha-512_ell2_ring.json")[0]
keys = split_keys(hx(v, "ring_pks"))
params = rp.Params(test_vectors=True)
ring = native_ring(srs, keys, params)
assert ring.root().hex() == v["ring_pks_com"]
k = rp.Ring(keys, params).index_of(hx(v, "pk"))
proofs, st = ring.prove_batch([hx(v, "alpha")] * 3, [hx(v, "ad")] * 3, [hx(v, "sk")] * 3, [k] * 3)
assert st == [0, 0, 0] and proofs[0] == ring_proof_bytes(v)
ctx.set_dense_witness_commit(True)
assert ring.prove_batch([hx(v, "alpha")], [hx(v, "ad")], [hx(v, "sk")], [k])[0][0] == proofs[0]
ctx.set_dense_witness_commit(False)
assert ring.verify_batch([hx(v, "alpha")] * 3, [hx(v, "ad")] * 3, proofs, cases.coeffs_for(3)) == ([1, 1, 1], True)
assert ring.verify_batch([hx(v, "alpha")] * 3, [hx(v, "ad")] * 3, proofs, cases.coeffs_for(3, 2, False), aggregate=True)[1]
cases.pedersen_vectors(ctx)
cases.tiny_vectors(ctx)
msm_cases.synthetic_property(ctx, [300], (0, 2))
rng = random.Random(1)
vals = [rng.randrange(fr.R) for _ in range(512)]
assert ctx.fr_ntt(ctx.fr_ntt(vals, 512, params.omega), 512, params.omega, inverse=True) == vals
print("sanitize_small ok, launches", ctx.library.launch_count())
