"""Domains above 4096 (outside the reference, which rejects them: ring_proof/params.py:172-173): a synthetic SRS tau^i * G with
3N + 1 points, the engine's large-domain route against the CPU oracle run with the same SRS, and prove -> verify round trips."""

from __future__ import annotations

import hashlib
import random
import time

from dot_ring_b200 import _native
from oracle import bandersnatch as bs
from oracle import bls12_381 as bls
from oracle import fr
from oracle import ring_proof as rp
from oracle import transcript as tr
from oracle import vrf as ovrf
from tests import verify_cases as cases
from tests.msm_cases import TAU
from tests.ring_fixtures import native_ring


def synthetic_srs(ctx, n_points: int, window_bits: int):
    """(NativeSrs, oracle SRS) over tau^i * G, i < n_points, and [1]_2, [tau]_2."""
    g1 = ctx.g1_synthetic_srs(TAU, 0, n_points)
    g2_gen = rp.load_srs().g2[0]
    g2 = [g2_gen, bls.g2_mul(g2_gen, TAU)]
    g2_bytes = b"".join(bls.g2_serialize(p) for p in g2)
    native = _native.NativeSrs(ctx, g1, g2_bytes, window_bits)
    oracle = rp.SRS([(int.from_bytes(g1[96 * i : 96 * i + 48], "big"), int.from_bytes(g1[96 * i + 48 : 96 * i + 96], "big")) for i in range(n_points)], g2)
    return native, oracle


def prove_verify_against_oracle(ctx, domain: int, n_keys: int, n_proofs: int, oracle_proofs: int, window_bits: int = 8, log=None):
    """Ring of `n_keys` real keys inside a `domain`-row ring (the rest is padding): root and the first `oracle_proofs` proofs
    byte-identical to the oracle with the same SRS; every proof accepted by the device verifier, per item and aggregated."""
    say = log or (lambda *a: None)
    t0 = time.time()
    srs, osrs = synthetic_srs(ctx, 3 * domain + 1, window_bits)
    say("srs + table", round(time.time() - t0, 1), "s; table GB", round(srs.table_bytes / 1e9, 1))
    params = rp.Params(domain_size=domain, max_ring_size=domain - rp.SCALAR_BITS - 4, max_domain_size=max(domain, 4096))
    pairs = [tr.secret_from_seed(bs.SHA512, hashlib.sha256(b"ext%d" % i).digest()) for i in range(min(n_keys, 64))]
    keys = [pk for pk, _ in pairs]
    if n_keys > len(keys):  # bulk keys: multiples of the generator computed on the device
        rng = random.Random(5)
        extra = ctx.te_mul([bs.point_to_string(bs.GENERATOR)], [rng.randrange(1, bs.N) for _ in range(n_keys - len(keys))])
        keys += extra
    t0 = time.time()
    ring = native_ring(srs, keys, params)
    say("ring create", round(time.time() - t0, 2), "s for", len(keys), "keys")
    rng = random.Random(11)
    signers = [rng.randrange(len(pairs)) for _ in range(n_proofs)]
    alphas = [b"ext-input-%d" % j for j in range(n_proofs)]
    ads = [b"ext-ad-%d" % j for j in range(n_proofs)]
    zk = [rng.randrange(fr.R) for _ in range(12 * n_proofs)]
    t0 = time.time()
    proofs, status = ring.prove_batch(alphas, ads, [pairs[s][1] for s in signers], signers, zk_rows=zk)
    dt = time.time() - t0
    assert status == [0] * n_proofs
    say("prove", n_proofs, "proofs", round(dt, 2), "s; phases ms", [round(x, 1) for x in ring.prove_phase_ms()])
    t0 = time.time()
    verdicts, ok = ring.verify_batch(alphas, ads, proofs, cases.coeffs_for(n_proofs))
    assert ok and verdicts == [1] * n_proofs
    assert ring.verify_batch(alphas, ads, proofs, cases.coeffs_for(n_proofs, 2, independent=False), aggregate=True)[1]
    bad = bytearray(proofs[0])
    bad[192 + 300] ^= 1
    assert ring.verify_batch(alphas[:1], ads[:1], [bytes(bad)], cases.coeffs_for(1))[0] == [0]
    say("verify", round(time.time() - t0, 2), "s")
    out = {"domain": domain, "keys": len(keys), "proofs": n_proofs, "prove_s": dt, "prove_phase_ms": ring.prove_phase_ms()}
    if oracle_proofs:
        t0 = time.time()
        oring = rp.Ring(keys, params)
        oroot = rp.RingRoot.from_ring(oring, params, osrs)
        assert ring.root() == oroot.encode(), "ring root differs from the oracle"
        for j in range(oracle_proofs):
            pk, sk = pairs[signers[j]]
            want = ovrf.ring_prove(alphas[j], ads[j], sk, pk, oring, oroot, zk_rows=zk[12 * j : 12 * j + 12]).encode()
            assert proofs[j] == want, f"proof {j} differs from the oracle"
        out["oracle_s"] = time.time() - t0
        say("oracle root +", oracle_proofs, "proofs identical;", round(out["oracle_s"], 1), "s of CPU")
    ring.close()
    srs.close()
    return out
