#!/usr/bin/env python
"""Round-2 golden vectors, produced by running the UNMODIFIED reference in the build container
(same mechanism as generate_golden.py: reference Python from /root/reference, its Cython extensions from
oracle/_ref/ext, blst / py_ecc / gmpy2 replaced by oracle/ref_shims).

    ./oracle/build_ref.sh && python tests/golden/generate_golden_r2.py

Writes tests/golden/pcs_reference.json (KZG.open / verify / batch_verify, RingRoot.verifier_transcript_prefix)
and tests/golden/small_ring_reference.json (max_ring_size below the domain capacity).
"""

from __future__ import annotations

import json
import random
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import ref_shims  # noqa: E402

OUT = Path(__file__).resolve().parent


def main() -> None:
    if not ref_shims.available():
        raise SystemExit("reference or oracle/_ref/ext missing: run ./oracle/build_ref.sh where /root/reference is mounted")
    ref_shims.install()
    import dot_ring.ring_proof.columns.columns as columns_mod
    from dot_ring.curve.specs.bandersnatch import Bandersnatch
    from dot_ring.ring_proof.params import RingProofParams
    from dot_ring.ring_proof.pcs.kzg import KZG
    from dot_ring.ring_proof.pcs.utils import LinearPcsVerification
    from dot_ring.vrf.ring.members import Ring
    from dot_ring.vrf.ring.root import RingRoot
    from dot_ring.vrf.ring.vrf import RingVRF

    prime = Bandersnatch.curve.params.field_modulus
    t0 = time.time()
    ser = lambda p: KZG.serialize_g1_uncompressed(p).hex()  # noqa: E731

    # ---- 1. PCS seam ---------------------------------------------------------------------------------
    pcs = {"open": [], "verify": [], "batch_verify": [], "linear": []}
    rng = random.Random(2024)
    opened = []
    for n in (1, 2, 17, 64):
        coeffs = [rng.randrange(prime) for _ in range(n)]
        x = rng.randrange(prime)
        c = KZG.commit(coeffs)
        o = KZG.open(coeffs, x)
        ok = bool(KZG.verify(c, o.proof, x, o.y))
        bad = bool(KZG.verify(c, o.proof, x, (o.y + 1) % prime))
        pcs["open"].append({"coeffs": [hex(v) for v in coeffs], "x": hex(x), "commitment": ser(c), "proof": ser(o.proof), "y": hex(o.y)})
        pcs["verify"].append({"commitment": ser(c), "proof": ser(o.proof), "x": hex(x), "y": hex(o.y), "valid": ok, "valid_with_y_plus_1": bad})
        opened.append((c, o.proof, x, o.y))
    pcs["batch_verify"].append({"items": [0, 1, 2, 3], "valid": bool(KZG.batch_verify(opened))})
    tampered = list(opened)
    tampered[2] = (opened[2][0], opened[2][1], opened[2][2], (opened[2][3] + 5) % prime)
    pcs["batch_verify"].append({"items": [0, 1, 2, 3], "tamper_item": 2, "y_delta": 5, "valid": bool(KZG.batch_verify(tampered))})
    # linear form: commitment 2 expressed as 3 * C_2 - 2 * C_2
    lin = [LinearPcsVerification(((opened[2][0], 3), (opened[2][0], prime - 2)), opened[2][1], opened[2][2], opened[2][3]),
           LinearPcsVerification(((opened[3][0], 1),), opened[3][1], opened[3][2], opened[3][3])]
    pcs["linear"].append({"terms": [[[2, 3], [2, -2]], [[3, 1]]], "valid": bool(KZG.batch_verify_linear_preconverted(lin))})
    print("pcs", time.time() - t0)

    # ---- 2. verifier transcript prefix (root.py:54-71) on ark-vrf vector 1 -----------------------------------
    v = json.loads((OUT / "reference_vectors" / "bandersnatch_sha-512_ell2_ring.json").read_text())[0]
    keys8 = RingVRF[Bandersnatch].parse_keys(bytes.fromhex(v["ring_pks"]))
    params8 = RingProofParams(test_vectors=True)
    ring8 = Ring(keys8, params8)
    root8 = RingRoot.from_ring(ring8, params8)
    tr = root8.verifier_transcript_prefix()
    pcs["verifier_transcript_prefix"] = {
        "ring_root": root8.encode().hex(),
        "challenge_label": "golden",
        "challenge": hex(tr.copy().challenge(b"golden")),
        "two_challenges": [hex(c) for c in tr.copy().challenges(b"pair", 2)],
        "custom_label_challenge": hex(root8.verifier_transcript_prefix(b"w3f-ring-proof-test").challenge(b"golden")),
    }
    (OUT / "pcs_reference.json").write_text(json.dumps(pcs, indent=1))
    print("prefix", time.time() - t0)

    # ---- 3. max_ring_size below the capacity of the domain ---------------------------------------------------
    pk, sk = bytes.fromhex(v["pk"]), bytes.fromhex(v["sk"])
    small = []
    stream = random.Random(7)
    draws: list[int] = []

    class _Secrets:
        @staticmethod
        def randbelow(n):
            val = stream.randrange(n)
            draws.append(val)
            return val

    columns_mod.secrets = _Secrets
    for domain, max_ring, tv in ((512, 100, True), (512, 100, False), (512, 1, True), (2048, 6, True)):
        keys = ([pk] + [k for k in keys8 if k != pk])[:max_ring]
        params = RingProofParams(domain_size=domain, max_ring_size=max_ring, test_vectors=tv)
        ring = Ring(keys, params)
        root = RingRoot.from_ring(ring, params)
        alpha, ad = b"small-ring" + bytes([max_ring & 0xFF]), b"ad"
        start = len(draws)
        proof = RingVRF[Bandersnatch].prove(alpha, ad, sk, pk, ring, root)
        assert proof.verify(alpha, ad, ring, root)
        small.append({"domain_size": domain, "max_ring_size": max_ring, "test_vectors": tv, "keys": [k.hex() for k in keys], "pk": pk.hex(), "sk": sk.hex(),
                      "alpha": alpha.hex(), "ad": ad.hex(), "zk_rows": [hex(d) for d in draws[start:]], "ring_root": root.encode().hex(), "proof": proof.encode().hex()})
        print("small", domain, max_ring, tv, time.time() - t0)
    (OUT / "small_ring_reference.json").write_text(json.dumps(small, indent=1))
    print("done", time.time() - t0)


if __name__ == "__main__":
    main()
