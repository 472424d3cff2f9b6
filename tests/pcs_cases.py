"""PCS seam (pcs/protocol.py:33-40) and RingRoot.verifier_transcript_prefix (root.py:54-71) against vectors produced by the
unmodified reference (tests/golden/pcs_reference.json, generate_golden_r2.py).  Shared by the emulation and GPU suites."""

from __future__ import annotations

from tests.helpers import hx, load, split_keys

H = lambda s: int(s, 16)  # noqa: E731


def check_pcs(api) -> None:
    from dot_ring_b200.kzg import LinearPcsVerification, PcsVerification

    g = load("pcs_reference.json")
    KZG = api.KZG
    prime = KZG.scalar_modulus
    opened = []
    for o in g["open"]:
        coeffs, x = [H(c) for c in o["coeffs"]], H(o["x"])
        assert KZG.commit(coeffs).hex() == o["commitment"]
        got = KZG.open(coeffs, x)
        assert got.proof.hex() == o["proof"] and got.y == H(o["y"]), len(coeffs)
        opened.append(PcsVerification(bytes.fromhex(o["commitment"]), got.proof, x, got.y))
    # batched opening gives the same proofs
    same_len = [o for o in g["open"] if len(o["coeffs"]) == 17] * 2
    both = KZG.open_batch([[H(c) for c in o["coeffs"]] for o in same_len], [H(o["x"]) for o in same_len])
    assert [b.proof.hex() for b in both] == [o["proof"] for o in same_len]
    for v, ver in zip(g["verify"], opened):
        assert KZG.verify(*ver) is v["valid"] is True
        assert KZG.verify(ver.commitment, ver.proof, ver.point, (ver.value + 1) % prime) is v["valid_with_y_plus_1"] is False
    assert KZG.batch_verify([]) is True
    for case in g["batch_verify"]:
        items = [opened[i] for i in case["items"]]
        if "tamper_item" in case:
            t = items[case["tamper_item"]]
            items[case["tamper_item"]] = PcsVerification(t.commitment, t.proof, t.point, (t.value + case["y_delta"]) % prime)
        assert KZG.batch_verify(items) is case["valid"]
    for case in g["linear"]:
        vers = [LinearPcsVerification(tuple((opened[i].commitment, k % prime) for i, k in terms), opened[terms[0][0]].proof, opened[terms[0][0]].point, opened[terms[0][0]].value)
                for terms in case["terms"]]  # fmt: skip
        assert KZG.batch_verify_linear_preconverted(vers) is case["valid"]
        bad = [vers[0]._replace(value=(vers[0].value + 1) % prime)] + vers[1:]
        assert KZG.batch_verify_linear_preconverted(bad) is False
    # malformed point -> ValueError like the reference's g1_to_blst
    import pytest

    with pytest.raises(ValueError):
        KZG.verify(b"\x01" * 96, opened[0].proof, 1, 1)


def check_transcript_prefix(api) -> None:
    g = load("pcs_reference.json")["verifier_transcript_prefix"]
    v = load("bandersnatch_sha-512_ell2_ring.json")[0]
    params = api.RingProofParams(test_vectors=True)
    ring = api.Ring(split_keys(hx(v, "ring_pks")), params)
    root = api.RingRoot.from_ring(ring, params)
    assert root.encode().hex() == g["ring_root"]
    tr = root.verifier_transcript_prefix()
    assert tr.copy().challenge(g["challenge_label"].encode()) == H(g["challenge"])
    assert tr.copy().challenges(b"pair", 2) == [H(c) for c in g["two_challenges"]]
    assert root.verifier_transcript_prefix(b"w3f-ring-proof-test").challenge(b"golden") == H(g["custom_label_challenge"])
    # a decoded root (no ring attached) gives the same prefix
    assert api.RingRoot.decode(root.encode(), params).verifier_transcript_prefix().challenge(b"golden") == H(g["challenge"])


def check_small_ring_golden(api, domains=(512,)) -> None:
    """Proofs of the unmodified reference at max_ring_size below the domain capacity (tests/golden/small_ring_reference.json)."""
    cls = api.RingVRF[api.Bandersnatch]
    for g in load("small_ring_reference.json"):
        if g["domain_size"] not in domains:
            continue
        params = api.RingProofParams(domain_size=g["domain_size"], max_ring_size=g["max_ring_size"], test_vectors=g["test_vectors"])
        keys = [bytes.fromhex(k) for k in g["keys"]]
        ring = api.Ring(keys, params)
        root = api.RingRoot.from_ring(ring, params)
        assert root.encode().hex() == g["ring_root"]
        zk = [H(z) for z in g["zk_rows"]] or None
        proof = cls.prove(bytes.fromhex(g["alpha"]), bytes.fromhex(g["ad"]), bytes.fromhex(g["sk"]), bytes.fromhex(g["pk"]), ring, root, zk_rows=zk)
        assert proof.encode().hex() == g["proof"], (g["domain_size"], g["max_ring_size"], g["test_vectors"])
        assert cls.decode(bytes.fromhex(g["proof"])).verify(bytes.fromhex(g["alpha"]), bytes.fromhex(g["ad"]), ring, root)
