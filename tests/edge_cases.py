"""Edge cases the reference's tests exercise (empty / ragged inputs, capacity limits, malformed keys), shared by the CPU
(emulation) and GPU suites.  Each case compares the C-ABI path with the CPU oracle on the same inputs."""

from __future__ import annotations

import hashlib
import random

from oracle import bandersnatch as bs
from oracle import fr
from oracle import ring_proof as rp
from oracle import transcript as tr
from oracle import vrf as ovrf
from tests import verify_cases as cases
from tests.ring_fixtures import native_ring


def _keys(n: int, tag: bytes = b"edge"):
    out = []
    for i in range(n):
        pk, sk = tr.secret_from_seed(bs.SHA512, hashlib.sha256(tag + i.to_bytes(4, "little")).digest())
        out.append((pk, sk))
    return out


def ragged_inputs_match_oracle(srs, domain_ring_sizes=((512, 5),), n_items: int = 4):
    """alpha / ad of length 0, 1, 111 (crosses a SHA-512 block), 300; signer anywhere in the ring; blinded rows: byte parity
    with the oracle prover and acceptance by both verifiers."""
    rng = random.Random(21)
    shapes = [(b"", b""), (b"\x00", b"x" * 111), (b"a" * 111, b""), (b"long-alpha" * 30, b"long-ad" * 43)]
    for _, ring_size in domain_ring_sizes:
        params = rp.Params.from_ring_size(ring_size)
        pairs = _keys(ring_size, b"ragged%d" % ring_size)
        keys = [pk for pk, _ in pairs]
        oring = rp.Ring(keys, params)
        oroot = rp.RingRoot.from_ring(oring, params)
        ring = native_ring(srs, keys, params)
        assert ring.root() == oroot.encode()
        items = []
        for j in range(n_items):
            alpha, ad = shapes[j % len(shapes)]
            signer = rng.randrange(ring_size)
            items.append((alpha, ad, signer, [rng.randrange(fr.R) for _ in range(12)]))
        proofs, status = ring.prove_batch([i[0] for i in items], [i[1] for i in items], [pairs[i[2]][1] for i in items], [i[2] for i in items],
                                          zk_rows=[v for i in items for v in i[3]])  # fmt: skip
        assert status == [0] * len(items)
        for (alpha, ad, signer, zk), got in zip(items, proofs):
            want = ovrf.ring_prove(alpha, ad, pairs[signer][1], pairs[signer][0], oring, oroot, zk_rows=zk).encode()
            assert got == want, (ring_size, signer, len(alpha), len(ad))
        verdicts, ok = ring.verify_batch([i[0] for i in items], [i[1] for i in items], proofs, cases.coeffs_for(len(items)))
        assert ok and verdicts == [1] * len(items)
        ring.close()


def empty_batches(ctx, srs):
    params = rp.Params(test_vectors=True)
    pairs = _keys(3, b"empty")
    ring = native_ring(srs, [pk for pk, _ in pairs], params)
    assert ring.prove_batch([], [], [], []) == ([], [])
    assert ring.verify_batch([], [], [], []) == ([], True)
    assert ring.verify_batch([], [], [], [], aggregate=True) == ([], True)
    su = cases.suite_struct()
    assert ctx.pedersen_verify(su, [], [], []) == []
    assert ctx.tiny_verify(su, [], [], [], []) == []
    assert ctx.vrf_prove("tiny", su, [], [], []) == []
    assert ctx.te_decode([]) == [] and ctx.te_mul([bs.point_to_string(bs.GENERATOR)], []) == []
    assert srs.commit([]) == []
    ring.close()


def ring_capacity_and_bad_keys(srs):
    """members.py:22-55: more keys than max_ring_size is an error; undecodable, identity, small-order and off-subgroup keys
    become the padding point; the root equals the oracle's for the same key list."""
    params = rp.Params(test_vectors=True)  # N = 512, max ring 255
    pairs = _keys(6, b"cap")
    keys = [pk for pk, _ in pairs]
    bad = [b"\xff" * 32, bs.point_to_string(bs.IDENTITY), bytes(32)]
    mixed = [keys[0], bad[0], keys[1], bad[1], bad[2], keys[2]]
    ring = native_ring(srs, mixed, params)
    oring = rp.Ring(mixed, params)
    assert tuple(ring.points()) == oring.nm_points
    assert ring.root() == rp.RingRoot.from_ring(oring, params).encode()
    pad = params.suite.padding_point
    assert ring.points()[1] == pad and ring.points()[3] == pad and ring.points()[4] == pad
    ring.close()
    try:
        native_ring(srs, keys * 43, params)  # 258 > 255
    except ValueError:
        pass
    else:
        raise AssertionError("oversized ring accepted")
    full = native_ring(srs, (keys * 43)[:255], params)  # exactly max_ring_size
    assert full.root() == rp.RingRoot.from_ring(rp.Ring((keys * 43)[:255], params), params).encode()
    # signer in the last row of a full ring
    alpha, ad = b"full-ring", b""
    idx = 254
    sk = pairs[idx % 6][1]
    # duplicate keys: the reference proves at the FIRST matching row (members.py:71-81); the ABI takes the row explicitly
    proofs, status = full.prove_batch([alpha], [ad], [sk], [idx])
    assert status == [0]
    v, ok = full.verify_batch([alpha], [ad], proofs, cases.coeffs_for(1))
    assert ok and v == [1]
    full.close()


def te_msm_matches_oracle(ctx):
    """`BandersnatchPoint.msm` sizes 0, 1, 2, 3 (GLV joint-window paths in the reference), 5 and 200 (its Pippenger path);
    negative / zero / over-order scalars; an undecodable point is a ValueError."""
    rng = random.Random(17)
    base = [bs.mul(bs.GENERATOR, rng.randrange(1, bs.N)) for _ in range(200)]
    enc = [bs.point_to_string(p) for p in base]
    for n in (0, 1, 2, 3, 5, 200):
        ks = [rng.randrange(bs.N) for _ in range(n)]
        if n >= 3:
            ks[0], ks[1], ks[2] = 0, bs.N - 1, bs.N + 5
        want = bs.msm(base[:n], [k % bs.N for k in ks]) if n else bs.IDENTITY
        assert ctx.te_msm(enc[:n], ks) == bs.point_to_string(want), n
    try:
        ctx.te_msm([b"\xff" * 32], [1])
    except ValueError:
        pass
    else:
        raise AssertionError("undecodable point accepted")
