"""max_ring_size below the domain's capacity (the reference's tests use 1..6 at domain 512 / 2048): the public vector then
holds more than 253 blinding-base rows (members.py:46-51) and the accumulator columns stay constant after the bit rows
(columns.py:111-146).  Shared by the CPU-emulation suite and the GPU suite: proofs must equal the oracle's byte for byte
and verify on both sides."""

from __future__ import annotations

import random

from oracle import ring_proof as rp
from oracle import vrf as ovrf
from tests.helpers import hx, load, split_keys


def check_small_max_ring(api, cases=((512, 100), (512, 1)), blinded: bool = True) -> None:
    v = load("bandersnatch_sha-512_ell2_ring.json")[0]
    all_keys = split_keys(hx(v, "ring_pks"))
    pk, sk = hx(v, "pk"), hx(v, "sk")
    cls = api.RingVRF[api.Bandersnatch]
    for domain, max_ring in cases:
        keys = ([pk] + [k for k in all_keys if k != pk])[:max_ring]
        for test_vectors in ([True, False] if blinded else [True]):
            params = api.RingProofParams(domain_size=domain, max_ring_size=max_ring, test_vectors=test_vectors)
            oparams = rp.Params(domain_size=domain, max_ring_size=max_ring, test_vectors=test_vectors)
            assert params.max_ring_size == oparams.max_ring_size == max_ring
            ring = api.Ring(keys, params)
            root = api.RingRoot.from_ring(ring, params)
            oring = rp.Ring(keys, oparams)
            oroot = rp.RingRoot.from_ring(oring, oparams)
            assert root.encode() == oroot.encode()
            assert [tuple(p) for p in ring.nm_points] == [tuple(p) for p in oring.nm_points]
            rng = random.Random(domain + max_ring)
            zk = None if test_vectors else [rng.randrange(params.prime) for _ in range(12)]
            alpha, ad = b"small-ring" + bytes([max_ring & 0xff]), b"ad"
            proof = cls.prove(alpha, ad, sk, pk, ring, root, zk_rows=zk)
            expect = ovrf.ring_prove(alpha, ad, sk, pk, oring, oroot, zk_rows=zk)
            assert proof.encode() == bytes(expect.encode()), (domain, max_ring, test_vectors)
            assert proof.verify(alpha, ad, ring, root)
            assert not proof.verify(alpha, b"other", ring, root)
            assert ovrf.ring_verify(ovrf.RingVrfProof.decode(proof.encode()), alpha, ad, oring, oroot, ring_matches=True)
