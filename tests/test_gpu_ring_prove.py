"""GPU parity: ring root and Ring VRF proofs, byte-for-byte against the reference's golden vectors
(tests/vectors/ark-vrf/bandersnatch_sha-512_ell2_ring.json; reference test tests/test_ark_vrf.py:118-132)
and against outputs of the unmodified reference at ring 1023 / N=2048 (tests/golden/ring1023_reference.json)."""

import hashlib
import random

import pytest

from oracle import fr
from oracle import ring_proof as rp
from oracle import vrf as ovrf
from tests.helpers import bench_ring_keys, hx, le64, load, ring_proof_bytes, split_keys
from tests.ring_fixtures import native_ring, native_srs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from dot_ring_b200 import _native

    c = _native.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def srs(ctx):
    s = native_srs(ctx, None, 10)  # full 6145-point SRS, 7.9 GB table
    yield s
    s.close()


def test_ark_vrf_ring_vectors_root_and_proof(srs):
    params = rp.Params(test_vectors=True)
    for v in load("bandersnatch_sha-512_ell2_ring.json"):
        keys = split_keys(hx(v, "ring_pks"))
        ring = native_ring(srs, keys, params)
        assert ring.root().hex() == v["ring_pks_com"]
        oring = rp.Ring(keys, params)
        assert tuple(ring.points()) == oring.nm_points
        k = oring.index_of(hx(v, "pk"))
        proofs, status = ring.prove_batch([hx(v, "alpha")], [hx(v, "ad")], [hx(v, "sk")], [k])
        assert status == [0]
        assert proofs[0] == ring_proof_bytes(v)
        ring.close()


def test_ring8_blinded_rows_match_reference(srs):
    g = load("ring8_reference.json")
    v = load("bandersnatch_sha-512_ell2_ring.json")[0]
    params = rp.Params()
    keys = split_keys(hx(v, "ring_pks"))
    ring = native_ring(srs, keys, params)
    k = rp.Ring(keys, params).index_of(hx(v, "pk"))
    proofs, status = ring.prove_batch([hx(v, "alpha")], [hx(v, "ad")], [hx(v, "sk")], [k], zk_rows=g["zk_rows"])
    assert status == [0] and proofs[0].hex() == g["proof"]
    ring.close()


def test_glv_table_gives_the_reference_root_and_proofs(ctx):
    """Same vectors through a table that covers 128 bits with the scalars split by the G1 endomorphism (the bench default)."""
    from dot_ring_b200 import _native
    from dot_ring_b200.srs import read_srs_file

    raw = read_srs_file(None, None)
    glv_srs = _native.NativeSrs(ctx, raw.g1_be96, raw.g2_be192, 12, 0, True)
    try:
        assert glv_srs.geometry == (12, 0, 1, 22)
        g = load("ring8_reference.json")
        for i, v in enumerate(load("bandersnatch_sha-512_ell2_ring.json")):
            keys = split_keys(hx(v, "ring_pks"))
            ring = native_ring(glv_srs, keys, rp.Params(test_vectors=True))
            assert ring.root().hex() == v["ring_pks_com"]
            k = rp.Ring(keys, rp.Params(test_vectors=True)).index_of(hx(v, "pk"))
            proofs, status = ring.prove_batch([hx(v, "alpha")], [hx(v, "ad")], [hx(v, "sk")], [k])
            assert status == [0] and proofs[0] == ring_proof_bytes(v)
            ring.close()
            if i == 0:
                ring = native_ring(glv_srs, keys, rp.Params())
                proofs, status = ring.prove_batch([hx(v, "alpha")], [hx(v, "ad")], [hx(v, "sk")], [k], zk_rows=g["zk_rows"])
                assert status == [0] and proofs[0].hex() == g["proof"]
                ring.close()
    finally:
        glv_srs.close()


def test_ring1023_root_and_proofs_match_reference(srs):
    g = load("ring1023_reference.json")
    pk, sk, keys = bench_ring_keys(1023)
    params = rp.Params.from_ring_size(1023, test_vectors=True)
    ring = native_ring(srs, keys, params)
    assert ring.root().hex() == g["ring_root"]
    n = len(g["proofs_test_vectors"])
    alphas = [bytes.fromhex(e["alpha"]) for e in g["proofs_test_vectors"]]
    ads = [bytes.fromhex(e["ad"]) for e in g["proofs_test_vectors"]]
    proofs, status = ring.prove_batch(alphas, ads, [sk] * n, [3] * n)
    assert status == [0] * n
    for e, p in zip(g["proofs_test_vectors"], proofs):
        assert p.hex() == e["proof"]
    # blinded rows (secrets.randbelow stream replayed)
    zk = [int(x, 16) for e in g["proofs_blinded"] for x in e["zk_rows"]]
    m = len(g["proofs_blinded"])
    proofs, status = ring.prove_batch(
        [bytes.fromhex(e["alpha"]) for e in g["proofs_blinded"]], [bytes.fromhex(e["ad"]) for e in g["proofs_blinded"]], [sk] * m, [3] * m, zk_rows=zk
    )
    assert status == [0] * m
    for e, p in zip(g["proofs_blinded"], proofs):
        assert p.hex() == e["proof"]
    # a wrong producer index / foreign secret key is reported per item, not silently proven
    other_sk = hashlib.sha256(b"not-a-member").digest()
    _, status = ring.prove_batch([b"x", b"y"], [b"", b""], [sk, other_sk], [4, 3])
    assert status[0] != 0 and status[1] != 0
    ring.close()


def test_ring1023_batch_proofs_verify_with_oracle(srs):
    """Full-size batch: every proof of a 64-proof batch (blinded rows from random.Random(0)) must verify
    under the CPU oracle's verifier for a sample, and proofs must differ across inputs."""
    pk, sk, keys = bench_ring_keys(1023)
    params = rp.Params.from_ring_size(1023)
    ring = native_ring(srs, keys, params)
    n = 64
    rng = random.Random(0)
    zk = [rng.randrange(fr.R) for _ in range(12 * n)]
    alphas = [b"bench-batch-input" + le64(j) for j in range(n)]
    ads = [b"bench-batch-ad" + le64(j) for j in range(n)]
    proofs, status = ring.prove_batch(alphas, ads, [sk] * n, [3] * n, zk_rows=zk)
    assert status == [0] * n and len(set(proofs)) == n
    oring = rp.Ring(keys, params)
    srs_o = rp.load_srs()
    # root from the device is the verifier key here (decode to oracle commitments)
    from oracle import bls12_381 as bls

    fixed = rp.decode_ring_root(ring.root())
    prefix = __import__("oracle.transcript", fromlist=["RingTranscript"]).RingTranscript(params.suite.suite_id)
    prefix.absorb_labeled(b"vk", rp.srs_vk_prefix(srs_o) + b"".join(bls.g1_serialize(c) for c in fixed))
    for j in (0, 17, 63):
        pr = ovrf.RingVrfProof.decode(proofs[j])
        assert ovrf.pedersen_verify(params.suite, pr.pedersen, alphas[j], ads[j])
        assert rp.verify_ring(params, fixed, prefix, pr.pedersen.blinded_pk, pr.ring, srs_o)
    assert oring.nm_points[3] == tuple(ring.points())[3]
    ring.close()


def test_sparse_and_dense_witness_commitments_agree_at_full_size(ctx, srs):
    """Same proofs from the Lagrange-prefix (sparse) witness commitments and from KZG.commit of the interpolated columns."""
    pk, sk, keys = bench_ring_keys(1023)
    params = rp.Params.from_ring_size(1023)
    ring = native_ring(srs, keys, params)
    n = 96
    rng = random.Random(3)
    zk = [rng.randrange(fr.R) for _ in range(12 * n)]
    alphas = [b"x" * (j % 7) + le64(j) for j in range(n)]
    ads = [b"ad" + le64(j) for j in range(n)]
    sparse, st1 = ring.prove_batch(alphas, ads, [sk] * n, [3] * n, zk_rows=zk)
    ctx.set_dense_witness_commit(True)
    try:
        dense, st2 = ring.prove_batch(alphas, ads, [sk] * n, [3] * n, zk_rows=zk)
    finally:
        ctx.set_dense_witness_commit(False)
    assert st1 == st2 == [0] * n and sparse == dense
    ctx.set_commit_mode(1)  # batched-affine summation of the table entries
    try:
        affine, st3 = ring.prove_batch(alphas, ads, [sk] * n, [3] * n, zk_rows=zk)
    finally:
        ctx.set_commit_mode(0)
    assert st3 == [0] * n and affine == sparse
    ring.close()
