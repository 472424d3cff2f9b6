"""GPU parity for the verification side (C ABI on cuda:0): pairing, Tiny / Pedersen verification, ring-proof verification
against the reference's vectors, plus prove -> verify round trips at ring 1023 / N = 2048 with per-item verdicts."""

import random

import pytest

from oracle import fr
from oracle import ring_proof as rp
from tests import verify_cases as cases
from tests.helpers import bench_ring_keys, le64, load
from tests.ring_fixtures import native_ring, native_srs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from dot_ring_b200 import _native

    c = _native.Context(0)
    assert c.library.is_cuda
    yield c
    c.close()


@pytest.fixture(scope="module")
def srs(ctx):
    s = native_srs(ctx, None, 10)
    yield s
    s.close()


def test_pairing(ctx):
    cases.pairing_kats(ctx)


def test_pedersen_vectors(ctx):
    cases.pedersen_vectors(ctx)


def test_tiny_vectors(ctx):
    cases.tiny_vectors(ctx)


def test_thin_vectors(ctx):
    cases.thin_vectors(ctx)


def test_vrf_batch_fixtures(ctx):
    cases.vrf_batch_fixtures(ctx)


def test_ring8_verify(srs):
    cases.ring8_vectors(srs)


def test_w3f_verifier_vectors(ctx):
    cases.w3f_vectors(ctx)


def test_ring1023_prove_then_verify_round_trip(srs):
    """512 blinded proofs at the headline shape: all verify individually and aggregated; 1 in 16 corrupted -> exactly
    those fail; reference-generated N=2048 proofs verify too."""
    pk, sk, keys = bench_ring_keys(1023)
    params = rp.Params.from_ring_size(1023)
    ring = native_ring(srs, keys, params)
    n = 512
    rng = random.Random(11)
    zk = [rng.randrange(fr.R) for _ in range(12 * n)]
    alphas = [b"bench-batch-input" + le64(j) for j in range(n)]
    ads = [b"bench-batch-ad" + le64(j) for j in range(n)]
    proofs, status = ring.prove_batch(alphas, ads, [sk] * n, [3] * n, zk_rows=zk)
    assert status == [0] * n
    verdicts, all_ok = ring.verify_batch(alphas, ads, proofs, cases.coeffs_for(n, 1))
    assert verdicts == [1] * n and all_ok
    assert ring.verify_batch(alphas, ads, proofs, cases.coeffs_for(n, 2, independent=False), aggregate=True) == ([1] * n, True)
    bad = [bytearray(p) for p in proofs]
    expect = []
    for j in range(n):
        if j % 16 == 5:
            field = (j // 16) % 4
            if field == 0:
                bad[j][192 + 250] ^= 1  # an evaluation
            elif field == 1:
                bad[j][140] ^= 1  # Pedersen s
            elif field == 2:
                bad[j][192 + 500] ^= 1  # opening proof x coordinate: off the curve or a different point
            else:
                ads[j] = b"tampered"
            expect.append(None)
        else:
            expect.append(1)
    verdicts, all_ok = ring.verify_batch(alphas, ads, [bytes(b) for b in bad], cases.coeffs_for(n, 3))
    assert not all_ok
    for got, want in zip(verdicts, expect):
        assert got == 1 if want == 1 else got in (0, 2)
    assert not ring.verify_batch(alphas, ads, [bytes(b) for b in bad], cases.coeffs_for(n, 4, independent=False), aggregate=True)[1]
    # proofs produced by the unmodified reference at N = 2048
    g = load("ring1023_reference.json")
    es = g["proofs_test_vectors"] + g["proofs_blinded"]
    v, ok = ring.verify_batch([bytes.fromhex(e["alpha"]) for e in es], [bytes.fromhex(e["ad"]) for e in es], [bytes.fromhex(e["proof"]) for e in es], cases.coeffs_for(len(es)))
    assert v == [1] * len(es) and ok
    ring.close()


def test_python_api_end_to_end_on_gpu():
    """The drop-in API on the CUDA engine: README flow of the reference (ring 8) incl. Tiny / Pedersen."""
    from dot_ring_b200 import Bandersnatch, PedersenVRF, Ring, RingProofParams, RingRoot, RingVRF, TinyVRF
    from dot_ring_b200 import engine as engine_mod
    from tests.helpers import hx, ring_proof_bytes, split_keys

    eng = engine_mod.Engine(0, window_bits=8, srs_points=1537)
    engine_mod.set_default_engine(eng, 0)
    try:
        v = load("bandersnatch_sha-512_ell2_ring.json")[2]
        keys = split_keys(hx(v, "ring_pks"))
        params = RingProofParams(test_vectors=True)
        ring = Ring(keys, params)
        root = RingRoot.from_ring(ring, params)
        assert root.encode().hex() == v["ring_pks_com"]
        R = RingVRF[Bandersnatch]
        proof = R.prove(hx(v, "alpha"), hx(v, "ad"), hx(v, "sk"), hx(v, "pk"), ring, root)
        assert proof.encode() == ring_proof_bytes(v)
        assert proof.verify(hx(v, "alpha"), hx(v, "ad"), ring, root)
        assert not proof.verify(hx(v, "alpha"), b"no", ring, root)
        assert R.batch_verify([proof, R.decode(proof.encode())], [hx(v, "alpha")] * 2, [hx(v, "ad")] * 2, ring, root)
        assert R.proof_to_hash(proof.pedersen_proof.output_point).hex() == v["beta"]
        vp = load("bandersnatch_sha-512_ell2_pedersen.json")[1]
        pp = PedersenVRF[Bandersnatch].prove(hx(vp, "alpha"), hx(vp, "sk"), hx(vp, "ad"))
        assert pp.encode() == hx(vp, "gamma", "proof_pk_com", "proof_r", "proof_ok", "proof_s", "proof_sb")
        assert pp.verify(hx(vp, "alpha"), hx(vp, "ad")) and not pp.verify(b"other", hx(vp, "ad"))
        assert PedersenVRF[Bandersnatch].batch_verify([pp, pp], [hx(vp, "alpha")] * 2, [hx(vp, "ad")] * 2)
        vt = load("bandersnatch_sha-512_ell2_tiny.json")[1]
        tp = TinyVRF[Bandersnatch].prove(hx(vt, "alpha"), hx(vt, "sk"), hx(vt, "ad"))
        assert tp.encode() == hx(vt, "gamma", "proof_c", "proof_s")
        assert tp.verify(hx(vt, "pk"), hx(vt, "alpha"), hx(vt, "ad"))
        assert TinyVRF[Bandersnatch].proof_to_hash(tp.output_point).hex() == vt["beta"]
    finally:
        engine_mod.set_default_engine(None, 0)
        eng.close()


def test_edge_cases_and_domain_1024(ctx, srs):
    """Empty / ragged inputs, capacity limits, malformed keys; byte parity with the oracle prover at N = 512 and N = 1024."""
    from tests import edge_cases

    edge_cases.empty_batches(ctx, srs)
    edge_cases.te_msm_matches_oracle(ctx)
    edge_cases.ring_capacity_and_bad_keys(srs)
    edge_cases.ragged_inputs_match_oracle(srs, ((512, 5), (1024, 300)), n_items=4)


def test_extended_domain_8192_matches_oracle(ctx):
    """N = 8192 (a domain the reference rejects): the large-domain route (two-pass NTT, element-wise coset twists, dense
    witness commitments) against the CPU oracle on a synthetic 24 577-point SRS, plus prove -> verify round trips."""
    from tests import extended_domain

    extended_domain.prove_verify_against_oracle(ctx, domain=8192, n_keys=40, n_proofs=3, oracle_proofs=1, window_bits=8)


def test_domain_4096_with_a_larger_srs_matches_oracle(ctx):
    """N = 4096 is the largest domain the reference's parameters allow (params.py:20), but its bundled SRS (6145 points) cannot
    serve it; with a 12 289-point SRS (DOT_RING_BLS12_381_SRS in the reference, synthetic here) the fused single-CTA kernels
    and the sparse witness commitments run at their maximum size."""
    from tests import extended_domain

    extended_domain.prove_verify_against_oracle(ctx, domain=4096, n_keys=40, n_proofs=3, oracle_proofs=1, window_bits=8)


def test_vrf_vectors_through_the_large_batch_kernels(ctx):
    """Tiny / Thin / Pedersen batches above the cross-over use the one-thread-per-item kernels: same vectors, same verdicts."""
    lib = ctx.library.lib
    lib.dr_vrf_verify_set_coop_threshold(0)
    try:
        cases.pedersen_vectors(ctx)
        cases.tiny_vectors(ctx)
        cases.thin_vectors(ctx)
        cases.vrf_batch_fixtures(ctx)
    finally:
        lib.dr_vrf_verify_set_coop_threshold(8192)
