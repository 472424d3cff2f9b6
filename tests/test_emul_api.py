"""CPU checks of the host-side mirror of the reference API (dot_ring_b200.Ring / RingRoot / RingVRF) and of the
C-ABI surface.  The kernels run here through the *test-only* emulation build (tests/host/emul.py: same .cu
sources compiled by g++, block after block on the CPU); the product package never selects it by itself.

Mirrors the reference's tests/test_ark_vrf.py:118-132 (byte-exact ring root and proof) and the malformed-input
cases of tests/test_ring_vrf/test_audit_regressions.py:30-114."""

import ctypes

import pytest

import __graft_entry__ as entry
from dot_ring_b200 import _native
from dot_ring_b200 import engine as engine_mod
from tests.helpers import hx, load, ring_proof_bytes, split_keys
from tests.host.emul import emulation_library


@pytest.fixture(scope="module")
def api():
    lib = emulation_library()
    old = _native._default
    _native.set_default_library(lib)
    eng = engine_mod.Engine(0, window_bits=4, library=lib, srs_points=1537)
    engine_mod.set_default_engine(eng, 0)
    import dot_ring_b200 as pkg

    yield pkg
    engine_mod.set_default_engine(None, 0)
    eng.close()
    _native.set_default_library(old)


def test_cuda_library_exports_every_declared_symbol():
    """No compute call: the CUDA build loads (libcudart only) and exports what include/dot_ring_b200.h declares."""
    path = _native.DEFAULT_LIBRARY
    if not path.exists():
        pytest.skip("CUDA library not built (run ./build.sh)")
    lib = ctypes.CDLL(str(path))
    names = entry.declared_symbols()
    assert len(names) >= 30
    assert [n for n in names if not hasattr(lib, n)] == []
    assert lib.dr_is_cuda_build() == 1


def test_product_refuses_non_cuda_library():
    lib = emulation_library()
    with pytest.raises(ImportError):
        _native.Library(lib.path, require_cuda=True)
    with pytest.raises(ImportError):
        _native.Library("/nonexistent/libdotring_b200.so")


def test_ring_api_matches_reference_vectors(api):
    vectors = load("bandersnatch_sha-512_ell2_ring.json")
    params = api.RingProofParams(test_vectors=True)
    for v in vectors[:2]:
        keys = split_keys(hx(v, "ring_pks"))
        ring = api.Ring(keys, params)
        root = api.RingRoot.from_ring(ring, params)
        assert root.encode().hex() == v["ring_pks_com"]
        assert api.RingRoot.decode(root.encode(), params).encode() == root.encode()
        proof = api.RingVRF[api.Bandersnatch].prove(hx(v, "alpha"), hx(v, "ad"), hx(v, "sk"), hx(v, "pk"), ring, root)
        assert proof.encode() == ring_proof_bytes(v)
        again = api.RingVRF[api.Bandersnatch].decode(proof.encode())
        assert again.encode() == proof.encode() and again.l_zeta_omega == proof.l_zeta_omega


def test_ring_api_error_behaviour(api):
    v = load("bandersnatch_sha-512_ell2_ring.json")[0]
    keys = split_keys(hx(v, "ring_pks"))
    params = api.RingProofParams(test_vectors=True)
    ring = api.Ring(keys, params)
    cls = api.RingVRF[api.Bandersnatch]
    other_pk = keys[(keys.index(hx(v, "pk")) + 1) % len(keys)]
    with pytest.raises(ValueError):  # vrf/ring/vrf.py:196-197
        cls.prove(hx(v, "alpha"), hx(v, "ad"), hx(v, "sk"), other_pk, ring)
    outsider_pk, outsider_sk = api.Bandersnatch.secret_from_seed(b"\x07" * 32)
    with pytest.raises(ValueError, match="not in ring"):  # members.py:71-81
        cls.prove(b"a", b"b", outsider_sk, outsider_pk, ring)
    with pytest.raises(ValueError):
        cls.decode(b"\x00" * 783)
    good = ring_proof_bytes(v)
    bad_scalar = bytearray(good)
    bad_scalar[192 + 192 : 192 + 224] = b"\xff" * 32  # px_zeta >= r
    with pytest.raises(ValueError):
        cls.decode(bytes(bad_scalar))
    bad_g1 = bytearray(good)
    bad_g1[192] ^= 0x80  # clears the compression flag of C_b
    with pytest.raises(ValueError):
        cls.decode(bytes(bad_g1))
    with pytest.raises(ValueError):
        api.RingRoot.decode(b"\x00" * 143, params)
    with pytest.raises(ValueError):
        api.Ring(keys * 40, params)  # 320 keys > max_ring_size 255
    # undecodable keys become the padding point (members.py:36-41)
    padded = api.Ring([b"\xff" * 32, keys[0]], params)
    assert padded.nm_points[0] == params.cv.curve.params.auxiliary_points.padding_point


def test_params_table(api):
    """tests/test_coverage/test_params.py:66-88 size table."""
    P = api.RingProofParams
    for ring_size, domain, max_ring in [(8, 512, 255), (255, 512, 255), (256, 1024, 767), (1023, 2048, 1791), (1791, 2048, 1791), (1792, 4096, 3839)]:
        p = P.from_ring_size(ring_size)
        assert (p.domain_size, p.max_ring_size, p.radix_domain_size) == (domain, max_ring, 4 * domain)
    with pytest.raises(ValueError):
        P.from_ring_size(3840)  # domain 8192 > 4096
    with pytest.raises(ValueError):
        P.from_ring_size(0)
    p = P.from_ring_size(1023)
    assert pow(p.omega, 2048, p.prime) == 1 and pow(p.radix_omega, 4, p.prime) == p.omega


def test_sparse_and_dense_witness_commitments_give_identical_proofs(api):
    """The witness columns are committed from their evaluation form over Lagrange prefix-sum bases; the reference's route
    (interpolate, then KZG.commit of the coefficients; columns.py:29-60) must give byte-identical proofs, blinded rows included."""
    import random

    from dot_ring_b200 import engine as engine_mod

    v = load("bandersnatch_sha-512_ell2_ring.json")[3]
    keys = split_keys(hx(v, "ring_pks"))
    params = api.RingProofParams()  # blinded rows
    ring = api.Ring(keys, params)
    rng = random.Random(99)
    zk = [rng.randrange(params.prime) for _ in range(24)]
    cls = api.RingVRF[api.Bandersnatch]
    args = ([hx(v, "alpha"), b"second"], [hx(v, "ad"), b""], hx(v, "sk"), hx(v, "pk"), ring)
    sparse = cls.prove_batch(*args, zk_rows=zk, as_bytes=True)
    ctx = engine_mod.default_engine().ctx
    ctx.set_dense_witness_commit(True)
    try:
        dense = cls.prove_batch(*args, zk_rows=zk, as_bytes=True)
    finally:
        ctx.set_dense_witness_commit(False)
    assert sparse == dense and sparse[0] != sparse[1]
    ctx.set_commit_mode(1)  # batched-affine summation of the table entries (msm.cuh CommitAffineBody)
    try:
        affine = cls.prove_batch(*args, zk_rows=zk, as_bytes=True)
    finally:
        ctx.set_commit_mode(0)
    assert affine == sparse
    ctx.set_generic_ntt_path(True)  # the route domains above 4096 take (element-wise twists around the batched NTT)
    try:
        ring2 = api.Ring(keys, params)
        assert api.RingRoot.from_ring(ring2, params).encode() == api.RingRoot.from_ring(ring, params).encode()
        args2 = ([hx(v, "alpha"), b"second"], [hx(v, "ad"), b""], hx(v, "sk"), hx(v, "pk"), ring2)
        generic = cls.prove_batch(*args2, zk_rows=zk, as_bytes=True)
    finally:
        ctx.set_generic_ntt_path(False)
    assert generic == sparse
    # batches larger than one internal pass: same proofs whatever the pass size
    five = ([hx(v, "alpha"), b"b", b"c", b"", b"e" * 70], [hx(v, "ad"), b"", b"x", b"y" * 200, b""], hx(v, "sk"), hx(v, "pk"), ring)
    zk5 = [rng.randrange(params.prime) for _ in range(60)]
    whole = cls.prove_batch(*five, zk_rows=zk5, as_bytes=True)
    ctx.set_prove_chunk(2)
    try:
        chunked = cls.prove_batch(*five, zk_rows=zk5, as_bytes=True)
    finally:
        ctx.set_prove_chunk(0)
    assert chunked == whole and len(set(whole)) == 5
    root = api.RingRoot.from_ring(ring, params)
    assert cls.verify_batch(sparse, [hx(v, "alpha"), b"second"], [hx(v, "ad"), b""], ring, root) == [1, 1]


def test_shake128_suite_through_the_api(api):
    """Bandersnatch_SHAKE128 (specs/bandersnatch.py:108-144): SHAKE128 transcripts + XOF hash-to-curve, ark-vrf vectors."""
    cv = api.Bandersnatch_SHAKE128
    v = load("bandersnatch_shake128_ell2_ring.json")[1]
    keys = split_keys(hx(v, "ring_pks"))
    params = api.RingProofParams(test_vectors=True, cv=cv)
    ring = api.Ring(keys, params)
    root = api.RingRoot.from_ring(ring, params)
    assert root.encode().hex() == v["ring_pks_com"]
    assert cv.public_key_from_secret(hx(v, "sk")) == hx(v, "pk")
    proof = api.RingVRF[cv].prove(hx(v, "alpha"), hx(v, "ad"), hx(v, "sk"), hx(v, "pk"), ring, root)
    assert proof.encode() == ring_proof_bytes(v)
    assert proof.verify(hx(v, "alpha"), hx(v, "ad"), ring, root)
    assert api.RingVRF[cv].proof_to_hash(proof.pedersen_proof.output_point).hex() == v["beta"]
    vt = load("bandersnatch_shake128_ell2_tiny.json")[0]
    tp = api.TinyVRF[cv].prove(hx(vt, "alpha"), hx(vt, "sk"), hx(vt, "ad"))
    assert tp.encode() == hx(vt, "gamma", "proof_c", "proof_s") and tp.verify(hx(vt, "pk"), hx(vt, "alpha"), hx(vt, "ad"))
    assert not api.TinyVRF[api.Bandersnatch].decode(tp.encode()).verify(hx(vt, "pk"), hx(vt, "alpha"), hx(vt, "ad"))


def test_thin_vrf_api(api):
    cls = api.ThinVRF[api.Bandersnatch]
    vs = load("bandersnatch_sha-512_ell2_thin.json")
    v = vs[2]
    proof = cls.prove(hx(v, "alpha"), hx(v, "sk"), hx(v, "ad"))
    assert proof.encode() == hx(v, "gamma", "proof_r", "proof_s")
    assert cls.decode(proof.encode()).verify(hx(v, "pk"), hx(v, "alpha"), hx(v, "ad"))
    assert not proof.verify(hx(v, "pk"), b"other", hx(v, "ad"))
    assert cls.proof_to_hash(proof.output_point).hex() == v["beta"]
    two = [cls.decode(hx(x, "gamma", "proof_r", "proof_s")) for x in vs[:2]]
    args = ([hx(x, "pk") for x in vs[:2]], [hx(x, "alpha") for x in vs[:2]], [hx(x, "ad") for x in vs[:2]])
    assert cls.batch_verify(two, *args)
    flipped = cls(two[1].output_point, two[1].r, (two[1].s + 1) % api.Bandersnatch.curve.params.subgroup_order)
    assert not cls.batch_verify([two[0], flipped], *args)  # tests/test_ark_vrf.py:135-165
    with pytest.raises(ValueError):
        cls.decode(b"\x00" * 95)
    with pytest.raises(ValueError):
        proof.verify(b"\xff" * 32, hx(v, "alpha"), hx(v, "ad"))


def test_pedersen_blinding_factor_and_unblinding(api):
    """pedersen/vrf.py:111-126,144-162: the prover keeps its blinding factor; `verify_unblinding` checks Y_bar = Y + b*B."""
    cls = api.PedersenVRF[api.Bandersnatch]
    v = load("bandersnatch_sha-512_ell2_pedersen.json")[0]
    proof = cls.prove(hx(v, "alpha"), hx(v, "sk"), hx(v, "ad"))
    assert proof._blinding_factor == int.from_bytes(hx(v, "blinding"), "little")
    assert proof.verify_unblinding(hx(v, "pk"), proof._blinding_factor)
    assert not proof.verify_unblinding(hx(v, "pk"), proof._blinding_factor + 1)
    assert cls.decode(proof.encode())._blinding_factor is None


def test_mixed_window_geometries_commit_like_the_oracle(api):
    from tests import window_cases

    window_cases.check_commit_geometries(engine_mod.default_engine().ctx, [(4, 4), (5, 1), (7, 4), (4, 0, True), (6, 2, True), (7, 0, True), (11, 0, True)], n=5)


def test_max_ring_size_below_domain_capacity(api):
    """max_ring_size=100 / 1 at domain 512 must prove like the oracle (and the reference, whose tests use such sizes)."""
    from tests import small_ring_cases

    small_ring_cases.check_small_max_ring(api, cases=((512, 100), (512, 1)))
