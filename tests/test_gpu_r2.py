"""GPU parity, round 2 (through the C ABI / the Python mirror on the CUDA build):

* the HEADLINE configuration itself -- GLV table with 16-bit windows over the full 6145-point SRS (161 GB), 8192 proofs in one
  device pass, blinded rows -- against the unmodified reference's ring-1023 proofs, the oracle verifier and the device verifier,
  plus full-length (6145 / 6144 / 2047-term) KZG commitments against the oracle at that geometry;
* the multi-device dispatcher (EnginePool) against a single device;
* the PCS seam, the verifier transcript prefix, max_ring_size below capacity -- reference-generated vectors;
* dr_te_decode_batch / dr_te_mul_batch against reference-generated Bandersnatch vectors; the SHAKE128 ring suite's ark-vrf vectors.
"""

import random

import pytest

from oracle import bandersnatch as bs
from oracle import bls12_381 as bls
from oracle import fr
from oracle import ring_proof as rp
from oracle import vrf as ovrf
from tests.helpers import bench_ring_keys, hx, le64, load, ring_proof_bytes, split_keys
from tests.ring_fixtures import native_ring

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    """The Python mirror on a CUDA engine with the full SRS and a small (2.4 GB) window table."""
    from dot_ring_b200 import engine as engine_mod

    eng = engine_mod.Engine(0, window_bits=8)
    assert eng.ctx.library.is_cuda
    engine_mod.set_default_engine(eng, 0)
    import dot_ring_b200 as pkg

    yield pkg
    engine_mod.set_default_engine(None, 0)
    eng.close()


def test_headline_configuration_full_size():
    """bench.py's exact configuration: (16-bit windows, GLV) table on the full SRS, one 8192-proof pass with blinded rows."""
    from dot_ring_b200 import _native
    from dot_ring_b200.srs import read_srs_file

    ctx = _native.Context(0)
    raw = read_srs_file(None, None)
    srs = _native.NativeSrs(ctx, raw.g1_be96, raw.g2_be192, 16, 0, True)
    try:
        assert srs.geometry == (16, 0, 1, 16) and srs.table_bytes > 160e9
        # full-length commitments at this geometry: the top window (msm.cuh TableGeom::top) sees real 255-bit scalars
        osrs = rp.load_srs()
        rng = random.Random(2)
        for n in (6145, 6144, 2047):
            vecs = [[rng.randrange(fr.R) for _ in range(n)], [fr.R - 1] * n, [rng.randrange(fr.R) if i % 97 == 0 else 0 for i in range(n)]]
            got = srs.commit(vecs)
            for vec, g in zip(vecs, got):
                assert g == bls.g1_serialize(rp.kzg_commit(rp.SRS(osrs.g1[:n], osrs.g2), vec)), n
        g = load("ring1023_reference.json")
        pk, sk, keys = bench_ring_keys(1023)
        params = rp.Params.from_ring_size(1023)
        ring = native_ring(srs, keys, params)
        assert ring.root().hex() == g["ring_root"]
        golden = [(bytes.fromhex(e["alpha"]), bytes.fromhex(e["ad"]), [0] * 12, e["proof"]) for e in g["proofs_test_vectors"]]
        golden += [(bytes.fromhex(e["alpha"]), bytes.fromhex(e["ad"]), [int(z, 16) for z in e["zk_rows"]], e["proof"]) for e in g["proofs_blinded"]]
        n = 8192
        rng = random.Random(0)
        alphas = [it[0] for it in golden] + [b"bench-batch-input" + le64(j) for j in range(len(golden), n)]
        ads = [it[1] for it in golden] + [b"bench-batch-ad" + le64(j) for j in range(len(golden), n)]
        zk = b"".join(z.to_bytes(32, "little") for it in golden for z in it[2]) + b"".join(rng.randrange(fr.R).to_bytes(32, "little") for _ in range(12 * (n - len(golden))))
        proofs, status = ring.prove_batch(alphas, ads, [sk] * n, [3] * n, zk_rows=zk)
        assert status == [0] * n and len(set(proofs)) == n
        for i, it in enumerate(golden):
            assert proofs[i].hex() == it[3], f"golden proof {i} differs in the 8192-wide pass"
        # device verifier over a sample spread across the pass, both modes
        from tests import verify_cases as cases

        picks = sorted(random.Random(1).sample(range(n), 253) + [0, n // 2, n - 1])
        v, ok = ring.verify_batch([alphas[i] for i in picks], [ads[i] for i in picks], [proofs[i] for i in picks], cases.coeffs_for(len(picks)))
        assert ok and v == [1] * len(picks)
        v, ok = ring.verify_batch([alphas[i] for i in picks], [ads[i] for i in picks], [proofs[i] for i in picks], cases.coeffs_for(len(picks), independent=False), aggregate=True)
        assert ok
        # the CPU oracle's verifier accepts proofs from the start, the middle and the end of the pass
        oring = rp.Ring(keys, params)
        oroot = rp.RingRoot.from_ring(oring, params)
        assert oroot.encode() == ring.root()
        for i in (7, n // 2 + 1, n - 1):
            assert ovrf.ring_verify(ovrf.RingVrfProof.decode(proofs[i]), alphas[i], ads[i], oring, oroot, ring_matches=True), i
        ring.close()
    finally:
        srs.close()
        ctx.trim()
        ctx.close()


def test_engine_pool_matches_single_device(api):
    from dot_ring_b200 import _native
    from dot_ring_b200 import engine as engine_mod
    from tests import pool_cases

    devices = [0, 1] if _native.default_library().device_count() >= 2 else [0, 0]
    pool = engine_mod.EnginePool(devices=devices, window_bits=8, srs_points=1537)
    try:
        pool_cases.check_pool_matches_single(api, pool, engine_mod.default_engine(), n=37)
        pool_cases.check_range_split_msm(pool, engine_mod.default_engine().ctx, n=3000)
    finally:
        pool.close()


def test_pcs_seam_transcript_prefix_and_small_rings(api):
    from tests import pcs_cases, small_ring_cases

    pcs_cases.check_pcs(api)
    pcs_cases.check_transcript_prefix(api)
    pcs_cases.check_small_ring_golden(api, domains=(512, 2048))
    small_ring_cases.check_small_max_ring(api, cases=((512, 100), (512, 1), (2048, 6), (2048, 1000)))


def test_te_batch_ops_match_reference_goldens(api):
    """dr_te_decode_batch / dr_te_mul_batch (and the Elligator 2 map behind the provers) against vectors written by the
    unmodified reference (tests/golden/bandersnatch_reference.json)."""
    from dot_ring_b200 import engine as engine_mod

    ctx = engine_mod.default_engine().ctx
    g = load("bandersnatch_reference.json")
    dec = ctx.te_decode([bytes.fromhex(e["raw"]) for e in g["dec_point"]], checked=True)
    for e, got in zip(g["dec_point"], dec):
        assert (got is not None) == e["ok"]
        if e["ok"]:
            assert got == (int(e["x"], 16), int(e["y"], 16))
    assert ctx.te_decode([bs.point_to_string((0, bs.P - 1)), bs.point_to_string(bs.IDENTITY)], checked=True) == [None, None]
    outs = ctx.te_mul([bytes.fromhex(e["base"]) for e in g["scalar_mul"]], [int(e["k"], 16) for e in g["scalar_mul"]])
    assert [o.hex() for o in outs] == [e["out"] for e in g["scalar_mul"]]
    # one base, many scalars (n_points == 1 form), incl. 0, order - 1, order + 5
    base = bytes.fromhex(g["scalar_mul"][0]["base"])
    ks = [0, 1, bs.N - 1, bs.N + 5, 1 << 252]
    bp = bs.string_to_point(base) if hasattr(bs, "string_to_point") else None
    got = ctx.te_mul([base], ks)
    if bp is not None:
        assert got == [bs.point_to_string(bs.mul(bp, k % bs.N)) for k in ks]
    # the hash-to-curve points of the golden file, through the Tiny prover's output: gamma = sk * H(alpha) with sk = 1
    su = __import__("tests.verify_cases", fromlist=["suite_struct"]).suite_struct()
    proofs = ctx.vrf_prove("tiny", su, [bytes.fromhex(e["alpha"]) for e in g["encode_to_curve"]], [b""] * len(g["encode_to_curve"]), [(1).to_bytes(32, "little")] * len(g["encode_to_curve"]))
    assert [p[:32].hex() for p in proofs] == [e["point"] for e in g["encode_to_curve"]]


def test_shake128_ring_vectors_on_gpu(api):
    """All ark-vrf vectors of the SHAKE128 ring suite (specs/bandersnatch.py:108-144) on the CUDA path: root, proof bytes, verify."""
    cv = api.Bandersnatch_SHAKE128
    params = api.RingProofParams(test_vectors=True, cv=cv)
    for v in load("bandersnatch_shake128_ell2_ring.json"):
        keys = split_keys(hx(v, "ring_pks"))
        ring = api.Ring(keys, params)
        root = api.RingRoot.from_ring(ring, params)
        assert root.encode().hex() == v["ring_pks_com"]
        proof = api.RingVRF[cv].prove(hx(v, "alpha"), hx(v, "ad"), hx(v, "sk"), hx(v, "pk"), ring, root)
        assert proof.encode() == ring_proof_bytes(v)
        assert proof.verify(hx(v, "alpha"), hx(v, "ad"), ring, root)
        assert not proof.verify(hx(v, "alpha") + b"x", hx(v, "ad"), ring, root)
        assert api.RingVRF[cv].proof_to_hash(proof.pedersen_proof.output_point).hex() == v["beta"]
        # the SHA-512 suite must reject the same bytes (different transcript hash)
        sha_ring = api.Ring(keys, api.RingProofParams(test_vectors=True))
        assert not api.RingVRF[api.Bandersnatch].decode(proof.encode()).verify(hx(v, "alpha"), hx(v, "ad"), sha_ring, api.RingRoot.from_ring(sha_ring))
