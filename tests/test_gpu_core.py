"""GPU parity tests for the arithmetic layer, KZG commit, NTT and G1 codecs (through the C ABI).

Run on the B200 box:  python -m pytest tests -m gpu -x -q
Each case compares the CUDA path with the CPU oracle on the same seeded inputs."""

import random

import pytest

from oracle import bandersnatch as bs
from oracle import bls12_381 as bls
from oracle import fr
from oracle import ring_proof as rp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from dot_ring_b200 import _native

    lib = _native.default_library()
    assert lib.is_cuda
    c = _native.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def small_srs(ctx):
    from dot_ring_b200 import _native
    from dot_ring_b200.srs import read_srs_file

    raw = read_srs_file(None, 600)
    s = _native.NativeSrs(ctx, raw.g1_be96, raw.g2_be192, 8)
    yield s
    s.close()


@pytest.mark.parametrize("field,mod", [("fq", bls.P), ("fr", fr.R), ("fn", bs.N)])
def test_field_kernels_match_python_ints(ctx, field, mod):
    rng = random.Random(hash(field) & 0xFFFF)
    edge = [0, 1, 2, mod - 1, mod - 2, (1 << 32) - 1, 1 << 32, mod >> 1, (mod >> 1) + 1, 1 << (mod.bit_length() - 1)]
    a = [x for x in edge for _ in edge] + [rng.randrange(mod) for _ in range(4000)]
    b = [y for _ in edge for y in edge] + [rng.randrange(mod) for _ in range(4000)]
    assert ctx.field_op(field, "mul", a, b) == [x * y % mod for x, y in zip(a, b)]
    assert ctx.field_op(field, "add", a, b) == [(x + y) % mod for x, y in zip(a, b)]
    assert ctx.field_op(field, "sub", a, b) == [(x - y) % mod for x, y in zip(a, b)]
    assert ctx.field_op(field, "sqr", a) == [x * x % mod for x in a]
    assert ctx.field_op(field, "neg", a) == [(-x) % mod for x in a]
    assert ctx.field_op(field, "inv", a[:300]) == [pow(x, -1, mod) if x else 0 for x in a[:300]]


def test_kzg_commit_matches_oracle(ctx, small_srs):
    srs = rp.load_srs()
    rng = random.Random(11)
    for n in (1, 2, 33, 600):
        sub = rp.SRS(srs.g1[:n], srs.g2)
        vecs = [[rng.randrange(fr.R) for _ in range(n)] for _ in range(3)]
        vecs += [[0] * n, [1] + [0] * (n - 1), [fr.R - 1] * n, [rng.randrange(4) for _ in range(n)]]
        got = small_srs.commit(vecs)
        for v, g in zip(vecs, got):
            assert g == bls.g1_serialize(rp.kzg_commit(sub, v))


def test_mixed_window_geometries_commit_like_the_oracle(ctx):
    """dr_srs_load geometries with some windows one bit wider (the bench default is 14 bits with four 15-bit windows)."""
    from tests import window_cases

    window_cases.check_commit_geometries(ctx, [(8, 8), (10, 6), (13, 9), (14, 4), (8, 0, True), (13, 2, True), (16, 0, True)], n=40)
    ctx.set_commit_mode(1)  # batched-affine rounds read the same table
    try:
        window_cases.check_commit_geometries(ctx, [(8, 8)], n=700, seed=6)
    finally:
        ctx.set_commit_mode(0)


def test_kzg_commit_linearity_full_size(ctx, small_srs):
    """Size-independent property: commit(a) + commit(b) == commit(a + b) at the full table width."""
    rng = random.Random(12)
    n = 600
    a = [rng.randrange(fr.R) for _ in range(n)]
    b = [rng.randrange(fr.R) for _ in range(n)]
    ca, cb, cab = small_srs.commit([a, b, [(x + y) % fr.R for x, y in zip(a, b)]])
    s = bls.g1_add(bls.g1_decompress(ca), bls.g1_decompress(cb))
    assert bls.g1_serialize(s) == cab


def test_ntt_matches_oracle_and_round_trips(ctx):
    rng = random.Random(13)
    for n in (2, 8, 512, 2048, 4096):
        params_root, size = fr.extend_root_to_size(fr.ROOT_OF_UNITY_2048, 2048, max(n, 2048))
        omega = pow(params_root, size // n, fr.R)
        vals = [rng.randrange(fr.R) for _ in range(3 * n)]
        fwd = ctx.fr_ntt(vals, n, omega, False)
        if n <= 2048:
            assert fwd == sum((fr.ntt(vals[i * n : (i + 1) * n], omega) for i in range(3)), [])
        assert ctx.fr_ntt(fwd, n, omega, True) == vals
    g = __import__("tests.helpers", fromlist=["load"]).load("ntt_reference.json")
    for e in g:
        if e["n"] > 4096:
            continue
        r = random.Random(e["seed"])
        vals = [r.randrange(fr.R) for _ in range(e["n"])]
        inv = ctx.fr_ntt(vals, e["n"], int(e["omega"], 16), True)
        assert [hex(v) for v in inv[:4]] == e["inverse_head"]


def test_g1_codecs(ctx):
    rng = random.Random(14)
    g = (bls.G1_GEN[0], bls.G1_GEN[1], 1)
    pts = [bls.g1_mul(g, rng.randrange(fr.R)) for _ in range(16)] + [None]
    ser = b"".join(bls.g1_serialize(p) for p in pts)
    comp = ctx.g1_compress(ser)
    assert comp == b"".join(bls.g1_compress(p) for p in pts)
    back, ok = ctx.g1_decompress(comp)
    assert back == ser and ok == b"\x01" * len(pts)
    _, ok = ctx.g1_decompress(b"\xff" * 48 + bytes(48))
    assert ok == b"\x00\x00"


def test_g1_msm_vs_oracle_and_sweep_property(ctx):
    """Variable-base MSM: against the oracle on SRS points up to 6145, and at 2^11 .. 2^20 points over the synthetic SRS
    tau^i * G where the result must equal (sum k_i tau^i) * G (uniform, all-ones and random-bit scalars)."""
    from tests import msm_cases

    msm_cases.msm_vs_oracle(ctx, [1, 33, 2048, 6145])
    msm_cases.synthetic_property(ctx, [1 << 11, 1 << 14], (0, 1, 2))
    msm_cases.synthetic_property(ctx, [1 << 17, 1 << 20], (0, 2))


def test_large_ntt_two_pass(ctx):
    """n = 2^13, 2^16 against the oracle; 2^18 forward/inverse round trip (the 4x LDE domain of a 65k-key ring)."""
    from tests import msm_cases

    msm_cases.large_ntt(ctx, [1 << 13, 1 << 16, 1 << 18], {1 << 13, 1 << 16})
