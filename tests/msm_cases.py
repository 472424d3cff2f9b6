"""G1 MSM parity cases (shared by the CPU emulation suite and the GPU suite)."""

from __future__ import annotations

import random

from dot_ring_b200.srs import read_srs_file
from oracle import bls12_381 as bls
from oracle import fr

TAU = 0x5EED5EED5EED5EED5EED  # public test value for the synthetic SRS


def _splitmix(st: int):
    st = (st + 0x9E3779B97F4A7C15) & (2**64 - 1)
    z = st
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
    return st, z ^ (z >> 31)


def expected_synthetic(n: int, seed: int, distribution: int) -> bytes:
    """(sum_i k_i tau^i) * G for the scalar stream dr_g1_msm_bench generates."""
    st, s, tp = seed, 0, 1
    for _ in range(n):
        if distribution == 0:
            w = []
            for _ in range(4):
                st, z = _splitmix(st)
                w.append(z)
            k = w[0] | w[1] << 64 | w[2] << 128 | (w[3] >> 2) << 192
        elif distribution == 1:
            k = 1
        else:
            st, z = _splitmix(st)
            k = z & 1
        s = (s + k * tp) % fr.R
        tp = tp * TAU % fr.R
    return bls.g1_serialize(bls.g1_mul((bls.G1_GEN[0], bls.G1_GEN[1], 1), s))


def msm_vs_oracle(ctx, sizes):
    """dr_g1_msm over real SRS points against the oracle MSM: uniform, unit, bit, near-modulus, zero and unreduced scalars."""
    top = max(sizes)
    raw = read_srs_file(None, top)
    pts = [(int.from_bytes(raw.g1_be96[96 * i : 96 * i + 48], "big"), int.from_bytes(raw.g1_be96[96 * i + 48 : 96 * i + 96], "big")) for i in range(top)]
    rng = random.Random(3)
    for n in sizes:
        for mode in ("uniform", "ones", "bits", "near_r"):
            if mode == "uniform":
                ks = [rng.randrange(fr.R) for _ in range(n)]
            elif mode == "ones":
                ks = [1] * n
            elif mode == "bits":
                ks = [rng.randrange(2) for _ in range(n)]
            else:
                ks = [fr.R - 1 - rng.randrange(3) for _ in range(n)]
            assert ctx.g1_msm(raw.g1_be96[: 96 * n], ks) == bls.g1_serialize(bls.g1_msm(pts[:n], ks)), (n, mode)
    assert ctx.g1_msm(raw.g1_be96[: 96 * 4], [0, 0, 0, 0]) == bytes([0x40]) + bytes(95)
    assert ctx.g1_msm(b"", []) == bytes([0x40]) + bytes(95)
    # P - P, and a scalar >= r is reduced (kzg.py passes unreduced coefficients)
    assert ctx.g1_msm(raw.g1_be96[:96] * 2, [5, fr.R - 5]) == bytes([0x40]) + bytes(95)
    assert ctx.g1_msm(raw.g1_be96[:96], [fr.R + 7]) == bls.g1_serialize(bls.g1_mul((pts[0][0], pts[0][1], 1), 7))
    bad = bytearray(raw.g1_be96[:96])
    bad[95] ^= 1
    try:
        ctx.g1_msm(bytes(bad), [1])
    except ValueError:
        pass
    else:
        raise AssertionError("malformed point accepted")


def synthetic_property(ctx, sizes, distributions=(0, 1, 2), seed: int = 7):
    for n in sizes:
        for dist in distributions:
            _, _, out = ctx.g1_msm_bench(n, 1, seed, dist, TAU)
            assert out == expected_synthetic(n, seed, dist), (n, dist)


def large_ntt(ctx, sizes, oracle_sizes):
    """dr_fr_ntt beyond one CTA's shared memory (two-pass transform): against the oracle NTT with the root the reference's
    `_extend_root_to_size` walk gives (params.py:63-115), plus forward/inverse round trips at the larger sizes."""
    import random as _random

    from dot_ring_b200.params import ROOT_OF_UNITY_2048, _extend_root_to_size, _omega_for_domain

    rng = _random.Random(4)
    for n in sizes:
        root, size = _extend_root_to_size(ROOT_OF_UNITY_2048, 2048, n, fr.R)
        om = _omega_for_domain(n, fr.R, root, size)
        vals = [rng.randrange(fr.R) for _ in range(2 * n)]  # batch of two
        got = ctx.fr_ntt(vals, n, om)
        if n in oracle_sizes:
            assert got == fr.ntt(vals[:n], om) + fr.ntt(vals[n:], om), n
        assert ctx.fr_ntt(got, n, om, inverse=True) == vals, n
    pts = ctx.g1_synthetic_srs(TAU, 5, 3)
    gen = (bls.G1_GEN[0], bls.G1_GEN[1], 1)
    assert pts == b"".join(bls.g1_serialize(bls.g1_mul(gen, pow(TAU, 5 + i, fr.R))) for i in range(3))
