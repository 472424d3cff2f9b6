"""Fixed-base commit under mixed window geometries (the `wide` lowest windows one bit wider, include/dot_ring_b200.h dr_srs_load),
shared by the CPU-emulation and GPU suites: every geometry must give the oracle's commitment for scalars that exercise the
carries between windows and the top window."""

from __future__ import annotations

import random

from dot_ring_b200 import _native
from dot_ring_b200.srs import read_srs_file
from oracle import bls12_381 as bls
from oracle import fr
from oracle import ring_proof as rp


def expected_windows(c: int, wide: int, glv: bool = False) -> int:
    return -(-((128 if glv else 256) - wide) // c)


def check_commit_geometries(ctx, geometries, n: int, seed: int = 5) -> None:
    raw = read_srs_file(None, n)
    srs = rp.load_srs()
    sub = rp.SRS(srs.g1[:n], srs.g2)
    rng = random.Random(seed)
    vecs = [[rng.randrange(fr.R) for _ in range(n)] for _ in range(2)]
    # all-ones digits (carry chains through every window), the largest scalar, single high bits, small values
    lam = 0xAC45A4010001A40200000000FFFFFFFF  # GLV eigenvalue: halves at the edges of their range
    vecs += [[(lam * lam + lam - i) % fr.R for i in range(n)], [lam - 1 + i for i in range(n)], [(lam + 1) * lam - 1 + i for i in range(n)]]
    vecs += [[fr.R - 1] * n, [(1 << 254) + i for i in range(n)], [(1 << 255) - 1 - i for i in range(n)], [1 << (7 * i % 255) for i in range(n)],
             [rng.randrange(3) for _ in range(n)], [0] * n]  # fmt: skip
    want = [bls.g1_serialize(rp.kzg_commit(sub, [x % fr.R for x in v])) for v in vecs]
    for geom in geometries:
        c, wide, glv = geom if len(geom) == 3 else (*geom, False)
        native = _native.NativeSrs(ctx, raw.g1_be96, raw.g2_be192, c, wide, glv)
        try:
            windows = expected_windows(c, wide, glv)
            assert native.geometry == (c, wide, int(glv), windows * (2 if glv else 1))
            if not glv:
                assert native.table_bytes == n * (windows + wide) * (1 << (c - 1)) * 96
            assert native.commit(vecs) == want, geom
        finally:
            native.close()
