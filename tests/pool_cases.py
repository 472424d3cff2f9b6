"""EnginePool: a caller's batch sharded over several devices gives exactly what one device gives (engine.py)."""

from __future__ import annotations

import random

from tests.helpers import hx, load, split_keys


def check_pool_matches_single(api, pool, single_engine, n: int = 5) -> None:
    v = load("bandersnatch_sha-512_ell2_ring.json")[2]
    keys = split_keys(hx(v, "ring_pks"))
    params = api.RingProofParams()
    cls = api.RingVRF[api.Bandersnatch]
    ring_one = api.Ring(keys, params, single_engine)
    ring_pool = api.Ring(keys, params, pool)
    assert api.RingRoot.from_ring(ring_pool, params).encode() == api.RingRoot.from_ring(ring_one, params).encode()
    assert ring_pool.nm_points == ring_one.nm_points
    rng = random.Random(5)
    alphas = [b"pool-%d" % i for i in range(n)]
    ads = [b"ad" * (i % 3) for i in range(n)]
    zk = [rng.randrange(params.prime) for _ in range(12 * n)]
    one = cls.prove_batch(alphas, ads, hx(v, "sk"), hx(v, "pk"), ring_one, zk_rows=zk, as_bytes=True)
    many = cls.prove_batch(alphas, ads, hx(v, "sk"), hx(v, "pk"), ring_pool, zk_rows=zk, as_bytes=True)
    assert one == many and len(set(many)) == n
    zk_bytes = b"".join(z.to_bytes(32, "little") for z in zk)
    assert cls.prove_batch(alphas, ads, hx(v, "sk"), hx(v, "pk"), ring_pool, zk_rows=zk_bytes, as_bytes=True) == one
    root = api.RingRoot.from_ring(ring_pool, params)
    assert cls.verify_batch(many, alphas, ads, ring_pool, root) == [1] * n
    assert cls.batch_verify(many, alphas, ads, ring_pool, root) is True
    broken = list(many)
    broken[n - 1] = broken[n - 1][:200] + bytes([broken[n - 1][200] ^ 1]) + broken[n - 1][201:]
    verdicts = cls.verify_batch(broken, alphas, ads, ring_pool, root)
    assert verdicts[: n - 1] == [1] * (n - 1) and verdicts[n - 1] != 1
    assert cls.batch_verify(broken, alphas, ads, ring_pool, root) is False
    wrong_ad = list(ads)
    wrong_ad[0] = b"not the ad"
    assert cls.verify_batch(many, alphas, wrong_ad, ring_pool, root) == [0] + [1] * (n - 1)
    # shard arithmetic: contiguous, balanced, order-preserving
    assert pool.shard_bounds(5, 2) == [(0, 3), (3, 5)] and pool.shard_bounds(1, 2) == [(0, 1)] and pool.shard_bounds(0, 2) == []
    assert pool.shard_bounds(4096, 8) == [(512 * i, 512 * (i + 1)) for i in range(8)]


def check_range_split_msm(pool, ctx, n: int = 40) -> None:
    """EnginePool.g1_msm: the point-range split gives the same 96 bytes as one device."""
    from dot_ring_b200.srs import read_srs_file

    raw = read_srs_file(None, n)
    rng = random.Random(8)
    ks = [rng.randrange(1 << 255) for _ in range(n)]
    ks[0], ks[1] = 0, 1
    whole = ctx.g1_msm(raw.g1_be96, ks)
    assert pool.g1_msm(raw.g1_be96, ks, min_points_per_device=7) == whole  # forces a split into len(pool) ranges
    assert pool.g1_msm(raw.g1_be96, ks) == whole  # default threshold: not split
