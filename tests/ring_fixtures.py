"""Helpers shared by the CPU-emulation and GPU ring-proof tests: build native rings from suite constants."""

from __future__ import annotations

from dot_ring_b200 import _native
from dot_ring_b200.srs import read_srs_file
from oracle import bandersnatch as bs
from oracle import ring_proof as rp


def native_srs(ctx, n_points: int | None, window_bits: int):
    raw = read_srs_file(None, n_points)
    return _native.NativeSrs(ctx, raw.g1_be96, raw.g2_be192, window_bits)


def native_ring(srs, keys, params: rp.Params):
    s = params.suite
    return _native.NativeRing(
        srs, list(keys), domain_size=params.domain_size, max_ring_size=params.max_ring_size, padding_rows=params.padding_rows,
        omega=params.omega, radix_omega=params.radix_omega, seed=s.accumulator_base, blinding_base=s.blinding_base,
        padding_point=s.padding_point, generator=bs.GENERATOR, suite_id=s.suite_id, h2c_dst=s.dst, hash_name=s.hash_name,
    )  # fmt: skip
