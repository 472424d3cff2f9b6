"""SRS file loader (host side mirror of dot_ring/ring_proof/pcs/srs.py:20-90).

File layout (little-endian u64 counts, zcash uncompressed big-endian points):
  u64 n_g1 | n_g1 x 96-byte G1 | u64 n_g2 | n_g2 x 192-byte G2 (x.c1 | x.c0 | y.c1 | y.c0).
The bundled file is the 2^11 Zcash powers-of-tau SRS the reference ships
(dot_ring/vrf/data/bls12-381-srs-2-11-uncompressed-zcash.bin: 6145 G1 + 2 G2 points); the
environment variable DOT_RING_BLS12_381_SRS overrides it exactly as in the reference (srs.py:20-22).
"""

from __future__ import annotations

import os
from dataclasses import dataclass
from functools import lru_cache
from pathlib import Path

BUNDLED_SRS = Path(__file__).resolve().parent / "data" / "bls12-381-srs-2-11-uncompressed-zcash.bin"


@dataclass(frozen=True)
class SrsBytes:
    g1_be96: bytes  # n x 96
    g2_be192: bytes  # 2 x 192

    @property
    def n_g1(self) -> int:
        return len(self.g1_be96) // 96


@lru_cache(maxsize=4)
def read_srs_file(path: str | None = None, g1_limit: int | None = None) -> SrsBytes:
    p = Path(path or os.environ.get("DOT_RING_BLS12_381_SRS") or BUNDLED_SRS)
    data = p.read_bytes()
    if len(data) < 8:
        raise ValueError("File too short to contain header.")
    n1 = int.from_bytes(data[:8], "little")
    take = n1 if g1_limit is None else min(n1, g1_limit)
    if len(data) < 8 + 96 * n1 + 8:
        raise ValueError("File too short to contain G2 vector length header.")
    off = 8 + 96 * n1
    n2 = int.from_bytes(data[off : off + 8], "little")
    if n2 < 2:
        raise ValueError("SRS file must contain at least two G2 points")
    g2 = data[off + 8 : off + 8 + 384]
    if len(g2) != 384:
        raise ValueError("Unexpected end-of-file when reading G2 points.")
    return SrsBytes(data[8 : 8 + 96 * take], g2)
