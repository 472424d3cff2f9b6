"""Shared test helpers: the reference's own benchmark seeding (tests/benchmark/bench_ring_proof.py:47-77)."""

from __future__ import annotations

import hashlib
import json
from functools import lru_cache
from pathlib import Path

GOLDEN = Path(__file__).resolve().parent / "golden"
REFVEC = GOLDEN / "reference_vectors"


def seed(*parts) -> bytes:
    h = hashlib.sha256()
    for part in parts:
        if isinstance(part, bytes):
            h.update(part)
        elif isinstance(part, int):
            h.update(part.to_bytes(8, "little"))
        else:
            h.update(part.encode())
        h.update(b"\0")
    return h.digest()


def le64(i: int) -> bytes:
    return i.to_bytes(8, "little")


def load(name: str):
    path = GOLDEN / name
    if not path.exists():
        path = REFVEC / name
    return json.loads(path.read_text())


def hx(vec, *fields) -> bytes:
    return bytes.fromhex("".join(vec[f] for f in fields))


def ring_proof_bytes(vec) -> bytes:
    return hx(vec, "gamma", "proof_pk_com", "proof_r", "proof_ok", "proof_s", "proof_sb", "ring_proof")


def split_keys(blob: bytes) -> list[bytes]:
    return [blob[i : i + 32] for i in range(0, len(blob), 32)]


@lru_cache(maxsize=1)
def bench_ring_keys(ring_size: int = 1023):
    """(signer_pk, signer_sk, keys) exactly as bench_ring_proof.py:140-152 builds them (oracle key derivation)."""
    from oracle import bandersnatch as bs
    from oracle import transcript as tr

    pk, sk = tr.secret_from_seed(bs.SHA512, seed("batch-signer", 0, 0))
    idx = min(3, ring_size - 1)
    keys = [pk if i == idx else tr.secret_from_seed(bs.SHA512, seed("ring-member", 0, i))[0] for i in range(ring_size)]
    return pk, sk, keys
