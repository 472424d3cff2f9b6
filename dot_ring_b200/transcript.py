"""Host-side SHA-512 VRF transcript pieces needed outside the batched kernels.

Only key derivation (``secret_from_seed``) runs here: dot_ring/vrf/primitives.py:26-82,147-174.
Everything on the proving / verifying hot path hashes on the device (csrc/hash.cuh).
"""

from __future__ import annotations

import hashlib

NONCE_EXPAND, NONCE = 0x10, 0x11


def _squeeze(absorbed: bytes, size: int, hash_name: str = "sha512") -> bytes:
    """primitives.py:165-174: SHA-512 suites squeeze in counter mode, the SHAKE128 suite squeezes the XOF."""
    if hash_name == "shake128":
        return hashlib.shake_128(absorbed).digest(size)
    seed = hashlib.sha512(absorbed).digest()
    blocks = -(-size // 64)
    return b"".join(hashlib.sha512(seed + c.to_bytes(8, "little")).digest() for c in range(blocks))[:size]


def _nonce(order: int, absorbed: bytes, secret_scalar: int, hash_name: str = "sha512") -> int:
    secret_hash = _squeeze(absorbed + bytes([NONCE_EXPAND]) + (secret_scalar % order).to_bytes(32, "little"), 64, hash_name)
    wide = _squeeze(absorbed + bytes([NONCE]) + secret_hash, (order.bit_length() + 128 + 7) // 8, hash_name)
    value = int.from_bytes(wide, "little") % order
    if value == 0:
        raise ValueError("nonce scalar is zero")
    return value


def secret_scalar_from_seed(cv, seed: bytes) -> int:
    if len(seed) != 32:
        raise ValueError("seed must be exactly 32 bytes")
    order = cv.curve.params.subgroup_order
    base_secret = int.from_bytes(seed, "little") % order
    counter = 0
    while True:
        absorbed = cv.curve.params.suite_id + seed + (bytes([counter]) if counter else b"")
        try:
            return _nonce(order, absorbed, base_secret, cv.curve.params.hash_name)
        except ValueError:
            counter += 1
            if counter > 255:
                raise RuntimeError("failed to derive non-zero secret scalar") from None
