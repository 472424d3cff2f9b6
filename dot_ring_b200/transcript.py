"""Host-side SHA-512 VRF transcript pieces needed outside the batched kernels.

Only key derivation (``secret_from_seed``) runs here: dot_ring/vrf/primitives.py:26-82,147-174.
Everything on the proving / verifying hot path hashes on the device (csrc/hash.cuh).

``FiatShamirTranscript`` is the host-side handle ``RingRoot.verifier_transcript_prefix()`` returns
(dot_ring/ring_proof/transcript/transcript.py:21-136): callers outside the engine can continue the
ring transcript from the verifier-key prefix; the prover / verifier kernels keep their own copy of the
same state on the device (csrc/ring.cuh ``Shake128``).
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


class FiatShamirTranscript:
    """SHAKE128 transcript with 4-byte big-endian length framing: ``label | be32(len(label))`` opens an item, its payload is
    closed by ``be32(len(payload))``; a challenge is 48 squeezed bytes (big-endian) mod ``modulus`` taken from a snapshot of
    the sponge, after which ``be32(9)`` (the length of ``b"challenge"``) is absorbed."""

    _TAG = b"challenge"

    def __init__(self, modulus: int, initial: bytes):
        self.modulus = int(modulus)
        self._sponge = hashlib.shake_128()
        self._squeeze_len = (self.modulus.bit_length() + 128 + 7) // 8
        self.label(initial)

    def copy(self) -> "FiatShamirTranscript":
        other = object.__new__(FiatShamirTranscript)
        other.modulus, other._squeeze_len, other._sponge = self.modulus, self._squeeze_len, self._sponge.copy()
        return other

    @staticmethod
    def _framed(label: bytes) -> bytes:
        return bytes(label) + len(label).to_bytes(4, "big")

    def label(self, lbl: bytes) -> None:
        self._sponge.update(self._framed(lbl))

    def absorb_labeled(self, label: bytes, data: bytes) -> None:
        data = bytes(data)
        if len(data) >> 31:
            raise ValueError("transcript items are limited to 2^31 - 1 bytes")
        self._sponge.update(self._framed(label) + data + len(data).to_bytes(4, "big"))

    def challenge(self, label: bytes) -> int:
        return self.challenges(label, 1)[0]

    def challenges(self, label: bytes, n: int) -> list[int]:
        out = []
        for _ in range(max(n, 0)):
            self._sponge.update(self._framed(label) + self._TAG)
            out.append(int.from_bytes(self._sponge.digest(self._squeeze_len), "big") % self.modulus)
            self._sponge.update(len(self._TAG).to_bytes(4, "big"))
        return out
