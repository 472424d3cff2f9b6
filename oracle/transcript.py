"""Transcripts (oracle; test infrastructure only).

* :class:`RingTranscript` restates the ark-transcript compatible SHAKE128 Fiat-Shamir
  transcript of dot_ring/ring_proof/transcript/transcript.py:21-136 and the phase helpers
  of dot_ring/ring_proof/transcript/phases.py:18-131.  Python's ``shake.digest`` does not
  finalise, so every challenge is SHAKE128(everything absorbed so far)[:48] as a big-endian
  integer mod r, after which absorption simply continues.
* :class:`VrfTranscript` and helpers restate dot_ring/vrf/primitives.py:26-174.
"""

from __future__ import annotations

import hashlib
import struct

from . import bandersnatch as bs

FR = bs.P


def _be32(n: int) -> bytes:
    return struct.pack(">I", n)


class RingTranscript:
    """transcript.py:21-136 restricted to inputs < 2^31 bytes (the only case on the path)."""

    def __init__(self, initial: bytes | None = None):
        self.buf = bytearray()
        self._length: int | None = None
        if initial is not None:
            self.label(initial)

    def copy(self) -> "RingTranscript":
        other = RingTranscript()
        other.buf = bytearray(self.buf)
        other._length = self._length
        return other

    def _separate(self) -> None:
        if self._length is not None:
            self.buf += _be32(self._length)
        self._length = None

    def label(self, lbl: bytes) -> None:
        self._separate()
        self.buf += lbl + _be32(len(lbl))
        self._length = None

    def absorb_labeled(self, label: bytes, data: bytes) -> None:
        self._separate()
        self.buf += label + _be32(len(label)) + data + _be32(len(data))
        self._length = None

    def _squeeze(self) -> int:
        return int.from_bytes(hashlib.shake_128(bytes(self.buf)).digest(48), "big") % FR

    def challenge(self, label: bytes) -> int:
        self._separate()
        self.buf += label + _be32(len(label)) + b"challenge"
        value = self._squeeze()
        self.buf += b"\x00\x00\x00\x09"
        return value

    def challenges(self, label: bytes, n: int) -> list[int]:
        if n <= 0:
            return []
        self._separate()
        prefix = label + _be32(len(label)) + b"challenge"
        out = []
        self.buf += prefix
        for i in range(n):
            out.append(self._squeeze())
            self.buf += b"\x00\x00\x00\x09" + (b"" if i == n - 1 else prefix)
        return out


# ---- VRF transcript (primitives.py) ----------------------------------------

TINY_VRF, THIN_VRF, PEDERSEN_VRF = 0x00, 0x01, 0x02
NONCE_EXPAND, NONCE, PEDERSEN_BLINDING = 0x10, 0x11, 0x12
POINT_TO_HASH, DELINEARIZE, CHALLENGE, BATCH_VERIFY = 0x20, 0x30, 0x40, 0x50
CHALLENGE_LEN = 16


def squeeze_transcript_bytes(suite: bs.Suite, absorbed: bytes, size: int) -> bytes:
    """primitives.py:165-174."""
    if suite.hash_name == "shake128":
        return hashlib.shake_128(absorbed).digest(size)
    seed = hashlib.sha512(absorbed).digest()
    blocks = -(-size // 64)
    return b"".join(hashlib.sha512(seed + c.to_bytes(8, "little")).digest() for c in range(blocks))[:size]


class VrfTranscript:
    """primitives.py:26-55: append-only, counter-mode squeeze, no absorb after squeeze."""

    def __init__(self, suite: bs.Suite, label: bytes | None = None):
        self.suite = suite
        self.absorbed = bytearray(suite.suite_id if label is None else label)
        self.seed: bytes | None = None
        self.offset = 0

    def copy(self) -> "VrfTranscript":
        other = VrfTranscript(self.suite, b"")
        other.absorbed = bytearray(self.absorbed)
        other.seed = self.seed
        other.offset = self.offset
        return other

    def absorb(self, data: bytes) -> None:
        if self.seed is not None:
            raise ValueError("cannot absorb after squeeze")
        self.absorbed += data

    def squeeze(self, size: int) -> bytes:
        if self.seed is None:
            self.seed = bytes(self.absorbed)
        start, end = self.offset, self.offset + size
        self.offset = end
        return squeeze_transcript_bytes(self.suite, self.seed, end)[start:end]


def enc_scalar(k: int) -> bytes:
    return (k % bs.N).to_bytes(32, "little")


def dec_scalar_mod(data: bytes) -> int:
    return int.from_bytes(data, "little") % bs.N


def nonce(suite: bs.Suite, secret_scalar: int, transcript: VrfTranscript | None = None) -> int:
    """primitives.py:66-82."""
    t = transcript.copy() if transcript is not None else VrfTranscript(suite)
    t_exp = t.copy()
    t_exp.absorb(bytes([NONCE_EXPAND]))
    t_exp.absorb(enc_scalar(secret_scalar))
    secret_hash = t_exp.squeeze(64)
    t.absorb(bytes([NONCE]))
    t.absorb(secret_hash)
    k = dec_scalar_mod(t.squeeze((bs.N.bit_length() + 128 + 7) // 8))
    if k == 0:
        raise ValueError("nonce scalar is zero")
    return k


def challenge(suite: bs.Suite, points, transcript: VrfTranscript | None = None) -> int:
    """primitives.py:85-91."""
    t = transcript.copy() if transcript is not None else VrfTranscript(suite)
    t.absorb(bytes([CHALLENGE]))
    for pt in points:
        t.absorb(bs.point_to_string(pt))
    return dec_scalar_mod(t.squeeze(CHALLENGE_LEN))


def point_to_hash(suite: bs.Suite, pt, size: int = 32) -> bytes:
    """primitives.py:94-99."""
    t = VrfTranscript(suite)
    t.absorb(bytes([POINT_TO_HASH]))
    t.absorb(bs.point_to_string(pt))
    return t.squeeze(size)


def vrf_transcript(suite: bs.Suite, scheme: int, ios, ad: bytes):
    """primitives.py:102-144: returns (transcript, merged (input, output))."""
    t = VrfTranscript(suite)
    t.absorb(bytes([scheme]))
    t.absorb(len(ios).to_bytes(8, "little"))
    for inp, out in ios:
        t.absorb(bs.point_to_string(inp) + bs.point_to_string(out))
    t.absorb(len(ad).to_bytes(8, "little"))
    t.absorb(ad)
    if not ios:
        return t, (bs.IDENTITY, bs.IDENTITY)
    if len(ios) == 1:
        return t, ios[0]
    td = t.copy()
    td.absorb(bytes([DELINEARIZE]))
    zs = [1] + [dec_scalar_mod(td.squeeze(CHALLENGE_LEN)) for _ in range(len(ios) - 1)]
    return t, (bs.msm([io[0] for io in ios], zs), bs.msm([io[1] for io in ios], zs))


def secret_from_seed(suite: bs.Suite, seed: bytes) -> tuple[bytes, bytes]:
    """curve.py:386-399 + primitives.py:147-162: returns (public_key, secret_key) bytes."""
    if len(seed) != 32:
        raise ValueError("seed must be exactly 32 bytes")
    base = dec_scalar_mod(seed)
    counter = 0
    while True:
        t = VrfTranscript(suite)
        t.absorb(seed)
        if counter:
            t.absorb(bytes([counter]))
        secret = nonce(suite, base, t)
        if secret:
            break
        counter += 1
    sk = enc_scalar(secret)
    return public_key_from_secret(sk), sk


def public_key_from_secret(secret_key: bytes) -> bytes:
    """curve.py:377-384."""
    return bs.point_to_string(bs.mul(bs.GENERATOR, int.from_bytes(secret_key, "little")))
