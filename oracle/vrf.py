"""Tiny / Pedersen / Ring VRF on Bandersnatch (oracle; test infrastructure only).

Restates:
  dot_ring/vrf/ietf/tiny.py:35-89           TinyVRF prove / verify / codec
  dot_ring/vrf/pedersen/vrf.py:44-242       PedersenVRF prove / verify / batch_verify / codec
  dot_ring/vrf/ring/vrf.py:51-294           RingVRF prove / verify / batch_verify / codec
  dot_ring/vrf/codec.py:9-51                scalar / point codecs
"""

from __future__ import annotations

from dataclasses import dataclass

from . import bandersnatch as bs
from . import ring_proof as rp
from . import transcript as tr


def dec_scalar(data: bytes) -> int:
    if len(data) != 32:
        raise ValueError("scalar must be exactly 32 bytes")
    v = int.from_bytes(data, "little")
    if v >= bs.N:
        raise ValueError("scalar is not canonical")
    return v


# ---- Tiny --------------------------------------------------------------------


@dataclass
class TinyProof:
    output_point: tuple
    c: int
    s: int

    def encode(self) -> bytes:
        return bs.point_to_string(self.output_point) + self.c.to_bytes(16, "little") + tr.enc_scalar(self.s)

    @classmethod
    def decode(cls, data: bytes) -> "TinyProof":
        if len(data) != 80:
            raise ValueError(f"invalid Tiny VRF proof length: expected 80, got {len(data)}")
        try:
            out = bs.dec_point(data[:32])
        except ValueError as exc:
            raise ValueError("Invalid output point") from exc
        return cls(out, tr.dec_scalar_mod(data[32:48]), dec_scalar(data[48:]))


def tiny_prove(suite: bs.Suite, alpha: bytes, secret_key: bytes, ad: bytes, salt: bytes = b"") -> TinyProof:
    x = tr.dec_scalar_mod(secret_key)
    pk = bs.mul(bs.GENERATOR, x)
    inp = bs.encode_to_curve(suite, alpha, salt)
    out = bs.mul(inp, x)
    t, merged = tr.vrf_transcript(suite, tr.TINY_VRF, [(bs.GENERATOR, pk), (inp, out)], ad)
    k = tr.nonce(suite, x, t)
    c = tr.challenge(suite, [bs.mul(merged[0], k)], t)
    return TinyProof(out, c, (k + c * x) % bs.N)


def tiny_verify(suite: bs.Suite, proof: TinyProof, public_key: bytes, alpha: bytes, ad: bytes, salt: bytes = b"") -> bool:
    inp = bs.encode_to_curve(suite, alpha, salt)
    try:
        pk = bs.dec_point(public_key)
    except ValueError as exc:
        raise ValueError("Invalid public key") from exc
    t, merged = tr.vrf_transcript(suite, tr.TINY_VRF, [(bs.GENERATOR, pk), (inp, proof.output_point)], ad)
    r = bs.msm([merged[0], merged[1]], [proof.s, -proof.c])
    return proof.c == tr.challenge(suite, [r], t)


# ---- Thin (ietf/thin.py:38-106) ---------------------------------------------------


@dataclass
class ThinProof:
    output_point: tuple
    r: tuple
    s: int

    def encode(self) -> bytes:
        return bs.point_to_string(self.output_point) + bs.point_to_string(self.r) + tr.enc_scalar(self.s)

    @classmethod
    def decode(cls, data: bytes) -> "ThinProof":
        if len(data) != 96:
            raise ValueError(f"invalid Thin VRF proof length: expected 96, got {len(data)}")
        return cls(bs.dec_point(data[:32]), bs.dec_point(data[32:64]), dec_scalar(data[64:]))


def thin_prove(suite: bs.Suite, alpha: bytes, secret_key: bytes, ad: bytes, salt: bytes = b"") -> ThinProof:
    x = tr.dec_scalar_mod(secret_key)
    pk = bs.mul(bs.GENERATOR, x)
    inp = bs.encode_to_curve(suite, alpha, salt)
    out = bs.mul(inp, x)
    t, merged = tr.vrf_transcript(suite, tr.THIN_VRF, [(bs.GENERATOR, pk), (inp, out)], ad)
    k = tr.nonce(suite, x, t)
    r = bs.mul(merged[0], k)
    c = tr.challenge(suite, [r], t)
    return ThinProof(out, r, (k + c * x) % bs.N)


def thin_verify(suite: bs.Suite, proof: ThinProof, public_key: bytes, alpha: bytes, ad: bytes, salt: bytes = b"") -> bool:
    inp = bs.encode_to_curve(suite, alpha, salt)
    try:
        pk = bs.dec_point(public_key)
    except ValueError as exc:
        raise ValueError("Invalid public key") from exc
    t, merged = tr.vrf_transcript(suite, tr.THIN_VRF, [(bs.GENERATOR, pk), (inp, proof.output_point)], ad)
    c = tr.challenge(suite, [proof.r], t)
    return bs.msm([merged[0], merged[1]], [proof.s, -c]) == proof.r


# ---- Pedersen ------------------------------------------------------------------


@dataclass
class PedersenProof:
    output_point: tuple
    blinded_pk: tuple
    result_point: tuple
    ok: tuple
    s: int
    sb: int
    blinding_factor: int = 0

    def encode(self) -> bytes:
        return b"".join(bs.point_to_string(p) for p in (self.output_point, self.blinded_pk, self.result_point, self.ok)) + tr.enc_scalar(
            self.s
        ) + tr.enc_scalar(self.sb)

    @classmethod
    def decode(cls, data: bytes) -> "PedersenProof":
        if len(data) != 192:
            raise ValueError(f"invalid Pedersen VRF proof length: expected 192, got {len(data)}")
        try:
            pts = [bs.dec_point(data[32 * i : 32 * i + 32]) for i in range(4)]
        except ValueError as exc:
            raise ValueError("Invalid point in proof") from exc
        return cls(*pts, dec_scalar(data[128:160]), dec_scalar(data[160:192]))


def pedersen_prove(suite: bs.Suite, alpha: bytes, secret_key: bytes, ad: bytes, salt: bytes = b"") -> PedersenProof:
    x = tr.dec_scalar_mod(secret_key)
    pk = bs.mul(bs.GENERATOR, x)
    inp = bs.encode_to_curve(suite, alpha, salt)
    out = bs.mul(inp, x)
    t, merged = tr.vrf_transcript(suite, tr.PEDERSEN_VRF, [(inp, out)], ad)
    tb = t.copy()
    tb.absorb(bytes([tr.PEDERSEN_BLINDING]))
    b = tr.nonce(suite, x, tb)
    blinded = bs.add(pk, bs.mul(suite.blinding_base, b))
    t.absorb(bs.point_to_string(blinded))
    k = tr.nonce(suite, x, t)
    kb = tr.nonce(suite, b, t)
    rpt = bs.msm([bs.GENERATOR, suite.blinding_base], [k, kb])
    ok = bs.mul(merged[0], k)
    c = tr.challenge(suite, [rpt, ok], t)
    return PedersenProof(out, blinded, rpt, ok, (k + c * x) % bs.N, (kb + c * b) % bs.N, b)


def _pedersen_challenge(suite: bs.Suite, proof: PedersenProof, alpha: bytes, ad: bytes, salt: bytes = b""):
    inp = bs.encode_to_curve(suite, alpha, salt)
    t, _ = tr.vrf_transcript(suite, tr.PEDERSEN_VRF, [(inp, proof.output_point)], ad)
    t.absorb(bs.point_to_string(proof.blinded_pk))
    return inp, tr.challenge(suite, [proof.result_point, proof.ok], t)


def pedersen_verify(suite: bs.Suite, proof: PedersenProof, alpha: bytes, ad: bytes, salt: bytes = b"") -> bool:
    inp, c = _pedersen_challenge(suite, proof, alpha, ad, salt)
    if bs.msm([inp, proof.output_point], [proof.s, -c]) != proof.ok:
        return False
    return bs.msm([bs.GENERATOR, suite.blinding_base, proof.blinded_pk], [proof.s, proof.sb, -c]) == proof.result_point


def pedersen_batch_verify(suite: bs.Suite, proofs, inputs, ads, salts=None) -> bool:
    """pedersen/vrf.py:171-242: one random-weighted MSM that must land on the identity."""
    salts = salts or [b""] * len(proofs)
    items, coeff_bytes = [], bytearray()
    for proof, alpha, ad, salt in zip(proofs, inputs, ads, salts, strict=True):
        inp, c = _pedersen_challenge(suite, proof, alpha, ad, salt)
        items.append((proof, inp, c))
        coeff_bytes += tr.enc_scalar(c) + tr.enc_scalar(proof.s) + tr.enc_scalar(proof.sb)
    if not items:
        return True
    weights = tr.squeeze_transcript_bytes(suite, suite.suite_id + bytes([tr.BATCH_VERIFY]) + bytes(coeff_bytes), 32 * len(items))
    pts, scs, gsc, bsc = [], [], 0, 0
    for i, (proof, inp, c) in enumerate(items):
        wio = tr.dec_scalar_mod(weights[32 * i : 32 * i + 16])
        wc = tr.dec_scalar_mod(weights[32 * i + 16 : 32 * i + 32])
        pts += [proof.ok, proof.output_point, inp, proof.result_point, proof.blinded_pk]
        scs += [wio, wio * c, -wio * proof.s, wc, wc * c]
        gsc = (gsc - wc * proof.s) % bs.N
        bsc = (bsc - wc * proof.sb) % bs.N
    pts += [bs.GENERATOR, suite.blinding_base]
    scs += [gsc, bsc]
    return bs.is_identity(bs.msm(pts, scs))


# ---- Ring ----------------------------------------------------------------------


@dataclass
class RingVrfProof:
    pedersen: PedersenProof
    ring: rp.RingProof

    def encode(self) -> bytes:
        return self.pedersen.encode() + self.ring.encode()

    @classmethod
    def decode(cls, data: bytes) -> "RingVrfProof":
        if len(data) != 784:
            raise ValueError(f"invalid Ring VRF proof length: Ring VRF proof must be exactly 784 bytes, got {len(data)}")
        return cls(PedersenProof.decode(data[:192]), rp.RingProof.decode(data[192:]))


def ring_prove(alpha: bytes, ad: bytes, secret_key: bytes, producer_key: bytes, ring: rp.Ring, root: rp.RingRoot, zk_rows=None, salt: bytes = b"") -> RingVrfProof:
    """vrf/ring/vrf.py:185-209."""
    suite = ring.params.suite
    if producer_key != tr.public_key_from_secret(secret_key):
        raise ValueError("producer_key does not match secret_key")
    ped = pedersen_prove(suite, alpha, secret_key, ad, salt)
    return RingVrfProof(ped, rp.prove_ring(ring, root, producer_key, ped.blinding_factor, zk_rows=zk_rows))


def ring_verify(proof: RingVrfProof, alpha: bytes, ad: bytes, ring: rp.Ring, root: rp.RingRoot, ring_matches: bool | None = None) -> bool:
    """vrf/ring/vrf.py:226-232 (``matches_ring`` recomputes the root; pass ``ring_matches`` to skip)."""
    suite = ring.params.suite
    ok_p = pedersen_verify(suite, proof.pedersen, alpha, ad)
    if ring_matches is None:
        ring_matches = rp.RingRoot.from_ring(ring, ring.params, root.srs).encode() == root.encode()
    if not ring_matches:
        return False
    fixed = (root.px.commitment, root.py.commitment, root.s.commitment)
    ok_r = rp.verify_ring(ring.params, fixed, root.transcript_prefix(), proof.pedersen.blinded_pk, proof.ring, root.srs)
    return ok_p and ok_r


def ring_batch_verify(proofs, inputs, ads, ring: rp.Ring, root: rp.RingRoot) -> bool:
    """vrf/ring/vrf.py:239-283."""
    suite = ring.params.suite
    if not pedersen_batch_verify(suite, [p.pedersen for p in proofs], inputs, ads):
        return False
    fixed = (root.px.commitment, root.py.commitment, root.s.commitment)
    prefix = root.transcript_prefix()
    ver = []
    for p in proofs:
        ver += rp.linear_verifications(ring.params, fixed, prefix, p.pedersen.blinded_pk, p.ring)
    return rp.batch_verify_linear(root.srs, ver)
