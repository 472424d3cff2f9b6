"""Ring-proof PIOP + KZG: params, ring, root, prover, verifier (oracle; test infrastructure only).

Restates, on Python ints and the oracle's own BLS12-381 / Bandersnatch code:
  dot_ring/ring_proof/params.py:118-287            RingProofParams
  dot_ring/vrf/ring/members.py:22-91               Ring
  dot_ring/vrf/ring/root.py:21-173                 RingRoot
  dot_ring/ring_proof/columns/columns.py:29-167    Column / WitnessColumnBuilder
  dot_ring/ring_proof/constraints/constraints.py:43-151   c1..c7 on the 4N domain
  dot_ring/ring_proof/proof_builder.py:38-315      prover pipeline
  dot_ring/ring_proof/proof_payload.py:68-143      592-byte payload codec
  dot_ring/ring_proof/verify.py:51-324             verifier scalar algebra + linearised KZG
  dot_ring/ring_proof/pcs/kzg.py:27-108,152-191,304-338   commit/open/batch verify
  dot_ring/ring_proof/pcs/srs.py:42-90             SRS file format
"""

from __future__ import annotations

import hashlib
import os
import secrets
from dataclasses import dataclass, field
from functools import lru_cache
from pathlib import Path

from . import bandersnatch as bs
from . import bls12_381 as bls
from . import fr
from .transcript import RingTranscript

FR = fr.R
ZK_ROWS = 3
SCALAR_BITS = bs.N.bit_length()  # 253
DEFAULT_SRS_PATH = Path(__file__).resolve().parent.parent / "dot_ring_b200" / "data" / "bls12-381-srs-2-11-uncompressed-zcash.bin"


# ---- SRS (srs.py:42-148) ------------------------------------------------------


@dataclass
class SRS:
    g1: list  # affine (x, y)
    g2: list  # two affine G2 points: [1]_2, [tau]_2


@lru_cache(maxsize=2)
def load_srs(path: str | None = None) -> SRS:
    data = Path(path or os.environ.get("DOT_RING_BLS12_381_SRS") or DEFAULT_SRS_PATH).read_bytes()
    n1 = int.from_bytes(data[:8], "little")
    g1 = []
    for i in range(n1):
        chunk = data[8 + 96 * i : 8 + 96 * (i + 1)]
        g1.append((int.from_bytes(chunk[:48], "big"), int.from_bytes(chunk[48:], "big")))
    off = 8 + 96 * n1
    n2 = int.from_bytes(data[off : off + 8], "little")
    if n2 < 2:
        raise ValueError("SRS file must contain at least two G2 points")
    g2 = [bls.g2_from_uncompressed(data[off + 8 + 192 * i : off + 8 + 192 * (i + 1)]) for i in range(2)]
    return SRS(g1, g2)


def kzg_commit(srs: SRS, coeffs):
    """kzg.py:152-175: all-zero -> infinity, else MSM over the first len(coeffs) SRS points."""
    if len(coeffs) > len(srs.g1):
        raise ValueError("polynomial degree exceeds SRS size")
    if not any(coeffs):
        return None
    return bls.g1_msm(srs.g1[: len(coeffs)], coeffs)


def kzg_open(srs: SRS, coeffs, x: int):
    """kzg.py:178-191."""
    q, y = fr.synthetic_div_with_eval(coeffs, x)
    return kzg_commit(srs, q), y


# ---- params (params.py) -------------------------------------------------------


def _next_pow2(n: int) -> int:
    if n <= 0:
        return 1
    return n if n & (n - 1) == 0 else 1 << n.bit_length()


@dataclass
class Params:
    domain_size: int = 512
    max_ring_size: int = 255
    padding_rows: int = 4
    test_vectors: bool = False
    suite: bs.Suite = field(default_factory=lambda: bs.SHA512)
    base_root: int = fr.ROOT_OF_UNITY_2048
    base_root_size: int = 2048
    max_domain_size: int = 4096

    def __post_init__(self):
        n = self.domain_size
        if n <= 0 or n & (n - 1):
            raise ValueError(f"domain_size must be a power of two, got {n}")
        if n > self.max_domain_size:  # the reference's limit (params.py:20,172-173); tests raise it for domains it rejects
            raise ValueError(f"domain_size {n} exceeds supported SRS domain size {self.max_domain_size}")
        self.radix_domain_size = 4 * n
        if self.radix_domain_size > self.base_root_size:
            self.base_root, self.base_root_size = fr.extend_root_to_size(self.base_root, self.base_root_size, self.radix_domain_size)
        if self.padding_rows != ZK_ROWS + 1:
            raise ValueError(f"padding_rows must be {ZK_ROWS + 1} to match the {ZK_ROWS} hidden rows")
        max_supported = n - SCALAR_BITS - self.padding_rows
        if max_supported <= 0:
            raise ValueError("domain_size is too small for the scalar bit decomposition")
        if self.max_ring_size == 255 and max_supported != 255:
            self.max_ring_size = max_supported
        elif self.max_ring_size > max_supported:
            raise ValueError(f"max_ring_size {self.max_ring_size} exceeds supported size {max_supported}")

    @classmethod
    def from_ring_size(cls, ring_size: int, test_vectors: bool = False, suite: bs.Suite = bs.SHA512, max_domain_size: int = 4096) -> "Params":
        if ring_size <= 0:
            raise ValueError(f"ring_size must be positive, got {ring_size}")
        overhead = SCALAR_BITS + 4
        n = _next_pow2(ring_size + overhead)
        return cls(domain_size=n, max_ring_size=n - overhead, test_vectors=test_vectors, suite=suite, max_domain_size=max_domain_size)

    @property
    def omega(self) -> int:
        return fr.omega_for_domain(self.domain_size, self.base_root, self.base_root_size)

    @property
    def radix_omega(self) -> int:
        return fr.omega_for_domain(self.radix_domain_size, self.base_root, self.base_root_size)

    @property
    def last_index(self) -> int:
        return self.domain_size - self.padding_rows

    def domain(self) -> list[int]:
        w, out, cur = self.omega, [], 1
        for _ in range(self.domain_size):
            out.append(cur)
            cur = cur * w % FR
        return out


# ---- ring + root (members.py, root.py) ---------------------------------------


class Ring:
    def __init__(self, keys, params: Params | None = None):
        if params is None:
            params = Params.from_ring_size(len(keys))
        self.params = params
        suite = params.suite
        if len(keys) > params.max_ring_size:
            raise ValueError(f"ring size {len(keys)} exceeds max supported size {params.max_ring_size}")
        pts = []
        for key in keys:
            pt = self._decode_key(key)
            pts.append(suite.padding_point if pt is None else pt)
        while len(pts) < params.max_ring_size:
            pts.append(suite.padding_point)
        fill = params.domain_size - params.padding_rows - len(pts)
        cur = suite.blinding_base
        for _ in range(max(fill, 0)):
            pts.append(cur)
            cur = bs.add(cur, cur) if cur != bs.IDENTITY else cur
        pts.extend([(0, 0)] * params.padding_rows)
        self.nm_points = tuple(pts)

    @staticmethod
    def _decode_key(key: bytes):
        try:
            pt = bs.dec_point(key)
        except ValueError:
            return None
        return None if bs.is_identity(pt) else pt

    def index_of(self, key: bytes) -> int:
        pt = self._decode_key(key)
        if pt is None:
            raise ValueError("invalid ring key")
        if pt == self.params.suite.padding_point:
            raise ValueError("producer key is not in ring")
        try:
            return self.nm_points[: self.params.max_ring_size].index(pt)
        except ValueError as exc:
            raise ValueError("producer key is not in ring") from exc


@dataclass
class Column:
    evals: list
    coeffs: list
    commitment: object  # Jacobian G1 or None (infinity)


def _interpolate(evals, params: Params, hidden: bool, zk_rows=None) -> Column:
    """columns.py:29-53.  ``zk_rows`` (3 ints) replaces secrets.randbelow for reproducible blinding."""
    n = params.domain_size
    ev = list(evals)
    if hidden and not params.test_vectors:
        cap = n - ZK_ROWS
        if len(ev) > cap:
            raise ValueError("evals length exceeds capacity")
        ev += [0] * (cap - len(ev))
        ev += list(zk_rows) if zk_rows is not None else [secrets.randbelow(FR) for _ in range(ZK_ROWS)]
    else:
        if len(ev) > n:
            raise ValueError("evals length exceeds column size")
        ev += [0] * (n - len(ev))
    return Column(ev, fr.inverse_fft(ev, params.omega), None)


@dataclass
class RingRoot:
    px: Column
    py: Column
    s: Column
    params: Params
    srs: SRS

    @classmethod
    def from_ring(cls, ring: Ring, params: Params | None = None, srs: SRS | None = None) -> "RingRoot":
        params = params or ring.params
        srs = srs or load_srs()
        cols = []
        sel = [1 if i < params.max_ring_size else 0 for i in range(params.domain_size)]
        for ev in ([p[0] for p in ring.nm_points], [p[1] for p in ring.nm_points], sel):
            col = _interpolate(ev, params, hidden=False)
            col.commitment = kzg_commit(srs, col.coeffs)
            cols.append(col)
        return cls(cols[0], cols[1], cols[2], params, srs)

    def encode(self) -> bytes:
        return b"".join(bls.g1_compress(c.commitment) for c in (self.px, self.py, self.s))

    def verifier_key_bytes(self) -> bytes:
        """phases.py:72-74 + root.py:54-71: G1[0] (96 B) | G2[0..1] (192 B each) | 3 fixed commitments (96 B each)."""
        out = srs_vk_prefix(self.srs)
        for c in (self.px, self.py, self.s):
            out += bls.g1_serialize(c.commitment)
        return out

    def transcript_prefix(self, label: bytes | None = None) -> RingTranscript:
        t = RingTranscript(label if label is not None else self.params.suite.suite_id)
        t.absorb_labeled(b"vk", self.verifier_key_bytes())
        return t


def srs_vk_prefix(srs: SRS) -> bytes:
    out = srs.g1[0][0].to_bytes(48, "big") + srs.g1[0][1].to_bytes(48, "big")
    for g2 in srs.g2:
        out += bls.g2_serialize(g2)
    return out


def decode_ring_root(data: bytes):
    if len(data) != 144:
        raise ValueError(f"invalid ring root length: ring root must be exactly 144 bytes, got {len(data)}")
    return [bls.g1_decompress(data[48 * i : 48 * i + 48]) for i in range(3)]


# ---- prover (columns.py, constraints.py, proof_builder.py) ------------------


@dataclass
class RingProof:
    c_b: object
    c_accip: object
    c_accx: object
    c_accy: object
    px_zeta: int
    py_zeta: int
    s_zeta: int
    b_zeta: int
    accip_zeta: int
    accx_zeta: int
    accy_zeta: int
    c_q: object
    l_zeta_omega: int
    phi_zeta: object
    phi_zeta_omega: object

    def encode(self) -> bytes:
        """proof_payload.py:68-91."""
        le = lambda v: int(v).to_bytes(32, "little")  # noqa: E731
        return b"".join(
            [bls.g1_compress(c) for c in (self.c_b, self.c_accip, self.c_accx, self.c_accy)]
            + [le(v) for v in (self.px_zeta, self.py_zeta, self.s_zeta, self.b_zeta, self.accip_zeta, self.accx_zeta, self.accy_zeta)]
            + [bls.g1_compress(self.c_q), le(self.l_zeta_omega), bls.g1_compress(self.phi_zeta), bls.g1_compress(self.phi_zeta_omega)]
        )

    @classmethod
    def decode(cls, data: bytes) -> "RingProof":
        """proof_payload.py:93-143."""
        if len(data) != 592:
            raise ValueError(f"invalid Ring VRF proof length: expected 592, got {len(data)}")
        off = 0

        def g1():
            nonlocal off
            pt = bls.g1_decompress(data[off : off + 48])
            off += 48
            return pt

        def sc():
            nonlocal off
            v = int.from_bytes(data[off : off + 32], "little")
            off += 32
            if v >= FR:
                raise ValueError("scalar is not canonical")
            return v

        c = [g1() for _ in range(4)]
        ev = [sc() for _ in range(7)]
        c_q = g1()
        lzw = sc()
        phi1, phi2 = g1(), g1()
        return cls(*c, *ev, c_q, lzw, phi1, phi2)

    @property
    def evaluations(self):
        return (self.px_zeta, self.py_zeta, self.s_zeta, self.b_zeta, self.accip_zeta, self.accx_zeta, self.accy_zeta)


def witness_vectors(ring: Ring, producer_index: int, secret_t: int):
    """columns.py:111-146: b (N-3 rows), acc_x/acc_y/acc_ip (N-3 rows each)."""
    p = ring.params
    n = p.domain_size
    bv = [1 if i == producer_index else 0 for i in range(p.max_ring_size)]
    bv += [int(ch) for ch in bin(secret_t)[2:][::-1]]
    pad_to = n - p.padding_rows
    if len(bv) > pad_to:
        raise ValueError("b vector length exceeds available rows")
    bv += [0] * (pad_to - len(bv)) + [0]
    acc = [p.suite.accumulator_base]
    accip = [0]
    sel = [1 if i < p.max_ring_size else 0 for i in range(n)]
    for i in range(1, pad_to + 1):
        acc.append(bs.add(acc[-1], ring.nm_points[i - 1]) if bv[i - 1] else acc[-1])
        accip.append(accip[-1] + bv[i - 1] * sel[i - 1])
    return bv, [a[0] for a in acc], [a[1] for a in acc], accip


def constraint_evals(params: Params, cols4, result_plus_seed):
    """constraints.py:43-151.  ``cols4`` = 4N-domain evaluations of (px, py, s, b, accx, accy, accip)."""
    n, n4 = params.domain_size, params.radix_domain_size
    w, w4 = params.omega, params.radix_omega
    px4, py4, s4, b4, ax4, ay4, aip4 = cols4
    last_root = pow(w, params.last_index, FR)
    radix_domain = [1] * n4
    for i in range(1, n4):
        radix_domain[i] = radix_domain[i - 1] * w4 % FR
    not_last = [(x - last_root) % FR for x in radix_domain]
    l0 = fr.evaluate_poly_fft(fr.lagrange_basis_coeffs(n, w, 0), n4, w4)
    ln = fr.evaluate_poly_fft(fr.lagrange_basis_coeffs(n, w, params.last_index), n4, w4)
    shift = n4 // n
    seed_x, seed_y = params.suite.accumulator_base
    rx, ry = result_plus_seed
    a = bs.A
    c = [[0] * n4 for _ in range(7)]
    for i in range(n4):
        j = (i + shift) % n4
        x1, y1, x2, y2, x3, y3 = ax4[i], ay4[i], px4[i], py4[i], ax4[j], ay4[j]
        bi, nl = b4[i], not_last[i]
        omb = (1 - bi) % FR
        x1y1, y2x2 = x1 * y1 % FR, y2 * x2 % FR
        c[0][i] = (aip4[j] - aip4[i] - bi * s4[i]) * nl % FR
        xt = (x3 * ((y1 * y2 + a * x1 * x2) % FR) - (x1y1 + y2x2)) % FR
        c[1][i] = (bi * xt + omb * (x3 - x1)) % FR * nl % FR
        yt = (y3 * ((x1 * y2 - x2 * y1) % FR) - (x1y1 - y2x2)) % FR
        c[2][i] = (bi * yt + omb * (y3 - y1)) % FR * nl % FR
        c[3][i] = bi * (1 - bi) % FR
        c[4][i] = ((x1 - seed_x) * l0[i] + (x1 - rx) * ln[i]) % FR
        c[5][i] = ((y1 - seed_y) * l0[i] + (y1 - ry) * ln[i]) % FR
        c[6][i] = (aip4[i] * l0[i] + (aip4[i] - 1) * ln[i]) % FR
    return c


def prove_ring(ring: Ring, root: RingRoot, producer_key: bytes, blinding: int, zk_rows=None, transcript_label: bytes | None = None) -> RingProof:
    """proof_builder.py:38-142.  ``zk_rows``: optional 12 ints in draw order b, accx, accy, accip (3 each)."""
    p = ring.params
    srs = root.srs
    n, n4 = p.domain_size, p.radix_domain_size
    w, w4 = p.omega, p.radix_omega
    k = ring.index_of(producer_key)

    bv, ax, ay, aip = witness_vectors(ring, k, blinding)
    zk = [None] * 4 if zk_rows is None else [list(zk_rows[3 * i : 3 * i + 3]) for i in range(4)]
    col_b, col_ax, col_ay, col_aip = (_interpolate(ev, p, hidden=True, zk_rows=z) for ev, z in zip((bv, ax, ay, aip), zk, strict=True))
    for col in (col_b, col_ax, col_ay, col_aip):
        col.commitment = kzg_commit(srs, col.coeffs)

    relation = bs.add(ring.nm_points[k], bs.mul(p.suite.blinding_base, blinding))
    result_plus_seed = bs.add(relation, p.suite.accumulator_base)

    # phase 1 (phases.py:18-26)
    t = root.transcript_prefix(transcript_label).copy()
    t.absorb_labeled(b"instance", relation[0].to_bytes(32, "little") + relation[1].to_bytes(32, "little"))
    t.absorb_labeled(b"committed_cols", b"".join(bls.g1_serialize(c.commitment) for c in (col_b, col_aip, col_ax, col_ay)))
    alphas = t.challenges(b"constraints_aggregation", 7)

    cols4 = [fr.evaluate_poly_fft(c.coeffs, n4, w4) for c in (root.px, root.py, root.s, col_b, col_ax, col_ay, col_aip)]
    cons = constraint_evals(p, cols4, result_plus_seed)

    # proof_builder.py:165-195
    agg = [0] * n4
    for cvec, alpha in zip(cons, alphas, strict=True):
        for i, v in enumerate(cvec):
            agg[i] = (agg[i] + v * alpha) % FR
    agg_poly = fr.inverse_fft(agg, w4)
    tail = [1]
    for off in range(1, ZK_ROWS + 1):
        tail = fr.poly_mul_small(tail, [(-pow(w, n - off, FR)) % FR, 1])
    c_agg = fr.poly_mul_small(agg_poly, tail)
    while c_agg and c_agg[-1] == 0:
        c_agg.pop()
    q_poly = fr.poly_divide_by_vanishing(c_agg, n)
    c_q = kzg_commit(srs, q_poly)

    # phase 2
    t.absorb_labeled(b"quotient", bls.g1_serialize(c_q))
    zeta = t.challenge(b"evaluation_point")
    zeta_omega = zeta * w % FR
    scalar_term = (zeta - pow(w, p.last_index, FR)) % FR

    ev = [fr.poly_evaluate_single(c.coeffs, zeta) for c in (root.px, root.py, root.s, col_b, col_aip, col_ax, col_ay)]
    px_z, py_z, s_z, b_z, aip_z, ax_z, ay_z = ev
    fx = (b_z * (ay_z * py_z + bs.A * ax_z * px_z) + (1 - b_z)) * scalar_term % FR
    fy = (b_z * (ax_z * py_z - px_z * ay_z) + (1 - b_z)) * scalar_term % FR
    l_agg = [0]
    for poly, f, alpha in ((col_aip.coeffs, scalar_term, alphas[0]), (col_ax.coeffs, fx, alphas[1]), (col_ay.coeffs, fy, alphas[2])):
        l_agg = fr.poly_add(l_agg, fr.poly_scalar_mul(fr.poly_scalar_mul(poly, f), alpha))
    l_zeta_omega = fr.poly_evaluate_single(l_agg, zeta_omega)

    # phase 3
    t.absorb_labeled(b"register_evaluations", b"".join(v.to_bytes(32, "little") for v in ev))
    t.absorb_labeled(b"shifted_linearization_evaluation", l_zeta_omega.to_bytes(32, "little"))
    nus = t.challenges(b"kzg_aggregation", 8)

    agg_open = [0]
    for poly, nu in zip((root.px.coeffs, root.py.coeffs, root.s.coeffs, col_b.coeffs, col_aip.coeffs, col_ax.coeffs, col_ay.coeffs, q_poly), nus, strict=True):
        agg_open = fr.poly_add(agg_open, fr.poly_scalar_mul(poly, nu))
    phi_zeta, _ = kzg_open(srs, agg_open, zeta)
    phi_zeta_omega, _ = kzg_open(srs, l_agg, zeta_omega)

    return RingProof(
        col_b.commitment, col_aip.commitment, col_ax.commitment, col_ay.commitment,
        px_z, py_z, s_z, b_z, aip_z, ax_z, ay_z, c_q, l_zeta_omega, phi_zeta, phi_zeta_omega,
    )  # fmt: skip


# ---- verifier (verify.py, kzg.py:56-108,304-338) ----------------------------


def verifier_challenges(prefix: RingTranscript, relation, proof: RingProof):
    """phases.py:46-69."""
    t = prefix.copy()
    t.absorb_labeled(b"instance", relation[0].to_bytes(32, "little") + relation[1].to_bytes(32, "little"))
    t.absorb_labeled(b"committed_cols", b"".join(bls.g1_serialize(c) for c in (proof.c_b, proof.c_accip, proof.c_accx, proof.c_accy)))
    alphas = t.challenges(b"constraints_aggregation", 7)
    t.absorb_labeled(b"quotient", bls.g1_serialize(proof.c_q))
    zeta = t.challenge(b"evaluation_point")
    t.absorb_labeled(b"register_evaluations", b"".join(v.to_bytes(32, "little") for v in proof.evaluations))
    t.absorb_labeled(b"shifted_linearization_evaluation", proof.l_zeta_omega.to_bytes(32, "little"))
    return alphas, zeta, t.challenges(b"kzg_aggregation", 8)


def linear_terms(params: Params, proof: RingProof, alphas, zeta: int, nus, seed, result_plus_seed):
    """verify.py:51-144 -> (agg_zeta, scalar_accip, scalar_accx, scalar_accy, zeta_omega)."""
    n, w = params.domain_size, params.omega
    d_last = pow(w, n - 4, FR)
    zn1 = (pow(zeta, n, FR) - 1) % FR
    zm1 = (zeta - 1) % FR
    zml = (zeta - d_last) % FR
    inv_n = pow(n, -1, FR)
    # L_0(zeta), L_{N-4}(zeta) and 1/(zeta^N - 1) (a zero denominator raises, as pow(x, -1, p) does in the reference)
    inv_zn1 = pow(zn1, -1, FR)
    l0 = 1 if zm1 == 0 else inv_n * zn1 % FR * pow(zm1, -1, FR) % FR
    ln = 1 if zml == 0 else d_last * inv_n % FR * zn1 % FR * pow(zml, -1, FR) % FR
    b, x1, y1, x2, y2 = proof.b_zeta, proof.accx_zeta, proof.accy_zeta, proof.px_zeta, proof.py_zeta
    omb = (1 - b) % FR
    x1y1, x2y2 = x1 * y1 % FR, x2 * y2 % FR
    cv = [
        -(proof.accip_zeta + b * proof.s_zeta) * zml % FR,
        (b * (-(x1y1 + x2y2)) + omb * (-x1)) * zml % FR,
        (b * (-(x1y1 - x2y2)) + omb * (-y1)) * zml % FR,
        b * omb % FR,
        ((x1 - seed[0]) * l0 + (x1 - result_plus_seed[0]) * ln) % FR,
        ((y1 - seed[1]) * l0 + (y1 - result_plus_seed[1]) * ln) % FR,
        (proof.accip_zeta * l0 + (proof.accip_zeta - 1) * ln) % FR,
    ]
    lin = sum(a * c for a, c in zip(alphas, cv, strict=True)) % FR
    prod = (zeta - pow(w, n - 1, FR)) * (zeta - pow(w, n - 2, FR)) % FR * (zeta - pow(w, n - 3, FR)) % FR
    q_zeta = (lin + proof.l_zeta_omega) * prod % FR * inv_zn1 % FR
    terms = (x2, y2, proof.s_zeta, b, proof.accip_zeta, x1, y1, q_zeta)
    agg_zeta = sum(v * tm for v, tm in zip(nus, terms, strict=True)) % FR
    cx = (b * ((y1 * y2 + bs.A * x1 * x2) % FR) + omb) % FR
    cy = (b * ((x1 * y2 - x2 * y1) % FR) + omb) % FR
    return agg_zeta, alphas[0] * zml % FR, alphas[1] * (cx * zml % FR) % FR, alphas[2] * (cy * zml % FR) % FR, zeta * w % FR


def linear_verifications(params: Params, fixed_commitments, prefix: RingTranscript, relation, proof: RingProof):
    """verify.py:147-210: two (commitment_terms, proof_point, eval_point, value) tuples."""
    seed = params.suite.accumulator_base
    rps = bs.add(seed, relation)
    alphas, zeta, nus = verifier_challenges(prefix, relation, proof)
    agg_zeta, s_ip, s_x, s_y, zeta_omega = linear_terms(params, proof, alphas, zeta, nus, seed, rps)
    cpx, cpy, cs = fixed_commitments
    quotient_terms = list(zip((cpx, cpy, cs, proof.c_b, proof.c_accip, proof.c_accx, proof.c_accy, proof.c_q), nus, strict=True))
    lin_terms = [(proof.c_accip, s_ip), (proof.c_accx, s_x), (proof.c_accy, s_y)]
    return [(quotient_terms, proof.phi_zeta, zeta, agg_zeta), (lin_terms, proof.phi_zeta_omega, zeta_omega, proof.l_zeta_omega)]


def random_batch_coefficients(count: int, seed: bytes | None = None) -> list[int]:
    """kzg.py:84-108 (first coefficient 1, rest rejection-sampled from SHAKE256(seed | ctr))."""
    if count <= 0:
        return []
    coeffs = [1]
    limit = (1 << 256) - ((1 << 256) % FR)
    seed = seed if seed is not None else secrets.token_bytes(32)
    counter = 0
    while len(coeffs) < count:
        raw = hashlib.shake_256(seed + counter.to_bytes(8, "little")).digest(32 * (count - len(coeffs)) * 2)
        counter += 1
        for off in range(0, len(raw), 32):
            cand = int.from_bytes(raw[off : off + 32], "big")
            if cand >= limit:
                continue
            if cand % FR:
                coeffs.append(cand % FR)
                if len(coeffs) == count:
                    break
    return coeffs


def batch_verify_linear(srs: SRS, verifications, coeffs=None) -> bool:
    """kzg.py:56-81,304-338: e(lhs, [1]_2) == e(rhs, [tau]_2)."""
    if not verifications:
        return True
    coeffs = coeffs or random_batch_coefficients(len(verifications))
    lhs_pts, lhs_sc, rhs_pts, rhs_sc = [], [], [], []
    sum_v = 0
    for coeff, (terms, proof_pt, point, value) in zip(coeffs, verifications, strict=False):
        for commitment, scalar in terms:
            lhs_pts.append(commitment)
            lhs_sc.append(coeff * scalar % FR)
        sum_v = (sum_v + coeff * value) % FR
        lhs_pts.append(proof_pt)
        lhs_sc.append(coeff * point % FR)
        rhs_pts.append(proof_pt)
        rhs_sc.append(coeff)
    lhs_pts.append((srs.g1[0][0], srs.g1[0][1], 1))
    lhs_sc.append((-sum_v) % FR)
    lhs = bls.g1_msm([bls.g1_to_affine(x) for x in lhs_pts], lhs_sc)
    rhs = bls.g1_msm([bls.g1_to_affine(x) for x in rhs_pts], rhs_sc)
    return bls.final_verify(bls.miller_loop(srs.g2[0], bls.g1_to_affine(lhs)), bls.miller_loop(srs.g2[1], bls.g1_to_affine(rhs)))


def verify_ring(params: Params, fixed_commitments, prefix: RingTranscript, relation, proof: RingProof, srs: SRS) -> bool:
    """verify.py:213-324 (``Verify.is_valid``)."""
    return batch_verify_linear(srs, linear_verifications(params, fixed_commitments, prefix, relation, proof))
