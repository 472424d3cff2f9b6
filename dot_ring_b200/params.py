"""RingProofParams for the B200 engine.

Drop-in for the parameter object of the reference (dot_ring/ring_proof/params.py:118-287): same field names, defaults,
derived properties and ValueError conditions, because callers construct it directly.  The only arithmetic that matters here
is the choice of domain generators: w_N is a power of the fixed 2048-th root of unity, and when the 4N evaluation domain is
larger than 2048 the root is extended by repeated square roots.  Which of the two square roots is taken decides every byte
of a proof, so `_tonelli_shanks` below follows the reference's walk (params.py:63-115: least quadratic non-residue as the
generator of the 2-Sylow subgroup, the textbook order-halving loop) and the tests compare the resulting roots with it.
"""

from __future__ import annotations

import dataclasses
import functools
from typing import Any

from .curve import Bandersnatch, CurveVariant

ROOT_OF_UNITY_2048 = 49307615728544765012166121802278658070711169839041683575071795236746050763237
DEFAULT_DOMAIN_SIZE = 512
DEFAULT_MAX_RING_SIZE = 255
ZK_ROWS = 3
MAX_PIOP_DOMAIN_SIZE = 4096  # the reference's limit (its bundled SRS has 3 * 2048 + 1 points)
EXTENDED_MAX_DOMAIN_SIZE = 65536  # what the engine itself handles when the caller brings a 3N + 1 point SRS


def _pow2(n: int) -> bool:
    return n >= 1 and (n & (n - 1)) == 0


def _ceil_pow2(n: int) -> int:
    return 1 if n <= 1 else 1 << (n - 1).bit_length()


# kept under the reference's helper names: its tests import them
_is_power_of_two = _pow2


def _next_power_of_two(n: int) -> int:
    return _ceil_pow2(n)


def _tonelli_shanks(a: int, p: int) -> int:
    """A square root of `a` mod the odd prime `p` -- specifically the one the reference's routine returns."""
    a %= p
    if a == 0:
        return 0
    if p & 3 == 3:
        return pow(a, (p + 1) >> 2, p)
    legendre = functools.partial(pow, exp=(p - 1) >> 1, mod=p)
    if legendre(a) != 1:
        raise ValueError("No square root exists for provided value")
    odd, twos = p - 1, 0
    while not odd & 1:
        odd >>= 1
        twos += 1
    non_residue = next(z for z in range(2, p) if legendre(z) == p - 1)
    gen, root, rest = pow(non_residue, odd, p), pow(a, (odd + 1) >> 1, p), pow(a, odd, p)
    level = twos
    while rest != 1:
        # order of `rest` is 2^k with k < level; multiply by the matching power of `gen` to halve it
        k, probe = 0, rest
        while probe != 1:
            probe = probe * probe % p
            k += 1
        step = pow(gen, 1 << (level - k - 1), p)
        gen = step * step % p
        root, rest, level = root * step % p, rest * gen % p, k
    return root


_sqrt_mod_prime = _tonelli_shanks


@functools.lru_cache(maxsize=8)
def _extend_root_to_size(base_root: int, base_size: int, target_size: int, prime: int) -> tuple[int, int]:
    """(root, size) with size >= target_size, obtained from (base_root, base_size) by successive square roots."""
    doublings = max(0, (target_size - 1).bit_length() - (base_size - 1).bit_length()) if target_size > base_size else 0
    root = base_root
    for _ in range(doublings):
        root = _tonelli_shanks(root, prime)
    return root, base_size << doublings


def _omega_for_domain(domain_size: int, prime: int, base_root: int, base_size: int = 2048) -> int:
    quotient, remainder = divmod(base_size, domain_size)
    if remainder:
        raise ValueError(f"Domain size {domain_size} must divide {base_size}")
    return pow(base_root, quotient, prime)


def _powers(generator: int, count: int, prime: int) -> list[int]:
    out, cur = [], 1
    for _ in range(count):
        out.append(cur)
        cur = cur * generator % prime
    return out


def _default_pcs():
    from .kzg import KZG

    return KZG


@dataclasses.dataclass
class RingProofParams:
    domain_size: int = DEFAULT_DOMAIN_SIZE
    max_ring_size: int = DEFAULT_MAX_RING_SIZE
    padding_rows: int = 4
    radix_domain_size: int | None = None
    base_root: int = ROOT_OF_UNITY_2048
    base_root_size: int = 2048
    pcs: Any = dataclasses.field(default_factory=_default_pcs, compare=False, hash=False, repr=False)
    test_vectors: bool = False
    cv: CurveVariant = dataclasses.field(default_factory=lambda: Bandersnatch, compare=False, hash=False)
    # Raise explicitly (up to EXTENDED_MAX_DOMAIN_SIZE) to use domains the reference rejects; needs an SRS with 3N + 1 points.
    max_domain_size: int = MAX_PIOP_DOMAIN_SIZE

    # ---- derived quantities ------------------------------------------------------------------------------------
    @property
    def prime(self) -> int:
        return self.cv.curve.params.field_modulus

    @property
    def scalar_bits(self) -> int:
        return self.cv.curve.params.subgroup_order.bit_length()

    @property
    def row_overhead(self) -> int:
        return self.scalar_bits + self.padding_rows

    @property
    def omega(self) -> int:
        return _omega_for_domain(self.domain_size, self.prime, self.base_root, self.base_root_size)

    @property
    def radix_omega(self) -> int:
        return _omega_for_domain(self.radix_domain_size, self.prime, self.base_root, self.base_root_size)

    @property
    def domain(self) -> list[int]:
        return _powers(self.omega, self.domain_size, self.prime)

    @property
    def radix_domain(self) -> list[int]:
        return _powers(self.radix_omega, self.radix_domain_size, self.prime)

    @property
    def radix_shift(self) -> int:
        return self.radix_domain_size // self.domain_size

    @property
    def last_index(self) -> int:
        return self.domain_size - self.padding_rows

    @property
    def max_effective_ring_size(self) -> int:
        return self.domain_size - self.row_overhead

    @property
    def required_srs_degree(self) -> int:
        return max(self.domain_size - 1, self.radix_domain_size - self.domain_size)

    # ---- validation (same conditions and messages as the reference, params.py:142-203) --------------------------------
    def __post_init__(self) -> None:
        aux = self.cv.curve.params.auxiliary_points
        missing = [name for name in ("blinding_base", "accumulator_base", "padding_point") if getattr(aux, name) is None]
        if missing:
            raise ValueError(f"{self.cv.name} ring proofs require auxiliary point {missing[0]}")
        if self.radix_domain_size is None:
            self.radix_domain_size = 4 * self.domain_size
        n, radix = self.domain_size, self.radix_domain_size
        limit = min(self.max_domain_size, EXTENDED_MAX_DOMAIN_SIZE)
        self._require(_pow2(n), f"domain_size must be a power of two, got {n}")
        self._require(_pow2(radix), f"radix_domain_size must be a power of two, got {radix}")
        self._require(radix % n == 0, f"domain_size {n} must divide radix_domain_size {radix}")
        self._require(n <= limit, f"domain_size {n} exceeds supported SRS domain size {limit}")
        self._require(radix == 4 * n, "the B200 engine evaluates constraints on exactly the 4N domain")
        self._require(radix > self.base_root_size or self.base_root_size % radix == 0, f"radix_domain_size {radix} must divide base_root_size {self.base_root_size}")
        primitive = pow(self.base_root, self.base_root_size, self.prime) == 1 and pow(self.base_root, self.base_root_size // 2, self.prime) != 1
        self._require(primitive, f"{self.cv.name} ring proofs require a primitive {self.base_root_size}-th root of unity")
        if radix > self.base_root_size:
            self.base_root, self.base_root_size = _extend_root_to_size(self.base_root, self.base_root_size, radix, self.prime)
        self._require(self.base_root_size % radix == 0, f"radix_domain_size {radix} must divide base_root_size {self.base_root_size}")
        self._require(self.padding_rows >= 1, "padding_rows must be >= 1 to preserve accumulator structure")
        self._require(self.padding_rows < n, "padding_rows must be less than domain_size")
        self._require(self.padding_rows == ZK_ROWS + 1, f"padding_rows must be {ZK_ROWS + 1} to match the {ZK_ROWS} hidden rows")
        capacity = n - self.row_overhead
        self._require(
            capacity > 0,
            "domain_size is too small for the scalar bit decomposition: "
            f"domain_size={n}, scalar_bits={self.scalar_bits}, padding_rows={self.padding_rows}",
        )
        if self.max_ring_size == DEFAULT_MAX_RING_SIZE != capacity:
            self.max_ring_size = capacity  # the default tracks the domain, as in the reference
        self._require(self.max_ring_size <= capacity, f"max_ring_size {self.max_ring_size} exceeds supported size {capacity}")

    @staticmethod
    def _require(condition: bool, message: str) -> None:
        if not condition:
            raise ValueError(message)

    @classmethod
    def from_ring_size(
        cls,
        ring_size: int,
        padding_rows: int = 4,
        base_root: int = ROOT_OF_UNITY_2048,
        base_root_size: int = 2048,
        test_vectors: bool = False,
        cv: CurveVariant = Bandersnatch,
        max_domain_size: int = MAX_PIOP_DOMAIN_SIZE,
    ) -> "RingProofParams":
        """Smallest power-of-two domain holding the ring, the scalar bits and the padding rows (params.py:244-287)."""
        if ring_size <= 0:
            raise ValueError(f"ring_size must be positive, got {ring_size}")
        rows_besides_keys = cv.curve.params.subgroup_order.bit_length() + padding_rows
        n = _ceil_pow2(ring_size + rows_besides_keys)
        return cls(domain_size=n, max_ring_size=n - rows_besides_keys, padding_rows=padding_rows, base_root=base_root, base_root_size=base_root_size,
                   test_vectors=test_vectors, cv=cv, max_domain_size=max_domain_size)  # fmt: skip
