"""BLS12-381 scalar field Fr: NTT and polynomial helpers (oracle; test infrastructure only).

Restates, on Python ints:
  dot_ring/ring_proof/polynomial/fft.py:14-144      (bit-reverse, twiddles, DIT NTT, iNTT, LDE)
  dot_ring/ring_proof/polynomial/ntt.pyx:116-163    (gather by bit-reverse, log2 n rounds, optional scale)
  dot_ring/curve/native_field/bls12_381_scalar.c:333-356  (one radix-2 DIT round)
  dot_ring/ring_proof/polynomial/ops.py:51-224      (poly add / scalar mul / Horner / Lagrange / divide by X^N-1)
  dot_ring/ring_proof/params.py:12,35-115           (roots of unity, sqrt-extension of the base root)
"""

from __future__ import annotations

from functools import lru_cache

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
ROOT_OF_UNITY_2048 = 49307615728544765012166121802278658070711169839041683575071795236746050763237


@lru_cache(maxsize=64)
def bit_reverse_table(n: int) -> tuple[int, ...]:
    """fft.py:14-27."""
    bits = n.bit_length() - 1
    return tuple(int(format(i, f"0{bits}b")[::-1], 2) if bits else 0 for i in range(n))


def ntt(values, omega: int, scale: int = 1) -> list[int]:
    """Natural-order in/out radix-2 DIT NTT (fft.py:58-84, ntt.pyx:116-163).

    out[k] = scale * sum_j values[j] * omega^(j*k) mod r.
    """
    n = len(values)
    a = [values[r] % R for r in bit_reverse_table(n)] if n > 1 else [values[0] % R]
    m = 2
    while m <= n:
        half = m >> 1
        w_step = pow(omega, n // m, R)
        tw = [1] * half
        for j in range(1, half):
            tw[j] = tw[j - 1] * w_step % R
        for start in range(0, n, m):
            for j in range(half):
                u = a[start + j]
                t = a[start + j + half] * tw[j] % R
                a[start + j] = (u + t) % R
                a[start + j + half] = (u - t) % R
        m <<= 1
    if scale != 1:
        a = [x * scale % R for x in a]
    return a


def inverse_fft(values, omega: int) -> list[int]:
    """fft.py:87-103: forward transform with omega^-1, scaled by n^-1."""
    n = len(values)
    return ntt(list(values), pow(omega, -1, R), pow(n, -1, R))


def evaluate_poly_fft(poly, domain_size: int, omega: int) -> list[int]:
    """fft.py:106-144 with coset_offset == 1: fold mod X^n - 1, then NTT."""
    coeffs = [0] * domain_size
    for i, c in enumerate(poly):
        coeffs[i % domain_size] = (coeffs[i % domain_size] + c) % R
    return ntt(coeffs, omega)


def poly_evaluate_single(poly, x: int) -> int:
    """ops.py:170-176 (Horner)."""
    acc = 0
    for coef in reversed(poly):
        acc = (acc * x + coef) % R
    return acc


def poly_scalar_mul(poly, k: int) -> list[int]:
    """ops.py:68-86."""
    k %= R
    return [c % R * k % R for c in poly]


def poly_add(p1, p2) -> list[int]:
    """ops.py:51-65."""
    out = [0] * max(len(p1), len(p2))
    for i, c in enumerate(p1):
        out[i] = c
    for i, c in enumerate(p2):
        out[i] = (out[i] + c) % R
    return out


def poly_mul_small(p1, p2) -> list[int]:
    """Schoolbook product (ops.py:125-157 small path)."""
    out = [0] * (len(p1) + len(p2) - 1)
    for i, a in enumerate(p1):
        for j, b in enumerate(p2):
            out[i + j] = (out[i + j] + a * b) % R
    return out


def lagrange_basis_coeffs(n: int, omega: int, i: int) -> list[int]:
    """ops.py:179-204 fast path: L_i(X) = (1/n) * sum_j (X / w^i)^j."""
    inv_xi = pow(pow(omega, i, R), -1, R)
    cur = pow(n, -1, R)
    out = []
    for _ in range(n):
        out.append(cur)
        cur = cur * inv_xi % R
    return out


def poly_divide_by_vanishing(poly, n: int) -> list[int]:
    """ops.py:207-224: quotient by X^n - 1 via fold-add; coefficients are left unreduced
    (ints < 4r), exactly as the reference hands them to the MSM."""
    if len(poly) < n:
        return [0]
    q = list(poly[n:])
    for i in range(1, len(poly) // n):
        for j in range(len(q)):
            src = n * (i + 1) + j
            if src < len(poly):
                q[j] += poly[src]
    while q and q[-1] == 0:
        q.pop()
    return q


def synthetic_div_with_eval(poly, x: int):
    """pcs/utils.py:27-35: quotient by (X - x) and f(x) in one Horner pass."""
    n = len(poly)
    q = [0] * (n - 1)
    rem = poly[-1]
    for i in range(n - 2, -1, -1):
        q[i] = rem
        rem = (rem * x + poly[i]) % R
    return q, rem


# ---- roots of unity (params.py:35-115) --------------------------------------


def sqrt_mod_prime(n: int, prime: int = R) -> int:
    """Tonelli-Shanks exactly as params.py:63-103 walks it (the root it returns fixes the
    4N-domain generator for 4N > 2048, so the branch order is part of the contract)."""
    if n == 0:
        return 0
    if prime % 4 == 3:
        return pow(n, (prime + 1) // 4, prime)
    if pow(n, (prime - 1) // 2, prime) != 1:
        raise ValueError("No square root exists for provided value")
    q, s = prime - 1, 0
    while q % 2 == 0:
        s += 1
        q //= 2
    z = 2
    while pow(z, (prime - 1) // 2, prime) != prime - 1:
        z += 1
    m = s
    c = pow(z, q, prime)
    x = pow(n, (q + 1) // 2, prime)
    t = pow(n, q, prime)
    while t != 1:
        i = 1
        t2i = t * t % prime
        while i < m:
            if t2i == 1:
                break
            t2i = t2i * t2i % prime
            i += 1
        b = pow(c, 1 << (m - i - 1), prime)
        x = x * b % prime
        t = t * b * b % prime
        c = b * b % prime
        m = i
    return x


@lru_cache(maxsize=8)
def extend_root_to_size(base_root: int, base_size: int, target_size: int) -> tuple[int, int]:
    """params.py:107-115."""
    root, size = base_root, base_size
    while size < target_size:
        root = sqrt_mod_prime(root, R)
        size *= 2
    return root, size


def omega_for_domain(domain_size: int, base_root: int, base_size: int) -> int:
    """params.py:35-44."""
    if base_size % domain_size:
        raise ValueError(f"Domain size {domain_size} must divide {base_size}")
    return pow(base_root, base_size // domain_size, R)
