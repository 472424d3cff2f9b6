"""gmpy2 stand-in: the reference only uses mpz / invert / powmod as big-int speedups."""

mpz = int


def invert(a, m):
    return pow(int(a), -1, int(m))


def powmod(a, e, m):
    return pow(int(a), int(e), int(m))
