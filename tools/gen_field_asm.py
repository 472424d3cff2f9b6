#!/usr/bin/env python
"""Generate the PTX carry-chain kernels for Montgomery arithmetic on 32-bit limbs.

For each prime field used on the hot path (BLS12-381 Fq: 12 limbs, BLS12-381 Fr = Bandersnatch
base field: 8 limbs) this script builds straight-line programs over PTX's extended-precision
integer instructions (mad.lo.cc / madc.hi.cc / addc / subc ...), *executes them in Python* against
big-integer ground truth on random and edge-case operands, and only then renders them as one
inline-asm block per operation into ``dot_ring_b200/csrc/gen/field_asm.inc``.

Keeping each operation inside a single asm block keeps the condition-code carry chain intact
(nothing can be scheduled between two chained instructions at the PTX level), and ptxas fuses each
(mad.lo.cc, madc.hi.cc) pair on the same operands into one IMAD.WIDE.U32.X, so a 12-limb
multiplication is ~300 IMAD.WIDE plus a handful of adds.

Multiplication layout ("even/odd" operand scanning): the running value is kept in two limb arrays,
E at limb positions k and O at positions k+1.  For one multiplier limb b_i the products a_j*b_i
with even j chain cleanly through one array (lo -> position j, hi -> position j+1) and those with
odd j through the other, so every product is one carry-chained (lo, hi) pair.  The Montgomery
reduction step adds m*p the same way and the division by 2^32 is a role swap of the two arrays.
"""

from __future__ import annotations

import random
import sys
from pathlib import Path

M32 = 0xFFFFFFFF

FIELDS = {
    "fq": 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
    "fr": 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
    # Bandersnatch prime-subgroup order (scalar arithmetic of the VRF layer: s = k + c*x mod n ...)
    "fn": 0x1CFB69D4CA675F520CCE760202687600FF8F87007419047174FD06B52876E7E1,
}


def limbs(x: int, n: int) -> list[int]:
    return [(x >> (32 * i)) & M32 for i in range(n)]


class Prog:
    """A straight-line PTX program over named u32 registers with one carry flag."""

    def __init__(self):
        self.ins: list[tuple] = []

    def emit(self, op, dst, *src):
        self.ins.append((op, dst) + src)

    # -- interpreter ---------------------------------------------------------------
    def run(self, regs: dict[str, int]) -> dict[str, int]:
        r = dict(regs)
        cf = 0

        def val(x):
            return x if isinstance(x, int) else r[x]

        for ins in self.ins:
            op, dst = ins[0], ins[1]
            s = [val(x) for x in ins[2:]]
            if op == "mul.lo":
                r[dst] = (s[0] * s[1]) & M32
            elif op == "mul.hi":
                r[dst] = (s[0] * s[1]) >> 32
            elif op in ("mad.lo.cc", "madc.lo.cc", "madc.lo", "mad.hi.cc", "madc.hi.cc", "madc.hi", "mad.lo", "mad.hi"):
                prod = s[0] * s[1]
                part = (prod & M32) if ".lo" in op else (prod >> 32)
                t = part + s[2] + (cf if op.startswith("madc") else 0)
                r[dst] = t & M32
                if op.endswith(".cc"):
                    cf = t >> 32
            elif op in ("add.cc", "addc.cc", "addc", "add"):
                t = s[0] + s[1] + (cf if op.startswith("addc") else 0)
                r[dst] = t & M32
                if op.endswith(".cc"):
                    cf = t >> 32
            elif op in ("sub.cc", "subc.cc", "subc", "sub"):
                t = s[0] - s[1] - (cf if op.startswith("subc") else 0)
                r[dst] = t & M32
                if op.endswith(".cc"):
                    cf = 1 if t < 0 else 0
            elif op == "and":
                r[dst] = s[0] & s[1]
            elif op == "mov":
                r[dst] = s[0]
            elif op == "selnz":  # dst = s2 != 0 ? s0 : s1
                r[dst] = s[0] if s[2] != 0 else s[1]
            else:
                raise ValueError(op)
        return r

    # -- PTX rendering -------------------------------------------------------------
    def render(self, operand_index: dict[str, int]) -> list[str]:
        def nm(x):
            if isinstance(x, int):
                return f"0x{x:08x}"
            if x in operand_index:
                return f"%{operand_index[x]}"
            return x

        out = []
        for ins in self.ins:
            op, dst = ins[0], ins[1]
            s = [nm(x) for x in ins[2:]]
            d = nm(dst)
            if op == "and":
                out.append(f"and.b32 {d}, {s[0]}, {s[1]};")
            elif op == "mov":
                out.append(f"mov.u32 {d}, {s[0]};")
            elif op == "selnz":
                out.append(f"setp.ne.u32 pb, {s[2]}, 0;")
                out.append(f"selp.u32 {d}, {s[0]}, {s[1]}, pb;")
            else:
                out.append(f"{op}.u32 {d}, {', '.join(s)};")
        return out


def final_sub(p: Prog, n: int, src, dst, pl):
    """dst = src - p if src >= p else src  (src < 2p)."""
    for k in range(n):
        p.emit("sub.cc" if k == 0 else "subc.cc", f"s{k}", src[k], pl[k])
    p.emit("subc", "bm", 0, 0)  # 0xffffffff when the subtraction borrowed (src < p)
    for k in range(n):
        p.emit("selnz", dst[k], src[k], f"s{k}", "bm")


def prog_mul(n: int, pl: list[int], m0: int, square: bool = False) -> Prog:
    p = Prog()
    A = [f"a{k}" for k in range(n)]
    B = A if square else [f"b{k}" for k in range(n)]
    E = [f"e{k}" for k in range(n)]
    O = [f"o{k}" for k in range(n)]

    def cmad(acc, src, src_off, m):
        # acc[j], acc[j+1] += src[src_off + j] * m for even j; returns with the carry in CC
        for j in range(0, n, 2):
            p.emit("mad.lo.cc" if j == 0 else "madc.lo.cc", acc[j], src[src_off + j], m, acc[j])
            p.emit("madc.hi.cc", acc[j + 1], src[src_off + j], m, acc[j + 1])

    def reduce_step(lo_arr, hi_arr):
        # lo_arr sits at positions k, hi_arr at k+1; add mi*p so that lo_arr[0] becomes 0
        p.emit("mul.lo", "mi", lo_arr[0], m0)
        cmad(hi_arr, pl, 1, "mi")  # odd limbs of p; cannot carry out (value fits)
        cmad(lo_arr, pl, 0, "mi")
        p.emit("addc", hi_arr[n - 1], hi_arr[n - 1], 0)

    # first multiplier limb: plain products, no carries needed (disjoint lo/hi slots)
    for j in range(0, n, 2):
        p.emit("mul.lo", O[j], A[j + 1], B[0])
        p.emit("mul.hi", O[j + 1], A[j + 1], B[0])
    for j in range(0, n, 2):
        p.emit("mul.lo", E[j], A[j], B[0])
        p.emit("mul.hi", E[j + 1], A[j], B[0])
    reduce_step(E, O)
    lo_arr, hi_arr = O, E  # roles swap = division by 2^32 (old E[0] is zero)
    for i in range(1, n):
        bi = B[i]
        # old lo limb 1 lands on new position 0
        p.emit("add.cc", lo_arr[0], lo_arr[0], hi_arr[1])
        # new hi array = old lo array shifted down two limbs, plus odd-limb products
        for j in range(0, n - 2, 2):
            p.emit("madc.lo.cc", hi_arr[j], A[j + 1], bi, hi_arr[j + 2])
            p.emit("madc.hi.cc", hi_arr[j + 1], A[j + 1], bi, hi_arr[j + 3])
        p.emit("madc.lo.cc", hi_arr[n - 2], A[n - 1], bi, 0)
        p.emit("madc.hi", hi_arr[n - 1], A[n - 1], bi, 0)
        cmad(lo_arr, A, 0, bi)
        p.emit("addc", hi_arr[n - 1], hi_arr[n - 1], 0)
        reduce_step(lo_arr, hi_arr)
        lo_arr, hi_arr = hi_arr, lo_arr
    # merge: result = hi-array-of-last-step (now lo_arr, positions k) + other array shifted one limb
    res, oth = lo_arr, hi_arr
    p.emit("add.cc", res[0], res[0], oth[1])
    for k in range(1, n - 1):
        p.emit("addc.cc", res[k], res[k], oth[k + 1])
    p.emit("addc", res[n - 1], res[n - 1], 0)
    final_sub(p, n, res, [f"r{k}" for k in range(n)], pl)
    return p


def prog_sqr(n: int, pl: list[int], m0: int) -> Prog:
    """Montgomery squaring by product scanning (FIPS): column k of a^2 + m p is accumulated in three registers.  The cross products
    a_i a_j (i < j) are formed once and doubled, so a squaring issues (n^2 + n) / 2 + n^2 word products (lo + hi multiply-adds each)
    instead of the 2 n^2 of a multiplication: 456 vs 600 IMAD for 12 limbs, 208 vs 272 for 8; the extra work is additions with
    carry, which run on the ALU pipe next to the multiply-add pipe the kernels are bound by."""
    p = Prog()
    A = [f"a{k}" for k in range(n)]
    Q = [f"q{k}" for k in range(2 * n + 3)]  # column k lives in (q_k, q_k+1, q_k+2)
    M = [f"m{k}" for k in range(n)]

    def acc3(t, x, y):  # (t0, t1, t2) += x * y
        p.emit("mad.lo.cc", t[0], x, y, t[0])
        p.emit("madc.hi.cc", t[1], x, y, t[1])
        p.emit("addc", t[2], t[2], 0)

    for r in Q[:3]:
        p.emit("mov", r, 0)
    for k in range(2 * n):
        c = Q[k : k + 3]
        if k > 0:
            p.emit("mov", c[2], 0)
        pairs = [(i, k - i) for i in range(max(0, k - n + 1), (k + 1) // 2) if i < k - i]
        if pairs:
            X = ["x0", "x1", "x2"]
            i0, j0 = pairs[0]
            p.emit("mul.lo", X[0], A[i0], A[j0])
            p.emit("mul.hi", X[1], A[i0], A[j0])
            p.emit("mov", X[2], 0)
            for i, j in pairs[1:]:
                acc3(X, A[i], A[j])
            p.emit("add.cc", X[0], X[0], X[0])  # 2 x
            p.emit("addc.cc", X[1], X[1], X[1])
            p.emit("addc", X[2], X[2], X[2])
            p.emit("add.cc", c[0], c[0], X[0])  # c += x
            p.emit("addc.cc", c[1], c[1], X[1])
            p.emit("addc", c[2], c[2], X[2])
        if k % 2 == 0 and k // 2 < n:
            acc3(c, A[k // 2], A[k // 2])
        if k < n:
            for j in range(k):
                acc3(c, M[j], pl[k - j])
            p.emit("mul.lo", M[k], c[0], m0)
            acc3(c, M[k], pl[0])  # makes c[0] zero
        else:
            for j in range(k - n + 1, n):
                acc3(c, M[j], pl[k - j])
    final_sub(p, n, Q[n : 2 * n], [f"r{k}" for k in range(n)], pl)
    return p


def prog_add(n: int, pl) -> Prog:
    p = Prog()
    for k in range(n):
        p.emit("add.cc" if k == 0 else ("addc.cc" if k < n - 1 else "addc"), f"e{k}", f"a{k}", f"b{k}")
    final_sub(p, n, [f"e{k}" for k in range(n)], [f"r{k}" for k in range(n)], pl)
    return p


def prog_sub(n: int, pl) -> Prog:
    p = Prog()
    for k in range(n):
        p.emit("sub.cc" if k == 0 else "subc.cc", f"e{k}", f"a{k}", f"b{k}")
    p.emit("subc", "bm", 0, 0)
    for k in range(n):
        p.emit("and", f"s{k}", "bm", pl[k])
    for k in range(n):
        p.emit("add.cc" if k == 0 else ("addc.cc" if k < n - 1 else "addc"), f"r{k}", f"e{k}", f"s{k}")
    return p


def check(name: str, prime: int, n: int, trials: int = 3000) -> dict[str, Prog]:
    pl = limbs(prime, n)
    m0 = (-pow(prime, -1, 1 << 32)) & M32
    rinv = pow(1 << (32 * n), -1, prime)
    progs = {"mul": prog_mul(n, pl, m0), "sqr": prog_sqr(n, pl, m0), "add": prog_add(n, pl), "sub": prog_sub(n, pl)}
    rng = random.Random(0xD07)
    edge = [0, 1, 2, prime - 1, prime - 2, (1 << 32) - 1, 1 << 32, (1 << (32 * n - 1)) % prime, prime >> 1, (prime >> 1) + 1]
    cases = [(x, y) for x in edge for y in edge] + [(rng.randrange(prime), rng.randrange(prime)) for _ in range(trials)]
    for a, b in cases:
        regs = {f"a{k}": v for k, v in enumerate(limbs(a, n))}
        regs.update({f"b{k}": v for k, v in enumerate(limbs(b, n))})
        for op, prog in progs.items():
            out = prog.run(regs)
            got = sum(out[f"r{k}"] << (32 * k) for k in range(n))
            want = {"mul": a * b * rinv % prime, "sqr": a * a * rinv % prime, "add": (a + b) % prime, "sub": (a - b) % prime}[op]
            if got != want:
                raise SystemExit(f"{name}.{op} mismatch for a={a:#x} b={b:#x}: got {got:#x} want {want:#x}")
    imad = lambda pr: sum(1 for ins in pr.ins if ins[0].startswith(("mul.", "mad", "madc")))  # noqa: E731
    print(f"{name}: {len(cases)} cases x {len(progs)} ops verified; mul = {len(progs['mul'].ins)} PTX instructions ({imad(progs['mul'])} multiply-adds), "
          f"sqr = {len(progs['sqr'].ins)} ({imad(progs['sqr'])} multiply-adds)", file=sys.stderr)
    return progs


def render_fn(fname: str, n: int, prog: Prog, nin: int) -> str:
    opidx = {f"r{k}": k for k in range(n)}
    opidx.update({f"a{k}": n + k for k in range(n)})
    if nin == 2:
        opidx.update({f"b{k}": 2 * n + k for k in range(n)})
    body = prog.render(opidx)
    lines = [f"__device__ __forceinline__ void {fname}(uint32_t* __restrict__ r, const uint32_t* a" + (", const uint32_t* b" if nin == 2 else "") + ") {"]
    lines.append("    asm(\"{\\n\\t\"")
    lines.append(f"        \".reg .u32 e<{n}>, o<{n}>, s<{n}>, q<{2 * n + 3}>, m<{n}>, x<3>, mi, bm;\\n\\t\"")
    lines.append("        \".reg .pred pb;\\n\\t\"")
    for ln in body:
        lines.append(f"        \"{ln}\\n\\t\"")
    lines.append("        \"}\"")
    outs = ", ".join(f"\"=r\"(r[{k}])" for k in range(n))
    ins = ", ".join(f"\"r\"(a[{k}])" for k in range(n))
    if nin == 2:
        ins += ", " + ", ".join(f"\"r\"(b[{k}])" for k in range(n))
    lines.append(f"        : {outs}")
    lines.append(f"        : {ins});")
    lines.append("}")
    return "\n".join(lines)


def main() -> None:
    out_dir = Path(__file__).resolve().parents[1] / "dot_ring_b200" / "csrc" / "gen"
    out_dir.mkdir(parents=True, exist_ok=True)
    parts = [
        "// GENERATED by tools/gen_field_asm.py -- do not edit.  Every program below was executed in",
        "// Python against big-integer ground truth before being rendered.",
        "#pragma once",
        "#include <cstdint>",
        "#if defined(__CUDA_ARCH__)",
    ]
    for name, prime in FIELDS.items():
        n = (prime.bit_length() + 31) // 32
        progs = check(name, prime, n)
        parts.append(f"// ---- {name}: {n} limbs ----")
        parts.append(render_fn(f"{name}_mul_ptx", n, progs["mul"], 2))
        parts.append(render_fn(f"{name}_sqr_ptx", n, progs["sqr"], 1))
        parts.append(render_fn(f"{name}_add_ptx", n, progs["add"], 2))
        parts.append(render_fn(f"{name}_sub_ptx", n, progs["sub"], 2))
    parts.append("#endif  // __CUDA_ARCH__")
    (out_dir / "field_asm.inc").write_text("\n".join(parts) + "\n")

    # constants header (Montgomery R, R^2, -p^-1 mod 2^32, ...)
    c = ["// GENERATED by tools/gen_field_asm.py -- do not edit.", "#pragma once", "#include <cstdint>"]
    for name, prime in FIELDS.items():
        n = (prime.bit_length() + 31) // 32
        rr = 1 << (32 * n)

        def arr(x):
            return "{" + ", ".join(f"0x{v:08x}u" for v in limbs(x, n)) + "}"

        c.append(f"#define DR_{name.upper()}_LIMBS {n}")
        c.append(f"#define DR_{name.upper()}_MOD {arr(prime)}")
        c.append(f"#define DR_{name.upper()}_R1 {arr(rr % prime)}")
        c.append(f"#define DR_{name.upper()}_R2 {arr(rr * rr % prime)}")
        c.append(f"#define DR_{name.upper()}_M0 0x{(-pow(prime, -1, 1 << 32)) & M32:08x}u")
    # Bandersnatch / Elligator2 / Tonelli-Shanks constants over Fr (Montgomery form unless noted)
    P = FIELDS["fr"]
    RR = 1 << 256

    def m8(x):
        return "{" + ", ".join(f"0x{v:08x}u" for v in limbs(x % P * RR % P, 8)) + "}"

    def raw8(x):
        return "{" + ", ".join(f"0x{v:08x}u" for v in limbs(x, 8)) + "}"

    te_a = -5 % P
    te_d = 0x6389C12633C267CBC66E3BF86BE3B6D8CB66677177E54F92B369F2F5188D58E7
    inv_amd = pow((te_a - te_d) % P, -1, P)
    mont_a = 2 * (te_a + te_d) * inv_amd % P
    mont_b = 4 * inv_amd % P
    q, s2 = P - 1, 0
    while q % 2 == 0:
        q //= 2
        s2 += 1
    assert s2 == 32 and pow(5, (P - 1) // 2, P) == P - 1
    # prime-subgroup test by 2-descent (te.cuh te_in_prime_subgroup): s = sqrt(a d) exists because a and d are both non-squares
    s_ad = pow(te_a * te_d % P, (q + 1) // 2, P)
    t_ad, cc, mm = pow(te_a * te_d % P, q, P), pow(5, q, P), 32
    while t_ad != 1:  # Tonelli-Shanks
        i, t2 = 0, t_ad
        while t2 != 1:
            t2, i = t2 * t2 % P, i + 1
        b = pow(cc, 1 << (mm - i - 1), P)
        s_ad, cc, mm = s_ad * b % P, b * b % P, i
        t_ad = t_ad * cc % P
    assert s_ad * s_ad % P == te_a * te_d % P
    c.append(f"#define DR_TE_A_MINUS_D {m8(te_a - te_d)}")
    c.append(f"#define DR_TE_A_MINUS_S {m8(te_a - s_ad)}  // a - sqrt(a d)")
    c.append(f"#define DR_TE_D_MINUS_S {m8(te_d - s_ad)}  // d - sqrt(a d)")
    # GLV on Bandersnatch (curve/glv.py:128-189, specs/bandersnatch.py:65-67): endomorphism constants (Montgomery) and the short
    # lattice basis v1 = (a1, b1), v2 = (a2, -a1) of {(x, y): x + y lambda = 0 mod n} with the two Barrett factors floor(2^256 |b| / n)
    glv_lambda = 0x13B4F3DC4A39A493EDF849562B38C72BCFC49DB970A5056ED13D21408783DF05
    glv_b = 0x52C9F28B828426A561F00D3A63511A882EA712770D9AF4D6EE0F014D172510B4
    glv_c = 0x6CC624CF865457C3A97C6EFD6C17D1078456ABCFFF36F4E9515C806CDF650B3D
    order = FIELDS["fn"]
    a1, b1, a2 = 0x555FE2004BE6928E4B02F94A9789181F, 0x0814B3EEE55E8F5DF8E2591A23D61F44, 0x102967DDCABD1EBBF1C4B23447AC3E88
    assert (a1 + b1 * glv_lambda) % order == 0 and (a2 - a1 * glv_lambda) % order == 0 and a1 * a1 + b1 * a2 == order
    raw = lambda x, k: "{" + ", ".join(f"0x{v:08x}u" for v in limbs(x, k)) + "}"  # noqa: E731
    c.append(f"#define DR_TE_GLV_B {m8(glv_b)}")
    c.append(f"#define DR_TE_GLV_C {m8(glv_c)}")
    c.append(f"#define DR_TE_GLV_A1 {raw(a1, 4)}")
    c.append(f"#define DR_TE_GLV_B1 {raw(b1, 4)}")
    c.append(f"#define DR_TE_GLV_A2 {raw(a2, 4)}")
    c.append(f"#define DR_TE_GLV_G1 {raw((a1 << 256) // order, 5)}  // floor(2^256 a1 / n)")
    c.append(f"#define DR_TE_GLV_G2 {raw((b1 << 256) // order, 4)}  // floor(2^256 b1 / n)")
    c.append(f"#define DR_TE_D {m8(te_d)}")
    c.append(f"#define DR_TE_2D {m8(2 * te_d)}")
    c.append(f"#define DR_ELL2_A_OVER_B {m8(mont_a * pow(mont_b, -1, P))}")
    c.append(f"#define DR_ELL2_INV_B2 {m8(pow(mont_b * mont_b, -1, P))}")
    c.append(f"#define DR_ELL2_B {m8(mont_b)}")
    c.append(f"#define DR_FR_TS_C {m8(pow(5, q, P))}  // 5^q: generator of the 2^32 roots of unity")
    c.append(f"#define DR_FR_TS_Z_HALF {m8(pow(5, (q + 1) // 2, P))}  // 5^((q+1)/2): turns the Tonelli-Shanks state of a into that of 5a")
    c.append(f"#define DR_FR_TS_QM1_HALF {raw8((q - 1) // 2)}  // (q-1)/2, raw limbs")
    c.append(f"#define DR_FR_PM1_HALF {raw8((P - 1) // 2)}  // (p-1)/2, raw limbs")
    c.append(f"#define DR_FN_RAW {raw8(FIELDS['fn'])}")
    c.append(f"#define DR_FQ_SQRT_EXP " + "{" + ", ".join(f"0x{v:08x}u" for v in limbs((FIELDS['fq'] + 1) // 4, 12)) + "}  // (p+1)/4")
    (out_dir / "field_consts.h").write_text("\n".join(c) + "\n")


if __name__ == "__main__":
    main()
