#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference in the build container.

Run (only where /root/reference is mounted):
    ./oracle/build_ref.sh && python tests/golden/generate_golden.py

The reference's Python sources are imported from /root/reference, its three Cython extensions
from oracle/_ref/ext (built by oracle/build_ref.sh from the sources where they lie), and the
absent third-party packages (blst, py_ecc, gmpy2) are replaced by oracle/ref_shims.  Outputs are
small JSON fixtures under tests/golden/ which travel with the repository; nothing in the GPU
tests, smoke() or bench.py reads /root/reference at run time.

Inputs follow the reference's own benchmark seeding (tests/benchmark/bench_ring_proof.py:47-77,
140-152; tests/benchmark/bench_ietf.py:40-45) so the same synthetic workload is used everywhere.
"""

from __future__ import annotations

import hashlib
import json
import random
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import ref_shims  # noqa: E402

OUT = Path(__file__).resolve().parent


def _seed(*parts) -> bytes:
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


def main() -> None:
    if not ref_shims.available():
        raise SystemExit("reference or oracle/_ref/ext missing: run ./oracle/build_ref.sh where /root/reference is mounted")
    ref_shims.install()
    import dot_ring.ring_proof.columns.columns as columns_mod
    from dot_ring.curve.specs.bandersnatch import Bandersnatch
    from dot_ring.ring_proof.params import RingProofParams
    from dot_ring.ring_proof.polynomial.fft import evaluate_poly_fft, inverse_fft
    from dot_ring.vrf.ietf.tiny import TinyVRF
    from dot_ring.vrf.pedersen.vrf import PedersenVRF
    from dot_ring.vrf.ring.members import Ring
    from dot_ring.vrf.ring.root import RingRoot
    from dot_ring.vrf.ring.vrf import RingVRF

    prime = Bandersnatch.curve.params.field_modulus
    t0 = time.time()

    # ---- 1. NTT known answers from the reference's Cython plan -------------------------------
    ntt = []
    for n in (8, 512, 2048, 8192):
        rng = random.Random(1000 + n)
        vals = [rng.randrange(prime) for _ in range(n)]
        params = RingProofParams.from_ring_size(1023) if n > 2048 else RingProofParams()
        omega = pow(params.base_root, params.base_root_size // n, prime)
        inv = inverse_fft(list(vals), omega, prime)
        fwd = evaluate_poly_fft(list(vals), n, omega, prime)
        ntt.append(
            {
                "n": n,
                "seed": 1000 + n,
                "omega": hex(omega),
                "inverse_sha256": hashlib.sha256(b"".join(v.to_bytes(32, "little") for v in inv)).hexdigest(),
                "forward_sha256": hashlib.sha256(b"".join(v.to_bytes(32, "little") for v in fwd)).hexdigest(),
                "inverse_head": [hex(v) for v in inv[:4]],
                "forward_head": [hex(v) for v in fwd[:4]],
            }
        )
    (OUT / "ntt_reference.json").write_text(json.dumps(ntt, indent=1))
    print("ntt", time.time() - t0)

    # ---- 2. Bandersnatch point operations ---------------------------------------------------------
    pt = Bandersnatch.point_type
    te = {"encode_to_curve": [], "scalar_mul": [], "dec_point": []}
    for i in range(16):
        alpha = b"golden-h2c" + le64(i)
        te["encode_to_curve"].append({"alpha": alpha.hex(), "point": pt.encode_to_curve(alpha).point_to_string().hex()})
    rng = random.Random(77)
    g = pt.generator_point()
    for i in range(16):
        k = rng.randrange(Bandersnatch.curve.params.subgroup_order)
        base = g * (i + 2)
        te["scalar_mul"].append({"base": base.point_to_string().hex(), "k": hex(k), "out": (base * k).point_to_string().hex()})
    from dot_ring.vrf.codec import dec_point

    for i in range(24):
        raw = hashlib.sha256(b"golden-dec" + le64(i)).digest()
        try:
            p = dec_point(Bandersnatch, raw)
            te["dec_point"].append({"raw": raw.hex(), "ok": True, "x": hex(int(p.x)), "y": hex(int(p.y))})
        except ValueError:
            te["dec_point"].append({"raw": raw.hex(), "ok": False})
    (OUT / "bandersnatch_reference.json").write_text(json.dumps(te, indent=1))
    print("te", time.time() - t0)

    # ---- 3. Tiny / Pedersen proofs for batched verification (bench_ietf.py seeding) -----------------
    vb = {"tiny": [], "pedersen": []}
    for i in range(32):
        pk, sk = Bandersnatch.secret_from_seed(_seed("ietf-signer", i))
        alpha, ad = b"bench-ietf-input" + le64(i), b"bench-ietf-ad" + le64(i)
        tp = TinyVRF[Bandersnatch].prove(alpha, sk, ad)
        pp = PedersenVRF[Bandersnatch].prove(alpha, sk, ad)
        assert tp.verify(pk, alpha, ad) and pp.verify(alpha, ad)
        vb["tiny"].append({"pk": pk.hex(), "sk": sk.hex(), "alpha": alpha.hex(), "ad": ad.hex(), "proof": tp.encode().hex()})
        vb["pedersen"].append({"pk": pk.hex(), "sk": sk.hex(), "alpha": alpha.hex(), "ad": ad.hex(), "proof": pp.encode().hex()})
    ped = [PedersenVRF[Bandersnatch].decode(bytes.fromhex(v["proof"])) for v in vb["pedersen"]]
    vb["pedersen_batch_all_valid"] = bool(
        PedersenVRF[Bandersnatch].batch_verify(ped, [bytes.fromhex(v["alpha"]) for v in vb["pedersen"]], [bytes.fromhex(v["ad"]) for v in vb["pedersen"]])
    )
    (OUT / "vrf_batch_reference.json").write_text(json.dumps(vb, indent=1))
    print("vrf", time.time() - t0)

    # ---- 4. Ring 1023 / N=2048 (bench_ring_proof.py:140-152 seeding) --------------------------------
    pk, sk = Bandersnatch.secret_from_seed(_seed("batch-signer", 0, 0))
    keys = [pk if i == 3 else Bandersnatch.secret_from_seed(_seed("ring-member", 0, i))[0] for i in range(1023)]
    out = {"signer_pk": pk.hex(), "signer_sk": sk.hex(), "signer_index": 3, "ring_size": 1023,
           "keys_sha256": hashlib.sha256(b"".join(keys)).hexdigest(), "first_keys": [k.hex() for k in keys[:8]]}
    params_tv = RingProofParams.from_ring_size(1023, test_vectors=True)
    ring = Ring(keys, params_tv)
    root = RingRoot.from_ring(ring, params_tv)
    out["domain_size"], out["max_ring_size"] = params_tv.domain_size, params_tv.max_ring_size
    out["radix_omega"] = hex(params_tv.radix_omega)
    out["ring_root"] = root.encode().hex()
    print("root1023", time.time() - t0)
    out["proofs_test_vectors"] = []
    for j in range(4):
        alpha, ad = b"bench-batch-input" + le64(j), b"bench-batch-ad" + le64(j)
        proof = RingVRF[Bandersnatch].prove(alpha, ad, sk, pk, ring, root)
        enc = proof.encode()
        if j == 0:
            assert proof.verify(alpha, ad, ring, root)
        out["proofs_test_vectors"].append({"j": j, "alpha": alpha.hex(), "ad": ad.hex(), "proof": enc.hex()})
        print("proof tv", j, time.time() - t0)
    # blinded rows: secrets.randbelow replaced by random.Random(0) (12 draws per proof: b, accx, accy, accip)
    params_zk = RingProofParams.from_ring_size(1023)
    ring_zk = Ring(keys, params_zk)
    root_zk = RingRoot.from_ring(ring_zk, params_zk)
    stream = random.Random(0)
    draws: list[int] = []

    class _Secrets:
        @staticmethod
        def randbelow(n):
            v = stream.randrange(n)
            draws.append(v)
            return v

    columns_mod.secrets = _Secrets
    out["proofs_blinded"] = []
    for j in range(2):
        alpha, ad = b"bench-batch-input" + le64(j), b"bench-batch-ad" + le64(j)
        start = len(draws)
        proof = RingVRF[Bandersnatch].prove(alpha, ad, sk, pk, ring_zk, root_zk)
        out["proofs_blinded"].append(
            {"j": j, "alpha": alpha.hex(), "ad": ad.hex(), "zk_rows": [hex(v) for v in draws[start:]], "proof": proof.encode().hex()}
        )
        print("proof zk", j, time.time() - t0)
    (OUT / "ring1023_reference.json").write_text(json.dumps(out, indent=1))

    # ---- 5. Ring 8 / N=512 blinded (ark-vrf vector-1 inputs, randbelow stream 1001..1012) ---------
    v = json.loads((OUT / "reference_vectors" / "bandersnatch_sha-512_ell2_ring.json").read_text())[0]
    keys8 = RingVRF[Bandersnatch].parse_keys(bytes.fromhex(v["ring_pks"]))
    params8 = RingProofParams()
    ring8 = Ring(keys8, params8)
    root8 = RingRoot.from_ring(ring8, params8)
    counter = iter(range(1001, 1013))

    class _Secrets2:
        @staticmethod
        def randbelow(n):
            return next(counter)

    columns_mod.secrets = _Secrets2
    proof8 = RingVRF[Bandersnatch].prove(bytes.fromhex(v["alpha"]), bytes.fromhex(v["ad"]), bytes.fromhex(v["sk"]), bytes.fromhex(v["pk"]), ring8, root8)
    r8 = {"vector": v["comment"], "zk_rows": list(range(1001, 1013)), "proof": proof8.encode().hex()}
    # verdict table from the reference verifier on tampered copies
    enc = bytearray(proof8.encode())
    verdicts = []
    for name, pos in (("intact", None), ("flip_l_zeta_omega", 192 + 4 * 48 + 7 * 32 + 48), ("flip_s", 128)):
        buf = bytearray(enc)
        if pos is not None:
            buf[pos] ^= 1
        try:
            ok = RingVRF[Bandersnatch].decode(bytes(buf)).verify(bytes.fromhex(v["alpha"]), bytes.fromhex(v["ad"]), ring8, root8)
        except ValueError:
            ok = "ValueError"
        verdicts.append({"case": name, "byte": pos, "verdict": ok})
    r8["verdicts"] = verdicts
    two = [RingVRF[Bandersnatch].decode(bytes(enc)), RingVRF[Bandersnatch].decode(bytes.fromhex("".join(v[k] for k in ("gamma", "proof_pk_com", "proof_r", "proof_ok", "proof_s", "proof_sb", "ring_proof"))))]
    r8["batch_verify_two"] = bool(RingVRF[Bandersnatch].batch_verify(two, [bytes.fromhex(v["alpha"])] * 2, [bytes.fromhex(v["ad"])] * 2, ring8, root8))
    (OUT / "ring8_reference.json").write_text(json.dumps(r8, indent=1))
    print("done", time.time() - t0)


if __name__ == "__main__":
    main()
