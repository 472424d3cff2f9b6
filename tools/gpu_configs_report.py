#!/usr/bin/env python
"""BASELINE.json configs 1-4 on one B200 with the CPU oracle timed beside each (bounded samples) -> gpurun_out/configs_report.json.

config 1  ring 8 / N=512: the 7 ark-vrf vectors, byte parity + timing of root / prove / verify
config 2  ring 1023 / N=2048: 4096 proofs proved then verified (per-item and aggregated)
config 3  RingRoot.from_ring for 1023 keys; fixed-base commit rate; variable-base MSM sweep 2^11..2^20
config 4  100k Tiny + 100k Pedersen verifications, 1 % corrupted, per-item verdicts
CPU column: the oracle port on ONE host core over a bounded sample (the reference's own CPU path cannot travel to the GPU box)."""
import ctypes, hashlib, json, os, random, sys, time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dot_ring_b200 import _native  # noqa: E402
from oracle import bandersnatch as bs, fr, ring_proof as rp, vrf as ovrf  # noqa: E402
from tests import msm_cases, verify_cases as cases  # noqa: E402
from tests.helpers import bench_ring_keys, hx, le64, load, ring_proof_bytes, split_keys  # noqa: E402
from tests.ring_fixtures import native_ring, native_srs  # noqa: E402

T = time.perf_counter
out = {"cpu": {"cores_used": 1, "kind": "port (oracle, C Pippenger for the MSMs)"}}
ctx = _native.Context(0)
out["device"] = ctx.device_info()
srs = native_srs(ctx, None, int(os.environ.get("DR_WINDOW_BITS", "14")))


def best(fn, reps=3):
    ts = []
    for _ in range(reps):
        t0 = T()
        r = fn()
        ts.append(T() - t0)
    return min(ts), r


# ---- config 1 ---------------------------------------------------------------------------------------------------
vs = load("bandersnatch_sha-512_ell2_ring.json")
params8 = rp.Params(test_vectors=True)
c1 = {"vectors": len(vs), "parity": True}
t_root = t_prove = t_verify = 0.0
for v in vs:
    keys = split_keys(hx(v, "ring_pks"))
    dt, ring = best(lambda: native_ring(srs, keys, params8), 1)
    t_root += dt
    c1["parity"] &= ring.root().hex() == v["ring_pks_com"]
    k = rp.Ring(keys, params8).index_of(hx(v, "pk"))
    dt, (proofs, st) = best(lambda: ring.prove_batch([hx(v, "alpha")], [hx(v, "ad")], [hx(v, "sk")], [k]))
    t_prove += dt
    c1["parity"] &= proofs[0] == ring_proof_bytes(v)
    dt, (verd, ok) = best(lambda: ring.verify_batch([hx(v, "alpha")], [hx(v, "ad")], proofs, cases.coeffs_for(1)))
    t_verify += dt
    c1["parity"] &= bool(ok)
    ring.close()
c1.update(gpu_ms_per_item={"ring_create+root": 1e3 * t_root / len(vs), "prove(1)": 1e3 * t_prove / len(vs), "verify(1)": 1e3 * t_verify / len(vs)})
v = vs[0]
keys = split_keys(hx(v, "ring_pks"))
t0 = T(); oring = rp.Ring(keys, params8); oroot = rp.RingRoot.from_ring(oring, params8); t_or = T() - t0
t0 = T(); op = ovrf.ring_prove(hx(v, "alpha"), hx(v, "ad"), hx(v, "sk"), hx(v, "pk"), oring, oroot); t_op = T() - t0
t0 = T(); okv = ovrf.ring_verify(op, hx(v, "alpha"), hx(v, "ad"), oring, oroot, ring_matches=True); t_ov = T() - t0
c1["cpu_ms_per_item"] = {"ring+root": 1e3 * t_or, "prove": 1e3 * t_op, "verify": 1e3 * t_ov, "sample": "vector 1"}
c1["reference_published_ms"] = {"root": 27.0, "prove": 152.3, "verify": 3.7, "hardware": "M1 Max, docs/BENCHMARK.md:63-65"}
out["config1_ring8"] = c1
print("config1", c1, flush=True)

# ---- config 2 / 3 (ring 1023) -----------------------------------------------------------------------------------
pk, sk, keys = bench_ring_keys(1023)
params = rp.Params.from_ring_size(1023)
dt_ring, ring = best(lambda: native_ring(srs, keys, params), 1)
g = load("ring1023_reference.json")
n = 4096
rng = random.Random(0)
al = [b"bench-batch-input" + le64(j) for j in range(n)]
ad = [b"bench-batch-ad" + le64(j) for j in range(n)]
zk = b"".join(rng.randrange(fr.R).to_bytes(32, "little") for _ in range(12 * n))
ring.prove_batch(al[:64], ad[:64], [sk] * 64, [3] * 64, zk_rows=zk[: 64 * 384])
dt_p, (proofs, st) = best(lambda: ring.prove_batch(al, ad, [sk] * n, [3] * n, zk_rows=zk), 2)
dev_ms = sum(ring.prove_phase_ms())
dt_v, (verd, ok) = best(lambda: ring.verify_batch(al, ad, proofs, cases.coeffs_for(n, 1)))
dt_a, (verd2, ok2) = best(lambda: ring.verify_batch(al, ad, proofs, cases.coeffs_for(n, 2, independent=False), aggregate=True))
out["config2_ring1023"] = {
    "root_parity": ring.root().hex() == g["ring_root"], "proofs": n, "prove_wall_s": dt_p, "prove_device_ms": dev_ms, "proofs_per_s_device": n / (dev_ms * 1e-3),
    "proofs_per_s_wall": n / dt_p, "verify_per_item_wall_s": dt_v, "verifies_per_s_per_item": n / dt_v, "verify_aggregated_wall_s": dt_a,
    "verifies_per_s_aggregated": n / dt_a, "all_valid": bool(ok and ok2 and not any(st)),
    "reference_published_ms": {"prove": 527.0, "verify": 3.81, "hardware": "M1 Max, docs/BENCHMARK.md:72-73"},
}
print("config2", out["config2_ring1023"], flush=True)
t0 = T(); oring = rp.Ring(keys, params); t_oring = T() - t0
t0 = T(); oroot = rp.RingRoot.from_ring(oring, params); t_oroot = T() - t0
c3 = {"ring_create_incl_root_gpu_s": dt_ring, "cpu_ring_ingest_s": t_oring, "cpu_root_s": t_oroot, "reference_published_root_ms": 327.1, "commit": [], "msm_sweep": []}
for nn, batch in ((2048, 4096), (6145, 1024), (2048, 1)):
    ms, _ = srs.commit_bench(nn, batch, 3, seed=nn)
    c3["commit"].append({"n": nn, "batch": batch, "ms": ms, "coefficients_per_s": nn * batch / (ms * 1e-3)})
imad_peak = ctx.microbench("imad", 20000)[0]
out["imad_peak_measured"] = imad_peak
for k in range(11, 21):
    nn = 1 << k
    ms, c, res = ctx.g1_msm_bench(nn, 3, 7, 0, msm_cases.TAU)
    # canonical Pippenger work: min_c ceil(255/c) (10 n + 14 2^c) Fq mul, 600 IMAD each
    canon = min(-(-255 // cc) * (nn * 10 + (1 << cc) * 14) for cc in range(2, 21)) * 600
    c3["msm_sweep"].append({"log2_n": k, "window_bits": c, "ms": ms, "points_per_s": nn / (ms * 1e-3), "parity": res == msm_cases.expected_synthetic(nn, 7, 0),
                            "roofline": {"bound": "imad", "achieved": canon / (ms * 1e-3) / 1e12, "peak": imad_peak / 1e12, "unit": "T IMAD/s (canonical Pippenger count, SURVEY 8d)",
                                         "frac": canon / (ms * 1e-3) / imad_peak}})
# CPU: the oracle's C Pippenger on 2^14 real points would need 16k SRS points; time 6145
from oracle import bls12_381 as bls  # noqa: E402
osrs = rp.load_srs()
ks = [rng.randrange(fr.R) for _ in range(6145)]
t0 = T(); bls.g1_msm([(p[0], p[1]) for p in osrs.g1[:6145]], ks); c3["cpu_msm_6145_points_per_s"] = 6145 / (T() - t0)
out["config3_root_and_msm"] = c3
print("config3", {k: v for k, v in c3.items() if k != "msm_sweep"}, flush=True)
for r in c3["msm_sweep"]:
    print("   ", r, flush=True)

# ---- config 4 ---------------------------------------------------------------------------------------------------
N = int(os.environ.get("N_VRF", "100000"))
su = cases.suite_struct()
sks = [hashlib.sha256(b"ietf-signer" + le64(i)).digest()[:31] + b"\x00" for i in range(N)]
alphas = [b"bench-ietf-input" + le64(i) for i in range(N)]
ads = [b"bench-ietf-ad" + le64(i) for i in range(N)]
pks = ctx.te_mul([bs.point_to_string(bs.GENERATOR)], [int.from_bytes(k, "little") for k in sks])
blob, a, b, c, d = _native.pack_items(alphas, ads)
L = ctx.library.lib
c4 = {}
for kind in ("tiny", "pedersen"):
    proofs = ctx.vrf_prove(kind, su, alphas, ads, sks)
    bad = [bytearray(p) for p in proofs]
    for i in range(0, N, 100):
        bad[i][-9] ^= 1
    joined = b"".join(bytes(x) for x in bad)
    outbuf = ctypes.create_string_buffer(N)
    def call():
        if kind == "tiny":
            return L.dr_tiny_verify_batch(ctx.handle, ctypes.byref(su), N, blob, a, b, c, d, b"".join(pks), joined, outbuf)
        return L.dr_pedersen_verify_batch(ctx.handle, ctypes.byref(su), N, blob, a, b, c, d, joined, outbuf)
    dt, rc = best(call, 4)
    verd = outbuf.raw
    good = rc == 0 and all((verd[i] != 1) == (i % 100 == 0) for i in range(N))
    # CPU: oracle verify of 5 items
    t0 = T()
    for i in range(1, 6):
        if kind == "tiny":
            ovrf.tiny_verify(bs.SHA512, ovrf.TinyProof.decode(proofs[i]), pks[i], alphas[i], ads[i])
        else:
            ovrf.pedersen_verify(bs.SHA512, ovrf.PedersenProof.decode(proofs[i]), alphas[i], ads[i])
    cpu = 5 / (T() - t0)
    canon_imad = 9000 * 272  # ~9 k Fr multiplications per verification (2 full + 2 half-length double-mults + ~4 sqrt / inversions), 272 IMAD each
    c4[kind] = {"n": N, "corrupted": len(range(0, N, 100)), "verdicts_exact": bool(good), "c_abi_host_buffers_s": dt, "verifies_per_s": N / dt, "cpu_verifies_per_s_1core": cpu,
                "reference_published_ms": {"tiny": 1.97, "pedersen": 1.74}[kind],
                "roofline": {"bound": "imad", "achieved": canon_imad * N / dt / 1e12, "peak": imad_peak / 1e12, "unit": "T IMAD/s (canonical 2.45 M IMAD per verification, SURVEY 8d; host buffers in the timed region)",
                             "frac": canon_imad * N / dt / imad_peak}}
    print("config4", kind, c4[kind], flush=True)
out["config4_vrf_batch"] = c4
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(os.environ.get("REPORT_OUT", "gpurun_out/configs_report.json"), "w"), indent=1)
