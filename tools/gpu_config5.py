#!/usr/bin/env python
"""BASELINE configs[4] kernels: the 2^16 domain of a ~65k-key ring (outside the reference, which rejects domain_size > 4096:
ring_proof/params.py:172-173).  Fr NTT / 4x LDE sizes on one GPU and the (3N+1)-point KZG MSM split by point range across the
ranks with one 96-byte all-gather.  Run alone or under torchrun:

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/gpu_config5.py

Correctness: the MSM over the synthetic SRS tau^i * G must equal (sum k_i tau^i) * G (oracle scalar multiplication); the large
NTT is checked against the oracle in tests/."""
import json, os, random, sys, time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, local, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
dist = None
if world > 1:
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
from dot_ring_b200 import _native  # noqa: E402
from dot_ring_b200.params import ROOT_OF_UNITY_2048, _extend_root_to_size, _omega_for_domain  # noqa: E402
from oracle import bls12_381 as bls, fr  # noqa: E402
from tests import msm_cases  # noqa: E402

ctx = _native.Context(local)
N = 1 << 16
out = {"world": world, "domain": N}
if rank == 0:
    imad_peak = ctx.microbench("imad", 20000)[0]
    rows = []
    for n, batch in ((2048, 16384), (4096, 8192), (8192, 4096), (N, 64), (4 * N, 16)):
        root, size = _extend_root_to_size(ROOT_OF_UNITY_2048, 2048, max(n, 2048), fr.R)
        om = _omega_for_domain(n, fr.R, root, size)
        ms, _ = ctx.fr_ntt_bench(n, batch, 5, om)
        elems = n * batch
        logn = n.bit_length() - 1
        passes = 1 if n <= 4096 else 2
        rows.append({"n": n, "batch": batch, "ms": ms, "elements_per_s": elems / (ms * 1e-3), "algorithmic_GBps": elems * 64 / (ms * 1e-3) / 1e9,
                     "hbm_traffic_GBps": elems * 64 * passes / (ms * 1e-3) / 1e9, "imad_per_s": elems / 2 * logn * 272 / (ms * 1e-3),
                     "frac_of_imad_peak": elems / 2 * logn * 272 / (ms * 1e-3) / imad_peak, "passes": passes})
        print(rows[-1], flush=True)
    out["ntt"] = rows
    out["imad_peak"] = imad_peak

# ---- MSM of 3N + 1 points split by point range --------------------------------------------------------------
total = 3 * N + 1
lo, hi = rank * total // world, (rank + 1) * total // world
rng = random.Random(1234)
scalars_all = [rng.randrange(fr.R) for _ in range(total)]  # same stream on every rank
pts = ctx.g1_synthetic_srs(msm_cases.TAU, lo, hi - lo)
ks = b"".join(k.to_bytes(32, "little") for k in scalars_all[lo:hi])  # wire encoding, packed outside the timed region
ctx.g1_msm(pts[: 96 * 64], ks[: 32 * 64])  # warm-up
if dist is not None:
    dist.barrier(device_ids=[local])
t0 = time.perf_counter()
partial = ctx.g1_msm(pts, ks)
if dist is not None:
    import torch

    mine = torch.tensor(list(partial), dtype=torch.uint8, device=torch.device("cuda", local))
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    parts = b"".join(bytes(g.cpu().tolist()) for g in gathered)
else:
    parts = partial
result = ctx.g1_msm(parts, [1] * world) if world > 1 else partial
dt = time.perf_counter() - t0
if rank == 0:
    s = sum(k * pow(msm_cases.TAU, i, fr.R) for i, k in enumerate(scalars_all)) % fr.R
    want = bls.g1_serialize(bls.g1_mul((bls.G1_GEN[0], bls.G1_GEN[1], 1), s))
    out["msm_split"] = {"points": total, "ranks": world, "wall_s_incl_h2d_and_gather": dt, "points_per_s": total / dt, "parity": result == want}
    print(out["msm_split"], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/config5_n{world}.json", "w"), indent=1)
if dist is not None:
    dist.destroy_process_group()
