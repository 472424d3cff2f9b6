#!/usr/bin/env python
"""Headline benchmark: Ring VRF proofs/s at ring size 1023 (domain 2^11), batched on B200.

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference --gpus N --steps K ...   # reference algorithm on the host CPU cores

Workload (BASELINE.json configs[1]; seeding of the reference's own tests/benchmark/bench_ring_proof.py:
47-77,140-152): one 1023-key ring, signer at index 3, per step a batch of `--batch` proofs with
alpha = "bench-batch-input" | le64(j), ad = "bench-batch-ad" | le64(j) and 12 blinding rows per proof from
random.Random(0).  Every rank proves its own batch per step (weak scaling, no data-path collective);
`value` = proofs of all ranks / max-over-ranks device time; `e2e` = the same through the public Python API
with host buffers (H2D of the inputs and D2H of the 784-byte proofs inside the timed region).

A number printed by this script under a profiler is not a bench value.
"""

from __future__ import annotations

import argparse
import hashlib
import json
import os
import random
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FR = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
RING_SIZE = 1023
SIGNER_INDEX = 3
MSM_SIZES = (2048, 2048, 2048, 2048, 6145, 6144, 2047)  # per proof at N = 2048
DENSE_MSM_SIZES = (6145, 6144, 2047)  # quotient + two openings: dense coefficient vectors
def sparse_witness_madds(table_bits: int) -> int:
    """Witness columns are committed from their evaluation form: ~131.5 non-zero steps for acc_x / acc_y (one table addition per
    window each), ~135 unit steps for b, 1 for acc_ip, plus 3 blinding rows per column."""
    w = -(-256 // max(table_bits, 1))
    return int(2 * 131.5 * w + 135 + 3 * w + 1 + 4 * w)


def seed_bytes(*parts) -> bytes:
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


def canonical_fq_mul_per_msm(n: int) -> int:
    """Pippenger, signed c-bit windows: min_c ceil(255/c) * (n*10 + 2^c*14) Fq multiplications."""
    return min(-(-255 // c) * (n * 10 + (1 << c) * 14) for c in range(2, 21))


CANONICAL_IMAD_PER_PROOF = sum(canonical_fq_mul_per_msm(n) for n in MSM_SIZES) * 600  # 600 IMAD per 12-limb Montgomery mul


class ClockSampler(threading.Thread):
    """SM clock, power and throttle reasons sampled during the timed region: NVML in-process (no child process next to the
    timed calls); `nvidia-smi` only when the NVML binding is missing."""

    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, device: int):
        super().__init__(daemon=True)
        visible = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v.strip() for v in visible.split(",") if v.strip()]
        self.device = int(ids[device]) if device < len(ids) and ids[device].isdigit() else device
        self.samples: list[dict] = []
        self._halt = threading.Event()

    def _nvml_loop(self) -> bool:
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            return False
        while not self._halt.is_set():
            try:
                mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.samples.append({
                    "sm": float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), "max": sm_max, "power": pynvml.nvmlDeviceGetPowerUsage(h) / 1e3,
                    "reasons": ["Active" if mask & bit else "Not Active" for bit, _ in self.REASONS],
                })  # fmt: skip
            except Exception:
                pass
            self._halt.wait(0.1)
        return True

    def run(self) -> None:
        if self._nvml_loop():
            return
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._halt.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5
                ).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                if len(f) >= 7:
                    self.samples.append({"sm": float(f[0]), "max": float(f[1]), "power": float(f[2]), "reasons": f[3:7]})
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self) -> dict:
        self._halt.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(s["sm"] for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i, r in enumerate(s["reasons"]) if r.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0]["max"], "power_w_max": max(s["power"] for s in self.samples), "reasons": reasons, "samples": len(sm)}


DRY_RUN = os.environ.get("DOT_RING_B200_BENCH_DRYRUN") == "1"  # tests only: CPU emulation build + gloo


def dist_setup(n_gpus: int):
    """One process per GPU; torch.distributed (NCCL) is plumbing for the barrier and the max-over-ranks only."""
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return rank, local, 1, None
    import torch
    import torch.distributed as dist

    if DRY_RUN:
        dist.init_process_group(backend="gloo")
        return rank, local, world, dist
    torch.cuda.set_device(local)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    return rank, local, world, dist


def _device(local):
    import torch

    return torch.device("cpu") if DRY_RUN else torch.device("cuda", local)


def barrier(dist, local):
    if dist is not None:
        import torch

        if DRY_RUN:
            dist.barrier()
            return
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize()


def _reduce(dist, local, values: list[float], op) -> list[float]:
    if dist is None:
        return values
    import torch

    t = torch.tensor(values, dtype=torch.float64, device=_device(local))
    dist.all_reduce(t, op=op)
    return [float(x) for x in t.tolist()]


def reduce_max(dist, local, values: list[float]) -> list[float]:
    return _reduce(dist, local, values, None if dist is None else dist.ReduceOp.MAX)


def reduce_sum(dist, local, values: list[float]) -> list[float]:
    return _reduce(dist, local, values, None if dist is None else dist.ReduceOp.SUM)


# -------------------------------------------------------------------------------------------------- ours
def run_ours(args) -> None:
    rank, local, world, dist = dist_setup(args.gpus)
    os.environ["DOT_RING_B200_DEVICE"] = str(local)
    from dot_ring_b200 import Bandersnatch, Ring, RingProofParams, RingRoot, RingVRF
    from dot_ring_b200 import engine as eng_mod

    t0 = time.perf_counter()
    dry_run = DRY_RUN
    library = None
    if dry_run:
        from tests.host.emul import emulation_library

        library = emulation_library()
    eng = eng_mod.Engine(local, window_bits=args.window_bits, library=library, wide_windows=args.wide_windows, glv=bool(args.glv))
    if not dry_run and not eng.ctx.library.is_cuda:
        raise SystemExit("bench.py measures the CUDA build only")
    eng_mod.set_default_engine(eng, local)
    # the requested table first; if the device cannot hold it (a smaller part, memory in use) fall back to the next smaller geometry
    requested = (args.window_bits, args.wide_windows, bool(args.glv))
    size = lambda g: (-(-((128 if g[2] else 256) - g[1]) // g[0]) + g[1]) << (g[0] - 1) if g[0] else 0  # ~table entries per SRS point  # noqa: E731
    candidates = [requested] + [g for g in ((14, 4, False), (14, 0, False), (13, 0, False), (12, 0, False)) if size(g) < size(requested)]
    for idx, (c, k, glv) in enumerate(candidates):
        eng.window_bits, eng.wide_windows, eng.glv = c, k, glv
        try:
            _ = eng.srs
            # the ring tables (2.6 GB) and a 4096-proof pass (9.5 GB) must still fit next to the table
            left = eng.ctx.device_info()["free_bytes"]
            if not dry_run and left < 14e9 and idx + 1 < len(candidates):
                print(f"[bench] window table ({c}, {k}, glv={glv}) leaves only {left / 1e9:.1f} GB: falling back", file=sys.stderr, flush=True)
                eng._srs.close()
                eng._srs = None
                eng.ctx.trim()
                continue
            break
        except MemoryError as e:
            print(f"[bench] window table ({c}, {k}, glv={glv}) does not fit: {e}", file=sys.stderr, flush=True)
            eng.ctx.trim()
    else:
        raise SystemExit("bench.py: no window table fits on this device")
    eng.ctx.set_commit_mode(args.commit_mode)
    if args.prove_chunk:
        eng.ctx.set_prove_chunk(args.prove_chunk)
    eng.ctx.sync()
    table_s = time.perf_counter() - t0
    info = eng.ctx.device_info()

    # synthetic ring: keys derived with the product's own key derivation (no oracle on this path)
    t0 = time.perf_counter()
    seeds = [seed_bytes("batch-signer", 0, 0) if i == SIGNER_INDEX else seed_bytes("ring-member", 0, i) for i in range(RING_SIZE)]
    from dot_ring_b200.transcript import secret_scalar_from_seed

    sks = [secret_scalar_from_seed(Bandersnatch, s).to_bytes(32, "little") for s in seeds]
    keys = Bandersnatch.public_keys_from_secrets(sks)
    sk, pk = sks[SIGNER_INDEX], keys[SIGNER_INDEX]
    params = RingProofParams.from_ring_size(RING_SIZE)
    ring = Ring(keys, params, eng)
    root = RingRoot.from_ring(ring, params)
    root_bytes = root.encode()
    ring_s = time.perf_counter() - t0

    batch = args.batch
    rng = random.Random(rank)

    def inputs(step: int):
        base = (rank * 1_000_000 + step) * batch
        return (
            [b"bench-batch-input" + le64(base + j) for j in range(batch)],
            [b"bench-batch-ad" + le64(base + j) for j in range(batch)],
            b"".join(rng.randrange(FR).to_bytes(32, "little") for _ in range(12 * batch)),  # blinding rows as wire bytes
        )

    launches0 = eng.ctx.library.launch_count()
    for w in range(args.warmup):
        a, d, zk = inputs(500_000 + w)
        RingVRF[Bandersnatch].prove_batch(a, d, sk, pk, ring, None, zk_rows=zk, as_bytes=True)
    launches_warm = eng.ctx.library.launch_count()

    prepared = [inputs(s) for s in range(args.steps)]
    sampler = ClockSampler(local)
    sampler.start()
    barrier(dist, local)
    eng.ctx.sync()
    dev_ms, phases = 0.0, [0.0] * 6
    t_start = time.perf_counter()
    last = None
    for a, d, zk in prepared:
        last = RingVRF[Bandersnatch].prove_batch(a, d, sk, pk, ring, None, zk_rows=zk, as_bytes=True)
        ph = ring.native.prove_phase_ms()
        dev_ms += sum(ph)
        phases = [x + y for x, y in zip(phases, ph)]
    eng.ctx.sync()
    barrier(dist, local)
    wall_s = time.perf_counter() - t_start
    clocks = sampler.stop()
    launches = eng.ctx.library.launch_count() - launches_warm

    # integer-pipe ceiling measured live on this GPU (dependent-free mad.lo.u32, all SMs)
    if dry_run:
        imad_peak = imad_wide_peak = 148 * 64 * 1.965e9  # nominal; a dry run is not a measurement
    else:
        imad_peak, _ = eng.ctx.microbench("imad", 20000)
        imad_wide_peak, _ = eng.ctx.microbench("imad_wide", 20000)

    mx = reduce_max(dist, local, [dev_ms, wall_s])
    proofs_total = batch * args.steps * world
    value = proofs_total / (mx[0] * 1e-3)
    e2e = proofs_total / mx[1]
    commit_ms = phases[2]
    achieved = CANONICAL_IMAD_PER_PROOF * batch * args.steps / (commit_ms * 1e-3)
    window_bits, wide_windows, glv, windows = eng.srs.geometry
    madds_per_proof = sum(DENSE_MSM_SIZES) * windows + sparse_witness_madds(ring.native.witness_table_bits() or 10)
    executed = madds_per_proof * 10 * 600 * batch * args.steps / (commit_ms * 1e-3)  # 8M + 2S per mixed addition, 600 IMAD per Fq mul
    table_traffic = madds_per_proof * 96 * batch * args.steps  # algorithmic table bytes read

    if dist is not None:
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": "ring_vrf_proofs_per_s",
        "value": value,
        "unit": "proofs/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": mx[0] / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u32 limbs (381-bit Fq / 255-bit Fr Montgomery)",
        "data": "synthetic",
        "config": {
            "workload": f"Ring VRF prove, Bandersnatch, ring {RING_SIZE} / domain 2048, batch {batch} proofs per GPU per step (BASELINE configs[1])",
            "batch_per_gpu": batch,
            "window_bits": window_bits,
            "wide_windows": wide_windows,
            "glv_split": bool(glv),
            "table_additions_per_coefficient": windows,
            "table_gb": round(eng.srs.table_bytes / 1e9, 2),
            "l2": "per-step working set (window table + 2.3 MB of scratch per proof) is far larger than the 126 MB L2; no flush needed",
            "parity": "ring root sha256 " + hashlib.sha256(root_bytes).hexdigest()[:16],
        },
        "e2e": {"value": e2e, "unit": "proofs/s", "h2d_bytes_per_step": batch * (12 * 32 + 32 + 4 * 5 + 25 + 22), "d2h_bytes_per_step": batch * (784 + 4), "ms_per_step": mx[1] * 1e3 / args.steps},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {
            "kernel": "kernel_entry_lb<CommitBodyT> (fixed-base KZG commit) + WitnessCommitBody (sparse witness commitments)",
            "bound": "imad (int32 multiply-add pipe; neither hbm nor tensor: the path is 381-bit modular arithmetic)",
            "achieved": achieved / 1e12,
            "peak": imad_peak / 1e12,
            "unit": "T IMAD/s (canonical Pippenger count, SURVEY.md 8d: 4.674 G IMAD per proof)",
            "frac": achieved / imad_peak,
            "executed": executed / 1e12,
            "executed_frac": executed / imad_peak,
            "executed_note": f"{madds_per_proof} mixed G1 additions per proof actually issued (fixed-base tables + sparse witness columns) x 6000 IMAD; 'achieved' counts the canonical Pippenger work of all 7 MSMs",
            "peak_source": "measured live: dependent-free mad.lo.u32 on all SMs (dr_microbench); IMAD.WIDE measured " + f"{imad_wide_peak / 1e12:.2f} T/s",
            # ncu dram__bytes_read + write of the largest commit launch (1024 x 6145 coefficients, GLV table, 16 additions per
            # coefficient; profiles/r01_ncu_full_CommitBody_glv16.csv) against 9.7 GB of table entries it must touch: a 96-byte entry at
            # a random address straddles 64-byte DRAM atoms (2 or 3 of them), and the kernel is bound by the integer pipe, not by these
            # reads.  (106 GB table: 23.0e9 for 10.9e9 algorithmic, r01_ncu_full_CommitBody_w18.csv.)
            "traffic": 22.0e9 if glv else 23.0e9,
            "traffic_launch": ("CommitBodyT<true> grid (2, 1024) x 128 threads, 40.3 ms under ncu; algorithmic table bytes of that launch 9.7e9" if glv else
                               "CommitBody grid (2, 1024) x 128 threads, 44.6 ms under ncu; algorithmic table bytes of that launch 10.9e9"),
            "algorithmic_table_bytes": table_traffic,
            "hbm_gbs_for_table_reads": table_traffic / (commit_ms * 1e-3) / 1e9,
            "kernel_share_of_step": commit_ms / sum(phases),
        },
        "phase_ms_per_step": {k: v / args.steps for k, v in zip(["pedersen+witness", "interpolate", "commit(msm)", "lde+constraints+quotient", "evals+openings", "transcripts+assembly"], phases)},
        "setup": {"srs_table_s": table_s, "ring_s": ring_s, "device": info["name"], "sm_count": info["sm_count"]},
    }
    if dry_run:
        line["dry_run"] = "CPU emulation of the kernels (tests only); not a measurement"
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_single()
    emit(line)
    if last:
        sys.stderr.write(f"[bench] last proof sha256 {hashlib.sha256(last[-1]).hexdigest()[:16]}\n")


# ------------------------------------------------------------------------------------ CPU baseline (oracle)
def _oracle_ring():
    from oracle import bandersnatch as bs
    from oracle import ring_proof as rp
    from oracle import transcript as tr

    pk, sk = tr.secret_from_seed(bs.SHA512, seed_bytes("batch-signer", 0, 0))
    keys = [pk if i == SIGNER_INDEX else tr.secret_from_seed(bs.SHA512, seed_bytes("ring-member", 0, i))[0] for i in range(RING_SIZE)]
    params = rp.Params.from_ring_size(RING_SIZE)
    ring = rp.Ring(keys, params)
    root = rp.RingRoot.from_ring(ring, params)
    return pk, sk, ring, root


def _oracle_prove_n(args):
    """Worker: build the ring once, then time `count` proofs."""
    count, start = args
    from oracle import fr as ofr
    from oracle import vrf as ovrf

    pk, sk, ring, root = _oracle_ring()
    rng = random.Random(start)
    t0 = time.perf_counter()
    for j in range(count):
        zk = [rng.randrange(ofr.R) for _ in range(12)]
        ovrf.ring_prove(b"bench-batch-input" + le64(start + j), b"bench-batch-ad" + le64(start + j), sk, pk, ring, root, zk_rows=zk)
    return time.perf_counter() - t0


def cpu_baseline_single() -> dict:
    from oracle import backend_name

    count = int(os.environ.get("DOT_RING_B200_CPU_BASELINE_PROOFS", "10"))  # about 11 s of CPU work
    dt = _oracle_prove_n((count, 0))
    return {
        "value": count / dt,
        "unit": "proofs/s",
        "cores": 1,
        "kind": "port",
        "sample": f"{count} proofs, ring {RING_SIZE} / domain 2048, oracle port ({backend_name()}), ring set-up excluded",
    }


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ProcessPoolExecutor

    from oracle import backend_name

    cores = os.cpu_count() or 1
    workers = max(1, min(cores, args.ref_workers or cores))
    total_steps = args.steps + args.warmup
    with ProcessPoolExecutor(max_workers=workers) as ex:
        times = list(ex.map(_oracle_prove_n, [(total_steps, 1000 * w) for w in range(workers)]))
    # every worker proves one proof per step; warm-up steps are part of each worker's loop, so scale them out
    per_worker_s = [t * args.steps / total_steps for t in times]
    wall = max(per_worker_s)
    value = workers * args.steps / wall
    line = {
        "impl": "reference",
        "metric": "ring_vrf_proofs_per_s",
        "value": value,
        "unit": "proofs/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": wall * 1e3 / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "python ints / u64 limbs",
        "data": "synthetic",
        "config": {"workload": f"Ring VRF prove, Bandersnatch, ring {RING_SIZE} / domain 2048; each step = 1 proof on each of {workers} host processes"},
        "cpu_baseline": {"value": value, "unit": "proofs/s", "cores": workers, "kind": "port", "sample": f"{workers} x {args.steps} proofs, oracle port ({backend_name()})"},
        "e2e": {"value": value, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the process's original stdout; everything else (NCCL banners, library chatter) was
    re-routed to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main() -> None:
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # fd 1 -> stderr for native libraries (NCCL prints its version banner on stdout)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8192, help="proofs per GPU per step (one device pass: 19 GB of scratch at ring 1023)")
    ap.add_argument("--window-bits", type=int, default=int(os.environ.get("DOT_RING_B200_WINDOW_BITS", "16")),
                    help="fixed-base table window width; 0 = sized by the library. Default: 16-bit windows over GLV halves (16 additions per coefficient, 161 GB); "
                    "if that does not fit the run falls back to 14-bit windows with four 15-bit ones (18 additions, 106 GB)")
    ap.add_argument("--wide-windows", type=int, default=None, help="how many low windows take one more bit (default 4 with 14-bit windows)")
    ap.add_argument("--glv", type=int, default=None, help="1: table over 128 bits, scalars split with the G1 endomorphism (2 x windows additions per coefficient)")
    ap.add_argument("--prove-chunk", type=int, default=int(os.environ.get("DOT_RING_B200_PROVE_CHUNK", "0")), help="proofs per device pass (0: the library's choice, at most 4096)")
    ap.add_argument("--commit-mode", type=int, default=int(os.environ.get("DOT_RING_B200_COMMIT_MODE", "0")), help="0 XYZZ accumulation, 1 batched-affine rounds")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-workers", type=int, default=0)
    args = ap.parse_args()
    if args.glv is None:
        args.glv = int(os.environ.get("DOT_RING_B200_GLV", "1" if args.window_bits == 16 else "0"))
    if args.wide_windows is None:
        args.wide_windows = int(os.environ.get("DOT_RING_B200_WIDE_WINDOWS", "4" if args.window_bits == 14 and not args.glv else "0"))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
