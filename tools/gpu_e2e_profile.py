#!/usr/bin/env python
"""Where the host-side time of one 4096-proof RingVRF.prove_batch call goes (cProfile, after two warm-up calls)."""
import cProfile
import os
import pstats
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dot_ring_b200 import Bandersnatch, Ring, RingProofParams, RingVRF  # noqa: E402
from dot_ring_b200 import engine as eng_mod  # noqa: E402
from dot_ring_b200.transcript import secret_scalar_from_seed  # noqa: E402

eng = eng_mod.Engine(0, window_bits=int(os.environ.get("WB", "14")), wide_windows=int(os.environ.get("WW", "4")), glv=bool(int(os.environ.get("GLV", "0"))))
eng_mod.set_default_engine(eng, 0)
seeds = [bench.seed_bytes("batch-signer", 0, 0) if i == bench.SIGNER_INDEX else bench.seed_bytes("ring-member", 0, i) for i in range(bench.RING_SIZE)]
sks = [secret_scalar_from_seed(Bandersnatch, s).to_bytes(32, "little") for s in seeds]
keys = Bandersnatch.public_keys_from_secrets(sks)
sk, pk = sks[bench.SIGNER_INDEX], keys[bench.SIGNER_INDEX]
ring = Ring(keys, RingProofParams.from_ring_size(bench.RING_SIZE), eng)
n = 4096
rng = random.Random(0)


def inputs(step):
    return ([b"bench-batch-input" + bench.le64(step * n + j) for j in range(n)], [b"bench-batch-ad" + bench.le64(step * n + j) for j in range(n)],
            b"".join(rng.randrange(bench.FR).to_bytes(32, "little") for _ in range(12 * n)))  # fmt: skip


for w in range(2):
    a, d, zk = inputs(w)
    RingVRF[Bandersnatch].prove_batch(a, d, sk, pk, ring, None, zk_rows=zk, as_bytes=True)
for rep in range(3):
    a, d, zk = inputs(10 + rep)
    eng.ctx.sync()
    t0 = time.perf_counter()
    prof = cProfile.Profile()
    prof.enable()
    RingVRF[Bandersnatch].prove_batch(a, d, sk, pk, ring, None, zk_rows=zk, as_bytes=True)
    prof.disable()
    wall = time.perf_counter() - t0
    dev = sum(ring.native.prove_phase_ms())
    print(f"rep {rep}: wall {wall * 1e3:.1f} ms, device phases {dev:.1f} ms, gap {wall * 1e3 - dev:.1f} ms")
pstats.Stats(prof).sort_stats("tottime").print_stats(12)
