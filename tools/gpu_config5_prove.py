#!/usr/bin/env python
"""BASELINE configs[4] end to end: Ring VRF proofs at domain 2^16 (65 279-row ring) on a synthetic 196 609-point SRS.
  part 1  oracle parity: 48 real keys inside the 2^16-row ring, root + 1 proof byte-identical to the CPU oracle (minutes of CPU)
  part 2  a 65 000-key ring: ring creation, a batch of proofs, per-item and aggregated verification (device only)
Writes gpurun_out/config5_prove.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dot_ring_b200 import _native
from tests import extended_domain

ctx = _native.Context(0)
say = lambda *a: print(*a, flush=True)
out = {}
if os.environ.get("SKIP_ORACLE") != "1":
    out["oracle_parity"] = extended_domain.prove_verify_against_oracle(ctx, 65536, 48, 2, 1, window_bits=8, log=say)
out["full_ring"] = extended_domain.prove_verify_against_oracle(ctx, 65536, 65000, int(os.environ.get("N_PROOFS", "64")), 0, window_bits=8, log=say)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/config5_prove.json", "w"), indent=1)
print("config5 prove ok")
