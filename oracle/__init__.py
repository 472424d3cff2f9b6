"""CPU oracle for the dot-ring ring-proof hot path.

TEST INFRASTRUCTURE ONLY.  This package is a plain-Python restatement of the
algorithms the reference (Chainscore/dot-ring, mounted read-only at
/root/reference while developing) runs on the CPU for the path named in
BASELINE.json.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the
product package ``dot_ring_b200`` never does and fails loudly when its CUDA
library is missing.

Third-party code the reference depends on but which is absent from
/root/reference is restated from its published algorithm:

* blst (https://github.com/Chainscore/blst, branch
  fix/python-as-memory-refcount, un-pinned; reference setup.py:24-25,102-160):
  BLS12-381 G1/G2 arithmetic, Pippenger MSM, zcash point codecs, ate pairing.
* py_ecc 8.0.0 (uv.lock:554-555): curve constants, G1/G2 compression.
* gmpy2 2.2.1 (uv.lock:408-409): big-int helpers only.

Parity pinning: ``tests/test_oracle_golden.py`` checks this oracle against the
reference's own golden vectors (tests/vectors/ark-vrf/bandersnatch_*_ring,
_pedersen, _tiny; tests/vectors/others/ring_proof_*.json) copied as fixtures
under ``tests/golden/reference_vectors/`` and against outputs of the unmodified
reference run in the build container (``tests/golden/generate_golden.py``).
"""


def backend_name() -> str:
    """Which MSM the oracle is using: the plain-C restatement (oracle/c) or pure Python."""
    from . import bls12_381

    return "C Pippenger MSM + Python" if bls12_381._load_c_msm() else "pure Python"
