#!/usr/bin/env python
"""One variable-base MSM of 2^K points (K from argv) for profiling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dot_ring_b200 import _native
from tests import msm_cases
k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = _native.Context(0)
ms, c, out = ctx.g1_msm_bench(1 << k, 2, 7, 0, msm_cases.TAU)
print("n=2^%d c=%d %.3f ms %.1f Mpts/s" % (k, c, ms, (1 << k) / ms / 1e3))
