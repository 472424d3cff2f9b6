#!/usr/bin/env python
"""Batched Fr NTT for profiling: python tools/ntt_one.py [n] [batch]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dot_ring_b200 import _native
from dot_ring_b200.params import ROOT_OF_UNITY_2048, _extend_root_to_size, _omega_for_domain
from oracle import fr
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
ctx = _native.Context(0)
root, size = _extend_root_to_size(ROOT_OF_UNITY_2048, 2048, max(n, 2048), fr.R)
ms, first = ctx.fr_ntt_bench(n, batch, 3, _omega_for_domain(n, fr.R, root, size))
print("n=%d batch=%d %.3f ms %.1f GB/s (64 B/element) first=%x" % (n, batch, ms, n * batch * 64 / ms / 1e6, first & 0xffffffff))
