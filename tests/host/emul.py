"""Bind the CPU emulation build of the kernels (tests only).

``./build.sh --emul`` compiles the same ``dot_ring_b200/csrc/api_*.cu`` sources with g++ and
-DDR_HOST_EMULATION: kernel bodies run block by block on the CPU (see csrc/rt.cuh).  The product
package never loads this library; tests construct it explicitly to check kernel logic on CPU.
"""

from __future__ import annotations

import subprocess
from functools import lru_cache
from pathlib import Path

from dot_ring_b200 import _native

ROOT = Path(__file__).resolve().parents[2]
SO = Path(__file__).resolve().parent / "libdotring_emul.so"


@lru_cache(maxsize=1)
def emulation_library() -> _native.Library:
    srcs = list((ROOT / "dot_ring_b200" / "csrc").rglob("*.cu*")) + list((ROOT / "dot_ring_b200" / "csrc" / "gen").glob("*")) + [ROOT / "include" / "dot_ring_b200.h"]
    if not SO.exists() or SO.stat().st_mtime < max(p.stat().st_mtime for p in srcs):
        subprocess.check_call([str(ROOT / "build.sh"), "--emul"])
    return _native.Library(SO, require_cuda=False)
