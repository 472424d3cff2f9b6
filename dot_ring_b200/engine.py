"""Process-wide engine: one CUDA context + SRS (with its fixed-base table) per GPU.

The reference keeps a module-level ``KZG.srs`` (dot_ring/ring_proof/pcs/kzg.py:114-119); the
equivalent here is a lazily created :class:`Engine` per device, holding the ``dr_ctx`` and ``dr_srs``
handles.  Device selection is a property of the engine (``DOT_RING_B200_DEVICE`` or
``Engine(device=...)``), not of ``RingProofParams``.
"""

from __future__ import annotations

import os
import threading

from . import _native
from .srs import SrsBytes, read_srs_file

_lock = threading.Lock()
_engines: dict[int, "Engine"] = {}


class Engine:
    def __init__(self, device: int = 0, window_bits: int | None = None, library: _native.Library | None = None, srs_points: int | None = None,
                 srs: "SrsBytes | None" = None):
        self.device = device
        self.ctx = _native.Context(device, library)
        env = os.environ.get("DOT_RING_B200_WINDOW_BITS")
        self.window_bits = int(window_bits) if window_bits is not None else (int(env) if env else None)
        self._srs: _native.NativeSrs | None = None
        self._srs_points = srs_points
        # `srs` overrides the bundled 6145-point file (needed for domains above 2048, whose quotient has 3N + 1 coefficients)
        self.srs_bytes = srs if srs is not None else read_srs_file(None, srs_points)

    @property
    def srs(self) -> _native.NativeSrs:
        """SRS points + window table in HBM, built on first use (about 0.5 s for the default 26.6 GB table)."""
        if self._srs is None:
            if self.window_bits is None:
                self.window_bits = self._auto_window_bits()
            self._srs = _native.NativeSrs(self.ctx, self.srs_bytes.g1_be96, self.srs_bytes.g2_be192, self.window_bits)
        return self._srs

    def _auto_window_bits(self) -> int:
        """Largest window whose table (n * ceil(256/c) * 2^(c-1) * 96 B) fits in half of the free HBM: 14 bits = 92 GB for
        the bundled 6145-point SRS on a 180 GB B200 (19 additions per coefficient; 12 bits = 27 GB, 22 additions)."""
        free = self.ctx.device_info()["free_bytes"]
        n = self.srs_bytes.n_g1
        for c in (14, 13, 12, 11, 10, 9, 8):
            if n * -(-256 // c) * (1 << (c - 1)) * 96 <= free // 2:
                return c
        return 8

    def close(self) -> None:
        if self._srs is not None:
            self._srs.close()
            self._srs = None
        self.ctx.close()


def default_engine() -> Engine:
    device = int(os.environ.get("DOT_RING_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    with _lock:
        eng = _engines.get(device)
        if eng is None:
            eng = Engine(device)
            _engines[device] = eng
        return eng


def set_default_engine(engine: Engine | None, device: int = 0) -> None:
    """Install (or drop) the engine used by the module-level API; tests use it to inject the emulation build."""
    with _lock:
        if engine is None:
            _engines.pop(device, None)
        else:
            _engines[device] = engine
