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
                 srs: "SrsBytes | None" = None, wide_windows: int = 0, glv: bool = False):
        self.device = device
        self.ctx = _native.Context(device, library)
        env = os.environ.get("DOT_RING_B200_WINDOW_BITS")
        # None / 0 = let the library size the table from the free HBM (dr_srs_load); wide_windows only with an explicit width
        self.window_bits = int(window_bits) if window_bits is not None else (int(env) if env else 0)
        self.wide_windows = int(wide_windows) if self.window_bits else 0
        self.glv = bool(glv) if self.window_bits else False
        self._srs: _native.NativeSrs | None = None
        self._srs_points = srs_points
        # `srs` overrides the bundled 6145-point file (needed for domains above 2048, whose quotient has 3N + 1 coefficients)
        self.srs_bytes = srs if srs is not None else read_srs_file(None, srs_points)

    @property
    def srs(self) -> _native.NativeSrs:
        """SRS points + window table in HBM, built on first use (about 3.5 s for the 106 GB table a B200 gets by default)."""
        if self._srs is None:
            self._srs = _native.NativeSrs(self.ctx, self.srs_bytes.g1_be96, self.srs_bytes.g2_be192, self.window_bits, self.wide_windows, self.glv)
            self.window_bits, self.wide_windows, glv, self.table_additions = self._srs.geometry
            self.glv = bool(glv)
        return self._srs

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
