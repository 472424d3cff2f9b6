"""Import shims that let the UNMODIFIED reference package run in the build container.

TEST INFRASTRUCTURE ONLY (used by tests/golden/generate_golden.py to produce committed golden
vectors).  The reference needs three third-party packages that are not installed and cannot be
fetched (no network): gmpy2, py_ecc and its own SWIG build of blst.  ``install()`` registers
minimal stand-ins for exactly the API subset the reference calls,
backed by ``oracle.bls12_381``, and maps the reference's three Cython extension modules to the
binaries built by ``oracle/build_ref.sh`` into ``oracle/_ref/ext``.
"""

from __future__ import annotations

import importlib.abc
import importlib.machinery
import importlib.util
import sys
import sysconfig
import types
from pathlib import Path

REF_ROOT = Path("/root/reference")
EXT_DIR = Path(__file__).resolve().parent.parent / "_ref" / "ext"
_EXT = {
    "dot_ring.curve.native_field.scalar": "scalar",
    "dot_ring.curve.native_field.bandersnatch_te": "bandersnatch_te",
    "dot_ring.ring_proof.polynomial.ntt": "ntt",
}


class _ExtFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        base = _EXT.get(fullname)
        if base is None:
            return None
        so = EXT_DIR / (base + sysconfig.get_config_var("EXT_SUFFIX"))
        loader = importlib.machinery.ExtensionFileLoader(fullname, str(so))
        return importlib.util.spec_from_file_location(fullname, str(so), loader=loader)


def available() -> bool:
    return (REF_ROOT / "dot_ring").is_dir() and all(
        (EXT_DIR / (b + sysconfig.get_config_var("EXT_SUFFIX"))).exists() for b in _EXT.values()
    )


def install() -> None:
    """Make ``import dot_ring.vrf.ring`` etc. resolve to the reference sources + shims."""
    if "dot_ring" in sys.modules:
        return
    from . import blst_shim, gmpy2_shim, py_ecc_shim

    sys.modules["gmpy2"] = gmpy2_shim
    py_ecc_shim.register(sys.modules)
    # A bare package object: the reference's top-level __init__ imports every curve suite
    # (most need more of py_ecc); the hot path only needs the sub-packages.
    pkg = types.ModuleType("dot_ring")
    pkg.__path__ = [str(REF_ROOT / "dot_ring")]
    sys.modules["dot_ring"] = pkg
    sys.modules["dot_ring.blst"] = blst_shim
    pkg.blst = blst_shim
    sys.meta_path.insert(0, _ExtFinder())
