#!/usr/bin/env bash
# Build the reference's own three Cython extensions (unmodified sources, read where they lie
# under /root/reference) into oracle/_ref/ so that the UNMODIFIED reference Python package can
# be imported in the build container by tests/golden/generate_golden.py.
#
# oracle/_ref/ is git-ignored build output; no reference source is copied into the repository.
# The reference's own build system (setup.py) is not run: it clones blst over the network.
# blst / py_ecc / gmpy2 (absent third-party dependencies) are replaced at import time by the
# shims in oracle/ref_shims/, which are backed by the oracle's own BLS12-381 restatement.
set -euo pipefail
REF="${REF:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/dot_ring" ]; then echo "reference not present at $REF; nothing to build"; exit 0; fi
mkdir -p "$OUT/build" "$OUT/ext"
PYINC="$(python -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
SUFFIX="$(python -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
NF="$REF/dot_ring/curve/native_field"
build_one() { # <module dotted name> <pyx path>
  local mod="$1" pyx="$2" base
  base="$(basename "$pyx" .pyx)"
  python -m cython -3 -I "$REF" "$pyx" -o "$OUT/build/$base.c" --module-name "$mod" >/dev/null
  gcc -O3 -fPIC -shared -I"$PYINC" -I"$NF" "$OUT/build/$base.c" "$NF/bls12_381_scalar.c" -o "$OUT/ext/$base$SUFFIX"
}
build_one dot_ring.curve.native_field.scalar "$NF/scalar.pyx"
build_one dot_ring.curve.native_field.bandersnatch_te "$NF/bandersnatch_te.pyx"
build_one dot_ring.ring_proof.polynomial.ntt "$REF/dot_ring/ring_proof/polynomial/ntt.pyx"
ls -la "$OUT/ext"
