#!/usr/bin/env bash
# Memory / UB check of the kernel logic on the CPU: the emulation build (same .cu sources, g++ -DDR_HOST_EMULATION) compiled with
# AddressSanitizer + UBSan, then tools/sanitize_small.py (SRS table, ring 8 prove + verify, VRF verify, MSM, NTT) through it.
# compute-sanitizer is closed on the GPU pool, so this is the out-of-bounds check for indexing shared by both builds.
set -euo pipefail
cd "$(dirname "$0")/.."
OUT=${OUT:-/tmp/dotring_asan}
mkdir -p "$OUT"
for f in dot_ring_b200/csrc/api_*.cu; do
  g++ -O1 -g -std=c++17 -fPIC -fsanitize=address,undefined -fno-omit-frame-pointer -DDR_HOST_EMULATION -x c++ -c "$f" -o "$OUT/$(basename "$f" .cu).o" &
done
wait
g++ -shared -fsanitize=address,undefined -o "$OUT/libdotring_asan.so" "$OUT"/*.o -lpthread
ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" DR_EMUL_LIB="$OUT/libdotring_asan.so" python tools/sanitize_small.py
