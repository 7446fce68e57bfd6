#!/usr/bin/env bash
# ThreadSanitizer over the multi-threaded host code (fixture builder: parallel heightmap, tunnel bitmap, per-subcell
# sub-DAGs merged in order; flatten; bulk and looped edits).  Pure C++ driver -- no Python, no CUDA: the three upload
# entry points the host files call are stubbed.  Usage: tools/tsan_host.sh [depth] [log2cap]   (9 21 by default, ~10 s;
# 11 23 takes ~8 min under TSan).  Exit code 0 and no "WARNING: ThreadSanitizer" = no data race seen.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="${TMPDIR:-/tmp}/ort_tsan_driver"
g++ -std=c++17 -O1 -g -fsanitize=thread -fPIE -pie -ffp-contract=off -I "$ROOT/include" -I "$ROOT/octree_ray_tracing_b200/csrc" \
    "$ROOT/tools/tsan_host_driver.cpp" "$ROOT/octree_ray_tracing_b200/csrc/ort_host_tree.cpp" "$ROOT/octree_ray_tracing_b200/csrc/ort_fixture.cpp" \
    -lpthread -o "$OUT"
TSAN_OPTIONS=halt_on_error=0 "$OUT" "${1:-9}" "${2:-21}" 2>&1 | tee "$OUT.log"
! grep -q "WARNING: ThreadSanitizer" "$OUT.log"
echo "host threads: no data race reported"
