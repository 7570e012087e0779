#!/bin/bash
# Per-phase cycle counts of the step loops (tiled: first-stage wait / K loop / epilogue / prefetch /
# exchange wait; split and cluster kernels likewise).  `build` compiles a variant of the library with
# TORBI_TILED_PROFILE into tools/_build/libtorbi_b200_prof.so (no GPU needed; the file travels with
# gpurun); `run` (on the GPU box) decodes the headline shape once per mode and prints what a few
# CTAs measured.
set -e
cd "$(dirname "$0")/.."
lib=tools/_build/libtorbi_b200_prof.so
if [ "$1" = build ]; then
  mkdir -p tools/_build
  make -s -j8 -C torbi_b200/csrc BUILD=_build/prof OUT=../../$lib EXTRA=-DTORBI_TILED_PROFILE >/dev/null
  ls -la "$lib"
  exit 0
fi
for mode in ${MODES:-"" "--backpointers"} ${EXTRA_MODES}; do
  echo "== mode: ${mode:-score history}"
  TORBI_B200_LIBRARY="$PWD/$lib" python bench.py --steps 1 --warmup 1 --no-e2e \
      --no-cpu-baseline --no-configs --no-strong $mode --batch ${BATCH:-512} --frames ${FRAMES:-500} | grep "^cta" | sort | uniq | head -8
done
