#!/bin/bash
# Per-phase cycle counts of the tiled kernel's step loop (first-stage wait / K loop / epilogue /
# prefetch / exchange wait).  `build` compiles a throw-away copy of the library with
# TORBI_TILED_PROFILE into torbi_b200/csrc/_build/prof (no GPU needed; the directory travels with
# gpurun); `run` (on the GPU box) decodes the headline shape once per mode and prints what a few
# CTAs measured.
set -e
cd "$(dirname "$0")/.."
prof=torbi_b200/csrc/_build/prof
if [ "$1" = build ]; then
  rm -rf "$prof"; mkdir -p "$prof/torbi_b200" 
  mkdir -p "$prof/torbi_b200/csrc"; cp torbi_b200/csrc/*.cu torbi_b200/csrc/*.cpp torbi_b200/csrc/*.cuh torbi_b200/csrc/Makefile "$prof/torbi_b200/csrc/"
  cp -r include "$prof"/
  make -s -j8 -C "$prof/torbi_b200/csrc" NVCC="/usr/local/cuda/bin/nvcc -DTORBI_TILED_PROFILE" >/dev/null
  ls -la "$prof/torbi_b200/libtorbi_b200.so"
  exit 0
fi
for mode in ${MODES:-"" "--backpointers"} ${EXTRA_MODES}; do
  echo "== mode: ${mode:-score history}"
  TORBI_B200_LIBRARY="$prof/torbi_b200/libtorbi_b200.so" python bench.py --steps 1 --warmup 1 --no-e2e \
      --no-cpu-baseline $mode --batch ${BATCH:-512} --frames ${FRAMES:-500} | grep "^cta" | sort | uniq | head -8
done
