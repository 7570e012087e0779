#!/bin/bash
# compute-sanitizer over small decodes through every kernel (tests/sanitizer_workload.py).
#   tools/sanitize.sh build   (no GPU) also builds the -DTORBI_RING_SYNCTHREADS variant of the library
#                             into tools/_build/libtorbi_b200_ringsync.so: the tiled / streamed-split
#                             copy rings synchronised by cp.async.wait_all + __syncthreads instead of
#                             mbarriers -- the same schedule in a form racecheck can model
#   tools/sanitize.sh run     (GPU box) memcheck + synccheck + racecheck on the shipped library,
#                             racecheck again on the ring-sync variant; summary lines to stdout
set -e
cd "$(dirname "$0")/.."
variant=tools/_build/libtorbi_b200_ringsync.so
if [ "$1" = build ]; then
  mkdir -p tools/_build
  make -s -j8 -C torbi_b200/csrc BUILD=_build/ringsync OUT=../../$variant EXTRA=-DTORBI_RING_SYNCTHREADS >/dev/null
  ls -la "$variant"
  exit 0
fi
summary() {
  tee /tmp/sanitizer_raw.txt > /dev/null
  echo "decodes checked against the oracle: $(grep -c '^ok ' /tmp/sanitizer_raw.txt)"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" /tmp/sanitizer_raw.txt | sort | uniq -c
  grep -E "hazard|Invalid|Uninitialized|Barrier error" /tmp/sanitizer_raw.txt | sed 's/+0x[0-9a-f]*//; s/=========//' | sort | uniq -c | sort -rn | head -${2:-6}
}
for tool in ${TOOLS:-memcheck synccheck racecheck}; do
  echo "== $tool, shipped library"
  compute-sanitizer --tool $tool --print-limit 5 python tests/sanitizer_workload.py 2>&1 | summary
done
echo "== racecheck, ring synchronised by __syncthreads (-DTORBI_RING_SYNCTHREADS)"
TORBI_B200_LIBRARY="$PWD/$variant" compute-sanitizer --tool racecheck --print-limit 5 python tests/sanitizer_workload.py 2>&1 | summary
