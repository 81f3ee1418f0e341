#!/usr/bin/env bash
# compute-sanitizer over the small-shape pass of every kernel family (scripts/sanitize_small.py).
#   scripts/sanitize.sh [outdir]      (run on a GPU box; logs + a summary go to outdir, default gpurun_out/)
# memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards between the warp roles
# of the hand-rolled mbarrier pipelines; synccheck: invalid barrier usage.  The global-memory
# data-as-flag protocols (car2.cu, nystrom.cu) are outside racecheck's model (it tracks shared memory only).
set -u
cd "$(dirname "$0")/.."
OUT="${1:-gpurun_out}"
mkdir -p "$OUT"
SAN=/usr/local/cuda/bin/compute-sanitizer
PER_TOOL_TIMEOUT="${SANITIZE_TIMEOUT:-900}"
: > "$OUT/sanitize_summary.txt"
for tool in memcheck racecheck synccheck; do
  log="$OUT/sanitize_${tool}.log"
  extra=""
  [ "$tool" = racecheck ] && extra="--racecheck-report all"
  timeout "$PER_TOOL_TIMEOUT" "$SAN" --tool "$tool" $extra --print-limit 20 --error-exitcode 3 \
      python scripts/sanitize_small.py ${SANITIZE_WHICH:-} > "$log" 2>&1
  rc=$?
  {
    echo "== $tool: exit code $rc (0 = clean, 3 = errors reported, 124 = timeout after ${PER_TOOL_TIMEOUT}s)"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|\[sanitize\]" "$log" | tail -n 20
  } >> "$OUT/sanitize_summary.txt"
done
cat "$OUT/sanitize_summary.txt"
