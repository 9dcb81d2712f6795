#!/bin/bash
# Pins the oracle (and the CUDA path) to the REAL pbrt-r3 binary, for whoever has a Rust toolchain: this image has no
# rustc / cargo and the crates are not vendored, so the script cannot run here - it turns "parity unpinned
# at image level" into one command elsewhere.
#
#   tools/run_reference.sh [/path/to/pbrt-r3-checkout | /path/to/pbrt-r3-binary] [out_dir]
#
# 1. builds the reference (cargo build --release; add --features sampled-spectrum for the 60-bin scene) unless a binary
#    is given,
# 2. writes the five synthetic BASELINE scenes (tools/make_reference_scenes.py: the same generators the tests use, at
#    sizes a CPU renders in minutes) as .pbrt files,
# 3. renders each with the reference: RAYON_NUM_THREADS=1 for the Halton scenes (the reference's tile clones share one
#    pixel-offset cache through Arc<Mutex<..>>, halton.rs:52-53,267-269: a data race under rayon; one thread gives the
#    intended per-pixel offsets),
# 4. converts the EXR images to tests/golden/ref_<scene>.npy and compares them with the oracle's render of the same
#    file (relative RMSE <= 1e-3, the north-star bar) - and with the CUDA library when a GPU is visible.
set -e
REF=${1:-/root/reference}
OUT=${2:-tests/golden/_ref_run}
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
mkdir -p "$OUT"
if [ -d "$REF" ]; then
  command -v cargo >/dev/null 2>&1 || { echo "cargo not found: install a Rust toolchain (and let it fetch the crates of $REF/Cargo.toml), or pass a pbrt-r3 binary"; exit 2; }
  (cd "$REF" && cargo build --release)
  BIN_RGB="$REF/target/release/pbrt-r3"
  (cd "$REF" && cargo build --release --features sampled-spectrum --target-dir target-spectral) && BIN_SPEC="$REF/target-spectral/release/pbrt-r3" || BIN_SPEC=""
else
  BIN_RGB="$REF"; BIN_SPEC=""
fi
python "$ROOT/tools/make_reference_scenes.py" "$OUT"
for f in "$OUT"/*.pbrt; do
  name=$(basename "$f" .pbrt)
  bin="$BIN_RGB"
  case "$name" in *spectral*) [ -n "$BIN_SPEC" ] || { echo "skip $name (no sampled-spectrum build)"; continue; }; bin="$BIN_SPEC";; esac
  echo "== $name"
  RAYON_NUM_THREADS=1 "$bin" -i "$f" -o "$OUT/$name.exr" --quiet
done
python "$ROOT/tools/compare_reference.py" "$OUT"
