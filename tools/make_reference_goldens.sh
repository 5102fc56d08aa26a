#!/usr/bin/env bash
# Builds the UNMODIFIED reference (kkrt-labs/stwo-brainfuck, Rust, stwo-prover 0.1.1 @ 31e8dbc) on a machine that has cargo and
# network access, and turns its outputs into golden files for this repo — the step that closes "parity unpinned" (SURVEY.md
# §8c, §7 "Hard parts").  Neither the build container nor the GPU box of this project has a Rust toolchain
# (profiles/r2_gpu_box_probe.txt), so this script has never run there; it is the recipe a maintainer runs once.
#
#   tools/make_reference_goldens.sh <path-to-stwo-brainfuck-checkout> [<path-to-this-repo>]
#
# Produces, under <repo>/tests/golden/ref/:
#   <program>.proof.json    serde_json of the reference's BrainfuckProof  (bin/brainfuck_prover.rs:127-131)
#   <program>.time.txt      "Proof generation time" as the reference logs it (:138), RAYON threads stated
#   toolchain.txt           rustc/cargo versions, the Stwo revision from Cargo.lock, `nproc`
# and, when <repo>/gpurun_out/cuda_proofs/<program>.proof.json exist (written by `python tools/dump_cuda_proofs.py` on a GPU
# box), runs the reference's own `verify` (bin/brainfuck_prover.rs:145-152 -> verify_brainfuck, brainfuck_air/mod.rs:738-797)
# on every CUDA proof and records accept / reject in cuda_proofs_verified_by_reference.txt.
#
# tests/test_reference_goldens.py then compares (skipped while tests/golden/ref/ is empty):
#   * the CUDA prover's JSON with <program>.proof.json byte for byte;
#   * every commitment, claimed sum and FRI layer root individually, so that a mismatch names the first stage that differs.
set -euo pipefail
REF=${1:?usage: make_reference_goldens.sh <stwo-brainfuck checkout> [<this repo>]}
REPO=${2:-$(cd "$(dirname "$0")/.." && pwd)}
OUT="$REPO/tests/golden/ref"
PROGS="$REPO/tests/golden/programs"
mkdir -p "$OUT"

cd "$REF"
{
  echo "date: $(date -u +%FT%TZ)"
  rustc --version; cargo --version
  echo "nproc: $(nproc)"
  grep -A2 'name = "stwo-prover"' Cargo.lock
  git -C "$REF" rev-parse HEAD 2>/dev/null | sed 's/^/reference HEAD: /' || true
} > "$OUT/toolchain.txt"

# README.md:32-36 — the parallel feature is the reference's fastest CPU configuration and the one BASELINE.json names
cargo build --package brainfuck_prover --features parallel --release
BIN="$REF/target/release/brainfuck_prover"

prove() {  # name, stdin bytes (printf format)
  local name=$1 input=$2
  printf "$input" | "$BIN" prove --file "$PROGS/$name.bf" --output "$OUT/$name.proof.json" 2>&1 | tee "$OUT/$name.log" \
    | grep -E "Steps|Proof generation time|Execution trace time" > "$OUT/$name.time.txt" || true
  echo "RAYON_NUM_THREADS=${RAYON_NUM_THREADS:-unset} nproc=$(nproc)" >> "$OUT/$name.time.txt"
  "$BIN" verify "$OUT/$name.proof.json"
}
# BASELINE.json configs[0..3]; LOG_MAX_ROWS is the reference's shipped 24 (brainfuck_air/mod.rs:427-428)
prove hello_kakarot ""
prove fib19 ""
prove collatz '7\n'
prove synthetic_2p24 ""

# the reference's verifier on CUDA proofs
CUDA="$REPO/gpurun_out/cuda_proofs"
if [ -d "$CUDA" ]; then
  : > "$OUT/cuda_proofs_verified_by_reference.txt"
  for f in "$CUDA"/*.proof.json; do
    if "$BIN" verify "$f" > /dev/null 2>&1; then echo "$(basename "$f") accepted" >> "$OUT/cuda_proofs_verified_by_reference.txt"
    else echo "$(basename "$f") REJECTED" >> "$OUT/cuda_proofs_verified_by_reference.txt"; fi
  done
  cat "$OUT/cuda_proofs_verified_by_reference.txt"
fi
sha256sum "$OUT"/*.proof.json > "$OUT/SHA256SUMS"
echo "goldens written to $OUT"
