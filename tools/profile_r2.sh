#!/usr/bin/env bash
# Round-2 profiling pass on the GPU box (one GPU).  usage: tools/profile_r2.sh ncu | sanitizers
#   ncu        : the launch list of the bench command, the DRAM traffic of every Merkle launch of one proof, --set full captures
#                of the top kernels.  Every .ncu-rep is turned into its raw-page CSV on the box and deleted (gpurun copies at
#                most 64 MiB back); tools/summarize_profiles.py turns the CSVs into the tracked summaries under profiles/.
#   sanitizers : compute-sanitizer memcheck / racecheck over kernel-level parity tests and one whole small proof.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
mode="${1:-ncu}"
full() {  # full <name> <ncu args...> -- <command...>
  local name="$1"; shift
  ncu --set full --clock-control none --import-source on -f -o "gpurun_out/$name" "$@" > /dev/null 2>&1
  ncu -i "gpurun_out/$name.ncu-rep" --page raw --csv > "gpurun_out/$name.raw.csv" 2> /dev/null
  rm -f "gpurun_out/$name.ncu-rep"
}
if [ "$mode" = "sanitizers" ]; then
  compute-sanitizer --tool memcheck --error-exitcode 1 --log-file gpurun_out/r2_sanitizer_memcheck.log \
    python -m pytest tests/test_backend_gpu.py tests/test_repeated_gpu.py -m gpu -q -x \
    -k "not 2p24 and not fib19 and not random and not 19 and not 20 and not 21" > gpurun_out/r2_sanitizer_memcheck.pytest 2>&1
  echo "memcheck pytest rc=$?" >> gpurun_out/r2_sanitizer_memcheck.pytest
  compute-sanitizer --tool memcheck --error-exitcode 1 --log-file gpurun_out/r2_sanitizer_memcheck_proof.log \
    python tools/prove_once.py hello_kakarot 1 17 > gpurun_out/r2_sanitizer_memcheck_proof.out 2>&1
  echo "memcheck proof rc=$?" >> gpurun_out/r2_sanitizer_memcheck_proof.out
  compute-sanitizer --tool racecheck --error-exitcode 1 --log-file gpurun_out/r2_sanitizer_racecheck_proof.log \
    python tools/prove_once.py hello_kakarot 1 17 > gpurun_out/r2_sanitizer_racecheck_proof.out 2>&1
  echo "racecheck proof rc=$?" >> gpurun_out/r2_sanitizer_racecheck_proof.out
  compute-sanitizer --tool racecheck --error-exitcode 1 --log-file gpurun_out/r2_sanitizer_racecheck_tests.log \
    python -m pytest tests/test_backend_gpu.py -m gpu -q -x -k "(interpolate or prefix or merkle or is_first) and not 19 and not 20 and not 21" \
    > gpurun_out/r2_sanitizer_racecheck_tests.pytest 2>&1
  echo "racecheck pytest rc=$?" >> gpurun_out/r2_sanitizer_racecheck_tests.pytest
  exit 0
fi
# ---- ncu: launch list of the bench command (per-launch times are serialised and cold-cache: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/r2_launches_bench_prove.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_launches_bench.out 2>&1
# ---- ncu: DRAM traffic of every Merkle launch of the second (warm) proof
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:commit_ -c 400 --csv \
  --log-file gpurun_out/r2_merkle_traffic.csv python tools/prove_once.py fib19 2 > gpurun_out/r2_merkle_traffic.out 2>&1
# ---- ncu --set full on the top kernels (one proof; a few launches of each)
full r2_ncu_full_merkle -k regex:commit_layer_kernel --launch-skip 60 -c 6 python tools/prove_once.py fib19 1
full r2_ncu_full_subtree -k regex:commit_subtree_kernel -c 3 python tools/prove_once.py fib19 1
full r2_ncu_full_other -k 'regex:quotients_kernel|constraint_kernel|fri_tail_kernel|is_first_lde' -c 16 python tools/prove_once.py fib19 1
# the three passes of a 2^25 interpolate and of a 2^25 -> 2^26 extension, four columns, on their own (second iteration)
full r2_ncu_full_fft25 -k regex:fft_kernel --launch-skip 6 -c 6 python tools/fft_one.py 25
ls -la gpurun_out | tail -20
