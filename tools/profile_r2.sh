#!/usr/bin/env bash
# Round-2 profiling pass on the GPU box (one GPU): sanitizers, the ncu launch list of the bench command, full-set captures of
# the top kernels and the DRAM traffic of every Merkle launch of one proof.  Everything lands in gpurun_out/; the summaries
# that are judged are copied to profiles/ by tools/summarize_profiles.py.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
# ---- compute-sanitizer: memcheck over the kernel-level parity tests + one whole small proof; racecheck over the smem kernels
compute-sanitizer --tool memcheck --error-exitcode 1 --log-file gpurun_out/r2_sanitizer_memcheck.log \
  python -m pytest tests/test_backend_gpu.py tests/test_repeated_gpu.py tests/test_device_tables.py -m gpu -q -x \
  -k "not 2p24 and not fib19 and not random" > gpurun_out/r2_sanitizer_memcheck.pytest 2>&1
echo "memcheck pytest rc=$?" >> gpurun_out/r2_sanitizer_memcheck.pytest
compute-sanitizer --tool memcheck --error-exitcode 1 --log-file gpurun_out/r2_sanitizer_memcheck_proof.log \
  python tools/prove_once.py hello_kakarot 1 17 > gpurun_out/r2_sanitizer_memcheck_proof.out 2>&1
echo "memcheck proof rc=$?" >> gpurun_out/r2_sanitizer_memcheck_proof.out
compute-sanitizer --tool racecheck --error-exitcode 1 --log-file gpurun_out/r2_sanitizer_racecheck_proof.log \
  python tools/prove_once.py hello_kakarot 1 14 > gpurun_out/r2_sanitizer_racecheck_proof.out 2>&1
echo "racecheck proof rc=$?" >> gpurun_out/r2_sanitizer_racecheck_proof.out
compute-sanitizer --tool racecheck --error-exitcode 1 --log-file gpurun_out/r2_sanitizer_racecheck_tests.log \
  python -m pytest tests/test_backend_gpu.py tests/test_device_tables.py -m gpu -q -x -k "interpolate or prefix or hello_kakarot or merkle" \
  > gpurun_out/r2_sanitizer_racecheck_tests.pytest 2>&1
echo "racecheck pytest rc=$?" >> gpurun_out/r2_sanitizer_racecheck_tests.pytest
# ---- ncu: launch list of the bench command (per-launch times are serialised and cold-cache: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_bench_prove.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_launches_bench.out 2>&1
# ---- ncu: DRAM traffic of every Merkle launch of the second (warm) proof
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:commit_ -c 400 --csv \
  --log-file gpurun_out/r2_merkle_traffic.csv python tools/prove_once.py fib19 2 > gpurun_out/r2_merkle_traffic.out 2>&1
# ---- ncu --set full on the top kernels (one proof; a few launches of each)
ncu --set full --clock-control none --import-source on -k regex:commit_layer_kernel --launch-skip 60 -c 6 -f -o gpurun_out/r2_ncu_full_merkle \
  python tools/prove_once.py fib19 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:commit_subtree_kernel -c 3 -f -o gpurun_out/r2_ncu_full_subtree \
  python tools/prove_once.py fib19 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_kernel --launch-skip 20 -c 6 -f -o gpurun_out/r2_ncu_full_fft \
  python tools/prove_once.py fib19 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"quotients_kernel|constraint_kernel|fri_tail_kernel|rs_scatter" -c 8 -f \
  -o gpurun_out/r2_ncu_full_other python tools/prove_once.py fib19 1 > /dev/null 2>&1
ls -la gpurun_out | tail -20
