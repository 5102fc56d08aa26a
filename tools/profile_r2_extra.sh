set -u
cd "$(dirname "$0")/.."
ncu --set full --clock-control none --import-source on -f -o gpurun_out/r2_ncu_full_merkle -k regex:commit_layer_kernel --launch-skip 22 -c 6 python tools/prove_once.py fib19 1 > /dev/null 2>&1
ncu -i gpurun_out/r2_ncu_full_merkle.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_merkle.raw.csv 2>/dev/null
rm -f gpurun_out/r2_ncu_full_merkle.ncu-rep
compute-sanitizer --tool memcheck --error-exitcode 1 --log-file gpurun_out/r2_sanitizer_memcheck_proof.log python tools/prove_once.py hello_kakarot 1 17 > gpurun_out/r2_sanitizer_memcheck_proof.out 2>&1
echo "memcheck proof rc=$?" >> gpurun_out/r2_sanitizer_memcheck_proof.out
compute-sanitizer --tool racecheck --error-exitcode 1 --log-file gpurun_out/r2_sanitizer_racecheck_proof.log python tools/prove_once.py hello_kakarot 1 17 > gpurun_out/r2_sanitizer_racecheck_proof.out 2>&1
echo "racecheck proof rc=$?" >> gpurun_out/r2_sanitizer_racecheck_proof.out
compute-sanitizer --tool racecheck --error-exitcode 1 --log-file gpurun_out/r2_sanitizer_racecheck_tests.log python -m pytest tests/test_backend_gpu.py -m gpu -q -x -k "(interpolate or evaluate or prefix or merkle or is_first) and not 19 and not 20 and not 21 and not 22 and not 23" > gpurun_out/r2_sanitizer_racecheck_tests.pytest 2>&1
echo "racecheck pytest rc=$?" >> gpurun_out/r2_sanitizer_racecheck_tests.pytest
compute-sanitizer --tool memcheck --error-exitcode 1 --log-file gpurun_out/r2_sanitizer_memcheck.log python -m pytest tests/test_backend_gpu.py -m gpu -q -x -k "(interpolate or evaluate or quotient or is_first or fold or merkle) and not 19 and not 20 and not 21 and not 22 and not 23" > gpurun_out/r2_sanitizer_memcheck.pytest 2>&1
echo "memcheck pytest rc=$?" >> gpurun_out/r2_sanitizer_memcheck.pytest
tail -n 3 gpurun_out/r2_sanitizer_*.pytest gpurun_out/r2_sanitizer_*proof.out; cat gpurun_out/r2_sanitizer_*.log | tail -12; ls -la gpurun_out
