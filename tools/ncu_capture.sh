#!/bin/bash
# ncu --set full captures of a few launches of each hot kernel class inside ONE real fib19 proof (run on the GPU box; the
# reports land in gpurun_out/ and are summarised into profiles/ with tools/ncu_summary.py).  Kernel replay restores GBs of
# device memory per pass, so every capture costs seconds: keep the counts small.
N="ncu --set full --clock-control none"
# Merkle: preprocessed leaves (2^25, one column), then composition / FRI-first-layer region (skip counts from the launch list)
timeout 200 $N -k regex:commit_layer_kernel -c 3 -f -o gpurun_out/r1f_merkle_a python tools/run_one_proof.py
timeout 200 $N -k regex:commit_layer_kernel --launch-skip 50 -c 4 -f -o gpurun_out/r1f_merkle_b python tools/run_one_proof.py
# FFT: main-trace line transforms and the interaction / composition passes
timeout 200 $N -k regex:fft_kernel --launch-skip 95 -c 8 -f -o gpurun_out/r1f_fft_a python tools/run_one_proof.py
timeout 200 $N -k regex:fft_kernel --launch-skip 150 -c 6 -f -o gpurun_out/r1f_fft_b python tools/run_one_proof.py
timeout 200 $N -k regex:"quotients_kernel2" -c 2 -f -o gpurun_out/r1f_quot python tools/run_one_proof.py
timeout 200 $N -k regex:"constraint_kernel|ps_tile_kernel" -c 3 -f -o gpurun_out/r1f_misc python tools/run_one_proof.py
ls -la gpurun_out/*.ncu-rep
