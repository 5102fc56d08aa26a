// Launch parameters of the AIR kernels (air_kernels.cu), filled by the C ABI (capi.cu).
#pragma once
#include "host/air.hpp"
#include "kernels.cuh"

namespace sb {
using namespace sbf;

struct AirParams {
  const uint32_t* const* main;   // main-trace columns (trace domain for LogUp generation, LDE for constraints)
  const uint32_t* const* inter;  // interaction columns on the LDE (constraints only)
  uint32_t* const* out;          // LogUp output columns (generation only)
  const uint32_t* is_first;      // IsFirst(log_size) on the LDE
  const QM31* coeff;             // per-constraint random-coefficient power
  InteractionElements el;
  QM31 total_sum;
  uint32_t log_size;             // trace log size
  uint32_t main_shift;           // LogUp generation: main columns hold one value per 2^main_shift rows (lane broadcast)
  uint32_t denom_inv[2];         // 1 / coset_vanishing on the two halves of the bit-reversed LDE
  uint32_t* acc[4];
  // row-range evaluation (multi-GPU): the column pointers cover rows [row_off, row_off + n_rows) of the LDE; the value of
  // the last LogUp column at coset offset -1 then comes from `prev` (that column pre-shifted by the owner, same rows).
  uint32_t row_off, n_rows;      // n_rows == 0: the whole domain
  const uint32_t* prev[4];       // nullable
};

int launch_air(bool constraints, int comp, const AirParams& p, cudaStream_t st);
void vanishing_denom_inv(uint32_t log_size, uint32_t out[2]);

}  // namespace sb
