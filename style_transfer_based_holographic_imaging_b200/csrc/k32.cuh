// k32.cuh -- the FFT-size-1024 fast path: 32 points per thread, 1024 = 32 x 32.
//
// Shared-memory bandwidth (128 B/clk/SM) is the scarcest resource of this pipeline after the FP32 pipe, so this
// path makes ONE shared-memory exchange per 1-D FFT and moves everything else register <-> global directly:
//   k32_rows_fwd : one WARP per row.  lane l loads x[l + 32 i] (coalesced), radix-32 in registers, exchange
//                  through the warp's private 8.25 KB line (syncwarp only, no CTA barrier), radix-32 with table
//                  twiddles, and stores X[l + 32 i] straight from registers: register i of lane l holds frequency
//                  l + 32 i, so the workspace row is in NATURAL frequency order and the store is coalesced.
//   k32_cols     : 8 columns x 32 threads.  ws -> registers (64 B segments), radix-32, exchange, radix-32,
//                  x H(z) (kappa slab staged once per CTA with cp.async), inverse radix-32, exchange, radix-32,
//                  registers -> ws.
//   k32_rows_inv : mirror of k32_rows_fwd with the output stage fused.
// Included by asm_b200.cu (needs Params, load_one, emit_one, ...).
#pragma once
#include "k32_common.cuh"
#include "k32_rows.cuh"
#include "k32_cols.cuh"
#include "k32_persistent.cuh"
