// k32.cuh -- the FFT-size-1024 fast path: 32 points per thread, 1024 = 32 x 32.
//
// Shared-memory bandwidth (128 B/clk/SM) is the scarcest resource of this pipeline after the FP32 pipe, so this
// path makes ONE shared-memory exchange per 1-D FFT:
//   row kernels (k32_rows.cuh) : one WARP per row.  lane l takes x[l + 32 i], radix-32 in registers, exchange
//                  through the warp's private 8.25 KB line (syncwarp only, no CTA barrier), radix-32 with table
//                  twiddles; register i of lane l then holds frequency l + 32 i, so the workspace row is in NATURAL
//                  frequency order.  Rows move HBM/L2 <-> shared memory as single TMA bulk copies.
//   column kernel (k32_cols.cuh): CC columns x 32 threads.  cp.async-staged slab -> registers, radix-32, exchange,
//                  radix-32, x H(z) (kappa slab staged with cp.async), inverse radix-32, exchange, radix-32,
//                  registers -> workspace.
// Included by asm_b200.cu (needs Params, load_one, emit_one, ...).
#pragma once
#include "k32_common.cuh"
#include "k32_rows.cuh"
#include "k32_cols.cuh"
