#pragma once
#include "quotient.cuh"
// z_out[z][r], r < 2^logn, for every batch in `perm` (trace is column-major values, stride 2^logn).
void compute_z_polys(sbn_ctx* ctx, const u64* trace, int logn, const PermInstances& perm, u64* z_out);
