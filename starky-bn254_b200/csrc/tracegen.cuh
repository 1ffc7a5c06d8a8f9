#pragma once
#include "air.cuh"
// K1: `stark.generate_trace(&inputs)` on the device.  ios: array of the AIR's input records (host, or device if ios_on_device).
// d_cols: num_columns x num_rows column-major device buffer.  h_results: result_words u64 per io (host).
void generate_trace(sbn_ctx* ctx, const AirDesc& air, const void* ios, bool ios_on_device, u64* d_cols, u64* h_results);
// `stark.generate_public_inputs(&inputs)`; host-side formatting only.
void format_public_inputs(const AirDesc& air, const void* ios, u64* out);
// Range-check lookup columns over device-resident trace columns (shared by all AIRs).
void generate_u16_range_check_cols(sbn_ctx* ctx, u64* d_cols, size_t N, int t0, int ntargets, int start_lookups);
void generate_split_u16_range_check_cols(sbn_ctx* ctx, u64* d_cols, size_t N, int t0, int ntargets, int main_col);
