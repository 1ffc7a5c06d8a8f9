#pragma once
#include "common.cuh"
const u64* get_pow_table(sbn_ctx* ctx, u64 base, int logn);
NttTables get_ntt_tables(sbn_ctx* ctx, int logn);
void ntt_batch(sbn_ctx* ctx, const u64* in, size_t in_stride, u64* out, size_t out_stride, int ncols, int logn, bool inverse,
               u64 pre_base, const u64* postscale);
void intt_columns(sbn_ctx* ctx, const u64* values, u64* coeffs, int ncols, int logn);
void lde_columns(sbn_ctx* ctx, const u64* coeffs, u64* lde, int ncols, int logn, int rate_bits);
void lde_sub_coset(sbn_ctx* ctx, const u64* coeffs, u64* out, int ncols, int logn, int rate_bits, int b);
// Values of every column on LDE class rho of G = 2^m (natural LDE indices i = rho mod G): out[col][b'][k'], L / G points per column.
void lde_class(sbn_ctx* ctx, const u64* coeffs, u64* out, int ncols, int logn, int rate_bits, int m, u32 rho);
