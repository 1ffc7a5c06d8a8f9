#pragma once
#include "common.cuh"
#include "merkle.cuh"
#include <functional>
#include <vector>

// Extension elements are stored struct-of-arrays: comp[0][i] = a-part, comp[1][i] = b-part.

// out[2*c], out[2*c+1] = sum_j coeffs[c][j] * z^j for z in {zeta, zeta_next}:  StarkOpeningSet::new
void eval_columns_at_two_points(sbn_ctx* ctx, const u64* coeffs, int ncols, int logn, gl2 zeta, gl2 zeta_next, u64* d_out /* ncols x 4 */);
// The same with a power table built once per proof: pw = [z0^j].a | [z0^j].b | [z1^j].a | [z1^j].b, N entries each.
DevBuf<u64> two_point_power_table(sbn_ctx* ctx, int logn, gl2 z0, gl2 z1);
void eval_columns_at_two_points(sbn_ctx* ctx, const u64* coeffs, int ncols, int logn, const u64* pw, u64* d_out);

struct OracleView { const u64* coeffs; int ncols; };  // coefficient columns, stride N
// FRI batch reduction + division (PolynomialBatch::prove_openings up to `final_poly`): writes the
// zero-padded coefficient vector [2][N << rate_bits] of the polynomial FRI is run on.
// `split` (optional): this rank reduces only its slice of every oracle's columns and the partial sums of all ranks are
// all-gathered on the device and added (exact field additions, so the result is the same element).
struct ColumnSplit { int rank, world; std::function<void(const void* d_send, size_t nbytes, void* d_recv)> gather_device; };
void fri_final_poly(sbn_ctx* ctx, const std::vector<OracleView>& oracles, int logn, int rate_bits, gl2 alpha, gl2 zeta, gl2 zeta_next,
                    u64* d_final_coeffs, const ColumnSplit* split = nullptr, const u64* pw /* table of (zeta, zeta_next) or null */ = nullptr);

struct FriLayer { DevMerkleTree tree; DevBuf<u64> leaves; size_t nleaves; int arity_bits; };  // leaves[l][2*arity]
// Commit one layer: values [2][n] natural order -> leaves in bit-reversed order, chunks of 2^arity_bits.
void fri_commit_layer(sbn_ctx* ctx, const u64* d_values, int logsize, int arity_bits, int cap_height, FriLayer* out);
// coeffs [2][n] -> folded [2][n >> arity_bits] with beta:  c'_i = sum_k c_{i*arity+k} beta^k
void fri_fold_coeffs(sbn_ctx* ctx, const u64* d_coeffs, size_t n, int arity_bits, gl2 beta, u64* d_out);
// smallest w with leading_zeros(poseidon(state with state[pos] = w)[7]) >= pow_bits
u64 fri_pow_search(sbn_ctx* ctx, const u64 state[12], int pos, int pow_bits);

// sub_coset < 0: `lde` is the whole batch [col][b][k].  sub_coset = b (streamed commitment): `lde` holds sub-coset b only
// ([col][k], column stride N); the rows of queries that fall into another sub-coset are left untouched in the record, so the
// caller gathers into the same device records once per sub-coset.  lde == nullptr: paths only.
struct QueryOracle { const u64* lde; int ncols; const DevMerkleTree* tree; int sub_coset = -1; };
// Gathers all query openings into one flat u64 record per query (layout documented in fri.cu).
size_t fri_query_record_words(const std::vector<QueryOracle>& oracles, const std::vector<FriLayer*>& layers);
void fri_gather_queries(sbn_ctx* ctx, const std::vector<QueryOracle>& oracles, int logn, int rate_bits, const std::vector<FriLayer*>& layers,
                        const std::vector<u64>& indices, u64* h_out /* pinned or pageable host, nq * record_words; may be null */,
                        u64* d_dst = nullptr /* optional device destination of the same shape (kept there for a device exchange) */);
