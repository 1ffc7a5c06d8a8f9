#pragma once
#include "common.cuh"
#include <vector>

// Device-resident Merkle tree (plonky2 `MerkleTree::new(leaves, cap_height)`): all digest levels
// from the leaf digests (level 0) up to the cap level are kept for query-time path extraction.
struct DevMerkleTree {
  DevBuf<u64> digests;             // concatenated levels, 4 u64 per digest
  std::vector<size_t> level_off;   // offset (in digests) of each level
  size_t nleaves = 0; int cap_height = 0;
  std::vector<u64> cap;            // host copy of the cap (2^cap_height x 4)
  int num_levels() const { return (int)level_off.size(); }
  int proof_len() const { return num_levels() - 1; }
};

// Hash the rows of a column-major LDE batch lde[col][b][k] (see ntt.cu) into leaf digests placed in
// plonky2's bit-reversed leaf order, then build the tree and copy the cap to the host.
void merkle_commit_lde(sbn_ctx* ctx, const u64* lde, int ncols, int logn, int rate_bits, int cap_height, DevMerkleTree* out);
// Build the upper levels from already-written leaf digests (tree->digests level 0) and fetch the cap.
void merkle_build_from_leaf_digests(sbn_ctx* ctx, DevMerkleTree* t);
void merkle_alloc(sbn_ctx* ctx, DevMerkleTree* t, size_t nleaves, int cap_height);
// Leaf digests of ONE sub-coset b (values sub[col][k], column stride N) of a streamed commitment: the LDE is produced, hashed
// and dropped one sub-coset at a time (config 5 at 2^22 rows does not fit otherwise); same digest positions as the full batch.
void merkle_leaf_hash_sub_coset(sbn_ctx* ctx, const u64* sub, int ncols, int logn, int rate_bits, int b, DevMerkleTree* t);
// Leaf digests only (level 0 of an already allocated tree); used by the commitment micro-benchmark.
void merkle_leaf_hash_only(sbn_ctx* ctx, const u64* lde, int ncols, int logn, int rate_bits, DevMerkleTree* t);
