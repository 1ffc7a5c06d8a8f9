// Host-side description of the AIRs the library can prove: column counts, permutation pairs and the
// ordered list of constraint segments (see constraints.cuh).  Mirrors the reference's
// `ExpStarkConstants` (src/constants.rs:4-16) / per-AIR `constants(num_io)` and
// `permutation_pairs()` (e.g. src/curves/g1/exp.rs:6-34, :735-741).
#pragma once
#include "common.cuh"
#include "../../include/starky_bn254_b200.h"
#include <utility>
#include <vector>


enum SegKind {
  SEG_SPLIT_RANGE_CHECK, SEG_MODULAR_CORE, SEG_G1_CORE, SEG_FLAGS, SEG_G1_ADD, SEG_G1_DOUBLE, SEG_PERIODIC_PULSE, SEG_PULSE,
  SEG_U16_RANGE_CHECK, SEG_PERMUTATION,
  SEG_FQ_CORE, SEG_FQ_MUL, SEG_G2_CORE, SEG_G2_ADD, SEG_G2_DOUBLE, SEG_FQ12_CORE, SEG_FQ12_MUL, SEG_FLAGS_U64
};
struct Segment { SegKind kind; int p0, p1, p2, p3; size_t num_constraints; };

struct AirDesc {
  int air_id = 0; size_t num_io = 0;
  size_t num_columns = 0, num_public_inputs = 0, num_rows = 0, io_size = 0, result_words = 0;
  int constraint_degree = 3;
  std::vector<std::pair<u32, u32>> perm_pairs;
  std::vector<Segment> segments;   // AIR segments in emission order (the permutation segment is appended by the prover)
  size_t num_air_constraints() const { size_t n = 0; for (auto& s : segments) n += s.num_constraints; return n; }
  int quotient_degree_factor() const { int d = constraint_degree - 1; return d < 1 ? 1 : d; }
};

AirDesc make_air(int air_id, size_t num_io);
