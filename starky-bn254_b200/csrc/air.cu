#include "air.cuh"

static void add_split_pairs(AirDesc& a, u32 main_col, u32 nt) {  // reference src/utils/range_check.rs:230-246
  for (u32 i = main_col + 1; i < main_col + 1 + 6 * nt; i += 6) {
    a.perm_pairs.push_back({main_col, i + 2}); a.perm_pairs.push_back({main_col, i + 5});
    a.perm_pairs.push_back({i, i + 1}); a.perm_pairs.push_back({i + 3, i + 4});
  }
}
static void add_u16_pairs(AirDesc& a, u32 start_lookups, u32 t0, u32 nt) {  // reference src/utils/range_check.rs:96-113
  for (u32 i = 0; i < nt; i++) {
    a.perm_pairs.push_back({start_lookups, start_lookups + 1 + 2 * i + 1});
    a.perm_pairs.push_back({t0 + i, start_lookups + 1 + 2 * i});
  }
}

AirDesc make_air(int air_id, size_t num_io) {
  AirDesc a; a.air_id = air_id; a.num_io = num_io;
  SBN_REQUIRE(num_io >= 1, "num_io must be positive");
  switch (air_id) {
    case SBN_AIR_MODULAR: {  // reference src/modular/modular.rs:361-369 (rows = num_io, a power of two >= 256)
      const u32 MAIN = 145, T0 = 32, NT = 111;
      SBN_REQUIRE((num_io & (num_io - 1)) == 0 && num_io >= 256, "ModularStark: rows must be a power of two >= 256");
      a.num_columns = MAIN + 1 + 6 * NT; a.num_public_inputs = 0; a.num_rows = num_io; a.io_size = 64; a.result_words = 0;
      add_split_pairs(a, MAIN, NT);
      a.segments.push_back({SEG_SPLIT_RANGE_CHECK, (int)MAIN, (int)T0, (int)NT, 0, 5 * (size_t)NT + 3});
      a.segments.push_back({SEG_MODULAR_CORE, 0, 0, 0, 0, 66});
      break;
    }
    case SBN_AIR_G1_EXP: {  // reference src/curves/g1/exp.rs:6-34
      SBN_REQUIRE((num_io & (num_io - 1)) == 0 && num_io >= 128, "G1ExpStark: num_io must be a power of two >= 128 (u16 lookup table needs 2^16 rows)");
      const u32 sf = 24 * 16, main_cols = sf + 14, pp = main_cols, iop = pp + 2, lookups = iop + 1 + 4 * (u32)num_io, nrc = 24 * 16 - 3;
      a.num_columns = lookups + 1 + 2 * nrc; a.num_public_inputs = 56 * num_io; a.num_rows = 512 * num_io; a.io_size = 224; a.result_words = 8;
      add_u16_pairs(a, lookups, 0, nrc);
      a.segments.push_back({SEG_G1_CORE, (int)num_io, 64 + 16, (int)sf, 0, 1 + 56 * num_io + 192});
      a.segments.push_back({SEG_FLAGS, (int)sf, 0, 0, 0, 26});
      a.segments.push_back({SEG_G1_ADD, 64, (int)sf + 4, 0, 0, 165});
      a.segments.push_back({SEG_G1_DOUBLE, 64, (int)sf + 2, 0, 0, 165});
      a.segments.push_back({SEG_FLAGS, (int)sf, 0, 0, 0, 26});  // emitted twice: exp.rs:462 and :467-472
      a.segments.push_back({SEG_PERIODIC_PULSE, (int)sf + 1, (int)pp, 64, 62, 5});
      a.segments.push_back({SEG_PULSE, (int)iop, (int)num_io, 512, 0, 2 + 4 * num_io});
      a.segments.push_back({SEG_U16_RANGE_CHECK, (int)lookups, (int)nrc, 0, 0, 2 * (size_t)nrc + 3});
      break;
    }
    case SBN_AIR_FQ_EXP: {  // reference src/fields/fq/exp.rs:6-34
      SBN_REQUIRE((num_io & (num_io - 1)) == 0 && num_io >= 128, "FqExpStark: num_io must be a power of two >= 128 (u16 lookup table needs 2^16 rows)");
      const u32 sf = 9 * 16, main_cols = sf + 14, pp = main_cols, iop = pp + 2, lookups = iop + 1 + 4 * (u32)num_io, nrc = 9 * 16 - 1;
      a.num_columns = lookups + 1 + 2 * nrc; a.num_public_inputs = 32 * num_io; a.num_rows = 512 * num_io; a.io_size = sizeof(sbn_fq_exp_io); a.result_words = 4;
      add_u16_pairs(a, lookups, 0, nrc);
      a.segments.push_back({SEG_FQ_CORE, (int)num_io, 32, (int)sf, 0, 1 + 32 * num_io + 96});
      a.segments.push_back({SEG_FLAGS, (int)sf, 0, 0, 0, 26});
      a.segments.push_back({SEG_FQ_MUL, (int)sf + 2, 1, 0, 0, 66});   // eval_fq_mul(is_sq, a, a)
      a.segments.push_back({SEG_FQ_MUL, (int)sf + 4, 0, 0, 0, 66});   // eval_fq_mul(is_mul, a, b)
      a.segments.push_back({SEG_FLAGS, (int)sf, 0, 0, 0, 26});        // emitted twice: fq/exp.rs:361 and :366-371
      a.segments.push_back({SEG_PERIODIC_PULSE, (int)sf + 1, (int)pp, 64, 62, 5});
      a.segments.push_back({SEG_PULSE, (int)iop, (int)num_io, 512, 0, 2 + 4 * num_io});
      a.segments.push_back({SEG_U16_RANGE_CHECK, (int)lookups, (int)nrc, 0, 0, 2 * (size_t)nrc + 3});
      break;
    }
    case SBN_AIR_G2_EXP: {  // reference src/curves/g2/exp.rs:6-34
      SBN_REQUIRE((num_io & (num_io - 1)) == 0 && num_io >= 128, "G2ExpStark: num_io must be a power of two >= 128 (u16 lookup table needs 2^16 rows)");
      const u32 sf = 48 * 16, main_cols = sf + 14, pp = main_cols, iop = pp + 2, lookups = iop + 1 + 4 * (u32)num_io, nrc = 48 * 16 - 6;
      a.num_columns = lookups + 1 + 2 * nrc; a.num_public_inputs = 104 * num_io; a.num_rows = 512 * num_io; a.io_size = sizeof(sbn_g2_exp_io); a.result_words = 16;
      add_u16_pairs(a, lookups, 0, nrc);
      a.segments.push_back({SEG_G2_CORE, (int)num_io, 128 + 32, (int)sf, 0, 1 + 104 * num_io + 384});
      a.segments.push_back({SEG_FLAGS, (int)sf, 0, 0, 0, 26});
      a.segments.push_back({SEG_G2_ADD, 128, (int)sf + 4, 0, 0, 330});
      a.segments.push_back({SEG_G2_DOUBLE, 128, (int)sf + 2, 0, 0, 330});
      a.segments.push_back({SEG_FLAGS, (int)sf, 0, 0, 0, 26});        // emitted twice: g2/exp.rs:474 and :479-484
      a.segments.push_back({SEG_PERIODIC_PULSE, (int)sf + 1, (int)pp, 64, 62, 5});
      a.segments.push_back({SEG_PULSE, (int)iop, (int)num_io, 512, 0, 2 + 4 * num_io});
      a.segments.push_back({SEG_U16_RANGE_CHECK, (int)lookups, (int)nrc, 0, 0, 2 * (size_t)nrc + 3});
      break;
    }
    case SBN_AIR_FQ12_EXP: {  // reference src/fields/fq12/exp.rs:6-34
      SBN_REQUIRE((num_io & (num_io - 1)) == 0, "Fq12ExpStark: num_io must be a power of two");
      const u32 sf = 108 * 16, main_cols = sf + 14, pp = main_cols, iop = pp + 2, lookups = iop + 1 + 4 * (u32)num_io, t0 = 24 * 16, nrc = 84 * 16 - 12;
      a.num_columns = lookups + 1 + 6 * nrc; a.num_public_inputs = 584 * num_io; a.num_rows = 512 * num_io; a.io_size = sizeof(sbn_fq12_exp_io); a.result_words = 48;
      add_split_pairs(a, lookups, nrc);
      a.segments.push_back({SEG_FQ12_CORE, (int)num_io, (int)sf, 0, 0, 1 + 584 * num_io + 1152});
      a.segments.push_back({SEG_FLAGS, (int)sf, 0, 0, 0, 26});
      a.segments.push_back({SEG_FQ12_MUL, (int)sf + 2, 1, 0, 0, 12 * 66});   // eval_fq12_mul(is_sq, a, a)
      a.segments.push_back({SEG_FQ12_MUL, (int)sf + 4, 0, 0, 0, 12 * 66});   // eval_fq12_mul(is_mul, a, b)
      a.segments.push_back({SEG_FLAGS, (int)sf, 0, 0, 0, 26});               // emitted twice: fq12/exp.rs:395 and :400-405
      a.segments.push_back({SEG_PERIODIC_PULSE, (int)sf + 1, (int)pp, 64, 62, 5});
      a.segments.push_back({SEG_PULSE, (int)iop, (int)num_io, 512, 0, 2 + 4 * num_io});
      a.segments.push_back({SEG_SPLIT_RANGE_CHECK, (int)lookups, (int)t0, (int)nrc, 0, 5 * (size_t)nrc + 3});
      break;
    }
    case SBN_AIR_FQ12_EXP_U64: {  // reference src/fields/fq12_u64/exp_u64.rs:19-45 (128 rows per instance, 6 flag columns, no rotation)
      SBN_REQUIRE((num_io & (num_io - 1)) == 0 && num_io >= 2, "Fq12ExpU64Stark: num_io must be a power of two >= 2 (split lookup table needs 256 rows)");
      const u32 sf = 108 * 16, main_cols = sf + 6, iop = main_cols, lookups = iop + 1 + 4 * (u32)num_io, t0 = 24 * 16, nrc = 84 * 16 - 12;
      a.num_columns = lookups + 1 + 6 * nrc; a.num_public_inputs = 577 * num_io; a.num_rows = 128 * num_io; a.io_size = sizeof(sbn_fq12_exp_u64_io); a.result_words = 48;
      add_split_pairs(a, lookups, nrc);
      a.segments.push_back({SEG_FQ12_CORE, (int)num_io, (int)sf, 1, 0, 1 + 577 * num_io + 1152});
      a.segments.push_back({SEG_FLAGS_U64, (int)sf, 0, 0, 0, 9});
      a.segments.push_back({SEG_FQ12_MUL, (int)sf + 1, 1, 0, 0, 12 * 66});
      a.segments.push_back({SEG_FQ12_MUL, (int)sf + 3, 0, 0, 0, 12 * 66});
      a.segments.push_back({SEG_FLAGS_U64, (int)sf, 0, 0, 0, 9});            // emitted twice: exp_u64.rs:385 and :390-395
      a.segments.push_back({SEG_PULSE, (int)iop, (int)num_io, 128, 0, 2 + 4 * num_io});
      a.segments.push_back({SEG_SPLIT_RANGE_CHECK, (int)lookups, (int)t0, (int)nrc, 0, 5 * (size_t)nrc + 3});
      break;
    }
    case SBN_AIR_G1_MULADD: {  // reference src/curves/g1/muladd.rs:462-468 (rows = num_io, a power of two >= 256)
      const u32 MAIN = 24 * 16 + 2, T0 = 4 * 16, NT = 20 * 16 - 4;
      SBN_REQUIRE((num_io & (num_io - 1)) == 0 && num_io >= 256, "G1Stark: rows must be a power of two >= 256");
      a.num_columns = MAIN + 1 + 6 * NT; a.num_public_inputs = 0; a.num_rows = num_io; a.io_size = sizeof(sbn_g1_muladd_io); a.result_words = 0;
      add_split_pairs(a, MAIN, NT);
      a.segments.push_back({SEG_SPLIT_RANGE_CHECK, (int)MAIN, (int)T0, (int)NT, 0, 5 * (size_t)NT + 3});
      a.segments.push_back({SEG_G1_ADD, 64, 384, 0, 0, 165});      // eval_g1_add(is_add, ...)       muladd.rs:579
      a.segments.push_back({SEG_G1_DOUBLE, 64, 385, 0, 0, 165});   // eval_g1_double(is_double, ...) muladd.rs:580
      break;
    }
    case SBN_AIR_FQ12_MUL: {   // reference src/fields/fq12/mul.rs:355-363 (rows = num_io, a power of two >= 256)
      const u32 MAIN = 108 * 16 + 1, T0 = 24 * 16, NT = 84 * 16 - 12;
      SBN_REQUIRE((num_io & (num_io - 1)) == 0 && num_io >= 256, "Fq12Stark: rows must be a power of two >= 256");
      a.num_columns = MAIN + 1 + 6 * NT; a.num_public_inputs = 0; a.num_rows = num_io; a.io_size = sizeof(sbn_fq12_mul_io); a.result_words = 0;
      add_split_pairs(a, MAIN, NT);
      a.segments.push_back({SEG_SPLIT_RANGE_CHECK, (int)MAIN, (int)T0, (int)NT, 0, 5 * (size_t)NT + 3});
      a.segments.push_back({SEG_FQ12_MUL, 108 * 16, 0, 0, 0, 12 * 66});   // eval_fq12_mul(filter, x, y)  mul.rs:447
      break;
    }
    default: throw SbnError(SBN_ERR_UNSUPPORTED, "unknown AIR identifier");
  }
  return a;
}
