/* Pure-C consumer of include/starky_bn254_b200.h: the header must compile as C99, the record layouts must be the ones the
 * reference-side shim mirrors, and without a device the library must fail loudly (no CPU fallback).
 * With a device it proves one small ModularStark trace end to end through the C ABI alone.  Exit code 0 = all checks passed. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "starky_bn254_b200.h"

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "abi_check: %s failed (line %d)\n", #c, __LINE__); return 1; } } while (0)

int main(void) {
  CHECK(sizeof(sbn_g1_exp_io) == 224 && sizeof(sbn_fq_exp_io) == 128 && sizeof(sbn_g2_exp_io) == 416);
  CHECK(sizeof(sbn_fq12_exp_io) == 1184 && sizeof(sbn_fq12_exp_u64_io) == 1160 && sizeof(sbn_modular_io) == 64);
  CHECK(sizeof(sbn_g1_muladd_io) == 128 && sizeof(sbn_fq12_mul_io) == 768);
  sbn_config cfg;
  CHECK(sbn_config_standard_fast(&cfg) == 0);
  CHECK(cfg.security_bits == 100 && cfg.num_challenges == 2 && cfg.rate_bits == 1 && cfg.cap_height == 4 && cfg.pow_bits == 16);
  CHECK(cfg.fri_arity_bits == 4 && cfg.fri_final_poly_bits == 5 && cfg.num_query_rounds == 84);
  size_t ncols = 0, npis = 0, nrows = 0, io_size = 0, v4 = 0, v5 = 0;
  CHECK(sbn_air_info(SBN_AIR_G1_EXP, 128, &ncols, &npis, &nrows, &io_size, &v4, &v5) == 0);
  CHECK(ncols == 1676 && npis == 7168 && nrows == 65536 && io_size == sizeof(sbn_g1_exp_io));
  sbn_ctx* ctx = NULL;
  int rc = sbn_ctx_create(0, NULL, &ctx);
  if (rc != 0) {   /* no device: the error must say so */
    CHECK(ctx == NULL && strstr(sbn_last_error(NULL), "no CPU fallback") != NULL);
    printf("abi_check: ok (no device: %s)\n", sbn_last_error(NULL));
    return 0;
  }
  /* with a device: 256 rows of ModularStark, trace generation + prove + serialize */
  enum { ROWS = 256 };
  sbn_modular_io* ios = (sbn_modular_io*)calloc(ROWS, sizeof *ios);
  for (int r = 0; r < ROWS; r++) { ios[r].input0[0] = 3u + (unsigned)r; ios[r].input1[0] = 5u + 7u * (unsigned)r; ios[r].input1[1] = 11; }
  sbn_trace* tr = NULL; sbn_proof* pf = NULL;
  CHECK(sbn_trace_generate(ctx, SBN_AIR_MODULAR, ios, ROWS, &tr) == 0);
  CHECK(sbn_prove(ctx, &cfg, tr, NULL, 0, &pf) == 0);
  size_t len = 0;
  CHECK(sbn_proof_serialize(pf, NULL, &len) == 0 && len > 100000);
  unsigned char* buf = (unsigned char*)malloc(len);
  CHECK(sbn_proof_serialize(pf, buf, &len) == 0);
  CHECK(buf[0] == 16 && buf[1] == 0 && buf[2] == 0 && buf[3] == 0);   /* trace cap: u32 length prefix = 2^cap_height digests */
  printf("abi_check: ok (%zu-byte proof, %llu kernel launches)\n", len, (unsigned long long)sbn_ctx_launch_count(ctx));
  free(buf); free(ios);
  sbn_proof_free(pf); sbn_trace_free(tr); sbn_ctx_destroy(ctx);
  return 0;
}
