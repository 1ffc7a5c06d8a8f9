/* Driver for compute-sanitizer runs (not product code): proves one trace through the C ABI from an input-record file.
 *   prove_file <air id> <num_io> <ios file> [lanes]      lanes > 0: the same inputs twice through sbn_prove_batch
 * Build: gcc -std=c99 -I include tools/sanitize/prove_file.c -o prove_file -L starky-bn254_b200 -lstarkybn254_b200 -Wl,-rpath,$PWD/starky-bn254_b200 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "starky_bn254_b200.h"

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: prove_file air num_io ios_file [lanes]\n"); return 2; }
  int air = atoi(argv[1]); size_t num_io = (size_t)atol(argv[2]); int lanes = argc > 4 ? atoi(argv[4]) : 0;
  size_t ncols, npis, nrows, io_size, rw, pairs;
  if (sbn_air_info(air, num_io, &ncols, &npis, &nrows, &io_size, &rw, &pairs) != 0) { fprintf(stderr, "air_info: %s\n", sbn_last_error(NULL)); return 1; }
  unsigned char* ios = (unsigned char*)malloc(io_size * num_io);
  FILE* f = fopen(argv[3], "rb");
  if (!f || fread(ios, 1, io_size * num_io, f) != io_size * num_io) { fprintf(stderr, "cannot read %zu bytes from %s\n", io_size * num_io, argv[3]); return 1; }
  fclose(f);
  sbn_config cfg; sbn_config_standard_fast(&cfg);
  if (lanes > 0) {
    sbn_batch* b = NULL;
    if (sbn_batch_create(0, (uint32_t)lanes, &b) != 0) { fprintf(stderr, "batch_create: %s\n", sbn_batch_last_error(NULL)); return 1; }
    const void* in[2] = {ios, ios}; sbn_proof* out[2] = {NULL, NULL};
    if (sbn_prove_batch(b, air, num_io, &cfg, in, 2, SBN_BATCH_FILL_OUTPUTS, out) != 0) { fprintf(stderr, "prove_batch: %s\n", sbn_batch_last_error(b)); return 1; }
    size_t l0 = 0, l1 = 0; sbn_proof_serialize(out[0], NULL, &l0); sbn_proof_serialize(out[1], NULL, &l1);
    printf("prove_file: batch ok, proofs of %zu and %zu bytes, %llu launches\n", l0, l1, (unsigned long long)sbn_batch_launch_count(b));
    sbn_proof_free(out[0]); sbn_proof_free(out[1]); sbn_batch_destroy(b);
    return l0 == l1 ? 0 : 1;
  }
  sbn_ctx* ctx = NULL;
  if (sbn_ctx_create(0, NULL, &ctx) != 0) { fprintf(stderr, "ctx_create: %s\n", sbn_last_error(NULL)); return 1; }
  sbn_trace* tr = NULL; sbn_proof* pf = NULL;
  if (sbn_trace_generate(ctx, air, ios, num_io, &tr) != 0) { fprintf(stderr, "trace: %s\n", sbn_last_error(ctx)); return 1; }
  if (rw) {   /* output field = the chain result (trailing result_words u64 of every record) */
    uint64_t* res = (uint64_t*)malloc(rw * num_io * 8);
    sbn_trace_results(tr, res);
    for (size_t i = 0; i < num_io; i++) memcpy(ios + (i + 1) * io_size - rw * 8, res + i * rw, rw * 8);
    free(res);
  }
  uint64_t* pis = (uint64_t*)malloc((npis ? npis : 1) * 8);
  if (sbn_public_inputs(air, ios, num_io, pis, npis) != 0) { fprintf(stderr, "public inputs: %s\n", sbn_last_error(NULL)); return 1; }
  if (sbn_prove(ctx, &cfg, tr, pis, npis, &pf) != 0) { fprintf(stderr, "prove: %s\n", sbn_last_error(ctx)); return 1; }
  size_t len = 0; sbn_proof_serialize(pf, NULL, &len);
  printf("prove_file: ok, %zu-byte proof, %llu launches\n", len, (unsigned long long)sbn_ctx_launch_count(ctx));
  sbn_proof_free(pf); sbn_trace_free(tr); sbn_ctx_destroy(ctx); free(pis); free(ios);
  return 0;
}
