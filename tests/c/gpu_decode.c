/* A plain C99 consumer of include/sdfb200.h that actually computes on the GPU: builds a decoder from a parameter blob,
 * decodes a 32^3 grid through sdfb_decode_grid_host (fp32 path and bf16 tensor-core path, with the sign-change mask) and
 * compares with a committed golden field (the fp32 CPU oracle's output).  Built and run by tests/test_gpu_boundary.py.
 *   usage: gpu_decode params.bin latent.bin golden32.bin */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "sdfb200.h"

#define RES 32
#define NQ (RES * RES * RES)
#define NC ((RES - 1) * (RES - 1) * (RES - 1))

static float* read_floats(const char* path, size_t n) {
  FILE* f = fopen(path, "rb");
  float* p;
  if (!f) return NULL;
  p = (float*)malloc(n * sizeof(float));
  if (p && fread(p, sizeof(float), n, f) != n) { free(p); p = NULL; }
  fclose(f);
  return p;
}

static int inside(float v) { return v < 0.0f; }

int main(int argc, char** argv) {
  sdfb_decoder* dec = NULL;
  float *params, *latent, *golden, *sdf32, *sdf16;
  uint8_t* mask;
  double e32 = 0.0, e16 = 0.0;
  long bad_mask = 0, active = 0;
  int i, rc;
  if (argc != 4) { fprintf(stderr, "usage: %s params.bin latent.bin golden32.bin\n", argv[0]); return 2; }
  params = read_floats(argv[1], SDFB_DECODER_PARAM_FLOATS);
  latent = read_floats(argv[2], SDFB_LATENT_DIM);
  golden = read_floats(argv[3], NQ);
  sdf32 = (float*)malloc(NQ * sizeof(float));
  sdf16 = (float*)malloc(NQ * sizeof(float));
  mask = (uint8_t*)malloc(NC);
  if (!params || !latent || !golden || !sdf32 || !sdf16 || !mask) { fprintf(stderr, "cannot read the inputs\n"); return 3; }
  rc = sdfb_decoder_create(params, SDFB_DECODER_PARAM_FLOATS, 0, &dec);
  if (rc != SDFB_OK) { fprintf(stderr, "create: %d %s\n", rc, sdfb_last_error()); return 4; }
  rc = sdfb_decode_grid_host(dec, latent, RES, 0, RES, sdf32, NULL, SDFB_PREC_FP32);
  if (rc != SDFB_OK) { fprintf(stderr, "fp32 decode: %d %s\n", rc, sdfb_last_error()); return 5; }
  rc = sdfb_decode_grid_host(dec, latent, RES, 0, RES, sdf16, mask, SDFB_PREC_BF16);
  if (rc != SDFB_OK) { fprintf(stderr, "bf16 decode: %d %s\n", rc, sdfb_last_error()); return 6; }
  for (i = 0; i < NQ; ++i) {
    double d32 = fabs((double)sdf32[i] - (double)golden[i]), d16 = fabs((double)sdf16[i] - (double)golden[i]);
    if (d32 > e32) e32 = d32;
    if (d16 > e16) e16 = d16;
  }
  /* the mask must be exactly the sign-change mask of the bf16 field the same call returned (rule A4) */
  for (i = 0; i < NC; ++i) {
    int x = i % (RES - 1), y = (i / (RES - 1)) % (RES - 1), z = i / ((RES - 1) * (RES - 1));
    int any = 0, all = 1, dz, dy, dx;
    for (dz = 0; dz < 2; ++dz)
      for (dy = 0; dy < 2; ++dy)
        for (dx = 0; dx < 2; ++dx) {
          int in = inside(sdf16[((z + dz) * RES + (y + dy)) * RES + (x + dx)]);
          any |= in;
          all &= in;
        }
    if ((uint8_t)(any && !all) != mask[i]) ++bad_mask;
    active += mask[i];
  }
  if (sdfb_decoder_check(dec, NULL) != SDFB_OK) { fprintf(stderr, "check: %s\n", sdfb_last_error()); return 7; }
  sdfb_decoder_destroy(dec);
  printf("fp32 path max |sdf - golden| = %.3e, bf16 path = %.3e, active cells %ld, mask mismatches %ld\n", e32, e16, active,
         bad_mask);
  if (!(e32 < 1e-5)) return 8;       /* north star: 1e-5 on the fp32 path */
  if (!(e16 < 2e-2)) return 9;       /* bf16 operands: reported distance to fp32, bounded */
  if (bad_mask != 0 || active == 0) return 10;
  printf("gpu_decode ok\n");
  free(params); free(latent); free(golden); free(sdf32); free(sdf16); free(mask);
  return 0;
}
