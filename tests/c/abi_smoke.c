/* A plain C99 consumer of include/sdfb200.h: proves the header is valid C (not only C++) and that the
 * library links and answers without a device.  Built and run by tests/test_abi.py. */
#include <stdio.h>
#include <string.h>

#include "sdfb200.h"

int main(void) {
  sdfb_decoder* dec = NULL;
  sdfb_ddpm* ddpm = NULL;
  float blob[4] = {0.f, 0.f, 0.f, 0.f};
  size_t bytes = 0;
  int64_t n = 0;
  if (sdfb_version() < 100) return 1;
  if (sdfb_decoder_create(blob, 4, 0, &dec) != SDFB_E_INVALID || dec != NULL) return 2;   /* wrong blob size */
  if (strstr(sdfb_last_error(), "1839358") == NULL) return 3;
  if (sdfb_ddpm_create(blob, 4, 0, &ddpm) != SDFB_E_INVALID) return 4;
  if (sdfb_decoder_destroy(NULL) != SDFB_OK || sdfb_ddpm_destroy(NULL) != SDFB_OK) return 5;
  if (sdfb_mc_workspace_bytes(32, 32, 32, &bytes) != SDFB_OK || bytes == 0) return 6;
  if (sdfb_mc_count(NULL, NULL, 4, 4, 4, NULL, 0, &n, NULL) != SDFB_E_INVALID) return 7;
  if (sdfb_philox_normal(1u, -1, 4, 0, 1, NULL, NULL) != SDFB_E_INVALID) return 8;
  printf("sdfb C ABI ok: version %d, SDFB_DECODER_PARAM_FLOATS %d, SDFB_DDPM_STEPS %d\n", sdfb_version(),
         SDFB_DECODER_PARAM_FLOATS, SDFB_DDPM_STEPS);
  return 0;
}
