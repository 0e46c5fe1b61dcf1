/* Plain C99 client of include/ltxcuda.h -- what the SwiftPM C target (CLTXCuda) imports.  Built and run by
 * tests/test_abi.py::test_header_is_plain_c_and_links on the GPU-less host: the header must compile as C (no C++ types
 * in the signatures), the library must link, the pure host entry points must work and, without an sm_100a device,
 * ltx_ctx_create must fail with LTX_ERR_CUDA and a message (no CPU fallback).  With a device it creates and destroys a
 * context instead. */
#include <stdio.h>
#include <string.h>

#include "ltxcuda.h"

int main(void) {
  ltx_config cfg;
  ltx_ctx* ctx = NULL;
  int rc;
  if (strncmp(ltx_version(), "ltxcuda", 7) != 0) return 10;
  ltx_config_default(&cfg);
  if (cfg.num_layers != 48 || cfg.num_heads != 32 || cfg.head_dim != 128 || cfg.caption_channels != 3840) return 11;
  if (ltx_vae_tiled_frames(16, 8, 1) != 107 || ltx_vae_tiled_frames(4, 8, 1) != 25 || ltx_vae_tiled_frames(16, 4, 4) != -1) return 12;
  rc = ltx_ctx_create(&cfg, 0, &ctx);
  if (rc == LTX_OK) {
    printf("device present: context created\n");
    if (ltx_ctx_destroy(ctx) != LTX_OK) return 13;
    return 0;
  }
  if (rc != LTX_ERR_CUDA || ctx != NULL) return 14;
  if (strstr(ltx_last_error(NULL), "no CPU fallback") == NULL) return 15;
  printf("no device: %s\n", ltx_last_error(NULL));
  return 0;
}
