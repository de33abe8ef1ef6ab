/* A plain-C consumer of include/msmd_b200.h: compiled with gcc -std=c99 (no C++, no torch), linked against
 * libmsmd_b200.so by tests/test_abi.py.  Exercises only calls that need no GPU: version string, argument
 * validation and the error channel. */
#include <stdio.h>
#include <string.h>
#include "msmd_b200.h"

int main(void) {
  const char* v = msmd_version();
  if (v == NULL || strstr(v, "sm_100a") == NULL) { printf("bad version\n"); return 1; }
  float in[3] = {0.f, 0.f, 0.f}, out[9];
  if (msmd_rot_convert(999, in, out, 1, 0, NULL) != MSMD_ERR_INVALID) { printf("unknown kind accepted\n"); return 2; }
  if (strstr(msmd_last_error(), "unknown kind") == NULL) { printf("no message: %s\n", msmd_last_error()); return 3; }
  if (msmd_rot_convert(0, NULL, NULL, 0, 0, NULL) != MSMD_OK) { printf("empty input rejected\n"); return 4; }
  msmd_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  msmd_model* m = NULL;
  if (msmd_create(&cfg, 0, &m) == MSMD_OK) { printf("all-zero config accepted\n"); return 5; }
  if (msmd_audio_normalize(NULL, NULL, 0, 0, NULL) != MSMD_OK) { printf("empty normalise rejected\n"); return 6; }
  msmd_sample_extras ex;
  memset(&ex, 0, sizeof(ex));
  ex.precise_last_steps = -1;
  printf("ok %s\n", v);
  return 0;
}
