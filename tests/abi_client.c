/* Plain C99 client of include/kml.h: what a cgo / JNI / ROS-node binding of the reference side
 * would compile against.  Built and run by tests/test_abi.py (no Python, no torch in the process).
 * Without a usable GPU it checks that creation fails loudly; with one it runs a one-pair
 * add -> computeMatchedIndices round trip through the kernels. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kml.h"

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "abi_client: %s failed (line %d)\n", #c, __LINE__); return 1; } } while (0)

int main(void) {
  kml_params p;
  kml_handle* h = NULL;
  int rc, i, n = 0;
  memset(&p, 0xff, sizeof p);
  kml_default_params(&p);
  CHECK(p.max_db_results == 50 && p.top_k_verify == 16 && p.ransac_seed == 12345);
  CHECK(kml_create(&p, 0, NULL) == KML_ERR_ARG);
  rc = kml_create(&p, 0, &h);
  if (kml_device_count() <= 0) {
    CHECK(rc == KML_ERR_CUDA && h == NULL);
    CHECK(strstr(kml_last_error(NULL), "no CPU fallback") != NULL);
    printf("abi_client: no device, create refused: %s\n", kml_last_error(NULL));
    return 0;
  }
  CHECK(rc == KML_OK && h != NULL);
  {
    enum { F = 64 };
    static uint8_t da[F * 32], db[F * 32];
    static double bear[F * 3], pts[F * 3];
    static uint32_t iq[F], im[F];
    unsigned s = 12345u;
    for (i = 0; i < F * 32; ++i) { s = s * 1664525u + 1013904223u; da[i] = (uint8_t)(s >> 24); }
    for (i = 0; i < F; ++i)      /* frame b = frame a with its rows reversed */
      memcpy(db + (size_t)(F - 1 - i) * 32, da + (size_t)i * 32, 32);
    for (i = 0; i < F; ++i) { bear[3 * i] = 0; bear[3 * i + 1] = 0; bear[3 * i + 2] = 1; pts[3 * i] = 0; pts[3 * i + 1] = 0; pts[3 * i + 2] = 1; }
    CHECK(kml_add_frame(h, 0, 1, da, bear, pts, F) == KML_OK);
    CHECK(kml_add_frame(h, 1, 2, db, bear, pts, F) == KML_OK);
    CHECK(kml_frame_exists(h, 0, 1) == 1 && kml_frame_exists(h, 0, 2) == 0);
    CHECK(kml_compute_matched_indices(h, 0, 1, 1, 2, iq, im, F, &n) == KML_OK);
    CHECK(n == F);
    for (i = 0; i < n; ++i) CHECK(im[i] == (uint32_t)(F - 1) - iq[i]);
    printf("abi_client: %d matches through the CUDA matcher\n", n);
  }
  CHECK(kml_destroy(h) == KML_OK);
  return 0;
}
