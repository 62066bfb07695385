/* Plain C99 client of the C ABI (include/noisediff_b200.h): proves the boundary needs nothing but a C compiler — no torch, no
 * C++ — and that the argument checks answer with status codes + ndiff_last_error() instead of crashing.  Compiled and run by
 * tests/test_host.py::test_c_abi_from_plain_c (no GPU needed: every call here fails validation before touching CUDA). */
#include "noisediff_b200.h"
#include <stdio.h>
#include <string.h>
int main(void) {
    if (ndiff_abi_version() != NDIFF_ABI_VERSION) return 1;
    ndiff_engine* e = NULL;
    if (ndiff_engine_create(NULL, &e) == 0) return 2;
    if (!ndiff_last_error() || !strlen(ndiff_last_error())) return 3;
    ndiff_config cfg = {72, 1, 32, 32, 0, 0};           /* widths above the kernels' 64-channel base are refused */
    if (ndiff_engine_create(&cfg, &e) == 0) return 4;
    printf("%s\n", ndiff_last_error());
    ndiff_trainer* t = NULL;
    cfg.dim = 48;                                       /* runs on the sampling path (embedded), not on the training path */
    if (ndiff_trainer_create(&cfg, &t) == 0) return 6;
    printf("%s\n", ndiff_last_error());
    cfg.dim = 64; cfg.height = 12;
    if (ndiff_engine_create(&cfg, &e) == 0) return 5;
    printf("%s\n", ndiff_last_error());
    ndiff_engine_destroy(NULL);
    return 0;
}
