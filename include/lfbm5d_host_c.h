/* lfbm5d_host_c.h — C exports of liblfbm5d_host.so beside the C++ adapters (lfbm5d_b200/csrc/lfbm5d_host.h):
 * the host-side noise generator and PSNR of the reference's command lines, for callers without a C++ toolchain
 * (bench.py generates its noisy input with them, as north_star prescribes: reference mt19937ar on the host).
 */
#ifndef LFBM5D_HOST_C_H
#define LFBM5D_HOST_C_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* add_noise (utilities.cpp:154-185) on one image of n samples with its own mt19937ar generator seeded `seed`
 * (mt19937ar.c init_genrand / genrand_res53; the reference seeds from time + pid, utilities.cpp:165-175):
 * out[k] = img[k] + (float)(sigma * sqrt(-2 log a) * cos(2 pi b)), a, b = consecutive genrand_res53(). Unclipped. */
void lfio_add_noise(const float *img, float *out, size_t n, float sigma, unsigned long seed);
/* compute_psnr (utilities.cpp:412-435): float accumulator, psnr = 20 log10(255 / rmse) */
void lfio_psnr(const float *a, const float *b, size_t n, float *psnr, float *rmse);

#ifdef __cplusplus
}
#endif
#endif
