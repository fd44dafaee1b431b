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


/* PNG files as the command lines read and write them (io_png.c:116-260 read_png_f32: channels of the file, 16 -> 8 bits, 1/2/4-bit
 * samples unpacked, any interlace; io_png.c:560-700 write_png_f32: 8 bits, floor(x + .5) clamped). Planar floats c*W*H + i*W + j.
 * lfio_png_read with out == NULL only reports the size. Return 0 on success, 1 on error. */
int lfio_png_read(const char *name, float *out, size_t capacity, size_t *w, size_t *h, size_t *c);
int lfio_png_write(const char *name, const float *data, size_t w, size_t h, size_t c);
/* load_LF / save_LF (utilities_LF.cpp:72-231): <dir>/<sub><sep>%02d<sep>%02d.png for s in [s_start, s_start + aheight), t likewise; the
 * light field as [asize][c*W*H] floats with st ordered per ang_major (LFBM5D_ROWMAJOR / LFBM5D_COLMAJOR); mask[st] = 1 where an image
 * holds a non-zero sample; a gray image stored as RGB counts one channel. Files are decoded / encoded on the host cores
 * (LFBM5D_IO_THREADS). lfio_load_LF with out == NULL only reports width / height / chnls of the first image and the mask. */
int lfio_load_LF(const char *dir, const char *sub, const char *sep, unsigned ang_major, unsigned awidth, unsigned aheight, unsigned s_start,
                 unsigned t_start, float *out, size_t capacity, unsigned *mask, unsigned *width, unsigned *height, unsigned *chnls);
int lfio_save_LF(const char *dir, const char *sub, const char *sep, const float *lf, const unsigned *mask, unsigned ang_major, unsigned awidth,
                 unsigned aheight, unsigned s_start, unsigned t_start, unsigned width, unsigned height, unsigned chnls);
/* compute_psnr_LF (utilities_LF.cpp:639-700), compute_diff_LF (:702-745) and write_psnr_LF (:782-869) on light fields stored as
 * [asize][each] floats; stats4 = (avg psnr, std psnr, avg rmse, std rmse). The report is appended to file_name. */
int lfio_psnr_LF(const float *lf1, const float *lf2, const unsigned *mask, unsigned asize, size_t each, float *psnr, float *rmse, float *stats4);
int lfio_diff_LF(const float *lf1, const float *lf2, const unsigned *mask, unsigned asize, size_t each, float sigma, float *diff);
int lfio_write_psnr_LF(const char *file_name, const char *LF_name, const unsigned *mask, unsigned ang_major, unsigned awidth, unsigned aheight,
                       const float *psnr, float avg_psnr, float std_psnr, const float *rmse, float avg_rmse, float std_rmse);

#ifdef __cplusplus
}
#endif
#endif
