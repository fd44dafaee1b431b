/* TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
 *
 * Minimal stand-in for the subset of the FFTW3 single-precision API that the
 * reference (V-Sense/LFBM5D) calls: batched, contiguous r2r plans of rank 1 or
 * 2 with kinds REDFT10 (DCT-II) and REDFT01 (DCT-III), unnormalised, exactly as
 * the FFTW manual defines them:
 *     REDFT10:  Y_k = 2 * sum_{j=0}^{n-1} X_j cos(pi (j+1/2) k / n)
 *     REDFT01:  Y_k = X_0 + 2 * sum_{j=1}^{n-1} X_j cos(pi j (k+1/2) / n)
 * FFTW itself (libfftw3f, unpinned in the reference's CMakeLists.txt:22; the
 * shipped Windows DLL is fftw-3.3.5) is not installed in this image, so the
 * reference is linked against this file instead. Call sites in the reference:
 * utilities.cpp:748-817 (plan creation), bm3d.cpp:745,794,1061,
 * bm5d_core_processing.cpp:1791,1889,1941,2022,2088,2186,2238,... (execution).
 *
 * Two arithmetic modes (lfbm5d_shim_set_dct_mode):
 *   0 = "f32-ordered": every 1-D pass is acc = fmaf(x_j, c_kj, acc), j ascending,
 *       cosine table rounded to float. This is the arithmetic the CUDA kernels use,
 *       so a bit-level comparison is meaningful.
 *   1 = "f64": both passes of a transform accumulate in double with double
 *       cosines and round once at the end (a close-to-correctly-rounded DCT).
 */
#ifndef LFBM5D_ORACLE_FFTW3_SHIM_H
#define LFBM5D_ORACLE_FFTW3_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum { FFTW_R2HC = 0, FFTW_HC2R = 1, FFTW_DHT = 2, FFTW_REDFT00 = 3,
               FFTW_REDFT01 = 4, FFTW_REDFT10 = 5, FFTW_REDFT11 = 6 } fftwf_r2r_kind;
#define FFTW_ESTIMATE (1U << 6)

typedef struct lfbm5d_shim_plan_s *fftwf_plan;

fftwf_plan fftwf_plan_many_r2r(int rank, const int *n, int howmany,
                               float *in, const int *inembed, int istride, int idist,
                               float *out, const int *onembed, int ostride, int odist,
                               const fftwf_r2r_kind *kind, unsigned flags);
void  fftwf_execute_r2r(const fftwf_plan p, float *in, float *out);
void  fftwf_destroy_plan(fftwf_plan p);
void  fftwf_cleanup(void);
void *fftwf_malloc(size_t n);
void  fftwf_free(void *p);

void lfbm5d_shim_set_dct_mode(int mode);
int  lfbm5d_shim_get_dct_mode(void);

#ifdef __cplusplus
}
#endif
#endif
