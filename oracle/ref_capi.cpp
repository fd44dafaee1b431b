/* TEST INFRASTRUCTURE ONLY (oracle).
 *
 * Thin extern "C" wrapper that exposes functions of the UNMODIFIED reference
 * (compiled from /root/reference/src where it lies, see oracle/Makefile) to
 * ctypes, so that tests can (a) pin oracle/lfbm5d_oracle.c against the real
 * reference and (b) generate golden fixtures. Nothing in the product path links
 * or loads this file. Outputs go to oracle/_ref/ only.
 *
 * Every wrapper only marshals flat arrays into the std::vector arguments the
 * reference declares (bm5d.h:11-62, bm5d_core_processing.h:6-80, :377-400,
 * bm3d.h:11-34, bm3d_LF.h:10-35, utilities.h, utilities_LF.h, lib_transforms.h).
 */
#include <vector>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <utility>
#include <iostream>
#include <streambuf>
#include <fftw3.h>
#include <omp.h>

#include "bm5d.h"
#include "bm3d.h"
#include "bm3d_LF.h"
#include "bm5d_core_processing.h"
#include "utilities.h"
#include "utilities_LF.h"
#include "lib_transforms.h"
#include "mt19937ar.h"

using std::vector;

namespace {
/* The reference prints progress on cout; silence it unless asked. */
struct NullBuf : std::streambuf { int overflow(int c) override { return c; } };
/* The reference's SAI-selection loops race on shared variables under OpenMP (bm5d.cpp:190-202, 322-334), so
 * its window schedule is only deterministic with one OpenMP thread: force that unless told otherwise. */
struct CoutSilencer {
    std::streambuf *old; NullBuf nb;
    CoutSilencer() : old(nullptr) {
        if (!getenv("LFBM5D_REF_VERBOSE")) old = std::cout.rdbuf(&nb);
        const char *t = getenv("LFBM5D_REF_OMP_THREADS");
        omp_set_num_threads(t ? atoi(t) : 1);
    }
    ~CoutSilencer() { if (old) std::cout.rdbuf(old); }
};

vector<vector<float> > to_vv(const float *p, unsigned n, size_t each)
{
    vector<vector<float> > v(n);
    for (unsigned i = 0; i < n; i++) v[i].assign(p + (size_t) i * each, p + (size_t) (i + 1) * each);
    return v;
}
void from_vv(const vector<vector<float> > &v, float *p, size_t each)
{
    for (size_t i = 0; i < v.size(); i++)
        if (v[i].size() == each) memcpy(p + i * each, v[i].data(), each * sizeof(float));
}

const unsigned kDCT = 5, kSADCT = 6;

struct Plans {
    fftwf_plan p2_1 = nullptr, p2_2 = nullptr, p2_3 = nullptr, p2_inv = nullptr, p4 = nullptr, p4_inv = nullptr, p5 = nullptr, p5_inv = nullptr;
    vector<fftwf_plan> sa, sa_inv;
};
/* Plan set-up as the step drivers do it (bm5d.cpp:283-308 / :982-1007). */
void make_plans(Plans &P, unsigned tau_2D, unsigned tau_4D, unsigned asw, unsigned w_b, unsigned n, unsigned k, unsigned N, unsigned p, unsigned chnls)
{
    if (tau_2D == kDCT) {
        const unsigned nb_cols = ind_size(w_b - k + 1, n, p);
        allocate_plan_2d(&P.p2_1, k, FFTW_REDFT10, w_b * (2 * n + 1) * chnls);
        allocate_plan_2d(&P.p2_2, k, FFTW_REDFT10, w_b * p * chnls);
        allocate_plan_2d(&P.p2_3, k, FFTW_REDFT10, (2 * n + 1) * (2 * n + 1) * chnls);
        allocate_plan_2d(&P.p2_inv, k, FFTW_REDFT01, N * nb_cols * chnls);
    }
    if (tau_4D == kDCT || tau_4D == kSADCT) {
        allocate_plan_2d(&P.p4, asw, asw, FFTW_REDFT10, k * k * chnls);
        allocate_plan_2d(&P.p4_inv, asw, asw, FFTW_REDFT01, k * k * chnls);
    }
    P.sa.assign(asw > 1 ? asw - 1 : 1, nullptr);
    P.sa_inv.assign(asw > 1 ? asw - 1 : 1, nullptr);
    if (tau_4D == kSADCT)
        for (unsigned i = 0; i + 1 < asw; i++) {
            allocate_plan_1d(&P.sa[i], i + 2, FFTW_REDFT10, k * k * chnls);
            allocate_plan_1d(&P.sa_inv[i], i + 2, FFTW_REDFT01, k * k * chnls);
        }
}
void free_plans(Plans &P)
{
    fftwf_destroy_plan(P.p2_1); fftwf_destroy_plan(P.p2_2); fftwf_destroy_plan(P.p2_3); fftwf_destroy_plan(P.p2_inv);
    fftwf_destroy_plan(P.p4); fftwf_destroy_plan(P.p4_inv);
    for (auto q : P.sa) fftwf_destroy_plan(q);
    for (auto q : P.sa_inv) fftwf_destroy_plan(q);
}
} // namespace

extern "C" {

void ref_set_dct_mode(int mode) { lfbm5d_shim_set_dct_mode(mode); }

/* ---- step drivers (bm5d.h:11-62) ------------------------------------------------ */
int ref_run_bm5d_1st_step(float sigma, float lambdaHard5D, float *noisy_io, const unsigned *mask, float *basic_out,
                          unsigned ang_major, unsigned awidth, unsigned aheight, unsigned anHard,
                          unsigned width, unsigned height, unsigned chnls, unsigned NHard, unsigned nSim, unsigned nDisp,
                          unsigned kHard, unsigned pHard, int useSD, unsigned tau_2D, unsigned tau_4D, unsigned tau_5D,
                          unsigned color_space, unsigned nb_threads)
{
    CoutSilencer s;
    const unsigned asize = awidth * aheight;
    const size_t each = (size_t) width * height * chnls;
    vector<vector<float> > LF_noisy = to_vv(noisy_io, asize, each), LF_basic(asize);
    vector<unsigned> m(mask, mask + asize);
    int rc = run_bm5d_1st_step(sigma, lambdaHard5D, LF_noisy, m, LF_basic, ang_major, awidth, aheight, anHard, width, height, chnls,
                               NHard, nSim, nDisp, kHard, pHard, useSD != 0, tau_2D, tau_4D, tau_5D, color_space, nb_threads);
    from_vv(LF_noisy, noisy_io, each);
    from_vv(LF_basic, basic_out, each);
    return rc;
}

int ref_run_bm5d_2nd_step(float sigma, float *noisy_io, const unsigned *mask, float *basic_io, float *denoised_out,
                          unsigned ang_major, unsigned awidth, unsigned aheight, unsigned anWien,
                          unsigned width, unsigned height, unsigned chnls, unsigned NWien, unsigned nSim, unsigned nDisp,
                          unsigned kWien, unsigned pWien, int useSD, unsigned tau_2D, unsigned tau_4D, unsigned tau_5D,
                          unsigned color_space, unsigned nb_threads)
{
    CoutSilencer s;
    const unsigned asize = awidth * aheight;
    const size_t each = (size_t) width * height * chnls;
    vector<vector<float> > LF_noisy = to_vv(noisy_io, asize, each), LF_basic = to_vv(basic_io, asize, each), LF_den(asize);
    vector<unsigned> m(mask, mask + asize);
    int rc = run_bm5d_2nd_step(sigma, LF_noisy, m, LF_basic, LF_den, ang_major, awidth, aheight, anWien, width, height, chnls,
                               NWien, nSim, nDisp, kWien, pWien, useSD != 0, tau_2D, tau_4D, tau_5D, color_space, nb_threads);
    from_vv(LF_noisy, noisy_io, each);
    from_vv(LF_basic, basic_io, each);
    from_vv(LF_den, denoised_out, each);
    return rc;
}

/* bm3d_LF.h:10-35 */
int ref_run_bm3d_LF(float sigma, float *noisy_io, const unsigned *mask, float *basic_out, float *denoised_out,
                    unsigned asize, unsigned width, unsigned height, unsigned chnls,
                    unsigned nHard, unsigned nWien, unsigned kHard, unsigned kWien, unsigned NHard, unsigned NWien,
                    unsigned pHard, unsigned pWien, int useSD_h, int useSD_w, unsigned tau_2D_hard, unsigned tau_2D_wien,
                    float lambdaHard3D, unsigned color_space, unsigned nb_threads)
{
    CoutSilencer s;
    const size_t each = (size_t) width * height * chnls;
    vector<vector<float> > LF_noisy = to_vv(noisy_io, asize, each), LF_basic(asize), LF_den(asize);
    vector<unsigned> m(mask, mask + asize);
    char name[8] = "sai";
    int rc = run_bm3d_LF(sigma, LF_noisy, m, LF_basic, LF_den, width, height, chnls, nHard, nWien, kHard, kWien, NHard, NWien,
                         pHard, pWien, useSD_h != 0, useSD_w != 0, tau_2D_hard, tau_2D_wien, lambdaHard3D, color_space, nb_threads, name);
    from_vv(LF_noisy, noisy_io, each);
    from_vv(LF_basic, basic_out, each);
    from_vv(LF_den, denoised_out, each);
    return rc;
}

/* ---- one window pass on padded buffers (bm5d_core_processing.h:6-80) -------------- */
int ref_bm5d_1st_step_pass(float sigma, float lambdaHard5D, const float *noisy_sym, float *num_sym_io, float *den_sym_io,
                           const unsigned *mask_asw, const unsigned *procSAI_asw, unsigned cst, unsigned pst,
                           unsigned asw, unsigned w_b, unsigned h_b, unsigned chnls, unsigned nSim, unsigned nDisp,
                           unsigned kHard, unsigned NHard, unsigned pHard, int useSD, unsigned color_space,
                           unsigned tau_2D, unsigned tau_4D, unsigned tau_5D)
{
    CoutSilencer s;
    const unsigned A = asw * asw;
    const size_t each = (size_t) w_b * h_b * chnls;
    vector<vector<float> > N = to_vv(noisy_sym, A, each), num = to_vv(num_sym_io, A, each), den = to_vv(den_sym_io, A, each);
    vector<unsigned> m(mask_asw, mask_asw + A), proc(procSAI_asw, procSAI_asw + A);
    Plans P; make_plans(P, tau_2D, tau_4D, asw, w_b, nSim + nDisp, kHard, NHard, pHard, chnls);
    float bm_secs = 0.f;
    bm5d_1st_step(sigma, lambdaHard5D, N, num, den, m, proc, cst, pst, asw, asw, w_b, h_b, chnls, nSim, nDisp, kHard, NHard, pHard,
                  useSD != 0, color_space, tau_2D, tau_4D, tau_5D, &P.p2_1, &P.p2_2, &P.p2_3, &P.p2_inv, &P.p4, &P.p4_inv,
                  P.sa.data(), P.sa_inv.data(), &P.p5, &P.p5_inv, bm_secs);
    free_plans(P);
    from_vv(num, num_sym_io, each);
    from_vv(den, den_sym_io, each);
    return 0;
}

int ref_bm5d_2nd_step_pass(float sigma, const float *noisy_sym, const float *basic_sym, float *num_sym_io, float *den_sym_io,
                           const unsigned *mask_asw, const unsigned *procSAI_asw, unsigned cst, unsigned pst,
                           unsigned asw, unsigned w_b, unsigned h_b, unsigned chnls, unsigned nSim, unsigned nDisp,
                           unsigned kWien, unsigned NWien, unsigned pWien, int useSD, unsigned color_space,
                           unsigned tau_2D, unsigned tau_4D, unsigned tau_5D)
{
    CoutSilencer s;
    const unsigned A = asw * asw;
    const size_t each = (size_t) w_b * h_b * chnls;
    vector<vector<float> > N = to_vv(noisy_sym, A, each), B = to_vv(basic_sym, A, each),
                           num = to_vv(num_sym_io, A, each), den = to_vv(den_sym_io, A, each);
    vector<unsigned> m(mask_asw, mask_asw + A), proc(procSAI_asw, procSAI_asw + A);
    Plans P; make_plans(P, tau_2D, tau_4D, asw, w_b, nSim + nDisp, kWien, NWien, pWien, chnls);
    float bm_secs = 0.f;
    bm5d_2nd_step(sigma, N, B, num, den, m, proc, cst, pst, asw, asw, w_b, h_b, chnls, nSim, nDisp, kWien, NWien, pWien,
                  useSD != 0, color_space, tau_2D, tau_4D, tau_5D, &P.p2_1, &P.p2_2, &P.p2_3, &P.p2_inv, &P.p4, &P.p4_inv,
                  P.sa.data(), P.sa_inv.data(), &P.p5, &P.p5_inv, bm_secs);
    free_plans(P);
    from_vv(num, num_sym_io, each);
    from_vv(den, den_sym_io, each);
    return 0;
}

/* ---- block matching (bm5d_core_processing.h:377-400) ----------------------------- */
/* out_count[w*h] (0 where the reference leaves the entry empty), out_idx[w*h*maxN] */
void ref_precompute_BM(const float *img, unsigned width, unsigned height, unsigned kHW, unsigned NHW, unsigned nHW,
                       unsigned nHW_sim, unsigned pHW, float tauMatch, unsigned *out_count, unsigned *out_idx, unsigned maxN)
{
    vector<float> im(img, img + (size_t) width * height);
    vector<vector<unsigned> > table;
    precompute_BM(table, im, width, height, kHW, NHW, nHW, nHW_sim, pHW, tauMatch);
    for (size_t k = 0; k < (size_t) width * height; k++) {
        const unsigned c = k < table.size() ? (unsigned) table[k].size() : 0;
        out_count[k] = c;
        for (unsigned n = 0; n < c && n < maxN; n++) out_idx[k * maxN + n] = table[k][n];
    }
}

/* out_first[w*h] = patch_table[k][0] (0xFFFFFFFF where empty), out_shape[w*h]; out_all (optional) [w*h*Ns*Ns] */
int ref_precompute_BM_stereo(const float *img1, const float *img2, unsigned width, unsigned height, unsigned kHW, unsigned nHW,
                             unsigned nHW_disp, unsigned pHW, float tauMatch, unsigned *out_first, unsigned *out_shape, unsigned *out_all)
{
    vector<float> a(img1, img1 + (size_t) width * height), b(img2, img2 + (size_t) width * height);
    vector<vector<unsigned> > table;
    vector<unsigned> shape;
    int rc = precompute_BM_stereo(table, shape, a, b, width, height, kHW, nHW, nHW_disp, pHW, tauMatch);
    if (rc != 0) return rc;
    const unsigned Ns2 = (2 * nHW_disp + 1) * (2 * nHW_disp + 1);
    for (size_t k = 0; k < (size_t) width * height; k++) {
        const bool has = k < table.size() && !table[k].empty();
        out_first[k] = has ? table[k][0] : 0xFFFFFFFFu;
        out_shape[k] = has ? shape[k] : 0;
        if (out_all)
            for (unsigned n = 0; n < Ns2; n++) out_all[k * Ns2 + n] = (has && n < table[k].size()) ? table[k][n] : 0xFFFFFFFFu;
    }
    return 0;
}

/* BM3D flavour (bm3d.h:188) */
void ref_bm3d_precompute_BM(const float *img, unsigned width, unsigned height, unsigned kHW, unsigned NHW, unsigned nHW,
                            unsigned pHW, float tauMatch, unsigned *out_count, unsigned *out_idx, unsigned maxN)
{
    vector<float> im(img, img + (size_t) width * height);
    vector<vector<unsigned> > table;
    precompute_BM(table, im, width, height, kHW, NHW, nHW, pHW, tauMatch);
    for (size_t k = 0; k < (size_t) width * height; k++) {
        const unsigned c = k < table.size() ? (unsigned) table[k].size() : 0;
        out_count[k] = c;
        for (unsigned n = 0; n < c && n < maxN; n++) out_idx[k * maxN + n] = table[k][n];
    }
}

/* ---- small helpers --------------------------------------------------------------- */
int ref_color_space_transform(float *img_io, unsigned color_space, unsigned width, unsigned height, unsigned chnls, int rgb2yuv)
{
    CoutSilencer s;
    vector<float> im(img_io, img_io + (size_t) width * height * chnls);
    int rc = color_space_transform(im, color_space, width, height, chnls, rgb2yuv != 0);
    memcpy(img_io, im.data(), im.size() * sizeof(float));
    return rc;
}
int ref_estimate_sigma(float sigma, float *table, unsigned chnls, unsigned color_space)
{
    vector<float> t(chnls);
    int rc = estimate_sigma(sigma, t, chnls, color_space);
    for (unsigned c = 0; c < chnls; c++) table[c] = t[c];
    return rc;
}
void ref_symetrize(const float *img, float *out, unsigned width, unsigned height, unsigned chnls, unsigned N)
{
    vector<float> im(img, img + (size_t) width * height * chnls), sym;
    symetrize(im, sym, width, height, chnls, N);
    memcpy(out, sym.data(), sym.size() * sizeof(float));
}
unsigned ref_ind_initialize(unsigned *out, unsigned max_size, unsigned N, unsigned step)
{
    vector<unsigned> v;
    ind_initialize(v, max_size, N, step);
    for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
    return (unsigned) v.size();
}
void ref_angular_search_window(int *c_asw, int *min_asw, int *max_asw, unsigned aidx, unsigned asize, unsigned asize_sw)
{
    compute_LF_angular_search_window(*c_asw, *min_asw, *max_asw, aidx, asize, asize_sw);
}
float ref_LF_denoised_percent(const float *den_sym, const unsigned *mask, unsigned A, unsigned width, unsigned height, unsigned chnls, unsigned N, unsigned kHW)
{
    const size_t each = (size_t) (width + 2 * N) * (height + 2 * N) * chnls;
    vector<vector<float> > d = to_vv(den_sym, A, each);
    vector<unsigned> m(mask, mask + A);
    return LF_denoised_percent(d, m, width, height, chnls, N, kHW);
}
void ref_preProcess(float *kaiser, float *coef_norm, float *coef_norm_inv, unsigned kHW)
{
    vector<float> a(kHW * kHW), b(kHW * kHW), c(kHW * kHW);
    preProcess(a, b, c, kHW);
    memcpy(kaiser, a.data(), a.size() * 4); memcpy(coef_norm, b.data(), b.size() * 4); memcpy(coef_norm_inv, c.data(), c.size() * 4);
}
void ref_preProcess_4d(float *coef_norm, float *coef_norm_inv, unsigned awidth, unsigned aheight)
{
    vector<float> a(awidth * aheight), b(awidth * aheight);
    preProcess_4d(a, b, awidth, aheight);
    memcpy(coef_norm, a.data(), a.size() * 4); memcpy(coef_norm_inv, b.data(), b.size() * 4);
}
/* tables for dct sizes 2..max, packed [size-2][max] */
void ref_preProcess_4d_sadct(float *coef_norm, float *coef_norm_inv, unsigned max_dct_size)
{
    vector<vector<float> > a(max_dct_size - 1), b(max_dct_size - 1);
    preProcess_4d_sadct(a, b, max_dct_size);
    for (unsigned k = 0; k + 1 < max_dct_size; k++)
        for (unsigned i = 0; i < k + 2; i++) { coef_norm[k * max_dct_size + i] = a[k][i]; coef_norm_inv[k * max_dct_size + i] = b[k][i]; }
}
void ref_bior15_coef(float *lpd, float *hpd, float *lpr, float *hpr)
{
    vector<float> a, b, c, d;
    bior15_coef(a, b, c, d);
    for (int i = 0; i < 10; i++) { lpd[i] = a[i]; hpd[i] = b[i]; lpr[i] = c[i]; hpr[i] = d[i]; }
}
void ref_bior_2d_forward(const float *patch, float *out, unsigned N)
{
    vector<float> a, b, c, d; bior15_coef(a, b, c, d);
    vector<float> in(patch, patch + N * N), o(N * N);
    bior_2d_forward(in, o, N, 0, N, 0, a, b);
    memcpy(out, o.data(), o.size() * 4);
}
void ref_bior_2d_inverse(float *patch_io, unsigned N)
{
    vector<float> a, b, c, d; bior15_coef(a, b, c, d);
    vector<float> sig(patch_io, patch_io + N * N);
    bior_2d_inverse(sig, N, 0, c, d);
    memcpy(patch_io, sig.data(), sig.size() * 4);
}
void ref_haar_forward(float *v, unsigned N) { vector<float> a(v, v + N), t(N); haar_forward(a, t, N, 0); memcpy(v, a.data(), N * 4); }
void ref_haar_inverse(float *v, unsigned N) { vector<float> a(v, v + N), t(N); haar_inverse(a, t, 1, N, 0); memcpy(v, a.data(), N * 4); }
void ref_hadamard(float *v, unsigned N) { vector<float> a(v, v + N), t(N); hadamard_transform(a, t, N, 0); memcpy(v, a.data(), N * 4); }

/* 2-D DCT of one k x k patch the way the tables are built (bm3d.cpp:745-757) and inverted (bm3d.cpp:1039-1071) */
void ref_dct_2d_patch(const float *patch, float *out, unsigned k)
{
    vector<float> kw(k * k), cn(k * k), cni(k * k);
    preProcess(kw, cn, cni, k);
    fftwf_plan p; allocate_plan_2d(&p, k, FFTW_REDFT10, 1);
    vector<float> in(patch, patch + k * k), o(k * k);
    fftwf_execute_r2r(p, in.data(), o.data());
    for (unsigned i = 0; i < k * k; i++) out[i] = o[i] * cn[i];
    fftwf_destroy_plan(p);
}
void ref_dct_2d_inverse_patch(float *patch_io, unsigned k)
{
    vector<float> kw(k * k), cn(k * k), cni(k * k);
    preProcess(kw, cn, cni, k);
    fftwf_plan p; allocate_plan_2d(&p, k, FFTW_REDFT01, 1);
    vector<float> g(patch_io, patch_io + k * k);
    dct_2d_inverse(g, k, 1, cni, &p);
    memcpy(patch_io, g.data(), g.size() * 4);
    fftwf_destroy_plan(p);
}

/* Deterministic noise: seeds the reference's own mt19937ar and applies the Box-Muller expression of
 * add_noise (utilities.cpp:177-184) through the reference's mt_genrand_res53. */
void ref_mt_seed(unsigned long s) { mt_init_genrand(s); }
double ref_mt_res53(void) { return mt_genrand_res53(); }

int ref_compute_psnr(const float *a, const float *b, size_t n, float *psnr, float *rmse)
{
    vector<float> x(a, a + n), y(b, b + n);
    return compute_psnr(x, y, psnr, rmse);
}

/* compute_psnr_LF / compute_diff_LF / write_psnr_LF (utilities_LF.cpp:639-700, :702-745, :782-869) on [asize][each] arrays */
int ref_compute_psnr_LF(const float *lf1, const float *lf2, const unsigned *mask, unsigned asize, size_t each, float *psnr, float *rmse, float *stats4)
{
    CoutSilencer quiet;
    vector<unsigned> m(mask, mask + asize);
    vector<float> ps, rm;
    const int rc = compute_psnr_LF(to_vv(lf1, asize, each), to_vv(lf2, asize, each), m, ps, &stats4[0], &stats4[1], rm, &stats4[2], &stats4[3]);
    if (rc == EXIT_SUCCESS) { memcpy(psnr, ps.data(), asize * sizeof(float)); memcpy(rmse, rm.data(), asize * sizeof(float)); }
    return rc;
}
int ref_compute_diff_LF(const float *lf1, const float *lf2, const unsigned *mask, unsigned asize, size_t each, float sigma, float *diff)
{
    CoutSilencer quiet;
    vector<unsigned> m(mask, mask + asize);
    vector<vector<float> > d;
    const int rc = compute_diff_LF(to_vv(lf1, asize, each), to_vv(lf2, asize, each), m, d, sigma);
    if (rc == EXIT_SUCCESS) from_vv(d, diff, each);
    return rc;
}
int ref_write_psnr_LF(const char *file_name, const char *LF_name, const unsigned *mask, unsigned ang_major, unsigned awidth, unsigned aheight,
                      const float *psnr, float avg_psnr, float std_psnr, const float *rmse, float avg_rmse, float std_rmse)
{
    CoutSilencer quiet;
    const unsigned asize = awidth * aheight;
    vector<unsigned> m(mask, mask + asize);
    return write_psnr_LF(file_name, LF_name, m, ang_major, awidth, aheight, vector<float>(psnr, psnr + asize), avg_psnr, std_psnr,
                         vector<float>(rmse, rmse + asize), avg_rmse, std_rmse);
}

} // extern "C"
