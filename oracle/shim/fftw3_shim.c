/* TEST INFRASTRUCTURE ONLY (oracle). See fftw3.h in this directory. */
#include "fftw3.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

struct lfbm5d_shim_plan_s {
    int rank, n0, n1, howmany, idist, odist;
    fftwf_r2r_kind kind;
    float  *tf0, *tf1;   /* float tables, [k*n + j] */
    double *td0, *td1;   /* double tables */
};

static int g_mode = 0;
void lfbm5d_shim_set_dct_mode(int mode) { g_mode = mode; }
int  lfbm5d_shim_get_dct_mode(void) { return g_mode; }

static void make_tables(int n, fftwf_r2r_kind kind, float **tf, double **td)
{
    *tf = (float *) malloc(sizeof(float) * n * n);
    *td = (double *) malloc(sizeof(double) * n * n);
    for (int k = 0; k < n; k++)
        for (int j = 0; j < n; j++) {
            double v;
            if (kind == FFTW_REDFT10)
                v = 2.0 * cos(M_PI * ((double) j + 0.5) * (double) k / (double) n);
            else
                v = (j == 0) ? 1.0 : 2.0 * cos(M_PI * (double) j * ((double) k + 0.5) / (double) n);
            (*td)[k * n + j] = v;
            (*tf)[k * n + j] = (float) v;
        }
}

fftwf_plan fftwf_plan_many_r2r(int rank, const int *n, int howmany,
                               float *in, const int *inembed, int istride, int idist,
                               float *out, const int *onembed, int ostride, int odist,
                               const fftwf_r2r_kind *kind, unsigned flags)
{
    (void) in; (void) out; (void) inembed; (void) onembed; (void) flags;
    if (rank < 1 || rank > 2 || istride != 1 || ostride != 1) return NULL;
    if (kind[0] != FFTW_REDFT10 && kind[0] != FFTW_REDFT01) return NULL;
    struct lfbm5d_shim_plan_s *p = (struct lfbm5d_shim_plan_s *) calloc(1, sizeof(*p));
    p->rank = rank; p->howmany = howmany; p->idist = idist; p->odist = odist;
    p->kind = kind[0];
    p->n0 = n[0]; p->n1 = rank == 2 ? n[1] : 1;
    make_tables(p->n0, p->kind, &p->tf0, &p->td0);
    if (rank == 2) make_tables(p->n1, p->kind, &p->tf1, &p->td1);
    return p;
}

void fftwf_execute_r2r(const fftwf_plan p, float *in, float *out)
{
    const int n0 = p->n0, n1 = p->n1;
    if (p->rank == 1) {
        for (int b = 0; b < p->howmany; b++) {
            const float *x = in + (size_t) b * p->idist;
            float *y = out + (size_t) b * p->odist;
            float tmp[64];
            for (int k = 0; k < n0; k++) {
                if (g_mode == 0) {
                    float acc = 0.0f;
                    for (int j = 0; j < n0; j++) acc = fmaf(x[j], p->tf0[k * n0 + j], acc);
                    tmp[k] = acc;
                } else {
                    double acc = 0.0;
                    for (int j = 0; j < n0; j++) acc += (double) x[j] * p->td0[k * n0 + j];
                    tmp[k] = (float) acc;
                }
            }
            for (int k = 0; k < n0; k++) y[k] = tmp[k];
        }
        return;
    }
    /* rank 2: n0 rows x n1 columns, row-major; pass 1 along the contiguous
       dimension (length n1), pass 2 along the strided dimension (length n0). */
    float  *tf = (float *) malloc(sizeof(float) * n0 * n1);
    double *td = (double *) malloc(sizeof(double) * n0 * n1);
    for (int b = 0; b < p->howmany; b++) {
        const float *x = in + (size_t) b * p->idist;
        float *y = out + (size_t) b * p->odist;
        if (g_mode == 0) {
            for (int r = 0; r < n0; r++)
                for (int k = 0; k < n1; k++) {
                    float acc = 0.0f;
                    for (int j = 0; j < n1; j++) acc = fmaf(x[r * n1 + j], p->tf1[k * n1 + j], acc);
                    tf[r * n1 + k] = acc;
                }
            for (int c = 0; c < n1; c++)
                for (int k = 0; k < n0; k++) {
                    float acc = 0.0f;
                    for (int j = 0; j < n0; j++) acc = fmaf(tf[j * n1 + c], p->tf0[k * n0 + j], acc);
                    y[k * n1 + c] = acc;
                }
        } else {
            for (int r = 0; r < n0; r++)
                for (int k = 0; k < n1; k++) {
                    double acc = 0.0;
                    for (int j = 0; j < n1; j++) acc += (double) x[r * n1 + j] * p->td1[k * n1 + j];
                    td[r * n1 + k] = acc;
                }
            for (int c = 0; c < n1; c++)
                for (int k = 0; k < n0; k++) {
                    double acc = 0.0;
                    for (int j = 0; j < n0; j++) acc += td[j * n1 + c] * p->td0[k * n0 + j];
                    y[k * n1 + c] = (float) acc;
                }
        }
    }
    free(tf); free(td);
}

void fftwf_destroy_plan(fftwf_plan p)
{
    if (!p) return;
    free(p->tf0); free(p->tf1); free(p->td0); free(p->td1); free(p);
}
void fftwf_cleanup(void) {}
void *fftwf_malloc(size_t n) { return malloc(n ? n : 1); }
void fftwf_free(void *p) { free(p); }
