// Step-2 (Wiener) group kernel specialised for k = 8, 2-D DCT, 3x3 angular window, Haar along the similar patches:
// configs 1-3 and 5 of BASELINE.json. Same arithmetic as k_groups<2,3> (core:1054-1329) with two changes of mechanics:
//   * the noisy patch X and the basic-estimate patch E go through identical forward transforms, so they travel as the two
//     halves of one packed FP32x2 value (FFMA2 / FADD2: two independent IEEE operations per instruction);
//   * after the shrinkage only one signal is left; the inverse transforms pack two rows (p, p+4) of a patch instead.
// One CTA of 288 threads (the item counts 1152 / 576 of the phases divide by it) per reference patch, one colour channel
// at a time in shared memory. Every product that is followed by an addition is written as fma(a, b, -0) with a run-time -0
// (lf_mul2) so that ptxas cannot contract it.
#pragma once
#include "groups.cuh"

#define W8_RS 10      // packed values per patch row: 8 + 2 (row reads of eight lanes hit distinct banks)
#define W8_PS 88      // packed values per patch (8 rows * 10 + 8; = 16 banks mod 32: two patches per half warp are disjoint)
#define W8_NT 288
#define W8_ZRS 10     // filtered signal: packed (row p, row p+4) values per row pair, 4 row pairs per patch
#define W8_ZPS 40     // packed values per patch of the filtered signal (= 16 banks mod 32 as well)

// out[kk] = sum_j v[j] * T[kk*8 + j], ascending j from +0 (oracle DCT mode 0), on packed pairs
__device__ __forceinline__ void w8_dct8(const lf_f2 (&v)[8], lf_f2 (&o)[8], const float *T)
{
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
        lf_f2 acc = 0ull;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc = lf_fma2(v[j], lf_dup(T[kk * 8 + j]), acc);
        o[kk] = acc;
    }
}

// rare shape-adaptive path: both halves through the scalar routines
__device__ __forceinline__ void w8_sadct(lf_f2 (&v)[9], const GroupShape &sh, bool fwd)
{
    float a[9], b[9];
#pragma unroll
    for (int st = 0; st < 9; ++st) lf_upk(v[st], a[st], b[st]);
    if (fwd) { lf_sadct_fwd(a, sh, 3); lf_sadct_fwd(b, sh, 3); }
    else     { lf_sadct_inv(a, sh, 3); lf_sadct_inv(b, sh, 3); }
#pragma unroll
    for (int st = 0; st < 9; ++st) v[st] = lf_pk(a[st], b[st]);
}

template <int NS> __device__ __forceinline__ void w8_haar_fwd(lf_f2 *v, lf_f2 nz2)
{
    lf_f2 tmp[NS > 1 ? NS : 1];
    const lf_f2 c = lf_dup(LF_SQRT2_INV_F);
#pragma unroll
    for (int N = NS; N >= 2; N >>= 1) {
        const int n = N / 2;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const lf_f2 a = v[2 * i], b = v[2 * i + 1];
            tmp[i] = lf_mul2(lf_add2(a, b), c, nz2);
            tmp[n + i] = lf_mul2(lf_sub2(a, b), c, nz2);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = tmp[i];
    }
}
template <int NS> __device__ __forceinline__ void w8_haar_inv(lf_f2 *v, lf_f2 nz2)
{
    lf_f2 tmp[NS > 1 ? NS : 1];
    const lf_f2 c = lf_dup(LF_SQRT2_INV_F);
#pragma unroll
    for (int n = 1; n < NS; n *= 2) {
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const lf_f2 a = v[i], b = v[n + i];
            tmp[2 * i] = lf_mul2(lf_add2(a, b), c, nz2);
            tmp[2 * i + 1] = lf_mul2(lf_sub2(a, b), c, nz2);
        }
#pragma unroll
        for (int i = 0; i < 2 * n; ++i) v[i] = tmp[i];
    }
}

// position of coefficient index i (0..63) inside a packed patch: consecutive lanes of a half warp take rows (p, p+4),
// which are 16 banks apart with W8_RS = 10
__device__ __forceinline__ void w8_pos(int i, int &p, int &q)
{
    q = i & 7;
    p = ((i >> 3) & 1) * 4 + (i >> 4);
}

// 5th dimension + Wiener shrinkage for the two (st, coefficient) vectors of a thread (core:2706-2925).
// Returns the thread's part of weight_table[c]; out[n] = (filtered item 0, filtered item 1) after the inverse Haar.
template <int NS>
__device__ __forceinline__ float w8_filter(const lf_f2 *P, int pos0, int pos1, bool shrink0, bool shrink1, float s2, lf_f2 nz2, lf_f2 (&out)[16])
{
    float wsum = 0.f;
    float r0[NS], r1[NS];
#pragma unroll
    for (int item = 0; item < 2; ++item) {
        lf_f2 v[NS];
        const lf_f2 *src = P + (item ? pos1 : pos0);
#pragma unroll
        for (int n = 0; n < NS; ++n) v[n] = src[n * 9 * W8_PS];
        if (NS > 1) w8_haar_fwd<NS>(v, nz2);
        const bool shrink = item ? shrink1 : shrink0;
#pragma unroll
        for (int n = 0; n < NS; ++n) {
            float x, e;
            lf_upk(v[n], x, e);
            if (shrink) {
                float value = e * e;
                value = value / (value + s2);
                e = x * value;
                wsum += value;
            }
            if (item) r1[n] = e; else r0[n] = e;
        }
    }
#pragma unroll
    for (int n = 0; n < NS; ++n) out[n] = lf_pk(r0[n], r1[n]);
    if (NS > 1) w8_haar_inv<NS>(out, nz2);
    return wsum;
}

template <int CC>
__global__ void __launch_bounds__(W8_NT, 2) k_groups_w8(GroupArgs g, unsigned long long nz2)
{
    extern __shared__ __align__(16) unsigned char w8_smem[];
    lf_f2 *P = reinterpret_cast<lf_f2 *>(w8_smem);       // packed (X, E) patches [pa][8 rows * W8_RS]; later the filtered signal
    float *Pf = reinterpret_cast<float *>(w8_smem);
    __shared__ GroupShape sh;
    __shared__ float red[W8_NT / 32];
    __shared__ unsigned sofs[16 * 9];
    constexpr int A = 9, K2 = 64;
    const int C = CC ? CC : g.C;
    const int tid = threadIdx.x, lane = tid & 31;
    const int r = blockIdx.x + g.r0;
    const int w = g.w;
    const unsigned plane = (unsigned) g.w * (unsigned) g.h;
    const int nSx = (int) g.bm_count[r];
    const int npatch = nSx * A;

    if (!lf_group_setup<true>(g, r, nSx, sh, sofs, nullptr)) return;
    const bool use_sadct = sh.use_sadct != 0 && g.tau_4D == 6;

    // per-thread constants of the phases
    const int q8 = tid & 7;                                    // row index (row pass) / column index (column pass)
    float cn_col[8];                                           // forward normalisation of column q8 (bm3d.cpp:1160)
#pragma unroll
    for (int j = 0; j < 8; ++j) cn_col[j] = c_tab.cn2[j * 8 + q8];
    int gp, gq;                                                // gather: thread <-> pixel, patches sub, sub + 4, ...
    gp = (tid & 63) >> 3; gq = tid & 7;
    const int gsub = tid >> 6;
    // the two (st, coefficient) vectors of the thread in the shrinkage phase
    int p0, q0, p1, q1;
    w8_pos(tid & 63, p0, q0);
    w8_pos((tid + W8_NT) & 63, p1, q1);
    const int st0 = tid >> 6, st1 = (tid + W8_NT) >> 6;
    const int pos0 = st0 * W8_PS + p0 * W8_RS + q0, pos1 = st1 * W8_PS + p1 * W8_RS + q1;
    // same items in the layout of the filtered signal: float index ((p & 3) * W8_ZRS + q) * 2 + (p >> 2) inside the patch
    const int zo0 = st0 * (2 * W8_ZPS) + ((p0 & 3) * W8_ZRS + q0) * 2 + (p0 >> 2);
    const int zo1 = st1 * (2 * W8_ZPS) + ((p1 & 3) * W8_ZRS + q1) * 2 + (p1 >> 2);
    const int pr = tid & 3;                                    // inverse 2-D: rows (pr, pr + 4) / columns (2 pr, 2 pr + 1)
    lf_f2 cni_row[8];                                          // inverse pre-scaling of rows pr (low) and pr + 4 (high)
#pragma unroll
    for (int j = 0; j < 8; ++j) cni_row[j] = lf_pk(c_tab.cni2[pr * 8 + j], c_tab.cni2[(pr + 4) * 8 + j]);
    float *zdst = g.zbuf + (size_t) r * g.N * A * C * K2;

    for (int c = 0; c < C; ++c) {
        // ---- gather (core:1062-1083): noisy -> low half, basic -> high half ----
        if (tid < 256) {
            const unsigned tofs = (unsigned) c * plane + (unsigned) (gp * w + gq);
            float *dst = Pf + 2 * (gsub * W8_PS + gp * W8_RS + gq);
            for (int pa = gsub; pa < npatch; pa += 4) {
                const unsigned src = sofs[pa] + tofs;
                lf_cp_async4(dst, g.nsym + src);
                lf_cp_async4(dst + 1, g.bsym + src);
                dst += 2 * 4 * W8_PS;
            }
        }
        lf_cp_async_wait_all();
        __syncthreads();
        // ---- 2-D DCT (rows, then columns; lib_transforms / bm3d.cpp:1011-1050): a patch stays inside eight lanes ----
        for (int base = 0; base < npatch * 8; base += W8_NT) {
            const int it = base + tid;
            const bool act = it < npatch * 8;
            lf_f2 *pb = P + (it >> 3) * W8_PS;
            if (act) {
                lf_f2 v[8], o[8];
                ulonglong2 *row = reinterpret_cast<ulonglong2 *>(pb + q8 * W8_RS);
#pragma unroll
                for (int j = 0; j < 4; ++j) { const ulonglong2 t = row[j]; v[2 * j] = t.x; v[2 * j + 1] = t.y; }
                w8_dct8(v, o, c_tab.dct2f);
#pragma unroll
                for (int j = 0; j < 4; ++j) row[j] = make_ulonglong2(o[2 * j], o[2 * j + 1]);
            }
            __syncwarp();
            if (act) {
                lf_f2 v[8], o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = pb[j * W8_RS + q8];
                w8_dct8(v, o, c_tab.dct2f);
#pragma unroll
                for (int j = 0; j < 8; ++j) pb[j * W8_RS + q8] = lf_mul2(o[j], lf_dup(cn_col[j]), nz2);
            }
        }
        __syncthreads();
        // ---- angular transform (core:1096-1110) ----
        if (g.tau_4D != 4) {
            for (int it = tid; it < nSx * 64; it += W8_NT) {
                int p, q;
                w8_pos(it & 63, p, q);
                lf_f2 *b = P + (it >> 6) * 9 * W8_PS + p * W8_RS + q;
                lf_f2 v[9];
#pragma unroll
                for (int st = 0; st < 9; ++st) v[st] = b[st * W8_PS];
                if (use_sadct) w8_sadct(v, sh, true); else w8_dct4_fwd(v, nz2);
#pragma unroll
                for (int st = 0; st < 9; ++st) b[st * W8_PS] = v[st];
            }
            __syncthreads();
        }
        // ---- Haar along the similar patches + Wiener shrinkage + inverse Haar (core:1170-1210, :2706-2925) ----
        lf_f2 out[16];
        float wpart;
        {
            const bool sh0 = !use_sadct || sh.mask_dct[st0], sh1 = !use_sadct || sh.mask_dct[st1];
            const float s2 = c_tab.sigma2[c];
            switch (nSx) {
                case 1:  wpart = w8_filter<1>(P, pos0, pos1, sh0, sh1, s2, nz2, out); break;
                case 2:  wpart = w8_filter<2>(P, pos0, pos1, sh0, sh1, s2, nz2, out); break;
                case 4:  wpart = w8_filter<4>(P, pos0, pos1, sh0, sh1, s2, nz2, out); break;
                case 8:  wpart = w8_filter<8>(P, pos0, pos1, sh0, sh1, s2, nz2, out); break;
                default: wpart = w8_filter<16>(P, pos0, pos1, sh0, sh1, s2, nz2, out); break;
            }
        }
        for (int o = 16; o > 0; o >>= 1) wpart += __shfl_down_sync(0xffffffffu, wpart, o);
        if (lane == 0) red[tid >> 5] = wpart;
        __syncthreads();                 // every packed value has been read: the patch memory can take the filtered signal
#pragma unroll
        for (int n = 0; n < 16; ++n)
            if (n < nSx) {
                float a, b;
                lf_upk(out[n], a, b);
                Pf[n * 9 * (2 * W8_ZPS) + zo0] = a;
                Pf[n * 9 * (2 * W8_ZPS) + zo1] = b;
            }
        if (tid == 0) {
            float wsum = 0.f;
            for (int i = 0; i < W8_NT / 32; ++i) wsum += red[i];
            const float sg = c_tab.sigma[c];
            g.wbuf[(size_t) r * C + c] = wsum > 0.0f ? (sg > 0.0f ? 1.0f / (c_tab.sigma2[c] * wsum) : 1.0f / wsum) : 1.0f;   // core:1218-1219
        }
        __syncthreads();
        // ---- inverse angular transform (core:1231-1250), rows (p, p+4) of a patch packed ----
        if (g.tau_4D != 4) {
            // a half warp = one row pair of two consecutive n (their patches are 16 banks apart)
            for (int it = tid; it < ((nSx + 1) & ~1) * 32; it += W8_NT) {
                const int n = 2 * (it >> 6) + ((lane >> 3) & 1), prd = ((it >> 5) & 1) * 2 + (lane >> 4);
                if (n < nSx) {
                    lf_f2 *b = P + n * 9 * W8_ZPS + prd * W8_ZRS + (lane & 7);
                    lf_f2 v[9];
#pragma unroll
                    for (int st = 0; st < 9; ++st) v[st] = b[st * W8_ZPS];
                    if (use_sadct) w8_sadct(v, sh, false); else w8_dct4_inv(v, nz2);
#pragma unroll
                    for (int st = 0; st < 9; ++st) b[st * W8_ZPS] = v[st];
                }
            }
            __syncthreads();
        }
        // ---- inverse 2-D DCT (rows, then columns) and staging of the filtered patches for k_aggregate ----
        for (int base = 0; base < npatch * 4; base += W8_NT) {
            const int it = base + tid;
            const bool act = it < npatch * 4;
            const int pa = it >> 2;
            lf_f2 *pb = P + pa * W8_ZPS;
            if (act) {
                lf_f2 v[8], o[8];
                ulonglong2 *row = reinterpret_cast<ulonglong2 *>(pb + pr * W8_ZRS);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const ulonglong2 t = row[j];
                    v[2 * j] = lf_mul2(t.x, cni_row[2 * j], nz2);
                    v[2 * j + 1] = lf_mul2(t.y, cni_row[2 * j + 1], nz2);
                }
                w8_dct8(v, o, c_tab.dct2i);
#pragma unroll
                for (int j = 0; j < 4; ++j) row[j] = make_ulonglong2(o[2 * j], o[2 * j + 1]);
            }
            __syncwarp();
            if (act) {
                lf_f2 v[8], o[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {      // rows j (low halves) and j + 4 (high halves) at columns 2 pr, 2 pr + 1
                    const ulonglong2 t = *reinterpret_cast<const ulonglong2 *>(pb + j * W8_ZRS + 2 * pr);
                    float a0, a4, b0, b4;
                    lf_upk(t.x, a0, a4);
                    lf_upk(t.y, b0, b4);
                    v[j] = lf_pk(a0, b0);
                    v[j + 4] = lf_pk(a4, b4);
                }
                w8_dct8(v, o, c_tab.dct2i);
                float *zp = zdst + ((size_t) pa * C + c) * K2 + 2 * pr;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                    *reinterpret_cast<lf_f2 *>(zp + kk * 8) = lf_mul2(lf_dup(c_tab.coef2inv), o[kk], nz2);
            }
        }
        __syncthreads();
    }
}
