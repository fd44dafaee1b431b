// Colour transform, mirror padding / cropping, running estimate and the small reductions of the step drivers.
#pragma once
#include "common.cuh"

// utilities.cpp:482-599, evaluated left to right in float (no contraction). One thread per pixel of one SAI.
__device__ __forceinline__ void lf_color_px(unsigned cs, bool fwd, float x, float y, float z, float &o0, float &o1, float &o2)
{
    if (cs == 0) {            // YUV
        if (fwd) {
            o0 = 0.299f * x + 0.587f * y + 0.114f * z;
            o1 = -0.14713f * x - 0.28886f * y + 0.436f * z;
            o2 = 0.615f * x - 0.51498f * y - 0.10001f * z;
        } else {
            o0 = x + 1.13983f * z;
            o1 = x - 0.39465f * y - 0.5806f * z;
            o2 = x + 2.03211f * y;
        }
    } else if (cs == 1) {     // YCbCr
        if (fwd) {
            o0 = 0.299f * x + 0.587f * y + 0.114f * z;
            o1 = -0.169f * x - 0.331f * y + 0.500f * z;
            o2 = 0.500f * x - 0.419f * y - 0.081f * z;
        } else {
            o0 = 1.000f * x + 0.000f * y + 1.402f * z;
            o1 = 1.000f * x - 0.344f * y - 0.714f * z;
            o2 = 1.000f * x + 1.772f * y + 0.000f * z;
        }
    } else {                  // OPP
        if (fwd) {
            o0 = 0.333f * x + 0.333f * y + 0.333f * z;
            o1 = 0.500f * x + 0.000f * y - 0.500f * z;
            o2 = 0.250f * x - 0.500f * y + 0.250f * z;
        } else {
            o0 = 1.0f * x + 1.0f * y + 0.666f * z;
            o1 = 1.0f * x + 0.0f * y - 1.333f * z;
            o2 = 1.0f * x - 1.0f * y + 0.666f * z;
        }
    }
}

// In-place colour transform of every non-masked SAI of a [nsai][3][HW] light field.
__global__ void k_color(float *lf, const unsigned *mask, unsigned nsai, size_t HW, unsigned cs, int fwd)
{
    const size_t total = (size_t) nsai * HW;
    for (size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t) gridDim.x * blockDim.x) {
        const unsigned st = (unsigned) (t / HW);
        if (!mask[st]) continue;
        const size_t px = t - (size_t) st * HW;
        float *b = lf + (size_t) st * 3 * HW + px;
        float o0, o1, o2;
        lf_color_px(cs, fwd != 0, b[0], b[HW], b[2 * HW], o0, o1, o2);
        b[0] = o0; b[HW] = o1; b[2 * HW] = o2;
    }
}

// rt = inverse(forward(x)): what the reference leaves in LF_noisy / LF_basic after a step (bm5d.cpp:133, 711-714; not the
// identity for OPP / YUV / YCbCr). Computed up front so that the host entry points can return it while the passes run.
__global__ void k_roundtrip(const float *__restrict__ lf, float *__restrict__ rt, const unsigned *mask, unsigned nsai, size_t HW, unsigned cs)
{
    const size_t total = (size_t) nsai * HW;
    for (size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t) gridDim.x * blockDim.x) {
        const unsigned st = (unsigned) (t / HW);
        if (!mask[st]) continue;
        const size_t px = t - (size_t) st * HW;
        const float *b = lf + (size_t) st * 3 * HW + px;
        float f0, f1, f2, o0, o1, o2;
        lf_color_px(cs, true, b[0], b[HW], b[2 * HW], f0, f1, f2);
        lf_color_px(cs, false, f0, f1, f2, o0, o1, o2);
        float *r = rt + (size_t) st * 3 * HW + px;
        r[0] = o0; r[HW] = o1; r[2 * HW] = o2;
    }
}

__device__ __forceinline__ int lf_mirror(int v, int n)   // utilities.cpp:215-263 (edge pixel repeated)
{
    return v < 0 ? -v - 1 : (v >= n ? 2 * n - 1 - v : v);
}

// Build the padded working set of one angular window (bm5d.cpp:254-265) and the channel-0 running estimate
// block matching reads (utilities_LF.cpp:913-954 via core:169 / :937). sub = noisy (step 1) or basic (step 2).
// Outputs: nsym/bsym/numsym/densym [A][C][hb][wb], est0 [A][hb][wb]. bsym/basic may be null.
__global__ void k_pad_window(const float *__restrict__ noisy, const float *__restrict__ basic, const float *__restrict__ num,
                             const float *__restrict__ den, float *__restrict__ nsym, float *__restrict__ bsym,
                             float *__restrict__ numsym, float *__restrict__ densym, float *__restrict__ est0,
                             LfWindow win, int W, int H, int C, int n)
{
    const int wb = W + 2 * n, hb = H + 2 * n;
    const size_t plane_b = (size_t) wb * hb, plane = (size_t) W * H;
    const size_t total = (size_t) win.A * plane_b;
    for (size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t) gridDim.x * blockDim.x) {
        const int a = (int) (t / plane_b);
        if (!win.mask[a]) continue;
        const size_t r = t - (size_t) a * plane_b;
        const int i = (int) (r / wb), j = (int) (r - (size_t) i * wb);
        const int si = lf_mirror(i - n, H), sj = lf_mirror(j - n, W);
        const size_t src = (size_t) win.st[a] * C * plane + (size_t) si * W + sj;
        const size_t dst = (size_t) a * C * plane_b + r;
        for (int c = 0; c < C; c++) {
            const float nv = noisy[src + c * plane], uv = num[src + c * plane], dv = den[src + c * plane];
            nsym[dst + c * plane_b] = nv;
            numsym[dst + c * plane_b] = uv;
            densym[dst + c * plane_b] = dv;
            float bv = 0.f;
            if (basic) { bv = basic[src + c * plane]; bsym[dst + c * plane_b] = bv; }
            if (c == 0) est0[(size_t) a * plane_b + r] = dv ? uv / dv : (basic ? bv : nv);
        }
    }
}

// Same estimate for already padded host-provided buffers (debug single-pass entry).
__global__ void k_est0(const float *__restrict__ sub, const float *__restrict__ numsym, const float *__restrict__ densym,
                       float *__restrict__ est0, LfWindow win, size_t plane_b, int C)
{
    const size_t total = (size_t) win.A * plane_b;
    for (size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t) gridDim.x * blockDim.x) {
        const int a = (int) (t / plane_b);
        if (!win.mask[a]) continue;
        const size_t r = t - (size_t) a * plane_b, o = (size_t) a * C * plane_b + r;
        const float dv = densym[o];
        est0[t] = dv ? numsym[o] / dv : sub[o];
    }
}

// Crop the accumulators back (bm5d.cpp:388-396).
__device__ __forceinline__ unsigned long long lf_block_sum_u64(unsigned long long v);

// Crop the padded accumulators of the window back into the light field (bm5d.cpp:388-396). With `count` it also counts the
// entries with den > 0 in the top-left (H-k+1) x (W-k+1) of every SAI (LF_denoised_percent, utilities_LF.cpp:967-995): the
// step driver runs it after every core call, so the coverage test costs no extra pass over the accumulators.
__global__ void k_unpad_window(float *__restrict__ num, float *__restrict__ den, const float *__restrict__ numsym,
                               const float *__restrict__ densym, LfWindow win, int W, int H, int C, int n, int k = 0,
                               unsigned long long *count = nullptr)
{
    unsigned long long cnt = 0;
    const int wb = W + 2 * n, hb = H + 2 * n;
    const size_t plane_b = (size_t) wb * hb, plane = (size_t) W * H;
    const size_t total = (size_t) win.A * C * plane;
    for (size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t) gridDim.x * blockDim.x) {
        const int a = (int) (t / (C * plane));
        if (!win.mask[a]) continue;
        const size_t r = t - (size_t) a * C * plane;
        const int c = (int) (r / plane);
        const size_t px = r - (size_t) c * plane;
        const int i = (int) (px / W), j = (int) (px - (size_t) i * W);
        const size_t src = ((size_t) a * C + c) * plane_b + (size_t) (i + n) * wb + (j + n);
        const size_t dst = ((size_t) win.st[a] * C + c) * plane + px;
        const float dv = densym[src];
        num[dst] = numsym[src];
        den[dst] = dv;
        if (count && i < H - k + 1 && j < W - k + 1 && dv > 0.0f) cnt++;
    }
    if (count) {
        cnt = lf_block_sum_u64(cnt);
        if (threadIdx.x == 0 && cnt) atomicAdd(count, cnt);
    }
}

__device__ __forceinline__ unsigned long long lf_block_sum_u64(unsigned long long v)
{
    __shared__ unsigned long long sh[32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = threadIdx.x < (blockDim.x + 31) / 32 ? sh[threadIdx.x] : 0ull;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    }
    return v;
}

// Number of entries equal to 0.0 in the accumulator of one SAI (bm5d.cpp:195).
__global__ void k_count_zero(const float *__restrict__ den, size_t nelem, unsigned long long *out)
{
    unsigned long long cnt = 0;
    for (size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x; t < nelem; t += (size_t) gridDim.x * blockDim.x)
        if (den[t] == 0.0f) cnt++;
    cnt = lf_block_sum_u64(cnt);
    if (threadIdx.x == 0 && cnt) atomicAdd(out, cnt);
}

// Final estimate (bm5d.cpp:405 / :706) fused with the inverse colour transforms of the outputs (bm5d.cpp:711-714,
// 1414-1419): out = den ? num/den : sub, then out, noisy (and basic) go back to RGB. C == 3 path handles colour;
// for C == 1 or RGB colour_space `docolor` is 0.
__global__ void k_final(const float *__restrict__ num, const float *__restrict__ den, float *noisy, float *basic, float *out,
                        const unsigned *mask, unsigned nsai, size_t HW, int C, int step, unsigned cs, int docolor, int inputs_back = 1)
{
    // inputs_back = 0: noisy / basic stay in the working colour space (host entry points: their round-tripped copies went back early)
    const size_t total = (size_t) nsai * HW;
    for (size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t) gridDim.x * blockDim.x) {
        const unsigned st = (unsigned) (t / HW);
        if (!mask[st]) continue;
        const size_t px = t - (size_t) st * HW, base = (size_t) st * C * HW + px;
        float e[3], nz[3], bs[3];
        for (int c = 0; c < C; c++) {
            const float nv = noisy[base + c * HW], dv = den[base + c * HW];
            nz[c] = nv;
            bs[c] = step == 2 ? basic[base + c * HW] : 0.f;
            e[c] = dv ? num[base + c * HW] / dv : (step == 2 ? bs[c] : nv);
        }
        if (docolor) {
            float a, b, c2;
            lf_color_px(cs, false, e[0], e[1], e[2], a, b, c2); e[0] = a; e[1] = b; e[2] = c2;
            if (inputs_back) {
                lf_color_px(cs, false, nz[0], nz[1], nz[2], a, b, c2); nz[0] = a; nz[1] = b; nz[2] = c2;
                if (step == 2) { lf_color_px(cs, false, bs[0], bs[1], bs[2], a, b, c2); bs[0] = a; bs[1] = b; bs[2] = c2; }
            }
        }
        for (int c = 0; c < C; c++) {
            out[base + c * HW] = e[c];
            if (docolor && inputs_back) {
                noisy[base + c * HW] = nz[c];
                if (step == 2) basic[base + c * HW] = bs[c];
            }
        }
    }
}
