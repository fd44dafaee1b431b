// Block matching: exact float32 summed-area recurrences (one warp sweeps one offset plane as a skewed
// wavefront), top-N selection for self similarity and argmin for disparity matching, with the libstdc++
// sort algorithms emulated where exact float ties make the reference's result algorithm-dependent.
//
// Reference behaviour restated: precompute_BM (bm5d_core_processing.cpp:3301-3461) and precompute_BM_stereo
// (:3479-3611): per offset, a table of squared differences, then patch sums by the recurrence
//   s[k] = s[k-1] + s[k-w] - s[k-1-w] + d[pq] - d[pq-kHW] - d[pq-kHW*w] + d[pq-kHW-kHW*w]
// evaluated left to right in float, with dedicated formulas for the first patch, first row and first column.
// The rounding of that recurrence decides ~1 % of the match lists, so it is reproduced operation for operation.
#pragma once
#include "common.cuh"

struct SatDesc {
    const float *img1;      // reference image (channel 0 of the running estimate)
    const float *img2;      // image the offset is applied to
    int          dk;        // flat offset: d(y,x) = (img2[y*w+x+dk] - img1[y*w+x])^2
    float       *out_plane; // stereo: full plane of sums [h*w] (written on the computed region only)
    float       *out_at;    // self: sums sampled at the reference patches, [nr*nc]
    float       *out_mir;   // self: sums sampled at (ref - (mir_di, -mir_dc)), [nr*nc] (pre-filled with 2*threshold)
    int          mir_di, mir_dc;
    float       *bnd;       // scratch, 2*h floats: last column of the previous / current 32-column strip
};

struct SatGeom {
    int w, h, k;
    int lo;                 // first row/column of the summed region
    int row_end, col_end;   // one past its last row/column
    int dlo;                // squared differences are non-zero only on [dlo, h-dlo) x [dlo, w-dlo)
    int nc;                 // number of reference-patch columns (self)
    const int *rowmap;      // [h] row -> reference row index or -1 (self)
    const int *colmap;      // [w]
};

// One warp per offset plane. The plane is swept in 32-column strips; inside a strip lane l owns column c0+l and
// runs one row behind lane l-1, so that s(i, j-1) arrives by one shuffle per step. Squared differences are
// produced one image row per step with coalesced loads into a 64-row shared-memory ring whose row stride (64
// floats) makes the skewed reads bank-conflict free; finished rows are transposed through a 32x32 tile so that
// results leave with coalesced 128-byte stores.
template <bool SELF>
__global__ void __launch_bounds__(32) k_sat_planes(SatGeom g, const SatDesc *__restrict__ descs)
{
    __shared__ float dring[64 * 64];
    __shared__ float tile[32 * 32];
    __shared__ float frow[32];
    const SatDesc D = descs[blockIdx.x];
    const int lane = threadIdx.x;
    const int w = g.w, k = g.k, lo = g.lo;
    const int W = g.col_end - lo, Hh = g.row_end - lo;
    const int nstrips = (W + 31) >> 5;
    float *bnd_prev = D.bnd, *bnd_next = D.bnd + g.h;
    const unsigned FULL = 0xffffffffu;

    for (int strip = 0; strip < nstrips; ++strip) {
        const int c0 = lo + (strip << 5);
        const int j = c0 + lane;
        const bool valid = j < g.col_end;
        const int lastlane = min(31, W - 1 - (strip << 5));
        const bool has_next = strip + 1 < nstrips;

        auto load_drow = [&](int y) {
            for (int t = lane; t < 32 + k; t += 32) {
                const int x = c0 - 1 + t;
                float v = 0.f;
                if (y >= g.dlo && y < g.h - g.dlo && x >= g.dlo && x < w - g.dlo) {
                    const float df = D.img2[y * w + x + D.dk] - D.img1[y * w + x];
                    v = df * df;
                }
                dring[(y & 63) * 64 + t] = v;
            }
        };
        auto emit_row = [&](int i, float v) {   // v = s(i, j) of this lane
            if (SELF) {
                if (!valid) return;
                const int a = g.rowmap[i];
                if (a >= 0) {
                    const int b = g.colmap[j];
                    if (b >= 0) D.out_at[a * g.nc + b] = v;
                }
                if (D.mir_di > 0) {
                    const int ir = i + D.mir_di, jr = j - D.mir_dc;
                    if (ir < g.h && jr >= 0 && jr < w) {
                        const int a2 = g.rowmap[ir], b2 = g.colmap[jr];
                        if (a2 >= 0 && b2 >= 0) D.out_mir[a2 * g.nc + b2] = v;
                    }
                }
            } else {
                if (valid) D.out_plane[i * w + j] = v;
            }
        };

        // ---- first row of the strip (core:3345-3362 / :3530-3547) ----
        for (int y = lo; y < lo + k; ++y) load_drow(y);
        __syncwarp();
        for (int p = 0; p < k; ++p)
            tile[p * 32 + lane] = dring[((lo + p) & 63) * 64 + lane + k] - dring[((lo + p) & 63) * 64 + lane];
        __syncwarp();
        if (lane == 0) {
            float left;
            int l0 = 0;
            if (strip == 0) {
                float v = 0.0f;
                for (int p = 0; p < k; ++p)
                    for (int q = 0; q < k; ++q) v += dring[((lo + p) & 63) * 64 + 1 + q];
                frow[0] = v; left = v; l0 = 1;
            } else left = __ldcg(&bnd_prev[lo]);
            for (int l = l0; l <= lastlane; ++l) {
                float s = left;
                for (int p = 0; p < k; ++p) s += tile[p * 32 + l];
                frow[l] = s; left = s;
            }
        }
        __syncwarp();
        float cur = valid ? frow[lane] : 0.f;
        float prevL;
        {
            const float up = __shfl_up_sync(FULL, cur, 1);
            prevL = lane == 0 ? (strip == 0 ? 0.f : __ldcg(&bnd_prev[lo])) : up;
        }
        emit_row(lo, cur);
        if (has_next && lane == 31) __stcg(&bnd_next[lo], cur);
        __syncwarp();

        // ---- wavefront over the remaining rows (core:3365-3387 / :3550-3572) ----
        const int nsteps = (Hh - 1) + 31;
        float bchunk = 0.f;
        for (int s = 1; s <= nsteps; ++s) {
            const int i0 = lo + s;
            if (i0 < g.row_end) load_drow(i0 + k - 1);
            if (strip > 0 && ((s - 1) & 31) == 0) {
                const int r = i0 + lane;
                bchunk = r < g.row_end ? __ldcg(&bnd_prev[r]) : 0.f;
            }
            __syncwarp();
            const float Lsh = __shfl_up_sync(FULL, cur, 1);
            const float bL = __shfl_sync(FULL, bchunk, (s - 1) & 31);
            const int i = i0 - lane;
            const bool active = valid && i > lo && i < g.row_end;
            if (active) {
                float nv;
                if (strip == 0 && lane == 0) {
                    nv = cur;
                    const float *ra = &dring[((i - 1 + k) & 63) * 64 + 1];
                    const float *rb = &dring[((i - 1) & 63) * 64 + 1];
                    for (int q = 0; q < k; ++q) nv += ra[q] - rb[q];
                } else {
                    const float L = lane == 0 ? bL : Lsh;
                    const float *r1 = &dring[((i + k - 1) & 63) * 64 + lane];
                    const float *r0 = &dring[((i - 1) & 63) * 64 + lane];
                    nv = L + cur;
                    nv = nv - prevL;
                    nv = nv + r1[k];
                    nv = nv - r1[0];
                    nv = nv - r0[k];
                    nv = nv + r0[0];
                    prevL = L;
                }
                cur = nv;
                tile[(i & 31) * 32 + lane] = nv;
                if (has_next && lane == 31) __stcg(&bnd_next[i], nv);
            }
            __syncwarp();
            const int idone = i0 - 31;      // the row lane 31 has just finished
            if (idone > lo && idone < g.row_end) emit_row(idone, tile[(idone & 31) * 32 + lane]);
        }
        __syncwarp();
        float *t = bnd_prev; bnd_prev = bnd_next; bnd_next = t;
    }
}

__global__ void k_fill(float *p, float v, size_t n)
{
    for (size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t) gridDim.x * blockDim.x) p[t] = v;
}

// ------------------------------------------------------------------------------------------------------------
// libstdc++ (GCC 13 bits/stl_algo.h, bits/stl_heap.h) partial_sort / sort on (distance, index) pairs compared on the
// distance only (bm3d.cpp:1377-1380), restated for one thread. Only executed where exact ties make the reference's
// output depend on the algorithm.
// ------------------------------------------------------------------------------------------------------------
struct LfPair { float d; unsigned i; };
#define LF_LESS(a, b) ((a).d < (b).d)

__device__ inline void lfs_push_heap(LfPair *first, int hole, int top, LfPair value)
{
    int parent = (hole - 1) / 2;
    while (hole > top && LF_LESS(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
__device__ inline void lfs_adjust_heap(LfPair *first, int hole, int len, LfPair value)
{
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (LF_LESS(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    lfs_push_heap(first, hole, top, value);
}
__device__ inline void lfs_make_heap(LfPair *first, int len)
{
    if (len < 2) return;
    int parent = (len - 2) / 2;
    for (;;) {
        const LfPair v = first[parent];
        lfs_adjust_heap(first, parent, len, v);
        if (parent == 0) return;
        parent--;
    }
}
// pop_heap(first, first+len, result): *result receives the top, its old value is sifted in
__device__ inline void lfs_pop_heap(LfPair *first, int len, LfPair *result)
{
    const LfPair v = *result;
    *result = *first;
    lfs_adjust_heap(first, 0, len, v);
}
__device__ inline void lfs_partial_sort(LfPair *first, int middle, int last)
{
    lfs_make_heap(first, middle);
    for (int i = middle; i < last; ++i)
        if (LF_LESS(first[i], first[0])) lfs_pop_heap(first, middle, &first[i]);
    while (middle > 1) { --middle; lfs_pop_heap(first, middle, &first[middle]); }
}
__device__ inline void lfs_swap(LfPair *a, LfPair *b) { const LfPair t = *a; *a = *b; *b = t; }
__device__ inline void lfs_unguarded_linear_insert(LfPair *v, int last)
{
    const LfPair val = v[last];
    int next = last - 1;
    while (LF_LESS(val, v[next])) { v[last] = v[next]; last = next; --next; }
    v[last] = val;
}
__device__ inline void lfs_insertion_sort(LfPair *v, int first, int last)
{
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (LF_LESS(v[i], v[first])) {
            const LfPair val = v[i];
            for (int t = i; t > first; --t) v[t] = v[t - 1];
            v[first] = val;
        } else lfs_unguarded_linear_insert(v, i);
    }
}
__device__ inline void lfs_sort(LfPair *v, int n)
{
    if (n == 0) return;
    int lg = 0;
    while ((n >> (lg + 1)) > 0) lg++;
    // introsort loop with an explicit stack for the recursion on the right part
    int stf[64], stl[64], std_[64], sp = 0;
    stf[0] = 0; stl[0] = n; std_[0] = lg * 2; sp = 1;
    while (sp > 0) {
        --sp;
        int first = stf[sp], last = stl[sp], depth = std_[sp];
        // emulate: while (last - first > 16) {...; introsort_loop(cut, last, depth); last = cut;}
        // The recursive call on [cut, last) runs BEFORE the loop continues on [first, cut): keep that order by
        // pushing the left continuation and then processing the right part first.
        while (last - first > 16) {
            if (depth == 0) { lfs_partial_sort(v + first, last - first, last - first); break; }
            --depth;
            const int mid = first + (last - first) / 2;
            {   // move_median_to_first(first, first+1, mid, last-1)
                LfPair *r = &v[first], *a = &v[first + 1], *b = &v[mid], *c = &v[last - 1];
                if (LF_LESS(*a, *b)) {
                    if (LF_LESS(*b, *c)) lfs_swap(r, b);
                    else if (LF_LESS(*a, *c)) lfs_swap(r, c);
                    else lfs_swap(r, a);
                } else if (LF_LESS(*a, *c)) lfs_swap(r, a);
                else if (LF_LESS(*b, *c)) lfs_swap(r, c);
                else lfs_swap(r, b);
            }
            int f = first + 1, l = last;
            for (;;) {   // unguarded_partition(first+1, last, pivot = first)
                while (LF_LESS(v[f], v[first])) ++f;
                --l;
                while (LF_LESS(v[first], v[l])) --l;
                if (!(f < l)) break;
                lfs_swap(&v[f], &v[l]);
                ++f;
            }
            const int cut = f;
            // right part [cut, last) first (recursion), left part [first, cut) afterwards (loop continuation).
            // The two ranges are disjoint and nothing else touches them, so deferring the left part is equivalent.
            stf[sp] = first; stl[sp] = cut; std_[sp] = depth; ++sp;
            first = cut;
        }
    }
    if (n > 16) {
        lfs_insertion_sort(v, 0, 16);
        for (int i = 16; i != n; ++i) lfs_unguarded_linear_insert(v, i);
    } else lfs_insertion_sort(v, 0, n);
}

// ------------------------------------------------------------------------------------------------------------
// Self-similarity selection (core:3397-3445): one warp per reference patch.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned lf_fkey(float v)     // monotonic float -> unsigned
{
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float lf_fkey_inv(unsigned key)
{
    return __uint_as_float((key & 0x80000000u) ? (key & 0x7fffffffu) : ~key);
}

struct SelGeom {
    int w, nSim, Ns, N, R, nc;
    float threshold;
    const int *rows, *cols;       // reference rows / columns
};

__global__ void __launch_bounds__(32) k_bm_select(SelGeom g, const float *__restrict__ s_at, const float *__restrict__ s_mir,
                                                  unsigned *__restrict__ out_count, unsigned *__restrict__ out_idx)
{
    extern __shared__ unsigned long long keys[];      // up to Ns*Ns entries, later reused as LfPair[]
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x, r = blockIdx.x;
    const int k_r = g.rows[r / g.nc] * g.w + g.cols[r % g.nc];
    const int Ns = g.Ns, nSim = g.nSim, total = Ns * Ns;

    // candidates in the reference's push order: for dj { di = 0..nSim ; di = -nSim..-1 }
    int cnt = 0;
    for (int base = 0; base < total; base += 32) {
        const int o = base + lane;
        bool keep = false;
        float val = 0.f;
        if (o < total) {
            const int djx = o / Ns, rem = o - djx * Ns;
            float test;
            if (rem <= nSim) {
                const int ddk = djx + rem * Ns;
                test = val = s_at[(size_t) ddk * g.R + r];
            } else {
                const int a = nSim - (rem - nSim - 1);            // a = -di, di = -nSim + (rem - nSim - 1)
                const int ddk = (Ns - 1 - djx) + a * Ns;
                test = s_at[(size_t) ddk * g.R + r];
                val = s_mir[(size_t) ddk * g.R + r];
            }
            keep = test < g.threshold;
        }
        const unsigned m = __ballot_sync(FULL, keep);
        if (keep) {
            const int slot = cnt + __popc(m & ((1u << lane) - 1u));
            keys[slot] = ((unsigned long long) lf_fkey(val + 0.0f) << 32) | (unsigned) o;
        }
        cnt += __popc(m);
    }
    __syncwarp();
    unsigned nSx;
    if ((unsigned) g.N > (unsigned) cnt) { nSx = 1; while (nSx * 2 <= (unsigned) cnt) nSx *= 2; } else nSx = g.N;
    unsigned *dst = out_idx + (size_t) r * (g.N + 1);
    if (cnt == 0) {      // nSx == 1: the reference pushes (0, k_r) and duplicates it
        if (lane == 0) { dst[0] = k_r; dst[1] = k_r; out_count[r] = 2; }
        return;
    }
    // the M = min(cnt, nSx+1) smallest keys, ascending; lane t keeps the t-th (t < 32), `extra` the 33rd
    const int M = min(cnt, (int) nSx + 1);
    unsigned long long prev = 0, mine = ~0ull, extra = ~0ull;
    for (int t = 0; t < M; ++t) {
        unsigned long long best = ~0ull;
        for (int q = lane; q < cnt; q += 32) {
            const unsigned long long kq = keys[q];
            if ((t == 0 || kq > prev) && kq < best) best = kq;
        }
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(FULL, best, o);
            best = other < best ? other : best;
        }
        prev = best;
        if (t < 32) { if (lane == t) mine = best; } else extra = best;
    }
    // exact float ties among the M smallest make the result depend on the heap algorithm
    unsigned long long nxt = __shfl_down_sync(FULL, mine, 1);
    if (lane == 31) nxt = extra;
    const bool tie_here = (lane + 1 < M) && ((unsigned) (mine >> 32) == (unsigned) (nxt >> 32));
    const bool tie = __any_sync(FULL, tie_here);
    auto idx_of = [&](unsigned o) -> unsigned {
        const int djx = (int) o / Ns, rem = (int) o - djx * Ns;
        const int di = rem <= nSim ? rem : -nSim + (rem - nSim - 1);
        return (unsigned) (k_r + di * g.w + (djx - nSim));
    };
    if (!tie) {
        if (lane < (int) nSx) dst[lane] = idx_of((unsigned) (mine & 0xffffffffu));
        if (lane == 0) {
            if (nSx == 1) { dst[1] = dst[0] = idx_of((unsigned) (mine & 0xffffffffu)); out_count[r] = 2; }
            else out_count[r] = nSx;
        }
        return;
    }
    // slow path: one thread runs the reference's partial_sort on the pairs in push order
    LfPair *pairs = reinterpret_cast<LfPair *>(keys);
    for (int q = lane; q < cnt; q += 32) {
        const unsigned long long kq = keys[q];
        LfPair pr;
        pr.d = lf_fkey_inv((unsigned) (kq >> 32));
        pr.i = idx_of((unsigned) (kq & 0xffffffffu));
        pairs[q] = pr;          // same 8-byte slot, each lane converts its own entries
    }
    __syncwarp();
    if (lane == 0) {
        lfs_partial_sort(pairs, (int) nSx, cnt);
        for (unsigned t = 0; t < nSx; ++t) dst[t] = pairs[t].i;
        if (nSx == 1) { dst[1] = pairs[0].i; out_count[r] = 2; } else out_count[r] = nSx;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Disparity matching (core:3576-3608): one thread per position of the dense grid. Only element [0] of the sorted
// list and the shape flag are consumed downstream (core:294, 310, 503, 510).
// ------------------------------------------------------------------------------------------------------------
__global__ void k_stereo_argmin(const float *__restrict__ sums, int w, int h, int nDisp, int row_end, int col_end, float threshold,
                                unsigned *__restrict__ out_first, unsigned char *__restrict__ out_shape, unsigned *tie_counter)
{
    const int Ns = 2 * nDisp + 1, np = Ns * Ns;
    const size_t plane = (size_t) w * h;
    const int ww = col_end - nDisp, hh = row_end - nDisp;
    const size_t total = (size_t) ww * hh;
    for (size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t) gridDim.x * blockDim.x) {
        const int i = nDisp + (int) (t / ww), j = nDisp + (int) (t % ww);
        const int k_r = i * w + j;
        float best = 0.f;
        int nmin = 0, amin = 0, c = 0;
        for (int djx = 0; djx < Ns; ++djx)
            for (int dix = 0; dix < Ns; ++dix, ++c) {
                const float v = sums[(size_t) (djx + dix * Ns) * plane + k_r];
                if (c == 0 || v < best) { best = v; nmin = 1; amin = c; }
                else if (v == best) nmin++;
            }
        unsigned first = (unsigned) (k_r + (amin % Ns - nDisp) * w + (amin / Ns - nDisp));
        if (nmin > 1) {
            LfPair td[LF_MAXNS2];
            c = 0;
            for (int djx = 0; djx < Ns; ++djx)
                for (int dix = 0; dix < Ns; ++dix, ++c) {
                    td[c].d = sums[(size_t) (djx + dix * Ns) * plane + k_r];
                    td[c].i = (unsigned) (k_r + (dix - nDisp) * w + (djx - nDisp));
                }
            lfs_sort(td, np);
            first = td[0].i;
            if (tie_counter) atomicAdd(tie_counter, 1u);
        }
        out_first[k_r] = first;
        out_shape[k_r] = best < threshold ? 1 : 0;
    }
}
