// Probe: duration of the summed-area plane kernel as a function of the number of plane groups in the launch
// (one disparity slot = 13 groups of 13 planes; self groups = 14 planes). Not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -std=c++17 -I lfbm5d_b200/csrc -o gpurun_out/sat_probe tools/probe/sat_probe.cu
#include "block_matching.cuh"
#include <cstdio>
#include <cstring>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

template <bool SELF, int K, bool TMA = true>
int run(int w, int h, int nDisp, int nSim, int nslots_or_groups, int reps)
{
    const int n = nSim + nDisp;
    const size_t plane = (size_t) w * h;
    float *img;
    CK(cudaMalloc(&img, plane * 9 * 4));
    std::vector<float> hi(plane * 9);
    unsigned s = 12345;
    for (auto &v : hi) { s = s * 1664525u + 1013904223u; v = (float) (s >> 8) / 65536.0f; }
    CK(cudaMemcpy(img, hi.data(), plane * 9 * 4, cudaMemcpyHostToDevice));
    std::vector<SatPlane> planes;
    std::vector<SatGroup> groups;
    SatGeom g{};
    const int Nd = 2 * nDisp + 1, Ns = 2 * nSim + 1;
    float *sums = nullptr, *s_at = nullptr, *s_mir = nullptr;
    int *rowmap = nullptr, *colmap = nullptr;
    if (!SELF) {
        const int st_lo = nDisp, row_end = h - nDisp - K + 1, col_end = w - nDisp - K + 1;
        const int strips = (col_end - st_lo + 31) / 32, SR = (row_end - st_lo) + 31;
        const size_t stride = (size_t) strips * SR * 32;
        CK(cudaMalloc(&sums, (size_t) nslots_or_groups * Nd * Nd * stride * 4));
        for (int slot = 0; slot < nslots_or_groups; slot++)
            for (int di = 0; di < Nd; di++) {
                SatGroup G{};
                G.img1 = img; G.img2 = img + (size_t) (1 + slot % 8) * plane; G.z1 = 0; G.z2 = 1 + slot % 8; G.oy = di - nDisp; G.oxmin = -nDisp; G.first_plane = (int) planes.size();
                for (int dj = 0; dj < Nd; dj++) { SatPlane P{}; P.ox = dj - nDisp; P.out_skew = sums + ((size_t) slot * Nd * Nd + di * Nd + dj) * stride; planes.push_back(P); G.nplanes++; }
                groups.push_back(G);
            }
        g.w = w; g.h = h; g.k = K; g.lo = st_lo; g.row_end = row_end; g.col_end = col_end; g.ylim = h; g.xlim = w; g.nstrips = strips; g.pstrips = strips; g.SR = SR; g.gp = 1;
    } else {
        const int p = 4;
        std::vector<int> rows, cols, rm(h, -1), cm(w, -1);
        for (int i = n; i < h - K + 1 - n; i += p) rows.push_back(i);
        if (rows.back() < h - K + 1 - n - 1) rows.push_back(h - K + 1 - n - 1);
        for (int i = n; i < w - K + 1 - n; i += p) cols.push_back(i);
        if (cols.back() < w - K + 1 - n - 1) cols.push_back(w - K + 1 - n - 1);
        for (size_t a = 0; a < rows.size(); a++) rm[rows[a]] = (int) a;
        for (size_t a = 0; a < cols.size(); a++) cm[cols[a]] = (int) a;
        const size_t R = rows.size() * cols.size();
        CK(cudaMalloc(&s_at, (size_t) (nSim + 1) * Ns * R * 4)); CK(cudaMalloc(&s_mir, (size_t) (nSim + 1) * Ns * R * 4));
        CK(cudaMalloc(&rowmap, h * 4)); CK(cudaMalloc(&colmap, w * 4));
        CK(cudaMemcpy(rowmap, rm.data(), h * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(colmap, cm.data(), w * 4, cudaMemcpyHostToDevice));
        int ng = 0;
        for (int di = 0; di <= nSim && ng < nslots_or_groups; di++)
            for (int djx0 = 0; djx0 < Ns && ng < nslots_or_groups; djx0 += 2 * SAT_NW, ng++) {
                SatGroup G{};
                G.img1 = img; G.img2 = img; G.z1 = G.z2 = 0; G.oy = di; G.oxmin = djx0 - nSim; G.first_plane = (int) planes.size();
                for (int djx = djx0; djx < std::min(Ns, djx0 + 2 * SAT_NW); djx++) {
                    SatPlane P{}; const int ddk = di * Ns + djx;
                    P.ox = djx - nSim; P.out_at = s_at + (size_t) ddk * R; P.out_mir = s_mir + (size_t) ddk * R; P.mir_di = di; P.mir_dc = nSim - djx;
                    planes.push_back(P); G.nplanes++;
                }
                groups.push_back(G);
            }
        g.w = w; g.h = h; g.k = K; g.lo = n; g.row_end = h - n; g.col_end = w - n; g.ylim = h - n; g.xlim = w - n;
        g.nstrips = (g.col_end - n + 31) / 32; g.pstrips = g.nstrips; g.SR = 0; g.nc = (int) cols.size(); g.rowmap = rowmap; g.colmap = colmap;
        g.gp = p; g.nr = (int) rows.size(); g.rlast = rows.back(); g.nreg = 0;
        for (int ind = n; ind < h - K + 1 - n; ind += p) g.nreg++;
    }
    g.negzero2 = 0x8000000080000000ull;
    SatPlane *dp; SatGroup *dg; unsigned long long *bnd; int *prog;
    CK(cudaMalloc(&dp, planes.size() * sizeof(SatPlane))); CK(cudaMalloc(&dg, groups.size() * sizeof(SatGroup)));
    CK(cudaMemcpy(dp, planes.data(), planes.size() * sizeof(SatPlane), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dg, groups.data(), groups.size() * sizeof(SatGroup), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&bnd, planes.size() * (size_t) g.pstrips * h * 8)); CK(cudaMemset(bnd, 0, planes.size() * (size_t) g.pstrips * h * 8));
    const size_t pbytes = (4 + planes.size() * (size_t) g.pstrips) * 4;
    CK(cudaMalloc(&prog, pbytes));
    float *frow, *fcol;
    CK(cudaMalloc(&frow, planes.size() * (size_t) w * 4)); CK(cudaMalloc(&fcol, planes.size() * (size_t) h * 4));
    g.frow = frow; g.fcol = fcol;
    const size_t smem = 2 * (128 + K) * 64 * 4;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    const bool tma = TMA && (w % 4 == 0);
    if (tma) {
        typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                      const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void *fp = nullptr;
        cudaDriverEntryPointQueryResult qr;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr));
        const cuuint64_t dims[3] = { (cuuint64_t) w, (cuuint64_t) h, 9 }, strides[2] = { (cuuint64_t) w * 4, (cuuint64_t) w * h * 4 };
        const cuuint32_t box[3] = { 64, 8, 1 }, estr[3] = { 1, 1, 1 };
        if (((encode_fn) fp)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, img, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("tensor map failed\n"); return 1; }
    }
    auto kfn = tma ? k_sat2<SELF, K, true> : k_sat2<SELF, K, false>;
    CK(cudaFuncSetAttribute((const void *) kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    float best_edge = 1e9f;
    float best = 1e9f;
    for (int r = 0; r < reps; r++) {
        CK(cudaMemset(prog, 0, pbytes));
        CK(cudaEventRecord(e0));
        {
            const size_t smem_e = sate_smem(K);
            CK(cudaFuncSetAttribute((const void *) k_sat_edges<SELF, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_e));
            k_sat_edges<SELF, K><<<dim3((int) groups.size(), 2), SATE_NT, smem_e>>>(g, dg, dp, frow, fcol);
        }
        CK(cudaEventRecord(e2));
        g.epoch = (unsigned) (r + 1);
        kfn<<<(int) groups.size() * g.nstrips, SAT_NW * 32, smem>>>(g, dg, dp, (int) groups.size(), bnd, prog, tmap);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
        cudaEventElapsedTime(&ms, e0, e2);
        best_edge = std::min(best_edge, ms);
    }
    printf("%s%s K=%d groups=%zu planes=%zu strips=%d CTAs=%zu : %.3f ms (edges %.3f)  (%.2f us per plane)\n", SELF ? "self  " : "stereo", tma ? " tma" : " cpa", K, groups.size(), planes.size(), g.nstrips,
           groups.size() * g.nstrips, best, best_edge, 1e3 * best / planes.size());
    cudaFree(frow); cudaFree(fcol);
    cudaFree(img); cudaFree(sums); cudaFree(s_at); cudaFree(s_mir); cudaFree(rowmap); cudaFree(colmap); cudaFree(dp); cudaFree(dg); cudaFree(bnd); cudaFree(prog);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc > 1 && argv[1][0] == 't') return run<true, 16>(1072, 1072, 6, 18, 1, 1);      // one self group through the TMA variant
    if (argc > 1 && argv[1][0] == 'u') return run<false, 8>(1072, 1072, 6, 18, 1, 1);      // one stereo slot through the TMA variant
    if (argc > 1 && argv[1][0] == 'o') return run<true, 16>(1072, 1072, 6, 18, 1, 1);      // one launch (for ncu)
    if (argc > 1 && argv[1][0] == 'f') return run<true, 16>(1072, 1072, 6, 18, 57, 1);     // full self launch (for ncu)
    if (argc > 1) {      // lone strips: unit time and per-strip lag
        for (int w : { 80, 112, 176, 304, 560 }) if (run<true, 16>(w, 1072, 6, 18, 1, 3)) return 1;
        for (int w : { 80, 112, 176, 304, 560 }) if (run<true, 8>(w, 1072, 6, 18, 1, 3)) return 1;
        for (int w : { 60, 92, 156, 284 }) if (run<false, 16>(w, 1072, 6, 18, 1, 3)) return 1;
        for (int hh : { 272, 528 }) if (run<true, 16>(1072, hh, 6, 18, 1, 3)) return 1;
        return 0;
    }
    for (int s : { 1, 8 }) if (run<false, 16, false>(1072, 1072, 6, 18, s, 3)) return 1;
    for (int gq : { 1, 7, 57 }) if (run<true, 16, false>(1072, 1072, 6, 18, gq, 3)) return 1;
    for (int s : { 1, 2, 4, 8 }) if (run<false, 16>(1072, 1072, 6, 18, s, 3)) return 1;
    for (int s : { 1, 2, 4, 8 }) if (run<false, 8>(1072, 1072, 6, 18, s, 3)) return 1;
    for (int gq : { 1, 4, 7, 13, 26, 57 }) if (run<true, 16>(1072, 1072, 6, 18, gq, 3)) return 1;
    for (int gq : { 1, 7, 13, 57 }) if (run<true, 8>(1072, 1072, 6, 18, gq, 3)) return 1;
    return 0;
}
