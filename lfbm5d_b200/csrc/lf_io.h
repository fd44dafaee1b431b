// Host glue of the command-line drivers: a self-contained PNG codec (all bit depths, Adam7) over
// zlib (the reference wraps libpng, io_png.c:379 / :700; libpng is not in this image), the light-field loader / saver
// with the reference's file naming (utilities_LF.cpp:105-112), noise (utilities.cpp:154-185, mt19937ar), PSNR / RMSE
// (utilities.cpp:412-435, utilities_LF.cpp:639-700), difference images (utilities.cpp:440-470) and the PSNR report
// (utilities_LF.cpp:782-869). Header only.
#pragma once
#include <zlib.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <numeric>
#include <sstream>
#include <string>
#include <vector>
#include <atomic>
#include <mutex>
#include <thread>
#include <sys/time.h>
#include <unistd.h>

namespace lfio {

// SAIs are independent files / arrays: decode, encode and noise generation run on the host cores (the reference does the same with
// `#pragma omp parallel for`, utilities_LF.cpp:102, :255). LFBM5D_IO_THREADS=<n> sets the number of threads (1 = sequential).
template <class F> inline void parallel_for(unsigned n, F fn)
{
    unsigned nt = std::thread::hardware_concurrency();
    if (const char *e = getenv("LFBM5D_IO_THREADS")) nt = (unsigned) atoi(e);
    nt = std::max(1u, std::min(std::min(nt, 32u), n));
    if (nt <= 1) { for (unsigned i = 0; i < n; i++) fn(i); return; }
    std::atomic<unsigned> next(0);
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; t++)
        pool.emplace_back([&]() { for (unsigned i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i); });
    for (auto &th : pool) th.join();
}

// ---------------------------------------------------------------- PNG
inline uint32_t be32(const unsigned char *p) { return ((uint32_t) p[0] << 24) | ((uint32_t) p[1] << 16) | ((uint32_t) p[2] << 8) | p[3]; }
inline int paeth(int a, int b, int c) { const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c); return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); }

// One pass of a PNG image (the whole image, or one of the seven Adam7 sub-images): pw x ph pixels of `bits` bits, every scanline
// preceded by its filter type; unfiltered in place (PNG specification section 9: the filters work on bytes, `fb` = bytes per complete
// pixel, at least 1) and unpacked to one byte per sample at (x0 + x * dx, y0 + y * dy) of the full image. Returns the bytes consumed.
inline size_t png_pass(unsigned char *raw, size_t avail, size_t pw, size_t ph, unsigned nch, unsigned depth, std::vector<unsigned char> &img, size_t w,
                       size_t x0, size_t y0, size_t dx, size_t dy, bool *ok)
{
    if (!pw || !ph) return 0;
    const size_t bits = (size_t) nch * depth, rowbytes = (pw * bits + 7) / 8, fb = bits >= 8 ? bits / 8 : 1;
    if (avail < (rowbytes + 1) * ph) { *ok = false; return 0; }
    for (size_t y = 0; y < ph; y++) {
        unsigned char *line = raw + y * (rowbytes + 1) + 1;
        const unsigned char ft = line[-1], *up = y ? line - (rowbytes + 1) : nullptr;
        if (ft > 4) { *ok = false; return 0; }
        for (size_t x = 0; x < rowbytes; x++) {
            const int a = x >= fb ? line[x - fb] : 0, b = up ? up[x] : 0, cc = (up && x >= fb) ? up[x - fb] : 0;
            int v = line[x];
            switch (ft) { case 1: v += a; break; case 2: v += b; break; case 3: v += (a + b) / 2; break; case 4: v += paeth(a, b, cc); break; default: break; }
            line[x] = (unsigned char) v;
        }
        unsigned char *dst = &img[((y0 + y * dy) * w + x0) * nch];
        for (size_t x = 0; x < pw; x++, dst += dx * nch)
            for (unsigned ch = 0; ch < nch; ch++) {
                const size_t sidx = x * nch + ch;
                if (depth == 8) dst[ch] = line[sidx];
                else if (depth == 16) dst[ch] = line[2 * sidx];      // most significant byte, as PNG_TRANSFORM_STRIP_16 does
                else {      // 1, 2, 4 bits, leftmost sample in the high-order bits; unpacked, not scaled (PNG_TRANSFORM_PACKING)
                    const size_t bit = sidx * depth;
                    dst[ch] = (unsigned char) ((line[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1));
                }
            }
    }
    return (rowbytes + 1) * ph;
}

// Reads a PNG into planar float (c*W*H + i*W + j) the way read_png_f32 does (io_png.c:379 -> :116-260: png_read_png with
// PNG_TRANSFORM_STRIP_16 | PNG_TRANSFORM_PACKING and nothing else): c = the channels of the file (1 gray or palette INDEX, 2
// gray + alpha, 3 RGB, 4 RGBA; load_LF then keeps 3 of 4), 16-bit samples cut to their high byte, 1/2/4-bit samples unpacked
// without scaling, any interlace method (Adam7 sub-images are put back in place). Returns false on error.
inline bool read_png_f32(const std::string &name, std::vector<float> &out, size_t &w, size_t &h, size_t &c)
{
    std::ifstream f(name.c_str(), std::ios::binary);
    if (!f) return false;
    std::vector<unsigned char> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    static const unsigned char sig[8] = { 137, 80, 78, 71, 13, 10, 26, 10 };
    if (buf.size() < 33 || memcmp(buf.data(), sig, 8) != 0) return false;
    size_t pos = 8;
    unsigned depth = 0, ctype = 0, interlace = 0;
    std::vector<unsigned char> idat;
    w = h = 0;
    while (pos + 12 <= buf.size()) {
        const uint32_t len = be32(&buf[pos]);
        const std::string type((const char *) &buf[pos + 4], 4);
        if (pos + 12 + (size_t) len > buf.size()) return false;
        const unsigned char *d = &buf[pos + 8];
        if (type == "IHDR") { if (len < 13) return false; w = be32(d); h = be32(d + 4); depth = d[8]; ctype = d[9]; interlace = d[12]; }
        else if (type == "IDAT") idat.insert(idat.end(), d, d + len);
        else if (type == "IEND") break;
        pos += 12 + (size_t) len;
    }
    const unsigned nch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    const bool depth_ok = ctype == 0 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)
                        : ctype == 3 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8) : (depth == 8 || depth == 16);
    if (!w || !h || !nch || !depth_ok || interlace > 1 || w > (1u << 20) || h > (1u << 20)) return false;
    // Adam7: start / step of the seven passes (PNG specification section 8.2); pass 0 of a non-interlaced file is the image
    static const unsigned ax0[7] = { 0, 4, 0, 2, 0, 1, 0 }, ay0[7] = { 0, 0, 4, 0, 2, 0, 1 }, adx[7] = { 8, 8, 4, 4, 2, 2, 1 }, ady[7] = { 8, 8, 8, 4, 4, 2, 2 };
    const size_t bits = (size_t) nch * depth;
    size_t rawlen = 0;
    const int npass = interlace ? 7 : 1;
    for (int q = 0; q < npass; q++) {
        const size_t pw = interlace ? (w + adx[q] - 1 - ax0[q]) / adx[q] : w, ph = interlace ? (h + ady[q] - 1 - ay0[q]) / ady[q] : h;
        if (pw && ph) rawlen += ((pw * bits + 7) / 8 + 1) * ph;
    }
    std::vector<unsigned char> raw(rawlen);
    uLongf got = (uLongf) rawlen;
    if (idat.empty() || uncompress(raw.data(), &got, idat.data(), (uLong) idat.size()) != Z_OK || got != rawlen) return false;
    std::vector<unsigned char> img(w * h * nch);
    bool ok = true;
    size_t off = 0;
    for (int q = 0; q < npass && ok; q++) {
        const size_t pw = interlace ? (w + adx[q] - 1 - ax0[q]) / adx[q] : w, ph = interlace ? (h + ady[q] - 1 - ay0[q]) / ady[q] : h;
        off += png_pass(raw.data() + off, rawlen - off, pw, ph, nch, depth, img, w, interlace ? ax0[q] : 0, interlace ? ay0[q] : 0,
                        interlace ? adx[q] : 1, interlace ? ady[q] : 1, &ok);
    }
    if (!ok) return false;
    c = nch;
    out.assign(w * h * c, 0.0f);
    for (size_t ch = 0; ch < c; ch++)
        for (size_t y = 0; y < h; y++)
            for (size_t x = 0; x < w; x++) out[ch * w * h + y * w + x] = (float) img[(y * w + x) * nch + ch];
    return true;
}

inline void put32(std::vector<unsigned char> &v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
inline void chunk(std::vector<unsigned char> &png, const char *type, const std::vector<unsigned char> &data)
{
    put32(png, (uint32_t) data.size());
    const size_t s = png.size();
    png.insert(png.end(), type, type + 4);
    png.insert(png.end(), data.begin(), data.end());
    put32(png, (uint32_t) crc32(0, &png[s], (uInt) (png.size() - s)));
}
// 8-bit PNG from planar float, rounded floor(x + .5) and clamped like write_png_f32 (io_png.c:648-650).
inline bool write_png_f32(const std::string &name, const float *data, size_t w, size_t h, size_t c)
{
    const size_t stride = w * c;
    std::vector<unsigned char> raw((stride + 1) * h);
    for (size_t y = 0; y < h; y++) {
        raw[y * (stride + 1)] = 0;
        for (size_t x = 0; x < w; x++)
            for (size_t ch = 0; ch < c; ch++) {
                float v = floorf(data[ch * w * h + y * w + x] + 0.5f);
                v = v < 0.f ? 0.f : (v > 255.f ? 255.f : v);
                raw[y * (stride + 1) + 1 + x * c + ch] = (unsigned char) v;
            }
    }
    uLongf zlen = compressBound(raw.size());
    std::vector<unsigned char> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), raw.size(), 6) != Z_OK) return false;
    z.resize(zlen);
    std::vector<unsigned char> png = { 137, 80, 78, 71, 13, 10, 26, 10 }, ihdr;
    put32(ihdr, (uint32_t) w); put32(ihdr, (uint32_t) h);
    ihdr.push_back(8); ihdr.push_back(c == 1 ? 0 : 2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk(png, "IHDR", ihdr); chunk(png, "IDAT", z); chunk(png, "IEND", {});
    std::ofstream f(name.c_str(), std::ios::binary);
    if (!f) return false;
    f.write((const char *) png.data(), png.size());
    return (bool) f;
}

// ---------------------------------------------------------------- light field files
inline std::string sai_path(const char *dir, const char *sub, const char *sep, unsigned s, unsigned t)
{   // utilities_LF.cpp:105-112: <dir>/<name><sep>%02d<sep>%02d.png
    std::ostringstream o;
    o << dir << "/" << sub << sep << std::setw(2) << std::setfill('0') << s << sep << std::setw(2) << std::setfill('0') << t << ".png";
    return o.str();
}
inline int load_LF(const char *dir, const char *sub, const char *sep, std::vector<std::vector<float> > &LF, std::vector<unsigned> &mask,
                   unsigned ang_major, unsigned awidth, unsigned aheight, unsigned s_start, unsigned t_start, unsigned *width,
                   unsigned *height, unsigned *chnls, unsigned ROW)
{
    const unsigned asize = awidth * aheight;
    LF.assign(asize, std::vector<float>());
    mask.assign(asize, 0u);
    std::cout << std::endl;
    std::mutex out_mu;
    std::vector<std::string> failed(asize);
    std::vector<unsigned> dims(3 * (size_t) asize, 0u);
    parallel_for(asize, [&](unsigned q) {
        const unsigned s = q / awidth, t = q % awidth;
        const std::string name = sai_path(dir, sub, sep, s + s_start, t + t_start);
        { std::lock_guard<std::mutex> lk(out_mu); std::cout << "\rRead input image " << name << std::flush; }
        size_t w = 0, h = 0, c = 0;
        std::vector<float> tmp;
        if (!read_png_f32(name, tmp, w, h, c)) { failed[q] = name; return; }
        if (c > 2) {       // grey image stored as colour (utilities_LF.cpp:123-130)
            size_t k = 0;
            float acc = 0.0f;
            while (k < w * h && tmp[k] == tmp[w * h + k] && tmp[k] == tmp[2 * w * h + k]) { acc += tmp[k] + tmp[w * h + k] + tmp[2 * w * h + k]; k++; }
            c = (k == w * h && acc > 0.0f) ? 1 : 3;
        }
        dims[3 * q] = (unsigned) w; dims[3 * q + 1] = (unsigned) h; dims[3 * q + 2] = (unsigned) c;
        const unsigned st = ang_major == ROW ? s * awidth + t : s + t * aheight;
        LF[st].assign(tmp.begin(), tmp.begin() + w * h * c);
        for (size_t k = 0; k < w * h * c; k++) if (tmp[k]) { mask[st] = 1; break; }      // utilities_LF.cpp:149-154
    });
    for (unsigned q = 0; q < asize; q++)
        if (!failed[q].empty()) {
            std::cout << std::endl << "error :: " << failed[q] << " not found or not a correct png image." << dir << " folder might not exist." << std::endl;
            return EXIT_FAILURE;      // (the reference prints and then dereferences a null pointer here, utilities_LF.cpp:117-124)
        }
    *width = dims[0]; *height = dims[1]; *chnls = dims[2];      // of the first image (s = t = 0), utilities_LF.cpp:133-138
    std::cout << std::endl << " Light field size :" << std::endl << " - awidth         = " << awidth << std::endl << " - aheight        = " << aheight << std::endl
              << " - width          = " << *width << std::endl << " - height         = " << *height << std::endl << " - nb of channels = " << *chnls << std::endl;
    return EXIT_SUCCESS;
}
inline int save_LF(const char *dir, const char *sub, const char *sep, const std::vector<std::vector<float> > &LF, const std::vector<unsigned> &mask,
                   unsigned ang_major, unsigned awidth, unsigned aheight, unsigned s_start, unsigned t_start, unsigned width,
                   unsigned height, unsigned chnls, unsigned ROW)
{
    const unsigned asize = awidth * aheight;
    std::vector<char> bad(asize, 0);
    parallel_for(asize, [&](unsigned q) {
        const unsigned s = q / awidth, t = q % awidth;
        const unsigned st = ang_major == ROW ? s * awidth + t : s + t * aheight;
        if (!mask[st]) return;
        if (!write_png_f32(sai_path(dir, sub, sep, s + s_start, t + t_start), LF[st].data(), width, height, chnls)) bad[q] = 1;
    });
    for (unsigned q = 0; q < asize; q++)
        if (bad[q]) {
            std::cout << "... failed to save png image " << sai_path(dir, sub, sep, q / awidth + s_start, q % awidth + t_start) << std::endl;
            return EXIT_FAILURE;
        }
    return EXIT_SUCCESS;
}

// ---------------------------------------------------------------- noise (mt19937ar + Box-Muller, utilities.cpp:154-185)
struct MT {
    unsigned long mt[624]; int mti;
    void seed(unsigned long s) { mt[0] = s & 0xffffffffUL; for (mti = 1; mti < 624; mti++) { mt[mti] = (1812433253UL * (mt[mti - 1] ^ (mt[mti - 1] >> 30)) + (unsigned long) mti); mt[mti] &= 0xffffffffUL; } }
    unsigned long next()
    {
        static const unsigned long mag01[2] = { 0x0UL, 0x9908b0dfUL };
        unsigned long y;
        if (mti >= 624) {
            int kk;
            for (kk = 0; kk < 624 - 397; kk++) { y = (mt[kk] & 0x80000000UL) | (mt[kk + 1] & 0x7fffffffUL); mt[kk] = mt[kk + 397] ^ (y >> 1) ^ mag01[y & 1UL]; }
            for (; kk < 623; kk++) { y = (mt[kk] & 0x80000000UL) | (mt[kk + 1] & 0x7fffffffUL); mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1UL]; }
            y = (mt[623] & 0x80000000UL) | (mt[0] & 0x7fffffffUL); mt[623] = mt[396] ^ (y >> 1) ^ mag01[y & 1UL];
            mti = 0;
        }
        y = mt[mti++];
        y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680UL; y ^= (y << 15) & 0xefc60000UL; y ^= (y >> 18);
        return y & 0xffffffffUL;
    }
    double res53() { const unsigned long a = next() >> 5, b = next() >> 6; return (1.0 * a * 67108864.0 + b) * (1.0 / 9007199254740992.0); }
};
// One generator per SAI. The reference seeds from time + pid (not reproducible); LFBM5D_SEED=<n> fixes seed n + st instead.
inline void add_noise_LF(const std::vector<std::vector<float> > &LF, const std::vector<unsigned> &mask, std::vector<std::vector<float> > &noisy, float sigma)
{
    const char *fixed = getenv("LFBM5D_SEED");
    struct timeval tp;
    gettimeofday(&tp, nullptr);
    const unsigned long seed0 = fixed ? strtoul(fixed, nullptr, 10) : (unsigned long) (tp.tv_sec * 1000 + tp.tv_usec / 1000) + (unsigned long) getpid();
    if (noisy.size() < LF.size()) noisy.resize(LF.size());
    parallel_for((unsigned) LF.size(), [&](unsigned st) {
        if (!mask[st]) return;
        MT g;
        g.seed(seed0 + st);
        noisy[st].resize(LF[st].size());
        for (size_t k = 0; k < LF[st].size(); k++) {
            const double a = g.res53(), b = g.res53();
            const double z = (double) sigma * sqrt(-2.0 * log(a)) * cos(2.0 * M_PI * b);
            noisy[st][k] = LF[st][k] + (float) z;
        }
    });
}

// ---------------------------------------------------------------- metrics
inline void compute_psnr(const std::vector<float> &a, const std::vector<float> &b, float *psnr, float *rmse)
{
    float tmp = 0.0f;
    for (size_t k = 0; k < a.size(); k++) tmp += (a[k] - b[k]) * (a[k] - b[k]);
    *rmse = sqrtf(tmp / (float) a.size());
    *psnr = 20.0f * log10f(255.0f / (*rmse));
}
inline int compute_psnr_LF(const std::vector<std::vector<float> > &A, const std::vector<std::vector<float> > &B, const std::vector<unsigned> &mask,
                           std::vector<float> &psnr, float *avg_psnr, float *std_psnr, std::vector<float> &rmse, float *avg_rmse, float *std_rmse)
{
    if (A.size() != B.size()) { std::cout << "Can't compute PSNR & RMSE, LF_1 and LF_2 don't have the same size" << std::endl; return EXIT_FAILURE; }
    const size_t asize = mask.size();
    psnr.assign(asize, 0.0f); rmse.assign(asize, 0.0f);
    float n = 0;
    for (size_t st = 0; st < asize; st++) if (mask[st]) { compute_psnr(A[st], B[st], &psnr[st], &rmse[st]); n += 1.0f; }
    *avg_psnr = (float) (std::accumulate(psnr.begin(), psnr.end(), 0.0) / n);
    *avg_rmse = (float) (std::accumulate(rmse.begin(), rmse.end(), 0.0) / n);
    float sp = 0.0f, sr = 0.0f;
    for (size_t st = 0; st < asize; st++) if (mask[st]) { sp += (psnr[st] - *avg_psnr) * (psnr[st] - *avg_psnr); sr += (rmse[st] - *avg_rmse) * (rmse[st] - *avg_rmse); }
    *std_psnr = sqrtf(sp / n); *std_rmse = sqrtf(sr / n);
    return EXIT_SUCCESS;
}
inline void compute_diff_LF(const std::vector<std::vector<float> > &A, const std::vector<std::vector<float> > &B, const std::vector<unsigned> &mask,
                            std::vector<std::vector<float> > &D, float sigma)
{
    const float s = 4.0f * sigma;
    D.resize(A.size());
    for (size_t st = 0; st < A.size(); st++) {
        if (!mask[st]) continue;
        D[st].resize(A[st].size());
        for (size_t k = 0; k < A[st].size(); k++) {
            const float v = s > 0.0 ? (A[st][k] - B[st][k] + s) * 255.0f / (2.0f * s) : fabsf(A[st][k] - B[st][k]);
            D[st][k] = v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v);
        }
    }
}
inline int write_psnr_LF(const char *file_name, const char *LF_name, const std::vector<unsigned> &mask, unsigned ang_major, unsigned awidth, unsigned aheight,
                         const std::vector<float> &psnr, float avg_psnr, float std_psnr, const std::vector<float> &rmse, float avg_rmse, float std_rmse, unsigned ROW)
{
    std::ofstream file(file_name, std::ios::out | std::ios::app);
    if (!file) { std::cout << "Can't open " << file_name << std::endl; return EXIT_FAILURE; }
    file << std::endl << "******************************************" << std::endl;
    file << "-> Average PSNR " << LF_name << " = " << avg_psnr << std::endl;
    file << "-> Standard deviation PSNR " << LF_name << " = " << std_psnr << std::endl;
    file << "PSNR for all " << LF_name << " SAIs:" << std::endl;
    for (unsigned s = 0; s < aheight; s++) {
        for (unsigned t = 0; t < awidth; t++) {
            const unsigned st = ang_major == ROW ? s * awidth + t : s + t * aheight;
            if (mask[st]) file << psnr[st] << " "; else file << "No SAI ";
        }
        file << std::endl;
    }
    file << std::endl;
    file << "-> Average RMSE " << LF_name << " = " << avg_rmse << std::endl;
    file << "-> Standard deviation RMSE " << LF_name << " = " << std_rmse << std::endl;
    file << "RMSE for all " << LF_name << " SAIs:" << std::endl;
    for (unsigned s = 0; s < aheight; s++) {
        for (unsigned t = 0; t < awidth; t++) file << rmse[ang_major == ROW ? s * awidth + t : s + t * aheight] << " ";
        file << std::endl;
    }
    file << "******************************************" << std::endl;
    return EXIT_SUCCESS;
}
inline double now() { struct timeval tp; gettimeofday(&tp, nullptr); return tp.tv_sec + tp.tv_usec * 1e-6; }

} // namespace lfio
