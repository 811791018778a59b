// Core arithmetic of the on-device crop generation, shared by the CUDA kernels (csrc/crops.cu) and by the
// host-side emulation the CPU tests build with g++ (tests/host_emul/crops_emul.cpp) — the emulation is test
// infrastructure only and is NOT part of libclipself_b200.so.  See crops.cu for what this restates
// (Pillow ImagingResample + the reference's crop transforms).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define CS_HD __host__ __device__ __forceinline__
#else
#define CS_HD inline
#endif

namespace cs {
namespace crops {

constexpr int PRECISION_BITS = 32 - 8 - 2;

struct Desc {            // == cs_crop_desc_t
    int x0, y0, x1, y1;  // source rectangle in image pixels (may exceed the image: zeros, like Image.crop)
    int out_w, out_h;    // resized size
    int pad_left, pad_top;
};

// IEEE round-to-nearest steps WITHOUT fused multiply-add, so device and host round exactly like Pillow's C code
#if defined(__CUDA_ARCH__)
CS_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
CS_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
CS_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
CS_HD double ddiv(double a, double b) { return __ddiv_rn(a, b); }
CS_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
CS_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
#else
CS_HD double dadd(double a, double b) { volatile double r = a + b; return r; }
CS_HD double dsub(double a, double b) { volatile double r = a - b; return r; }
CS_HD double dmul(double a, double b) { volatile double r = a * b; return r; }
CS_HD double ddiv(double a, double b) { volatile double r = a / b; return r; }
CS_HD float fsub(float a, float b) { volatile float r = a - b; return r; }
CS_HD float fdiv(float a, float b) { volatile float r = a / b; return r; }
#endif

CS_HD double bicubic(double x) {
    // Resample.c bicubic_filter with a = -0.5:  ((a+2) x - (a+3)) x x + 1   |   (((x-5) x + 8) x - 4) a
    x = x < 0.0 ? -x : x;
    if (x < 1.0) return dadd(dmul(dmul(dsub(dmul(1.5, x), 2.5), x), x), 1.0);
    if (x < 2.0) return dmul(dsub(dmul(dadd(dmul(dsub(x, 5.0), x), 8.0), x), 4.0), -0.5);
    return 0.0;
}

// precompute_coeffs + normalize_coeffs_8bpc for output index xx of one (crop, axis):
// b = bounds row [size][2] (first source index, taps), kbase = coefficient rows [size][ksize_max]
CS_HD void coeffs_one(int in_size, int out_size, int xx, int ksize_max, int* b, int* kbase) {
    const double scale = ddiv((double)in_size, (double)out_size);        // (double)(in1 - in0) / outSize, in0 = 0
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = dmul(2.0, filterscale);
    const double ss = ddiv(1.0, filterscale);
    const double center = dmul(dadd((double)xx, 0.5), scale);            // in0 + (xx + 0.5) * scale with in0 = 0.0
    int xmin = (int)dadd(dsub(center, support), 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)dadd(dadd(center, support), 0.5);
    if (xmax > in_size) xmax = in_size;
    const int n = xmax - xmin;
    int* ko = kbase + (long long)xx * ksize_max;
    double ww = 0.0;
    for (int x = 0; x < n; ++x) ww = dadd(ww, bicubic(dmul(dadd(dsub((double)(x + xmin), center), 0.5), ss)));
    for (int x = 0; x < ksize_max; ++x) {
        int ki = 0;
        if (x < n) {
            double w = bicubic(dmul(dadd(dsub((double)(x + xmin), center), 0.5), ss));
            if (ww != 0.0) w = ddiv(w, ww);
            const double fx = dmul(w, (double)(1 << PRECISION_BITS));
            ki = w < 0.0 ? (int)dadd(-0.5, fx) : (int)dadd(0.5, fx);     // C cast: toward zero
        }
        ko[x] = ki;
    }
    b[xx * 2 + 0] = xmin;
    b[xx * 2 + 1] = n;
}

CS_HD int clip8(int acc) {
    const int v = acc >> PRECISION_BITS;          // arithmetic shift, like the C code
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// one byte of the horizontal pass: intermediate row tr (source row ybox_first + tr of the crop), column xx, channel c
CS_HD uint8_t horizontal_one(const uint8_t* image, int H, int W, const Desc& d, const int* bh, const int* kh, int ksize_max,
                             int ybox_first, int tr, int xx, int c) {
    const int xmin = bh[xx * 2], n = bh[xx * 2 + 1];
    const int* ko = kh + (long long)xx * ksize_max;
    const int sy = d.y0 + ybox_first + tr;                            // image row
    int acc = 1 << (PRECISION_BITS - 1);
    if (sy >= 0 && sy < H) {
        const uint8_t* srow = image + ((long long)sy * W) * 3 + c;
        for (int j = 0; j < n; ++j) {
            const int sx = d.x0 + xmin + j;
            const int p = (sx >= 0 && sx < W) ? srow[(long long)sx * 3] : 0;     // Image.crop zero-fills outside
            acc += p * ko[j];
        }
    }
    return (uint8_t)clip8(acc);
}

// one float of the output canvas [3][size][size]: vertical pass + pad + ToTensor + Normalize
CS_HD float vertical_one(const Desc& d, bool empty, const int* bv, const int* kv, int ksize_max, int ybox_first,
                         const uint8_t* t, int c, int y, int x, float mean, float stdv) {
    const int xx = x - d.pad_left, yy = y - d.pad_top;
    int v = 0;                                                        // padding: fill = 0 before ToTensor
    if (!empty && xx >= 0 && xx < d.out_w && yy >= 0 && yy < d.out_h) {
        const int ymin = bv[yy * 2] - ybox_first, n = bv[yy * 2 + 1];
        const int* ko = kv + (long long)yy * ksize_max;
        int acc = 1 << (PRECISION_BITS - 1);
        for (int j = 0; j < n; ++j) acc += (int)t[((long long)(ymin + j) * d.out_w + xx) * 3 + c] * ko[j];
        v = clip8(acc);
    }
    return fdiv(fsub(fdiv((float)v, 255.0f), mean), stdv);             // ToTensor (/255) then Normalize, f32
}

CS_HD bool desc_empty(const Desc& d) { return d.out_w <= 0 || d.out_h <= 0 || d.x1 <= d.x0 || d.y1 <= d.y0; }

}  // namespace crops
}  // namespace cs
