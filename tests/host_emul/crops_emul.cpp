// Host-side emulation of csrc/crops.cu for the CPU tests: the SAME core functions (crops_core.cuh) driven by
// plain loops in the kernels' index order.  Test infrastructure only (built by tests/test_crops_emulation.py with
// g++ into a scratch directory); it is not linked into libclipself_b200.so and nothing in the product calls it.
#include <cstring>
#include <vector>

#include "../../clipself_b200/csrc/crops_core.cuh"

using namespace cs::crops;

extern "C" int crops_emulate(const uint8_t* image, int H, int W, const int32_t* descs_i32, int K, int size, int ksize_max,
                             int tmp_rows_max, const float* mean3, const float* std3, float* out) {
    const Desc* descs = reinterpret_cast<const Desc*>(descs_i32);
    std::vector<int> bounds((size_t)K * 2 * size * 2, 0), kk((size_t)K * 2 * size * ksize_max, 0);
    std::vector<uint8_t> tmp((size_t)K * tmp_rows_max * size * 3 + 16, 0);
    for (int k = 0; k < K; ++k)
        for (int axis = 0; axis < 2; ++axis) {                                    // crop_coeffs_kernel
            const Desc d = descs[k];
            const int in_size = axis == 0 ? d.x1 - d.x0 : d.y1 - d.y0;
            const int out_size = axis == 0 ? d.out_w : d.out_h;
            if (in_size <= 0 || out_size <= 0) continue;
            if (out_size > size) return 1;
            int* b = bounds.data() + ((long long)(k * 2 + axis) * size) * 2;
            int* kbase = kk.data() + ((long long)(k * 2 + axis) * size) * ksize_max;
            for (int xx = 0; xx < out_size; ++xx) {
                coeffs_one(in_size, out_size, xx, ksize_max, b, kbase);
                if (b[xx * 2 + 1] > ksize_max) return 2;                          // the host-side bound must hold
            }
        }
    for (int k = 0; k < K; ++k) {                                                 // crop_horizontal_kernel
        const Desc d = descs[k];
        if (desc_empty(d)) continue;
        const int* bh = bounds.data() + ((long long)(k * 2 + 0) * size) * 2;
        const int* bv = bounds.data() + ((long long)(k * 2 + 1) * size) * 2;
        const int* kh = kk.data() + ((long long)(k * 2 + 0) * size) * ksize_max;
        const int ybox_first = bv[0];
        const int rows = bv[(d.out_h - 1) * 2] + bv[(d.out_h - 1) * 2 + 1] - ybox_first;
        if (rows > tmp_rows_max) return 3;
        uint8_t* t = tmp.data() + (long long)k * tmp_rows_max * size * 3;
        for (int i = 0; i < rows * d.out_w * 3; ++i) {
            const int c = i % 3, xx = (i / 3) % d.out_w, tr = i / (3 * d.out_w);
            t[i] = horizontal_one(image, H, W, d, bh, kh, ksize_max, ybox_first, tr, xx, c);
        }
    }
    for (int k = 0; k < K; ++k) {                                                 // crop_vertical_kernel
        const Desc d = descs[k];
        const bool empty = desc_empty(d);
        const int* bv = bounds.data() + ((long long)(k * 2 + 1) * size) * 2;
        const int* kv = kk.data() + ((long long)(k * 2 + 1) * size) * ksize_max;
        const int ybox_first = empty ? 0 : bv[0];
        const uint8_t* t = tmp.data() + (long long)k * tmp_rows_max * size * 3;
        float* o = out + (long long)k * 3 * size * size;
        for (int i = 0; i < 3 * size * size; ++i) {
            const int x = i % size, y = (i / size) % size, c = i / (size * size);
            o[i] = vertical_one(d, empty, bv, kv, ksize_max, ybox_first, t, c, y, x, mean3[c], std3[c]);
        }
    }
    return 0;
}
