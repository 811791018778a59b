// On-device crop generation, bit-exact on a B200 against the CPU oracle
// (oracle/crops_oracle.py, pinned to Pillow and to the reference's transform objects) and the reference-generated
// fixtures (tests/test_gpu_region.py).
//
// Replaces the CPU PIL path of the distill datasets: GridDistillDataset._obtain_image_crops
// (src/training/data.py:226-245: image.crop(box) -> transforms[1]) with
// transforms[1] = ResizeMaxSize(s, BICUBIC) + pad + ToTensor + Normalize (src/open_clip/transform.py:26-49,
// 119-133) and, with pad_left = pad_top = 0, the detector-image transform ResizeLongest (transform.py:169-191).
// The resampler restates Pillow's ImagingResample (libImaging/Resample.c): separable, HORIZONTAL pass first into
// a uint8 intermediate, then the VERTICAL pass; weights from the a = -0.5 bicubic kernel over a support of
// 2 * max(in/out, 1) pixels, normalised in double precision, converted to 22-bit fixed point, accumulator
// seeded with 2^21, result clip(acc >> 22).  All double / float steps use round-to-nearest intrinsics without
// FMA contraction so they round exactly like the C library on the host.
//
// Integer / byte work, HBM bound: per crop the source rectangle is read once (through L2 for the taps), the
// uint8 intermediate (rows x out_w x 3) is written and read once, the f32 [3, s, s] crop is written once.
#include "common.cuh"
#include "crops_core.cuh"

namespace cs {
namespace crops {

// coefficient tables: bounds [K][2][size][2] (first source index, taps), kk [K][2][size][ksize_max] int32
__global__ void crop_coeffs_kernel(const Desc* __restrict__ descs, int size, int ksize_max, int* __restrict__ bounds,
                                   int* __restrict__ kk) {
    const int k = blockIdx.x, axis = blockIdx.y;
    const Desc d = descs[k];
    const int in_size = axis == 0 ? d.x1 - d.x0 : d.y1 - d.y0;
    const int out_size = axis == 0 ? d.out_w : d.out_h;
    int* b = bounds + ((long long)(k * 2 + axis) * size) * 2;
    int* kbase = kk + ((long long)(k * 2 + axis) * size) * ksize_max;
    if (in_size <= 0 || out_size <= 0) return;
    for (int xx = threadIdx.x; xx < out_size; xx += blockDim.x) coeffs_one(in_size, out_size, xx, ksize_max, b, kbase);
}

// horizontal pass: tmp[k][t][xx][c] for source rows ybox_first + t;  grid (row chunks, K)
// A batch of crops may come from several images packed in one uint8 blob: desc_image[k] selects the image of crop k,
// image_offsets / image_hw its byte offset and (H, W).  desc_image == nullptr: one image (H, W) at `image`.
__global__ void crop_horizontal_kernel(const uint8_t* __restrict__ image, int H, int W, const int* __restrict__ desc_image,
                                       const long long* __restrict__ image_offsets, const int* __restrict__ image_hw,
                                       const Desc* __restrict__ descs,
                                       int size, int ksize_max, int tmp_rows_max, const int* __restrict__ bounds,
                                       const int* __restrict__ kk, uint8_t* __restrict__ tmp) {
    const int k = blockIdx.y;
    const Desc d = descs[k];
    if (desc_empty(d)) return;
    if (desc_image != nullptr) {
        const int im = desc_image[k];
        image += image_offsets[im];
        H = image_hw[2 * im];
        W = image_hw[2 * im + 1];
    }
    const int* bh = bounds + ((long long)(k * 2 + 0) * size) * 2;
    const int* bv = bounds + ((long long)(k * 2 + 1) * size) * 2;
    const int* kh = kk + ((long long)(k * 2 + 0) * size) * ksize_max;
    const int ybox_first = bv[0];
    const int ybox_last = bv[(d.out_h - 1) * 2] + bv[(d.out_h - 1) * 2 + 1];
    const int rows = ybox_last - ybox_first;
    uint8_t* t = tmp + (long long)k * tmp_rows_max * size * 3;
    const int total = rows * d.out_w * 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i % 3, xx = (i / 3) % d.out_w, tr = i / (3 * d.out_w);
        t[i] = horizontal_one(image, H, W, d, bh, kh, ksize_max, ybox_first, tr, xx, c);
    }
}

// vertical pass + pad + ToTensor + Normalize: out[k][c][y][x] f32, every pixel of the s x s canvas written
__global__ void crop_vertical_kernel(const Desc* __restrict__ descs, int size, int ksize_max, int tmp_rows_max,
                                     const int* __restrict__ bounds, const int* __restrict__ kk,
                                     const uint8_t* __restrict__ tmp, float3 mean, float3 stdv, float* __restrict__ out) {
    const int k = blockIdx.y;
    const Desc d = descs[k];
    const bool empty = desc_empty(d);
    const int* bv = bounds + ((long long)(k * 2 + 1) * size) * 2;
    const int* kv = kk + ((long long)(k * 2 + 1) * size) * ksize_max;
    const int ybox_first = empty ? 0 : bv[0];
    const uint8_t* t = tmp + (long long)k * tmp_rows_max * size * 3;
    float* o = out + (long long)k * 3 * size * size;
    const int total = 3 * size * size;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int x = i % size, y = (i / size) % size, c = i / (size * size);
        const float m = c == 0 ? mean.x : (c == 1 ? mean.y : mean.z);
        const float s = c == 0 ? stdv.x : (c == 1 ? stdv.y : stdv.z);
        o[i] = vertical_one(d, empty, bv, kv, ksize_max, ybox_first, t, c, y, x, m, s);
    }
}

}  // namespace crops
}  // namespace cs

using namespace cs;
using namespace cs::crops;

extern "C" int cs_crop_workspace_bytes(int K, int size, int ksize_max, int tmp_rows_max, int64_t* bytes) {
    CS_CHECK_ARG(bytes && K >= 0 && size > 0 && ksize_max > 0 && tmp_rows_max >= 0, "cs_crop_workspace_bytes: bad argument");
    const int64_t coeff = (int64_t)K * 2 * size * ((int64_t)ksize_max + 2) * 4;
    const int64_t tmp = ((int64_t)K * tmp_rows_max * size * 3 + 15) / 16 * 16;
    *bytes = coeff + tmp;
    return CS_OK;
}

static int crop_launch(const uint8_t* image, int H, int W, const int* desc_image, const long long* image_offsets,
                       const int* image_hw, const void* descs, int K, int size, int ksize_max, int tmp_rows_max,
                       const float* mean3, const float* std3, float* out, void* workspace, int64_t workspace_bytes,
                       cudaStream_t st) {
    int64_t need = 0;
    cs_crop_workspace_bytes(K, size, ksize_max, tmp_rows_max, &need);
    CS_CHECK_ARG(workspace_bytes >= need, "cs_crop_resize_normalize: workspace too small");
    if (K == 0) return CS_OK;
    int* bounds = (int*)workspace;
    int* kk = bounds + (int64_t)K * 2 * size * 2;
    uint8_t* tmp = (uint8_t*)(kk + (int64_t)K * 2 * size * ksize_max);
    const Desc* d = (const Desc*)descs;
    crop_coeffs_kernel<<<dim3(K, 2), 256, 0, st>>>(d, size, ksize_max, bounds, kk);
    CS_LAUNCH_CHECK();
    const int hblocks = ceil_div((int64_t)tmp_rows_max * size * 3, 256 * 4);
    crop_horizontal_kernel<<<dim3(hblocks > 0 ? hblocks : 1, K), 256, 0, st>>>(image, H, W, desc_image, image_offsets, image_hw, d,
                                                                              size, ksize_max, tmp_rows_max, bounds, kk, tmp);
    CS_LAUNCH_CHECK();
    const float3 mean = make_float3(mean3[0], mean3[1], mean3[2]);
    const float3 stdv = make_float3(std3[0], std3[1], std3[2]);
    crop_vertical_kernel<<<dim3(ceil_div((int64_t)3 * size * size, 256 * 4), K), 256, 0, st>>>(d, size, ksize_max, tmp_rows_max,
                                                                                               bounds, kk, tmp, mean, stdv, out);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

extern "C" int cs_crop_resize_normalize(const uint8_t* image_hwc, int H, int W, const void* descs, int K, int size,
                                        int ksize_max, int tmp_rows_max, const float* mean3, const float* std3,
                                        float* out, void* workspace, int64_t workspace_bytes, void* stream) {
    CS_CHECK_ARG(image_hwc && descs && mean3 && std3 && out && workspace, "cs_crop_resize_normalize: null pointer");
    CS_CHECK_ARG(H > 0 && W > 0 && K >= 0 && size > 0 && ksize_max > 0 && tmp_rows_max >= 0, "cs_crop_resize_normalize: bad shape");
    return crop_launch(image_hwc, H, W, nullptr, nullptr, nullptr, descs, K, size, ksize_max, tmp_rows_max, mean3, std3, out,
                       workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int cs_crop_resize_normalize_batched(const uint8_t* images_blob, const int64_t* image_offsets, const int32_t* image_hw,
                                                const int32_t* desc_image, const void* descs, int K, int size, int ksize_max,
                                                int tmp_rows_max, const float* mean3, const float* std3, float* out,
                                                void* workspace, int64_t workspace_bytes, void* stream) {
    CS_CHECK_ARG(images_blob && image_offsets && image_hw && desc_image && descs && mean3 && std3 && out && workspace,
                 "cs_crop_resize_normalize_batched: null pointer");
    CS_CHECK_ARG(K >= 0 && size > 0 && ksize_max > 0 && tmp_rows_max >= 0, "cs_crop_resize_normalize_batched: bad shape");
    static_assert(sizeof(long long) == sizeof(int64_t), "offset table type");
    return crop_launch(images_blob, 0, 0, desc_image, reinterpret_cast<const long long*>(image_offsets), image_hw, descs, K, size,
                       ksize_max, tmp_rows_max, mean3, std3, out, workspace, workspace_bytes, (cudaStream_t)stream);
}
