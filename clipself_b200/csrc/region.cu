// Region path of the CLIPSelf step: index extraction, RoIAlign (1x1, adaptive, aligned) forward /
// backward on NHWC maps, mask pooling, L2-normalise + cosine loss.  HBM-bound CUDA-core kernels:
// every global access is coalesced over the channel axis; each feature map is read from HBM once
// (staged in shared memory per image x channel-slice) and box results are written once.
#include "common.cuh"

namespace cs {
namespace region {

// ------------------------------------------------------------------------------------------
// Index extraction (clipself.py:29-36): stable, image-major compaction of rows with box[4] > 0.5
// ------------------------------------------------------------------------------------------
__global__ void extract_rois_kernel(const float* __restrict__ boxes, int B, int K, float* __restrict__ rois,
                                    int* __restrict__ crop_index, int* __restrict__ roi_batch,
                                    int* __restrict__ img_offsets) {
    extern __shared__ int s_off[];   // B + 1
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        int n = 0;
        for (int k = 0; k < K; ++k) n += boxes[((long long)b * K + k) * 5 + 4] > 0.5f;
        s_off[b + 1] = n;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        s_off[0] = 0;
        for (int b = 0; b < B; ++b) s_off[b + 1] += s_off[b];
    }
    __syncthreads();
    for (int b = threadIdx.x; b <= B; b += blockDim.x) img_offsets[b] = s_off[b];
    for (int i = threadIdx.x; i < B * K; i += blockDim.x) {
        const int b = i / K, k = i % K;
        const float* row = boxes + (long long)i * 5;
        if (!(row[4] > 0.5f)) continue;
        int before = 0;
        for (int j = 0; j < k; ++j) before += boxes[((long long)b * K + j) * 5 + 4] > 0.5f;
        const int dst = s_off[b] + before;
        rois[dst * 4 + 0] = row[0];
        rois[dst * 4 + 1] = row[1];
        rois[dst * 4 + 2] = row[2];
        rois[dst * 4 + 3] = row[3];
        crop_index[dst] = i;
        roi_batch[dst] = b;
    }
    for (int i = s_off[B] + threadIdx.x; i < B * K; i += blockDim.x) {     // rows past the valid count: zeros
        rois[i * 4 + 0] = rois[i * 4 + 1] = rois[i * 4 + 2] = rois[i * 4 + 3] = 0.f;
        crop_index[i] = 0;
        roi_batch[i] = 0;
    }
}

__global__ void gather_rows_kernel(const uint4* __restrict__ src, const int* __restrict__ index,
                                   long long row_vecs, uint4* __restrict__ dst) {
    const long long r = blockIdx.y;
    const uint4* s = src + (long long)index[r] * row_vecs;
    uint4* d = dst + r * row_vecs;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < row_vecs;
         i += (long long)gridDim.x * blockDim.x)
        d[i] = s[i];
}

// ------------------------------------------------------------------------------------------
// RoIAlign separable weights.  For output 1x1 / aligned=True the sample grid is a tensor product
// so the pooled value is  sum_y sum_x Wy[y] Wx[x] f[y,x]  with Wy/Wx accumulated over the 1-D
// sample positions (torchvision semantics restated in oracle/clipself_oracle.py).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void axis_weights(float lo_n, float hi_n, int size, float* __restrict__ w,
                                             int* count_out) {
    // denormalise exactly like eva_vit_model.py:660-662 (f32 multiply), then torchvision's
    // aligned=True offset.
    // (explicit _rn intrinsics: no FMA contraction, so the sample positions round exactly like
    //  the separate multiply / subtract kernels of the reference)
    const float lo = __fmul_rn(lo_n, (float)size);
    const float hi = __fmul_rn(hi_n, (float)size);
    const float start = __fsub_rn(lo, 0.5f);
    const float end = __fsub_rn(hi, 0.5f);
    const float roi = __fsub_rn(end, start);
    const int grid = (int)ceilf(roi);
    for (int i = 0; i < size; ++i) w[i] = 0.f;
    for (int i = 0; i < grid; ++i) {
        float p = __fadd_rn(start, __fdiv_rn(__fmul_rn((float)i + 0.5f, roi), (float)grid));
        if (p < -1.0f || p > (float)size) continue;
        if (p <= 0.f) p = 0.f;
        int lo_i = (int)p, hi_i;
        if (lo_i >= size - 1) {
            hi_i = lo_i = size - 1;
            p = (float)lo_i;
        } else {
            hi_i = lo_i + 1;
        }
        const float l = __fsub_rn(p, (float)lo_i);
        const float h = __fsub_rn(1.0f, l);
        w[lo_i] += h;
        w[hi_i] += l;
    }
    *count_out = grid;
}

// 32 boxes per CTA, one thread per box; the weight rows are accumulated in shared memory (the scatter
// `w[lo] += h` is a read-modify-write chain) and written back coalesced: the CTA's rows are contiguous.
constexpr int RW_BOXES = 32;
__global__ void __launch_bounds__(RW_BOXES)
roi_weights_kernel(const float* __restrict__ rois, int R, int H, int W, float* __restrict__ wy,
                   float* __restrict__ wx) {
    extern __shared__ float s_w[];                   // [RW_BOXES][W] then [RW_BOXES][H]
    float* s_wx = s_w;
    float* s_wy = s_w + RW_BOXES * W;
    const int r0 = blockIdx.x * RW_BOXES;
    const int r = r0 + threadIdx.x;
    const int n = min(RW_BOXES, R - r0);
    if (r < R) {
        const float4 b = *reinterpret_cast<const float4*>(rois + (long long)r * 4);
        int gw, gh;
        float* mx = s_wx + threadIdx.x * W;
        float* my = s_wy + threadIdx.x * H;
        axis_weights(b.x, b.z, W, mx, &gw);
        axis_weights(b.y, b.w, H, my, &gh);
        const float count = (float)max(gw * gh, 1);
        const float inv = 1.0f / count;
        for (int i = 0; i < H; ++i) my[i] *= inv;
    }
    __syncwarp();
    for (int i = threadIdx.x; i < n * W; i += RW_BOXES) wx[(long long)r0 * W + i] = s_wx[i];
    for (int i = threadIdx.x; i < n * H; i += RW_BOXES) wy[(long long)r0 * H + i] = s_wy[i];
}

constexpr int ROI_CS = 64;       // channel slice per CTA
constexpr int ROI_LANES = 8;     // roi (or pixel) lanes per CTA (each lane = two warps of 32 channels)
constexpr int ROI_THREADS = ROI_CS * ROI_LANES;
constexpr int ROI_CHUNK = 64;    // rois whose separable weights are staged in shared memory at once

// first / last index with a non-zero weight (the support of a box along one axis is contiguous).
// Warp-cooperative: the 32 threads of a warp always work on the same box (they are 32 channels of it).
__device__ __forceinline__ void support(const float* __restrict__ w, int n, int& lo, int& hi) {
    const int l = threadIdx.x & 31;
    lo = n;
    hi = -1;
    for (int base = 0; base < n; base += 32) {
        const unsigned m = __ballot_sync(0xffffffffu, base + l < n && w[base + l] != 0.f);
        if (m) {
            if (lo == n) lo = base + __ffs(m) - 1;
            hi = base + 31 - __clz(m);
        }
    }
}

// out[r, c] = sum_{y,x} wy[r,y] wx[r,x] f[b,y,x,c];  grid (C/64, B).
// The image's map slice [H*W][64] is staged in shared memory once (kStage) — every map byte crosses HBM
// exactly once — together with the separable weights of up to 64 of its boxes; each box then only walks
// the pixels of its support window.
template <bool kStage>
__global__ void __launch_bounds__(ROI_THREADS)
roi_align_fwd_kernel(const float* __restrict__ fmap, int H, int W, int C, const int* __restrict__ img_offsets,
                     const float* __restrict__ wy, const float* __restrict__ wx, float* __restrict__ out) {
    extern __shared__ float s_dyn[];
    float* s_wy = s_dyn;                             // [ROI_CHUNK][H]
    float* s_wx = s_wy + ROI_CHUNK * H;              // [ROI_CHUNK][W]
    float* s_map = s_wx + ROI_CHUNK * W;             // [H*W][ROI_CS] when staged
    const int b = blockIdx.y;
    const int c0 = blockIdx.x * ROI_CS;
    const int c = threadIdx.x % ROI_CS;
    const int lane = threadIdx.x / ROI_CS;
    const int begin = img_offsets[b], end = img_offsets[b + 1];
    if (begin == end) return;
    const float* g = fmap + (long long)b * H * W * C + c0;
    const bool c_ok = c0 + c < C;
    if (kStage) {
        for (int i = threadIdx.x; i < H * W * (ROI_CS / 4); i += ROI_THREADS) {      // float4 per thread
            const int p = i / (ROI_CS / 4), c4 = (i % (ROI_CS / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c0 + c4 + 3 < C) v = *reinterpret_cast<const float4*>(g + (long long)p * C + c4);
            else
                for (int z = 0; z < 4; ++z)
                    if (c0 + c4 + z < C) (&v.x)[z] = g[(long long)p * C + c4 + z];
            *reinterpret_cast<float4*>(s_map + p * ROI_CS + c4) = v;
        }
    }
    for (int r0 = begin; r0 < end; r0 += ROI_CHUNK) {
        const int n = min(ROI_CHUNK, end - r0);
        __syncthreads();
        for (int i = threadIdx.x; i < n * H; i += ROI_THREADS) s_wy[i] = wy[(long long)r0 * H + i];
        for (int i = threadIdx.x; i < n * W; i += ROI_THREADS) s_wx[i] = wx[(long long)r0 * W + i];
        __syncthreads();
        for (int rr = lane; rr < n; rr += ROI_LANES) {
            const float* wyr = s_wy + rr * H;
            const float* wxr = s_wx + rr * W;
            int ylo, yhi, xlo, xhi;
            support(wyr, H, ylo, yhi);
            support(wxr, W, xlo, xhi);
            float acc = 0.f;
            for (int y = ylo; y <= yhi; ++y) {
                float row = 0.f;
                for (int x = xlo; x <= xhi; ++x) {
                    const float f = kStage ? s_map[(y * W + x) * ROI_CS + c] : (c_ok ? g[(long long)(y * W + x) * C + c] : 0.f);
                    row = fmaf(wxr[x], f, row);
                }
                acc = fmaf(wyr[y], row, acc);
            }
            if (c_ok) out[(long long)(r0 + rr) * C + c0 + c] = acc;
        }
    }
}

// d_fmap[b,y,x,c] = sum_{r in image b} wy[r,y] wx[r,x] d_out[r,c];  grid (C/64, B).
// The image's gradient slice [H*W][64] is accumulated in shared memory, box after box in index order
// (deterministic, no atomics: inside one box every (pixel, channel) belongs to one thread), and written
// to HBM once.  Per chunk of 64 boxes the weights, the d_out rows and the support windows are staged /
// precomputed in shared memory, so the serial per-box loop touches no global memory.
template <bool kStage>
__global__ void __launch_bounds__(ROI_THREADS)
roi_align_bwd_kernel(const float* __restrict__ d_out, int H, int W, int C, const int* __restrict__ img_offsets,
                     const float* __restrict__ wy, const float* __restrict__ wx, float* __restrict__ d_fmap) {
    extern __shared__ float s_dyn[];
    float* s_wy = s_dyn;                             // [ROI_CHUNK][H]
    float* s_wx = s_wy + ROI_CHUNK * H;              // [ROI_CHUNK][W]
    float* s_go = s_wx + ROI_CHUNK * W;              // [ROI_CHUNK][ROI_CS]
    int4* s_win = reinterpret_cast<int4*>(s_go + ROI_CHUNK * ROI_CS);   // [ROI_CHUNK] (ylo, yhi, xlo, xhi)
    float* s_acc = reinterpret_cast<float*>(s_win + ROI_CHUNK);         // [H*W][ROI_CS] when staged
    const int b = blockIdx.y;
    const int c0 = blockIdx.x * ROI_CS;
    const int c = threadIdx.x % ROI_CS;
    const int lane = threadIdx.x / ROI_CS;
    const int warp = threadIdx.x / 32;
    const int begin = img_offsets[b], end = img_offsets[b + 1];
    const bool c_ok = c0 + c < C;
    float* g = d_fmap + (long long)b * H * W * C + c0;
    const int HW = H * W;
    if (kStage) {
        for (int i = threadIdx.x; i < HW * ROI_CS; i += ROI_THREADS) s_acc[i] = 0.f;
    } else {
        for (int p = lane; p < HW; p += ROI_LANES)
            if (c_ok) g[(long long)p * C + c] = 0.f;
    }
    for (int r0 = begin; r0 < end; r0 += ROI_CHUNK) {
        const int n = min(ROI_CHUNK, end - r0);
        __syncthreads();
        for (int i = threadIdx.x; i < n * H; i += ROI_THREADS) s_wy[i] = wy[(long long)r0 * H + i];
        for (int i = threadIdx.x; i < n * W; i += ROI_THREADS) s_wx[i] = wx[(long long)r0 * W + i];
        for (int i = threadIdx.x; i < n * ROI_CS; i += ROI_THREADS) {
            const int rr = i / ROI_CS, cc = i % ROI_CS;
            s_go[i] = (c0 + cc < C) ? d_out[(long long)(r0 + rr) * C + c0 + cc] : 0.f;
        }
        __syncthreads();
        for (int rr = warp; rr < n; rr += ROI_THREADS / 32) {       // one warp per box: support windows
            int ylo, yhi, xlo, xhi;
            support(s_wy + rr * H, H, ylo, yhi);
            support(s_wx + rr * W, W, xlo, xhi);
            if ((threadIdx.x & 31) == 0) s_win[rr] = make_int4(ylo, yhi, xlo, xhi);
        }
        __syncthreads();
        for (int rr = 0; rr < n; ++rr) {            // boxes in order; the lanes split the window's pixels
            const int4 win = s_win[rr];
            const int ww = max(win.w - win.z + 1, 0);               // degenerate boxes have an empty support
            const int npx = max(win.y - win.x + 1, 0) * ww;
            if (npx == 0) continue;
            const float* wyr = s_wy + rr * H;
            const float* wxr = s_wx + rr * W;
            const float go = s_go[rr * ROI_CS + c];
            for (int q = lane; q < npx; q += ROI_LANES) {
                const int y = win.x + q / ww, x = win.z + q % ww;
                const float v = wyr[y] * wxr[x] * go;
                if (kStage) s_acc[(y * W + x) * ROI_CS + c] += v;
                else if (c_ok) g[(long long)(y * W + x) * C + c] += v;
            }
            __syncthreads();                        // the next box may overlap this one's pixels (other lanes)
        }
    }
    if (kStage) {
        __syncthreads();
        for (int i = threadIdx.x; i < HW * (ROI_CS / 4); i += ROI_THREADS) {
            const int p = i / (ROI_CS / 4), c4 = (i % (ROI_CS / 4)) * 4;
            const float4 v = *reinterpret_cast<const float4*>(s_acc + p * ROI_CS + c4);
            if (c0 + c4 + 3 < C) *reinterpret_cast<float4*>(g + (long long)p * C + c4) = v;
            else
                for (int z = 0; z < 4; ++z)
                    if (c0 + c4 + z < C) g[(long long)p * C + c4 + z] = (&v.x)[z];
        }
    }
}

// out[r,c] = sum_p m[r,p] f[b,p,c] / (sum_p m[r,p] + 1e-12);  grid (C/128, B).
// A small [boxes x pixels] x [pixels x channels] product on the CUDA cores: pixel chunks of the map slice
// and of the image's masks are staged in shared memory; every thread keeps an 8-box x 4-channel register
// tile (one LDS.128 of the map and 8 broadcast mask reads feed 32 FMAs).
constexpr int MP_CS = 128;       // channels per CTA (32 threads x float4)
constexpr int MP_PX = 32;        // pixels per staged chunk
constexpr int MP_ACC = 8;        // boxes per thread per pass
constexpr int MP_LANES = 8;      // box lanes (warps) per CTA -> 64 boxes per pass
constexpr int MP_THREADS = 32 * MP_LANES;
__global__ void __launch_bounds__(MP_THREADS)
mask_pool_kernel(const float* __restrict__ fmap, int HW, int C, const float* __restrict__ masks,
                 const int* __restrict__ img_offsets, float* __restrict__ out) {
    __shared__ __align__(16) float s_f[MP_PX][MP_CS];
    __shared__ float s_m[MP_ACC * MP_LANES][MP_PX + 1];
    const int b = blockIdx.y;
    const int c0 = blockIdx.x * MP_CS;
    const int ct = threadIdx.x % 32;                 // channel quad
    const int lane = threadIdx.x / 32;               // = warp: all 32 threads read the same mask value
    const int begin = img_offsets[b], end = img_offsets[b + 1];
    const float* g = fmap + (long long)b * HW * C + c0;
    for (int r0 = begin; r0 < end; r0 += MP_ACC * MP_LANES) {
        const int n = min(MP_ACC * MP_LANES, end - r0);
        float acc[MP_ACC][4], msum[MP_ACC];
#pragma unroll
        for (int k = 0; k < MP_ACC; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = msum[k] = 0.f;
        for (int p0 = 0; p0 < HW; p0 += MP_PX) {
            const int np = min(MP_PX, HW - p0);
            __syncthreads();
            for (int i = threadIdx.x; i < np * (MP_CS / 4); i += MP_THREADS) {
                const int p = i / (MP_CS / 4), c4 = (i % (MP_CS / 4)) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c0 + c4 + 3 < C) v = *reinterpret_cast<const float4*>(g + (long long)(p0 + p) * C + c4);
                else
                    for (int z = 0; z < 4; ++z)
                        if (c0 + c4 + z < C) (&v.x)[z] = g[(long long)(p0 + p) * C + c4 + z];
                *reinterpret_cast<float4*>(&s_f[p][c4]) = v;
            }
            for (int i = threadIdx.x; i < MP_ACC * MP_LANES * np; i += MP_THREADS) {
                const int rr = i / np, p = i % np;
                s_m[rr][p] = rr < n ? masks[(long long)(r0 + rr) * HW + p0 + p] : 0.f;
            }
            __syncthreads();
            for (int p = 0; p < np; ++p) {
                const float4 f = *reinterpret_cast<const float4*>(&s_f[p][ct * 4]);
#pragma unroll
                for (int k = 0; k < MP_ACC; ++k) {
                    const float w = s_m[k * MP_LANES + lane][p];
                    acc[k][0] = fmaf(w, f.x, acc[k][0]);
                    acc[k][1] = fmaf(w, f.y, acc[k][1]);
                    acc[k][2] = fmaf(w, f.z, acc[k][2]);
                    acc[k][3] = fmaf(w, f.w, acc[k][3]);
                    msum[k] += w;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < MP_ACC; ++k) {
            const int rr = k * MP_LANES + lane;
            if (rr >= n) continue;
            const float inv = 1.f / (msum[k] + 1e-12f);
            float* o = out + (long long)(r0 + rr) * C + c0 + ct * 4;
            for (int z = 0; z < 4; ++z)
                if (c0 + ct * 4 + z < C) o[z] = acc[k][z] * inv;
        }
    }
}

// ------------------------------------------------------------------------------------------
// cosine loss (clipself.py:42-47) and row L2 normalisation (eva_vit_model.py:620)
// ------------------------------------------------------------------------------------------
__global__ void cosine_rows_kernel(const float* __restrict__ s, const float* __restrict__ t, int R, int C,
                                   float* __restrict__ stats) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const float4* sp = reinterpret_cast<const float4*>(s + (long long)r * C);
    const float4* tp = reinterpret_cast<const float4*>(t + (long long)r * C);
    float ss = 0.f, tt = 0.f, st = 0.f;
    for (int i = lane; i < C / 4; i += 32) {
        const float4 a = sp[i], b = tp[i];
        ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
        tt += b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
        st += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    }
    ss = warp_sum(ss);
    tt = warp_sum(tt);
    st = warp_sum(st);
    if (lane == 0) {
        const float is = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
        const float it = 1.0f / fmaxf(sqrtf(tt), 1e-12f);
        stats[r * 3 + 0] = is;
        stats[r * 3 + 1] = it;
        stats[r * 3 + 2] = st * is * it;
    }
}

// single CTA, fixed-order tree: loss = (1 - sum(cos)/R) * weight
__global__ void cosine_reduce_kernel(const float* __restrict__ stats, int R, float weight, float* __restrict__ loss) {
    __shared__ float s_part[1024];
    float acc = 0.f;
    for (int r = threadIdx.x; r < R; r += blockDim.x) acc += stats[r * 3 + 2];
    s_part[threadIdx.x] = acc;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if (threadIdx.x < o) s_part[threadIdx.x] += s_part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = (1.0f - s_part[0] / (float)R) * weight;
}

__global__ void cosine_bwd_kernel(const float* __restrict__ s, const float* __restrict__ t,
                                  const float* __restrict__ stats, int R, int C, float weight,
                                  const float* __restrict__ d_loss, float* __restrict__ d_s) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const float is = stats[r * 3 + 0], it = stats[r * 3 + 1], cs_ = stats[r * 3 + 2];
    const float k = -d_loss[0] * weight / (float)R;
    const float4* sp = reinterpret_cast<const float4*>(s + (long long)r * C);
    const float4* tp = reinterpret_cast<const float4*>(t + (long long)r * C);
    float4* dp = reinterpret_cast<float4*>(d_s + (long long)r * C);
    for (int i = lane; i < C / 4; i += 32) {
        const float4 a = sp[i], b = tp[i];
        float4 d;
        d.x = k * is * (b.x * it - cs_ * a.x * is);
        d.y = k * is * (b.y * it - cs_ * a.y * is);
        d.z = k * is * (b.z * it - cs_ * a.z * is);
        d.w = k * is * (b.w * it - cs_ * a.w * is);
        dp[i] = d;
    }
}

__global__ void l2norm_fwd_kernel(const float* __restrict__ x, long long M, int C, float* __restrict__ y,
                                  float* __restrict__ inv_norm) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= M) return;
    const float4* xp = reinterpret_cast<const float4*>(x + r * C);
    float4* yp = reinterpret_cast<float4*>(y + r * C);
    float ss = 0.f;
    for (int i = lane; i < C / 4; i += 32) {
        const float4 a = xp[i];
        ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
    ss = warp_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    for (int i = lane; i < C / 4; i += 32) {
        float4 a = xp[i];
        a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
        yp[i] = a;
    }
    if (lane == 0 && inv_norm) inv_norm[r] = inv;
}

// dx = inv * (dy - y * <y, dy>)
__global__ void l2norm_bwd_kernel(const float* __restrict__ y, const float* __restrict__ inv_norm,
                                  const float* __restrict__ dy, long long M, int C, float* __restrict__ dx) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= M) return;
    const float4* yp = reinterpret_cast<const float4*>(y + r * C);
    const float4* gp = reinterpret_cast<const float4*>(dy + r * C);
    float4* dp = reinterpret_cast<float4*>(dx + r * C);
    float dot = 0.f;
    for (int i = lane; i < C / 4; i += 32) {
        const float4 a = yp[i], g = gp[i];
        dot += a.x * g.x + a.y * g.y + a.z * g.z + a.w * g.w;
    }
    dot = warp_sum(dot);
    const float inv = inv_norm[r];
    for (int i = lane; i < C / 4; i += 32) {
        const float4 a = yp[i], g = gp[i];
        float4 d;
        d.x = inv * (g.x - a.x * dot);
        d.y = inv * (g.y - a.y * dot);
        d.z = inv * (g.z - a.z * dot);
        d.w = inv * (g.w - a.w * dot);
        dp[i] = d;
    }
}

}  // namespace region
}  // namespace cs

using namespace cs;
using namespace cs::region;

extern "C" int cs_extract_rois(const float* normed_boxes, int B, int K, float* rois, int32_t* crop_index,
                               int32_t* roi_batch, int32_t* img_offsets, void* stream) {
    CS_CHECK_ARG(normed_boxes && rois && crop_index && roi_batch && img_offsets, "cs_extract_rois: null pointer");
    CS_CHECK_ARG(B > 0 && K > 0 && B <= 8192, "cs_extract_rois: bad shape B=%d K=%d", B, K);
    extract_rois_kernel<<<1, 1024, (B + 1) * sizeof(int), (cudaStream_t)stream>>>(normed_boxes, B, K, rois,
                                                                                  crop_index, roi_batch,
                                                                                  img_offsets);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

extern "C" int cs_gather_rows(const void* src, const int32_t* index, int R, int64_t row_bytes, void* dst,
                              void* stream) {
    CS_CHECK_ARG(src && index && dst, "cs_gather_rows: null pointer");
    CS_CHECK_ARG(row_bytes > 0 && row_bytes % 16 == 0, "cs_gather_rows: row_bytes must be a multiple of 16");
    if (R == 0) return CS_OK;
    const long long vecs = row_bytes / 16;
    dim3 grid((unsigned)min((long long)64, (vecs + 255) / 256), R);
    gather_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)src, index, vecs, (uint4*)dst);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

static int roi_smem_bytes(int H, int W, bool stage, bool bwd = false) {
    return (ROI_CHUNK * (H + W) + (stage ? H * W * ROI_CS : 0) + (bwd ? ROI_CHUNK * (ROI_CS + 4) : 0)) * (int)sizeof(float);
}
constexpr int kMaxStage = 200 * 1024;

extern "C" int cs_roi_align_fwd(const float* fmap, int B, int H, int W, int C, const float* rois,
                                const int32_t* img_offsets, int R, float* wy, float* wx, float* out,
                                void* stream) {
    CS_CHECK_ARG(fmap && rois && img_offsets && wy && wx && out, "cs_roi_align_fwd: null pointer");
    CS_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && R >= 0 && H + W <= 384, "cs_roi_align_fwd: bad shape");
    CS_CHECK_ARG((uintptr_t)fmap % 16 == 0 && (uintptr_t)rois % 16 == 0 && C % 4 == 0,
                 "cs_roi_align_fwd: fmap and rois must be 16 B aligned, C %% 4 == 0");
    if (R == 0) return CS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    roi_weights_kernel<<<ceil_div(R, RW_BOXES), RW_BOXES, RW_BOXES * (H + W) * sizeof(float), st>>>(rois, R, H, W, wy, wx);
    CS_LAUNCH_CHECK();
    dim3 grid(ceil_div(C, ROI_CS), B);
    const bool stage = roi_smem_bytes(H, W, true) <= kMaxStage;
    const int smem = roi_smem_bytes(H, W, stage);
    static bool configured = false;
    if (!configured) {
        CS_CUDA(cudaFuncSetAttribute(roi_align_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxStage));
        CS_CUDA(cudaFuncSetAttribute(roi_align_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxStage));
        configured = true;
    }
    if (stage) roi_align_fwd_kernel<true><<<grid, ROI_THREADS, smem, st>>>(fmap, H, W, C, img_offsets, wy, wx, out);
    else roi_align_fwd_kernel<false><<<grid, ROI_THREADS, smem, st>>>(fmap, H, W, C, img_offsets, wy, wx, out);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

extern "C" int cs_roi_align_bwd(const float* d_out, int B, int H, int W, int C, const int32_t* img_offsets,
                                int R, const float* wy, const float* wx, float* d_fmap, void* stream) {
    CS_CHECK_ARG(d_out && img_offsets && wy && wx && d_fmap, "cs_roi_align_bwd: null pointer");
    CS_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && R >= 0 && H + W <= 384, "cs_roi_align_bwd: bad shape");
    CS_CHECK_ARG((uintptr_t)d_fmap % 16 == 0 && C % 4 == 0, "cs_roi_align_bwd: d_fmap must be 16 B aligned, C %% 4 == 0");
    dim3 grid(ceil_div(C, ROI_CS), B);
    const bool stage = roi_smem_bytes(H, W, true, true) <= kMaxStage;
    const int smem = roi_smem_bytes(H, W, stage, true);
    static bool configured = false;
    if (!configured) {
        CS_CUDA(cudaFuncSetAttribute(roi_align_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxStage));
        CS_CUDA(cudaFuncSetAttribute(roi_align_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxStage));
        configured = true;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (stage) roi_align_bwd_kernel<true><<<grid, ROI_THREADS, smem, st>>>(d_out, H, W, C, img_offsets, wy, wx, d_fmap);
    else roi_align_bwd_kernel<false><<<grid, ROI_THREADS, smem, st>>>(d_out, H, W, C, img_offsets, wy, wx, d_fmap);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

extern "C" int cs_mask_pool_fwd(const float* fmap, int B, int HW, int C, const float* masks,
                                const int32_t* img_offsets, int R, float* out, void* stream) {
    CS_CHECK_ARG(fmap && masks && img_offsets && out, "cs_mask_pool_fwd: null pointer");
    CS_CHECK_ARG(B > 0 && HW > 0 && C > 0 && R >= 0, "cs_mask_pool_fwd: bad shape");
    if (R == 0) return CS_OK;
    CS_CHECK_ARG((uintptr_t)fmap % 16 == 0 && C % 4 == 0, "cs_mask_pool_fwd: fmap must be 16 B aligned, C %% 4 == 0");
    dim3 grid(ceil_div(C, MP_CS), B);
    mask_pool_kernel<<<grid, MP_THREADS, 0, (cudaStream_t)stream>>>(fmap, HW, C, masks, img_offsets, out);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

extern "C" int cs_cosine_loss_fwd(const float* s, const float* t, int R, int C, float weight, float* loss,
                                  float* row_stats, void* stream) {
    CS_CHECK_ARG(s && t && loss && row_stats, "cs_cosine_loss_fwd: null pointer");
    CS_CHECK_ARG(R > 0 && C > 0 && C % 4 == 0, "cs_cosine_loss_fwd: bad shape R=%d C=%d", R, C);
    cudaStream_t st = (cudaStream_t)stream;
    cosine_rows_kernel<<<ceil_div(R, 8), 256, 0, st>>>(s, t, R, C, row_stats);
    CS_LAUNCH_CHECK();
    cosine_reduce_kernel<<<1, 1024, 0, st>>>(row_stats, R, weight, loss);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

extern "C" int cs_cosine_loss_bwd(const float* s, const float* t, const float* row_stats, int R, int C,
                                  float weight, const float* d_loss, float* d_s, void* stream) {
    CS_CHECK_ARG(s && t && row_stats && d_loss && d_s, "cs_cosine_loss_bwd: null pointer");
    CS_CHECK_ARG(R > 0 && C > 0 && C % 4 == 0, "cs_cosine_loss_bwd: bad shape R=%d C=%d", R, C);
    cosine_bwd_kernel<<<ceil_div(R, 8), 256, 0, (cudaStream_t)stream>>>(s, t, row_stats, R, C, weight, d_loss, d_s);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

extern "C" int cs_l2norm_fwd(const float* x, int64_t M, int C, float* y, float* inv_norm, void* stream) {
    CS_CHECK_ARG(x && y, "cs_l2norm_fwd: null pointer");
    CS_CHECK_ARG(M > 0 && C > 0 && C % 4 == 0, "cs_l2norm_fwd: bad shape");
    l2norm_fwd_kernel<<<ceil_div(M, 8), 256, 0, (cudaStream_t)stream>>>(x, M, C, y, inv_norm);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

extern "C" int cs_l2norm_bwd(const float* y, const float* inv_norm, const float* d_y, int64_t M, int C,
                             float* d_x, void* stream) {
    CS_CHECK_ARG(y && inv_norm && d_y && d_x, "cs_l2norm_bwd: null pointer");
    CS_CHECK_ARG(M > 0 && C > 0 && C % 4 == 0, "cs_l2norm_bwd: bad shape");
    l2norm_bwd_kernel<<<ceil_div(M, 8), 256, 0, (cudaStream_t)stream>>>(y, inv_norm, d_y, M, C, d_x);
    CS_LAUNCH_CHECK();
    return CS_OK;
}
