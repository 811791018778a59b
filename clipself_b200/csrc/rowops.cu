// Row / element-wise kernels of the EVA02 tower: LayerNorm, patch gather (im2col), CLS rows,
// casts and weight packing.  All HBM-bound: one pass over the data, 128-bit accesses, f32 math.
#include <cstdarg>

#include "common.cuh"

namespace cs {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

namespace rowops {

constexpr int LN_WARPS = 8;

// One warp per row, any D.  The row is staged in shared memory (single HBM read), statistics are the
// two-pass mean / centred variance in f32 (same numerics as ATen's layer_norm).  Used for the widths
// without a register-resident instantiation (e.g. ViT-L's 2730-wide ffn_ln).
template <typename TIn>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_kernel(const TIn* __restrict__ x, long long ldx, long long M, int D, int row_div, int row_mul,
                     int row_off, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                     __nv_bfloat16* __restrict__ y, long long ldy, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out) {
    extern __shared__ float s_rows[];   // LN_WARPS * D
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long m = (long long)blockIdx.x * LN_WARPS + warp;
    if (m >= M) return;
    const long long prow = m * row_mul + (row_div > 0 ? m / row_div : 0) + row_off;
    const TIn* xr = x + prow * ldx;
    float* s = s_rows + warp * D;
    float sum = 0.f;
    for (int i = lane; i < D; i += 32) {
        const float v = (float)xr[i];
        s[i] = v;
        sum += v;
    }
    sum = warp_sum(sum);
    const float mean = sum / (float)D;
    float var = 0.f;
    for (int i = lane; i < D; i += 32) {
        const float d = s[i] - mean;
        var += d * d;
    }
    var = warp_sum(var);
    const float rstd = rsqrtf(var / (float)D + eps);
    __nv_bfloat16* yr = y + m * ldy;
    for (int i = lane; i < D; i += 32) yr[i] = __float2bfloat16((s[i] - mean) * rstd * gamma[i] + beta[i]);
    if (lane == 0) {
        if (mean_out) mean_out[m] = mean;
        if (rstd_out) rstd_out[m] = rstd;
    }
}

// Register-resident variant for the row widths of the towers (D = 32 lanes x VPL 16-byte vectors):
// every load of the row is in flight at once, statistics and normalisation never leave registers.
template <typename TIn, int VPL>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_reg_kernel(const TIn* __restrict__ x, long long ldx, long long M, int row_div, int row_mul, int row_off,
                         const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                         __nv_bfloat16* __restrict__ y, long long ldy, float* __restrict__ mean_out,
                         float* __restrict__ rstd_out) {
    constexpr int EPV = 16 / (int)sizeof(TIn);          // elements per 16-byte vector (4 f32 / 8 bf16)
    constexpr int D = 32 * VPL * EPV;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long m = (long long)blockIdx.x * LN_WARPS + warp;
    if (m >= M) return;
    const long long prow = m * row_mul + (row_div > 0 ? m / row_div : 0) + row_off;
    const TIn* xr = x + prow * ldx;
    float v[VPL * EPV];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int e0 = (i * 32 + lane) * EPV;
        if constexpr (sizeof(TIn) == 4) {
            const float4 f = *reinterpret_cast<const float4*>(xr + e0);
            v[i * 4 + 0] = f.x; v[i * 4 + 1] = f.y; v[i * 4 + 2] = f.z; v[i * 4 + 3] = f.w;
        } else {
            const uint4 u = *reinterpret_cast<const uint4*>(xr + e0);
            const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
            v[i * 8 + 0] = a.x; v[i * 8 + 1] = a.y; v[i * 8 + 2] = b.x; v[i * 8 + 3] = b.y;
            v[i * 8 + 4] = c.x; v[i * 8 + 5] = c.y; v[i * 8 + 6] = d.x; v[i * 8 + 7] = d.y;
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPL * EPV; ++i) sum += v[i];
    const float mean = warp_sum(sum) / (float)D;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < VPL * EPV; ++i) {
        const float d = v[i] - mean;
        var += d * d;
    }
    const float rstd = rsqrtf(warp_sum(var) / (float)D + eps);
    __nv_bfloat16* yr = y + m * ldy;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int e0 = (i * 32 + lane) * EPV;
        float o[EPV];
#pragma unroll
        for (int j = 0; j < EPV; j += 4) {
            const float4 g = *reinterpret_cast<const float4*>(gamma + e0 + j);
            const float4 b = *reinterpret_cast<const float4*>(beta + e0 + j);
            o[j + 0] = (v[i * EPV + j + 0] - mean) * rstd * g.x + b.x;
            o[j + 1] = (v[i * EPV + j + 1] - mean) * rstd * g.y + b.y;
            o[j + 2] = (v[i * EPV + j + 2] - mean) * rstd * g.z + b.z;
            o[j + 3] = (v[i * EPV + j + 3] - mean) * rstd * g.w + b.w;
        }
        if constexpr (EPV == 4) {
            uint2 pk;
            pk.x = pack_bf16(o[0], o[1]);
            pk.y = pack_bf16(o[2], o[3]);
            *reinterpret_cast<uint2*>(yr + e0) = pk;
        } else {
            uint4 pk;
            pk.x = pack_bf16(o[0], o[1]);
            pk.y = pack_bf16(o[2], o[3]);
            pk.z = pack_bf16(o[4], o[5]);
            pk.w = pack_bf16(o[6], o[7]);
            *reinterpret_cast<uint4*>(yr + e0) = pk;
        }
    }
    if (lane == 0) {
        if (mean_out) mean_out[m] = mean;
        if (rstd_out) rstd_out[m] = rstd;
    }
}

template <typename TIn, int VPL>
static void launch_ln_reg(const void* x, long long ldx, long long M, int row_div, int row_mul, int row_off,
                          const float* gamma, const float* beta, float eps, void* y, long long ldy, float* mean,
                          float* rstd, cudaStream_t st) {
    layernorm_fwd_reg_kernel<TIn, VPL><<<ceil_div(M, LN_WARPS), LN_WARPS * 32, 0, st>>>(
        (const TIn*)x, ldx, M, row_div, row_mul, row_off, gamma, beta, eps, (__nv_bfloat16*)y, ldy, mean, rstd);
}

// images [B,3,S,S] -> patches [B*g*g, ldp] bf16, column = c*P*P + py*P + px  (flattened conv weight
// order).  One CTA per (image, patch row): each image row segment is read contiguously.
template <typename TIn>
__global__ void im2col_kernel(const TIn* __restrict__ img, int S, int P, __nv_bfloat16* __restrict__ out,
                              long long ldp) {
    const int g = S / P;
    const int b = blockIdx.y, gy = blockIdx.x;
    const int kreal = 3 * P * P;
    // iterate over (c, py, x) with x the full image row -> coalesced reads
    for (int i = threadIdx.x; i < 3 * P * S; i += blockDim.x) {
        const int x = i % S;
        const int py = (i / S) % P;
        const int c = i / (S * P);
        const float v = (float)img[(((long long)b * 3 + c) * S + (gy * P + py)) * S + x];
        const int gx = x / P, px = x % P;
        if (gx < g) out[((long long)(b * g + gy) * g + gx) * ldp + c * P * P + py * P + px] = __float2bfloat16(v);
    }
    // zero the padding columns [kreal, ldp)
    const int pad = (int)ldp - kreal;
    for (int i = threadIdx.x; i < g * pad; i += blockDim.x) {
        const int gx = i / pad, j = i % pad;
        out[((long long)(b * g + gy) * g + gx) * ldp + kreal + j] = __float2bfloat16(0.f);
    }
}

// Fast path (P % 8 == 0): one thread moves 8 consecutive pixels of one patch row:
// 32 B (f32) / 16 B (bf16) read -> one 16 B bf16 store.
template <typename TIn>
__global__ void im2col_vec8_kernel(const TIn* __restrict__ img, int B, int S, int P, __nv_bfloat16* __restrict__ out,
                                   long long ldp) {
    const int g = S / P;
    const int seg = S / 8;                                   // 8-pixel segments per image row
    const long long total = (long long)B * 3 * S * seg;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int xs = (int)(i % seg);
        const int y = (int)((i / seg) % S);
        const int c = (int)((i / ((long long)seg * S)) % 3);
        const int b = (int)(i / ((long long)seg * S * 3));
        const int x = xs * 8;
        const TIn* src = img + (((long long)b * 3 + c) * S + y) * S + x;
        float v[8];
        if constexpr (sizeof(TIn) == 4) {
            const float4 a = *reinterpret_cast<const float4*>(src), d = *reinterpret_cast<const float4*>(src + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = d.x; v[5] = d.y; v[6] = d.z; v[7] = d.w;
        } else {
            const uint4 u = *reinterpret_cast<const uint4*>(src);
            const float2 a = unpack_bf16(u.x), d = unpack_bf16(u.y), e = unpack_bf16(u.z), f = unpack_bf16(u.w);
            v[0] = a.x; v[1] = a.y; v[2] = d.x; v[3] = d.y; v[4] = e.x; v[5] = e.y; v[6] = f.x; v[7] = f.y;
        }
        uint4 pk;
        pk.x = pack_bf16(v[0], v[1]);
        pk.y = pack_bf16(v[2], v[3]);
        pk.z = pack_bf16(v[4], v[5]);
        pk.w = pack_bf16(v[6], v[7]);
        const int gy = y / P, py = y % P, gx = x / P, px = x % P;
        *reinterpret_cast<uint4*>(out + ((long long)(b * g + gy) * g + gx) * ldp + c * P * P + py * P + px) = pk;
    }
}

__global__ void fill_cls_kernel(const float* __restrict__ cls, const float* __restrict__ pos, int N, int D,
                                float* __restrict__ x) {
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < D; i += blockDim.x) x[(long long)b * N * D + i] = cls[i] + pos[i];
}

__global__ void cast_pad_kernel(const float* __restrict__ src, long long rows, long long cols, long long lds,
                                __nv_bfloat16* __restrict__ dst, long long ldd) {
    const long long total = rows * ldd;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / ldd, c = i % ldd;
        dst[i] = __float2bfloat16(c < cols ? src[r * lds + c] : 0.f);
    }
}

// packed row (t*256 + j) = w1 row (t*128 + j); packed row (t*256 + 128 + j) = w2 row (t*128 + j)
template <typename TIn>
__global__ void pack_swiglu_kernel(const TIn* __restrict__ w1, const TIn* __restrict__ w2, int Hd, int K,
                                   const float* __restrict__ b1, const float* __restrict__ b2,
                                   __nv_bfloat16* __restrict__ packed, long long ldk, float* __restrict__ bias12) {
    const int prow = blockIdx.x;                 // packed row
    const int t = prow / 256, j = prow % 256;
    const bool gate = j < 128;
    const int src = t * 128 + (gate ? j : j - 128);
    const TIn* w = gate ? w1 : w2;
    const bool ok = src < Hd;
    for (int k = threadIdx.x; k < ldk; k += blockDim.x)
        packed[(long long)prow * ldk + k] = __float2bfloat16((ok && k < K) ? (float)w[(long long)src * K + k] : 0.f);
    if (threadIdx.x == 0 && bias12) bias12[prow] = ok ? (gate ? b1[src] : b2[src]) : 0.f;
}


// ------------------------------------------------------------------------------------------
// bilinear resize of image planes (align_corners = false, no antialiasing): the --multiscale student
// input, training/clipself.py:17-27 (F.interpolate(images, size, mode='bilinear')).
// Index arithmetic follows ATen's upsample_bilinear2d: scale = in/out (f32), src = scale*(dst+.5)-.5
// clamped at 0, neighbour = +1 unless on the last row/column, weights (1-l, l).
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void resize_bilinear_kernel(const T* __restrict__ src, int Hin, int Win, int Hout, int Wout,
                                       float sy, float sx, T* __restrict__ dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const long long plane = blockIdx.z;
    if (x >= Wout) return;
    float fy = fmaf(sy, (float)y + 0.5f, -0.5f);            // ATen's kernel is compiled with FMA contraction
    float fx = fmaf(sx, (float)x + 0.5f, -0.5f);
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    const int y0 = (int)fy, x0 = (int)fx;
    const int yp = y0 < Hin - 1 ? 1 : 0, xp = x0 < Win - 1 ? 1 : 0;
    const float ly1 = fy - (float)y0, lx1 = fx - (float)x0;
    const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    const T* p = src + plane * Hin * Win + (long long)y0 * Win + x0;
    const float v00 = (float)p[0], v01 = (float)p[xp];
    const float v10 = (float)p[(long long)yp * Win], v11 = (float)p[(long long)yp * Win + xp];
    const float top = __fadd_rn(__fmul_rn(lx0, v00), __fmul_rn(lx1, v01));
    const float bot = __fadd_rn(__fmul_rn(lx0, v10), __fmul_rn(lx1, v11));
    const float val = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
    dst[plane * Hout * Wout + (long long)y * Wout + x] = (T)val;
}

}  // namespace rowops
}  // namespace cs

using namespace cs;
using namespace cs::rowops;

extern "C" const char* cs_last_error(void) { return cs::g_err; }
extern "C" int cs_abi_version(void) { return 3; }

extern "C" int cs_device_info(int* sm_out, int* num_sms_out, int64_t* hbm_bytes_out) {
    int dev = 0;
    CS_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    CS_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sm_out) *sm_out = p.major * 10 + p.minor;
    if (num_sms_out) *num_sms_out = p.multiProcessorCount;
    if (hbm_bytes_out) *hbm_bytes_out = (int64_t)p.totalGlobalMem;
    if (p.major != 10) {
        set_error("clipself_b200 needs an sm_100 (B200) device, found sm_%d%d", p.major, p.minor);
        return CS_ERR_UNSUPPORTED;
    }
    return CS_OK;
}

extern "C" int cs_layernorm_fwd(const void* x, cs_dtype_t x_dtype, int64_t ldx, int64_t M, int D, int row_div,
                                int row_mul, int row_off, const float* gamma, const float* beta, float eps,
                                void* y_bf16, int64_t ldy, float* mean, float* rstd, void* stream) {
    CS_CHECK_ARG(x && gamma && beta && y_bf16, "cs_layernorm_fwd: null pointer");
    CS_CHECK_ARG(M > 0 && D > 0 && ldx >= D && ldy >= D && row_mul >= 1, "cs_layernorm_fwd: bad shape (D=%d ldx=%lld ldy=%lld)", D,
                 (long long)ldx, (long long)ldy);
    const bool vec_ok = ldx % 8 == 0 && ldy % 8 == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y_bf16 % 16 == 0) &&
                        ((uintptr_t)gamma % 16 == 0) && ((uintptr_t)beta % 16 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    // register-resident fast paths for the widths of the EVA02 towers
#define LN_REG(T, VPL)                                                                                          \
    {                                                                                                           \
        launch_ln_reg<T, VPL>(x, ldx, M, row_div, row_mul, row_off, gamma, beta, eps, y_bf16, ldy, mean, rstd, st); \
        CS_LAUNCH_CHECK();                                                                                      \
        return CS_OK;                                                                                           \
    }
    if (vec_ok && x_dtype == CS_F32) {
        if (D == 768) LN_REG(float, 6)
        if (D == 1024) LN_REG(float, 8)
        if (D == 2048) LN_REG(float, 16)
        if (D == 128) LN_REG(float, 1)
    } else if (vec_ok) {
        if (D == 768) LN_REG(__nv_bfloat16, 3)
        if (D == 1024) LN_REG(__nv_bfloat16, 4)
        if (D == 2048) LN_REG(__nv_bfloat16, 8)
        if (D == 256) LN_REG(__nv_bfloat16, 1)
    }
#undef LN_REG
    const int smem = LN_WARPS * D * (int)sizeof(float);
    CS_CHECK_ARG(smem <= 160 * 1024, "cs_layernorm_fwd: D=%d too large", D);
    const int grid = ceil_div(M, LN_WARPS);
    if (x_dtype == CS_F32) {
        static bool cfg = false;
        if (!cfg) {
            CS_CUDA(cudaFuncSetAttribute(layernorm_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            cfg = true;
        }
        layernorm_fwd_kernel<float><<<grid, LN_WARPS * 32, smem, st>>>(
            (const float*)x, ldx, M, D, row_div, row_mul, row_off, gamma, beta, eps, (__nv_bfloat16*)y_bf16, ldy, mean, rstd);
    } else {
        static bool cfg = false;
        if (!cfg) {
            CS_CUDA(cudaFuncSetAttribute(layernorm_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            cfg = true;
        }
        layernorm_fwd_kernel<__nv_bfloat16><<<grid, LN_WARPS * 32, smem, st>>>(
            (const __nv_bfloat16*)x, ldx, M, D, row_div, row_mul, row_off, gamma, beta, eps, (__nv_bfloat16*)y_bf16, ldy, mean, rstd);
    }
    CS_LAUNCH_CHECK();
    return CS_OK;
}

// f32 rows -> bf16 copy + (sum, sum of squares) partials in the layout the LayerNorm-folded GEMM epilogues read
// (cs_gemm_epilogue_t.ln_stats): the whole-row statistics go to part 0, the remaining parts are zero.
namespace cs {
namespace rowops {
__global__ void __launch_bounds__(256)
row_stats_cast_kernel(const float* __restrict__ x, long long ldx, long long M, int D, __nv_bfloat16* __restrict__ xb,
                      long long ldb, float* __restrict__ stats, int parts) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long m = (long long)blockIdx.x * 8 + warp;
    if (m >= M) return;
    const float4* xr = reinterpret_cast<const float4*>(x + m * ldx);
    uint2* br = reinterpret_cast<uint2*>(xb + m * ldb);
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < D / 4; i += 32) {
        const float4 v = xr[i];
        s1 += (v.x + v.y) + (v.z + v.w);
        s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s2))));
        uint2 pk;
        pk.x = pack_bf16(v.x, v.y);
        pk.y = pack_bf16(v.z, v.w);
        br[i] = pk;
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    float2* sr = reinterpret_cast<float2*>(stats + m * parts * 2);
    for (int q = lane; q < parts; q += 32) sr[q] = q == 0 ? make_float2(s1, s2) : make_float2(0.f, 0.f);
}
}  // namespace rowops
}  // namespace cs

extern "C" int cs_row_stats_cast(const float* x, int64_t ldx, int64_t M, int D, void* xb_bf16, int64_t ldb,
                                 float* stats, int parts, void* stream) {
    using namespace cs;
    CS_CHECK_ARG(x && xb_bf16 && stats && M > 0 && D > 0 && D % 4 == 0 && ldx % 4 == 0 && ldb % 4 == 0 && parts > 0,
                 "cs_row_stats_cast: bad arguments (D, ldx, ldb multiples of 4; parts > 0)");
    CS_CHECK_ARG(((uintptr_t)x % 16 == 0) && ((uintptr_t)xb_bf16 % 8 == 0) && ((uintptr_t)stats % 8 == 0),
                 "cs_row_stats_cast: alignment");
    rowops::row_stats_cast_kernel<<<ceil_div(M, 8), 256, 0, (cudaStream_t)stream>>>(
        x, ldx, M, D, (__nv_bfloat16*)xb_bf16, ldb, stats, parts);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

extern "C" int cs_im2col_patches(const void* images, cs_dtype_t dtype, int B, int S, int P, void* patches_bf16,
                                 int64_t ldp, void* stream) {
    CS_CHECK_ARG(images && patches_bf16, "cs_im2col_patches: null pointer");
    CS_CHECK_ARG(B > 0 && S > 0 && P > 0 && S % P == 0 && ldp >= 3 * P * P, "cs_im2col_patches: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (P % 8 == 0 && ldp == 3 * P * P && ((uintptr_t)images % 16 == 0) && ((uintptr_t)patches_bf16 % 16 == 0)) {
        const long long total = (long long)B * 3 * S * (S / 8);
        const int blocks = (int)((total + 255) / 256 < (long long)num_sms() * 32 ? (total + 255) / 256 : (long long)num_sms() * 32);
        if (dtype == CS_F32)
            im2col_vec8_kernel<float><<<blocks, 256, 0, st>>>((const float*)images, B, S, P, (__nv_bfloat16*)patches_bf16, ldp);
        else
            im2col_vec8_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)images, B, S, P,
                                                                      (__nv_bfloat16*)patches_bf16, ldp);
        CS_LAUNCH_CHECK();
        return CS_OK;
    }
    dim3 grid(S / P, B);
    if (dtype == CS_F32)
        im2col_kernel<float><<<grid, 256, 0, st>>>((const float*)images, S, P, (__nv_bfloat16*)patches_bf16, ldp);
    else
        im2col_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)images, S, P, (__nv_bfloat16*)patches_bf16, ldp);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

extern "C" int cs_resize_bilinear(const void* src, cs_dtype_t dtype, int64_t planes, int Hin, int Win, int Hout,
                                  int Wout, void* dst, void* stream) {
    CS_CHECK_ARG(src && dst, "cs_resize_bilinear: null pointer");
    CS_CHECK_ARG(planes > 0 && planes < 65536 && Hin > 0 && Win > 0 && Hout > 0 && Hout < 65536 && Wout > 0,
                 "cs_resize_bilinear: bad shape");
    CS_CHECK_ARG(dtype == CS_F32 || dtype == CS_BF16, "cs_resize_bilinear: dtype must be f32 or bf16");
    const float sy = (float)Hin / (float)Hout, sx = (float)Win / (float)Wout;
    dim3 grid(ceil_div(Wout, 128), Hout, (unsigned)planes);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CS_F32)
        resize_bilinear_kernel<float><<<grid, 128, 0, st>>>((const float*)src, Hin, Win, Hout, Wout, sy, sx, (float*)dst);
    else
        resize_bilinear_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>((const __nv_bfloat16*)src, Hin, Win, Hout, Wout, sy, sx,
                                                                   (__nv_bfloat16*)dst);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

extern "C" int cs_fill_cls_rows(const float* cls_token, const float* pos_embed, int B, int N, int D, float* x,
                                void* stream) {
    CS_CHECK_ARG(cls_token && pos_embed && x && B > 0 && N > 0 && D > 0, "cs_fill_cls_rows: bad argument");
    fill_cls_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(cls_token, pos_embed, N, D, x);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

extern "C" int cs_cast_pad_bf16(const float* src, int64_t rows, int64_t cols, int64_t lds, void* dst_bf16,
                                int64_t ldd, void* stream) {
    CS_CHECK_ARG(src && dst_bf16 && rows > 0 && cols > 0 && lds >= cols && ldd >= cols, "cs_cast_pad_bf16: bad argument");
    const long long total = rows * ldd;
    const int grid = (int)min((long long)num_sms() * 8, (total + 255) / 256);
    cast_pad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, rows, cols, lds, (__nv_bfloat16*)dst_bf16, ldd);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

extern "C" int cs_pack_swiglu_weights(const void* w1, const void* w2, cs_dtype_t dtype, int Hd, int K,
                                      const float* b1, const float* b2, void* packed_bf16, int64_t ldk,
                                      float* bias12, void* stream) {
    CS_CHECK_ARG(w1 && w2 && packed_bf16, "cs_pack_swiglu_weights: null pointer");
    CS_CHECK_ARG(Hd > 0 && K > 0 && ldk >= K && ldk % 8 == 0, "cs_pack_swiglu_weights: bad shape");
    CS_CHECK_ARG(!bias12 || (b1 && b2), "cs_pack_swiglu_weights: bias12 needs b1 and b2");
    const int rows = ceil_div(Hd, 128) * 256;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CS_F32)
        pack_swiglu_kernel<float><<<rows, 256, 0, st>>>((const float*)w1, (const float*)w2, Hd, K, b1, b2,
                                                        (__nv_bfloat16*)packed_bf16, ldk, bias12);
    else
        pack_swiglu_kernel<__nv_bfloat16><<<rows, 256, 0, st>>>((const __nv_bfloat16*)w1, (const __nv_bfloat16*)w2, Hd, K,
                                                                b1, b2, (__nv_bfloat16*)packed_bf16, ldk, bias12);
    CS_LAUNCH_CHECK();
    return CS_OK;
}
