// Student-side backward helpers (autograd of eva_vit_model.py:98-105, 174-256, 300-324 as
// invoked by train.py:96): cast/transpose staging for the dgrad / wgrad GEMMs, LayerNorm
// backward, column sums (bias gradients), SwiGLU forward/backward on the packed gate|up layout,
// fused AdamW.  All HBM-bound, f32 math, deterministic (no atomics).
#include "common.cuh"

namespace cs {
namespace bwd {

// --------------------------------------------------------------------------------------------
// src [M,N] (f32 or bf16) -> dst [M,ldd] bf16 (optional) and dst_t [N,ldt] bf16 (optional)
// 32x32 tiles through shared memory, both sides coalesced.
// --------------------------------------------------------------------------------------------
template <typename TIn>
__global__ void cast_transpose_kernel(const TIn* __restrict__ src, long long M, int N, long long lds,
                                      __nv_bfloat16* __restrict__ dst, long long ldd,
                                      __nv_bfloat16* __restrict__ dst_t, long long ldt) {
    __shared__ __nv_bfloat16 tile[32][33];
    const long long m0 = (long long)blockIdx.y * 32;
    const int n0 = blockIdx.x * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const long long m = m0 + ty + i;
        const int n = n0 + tx;
        __nv_bfloat16 v = __float2bfloat16(0.f);
        if (m < M && n < N) v = __float2bfloat16((float)src[m * lds + n]);
        tile[ty + i][tx] = v;
        if (dst != nullptr && m < M && n < N) dst[m * ldd + n] = v;
    }
    if (dst_t == nullptr) return;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int n = n0 + ty + i;
        const long long m = m0 + tx;
        if (n < N && m < M) dst_t[(long long)n * ldt + m] = tile[tx][ty + i];
    }
}

// --------------------------------------------------------------------------------------------
// LayerNorm backward, input gradient:  dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat))
// one warp per row; x may be f32 or bf16 with the same row map as the forward.
// out: f32 (optionally accumulated onto `add` = upstream residual gradient) or bf16.
// --------------------------------------------------------------------------------------------
template <typename TX, typename TDY>
__global__ void __launch_bounds__(256)
layernorm_bwd_dx_kernel(const TDY* __restrict__ dy, long long lddy, const TX* __restrict__ x, long long ldx,
                        long long M, int D, int row_div, int row_mul, int row_off,
                        const float* __restrict__ mean, const float* __restrict__ rstd,
                        const float* __restrict__ gamma, const float* __restrict__ add, long long ldadd,
                        void* __restrict__ dx, int dx_bf16, long long lddx, int dx_row_mapped) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long m = (long long)blockIdx.x * 8 + warp;
    if (m >= M) return;
    const long long prow = m * row_mul + (row_div > 0 ? m / row_div : 0) + row_off;
    const TX* xr = x + prow * ldx;
    const TDY* dyr = dy + m * lddy;
    const float mu = mean[m], rs = rstd[m];
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < D; i += 32) {
        const float gdy = gamma[i] * (float)dyr[i];
        const float xh = ((float)xr[i] - mu) * rs;
        s1 += gdy;
        s2 += gdy * xh;
    }
    s1 = warp_sum(s1) / (float)D;
    s2 = warp_sum(s2) / (float)D;
    const long long orow = dx_row_mapped ? prow : m;
    for (int i = lane; i < D; i += 32) {
        const float gdy = gamma[i] * (float)dyr[i];
        const float xh = ((float)xr[i] - mu) * rs;
        float v = rs * (gdy - s1 - xh * s2);
        if (add != nullptr) v += add[orow * ldadd + i];
        if (dx_bf16)
            reinterpret_cast<__nv_bfloat16*>(dx)[orow * lddx + i] = __float2bfloat16(v);
        else
            reinterpret_cast<float*>(dx)[orow * lddx + i] = v;
    }
}

// partial[s][0][c] = sum_{rows of split s} dy*xhat ; partial[s][1][c] = sum dy    (x == nullptr: only [1])
template <typename TX, typename TDY>
__global__ void __launch_bounds__(128)
col_partials_kernel(const TDY* __restrict__ dy, long long lddy, const TX* __restrict__ x, long long ldx, long long M,
                    int D, int row_div, int row_mul, int row_off, const float* __restrict__ mean,
                    const float* __restrict__ rstd, float* __restrict__ partial) {
    const int c = blockIdx.x * 128 + threadIdx.x;
    const int S = gridDim.y, s = blockIdx.y;
    const long long rows_per = (M + S - 1) / S;
    const long long r0 = s * rows_per, r1 = min(M, r0 + rows_per);
    float dg = 0.f, db = 0.f;
    if (c < D) {
        long long m = r0;
        // 8 rows per trip with every load issued before the first use: the loop is latency bound otherwise (one dependent
        // global load per row and thread); the summation order stays row-ascending, i.e. deterministic
        for (; m + 8 <= r1; m += 8) {
            float g[8], xv[8], mu[8], rs[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) g[u] = (float)dy[(m + u) * lddy + c];
            if (x != nullptr) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const long long mm = m + u;
                    const long long prow = mm * row_mul + (row_div > 0 ? mm / row_div : 0) + row_off;
                    xv[u] = (float)x[prow * ldx + c];
                    mu[u] = mean[mm];
                    rs[u] = rstd[mm];
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                db += g[u];
                if (x != nullptr) dg += g[u] * (xv[u] - mu[u]) * rs[u];
            }
        }
        for (; m < r1; ++m) {
            const float g = (float)dy[m * lddy + c];
            db += g;
            if (x != nullptr) {
                const long long prow = m * row_mul + (row_div > 0 ? m / row_div : 0) + row_off;
                dg += g * ((float)x[prow * ldx + c] - mean[m]) * rstd[m];
            }
        }
        partial[((long long)s * 2 + 0) * D + c] = dg;
        partial[((long long)s * 2 + 1) * D + c] = db;
    }
}
__global__ void col_reduce_kernel(const float* __restrict__ partial, int S, int D, float* __restrict__ dgamma,
                                  float* __restrict__ dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= D) return;
    float dg = 0.f, db = 0.f;
    for (int s = 0; s < S; ++s) {
        dg += partial[((long long)s * 2 + 0) * D + c];
        db += partial[((long long)s * 2 + 1) * D + c];
    }
    if (dgamma) dgamma[c] = dg;
    if (dbeta) dbeta[c] = db;
}

// --------------------------------------------------------------------------------------------
// SwiGLU on the packed layout of cs_pack_swiglu_weights: hidden column j = t*128 + jj has its gate
// at packed column t*256 + jj and its up value at t*256 + 128 + jj.
// --------------------------------------------------------------------------------------------
// split: 0 = packed (gate|up interleaved by 128), 1 = up columns start at Hd, >= 2 = up columns start at `split`
__device__ __forceinline__ void swiglu_cols(int j, int Hd, int split, int& gate_col, int& up_col) {
    if (split) {
        gate_col = j;
        up_col = (split == 1 ? Hd : split) + j;
    } else {
        gate_col = (j >> 7) * 256 + (j & 127);
        up_col = gate_col + 128;
    }
}
__global__ void swiglu_fwd_kernel(const __nv_bfloat16* __restrict__ x12, long long M, int Hd, long long ld12,
                                  __nv_bfloat16* __restrict__ h, long long ldh, int split) {
    const long long total = M * (Hd / 2);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / (Hd / 2);
        const int j = (int)(i % (Hd / 2)) * 2;
        int pc, uc;
        swiglu_cols(j, Hd, split, pc, uc);
        const float2 g = unpack_bf16(*reinterpret_cast<const uint32_t*>(x12 + m * ld12 + pc));
        const float2 u = unpack_bf16(*reinterpret_cast<const uint32_t*>(x12 + m * ld12 + uc));
        const float h0 = g.x / (1.f + __expf(-g.x)) * u.x;
        const float h1 = g.y / (1.f + __expf(-g.y)) * u.y;
        *reinterpret_cast<uint32_t*>(h + m * ldh + j) = pack_bf16(h0, h1);
    }
}
__global__ void swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ x12, const __nv_bfloat16* __restrict__ dh,
                                  long long M, int Hd, long long ld12, long long lddh,
                                  __nv_bfloat16* __restrict__ dx12, int split) {
    const long long total = M * (Hd / 2);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / (Hd / 2);
        const int j = (int)(i % (Hd / 2)) * 2;
        int pc, uc;
        swiglu_cols(j, Hd, split, pc, uc);
        const float2 g = unpack_bf16(*reinterpret_cast<const uint32_t*>(x12 + m * ld12 + pc));
        const float2 u = unpack_bf16(*reinterpret_cast<const uint32_t*>(x12 + m * ld12 + uc));
        const float2 d = unpack_bf16(*reinterpret_cast<const uint32_t*>(dh + m * lddh + j));
        const float s0 = 1.f / (1.f + __expf(-g.x)), s1 = 1.f / (1.f + __expf(-g.y));
        const float dg0 = d.x * u.x * s0 * (1.f + g.x * (1.f - s0));
        const float dg1 = d.y * u.y * s1 * (1.f + g.y * (1.f - s1));
        const float du0 = d.x * g.x * s0, du1 = d.y * g.y * s1;
        *reinterpret_cast<uint32_t*>(dx12 + m * ld12 + pc) = pack_bf16(dg0, dg1);
        *reinterpret_cast<uint32_t*>(dx12 + m * ld12 + uc) = pack_bf16(du0, du1);
    }
}

// --------------------------------------------------------------------------------------------
// AdamW, torch.optim.AdamW semantics (main.py:205-213): decoupled decay, bias correction.
//   p *= 1 - lr*wd;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
//   p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
// --------------------------------------------------------------------------------------------
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float wd,
                             float bc1, float bc2_sqrt, float grad_scale) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i] * grad_scale;
        float pi = p[i] * (1.f - lr * wd);
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        pi -= (lr / bc1) * (mi / denom);
        p[i] = pi;
    }
}

}  // namespace bwd
}  // namespace cs

using namespace cs;
using namespace cs::bwd;

extern "C" int cs_cast_transpose_bf16(const void* src, cs_dtype_t dtype, int64_t M, int N, int64_t lds,
                                      void* dst_bf16, int64_t ldd, void* dst_t_bf16, int64_t ldt, void* stream) {
    CS_CHECK_ARG(src && (dst_bf16 || dst_t_bf16), "cs_cast_transpose_bf16: null pointer");
    CS_CHECK_ARG(M > 0 && N > 0 && lds >= N, "cs_cast_transpose_bf16: bad shape");
    CS_CHECK_ARG(!dst_bf16 || ldd >= N, "cs_cast_transpose_bf16: ldd < N");
    CS_CHECK_ARG(!dst_t_bf16 || ldt >= M, "cs_cast_transpose_bf16: ldt < M");
    dim3 grid(ceil_div(N, 32), ceil_div(M, 32)), block(32, 8);
    CS_CHECK_ARG(grid.y <= 65535u * 32u, "cs_cast_transpose_bf16: M too large");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CS_F32)
        cast_transpose_kernel<float><<<grid, block, 0, st>>>((const float*)src, M, N, lds, (__nv_bfloat16*)dst_bf16, ldd,
                                                             (__nv_bfloat16*)dst_t_bf16, ldt);
    else
        cast_transpose_kernel<__nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)src, M, N, lds,
                                                                     (__nv_bfloat16*)dst_bf16, ldd,
                                                                     (__nv_bfloat16*)dst_t_bf16, ldt);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

extern "C" int cs_layernorm_bwd_dx(const void* dy, cs_dtype_t dy_dtype, int64_t lddy, const void* x,
                                   cs_dtype_t x_dtype, int64_t ldx, int64_t M, int D, int row_div, int row_mul,
                                   int row_off, const float* mean, const float* rstd, const float* gamma,
                                   const float* add, int64_t ldadd, void* dx, cs_dtype_t dx_dtype, int64_t lddx,
                                   int dx_row_mapped, void* stream) {
    CS_CHECK_ARG(dy && x && mean && rstd && gamma && dx, "cs_layernorm_bwd_dx: null pointer");
    CS_CHECK_ARG(M > 0 && D > 0 && row_mul >= 1, "cs_layernorm_bwd_dx: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ceil_div(M, 8);
    const int obf = dx_dtype == CS_BF16;
#define LNB(TX, TDY)                                                                                             \
    layernorm_bwd_dx_kernel<TX, TDY><<<grid, 256, 0, st>>>((const TDY*)dy, lddy, (const TX*)x, ldx, M, D, row_div, \
                                                           row_mul, row_off, mean, rstd, gamma, add, ldadd, dx, obf, \
                                                           lddx, dx_row_mapped)
    if (x_dtype == CS_F32 && dy_dtype == CS_BF16) LNB(float, __nv_bfloat16);
    else if (x_dtype == CS_BF16 && dy_dtype == CS_BF16) LNB(__nv_bfloat16, __nv_bfloat16);
    else if (x_dtype == CS_F32 && dy_dtype == CS_F32) LNB(float, float);
    else LNB(__nv_bfloat16, float);
#undef LNB
    CS_LAUNCH_CHECK();
    return CS_OK;
}

extern "C" int cs_col_reduce(const void* dy, cs_dtype_t dy_dtype, int64_t lddy, const void* x, cs_dtype_t x_dtype,
                             int64_t ldx, int64_t M, int D, int row_div, int row_mul, int row_off,
                             const float* mean, const float* rstd, float* dgamma, float* dbeta,
                             float* workspace, int64_t workspace_floats, void* stream) {
    CS_CHECK_ARG(dy && dbeta && workspace, "cs_col_reduce: null pointer");
    CS_CHECK_ARG(!x || (mean && rstd && dgamma), "cs_col_reduce: x needs mean, rstd and dgamma");
    CS_CHECK_ARG(M > 0 && D > 0, "cs_col_reduce: bad shape");
    int S = (int)((M + 63) / 64 < 128 ? (M + 63) / 64 : 128);
    if (S < 1) S = 1;
    CS_CHECK_ARG(workspace_floats >= (int64_t)S * 2 * D, "cs_col_reduce: workspace too small (need %lld floats)",
                 (long long)S * 2 * D);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(ceil_div(D, 128), S);
#define CP(TX, TDY)                                                                                               \
    col_partials_kernel<TX, TDY><<<grid, 128, 0, st>>>((const TDY*)dy, lddy, (const TX*)x, ldx, M, D, row_div, row_mul, \
                                                       row_off, mean, rstd, workspace)
    if (x_dtype == CS_F32 && dy_dtype == CS_BF16) CP(float, __nv_bfloat16);
    else if (x_dtype == CS_BF16 && dy_dtype == CS_BF16) CP(__nv_bfloat16, __nv_bfloat16);
    else if (x_dtype == CS_F32 && dy_dtype == CS_F32) CP(float, float);
    else CP(__nv_bfloat16, float);
#undef CP
    CS_LAUNCH_CHECK();
    col_reduce_kernel<<<ceil_div(D, 128), 128, 0, st>>>(workspace, S, D, x ? dgamma : nullptr, dbeta);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

extern "C" int cs_swiglu_fwd(const void* x12_bf16, int64_t M, int Hd, int64_t ld12, void* h_bf16, int64_t ldh,
                             int split_layout, void* stream) {
    CS_CHECK_ARG(x12_bf16 && h_bf16 && M > 0 && Hd > 0 && Hd % 2 == 0 && ld12 >= 2 * Hd && ldh >= Hd &&
                     (split_layout || Hd % 128 == 0) && (split_layout < 2 || (split_layout >= Hd && split_layout % 2 == 0 && ld12 >= split_layout + Hd)),
                 "cs_swiglu_fwd: bad argument (packed layout needs Hd %% 128 == 0)");
    const long long total = M * (Hd / 2);
    const int grid = (int)(((total + 255) / 256) < (long long)num_sms() * 16 ? ((total + 255) / 256) : (long long)num_sms() * 16);
    swiglu_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x12_bf16, M, Hd, ld12,
                                                              (__nv_bfloat16*)h_bf16, ldh, split_layout);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

extern "C" int cs_swiglu_bwd(const void* x12_bf16, const void* dh_bf16, int64_t M, int Hd, int64_t ld12,
                             int64_t lddh, void* dx12_bf16, int split_layout, void* stream) {
    CS_CHECK_ARG(x12_bf16 && dh_bf16 && dx12_bf16 && M > 0 && Hd > 0 && Hd % 2 == 0 && ld12 >= 2 * Hd && lddh >= Hd &&
                     (split_layout || Hd % 128 == 0) && (split_layout < 2 || (split_layout >= Hd && split_layout % 2 == 0 && ld12 >= split_layout + Hd)),
                 "cs_swiglu_bwd: bad argument");
    const long long total = M * (Hd / 2);
    const int grid = (int)(((total + 255) / 256) < (long long)num_sms() * 16 ? ((total + 255) / 256) : (long long)num_sms() * 16);
    swiglu_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x12_bf16,
                                                              (const __nv_bfloat16*)dh_bf16, M, Hd, ld12, lddh,
                                                              (__nv_bfloat16*)dx12_bf16, split_layout);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

extern "C" int cs_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                             double beta1, double beta2, double eps, double weight_decay, int step,
                             double grad_scale, void* stream) {
    CS_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "cs_adamw_step: bad argument");
    // bias corrections in double like torch.optim (python floats)
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2 = 1.0 - pow(beta2, (double)step);
    const long long blocks = (n + 255) / 256;
    const int grid = (int)(blocks < (long long)num_sms() * 16 ? blocks : (long long)num_sms() * 16);
    adamw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, (float)lr, (float)beta1,
                                                         (float)beta2, (float)eps, (float)weight_decay, (float)bc1,
                                                         (float)sqrt(bc2), (float)grad_scale);
    CS_LAUNCH_CHECK();
    return CS_OK;
}
