// Non-causal softmax attention for head_dim 64 over the packed q|k|v projections
// (eva_vit_model.py:206-217; math of the fallback branch :221-246).
//
// Forward: flash-style, one CTA = 64 query rows of one (image, head); 4 warps x 16 rows;
// K/V streamed in 64-key tiles through a cp.async double buffer; S and O live in registers,
// online softmax in f32 (exp2 domain).  Tensor-core path here is mma.sync (legacy HMMA) — the
// attention contractions are ~4 % (B/16) of the tower FLOPs; the tcgen05 rewrite is tracked in
// DESIGN.md §roadmap.
#include <cstdlib>

#include "common.cuh"

namespace cs {
namespace attn {

constexpr int HD = 64;        // head dim
constexpr int BQ = 64;        // query rows per CTA
constexpr int BK = 64;        // keys per tile
constexpr int THREADS = 128;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
    const int sz = pred ? 16 : 0;   // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// smem tile [64 rows][64 bf16] = 128 B rows, 16 B chunks XOR-swizzled by (row & 7)
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
    return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
}

// load a [64 x 64] bf16 tile: rows row0.. of a matrix with row stride ld (elements); rows >= nrows zero-filled
__device__ __forceinline__ void load_tile_async(uint32_t smem_tile, const __nv_bfloat16* base, long long ld,
                                                int row0, int nrows) {
    for (int i = threadIdx.x; i < 64 * 8; i += THREADS) {
        const int r = i >> 3, ch = i & 7;
        const bool ok = row0 + r < nrows;
        const __nv_bfloat16* src = base + (long long)(ok ? row0 + r : 0) * ld + ch * 8;
        cp_async16(smem_tile + tile_off(r, ch), src, ok);
    }
}

__global__ void __launch_bounds__(THREADS)
attention_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, int N, int H, float scale_log2,
                     __nv_bfloat16* __restrict__ out, float* __restrict__ lse) {
    __shared__ __align__(128) uint8_t s_q[BQ * 128];
    __shared__ __align__(128) uint8_t s_k[2][BK * 128];
    __shared__ __align__(128) uint8_t s_v[2][BK * 128];

    const int D = H * HD;
    const long long ld = 3ll * D;
    const int num_qt = (N + BQ - 1) / BQ;
    const int bh = blockIdx.x / num_qt;
    const int b = bh / H, h = bh % H;
    const int q0 = (blockIdx.x % num_qt) * BQ;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;

    const __nv_bfloat16* qbase = qkv + (long long)b * N * ld + h * HD;
    const __nv_bfloat16* kbase = qbase + D;
    const __nv_bfloat16* vbase = qbase + 2 * D;

    const uint32_t sq = smem_u32(s_q);
    const uint32_t sk[2] = {smem_u32(s_k[0]), smem_u32(s_k[1])};
    const uint32_t sv[2] = {smem_u32(s_v[0]), smem_u32(s_v[1])};

    const int num_kt = (N + BK - 1) / BK;
    load_tile_async(sq, qbase, ld, q0, N);
    load_tile_async(sk[0], kbase, ld, 0, N);
    load_tile_async(sv[0], vbase, ld, 0, N);
    cp_async_commit();

    uint32_t qf[4][4];
    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
    float row_max[2] = {-INFINITY, -INFINITY};
    float row_sum[2] = {0.f, 0.f};

    for (int kt = 0; kt < num_kt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < num_kt) {
            load_tile_async(sk[buf ^ 1], kbase, ld, (kt + 1) * BK, N);
            load_tile_async(sv[buf ^ 1], vbase, ld, (kt + 1) * BK, N);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (kt == 0) {
            // Q fragments of this warp's 16 rows (kept in registers for all key tiles)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int chunk = ks * 2 + (lane >> 4);
                ldsm_x4(sq + tile_off(row, chunk), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
            }
        }
        // S = Q K^T  (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {     // pairs of 8-key tiles
                uint32_t b0, b1, b2, b3;
                const int row = jp * 16 + (lane & 7) + (lane >> 4) * 8;
                const int chunk = ks * 2 + ((lane >> 3) & 1);
                ldsm_x4(sk[buf] + tile_off(row, chunk), b0, b1, b2, b3);
                mma_bf16(s[2 * jp], qf[ks], b0, b1);
                mma_bf16(s[2 * jp + 1], qf[ks], b2, b3);
            }
        }
        // mask keys beyond N, online softmax
        const int key0 = kt * BK;
        float tmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int key = key0 + j * 8 + 2 * t + (e & 1);
                if (key >= N) s[j][e] = -INFINITY;
                tmax[e >> 1] = fmaxf(tmax[e >> 1], s[j][e]);
            }
        }
        float corr[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 1));
            tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 2));
            const float new_max = fmaxf(row_max[r], tmax[r]);
            corr[r] = exp2f((row_max[r] - new_max) * scale_log2);
            row_max[r] = new_max;
            row_sum[r] *= corr[r];
        }
        uint32_t pf[4][4];   // P as A-operand fragments for the 4 k-steps (16 keys each)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float m0 = row_max[0] * scale_log2, m1 = row_max[1] * scale_log2;
            const float p0 = exp2f(s[j][0] * scale_log2 - m0);
            const float p1 = exp2f(s[j][1] * scale_log2 - m0);
            const float p2 = exp2f(s[j][2] * scale_log2 - m1);
            const float p3 = exp2f(s[j][3] * scale_log2 - m1);
            row_sum[0] += p0 + p1;
            row_sum[1] += p2 + p3;
            // accumulator layout of n-tile j -> A fragment: k-step j/2, regs {0,1} for even j, {2,3} for odd j
            pf[j >> 1][(j & 1) * 2 + 0] = pack_bf16(p0, p1);
            pf[j >> 1][(j & 1) * 2 + 1] = pack_bf16(p2, p3);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            o[j][0] *= corr[0];
            o[j][1] *= corr[0];
            o[j][2] *= corr[1];
            o[j][3] *= corr[1];
        }
        // O += P V
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {          // 16 keys
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {      // pairs of 8-dim tiles
                uint32_t b0, b1, b2, b3;
                const int row = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int chunk = jp * 2 + (lane >> 4);
                ldsm_x4_trans(sv[buf] + tile_off(row, chunk), b0, b1, b2, b3);
                mma_bf16(o[2 * jp], pf[ks], b0, b1);
                mma_bf16(o[2 * jp + 1], pf[ks], b2, b3);
            }
        }
        __syncthreads();   // everyone done with buf before it is refilled two iterations later
    }

    // finalise: divide by the row sums, store bf16; optional log-sum-exp for the backward
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        row_sum[r] += __shfl_xor_sync(0xffffffffu, row_sum[r], 1);
        row_sum[r] += __shfl_xor_sync(0xffffffffu, row_sum[r], 2);
    }
    const float inv0 = 1.0f / row_sum[0], inv1 = 1.0f / row_sum[1];
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
    __nv_bfloat16* ob = out + (long long)b * N * D + h * HD;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int col = j * 8 + 2 * t;
        if (r0 < N) *reinterpret_cast<uint32_t*>(ob + (long long)r0 * D + col) = pack_bf16(o[j][0] * inv0, o[j][1] * inv0);
        if (r1 < N) *reinterpret_cast<uint32_t*>(ob + (long long)r1 * D + col) = pack_bf16(o[j][2] * inv1, o[j][3] * inv1);
    }
    if (lse != nullptr && t == 0) {
        const float ln2 = 0.6931471805599453f;
        float* l = lse + ((long long)b * H + h) * N;
        if (r0 < N) l[r0] = (row_max[0] * scale_log2 + log2f(row_sum[0])) * ln2;
        if (r1 < N) l[r1] = (row_max[1] * scale_log2 + log2f(row_sum[1])) * ln2;
    }
}


// ============================================================================================
// Backward (student only).  Two deterministic passes, both recompute P from the saved
// log-sum-exp:   P = exp(S*scale - lse),  dS = P o (dP - delta) * scale,  delta = rowsum(dO o O)
//   dq pass : CTA = 64 query rows, loops over key tiles       dQ = dS K
//   dkv pass: CTA = 64 keys,       loops over query tiles     dV = P^T dO,  dK = dS^T Q
// The RoPE of the forward epilogue (rope.py:148-164) is undone on dQ / dK before the store, so the
// result is the gradient w.r.t. the raw projections q = x Wq^T + b (eva_vit_model.py:177-204).
// ============================================================================================
constexpr float LOG2E = 1.4426950408889634f;

__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                                  long long rows, int N, int H, float* __restrict__ delta) {
    // one warp per (token row, head)
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= rows * H) return;
    const long long row = w / H;
    const int h = (int)(w % H);
    const long long off = row * (long long)(H * HD) + h * HD + lane * 2;
    const float2 a = unpack_bf16(*reinterpret_cast<const uint32_t*>(o + off));
    const float2 b = unpack_bf16(*reinterpret_cast<const uint32_t*>(d_o + off));
    const float s = warp_sum(a.x * b.x + a.y * b.y);
    if (lane == 0) {
        const long long bi = row / N;
        const int n = (int)(row % N);
        delta[(bi * H + h) * N + n] = s;
    }
}

// inverse rotation of one adjacent (even, odd) pair:  forward was y0 = x0 c0 - x1 s0, y1 = x1 c1 + x0 s1
__device__ __forceinline__ void unrope_pair(float& d0, float& d1, const float* __restrict__ cs_row,
                                            const float* __restrict__ sn_row, int d) {
    const float c0 = cs_row[d], c1 = cs_row[d + 1], s0 = sn_row[d], s1 = sn_row[d + 1];
    const float x0 = d0 * c0 + d1 * s1;
    const float x1 = d1 * c1 - d0 * s0;
    d0 = x0;
    d1 = x1;
}

constexpr int BWD_SMEM = (2 + 4) * 64 * 128 + 4 * 64 * 4;   // 2 resident tiles + 2x2 streamed tiles + lse/delta

__global__ void __launch_bounds__(THREADS)
attention_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ d_out,
                        const float* __restrict__ lse, const float* __restrict__ delta, int N, int H,
                        float scale, const float* __restrict__ rope_cos, const float* __restrict__ rope_sin,
                        __nv_bfloat16* __restrict__ dqkv) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int D = H * HD;
    const long long ld = 3ll * D;
    const int num_qt = (N + BQ - 1) / BQ;
    const int bh = blockIdx.x / num_qt;
    const int b = bh / H, h = bh % H;
    const int q0 = (blockIdx.x % num_qt) * BQ;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;

    const __nv_bfloat16* qbase = qkv + (long long)b * N * ld + h * HD;
    const __nv_bfloat16* kbase = qbase + D;
    const __nv_bfloat16* vbase = qbase + 2 * D;
    const __nv_bfloat16* dobase = d_out + (long long)b * N * D + h * HD;

    const uint32_t sq = smem_u32(smem);
    const uint32_t sdo = sq + 64 * 128;
    const uint32_t sk[2] = {sdo + 64 * 128, sdo + 3 * 64 * 128};
    const uint32_t sv[2] = {sdo + 2 * 64 * 128, sdo + 4 * 64 * 128};

    const int num_kt = (N + BK - 1) / BK;
    load_tile_async(sq, qbase, ld, q0, N);
    load_tile_async(sdo, dobase, D, q0, N);
    load_tile_async(sk[0], kbase, ld, 0, N);
    load_tile_async(sv[0], vbase, ld, 0, N);
    cp_async_commit();

    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
    const float* lrow = lse + ((long long)b * H + h) * N;
    const float* drow = delta + ((long long)b * H + h) * N;
    const float lse2[2] = {r0 < N ? lrow[r0] * LOG2E : 0.f, r1 < N ? lrow[r1] * LOG2E : 0.f};
    const float dlt[2] = {r0 < N ? drow[r0] : 0.f, r1 < N ? drow[r1] : 0.f};
    const float scale_log2 = scale * LOG2E;

    uint32_t qf[4][4], dof[4][4];
    float dq[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;

    for (int kt = 0; kt < num_kt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < num_kt) {
            load_tile_async(sk[buf ^ 1], kbase, ld, (kt + 1) * BK, N);
            load_tile_async(sv[buf ^ 1], vbase, ld, (kt + 1) * BK, N);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (kt == 0) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int chunk = ks * 2 + (lane >> 4);
                ldsm_x4(sq + tile_off(row, chunk), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
                ldsm_x4(sdo + tile_off(row, chunk), dof[ks][0], dof[ks][1], dof[ks][2], dof[ks][3]);
            }
        }
        const int key0 = kt * BK;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {          // 16-key sub-block = k-step of the dQ product
            float s2[2][4], dp2[2][4];
#pragma unroll
            for (int j = 0; j < 2; ++j) s2[j][0] = s2[j][1] = s2[j][2] = s2[j][3] = dp2[j][0] = dp2[j][1] = dp2[j][2] = dp2[j][3] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                uint32_t b0, b1, b2, b3;
                const int row = kk * 16 + (lane & 7) + (lane >> 4) * 8;
                const int chunk = ks * 2 + ((lane >> 3) & 1);
                ldsm_x4(sk[buf] + tile_off(row, chunk), b0, b1, b2, b3);
                mma_bf16(s2[0], qf[ks], b0, b1);
                mma_bf16(s2[1], qf[ks], b2, b3);
                ldsm_x4(sv[buf] + tile_off(row, chunk), b0, b1, b2, b3);
                mma_bf16(dp2[0], dof[ks], b0, b1);
                mma_bf16(dp2[1], dof[ks], b2, b3);
            }
            uint32_t dsf[4];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float ds[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int key = key0 + kk * 16 + j * 8 + 2 * t + (e & 1);
                    const float p = key < N ? exp2f(s2[j][e] * scale_log2 - lse2[e >> 1]) : 0.f;
                    ds[e] = p * (dp2[j][e] - dlt[e >> 1]) * scale;
                }
                dsf[j * 2 + 0] = pack_bf16(ds[0], ds[1]);
                dsf[j * 2 + 1] = pack_bf16(ds[2], ds[3]);
            }
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {      // dQ += dS K   (B = K[key k][dim n], transposed load)
                uint32_t b0, b1, b2, b3;
                const int row = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int chunk = jp * 2 + (lane >> 4);
                ldsm_x4_trans(sk[buf] + tile_off(row, chunk), b0, b1, b2, b3);
                mma_bf16(dq[2 * jp], dsf, b0, b1);
                mma_bf16(dq[2 * jp + 1], dsf, b2, b3);
            }
        }
        __syncthreads();
    }

    __nv_bfloat16* ob = dqkv + (long long)b * N * ld + h * HD;     // q section
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int col = j * 8 + 2 * t;
        if (rope_cos != nullptr) {
            if (r0 < N && r0 > 0) unrope_pair(dq[j][0], dq[j][1], rope_cos + (long long)(r0 - 1) * HD, rope_sin + (long long)(r0 - 1) * HD, col);
            if (r1 < N && r1 > 0) unrope_pair(dq[j][2], dq[j][3], rope_cos + (long long)(r1 - 1) * HD, rope_sin + (long long)(r1 - 1) * HD, col);
        }
        if (r0 < N) *reinterpret_cast<uint32_t*>(ob + (long long)r0 * ld + col) = pack_bf16(dq[j][0], dq[j][1]);
        if (r1 < N) *reinterpret_cast<uint32_t*>(ob + (long long)r1 * ld + col) = pack_bf16(dq[j][2], dq[j][3]);
    }
}

__device__ __forceinline__ void load_vec64_async(uint32_t dst, const float* base, int row0, int nrows) {
    // 64 floats (one per query row of the tile); rows beyond nrows zero-filled
    if (threadIdx.x < 64) {
        const bool ok = row0 + (int)threadIdx.x < nrows;
        const float* src = base + (ok ? row0 + threadIdx.x : 0);
        const int sz = ok ? 4 : 0;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + threadIdx.x * 4), "l"(src), "r"(sz) : "memory");
    }
}

__global__ void __launch_bounds__(THREADS)
attention_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ d_out,
                         const float* __restrict__ lse, const float* __restrict__ delta, int N, int H,
                         float scale, const float* __restrict__ rope_cos, const float* __restrict__ rope_sin,
                         __nv_bfloat16* __restrict__ dqkv) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int D = H * HD;
    const long long ld = 3ll * D;
    const int num_kt = (N + BK - 1) / BK;
    const int bh = blockIdx.x / num_kt;
    const int b = bh / H, h = bh % H;
    const int k0 = (blockIdx.x % num_kt) * BK;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;

    const __nv_bfloat16* qbase = qkv + (long long)b * N * ld + h * HD;
    const __nv_bfloat16* kbase = qbase + D;
    const __nv_bfloat16* vbase = qbase + 2 * D;
    const __nv_bfloat16* dobase = d_out + (long long)b * N * D + h * HD;
    const float* lrow = lse + ((long long)b * H + h) * N;
    const float* drow = delta + ((long long)b * H + h) * N;

    const uint32_t sk = smem_u32(smem);
    const uint32_t sv = sk + 64 * 128;
    const uint32_t sq[2] = {sv + 64 * 128, sv + 3 * 64 * 128};
    const uint32_t sdo[2] = {sv + 2 * 64 * 128, sv + 4 * 64 * 128};
    const uint32_t sl_addr = sv + 5 * 64 * 128;                  // [2][64] lse, then [2][64] delta
    const float* sl = reinterpret_cast<const float*>(smem + 6 * 64 * 128);

    const int num_qt = (N + BQ - 1) / BQ;
    load_tile_async(sk, kbase, ld, k0, N);
    load_tile_async(sv, vbase, ld, k0, N);
    load_tile_async(sq[0], qbase, ld, 0, N);
    load_tile_async(sdo[0], dobase, D, 0, N);
    load_vec64_async(sl_addr, lrow, 0, N);
    load_vec64_async(sl_addr + 2 * 64 * 4, drow, 0, N);
    cp_async_commit();

    const float scale_log2 = scale * LOG2E;
    const int key_r0 = k0 + warp * 16 + g, key_r1 = key_r0 + 8;
    uint32_t kf[4][4], vf[4][4];
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;

    for (int qt = 0; qt < num_qt; ++qt) {
        const int buf = qt & 1;
        if (qt + 1 < num_qt) {
            load_tile_async(sq[buf ^ 1], qbase, ld, (qt + 1) * BQ, N);
            load_tile_async(sdo[buf ^ 1], dobase, D, (qt + 1) * BQ, N);
            load_vec64_async(sl_addr + (buf ^ 1) * 64 * 4, lrow, (qt + 1) * BQ, N);
            load_vec64_async(sl_addr + (2 + (buf ^ 1)) * 64 * 4, drow, (qt + 1) * BQ, N);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (qt == 0) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int chunk = ks * 2 + (lane >> 4);
                ldsm_x4(sk + tile_off(row, chunk), kf[ks][0], kf[ks][1], kf[ks][2], kf[ks][3]);
                ldsm_x4(sv + tile_off(row, chunk), vf[ks][0], vf[ks][1], vf[ks][2], vf[ks][3]);
            }
        }
        const float* s_lse = sl + buf * 64;
        const float* s_dlt = sl + (2 + buf) * 64;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {          // 16-query sub-block
            float st[2][4], dpt[2][4];
#pragma unroll
            for (int j = 0; j < 2; ++j) st[j][0] = st[j][1] = st[j][2] = st[j][3] = dpt[j][0] = dpt[j][1] = dpt[j][2] = dpt[j][3] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {      // S^T = K Q^T, dP^T = V dO^T
                uint32_t b0, b1, b2, b3;
                const int row = qq * 16 + (lane & 7) + (lane >> 4) * 8;
                const int chunk = ks * 2 + ((lane >> 3) & 1);
                ldsm_x4(sq[buf] + tile_off(row, chunk), b0, b1, b2, b3);
                mma_bf16(st[0], kf[ks], b0, b1);
                mma_bf16(st[1], kf[ks], b2, b3);
                ldsm_x4(sdo[buf] + tile_off(row, chunk), b0, b1, b2, b3);
                mma_bf16(dpt[0], vf[ks], b0, b1);
                mma_bf16(dpt[1], vf[ks], b2, b3);
            }
            uint32_t pf[4], dsf[4];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float p[4], ds[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int qi = qq * 16 + j * 8 + 2 * t + (e & 1);      // query index within the tile
                    const int key = (e >> 1) ? key_r1 : key_r0;
                    p[e] = key < N ? exp2f(st[j][e] * scale_log2 - s_lse[qi] * LOG2E) : 0.f;
                    ds[e] = p[e] * (dpt[j][e] - s_dlt[qi]) * scale;
                }
                pf[j * 2 + 0] = pack_bf16(p[0], p[1]);
                pf[j * 2 + 1] = pack_bf16(p[2], p[3]);
                dsf[j * 2 + 0] = pack_bf16(ds[0], ds[1]);
                dsf[j * 2 + 1] = pack_bf16(ds[2], ds[3]);
            }
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {      // dV += P^T dO ; dK += dS^T Q   (B[k=query][n=dim], transposed loads)
                uint32_t b0, b1, b2, b3;
                const int row = qq * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int chunk = jp * 2 + (lane >> 4);
                ldsm_x4_trans(sdo[buf] + tile_off(row, chunk), b0, b1, b2, b3);
                mma_bf16(dv[2 * jp], pf, b0, b1);
                mma_bf16(dv[2 * jp + 1], pf, b2, b3);
                ldsm_x4_trans(sq[buf] + tile_off(row, chunk), b0, b1, b2, b3);
                mma_bf16(dk[2 * jp], dsf, b0, b1);
                mma_bf16(dk[2 * jp + 1], dsf, b2, b3);
            }
        }
        __syncthreads();
    }

    __nv_bfloat16* okb = dqkv + (long long)b * N * ld + D + h * HD;        // k section
    __nv_bfloat16* ovb = okb + D;                                          // v section
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int col = j * 8 + 2 * t;
        if (rope_cos != nullptr) {
            if (key_r0 < N && key_r0 > 0) unrope_pair(dk[j][0], dk[j][1], rope_cos + (long long)(key_r0 - 1) * HD, rope_sin + (long long)(key_r0 - 1) * HD, col);
            if (key_r1 < N && key_r1 > 0) unrope_pair(dk[j][2], dk[j][3], rope_cos + (long long)(key_r1 - 1) * HD, rope_sin + (long long)(key_r1 - 1) * HD, col);
        }
        if (key_r0 < N) {
            *reinterpret_cast<uint32_t*>(okb + (long long)key_r0 * ld + col) = pack_bf16(dk[j][0], dk[j][1]);
            *reinterpret_cast<uint32_t*>(ovb + (long long)key_r0 * ld + col) = pack_bf16(dv[j][0], dv[j][1]);
        }
        if (key_r1 < N) {
            *reinterpret_cast<uint32_t*>(okb + (long long)key_r1 * ld + col) = pack_bf16(dk[j][2], dk[j][3]);
            *reinterpret_cast<uint32_t*>(ovb + (long long)key_r1 * ld + col) = pack_bf16(dv[j][2], dv[j][3]);
        }
    }
}

}  // namespace attn
}  // namespace cs

namespace cs {
int attention_fwd_tc(const void* qkv, int B, int N, int H, float scale, void* out, float* lse, float* row_stats,
                     cudaStream_t st);
int attention_fwd_tc3(const void* qkv, int B, int N, int H, float scale, void* out, float* lse, float* row_stats,
                      cudaStream_t st);
int attention_fwd_tc_long(const void* qkv, int B, int N, int H, float scale, void* out, float* lse, float* row_stats,
                          cudaStream_t st);
int attention_fwd_tc4(const void* qkv, int B, int N, int H, float scale, void* out, float* lse, float* row_stats,
                      cudaStream_t st);
int attention_bwd_tc(const void* qkv, const void* d_out, const float* lse, const float* delta, int B, int N, int H,
                     float scale, const float* rope_cos, const float* rope_sin, void* dqkv, cudaStream_t st);
}

// opt-in switch: set and neither empty nor "0"
static bool env_on(const char* name) {
    const char* v = getenv(name);
    return v != nullptr && v[0] != '\0' && !(v[0] == '0' && v[1] == '\0');
}

extern "C" int cs_attention_fwd(const void* qkv_bf16, int B, int N, int H, float scale, void* out_bf16,
                                float* lse, float* row_stats, void* stream) {
    using namespace cs;
    using namespace cs::attn;
    CS_CHECK_ARG(qkv_bf16 && out_bf16, "cs_attention_fwd: null pointer");
    CS_CHECK_ARG(B > 0 && N > 0 && H > 0, "cs_attention_fwd: bad shape");
    {
        // tcgen05 kernels: the single-pass kernel for N <= 224 (B/16: 197 tokens), the streaming ping-pong kernel for
        // longer sequences (ViT-L/14-336: 577, the 1024 px student: 4097).  CS_ATTN_LEGACY=1 selects the mma.sync kernel
        // below for A/B measurements; CS_ATTN_FORCE_LONG=1 routes short sequences through the streaming kernel.
        static const bool legacy = env_on("CS_ATTN_LEGACY");
        if (!legacy && !env_on("CS_ATTN_FORCE_LONG") && !env_on("CS_ATTN_V1") && !env_on("CS_ATTN_V3")) {   // 129 <= N <= 208: two ping-pong slots (attention_tc4.cu)
            const int rc = attention_fwd_tc4(qkv_bf16, B, N, H, scale, out_bf16, lse, row_stats, (cudaStream_t)stream);
            if (rc != CS_ERR_UNSUPPORTED) return rc;
        }
        if (!legacy && !env_on("CS_ATTN_FORCE_LONG") && !env_on("CS_ATTN_V1")) {     // N <= 208: 16-softmax-warp kernel (attention_tc3.cu)
            const int rc = attention_fwd_tc3(qkv_bf16, B, N, H, scale, out_bf16, lse, row_stats, (cudaStream_t)stream);
            if (rc != CS_ERR_UNSUPPORTED) return rc;
        }
        if (!legacy && !env_on("CS_ATTN_FORCE_LONG") && row_stats == nullptr && env_on("CS_ATTN_V1")) {   // first generation (A/B timing)
            const int rc = attention_fwd_tc(qkv_bf16, B, N, H, scale, out_bf16, lse, nullptr, (cudaStream_t)stream);
            if (rc != CS_ERR_UNSUPPORTED) return rc;
        }
        if (!legacy) {
            const int rc = attention_fwd_tc_long(qkv_bf16, B, N, H, scale, out_bf16, lse, row_stats, (cudaStream_t)stream);
            if (rc != CS_ERR_UNSUPPORTED) return rc;
        }
        CS_CHECK_ARG(row_stats == nullptr, "cs_attention_fwd: row_stats is only produced by the tcgen05 kernels");
    }
    const long long blocks = (long long)B * H * ceil_div(N, BQ);
    CS_CHECK_ARG(blocks < (1ll << 31), "cs_attention_fwd: grid too large");
    dim3 grid((unsigned)blocks);
    const float scale_log2 = scale * 1.4426950408889634f;
    attention_fwd_kernel<<<grid, THREADS, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)qkv_bf16, N, H, scale_log2,
                                                                     (__nv_bfloat16*)out_bf16, lse);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

extern "C" int cs_attention_bwd(const void* qkv_bf16, const void* out_bf16, const void* d_out_bf16, const float* lse,
                                int B, int N, int H, float scale, const float* rope_cos, const float* rope_sin,
                                float* delta_ws, void* dqkv_bf16, void* stream) {
    using namespace cs;
    using namespace cs::attn;
    CS_CHECK_ARG(qkv_bf16 && out_bf16 && d_out_bf16 && lse && delta_ws && dqkv_bf16, "cs_attention_bwd: null pointer");
    CS_CHECK_ARG(B > 0 && N > 0 && H > 0, "cs_attention_bwd: bad shape");
    CS_CHECK_ARG((rope_cos == nullptr) == (rope_sin == nullptr), "cs_attention_bwd: need both rope tables or none");
    const long long blocks = (long long)B * H * ceil_div(N, BQ);
    CS_CHECK_ARG(blocks < (1ll << 31), "cs_attention_bwd: grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    static bool configured = false;
    if (!configured) {
        CS_CUDA(cudaFuncSetAttribute(attention_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
        CS_CUDA(cudaFuncSetAttribute(attention_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
        configured = true;
    }
    const long long rows = (long long)B * N;
    attn_delta_kernel<<<ceil_div(rows * H, 8), 256, 0, st>>>((const __nv_bfloat16*)out_bf16, (const __nv_bfloat16*)d_out_bf16,
                                                             rows, N, H, delta_ws);
    CS_LAUNCH_CHECK();
    // tcgen05 backward (attention_bwd_tc.cu); CS_ATTN_LEGACY=1 selects the mma.sync kernels below for A/B measurements
    if (!env_on("CS_ATTN_LEGACY")) {
        const int rc = attention_bwd_tc(qkv_bf16, d_out_bf16, lse, delta_ws, B, N, H, scale, rope_cos, rope_sin, dqkv_bf16, st);
        if (rc != CS_ERR_UNSUPPORTED) return rc;
    }
    attention_bwd_dq_kernel<<<(unsigned)blocks, THREADS, BWD_SMEM, st>>>(
        (const __nv_bfloat16*)qkv_bf16, (const __nv_bfloat16*)d_out_bf16, lse, delta_ws, N, H, scale, rope_cos, rope_sin,
        (__nv_bfloat16*)dqkv_bf16);
    CS_LAUNCH_CHECK();
    attention_bwd_dkv_kernel<<<(unsigned)blocks, THREADS, BWD_SMEM, st>>>(
        (const __nv_bfloat16*)qkv_bf16, (const __nv_bfloat16*)d_out_bf16, lse, delta_ws, N, H, scale, rope_cos, rope_sin,
        (__nv_bfloat16*)dqkv_bf16);
    CS_LAUNCH_CHECK();
    return CS_OK;
}
