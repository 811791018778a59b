// Attention for the CLS query only: out[b, h, :] = softmax(q_cls k^T * scale) v for ONE query row per (image, head).
//
// encode_image (clipself.py:37-38 -> eva_vit_model.py:565-569) returns head(norm(x)[:, 0]): of the LAST block's output only
// the CLS row is ever read.  The frozen teacher therefore runs its last block on the CLS rows alone — K and V still need
// every token (the QKV GEMM is unchanged), but attention, proj, the SwiGLU MLP and w3 shrink from n x N rows to n rows
// (csrc/tower.cu: block_cls_tail).  This kernel is the attention of that block: 2 x N x 64 multiply-adds per head on the
// CUDA cores, HBM-bound on the K / V rows (N x 256 B per head), one warp per (image, head):
//   scores: lane j handles keys j, j+32, ... (a key row is one 128-byte line), q replicated in registers;
//   P = exp2((s - max) * scale * log2e) rounded to bf16 and normalised by the sum of the ROUNDED values, exactly the
//   rounding points of the tcgen05 kernels (oracle/device_arith_oracle.py: softmax_pv);
//   output: lane l owns dims 2l, 2l+1 (a value row is read as one coalesced 128-byte line), p_j broadcast by shuffle.
// Also emits the per-(row, head, 16-dim quarter) sum / sum of squares the folded inner_attn_ln needs, like the other
// attention kernels.  150-160 us at 512 crops x 197 tokens x 12 heads (tools/attn_cls_one.py; 0.6 ms of a cfg2 step); a variant
// with 16-byte value loads of four keys per instruction (8x the bytes in flight in the P V phase) measured the same, so the
// per-lane key-row reads of the score phase are what is left to restructure.
#include "common.cuh"

namespace cs {
namespace attn_cls {

constexpr int HD = 64;
constexpr int WARPS = 8;
constexpr int MAX_SLOTS = 32;           // keys per lane: N <= 1024 (checked on the host; the teacher's towers have 197 / 257 / 577 tokens)

__device__ __forceinline__ float ex2(float x) {     // the same instruction the tcgen05 kernels use
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(WARPS * 32)
attention_cls_kernel(const __nv_bfloat16* __restrict__ qkv, int B, int N, int H, float scale_log2, __nv_bfloat16* __restrict__ out,
                     float* __restrict__ row_stats) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * WARPS + warp;
    if (item >= B * H) return;
    const int b = item / H, h = item % H;
    const int D = H * HD;
    const long long ld = 3ll * D;
    const __nv_bfloat16* base = qkv + (long long)b * N * ld + h * HD;
    // q of the CLS row, replicated in every lane (8 broadcast 16-byte loads)
    float q[HD];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint4 u = *reinterpret_cast<const uint4*>(base + 8 * i);
        const float2 a = unpack_bf16(u.x), c = unpack_bf16(u.y), d = unpack_bf16(u.z), e = unpack_bf16(u.w);
        q[8 * i] = a.x; q[8 * i + 1] = a.y; q[8 * i + 2] = c.x; q[8 * i + 3] = c.y;
        q[8 * i + 4] = d.x; q[8 * i + 5] = d.y; q[8 * i + 6] = e.x; q[8 * i + 7] = e.y;
    }
    // scores of this lane's keys
    float s[MAX_SLOTS];
    float mx = -INFINITY;
    const int slots = (N + 31) >> 5;
#pragma unroll 1
    for (int t = 0; t < slots; ++t) {
        const int j = t * 32 + lane;
        float acc = -INFINITY;
        if (j < N) {
            const __nv_bfloat16* kr = base + D + (long long)j * ld;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint4 u = *reinterpret_cast<const uint4*>(kr + 8 * i);
                const float2 x0 = unpack_bf16(u.x), x1 = unpack_bf16(u.y), x2 = unpack_bf16(u.z), x3 = unpack_bf16(u.w);
                a0 = fmaf(q[8 * i], x0.x, a0);     a1 = fmaf(q[8 * i + 1], x0.y, a1);
                a0 = fmaf(q[8 * i + 2], x1.x, a0); a1 = fmaf(q[8 * i + 3], x1.y, a1);
                a0 = fmaf(q[8 * i + 4], x2.x, a0); a1 = fmaf(q[8 * i + 5], x2.y, a1);
                a0 = fmaf(q[8 * i + 6], x3.x, a0); a1 = fmaf(q[8 * i + 7], x3.y, a1);
            }
            acc = a0 + a1;
        }
#pragma unroll
        for (int u = 0; u < MAX_SLOTS; ++u)
            if (u == t) s[u] = acc;               // static register indexing
        mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    // P (bf16-rounded) and its sum
    float lsum = 0.f;
    const float mxs = mx * scale_log2;
#pragma unroll
    for (int u = 0; u < MAX_SLOTS; ++u) {
        if (u < slots) {
            const float p = (u * 32 + lane < N) ? ex2(fmaf(s[u], scale_log2, -mxs)) : 0.f;
            const float pr = __bfloat162float(__float2bfloat16_rn(p));
            s[u] = pr;
            lsum += pr;
        }
    }
    lsum = warp_sum(lsum);
    // O: lane l owns dims 2l, 2l+1
    float o0 = 0.f, o1 = 0.f;
    const __nv_bfloat16* vbase = base + 2 * D + 2 * lane;
#pragma unroll
    for (int u = 0; u < MAX_SLOTS; ++u) {
        if (u < slots) {
            const int jn = min(32, N - u * 32);
#pragma unroll 4
            for (int src = 0; src < jn; ++src) {
                const float p = __shfl_sync(0xffffffffu, s[u], src);
                const float2 v = unpack_bf16(*reinterpret_cast<const uint32_t*>(vbase + (long long)(u * 32 + src) * ld));
                o0 = fmaf(p, v.x, o0);
                o1 = fmaf(p, v.y, o1);
            }
        }
    }
    const float inv = 1.0f / lsum;
    o0 *= inv;
    o1 *= inv;
    *reinterpret_cast<uint32_t*>(out + (long long)b * D + h * HD + 2 * lane) = pack_bf16(o0, o1);
    if (row_stats != nullptr) {
        float s1 = o0 + o1, s2 = fmaf(o0, o0, o1 * o1);
#pragma unroll
        for (int m = 1; m < 8; m <<= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, m);
            s2 += __shfl_xor_sync(0xffffffffu, s2, m);
        }
        if ((lane & 7) == 0)
            *reinterpret_cast<float2*>(row_stats + ((long long)b * (4 * H) + 4 * h + (lane >> 3)) * 2) = make_float2(s1, s2);
    }
}

}  // namespace attn_cls
}  // namespace cs

// qkv [B*N, 3*H*64] bf16 (q|k|v, RoPE already applied by the QKV epilogue; the CLS row is row 0 of each image) ->
// out_cls [B, H*64] bf16 and, optionally, row_stats_cls [B, 4H, 2].  Replaces eva_vit_model.py:206-217 for the last block
// of a tower whose output is read at the CLS token only.
extern "C" int cs_attention_cls_fwd(const void* qkv_bf16, int B, int N, int H, float scale, void* out_cls_bf16, float* row_stats_cls,
                                    void* stream) {
    using namespace cs;
    using namespace cs::attn_cls;
    CS_CHECK_ARG(qkv_bf16 && out_cls_bf16, "cs_attention_cls_fwd: null pointer");
    CS_CHECK_ARG(B > 0 && N > 0 && H > 0 && N <= 32 * MAX_SLOTS, "cs_attention_cls_fwd: bad shape (N <= %d)", 32 * MAX_SLOTS);
    CS_CHECK_ARG((uintptr_t)qkv_bf16 % 16 == 0, "cs_attention_cls_fwd: qkv must be 16 B aligned");
    const long long items = (long long)B * H;
    attention_cls_kernel<<<(unsigned)((items + WARPS - 1) / WARPS), WARPS * 32, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)qkv_bf16, B, N, H, scale * 1.4426950408889634f, (__nv_bfloat16*)out_cls_bf16, row_stats_cls);
    CS_LAUNCH_CHECK();
    return CS_OK;
}
