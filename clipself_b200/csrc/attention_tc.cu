// tcgen05 softmax attention forward for head_dim 64 and N <= 224 tokens (EVA02-B/16: N = 197).
//   S = Q K^T  (tcgen05.mma, accumulator in TMEM)  ->  softmax in registers (one thread per query row,
//   tcgen05.ld)  ->  P (bf16) into swizzled shared memory  ->  O = P V (tcgen05.mma, V consumed
//   MN-major straight from its [key][dim] layout)  ->  O / rowsum -> bf16.
// Replaces xformers.memory_efficient_attention at eva_vit_model.py:206-217 for the teacher's crops.
//
// Persistent CTA per SM, 320 threads: warp 0 TMA producer, warp 1 MMA issuer + TMEM allocator,
// warps 2-9 softmax + epilogue: two warps per TMEM lane quarter split each row's key columns (and the
// output dims), exchanging row max / row sum through shared memory.  Work item = one (image, head); its K and V tiles are loaded once and
// reused by the ceil(N/128) query tiles.  S is double buffered in TMEM so Q K^T of the next tile runs
// under the softmax of the current one.
#include "tc_common.cuh"

namespace cs {
namespace attn_tc {
using namespace cs::tc;

constexpr int HD = 64;
constexpr int BM = 128;
constexpr int THREADS = 320;           // TMA warp, MMA warp, 8 softmax/epilogue warps (two per TMEM lane quarter)
constexpr int Q_BYTES = BM * 128;
constexpr int P_ATOM_BYTES = BM * 128;      // [128 rows][64 keys] bf16, SWIZZLE_128B K-major
constexpr int MAX_NKP = 224;
constexpr int O_COL = 448;                  // TMEM column of the O accumulator (S buffers at 0 and 224)
constexpr int S_STRIDE = 224;
constexpr int TMEM_COLS = 512;

// generic descriptor: SWIZZLE_128B, version 1, explicit LBO / SBO (bytes)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct Params {
    int B, N, H, nkp, ntm;          // nkp: keys padded to 16; ntm: query tiles per head
    float scale_log2, scale;
    __nv_bfloat16* out;
    float* lse;
    float* row_stats;               // optional [B*N, 2H, 2]: per (row, head, dim-half) sum and sum of squares of the bf16 output
};

__global__ void __launch_bounds__(THREADS, 1)
attention_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                        const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_addr);

    const int kv_bytes = p.nkp * 128;
    const int n_atoms = (p.nkp + 63) / 64;
    const uint32_t sQ = base;                                   // 2 stages
    const uint32_t sK = sQ + 2 * Q_BYTES;                       // 2 stages
    const uint32_t sV = sK + 2 * kv_bytes;                      // 2 stages
    const uint32_t sP = sV + 2 * kv_bytes;                      // n_atoms atoms
    const uint32_t bar = sP + n_atoms * P_ATOM_BYTES;
    uint8_t* sP_ptr = smem + (sP - base);
    // barriers (8 B each)
    auto q_full = [&](int s) { return bar + 8u * s; };
    auto q_empty = [&](int s) { return bar + 8u * (2 + s); };
    auto kv_full = [&](int s) { return bar + 8u * (4 + s); };
    auto kv_empty = [&](int s) { return bar + 8u * (6 + s); };
    auto s_full = [&](int s) { return bar + 8u * (8 + s); };
    auto s_empty = [&](int s) { return bar + 8u * (10 + s); };
    const uint32_t p_full = bar + 8u * 12, p_empty = bar + 8u * 13, o_full = bar + 8u * 14, o_empty = bar + 8u * 15;
    const uint32_t tmem_slot = bar + 8u * 16;
    float* xch = reinterpret_cast<float*>(smem + (bar - base) + 8 * 18);   // [2 halves][128 rows] max, then sums
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = p.H * HD;
    const int n_items = p.B * p.H;
    // items of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total_tiles = my_items * p.ntm;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_kv);
        for (int s = 0; s < 2; ++s) {
            mbar_init(q_full(s), 1);
            mbar_init(q_empty(s), 1);
            mbar_init(kv_full(s), 1);
            mbar_init(kv_empty(s), 1);
            mbar_init(s_full(s), 1);
            mbar_init(s_empty(s), 256);
        }
        mbar_init(p_full, 256);
        mbar_init(p_empty, 1);
        mbar_init(o_full, 1);
        mbar_init(o_empty, 256);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            for (int il = 0; il < my_items; ++il) {
                const int item = blockIdx.x + il * gridDim.x;
                const int b = item / p.H, h = item % p.H;
                const int st = il & 1;
                mbar_wait(kv_empty(st), (uint32_t)((il >> 1) & 1) ^ 1u);
                mbar_arrive_expect_tx(kv_full(st), 2u * kv_bytes);
                tma_load_2d(sK + st * kv_bytes, &map_kv, kv_full(st), D + h * HD, b * p.N);
                tma_load_2d(sV + st * kv_bytes, &map_kv, kv_full(st), 2 * D + h * HD, b * p.N);
                for (int mt = 0; mt < p.ntm; ++mt) {
                    const int tt = il * p.ntm + mt;
                    const int qs = tt & 1;
                    mbar_wait(q_empty(qs), (uint32_t)((tt >> 1) & 1) ^ 1u);
                    mbar_arrive_expect_tx(q_full(qs), Q_BYTES);
                    tma_load_2d(sQ + qs * Q_BYTES, &map_q, q_full(qs), h * HD, b * p.N + mt * BM);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer --------------------------------
        if (lane == 0 && total_tiles > 0) {
            // S: M=128, N=nkp, A/B K-major.  PV: M=128, N=64, A K-major (P), B MN-major (V) -> bit 16
            const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.nkp >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(HD >> 3) << 17) |
                                      ((uint32_t)(BM >> 4) << 24);
            auto issue_s = [&](int tt) {
                const int qs = tt & 1;
                const int il = tt / p.ntm;
                mbar_wait(q_full(qs), (uint32_t)(tt >> 1) & 1u);
                if (tt % p.ntm == 0) mbar_wait(kv_full(il & 1), (uint32_t)(il >> 1) & 1u);
                mbar_wait(s_empty(tt & 1), ((uint32_t)(tt >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint64_t dq = smem_desc(sQ + qs * Q_BYTES, 0, 1024);
                const uint64_t dk = smem_desc(sK + (il & 1) * kv_bytes, 0, 1024);
                const uint32_t d_tmem = tmem_base + (uint32_t)((tt & 1) * S_STRIDE);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_bf16(d_tmem, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, k > 0);
                umma_commit(s_full(tt & 1));
                umma_commit(q_empty(qs));
            };
            issue_s(0);
            for (int tt = 0; tt < total_tiles; ++tt) {
                if (tt + 1 < total_tiles) issue_s(tt + 1);
                const int il = tt / p.ntm;
                mbar_wait(p_full, (uint32_t)tt & 1u);
                mbar_wait(o_empty, ((uint32_t)tt & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t sv = sV + (il & 1) * kv_bytes;
                const int ksteps = p.nkp / 16;
                for (int kk = 0; kk < ksteps; ++kk) {
                    const uint64_t dp = smem_desc(sP + (kk >> 2) * P_ATOM_BYTES + (kk & 3) * 32, 0, 1024);
                    // V tile [keys][64 dims]: MN-major, one 64-wide atom, 16 keys per k-step = 2048 B
                    const uint64_t dv = smem_desc(sv + kk * 2048, (uint32_t)kv_bytes, 1024);
                    umma_bf16(tmem_base + O_COL, dp, dv, idesc_pv, kk > 0);
                }
                umma_commit(o_full);
                umma_commit(p_empty);
                if (tt % p.ntm == p.ntm - 1) umma_commit(kv_empty(il & 1));
            }
        }
    } else {
        // ------------------------------ softmax + epilogue ------------------------
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;                          // 0: first half of the key chunks, 1: the rest
        const int r = quarter * 32 + lane;                         // row of the tile owned by this thread
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const int nchunks = (p.nkp + 31) / 32;
        const int c_begin = half == 0 ? 0 : (nchunks + 1) / 2;
        const int c_end = half == 0 ? (nchunks + 1) / 2 : nchunks;
        const int n_keys = p.N, full_chunks = p.N / 32;
        const float scale_log2 = p.scale_log2;
        float* xmax = xch;                                         // [2][128]
        float* xsum = xch + 256;                                   // [2][128]
        // row max of one tile (this thread's share of the key columns, then exchanged with the partner warp)
        auto row_max = [&](int tt) -> float {
            mbar_wait(s_full(tt & 1), (uint32_t)(tt >> 1) & 1u);
            tc_fence_after();
            const uint32_t ts = tmem_base + lane_addr + (uint32_t)((tt & 1) * S_STRIDE);
            float mx = -INFINITY;
            for (int c = c_begin; c < c_end; ++c) {
                uint32_t v[32];
                tmem_ld32(ts + c * 32, v);
                tmem_ld_wait();
                if (c < full_chunks) {                    // all 32 keys valid: no per-element predicate
#pragma unroll
                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c * 32 + j < n_keys) mx = fmaxf(mx, __uint_as_float(v[j]));
                }
            }
            xmax[half * 128 + r] = mx;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            return fmaxf(xmax[r], xmax[128 + r]);
        };
        float mx = total_tiles > 0 ? row_max(0) : 0.f;
        for (int tt = 0; tt < total_tiles; ++tt) {
            const int il = tt / p.ntm, mt = tt % p.ntm;
            const int item = blockIdx.x + il * gridDim.x;
            const int b = item / p.H, h = item % p.H;
            const uint32_t ts = tmem_base + lane_addr + (uint32_t)((tt & 1) * S_STRIDE);
            const float mxs = mx * scale_log2;
            // P buffer must have been consumed by the previous tile's PV
            mbar_wait(p_empty, ((uint32_t)tt & 1u) ^ 1u);
            float sum = 0.f;
            for (int c = c_begin; c < c_end; ++c) {
                uint32_t v[32];
                tmem_ld32(ts + c * 32, v);
                tmem_ld_wait();
                float pr[32];
                if (c < full_chunks) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        pr[j] = ex2(__uint_as_float(v[j]) * scale_log2 - mxs);
                        sum += pr[j];
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        pr[j] = (c * 32 + j < n_keys) ? ex2(__uint_as_float(v[j]) * scale_log2 - mxs) : 0.f;
                        sum += pr[j];
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint4 pk;
                    pk.x = pack_bf16(pr[8 * i], pr[8 * i + 1]);
                    pk.y = pack_bf16(pr[8 * i + 2], pr[8 * i + 3]);
                    pk.z = pack_bf16(pr[8 * i + 4], pr[8 * i + 5]);
                    pk.w = pack_bf16(pr[8 * i + 6], pr[8 * i + 7]);
                    const int kb8 = c * 4 + i;                     // 8-key block index
                    *reinterpret_cast<uint4*>(sP_ptr + (kb8 >> 3) * P_ATOM_BYTES + r * 128 + (((kb8 & 7) ^ (r & 7)) << 4)) = pk;
                }
            }
            xsum[half * 128 + r] = sum;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to UMMA
            tc_fence_before();
            mbar_arrive(p_full);
            mbar_arrive(s_empty(tt & 1));
            // the next tile's S is already in TMEM: take its row max while the tensor core runs P V
            const float mx_this = mx;
            if (tt + 1 < total_tiles) mx = row_max(tt + 1);
            // epilogue: O / sum; this warp converts 32 of the 64 output dims
            mbar_wait(o_full, (uint32_t)tt & 1u);
            tc_fence_after();
            uint32_t o[32];
            tmem_ld32(tmem_base + lane_addr + O_COL + half * 32, o);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(o_empty);
            // o_full implies every softmax thread passed p_full, i.e. both halves of xsum are written
            sum = xsum[r] + xsum[128 + r];
            const int row = mt * BM + r;
            if (row < p.N) {
                const float inv = 1.0f / sum;
                __nv_bfloat16* dst = p.out + ((long long)b * p.N + row) * D + h * HD + half * 32;
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint4 pk;
                    pk.x = pack_bf16(__uint_as_float(o[8 * i]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
                    pk.y = pack_bf16(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
                    pk.z = pack_bf16(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
                    pk.w = pack_bf16(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
                    *reinterpret_cast<uint4*>(dst + 8 * i) = pk;
                    if (p.row_stats != nullptr) {
#pragma unroll
                        for (int z = 0; z < 8; ++z) {          // statistics for the folded inner_attn_ln (pre-rounding values)
                            const float f = __uint_as_float(o[8 * i + z]) * inv;
                            s1 += f;
                            s2 += f * f;
                        }
                    }
                }
                if (p.row_stats != nullptr)
                    *reinterpret_cast<float2*>(p.row_stats + (((long long)b * p.N + row) * (2 * p.H) + 2 * h + half) * 2) = make_float2(s1, s2);
                if (p.lse != nullptr && half == 0)
                    p.lse[((long long)b * p.H + h) * p.N + row] = mx_this * p.scale + logf(sum);
            }
            // xmax / xsum are rewritten by the next tile only after its bar.sync / p_full round
            asm volatile("bar.sync 2, 256;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace attn_tc

// Returns CS_ERR_UNSUPPORTED (without setting up anything) when the shape is outside this kernel's
// envelope; cs_attention_fwd then uses the generic mma.sync kernel.
int attention_fwd_tc(const void* qkv, int B, int N, int H, float scale, void* out, float* lse, float* row_stats,
                     cudaStream_t st) {
    using namespace attn_tc;
    if (N > MAX_NKP || N < 1) return CS_ERR_UNSUPPORTED;
    const int nkp = ceil_div(N, 16) * 16;
    const int D = H * HD;
    const long long rows = (long long)B * N;
    if (rows >= (1ll << 31)) return CS_ERR_UNSUPPORTED;
    CUtensorMap mq, mkv;
    int rc = make_map_bf16_2d(&mq, qkv, rows, 3 * D, 3 * D, HD, BM);
    if (rc) return rc;
    rc = make_map_bf16_2d(&mkv, qkv, rows, 3 * D, 3 * D, HD, nkp);
    if (rc) return rc;
    Params p;
    p.B = B; p.N = N; p.H = H; p.nkp = nkp; p.ntm = ceil_div(N, BM);
    p.scale = scale;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.out = (__nv_bfloat16*)out;
    p.lse = lse;
    p.row_stats = row_stats;
    const int n_atoms = (nkp + 63) / 64;
    const int smem = 2 * Q_BYTES + 4 * nkp * 128 + n_atoms * P_ATOM_BYTES + 256 + 2048 + 1024;   // + barriers, exchange, align
    static int configured = 0;
    if (configured < smem) {
        CS_CUDA(cudaFuncSetAttribute(attention_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    const int items = B * H;
    const int grid = items < num_sms() ? items : num_sms();
    attention_fwd_tc_kernel<<<grid, THREADS, smem, st>>>(mq, mkv, p);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

}  // namespace cs
