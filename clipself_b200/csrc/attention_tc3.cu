// tcgen05 softmax attention forward for head_dim 64 and N <= 208 tokens (EVA02-B/16: N = 197), third generation.
//   S = Q K^T  (tcgen05.mma, accumulator in TMEM)  ->  softmax in registers  ->  P (bf16) into swizzled shared memory  ->
//   O = P V (tcgen05.mma, V consumed MN-major straight from its [key][dim] layout) and the row sums = P x ones
//   (16 extra accumulator columns)  ->  O / rowsum -> bf16.
// Replaces xformers.memory_efficient_attention at eva_vit_model.py:206-217 for the teacher's crops and the student.
//
// Against attention_tc.cu (first generation, kept for A/B with CS_ATTN_V1=1): its ncu profile (profiles/r02_attn_*) shows
// the softmax warps issuing one instruction every ~3 cycles per scheduler — 2.5 warps per scheduler running dependent
// chains (row max, row sum) with the MUFU pipe at 29 % and the tensor pipe at 12 %: latency bound, not throughput bound.
// Measured here (tools/microbench): tcgen05.ld.x32 50 cycles per warp and scaling with warps, MUFU.EX2 16 lanes/clk/SM —
// neither limits.  So this kernel doubles the resident softmax warps (16: FOUR per TMEM lane quarter, each owning a
// quarter of the key columns and of the output dims), takes the row sum from the tensor core instead of a serial FADD
// chain, reduces the row max with four independent accumulators and keeps one exchange barrier per tile.
//
// Persistent CTA per SM, 576 threads: warp 0 TMA producer, warp 1 MMA issuer + TMEM allocator, warps 2-17 softmax +
// epilogue.  Work item = one (image, head); its K and V tiles are loaded once and reused by the ceil(N/128) query tiles.
// TMEM: S double buffered [0,208) [208,416), O [416,480), row sums [480,496).
#include "tc_common.cuh"

namespace cs {
namespace attn_tc3 {
using namespace cs::tc;
#define mbar_wait mbar_wait_spin

constexpr int HD = 64;
constexpr int BM = 128;
constexpr int THREADS = 576;           // TMA warp, MMA warp, 16 softmax/epilogue warps (four per TMEM lane quarter)
constexpr int SOFTMAX_THREADS = 512;
constexpr int PARTS = 4;               // column / output-dim split of a row among the 4 warps of a lane quarter
constexpr int Q_BYTES = BM * 128;
constexpr int P_ATOM_BYTES = BM * 128;      // [128 rows][64 keys] bf16, SWIZZLE_128B K-major
constexpr int MAX_NKP = 208;
constexpr int S_STRIDE = 208;
constexpr int O_COL = 416;                  // TMEM column of the O accumulator (S buffers at 0 and 208)
constexpr int SUM_COL = 480;                // 16 columns: P x ones = the row sums of the bf16-rounded probabilities
constexpr int TMEM_COLS = 512;
constexpr int ONES_BYTES = 16 * 128;        // B operand of the row-sum MMA: [16 rows][64 k] of 1.0 (every layout of ones is ones)

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
    int dbg;                        // timing experiments (CS_ATTN_DBG): 1 = skip the P V / row-sum MMAs, 2 = skip the exp pass
    __nv_bfloat16* out;
    float* lse;
    float* row_stats;               // optional [B*N, 4H, 2]: per (row, head, 16-dim quarter) sum and sum of squares of the f32 output
};

__global__ void __launch_bounds__(THREADS, 1)
attention_fwd_tc3_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
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
    const uint32_t sOnes = sP + n_atoms * P_ATOM_BYTES;
    const uint32_t bar = sOnes + ONES_BYTES;
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
    float* xch = reinterpret_cast<float*>(smem + (bar - base) + 8 * 18);   // [2 tile parities][4 parts][128 rows] partial row max
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
            mbar_init(s_empty(s), SOFTMAX_THREADS);
        }
        mbar_init(p_full, SOFTMAX_THREADS);
        mbar_init(p_empty, 1);
        mbar_init(o_full, 1);
        mbar_init(o_empty, SOFTMAX_THREADS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the block of ones for the row-sum MMA (bf16 1.0 = 0x3F80)
    for (int i = threadIdx.x; i < ONES_BYTES / 4; i += THREADS) reinterpret_cast<uint32_t*>(smem + (sOnes - base))[i] = 0x3F803F80u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
            const uint32_t idesc_sum = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            const uint64_t d_ones = smem_desc(sOnes, 0, 1024);
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
                for (int kk = 0; kk < ((p.dbg & 1) ? 0 : ksteps); ++kk) {
                    const uint64_t dp = smem_desc(sP + (kk >> 2) * P_ATOM_BYTES + (kk & 3) * 32, 0, 1024);
                    // V tile [keys][64 dims]: MN-major, one 64-wide atom, 16 keys per k-step = 2048 B
                    const uint64_t dv = smem_desc(sv + kk * 2048, (uint32_t)kv_bytes, 1024);
                    umma_bf16(tmem_base + O_COL, dp, dv, idesc_pv, kk > 0);
                    umma_bf16(tmem_base + SUM_COL, dp, d_ones, idesc_sum, kk > 0);
                }
                umma_commit(o_full);
                umma_commit(p_empty);
                if (tt % p.ntm == p.ntm - 1) umma_commit(kv_empty(il & 1));
            }
        }
    } else {
        // ------------------------------ softmax + epilogue ------------------------
        const int quarter = warp & 3;
        const int part = (warp - 2) >> 2;                          // 0..3: which 2 of the (up to 7) 32-key chunks / which 16 output dims
        const int r = quarter * 32 + lane;                         // row of the tile owned by this thread
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const int nchunks = (p.nkp + 31) / 32;
        const int c_begin = min(2 * part, nchunks), c_end = min(2 * part + 2, nchunks);
        const int n_keys = p.N, full_chunks = p.N / 32;
        const float scale_log2 = p.scale_log2;
        // partial row max of one tile (this thread's chunks), published for the 3 partner warps of the lane quarter
        auto row_max = [&](int tt) -> float {
            mbar_wait(s_full(tt & 1), (uint32_t)(tt >> 1) & 1u);
            tc_fence_after();
            const uint32_t ts = tmem_base + lane_addr + (uint32_t)((tt & 1) * S_STRIDE);
            float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
            for (int c = c_begin; c < ((p.dbg & 8) ? c_begin : c_end); ++c) {
                uint32_t v[32];
                tmem_ld32(ts + c * 32, v);
                tmem_ld_wait();
                if (c < full_chunks) {                    // all 32 keys valid: no per-element predicate, 4 independent chains
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        m0 = fmaxf(m0, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
                        m1 = fmaxf(m1, fmaxf(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
                        m2 = fmaxf(m2, fmaxf(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5])));
                        m3 = fmaxf(m3, fmaxf(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7])));
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c * 32 + j < n_keys) m0 = fmaxf(m0, __uint_as_float(v[j]));
                }
            }
            float* xm = xch + (tt & 1) * (PARTS * BM);
            xm[part * BM + r] = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            asm volatile("bar.sync 1, 512;" ::: "memory");
            return fmaxf(fmaxf(xm[r], xm[BM + r]), fmaxf(xm[2 * BM + r], xm[3 * BM + r]));
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
            for (int c = c_begin; c < ((p.dbg & 2) ? c_begin : c_end); ++c) {
                uint32_t v[32];
                tmem_ld32(ts + c * 32, v);
                tmem_ld_wait();
                float pr[32];
                if (c < full_chunks) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) pr[j] = ex2(fmaf(__uint_as_float(v[j]), scale_log2, -mxs));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) pr[j] = (c * 32 + j < n_keys) ? ex2(fmaf(__uint_as_float(v[j]), scale_log2, -mxs)) : 0.f;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (c * 32 + i * 8 < p.nkp) {
                        uint4 pk;
                        pk.x = pack_bf16(pr[8 * i], pr[8 * i + 1]);
                        pk.y = pack_bf16(pr[8 * i + 2], pr[8 * i + 3]);
                        pk.z = pack_bf16(pr[8 * i + 4], pr[8 * i + 5]);
                        pk.w = pack_bf16(pr[8 * i + 6], pr[8 * i + 7]);
                        const int kb8 = c * 4 + i;                     // 8-key block index
                        *reinterpret_cast<uint4*>(sP_ptr + (kb8 >> 3) * P_ATOM_BYTES + r * 128 + (((kb8 & 7) ^ (r & 7)) << 4)) = pk;
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to UMMA
            tc_fence_before();
            mbar_arrive(p_full);
            mbar_arrive(s_empty(tt & 1));
            // the next tile's S is already in TMEM: take its row max while the tensor core runs P V
            const float mx_this = mx;
            if (tt + 1 < total_tiles) mx = row_max(tt + 1);
            // epilogue: O / sum; this warp converts 16 of the 64 output dims
            mbar_wait(o_full, (uint32_t)tt & 1u);
            tc_fence_after();
            uint32_t o[16];
            tmem_ld16(tmem_base + lane_addr + O_COL + part * 16, o);
            uint32_t sum_bits;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(sum_bits) : "r"(tmem_base + lane_addr + SUM_COL) : "memory");
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(o_empty);
            const float sum = __uint_as_float(sum_bits);
            const int row = mt * BM + r;
            if (row < p.N && !(p.dbg & 4)) {
                const float inv = 1.0f / sum;
                __nv_bfloat16* dst = p.out + ((long long)b * p.N + row) * D + h * HD + part * 16;
                float f[16];
                float s1 = 0.f, s2 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
                for (int z = 0; z < 16; z += 2) {
                    f[z] = __uint_as_float(o[z]) * inv;
                    f[z + 1] = __uint_as_float(o[z + 1]) * inv;
                    s1 += f[z];
                    s2 = fmaf(f[z], f[z], s2);
                    t1 += f[z + 1];
                    t2 = fmaf(f[z + 1], f[z + 1], t2);
                }
                uint4 pk, qk;
                pk.x = pack_bf16(f[0], f[1]); pk.y = pack_bf16(f[2], f[3]); pk.z = pack_bf16(f[4], f[5]); pk.w = pack_bf16(f[6], f[7]);
                qk.x = pack_bf16(f[8], f[9]); qk.y = pack_bf16(f[10], f[11]); qk.z = pack_bf16(f[12], f[13]); qk.w = pack_bf16(f[14], f[15]);
                *reinterpret_cast<uint4*>(dst) = pk;
                *reinterpret_cast<uint4*>(dst + 8) = qk;
                if (p.row_stats != nullptr)        // statistics for the folded inner_attn_ln (f32 values before the bf16 rounding)
                    *reinterpret_cast<float2*>(p.row_stats + (((long long)b * p.N + row) * (4 * p.H) + 4 * h + part) * 2) = make_float2(s1 + t1, s2 + t2);
                if (p.lse != nullptr && part == 0)
                    p.lse[((long long)b * p.H + h) * p.N + row] = mx_this * p.scale + logf(sum);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace attn_tc3

// Returns CS_ERR_UNSUPPORTED (without setting up anything) when the shape is outside this kernel's
// envelope; cs_attention_fwd then uses the generic mma.sync kernel.
int attention_fwd_tc3(const void* qkv, int B, int N, int H, float scale, void* out, float* lse, float* row_stats,
                     cudaStream_t st) {
    using namespace attn_tc3;
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
    {
        const char* e = getenv("CS_ATTN_DBG");
        p.dbg = e != nullptr ? atoi(e) : 0;
    }
    p.out = (__nv_bfloat16*)out;
    p.lse = lse;
    p.row_stats = row_stats;
    const int n_atoms = (nkp + 63) / 64;
    const int smem = 2 * Q_BYTES + 4 * nkp * 128 + n_atoms * P_ATOM_BYTES + ONES_BYTES + 256 + 2 * PARTS * BM * 4 + 1024;   // + barriers, exchange, align
    static int configured = 0;
    if (configured < smem) {
        CS_CUDA(cudaFuncSetAttribute(attention_fwd_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    const int items = B * H;
    const int grid = items < num_sms() ? items : num_sms();
    attention_fwd_tc3_kernel<<<grid, THREADS, smem, st>>>(mq, mkv, p);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

}  // namespace cs
