// tcgen05 flash-attention forward for head_dim 64 and ANY sequence length: the long-sequence counterpart of
// attention_tc.cu (N <= 224) for ViT-L/14-336 (577 tokens) and the 1024 px student (4097 tokens); measured on a
// B200 against the mma.sync kernel of attention.cu: 1.41x at N = 577, 1.58x at N = 4097 (profiles/r02_*).  Replaces xformers.memory_efficient_attention
// (eva_vit_model.py:206-217).
//
// Persistent CTA per SM, 640 threads: warps 0-7 = softmax group A, warps 8-15 = softmax group B, warp 16 TMA producer,
// warp 17 MMA issuer + TMEM allocator (setmaxnreg moves the registers of the last warpgroup to the softmax groups).  A work
// item is one (image, head, PAIR of 128-row query tiles): group A owns the first tile, group B the second.  Inside a group
// TWO warps share each TMEM lane quarter (the lesson of attention_tc4.cu: one softmax warp per scheduler issues every ~4
// cycles): warp half h owns keys [64h, 64h+64) of every 128-key block and output dims [32h, 32h+32); the block's row
// maximum is exchanged through shared memory (one 256-thread named barrier per block), the row sums stay partial per warp
// until the end.  The KV sequence streams through a 3-stage TMA ring in blocks of 128 keys;
// per block and group:   S = Q K^T (tcgen05.mma -> TMEM)  ->  softmax group: running max, P = exp2(.) as
// bf16 into SWIZZLE_128B shared memory  ->  O_blk = P V (tcgen05.mma, V MN-major, fresh accumulator)  ->
// softmax group: O = O * 2^(m_old - m_new) + O_blk in registers (no TMEM read-modify-write).
// The two groups ping-pong on the tensor core: Q K^T / P V of one group run under the softmax of the other.
#include "tc_common.cuh"

namespace cs {
namespace attn_tcl {
using namespace cs::tc;

constexpr int HD = 64;
constexpr int BM = 128;                     // query rows per tile (= TMEM lanes)
constexpr int BK = 128;                     // keys per block
constexpr int THREADS = 640;
constexpr int TMA_WARP = 16, MMA_WARP = 17;
constexpr int GROUP_WARPS = 8;
constexpr int KV_STAGES = 3;
constexpr int TILE_BYTES = BM * 128;        // [128 rows][64 bf16], SWIZZLE_128B: Q, K and V tiles, one P atom
constexpr int P_BYTES = 2 * TILE_BYTES;     // P [128 rows][128 keys] = two 64-key atoms
constexpr int S_COL = 0;                    // TMEM columns: S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384)
constexpr int O_COL = 256;
constexpr int X_COL = 384;                   // 12 spare columns: the two warps of a lane quarter exchange their partial row max / row
                                            // sum here (shared memory is full): [parity][group][half] max at 384.., [group][half] sum at 392..
constexpr int TMEM_COLS = 512;

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, float v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(__float_as_uint(v)) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    return __uint_as_float(v);
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct Params {
    int B, N, H;
    int ntm;        // 128-row query tiles per head
    int npairs;     // ceil(ntm / 2)
    int nkb;        // 128-key blocks per head
    float scale_log2, scale;
    __nv_bfloat16* out;
    float* lse;
    float* row_stats;   // optional [B*N, 4H, 2]: per (row, head, 16-dim quarter) sum and sum of squares of the f32 output
};

__global__ void __launch_bounds__(THREADS, 1)
attention_fwd_tc_long_kernel(const __grid_constant__ CUtensorMap map, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_addr);

    const uint32_t sQ = base;                                       // [2 stages][2 groups]
    const uint32_t sKV = sQ + 4 * TILE_BYTES;                       // [KV_STAGES][K | V]
    const uint32_t sP = sKV + KV_STAGES * 2 * TILE_BYTES;           // [2 groups][2 atoms]
    const uint32_t bar = sP + 2 * P_BYTES;
    auto q_full = [&](int s, int g) { return bar + 8u * (2 * s + g); };          // item il uses stage il & 1
    auto q_empty = [&](int s, int g) { return bar + 8u * (4 + 2 * s + g); };
    auto kv_full = [&](int s) { return bar + 8u * (8 + s); };
    auto kv_empty = [&](int s) { return bar + 8u * (8 + KV_STAGES + s); };
    constexpr int B0 = 8 + 2 * KV_STAGES;
    auto s_full = [&](int g) { return bar + 8u * (B0 + g); };
    auto s_empty = [&](int g) { return bar + 8u * (B0 + 2 + g); };
    auto p_full = [&](int g) { return bar + 8u * (B0 + 4 + g); };
    auto p_empty = [&](int g) { return bar + 8u * (B0 + 6 + g); };
    auto o_full = [&](int g) { return bar + 8u * (B0 + 8 + g); };
    auto o_empty = [&](int g) { return bar + 8u * (B0 + 10 + g); };
    const uint32_t tmem_slot = bar + 8u * (B0 + 12);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));


    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = p.H * HD;
    const int n_items = p.B * p.H * p.npairs;
    const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map);
        for (int g = 0; g < 2; ++g) {
            for (int s = 0; s < 2; ++s) {
                mbar_init(q_full(s, g), 1);
                mbar_init(q_empty(s, g), 1);
            }
            mbar_init(s_full(g), 1);
            mbar_init(s_empty(g), GROUP_WARPS);        // one elected arrival per warp of the group
            mbar_init(p_full(g), GROUP_WARPS);
            mbar_init(p_empty(g), 1);
            mbar_init(o_full(g), 1);
            mbar_init(o_empty(g), GROUP_WARPS);
        }
        for (int s = 0; s < KV_STAGES; ++s) {
            mbar_init(kv_full(s), 1);
            mbar_init(kv_empty(s), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp >= 2 * GROUP_WARPS) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
      if (warp == TMA_WARP) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            int c = 0;                                              // KV blocks loaded so far (all items)
            for (int il = 0; il < my_items; ++il) {
                const int item = blockIdx.x + il * gridDim.x;
                const int pair = item % p.npairs, bh = item / p.npairs;
                const int b = bh / p.H, h = bh % p.H;
                const int qs = il & 1;
                for (int g = 0; g < 2; ++g) {
                    mbar_wait(q_empty(qs, g), (uint32_t)((il >> 1) & 1) ^ 1u);
                    mbar_arrive_expect_tx(q_full(qs, g), TILE_BYTES);
                    tma_load_2d(sQ + (2 * qs + g) * TILE_BYTES, &map, q_full(qs, g), h * HD, b * p.N + (2 * pair + g) * BM);
                }
                for (int j = 0; j < p.nkb; ++j, ++c) {
                    const int st = c % KV_STAGES;
                    mbar_wait(kv_empty(st), (uint32_t)((c / KV_STAGES) & 1) ^ 1u);
                    mbar_arrive_expect_tx(kv_full(st), 2u * TILE_BYTES);
                    tma_load_2d(sKV + st * 2 * TILE_BYTES, &map, kv_full(st), D + h * HD, b * p.N + j * BK);
                    tma_load_2d(sKV + st * 2 * TILE_BYTES + TILE_BYTES, &map, kv_full(st), 2 * D + h * HD, b * p.N + j * BK);
                }
            }
        }
      } else if (warp == MMA_WARP) {
        // ------------------------------ MMA issuer --------------------------------
        if (lane == 0) {
            // S: M=128, N=128, A/B K-major.  PV: M=128, N=64, A K-major (P), B MN-major (V) -> bit 16
            const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BK >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(HD >> 3) << 17) |
                                      ((uint32_t)(BM >> 4) << 24);
            auto issue_s = [&](int g, int c, int qs) {              // S_g of global block c (its K tile is in stage c % 3)
                mbar_wait(s_empty(g), (uint32_t)(c & 1) ^ 1u);      // the group has read S_g of block c-1
                tc_fence_after();
                const uint64_t dq = smem_desc(sQ + (2 * qs + g) * TILE_BYTES, 0, 1024);
                const uint64_t dk = smem_desc(sKV + (c % KV_STAGES) * 2 * TILE_BYTES, 0, 1024);
                const uint32_t d_tmem = tmem_base + (uint32_t)(S_COL + g * BK);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_bf16(d_tmem, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, k > 0);
                umma_commit(s_full(g));
            };
            int c0 = 0;
            for (int il = 0; il < my_items; ++il) {
                const int qs = il & 1;
                mbar_wait(q_full(qs, 0), (uint32_t)((il >> 1) & 1));
                mbar_wait(q_full(qs, 1), (uint32_t)((il >> 1) & 1));
                mbar_wait(kv_full(c0 % KV_STAGES), (uint32_t)((c0 / KV_STAGES) & 1));
                issue_s(0, c0, qs);
                issue_s(1, c0, qs);
                for (int j = 0; j < p.nkb; ++j) {
                    const int c = c0 + j;
                    const uint32_t sv = sKV + (c % KV_STAGES) * 2 * TILE_BYTES + TILE_BYTES;
                    for (int g = 0; g < 2; ++g) {
                        mbar_wait(p_full(g), (uint32_t)(c & 1));            // P_g of block c is in shared memory
                        mbar_wait(o_empty(g), (uint32_t)(c & 1) ^ 1u);      // O_g of block c-1 has been read
                        tc_fence_after();
#pragma unroll
                        for (int kk = 0; kk < BK / 16; ++kk) {
                            const uint64_t dp = smem_desc(sP + g * P_BYTES + (kk >> 2) * TILE_BYTES + (kk & 3) * 32, 0, 1024);
                            // V tile [128 keys][64 dims]: MN-major, one 64-wide atom, 16 keys per k-step = 2048 B
                            const uint64_t dv = smem_desc(sv + kk * 2048, (uint32_t)TILE_BYTES, 1024);
                            umma_bf16(tmem_base + (uint32_t)(O_COL + g * HD), dp, dv, idesc_pv, kk > 0);
                        }
                        umma_commit(o_full(g));
                        umma_commit(p_empty(g));
                        if (g == 1) umma_commit(kv_empty(c % KV_STAGES));  // both groups' MMAs on this stage are done
                        if (j + 1 < p.nkb) {
                            if (g == 0) mbar_wait(kv_full((c + 1) % KV_STAGES), (uint32_t)(((c + 1) / KV_STAGES) & 1));
                            issue_s(g, c + 1, qs);
                        }
                    }
                }
                umma_commit(q_empty(qs, 0));                        // every S MMA of this item has been issued
                umma_commit(q_empty(qs, 1));
                c0 += p.nkb;
            }
        }
      }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // ------------------------------ softmax groups ----------------------------
        const int g = warp >> 3;                                    // 0: group A (warps 0-7), 1: group B (warps 8-15)
        const int half = (warp >> 2) & 1;                           // keys [64 half, 64 half + 64) of a block, dims [32 half, 32 half + 32)
        const int quarter = warp & 3;                               // TMEM lane quarter this warp may access
        const int r = quarter * 32 + lane;                          // row of the tile owned by this thread (shared with its partner warp)
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const uint32_t ts = tmem_base + lane_addr + (uint32_t)(S_COL + g * BK + half * 64);
        const uint32_t to = tmem_base + lane_addr + (uint32_t)(O_COL + g * HD + half * 32);
        uint8_t* sP_ptr = smem + (sP - base) + g * P_BYTES + half * TILE_BYTES + r * 128;    // this warp's 64-key atom, this thread's row
        const float sl2 = p.scale_log2;
        const uint32_t tx = tmem_base + lane_addr + (uint32_t)X_COL;     // exchange columns of this lane quarter
        int c0 = 0;
        for (int il = 0; il < my_items; ++il) {
            const int item = blockIdx.x + il * gridDim.x;
            const int pair = item % p.npairs, bh = item / p.npairs;
            const int b = bh / p.H, h = bh % p.H;
            const int row = (2 * pair + g) * BM + r;
            float m_run = -INFINITY, m_acc = -INFINITY, l_run = 0.f;   // l_run: partial sum over this warp's keys
            float o_run[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) o_run[i] = 0.f;
            // O_blk of global block c (relative to max m_blk) folded into the running output (this warp's 32 dims)
            auto fold_o = [&](int c, float m_blk) {
                mbar_wait(o_full(g), (uint32_t)(c & 1));
                tc_fence_after();
                const float beta = ex2((m_acc - m_blk) * sl2);      // m_acc = -inf before the first block: beta = 0
                uint32_t o[32];
                tmem_ld32(to, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o_run[i] = fmaf(o_run[i], beta, __uint_as_float(o[i]));
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(o_empty(g));
                m_acc = m_blk;
            };
            for (int j = 0; j < p.nkb; ++j) {
                const int c = c0 + j;
                const int nvalid = min(BK, p.N - j * BK) - half * 64;   // keys of this warp's half that belong to the image (may be <= 0)
                mbar_wait(s_full(g), (uint32_t)(c & 1));
                tc_fence_after();
                // ---- partial block max over this warp's 64 keys, exchanged with the partner warp
                float m_part = -INFINITY;
#pragma unroll 1
                for (int cc = 0; cc < 2; ++cc) {
                    uint32_t v[32];
                    tmem_ld32(ts + cc * 32, v);
                    tmem_ld_wait();
                    if ((cc + 1) * 32 <= nvalid) {
                        float m0 = m_part, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            m0 = fmaxf(m0, __uint_as_float(v[i]));
                            m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
                            m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
                            m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
                        }
                        m_part = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (cc * 32 + i < nvalid) m_part = fmaxf(m_part, __uint_as_float(v[i]));
                    }
                }
                const uint32_t xm = tx + (uint32_t)((((c & 1) * 2 + g) * 2));         // double buffered by block parity
                tmem_st1(xm + half, m_part);
                tc_fence_before();
                asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");
                tc_fence_after();
                const float m_new = fmaxf(m_run, fmaxf(m_part, tmem_ld1(xm + (half ^ 1))));
                const float mxs = m_new * sl2;
                mbar_wait(p_empty(g), (uint32_t)(c & 1) ^ 1u);      // P V of block c-1 has consumed the P buffer
                float s0 = 0.f, s1 = 0.f;
#pragma unroll 1
                for (int cc = 0; cc < 2; ++cc) {
                    uint32_t v[32];
                    tmem_ld32(ts + cc * 32, v);
                    tmem_ld_wait();
                    if ((cc + 1) * 32 > nvalid) {                   // partial / empty chunk: keys beyond the image score -inf
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (cc * 32 + i >= nvalid) v[i] = 0xFF800000u;
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float pr[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) pr[q] = ex2(fmaf(__uint_as_float(v[8 * i + q]), sl2, -mxs));
                        s0 += (pr[0] + pr[1]) + (pr[2] + pr[3]);
                        s1 += (pr[4] + pr[5]) + (pr[6] + pr[7]);
                        uint4 pk;
                        pk.x = pack_bf16(pr[0], pr[1]);
                        pk.y = pack_bf16(pr[2], pr[3]);
                        pk.z = pack_bf16(pr[4], pr[5]);
                        pk.w = pack_bf16(pr[6], pr[7]);
                        const int kb8 = cc * 4 + i;                 // 8-key block index inside this warp's 64-key atom
                        *reinterpret_cast<uint4*>(sP_ptr + (((kb8 & 7) ^ (r & 7)) << 4)) = pk;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to UMMA
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(p_full(g));
                    mbar_arrive(s_empty(g));
                }
                l_run = fmaf(l_run, ex2((m_run - m_new) * sl2), s0 + s1);       // m_run = -inf on the first block: factor 0
                if (j > 0) fold_o(c - 1, m_run);                    // O of the previous block, relative to its max m_run
                m_run = m_new;
            }
            fold_o(c0 + p.nkb - 1, m_run);
            // total row sum = this warp's partial + the partner's (same running maximum in both)
            tmem_st1(tx + 8 + g * 2 + half, l_run);
            tc_fence_before();
            asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");
            tc_fence_after();
            const float l_tot = l_run + tmem_ld1(tx + 8 + g * 2 + (half ^ 1));
            if (2 * pair + g < p.ntm && row < p.N) {
                const float inv = 1.0f / l_tot;
                __nv_bfloat16* dst = p.out + ((long long)b * p.N + row) * D + h * HD + half * 32;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint4 pk;
                    pk.x = pack_bf16(o_run[8 * i] * inv, o_run[8 * i + 1] * inv);
                    pk.y = pack_bf16(o_run[8 * i + 2] * inv, o_run[8 * i + 3] * inv);
                    pk.z = pack_bf16(o_run[8 * i + 4] * inv, o_run[8 * i + 5] * inv);
                    pk.w = pack_bf16(o_run[8 * i + 6] * inv, o_run[8 * i + 7] * inv);
                    *reinterpret_cast<uint4*>(dst + 8 * i) = pk;
                }
                if (p.row_stats != nullptr) {                       // statistics for the folded inner_attn_ln: parts 2 half, 2 half + 1 of this head
                    float st[4];
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        float a1 = 0.f, a2 = 0.f;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float f = o_run[hh * 16 + i] * inv;
                            a1 += f;
                            a2 = fmaf(f, f, a2);
                        }
                        st[2 * hh] = a1;
                        st[2 * hh + 1] = a2;
                    }
                    *reinterpret_cast<float4*>(p.row_stats + (((long long)b * p.N + row) * (4 * p.H) + 4 * h + 2 * half) * 2) =
                        make_float4(st[0], st[1], st[2], st[3]);
                }
                if (p.lse != nullptr && half == 0) p.lse[((long long)b * p.H + h) * p.N + row] = m_run * p.scale + logf(l_tot);
            }
            // (the sum columns are rewritten only at the end of the next item, i.e. after all its block barriers)
            c0 += p.nkb;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace attn_tcl

// Long-sequence tcgen05 forward.  Returns CS_ERR_UNSUPPORTED when the shape is outside the kernel's envelope
// (cs_attention_fwd then falls through to the mma.sync kernel).
int attention_fwd_tc_long(const void* qkv, int B, int N, int H, float scale, void* out, float* lse, float* row_stats,
                          cudaStream_t st) {
    using namespace attn_tcl;
    if (N < 1) return CS_ERR_UNSUPPORTED;
    const int D = H * HD;
    const long long rows = (long long)B * N;
    if (rows >= (1ll << 31)) return CS_ERR_UNSUPPORTED;
    CUtensorMap map;
    const int rc = make_map_bf16_2d(&map, qkv, rows, 3 * D, 3 * D, HD, BM);     // Q, K and V tiles: [128 rows][64 dims]
    if (rc) return rc;
    Params p;
    p.B = B; p.N = N; p.H = H;
    p.ntm = ceil_div(N, BM);
    p.npairs = ceil_div(p.ntm, 2);
    p.nkb = ceil_div(N, BK);
    p.scale = scale;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.out = (__nv_bfloat16*)out;
    p.lse = lse;
    p.row_stats = row_stats;
    const int smem = 4 * TILE_BYTES + KV_STAGES * 2 * TILE_BYTES + 2 * P_BYTES + 512 + 1024;   // + barriers, alignment
    static bool configured = false;
    if (!configured) {
        CS_CUDA(cudaFuncSetAttribute(attention_fwd_tc_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    const long long items = (long long)B * H * p.npairs;
    if (items >= (1ll << 31)) return CS_ERR_UNSUPPORTED;
    const int grid = items < num_sms() ? (int)items : num_sms();
    attention_fwd_tc_long_kernel<<<grid, THREADS, smem, st>>>(map, p);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

}  // namespace cs
