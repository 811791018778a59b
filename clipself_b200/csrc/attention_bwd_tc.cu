// tcgen05 attention backward for head_dim 64 and any sequence length (autograd of eva_vit_model.py:206-217): the default
// behind cs_attention_bwd for every shape; the mma.sync kernels of round 1 (attention.cu) only run with CS_ATTN_LEGACY=1.
// Measured on a B200 (profiles/r02_attention_experiments.txt): 429 us at B=64 N=197 (slower than the mma.sync pair, 307 us),
// 451 us at B=16 N=577, 1052 us at B=2 N=4097 (1.6x faster).
//
// Like the mma.sync version it is two passes that recompute P from the saved log-sum-exp, so neither needs
// a running softmax or atomics; both are instances of ONE streamed-operand template:
//
//   resident tile (128 rows = TMEM lanes)   streamed blocks (64 rows each)     per block
//   dQ  pass: R1 = Q tile, R2 = dO tile     X1 = K block, X2 = V block         T1 = S = R1 X1^T, T2 = dP = R2 X2^T
//             dS = P o (dP - delta_row) * scale  (bf16 -> smem)                dQ  += dS X1        (X1 MN-major)
//   dKV pass: R1 = K tile, R2 = V tile      X1 = Q block, X2 = dO block        T1 = S^T,          T2 = dP^T
//             P^T, dS^T (bf16 -> smem)                                         dV += P^T X2,  dK += dS^T X1
//
// with P = exp2(S * scale * log2e - lse * log2e).  Persistent CTA per SM, 320 threads: warp 0 TMA producer (and,
// in the dKV pass, the per-query lse / delta vectors of each block), warp 1 MMA issuer + TMEM allocator,
// warps 2-9 elementwise: two warps per TMEM lane quarter split the 64 streamed columns — there is no
// reduction in the backward, so they never exchange anything.  T1 / T2 are double buffered in TMEM so the
// products of block j+1 run under the elementwise work of block j; the gradient tiles accumulate in TMEM
// over the whole streamed loop and get the inverse RoPE rotation in the epilogue.
#include "tc_common.cuh"

namespace cs {
namespace attn_bwd_tc {
using namespace cs::tc;

constexpr int HD = 64;
constexpr int BM = 128;                     // resident rows (TMEM lanes)
constexpr int BS = 64;                      // streamed rows per block
constexpr int THREADS = 320;
constexpr int STAGES = 4;                   // streamed ring
constexpr int RES_BYTES = BM * 128;         // [128 rows][64 bf16], SWIZZLE_128B
constexpr int STR_BYTES = BS * 128;         // [64 rows][64 bf16]
constexpr int E_BYTES = BM * 128;           // [128 rows][64 streamed] bf16 = one SWIZZLE_128B atom
constexpr int T1_COL = 0, T2_COL = 128;     // two 64-column buffers each
constexpr int ACC1_COL = 256, ACC2_COL = 320;
constexpr int TMEM_COLS = 512;
constexpr float LOG2E = 1.4426950408889634f;

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
    int B, N, H;
    int ntiles;         // 128-row resident tiles per head
    int nblk;           // 64-row streamed blocks per head
    int r1_col, r2_col, x1_col, x2_col;     // column offsets (elements) of head 0 inside the mapped tensors
    float scale, scale_log2;
    const float* lse;   // [B, H, N]
    const float* delta; // [B, H, N]
    const float* rope_cos;  // [N-1, 64] or null
    const float* rope_sin;
    __nv_bfloat16* dqkv;    // [B*N, 3D]
};

// inverse rotation of one adjacent (even, odd) pair:  forward was y0 = x0 c0 - x1 s0, y1 = x1 c1 + x0 s1
__device__ __forceinline__ void unrope_pair(float& d0, float& d1, const float* __restrict__ cs_row,
                                            const float* __restrict__ sn_row, int d) {
    const float c0 = cs_row[d], c1 = cs_row[d + 1], s0 = sn_row[d], s1 = sn_row[d + 1];
    const float x0 = d0 * c0 + d1 * s1;
    const float x1 = d1 * c1 - d0 * s0;
    d0 = x0;
    d1 = x1;
}

__device__ __forceinline__ void store_row32(__nv_bfloat16* dst, const float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 pk;
        pk.x = pack_bf16(v[8 * i], v[8 * i + 1]);
        pk.y = pack_bf16(v[8 * i + 2], v[8 * i + 3]);
        pk.z = pack_bf16(v[8 * i + 4], v[8 * i + 5]);
        pk.w = pack_bf16(v[8 * i + 6], v[8 * i + 7]);
        *reinterpret_cast<uint4*>(dst + 8 * i) = pk;
    }
}

template <bool kDKV>
__global__ void __launch_bounds__(THREADS, 1)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_r1, const __grid_constant__ CUtensorMap map_r2,
                        const __grid_constant__ CUtensorMap map_x1, const __grid_constant__ CUtensorMap map_x2,
                        const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_addr);

    constexpr int NE = kDKV ? 2 : 1;                                // elementwise result tiles per buffer
    const uint32_t sR = base;                                       // [2 item stages][R1 | R2]
    const uint32_t sX = sR + 4 * RES_BYTES;                         // [STAGES][X1 | X2]
    const uint32_t sE = sX + STAGES * 2 * STR_BYTES;                // [2 buffers][NE]
    const uint32_t sVec = sE + 2 * NE * E_BYTES;                    // dKV: [STAGES][lse2 | delta][64] f32
    const uint32_t bar = sVec + STAGES * 2 * BS * 4;
    auto r_full = [&](int s) { return bar + 8u * s; };
    auto r_empty = [&](int s) { return bar + 8u * (2 + s); };
    auto x_full = [&](int s) { return bar + 8u * (4 + s); };
    auto x_empty = [&](int s) { return bar + 8u * (4 + STAGES + s); };
    constexpr int B0 = 4 + 2 * STAGES;
    auto t_full = [&](int b) { return bar + 8u * (B0 + b); };
    auto t_empty = [&](int b) { return bar + 8u * (B0 + 2 + b); };
    auto e_full = [&](int b) { return bar + 8u * (B0 + 4 + b); };
    auto e_empty = [&](int b) { return bar + 8u * (B0 + 6 + b); };
    const uint32_t acc_full = bar + 8u * (B0 + 8), acc_empty = bar + 8u * (B0 + 9);
    const uint32_t tmem_slot = bar + 8u * (B0 + 10);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));
    float* vec = reinterpret_cast<float*>(smem + (sVec - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = p.H * HD;
    const int n_items = p.B * p.H * p.ntiles;
    const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_r1);
        tma_prefetch_desc(&map_r2);
        tma_prefetch_desc(&map_x1);
        tma_prefetch_desc(&map_x2);
        for (int s = 0; s < 2; ++s) {
            mbar_init(r_full(s), 1);
            mbar_init(r_empty(s), 1);
            mbar_init(t_full(s), 1);
            mbar_init(t_empty(s), 8);       // one elected arrival per softmax warp
            mbar_init(e_full(s), 8);
            mbar_init(e_empty(s), 1);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(x_full(s), kDKV ? 33 : 1);                    // TMA transaction (+ the 32 lanes that stage lse / delta)
            mbar_init(x_empty(s), 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0 || kDKV) {
            int c = 0;
            for (int il = 0; il < my_items; ++il) {
                const int item = blockIdx.x + il * gridDim.x;
                const int tile = item % p.ntiles, bh = item / p.ntiles;
                const int b = bh / p.H, h = bh % p.H;
                const int rs = il & 1;
                if (lane == 0) {
                    mbar_wait(r_empty(rs), (uint32_t)((il >> 1) & 1) ^ 1u);
                    mbar_arrive_expect_tx(r_full(rs), 2u * RES_BYTES);
                    tma_load_2d(sR + (2 * rs) * RES_BYTES, &map_r1, r_full(rs), p.r1_col + h * HD, b * p.N + tile * BM);
                    tma_load_2d(sR + (2 * rs + 1) * RES_BYTES, &map_r2, r_full(rs), p.r2_col + h * HD, b * p.N + tile * BM);
                }
                for (int j = 0; j < p.nblk; ++j, ++c) {
                    const int st = c % STAGES;
                    mbar_wait(x_empty(st), (uint32_t)((c / STAGES) & 1) ^ 1u);
                    if (lane == 0) {
                        mbar_arrive_expect_tx(x_full(st), 2u * STR_BYTES);
                        tma_load_2d(sX + st * 2 * STR_BYTES, &map_x1, x_full(st), p.x1_col + h * HD, b * p.N + j * BS);
                        tma_load_2d(sX + st * 2 * STR_BYTES + STR_BYTES, &map_x2, x_full(st), p.x2_col + h * HD, b * p.N + j * BS);
                    }
                    if (kDKV) {                                     // per-query scalars of the streamed block
                        const float* lrow = p.lse + ((long long)b * p.H + h) * p.N;
                        const float* drow = p.delta + ((long long)b * p.H + h) * p.N;
#pragma unroll
                        for (int z = 0; z < 2; ++z) {
                            const int col = z * 32 + lane, q = j * BS + col;
                            vec[(st * 2 + 0) * BS + col] = q < p.N ? lrow[q] * LOG2E : 0.f;
                            vec[(st * 2 + 1) * BS + col] = q < p.N ? drow[q] : 0.f;
                        }
                        mbar_arrive(x_full(st));
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer --------------------------------
        if (lane == 0) {
            // products: M=128, N=64, both operands K-major.  accumulations: B operand MN-major (bit 16)
            const uint32_t idesc_t = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BS >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            const uint32_t idesc_acc = idesc_t | (1u << 16);        // N = 64 dims as well
            auto issue_t = [&](int c, int rs) {
                const int buf = c & 1, st = c % STAGES;
                mbar_wait(x_full(st), (uint32_t)((c / STAGES) & 1));
                mbar_wait(t_empty(buf), (uint32_t)((c >> 1) & 1) ^ 1u);
                tc_fence_after();
                const uint64_t dr1 = smem_desc(sR + (2 * rs) * RES_BYTES, 0, 1024);
                const uint64_t dr2 = smem_desc(sR + (2 * rs + 1) * RES_BYTES, 0, 1024);
                const uint64_t dx1 = smem_desc(sX + st * 2 * STR_BYTES, 0, 1024);
                const uint64_t dx2 = smem_desc(sX + st * 2 * STR_BYTES + STR_BYTES, 0, 1024);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    umma_bf16(tmem_base + (uint32_t)(T1_COL + buf * BS), dr1 + (uint64_t)(2 * k), dx1 + (uint64_t)(2 * k), idesc_t, k > 0);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    umma_bf16(tmem_base + (uint32_t)(T2_COL + buf * BS), dr2 + (uint64_t)(2 * k), dx2 + (uint64_t)(2 * k), idesc_t, k > 0);
                umma_commit(t_full(buf));
            };
            int c0 = 0;
            for (int il = 0; il < my_items; ++il) {
                const int rs = il & 1;
                mbar_wait(r_full(rs), (uint32_t)((il >> 1) & 1));
                issue_t(c0, rs);
                for (int j = 0; j < p.nblk; ++j) {
                    const int c = c0 + j;
                    const int buf = c & 1, st = c % STAGES;
                    if (j + 1 < p.nblk) issue_t(c + 1, rs);
                    mbar_wait(e_full(buf), (uint32_t)((c >> 1) & 1));
                    if (j == 0) mbar_wait(acc_empty, (uint32_t)(il & 1) ^ 1u);   // previous item's gradient tile has been read
                    tc_fence_after();
                    const uint32_t sx1 = sX + st * 2 * STR_BYTES, sx2 = sx1 + STR_BYTES;
#pragma unroll
                    for (int kk = 0; kk < BS / 16; ++kk) {          // 16 streamed rows per k-step
                        const uint64_t de1 = smem_desc(sE + (buf * NE) * E_BYTES + kk * 32, 0, 1024);
                        if (!kDKV) {
                            // dQ += dS K      (K block [64 keys][64 dims] consumed MN-major: 16 keys = 2048 B per k-step)
                            umma_bf16(tmem_base + ACC1_COL, de1, smem_desc(sx1 + kk * 2048, (uint32_t)STR_BYTES, 1024), idesc_acc,
                                      (j > 0 || kk > 0));
                        } else {
                            const uint64_t de2 = smem_desc(sE + (buf * NE + 1) * E_BYTES + kk * 32, 0, 1024);
                            // dV += P^T dO ;  dK += dS^T Q
                            umma_bf16(tmem_base + ACC1_COL, de1, smem_desc(sx2 + kk * 2048, (uint32_t)STR_BYTES, 1024), idesc_acc,
                                      (j > 0 || kk > 0));
                            umma_bf16(tmem_base + ACC2_COL, de2, smem_desc(sx1 + kk * 2048, (uint32_t)STR_BYTES, 1024), idesc_acc,
                                      (j > 0 || kk > 0));
                        }
                    }
                    umma_commit(e_empty(buf));
                    umma_commit(x_empty(st));
                }
                umma_commit(acc_full);
                umma_commit(r_empty(rs));
                c0 += p.nblk;
            }
        }
    } else {
        // ------------------------------ elementwise + epilogue --------------------
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;                           // which 32 of the 64 streamed columns / output dims
        const int r = quarter * 32 + lane;                          // resident row owned by this thread
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        uint8_t* sE_ptr = smem + (sE - base);
        const float sl2 = p.scale_log2, scale = p.scale;
        int c0 = 0;
        for (int il = 0; il < my_items; ++il) {
            const int item = blockIdx.x + il * gridDim.x;
            const int tile = item % p.ntiles, bh = item / p.ntiles;
            const int b = bh / p.H, h = bh % p.H;
            const int row = tile * BM + r;                          // query (dQ pass) or key (dKV pass) index
            float lse2_r = 0.f, dl_r = 0.f;
            if (!kDKV && row < p.N) {
                lse2_r = p.lse[((long long)b * p.H + h) * p.N + row] * LOG2E;
                dl_r = p.delta[((long long)b * p.H + h) * p.N + row];
            }
            for (int j = 0; j < p.nblk; ++j) {
                const int c = c0 + j;
                const int buf = c & 1, st = c % STAGES;
                mbar_wait(t_full(buf), (uint32_t)((c >> 1) & 1));
                tc_fence_after();
                uint32_t sv[32], dv[32];
                tmem_ld32(tmem_base + lane_addr + (uint32_t)(T1_COL + buf * BS + half * 32), sv);
                tmem_ld32(tmem_base + lane_addr + (uint32_t)(T2_COL + buf * BS + half * 32), dv);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(t_empty(buf));                          // products of block c+2 may overwrite this buffer
                const float* v_lse = vec + (st * 2 + 0) * BS + half * 32;
                const float* v_dl = vec + (st * 2 + 1) * BS + half * 32;
                const int col0 = j * BS + half * 32;                // streamed index of this thread's first column
                float pr[32], ds[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float l2 = kDKV ? v_lse[i] : lse2_r;
                    const float dl = kDKV ? v_dl[i] : dl_r;
                    const float pv = (col0 + i < p.N) ? ex2(fmaf(__uint_as_float(sv[i]), sl2, -l2)) : 0.f;
                    pr[i] = pv;
                    ds[i] = pv * (__uint_as_float(dv[i]) - dl) * scale;
                }
                mbar_wait(e_empty(buf), (uint32_t)((c >> 1) & 1) ^ 1u);         // accumulations of block c-2 consumed the buffer
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int kb8 = half * 4 + i;                   // 8-column group inside the 64-column atom
                    const uint32_t off = r * 128 + (((kb8 & 7) ^ (r & 7)) << 4);
                    uint4 pk;
                    if (kDKV) {
                        pk.x = pack_bf16(pr[8 * i], pr[8 * i + 1]);
                        pk.y = pack_bf16(pr[8 * i + 2], pr[8 * i + 3]);
                        pk.z = pack_bf16(pr[8 * i + 4], pr[8 * i + 5]);
                        pk.w = pack_bf16(pr[8 * i + 6], pr[8 * i + 7]);
                        *reinterpret_cast<uint4*>(sE_ptr + (buf * NE) * E_BYTES + off) = pk;
                    }
                    pk.x = pack_bf16(ds[8 * i], ds[8 * i + 1]);
                    pk.y = pack_bf16(ds[8 * i + 2], ds[8 * i + 3]);
                    pk.z = pack_bf16(ds[8 * i + 4], ds[8 * i + 5]);
                    pk.w = pack_bf16(ds[8 * i + 6], ds[8 * i + 7]);
                    *reinterpret_cast<uint4*>(sE_ptr + (buf * NE + (kDKV ? 1 : 0)) * E_BYTES + off) = pk;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to UMMA
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(e_full(buf));
            }
            // ---- gradient tile(s) of this item
            mbar_wait(acc_full, (uint32_t)(il & 1));
            tc_fence_after();
            uint32_t a1[32], a2[32];
            tmem_ld32(tmem_base + lane_addr + (uint32_t)(ACC1_COL + half * 32), a1);
            if (kDKV) tmem_ld32(tmem_base + lane_addr + (uint32_t)(ACC2_COL + half * 32), a2);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
            if (row < p.N) {
                float g1[32], g2[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    g1[i] = __uint_as_float(a1[i]);
                    g2[i] = kDKV ? __uint_as_float(a2[i]) : 0.f;
                }
                // the rotated gradient: dQ (dQ pass) or dK (dKV pass); token 0 (CLS) is not rotated
                float(&rot)[32] = kDKV ? g2 : g1;
                if (p.rope_cos != nullptr && row > 0) {
                    const float* cr = p.rope_cos + (long long)(row - 1) * HD;
                    const float* sr = p.rope_sin + (long long)(row - 1) * HD;
#pragma unroll
                    for (int i = 0; i < 16; ++i) unrope_pair(rot[2 * i], rot[2 * i + 1], cr, sr, half * 32 + 2 * i);
                }
                __nv_bfloat16* orow = p.dqkv + ((long long)b * p.N + row) * (3ll * D) + h * HD + half * 32;
                if (!kDKV) {
                    store_row32(orow, g1);                          // q section
                } else {
                    store_row32(orow + D, g2);                      // k section
                    store_row32(orow + 2 * D, g1);                  // v section
                }
            }
            c0 += p.nblk;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <bool kDKV>
static int launch(const CUtensorMap& r1, const CUtensorMap& r2, const CUtensorMap& x1, const CUtensorMap& x2, const Params& p,
                  cudaStream_t st) {
    constexpr int NE = kDKV ? 2 : 1;
    const int smem = 4 * RES_BYTES + STAGES * 2 * STR_BYTES + 2 * NE * E_BYTES + STAGES * 2 * BS * 4 + 512 + 1024;
    static bool configured = false;
    if (!configured) {
        CS_CUDA(cudaFuncSetAttribute(attention_bwd_tc_kernel<kDKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    const long long items = (long long)p.B * p.H * p.ntiles;
    const int grid = items < num_sms() ? (int)items : num_sms();
    attention_bwd_tc_kernel<kDKV><<<grid, THREADS, smem, st>>>(r1, r2, x1, x2, p);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

}  // namespace attn_bwd_tc

// dqkv (bf16 [B*N, 3D]) from qkv (post-RoPE q|k|v), d_out, lse and delta = rowsum(out o d_out); see cs_attention_bwd.
int attention_bwd_tc(const void* qkv, const void* d_out, const float* lse, const float* delta, int B, int N, int H,
                     float scale, const float* rope_cos, const float* rope_sin, void* dqkv, cudaStream_t st) {
    using namespace attn_bwd_tc;
    const int D = H * HD;
    const long long rows = (long long)B * N;
    if (N < 1 || rows >= (1ll << 31) || (long long)B * H * ceil_div(N, BM) >= (1ll << 31)) return CS_ERR_UNSUPPORTED;
    CUtensorMap qkv128, qkv64, do128, do64;
    int rc = make_map_bf16_2d(&qkv128, qkv, rows, 3 * D, 3 * D, HD, BM);
    if (rc) return rc;
    rc = make_map_bf16_2d(&qkv64, qkv, rows, 3 * D, 3 * D, HD, BS);
    if (rc) return rc;
    rc = make_map_bf16_2d(&do128, d_out, rows, D, D, HD, BM);
    if (rc) return rc;
    rc = make_map_bf16_2d(&do64, d_out, rows, D, D, HD, BS);
    if (rc) return rc;
    Params p;
    p.B = B; p.N = N; p.H = H;
    p.ntiles = ceil_div(N, BM);
    p.nblk = ceil_div(N, BS);
    p.scale = scale;
    p.scale_log2 = scale * LOG2E;
    p.lse = lse;
    p.delta = delta;
    p.rope_cos = rope_cos;
    p.rope_sin = rope_sin;
    p.dqkv = (__nv_bfloat16*)dqkv;
    // dQ pass: resident Q | dO tiles, streamed K | V blocks
    p.r1_col = 0; p.r2_col = 0; p.x1_col = D; p.x2_col = 2 * D;
    rc = launch<false>(qkv128, do128, qkv64, qkv64, p, st);
    if (rc) return rc;
    // dKV pass: resident K | V tiles, streamed Q | dO blocks
    p.r1_col = D; p.r2_col = 2 * D; p.x1_col = 0; p.x2_col = 0;
    return launch<true>(qkv128, qkv128, qkv64, do64, p, st);
}

}  // namespace cs
