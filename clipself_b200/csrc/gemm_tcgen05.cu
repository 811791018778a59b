// tcgen05 / TMA / TMEM GEMM for sm_100a:  C[M,N] = A[M,K] · W[N,K]^T  (bf16 in, f32 accumulate)
//
// One persistent CTA per SM, 192 threads:
//   warp 0      : TMA producer   (cp.async.bulk.tensor 2D, SWIZZLE_128B, 64-wide K slabs)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BLOCK_N x 16)
//   warps 2..5  : epilogue       (tcgen05.ld 32x32b -> registers -> fused epilogue -> global)
// Pipelines: smem ring full/empty (TMA <-> MMA), double-buffered TMEM accumulator full/empty
// (MMA <-> epilogue) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Fused epilogues (cs_epilogue_mode_t): bias / residual / alpha, RoPE on the q|k columns
// (rope.py:148-164 semantics, angles recomputed in-kernel from the 1-D position / frequency
// vectors instead of reading [tokens,64] tables), SwiGLU gate*up on packed weights, patch-embed
// token assembly.  Every (mode, out dtype, residual kind) is its own template instantiation so the
// epilogue is straight-line code.
#include "tc_common.cuh"

namespace cs {
namespace gemm {
using namespace cs::tc;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 B = one SWIZZLE_128B atom row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;

template <int BLOCK_N>
struct Cfg {
    static constexpr int STAGES = (BLOCK_N == 256) ? 4 : 6;
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = 2 * BLOCK_N;  // two accumulator stages (power of two)
    static constexpr int BAR_BYTES = 256;
    static constexpr int EPI_TILE_BYTES = 32 * 128;          // one swizzled 32-row x 128 B staging tile per warp
    static constexpr int EPI_BIAS_BYTES = BLOCK_N * 4;       // this tile's bias slice, one private copy per warp
    static constexpr int EPI_ROWSTAT_BYTES = 256;            // (mean, rstd) of this warp's 32 rows (LN folding)
    static constexpr int EPI_WARP_BYTES = EPI_TILE_BYTES + 2 * EPI_BIAS_BYTES + EPI_ROWSTAT_BYTES;  // bias + ln_c1
    static constexpr int EPI_BYTES = 4 * EPI_WARP_BYTES;
    static constexpr int ROPE_BYTES = 2 * 16 * 64 * 4;       // cos / sin of pos[g] * freq[q]: [2][16 freqs][64 grid positions]
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + ROPE_BYTES + BAR_BYTES + 1024;  // +1024 align slack
};

struct EpiParams {
    int mode;
    int out_bf16;
    void* out;
    long long ldo;
    const float* bias;
    const float* residual;
    long long ldr;
    const float* rope_pos;    // [rope_grid]  position value of grid index i  (rope.py:127: i/ft*pt)
    const float* rope_freq;   // [16]         theta^(-2n/32)                  (rope.py:118)
    int rope_grid;
    int tokens;
    int rope_cols;
    const float* pos_embed;
    float alpha;
    int dbg;   // debug switches for epilogue ablations (cs_gemm_epilogue_t.reserved); 0 in production
    const float* ln_stats;   // LayerNorm folding (see cs_gemm_epilogue_t)
    const float* ln_c1;
    int ln_parts;
    float ln_inv_dim;
    float ln_eps;
    float* stats_out;        // SWIGLU: per (row, tile) partial sum / sumsq of the bf16 outputs
    int k_splits;            // split-K factor (RES_RED accumulate into a zeroed output), 1 = off
    int kb_per_split;
};

enum ResKind { RES_NONE = 0, RES_LOAD = 1, RES_RED = 2 };

// one epilogue warp: 32 accumulator rows of one tile
template <int BLOCK_N, int MODE, bool OUT_BF16, int RES, bool LNFOLD>
__device__ __forceinline__ void epilogue_tile(const EpiParams& ep, uint32_t taddr, int mw, int n0, int M, int N,
                                              uint8_t* st, float* sbias, const float* srope, int lane, bool use_bias) {
    using C = Cfg<BLOCK_N>;
    const int crow = lane >> 3;                   // coalesced phase: row within a group of 4
    const int cchunk = lane & 7;                  // coalesced phase: 16 B chunk of the 128 B row slice
    // this tile's bias slice -> private smem copy (replaces dependent global loads in the hot loop)
    if (ep.bias != nullptr && use_bias) {
#pragma unroll
        for (int j = lane * 4; j < BLOCK_N; j += 128) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + j < N) b = *reinterpret_cast<const float4*>(ep.bias + n0 + j);
            *reinterpret_cast<float4*>(sbias + j) = b;
        }
    } else {
#pragma unroll
        for (int j = lane * 4; j < BLOCK_N; j += 128) *reinterpret_cast<float4*>(sbias + j) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float* sc1 = sbias + BLOCK_N;
    float* srow = sc1 + BLOCK_N;
    if constexpr (LNFOLD) {
#pragma unroll
        for (int j = lane * 4; j < BLOCK_N; j += 128) {
            float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + j < N) c = *reinterpret_cast<const float4*>(ep.ln_c1 + n0 + j);
            *reinterpret_cast<float4*>(sc1 + j) = c;
        }
        // (mean, rstd) of input row mw + lane from its partial sums
        float mu = 0.f, rs = 0.f;
        if (mw + lane < M) {
            const float* ps = ep.ln_stats + (long long)(mw + lane) * ep.ln_parts * 2;
            float s1 = 0.f, s2 = 0.f;
            for (int q = 0; q < ep.ln_parts; q += 2) {
                const float4 v = *reinterpret_cast<const float4*>(ps + q * 2);
                s1 += v.x + v.z;
                s2 += v.y + v.w;
            }
            mu = s1 * ep.ln_inv_dim;
            const float var = fmaxf(s2 * ep.ln_inv_dim - mu * mu, 0.f);
            rs = rsqrtf(var + ep.ln_eps);
        }
        srow[lane * 2] = mu;
        srow[lane * 2 + 1] = rs;
    }
    __syncwarp();

    if constexpr (OUT_BF16) {
        // ---------------- bf16 outputs: math in registers (row per thread), packed staging --------
        const int m = mw + lane;
        int gi = 0, gj = 0;
        bool rot = false;
        if constexpr (MODE == CS_EPI_QKV_ROPE) {
            const int tok = m % ep.tokens;
            rot = tok > 0;
            const int p = tok > 0 ? tok - 1 : 0;
            gi = p / ep.rope_grid;
            gj = p % ep.rope_grid;
        }
        constexpr int OUT_COLS = (MODE == CS_EPI_SWIGLU) ? BLOCK_N / 2 : BLOCK_N;
        const int out_n0 = (MODE == CS_EPI_SWIGLU) ? (n0 >> 1) : n0;
        const int out_N = (MODE == CS_EPI_SWIGLU) ? (N >> 1) : N;
        float st1 = 0.f, st2 = 0.f;                          // SWIGLU stats_out: row sum / sum of squares of the stored values
#pragma unroll 1
        for (int gc = 0; gc < OUT_COLS; gc += 64) {          // 64 bf16 output columns = 128 B per row
            if (out_n0 + gc >= out_N) break;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int c = gc + half * 32;                // output column offset inside the tile
                float v[32];
                if constexpr (MODE == CS_EPI_SWIGLU) {
                    uint32_t rg[32], ru[32];
                    tmem_ld32(taddr + c, rg);
                    tmem_ld32(taddr + 128 + c, ru);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float g = __uint_as_float(rg[j]) * ep.alpha + sbias[c + j];
                        const float u = __uint_as_float(ru[j]) * ep.alpha + sbias[128 + c + j];
                        v[j] = __fdividef(g, 1.0f + __expf(-g)) * u;
                    }
                } else {
                    uint32_t r[32];
                    tmem_ld32(taddr + c, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * ep.alpha + sbias[c + j];
                    if constexpr (MODE == CS_EPI_QKV_ROPE) {
                        const int n = n0 + c;
                        if (n < ep.rope_cols && rot) {
                            // head-dim offset 0..31 rotates by the token's grid ROW, 32..63 by its COLUMN
                            const int g = (n & 32) ? gj : gi;
#pragma unroll
                            for (int q = 0; q < 16; ++q) {
                                const float cs_ = srope[q * 64 + g], sn = srope[1024 + q * 64 + g];
                                const float x0 = v[2 * q], x1 = v[2 * q + 1];
                                v[2 * q] = x0 * cs_ - x1 * sn;
                                v[2 * q + 1] = x1 * cs_ + x0 * sn;
                            }
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 pk;
                    pk.x = pack_bf16(v[8 * j], v[8 * j + 1]);
                    pk.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
                    pk.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
                    pk.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
                    const int chunk = half * 4 + j;
                    *reinterpret_cast<uint4*>(st + lane * 128 + ((chunk ^ (lane & 7)) << 4)) = pk;
                    if constexpr (MODE == CS_EPI_SWIGLU) {
                        // row statistics for the folded ffn_ln (from the f32 values: the bf16 rounding of
                        // the stored h is zero-mean noise ~2^-9/sqrt(n) on the mean, negligible)
#pragma unroll
                        for (int z = 0; z < 8; ++z) {
                            st1 += v[8 * j + z];
                            st2 += v[8 * j + z] * v[8 * j + z];
                        }
                    }
                }
            }
            __syncwarp();
            const int ocol = out_n0 + gc + cchunk * 8;
            uint4 val[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = i * 4 + crow;
                val[i] = *reinterpret_cast<const uint4*>(st + rr * 128 + ((cchunk ^ (rr & 7)) << 4));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int mm = mw + i * 4 + crow;
                if (mm < M && ocol < out_N)
                    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + (long long)mm * ep.ldo + ocol) = val[i];
            }
            __syncwarp();
        }
        if constexpr (MODE == CS_EPI_SWIGLU) {
            if (ep.stats_out != nullptr && m < M)
                *reinterpret_cast<float2*>(ep.stats_out + ((long long)m * (N / BLOCK_N) + n0 / BLOCK_N) * 2) = make_float2(st1, st2);
        }
    } else {
        // ---------------- f32 outputs: raw staging, math in the coalesced phase ----------------
        long long orow[8];
        int prow[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int mm = mw + i * 4 + crow;
            orow[i] = mm;
            prow[i] = 0;
            if constexpr (MODE == CS_EPI_TOKENS) {
                orow[i] = (long long)mm + mm / (ep.tokens - 1) + 1;
                prow[i] = mm % (ep.tokens - 1) + 1;
            }
            if (mm >= M) orow[i] = -1;
        }
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 32) {              // 32 f32 output columns = 128 B per row
            const int n = n0 + c;
            if (n >= N) break;
            const int nc = n + cchunk * 4;
            const bool col_ok = nc < N;
            // residual / pos_embed reads do not depend on the accumulator: issue them first
            float4 extra[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                extra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (orow[i] >= 0 && col_ok) {
                    if constexpr (RES == RES_LOAD) extra[i] = *reinterpret_cast<const float4*>(ep.residual + orow[i] * ep.ldr + nc);
                    if constexpr (MODE == CS_EPI_TOKENS) extra[i] = *reinterpret_cast<const float4*>(ep.pos_embed + (long long)prow[i] * N + nc);
                }
            }
            {
                uint32_t r[32];
                tmem_ld32(taddr + c, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 f = make_float4(__uint_as_float(r[4 * j]) * ep.alpha, __uint_as_float(r[4 * j + 1]) * ep.alpha,
                                                 __uint_as_float(r[4 * j + 2]) * ep.alpha, __uint_as_float(r[4 * j + 3]) * ep.alpha);
                    *reinterpret_cast<float4*>(st + lane * 128 + ((j ^ (lane & 7)) << 4)) = f;
                }
            }
            __syncwarp();
            const float4 b4 = *reinterpret_cast<const float4*>(sbias + c + cchunk * 4);
            float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if constexpr (LNFOLD) c4 = *reinterpret_cast<const float4*>(sc1 + c + cchunk * 4);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = i * 4 + crow;
                float4 f = *reinterpret_cast<const float4*>(st + rr * 128 + ((cchunk ^ (rr & 7)) << 4));
                if constexpr (LNFOLD) {
                    const float2 ms = *reinterpret_cast<const float2*>(srow + rr * 2);      // (mean, rstd) of the input row
                    f.x = ms.y * (f.x - ms.x * c4.x);
                    f.y = ms.y * (f.y - ms.x * c4.y);
                    f.z = ms.y * (f.z - ms.x * c4.z);
                    f.w = ms.y * (f.w - ms.x * c4.w);
                }
                f.x += b4.x + extra[i].x;
                f.y += b4.y + extra[i].y;
                f.z += b4.z + extra[i].z;
                f.w += b4.w + extra[i].w;
                if (orow[i] >= 0 && col_ok) {
                    float* dst = reinterpret_cast<float*>(ep.out) + orow[i] * ep.ldo + nc;
                    if constexpr (RES == RES_RED) {
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(f.x), "f"(f.y), "f"(f.z), "f"(f.w)
                                     : "memory");
                    } else {
                        *reinterpret_cast<float4*>(dst) = f;
                    }
                }
            }
            __syncwarp();
        }
    }
}

template <int BLOCK_N, int MODE, bool OUT_BF16, int RES, bool LNFOLD>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
            int M, int N, int K, const EpiParams ep) {
    using C = Cfg<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;   // SWIZZLE_128B tiles need 1024 B alignment
    uint8_t* smem = smem_raw + (base - raw_addr);

    constexpr int EPI_OFF = C::STAGES * C::STAGE_BYTES;
    constexpr int ROPE_OFF = EPI_OFF + C::EPI_BYTES;
    constexpr int BAR_OFF = ROPE_OFF + C::ROPE_BYTES;
    const uint32_t bar_base = base + BAR_OFF;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 4);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + BAR_OFF + 8 * (2 * C::STAGES + 4));
    float* srope = reinterpret_cast<float*>(smem + ROPE_OFF);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int num_m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
    const int num_n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
    const int num_mn = num_m_tiles * num_n_tiles;
    const int num_tiles = num_mn * ep.k_splits;          // split-K: tile t -> (mn = t % num_mn, split = t / num_mn)
    const int num_kb_total = (K + BLOCK_K - 1) / BLOCK_K;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 4 * 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if constexpr (MODE == CS_EPI_QKV_ROPE) {
        // rotary table of this launch: angle(g, q) = pos[g] * freq[q] (f32 product as in rope.py:129), accurate
        // sincosf once per CTA; layout [q][g] so that lanes with different grid positions hit different banks
        for (int i = threadIdx.x; i < 16 * 64; i += NUM_THREADS) {
            const int q = i >> 6, g = i & 63;
            float sn = 0.f, cs_ = 1.f;
            if (g < ep.rope_grid) sincosf(ep.rope_pos[g] * ep.rope_freq[q], &sn, &cs_);
            srope[i] = cs_;
            srope[1024 + i] = sn;
        }
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int mn = tile % num_mn, sp = tile / num_mn;
                const int m0 = (mn / num_n_tiles) * BLOCK_M;
                const int n0 = (mn % num_n_tiles) * BLOCK_N;
                const int kb0 = sp * ep.kb_per_split;
                const int kb1 = min(num_kb_total, kb0 + ep.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = base + stage * C::STAGE_BYTES;
                    const uint32_t sb = sa + C::A_BYTES;
                    mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
                    tma_load_2d(sa, &map_a, full_bar(stage), kb * BLOCK_K, m0);
                    tma_load_2d(sb, &map_b, full_bar(stage), kb * BLOCK_K, n0);
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer --------------------------------
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);   // epilogue drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
                const int kb0 = (tile / num_mn) * ep.kb_per_split;
                const int kb1 = min(num_kb_total, kb0 + ep.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(full_bar(stage), phase);         // TMA bytes landed
                    tc_fence_after();
                    const uint32_t sa = base + stage * C::STAGE_BYTES;
                    const uint32_t sb = sa + C::A_BYTES;
                    const uint64_t da = make_smem_desc(sa);
                    const uint64_t db = make_smem_desc(sb);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // advance 16 bf16 = 32 B along K inside the swizzle atom: +2 in (addr>>4)
                        umma_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                  (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage));             // frees the smem slot when MMAs retire
                    if (kb == kb1 - 1) umma_commit(tfull_bar(acc));
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else {
        // ------------------------------ epilogue -----------------------------------
        // Each warp owns 32 accumulator rows (its TMEM lane quarter); see epilogue_tile().
        const int quarter = warp & 3;
        uint8_t* st = smem + EPI_OFF + (warp - 2) * C::EPI_WARP_BYTES;
        float* sbias = reinterpret_cast<float*>(st + C::EPI_TILE_BYTES);
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int mn = tile % num_mn;
            const int m0 = (mn / num_n_tiles) * BLOCK_M;
            const int n0 = (mn % num_n_tiles) * BLOCK_N;
            const int acc = it & 1;
            const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            if (!(ep.dbg & 8))
                epilogue_tile<BLOCK_N, MODE, OUT_BF16, RES, LNFOLD>(ep, taddr, m0 + quarter * 32, n0, M, N, st, sbias, srope, lane,
                                                                    tile < num_mn);
            tc_fence_before();
            mbar_arrive(tempty_bar(acc));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ----------------------------------------------------------------------------------------
// Host side
// ----------------------------------------------------------------------------------------
template <int BLOCK_N, int MODE, bool OUT_BF16, int RES, bool LNFOLD = false>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, int M, int N, int K, const EpiParams& ep,
                  cudaStream_t stream) {
    using C = Cfg<BLOCK_N>;
    static bool configured = false;
    if (!configured) {
        CS_CUDA(cudaFuncSetAttribute(gemm_kernel<BLOCK_N, MODE, OUT_BF16, RES, LNFOLD>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        configured = true;
    }
    const int tiles = ceil_div(M, BLOCK_M) * ceil_div(N, BLOCK_N);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    gemm_kernel<BLOCK_N, MODE, OUT_BF16, RES, LNFOLD><<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(ma, mb, M, N, K, ep);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

template <int BLOCK_N>
static int dispatch(const CUtensorMap& ma, const CUtensorMap& mb, int M, int N, int K, const EpiParams& ep, int res,
                    cudaStream_t st) {
    switch (ep.mode) {
        case CS_EPI_STORE:
            if (ep.out_bf16) return launch<BLOCK_N, CS_EPI_STORE, true, RES_NONE>(ma, mb, M, N, K, ep, st);
            if (res == RES_RED && ep.ln_stats != nullptr) return launch<BLOCK_N, CS_EPI_STORE, false, RES_RED, true>(ma, mb, M, N, K, ep, st);
            if (res == RES_RED) return launch<BLOCK_N, CS_EPI_STORE, false, RES_RED>(ma, mb, M, N, K, ep, st);
            if (res == RES_LOAD) return launch<BLOCK_N, CS_EPI_STORE, false, RES_LOAD>(ma, mb, M, N, K, ep, st);
            return launch<BLOCK_N, CS_EPI_STORE, false, RES_NONE>(ma, mb, M, N, K, ep, st);
        case CS_EPI_QKV_ROPE:
            return launch<BLOCK_N, CS_EPI_QKV_ROPE, true, RES_NONE>(ma, mb, M, N, K, ep, st);
        case CS_EPI_TOKENS:
            return launch<BLOCK_N, CS_EPI_TOKENS, false, RES_NONE>(ma, mb, M, N, K, ep, st);
        case CS_EPI_SWIGLU:
            if constexpr (BLOCK_N == 256) return launch<256, CS_EPI_SWIGLU, true, RES_NONE>(ma, mb, M, N, K, ep, st);
    }
    set_error("cs_gemm_bf16: unsupported epilogue combination");
    return CS_ERR_UNSUPPORTED;
}

}  // namespace gemm
}  // namespace cs

extern "C" int cs_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int64_t M, int N,
                            int K, const cs_gemm_epilogue_t* e, void* stream) {
    using namespace cs;
    using namespace cs::gemm;
    CS_CHECK_ARG(A && W && e && e->out, "cs_gemm_bf16: null pointer");
    CS_CHECK_ARG(M > 0 && N > 0 && K > 0 && M < (1ll << 31), "cs_gemm_bf16: bad shape M=%lld N=%d K=%d",
                 (long long)M, N, K);
    CS_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K,
                 "cs_gemm_bf16: lda/ldw must be multiples of 8 and >= K (lda=%lld ldw=%lld K=%d)",
                 (long long)lda, (long long)ldw, K);
    CS_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0), "cs_gemm_bf16: A/W must be 16 B aligned");
    CS_CHECK_ARG(N % 32 == 0, "cs_gemm_bf16: N must be a multiple of 32 (N=%d)", N);
    CS_CHECK_ARG(e->mode >= CS_EPI_STORE && e->mode <= CS_EPI_TOKENS, "cs_gemm_bf16: bad epilogue mode %d", e->mode);
    CS_CHECK_ARG(((uintptr_t)e->out % 16 == 0) && (e->ldo % (e->out_dtype == CS_BF16 ? 8 : 4) == 0),
                 "cs_gemm_bf16: out must be 16 B aligned with 16 B aligned rows");
    if (e->residual) CS_CHECK_ARG(((uintptr_t)e->residual % 16 == 0) && e->ldr % 4 == 0, "cs_gemm_bf16: residual alignment");
    if (e->bias) CS_CHECK_ARG((uintptr_t)e->bias % 16 == 0, "cs_gemm_bf16: bias alignment");

    EpiParams ep;
    ep.mode = e->mode;
    ep.out_bf16 = e->out_dtype == CS_BF16;
    ep.out = e->out;
    ep.ldo = e->ldo;
    ep.bias = e->bias;
    ep.residual = e->residual;
    ep.ldr = e->ldr;
    ep.rope_pos = e->rope_pos;
    ep.rope_freq = e->rope_freq;
    ep.rope_grid = e->rope_grid;
    ep.tokens = e->tokens;
    ep.rope_cols = e->rope_cols;
    ep.pos_embed = e->pos_embed;
    ep.alpha = e->alpha;
    ep.dbg = e->reserved;
    ep.ln_stats = e->ln_stats;
    ep.ln_c1 = e->ln_c1;
    ep.ln_parts = e->ln_parts;
    ep.ln_inv_dim = e->ln_dim > 0 ? 1.0f / (float)e->ln_dim : 0.f;
    ep.ln_eps = e->ln_eps;
    ep.stats_out = e->stats_out;
    ep.k_splits = 1;
    ep.kb_per_split = ceil_div(K, BLOCK_K);

    bool use256 = (N % 256 == 0);
    if (e->mode == CS_EPI_SWIGLU) {
        CS_CHECK_ARG(N % 256 == 0 && e->bias && e->out_dtype == CS_BF16,
                     "cs_gemm_bf16: SWIGLU needs packed N %% 256 == 0, bias, bf16 out");
        use256 = true;
    }
    if (e->mode == CS_EPI_QKV_ROPE)
        CS_CHECK_ARG(e->rope_pos && e->rope_freq && e->rope_grid > 0 && e->rope_grid <= 64 &&
                         e->tokens == e->rope_grid * e->rope_grid + 1 && e->rope_cols % 64 == 0 && e->out_dtype == CS_BF16,
                     "cs_gemm_bf16: QKV_ROPE needs rope_pos/rope_freq, grid <= 64, tokens == grid^2+1, rope_cols %% 64 == 0, bf16 out");
    if (e->mode == CS_EPI_TOKENS)
        CS_CHECK_ARG(e->pos_embed && e->tokens > 1 && e->out_dtype == CS_F32 && ((uintptr_t)e->pos_embed % 16 == 0),
                     "cs_gemm_bf16: TOKENS needs pos_embed, tokens, f32 out");

    CUtensorMap ma, mb;
    int rc = make_map_bf16_2d(&ma, A, M, K, lda, BLOCK_K, BLOCK_M);
    if (rc) return rc;
    rc = make_map_bf16_2d(&mb, W, N, K, ldw, BLOCK_K, use256 ? 256 : 128);
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int res = RES_NONE;
    if (e->residual) {
        CS_CHECK_ARG(e->out_dtype == CS_F32 && e->mode == CS_EPI_STORE, "cs_gemm_bf16: residual needs f32 STORE output");
        res = (e->residual == e->out && e->ldr == e->ldo) ? RES_RED : RES_LOAD;
    }
    if (e->ln_stats)
        CS_CHECK_ARG(res == RES_RED && e->ln_c1 && e->ln_parts > 0 && e->ln_parts % 2 == 0 && e->ln_dim > 0 &&
                         ((uintptr_t)e->ln_stats % 16 == 0) && ((uintptr_t)e->ln_c1 % 16 == 0),
                     "cs_gemm_bf16: LN folding needs the in-place f32 residual epilogue, ln_c1, even ln_parts, ln_dim");
    if (e->stats_out) CS_CHECK_ARG(e->mode == CS_EPI_SWIGLU, "cs_gemm_bf16: stats_out is a SWIGLU output");
    if (e->reserved2 != 0) {
        // split-K (reserved2 = requested splits, -1 = choose): partial products are accumulated with
        // red.add into `out`, which the caller must have zeroed.  f32 STORE without residual only.
        CS_CHECK_ARG(e->mode == CS_EPI_STORE && e->out_dtype == CS_F32 && e->residual == nullptr && e->ln_stats == nullptr,
                     "cs_gemm_bf16: split-K needs a plain f32 STORE epilogue");
        const int num_kb = ceil_div(K, BLOCK_K);
        const int tiles = ceil_div(M, BLOCK_M) * ceil_div(N, use256 ? 256 : 128);
        int best = 1;
        if (e->reserved2 > 0) {
            best = e->reserved2;
        } else {
            long long best_cost = -1;
            for (int sp = 1; sp <= 8 && sp <= num_kb; ++sp) {
                const long long waves = ceil_div((long long)tiles * sp, num_sms());
                const long long cost = waves * (ceil_div(num_kb, sp) + 8);     // +8 k-blocks ~ per-tile epilogue/fill cost
                if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = sp; }
            }
        }
        if (best > 1) {
            ep.kb_per_split = ceil_div(num_kb, best);
            ep.k_splits = ceil_div(num_kb, ep.kb_per_split);        // no empty split
            ep.residual = reinterpret_cast<const float*>(ep.out);     // selects the red.add epilogue
            ep.ldr = ep.ldo;
            res = RES_RED;
        }
    }
    return use256 ? dispatch<256>(ma, mb, (int)M, N, K, ep, res, st) : dispatch<128>(ma, mb, (int)M, N, K, ep, res, st);
}
