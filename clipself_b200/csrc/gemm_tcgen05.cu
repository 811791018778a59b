// tcgen05 / TMA / TMEM GEMM for sm_100a:  C[M,N] = A[M,K] · W[N,K]^T  (bf16 in, f32 accumulate)
//
// One persistent CTA per SM, 320 threads; with BLOCK_N = 256 the CTAs work in PAIRS (cluster of 2, tcgen05 cta_group::2):
//   warp 0      : TMA producer   (cp.async.bulk.tensor 2D, SWIZZLE_128B, 64-wide K slabs)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128|256 x BLOCK_N x 16; pair: leader CTA only)
//   warps 2..9  : epilogue       (tcgen05.ld 32x32b -> registers -> fused epilogue -> global), two per TMEM lane quarter
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
constexpr int EPI_WARPS = 8;                       // two per TMEM lane quarter, splitting the tile's columns
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;   // + TMA warp + MMA warp
constexpr int ROPE_STRIDE = 65;                    // [16 freqs][64 grid positions + 1 identity slot]; odd stride: no bank conflicts

// CG = 1: one CTA per 128 x BLOCK_N tile (tcgen05 cta_group::1).
// CG = 2: a CTA pair (cluster of 2 on one TPC) per 256 x BLOCK_N tile: UMMA M = 256, each CTA owns 128 rows of the
//         accumulator in its own TMEM and stages only HALF of the B tile (the tensor cores of the pair share it), so the
//         operand bytes a CTA pulls from L2 per flop drop by a third and 6 instead of 4 stages fit.
// EMIT (new residual stream as f32 + bf16 + statistics): the outputs leave through double-buffered staging tiles and
//         cp.async.bulk.tensor stores (direct stores stalled the epilogue warps: 243 -> 143 us without them), paid for with
//         one mainloop stage.
template <int BLOCK_N, int CG = 1, bool EMIT = false>
struct Cfg {
    static constexpr int STAGES = (BLOCK_N == 256 && CG == 1) ? (EMIT ? 3 : 4) : (EMIT ? 5 : 6);
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_ROWS = BLOCK_N / CG;               // rows of W staged by one CTA
    static constexpr int B_BYTES = B_ROWS * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = 2 * BLOCK_N;  // two accumulator stages (power of two)
    static constexpr int HALF_N = BLOCK_N / 2;     // accumulator columns owned by one epilogue warp
    static constexpr int BAR_BYTES = 256;
    static constexpr int EPI_TILE_BYTES = EMIT ? 2 * (32 * 64 + 32 * 32) : 32 * 64;   // one swizzled 32-row x 64 B staging tile per warp
                                                             // (EMIT: two f32 tiles, then two 32-row x 32 B bf16 tiles)
    static constexpr int EPI_VEC_BYTES = HALF_N * 4;         // this warp's bias slice / ln_c1 slice
    static constexpr int EPI_WARP_BYTES = EPI_TILE_BYTES + 2 * EPI_VEC_BYTES;
    static constexpr int EPI_BYTES = EPI_WARPS * EPI_WARP_BYTES;
    static constexpr int ROPE_BYTES = 2 * 16 * ROPE_STRIDE * 4;   // cos / sin of pos[g] * freq[q]
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + ROPE_BYTES + BAR_BYTES + 1024;  // +1024 align slack
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
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
    float* stats_out;        // SWIGLU / EMIT: per (row, tile, column half) partial sum / sumsq of the outputs
    __nv_bfloat16* out2;     // EMIT: bf16 copy of the f32 output
    long long ldo2;
    int k_splits;            // split-K factor (RES_RED accumulate into a zeroed output), 1 = off
    int kb_per_split;
};

enum ResKind { RES_NONE = 0, RES_LOAD = 1, RES_RED = 2 };

// One epilogue warp: 32 accumulator rows (its TMEM lane quarter) x one column half of one tile.
//   bf16 outputs: math in the row-per-thread layout of tcgen05.ld.32x32b, 32 output columns (64 B per row) per round
//   f32 outputs : 16 columns (64 B per row) per round, bias / residual / statistics in the coalesced phase
// Either way a round goes through this warp's XOR-swizzled 32 x 64 B staging tile so that global accesses are
// whole 32 B sectors of consecutive rows.
template <int BLOCK_N, int MODE, bool OUT_BF16, int RES, bool LNFOLD, bool EMIT>
__device__ __forceinline__ void epilogue_tile(const EpiParams& ep, uint32_t taddr, int mw, int n0, int half, int M, int N,
                                              uint8_t* st, float* sbias, const float* srope, int lane, bool use_bias,
                                              uint32_t tfull, uint32_t tfull_phase, const CUtensorMap* map_o32,
                                              const CUtensorMap* map_o16) {
    using C = Cfg<BLOCK_N>;
    constexpr int HALF_N = C::HALF_N;
    constexpr bool SWI = (MODE == CS_EPI_SWIGLU);
    float* sc1 = sbias + HALF_N;
    // accumulator column of local index j of this warp's slice (SWIGLU: 64 gate columns then the 64 matching up columns)
    auto acc_col = [&](int j) { return SWI ? ((j < HALF_N / 2) ? half * (HALF_N / 2) + j : BLOCK_N / 2 + half * (HALF_N / 2) + (j - HALF_N / 2))
                                           : half * HALF_N + j; };
    // this warp's bias / ln_c1 slices -> private smem (broadcast reads in the hot loops)
#pragma unroll
    for (int j = lane * 4; j < HALF_N; j += 128) {
        const int n = n0 + acc_col(j);
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f), c = b;
        if (n < N) {
            if (ep.bias != nullptr && use_bias) b = *reinterpret_cast<const float4*>(ep.bias + n);
            if constexpr (LNFOLD) c = *reinterpret_cast<const float4*>(ep.ln_c1 + n);
        }
        *reinterpret_cast<float4*>(sbias + j) = b;
        if constexpr (LNFOLD) *reinterpret_cast<float4*>(sc1 + j) = c;
    }
    // (mean, rstd) of input row mw + lane from its partial sums
    float rs = 1.f, rm = 0.f;                     // rstd, rstd * mean
    if constexpr (LNFOLD) {
        float mu = 0.f;
        rs = 0.f;
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
        rm = rs * mu;
    }
    __syncwarp();
    const int swz = (lane >> 1) & 3;                     // row-per-thread phase: 16 B chunk j of row `lane` lives at j ^ swz
    const int crow = lane >> 2;                          // coalesced phase: row within a group of 8
    const int cch = lane & 3;                            // coalesced phase: 16 B chunk of the 64 B row slice
    // Everything above (bias / ln_c1 slices, row statistics) and the first residual reads below do not depend on the
    // accumulator: they are issued BEFORE waiting for the MMAs of this tile, so their latency hides under the mainloop.
    auto wait_accumulator = [&]() {
        mbar_wait(tfull, tfull_phase);
        tc_fence_after();
    };
    if (ep.dbg & 8) {                                    // ablation: mainloop + handshakes only
        wait_accumulator();
        return;
    }

    if constexpr (OUT_BF16) {
        wait_accumulator();
        const int m = mw + lane;
        int gi = 64, gj = 64;                            // slot 64 of the rotary table = identity (CLS rows)
        if constexpr (MODE == CS_EPI_QKV_ROPE) {
            const int tok = m % ep.tokens;
            if (tok > 0) {
                gi = (tok - 1) / ep.rope_grid;
                gj = (tok - 1) % ep.rope_grid;
            }
        }
        constexpr int OUT_HALF = SWI ? HALF_N / 2 : HALF_N;        // output columns of this warp
        const int out_n0 = (SWI ? (n0 >> 1) : n0) + half * OUT_HALF;
        const int out_N = SWI ? (N >> 1) : N;
        float st1 = 0.f, st2 = 0.f;                      // SWIGLU stats_out: row sum / sum of squares of the f32 values
#pragma unroll 1
        for (int r = 0; r < OUT_HALF; r += 32) {
            if (out_n0 + r >= out_N) break;
            float v[32];
            if constexpr (SWI) {
                uint32_t rg[32], ru[32];
                tmem_ld32(taddr + half * OUT_HALF + r, rg);
                tmem_ld32(taddr + BLOCK_N / 2 + half * OUT_HALF + r, ru);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float g, u;
                    if constexpr (LNFOLD) {
                        g = fmaf(rs, __uint_as_float(rg[j]), fmaf(-rm, sc1[r + j], sbias[r + j]));
                        u = fmaf(rs, __uint_as_float(ru[j]), fmaf(-rm, sc1[OUT_HALF + r + j], sbias[OUT_HALF + r + j]));
                    } else {
                        g = fmaf(__uint_as_float(rg[j]), ep.alpha, sbias[r + j]);
                        u = fmaf(__uint_as_float(ru[j]), ep.alpha, sbias[OUT_HALF + r + j]);
                    }
                    v[j] = __fdividef(g, 1.0f + __expf(-g)) * u;
                    st1 += v[j];
                    st2 = fmaf(v[j], v[j], st2);
                }
            } else {
                uint32_t a[32];
                if (ep.dbg & 4) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) a[j] = (uint32_t)(r + j + lane);
                } else {
                    tmem_ld32(taddr + half * HALF_N + r, a);
                    tmem_ld_wait();
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if constexpr (LNFOLD) v[j] = fmaf(rs, __uint_as_float(a[j]), fmaf(-rm, sc1[r + j], sbias[r + j]));
                    else v[j] = fmaf(__uint_as_float(a[j]), ep.alpha, sbias[r + j]);
                }
                if constexpr (MODE == CS_EPI_QKV_ROPE) {
                    const int n = n0 + half * HALF_N + r;
                    if (n < ep.rope_cols) {
                        // head-dim offset 0..31 rotates by the token's grid ROW, 32..63 by its COLUMN
                        const float* tc = srope + ((n & 32) ? gj : gi);
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            const float cs_ = tc[q * ROPE_STRIDE], sn = tc[16 * ROPE_STRIDE + q * ROPE_STRIDE];
                            const float x0 = v[2 * q], x1 = v[2 * q + 1];
                            v[2 * q] = fmaf(x0, cs_, -x1 * sn);
                            v[2 * q + 1] = fmaf(x1, cs_, x0 * sn);
                        }
                    }
                }
            }
            if (ep.dbg & 2) {            // ablation: math only (keep the values alive without touching smem / global)
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) acc += v[j];
                if (acc == 1.2345e-30f) *reinterpret_cast<float*>(ep.out) = acc;
                continue;
            }
            // the staging tile (32 rows x 64 B, SWIZZLE_64B pattern) leaves as ONE bulk tensor store: rows beyond M are clipped
            // by the map, no read-back and no per-thread global stores (they cost ~10 % of the kernel).  The previous
            // round's store must have read the tile before it is rewritten.
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 pk;
                pk.x = pack_bf16(v[8 * j], v[8 * j + 1]);
                pk.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
                pk.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
                pk.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
                *reinterpret_cast<uint4*>(st + lane * 64 + ((j ^ swz) << 4)) = pk;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0 && !(ep.dbg & 1)) {      // ablation bit 0: no global stores
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map_o16),
                             "r"(smem_u32(st)), "r"(out_n0 + r), "r"(mw)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if constexpr (SWI) {
            if (ep.stats_out != nullptr && m < M)
                *reinterpret_cast<float2*>(ep.stats_out + (((long long)m * (N / BLOCK_N) + n0 / BLOCK_N) * 2 + half) * 2) =
                    make_float2(st1, st2);
        }
    } else {
        // ---------------- f32 outputs ----------------
        constexpr int ROUNDS = HALF_N / 16;
        constexpr bool HAS_EXTRA = (RES == RES_LOAD) || (MODE == CS_EPI_TOKENS);
        constexpr int PF = HAS_EXTRA ? 4 : 1;            // rounds of residual / pos_embed reads kept in flight per thread
        long long orow[4];
        int prow[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int mm = mw + i * 8 + crow;
            orow[i] = mm;
            prow[i] = 0;
            if constexpr (MODE == CS_EPI_TOKENS) {
                orow[i] = (long long)mm + mm / (ep.tokens - 1) + 1;
                prow[i] = mm % (ep.tokens - 1) + 1;
            }
            if (mm >= M) orow[i] = -1;
        }
        const int nbase = n0 + half * HALF_N + cch * 4;  // this lane's first output column
        float4 extra[PF][4];
        auto fetch_extra = [&](int rd, float4 (&dst)[4]) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (orow[i] >= 0 && nbase + rd * 16 < N) {
                    if constexpr (RES == RES_LOAD) dst[i] = *reinterpret_cast<const float4*>(ep.residual + orow[i] * ep.ldr + nbase + rd * 16);
                    if constexpr (MODE == CS_EPI_TOKENS) dst[i] = *reinterpret_cast<const float4*>(ep.pos_embed + (long long)prow[i] * N + nbase + rd * 16);
                }
            }
        };
        if constexpr (HAS_EXTRA) {
#pragma unroll
            for (int p = 0; p < PF && p < ROUNDS; ++p) fetch_extra(p, extra[p]);
        }
        wait_accumulator();
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};      // EMIT: statistics of this lane's 4 rows
        const uint32_t st_s = smem_u32(st);
#pragma unroll
        for (int rd = 0; rd < ROUNDS; ++rd) {
            const int r = rd * 16;
            const int nc = nbase + r;
            const bool col_ok = n0 + half * HALF_N + r < N;          // warp-uniform (N % 32 == 0)
            // EMIT: staging buffer rd & 1 (f32 tile at 2048 * b, bf16 tile at 4096 + 1024 * b); the bulk stores of round
            // rd - 2 must have read it
            uint8_t* stf = EMIT ? st + (rd & 1) * 2048 : st;
            uint8_t* sth = st + 4096 + (rd & 1) * 1024;
            if constexpr (EMIT) {
                if (rd >= 2) {
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncwarp();
                }
            }
            float4 ex[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) ex[i] = HAS_EXTRA ? extra[rd % PF][i] : make_float4(0.f, 0.f, 0.f, 0.f);
            if constexpr (HAS_EXTRA) {
                if (rd + PF < ROUNDS) fetch_extra(rd + PF, extra[rd % PF]);
            }
            if (col_ok) {
                uint32_t a[16];
                tmem_ld16(taddr + half * HALF_N + r, a);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float4 f;
                    if constexpr (LNFOLD) {
                        const float4 c4 = *reinterpret_cast<const float4*>(sc1 + r + 4 * j);     // one 16-byte broadcast read
                        f.x = fmaf(rs, __uint_as_float(a[4 * j]), -rm * c4.x);
                        f.y = fmaf(rs, __uint_as_float(a[4 * j + 1]), -rm * c4.y);
                        f.z = fmaf(rs, __uint_as_float(a[4 * j + 2]), -rm * c4.z);
                        f.w = fmaf(rs, __uint_as_float(a[4 * j + 3]), -rm * c4.w);
                    } else {
                        f = make_float4(__uint_as_float(a[4 * j]) * ep.alpha, __uint_as_float(a[4 * j + 1]) * ep.alpha,
                                        __uint_as_float(a[4 * j + 2]) * ep.alpha, __uint_as_float(a[4 * j + 3]) * ep.alpha);
                    }
                    *reinterpret_cast<float4*>(stf + lane * 64 + ((j ^ swz) << 4)) = f;
                }
            }
            __syncwarp();
            if (col_ok) {
                const float4 b4 = *reinterpret_cast<const float4*>(sbias + r + cch * 4);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rr = i * 8 + crow;
                    float4* slot = reinterpret_cast<float4*>(stf + rr * 64 + ((cch ^ ((rr >> 1) & 3)) << 4));
                    float4 f = *slot;
                    f.x += b4.x + ex[i].x;
                    f.y += b4.y + ex[i].y;
                    f.z += b4.z + ex[i].z;
                    f.w += b4.w + ex[i].w;
                    if constexpr (EMIT) {
                        // final values back into the (SWIZZLE_64B) f32 tile, their bf16 copy into the dense 32-byte-row tile;
                        // rows beyond M are clipped by the store maps
                        *slot = f;
                        uint2 pk;
                        pk.x = pack_bf16(f.x, f.y);
                        pk.y = pack_bf16(f.z, f.w);
                        *reinterpret_cast<uint2*>(sth + rr * 32 + cch * 8) = pk;
                        if (orow[i] >= 0) {
                            s1[i] += (f.x + f.y) + (f.z + f.w);
                            s2[i] = fmaf(f.x, f.x, fmaf(f.y, f.y, fmaf(f.z, f.z, fmaf(f.w, f.w, s2[i]))));
                        }
                    } else if (orow[i] >= 0 && !(ep.dbg & 1)) {
                        float* dst = reinterpret_cast<float*>(ep.out) + orow[i] * ep.ldo + nc;
                        if constexpr (RES == RES_RED) {
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(f.x), "f"(f.y), "f"(f.z), "f"(f.w)
                                         : "memory");
                        } else {
                            *reinterpret_cast<float4*>(dst) = f;
                        }
                    }
                }
            }
            if constexpr (EMIT) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // staging writes -> visible to the bulk stores
                __syncwarp();
                if (lane == 0 && col_ok && !(ep.dbg & 1)) {
                    const int c0 = n0 + half * HALF_N + r;
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map_o32),
                                 "r"(st_s + (uint32_t)((rd & 1) * 2048)), "r"(c0), "r"(mw)
                                 : "memory");
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map_o16),
                                 "r"(st_s + (uint32_t)(4096 + (rd & 1) * 1024)), "r"(c0), "r"(mw)
                                 : "memory");
                }
                if (lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            } else {
                __syncwarp();
            }
        }
        if constexpr (EMIT) {       // the staging tiles are rewritten by the next tile's first rounds
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
        }
        if constexpr (EMIT) {
            // the 4 lanes of a row hold its partial sums: combine, lane cch == 0 writes (row, tile, half)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], 1);
                s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], 1);
                s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], 2);
                s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], 2);
                if (cch == 0 && orow[i] >= 0 && ep.stats_out != nullptr)
                    *reinterpret_cast<float2*>(ep.stats_out + ((orow[i] * ((N + BLOCK_N - 1) / BLOCK_N) + n0 / BLOCK_N) * 2 + half) * 2) =
                        make_float2(s1[i], s2[i]);
            }
        }
    }
}

// TN = true: both operands are consumed MN-major, i.e. A is stored [K][M] and W is stored [K][N] (the weight-gradient GEMM
// dW = dY^T X reads dY [tokens][out] and X [tokens][in] as they are: no transposed copies).  A stage then holds 64-wide
// atoms [64 k rows][64 m|n] of 8 KB each (UMMA canonical MN-major SWIZZLE_128B layout: LBO = atom pitch, SBO = 8 k rows).
template <int BLOCK_N, int MODE, bool OUT_BF16, int RES, bool LNFOLD, bool EMIT, int CG, bool TN = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
            const __grid_constant__ CUtensorMap map_o32, const __grid_constant__ CUtensorMap map_o16,
            int M, int N, int K, const EpiParams ep) {
    using C = Cfg<BLOCK_N, CG, EMIT>;
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
    static_assert(8 * (2 * C::STAGES + 5) <= C::BAR_BYTES, "barrier area");
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + BAR_OFF + 8 * (2 * C::STAGES + 4));
    float* srope = reinterpret_cast<float*>(smem + ROPE_OFF);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // work unit: one (CG*128) x BLOCK_N tile per CTA group; `rank` = this CTA's half of it
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const int group = (int)blockIdx.x / CG, num_groups = (int)gridDim.x / CG;
    const int num_m_tiles = (M + CG * BLOCK_M - 1) / (CG * BLOCK_M);
    const int num_n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
    const int num_mn = num_m_tiles * num_n_tiles;
    const int num_tiles = num_mn * ep.k_splits;          // split-K: tile t -> (mn = t % num_mn, split = t / num_mn)
    const int num_kb_total = (K + BLOCK_K - 1) / BLOCK_K;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        if constexpr (EMIT) tma_prefetch_desc(&map_o32);
        if constexpr (EMIT || OUT_BF16) tma_prefetch_desc(&map_o16);
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), CG * EPI_WARPS * 32);     // CG = 2: the leader's barrier collects both CTAs' epilogues
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if constexpr (MODE == CS_EPI_QKV_ROPE) {
        // rotary table of this launch: angle(g, q) = pos[g] * freq[q] (f32 product as in rope.py:129), accurate
        // sincosf once per CTA; layout [q][g], g = 64 is the identity rotation used by the CLS rows
        for (int i = threadIdx.x; i < 16 * ROPE_STRIDE; i += NUM_THREADS) {
            const int q = i / ROPE_STRIDE, g = i % ROPE_STRIDE;
            float sn = 0.f, cs_ = 1.f;
            if (g < ep.rope_grid) sincosf(ep.rope_pos[g] * ep.rope_freq[q], &sn, &cs_);
            srope[i] = cs_;
            srope[16 * ROPE_STRIDE + i] = sn;
        }
    }
    if (warp == 1) {
        if constexpr (CG == 2) tmem_alloc_cg2(tmem_slot, C::TMEM_COLS);       // one warp of EACH CTA of the pair
        else tmem_alloc(tmem_slot, C::TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();      // the peer's barriers are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        // CG = 2: both CTAs load their own A rows and their half of the B rows into their own shared memory; all bytes
        // are counted on the LEADER's full barrier (the only MMA issuer), which expects both CTAs' stage bytes.
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = group; tile < num_tiles; tile += num_groups) {
                const int mn = tile % num_mn, sp = tile / num_mn;
                const int m0 = (mn / num_n_tiles) * (CG * BLOCK_M) + (int)rank * BLOCK_M;
                const int n0 = (mn % num_n_tiles) * BLOCK_N + (int)rank * C::B_ROWS * (CG - 1);
                const int kb0 = sp * ep.kb_per_split;
                const int kb1 = min(num_kb_total, kb0 + ep.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = base + stage * C::STAGE_BYTES;
                    const uint32_t sb = sa + C::A_BYTES;
                    if constexpr (CG == 2) {
                        const uint32_t fb = map_to_cta(full_bar(stage), 0);
                        if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);
                        if constexpr (TN) {
#pragma unroll
                            for (int a = 0; a < BLOCK_M / 64; ++a) tma_load_2d_cg2(sa + a * 8192, &map_a, fb, m0 + a * 64, kb * BLOCK_K);
#pragma unroll
                            for (int b = 0; b < C::B_ROWS / 64; ++b) tma_load_2d_cg2(sb + b * 8192, &map_b, fb, n0 + b * 64, kb * BLOCK_K);
                        } else {
                            tma_load_2d_cg2(sa, &map_a, fb, kb * BLOCK_K, m0);
                            tma_load_2d_cg2(sb, &map_b, fb, kb * BLOCK_K, n0);
                        }
                    } else {
                        mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
                        if constexpr (TN) {
#pragma unroll
                            for (int a = 0; a < BLOCK_M / 64; ++a) tma_load_2d(sa + a * 8192, &map_a, full_bar(stage), m0 + a * 64, kb * BLOCK_K);
#pragma unroll
                            for (int b = 0; b < C::B_ROWS / 64; ++b) tma_load_2d(sb + b * 8192, &map_b, full_bar(stage), n0 + b * 64, kb * BLOCK_K);
                        } else {
                            tma_load_2d(sa, &map_a, full_bar(stage), kb * BLOCK_K, m0);
                            tma_load_2d(sb, &map_b, full_bar(stage), kb * BLOCK_K, n0);
                        }
                    }
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer (CG = 2: leader CTA only) --------
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = make_idesc(CG * BLOCK_M, BLOCK_N) | (TN ? ((1u << 15) | (1u << 16)) : 0u);   // bits 15 / 16: A / B MN-major
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = group; tile < num_tiles; tile += num_groups, ++it) {
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
                    // K-major: SBO = 1024 (8 rows), a k-step advances 16 bf16 = 32 B inside the swizzle atom: +2 in (addr >> 4).
                    // MN-major (TN): LBO = 8192 (next 64-wide m|n atom), SBO = 1024 (next 8 k rows), a k-step = 16 k rows = 2048 B
                    const uint64_t da = TN ? (make_smem_desc(sa) | ((uint64_t)(8192u >> 4) << 16)) : make_smem_desc(sa);
                    const uint64_t db = TN ? (make_smem_desc(sb) | ((uint64_t)(8192u >> 4) << 16)) : make_smem_desc(sb);
                    constexpr uint64_t kstep = TN ? (2048u >> 4) : 2u;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        if constexpr (CG == 2)
                            umma_bf16_cg2(tmem_d, da + kstep * k, db + kstep * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        else
                            umma_bf16(tmem_d, da + kstep * k, db + kstep * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    if constexpr (CG == 2) {
                        umma_commit_cg2(empty_bar(stage), 3);      // frees the slot in BOTH CTAs when the MMAs retire
                        if (kb == kb1 - 1) umma_commit_cg2(tfull_bar(acc), 3);
                    } else {
                        umma_commit(empty_bar(stage));             // frees the smem slot when MMAs retire
                        if (kb == kb1 - 1) umma_commit(tfull_bar(acc));
                    }
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else {
        // ------------------------------ epilogue -----------------------------------
        // warps 2..9: TMEM lane quarter = warp % 4 (hardware rule), column half = (warp - 2) / 4
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        uint8_t* st = smem + EPI_OFF + (warp - 2) * C::EPI_WARP_BYTES;
        float* sbias = reinterpret_cast<float*>(st + C::EPI_TILE_BYTES);
        const uint32_t tempty_leader0 = (CG == 2) ? map_to_cta(tempty_bar(0), 0) : 0u;
        int it = 0;
        for (int tile = group; tile < num_tiles; tile += num_groups, ++it) {
            const int mn = tile % num_mn;
            const int m0 = (mn / num_n_tiles) * (CG * BLOCK_M) + (int)rank * BLOCK_M;
            const int n0 = (mn % num_n_tiles) * BLOCK_N;
            const int acc = it & 1;
            const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            epilogue_tile<BLOCK_N, MODE, OUT_BF16, RES, LNFOLD, EMIT>(ep, taddr, m0 + quarter * 32, n0, half, M, N, st, sbias,
                                                                      srope, lane, tile < num_mn, tfull_bar(acc), acc_phase, &map_o32, &map_o16);
            tc_fence_before();
            if constexpr (CG == 2) mbar_arrive_cluster(tempty_leader0 + 8u * acc);
            else mbar_arrive(tempty_bar(acc));
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // staged tiles are read before the CTA retires
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();      // nobody leaves while the peer may still signal into its shared memory
    if (warp == 1) {
        if constexpr (CG == 2) tmem_dealloc_cg2(tmem_base, C::TMEM_COLS);
        else tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ----------------------------------------------------------------------------------------
// Host side
// ----------------------------------------------------------------------------------------
// BLOCK_N = 256 runs on CTA pairs (cta_group::2) unless CS_GEMM_1CTA=1 (A/B measurements against the single-CTA kernel)
static bool use_cta_pairs() {
    static const int v = [] {
        const char* e = getenv("CS_GEMM_1CTA");
        return (e != nullptr && e[0] != '\0' && e[0] != '0') ? 0 : 1;
    }();
    return v != 0;
}

template <int BLOCK_N, int MODE, bool OUT_BF16, int RES, bool LNFOLD, bool EMIT, int CG, bool TN = false>
static int launch_cg(const CUtensorMap& ma, const CUtensorMap& mb, int M, int N, int K, const EpiParams& ep,
                     cudaStream_t stream) {
    using C = Cfg<BLOCK_N, CG, EMIT>;
    // EMIT: store maps of the two outputs, one 32-row x 16-column box per epilogue round (f32: 64-byte rows in the
    // SWIZZLE_64B pattern of the staging tile, bf16: dense 32-byte rows); cached like the operand maps
    CUtensorMap mo32 = ma, mo16 = ma;
    if constexpr (EMIT) {
        int rc = make_map_2d(&mo32, ep.out, 4, M, N, ep.ldo, 16, 32, 64);
        if (rc) return rc;
        rc = make_map_2d(&mo16, ep.out2, 2, M, N, ep.ldo2, 16, 32, 0);
        if (rc) return rc;
    }
    if constexpr (OUT_BF16) {   // bf16 outputs: one 32-row x 32-column box (64-byte rows, SWIZZLE_64B) per epilogue round
        const int rc = make_map_2d(&mo16, ep.out, 2, M, MODE == CS_EPI_SWIGLU ? N / 2 : N, ep.ldo, 32, 32, 64);
        if (rc) return rc;
    }
    auto kern = gemm_kernel<BLOCK_N, MODE, OUT_BF16, RES, LNFOLD, EMIT, CG, TN>;
    static bool configured = false;
    if (!configured) {
        CS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        configured = true;
    }
    const long long tiles = (long long)ceil_div(M, CG * BLOCK_M) * ceil_div(N, BLOCK_N) * ep.k_splits;   // split-K tiles included
    const int max_groups = num_sms() / CG;
    const int groups = tiles < max_groups ? (int)tiles : max_groups;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(groups * CG));
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CS_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, mo32, mo16, M, N, K, ep));
    return CS_OK;
}

template <int BLOCK_N, int MODE, bool OUT_BF16, int RES, bool LNFOLD = false, bool EMIT = false>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb2, const CUtensorMap& mb1, int M, int N, int K, const EpiParams& ep,
                  cudaStream_t stream) {
    if constexpr (BLOCK_N == 256) {
        if (use_cta_pairs()) return launch_cg<BLOCK_N, MODE, OUT_BF16, RES, LNFOLD, EMIT, 2>(ma, mb2, M, N, K, ep, stream);
    }
    return launch_cg<BLOCK_N, MODE, OUT_BF16, RES, LNFOLD, EMIT, 1>(ma, mb1, M, N, K, ep, stream);
}

// weight-gradient form (both operands MN-major): plain f32 store or split-K red.add
template <int BLOCK_N>
static int dispatch_tn(const CUtensorMap& ma, const CUtensorMap& mb, int M, int N, int K, const EpiParams& ep, int res, cudaStream_t st) {
    constexpr int CG = BLOCK_N == 256 ? 2 : 1;
    if (res == RES_RED) return launch_cg<BLOCK_N, CS_EPI_STORE, false, RES_RED, false, false, CG, true>(ma, mb, M, N, K, ep, st);
    return launch_cg<BLOCK_N, CS_EPI_STORE, false, RES_NONE, false, false, CG, true>(ma, mb, M, N, K, ep, st);
}

// mb2: B map with the box of a CTA pair member (BLOCK_N / 2 rows); mb: box of BLOCK_N rows (single-CTA kernel)
template <int BLOCK_N>
static int dispatch(const CUtensorMap& ma, const CUtensorMap& mb2, const CUtensorMap& mb, int M, int N, int K, const EpiParams& ep,
                    int res, cudaStream_t st) {
    const bool fold = ep.ln_stats != nullptr, emit = ep.out2 != nullptr;
    switch (ep.mode) {
        case CS_EPI_STORE:
            if (ep.out_bf16) {
                if (fold) return launch<BLOCK_N, CS_EPI_STORE, true, RES_NONE, true>(ma, mb2, mb, M, N, K, ep, st);
                return launch<BLOCK_N, CS_EPI_STORE, true, RES_NONE>(ma, mb2, mb, M, N, K, ep, st);
            }
            if (emit) {         // new f32 residual stream + its bf16 copy + row statistics
                if (res == RES_LOAD && fold) return launch<BLOCK_N, CS_EPI_STORE, false, RES_LOAD, true, true>(ma, mb2, mb, M, N, K, ep, st);
                if (res == RES_LOAD) return launch<BLOCK_N, CS_EPI_STORE, false, RES_LOAD, false, true>(ma, mb2, mb, M, N, K, ep, st);
                break;
            }
            if (res == RES_RED && fold) return launch<BLOCK_N, CS_EPI_STORE, false, RES_RED, true>(ma, mb2, mb, M, N, K, ep, st);
            if (res == RES_LOAD && fold) return launch<BLOCK_N, CS_EPI_STORE, false, RES_LOAD, true>(ma, mb2, mb, M, N, K, ep, st);
            if (fold) break;
            if (res == RES_RED) return launch<BLOCK_N, CS_EPI_STORE, false, RES_RED>(ma, mb2, mb, M, N, K, ep, st);
            if (res == RES_LOAD) return launch<BLOCK_N, CS_EPI_STORE, false, RES_LOAD>(ma, mb2, mb, M, N, K, ep, st);
            return launch<BLOCK_N, CS_EPI_STORE, false, RES_NONE>(ma, mb2, mb, M, N, K, ep, st);
        case CS_EPI_QKV_ROPE:
            if (fold) return launch<BLOCK_N, CS_EPI_QKV_ROPE, true, RES_NONE, true>(ma, mb2, mb, M, N, K, ep, st);
            return launch<BLOCK_N, CS_EPI_QKV_ROPE, true, RES_NONE>(ma, mb2, mb, M, N, K, ep, st);
        case CS_EPI_TOKENS:
            return launch<BLOCK_N, CS_EPI_TOKENS, false, RES_NONE>(ma, mb2, mb, M, N, K, ep, st);
        case CS_EPI_SWIGLU:
            if constexpr (BLOCK_N == 256) {
                if (fold) return launch<256, CS_EPI_SWIGLU, true, RES_NONE, true>(ma, mb2, mb, M, N, K, ep, st);
                return launch<256, CS_EPI_SWIGLU, true, RES_NONE>(ma, mb2, mb, M, N, K, ep, st);
            }
    }
    set_error("cs_gemm_bf16: unsupported epilogue combination");
    return CS_ERR_UNSUPPORTED;
}

}  // namespace gemm
}  // namespace cs

// split-K factor for a plain f32 STORE: the one that minimises waves x (k-blocks per split + per-tile overhead)
static void choose_split_k(cs::gemm::EpiParams& ep, int& res, int M, int N, int K, bool use256, int cg, int requested) {
    using namespace cs;
    using namespace cs::gemm;
    const int num_kb = ceil_div(K, BLOCK_K);
    const int tiles = ceil_div(M, cg * BLOCK_M) * ceil_div(N, use256 ? 256 : 128);
    const int groups = num_sms() / cg;
    int best = 1;
    if (requested > 0) {
        best = requested;
    } else {
        long long best_cost = -1;
        for (int sp = 1; sp <= 16 && sp <= num_kb; ++sp) {
            const long long waves = ceil_div((long long)tiles * sp, groups);
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
    ep.out2 = reinterpret_cast<__nv_bfloat16*>(e->out2_bf16);
    ep.ldo2 = e->ldo2;
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

    CUtensorMap ma, mb, mb2;
    int rc = make_map_bf16_2d(&ma, A, M, K, lda, BLOCK_K, BLOCK_M);
    if (rc) return rc;
    rc = make_map_bf16_2d(&mb, W, N, K, ldw, BLOCK_K, use256 ? 256 : 128);
    if (rc) return rc;
    mb2 = mb;
    if (use256 && use_cta_pairs()) {
        rc = make_map_bf16_2d(&mb2, W, N, K, ldw, BLOCK_K, 128);       // each CTA of a pair stages 128 of the tile's 256 W rows
        if (rc) return rc;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int res = RES_NONE;
    if (e->residual) {
        CS_CHECK_ARG(e->out_dtype == CS_F32 && e->mode == CS_EPI_STORE, "cs_gemm_bf16: residual needs f32 STORE output");
        res = (e->residual == e->out && e->ldr == e->ldo && e->out2_bf16 == nullptr) ? RES_RED : RES_LOAD;
    }
    if (e->ln_stats)
        CS_CHECK_ARG(e->ln_c1 && e->ln_parts > 0 && e->ln_parts % 2 == 0 && e->ln_dim > 0 && e->alpha == 1.0f &&
                         e->mode != CS_EPI_TOKENS && ((uintptr_t)e->ln_stats % 16 == 0) && ((uintptr_t)e->ln_c1 % 16 == 0),
                     "cs_gemm_bf16: LN folding needs ln_c1, even ln_parts, ln_dim, alpha == 1 (any mode but TOKENS)");
    if (e->out2_bf16)
        CS_CHECK_ARG(e->mode == CS_EPI_STORE && e->out_dtype == CS_F32 && e->residual && ((uintptr_t)e->out2_bf16 % 16 == 0) &&
                         e->ldo2 % 8 == 0,
                     "cs_gemm_bf16: out2_bf16 (bf16 copy + row statistics of the new residual stream) needs the f32 STORE "
                     "epilogue with a residual");
    if (e->stats_out) CS_CHECK_ARG(e->mode == CS_EPI_SWIGLU || e->out2_bf16, "cs_gemm_bf16: stats_out is a SWIGLU / out2_bf16 output");
    if (e->reserved2 != 0) {
        // split-K (reserved2 = requested splits, -1 = choose): partial products are accumulated with
        // red.add into `out`, which the caller must have zeroed.  f32 STORE without residual only.
        CS_CHECK_ARG(e->mode == CS_EPI_STORE && e->out_dtype == CS_F32 && e->residual == nullptr && e->ln_stats == nullptr,
                     "cs_gemm_bf16: split-K needs a plain f32 STORE epilogue");
        choose_split_k(ep, res, (int)M, N, K, use256, (use256 && use_cta_pairs()) ? 2 : 1, e->reserved2);
    }
    return use256 ? dispatch<256>(ma, mb2, mb, (int)M, N, K, ep, res, st) : dispatch<128>(ma, mb2, mb, (int)M, N, K, ep, res, st);
}

// C[M,N] (f32) = At^T . Bt  with At stored [K][M] (row stride lda) and Bt stored [K][N] (row stride ldb), both bf16: the
// weight-gradient GEMM dW = dY^T X on dY [tokens][out] and X [tokens][in] as they lie in memory (train.py:104 backward of
// every nn.Linear of the student, torch's addmm with transposed operands).  Epilogue: plain f32 STORE, optional alpha is
// not applied; reserved2 selects split-K as in cs_gemm_bf16 (out must then be zeroed by the caller).
extern "C" int cs_gemm_bf16_tn(const void* At, int64_t lda, const void* Bt, int64_t ldb, int64_t M, int N, int K,
                               const cs_gemm_epilogue_t* e, void* stream) {
    using namespace cs;
    using namespace cs::gemm;
    CS_CHECK_ARG(At && Bt && e && e->out, "cs_gemm_bf16_tn: null pointer");
    CS_CHECK_ARG(M > 0 && N > 0 && K > 0 && M < (1ll << 31), "cs_gemm_bf16_tn: bad shape M=%lld N=%d K=%d", (long long)M, N, K);
    CS_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && lda >= M && ldb >= N, "cs_gemm_bf16_tn: lda/ldb must be multiples of 8 and >= M / N");
    CS_CHECK_ARG(((uintptr_t)At % 16 == 0) && ((uintptr_t)Bt % 16 == 0), "cs_gemm_bf16_tn: operands must be 16 B aligned");
    CS_CHECK_ARG(N % 32 == 0, "cs_gemm_bf16_tn: N must be a multiple of 32 (N=%d)", N);
    CS_CHECK_ARG(e->mode == CS_EPI_STORE && e->out_dtype == CS_F32 && e->residual == nullptr && e->bias == nullptr &&
                     e->ln_stats == nullptr && e->out2_bf16 == nullptr && e->alpha == 1.0f,
                 "cs_gemm_bf16_tn: plain f32 STORE epilogue only");
    CS_CHECK_ARG(((uintptr_t)e->out % 16 == 0) && e->ldo % 4 == 0, "cs_gemm_bf16_tn: out must be 16 B aligned with 16 B aligned rows");
    EpiParams ep = {};
    ep.mode = CS_EPI_STORE;
    ep.out = e->out;
    ep.ldo = e->ldo;
    ep.alpha = 1.0f;
    ep.dbg = e->reserved;
    ep.k_splits = 1;
    ep.kb_per_split = ceil_div(K, BLOCK_K);
    const bool use256 = (N % 256 == 0);
    CUtensorMap ma, mb;
    int rc = make_map_bf16_2d(&ma, At, K, M, lda, 64, 64);        // box = [64 k rows][64 m]: one MN-major atom
    if (rc) return rc;
    rc = make_map_bf16_2d(&mb, Bt, K, N, ldb, 64, 64);
    if (rc) return rc;
    int res = RES_NONE;
    if (e->reserved2 != 0) choose_split_k(ep, res, (int)M, N, K, use256, use256 ? 2 : 1, e->reserved2);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return use256 ? dispatch_tn<256>(ma, mb, (int)M, N, K, ep, res, st) : dispatch_tn<128>(ma, mb, (int)M, N, K, ep, res, st);
}
