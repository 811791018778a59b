// tcgen05 softmax attention forward for head_dim 64 and 129 <= N <= 208 tokens (EVA02-B/16: N = 197), fourth generation:
// the two query tiles of one (image, head) are two SLOTS that ping-pong on the tensor core and on the MUFU pipe.
//   S_s = Q_s K^T (tcgen05.mma -> TMEM)  ->  slot s: row max, P = exp2(.) as bf16 into swizzled shared memory  ->
//   O = P_s V (V consumed MN-major from its [key][dim] layout); row sums of the bf16-rounded P in registers  ->
//   O / rowsum -> bf16 rows staged in shared memory -> one TMA store per warp (full 128-byte lines).
// Replaces xformers.memory_efficient_attention at eva_vit_model.py:206-217 for the teacher's crops and the student.
//
// Why a fourth kernel (profiles/r02_attention_experiments.txt): attention_tc3 and its two siblings all ran at 5x their
// MUFU bound because every handshake was a 512-thread (or 128-thread) mbarrier arrival on one shared-memory word and the
// exp pass, the P V round trip and the epilogue of a tile ran back to back.  Here
//   * every thread owns one full query row (no cross-warp exchange, no named barrier), the tcgen05.ld of chunk c+1 is in
//     flight while chunk c is processed;
//   * handshakes are ONE elected arrival per warp (fence -> __syncwarp -> lane 0 arrives): barrier counts are 4, not 512;
//   * slot 0 (rows 0..127) and slot 1 (rows 128..N-1) are separate warp groups with their own S accumulator and P buffer,
//     so the exp pass of one slot runs under the S / P V MMAs, the TMEM read-back and the stores of the other;
//   * slot 1 only stages the rows it has (P atoms at a rows1 x 128 B pitch, the unused tensor-core rows read whatever
//     follows), which is what lets two P buffers, Q, K and a double-buffered V fit in 227 KB;
//   * warps whose 32 rows are all beyond N do nothing but the handshakes;
//   * the first S MMA of slot 1 is held back until slot 0 finished its first exp pass, so the slots start half a period
//     apart (measured: started together they stay in lockstep and nothing overlaps);
//   * the output leaves through shared memory and cp.async.bulk.tensor stores: 16-byte pieces scattered at a 1536-byte
//     stride cost 2 us per head as direct stores (32 L2 transactions per instruction), more than the exp pass.
//
// Persistent CTA per SM, 384 threads = three warpgroups: warps 0-3 slot 0, warps 4-7 slot 1, warp 8 TMA producer, warp 9 MMA
// issuer + TMEM allocator (warps 10-11 idle); setmaxnreg moves the registers of the third group to the softmax groups.  TMEM columns: S0 [0,208) S1 [208,416) O [416,480).
#include "tc_common.cuh"

namespace cs {
namespace attn_tc4 {
using namespace cs::tc;

constexpr int HD = 64;
constexpr int BM = 128;
constexpr int THREADS = 384;
constexpr int TMA_WARP = 8, MMA_WARP = 9;
constexpr int MIN_N = 129, MAX_NKP = 208;
constexpr int S_STRIDE = 208;
constexpr int O_COL = 416;
constexpr int TMEM_COLS = 512;
constexpr int Q0_BYTES = BM * 128;
constexpr int Q1_BYTES = 80 * 128;          // slot 1 has at most 208 - 128 = 80 rows
constexpr int ATOM_BYTES = BM * 128;        // [128 rows][64 keys] bf16, SWIZZLE_128B K-major

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// bounded wait without the diagnostic printf of tc::mbar_wait: ~25 wait sites, each inlined, were a third of the kernel's code
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();        // a protocol bug surfaces as a CUDA error, never as a hung GPU
    }
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float y;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
    return y;
}
// bulk tensor store shared -> global of one [32 rows][64 dims] sub-tile; coordinates (dim, token, image): rows >= N are clipped
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(smem_src), "r"(c0),
                 "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c_inner, int c_outer) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c_inner), "r"(c_outer) : "memory");
}

// CS_ATTN_DBG bit 16: CTA 0 records clock64() at the phase boundaries of its first items (tools/attn_timeline.py prints them)
constexpr int TL_ROWS = 12, TL_COLS = 256;
__device__ unsigned long long g_timeline[TL_ROWS][TL_COLS];
#define TL(rowi, idx)                                                                                  \
    do {                                                                                               \
        if ((p.dbg & 16) && blockIdx.x == 0 && (idx) < TL_COLS) g_timeline[rowi][idx] = clock64();     \
    } while (0)

struct Params {
    int B, N, H, nkp;               // nkp: keys padded to 16
    int rows1, pitch1;              // slot 1: rows staged (N - 128 rounded up to 8) and its P atom pitch in bytes
    float scale_log2, scale;
    int dbg;                        // timing experiments (CS_ATTN_DBG): 1 = no P V MMAs, 2 = no exp pass, 4 = no stores, 8 = no max pass
    __nv_bfloat16* out;
    float* lse;
    float* row_stats;               // optional [B*N, 4H, 2]: per (row, head, 16-dim quarter) sum and sum of squares of the f32 output
};

__global__ void __launch_bounds__(THREADS, 1)
attention_fwd_tc4_kernel(const __grid_constant__ CUtensorMap map_q0, const __grid_constant__ CUtensorMap map_q1,
                         const __grid_constant__ CUtensorMap map_kv, const __grid_constant__ CUtensorMap map_out,
                         const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw_addr);

    const int kv_bytes = p.nkp * 128;
    const int n_atoms = (p.nkp + 63) / 64;
    const uint32_t sQ0 = base;
    const uint32_t sQ1 = sQ0 + Q0_BYTES;
    const uint32_t sK = sQ1 + Q1_BYTES;                         // 1 stage (free again as soon as both S MMAs of the item ran)
    const uint32_t sV = sK + kv_bytes;                          // 2 stages
    const uint32_t sP1 = sV + 2 * kv_bytes;                     // slot 1: n_atoms atoms at pitch1 (tensor-core over-read lands in P0)
    const uint32_t sP0 = sP1 + n_atoms * p.pitch1;              // slot 0: n_atoms full atoms
    const uint32_t bar = sP0 + n_atoms * ATOM_BYTES;
    auto q_full = [&](int s) { return bar + 8u * s; };
    auto q_empty = [&](int s) { return bar + 8u * (2 + s); };
    const uint32_t k_full = bar + 8u * 4, k_empty = bar + 8u * 5;
    auto v_full = [&](int s) { return bar + 8u * (6 + s); };
    auto v_empty = [&](int s) { return bar + 8u * (8 + s); };
    auto s_full = [&](int s) { return bar + 8u * (10 + s); };
    auto s_empty = [&](int s) { return bar + 8u * (12 + s); };
    auto p_full = [&](int s) { return bar + 8u * (14 + s); };
    auto o_full = [&](int s) { return bar + 8u * (16 + s); };
    auto o_empty = [&](int s) { return bar + 8u * (18 + s); };
    const uint32_t tmem_slot = bar + 8u * 20;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = p.H * HD;
    const int n_items = p.B * p.H;
    const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_q0);
        tma_prefetch_desc(&map_q1);
        tma_prefetch_desc(&map_kv);
        tma_prefetch_desc(&map_out);
        for (int s = 0; s < 2; ++s) {
            mbar_init(q_full(s), 1);
            mbar_init(q_empty(s), 1);
            mbar_init(v_full(s), 1);
            mbar_init(v_empty(s), 1);
            mbar_init(s_full(s), 1);
            mbar_init(s_empty(s), 4);
            mbar_init(p_full(s), 4);
            mbar_init(o_full(s), 1);
            mbar_init(o_empty(s), 4);
        }
        mbar_init(k_full, 1);
        mbar_init(k_empty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
      if (warp == TMA_WARP) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            for (int il = 0; il < my_items; ++il) {
                const int item = blockIdx.x + il * gridDim.x;
                const int b = item / p.H, h = item % p.H;
                const uint32_t ph = (uint32_t)il & 1u;
                wait_bar(k_empty, ph ^ 1u);
                mbar_arrive_expect_tx(k_full, (uint32_t)kv_bytes);
                tma_load_2d(sK, &map_kv, k_full, D + h * HD, b * p.N);
                TL(8, 4 * il);
                wait_bar(q_empty(0), ph ^ 1u);
                mbar_arrive_expect_tx(q_full(0), Q0_BYTES);
                tma_load_2d(sQ0, &map_q0, q_full(0), h * HD, b * p.N);
                TL(8, 4 * il + 1);
                wait_bar(q_empty(1), ph ^ 1u);
                mbar_arrive_expect_tx(q_full(1), (uint32_t)(p.rows1 * 128));
                tma_load_2d(sQ1, &map_q1, q_full(1), h * HD, b * p.N + BM);
                TL(8, 4 * il + 2);
                const int st = il & 1;
                wait_bar(v_empty(st), (uint32_t)((il >> 1) & 1) ^ 1u);
                mbar_arrive_expect_tx(v_full(st), (uint32_t)kv_bytes);
                tma_load_2d(sV + st * kv_bytes, &map_kv, v_full(st), 2 * D + h * HD, b * p.N);
                TL(8, 4 * il + 3);
                if (il + 1 < my_items) {        // K and Q are single buffered: have the next item's tiles waiting in L2
                    const int nitem = item + gridDim.x;
                    const int nb = nitem / p.H, nh = nitem % p.H;
                    tma_prefetch_l2_2d(&map_kv, D + nh * HD, nb * p.N);
                    tma_prefetch_l2_2d(&map_q0, nh * HD, nb * p.N);
                    tma_prefetch_l2_2d(&map_q1, nh * HD, nb * p.N + BM);
                }
            }
        }
      } else if (warp == MMA_WARP) {
        // ------------------------------ MMA issuer --------------------------------
        if (lane == 0 && my_items > 0) {
            // S: M=128, N=nkp, A/B K-major.  PV: M=128, N=64, A K-major (P), B MN-major (V) -> bit 16
            const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.nkp >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(HD >> 3) << 17) |
                                      ((uint32_t)(BM >> 4) << 24);
            const uint64_t dk = smem_desc(sK, 0, 1024);
            const int ksteps = p.nkp / 16;
            auto issue_s = [&](int s, int il) {
                const uint32_t ph = (uint32_t)il & 1u;
                TL(9, 8 * il + 2 * s);
                wait_bar(q_full(s), ph);
                if (s == 0) wait_bar(k_full, ph);
                wait_bar(s_empty(s), ph ^ 1u);                       // the slot has read S of its previous item
                tc_fence_after();
                TL(9, 8 * il + 2 * s + 1);
                const uint64_t dq = smem_desc(s ? sQ1 : sQ0, 0, 1024);
                const uint32_t d_tmem = tmem_base + (uint32_t)(s * S_STRIDE);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) umma_bf16(d_tmem, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, k > 0);
                umma_commit(s_full(s));
                umma_commit(q_empty(s));
                if (s == 1) umma_commit(k_empty);
            };
            auto issue_pv = [&](int s, int il) {
                const uint32_t ph = (uint32_t)il & 1u;
                const int st = il & 1;
                TL(9, 8 * il + 4 + 2 * s);
                wait_bar(p_full(s), ph);
                if (s == 0) {
                    wait_bar(v_full(st), (uint32_t)(il >> 1) & 1u);
                    wait_bar(o_empty(1), ph ^ 1u);                   // O of slot 1's previous item has been read
                } else {
                    wait_bar(o_empty(0), ph);                        // O of slot 0's tile of this item has been read
                }
                tc_fence_after();
                TL(9, 8 * il + 5 + 2 * s);
                const uint32_t sv = sV + st * kv_bytes;
                const uint32_t sp = s ? sP1 : sP0;
                const uint32_t pitch = s ? (uint32_t)p.pitch1 : (uint32_t)ATOM_BYTES;
                for (int kk = 0; kk < ((p.dbg & 1) ? 0 : ksteps); ++kk) {
                    const uint64_t dp = smem_desc(sp + (kk >> 2) * pitch + (kk & 3) * 32, 0, 1024);
                    // V tile [keys][64 dims]: MN-major, one 64-wide atom, 16 keys per k-step = 2048 B
                    const uint64_t dv = smem_desc(sv + kk * 2048, (uint32_t)kv_bytes, 1024);
                    umma_bf16(tmem_base + O_COL, dp, dv, idesc_pv, kk > 0);
                }
                umma_commit(o_full(s));
                if (s == 1) umma_commit(v_empty(st));
            };
            issue_s(0, 0);
            for (int il = 0; il < my_items; ++il) {
                issue_pv(0, il);
                if (il == 0) issue_s(1, 0);        // slot 1 starts half a period after slot 0
                if (il + 1 < my_items) issue_s(0, il + 1);
                issue_pv(1, il);
                if (il + 1 < my_items) issue_s(1, il + 1);
            }
        }
      }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        // ------------------------------ softmax + epilogue slots ------------------
        const int slot = warp >> 2;
        const int quarter = warp & 3;                              // TMEM lane quarter this warp may access
        const int r = quarter * 32 + lane;                         // row of the tile owned by this thread
        const int row = slot * BM + r;                             // token index inside the image
        const bool warp_active = slot * BM + quarter * 32 < p.N;   // warp-uniform: any row of this warp inside the image
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const uint32_t ts = tmem_base + lane_addr + (uint32_t)(slot * S_STRIDE);
        const uint32_t to = tmem_base + lane_addr;
        const uint32_t sP_slot = slot ? sP1 : sP0;
        uint8_t* sP_ptr = smem + (sP_slot - base) + r * 128;
        const int pitch = slot ? p.pitch1 : ATOM_BYTES;
        const bool row_staged = r * 128 < pitch;                   // slot 1 stages rows1 rows per atom: the others must not write
        const int nchunks = (p.nkp + 31) / 32;
        const int n_keys = p.N, full_chunks = p.N / 32;
        const float scale_log2 = p.scale_log2;
        for (int il = 0; il < my_items; ++il) {
            const int item = blockIdx.x + il * gridDim.x;
            const int b = item / p.H, h = item % p.H;
            const uint32_t ph = (uint32_t)il & 1u;
            float mx = 0.f, lsum = 1.f;
            wait_bar(s_full(slot), ph);
            tc_fence_after();
            if (lane == 0) TL(warp, 6 * il);
            if (warp_active) {
                // Both passes are REAL loops over pairs of 32-key chunks (chunk c in va, c+1 in vb, the tcgen05.ld of the next
                // chunk in flight while one is processed): fully unrolled, the kernel was 92 KB of straight-line code and the
                // instruction cache hit rate 68 % ("no instruction" was the top stall).
                uint32_t va[32], vb[32];
                auto mask_fix = [&](uint32_t (&v)[32], int c) {     // the partial chunk: keys beyond N score -inf (exp2 -> 0)
                    if (c >= full_chunks) {
                        asm volatile("" ::: "memory");               // keep this a (warp-uniform) branch: if-converted it costs 99 instructions per chunk
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (c * 32 + j >= n_keys) v[j] = 0xFF800000u;
                    }
                };
                // ---- pass 1: row max (four independent chains, 3-input max)
                float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
                auto max_chunk = [&](uint32_t (&v)[32], int c) {
                    mask_fix(v, c);
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        m0 = max3(m0, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
                        m1 = max3(m1, __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                        m2 = max3(m2, __uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
                        m3 = max3(m3, __uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
                    }
                };
                if (!(p.dbg & 8)) {
                    tmem_ld32(ts, va);
                    tmem_ld_wait();
#pragma unroll 1
                    for (int c = 0; c < nchunks; c += 2) {
                        if (c + 1 < nchunks) tmem_ld32(ts + (c + 1) * 32, vb);
                        max_chunk(va, c);
                        tmem_ld_wait();
                        if (c + 1 < nchunks) {
                            if (c + 2 < nchunks) tmem_ld32(ts + (c + 2) * 32, va);
                            max_chunk(vb, c + 1);
                            tmem_ld_wait();
                        }
                    }
                    mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                }
                if (lane == 0) TL(warp, 6 * il + 1);
                // the previous item's output tile was staged in this slot's P buffer: its bulk store must have read it
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                // ---- pass 2: P = exp2(S * scale_log2 - max * scale_log2) as bf16 into this slot's P buffer; the row sum is
                //      taken from the ROUNDED values (the exact normaliser of the P V product), four independent chains
                const float mxs = mx * scale_log2;
                float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
                auto exp_chunk = [&](uint32_t (&v)[32], int c) {
                    mask_fix(v, c);
                    uint8_t* dst = sP_ptr + (c >> 1) * pitch;           // 64-key atom of this chunk
                    const int k8 = (c & 1) * 4;                         // first 8-key block of the chunk inside the atom
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float pr[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) pr[j] = ex2(fmaf(__uint_as_float(v[8 * i + j]), scale_log2, -mxs));
                        uint4 pk;
                        pk.x = pack_bf16(pr[0], pr[1]);
                        pk.y = pack_bf16(pr[2], pr[3]);
                        pk.z = pack_bf16(pr[4], pr[5]);
                        pk.w = pack_bf16(pr[6], pr[7]);
                        l0 += __uint_as_float(pk.x << 16) + __uint_as_float(pk.x & 0xFFFF0000u);
                        l1 += __uint_as_float(pk.y << 16) + __uint_as_float(pk.y & 0xFFFF0000u);
                        l2 += __uint_as_float(pk.z << 16) + __uint_as_float(pk.z & 0xFFFF0000u);
                        l3 += __uint_as_float(pk.w << 16) + __uint_as_float(pk.w & 0xFFFF0000u);
                        // (key blocks beyond nkp land in the unused tail of the last atom: never read by the MMA)
                        if (row_staged) *reinterpret_cast<uint4*>(dst + ((((k8 + i) ^ r) & 7) << 4)) = pk;
                    }
                };
                if (!(p.dbg & 2)) {
                    tmem_ld32(ts, va);
                    tmem_ld_wait();
#pragma unroll 1
                    for (int c = 0; c < nchunks; c += 2) {
                        if (c + 1 < nchunks) tmem_ld32(ts + (c + 1) * 32, vb);
                        exp_chunk(va, c);
                        tmem_ld_wait();
                        if (c + 1 < nchunks) {
                            if (c + 2 < nchunks) tmem_ld32(ts + (c + 2) * 32, va);
                            exp_chunk(vb, c + 1);
                            tmem_ld_wait();
                        }
                    }
                    lsum = (l0 + l1) + (l2 + l3);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to UMMA
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(p_full(slot));
                mbar_arrive(s_empty(slot));
                TL(warp, 6 * il + 2);
            }
            // ---- epilogue: O / sum (the other slot's exp pass runs meanwhile)
            wait_bar(o_full(slot), ph);
            tc_fence_after();
            if (lane == 0) TL(warp, 6 * il + 3);
            uint32_t o0[32], o1[32];
            if (warp_active) {
                tmem_ld32(to + O_COL, o0);
                tmem_ld32(to + O_COL + 32, o1);
                tmem_ld_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(o_empty(slot));
                TL(warp, 6 * il + 4);
            }
            if (warp_active && !(p.dbg & 4)) {
                // all P V MMAs of this slot completed (o_full): its P buffer is free and stages the bf16 output rows,
                // [row][64 dims] with the 128-byte swizzle the store's tensor map expects
                const float inv = 1.0f / lsum;
                float st[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t* o = q < 2 ? o0 + 16 * q : o1 + 16 * (q - 2);
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
                    st[2 * q] = s1 + t1;
                    st[2 * q + 1] = s2 + t2;
                    uint4 pk, qk;
                    pk.x = pack_bf16(f[0], f[1]); pk.y = pack_bf16(f[2], f[3]); pk.z = pack_bf16(f[4], f[5]); pk.w = pack_bf16(f[6], f[7]);
                    qk.x = pack_bf16(f[8], f[9]); qk.y = pack_bf16(f[10], f[11]); qk.z = pack_bf16(f[12], f[13]); qk.w = pack_bf16(f[14], f[15]);
                    *reinterpret_cast<uint4*>(sP_ptr + (((2 * q) ^ (r & 7)) << 4)) = pk;
                    *reinterpret_cast<uint4*>(sP_ptr + (((2 * q + 1) ^ (r & 7)) << 4)) = qk;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {        // this warp's 32 rows: one bulk tensor store, rows beyond the image clipped by the map
                    tma_store_3d(&map_out, sP_slot + (uint32_t)(quarter * 32 * 128), h * HD, slot * BM + quarter * 32, b);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (p.row_stats != nullptr) {
                    // statistics for the folded inner_attn_ln (f32 values before the bf16 rounding): 32 B per (row, head).  A lane
                    // pair writes the two halves of one row's sector in the same instruction (full 32-byte sectors).
                    const bool odd = lane & 1;
                    float4 mine_a = make_float4(st[0], st[1], st[2], st[3]), mine_b = make_float4(st[4], st[5], st[6], st[7]);
                    float4 send = odd ? mine_a : mine_b, recv;
                    recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
                    recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
                    recv.z = __shfl_xor_sync(0xffffffffu, send.z, 1);
                    recv.w = __shfl_xor_sync(0xffffffffu, send.w, 1);
                    const int row_e = row & ~1, row_o = row | 1;
                    float* base_e = p.row_stats + (((long long)b * p.N + row_e) * (4 * p.H) + 4 * h) * 2 + (odd ? 4 : 0);
                    float* base_o = p.row_stats + (((long long)b * p.N + row_o) * (4 * p.H) + 4 * h) * 2 + (odd ? 4 : 0);
                    if (row_e < p.N) *reinterpret_cast<float4*>(base_e) = odd ? recv : mine_a;     // even row: its a | its b (from the even lane)
                    if (row_o < p.N) *reinterpret_cast<float4*>(base_o) = odd ? mine_b : recv;     // odd row: its a (from the odd lane) | its b
                }
                if (p.lse != nullptr && row < p.N) p.lse[((long long)b * p.H + h) * p.N + row] = mx * p.scale + logf(lsum);
            }
            if (lane == 0) TL(warp, 6 * il + 5);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // staged rows must be read before the CTA retires
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace attn_tc4

// Returns CS_ERR_UNSUPPORTED (without setting up anything) when the shape is outside this kernel's envelope
// (two query tiles per head: 129 <= N <= 208); cs_attention_fwd then tries the other tcgen05 kernels.
int attention_fwd_tc4(const void* qkv, int B, int N, int H, float scale, void* out, float* lse, float* row_stats,
                      cudaStream_t st) {
    using namespace attn_tc4;
    if (N > MAX_NKP || N < MIN_N) return CS_ERR_UNSUPPORTED;
    const int nkp = ceil_div(N, 16) * 16;
    const int D = H * HD;
    const long long rows = (long long)B * N;
    if (rows >= (1ll << 31)) return CS_ERR_UNSUPPORTED;
    const int rows1 = ceil_div(N - BM, 8) * 8;
    CUtensorMap mq0, mq1, mkv, mout;
    int rc = make_map_bf16_3d(&mout, out, D, N, B, HD, 32);       // store map: [image][token][dim], box = 32 tokens x 64 dims
    if (rc) return rc;
    rc = make_map_bf16_2d(&mq0, qkv, rows, 3 * D, 3 * D, HD, BM);
    if (rc) return rc;
    rc = make_map_bf16_2d(&mq1, qkv, rows, 3 * D, 3 * D, HD, rows1);
    if (rc) return rc;
    rc = make_map_bf16_2d(&mkv, qkv, rows, 3 * D, 3 * D, HD, nkp);
    if (rc) return rc;
    Params p;
    p.B = B; p.N = N; p.H = H; p.nkp = nkp;
    p.rows1 = rows1;
    p.pitch1 = ceil_div(rows1 * 128, 1024) * 1024;
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
    const int smem = Q0_BYTES + Q1_BYTES + 3 * nkp * 128 + n_atoms * (p.pitch1 + ATOM_BYTES) + 256 + 1024;   // + barriers, align
    static int configured = 0;
    if (configured < smem) {
        CS_CUDA(cudaFuncSetAttribute(attention_fwd_tc4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    const int items = B * H;
    const int grid = items < num_sms() ? items : num_sms();
    attention_fwd_tc4_kernel<<<grid, THREADS, smem, st>>>(mq0, mq1, mkv, mout, p);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

}  // namespace cs

// debug: copies the phase timeline recorded by CTA 0 under CS_ATTN_DBG bit 16 ([12][256] clock64 values) to the host
extern "C" int cs_debug_attn_timeline(unsigned long long* host_out) {
    CS_CUDA(cudaDeviceSynchronize());
    CS_CUDA(cudaMemcpyFromSymbol(host_out, cs::attn_tc4::g_timeline, sizeof(unsigned long long) * cs::attn_tc4::TL_ROWS * cs::attn_tc4::TL_COLS));
    return CS_OK;
}
