// tcgen05 softmax attention forward for head_dim 64 and 129 <= N <= 208 tokens (EVA02-B/16: N = 197), fourth generation:
// the two query tiles of one (image, head) are two SLOTS that ping-pong on the tensor core and on the MUFU pipe.
//   S_s = Q_s K^T (tcgen05.mma -> TMEM)  ->  slot s: row max, P = exp2(.) as bf16 into swizzled shared memory  ->
//   O = P_s V (V consumed MN-major from its [key][dim] layout); row sums of the bf16-rounded P in registers  ->
//   O / rowsum -> bf16 rows staged in shared memory -> one TMA store per warp (full 128-byte lines).
// Replaces xformers.memory_efficient_attention at eva_vit_model.py:206-217 for the teacher's crops and the student.
//
// Why a fourth kernel (profiles/r02_attention_experiments.txt, r02_attn_tc4_*): attention_tc3 and its two siblings ran at
// 5x their MUFU bound.  Measured causes, in the order they were found and removed here:
//   * every handshake was a 512-thread (or 128-thread) mbarrier arrival on one shared-memory word -> ONE elected arrival
//     per warp (fence -> __syncwarp -> lane 0 arrives);
//   * the exp pass, the P V round trip and the epilogue of a tile ran back to back -> slot 0 (rows 0..127) and slot 1
//     (rows 128..N-1) are separate warp groups with their own S accumulator and P buffer; the first S MMA of slot 1 is
//     held back until slot 0 finished its first exp pass, so the slots run half a period apart;
//   * 16-byte output pieces scattered at a 1536-byte stride cost 2 us per head as direct stores (32 L2 transactions per
//     instruction) -> rows are staged in the dead P buffer and leave as cp.async.bulk.tensor stores;
//   * fully unrolled, the kernel was 92 KB of straight-line code with a 68 % instruction-cache hit rate -> real loops,
//     no printf in the wait path (46 KB -> the hot loops fit the L0/L1.5 instruction caches);
//   * one warp per scheduler in the exp pass issues every ~4 cycles (dependent FFMA -> MUFU -> F2FP -> FADD chains, ncu:
//     25 % issue utilisation) -> TWO warps per TMEM lane quarter and slot, each owning half of the key chunks and half of
//     the output dims (row max and row sum exchanged through shared memory, one 256-thread named barrier per tile);
//   * slot 1 only stages the rows its warps own (P atoms at a 32-row-granular pitch, the unused tensor-core rows read
//     whatever follows), which is what lets two P buffers, Q, K and a double-buffered V fit in 227 KB;
//   * warps whose 32 rows are all beyond N do nothing but the handshakes.
//
// Persistent CTA per SM, 640 threads = five warpgroups: warps 0-7 slot 0, warps 8-15 slot 1 (warp & 3 = TMEM lane quarter,
// (warp >> 2) & 1 = key / dim half), warp 16 TMA producer, warp 17 MMA issuer + TMEM allocator (18-19 idle); setmaxnreg moves
// the registers of the last group to the softmax groups.  TMEM columns: S0 [0,208) S1 [208,416) O [416,480).
#include "tc_common.cuh"

namespace cs {
namespace attn_tc4 {
using namespace cs::tc;

constexpr int HD = 64;
constexpr int BM = 128;
constexpr int THREADS = 640;
constexpr int TMA_WARP = 16, MMA_WARP = 17;
constexpr int SLOT_WARPS = 8;              // warps per slot: two per TMEM lane quarter
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
__device__ __forceinline__ uint32_t try_wait_nohint(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
    if (try_wait_nohint(bar, parity)) return;
    const long long t0 = clock64();
    while (!try_wait_nohint(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();        // a protocol bug surfaces as a CUDA error, never as a hung GPU
    }
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
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
constexpr int TL_ROWS = 20, TL_COLS = 256;
__device__ unsigned long long g_timeline[TL_ROWS][TL_COLS];
#define TL(rowi, idx)                                                                                  \
    do {                                                                                               \
        if ((p.dbg & 16) && blockIdx.x == 0 && (idx) < TL_COLS) g_timeline[rowi][idx] = clock64();     \
    } while (0)

struct Params {
    int B, N, H, nkp;               // nkp: keys padded to 16
    int rows1, pitch1;              // slot 1: Q rows loaded (N - 128 rounded up to 8) and its P atom pitch in bytes (32-row granular)
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
    float* xmax = reinterpret_cast<float*>(smem + (bar + 256u - base));      // [item parity][slot][half][128 rows] partial row max
    float* xsum = xmax + 2 * 2 * 2 * BM;                                      // [slot][half][128 rows] partial row sum

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
            mbar_init(s_empty(s), SLOT_WARPS);
            mbar_init(p_full(s), SLOT_WARPS);
            mbar_init(o_full(s), 1);
            mbar_init(o_empty(s), SLOT_WARPS);
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

    if (warp >= 2 * SLOT_WARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
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
                TL(16, 4 * il);
                wait_bar(q_empty(0), ph ^ 1u);
                mbar_arrive_expect_tx(q_full(0), Q0_BYTES);
                tma_load_2d(sQ0, &map_q0, q_full(0), h * HD, b * p.N);
                TL(16, 4 * il + 1);
                wait_bar(q_empty(1), ph ^ 1u);
                mbar_arrive_expect_tx(q_full(1), (uint32_t)(p.rows1 * 128));
                tma_load_2d(sQ1, &map_q1, q_full(1), h * HD, b * p.N + BM);
                TL(16, 4 * il + 2);
                const int st = il & 1;
                wait_bar(v_empty(st), (uint32_t)((il >> 1) & 1) ^ 1u);
                mbar_arrive_expect_tx(v_full(st), (uint32_t)kv_bytes);
                tma_load_2d(sV + st * kv_bytes, &map_kv, v_full(st), 2 * D + h * HD, b * p.N);
                TL(16, 4 * il + 3);
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
                TL(17, 8 * il + 2 * s);
                wait_bar(q_full(s), ph);
                if (s == 0) wait_bar(k_full, ph);
                wait_bar(s_empty(s), ph ^ 1u);                       // the slot has read S of its previous item
                tc_fence_after();
                TL(17, 8 * il + 2 * s + 1);
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
                TL(17, 8 * il + 4 + 2 * s);
                wait_bar(p_full(s), ph);
                if (s == 0) {
                    wait_bar(v_full(st), (uint32_t)(il >> 1) & 1u);
                    wait_bar(o_empty(1), ph ^ 1u);                   // O of slot 1's previous item has been read
                } else {
                    wait_bar(o_empty(0), ph);                        // O of slot 0's tile of this item has been read
                }
                tc_fence_after();
                TL(17, 8 * il + 5 + 2 * s);
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
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // ------------------------------ softmax + epilogue slots ------------------
        const int slot = warp >> 3;
        const int half = (warp >> 2) & 1;                          // which half of the key chunks / output dims of the row
        const int quarter = warp & 3;                              // TMEM lane quarter this warp may access
        const int r = quarter * 32 + lane;                         // row of the tile owned by this thread (shared with the partner warp)
        const int row = slot * BM + r;                             // token index inside the image
        const bool warp_active = slot * BM + quarter * 32 < p.N;   // warp-uniform: any row of this warp inside the image
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const uint32_t ts = tmem_base + lane_addr + (uint32_t)(slot * S_STRIDE);
        const uint32_t to = tmem_base + lane_addr + (uint32_t)(O_COL + half * 32);
        const uint32_t sP_slot = slot ? sP1 : sP0;
        const uint32_t pitch = slot ? (uint32_t)p.pitch1 : (uint32_t)ATOM_BYTES;
        const uint32_t sP_row = sP_slot + (uint32_t)r * 128u;      // shared-space address of this row inside atom 0
        const uint32_t sStage = sP_slot + (uint32_t)((quarter * 2 + half) * 2048);   // [32 rows][32 dims] bf16 output block of this warp
        const int nchunks = (p.nkp + 31) / 32;
        const int nsplit = (nchunks + 1) >> 1;
        const int c_begin = half ? nsplit : 0, c_end = half ? nchunks : nsplit;
        const int n_keys = p.N, full_chunks = p.N / 32;
        const float scale_log2 = p.scale_log2;
        float* xsum_mine = xsum + (slot * 2 + half) * BM + r;
        const float* xsum_other = xsum + (slot * 2 + (half ^ 1)) * BM + r;
        for (int il = 0; il < my_items; ++il) {
            const int item = blockIdx.x + il * gridDim.x;
            const int b = item / p.H, h = item % p.H;
            const uint32_t ph = (uint32_t)il & 1u;
            float mx = 0.f, lsum = 1.f;
            float* xmax_mine = xmax + (((il & 1) * 2 + slot) * 2 + half) * BM + r;
            const float* xmax_other = xmax + (((il & 1) * 2 + slot) * 2 + (half ^ 1)) * BM + r;
            wait_bar(s_full(slot), ph);
            tc_fence_after();
            if (lane == 0) TL(warp, 6 * il);
            uint32_t va[32], vb[32];
            auto mask_fix = [&](uint32_t (&v)[32], int c) {     // the partial chunk: keys beyond N score -inf (exp2 -> 0)
                if (c >= full_chunks) {
                    asm volatile("" ::: "memory");               // keep this a (warp-uniform) branch: if-converted it costs 99 instructions per chunk
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c * 32 + j >= n_keys) v[j] = 0xFF800000u;
                }
            };
            // ---- pass 1: partial row max over this warp's chunks (four independent chains, 3-input max).  Both passes are
            //      real loops over pairs of 32-key chunks: chunk c in va, c+1 in vb, the next tcgen05.ld in flight meanwhile.
            if (warp_active && !(p.dbg & 8)) {
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
                tmem_ld32(ts + c_begin * 32, va);
                tmem_ld_wait();
#pragma unroll 1
                for (int c = c_begin; c < c_end; c += 2) {
                    if (c + 1 < c_end) tmem_ld32(ts + (c + 1) * 32, vb);
                    max_chunk(va, c);
                    tmem_ld_wait();
                    if (c + 1 < c_end) {
                        if (c + 2 < c_end) tmem_ld32(ts + (c + 2) * 32, va);
                        max_chunk(vb, c + 1);
                        tmem_ld_wait();
                    }
                }
                mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                *xmax_mine = mx;
            }
            // the previous item's output block was staged in this slot's P buffer: its bulk store must have read it before ANY
            // warp of the slot writes P again, i.e. before the barrier below
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("bar.sync %0, 256;" ::"r"(1 + slot) : "memory");
            if (warp_active && !(p.dbg & 8)) mx = fmaxf(mx, *xmax_other);
            if (lane == 0) TL(warp, 6 * il + 1);
            // ---- pass 2: P = exp2(S * scale_log2 - max * scale_log2) as bf16 into this slot's P buffer; the row sum is taken
            //      from the ROUNDED values (the exact normaliser of the P V product), four independent chains
            if (warp_active && !(p.dbg & 2)) {
                const float mxs = mx * scale_log2;
                float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
                auto exp_chunk = [&](uint32_t (&v)[32], int c) {
                    mask_fix(v, c);
                    const uint32_t dst = sP_row + (uint32_t)(c >> 1) * pitch;      // 64-key atom of this chunk
                    const int k8 = (c & 1) * 4;                                    // first 8-key block of the chunk inside the atom
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
                        st_shared_v4(dst + (uint32_t)((((k8 + i) ^ r) & 7) << 4), pk);
                    }
                };
                tmem_ld32(ts + c_begin * 32, va);
                tmem_ld_wait();
#pragma unroll 1
                for (int c = c_begin; c < c_end; c += 2) {
                    if (c + 1 < c_end) tmem_ld32(ts + (c + 1) * 32, vb);
                    exp_chunk(va, c);
                    tmem_ld_wait();
                    if (c + 1 < c_end) {
                        if (c + 2 < c_end) tmem_ld32(ts + (c + 2) * 32, va);
                        exp_chunk(vb, c + 1);
                        tmem_ld_wait();
                    }
                }
                lsum = (l0 + l1) + (l2 + l3);
                *xsum_mine = lsum;                                             // read by the partner after o_full (ordered by the arrive below)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to UMMA
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(p_full(slot));
                mbar_arrive(s_empty(slot));
                TL(warp, 6 * il + 2);
            }
            // ---- epilogue: O / sum for this warp's 32 output dims (the other slot's exp pass runs meanwhile)
            wait_bar(o_full(slot), ph);
            tc_fence_after();
            if (lane == 0) TL(warp, 6 * il + 3);
            if (warp_active) {
                tmem_ld32(to, va);
                tmem_ld_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(o_empty(slot));
                TL(warp, 6 * il + 4);
            }
            if (warp_active && !(p.dbg & 4)) {
                // all P V MMAs of this slot completed (o_full): its P buffer is free and stages the bf16 output,
                // one [32 rows][32 dims] block per warp
                lsum += *xsum_other;
                const float inv = 1.0f / lsum;
                float st[4];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    float f[16];
                    float s1 = 0.f, s2 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
                    for (int z = 0; z < 16; z += 2) {
                        f[z] = __uint_as_float(va[16 * q + z]) * inv;
                        f[z + 1] = __uint_as_float(va[16 * q + z + 1]) * inv;
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
                    // 64-byte rows in the SWIZZLE_64B pattern of the store map (16-byte slot ^= (row >> 1) & 3): bank-conflict free
                    st_shared_v4(sStage + (uint32_t)(lane * 64) + (uint32_t)(((2 * q) ^ ((lane >> 1) & 3)) << 4), pk);
                    st_shared_v4(sStage + (uint32_t)(lane * 64) + (uint32_t)(((2 * q + 1) ^ ((lane >> 1) & 3)) << 4), qk);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {        // this warp's 32 rows x 32 dims: one bulk tensor store, rows beyond the image clipped by the map
                    tma_store_3d(&map_out, sStage, h * HD + half * 32, slot * BM + quarter * 32, b);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (row < p.N) {
                    // statistics for the folded inner_attn_ln (f32 values before the bf16 rounding): parts 2*half, 2*half+1 of this head
                    if (p.row_stats != nullptr)
                        *reinterpret_cast<float4*>(p.row_stats + (((long long)b * p.N + row) * (4 * p.H) + 4 * h + 2 * half) * 2) =
                            make_float4(st[0], st[1], st[2], st[3]);
                    if (p.lse != nullptr && half == 0) p.lse[((long long)b * p.H + h) * p.N + row] = mx * p.scale + logf(lsum);
                }
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
    int rc = make_map_bf16_3d(&mout, out, D, N, B, HD / 2, 32);   // store map: [image][token][dim], box = 32 tokens x 32 dims, dense in smem
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
    p.pitch1 = ceil_div(N - BM, 32) * 32 * 128;                     // every active warp of slot 1 owns 32 staged rows
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
    const int smem = Q0_BYTES + Q1_BYTES + 3 * nkp * 128 + n_atoms * (p.pitch1 + ATOM_BYTES) + 256 + 6144 + 1024;   // + barriers, align
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
