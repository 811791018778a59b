// Tower-level entry points of the C ABI: a frozen EVA02 vision tower as ONE handle and ONE call.
//
//   cs_pack_weights_bytes / _create / _update / _destroy   f32 state_dict tensors -> GEMM-ready bf16 operands with every
//                                                          LayerNorm folded (W' = W diag(gamma), c1 = rowsum(W'), c2 = W beta + b)
//   cs_query_workspace                                     activation scratch of one chunk (caller-owned)
//   cs_vit_forward_cls                                     EVAVisionTransformer.forward      (eva_vit_model.py:533-586): teacher
//   cs_vit_forward_dense                                   EVAVisionTransformer.encode_dense (eva_vit_model.py:588-623), no tape
//
// The kernel sequence of a block is the folded one of DESIGN.md §5.2 (5 launches, no LayerNorm pass):
//     qkv = rope(LN1-fold(xb Wqkv'^T))    att = softmax(q k^T / 8) v    x += LNi-fold(att Wproj'^T) -> x, xb, stats
//     h = silu(.)*(.) of LN2-fold(xb W12'^T)                            x += LNf-fold(h W3'^T)      -> x, xb, stats
// Everything the caller passes stays caller-owned (weights, workspace, images, outputs); the handle owns host bookkeeping,
// the cached TMA descriptors (tc_common.cu) and — after the first two calls with the same buffers — a CUDA graph of the
// chunk, so a steady-state call is one graph launch and makes no driver call besides it.
#include <cmath>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace cs {
namespace tower {

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct Cfg {
    int image_size, patch, width, heads, layers, hidden, embed_dim, pt_seq_len;
    float ln_eps;
    int grid() const { return image_size / patch; }
    int tokens() const { return grid() * grid() + 1; }
    int hidden_pad() const { return (hidden + 127) / 128 * 128; }
    int k_pe() const { return 3 * patch * patch; }
    int k_pe_pad() const { return (k_pe() + 7) / 8 * 8; }
};

static int stat_parts(int n) {
    const int t = (n % 256 == 0) ? 256 : 128;
    return 2 * ((n + t - 1) / t);
}

struct BlockPack {
    __nv_bfloat16 *wqkv_f, *wproj_f, *w12_f, *w3_f, *wproj;      // folded operands; wproj: un-folded (value-only block)
    float *c1_qkv, *c2_qkv, *c1_proj, *c2_proj, *c1_w12, *c2_w12, *c1_w3, *c2_w3;
    float *gi, *bi, *bproj;                                      // inner_attn_ln + proj bias of the value-only block
};

struct GridCtx {        // per-resolution constants (rope.py:179-214, eva_vit_model.py:631-643)
    float* rope_pos;    // [grid], device
    const float* pos;   // [1 + grid^2, D], device (native: inside the pack; other grids: caller-provided)
};

struct GraphKey {
    const void *images, *ws, *out;
    int n, grid, kind, dtype;
    bool operator<(const GraphKey& o) const {
        return std::tie(images, ws, out, n, grid, kind, dtype) < std::tie(o.images, o.ws, o.out, o.n, o.grid, o.kind, o.dtype);
    }
};

}  // namespace tower
}  // namespace cs

struct cs_tower {
    cs::tower::Cfg cfg;
    std::vector<cs::tower::BlockPack> blocks;
    __nv_bfloat16 *pe_w, *head_w;
    float *pe_b, *cls, *pos, *norm_g, *norm_b, *head_b, *rope_freq, *rope_pos_native;
    std::map<int, float*> rope_pos_other;                         // lazily cudaMalloc'ed [grid] vectors for other resolutions
    std::map<cs::tower::GraphKey, std::pair<int, cudaGraphExec_t>> graphs;     // (times seen, executable)
    std::mutex mu;
    bool use_graphs;
    bool cls_tail = true;                       // last block of cs_vit_forward_cls on the CLS rows only (CLIPSELF_FULL_LAST_BLOCK=1: off)
    cudaStream_t capture_stream = nullptr;      // graphs are captured here (the legacy default stream cannot be captured)
};

namespace cs {
namespace tower {

// ---------------------------------------------------------------------------------------------------------------
// packing kernels
// ---------------------------------------------------------------------------------------------------------------
// One warp per source row r of W [rows, K] (f32):  dst[row_map(r)][k] = bf16(W[r][k] * gamma[k]),  pad columns zero,
// c1[row_map(r)] = sum_k float(dst[.][k])  (from the ROUNDED values: what the tensor cores multiply),
// c2[row_map(r)] = sum_k W[r][k] * beta[k] + bias[r].   swiglu_half >= 0: row_map packs gate / up rows in 256-row tiles
// (packed row t*256 + half*128 + j  <-  source row t*128 + j), else row_map(r) = r + row_off.
__global__ void fold_rows_kernel(const float* __restrict__ W, long long ldw, int rows, int K, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, const float* __restrict__ bias, __nv_bfloat16* __restrict__ dst,
                                 long long ldd, int row_off, int swiglu_half, float* __restrict__ c1, float* __restrict__ c2) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + warp;
    if (r >= rows) return;
    const int dr = swiglu_half >= 0 ? (r / 128) * 256 + swiglu_half * 128 + (r % 128) : r + row_off;
    const float* w = W + (long long)r * ldw;
    __nv_bfloat16* d = dst + (long long)dr * ldd;
    float s1 = 0.f, s2 = 0.f;
    for (int k = lane; k < ldd; k += 32) {
        float v = 0.f;
        if (k < K) {
            const float x = w[k];
            v = gamma != nullptr ? x * gamma[k] : x;
            if (beta != nullptr) s2 = fmaf(x, beta[k], s2);
        }
        const __nv_bfloat16 b = __float2bfloat16_rn(v);
        d[k] = b;
        s1 += __bfloat162float(b);
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
        if (c1 != nullptr) c1[dr] = s1;
        if (c2 != nullptr) c2[dr] = s2 + (bias != nullptr ? bias[r] : 0.f);
    }
}

static int fold_rows(const float* W, long long ldw, int rows, int K, const float* gamma, const float* beta, const float* bias,
                     __nv_bfloat16* dst, long long ldd, int row_off, int swiglu_half, float* c1, float* c2, cudaStream_t st) {
    fold_rows_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(W, ldw, rows, K, gamma, beta, bias, dst, ldd, row_off, swiglu_half, c1, c2);
    CS_LAUNCH_CHECK();
    return CS_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// pack buffer layout (one carve function used by _bytes and _create so they cannot disagree)
// ---------------------------------------------------------------------------------------------------------------
struct Carver {
    uint8_t* base;
    int64_t off = 0;
    template <typename T>
    T* take(int64_t n) {
        off = align_up(off, 256);
        T* p = base != nullptr ? reinterpret_cast<T*>(base + off) : nullptr;
        off += n * (int64_t)sizeof(T);
        return p;
    }
};

static void carve(cs_tower* t, Carver& c) {
    const Cfg& g = t->cfg;
    const int D = g.width, Hp = g.hidden_pad(), C = g.embed_dim;
    t->pe_w = c.take<__nv_bfloat16>((int64_t)D * g.k_pe_pad());
    t->pe_b = c.take<float>(D);
    t->cls = c.take<float>(D);
    t->pos = c.take<float>((int64_t)g.tokens() * D);
    t->norm_g = c.take<float>(D);
    t->norm_b = c.take<float>(D);
    t->head_w = c.take<__nv_bfloat16>((int64_t)C * D);
    t->head_b = c.take<float>(C);
    t->rope_freq = c.take<float>(16);
    t->rope_pos_native = c.take<float>(64);
    t->blocks.resize(g.layers);
    for (auto& b : t->blocks) {
        b.wqkv_f = c.take<__nv_bfloat16>((int64_t)3 * D * D);
        b.c1_qkv = c.take<float>(3 * D);
        b.c2_qkv = c.take<float>(3 * D);
        b.wproj_f = c.take<__nv_bfloat16>((int64_t)D * D);
        b.wproj = c.take<__nv_bfloat16>((int64_t)D * D);
        b.c1_proj = c.take<float>(D);
        b.c2_proj = c.take<float>(D);
        b.w12_f = c.take<__nv_bfloat16>((int64_t)2 * Hp * D);
        b.c1_w12 = c.take<float>(2 * Hp);
        b.c2_w12 = c.take<float>(2 * Hp);
        b.w3_f = c.take<__nv_bfloat16>((int64_t)D * Hp);
        b.c1_w3 = c.take<float>(D);
        b.c2_w3 = c.take<float>(D);
        b.gi = c.take<float>(D);
        b.bi = c.take<float>(D);
        b.bproj = c.take<float>(D);
    }
    c.off = align_up(c.off, 256);
}

static int check_cfg(const cs_tower_cfg_t* c) {
    CS_CHECK_ARG(c != nullptr, "tower config: null pointer");
    CS_CHECK_ARG(c->patch > 0 && c->image_size > 0 && c->image_size % c->patch == 0 && c->image_size / c->patch <= 64,
                 "tower config: image_size must be a multiple of patch with a grid of at most 64 x 64");
    CS_CHECK_ARG(c->heads > 0 && c->width == 64 * c->heads, "tower config: head_dim must be 64 (width = 64 * heads)");
    CS_CHECK_ARG(c->layers > 0 && c->hidden > 0 && c->embed_dim > 0 && c->embed_dim % 32 == 0 && c->pt_seq_len > 0,
                 "tower config: bad layers / hidden / embed_dim / pt_seq_len");
    return CS_OK;
}

static Cfg to_cfg(const cs_tower_cfg_t* c) {
    return Cfg{c->image_size, c->patch, c->width, c->heads, c->layers, c->hidden, c->embed_dim, c->pt_seq_len, c->ln_eps};
}

static void rope_pos_host(int grid, int pt, float* out) {       // rope.py:127: arange(grid) / grid * pt, f32 steps
    for (int i = 0; i < grid; ++i) out[i] = (float)i / (float)grid * (float)pt;
}

static int pack(cs_tower* t, const char* const* names, const void* const* tensors, int count, cudaStream_t st) {
    const Cfg& g = t->cfg;
    std::unordered_map<std::string, const float*> sd;
    for (int i = 0; i < count; ++i) {
        CS_CHECK_ARG(names[i] != nullptr && tensors[i] != nullptr, "cs_pack_weights: null name / tensor at index %d", i);
        sd[names[i]] = static_cast<const float*>(tensors[i]);
    }
    const float* missing_sentinel = nullptr;
    std::string missing;
    auto get = [&](const std::string& k) -> const float* {
        auto it = sd.find(k);
        if (it == sd.end()) {
            if (missing.empty()) missing = k;
            return missing_sentinel;
        }
        return it->second;
    };
    const int D = g.width, Hd = g.hidden, Hp = g.hidden_pad(), C = g.embed_dim;
    std::vector<std::string> need = {"patch_embed.proj.weight", "patch_embed.proj.bias", "cls_token", "pos_embed", "norm.weight",
                                     "norm.bias", "head.weight", "head.bias"};
    for (int i = 0; i < g.layers; ++i) {
        const std::string p = "blocks." + std::to_string(i) + ".";
        for (const char* s : {"norm1.weight", "norm1.bias", "attn.q_proj.weight", "attn.k_proj.weight", "attn.v_proj.weight", "attn.q_bias",
                              "attn.v_bias", "attn.inner_attn_ln.weight", "attn.inner_attn_ln.bias", "attn.proj.weight", "attn.proj.bias",
                              "norm2.weight", "norm2.bias", "mlp.w1.weight", "mlp.w1.bias", "mlp.w2.weight", "mlp.w2.bias",
                              "mlp.ffn_ln.weight", "mlp.ffn_ln.bias", "mlp.w3.weight", "mlp.w3.bias"})
            need.push_back(p + s);
    }
    for (const auto& k : need) get(k);
    CS_CHECK_ARG(missing.empty(), "cs_pack_weights: state_dict entry '%s' is missing", missing.c_str());

    auto copy_f32 = [&](float* dst, const float* src, int64_t n) {
        return cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
    };
    int rc;
    // frozen head / embed / final norm
    if ((rc = fold_rows(get("patch_embed.proj.weight"), g.k_pe(), D, g.k_pe(), nullptr, nullptr, nullptr, t->pe_w, g.k_pe_pad(), 0, -1,
                        nullptr, nullptr, st))) return rc;
    if ((rc = fold_rows(get("head.weight"), D, C, D, nullptr, nullptr, nullptr, t->head_w, D, 0, -1, nullptr, nullptr, st))) return rc;
    CS_CUDA(copy_f32(t->pe_b, get("patch_embed.proj.bias"), D));
    CS_CUDA(copy_f32(t->cls, get("cls_token"), D));
    CS_CUDA(copy_f32(t->pos, get("pos_embed"), (int64_t)g.tokens() * D));
    CS_CUDA(copy_f32(t->norm_g, get("norm.weight"), D));
    CS_CUDA(copy_f32(t->norm_b, get("norm.bias"), D));
    CS_CUDA(copy_f32(t->head_b, get("head.bias"), C));
    float host[80];
    for (int n = 0; n < 16; ++n) host[n] = 1.0f / powf(10000.0f, (float)(2 * n) / 32.0f);       // rope.py:118 (dim = head_dim / 2 = 32)
    rope_pos_host(g.grid(), g.pt_seq_len, host + 16);
    CS_CUDA(cudaMemcpyAsync(t->rope_freq, host, 16 * sizeof(float), cudaMemcpyHostToDevice, st));
    CS_CUDA(cudaMemcpyAsync(t->rope_pos_native, host + 16, g.grid() * sizeof(float), cudaMemcpyHostToDevice, st));
    CS_CUDA(cudaStreamSynchronize(st));                                                           // `host` leaves scope
    for (int i = 0; i < g.layers; ++i) {
        const std::string p = "blocks." + std::to_string(i) + ".";
        BlockPack& b = t->blocks[i];
        const float *g1 = get(p + "norm1.weight"), *b1 = get(p + "norm1.bias");
        const float *gi = get(p + "attn.inner_attn_ln.weight"), *bi = get(p + "attn.inner_attn_ln.bias");
        const float *g2 = get(p + "norm2.weight"), *b2 = get(p + "norm2.bias");
        const float *gf = get(p + "mlp.ffn_ln.weight"), *bf = get(p + "mlp.ffn_ln.bias");
        // q | k | v with norm1 folded (k has no bias: eva_vit_model.py:127-129)
        if ((rc = fold_rows(get(p + "attn.q_proj.weight"), D, D, D, g1, b1, get(p + "attn.q_bias"), b.wqkv_f, D, 0, -1, b.c1_qkv, b.c2_qkv, st))) return rc;
        if ((rc = fold_rows(get(p + "attn.k_proj.weight"), D, D, D, g1, b1, nullptr, b.wqkv_f, D, D, -1, b.c1_qkv, b.c2_qkv, st))) return rc;
        if ((rc = fold_rows(get(p + "attn.v_proj.weight"), D, D, D, g1, b1, get(p + "attn.v_bias"), b.wqkv_f, D, 2 * D, -1, b.c1_qkv, b.c2_qkv, st))) return rc;
        // proj with inner_attn_ln folded, and plain (value-only block)
        if ((rc = fold_rows(get(p + "attn.proj.weight"), D, D, D, gi, bi, get(p + "attn.proj.bias"), b.wproj_f, D, 0, -1, b.c1_proj, b.c2_proj, st))) return rc;
        if ((rc = fold_rows(get(p + "attn.proj.weight"), D, D, D, nullptr, nullptr, nullptr, b.wproj, D, 0, -1, nullptr, nullptr, st))) return rc;
        CS_CUDA(copy_f32(b.gi, gi, D));
        CS_CUDA(copy_f32(b.bi, bi, D));
        CS_CUDA(copy_f32(b.bproj, get(p + "attn.proj.bias"), D));
        // w1 | w2 packed in 256-row tiles (128 gate + 128 up rows) with norm2 folded; padded rows stay zero
        CS_CUDA(cudaMemsetAsync(b.w12_f, 0, (size_t)2 * Hp * D * sizeof(__nv_bfloat16), st));
        CS_CUDA(cudaMemsetAsync(b.c1_w12, 0, (size_t)2 * Hp * sizeof(float), st));
        CS_CUDA(cudaMemsetAsync(b.c2_w12, 0, (size_t)2 * Hp * sizeof(float), st));
        if ((rc = fold_rows(get(p + "mlp.w1.weight"), D, Hd, D, g2, b2, get(p + "mlp.w1.bias"), b.w12_f, D, 0, 0, b.c1_w12, b.c2_w12, st))) return rc;
        if ((rc = fold_rows(get(p + "mlp.w2.weight"), D, Hd, D, g2, b2, get(p + "mlp.w2.bias"), b.w12_f, D, 0, 1, b.c1_w12, b.c2_w12, st))) return rc;
        // w3 with ffn_ln folded (K = hidden, rows padded to hidden_pad with zeros)
        if ((rc = fold_rows(get(p + "mlp.w3.weight"), Hd, D, Hd, gf, bf, get(p + "mlp.w3.bias"), b.w3_f, Hp, 0, -1, b.c1_w3, b.c2_w3, st))) return rc;
    }
    return CS_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------------------------------
struct Workspace {
    float *x, *stats_x, *stats_att, *stats_h, *head;
    __nv_bfloat16 *xb, *qkv, *att, *h, *u, *patches, *cls_ln;
};

static int64_t carve_ws(const Cfg& g, int images, int grid, uint8_t* base, Workspace* w) {
    Carver c{base};
    const int N = grid * grid + 1, D = g.width, Hp = g.hidden_pad();
    const int64_t rows = (int64_t)images * N;
    Workspace tmp;
    Workspace& o = w != nullptr ? *w : tmp;
    o.x = c.take<float>(rows * D);
    o.xb = c.take<__nv_bfloat16>(rows * D);
    o.qkv = c.take<__nv_bfloat16>(rows * 3 * D);
    o.att = c.take<__nv_bfloat16>(rows * D);
    o.u = c.take<__nv_bfloat16>(rows * D);
    o.h = c.take<__nv_bfloat16>(rows * Hp);
    o.stats_x = c.take<float>(rows * stat_parts(D) * 2);
    o.stats_att = c.take<float>(rows * 4 * g.heads * 2);
    o.stats_h = c.take<float>(rows * (Hp / 64) * 2);
    o.patches = c.take<__nv_bfloat16>((int64_t)images * (N - 1) * g.k_pe_pad());
    o.cls_ln = c.take<__nv_bfloat16>((int64_t)images * D);
    o.head = c.take<float>((int64_t)images * (N - 1) * g.embed_dim);
    return align_up(c.off, 256);
}

// ---------------------------------------------------------------------------------------------------------------
// kernel sequencing
// ---------------------------------------------------------------------------------------------------------------
static cs_gemm_epilogue_t epi0() {
    cs_gemm_epilogue_t e;
    memset(&e, 0, sizeof(e));
    e.alpha = 1.0f;
    return e;
}

static int embed(const cs_tower* t, const GridCtx& gc, int grid, const void* images, cs_dtype_t dtype, int n, const Workspace& w,
                 void* st) {
    const Cfg& g = t->cfg;
    const int N = grid * grid + 1, D = g.width;
    int rc;
    if ((rc = cs_im2col_patches(images, dtype, n, grid * g.patch, g.patch, w.patches, g.k_pe_pad(), st))) return rc;
    cs_gemm_epilogue_t e = epi0();
    e.mode = CS_EPI_TOKENS;
    e.out_dtype = CS_F32;
    e.out = w.x;
    e.ldo = D;
    e.bias = t->pe_b;
    e.pos_embed = gc.pos;
    e.tokens = N;
    if ((rc = cs_gemm_bf16(w.patches, g.k_pe_pad(), t->pe_w, g.k_pe_pad(), (int64_t)n * (N - 1), D, g.k_pe_pad(), &e, st))) return rc;
    if ((rc = cs_fill_cls_rows(t->cls, gc.pos, n, N, D, w.x, st))) return rc;
    return cs_row_stats_cast(w.x, D, (int64_t)n * N, D, w.xb, D, w.stats_x, stat_parts(D), st);
}

static int block(const cs_tower* t, const GridCtx& gc, int grid, int i, int n, const Workspace& w, bool with_attention, void* st) {
    const Cfg& g = t->cfg;
    const BlockPack& b = t->blocks[i];
    const int N = grid * grid + 1, D = g.width, Hp = g.hidden_pad();
    const int64_t M = (int64_t)n * N;
    int rc;
    auto fold = [&](cs_gemm_epilogue_t& e, const float* stats, const float* c1, int parts, int dim) {
        e.ln_stats = stats;
        e.ln_c1 = c1;
        e.ln_parts = parts;
        e.ln_dim = dim;
        e.ln_eps = g.ln_eps;
    };
    auto emit_x = [&](cs_gemm_epilogue_t& e) {          // x += ... in place, new stream also as bf16 + row statistics
        e.mode = CS_EPI_STORE;
        e.out_dtype = CS_F32;
        e.out = w.x;
        e.ldo = D;
        e.residual = w.x;
        e.ldr = D;
        e.out2_bf16 = w.xb;
        e.ldo2 = D;
        e.stats_out = w.stats_x;
    };
    if (with_attention) {
        cs_gemm_epilogue_t e = epi0();
        e.mode = CS_EPI_QKV_ROPE;
        e.out_dtype = CS_BF16;
        e.out = w.qkv;
        e.ldo = 3 * D;
        e.bias = b.c2_qkv;
        e.rope_pos = gc.rope_pos;
        e.rope_freq = t->rope_freq;
        e.rope_grid = grid;
        e.tokens = N;
        e.rope_cols = 2 * D;
        fold(e, w.stats_x, b.c1_qkv, stat_parts(D), D);
        if ((rc = cs_gemm_bf16(w.xb, D, b.wqkv_f, D, M, 3 * D, D, &e, st))) return rc;
        if ((rc = cs_attention_fwd(w.qkv, n, N, g.heads, 0.125f, w.att, nullptr, w.stats_att, st))) return rc;
        cs_gemm_epilogue_t p = epi0();
        emit_x(p);
        p.bias = b.c2_proj;
        fold(p, w.stats_att, b.c1_proj, 4 * g.heads, D);
        if ((rc = cs_gemm_bf16(w.att, D, b.wproj_f, D, M, D, D, &p, st))) return rc;
    } else {        // forward_without_attn (eva_vit_model.py:317-324, 249-256): v-projection, explicit inner LN, proj
        cs_gemm_epilogue_t e = epi0();
        e.mode = CS_EPI_STORE;
        e.out_dtype = CS_BF16;
        e.out = w.att;
        e.ldo = D;
        e.bias = b.c2_qkv + 2 * D;
        fold(e, w.stats_x, b.c1_qkv + 2 * D, stat_parts(D), D);
        if ((rc = cs_gemm_bf16(w.xb, D, b.wqkv_f + (int64_t)2 * D * D, D, M, D, D, &e, st))) return rc;
        if ((rc = cs_layernorm_fwd(w.att, CS_BF16, D, M, D, 0, 1, 0, b.gi, b.bi, g.ln_eps, w.u, D, nullptr, nullptr, st))) return rc;
        cs_gemm_epilogue_t p = epi0();
        emit_x(p);
        p.bias = b.bproj;
        if ((rc = cs_gemm_bf16(w.u, D, b.wproj, D, M, D, D, &p, st))) return rc;
    }
    cs_gemm_epilogue_t s = epi0();
    s.mode = CS_EPI_SWIGLU;
    s.out_dtype = CS_BF16;
    s.out = w.h;
    s.ldo = Hp;
    s.bias = b.c2_w12;
    s.stats_out = w.stats_h;
    fold(s, w.stats_x, b.c1_w12, stat_parts(D), D);
    if ((rc = cs_gemm_bf16(w.xb, D, b.w12_f, D, M, 2 * Hp, D, &s, st))) return rc;
    cs_gemm_epilogue_t o = epi0();
    emit_x(o);
    o.bias = b.c2_w3;
    fold(o, w.stats_h, b.c1_w3, Hp / 64, g.hidden);
    return cs_gemm_bf16(w.h, Hp, b.w3_f, Hp, M, D, Hp, &o, st);
}

// The LAST block of a tower that is read at the CLS token only (encode_image: head(norm(x)[:, 0]), eva_vit_model.py:565-569):
// K and V need every token, so the QKV GEMM is unchanged, but attention, proj, the SwiGLU MLP and w3 run on the n CLS rows
// instead of n x N rows (same kernels, same folding; ~72 % of one block's work disappears).  The compact CLS-row buffers
// live in scratch that the folded CLS path does not use otherwise (w.att, w.stats_att, w.u).  x_cls: the new residual
// stream of the CLS rows, [n, D] f32, which the caller normalises and projects.
struct ClsTail {
    float *x, *stats_x, *stats_h;
    __nv_bfloat16 *xb, *h;
};
static bool carve_cls_tail(const Cfg& g, int n, int N, const Workspace& w, ClsTail* c) {
    const int D = g.width, Hp = g.hidden_pad();
    Carver cv{reinterpret_cast<uint8_t*>(w.u)};
    c->x = cv.take<float>((int64_t)n * D);
    c->xb = cv.take<__nv_bfloat16>((int64_t)n * D);
    c->h = cv.take<__nv_bfloat16>((int64_t)n * Hp);
    c->stats_x = cv.take<float>((int64_t)n * stat_parts(D) * 2);
    c->stats_h = cv.take<float>((int64_t)n * (Hp / 64) * 2);
    return cv.off <= (int64_t)n * N * D * 2 && N <= 1024;        // w.u holds n x N x D bf16
}
static int block_cls_tail(const cs_tower* t, const GridCtx& gc, int grid, int i, int n, const Workspace& w, const ClsTail& c, void* st) {
    const Cfg& g = t->cfg;
    const BlockPack& b = t->blocks[i];
    const int N = grid * grid + 1, D = g.width, Hp = g.hidden_pad();
    const int64_t M = (int64_t)n * N;
    int rc;
    auto fold = [&](cs_gemm_epilogue_t& e, const float* stats, const float* c1, int parts, int dim) {
        e.ln_stats = stats;
        e.ln_c1 = c1;
        e.ln_parts = parts;
        e.ln_dim = dim;
        e.ln_eps = g.ln_eps;
    };
    auto emit_cls = [&](cs_gemm_epilogue_t& e, const float* residual, int64_t ldr) {
        e.mode = CS_EPI_STORE;
        e.out_dtype = CS_F32;
        e.out = c.x;
        e.ldo = D;
        e.residual = residual;
        e.ldr = ldr;
        e.out2_bf16 = c.xb;
        e.ldo2 = D;
        e.stats_out = c.stats_x;
    };
    cs_gemm_epilogue_t e = epi0();              // q | k | v of every token (k, v feed the CLS query)
    e.mode = CS_EPI_QKV_ROPE;
    e.out_dtype = CS_BF16;
    e.out = w.qkv;
    e.ldo = 3 * D;
    e.bias = b.c2_qkv;
    e.rope_pos = gc.rope_pos;
    e.rope_freq = t->rope_freq;
    e.rope_grid = grid;
    e.tokens = N;
    e.rope_cols = 2 * D;
    fold(e, w.stats_x, b.c1_qkv, stat_parts(D), D);
    if ((rc = cs_gemm_bf16(w.xb, D, b.wqkv_f, D, M, 3 * D, D, &e, st))) return rc;
    if ((rc = cs_attention_cls_fwd(w.qkv, n, N, g.heads, 0.125f, w.att, w.stats_att, st))) return rc;     // -> [n, D], [n, 4H, 2]
    cs_gemm_epilogue_t p = epi0();              // x_cls = x[CLS rows] + inner_attn_ln-fold(att_cls Wproj'^T)
    emit_cls(p, w.x, (int64_t)N * D);
    p.bias = b.c2_proj;
    fold(p, w.stats_att, b.c1_proj, 4 * g.heads, D);
    if ((rc = cs_gemm_bf16(w.att, D, b.wproj_f, D, n, D, D, &p, st))) return rc;
    cs_gemm_epilogue_t s = epi0();
    s.mode = CS_EPI_SWIGLU;
    s.out_dtype = CS_BF16;
    s.out = c.h;
    s.ldo = Hp;
    s.bias = b.c2_w12;
    s.stats_out = c.stats_h;
    fold(s, c.stats_x, b.c1_w12, stat_parts(D), D);
    if ((rc = cs_gemm_bf16(c.xb, D, b.w12_f, D, n, 2 * Hp, D, &s, st))) return rc;
    cs_gemm_epilogue_t o = epi0();              // x_cls += ffn_ln-fold(h_cls W3'^T), in place
    emit_cls(o, c.x, D);
    o.bias = b.c2_w3;
    fold(o, c.stats_h, b.c1_w3, Hp / 64, g.hidden);
    return cs_gemm_bf16(c.h, Hp, b.w3_f, Hp, n, D, Hp, &o, st);
}

enum Kind { KIND_CLS = 0, KIND_DENSE = 1 };

static int run_chunk(const cs_tower* t, const GridCtx& gc, int grid, int kind, const void* images, cs_dtype_t dtype, int n,
                     const Workspace& w, float* out, void* st) {
    const Cfg& g = t->cfg;
    const int N = grid * grid + 1, D = g.width;
    int rc;
    if ((rc = embed(t, gc, grid, images, dtype, n, w, st))) return rc;
    ClsTail ct;
    const bool cls_tail = kind == KIND_CLS && t->cls_tail && carve_cls_tail(g, n, N, w, &ct);
    for (int i = 0; i < g.layers; ++i) {
        if (cls_tail && i + 1 == g.layers) rc = block_cls_tail(t, gc, grid, i, n, w, ct, st);
        else rc = block(t, gc, grid, i, n, w, kind == KIND_CLS || i + 1 < g.layers, st);
        if (rc) return rc;
    }
    cs_gemm_epilogue_t e = epi0();
    e.mode = CS_EPI_STORE;
    e.out_dtype = CS_F32;
    e.bias = t->head_b;
    e.ldo = g.embed_dim;
    if (kind == KIND_CLS) {        // norm(x)[:, 0] -> head (eva_vit_model.py:565-569)
        if (cls_tail) rc = cs_layernorm_fwd(ct.x, CS_F32, D, n, D, 0, 1, 0, t->norm_g, t->norm_b, g.ln_eps, w.cls_ln, D, nullptr, nullptr, st);
        else rc = cs_layernorm_fwd(w.x, CS_F32, D, n, D, 0, N, 0, t->norm_g, t->norm_b, g.ln_eps, w.cls_ln, D, nullptr, nullptr, st);
        if (rc) return rc;
        e.out = out;
        return cs_gemm_bf16(w.cls_ln, D, t->head_w, D, n, g.embed_dim, D, &e, st);
    }
    // dense: drop CLS, norm, head, per-token L2 normalise -> NHWC (eva_vit_model.py:615-623)
    const int64_t Mp = (int64_t)n * (N - 1);
    if ((rc = cs_layernorm_fwd(w.x, CS_F32, D, Mp, D, N - 1, 1, 1, t->norm_g, t->norm_b, g.ln_eps, w.u, D, nullptr, nullptr, st))) return rc;
    e.out = w.head;
    if ((rc = cs_gemm_bf16(w.u, D, t->head_w, D, Mp, g.embed_dim, D, &e, st))) return rc;
    return cs_l2norm_fwd(w.head, Mp, g.embed_dim, out, nullptr, st);
}

static int grid_ctx(cs_tower* t, int grid, const float* pos_override, cudaStream_t st, GridCtx* gc) {
    const Cfg& g = t->cfg;
    if (grid == g.grid()) {
        gc->rope_pos = t->rope_pos_native;
        gc->pos = pos_override != nullptr ? pos_override : t->pos;
        return CS_OK;
    }
    CS_CHECK_ARG(pos_override != nullptr, "a resolution other than the tower's own needs the rescaled pos_embed [1 + grid^2, width] "
                                          "(eva_vit_model.py:631-643: bicubic, done by the caller once per resolution)");
    auto it = t->rope_pos_other.find(grid);
    if (it == t->rope_pos_other.end()) {
        float host[64];
        rope_pos_host(grid, g.pt_seq_len, host);
        float* d = nullptr;
        CS_CUDA(cudaMalloc(&d, 64 * sizeof(float)));
        CS_CUDA(cudaMemcpyAsync(d, host, grid * sizeof(float), cudaMemcpyHostToDevice, st));
        CS_CUDA(cudaStreamSynchronize(st));
        it = t->rope_pos_other.emplace(grid, d).first;
    }
    gc->rope_pos = it->second;
    gc->pos = pos_override;
    return CS_OK;
}

// One chunk, replayed from a CUDA graph once the same (buffers, shape) has been seen twice.
static int run_chunk_graphed(cs_tower* t, const GridCtx& gc, int grid, int kind, const void* images, cs_dtype_t dtype, int n,
                             void* workspace, const Workspace& w, float* out, cudaStream_t st) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    if (!t->use_graphs || cap != cudaStreamCaptureStatusNone) return run_chunk(t, gc, grid, kind, images, dtype, n, w, out, st);
    const GraphKey key{images, workspace, out, n, grid, kind, (int)dtype};
    std::unique_lock<std::mutex> lock(t->mu);
    auto& ent = t->graphs[key];
    if (ent.second != nullptr) return cudaGraphLaunch(ent.second, st) == cudaSuccess ? CS_OK : (set_error("cudaGraphLaunch failed"), CS_ERR_CUDA);
    if (ent.first++ == 0) {        // first sighting: eager (kernel attributes, descriptor cache)
        lock.unlock();
        return run_chunk(t, gc, grid, kind, images, dtype, n, w, out, st);
    }
    if (t->graphs.size() > 512) {       // bounded: drop everything but this key
        for (auto& kv : t->graphs)
            if (kv.second.second != nullptr) cudaGraphExecDestroy(kv.second.second);
        t->graphs.clear();
        t->graphs[key].first = 2;
    }
    // capture on a private stream (nothing executes during capture), launch on the caller's stream
    if (t->capture_stream == nullptr) CS_CUDA(cudaStreamCreateWithFlags(&t->capture_stream, cudaStreamNonBlocking));
    CS_CUDA(cudaStreamBeginCapture(t->capture_stream, cudaStreamCaptureModeThreadLocal));
    const int rc = run_chunk(t, gc, grid, kind, images, dtype, n, w, out, t->capture_stream);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(t->capture_stream, &graph);
    if (rc != CS_OK || ce != cudaSuccess || graph == nullptr) {
        if (graph != nullptr) cudaGraphDestroy(graph);
        if (rc == CS_OK) set_error("stream capture of the tower chunk failed: %s", cudaGetErrorString(ce));
        return rc != CS_OK ? rc : CS_ERR_CUDA;
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) {
        set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
        return CS_ERR_CUDA;
    }
    t->graphs[key].second = exec;
    CS_CUDA(cudaGraphLaunch(exec, st));
    return CS_OK;
}

static void drop_graphs(cs_tower* t) {
    for (auto& kv : t->graphs)
        if (kv.second.second != nullptr) cudaGraphExecDestroy(kv.second.second);
    t->graphs.clear();
}

static int forward(cs_tower* t, int kind, const void* images, cs_dtype_t dtype, int n_images, int image_size, const float* pos_override,
                   void* workspace, int64_t workspace_bytes, int chunk_images, float* out, void* stream) {
    CS_CHECK_ARG(t && images && workspace && out, "cs_vit_forward: null pointer");
    CS_CHECK_ARG(dtype == CS_F32 || dtype == CS_BF16, "cs_vit_forward: images must be f32 or bf16");
    const Cfg& g = t->cfg;
    if (image_size <= 0) image_size = g.image_size;
    CS_CHECK_ARG(image_size % g.patch == 0 && image_size / g.patch <= 64, "cs_vit_forward: image_size must be a multiple of the patch "
                                                                          "size with a grid of at most 64 x 64");
    CS_CHECK_ARG(n_images > 0 && chunk_images > 0, "cs_vit_forward: n_images and chunk_images must be positive");
    const int grid = image_size / g.patch, N = grid * grid + 1;
    if (chunk_images > n_images) chunk_images = n_images;
    Workspace w;
    const int64_t need = carve_ws(g, chunk_images, grid, (uint8_t*)workspace, &w);
    CS_CHECK_ARG(workspace_bytes >= need, "cs_vit_forward: workspace too small (%lld < %lld bytes, see cs_query_workspace)",
                 (long long)workspace_bytes, (long long)need);
    CS_CHECK_ARG((uintptr_t)workspace % 256 == 0, "cs_vit_forward: workspace must be 256 B aligned");
    cudaStream_t st = (cudaStream_t)stream;
    GridCtx gc;
    int rc = grid_ctx(t, grid, pos_override, st, &gc);
    if (rc) return rc;
    const int64_t img_elems = (int64_t)3 * image_size * image_size;
    const int64_t esize = dtype == CS_F32 ? 4 : 2;
    const int64_t out_per_image = kind == KIND_CLS ? g.embed_dim : (int64_t)(N - 1) * g.embed_dim;
    for (int s = 0; s < n_images; s += chunk_images) {
        const int n = n_images - s < chunk_images ? n_images - s : chunk_images;
        const uint8_t* img = (const uint8_t*)images + (int64_t)s * img_elems * esize;
        if ((rc = run_chunk_graphed(t, gc, grid, kind, img, dtype, n, workspace, w, out + (int64_t)s * out_per_image, st))) return rc;
    }
    return CS_OK;
}

}  // namespace tower
}  // namespace cs

using namespace cs;
using namespace cs::tower;

extern "C" int cs_pack_weights_bytes(const cs_tower_cfg_t* cfg, int64_t* bytes) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    CS_CHECK_ARG(bytes != nullptr, "cs_pack_weights_bytes: null pointer");
    cs_tower tmp;
    tmp.cfg = to_cfg(cfg);
    Carver c{nullptr};
    carve(&tmp, c);
    *bytes = c.off;
    return CS_OK;
}

extern "C" int cs_pack_weights_create(const cs_tower_cfg_t* cfg, const char* const* names, const void* const* tensors_f32, int count,
                                      void* pack_buffer, int64_t pack_bytes, void* stream, cs_tower_t** out) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    CS_CHECK_ARG(names && tensors_f32 && pack_buffer && out && count > 0, "cs_pack_weights_create: null pointer");
    CS_CHECK_ARG((uintptr_t)pack_buffer % 256 == 0, "cs_pack_weights_create: pack_buffer must be 256 B aligned");
    int64_t need = 0;
    cs_pack_weights_bytes(cfg, &need);
    CS_CHECK_ARG(pack_bytes >= need, "cs_pack_weights_create: pack buffer too small (%lld < %lld bytes)", (long long)pack_bytes, (long long)need);
    cs_tower* t = new cs_tower();
    t->cfg = to_cfg(cfg);
    const char* e = getenv("CLIPSELF_NO_GRAPH");
    t->use_graphs = !(e != nullptr && e[0] != '\0' && e[0] != '0');
    {
        const char* f = getenv("CLIPSELF_FULL_LAST_BLOCK");
        t->cls_tail = !(f != nullptr && f[0] != '\0' && f[0] != '0');
    }
    Carver c{(uint8_t*)pack_buffer};
    carve(t, c);
    rc = pack(t, names, tensors_f32, count, (cudaStream_t)stream);
    if (rc) {
        delete t;
        return rc;
    }
    *out = t;
    return CS_OK;
}

extern "C" int cs_pack_weights_update(cs_tower_t* t, const char* const* names, const void* const* tensors_f32, int count, void* stream) {
    CS_CHECK_ARG(t && names && tensors_f32 && count > 0, "cs_pack_weights_update: null pointer");
    std::lock_guard<std::mutex> lock(t->mu);
    return pack(t, names, tensors_f32, count, (cudaStream_t)stream);       // same buffers: captured graphs stay valid
}

extern "C" int cs_pack_weights_destroy(cs_tower_t* t) {
    if (t == nullptr) return CS_OK;
    drop_graphs(t);
    if (t->capture_stream != nullptr) cudaStreamDestroy(t->capture_stream);
    for (auto& kv : t->rope_pos_other) cudaFree(kv.second);
    delete t;
    return CS_OK;
}

extern "C" int cs_query_workspace(const cs_tower_cfg_t* cfg, int chunk_images, int image_size, int64_t* bytes) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    CS_CHECK_ARG(bytes != nullptr && chunk_images > 0, "cs_query_workspace: bad argument");
    const Cfg g = to_cfg(cfg);
    if (image_size <= 0) image_size = g.image_size;
    CS_CHECK_ARG(image_size % g.patch == 0 && image_size / g.patch <= 64, "cs_query_workspace: bad image_size");
    *bytes = carve_ws(g, chunk_images, image_size / g.patch, nullptr, nullptr);
    return CS_OK;
}

extern "C" int cs_vit_forward_cls(cs_tower_t* t, const void* images, cs_dtype_t dtype, int n_images, void* workspace,
                                  int64_t workspace_bytes, int chunk_images, float* out, void* stream) {
    return forward(t, KIND_CLS, images, dtype, n_images, 0, nullptr, workspace, workspace_bytes, chunk_images, out, stream);
}

extern "C" int cs_vit_forward_dense(cs_tower_t* t, const void* images, cs_dtype_t dtype, int n_images, int image_size,
                                    const float* pos_embed_rescaled, void* workspace, int64_t workspace_bytes, int chunk_images,
                                    float* out_nhwc, void* stream) {
    return forward(t, KIND_DENSE, images, dtype, n_images, image_size, pos_embed_rescaled, workspace, workspace_bytes, chunk_images,
                   out_nhwc, stream);
}
