// Host-side tensor-map construction shared by the TMA kernels.
#include "tc_common.cuh"

#include <mutex>
#include <unordered_map>

namespace cs {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// Descriptor cache: a tensor map depends only on (address, geometry, box), so steady-state steps (same workspace
// buffers every step) never call into the driver.  Bounded; dropped wholesale when full.
namespace {
struct MapKey {
    const void* ptr;
    int64_t rows, cols, ld;
    int box_cols, box_rows;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_cols == o.box_cols && box_rows == o.box_rows;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        uint64_t h = (uint64_t)(uintptr_t)k.ptr * 0x9E3779B97F4A7C15ull;
        h ^= ((uint64_t)k.rows * 0xC2B2AE3D27D4EB4Full) ^ ((uint64_t)k.cols << 17) ^ ((uint64_t)k.ld << 41) ^
             ((uint64_t)k.box_cols << 7) ^ ((uint64_t)k.box_rows << 29);
        return (size_t)(h ^ (h >> 31));
    }
};
std::mutex g_map_mutex;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;
long long g_map_encodes = 0;
}  // namespace

long long tensor_map_encode_count() { return g_map_encodes; }

static int encode_map_bf16_2d(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                              int box_rows);

int make_map_bf16_2d(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                     int box_rows) {
    const MapKey key{ptr, rows, cols, ld, box_cols, box_rows};
    std::lock_guard<std::mutex> lock(g_map_mutex);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) {
        *map = it->second;
        return CS_OK;
    }
    const int rc = encode_map_bf16_2d(map, ptr, rows, cols, ld, box_cols, box_rows);
    if (rc != CS_OK) return rc;
    if (g_map_cache.size() >= 8192) g_map_cache.clear();
    g_map_cache.emplace(key, *map);
    ++g_map_encodes;
    return CS_OK;
}

// 2D bf16 row-major [rows, cols] with leading dimension ld; box = [box_rows, box_cols], 128B swizzle.
static int encode_map_bf16_2d(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                              int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
        return CS_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): ptr=%p rows=%lld cols=%lld ld=%lld", (int)r, ptr,
                  (long long)rows, (long long)cols, (long long)ld);
        return CS_ERR_CUDA;
    }
    return CS_OK;
}

int make_map_2d(CUtensorMap* map, const void* ptr, int elem_bytes, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                int box_rows, int swizzle_bytes) {
    // the key cannot collide with the bf16 SWIZZLE_128B maps: box_cols carries the element size and swizzle in its high bits
    const MapKey key{ptr, rows, cols, ld, box_cols | (elem_bytes << 16) | (swizzle_bytes << 20) | (1 << 30), box_rows};
    std::lock_guard<std::mutex> lock(g_map_mutex);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) {
        *map = it->second;
        return CS_OK;
    }
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
        return CS_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * (cuuint64_t)elem_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = fn(map, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr),
                    dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (2D, %d-byte elements) failed (%d): ptr=%p rows=%lld cols=%lld ld=%lld", elem_bytes, (int)r,
                  ptr, (long long)rows, (long long)cols, (long long)ld);
        return CS_ERR_CUDA;
    }
    if (g_map_cache.size() >= 8192) g_map_cache.clear();
    g_map_cache.emplace(key, *map);
    ++g_map_encodes;
    return CS_OK;
}

int make_map_bf16_3d(CUtensorMap* map, const void* ptr, int64_t cols, int64_t n1, int64_t n2, int box_cols, int box_rows) {
    // same cache as the 2D maps; the key cannot collide with a 2D one (ld < 0 marks the 3D geometry)
    const MapKey key{ptr, n2, cols, -n1, box_cols, box_rows};
    std::lock_guard<std::mutex> lock(g_map_mutex);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) {
        *map = it->second;
        return CS_OK;
    }
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
        return CS_ERR_CUDA;
    }
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)n1, (cuuint64_t)n2};
    cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)cols * 2 * (cuuint64_t)n1};
    cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : box_cols * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (3D) failed (%d): ptr=%p cols=%lld n1=%lld n2=%lld", (int)r, ptr, (long long)cols,
                  (long long)n1, (long long)n2);
        return CS_ERR_CUDA;
    }
    if (g_map_cache.size() >= 8192) g_map_cache.clear();
    g_map_cache.emplace(key, *map);
    ++g_map_encodes;
    return CS_OK;
}

}  // namespace cs

extern "C" int64_t cs_tensor_map_encodes(void) { return (int64_t)cs::tensor_map_encode_count(); }
