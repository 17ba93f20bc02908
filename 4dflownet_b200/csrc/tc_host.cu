// Per-device host caches for the launch paths (see tc_host.h).
#include "tc_host.h"

#include <map>
#include <mutex>
#include <set>
#include <tuple>
#include <utility>

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

std::mutex g_mu;
std::map<int, int> g_sms;                                   // device -> SM count
std::set<std::pair<int, const void*>> g_attr;               // (device, func) whose smem attribute is set
typedef std::tuple<int, const void*, int, int, int, int, int> MapKey;
std::map<MapKey, CUtensorMap> g_maps;

int cur_dev() {
    int d = 0;
    cudaGetDevice(&d);
    return d;
}

}  // namespace

int tc_num_sms() {
    const int d = cur_dev();
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_sms.find(d);
    if (it != g_sms.end()) return it->second;
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d);
    g_sms[d] = n;
    return n;
}

cudaError_t tc_func_smem(const void* func, int bytes) {
    const int d = cur_dev();
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_attr.count({d, func})) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) g_attr.insert({d, func});
    return e;
}

bool tc_make_act_map(CUtensorMap* map, const __half* base, int B, int E, int by, int bz, int bx) {
    const MapKey key(cur_dev(), base, B, E, by, bz, bx);
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *map = it->second; return true; }
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[5] = {64, (cuuint64_t)E, (cuuint64_t)E, (cuuint64_t)E, (cuuint64_t)(2 * B)};
    cuuint64_t strides[4] = {128, (cuuint64_t)128 * E, (cuuint64_t)128 * E * E, (cuuint64_t)128 * E * E * E};
    cuuint32_t box[5] = {64, (cuuint32_t)bz, (cuuint32_t)by, (cuuint32_t)bx, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    if (enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<__half*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    if (g_maps.size() > 4096) g_maps.clear();     // bound the cache (single-layer test entry points allocate per call)
    g_maps[key] = *map;
    return true;
}

bool tc_make_rows_map(CUtensorMap* map, const __half* base, long long nrows, long long plane_elems, int box_rows) {
    // cached under the same key type: (device, base, B := nrows low bits, E := nrows high bits, by := box_rows, bz := -3, bx := plane stride hash)
    const MapKey key(cur_dev(), base, (int)(nrows & 0x7fffffff), (int)(nrows >> 31), box_rows, -3, (int)(plane_elems & 0x7fffffff));
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *map = it->second; return true; }
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[3] = {64, (cuuint64_t)nrows, 2};
    cuuint64_t strides[2] = {128, (cuuint64_t)plane_elems * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps[key] = *map;
    return true;
}

bool tc_make_gplanar_map(CUtensorMap* map, const float* base, int B, int D, int by) {
    const MapKey key(cur_dev(), base, B, D, by, -7, 0);
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *map = it->second; return true; }
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    const cuuint64_t d = (cuuint64_t)D;
    cuuint64_t dims[5] = {d, d, d, (cuuint64_t)B, 3};
    cuuint64_t strides[4] = {4 * d, 4 * d * d, 4 * d * d * d, 4 * d * d * d * (cuuint64_t)B};
    cuuint32_t box[5] = {(cuuint32_t)(D + 8), (cuuint32_t)by, 3, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    if (enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps[key] = *map;
    return true;
}

void tc_forget_maps(const void* base, size_t bytes) {
    const char* lo = static_cast<const char*>(base);
    const char* hi = lo + bytes;
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto it = g_maps.begin(); it != g_maps.end();) {
        const char* b = static_cast<const char*>(std::get<1>(it->first));
        if (b >= lo && b < hi) it = g_maps.erase(it);
        else ++it;
    }
}
