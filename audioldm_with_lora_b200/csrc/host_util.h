// Host-side helpers: error reporting for the C-ABI and TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

namespace b200 {

enum { B200_OK = 0, B200_ERR_ARG = -1, B200_ERR_CUDA = -2, B200_ERR_DRIVER = -3, B200_ERR_UNSUPPORTED = -4 };

inline char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
#define B200_CHECK_ARG(cond, ...) \
  do {                            \
    if (!(cond)) return ::b200::fail(::b200::B200_ERR_ARG, __VA_ARGS__); \
  } while (0)
#define B200_CHECK_LAUNCH(name)                                                                   \
  do {                                                                                            \
    cudaError_t e__ = cudaGetLastError();                                                         \
    if (e__ != cudaSuccess)                                                                       \
      return ::b200::fail(::b200::B200_ERR_CUDA, "%s: launch failed: %s", name, cudaGetErrorString(e__)); \
  } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// Encoded tensor maps are cached by VALUE of their inputs (base pointer, dims, strides, box, swizzle): a descriptor is a
// pure function of those, so a hit can never be stale.  Under CUDA-graph replay nothing is encoded at all; the cache is
// for the eager paths (the autograd seam of the fine-tuning step, B200AttnProcessor), where a step re-issues the same
// ~1000 launches on the same arena addresses and cuTensorMapEncodeTiled (~1.5 us, six per GEMM launch) would dominate
// the host side.  Process-wide, mutex-protected, bounded (cleared when full).
struct TmapKey {
  uint64_t v[12];
  bool operator==(const TmapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 12; ++i) {
      h ^= k.v[i];
      h *= 1099511628211ull;
    }
    return static_cast<size_t>(h);
  }
};
inline std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash>& tmap_cache() {
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> c;
  return c;
}
inline std::mutex& tmap_mutex() {
  static std::mutex m;
  return m;
}
inline unsigned long long& tmap_cache_hits() {
  static unsigned long long n = 0;
  return n;
}

inline int make_tmap_bf16_uncached(CUtensorMap* m, const void* base, int rank, const uint64_t* dims,
                                   const uint64_t* strides_el, const uint32_t* box, CUtensorMapSwizzle swz);

// bf16 tensor, `rank` dims (innermost first), strides in ELEMENTS for dims 1..rank-1.
inline int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_el,
                          const uint32_t* box, CUtensorMapSwizzle swz) {
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.v[0] = reinterpret_cast<uint64_t>(base);
  key.v[1] = (static_cast<uint64_t>(rank) << 32) | static_cast<uint64_t>(swz);
  for (int i = 0; i < rank && i < 4; ++i) {
    key.v[2 + i] = dims[i];
    key.v[6 + i] = i + 1 < rank ? strides_el[i] : 0;
  }
  key.v[10] = (static_cast<uint64_t>(box[0]) << 32) | (rank > 1 ? box[1] : 0);
  key.v[11] = (static_cast<uint64_t>(rank > 2 ? box[2] : 0) << 32) | (rank > 3 ? box[3] : 0);
  {
    std::lock_guard<std::mutex> lock(tmap_mutex());
    auto it = tmap_cache().find(key);
    if (it != tmap_cache().end()) {
      *m = it->second;
      ++tmap_cache_hits();
      return B200_OK;
    }
  }
  const int rc = make_tmap_bf16_uncached(m, base, rank, dims, strides_el, box, swz);
  if (rc == B200_OK) {
    std::lock_guard<std::mutex> lock(tmap_mutex());
    if (tmap_cache().size() > 65536) tmap_cache().clear();
    tmap_cache().emplace(key, *m);
  }
  return rc;
}

// Same with traversal (element) strides per dimension: every es[i]-th element of the box is fetched (Downsample2D's
// stride-2 convolution reads every second pixel).  Not cached: three launches per step, all under graph replay.
inline int make_tmap_bf16_strided(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_el,
                                  const uint32_t* box, const uint32_t* es_in, CUtensorMapSwizzle swz) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return fail(B200_ERR_DRIVER, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = es_in[i];
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_el[i] * 2;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(B200_ERR_ARG, "TMA base not 16B aligned");
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(B200_ERR_DRIVER, "cuTensorMapEncodeTiled (strided) failed: %d (box %u,%u,%u,%u)", (int)r, box[0], rank > 1 ? box[1] : 0,
                rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  return B200_OK;
}

inline int make_tmap_bf16_uncached(CUtensorMap* m, const void* base, int rank, const uint64_t* dims,
                                   const uint64_t* strides_el, const uint32_t* box, CUtensorMapSwizzle swz) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return fail(B200_ERR_DRIVER, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_el[i] * 2;   // bytes, for dims 1..rank-1
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(B200_ERR_ARG, "TMA base not 16B aligned");
  for (int i = 0; i + 1 < rank; ++i)
    if (gstr[i] % 16 != 0) return fail(B200_ERR_ARG, "TMA stride %d (%llu B) not multiple of 16", i, (unsigned long long)gstr[i]);
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(B200_ERR_DRIVER, "cuTensorMapEncodeTiled failed: %d (rank %d dims %llu,%llu,%llu,%llu box %u,%u,%u,%u)",
                (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  }
  return B200_OK;
}


// Launch with the programmatic-stream-serialization attribute (PDL): the kernel may start while its
// predecessor in the stream drains; kernels call pdl_wait() before touching dependent global memory.
// Captured into CUDA graphs as programmatic dependency edges.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  static const bool no_pdl = getenv("B200_NO_PDL") != nullptr;      // debugging knob
  if (!no_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster_x > 0) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster_x;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define B200_CHECK_PDL(name, expr)                                                                     \
  do {                                                                                                 \
    cudaError_t e__ = (expr);                                                                          \
    if (e__ != cudaSuccess)                                                                            \
      return ::b200::fail(::b200::B200_ERR_CUDA, "%s: launch failed: %s", name, cudaGetErrorString(e__)); \
  } while (0)

}  // namespace b200
