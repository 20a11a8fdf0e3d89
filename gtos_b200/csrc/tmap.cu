// TMA tensor-map construction.  cuTensorMapEncodeTiled is fetched through the runtime's
// driver-entry-point lookup so the library has no link-time dependency on libcuda (it must
// dlopen on a box without a driver for the "library loads and exports its symbols" check).
#include <mutex>

#include "common.cuh"

namespace gtos {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_once;

static void resolve_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  (void)cudaGetLastError();
}

int make_tmap_nd(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                 const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  std::call_once(g_once, resolve_encode);
  if (!g_encode) {
    set_error("cuTensorMapEncodeTiled not available (no CUDA driver?)");
    return GTOS_ERR_NO_DEVICE;
  }
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i];
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_error("TMA base address must be 16-byte aligned");
    return GTOS_ERR_ARG;
  }
  for (int i = 1; i < rank; ++i) {
    if (strides_bytes[i] % 16 != 0) {
      set_error("TMA stride %d = %llu bytes is not a multiple of 16", i, (unsigned long long)strides_bytes[i]);
      return GTOS_ERR_ARG;
    }
  }
  CUresult r = g_encode(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu, box %u %u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0],
              rank > 1 ? box[1] : 0);
    return GTOS_ERR_CUDA;
  }
  return GTOS_OK;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows) {
  uint64_t dims[2] = {cols, rows};
  uint64_t strides[2] = {0, ld_elems * 2};
  uint32_t box[2] = {64, box_rows};
  return make_tmap_nd(out, base, 2, 2, dims, strides, box, true);
}

}  // namespace gtos
