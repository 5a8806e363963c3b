// tc_host.cu -- host side of the TMA plumbing: driver entry point lookup + tensor-map encode.
#include <mutex>

#include "common.cuh"
#include "tc.cuh"

namespace sed {
namespace tc {

EncodeTiledFn encode_tiled_fn() {
  static std::once_flag once;
  static EncodeTiledFn fn = nullptr;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const char* what) {
  EncodeTiledFn enc = encode_tiled_fn();
  SED_REQUIRE(enc != nullptr, "%s: cuTensorMapEncodeTiled is not available from this driver", what);
  SED_REQUIRE(aligned(base, 16), "%s: TMA base pointer must be 16-byte aligned", what);
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    SED_REQUIRE(box[i] >= 1 && box[i] <= 256, "%s: TMA box dim %d = %u out of range", what, i, box[i]);
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    SED_REQUIRE(strides_bytes[i] % 16 == 0, "%s: TMA stride %d = %llu not a multiple of 16 B", what, i,
                (unsigned long long)strides_bytes[i]);
  }
  SED_REQUIRE(box[0] * 2 <= 128, "%s: inner box exceeds the 128 B swizzle span", what);
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SED_REQUIRE(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled failed with CUresult %d", what, (int)r);
  return 0;
}

}  // namespace tc
}  // namespace sed
