#include "common.cuh"
#include "tc.cuh"

#include <cudaTypedefs.h>
#include <stdlib.h>

namespace msm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

bool tc_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("MSM_DISABLE_TC");
    cached = (e != nullptr && e[0] != '\0' && e[0] != '0') ? 0 : 1;
  }
  return cached == 1;
}

bool pdl_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("MSM_DISABLE_PDL");
    cached = (e != nullptr && e[0] != '\0' && e[0] != '0') ? 0 : 1;
  }
  return cached == 1;
}

namespace tc {

// cuTensorMapEncodeTiled is a driver entry point; fetching it through the runtime keeps
// libmsmformer_b200.so free of a link-time libcuda dependency.
static PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  if (encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
      set_error("cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
      return nullptr;
    }
    encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  return encode;
}

int encode_tensor_map(CUtensorMap* map, TmapType type, TmapSwizzle swizzle, const void* base, int rank,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  PFN_cuTensorMapEncodeTiled_v12000 encode = tensor_map_encoder();
  if (encode == nullptr) return MSM_E_UNSUPPORTED;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUresult r = encode(map, type == TmapType::F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                      (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      swizzle == TmapSwizzle::B128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, inner dim %llu, box %u)", (int)r, rank,
              (unsigned long long)dims[0], box[0]);
    return MSM_E_BADARG;
  }
  return 0;
}

int encode_tensor_map_f32(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box) {
  return encode_tensor_map(map, TmapType::F32, TmapSwizzle::None, base, rank, dims, strides_bytes, box);
}

}  // namespace tc

}  // namespace msm

extern "C" int msm_abi_version(void) { return MSM_ABI_VERSION; }
extern "C" const char* msm_last_error(void) { return msm::g_err; }
extern "C" int msm_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return -1;
  return major * 10 + minor;
}

// EXPERIMENTAL (prefix msmx_, not in the public header; opt-in through MSM_L2_PERSIST=1): L2 persisting access window
// on [ptr, ptr + bytes) for the kernels launched on `stream` from now on (captured into graph kernel nodes). The mask
// features are re-read by every prediction head call of a step. bytes = 0 clears the window.
extern "C" int msmx_set_l2_persisting_window(const void* ptr, size_t bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaStreamAttrValue attr = {};
  if (bytes == 0 || ptr == nullptr) {
    attr.accessPolicyWindow.num_bytes = 0;
    MSM_CUDA(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr));
    return 0;
  }
  int dev = 0, max_window = 0, max_persist = 0;
  MSM_CUDA(cudaGetDevice(&dev));
  MSM_CUDA(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev));
  MSM_CUDA(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev));
  if (max_window <= 0 || max_persist <= 0) {
    msm::set_error("this device has no persisting L2 cache");
    return MSM_E_UNSUPPORTED;
  }
  static bool limit_set = false;
  if (!limit_set) {  // set-aside of the L2 for persisting lines: as much as the device allows
    MSM_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist));
    limit_set = true;
  }
  const size_t window = bytes < (size_t)max_window ? bytes : (size_t)max_window;
  attr.accessPolicyWindow.base_ptr = const_cast<void*>(ptr);
  attr.accessPolicyWindow.num_bytes = window;
  const float ratio = (float)((double)max_persist / (double)window);
  attr.accessPolicyWindow.hitRatio = ratio < 1.f ? ratio : 1.f;  // the fraction of the window that fits the set-aside
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  MSM_CUDA(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr));
  return 0;
}
