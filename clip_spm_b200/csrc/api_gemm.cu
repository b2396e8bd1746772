// C-ABI: error string + the stand-alone GEMM entry point (used by the parity tests to pin the
// tcgen05 kernel against torch.matmul before anything is stacked on top of it).
#include "../../include/clipspm_b200.h"
#include "api_common.cuh"
#include "gemm.cuh"

namespace spm {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* get_error() { return g_err.c_str(); }

int device_sm_count(int* out) {
  int dev = 0;
  SPM_CUDA(cudaGetDevice(&dev));
  SPM_CUDA(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
  return 0;
}
}  // namespace spm

extern "C" {

const char* spm_last_error(void) { return spm::get_error(); }

int spm_abi_version(void) { return SPM_ABI_VERSION; }

int spm_gemm(void* stream, int kind, const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
             const float* bias, int act, float slope, const float* residual, int ldr, int res_row_mod,
             int res_row_off, int out_row_group, int out_group_stride, int out_row_off, void* out, int ldo,
             int out_bf16) {
  static bool inited = false;
  const char* err = "";
  if (!inited) {
    if (spm::gemm_init(&err)) { spm::set_error(err); return 1; }
    inited = true;
  }
  int sms = 0;
  SPM_TRY(spm::device_sm_count(&sms));
  spm::GemmEpilogue ep;
  ep.bias = bias; ep.act = act; ep.slope = slope;
  ep.residual = residual; ep.ldr = ldr; ep.res_row_mod = res_row_mod; ep.res_row_off = res_row_off;
  ep.out_row_group = out_row_group; ep.out_group_stride = out_group_stride; ep.out_row_off = out_row_off;
  ep.out = out; ep.ldo = ldo; ep.out_bf16 = out_bf16;
  spm::GemmOp op;
  if (spm::gemm_plan(&op, kind, A, lda, B, ldb, M, N, K, ep, sms, &err)) { spm::set_error(err); return 1; }
  if (spm::gemm_run(&op, (cudaStream_t)stream, &err)) { spm::set_error(err); return 1; }
  return 0;
}

}  // extern "C"
