// C-ABI plumbing: error state, device info, GEMM engine dispatch.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace satk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int gemm_simt_launch(const satk_gemm_desc* d, cudaStream_t st);
int gemm_tc_launch(const satk_gemm_desc* d, cudaStream_t st, bool* supported);
int lstm_max_clusters_h256();
namespace arnn { int attn_rnn_max_clusters(); }

}  // namespace satk

extern "C" {

const char* satk_last_error(void) { return satk::g_err; }
int satk_version(void) { return 100; }

int satk_device_info(int* out5) {
  int dev = 0;
  SATK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  SATK_CUDA(cudaGetDeviceProperties(&p, dev));
  out5[0] = p.multiProcessorCount;
  out5[1] = p.major;
  out5[2] = p.minor;
  out5[3] = satk::arnn::attn_rnn_max_clusters();
  out5[4] = satk::lstm_max_clusters_h256();
  return 0;
}

int satk_struct_sizes(int* out5) {
  out5[0] = (int)sizeof(satk_gemm_desc);
  out5[1] = (int)sizeof(satk_lstm_fwd_desc);
  out5[2] = (int)sizeof(satk_lstm_bwd_desc);
  out5[3] = (int)sizeof(satk_attn_rnn_fwd_desc);
  out5[4] = (int)sizeof(satk_attn_rnn_bwd_desc);
  return 0;
}

int satk_struct_sizes_decode(int* out4) {
  out4[4] = (int)sizeof(satk_mlp_chain_desc);
  out4[0] = (int)sizeof(satk_rowgemm_desc);
  out4[1] = (int)sizeof(satk_attn_step_desc);
  out4[2] = (int)sizeof(satk_sa_step_desc);
  out4[3] = (int)sizeof(satk_sa_tail_desc);
  return 0;
}

int satk_gemm(const satk_gemm_desc* d, int engine, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (d->zcoord && engine == 1) {
    satk::set_error("satk_gemm: the zcoord form exists on the tcgen05 tile only");
    return SATK_ERR_UNSUPPORTED;
  }
  if (engine == 1) return satk::gemm_simt_launch(d, st);
  bool supported = false;
  int rc = satk::gemm_tc_launch(d, st, &supported);
  if (supported || rc != SATK_OK) return rc;
  if (engine == 2 || d->zcoord) {
    satk::set_error("satk_gemm: shape not supported by the tcgen05 tile (M=%d N=%d K=%d tA=%d tB=%d taps=%d)", d->M, d->N, d->K,
                    d->transA, d->transB, d->taps);
    return SATK_ERR_UNSUPPORTED;
  }
  return satk::gemm_simt_launch(d, st);
}

}  // extern "C"
