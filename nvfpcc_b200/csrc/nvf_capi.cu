// C ABI (include/nvf_b200.h) + CUDA launchers for sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <atomic>

#include "nvf_api_impl.h"

using namespace nvf;

namespace {

thread_local int g_last_cuda = 0;
std::atomic<long long> g_launches{0};  // kernels launched by this library (bench.py: gpu_launches)

template <class TS>
struct DevEnv {
  TS ts;
  template <class F>
  __device__ __forceinline__ void phase(F&& f) {
    f((int)threadIdx.x, ts);
    __syncthreads();
  }
};

// ------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(kThreads) k_pack(PackParams p) {
  pack_thread(p, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

__global__ void __launch_bounds__(kThreads, 1) k_decode_fused_A(FusedAParams p) {
  extern __shared__ __align__(16) float smem[];
  DevEnv<FusedATS> env;
  FusedABlock<DevEnv<FusedATS>>::run(env, p, smem, blockIdx.x, gridDim.x);
}

__global__ void __launch_bounds__(kThreads) k_scan_counts(EmitParams p) {
  __shared__ int64_t sm[2 * kThreads];
  DevEnv<int> env;
  ScanBlock<DevEnv<int>>::run(env, p, sm);
}

__global__ void __launch_bounds__(kThreads) k_emit_coords(EmitParams p) {
  __shared__ int sm[2 * kThreads];
  DevEnv<int> env;
  EmitBlock<DevEnv<int>>::run(env, p, sm, blockIdx.x);
}

template <class K>
__global__ void __launch_bounds__(kThreads) k_generic(typename K::Params p) {
  K::thread(p, blockIdx.x, threadIdx.x, gridDim.x);
}

template <int COT, int K>
__global__ void __launch_bounds__(kThreads) k_wgrad(WgradParams p) {
  extern __shared__ __align__(16) float smem[];
  DevEnv<int> env;
  WgradBlock<COT, K>::run(env, p, smem, blockIdx.x);
}

__global__ void __launch_bounds__(kThreads) k_chansum(ChanSumParams p) {
  __shared__ float sm[kThreads];
  DevEnv<int> env;
  ChanSumBlock::run(env, p, sm, blockIdx.x);
}

__global__ void __launch_bounds__(kThreads) k_mask(MaskParams p) {
  __shared__ int sm[kThreads];
  DevEnv<int> env;
  MaskBlock::run(env, p, sm, blockIdx.x);
}

__global__ void __launch_bounds__(kThreads) k_loss(LossParams p) {
  __shared__ double sm[kThreads * NVF_LOSS_SUMS];
  DevEnv<int> env;
  LossBlock::run(env, p, sm, blockIdx.x);
}

// FFMA throughput probes.  variant 0: scalar FFMA, 8x8 register tile (the shape
// of the conv inner loops); variant 1: packed fma.rn.f32x2 on the same tile.
__global__ void __launch_bounds__(kThreads) k_ffma(int variant, long long iters, float* sink) {
  float w[8], a[8];
  for (int i = 0; i < 8; ++i) {
    w[i] = 1.0f + 1e-6f * (threadIdx.x + i);
    a[i] = 0.5f + 1e-6f * (blockIdx.x + i);
  }
  if (variant == 0) {
    float acc[8][8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[c][j] = 0.f;
    for (long long it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[c][j] = fmaf(w[c], a[(j + r) & 7], acc[c][j]);
      }
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) s += acc[c][j];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else {
    unsigned long long acc[8][4], ww[8], aa[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      asm("mov.b64 %0, {%1, %1};" : "=l"(ww[c]) : "f"(w[c]));
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[c][j] = 0ull;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) asm("mov.b64 %0, {%1, %2};" : "=l"(aa[j]) : "f"(a[j]), "f"(a[(j + 1) & 7]));
    for (long long it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[c][j]) : "l"(ww[c]), "l"(aa[(2 * j + r) & 7]));
      }
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[c][j]));
        s += lo + hi;
      }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }
}

// ------------------------------------------------------------------ launcher
struct DevLauncher {
  cudaStream_t st;
  int n_sms = 0;
  int rc = NVF_OK;

  int init() {
    int dev = 0, major = 0;
    if (!chk(cudaGetDevice(&dev))) return rc;
    if (!chk(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev))) return rc;
    if (!chk(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev))) return rc;
    if (major != 10) rc = NVF_ERR_NO_DEVICE;
    return rc;
  }
  bool chk(cudaError_t e) {
    if (e != cudaSuccess) {
      g_last_cuda = (int)e;
      if (rc == NVF_OK) rc = (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? NVF_ERR_NO_DEVICE : NVF_ERR_CUDA;
      return false;
    }
    return true;
  }
  void post() {
    chk(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  int sms() const { return n_sms; }
  int error() const { return rc; }

  void pack(const PackParams& p) { k_pack<<<32, kThreads, 0, st>>>(p); post(); }
  void fusedA(const FusedAParams& p, int grid) {
    static bool attr_set = false;
    if (!attr_set) {
      if (!chk(cudaFuncSetAttribute(k_decode_fused_A, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedA::SMEM_BYTES)))
        return;
      attr_set = true;
    }
    k_decode_fused_A<<<grid, kThreads, FusedA::SMEM_BYTES, st>>>(p);
    post();
  }
  void scan(const EmitParams& p) { k_scan_counts<<<1, kThreads, 0, st>>>(p); post(); }
  void emit(const EmitParams& p, int grid) { k_emit_coords<<<grid, kThreads, 0, st>>>(p); post(); }
  template <int COT>
  void layer(const LayerParams& p) {
    const int64_t XT = (p.Dout + 3) >> 2;
    const int64_t total = (int64_t)(p.CO / COT) * p.Dout * p.Dout * XT * p.n;
    int64_t grid = (total + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)n_sms * 32;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    k_generic<LayerKernel<COT>><<<(unsigned)grid, kThreads, 0, st>>>(p);
    post();
  }
  template <int COT, int K>
  void wgrad(const WgradParams& p, int grid) {
    const int smem = kThreads * COT * K * (int)sizeof(float);
    k_wgrad<COT, K><<<grid, kThreads, smem, st>>>(p);
    post();
  }
  void chansum(const ChanSumParams& p, int grid) { k_chansum<<<grid, kThreads, 0, st>>>(p); post(); }
  template <class K>
  void generic(const typename K::Params& p, int grid) {
    k_generic<K><<<grid < 1 ? 1 : grid, kThreads, 0, st>>>(p);
    post();
  }
  void mask(const MaskParams& p, int grid) { k_mask<<<grid, kThreads, 0, st>>>(p); post(); }
  void loss(const LossParams& p, int grid) { k_loss<<<grid, kThreads, 0, st>>>(p); post(); }
};

}  // namespace

// ------------------------------------------------------------------ C ABI
extern "C" {

int nvf_abi_version(void) { return NVF_ABI_VERSION; }

const char* nvf_strerror(int code) {
  switch (code) {
    case NVF_OK: return "ok";
    case NVF_ERR_INVALID_ARG: return "invalid argument";
    case NVF_ERR_UNSUPPORTED: return "unsupported channel configuration";
    case NVF_ERR_WORKSPACE: return "workspace too small";
    case NVF_ERR_CUDA: return "CUDA runtime error";
    case NVF_ERR_NO_DEVICE: return "no sm_100 device";
    default: return "unknown error";
  }
}

int nvf_last_cuda_error(void) { return g_last_cuda; }

long long nvf_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int nvf_has_fused_decode(const NvfDesc* desc) { return desc && is_cfg_A(*desc) ? 1 : 0; }

int nvf_workspace_bytes(const NvfDesc* desc, int64_t n_blocks, int mode, size_t* bytes_out) {
  if (!desc || !bytes_out || n_blocks < 0) return NVF_ERR_INVALID_ARG;
  if (!generic_supported(*desc)) return NVF_ERR_UNSUPPORTED;
  if (mode == NVF_MODE_DECODE) *bytes_out = DecodeWs::make(*desc, n_blocks).total;
  else if (mode == NVF_MODE_TRAIN) *bytes_out = TrainWs::make(*desc, n_blocks).total;
  else return NVF_ERR_INVALID_ARG;
  return NVF_OK;
}

int nvf_decode(const NvfDesc* desc, const NvfWeights* w, const float* latent, const int32_t* origins,
               int64_t n_blocks, float thh, float* prob_out, uint32_t* mask_out, int32_t* counts_out,
               int32_t* coords_out, int64_t cap, int64_t* total_out, void* workspace, size_t workspace_bytes,
               void* stream) {
  DevLauncher l{(cudaStream_t)stream};
  if (l.init() != NVF_OK) return l.rc;
  return Api<DevLauncher>::decode(l, desc, w, latent, origins, n_blocks, thh, prob_out, mask_out, counts_out,
                                  coords_out, cap, total_out, workspace, workspace_bytes);
}

int nvf_emit_points(const uint32_t* mask, const int32_t* counts, const int32_t* origins, int64_t n_blocks,
                    int32_t* coords_out, int64_t cap, int64_t* total_out, void* workspace, size_t workspace_bytes,
                    void* stream) {
  DevLauncher l{(cudaStream_t)stream};
  if (l.init() != NVF_OK) return l.rc;
  return Api<DevLauncher>::emit_points(l, mask, counts, origins, n_blocks, coords_out, cap, total_out, workspace,
                              workspace_bytes);
}

int nvf_train_forward(const NvfDesc* desc, const NvfWeights* w, const float* latent, int64_t n_blocks, float* out,
                      float* cls1, float* cls0, void* workspace, size_t workspace_bytes, void* stream) {
  DevLauncher l{(cudaStream_t)stream};
  if (l.init() != NVF_OK) return l.rc;
  return Api<DevLauncher>::train_forward(l, desc, w, latent, n_blocks, out, cls1, cls0, workspace, workspace_bytes);
}

int nvf_loss_seeds(const float* out, const float* cls1, const float* cls0, const float* gt, const float* dist,
                   int64_t n_blocks, float alpha_main, float alpha_aux, float thh_metric, double* sums_out,
                   float* g_out, float* g_cls1, float* g_cls0, void* workspace, size_t workspace_bytes,
                   void* stream) {
  DevLauncher l{(cudaStream_t)stream};
  if (l.init() != NVF_OK) return l.rc;
  return Api<DevLauncher>::loss_seeds(l, out, cls1, cls0, gt, dist, n_blocks, alpha_main, alpha_aux, thh_metric,
                                      sums_out, g_out, g_cls1, g_cls0, workspace, workspace_bytes);
}

int nvf_train_backward(const NvfDesc* desc, const NvfWeights* w, const float* latent, int64_t n_blocks,
                       const float* g_out, const float* g_cls1, const float* g_cls0, int flags,
                       const NvfWeightGrads* gw, float* g_latent, void* workspace, size_t workspace_bytes,
                       void* stream) {
  DevLauncher l{(cudaStream_t)stream};
  if (l.init() != NVF_OK) return l.rc;
  return Api<DevLauncher>::train_backward(l, desc, w, latent, n_blocks, g_out, g_cls1, g_cls0, flags, gw, g_latent,
                                          workspace, workspace_bytes);
}

int nvf_ffma_microbench(int variant, int64_t iters, float* sink, double* flops_out, void* stream) {
  if (!sink || iters <= 0 || variant < 0 || variant > 1) return NVF_ERR_INVALID_ARG;
  DevLauncher l{(cudaStream_t)stream};
  if (l.init() != NVF_OK) return l.rc;
  const int grid = l.n_sms * 4;
  k_ffma<<<grid, kThreads, 0, l.st>>>(variant, (long long)iters, sink);
  l.post();
  if (flops_out) *flops_out = 2.0 * 256.0 * (double)iters * (double)grid * kThreads;
  return l.rc;
}

}  // extern "C"
