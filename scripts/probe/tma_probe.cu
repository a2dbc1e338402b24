// Probe which TMA 5-D box configurations are legal: ./tma_probe width pitch D C bx by bz bc x0 y0 z0
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../nvfpcc_b200/csrc/nvf_tma.cuh"
using namespace nvf;
__global__ void k(const __grid_constant__ CUtensorMap tmap, float* out, int n_floats, int x0, int y0, int z0) {
  extern __shared__ __align__(128) float smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (n_floats + 31) / 32 * 32);
  if (threadIdx.x == 0) { tma::mbar_init(bar, 1); tma::fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    tma::mbar_arrive_expect_tx(bar, n_floats * 4);
    tma::load_5d(smem, &tmap, bar, x0, y0, z0, 0, 0);
  }
  tma::mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < n_floats; i += blockDim.x) out[i] = smem[i];
}
int main(int argc, char** argv) {
  int width = atoi(argv[1]), pitch = atoi(argv[2]), D = atoi(argv[3]), C = atoi(argv[4]);
  int bx = atoi(argv[5]), by = atoi(argv[6]), bz = atoi(argv[7]), bc = atoi(argv[8]);
  int x0 = atoi(argv[9]), y0 = atoi(argv[10]), z0 = atoi(argv[11]);
  size_t n = (size_t)C * D * D * pitch;
  std::vector<float> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = (float)(i % 1000) + 1.f;
  float *d, *o;
  cudaMalloc(&d, n * 4); cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
  int nf = bx * by * bz * bc;
  cudaMalloc(&o, nf * 4);
  CUtensorMap map;
  if (!tma::make_map_5d(&map, d, 1, C, D, width, pitch, bx, by, bz, bc)) { printf("ENCODE FAILED\n"); return 2; }
  int smem = ((nf + 31) / 32 * 32) * 4 + 16;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<<<1, 128, smem>>>(map, o, nf, x0, y0, z0);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("KERNEL FAILED: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> r(nf);
  cudaMemcpy(r.data(), o, nf * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int c = 0; c < bc; ++c) for (int z = 0; z < bz; ++z) for (int y = 0; y < by; ++y) for (int x = 0; x < bx; ++x) {
    int gx = x0 + x, gy = y0 + y, gz = z0 + z;
    float want = 0.f;
    if (gx >= 0 && gx < width && gy >= 0 && gy < D && gz >= 0 && gz < D) want = h[((size_t)(c * D + gz) * D + gy) * pitch + gx];
    if (r[((c * bz + z) * by + y) * bx + x] != want) ++bad;
  }
  printf("OK mismatches=%d\n", bad);
  return bad ? 3 : 0;
}
