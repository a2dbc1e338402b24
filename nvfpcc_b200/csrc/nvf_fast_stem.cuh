// The "stem" of the NVF decoder - everything at <= 8^3 resolution - as ONE CTA
// per leaf block, activations in shared memory:
//   forward   latent -> up0 (convT k5 s2 p2 op1) -> IGDN -> conv0 (convT k5 s2 p2 op1, ReLU) -> conv0_cls + sigmoid
//             (utils/network.py:4759-4761, gdn_3d.py:153-157)
//   backward  d(conv0 pre-activation) -> conv0 wgrad/bias/dgrad -> IGDN backward (dx, dbeta, dgamma)
//             -> up0 wgrad/bias -> d_latent                       
// These layers hold < 0.5 % of the MACs; as separate grid-wide launches they were
// latency-bound (a handful of CTAs each).  Per-block weight-gradient partials are
// summed over blocks in fixed order by k_reduce_partials.
#pragma once
#include <cuda_runtime.h>
#include "nvf_common.h"

namespace nvf {
namespace fast {

constexpr int kStemThreads = 512;

struct StemFwdParams {
  const float* latent;  // [n][CH][8]
  const float* up0_wp;  // PACK_CONVT_FWD [ci][k][co]  (CH x 125 x C0)
  const float* up0_b;
  const float* beta;    // [C0]
  const float* gamma;   // [C0][C0]
  const float* conv0_wp;// PACK_CONVT_FWD [ci][k][co]  (C0 x 125 x C1)
  const float* conv0_b;
  const float* cls0_wp; // [C1][27] or null (decode: no aux head)
  const float* cls0_b;
  float* x0;            // [n][C0][64]
  float* a0;            // [n][C0][64]
  float* a1;            // [n][C1][512]
  float* cls0;          // [n][512] or null
  float* cls0_copy;     // [n][512] or null
  int32_t n, CH, C0, C1;
};

// convT k5 s2 p2 op1 taps of output coordinate o along one axis: k = (o & 1) + 2t, input (o + 2 - k) / 2
__device__ __forceinline__ int stem_tap_lo(int o) { return o & 1; }

__global__ void __launch_bounds__(kStemThreads) k_stem_fwd(StemFwdParams p) {
  extern __shared__ __align__(16) float smem[];
  const int CH = p.CH, C0 = p.C0, C1 = p.C1;
  float* s_lat = smem;                 // [CH][8]
  float* s_x0 = s_lat + CH * 8;        // [C0][64]
  float* s_a0 = s_x0 + C0 * 64;        // [C0][64]
  float* s_a1 = s_a0 + C0 * 64;        // [C1][512]
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid < CH * 8) s_lat[tid] = p.latent[(size_t)b * CH * 8 + tid];
  __syncthreads();
  // up0: CH x 2^3 -> C0 x 4^3
  for (int idx = tid; idx < C0 * 64; idx += kStemThreads) {
    const int co = idx >> 6, z = (idx >> 4) & 3, y = (idx >> 2) & 3, x = idx & 3;
    float v = p.up0_b[co];
    for (int ci = 0; ci < CH; ++ci)
      for (int kz = z & 1; kz < 5; kz += 2) {
        const int tz = z + 2 - kz;
        if (tz < 0 || tz >= 4) continue;
        for (int ky = y & 1; ky < 5; ky += 2) {
          const int ty = y + 2 - ky;
          if (ty < 0 || ty >= 4) continue;
          for (int kx = x & 1; kx < 5; kx += 2) {
            const int tx = x + 2 - kx;
            if (tx < 0 || tx >= 4) continue;
            v = fmaf(s_lat[ci * 8 + (tz >> 1) * 4 + (ty >> 1) * 2 + (tx >> 1)],
                     __ldg(p.up0_wp + ((size_t)ci * 125 + kz * 25 + ky * 5 + kx) * C0 + co), v);
          }
        }
      }
    s_x0[idx] = v;
    p.x0[(size_t)b * C0 * 64 + idx] = v;
  }
  __syncthreads();
  // IGDN
  for (int idx = tid; idx < C0 * 64; idx += kStemThreads) {
    const int c = idx >> 6, pos = idx & 63;
    float nn = p.beta[c];
    for (int j = 0; j < C0; ++j) {
      const float xj = s_x0[j * 64 + pos];
      nn = fmaf(p.gamma[c * C0 + j], xj * xj, nn);
    }
    const float v = s_x0[idx] * sqrtf(nn);
    s_a0[idx] = v;
    p.a0[(size_t)b * C0 * 64 + idx] = v;
  }
  __syncthreads();
  // conv0: C0 x 4^3 -> C1 x 8^3, ReLU.  item = (position, group of 8 output channels)
  const int groups = C1 / 8;
  for (int item = tid; item < 512 * groups; item += kStemThreads) {
    const int pos = item & 511, cg = item >> 9;
    const int z = pos >> 6, y = (pos >> 3) & 7, x = pos & 7;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    for (int kz = z & 1; kz < 5; kz += 2) {
      const int tz = z + 2 - kz;
      if (tz < 0 || tz >= 8) continue;
      for (int ky = y & 1; ky < 5; ky += 2) {
        const int ty = y + 2 - ky;
        if (ty < 0 || ty >= 8) continue;
        for (int kx = x & 1; kx < 5; kx += 2) {
          const int tx = x + 2 - kx;
          if (tx < 0 || tx >= 8) continue;
          const int ipos = (tz >> 1) * 16 + (ty >> 1) * 4 + (tx >> 1);
          const float* wk = p.conv0_wp + (size_t)(kz * 25 + ky * 5 + kx) * C1 + cg * 8;
          for (int ci = 0; ci < C0; ++ci) {
            const float a = s_a0[ci * 64 + ipos];
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(wk + (size_t)ci * 125 * C1));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(wk + (size_t)ci * 125 * C1) + 1);
            acc[0] = fmaf(a, w0.x, acc[0]); acc[1] = fmaf(a, w0.y, acc[1]);
            acc[2] = fmaf(a, w0.z, acc[2]); acc[3] = fmaf(a, w0.w, acc[3]);
            acc[4] = fmaf(a, w1.x, acc[4]); acc[5] = fmaf(a, w1.y, acc[5]);
            acc[6] = fmaf(a, w1.z, acc[6]); acc[7] = fmaf(a, w1.w, acc[7]);
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float v = acc[c] + p.conv0_b[cg * 8 + c];
      s_a1[(cg * 8 + c) * 512 + pos] = v > 0.f ? v : 0.f;
    }
  }
  __syncthreads();
  {
    const float4* src = reinterpret_cast<const float4*>(s_a1);
    float4* dst = reinterpret_cast<float4*>(p.a1 + (size_t)b * C1 * 512);
    for (int i = tid; i < C1 * 128; i += kStemThreads) dst[i] = src[i];
  }
  // conv0_cls (k3 p1) + sigmoid
  if (p.cls0_wp && p.cls0) {
    for (int pos = tid; pos < 512; pos += kStemThreads) {
      const int z = pos >> 6, y = (pos >> 3) & 7, x = pos & 7;
      float v = p.cls0_b[0];
      for (int ci = 0; ci < C1; ++ci)
        for (int kz = 0; kz < 3; ++kz) {
          const int iz = z + kz - 1;
          if (iz < 0 || iz >= 8) continue;
          for (int ky = 0; ky < 3; ++ky) {
            const int iy = y + ky - 1;
            if (iy < 0 || iy >= 8) continue;
            for (int kx = 0; kx < 3; ++kx) {
              const int ix = x + kx - 1;
              if (ix < 0 || ix >= 8) continue;
              v = fmaf(s_a1[ci * 512 + iz * 64 + iy * 8 + ix], __ldg(p.cls0_wp + ci * 27 + kz * 9 + ky * 3 + kx), v);
            }
          }
        }
      const float pr = sigmoidf(v);
      p.cls0[(size_t)b * 512 + pos] = pr;
      if (p.cls0_copy) p.cls0_copy[(size_t)b * 512 + pos] = pr;
    }
  }
}

struct StemBwdParams {
  const float* latent;   // [n][CH][8]
  const float* x0;       // [n][C0][64]
  const float* a0;       // [n][C0][64]
  const float* g1;       // [n][C1][512]  dL/d(conv0 pre-activation)
  const float* beta;     // [C0]
  const float* gamma;    // [C0][C0]
  const float* conv0_w;  // raw (C0, C1, 5,5,5)
  const float* up0_w;    // raw (CH, C0, 5,5,5)
  float* partial;        // [n][stem_partial_floats] or null (no weight gradients)
  float* g_latent;       // [n][CH][8] or null
  int32_t n, CH, C0, C1, need_w;
};
// per-block partial layout: conv0_w | conv0_b | igdn_gamma | igdn_beta | up0_w | up0_b
__host__ __device__ inline int stem_partial_floats(int CH, int C0, int C1) {
  return C0 * C1 * 125 + C1 + C0 * C0 + C0 + CH * C0 * 125 + C0;
}

__global__ void __launch_bounds__(kStemThreads) k_stem_bwd(StemBwdParams p) {
  extern __shared__ __align__(16) float smem[];
  const int CH = p.CH, C0 = p.C0, C1 = p.C1;
  float* s_g1 = smem;                   // [C1][512]
  float* s_a0 = s_g1 + C1 * 512;        // [C0][64]
  float* s_x0 = s_a0 + C0 * 64;         // [C0][64]
  float* s_gy = s_x0 + C0 * 64;         // [C0][64]  dL/d(IGDN out)
  float* s_gx = s_gy + C0 * 64;         // [C0][64]  dL/d(up0 out)
  float* s_n = s_gx + C0 * 64;          // [C0][64]  IGDN norm
  float* s_lat = s_n + C0 * 64;         // [CH][8]
  float* s_part = s_lat + CH * 8;       // [4][C0*64] co-split partial sums of the conv0 data gradient
  const int b = blockIdx.x, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  {
    const float4* src = reinterpret_cast<const float4*>(p.g1 + (size_t)b * C1 * 512);
    for (int i = tid; i < C1 * 128; i += kStemThreads) reinterpret_cast<float4*>(s_g1)[i] = __ldg(src + i);
    for (int i = tid; i < C0 * 64; i += kStemThreads) {
      s_a0[i] = p.a0[(size_t)b * C0 * 64 + i];
      s_x0[i] = p.x0[(size_t)b * C0 * 64 + i];
    }
    if (tid < CH * 8) s_lat[tid] = p.latent[(size_t)b * CH * 8 + tid];
  }
  __syncthreads();
  float* part = p.partial ? p.partial + (size_t)b * stem_partial_floats(CH, C0, C1) : nullptr;
  const int o_c0b = C0 * C1 * 125, o_gam = o_c0b + C1, o_bet = o_gam + C0 * C0, o_u0w = o_bet + C0,
            o_u0b = o_u0w + CH * C0 * 125;
  if (p.need_w && part) {
    // conv0 wgrad: dW[ci][co][k] = sum_{i in 4^3} a0[ci][i] * g1[co][2i + k - 2]
    for (int e = tid; e < C0 * C1 * 125; e += kStemThreads) {
      int q = e;
      const int kx = q % 5; q /= 5;
      const int ky = q % 5; q /= 5;
      const int kz = q % 5; q /= 5;
      const int co = q % C1, ci = q / C1;
      float s = 0.f;
      for (int iz = 0; iz < 4; ++iz) {
        const int oz = 2 * iz + kz - 2;
        if (oz < 0 || oz >= 8) continue;
        for (int iy = 0; iy < 4; ++iy) {
          const int oy = 2 * iy + ky - 2;
          if (oy < 0 || oy >= 8) continue;
#pragma unroll
          for (int ix = 0; ix < 4; ++ix) {
            const int ox = 2 * ix + kx - 2;
            if (ox < 0 || ox >= 8) continue;
            s = fmaf(s_a0[ci * 64 + iz * 16 + iy * 4 + ix], s_g1[co * 512 + oz * 64 + oy * 8 + ox], s);
          }
        }
      }
      part[e] = s;
    }
    // conv0 bias: one warp per channel, fixed-order shuffle tree
    for (int co = warp; co < C1; co += kStemThreads / 32) {
      float s = 0.f;
      for (int i = lane; i < 512; i += 32) s += s_g1[co * 512 + i];
      for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) part[o_c0b + co] = s;
    }
  }
  // conv0 dgrad: gy[ci][i] = sum_co sum_k g1[co][2i + k - 2] * W[ci][co][k]; co range split in 4 parts
  for (int item = tid; item < C0 * 64 * 4; item += kStemThreads) {
    const int h = item / (C0 * 64), e = item % (C0 * 64);
    const int ci = e >> 6, iz = (e >> 4) & 3, iy = (e >> 2) & 3, ix = e & 3;
    float s = 0.f;
    for (int co = h * (C1 / 4); co < (h + 1) * (C1 / 4); ++co) {
      const float* wk = p.conv0_w + ((size_t)ci * C1 + co) * 125;
      for (int kz = 0; kz < 5; ++kz) {
        const int oz = 2 * iz + kz - 2;
        if (oz < 0 || oz >= 8) continue;
        for (int ky = 0; ky < 5; ++ky) {
          const int oy = 2 * iy + ky - 2;
          if (oy < 0 || oy >= 8) continue;
#pragma unroll
          for (int kx = 0; kx < 5; ++kx) {
            const int ox = 2 * ix + kx - 2;
            if (ox < 0 || ox >= 8) continue;
            s = fmaf(s_g1[co * 512 + oz * 64 + oy * 8 + ox], __ldg(wk + kz * 25 + ky * 5 + kx), s);
          }
        }
      }
    }
    s_part[h * C0 * 64 + e] = s;
  }
  __syncthreads();
  for (int e = tid; e < C0 * 64; e += kStemThreads) {
    s_gy[e] = (s_part[e] + s_part[C0 * 64 + e]) + (s_part[2 * C0 * 64 + e] + s_part[3 * C0 * 64 + e]);
    const int c = e >> 6, pos = e & 63;
    float nn = p.beta[c];
    for (int j = 0; j < C0; ++j) {
      const float xj = s_x0[j * 64 + pos];
      nn = fmaf(p.gamma[c * C0 + j], xj * xj, nn);
    }
    s_n[e] = sqrtf(nn);
  }
  __syncthreads();
  // IGDN backward: dx_k = g_k n_k + x_k sum_i g_i x_i gamma_ik / n_i
  for (int e = tid; e < C0 * 64; e += kStemThreads) {
    const int k = e >> 6, pos = e & 63;
    float acc = 0.f;
    for (int i = 0; i < C0; ++i)
      acc += s_gy[i * 64 + pos] * s_x0[i * 64 + pos] * p.gamma[i * C0 + k] / s_n[i * 64 + pos];
    s_gx[e] = s_gy[e] * s_n[e] + s_x0[e] * acc;
  }
  if (p.need_w && part) {
    // dbeta_i = sum t_i, dgamma_ij = sum t_i x_j^2 with t_i = g_i x_i / (2 n_i)
    for (int e = tid; e < C0 * C0 + C0; e += kStemThreads) {
      float s = 0.f;
      if (e < C0 * C0) {
        const int i = e / C0, j = e % C0;
        for (int pos = 0; pos < 64; ++pos) {
          const float t = s_gy[i * 64 + pos] * s_x0[i * 64 + pos] / (2.f * s_n[i * 64 + pos]);
          const float xj = s_x0[j * 64 + pos];
          s = fmaf(t, xj * xj, s);
        }
        part[o_gam + e] = s;
      } else {
        const int i = e - C0 * C0;
        for (int pos = 0; pos < 64; ++pos) s += s_gy[i * 64 + pos] * s_x0[i * 64 + pos] / (2.f * s_n[i * 64 + pos]);
        part[o_bet + i] = s;
      }
    }
  }
  __syncthreads();
  if (p.need_w && part) {
    // up0 wgrad: dW[ci][co][k] = sum_{i in 2^3} lat[ci][i] * gx[co][2i + k - 2]
    for (int e = tid; e < CH * C0 * 125; e += kStemThreads) {
      int q = e;
      const int kx = q % 5; q /= 5;
      const int ky = q % 5; q /= 5;
      const int kz = q % 5; q /= 5;
      const int co = q % C0, ci = q / C0;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int oz = 2 * (i >> 2) + kz - 2, oy = 2 * ((i >> 1) & 1) + ky - 2, ox = 2 * (i & 1) + kx - 2;
        if (oz < 0 || oz >= 4 || oy < 0 || oy >= 4 || ox < 0 || ox >= 4) continue;
        s = fmaf(s_lat[ci * 8 + i], s_gx[co * 64 + oz * 16 + oy * 4 + ox], s);
      }
      part[o_u0w + e] = s;
    }
    for (int co = warp; co < C0; co += kStemThreads / 32) {
      float s = s_gx[co * 64 + lane] + s_gx[co * 64 + 32 + lane];
      for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) part[o_u0b + co] = s;
    }
  }
  // d_latent[ci][i] = sum_co sum_k gx[co][2i + k - 2] * W[ci][co][k]: one warp per output, lanes over (co,k)
  if (p.g_latent) {
    for (int e = warp; e < CH * 8; e += kStemThreads / 32) {
      const int ci = e >> 3, i = e & 7;
      const int iz = i >> 2, iy = (i >> 1) & 1, ix = i & 1;
      float s = 0.f;
      for (int t = lane; t < C0 * 125; t += 32) {
        const int co = t / 125, k = t % 125;
        const int oz = 2 * iz + k / 25 - 2, oy = 2 * iy + (k / 5) % 5 - 2, ox = 2 * ix + k % 5 - 2;
        if (oz < 0 || oz >= 4 || oy < 0 || oy >= 4 || ox < 0 || ox >= 4) continue;
        s = fmaf(s_gx[co * 64 + oz * 16 + oy * 4 + ox], __ldg(p.up0_w + ((size_t)ci * C0 + co) * 125 + k), s);
      }
      for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) p.g_latent[(size_t)b * CH * 8 + e] = s;
    }
  }
}

}  // namespace fast
}  // namespace nvf
