// Fused NVF decode kernel for chanstr 8,16,8,8 / ch 3 ("config A").
//
// ONE persistent CTA per SM walks over leaf blocks.  Per block:
//   stem  latent -> up0 -> IGDN -> conv0 -> up1 -> conv1        (13 % of MACs)
//         (utils/network.py:4759-4763) with conv0 output in shared memory and
//         the two larger tensors (up1 out 8x19^3, conv1 out 8x16^3) in a
//         per-CTA scratch that is rewritten for every block and therefore
//         lives in L2, not HBM.
//   head  up2 -> conv2 -> conv2_cls -> sigmoid -> threshold      (87 % of MACs)
//         (utils/network.py:4764-4768, NVFPCC.py:631-633) streamed along z:
//         one 8x35x35 up2 slice at a time is produced into shared memory and
//         consumed "input-stationary" by conv2, whose four partially summed
//         output slices live in registers (128 accumulators per thread); each
//         finished conv2 slice goes to shared memory once and is consumed the
//         same way by the 3x3x3 classifier (12 accumulators per thread).
//         Nothing of the 1.3 MB + 1.0 MB up2/conv2 activations leaves the SM.
// Output: dense probabilities (optional), a 32768-bit occupancy mask per block
// and the per-block point count; nvf_emit.h turns masks into ordered points.
#pragma once
#include "nvf_common.h"
#include "nvf_tiles.h"

namespace nvf {

struct FusedA {
  static constexpr int CH = 3, C0 = 8, C1 = 16, C2 = 8, C3 = 8;
  // ---- persistent shared memory (floats) --------------------------------
  static constexpr int W_UP2 = C2 * 125 * C3;  // [ci][kz][ky][kx][co]
  static constexpr int W_C2 = C3 * 64 * C3;    // [ci][kz][ky][kx][co]
  static constexpr int W_CLS = C3 * 36;        // [ci][ky][kz][4]
  static constexpr int OFF_W_UP2 = 0;
  static constexpr int OFF_W_C2 = OFF_W_UP2 + W_UP2;
  static constexpr int OFF_W_CLS = OFF_W_C2 + W_C2;
  static constexpr int OFF_BIAS = OFF_W_CLS + W_CLS;  // [0..7] up2_b, [8..15] conv2_b, [16] cls_b
  static constexpr int OFF_BITS = OFF_BIAS + 32;      // 64 x uint32 mask words (two z-slices), [64] count
  static constexpr int OFF_DYN = OFF_BITS + 96;
  // ---- head region --------------------------------------------------------
  static constexpr int U_PITCH = 36, U_ROWS = 35, U_PLANE = U_ROWS * U_PITCH, U_SIZE = C3 * U_PLANE;
  static constexpr int C_PITCH = 40, C_ROWS = 34, C_PLANE = C_ROWS * C_PITCH, C_SIZE = C3 * C_PLANE;
  static constexpr int IN_PITCH = 24, IN_ROWS = 20, IN_PLANE = IN_ROWS * IN_PITCH, IN_SLOT = C2 * IN_PLANE;
  static constexpr int IN_SIZE = 3 * IN_SLOT;
  static constexpr int OFF_U = OFF_DYN;
  static constexpr int OFF_C = OFF_U + U_SIZE;
  static constexpr int OFF_IN = OFF_C + C_SIZE;
  static constexpr int HEAD_END = OFF_IN + IN_SIZE;
  // ---- stem region (aliases the head region) ------------------------------
  static constexpr int W_UP1 = C1 * 125 * C2;  // 16000
  static constexpr int W_C1 = C2 * 64 * C2;    // 4096
  static constexpr int OFF_W_UP1 = OFF_DYN;
  static constexpr int OFF_W_C1 = OFF_W_UP1 + W_UP1;
  static constexpr int OFF_A1 = OFF_W_C1 + W_C1;       // conv0 out [16][8][8][8]
  static constexpr int OFF_X0 = OFF_A1 + C1 * 512;     // up0 out [8][4][4][4]
  static constexpr int OFF_A0 = OFF_X0 + C0 * 64;      // IGDN out
  static constexpr int OFF_LAT = OFF_A0 + C0 * 64;     // latent [3][8]
  static constexpr int STEM_END = OFF_LAT + 32;
  static constexpr int SMEM_FLOATS = HEAD_END > STEM_END ? HEAD_END : STEM_END;
  static constexpr int SMEM_BYTES = SMEM_FLOATS * 4;
  // ---- per-CTA global scratch (floats) ------------------------------------
  static constexpr int G2_PITCH = 20, G2_PLANE = 19 * G2_PITCH, G2_CS = 19 * G2_PLANE;  // up1 out
  static constexpr int OFF_G2 = 0;
  static constexpr int OFF_S = OFF_G2 + C2 * G2_CS;  // conv1 out [8][16][16][16]
  static constexpr int SCRATCH_FLOATS = OFF_S + C2 * 4096;
  // ---- packed weights in the workspace (floats) ----------------------------
  static constexpr int P_CONV0 = 0;                        // [8][125][16]
  static constexpr int P_UP1 = P_CONV0 + C0 * 125 * C1;    // [16][125][8]
  static constexpr int P_C1 = P_UP1 + W_UP1;               // [8][64][8]
  static constexpr int P_UP2 = P_C1 + W_C1;                // [8][125][8]
  static constexpr int P_C2 = P_UP2 + W_UP2;               // [8][64][8]
  static constexpr int P_CLS = P_C2 + W_C2;                // [8][3][3][4]
  static constexpr int PACKED_FLOATS = P_CLS + W_CLS;
};
static_assert(FusedA::SMEM_BYTES <= 227 * 1024, "fused decode smem");
static_assert(FusedA::OFF_DYN % 4 == 0 && FusedA::OFF_U % 4 == 0 && FusedA::OFF_C % 4 == 0 && FusedA::OFF_IN % 4 == 0,
              "16-byte alignment of smem regions");

struct FusedAParams {
  const float* latent;    // [n][3][2][2][2]
  const float* up0_w;     // raw (3,8,5,5,5)
  const float* up0_b;     // [8]
  const float* igdn_beta; // [8]
  const float* igdn_gamma;// [8][8]
  const float* conv0_b;   // [16]
  const float* up1_b;     // [8]
  const float* conv1_b;   // [8]
  const float* up2_b;     // [8]
  const float* conv2_b;   // [8]
  const float* cls2_b;    // [1]
  const float* packed;    // FusedA::PACKED_FLOATS
  float* scratch;         // gridDim * FusedA::SCRATCH_FLOATS
  float* prob_out;        // [n][32768] or null
  uint32_t* mask_out;     // [n][1024]
  int32_t* counts_out;    // [n]
  float thh;
  int32_t n_blocks;
};

// Per-thread state that lives across phases (registers on the GPU).
struct FusedATS {
  p2 acc[4][4][4];     // conv2 partial sums: [slot j: output slice z2-3+j][co pair][x] (lo = even co)
  float cacc[3][4];    // cls partial sums:   [slot g: output slice z1-1+g][x]
  int cnt;             // points counted by this thread (mask flush lanes)
};

template <class Env>
struct FusedABlock {
  using G = FusedA;

  // ---- helpers -------------------------------------------------------------
  static NVF_HD void copy4(float* dst, const float* src, int n, int tid) {
    for (int i = tid * 4; i < n; i += kThreads * 4) {
      f4 v = ld4(src + i);
      st4(dst + i, v.x, v.y, v.z, v.w);
    }
  }
  static NVF_HD void zero4(float* dst, int n, int tid) {
    for (int i = tid * 4; i < n; i += kThreads * 4) st4(dst + i, 0.f, 0.f, 0.f, 0.f);
  }
  // conv1-out slice iz (global scratch S) -> zero-padded smem slot iz%3
  static NVF_HD void load_in_slice(float* sm, const float* S, int iz, int tid) {
    const int ci = tid >> 5, rem = tid & 31, y = rem >> 1, xh = rem & 1;
    const float* src = S + ci * 4096 + iz * 256 + y * 16 + xh * 8;
    float* dst = sm + G::OFF_IN + (iz % 3) * G::IN_SLOT + ci * G::IN_PLANE + (y + 2) * G::IN_PITCH + 4 + xh * 8;
    f4 a = ld4(src), b = ld4(src + 4);
    st4(dst, a.x, a.y, a.z, a.w);
    st4(dst + 4, b.x, b.y, b.z, b.w);
  }

  // ---- head phase bodies ------------------------------------------------------
  // up2: one 8 x 35 x 35 output slice z2 (post-ReLU) into U.  175 thread tiles of
  // 8 co x 8 x (4 even + 4 odd outputs); rows of equal parity share a warp.
  static NVF_HD void up2_slice(float* sm, int z2, int tid) {
    const int cls = tid / 96, idx = tid - cls * 96;
    if (cls > 1) return;
    if (idx >= (cls == 0 ? 90 : 85)) return;
    const int yi = idx / 5, q = idx - yi * 5;
    const int y2 = 2 * yi + cls;
    p2 acc[4][8];  // [co pair][x]
    NVF_UNROLL
    for (int c = 0; c < 4; ++c) {
      NVF_UNROLL
      for (int j = 0; j < 8; ++j) acc[c][j] = p2_bcast(0.f);
    }
    const float* W = sm + G::OFF_W_UP2;
    NVF_NOUNROLL
    for (int ci = 0; ci < G::C2; ++ci) {
      NVF_NOUNROLL
      for (int kz = z2 & 1; kz < 5; kz += 2) {
        const int tz = z2 - kz;
        if (tz < 0) break;
        const int iz = tz >> 1;
        if (iz > 15) continue;
        const float* plane = sm + G::OFF_IN + (iz % 3) * G::IN_SLOT + ci * G::IN_PLANE;
        NVF_NOUNROLL
        for (int ky = cls; ky < 5; ky += 2) {
          const int row = ((y2 - ky) >> 1) + 2;  // y2-ky is even, >= -4
          const float* r = plane + row * G::IN_PITCH + 4 * q + 2;
          float a0, a1;
          ld2(r, a0, a1);
          f4 v = ld4(r + 2);
          const p2 a[6] = {p2_bcast(a0), p2_bcast(a1), p2_bcast(v.x), p2_bcast(v.y), p2_bcast(v.z), p2_bcast(v.w)};
          const float* wrow = W + (((ci * 5 + kz) * 5 + ky) * 5) * 8;
          NVF_UNROLL
          for (int kx = 0; kx < 5; ++kx) {
            p2 w[4];
            p2_load_w8(wrow + kx * 8, w);
            const int h = kx >> 1;
            NVF_UNROLL
            for (int c = 0; c < 4; ++c) {
              NVF_UNROLL
              for (int j = 0; j < 4; ++j) {
                if ((kx & 1) == 0) p2_fma(acc[c][2 * j], w[c], a[j + 2 - h]);
                else p2_fma(acc[c][2 * j + 1], w[c], a[j + 2 - h]);
              }
            }
          }
        }
      }
    }
    const float* bias = sm + G::OFF_BIAS;
    NVF_UNROLL
    for (int c = 0; c < 8; ++c) {
      float* o = sm + G::OFF_U + c * G::U_PLANE + y2 * G::U_PITCH + 8 * q;
      const float bb = bias[c];
      float v[8];
      NVF_UNROLL
      for (int j = 0; j < 8; ++j) v[j] = relu(((c & 1) ? p2_hi(acc[c >> 1][j]) : p2_lo(acc[c >> 1][j])) + bb);
      st4(o, v[0], v[1], v[2], v[3]);
      if (q < 4) st4(o + 4, v[4], v[5], v[6], v[7]);
    }
  }

  // conv2, input-stationary: slice z2 of U contributes to output slices z2-3+j with kz = 3-j.
  template <bool ALL>
  static NVF_HD void conv2_accum(const float* sm, FusedATS& ts, int z2, int y, int x0, int lo, int hi) {
    const float* W = sm + G::OFF_W_C2;
    NVF_NOUNROLL
    for (int ci = 0; ci < G::C3; ++ci) {
      NVF_UNROLL
      for (int ky = 0; ky < 4; ++ky) {
        const float* r = sm + G::OFF_U + ci * G::U_PLANE + (y + ky) * G::U_PITCH + x0;
        const f4 v0 = ld4(r), v1 = ld4(r + 4);
        const p2 a[8] = {p2_bcast(v0.x), p2_bcast(v0.y), p2_bcast(v0.z), p2_bcast(v0.w),
                         p2_bcast(v1.x), p2_bcast(v1.y), p2_bcast(v1.z), p2_bcast(v1.w)};
        NVF_UNROLL
        for (int j = 0; j < 4; ++j) {
          if (!ALL && (j < lo || j > hi)) continue;
          const int kz = 3 - j;
          const float* wrow = W + (((ci * 4 + kz) * 4 + ky) * 4) * 8;
          NVF_UNROLL
          for (int kx = 0; kx < 4; ++kx) {
            p2 w[4];
            p2_load_w8(wrow + kx * 8, w);
            NVF_UNROLL
            for (int c = 0; c < 4; ++c) {
              NVF_UNROLL
              for (int i = 0; i < 4; ++i) p2_fma(ts.acc[j][c][i], w[c], a[i + kx]);
            }
          }
        }
      }
    }
  }

  static NVF_HD void conv2_slice(float* sm, FusedATS& ts, int z2, int tid) {
    const int y = tid >> 3, x0 = (tid & 7) * 4;
    const int lo = 3 - z2 > 0 ? 3 - z2 : 0;
    const int hi = 34 - z2 < 3 ? 34 - z2 : 3;
    if (lo == 0 && hi == 3) conv2_accum<true>(sm, ts, z2, y, x0, lo, hi);
    else conv2_accum<false>(sm, ts, z2, y, x0, lo, hi);
    const int zo = z2 - 3;
    if (zo >= 0) {
      const float* bias = sm + G::OFF_BIAS + 8;
      NVF_UNROLL
      for (int c = 0; c < 8; ++c) {
        const float bb = bias[c];
        float v[4];
        NVF_UNROLL
        for (int i = 0; i < 4; ++i) v[i] = relu(((c & 1) ? p2_hi(ts.acc[0][c >> 1][i]) : p2_lo(ts.acc[0][c >> 1][i])) + bb);
        st4(sm + G::OFF_C + c * G::C_PLANE + (y + 1) * G::C_PITCH + 4 + x0, v[0], v[1], v[2], v[3]);
      }
    }
    NVF_UNROLL
    for (int c = 0; c < 4; ++c) {
      NVF_UNROLL
      for (int i = 0; i < 4; ++i) {
        ts.acc[0][c][i] = ts.acc[1][c][i];
        ts.acc[1][c][i] = ts.acc[2][c][i];
        ts.acc[2][c][i] = ts.acc[3][c][i];
        ts.acc[3][c][i] = p2_bcast(0.f);
      }
    }
  }

  static NVF_HD void emit_cls(float* sm, const FusedAParams& p, int b, int zo, int y, int x0, const float (&v)[4]) {
    const float bb = sm[G::OFF_BIAS + 16];
    float pr[4];
    uint32_t nib = 0;
    NVF_UNROLL
    for (int i = 0; i < 4; ++i) {
      pr[i] = sigmoidf(v[i] + bb);
      if (pr[i] > p.thh) nib |= 1u << i;
    }
    if (p.prob_out) st4(p.prob_out + (size_t)b * kVox + (zo * 32 + y) * 32 + x0, pr[0], pr[1], pr[2], pr[3]);
    if (nib) atomic_or_u32(reinterpret_cast<uint32_t*>(sm + G::OFF_BITS) + (zo & 1) * 32 + y, nib << x0);
  }

  // cls head, input-stationary: conv2-out slice z1 contributes to logits zo = z1-1+g with kz = 2-g.
  static NVF_HD void cls_slice(float* sm, FusedATS& ts, const FusedAParams& p, int b, int z1, int tid) {
    const int y = tid >> 3, x0 = (tid & 7) * 4;
    const float* W = sm + G::OFF_W_CLS;
    const int lo = z1 == 0 ? 1 : 0, hi = z1 == 31 ? 1 : 2;
    NVF_NOUNROLL
    for (int ci = 0; ci < G::C3; ++ci) {
      NVF_UNROLL
      for (int ky = 0; ky < 3; ++ky) {
        const float* r = sm + G::OFF_C + ci * G::C_PLANE + (y + ky) * G::C_PITCH + x0 + 3;
        const f4 v = ld4(r + 1);
        const float a[6] = {r[0], v.x, v.y, v.z, v.w, r[5]};
        const float* wr = W + (ci * 3 + ky) * 12;
        NVF_UNROLL
        for (int g = 0; g < 3; ++g) {
          if (g < lo || g > hi) continue;
          const f4 w = ld4(wr + (2 - g) * 4);
          NVF_UNROLL
          for (int i = 0; i < 4; ++i) {
            ts.cacc[g][i] = fmaf(w.x, a[i], ts.cacc[g][i]);
            ts.cacc[g][i] = fmaf(w.y, a[i + 1], ts.cacc[g][i]);
            ts.cacc[g][i] = fmaf(w.z, a[i + 2], ts.cacc[g][i]);
          }
        }
      }
    }
    if (z1 >= 1) emit_cls(sm, p, b, z1 - 1, y, x0, ts.cacc[0]);
    if (z1 == 31) emit_cls(sm, p, b, 31, y, x0, ts.cacc[1]);
    NVF_UNROLL
    for (int i = 0; i < 4; ++i) {
      ts.cacc[0][i] = ts.cacc[1][i];
      ts.cacc[1][i] = ts.cacc[2][i];
      ts.cacc[2][i] = 0.f;
    }
  }

  // mask words of finished logit slice zo: smem -> global, count points (lanes 224..255).
  static NVF_HD void flush_bits(float* sm, FusedATS& ts, const FusedAParams& p, int b, int zo, int tid) {
    const int r = tid - 224;
    if (r < 0) return;
    uint32_t* bits = reinterpret_cast<uint32_t*>(sm + G::OFF_BITS) + (zo & 1) * 32 + r;
    const uint32_t w = *bits;
    *bits = 0u;
    p.mask_out[(size_t)b * 1024 + zo * 32 + r] = w;
    ts.cnt += popc32(w);
  }

  // ---- stem phase bodies ------------------------------------------------------
  static NVF_HD void stem_up0(float* sm, const FusedAParams& p, int tid) {
    // up0: convT k5 s2 p2 op1, 3 x 2^3 -> 8 x 4^3 (utils/network.py:4671-4680), raw PyTorch layout
    for (int idx = tid; idx < G::C0 * 64; idx += kThreads) {
      const int co = idx >> 6, z = (idx >> 4) & 3, y = (idx >> 2) & 3, x = idx & 3;
      float v = p.up0_b[co];
      for (int ci = 0; ci < G::CH; ++ci)
        for (int kz = z & 1; kz < 5; kz += 2) {
          const int tz = z + 2 - kz;
          if (tz < 0 || (tz >> 1) >= 2) continue;
          for (int ky = y & 1; ky < 5; ky += 2) {
            const int ty = y + 2 - ky;
            if (ty < 0 || (ty >> 1) >= 2) continue;
            for (int kx = x & 1; kx < 5; kx += 2) {
              const int tx = x + 2 - kx;
              if (tx < 0 || (tx >> 1) >= 2) continue;
              const float a = sm[G::OFF_LAT + ci * 8 + (tz >> 1) * 4 + (ty >> 1) * 2 + (tx >> 1)];
              v = fmaf(a, p.up0_w[((ci * G::C0 + co) * 5 + kz) * 25 + ky * 5 + kx], v);
            }
          }
        }
      sm[G::OFF_X0 + idx] = v;
    }
  }
  static NVF_HD void stem_igdn(float* sm, const FusedAParams& p, int tid) {
    // IGDN: y_i = x_i * sqrt(beta_i + sum_j gamma_ij x_j^2)  (gdn_3d.py:153-157)
    for (int idx = tid; idx < G::C0 * 64; idx += kThreads) {
      const int c = idx >> 6, pos = idx & 63;
      float n = p.igdn_beta[c];
      for (int j = 0; j < G::C0; ++j) {
        const float xj = sm[G::OFF_X0 + j * 64 + pos];
        n = fmaf(p.igdn_gamma[c * G::C0 + j], xj * xj, n);
      }
      sm[G::OFF_A0 + idx] = sm[G::OFF_X0 + idx] * sqrtf(n);
    }
  }
  static NVF_HD void stem_conv0(float* sm, const FusedAParams& p, int tid) {
    // conv0: convT 8 -> 16, 4^3 -> 8^3, ReLU.  128 tiles of 16 co x 4 x.
    if (tid >= 128) return;
    const int oz = tid >> 4, oy = (tid >> 1) & 7, ox0 = (tid & 1) * 4;
    float acc[16][4];
    zero_acc<16>(acc);
    View in{sm + G::OFF_A0, 64, 16, 4, 4};
    convT5_tile<16>(in, G::C0, 2, p.packed + G::P_CONV0, G::C1, 0, oz, oy, ox0, acc);
    NVF_UNROLL
    for (int c = 0; c < 16; ++c) {
      const float bb = p.conv0_b[c];
      st4(sm + G::OFF_A1 + c * 512 + oz * 64 + oy * 8 + ox0, relu(acc[c][0] + bb), relu(acc[c][1] + bb),
          relu(acc[c][2] + bb), relu(acc[c][3] + bb));
    }
  }
  static NVF_HD void stem_up1(float* sm, const FusedAParams& p, float* scratch, int tid) {
    // up1: convT 16 -> 8, 8^3 -> 19^3, ReLU, to scratch G2 (row pitch 20).
    View in{sm + G::OFF_A1, 512, 64, 8, 8};
    for (int t = tid; t < 19 * 19 * 5; t += kThreads) {
      const int oz = t / 95, r = t - oz * 95, oy = r / 5, ox0 = (r - oy * 5) * 4;
      float acc[8][4];
      zero_acc<8>(acc);
      convT5_tile<8>(in, G::C1, 0, sm + G::OFF_W_UP1, G::C2, 0, oz, oy, ox0, acc);
      NVF_UNROLL
      for (int c = 0; c < 8; ++c) {
        const float bb = p.up1_b[c];
        st4(scratch + G::OFF_G2 + c * G::G2_CS + oz * G::G2_PLANE + oy * G::G2_PITCH + ox0, relu(acc[c][0] + bb),
            relu(acc[c][1] + bb), relu(acc[c][2] + bb), relu(acc[c][3] + bb));
      }
    }
  }
  static NVF_HD void stem_conv1(float* sm, const FusedAParams& p, float* scratch, int tid) {
    // conv1: conv k4 8 -> 8, 19^3 -> 16^3, ReLU, to scratch S.
    View in{scratch + G::OFF_G2, G::G2_CS, G::G2_PLANE, G::G2_PITCH, 19};
    for (int t = tid; t < 16 * 16 * 4; t += kThreads) {
      const int oz = t >> 6, oy = (t >> 2) & 15, ox0 = (t & 3) * 4;
      float acc[8][4];
      zero_acc<8>(acc);
      conv4_valid_tile<8>(in, G::C2, sm + G::OFF_W_C1, G::C2, 0, oz, oy, ox0, acc);
      NVF_UNROLL
      for (int c = 0; c < 8; ++c) {
        const float bb = p.conv1_b[c];
        st4(scratch + G::OFF_S + c * 4096 + oz * 256 + oy * 16 + ox0, relu(acc[c][0] + bb), relu(acc[c][1] + bb),
            relu(acc[c][2] + bb), relu(acc[c][3] + bb));
      }
    }
  }

  // ---- the CTA program ----------------------------------------------------------
  // env.phase(f): run f(tid, ts) for every thread of the CTA, then barrier.
  static NVF_HD void run(Env& env, const FusedAParams& p, float* sm, int cta, int ncta) {
    float* scratch = p.scratch + (size_t)cta * G::SCRATCH_FLOATS;
    // once per CTA: head weights + biases stay resident in shared memory
    env.phase([&](int tid, FusedATS& ts) {
      copy4(sm + G::OFF_W_UP2, p.packed + G::P_UP2, G::W_UP2, tid);
      copy4(sm + G::OFF_W_C2, p.packed + G::P_C2, G::W_C2, tid);
      copy4(sm + G::OFF_W_CLS, p.packed + G::P_CLS, G::W_CLS, tid);
      if (tid < 8) sm[G::OFF_BIAS + tid] = p.up2_b[tid];
      else if (tid < 16) sm[G::OFF_BIAS + tid] = p.conv2_b[tid - 8];
      else if (tid == 16) sm[G::OFF_BIAS + 16] = p.cls2_b[0];
      if (tid < 96) reinterpret_cast<uint32_t*>(sm + G::OFF_BITS)[tid] = 0u;
      ts.cnt = 0;
    });
    for (int b = cta; b < p.n_blocks; b += ncta) {
      // ------------------------------ stem ------------------------------
      env.phase([&](int tid, FusedATS& ts) {
        copy4(sm + G::OFF_W_UP1, p.packed + G::P_UP1, G::W_UP1, tid);
        copy4(sm + G::OFF_W_C1, p.packed + G::P_C1, G::W_C1, tid);
        if (tid < G::CH * 8) sm[G::OFF_LAT + tid] = p.latent[(size_t)b * G::CH * 8 + tid];
      });
      env.phase([&](int tid, FusedATS& ts) { stem_up0(sm, p, tid); });
      env.phase([&](int tid, FusedATS& ts) { stem_igdn(sm, p, tid); });
      env.phase([&](int tid, FusedATS& ts) { stem_conv0(sm, p, tid); });
      env.phase([&](int tid, FusedATS& ts) { stem_up1(sm, p, scratch, tid); });
      env.phase([&](int tid, FusedATS& ts) { stem_conv1(sm, p, scratch, tid); });
      // ------------------------------ head ------------------------------
      env.phase([&](int tid, FusedATS& ts) {
        zero4(sm + G::OFF_C, G::C_SIZE + G::IN_SIZE, tid);  // C and IN are contiguous: zero halos
        NVF_UNROLL
        for (int j = 0; j < 4; ++j) {
          NVF_UNROLL
          for (int c = 0; c < 4; ++c) {
            NVF_UNROLL
            for (int i = 0; i < 4; ++i) ts.acc[j][c][i] = p2_bcast(0.f);
          }
        }
        NVF_UNROLL
        for (int g = 0; g < 3; ++g) {
          NVF_UNROLL
          for (int i = 0; i < 4; ++i) ts.cacc[g][i] = 0.f;
        }
        ts.cnt = 0;
      });
      env.phase([&](int tid, FusedATS& ts) { load_in_slice(sm, scratch + G::OFF_S, 0, tid); });
      NVF_NOUNROLL
      for (int s = 0; s < 35; ++s) {
        env.phase([&](int tid, FusedATS& ts) {
          up2_slice(sm, s, tid);
          if (s >= 5) flush_bits(sm, ts, p, b, s - 5, tid);
        });
        env.phase([&](int tid, FusedATS& ts) { conv2_slice(sm, ts, s, tid); });
        env.phase([&](int tid, FusedATS& ts) {
          if (s >= 3) cls_slice(sm, ts, p, b, s - 3, tid);
          const int sn = s + 1;
          if ((sn & 1) == 0 && (sn >> 1) < 16) load_in_slice(sm, scratch + G::OFF_S, sn >> 1, tid);
        });
      }
      env.phase([&](int tid, FusedATS& ts) {
        flush_bits(sm, ts, p, b, 30, tid);
        flush_bits(sm, ts, p, b, 31, tid);
        if (tid >= 224) atomic_add_i32(reinterpret_cast<int*>(sm + G::OFF_BITS) + 64, ts.cnt);
      });
      env.phase([&](int tid, FusedATS& ts) {
        if (tid == 0) {
          int* c = reinterpret_cast<int*>(sm + G::OFF_BITS) + 64;
          p.counts_out[b] = *c;
          *c = 0;
        }
      });
    }
  }
};

}  // namespace nvf
