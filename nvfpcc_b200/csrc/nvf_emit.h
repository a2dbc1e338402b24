// Occupancy mask -> ordered point list.
//
// Replaces MinkowskiToSparseTensor + `F > thh` + MinkowskiPruning + `C[:,1:]`
// + origin (NVFPCC.py:631-637 decode, :520,532-539 encode).  The reference's
// point order is block order, then row-major (i,j,k) of out_dense[b,0,i,j,k];
// rc_enc.ply == rc_dec.ply (README.md:63) needs that order to be deterministic,
// so compaction is a count -> exclusive scan -> ordered write (no atomics):
//   scan_counts   one CTA: offsets[b] = sum_{b'<b} counts[b'], total
//   emit_coords   one CTA per block: bit k of mask word (i*32+j) -> (i,j,k)+origin
#pragma once
#include "nvf_common.h"

namespace nvf {

struct EmitParams {
  const uint32_t* mask;    // [n][1024]
  const int32_t* counts;   // [n]
  const int32_t* origins;  // [n][3] or null
  int64_t* offsets;        // [n+1] workspace
  int32_t* coords;         // [cap][3] or null
  int64_t* total_out;      // [1] or null
  int64_t cap;
  int32_t n_blocks;
};

constexpr int kScanSmemInts = 2 * kThreads + 4;

template <class Env>
struct ScanBlock {
  // single CTA
  static NVF_HD void run(Env& env, const EmitParams& p, int64_t* sm) {
    const int n = p.n_blocks;
    const int per = (n + kThreads - 1) / kThreads;
    env.phase([&](int tid, int&) {
      int64_t s = 0;
      const int lo = tid * per, hi = lo + per < n ? lo + per : n;
      for (int i = lo; i < hi; ++i) s += p.counts[i];
      sm[tid] = s;
    });
    env.phase([&](int tid, int&) {
      int64_t pre = 0;
      for (int i = 0; i < tid; ++i) pre += sm[i];
      sm[kThreads + tid] = pre;
    });
    env.phase([&](int tid, int&) {
      int64_t run = sm[kThreads + tid];
      const int lo = tid * per, hi = lo + per < n ? lo + per : n;
      for (int i = lo; i < hi; ++i) {
        p.offsets[i] = run;
        run += p.counts[i];
      }
      if (tid == kThreads - 1) {
        p.offsets[n] = run;  // per*kThreads >= n, so the last thread ends at n (or has an empty range)
        if (p.total_out) *p.total_out = run;
      }
    });
  }
};

template <class Env>
struct EmitBlock {
  static NVF_HD void run(Env& env, const EmitParams& p, int* sm, int b) {
    const uint32_t* m = p.mask + (size_t)b * 1024;
    env.phase([&](int tid, int&) {
      int c = 0;
      for (int w = 0; w < 4; ++w) c += popc32(m[tid * 4 + w]);
      sm[tid] = c;
    });
    env.phase([&](int tid, int&) {
      int pre = 0;
      for (int i = 0; i < tid; ++i) pre += sm[i];
      sm[kThreads + tid] = pre;
    });
    env.phase([&](int tid, int&) {
      if (!p.coords) return;
      int64_t o = p.offsets[b] + sm[kThreads + tid];
      int ox = 0, oy = 0, oz = 0;
      if (p.origins) {
        ox = p.origins[b * 3 + 0];
        oy = p.origins[b * 3 + 1];
        oz = p.origins[b * 3 + 2];
      }
      for (int w = 0; w < 4; ++w) {
        const int word = tid * 4 + w;
        uint32_t bits = m[word];
        const int i = word >> 5, j = word & 31;
        for (int k = 0; k < 32; ++k) {
          if ((bits >> k) & 1u) {
            if (o < p.cap) {
              p.coords[o * 3 + 0] = ox + i;
              p.coords[o * 3 + 1] = oy + j;
              p.coords[o * 3 + 2] = oz + k;
            }
            ++o;
          }
        }
      }
    });
  }
};

}  // namespace nvf
