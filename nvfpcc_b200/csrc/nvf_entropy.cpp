// Host-side entropy coding of the NVF bitstream, in process (include/nvf_prep_b200.h).
//
// * Latent code: adaptive-model binary arithmetic coder, bit-compatible with the stream the
//   reference produces by piping through its ./module_arithmeticcoding helper
//   (module_arithmeticcoding.cpp:368-432 driver, :119-173 Gaussian frequency model,
//   :175-233 interval update, :235-266 bit output, :268-360 decoder; called from
//   NVFPCC.py:446-477 and :588-607).  Differences in construction, not in the stream:
//   buffers instead of stdin/stdout, one cumulative table per distinct (mu, sigma) pair instead
//   of two erf() calls per symbol and ~20 per decoded symbol, 64-bit state with a 128-bit
//   product, and an exact integer symbol search in the decoder (the reference estimates the
//   target in double precision and asserts; the exact search returns the same symbol whenever
//   that assert holds).
// * Weight code: canonical bit walk over the Huffman codebook shipped in the pack
//   (util_code_quantized_weights.py:108-148; MSB-first bit packing).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "../../include/nvf_b200.h"
#include "../../include/nvf_prep_b200.h"

namespace {

typedef unsigned __int128 u128;

constexpr int kNumSymbols = 1025;            // module_arithmeticcoding.cpp:129 (symbols 0..1024)
constexpr int kMulFactor = 10000000;         // :129
constexpr uint32_t kTotal = kMulFactor + 1025;   // :135
constexpr uint64_t kTop = 1ull << 63, kSecond = 1ull << 62;

inline float mask_float(float v, int level) {   // set_FL_MASK / convert, :95-113
  uint32_t bits;
  std::memcpy(&bits, &v, 4);
  bits &= ~((1u << level) - 1u);
  float r;
  std::memcpy(&r, &bits, 4);
  return r;
}

// cumulative count below `symbol` (get_low, :152-157; get_high(s) == get_low(s + 1), :159-164).
// The arithmetic types follow the reference expression exactly: the argument of erf is formed in
// double from a float sigma + float TINY sum, the CDF value is rounded to float, and the product
// with the integer scale is a float product.
inline uint32_t model_low(float mu, float sigma, int symbol) {
  const float tiny = 1e-10f;
  const float s = sigma + tiny;
  const double arg = ((double)(symbol - 1) + 0.5 - (double)mu) / ((double)s * std::sqrt(2.0));
  const float c = (float)(0.5 * (1.0 + std::erf(arg)));
  const float scaled = std::floor(c * (float)kMulFactor);
  return (uint32_t)(int)(scaled + (float)symbol);
}

struct Model {
  const uint32_t* table;   // [kNumSymbols + 1] or null (computed on demand)
  float mu, sigma;
  uint32_t low(int s) const { return table ? table[s] : model_low(mu, sigma, s); }
};

class ModelCache {
 public:
  ModelCache(int level1, int level2) : l1_(level1), l2_(level2) {}
  Model get(float mu_raw, float sigma_raw) {
    const float mu = mask_float(mu_raw, l1_), sigma = mask_float(sigma_raw, l2_);
    uint32_t a, b;
    std::memcpy(&a, &mu, 4);
    std::memcpy(&b, &sigma, 4);
    const uint64_t key = ((uint64_t)a << 32) | b;
    auto it = map_.find(key);
    if (it == map_.end()) {
      if (map_.size() >= kMaxTables) return Model{nullptr, mu, sigma};
      std::vector<uint32_t> t(kNumSymbols + 1);
      for (int s = 0; s <= kNumSymbols; ++s) t[s] = model_low(mu, sigma, s);
      it = map_.emplace(key, std::move(t)).first;
    }
    return Model{it->second.data(), mu, sigma};
  }

 private:
  static constexpr size_t kMaxTables = 4096;
  int l1_, l2_;
  std::unordered_map<uint64_t, std::vector<uint32_t>> map_;
};

struct BitWriter {
  uint8_t* out;
  size_t cap, len = 0;
  int cur = 0, filled = 0;
  bool overflow = false;
  void put(int b) {
    cur = (cur << 1) | b;
    if (++filled == 8) {
      if (len < cap) out[len] = (uint8_t)cur; else overflow = true;
      ++len;
      cur = 0;
      filled = 0;
    }
  }
  void close() { while (filled != 0) put(0); }
};

struct BitReader {
  const uint8_t* in;
  size_t len, pos = 0;
  int left = 0;
  uint8_t cur = 0;
  int get() {   // past the end the stream reads as zeros (read_code_bit, :349-353)
    if (left == 0) {
      if (pos >= len) return 0;
      cur = in[pos++];
      left = 8;
    }
    --left;
    return (cur >> left) & 1;
  }
};

// interval update shared by both directions (:181-232); Sink receives shift / underflow events
template <class Sink>
inline bool coder_update(uint64_t& low, uint64_t& high, uint32_t symlow, uint32_t symhigh, Sink& sink) {
  if (symlow >= symhigh) return false;               // zero-frequency symbol
  const u128 range = (u128)high - low + 1;
  const uint64_t nl = low + (uint64_t)((u128)symlow * range / kTotal);
  const uint64_t nh = low + (uint64_t)((u128)symhigh * range / kTotal) - 1;
  low = nl;
  high = nh;
  while (((low ^ high) & kTop) == 0) {
    sink.shift(low);
    low <<= 1;
    high = (high << 1) | 1;
  }
  while ((low & ~high & kSecond) != 0) {
    sink.underflow();
    low = (low << 1) & (~0ull >> 1);
    high = ((high << 1) & (~0ull >> 1)) | kTop | 1;
  }
  return true;
}

struct EncSink {
  BitWriter* w;
  uint64_t pending = 0;
  void shift(uint64_t low) {
    const int bit = (int)(low >> 63);
    w->put(bit);
    for (; pending > 0; --pending) w->put(bit ^ 1);
  }
  void underflow() { ++pending; }
};

struct DecSink {
  BitReader* r;
  uint64_t code = 0;
  void shift(uint64_t) { code = (code << 1) | (uint64_t)r->get(); }
  void underflow() { code = (code & kTop) | ((code << 1) & (~0ull >> 1)) | (uint64_t)r->get(); }
};

}  // namespace

extern "C" {

int nvf_arith_encode_bound(int64_t n_symbols, size_t* bytes_out) {
  if (!bytes_out || n_symbols < 0) return NVF_ERR_INVALID_ARG;
  // a symbol costs at most log2(total / 1) < 24 bits; + terminator + pending/flush bits
  *bytes_out = (size_t)n_symbols * 3 + 64;
  return NVF_OK;
}

int nvf_arith_encode_host(const int16_t* symbols, const float* mu, const float* sigma, int64_t n, int level1,
                          int level2, uint8_t* out, size_t out_cap, size_t* out_len) {
  if (n < 0 || !out || !out_len || (n > 0 && (!symbols || !mu || !sigma)) || level1 < 0 || level1 > 23 ||
      level2 < 0 || level2 > 23)
    return NVF_ERR_INVALID_ARG;
  ModelCache cache(level1, level2);
  BitWriter w{out, out_cap};
  EncSink sink{&w};
  uint64_t low = 0, high = ~0ull;
  for (int64_t i = 0; i < n; ++i) {
    const int s = symbols[i];
    if (s < 0 || s >= kNumSymbols) return NVF_ERR_INVALID_ARG;           // "Symbol out of range", :166-173
    const Model m = cache.get(mu[i], sigma[i]);
    if (!coder_update(low, high, m.low(s), m.low(s + 1), sink)) return NVF_ERR_BITSTREAM;
  }
  {   // terminator: symbol 512 under N(255, 1) (:391-395), then the closing 1 bit (:251-253)
    const Model m = cache.get(255.f, 1.f);
    if (!coder_update(low, high, m.low(512), m.low(513), sink)) return NVF_ERR_BITSTREAM;
    w.put(1);
  }
  // The reference never closes its bit stream (:396-399 end without BitOutputStream::close): the bits of
  // the last, partially filled byte are dropped - the 23-bit terminator exists to push the payload out -
  // so only complete bytes belong to the stream.
  *out_len = w.len;
  return w.overflow ? NVF_ERR_WORKSPACE : NVF_OK;
}

int nvf_arith_decode_host(const uint8_t* stream, size_t stream_len, const float* mu, const float* sigma, int64_t n,
                          int level1, int level2, int16_t* symbols_out) {
  if (n < 0 || (stream_len > 0 && !stream) || (n > 0 && (!symbols_out || !mu || !sigma)) || level1 < 0 ||
      level1 > 23 || level2 < 0 || level2 > 23)
    return NVF_ERR_INVALID_ARG;
  ModelCache cache(level1, level2);
  BitReader r{stream, stream_len};
  DecSink sink{&r};
  for (int i = 0; i < 64; ++i) sink.code = (sink.code << 1) | (uint64_t)r.get();
  uint64_t low = 0, high = ~0ull;
  for (int64_t i = 0; i < n; ++i) {
    const Model m = cache.get(mu[i], sigma[i]);
    const u128 range = (u128)high - low + 1;
    const uint64_t offset = sink.code - low;
    // largest s with low(s) * range / total <= offset  (the containment test of :330)
    int a = 0, b = kNumSymbols;
    if ((uint64_t)((u128)m.low(0) * range / kTotal) > offset) return NVF_ERR_BITSTREAM;
    while (b - a > 1) {
      const int mid = (a + b) >> 1;
      if ((uint64_t)((u128)m.low(mid) * range / kTotal) > offset) b = mid; else a = mid;
    }
    const u128 hi_edge = (u128)m.low(a + 1) * range / kTotal;
    if (!((u128)offset < hi_edge)) return NVF_ERR_BITSTREAM;
    if (!coder_update(low, high, m.low(a), m.low(a + 1), sink)) return NVF_ERR_BITSTREAM;
    if (!(low <= sink.code && sink.code <= high)) return NVF_ERR_BITSTREAM;   // "Code out of range", :335-338
    symbols_out[i] = (int16_t)a;
  }
  return NVF_OK;
}

int nvf_huffman_encode_host(const int32_t* symbols, int64_t n, const int32_t* code_symbols,
                            const uint8_t* code_lengths, const uint64_t* code_bits, int32_t n_codes, uint8_t* out,
                            size_t out_cap, size_t* out_len) {
  if (n < 0 || n_codes < 1 || !code_symbols || !code_lengths || !code_bits || !out || !out_len || (n > 0 && !symbols))
    return NVF_ERR_INVALID_ARG;
  std::unordered_map<int32_t, int32_t> index;
  for (int32_t c = 0; c < n_codes; ++c) {
    if (code_lengths[c] > 64) return NVF_ERR_INVALID_ARG;
    index[code_symbols[c]] = c;
  }
  BitWriter w{out, out_cap};
  for (int64_t i = 0; i < n; ++i) {
    auto it = index.find(symbols[i]);
    if (it == index.end()) return NVF_ERR_INVALID_ARG;
    const int len = code_lengths[it->second];
    const uint64_t bits = code_bits[it->second];
    for (int b = len - 1; b >= 0; --b) w.put((int)((bits >> b) & 1));
  }
  w.close();   // zero padding to a byte boundary (util_code_quantized_weights.py:122-124)
  *out_len = w.len;
  return w.overflow ? NVF_ERR_WORKSPACE : NVF_OK;
}

int nvf_huffman_decode_host(const uint8_t* stream, size_t stream_len, const int32_t* code_symbols,
                            const uint8_t* code_lengths, const uint64_t* code_bits, int32_t n_codes,
                            int64_t n_symbols, int32_t* symbols_out) {
  if (n_symbols < 0 || n_codes < 1 || !code_symbols || !code_lengths || !code_bits || (stream_len > 0 && !stream) ||
      (n_symbols > 0 && !symbols_out))
    return NVF_ERR_INVALID_ARG;
  // binary trie: child[node][bit], leaf = ~index
  std::vector<int32_t> child(2, 0);
  auto new_node = [&]() { child.push_back(0); child.push_back(0); return (int32_t)(child.size() / 2 - 1); };
  for (int32_t c = 0; c < n_codes; ++c) {
    const int len = code_lengths[c];
    if (len > 64) return NVF_ERR_INVALID_ARG;
    if (len == 0) {                       // single-symbol alphabet: the empty codeword
      if (n_codes != 1) return NVF_ERR_INVALID_ARG;
      for (int64_t i = 0; i < n_symbols; ++i) symbols_out[i] = code_symbols[0];
      return NVF_OK;
    }
    int32_t node = 0;
    for (int b = len - 1; b >= 0; --b) {
      const int bit = (int)((code_bits[c] >> b) & 1);
      int32_t& slot = child[2 * node + bit];
      if (b == 0) {
        if (slot != 0) return NVF_ERR_INVALID_ARG;    // not prefix free
        slot = ~c;
      } else {
        if (slot < 0) return NVF_ERR_INVALID_ARG;
        if (slot == 0) { const int32_t nn = new_node(); child[2 * node + bit] = nn; node = nn; }
        else node = slot;
      }
    }
  }
  BitReader r{stream, stream_len};
  const uint64_t total_bits = (uint64_t)stream_len * 8;
  uint64_t used = 0;
  for (int64_t i = 0; i < n_symbols; ++i) {
    int32_t node = 0;
    for (;;) {
      if (used >= total_bits) return NVF_ERR_BITSTREAM;
      const int32_t nx = child[2 * node + r.get()];
      ++used;
      if (nx < 0) { symbols_out[i] = code_symbols[~nx]; break; }
      if (nx == 0) return NVF_ERR_BITSTREAM;
      node = nx;
    }
  }
  return NVF_OK;
}

}  // extern "C"
