// Channel concat of NHWC (channels_last) maps: the seam either side of the blocks (`C2f` chunk/cat,
// `Concat`, the Detect head's cat -- block.py C2f.forward, conv.py Concat.forward, head.py:66-76).
//
// Every source is a row-strided view `[rows = B*H*W, C_i]` (row stride S_i >= C_i elements: a channel slice of a wider
// channels_last tensor qualifies, which is what `chunk(2, 1)` and the gradient of a concat hand over), the destination is
// dense `[rows, sum C_i]`.  One thread moves one 16-byte vector; four independent vectors in flight per thread.  ATen's
// CatArrayBatchedCopy treats channels_last inputs as generic strided tensors and reaches ~15 % of the copy roofline on
// these shapes; this is a plain vectorised copy.
#include "common.cuh"

namespace b200 {
namespace {

constexpr int kMaxSrc = 8;
constexpr int kT = 256;
constexpr int kInFlight = 4;

struct FastDiv {   // exact 32-bit division by a run-time constant (multiply-high + shifts)
  uint32_t d, m, s1, s2;
  void init(uint32_t div) {
    d = div;
    uint32_t l = 0;
    while ((1ull << l) < div) ++l;
    m = (uint32_t)((((1ull << 32) * ((1ull << l) - div)) / div) + 1);
    s1 = l < 1 ? l : 1;
    s2 = l > 0 ? l - 1 : 0;
  }
  __device__ __forceinline__ uint32_t div(uint32_t n) const {
    const uint32_t t = __umulhi(m, n);
    return (t + ((n - t) >> s1)) >> s2;
  }
};

struct CatParams {
  const unsigned char* src[kMaxSrc];
  long long stride[kMaxSrc];   // source row stride in bytes
  uint32_t first[kMaxSrc + 1]; // first unit (16-byte vector or element) of source i inside a destination row
  unsigned char* dst;
  uint32_t n, upr;             // sources, units per destination row
  uint32_t items;              // rows * upr
  FastDiv dupr;
};

template <typename V>   // V = uint4 (16-byte vectors) or a 2/4-byte scalar
__global__ void __launch_bounds__(kT) nhwc_concat_kernel(const __grid_constant__ CatParams P) {
  const uint32_t step = gridDim.x * kT;
  for (uint32_t i0 = blockIdx.x * kT + threadIdx.x; i0 < P.items; i0 += kInFlight * step) {
    V v[kInFlight];
#pragma unroll
    for (int u = 0; u < kInFlight; ++u) {
      const uint32_t i = i0 + u * step;
      if (i < P.items) {
        const uint32_t row = P.dupr.div(i), col = i - row * P.upr;
        uint32_t j = 0;
#pragma unroll
        for (int k = 1; k < kMaxSrc; ++k) j += (k < (int)P.n && col >= P.first[k]) ? 1u : 0u;
        v[u] = *reinterpret_cast<const V*>(P.src[j] + (long long)row * P.stride[j] + (size_t)(col - P.first[j]) * sizeof(V));
      }
    }
#pragma unroll
    for (int u = 0; u < kInFlight; ++u) {
      const uint32_t i = i0 + u * step;
      if (i < P.items) *reinterpret_cast<V*>(P.dst + (size_t)i * sizeof(V)) = v[u];
    }
  }
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" B200_API int b200_nhwc_concat(const void* const* srcs, const int32_t* src_channels, const int64_t* src_row_stride,
                                         int32_t n_src, void* dst, int64_t rows, int32_t dtype, void* stream) {
  B200_REQUIRE(srcs && src_channels && src_row_stride && dst, B200_ERR_SHAPE, "nhwc_concat: null pointer");
  B200_REQUIRE(n_src >= 1 && n_src <= kMaxSrc, B200_ERR_UNSUPPORTED, "nhwc_concat: %d sources (1..%d supported)", n_src, kMaxSrc);
  B200_REQUIRE(dtype == B200_F32 || dtype == B200_BF16 || dtype == B200_F16, B200_ERR_DTYPE, "nhwc_concat: unsupported dtype code %d", dtype);
  B200_REQUIRE(rows > 0, B200_ERR_SHAPE, "nhwc_concat: rows=%lld", (long long)rows);
  const int es = dtype == B200_F32 ? 4 : 2;
  long long ctot = 0;
  bool vec = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
  for (int i = 0; i < n_src; ++i) {
    B200_REQUIRE(srcs[i] && src_channels[i] > 0 && src_row_stride[i] >= src_channels[i], B200_ERR_SHAPE,
                 "nhwc_concat: source %d: channels=%d row stride=%lld", i, src_channels[i], (long long)src_row_stride[i]);
    B200_REQUIRE((reinterpret_cast<uintptr_t>(srcs[i]) & (es - 1)) == 0, B200_ERR_ALIGN, "nhwc_concat: source %d is not element-aligned", i);
    ctot += src_channels[i];
    vec = vec && (reinterpret_cast<uintptr_t>(srcs[i]) & 15) == 0 && ((long long)src_channels[i] * es) % 16 == 0 &&
          ((long long)src_row_stride[i] * es) % 16 == 0;
  }
  const int unit = vec ? 16 : es;
  const long long upr = ctot * es / unit;
  B200_REQUIRE(rows * upr < (1ll << 31), B200_ERR_UNSUPPORTED, "nhwc_concat: %lld x %lld units exceed the 31-bit item range",
               (long long)rows, upr);
  CatParams P;
  uint32_t first = 0;
  for (int i = 0; i < kMaxSrc; ++i) {
    const int k = i < n_src ? i : n_src - 1;
    P.src[i] = static_cast<const unsigned char*>(srcs[k]);
    P.stride[i] = (long long)src_row_stride[k] * es;
    P.first[i] = first;
    if (i < n_src) first += (uint32_t)((long long)src_channels[i] * es / unit);
  }
  P.first[kMaxSrc] = first;
  P.dst = static_cast<unsigned char*>(dst);
  P.n = (uint32_t)n_src; P.upr = (uint32_t)upr; P.items = (uint32_t)(rows * upr);
  P.dupr.init((uint32_t)upr);
  long long blocks = ((long long)P.items + (long long)kT * kInFlight - 1) / ((long long)kT * kInFlight);
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  if (vec) nhwc_concat_kernel<uint4><<<(unsigned)blocks, kT, 0, st>>>(P);
  else if (es == 4) nhwc_concat_kernel<uint32_t><<<(unsigned)blocks, kT, 0, st>>>(P);
  else nhwc_concat_kernel<uint16_t><<<(unsigned)blocks, kT, 0, st>>>(P);
  return check_launch("nhwc_concat");
}
