// Shared by sppf_pool.cu (generic planes) and sppf_strip.cu (20x20 planes, strips in registers): the 32-bit lane word of the
// SPPF kernels and the exact scalar 1-D max pass with ATen's update rule.
#pragma once
#include "common.cuh"

namespace b200 {
namespace sppf {

template <typename T> struct Word;  // 32-bit lane word <-> EPL elements
template <> struct Word<float> {
  static constexpr int EPL = 1;
  __device__ static __forceinline__ float get(uint32_t w, int) { return __uint_as_float(w); }
  __device__ static __forceinline__ uint32_t neg_inf() { return 0xff800000u; }
  __device__ static __forceinline__ uint32_t put(uint32_t w, int, uint32_t bits) { return bits; }
  __device__ static __forceinline__ uint32_t bits(uint32_t w, int) { return w; }
  __device__ static __forceinline__ float bits_to_f(uint32_t b) { return __uint_as_float(b); }
  __device__ static __forceinline__ uint32_t f_to_bits(float f) { return __float_as_uint(f); }
  // exact-selection max, valid when the tile holds no NaN and no -0.0 (checked at load time)
  __device__ static __forceinline__ uint32_t vmax(uint32_t a, uint32_t b) {
    return __float_as_uint(fmaxf(__uint_as_float(a), __uint_as_float(b)));
  }
  __device__ static __forceinline__ bool special(uint32_t w) { return (w & 0x7fffffffu) > 0x7f800000u || w == 0x80000000u; }
};
template <> struct Word<__nv_bfloat16> {
  static constexpr int EPL = 2;
  __device__ static __forceinline__ uint32_t neg_inf() { return 0xff80ff80u; }
  __device__ static __forceinline__ uint32_t bits(uint32_t w, int e) { return (w >> (16 * e)) & 0xffffu; }
  __device__ static __forceinline__ float bits_to_f(uint32_t b) { return __uint_as_float(b << 16); }
  __device__ static __forceinline__ uint32_t put(uint32_t w, int e, uint32_t b) {
    return e ? ((w & 0x0000ffffu) | (b << 16)) : ((w & 0xffff0000u) | b);
  }
  __device__ static __forceinline__ uint32_t f_to_bits(float f) {
    return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(f));
  }
  __device__ static __forceinline__ uint32_t vmax(uint32_t a, uint32_t b) {
    const __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
  __device__ static __forceinline__ bool special(uint32_t w) {
    const uint32_t lo = w & 0xffffu, hi = w >> 16;
    return (lo & 0x7fffu) > 0x7f80u || (hi & 0x7fffu) > 0x7f80u || lo == 0x8000u || hi == 0x8000u;
  }
};
template <> struct Word<__half> {
  static constexpr int EPL = 2;
  __device__ static __forceinline__ uint32_t neg_inf() { return 0xfc00fc00u; }
  __device__ static __forceinline__ uint32_t bits(uint32_t w, int e) { return (w >> (16 * e)) & 0xffffu; }
  __device__ static __forceinline__ float bits_to_f(uint32_t b) { return __half2float(__ushort_as_half((unsigned short)b)); }
  __device__ static __forceinline__ uint32_t put(uint32_t w, int e, uint32_t b) {
    return e ? ((w & 0x0000ffffu) | (b << 16)) : ((w & 0xffff0000u) | b);
  }
  __device__ static __forceinline__ uint32_t f_to_bits(float f) {
    return (uint32_t)__half_as_ushort(__float2half_rn(f));
  }
  __device__ static __forceinline__ uint32_t vmax(uint32_t a, uint32_t b) {
    const __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
  __device__ static __forceinline__ bool special(uint32_t w) {
    const uint32_t lo = w & 0xffffu, hi = w >> 16;
    return (lo & 0x7fffu) > 0x7c00u || (hi & 0x7fffu) > 0x7c00u || lo == 0x8000u || hi == 0x8000u;
  }
};

// One 1-D max pass over a strip for one lane word.  src/dst are strided smem word arrays.
//   len    : strip length,  K: window, r = K/2
//   win_out: if non-null, receives the winner's OFFSET in the window [0,K) per element (4 bits each, packed
//            as (e*4) nibbles in a byte... one byte per element stored in a uint8 array with its own stride).
// Returns nothing; writes dst[i*dstride] for i in [0,len).
template <typename T, int K>
__device__ __forceinline__ void pass1d(const uint32_t* __restrict__ src, int sstride, uint32_t* __restrict__ dst,
                                       int dstride, int len, uint8_t* __restrict__ win_out, int wstride, int i0 = 0,
                                       int i1 = -1) {
  if (i1 < 0) i1 = len;   // [i0, i1): the outputs this thread produces (a strip may be split over several threads)
  using WD = Word<T>;
  constexpr int R = K / 2;
  constexpr int EPL = WD::EPL;
  uint32_t win[K];  // sliding window of words, win[j] = src[i - R + j] (or -inf outside)
#pragma unroll
  for (int j = 0; j < K; ++j) {
    int s = i0 + j - R;  // position for the first output i0 is i0 + j - R
    win[j] = (s >= 0 && s < len) ? src[s * sstride] : WD::neg_inf();
  }
  for (int i = i0; i < i1; ++i) {
    uint32_t outw = 0;
    uint32_t wsel = 0;
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      uint32_t bb = WD::bits(WD::neg_inf(), e);
      float bf = -INFINITY;
      int lo = i - R < 0 ? R - i : 0;  // first in-bounds window slot (ATen's initial index)
      int bj = lo;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        uint32_t vb = WD::bits(win[j], e);
        float vf = WD::bits_to_f(vb);
        bool take = (vf > bf) || (vf != vf);
        // out-of-range slots hold -inf and can never be taken (-inf > x is false, not NaN)
        bb = take ? vb : bb;
        bf = take ? vf : bf;
        bj = take ? j : bj;
      }
      outw = WD::put(outw, e, bb);
      wsel |= (uint32_t)bj << (4 * e);
    }
    dst[i * dstride] = outw;
    if (win_out) win_out[i * wstride] = (uint8_t)wsel;
    // slide
#pragma unroll
    for (int j = 0; j < K - 1; ++j) win[j] = win[j + 1];
    int s = i + 1 + R;
    win[K - 1] = (s < len) ? src[s * sstride] : WD::neg_inf();
  }
}

// sppf_strip.cu: the 20x20 / 16-bit kernels with strips in registers.  Return -1 when the shape is not theirs (the caller then
// runs the generic kernels of sppf_pool.cu), else B200_OK or an error code.
int strip_fwd(const void* y0, void* cat, int B, int C, int H, int W, int k, int dtype, cudaStream_t st);
int strip_bwd(const void* gcat, const void* y0, void* gy0, int B, int C, int H, int W, int k, int dtype, cudaStream_t st);

}  // namespace sppf
}  // namespace b200
