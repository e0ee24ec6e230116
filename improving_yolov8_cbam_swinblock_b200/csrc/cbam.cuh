// Shared pieces of the CBAM kernels (cbam.cu: streaming chain + backward; cbam_cluster.cu: resident cluster forward).
#pragma once
#include <algorithm>

#include "common.cuh"

namespace b200 {
namespace cbam {

constexpr int kT = 256;
constexpr int kWarps = kT / 32;
constexpr int KS = 7, PADK = 3, KROW = 8;   // conv taps are stored [2][7][8] (rows padded to 8 floats)
constexpr int NTAPS = 2 * KS * KROW;        // 112 slots, 98 used
constexpr int NT7 = 2 * KS * KS;            // 98

// Host-computed geometry (passed by value in the kernel parameters).
struct Geo {
  int B, C, H, W, HW, r, ksa, mode;
  int nch, lpp, S, np, groups, th, tw, one;
  float invC, invHW;
};

// VW contiguous channels <-> floats.  VW * sizeof(T) is 16 bytes on the fast path ("vectorised NHWC loads": one
// LDG.128 feeds 8 bf16 channels) and one 32-bit word (2 x 16-bit / 1 x f32) when C is not a multiple of that.
template <typename T, int VW> struct alignas(sizeof(T) * VW) VPack { T e[VW]; };
template <typename T, int VW> struct Vec {
  __device__ static __forceinline__ void load(const T* p, float (&v)[VW]) {
    VPack<T, VW> k;
    if constexpr (sizeof(T) * VW == 16) {
      const uint4 raw = ldg_stream16(p);
      k = *reinterpret_cast<const VPack<T, VW>*>(&raw);
    } else {
      k = *reinterpret_cast<const VPack<T, VW>*>(p);
    }
#pragma unroll
    for (int i = 0; i < VW; ++i) v[i] = DT<T>::to_f(k.e[i]);
  }
  __device__ static __forceinline__ void store(T* p, const float (&v)[VW]) {
    VPack<T, VW> k;
#pragma unroll
    for (int i = 0; i < VW; ++i) k.e[i] = DT<T>::from_f(v[i]);
    if constexpr (sizeof(T) * VW == 16) stg_stream16(p, *reinterpret_cast<const uint4*>(&k));
    else *reinterpret_cast<VPack<T, VW>*>(p) = k;
  }
};
template <typename T> struct Words { static constexpr int EPL = 4 / (int)sizeof(T), VE = 16 / (int)sizeof(T); };
template <typename T> struct Pair2 { using type = float2; };
template <> struct Pair2<__nv_bfloat16> { using type = __nv_bfloat162; };
template <> struct Pair2<__half> { using type = __half2; };
__device__ __forceinline__ __nv_bfloat162 make_pair(__nv_bfloat16, float a, float b) { return __floats2bfloat162_rn(a, b); }
__device__ __forceinline__ __half2 make_pair(__half, float a, float b) { return __floats2half2_rn(a, b); }
__device__ __forceinline__ float relu_nan(float a) { return (a != a) ? a : fmaxf(a, 0.f); }  // torch.relu keeps NaN

// conv weights -> smem [2][7][8], a 3x3 kernel centred in zeros
__device__ __forceinline__ void load_taps(float* wsas, const float* wsa, int ks) {
  for (int i = threadIdx.x; i < NTAPS; i += kT) {
    const int ch = i / (KS * KROW), u = (i / KROW) % KS, v = i % KROW;
    float w = 0.f;
    if (wsa && v < KS) {
      const int o = (KS - ks) / 2, uu = u - o, vv = v - o;
      if (uu >= 0 && uu < ks && vv >= 0 && vv < ks) w = wsa[(ch * ks + uu) * ks + vv];
    }
    wsas[i] = w;
  }
}


inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

// stash: pooled[B][2][C] f32 | amx[B][C] i32 | maps[B][3][HW] (mean_c, max_c f32, argmax_c i32)
struct Stash { float* pooled; int* amx; float* maps; size_t bytes; };
inline Stash carve_stash(void* base, int B, int C, int HW) {
  char* p = (char*)base;
  Stash s;
  s.pooled = (float*)p; p += up256((size_t)B * 2 * C * 4);
  s.amx = (int*)p; p += up256((size_t)B * C * 4);
  s.maps = (float*)p; p += up256((size_t)B * 3 * HW * 4);
  s.bytes = (size_t)(p - (char*)base);
  return s;
}

// Resident cluster forward (cbam_cluster.cu).  Returns -1 when the shape does not qualify (caller then runs the
// streaming chain), else a B200_* code.  maps/pooled/amx == NULL: nothing is stashed.
int cluster_fwd(const void* x, const float* w1, const float* w2, const float* wsa, void* out, float* ca_out, float* sa_out,
                const Stash* stash, int B, int C, int H, int W, int r, int ksa, int dtype, int mode, int vw, cudaStream_t st);

}  // namespace cbam
}  // namespace b200
