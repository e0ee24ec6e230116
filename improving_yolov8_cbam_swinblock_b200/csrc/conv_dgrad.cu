// Input gradient of the narrow stride-2 3x3 convolution at the top of the model (yaml backbone row 1 at scale n:
// nn.Conv2d(16, 32, 3, 2, 1, bias=False); conv.py:37-91 `self.conv`, autograd of F.conv2d w.r.t. the input).  cuDNN runs its
// generic strided-dgrad kernel there: 0.38 ms for 52 MB in / 210 MB out, 8x the HBM time, the slowest single launch of the step.
//
//   gx[b, iy, ix, ci] = sum over (ky, kx, oc) with iy = 2*oy - 1 + ky, ix = 2*ox - 1 + kx of gy[b, oy, ox, oc] * w[oc, ci, ky, kx]
//
// The parity of (iy, ix) decides which taps contribute: for iy = 2*py + dy, ix = 2*px + dx
//   dy = 0: ky = 1 (oy = py)              dy = 1: ky = 0 (oy = py + 1), ky = 2 (oy = py)        -- the same in x --
// so the four parity classes are four small implicit GEMMs with 1, 2, 2 and 4 taps: [16 pixels x 32 oc] x [32 oc x 16 ci] per tap
// on mma.sync m16n8k16.  A = gy rows as stored (pixel-major, oc contiguous) read by ldmatrix from a tile staged once per CTA with
// cp.async (rows py .. py + R, one zero pixel past the row end, zero rows past the map: no bounds checks); B = the weight
// re-laid as [tap][ci][oc] in shared memory.  A warp owns 16 consecutive px of one py row = a 2 x 32 block of input pixels: its
// four accumulator sets are interleaved through a per-warp staging tile and leave as 16-byte stores of whole pixel rows.
// No atomics, every output element written exactly once: deterministic.
#include <type_traits>

#include "common.cuh"

namespace b200 {
namespace {

constexpr int kCi = 16, kOc = 32, kR = 4, kWarps = 8, kThreads = kWarps * 32;
constexpr int kGP = 80;    // bytes per gy pixel in shared memory: 64 of data + 16 of padding (ldmatrix rows hit distinct banks)
constexpr int kWP = 80;    // bytes per (tap, ci) weight row: 32 oc + padding
constexpr int kSP = 48;    // bytes per staged output pixel: 32 of data + 16 of padding
constexpr int kStageWarp = 2 * 32 * kSP;

template <typename T> struct MmaD;
template <> struct MmaD<__nv_bfloat16> {
  __device__ static __forceinline__ void run(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  __device__ static __forceinline__ uint32_t pack(float lo, float hi) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&p);
  }
};
template <> struct MmaD<__half> {
  __device__ static __forceinline__ void run(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  __device__ static __forceinline__ uint32_t pack(float lo, float hi) {
    const __half2 p = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&p);
  }
};
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

struct WStride { long long oc, ci, ky, kx; };

// tap -> (A tile: row offset a, pixel offset b) and parity class (dy * 2 + dx); see the header
struct Tap { int ky, kx, a, b, cls; };
__device__ constexpr Tap kTaps[9] = {{1, 1, 0, 0, 0}, {1, 0, 0, 1, 1}, {1, 2, 0, 0, 1}, {0, 1, 1, 0, 2}, {2, 1, 0, 0, 2},
                                     {0, 0, 1, 1, 3}, {0, 2, 1, 0, 3}, {2, 0, 0, 1, 3}, {2, 2, 0, 0, 3}};

template <typename T>
__global__ void __launch_bounds__(kThreads, 2) conv3_dgrad_s2_kernel(const T* __restrict__ gy, const T* __restrict__ w, const WStride ws,
                                                                      T* __restrict__ gx, int H, int W, int Ho, int Wo, int tiles_per_img,
                                                                      int n_tiles) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char smem[];
  const int RP = (Wo + 1) * kGP;                       // bytes per staged gy row
  unsigned char* gs = smem;                            // [kR + 1][Wo + 1][kGP]
  unsigned char* wsm = gs + (size_t)(kR + 1) * RP;     // [9][kCi][kWP]
  unsigned char* stage = wsm + 9 * kCi * kWP;          // [kWarps][2][32][kSP]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = lane & 3, r = lane >> 2;
  // weight -> [tap][ci][oc] in the 16-bit dtype; the zero pixel past the end of every staged row
  for (int i = tid; i < 9 * kCi * kOc; i += kThreads) {
    const int tap = i / (kCi * kOc), rem = i - tap * kCi * kOc, ci = rem / kOc, oc = rem - ci * kOc;
    const int ky = tap / 3, kx = tap - ky * 3;
    *reinterpret_cast<T*>(wsm + (size_t)(tap * kCi + ci) * kWP + oc * 2) = w[oc * ws.oc + ci * ws.ci + ky * ws.ky + kx * ws.kx];
  }
  for (int i = tid; i < (kR + 1) * (kGP / 16); i += kThreads)
    *reinterpret_cast<uint4*>(gs + (size_t)(i / (kGP / 16)) * RP + (size_t)Wo * kGP + (i % (kGP / 16)) * 16) = make_uint4(0, 0, 0, 0);

  const uint32_t gs_u = smem_u32(gs), ws_u = smem_u32(wsm);
  const uint32_t a_lane = (uint32_t)(((lane & 7) + ((lane & 8) ? 8 : 0)) * kGP + ((lane & 16) ? 16 : 0));
  const uint32_t b_lane = (uint32_t)(((lane & 7) + ((lane & 16) ? 8 : 0)) * kWP + ((lane & 8) ? 16 : 0));
  unsigned char* st = stage + (size_t)warp * kStageWarp;
  const int gpr = Wo / 16;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int b = tile / tiles_per_img, py0 = (tile - b * tiles_per_img) * kR;
    __syncthreads();   // previous tile consumed (first pass: weights / zero pixels written)
    for (int i = 0; i <= kR; ++i) {
      const int oy = py0 + i;
      const bool ok = oy < Ho;
      const unsigned char* src = reinterpret_cast<const unsigned char*>(gy) + ((size_t)(b * Ho + (ok ? oy : 0)) * Wo) * (kOc * 2);
      unsigned char* dst = gs + (size_t)i * RP;
      for (int v = tid; v < Wo * 4; v += kThreads) cp_async16(dst + (size_t)(v >> 2) * kGP + (v & 3) * 16, src + (size_t)v * 16, ok);
    }
    cp_async_wait_all();
    __syncthreads();
    for (int g = warp; g < kR * gpr; g += kWarps) {
      const int pr = g / gpr, px0 = (g - pr * gpr) * 16;
      if (py0 + pr >= Ho) break;
      uint32_t A[2][2][2][4];   // [row offset a][pixel offset b][k step: oc 0-15 | 16-31]
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int bb = 0; bb < 2; ++bb)
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            ldsm_x4(A[a][bb][ks], gs_u + (uint32_t)((pr + a) * RP + (px0 + bb) * kGP + ks * 32) + a_lane);
      float acc[4][2][4];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) acc[c][nt][0] = acc[c][nt][1] = acc[c][nt][2] = acc[c][nt][3] = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const Tap tp = kTaps[t];
        const int tap = tp.ky * 3 + tp.kx;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          uint32_t B[4];   // b0 / b1 of ci 0-7, b0 / b1 of ci 8-15
          ldsm_x4(B, ws_u + (uint32_t)(tap * kCi * kWP + ks * 32) + b_lane);
          MmaD<T>::run(acc[tp.cls][0], A[tp.a][tp.b][ks], B[0], B[1]);
          MmaD<T>::run(acc[tp.cls][1], A[tp.a][tp.b][ks], B[2], B[3]);
        }
      }
      // interleave the four parity classes: staged pixel (dy, 2*m + dx), m = the warp's px index
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int dy = c >> 1, dx = c & 1;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          unsigned char* o = st + dy * (32 * kSP) + (2 * r + dx) * kSP + (nt * 8 + 2 * q) * 2;
          *reinterpret_cast<uint32_t*>(o) = MmaD<T>::pack(acc[c][nt][0], acc[c][nt][1]);
          *reinterpret_cast<uint32_t*>(o + 16 * kSP) = MmaD<T>::pack(acc[c][nt][2], acc[c][nt][3]);   // rows r + 8: 16 staged pixels on
        }
      }
      __syncwarp();
      unsigned char* orow = reinterpret_cast<unsigned char*>(gx) + (((size_t)b * H + 2 * (py0 + pr)) * W + 2 * px0) * (kCi * 2);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = lane + 32 * j, dy = c >> 6, rem = c & 63;   // 64 16-byte chunks = one row of 32 input pixels
        const uint4 v = *reinterpret_cast<const uint4*>(st + dy * (32 * kSP) + (rem >> 1) * kSP + (rem & 1) * 16);
        stg_stream16(orow + (size_t)dy * W * (kCi * 2) + (size_t)rem * 16, v);
      }
      __syncwarp();
    }
  }
}

bool dgrad_shape_ok(int H, int W, int cin, int cout, int dtype) {
  if (!(dtype == B200_BF16 || dtype == B200_F16) || cin != kCi || cout != kOc) return false;
  if (H <= 0 || W <= 0 || ((H | W) & 1)) return false;
  const int Wo = W / 2;
  const size_t smem = (size_t)(kR + 1) * (Wo + 1) * kGP + 9 * kCi * kWP + kWarps * kStageWarp;
  return Wo % 16 == 0 && smem <= (size_t)max_smem_optin();
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" B200_API int b200_conv3x3_dgrad_s2_supported(int32_t H, int32_t W, int32_t cin, int32_t cout, int32_t dtype) {
  return dgrad_shape_ok(H, W, cin, cout, dtype) ? 1 : 0;
}

extern "C" B200_API int b200_conv3x3_dgrad_s2(const void* gy, const void* w, const int64_t* w_stride, void* gx, int32_t B, int32_t H, int32_t W,
                                              int32_t cin, int32_t cout, int32_t dtype, void* stream) {
  B200_REQUIRE(gy && w && w_stride && gx, B200_ERR_SHAPE, "conv3x3_dgrad_s2: null pointer");
  B200_REQUIRE(B > 0 && dgrad_shape_ok(H, W, cin, cout, dtype), B200_ERR_UNSUPPORTED,
               "conv3x3_dgrad_s2: unsupported shape H=%d W=%d cin=%d cout=%d dtype=%d", H, W, cin, cout, dtype);
  B200_REQUIRE((((uintptr_t)gy | (uintptr_t)gx) & 15) == 0, B200_ERR_ALIGN, "conv3x3_dgrad_s2: tensors must be 16-byte aligned");
  const int Ho = H / 2, Wo = W / 2, tpi = (Ho + kR - 1) / kR, n_tiles = B * tpi;
  const size_t smem = (size_t)(kR + 1) * (Wo + 1) * kGP + 9 * kCi * kWP + kWarps * kStageWarp;
  const WStride ws{w_stride[0], w_stride[1], w_stride[2], w_stride[3]};
  cudaStream_t st = (cudaStream_t)stream;
  auto go = [&](auto kern, auto* tag) -> int {
    using T = typename std::remove_pointer<decltype(tag)>::type;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem);
    per_sm = per_sm < 1 ? 1 : per_sm > 4 ? 4 : per_sm;
    int grid = sm_count() * per_sm;   // persistent CTAs: one resident wave
    if (grid > n_tiles) grid = n_tiles;
    launch_k(kern, grid, kThreads, smem, st, (const T*)gy, (const T*)w, ws, (T*)gx, H, W, Ho, Wo, tpi, n_tiles);
    return check_launch("conv3x3_dgrad_s2");
  };
  if (dtype == B200_BF16) return go(conv3_dgrad_s2_kernel<__nv_bfloat16>, (__nv_bfloat16*)nullptr);
  return go(conv3_dgrad_s2_kernel<__half>, (__half*)nullptr);
}
