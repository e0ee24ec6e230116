// Fused BatchNorm2d (+ SiLU) epilogue of the stock `Conv` block, forward and backward, NHWC.
// Replaces `self.act(self.bn(...))` of Conv.forward (ultralytics/nn/modules/conv.py:65-79: BatchNorm2d then SiLU) for
// the SPPF cv1/cv2 convolutions (block.py:218-219) and, through the same module, the other
// Conv callers either side of the hot path.  Training mode uses per-GPU batch statistics like the reference (no
// SyncBN) and updates running_mean / running_var exactly like torch (momentum, unbiased variance).
//
// The conv output x is a row-major [M = B*H*W, C] matrix.  A thread owns one 16-byte channel vector (8 bf16 / 4 f32)
// and walks rows; a CTA owns a contiguous block of rows:
//   fwd  stats : per-channel shifted sums (shift = first row of the block, so E[(x-k)^2] - E[x-k]^2 does not cancel)
//                -> per-CTA (mean, M2) partials
//        final : warp per channel: Chan-merge the partials in a fixed order -> mean, rstd, running stats,
//                scale = gamma*rstd, shift = beta - mean*scale
//        apply : z = act(x*scale + shift)                       (x read twice, z written once: 2R + 1W;
//                                                                torch: stats R, transform R+W, SiLU R+W)
//   bwd  reduce: gy = gz * act'(y) recomputed from x; per-channel sum(gy), sum(gy*xhat) partials
//        final : warp per channel -> g_gamma, g_beta
//        apply : gx = gamma*rstd*(gy - sum(gy)/M - xhat*sum(gy*xhat)/M)     (4R + 1W; torch: 6R + 2W)
// Nothing but x, mean and rstd is kept for the backward (torch keeps x, the BN output and mean/rstd).
// Deterministic: no atomics, fixed summation order.
#include "common.cuh"

namespace b200 {
namespace {

constexpr int kT = 256;

struct BnGeo {
  long long M;      // rows
  int C, nch, tpr, ngrp, rpt, rpb, G;   // vectors/row, threads/row, row groups/CTA, rows/thread, rows/CTA, CTAs
  float invM;
};

template <typename T, int VW> struct alignas(sizeof(T) * VW) VP { T e[VW]; };
template <typename T, int VW>
__device__ __forceinline__ void ldv(const T* p, float (&v)[VW]) {
  const uint4 raw = ldg_stream16(p);
  const VP<T, VW> k = *reinterpret_cast<const VP<T, VW>*>(&raw);
#pragma unroll
  for (int i = 0; i < VW; ++i) v[i] = DT<T>::to_f(k.e[i]);
}
template <typename T, int VW>
__device__ __forceinline__ void stv(T* p, const float (&v)[VW]) {
  VP<T, VW> k;
#pragma unroll
  for (int i = 0; i < VW; ++i) k.e[i] = DT<T>::from_f(v[i]);
  stg_stream16(p, *reinterpret_cast<const uint4*>(&k));
}

__device__ __forceinline__ float silu_f(float y) { return y * __frcp_rn(1.f + __expf(-y)); }
__device__ __forceinline__ float dsilu_f(float y) {
  const float s = __frcp_rn(1.f + __expf(-y));
  return s * (1.f + y * (1.f - s));
}

// per-CTA additive reduction of NV values per channel held by the row groups: red[ngrp][NV][C] -> out via f(c, v[NV])
template <int NV, typename F>
__device__ __forceinline__ void cta_fold(float* red, int C, int ngrp, F emit) {
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kT) {
    float acc[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[j] = 0.f;
    for (int g = 0; g < ngrp; ++g) {
#pragma unroll
      for (int j = 0; j < NV; ++j) acc[j] += red[((size_t)g * NV + j) * C + c];
    }
    emit(c, acc);
  }
}

// ---- forward statistics: per-CTA (mean, M2) of every channel over the CTA's rows ------------------------------
template <typename T, int VW>
__global__ void __launch_bounds__(kT) bn_stats_kernel(const T* __restrict__ x, float* __restrict__ part, const BnGeo G) {
  extern __shared__ __align__(16) float red[];   // [ngrp][2][C]
  const int C = G.C, tpr = G.tpr;
  const int rg = threadIdx.x / tpr, ch = threadIdx.x - rg * tpr;
  const long long r0 = (long long)blockIdx.x * G.rpb;
  const int nrows = (int)min((long long)G.rpb, G.M - r0);
  float s[VW], q[VW], k[VW];
#pragma unroll
  for (int e = 0; e < VW; ++e) { s[e] = 0.f; q[e] = 0.f; }
  if (rg < G.ngrp) {
    const T* base = x + (size_t)r0 * C + ch * VW;
    ldv<T, VW>(base, k);   // the block's first row: common shift of this CTA
#pragma unroll 4
    for (int r = rg; r < nrows; r += G.ngrp) {
      float v[VW];
      ldv<T, VW>(base + (size_t)r * C, v);
#pragma unroll
      for (int e = 0; e < VW; ++e) { const float d = v[e] - k[e]; s[e] += d; q[e] += d * d; }
    }
    float* rs = red + (size_t)rg * 2 * C + ch * VW;
#pragma unroll
    for (int e = 0; e < VW; ++e) { rs[e] = s[e]; rs[C + e] = q[e]; }
  }
  const float n = (float)nrows;
  float* dst = part + (size_t)blockIdx.x * 2 * C;
  const T* krow = x + (size_t)r0 * C;
  cta_fold<2>(red, C, G.ngrp, [&](int c, const float (&a)[2]) {
    const float kk = DT<T>::to_f(krow[c]);
    const float m = a[0] / n;
    dst[c] = kk + m;
    dst[C + c] = fmaxf(a[1] - a[0] * m, 0.f);
  });
}

// ---- forward finalize: warp per channel, Chan merge of the G partials (fixed order) -----------------------------
__global__ void __launch_bounds__(kT) bn_final_kernel(const float* __restrict__ part, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, float* __restrict__ running_mean,
                                                      float* __restrict__ running_var, float* __restrict__ mean_out,
                                                      float* __restrict__ rstd_out, float* __restrict__ scsh, float eps,
                                                      float momentum, int training, const BnGeo G) {
  const int lane = threadIdx.x & 31, c = blockIdx.x * (kT / 32) + (threadIdx.x >> 5), C = G.C;
  if (c >= C) return;
  float mean, rstd;
  if (training) {
    float n = 0.f, mu = 0.f, m2 = 0.f;
    for (int g = lane; g < G.G; g += 32) {
      const float nb = (float)min((long long)G.rpb, G.M - (long long)g * G.rpb);
      const float mb = part[(size_t)g * 2 * C + c], qb = part[(size_t)g * 2 * C + C + c];
      const float nt = n + nb, d = mb - mu;
      mu += d * (nb / nt);
      m2 += qb + d * d * (n * nb / nt);
      n = nt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float nb = __shfl_xor_sync(0xffffffffu, n, o), mb = __shfl_xor_sync(0xffffffffu, mu, o),
                  qb = __shfl_xor_sync(0xffffffffu, m2, o);
      const float nt = n + nb;
      if (nt > 0.f) {
        // symmetric form: both partners compute the same merged triple
        const float d = mb - mu;
        const float mu_new = (n * mu + nb * mb) / nt;
        m2 = m2 + qb + d * d * (n * nb / nt);
        mu = mu_new;
        n = nt;
      }
    }
    mean = mu;
    const float var = m2 * G.invM;   // biased variance normalises (torch BatchNorm)
    rstd = rsqrtf(var + eps);
    if (lane == 0 && running_mean) {
      const float unb = G.M > 1 ? m2 / (float)(G.M - 1) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unb;
    }
  } else {
    mean = running_mean[c];
    rstd = rsqrtf(running_var[c] + eps);
  }
  if (lane == 0) {
    if (mean_out) { mean_out[c] = mean; rstd_out[c] = rstd; }
    const float sc = gamma[c] * rstd;
    scsh[c] = sc;
    scsh[C + c] = beta[c] - mean * sc;
  }
}

// ---- forward apply: z = act(x*scale + shift) ---------------------------------------------------------------
template <typename T, int VW, int ACT>
__global__ void __launch_bounds__(kT) bn_apply_kernel(const T* __restrict__ x, const float* __restrict__ scsh,
                                                      T* __restrict__ z, const BnGeo G) {
  const int C = G.C, tpr = G.tpr;
  const int rg = threadIdx.x / tpr, ch = threadIdx.x - rg * tpr;
  if (rg >= G.ngrp) return;
  const long long r0 = (long long)blockIdx.x * G.rpb;
  const int nrows = (int)min((long long)G.rpb, G.M - r0);
  float sc[VW], sh[VW];
#pragma unroll
  for (int e = 0; e < VW; ++e) { sc[e] = scsh[ch * VW + e]; sh[e] = scsh[C + ch * VW + e]; }
  const T* xb = x + (size_t)r0 * C + ch * VW;
  T* zb = z + (size_t)r0 * C + ch * VW;
#pragma unroll 4
  for (int r = rg; r < nrows; r += G.ngrp) {
    float v[VW];
    ldv<T, VW>(xb + (size_t)r * C, v);
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      const float y = v[e] * sc[e] + sh[e];
      v[e] = ACT ? silu_f(y) : y;
    }
    stv<T, VW>(zb + (size_t)r * C, v);
  }
}

// ---- backward reduce: per-CTA sum(gy), sum(gy*xhat) with gy = gz*act'(y) recomputed from x -----------------------
template <typename T, int VW, int ACT>
__global__ void __launch_bounds__(kT) bn_bwd_reduce_kernel(const T* __restrict__ x, const T* __restrict__ gz,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           float* __restrict__ part, const BnGeo G) {
  extern __shared__ __align__(16) float red[];   // [ngrp][2][C]
  const int C = G.C, tpr = G.tpr;
  const int rg = threadIdx.x / tpr, ch = threadIdx.x - rg * tpr;
  const long long r0 = (long long)blockIdx.x * G.rpb;
  const int nrows = (int)min((long long)G.rpb, G.M - r0);
  if (rg < G.ngrp) {
    float mu[VW], rs[VW], ga[VW], be[VW], a1[VW], a2[VW];
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      const int c = ch * VW + e;
      mu[e] = mean[c]; rs[e] = rstd[c]; ga[e] = gamma[c]; be[e] = beta[c]; a1[e] = 0.f; a2[e] = 0.f;
    }
    const T* xb = x + (size_t)r0 * C + ch * VW;
    const T* gb = gz + (size_t)r0 * C + ch * VW;
#pragma unroll 2
    for (int r = rg; r < nrows; r += G.ngrp) {
      float v[VW], g[VW];
      ldv<T, VW>(xb + (size_t)r * C, v);
      ldv<T, VW>(gb + (size_t)r * C, g);
#pragma unroll
      for (int e = 0; e < VW; ++e) {
        const float xh = (v[e] - mu[e]) * rs[e];
        const float gy = ACT ? g[e] * dsilu_f(xh * ga[e] + be[e]) : g[e];
        a1[e] += gy;
        a2[e] += gy * xh;
      }
    }
    float* rr = red + (size_t)rg * 2 * C + ch * VW;
#pragma unroll
    for (int e = 0; e < VW; ++e) { rr[e] = a1[e]; rr[C + e] = a2[e]; }
  }
  float* dst = part + (size_t)blockIdx.x * 2 * C;
  cta_fold<2>(red, C, G.ngrp, [&](int c, const float (&a)[2]) { dst[c] = a[0]; dst[C + c] = a[1]; });
}

// ---- backward finalize: warp per channel -> g_gamma = sum(gy*xhat), g_beta = sum(gy) --------------------------------
__global__ void __launch_bounds__(kT) bn_bwd_final_kernel(const float* __restrict__ part, float* __restrict__ ggamma,
                                                          float* __restrict__ gbeta, const BnGeo G) {
  const int lane = threadIdx.x & 31, c = blockIdx.x * (kT / 32) + (threadIdx.x >> 5), C = G.C;
  if (c >= C) return;
  float a1 = 0.f, a2 = 0.f;
  for (int g = lane; g < G.G; g += 32) { a1 += part[(size_t)g * 2 * C + c]; a2 += part[(size_t)g * 2 * C + C + c]; }
  a1 = warp_sum(a1);
  a2 = warp_sum(a2);
  if (lane == 0) { gbeta[c] = a1; ggamma[c] = a2; }
}

// ---- backward apply: gx = gamma*rstd*(gy - sum(gy)/M - xhat*sum(gy*xhat)/M) -----------------------------------------
template <typename T, int VW, int ACT>
__global__ void __launch_bounds__(kT) bn_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ gz,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          const float* __restrict__ mean, const float* __restrict__ rstd,
                                                          const float* __restrict__ ggamma, const float* __restrict__ gbeta,
                                                          T* __restrict__ gx, int training, const BnGeo G) {
  const int C = G.C, tpr = G.tpr;
  const int rg = threadIdx.x / tpr, ch = threadIdx.x - rg * tpr;
  if (rg >= G.ngrp) return;
  const long long r0 = (long long)blockIdx.x * G.rpb;
  const int nrows = (int)min((long long)G.rpb, G.M - r0);
  float mu[VW], rs[VW], ga[VW], be[VW], k1[VW], k2[VW];
#pragma unroll
  for (int e = 0; e < VW; ++e) {
    const int c = ch * VW + e;
    mu[e] = mean[c]; rs[e] = rstd[c]; ga[e] = gamma[c]; be[e] = beta[c];
    k1[e] = training ? gbeta[c] * G.invM : 0.f;    // eval mode: statistics are constants, gx = gamma*rstd*gy
    k2[e] = training ? ggamma[c] * G.invM : 0.f;
  }
  const T* xb = x + (size_t)r0 * C + ch * VW;
  const T* gb = gz + (size_t)r0 * C + ch * VW;
  T* ob = gx + (size_t)r0 * C + ch * VW;
#pragma unroll 2
  for (int r = rg; r < nrows; r += G.ngrp) {
    float v[VW], g[VW];
    ldv<T, VW>(xb + (size_t)r * C, v);
    ldv<T, VW>(gb + (size_t)r * C, g);
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      const float xh = (v[e] - mu[e]) * rs[e];
      const float gy = ACT ? g[e] * dsilu_f(xh * ga[e] + be[e]) : g[e];
      v[e] = ga[e] * rs[e] * (gy - k1[e] - xh * k2[e]);
    }
    stv<T, VW>(ob + (size_t)r * C, v);
  }
}

void fill_geo(BnGeo& G, long long M, int C, int vw) {
  G.M = M; G.C = C; G.nch = C / vw; G.tpr = G.nch; G.ngrp = kT / G.tpr;
  long long rpt = (M + (long long)592 * G.ngrp - 1) / ((long long)592 * G.ngrp);   // ~4 CTAs per SM ...
  if (rpt < 4) rpt = 4;
  while ((M + rpt * G.ngrp - 1) / (rpt * G.ngrp) > 2368) rpt *= 2;                  // ... and at most 16 per SM
  G.rpt = (int)rpt;
  G.rpb = G.rpt * G.ngrp;
  G.G = (int)((M + G.rpb - 1) / G.rpb);
  G.invM = 1.f / (float)M;
}

int make_geo(BnGeo& G, long long M, int C, int dtype, const void* p0, const void* p1, const void* p2) {
  const int vw = dtype == B200_F32 ? 4 : 8;
  B200_REQUIRE(M > 0 && C > 0, B200_ERR_SHAPE, "bn_silu: bad shape rows=%lld C=%d", M, C);
  B200_REQUIRE(dtype == B200_F32 || dtype == B200_BF16 || dtype == B200_F16, B200_ERR_DTYPE, "bn_silu: unsupported dtype code %d", dtype);
  B200_REQUIRE(C % vw == 0 && C / vw <= kT, B200_ERR_UNSUPPORTED, "bn_silu: C=%d must be a multiple of %d and <= %d", C, vw, vw * kT);
  B200_REQUIRE((((uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)p2) & 15) == 0, B200_ERR_ALIGN, "bn_silu: tensors must be 16-byte aligned");
  fill_geo(G, M, C, vw);
  return B200_OK;
}

inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace
}  // namespace b200

extern "C" B200_API int b200_bn_silu_supported(int64_t rows, int32_t C, int32_t dtype) {
  const int vw = dtype == B200_F32 ? 4 : 8;
  return rows > 0 && C > 0 && (dtype >= 0 && dtype <= 2) && C % vw == 0 && C / vw <= b200::kT;
}

extern "C" B200_API size_t b200_bn_silu_workspace_bytes(int64_t rows, int32_t C, int32_t dtype) {
  if (!b200_bn_silu_supported(rows, C, dtype)) return 0;
  b200::BnGeo G;
  b200::fill_geo(G, rows, C, dtype == B200_F32 ? 4 : 8);
  return b200::up256((size_t)G.G * 2 * C * 4) + b200::up256((size_t)2 * C * 4);   // per-CTA partials + scale/shift
}

extern "C" B200_API int b200_bn_silu_fwd(const void* x, const float* gamma, const float* beta, float* running_mean,
                                         float* running_var, void* z, float* mean_out, float* rstd_out, void* workspace,
                                         size_t workspace_bytes, int64_t rows, int32_t C, float eps, float momentum,
                                         int32_t training, int32_t act, int32_t dtype, void* stream) {
  using namespace b200;
  B200_REQUIRE(x && gamma && beta && z, B200_ERR_SHAPE, "bn_silu_fwd: null tensor pointer");
  B200_REQUIRE(training || (running_mean && running_var), B200_ERR_SHAPE, "bn_silu_fwd: eval mode needs running statistics");
  BnGeo G;
  if (int rc = make_geo(G, rows, C, dtype, x, z, nullptr)) return rc;
  const size_t need = up256((size_t)G.G * 2 * C * 4) + up256((size_t)2 * C * 4);
  B200_REQUIRE(workspace && workspace_bytes >= need, B200_ERR_WORKSPACE, "bn_silu_fwd: workspace %zu < %zu bytes", workspace_bytes, need);
  float* part = (float*)workspace;
  float* scsh = (float*)((char*)workspace + up256((size_t)G.G * 2 * C * 4));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)G.ngrp * 2 * C * 4;
  return B200_DISPATCH_DTYPE(dtype, [&]() -> int {
    constexpr int VW = 16 / (int)sizeof(T);
    if (training) bn_stats_kernel<T, VW><<<G.G, kT, smem, st>>>((const T*)x, part, G);
    bn_final_kernel<<<(C + 7) / 8, kT, 0, st>>>(part, gamma, beta, running_mean, running_var, mean_out, rstd_out, scsh, eps,
                                               momentum, training, G);
    if (act) bn_apply_kernel<T, VW, 1><<<G.G, kT, 0, st>>>((const T*)x, scsh, (T*)z, G);
    else bn_apply_kernel<T, VW, 0><<<G.G, kT, 0, st>>>((const T*)x, scsh, (T*)z, G);
    return check_launch("bn_silu_fwd");
  });
}

extern "C" B200_API int b200_bn_silu_bwd(const void* gz, const void* x, const float* gamma, const float* beta,
                                         const float* mean, const float* rstd, void* gx, float* ggamma, float* gbeta,
                                         void* workspace, size_t workspace_bytes, int64_t rows, int32_t C, int32_t training,
                                         int32_t act, int32_t dtype, void* stream) {
  using namespace b200;
  B200_REQUIRE(gz && x && gamma && beta && mean && rstd && gx && ggamma && gbeta, B200_ERR_SHAPE, "bn_silu_bwd: null tensor pointer");
  BnGeo G;
  if (int rc = make_geo(G, rows, C, dtype, x, gz, gx)) return rc;
  const size_t need = up256((size_t)G.G * 2 * C * 4);
  B200_REQUIRE(workspace && workspace_bytes >= need, B200_ERR_WORKSPACE, "bn_silu_bwd: workspace %zu < %zu bytes", workspace_bytes, need);
  float* part = (float*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)G.ngrp * 2 * C * 4;
  return B200_DISPATCH_DTYPE(dtype, [&]() -> int {
    constexpr int VW = 16 / (int)sizeof(T);
    if (act) bn_bwd_reduce_kernel<T, VW, 1><<<G.G, kT, smem, st>>>((const T*)x, (const T*)gz, gamma, beta, mean, rstd, part, G);
    else bn_bwd_reduce_kernel<T, VW, 0><<<G.G, kT, smem, st>>>((const T*)x, (const T*)gz, gamma, beta, mean, rstd, part, G);
    bn_bwd_final_kernel<<<(C + 7) / 8, kT, 0, st>>>(part, ggamma, gbeta, G);
    if (act) bn_bwd_apply_kernel<T, VW, 1><<<G.G, kT, 0, st>>>((const T*)x, (const T*)gz, gamma, beta, mean, rstd, ggamma, gbeta, (T*)gx, training, G);
    else bn_bwd_apply_kernel<T, VW, 0><<<G.G, kT, 0, st>>>((const T*)x, (const T*)gz, gamma, beta, mean, rstd, ggamma, gbeta, (T*)gx, training, G);
    return check_launch("bn_silu_bwd");
  });
}
