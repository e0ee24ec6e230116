// CBAM (channel attention -> spatial attention -> gate), forward and backward, one launch each.
// Replaces cbam.py:29-38 / :48-53 / :62-71 of the reference (and their autograd backward).
//
// One thread-block CLUSTER per image.  The image (NHWC, so a pixel range is one contiguous byte range) is split
// into CS contiguous pixel chunks, one per CTA of the cluster, and each chunk is pulled into shared memory ONCE
// by a 1-D bulk TMA copy (cp.async.bulk -> UBLKCP).  Everything else happens on chip:
//   A  per-channel sum / max over the chunk               -> partials in smem
//      [cluster.sync]  rank k reduces the partials of channel slice S_k over all ranks (DSMEM), multiplies with
//                      W1[:, S_k]                          -> partial hidden activations
//      [cluster.sync]  all ranks sum the partial hiddens, rank k computes ca for S_k with W2[S_k, :]
//      [cluster.sync]  all ranks gather the full ca vector (DSMEM)
//   B  per-pixel mean_c / max_c of x*ca over the chunk     -> 2-channel map chunk in smem
//      [cluster.sync]  gather the +-(pad*W+pad) halo of the map from neighbouring ranks (DSMEM)
//   C  k x k conv + sigmoid -> sa; out = x*ca*sa written in place into the staged chunk and pushed out with one
//      bulk shared->global store.
// HBM traffic = the algorithmic 1 read + 1 write of the feature map.  If a chunk does not fit in shared memory the
// same kernel runs "non-resident": phases A/B/C re-read the chunk from global memory (L2 at these sizes).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace b200 {
namespace {

constexpr int kThreads = 256;
constexpr int kRedBytes = 12288 + 1024;

struct CbamParams {
  const void* x;
  const void* g;  // bwd only
  void* out;      // fwd: out; bwd: gx
  const float* w1;
  const float* w2;
  const float* wsa;
  float* ca;  // fwd: out (nullable); bwd: in
  float* sa;
  float* part;  // bwd: per-image weight-grad partials
  int B, C, H, W, r, ksa, mode, pchunk;
};

// VW contiguous channels <-> floats.  VW * sizeof(T) is 16 bytes on the fast path ("vectorised NHWC loads": one
// LDS.128 / LDG.128 feeds 8 bf16 channels) and one 32-bit word (2 x 16-bit / 1 x f32) when C is not a multiple of that.
template <typename T, int VW> struct alignas(sizeof(T) * VW) VPack { T e[VW]; };
template <typename T, int VW> struct Vec {
  __device__ static __forceinline__ void load(const T* p, float (&v)[VW]) {
    const VPack<T, VW> k = *reinterpret_cast<const VPack<T, VW>*>(p);
#pragma unroll
    for (int i = 0; i < VW; ++i) v[i] = DT<T>::to_f(k.e[i]);
  }
  __device__ static __forceinline__ void store(T* p, const float (&v)[VW]) {
    VPack<T, VW> k;
#pragma unroll
    for (int i = 0; i < VW; ++i) k.e[i] = DT<T>::from_f(v[i]);
    *reinterpret_cast<VPack<T, VW>*>(p) = k;
  }
};
template <typename T> struct Words { static constexpr int EPL = 4 / (int)sizeof(T), VE = 16 / (int)sizeof(T); };

__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

// Shared-memory carve-up shared by host (sizing) and device.
struct SmemLayout {
  size_t xs, gs, psum, pmax, pidx, hpart, hid, caslice, ca, smap, shalo, sa, wsa, red, bar, total;
  int halo;
  __host__ __device__ SmemLayout(int C, int r, int W, int ksa, int pchunk, size_t esize, bool resident, bool bwd) {
    const int pad = ksa / 2;
    halo = pad * W + pad;
    size_t o = 0;
    xs = o; o += resident ? align16((size_t)pchunk * C * esize) : 0;
    gs = o; o += (resident && bwd) ? align16((size_t)pchunk * C * esize) : 0;
    psum = o; o += align16((size_t)C * 4);
    pmax = o; o += align16((size_t)C * 4);
    pidx = o; o += align16((size_t)C * 4);
    hpart = o; o += align16((size_t)2 * r * 4);       // this rank's partial hidden pre-activations [2][r]
    hid = o; o += align16((size_t)4 * r * 4);         // summed hidden [2][r] (+ bwd: grads [2][r])
    caslice = o; o += align16((size_t)C * 4 * (bwd ? 3 : 1));  // slice-owner results (peers gather from here)
    ca = o; o += align16((size_t)C * 4 * (bwd ? 4 : 1));       // gathered full-C vectors
    smap = o; o += align16((size_t)2 * pchunk * 4 * (bwd ? 2 : 1));  // own map chunk [2][pchunk] (+ bwd: gz / argmax)
    shalo = o; o += align16((size_t)2 * (pchunk + 2 * halo) * 4);     // gathered map with halo
    sa = o; o += align16((size_t)pchunk * 4 * (bwd ? 3 : 1));
    wsa = o; o += align16((size_t)2 * ksa * ksa * 4 * (bwd ? 2 : 1));
    red = o; o += kRedBytes;  // cross-group reduction scratch
    bar = o; o += 16;
    total = o;
  }
};

// Cooperative chunk load: bulk TMA when 16-byte aligned, plain loads otherwise.
template <typename T>
__device__ __forceinline__ void stage_chunk(T* dst, const T* src, size_t bytes, uint64_t* bar, uint32_t parity) {
  const bool aligned = ((bytes & 15) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  if (aligned) {
    if (threadIdx.x == 0 && bytes) {
      mbar_expect_tx(bar, (uint32_t)bytes);
      size_t off = 0;
      while (off < bytes) {
        const uint32_t n = (uint32_t)((bytes - off) > 32768 ? 32768 : (bytes - off));
        bulk_g2s(reinterpret_cast<char*>(dst) + off, reinterpret_cast<const char*>(src) + off, n, bar);
        off += n;
      }
    }
    if (bytes) mbar_wait(bar, parity);
  } else {
    const size_t n = bytes / sizeof(T);
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
  }
}

// ---- phase A: per-channel sum / max (+ first argmax pixel) over this CTA's pixels -----------------------------
template <typename T, int VW>
__device__ __forceinline__ void channel_partials(const T* xc, int np, int p0, int C, float* psum, float* pmax, int* pidx,
                                                 float* red) {
  constexpr int EPL = VW;
  const int nw = C / EPL;                       // chunks per pixel
  int groups = nw >= kThreads ? 1 : kThreads / nw;  // pixel groups working on the same chunk
  groups = min(groups, max(1, kRedBytes / (C * 12)));  // [groups][C] x {sum,max,idx} must fit the scratch
  const int tw = threadIdx.x % (nw < kThreads ? nw : kThreads);
  const int pg = threadIdx.x / (nw < kThreads ? nw : kThreads);
  // red layout: [groups][C] x {sum,max,idx}; groups*C*12 bytes must fit -> fall back to groups=1 otherwise
  for (int w = tw; w < nw; w += kThreads) {
    float s[EPL], m[EPL];
    int mi[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) { s[e] = 0.f; m[e] = -INFINITY; mi[e] = p0; }
    if (pg < groups) {
      for (int p = pg; p < np; p += groups) {
        float v[EPL];
        Vec<T, VW>::load(xc + (size_t)p * C + w * EPL, v);
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          s[e] += v[e];
          const bool take = (v[e] > m[e]) || (v[e] != v[e]);
          if (take) { m[e] = v[e]; mi[e] = p0 + p; }
        }
      }
    }
    if (groups == 1) {
      if (pg == 0) {
#pragma unroll
        for (int e = 0; e < EPL; ++e) { psum[w * EPL + e] = s[e]; pmax[w * EPL + e] = m[e]; pidx[w * EPL + e] = mi[e]; }
      }
    } else if (pg < groups) {
      float* rs = red + (size_t)pg * C * 3;
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        rs[w * EPL + e] = s[e];
        rs[C + w * EPL + e] = m[e];
        reinterpret_cast<int*>(rs)[2 * C + w * EPL + e] = mi[e];
      }
    }
  }
  if (groups > 1) {
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kThreads) {
      float s = 0.f, m = -INFINITY;
      int mi = p0;
      bool nan = false;
      for (int gi = 0; gi < groups; ++gi) {  // groups interleave pixels: combine by (value, then smaller index)
        const float* rs = red + (size_t)gi * C * 3;
        s += rs[c];
        const float v = rs[C + c];
        const int vi = reinterpret_cast<const int*>(rs)[2 * C + c];
        if (v != v) { if (!nan || vi > mi) { m = v; mi = vi; nan = true; } }
        else if (!nan && (v > m || (v == m && vi < mi && v != -INFINITY))) { m = v; mi = vi; }
      }
      psum[c] = s; pmax[c] = m; pidx[c] = mi;
    }
  }
}

template <typename T, bool RES, int VW>
__global__ void __launch_bounds__(kThreads) cbam_fwd_kernel(CbamParams P) {
  constexpr int EPL = VW;
  cg::cluster_group cluster = cg::this_cluster();
  const int CS = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / CS;
  const int C = P.C, HW = P.H * P.W, W = P.W, H = P.H, r = P.r, ks = P.ksa, pad = ks / 2;
  const int p0 = min(rank * P.pchunk, HW), p1 = min(p0 + P.pchunk, HW), np = p1 - p0;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const SmemLayout L(C, r, W, ks, P.pchunk, sizeof(T), RES, false);
  T* xs = reinterpret_cast<T*>(smem_raw + L.xs);
  float* psum = reinterpret_cast<float*>(smem_raw + L.psum);
  float* pmax = reinterpret_cast<float*>(smem_raw + L.pmax);
  int* pidx = reinterpret_cast<int*>(smem_raw + L.pidx);
  float* hpart = reinterpret_cast<float*>(smem_raw + L.hpart);
  float* hid = reinterpret_cast<float*>(smem_raw + L.hid);
  float* caslice = reinterpret_cast<float*>(smem_raw + L.caslice);
  float* ca = reinterpret_cast<float*>(smem_raw + L.ca);
  float* smap = reinterpret_cast<float*>(smem_raw + L.smap);
  float* shalo = reinterpret_cast<float*>(smem_raw + L.shalo);
  float* sas = reinterpret_cast<float*>(smem_raw + L.sa);
  float* wsas = reinterpret_cast<float*>(smem_raw + L.wsa);
  float* red = reinterpret_cast<float*>(smem_raw + L.red);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.bar);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  const T* xg = reinterpret_cast<const T*>(P.x) + ((size_t)b * HW + p0) * C;
  const T* xc = xg;
  if (RES) {
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    stage_chunk<T>(xs, xg, (size_t)np * C * sizeof(T), bar, 0);
    xc = xs;
  }
  for (int i = tid; i < 2 * ks * ks; i += kThreads) wsas[i] = P.wsa ? P.wsa[i] : 0.f;

  // channel slice owned by this rank for the cross-rank reductions
  const int cper = (C + CS - 1) / CS;
  const int cs0 = min(rank * cper, C), cs1 = min(cs0 + cper, C);

  if (P.mode != B200_CBAM_SA) {
    channel_partials<T, VW>(xc, np, p0, C, psum, pmax, pidx, red);
    cluster.sync();  // (1) partials visible cluster-wide
    // reduce slice S_rank over ranks -> pooled avg/max, then partial hidden = W1[:, S] . pooled[S]
    float* pav = caslice;  // reuse: pooled avg for the slice (overwritten by ca slice later)
    float* pmx = red;      // pooled max for the slice
    for (int c = cs0 + tid; c < cs1; c += kThreads) {
      float s = 0.f, m = -INFINITY;
      bool nan = false;
      for (int k2 = 0; k2 < CS; ++k2) {
        const float* rs = cluster.map_shared_rank(psum, k2);
        const float* rm = cluster.map_shared_rank(pmax, k2);
        s += rs[c];
        const float v = rm[c];
        if (v != v) nan = true;
        else if (v > m) m = v;
      }
      pav[c - cs0] = s / (float)HW;
      pmx[c - cs0] = nan ? __int_as_float(0x7fc00000) : m;
    }
    __syncthreads();
    for (int j = warp; j < 2 * r; j += kThreads / 32) {  // j < r: avg branch, j >= r: max branch
      const int jj = j < r ? j : j - r;
      const float* src = j < r ? pav : pmx;
      float acc = 0.f;
      for (int c = cs0 + lane; c < cs1; c += 32) acc += P.w1[(size_t)jj * C + c] * src[c - cs0];
      acc = warp_sum(acc);
      if (lane == 0) hpart[j] = acc;
    }
    cluster.sync();  // (2) partial hiddens visible
    for (int j = tid; j < 2 * r; j += kThreads) {
      float acc = 0.f;
      for (int k2 = 0; k2 < CS; ++k2) acc += cluster.map_shared_rank(hpart, k2)[j];
      hid[j] = fmaxf(acc, 0.f);  // ReLU (cbam.py:25)
    }
    __syncthreads();
    for (int c = cs0 + tid; c < cs1; c += kThreads) {
      float z = 0.f;
      for (int j = 0; j < r; ++j) z += P.w2[(size_t)c * r + j] * (hid[j] + hid[r + j]);
      const float a = sigmoidf_(z);
      ca[c] = a;  // own slice goes straight to its final place; peers read it from `ca` of this rank
      if (P.ca) P.ca[(size_t)b * C + c] = a;
    }
    cluster.sync();  // (3) every rank's ca slice visible
    for (int c = tid; c < C; c += kThreads) {
      const int owner = min(c / cper, CS - 1);
      if (owner != rank) ca[c] = cluster.map_shared_rank(ca, owner)[c];
    }
    __syncthreads();
    if (P.mode == B200_CBAM_CA) { cluster.sync(); return; }
  } else {
    for (int c = tid; c < C; c += kThreads) ca[c] = 1.f;
    __syncthreads();
  }

  // ---- phase B: per-pixel channel mean / max of x*ca -----------------------------------------------------
  const int nw = C / EPL;
  for (int p = warp; p < np; p += kThreads / 32) {
    float s = 0.f, m = -INFINITY;
    for (int w = lane; w < nw; w += 32) {
      float v[EPL];
      Vec<T, VW>::load(xc + (size_t)p * C + w * EPL, v);
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        const float t = v[e] * ca[w * EPL + e];
        s += t;
        m = (t > m || t != t) ? t : m;
      }
    }
    s = warp_sum(s);
    // NaN-propagating max across lanes
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float t = __shfl_xor_sync(0xffffffffu, m, o);
      m = (t > m || t != t) ? t : m;
    }
    if (lane == 0) { smap[p] = s / (float)C; smap[P.pchunk + p] = m; }
  }
  cluster.sync();  // (4) map chunks visible
  // gather map with halo: flat positions [p0 - halo, p1 + halo)
  const int halo = L.halo, span = np + 2 * halo;
  for (int i = tid; i < 2 * span; i += kThreads) {
    const int ch = i / span, q = p0 - halo + (i - ch * span);
    float v = 0.f;
    if (q >= 0 && q < HW) {
      const int owner = min(q / P.pchunk, CS - 1);
      v = cluster.map_shared_rank(smap, owner)[ch * P.pchunk + (q - owner * P.pchunk)];
    }
    shalo[ch * (P.pchunk + 2 * halo) + (i - ch * span)] = v;
  }
  __syncthreads();
  // ---- phase C: conv + sigmoid (warp per pixel, lanes own up to 4 of the 2*ks*ks taps), gate ---------------------
  {
    constexpr int TPL = 4;  // taps per lane: 2*7*7 = 98 <= 128
    int tdu[TPL], tdv[TPL], tch[TPL];
    float tw[TPL];
#pragma unroll
    for (int j = 0; j < TPL; ++j) {
      const int t = lane + 32 * j;
      const bool ok = t < 2 * ks * ks;
      tch[j] = ok ? t / (ks * ks) : 0;
      tdu[j] = ok ? (t / ks) % ks - pad : 0;
      tdv[j] = ok ? t % ks - pad : 0;
      tw[j] = ok ? wsas[t] : 0.f;
    }
    const int hp_ = P.pchunk + 2 * halo;
    for (int p = warp; p < np; p += kThreads / 32) {
      const int q = p0 + p, y = q / W, x = q - y * W;
      float z = 0.f;
#pragma unroll
      for (int j = 0; j < TPL; ++j) {
        const int yy = y + tdu[j], xx = x + tdv[j];
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) z += tw[j] * shalo[tch[j] * hp_ + (yy * W + xx) - (p0 - halo)];
      }
      z = warp_sum(z);
      if (lane == 0) {
        const float a = sigmoidf_(z);
        sas[p] = a;
        if (P.sa) P.sa[(size_t)b * HW + q] = a;
      }
    }
  }
  __syncthreads();
  if (P.mode == B200_CBAM_FULL) {
    T* og = reinterpret_cast<T*>(P.out) + ((size_t)b * HW + p0) * C;
    T* dst = RES ? xs : og;
    for (int p = warp; p < np; p += kThreads / 32) {
      const float sp = sas[p];
      for (int w = lane; w < nw; w += 32) {
        float v[EPL];
        Vec<T, VW>::load(xc + (size_t)p * C + w * EPL, v);
#pragma unroll
        for (int e = 0; e < EPL; ++e) v[e] = v[e] * ca[w * EPL + e] * sp;
        Vec<T, VW>::store(dst + (size_t)p * C + w * EPL, v);
      }
    }
    if (RES) {
      const size_t bytes = (size_t)np * C * sizeof(T);
      if (((bytes & 15) == 0) && ((reinterpret_cast<uintptr_t>(og) & 15) == 0)) {
        fence_proxy_async();
        __syncthreads();
        if (tid == 0 && bytes) {
          size_t off = 0;
          while (off < bytes) {
            const uint32_t n = (uint32_t)((bytes - off) > 32768 ? 32768 : (bytes - off));
            bulk_s2g(reinterpret_cast<char*>(og) + off, reinterpret_cast<char*>(xs) + off, n);
            off += n;
          }
          bulk_commit();
          bulk_wait_read_all();
        }
      } else {
        __syncthreads();
        for (size_t i = tid; i < (size_t)np * C; i += kThreads) og[i] = xs[i];
      }
    }
  }
  cluster.sync();  // keep smem alive until every peer finished its DSMEM reads
}

// =====================================================================================================
// backward.  Same cluster decomposition; x and g chunks both staged once.
// =====================================================================================================
template <typename T, bool RES, int VW>
__global__ void __launch_bounds__(kThreads) cbam_bwd_kernel(CbamParams P) {
  constexpr int EPL = VW;
  cg::cluster_group cluster = cg::this_cluster();
  const int CS = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / CS;
  const int C = P.C, HW = P.H * P.W, W = P.W, H = P.H, r = P.r, ks = P.ksa, pad = ks / 2;
  const int p0 = min(rank * P.pchunk, HW), p1 = min(p0 + P.pchunk, HW), np = p1 - p0;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const SmemLayout L(C, r, W, ks, P.pchunk, sizeof(T), RES, true);
  T* xs = reinterpret_cast<T*>(smem_raw + L.xs);
  T* gs = reinterpret_cast<T*>(smem_raw + L.gs);
  float* psum = reinterpret_cast<float*>(smem_raw + L.psum);
  float* pmax = reinterpret_cast<float*>(smem_raw + L.pmax);
  int* pidx = reinterpret_cast<int*>(smem_raw + L.pidx);
  float* hpart = reinterpret_cast<float*>(smem_raw + L.hpart);
  float* hid = reinterpret_cast<float*>(smem_raw + L.hid);       // [0,2r): pre-activations, [2r,4r): their grads
  float* slc = reinterpret_cast<float*>(smem_raw + L.caslice);   // [3][C]: slice-owner scratch
  float* ca = reinterpret_cast<float*>(smem_raw + L.ca);         // [0]=ca, [1]=g_pavg/HW, [2]=g_pmax, [3]=argmax_hw
  float* smap = reinterpret_cast<float*>(smem_raw + L.smap);     // [0..1]: s map chunk, [2]: g_z chunk, [3]: argmax_c
  float* shalo = reinterpret_cast<float*>(smem_raw + L.shalo);
  float* sas = reinterpret_cast<float*>(smem_raw + L.sa);        // [0]=sa, [1]=g_s0, [2]=g_s1
  float* wsas = reinterpret_cast<float*>(smem_raw + L.wsa);
  float* red = reinterpret_cast<float*>(smem_raw + L.red);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.bar);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nw = C / EPL;
  const int pc = P.pchunk;
  const int nwt = 2 * ks * ks;

  const T* xg = reinterpret_cast<const T*>(P.x) + ((size_t)b * HW + p0) * C;
  const T* gg = P.mode == B200_CBAM_FULL ? reinterpret_cast<const T*>(P.g) + ((size_t)b * HW + p0) * C : nullptr;
  const T* xc = xg;
  const T* gc = gg;
  if (RES) {
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    const size_t bytes = (size_t)np * C * sizeof(T);
    const bool aligned = ((bytes & 15) == 0) && ((reinterpret_cast<uintptr_t>(xg) & 15) == 0) &&
                         (!gg || (reinterpret_cast<uintptr_t>(gg) & 15) == 0);
    if (aligned) {
      if (tid == 0 && bytes) {
        mbar_expect_tx(bar, (uint32_t)(bytes * (gg ? 2 : 1)));
        for (size_t off = 0; off < bytes; off += 32768) {
          const uint32_t n = (uint32_t)((bytes - off) > 32768 ? 32768 : (bytes - off));
          bulk_g2s(reinterpret_cast<char*>(xs) + off, reinterpret_cast<const char*>(xg) + off, n, bar);
          if (gg) bulk_g2s(reinterpret_cast<char*>(gs) + off, reinterpret_cast<const char*>(gg) + off, n, bar);
        }
      }
      if (bytes) mbar_wait(bar, 0);
    } else {
      for (size_t i = tid; i < (size_t)np * C; i += kThreads) { xs[i] = xg[i]; if (gg) gs[i] = gg[i]; }
      __syncthreads();
    }
    xc = xs;
    gc = gg ? gs : nullptr;
  }
  for (int i = tid; i < nwt; i += kThreads) wsas[i] = P.wsa ? P.wsa[i] : 0.f;
  const int cper = (C + CS - 1) / CS;
  const int cs0 = min(rank * cper, C), cs1 = min(cs0 + cper, C);
  const bool use_ca = P.mode != B200_CBAM_SA;
  const bool use_sa = P.mode != B200_CBAM_CA;
  for (int c = tid; c < C; c += kThreads) ca[c] = use_ca ? P.ca[(size_t)b * C + c] : 1.f;
  for (int p = tid; p < np; p += kThreads) sas[p] = use_sa ? P.sa[(size_t)b * HW + p0 + p] : 1.f;
  __syncthreads();

  float* gw_part = P.part + (size_t)b * (2 * (size_t)r * C + nwt);  // per-image partials [r*C | C*r | nwt]

  if (use_sa) {
    // (1) recompute s map + channel argmax; g_sa[p] = sum_c g * x * ca   (SA mode: g_sa is the input itself)
    for (int p = warp; p < np; p += kThreads / 32) {
      float s = 0.f, m = -INFINITY, gsa = 0.f;
      int mi = 0;
      for (int w = lane; w < nw; w += 32) {
        float v[EPL], gv[EPL];
        Vec<T, VW>::load(xc + (size_t)p * C + w * EPL, v);
        if (gc) Vec<T, VW>::load(gc + (size_t)p * C + w * EPL, gv);
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          const float t = v[e] * ca[w * EPL + e];
          s += t;
          if (t > m || t != t) { m = t; mi = w * EPL + e; }
          if (gc) gsa += gv[e] * t;
        }
      }
      s = warp_sum(s);
      gsa = warp_sum(gsa);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {  // (value, first index) with NaN propagation: later NaN wins like ATen
        const float tm = __shfl_xor_sync(0xffffffffu, m, o);
        const int ti = __shfl_xor_sync(0xffffffffu, mi, o);
        const bool tnan = tm != tm, mnan = m != m;
        bool take;
        if (tnan || mnan) take = tnan && (!mnan || ti > mi);
        else take = (tm > m) || (tm == m && ti < mi);
        if (take) { m = tm; mi = ti; }
      }
      if (lane == 0) {
        smap[p] = s / (float)C;
        smap[pc + p] = m;
        smap[3 * pc + p] = __int_as_float(mi);
        const float a = sas[p];
        const float gsa_in = P.mode == B200_CBAM_SA ? reinterpret_cast<const float*>(P.g)[(size_t)b * HW + p0 + p] : gsa;
        smap[2 * pc + p] = gsa_in * a * (1.f - a);  // g_z
      }
    }
    cluster.sync();  // (1) s map + g_z chunks visible
    // gather g_z with halo (for the transposed conv) and s with halo (for the weight gradient)
    const int halo = L.halo, span = np + 2 * halo, hp = pc + 2 * halo;
    // --- g_s = conv^T(g_z): needs g_z halo
    for (int i = tid; i < span; i += kThreads) {
      const int q = p0 - halo + i;
      float v = 0.f;
      if (q >= 0 && q < HW) {
        const int owner = min(q / pc, CS - 1);
        v = cluster.map_shared_rank(smap, owner)[2 * pc + (q - owner * pc)];
      }
      shalo[i] = v;
    }
    __syncthreads();
    {
      constexpr int TPL = 2;  // taps per lane: ks*ks = 49 <= 64
      int tdu[TPL], tdv[TPL];
      float tw0[TPL], tw1[TPL];
#pragma unroll
      for (int j = 0; j < TPL; ++j) {
        const int t = lane + 32 * j;
        const bool ok = t < ks * ks;
        tdu[j] = ok ? t / ks - pad : 0;
        tdv[j] = ok ? t % ks - pad : 0;
        tw0[j] = ok ? wsas[t] : 0.f;
        tw1[j] = ok ? wsas[ks * ks + t] : 0.f;
      }
      for (int p = warp; p < np; p += kThreads / 32) {
        const int q = p0 + p, y = q / W, x = q - y * W;
        float g0 = 0.f, g1 = 0.f;
#pragma unroll
        for (int j = 0; j < TPL; ++j) {
          const int yy = y - tdu[j], xx = x - tdv[j];  // output position that used this tap on this input
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            const float gz = shalo[(yy * W + xx) - (p0 - halo)];
            g0 += tw0[j] * gz;
            g1 += tw1[j] * gz;
          }
        }
        g0 = warp_sum(g0);
        g1 = warp_sum(g1);
        if (lane == 0) {
          sas[pc + p] = g0 / (float)C;  // broadcast share of the channel mean
          sas[2 * pc + p] = g1;         // routed to the argmax channel
        }
      }
    }
    __syncthreads();
    // --- g_Wsa[j,u,v] = sum_p g_z[p] * s[j, p + (u-pad, v-pad)]: gather s halo, per-CTA partial
    for (int i = tid; i < 2 * span; i += kThreads) {
      const int ch = i / span, q = p0 - halo + (i - ch * span);
      float v = 0.f;
      if (q >= 0 && q < HW) {
        const int owner = min(q / pc, CS - 1);
        v = cluster.map_shared_rank(smap, owner)[ch * pc + (q - owner * pc)];
      }
      shalo[ch * hp + (i - ch * span)] = v;
    }
    __syncthreads();
    if (np <= 128) {  // lanes own up to 4 pixels each; coordinates and g_z decoded once, reused for every tap
      int py[4], px[4];
      float pgz[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = lane + 32 * j;
        const int q = p0 + (p < np ? p : 0);
        py[j] = q / W; px[j] = q - py[j] * W;
        pgz[j] = p < np ? smap[2 * pc + p] : 0.f;
      }
      for (int t = warp; t < nwt; t += kThreads / 32) {
        const int ch = t / (ks * ks), u = (t / ks) % ks - pad, v = t % ks - pad;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int yy = py[j] + u, xx = px[j] + v;
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) acc += pgz[j] * shalo[ch * hp + (yy * W + xx) - (p0 - halo)];
        }
        acc = warp_sum(acc);
        if (lane == 0) wsas[nwt + t] = acc;
      }
    } else {
      for (int t = warp; t < nwt; t += kThreads / 32) {
        const int ch = t / (ks * ks), u = (t / ks) % ks, v = t % ks;
        float acc = 0.f;
        for (int p = lane; p < np; p += 32) {
          const int q = p0 + p, y = q / W, x = q - y * W;
          const int yy = y + u - pad, xx = x + v - pad;
          if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
          acc += smap[2 * pc + p] * shalo[ch * hp + (yy * W + xx) - (p0 - halo)];
        }
        acc = warp_sum(acc);
        if (lane == 0) wsas[nwt + t] = acc;
      }
    }
    cluster.sync();  // (2) per-CTA g_Wsa partials visible; rank 0 folds them (fixed order -> deterministic)
    if (rank == 0)
      for (int t = tid; t < nwt; t += kThreads) {
        float acc = 0.f;
        for (int k2 = 0; k2 < CS; ++k2) acc += cluster.map_shared_rank(wsas, k2)[nwt + t];
        gw_part[2 * (size_t)r * C + t] = acc;
      }
  }

  if (P.mode == B200_CBAM_SA) {
    // gx = (g_s0 + [c == argmax] g_s1), no channel attention involved
    T* og = reinterpret_cast<T*>(P.out) + ((size_t)b * HW + p0) * C;
    for (int i = tid; i < np * nw; i += kThreads) {
      const int p = i / nw, w = i - p * nw;
      const int am = __float_as_int(smap[3 * pc + p]);
      float o[EPL];
#pragma unroll
      for (int e = 0; e < EPL; ++e) o[e] = sas[pc + p] + ((w * EPL + e) == am ? sas[2 * pc + p] : 0.f);
      Vec<T, VW>::store(og + (size_t)p * C + w * EPL, o);
    }
    cluster.sync();
    return;
  }

  // (3)+(4): g_x1 = g*sa + g_s0 + [c==argmax_c] g_s1 ;  g_ca[c] = sum_p g_x1 * x  (per-CTA partial -> psum)
  // also the pooled statistics again (avg / max / argmax_hw) for the MLP backward.
  channel_partials<T, VW>(xc, np, p0, C, psum, pmax, pidx, red);
  __syncthreads();
  float* gca_part = slc;  // [C] this CTA's partial of g_ca
  if (P.mode == B200_CBAM_FULL) {
    const int groups = min(nw >= kThreads ? 1 : kThreads / nw, max(1, kRedBytes / (C * 12)));
    const int tw = tid % (nw < kThreads ? nw : kThreads), pg = tid / (nw < kThreads ? nw : kThreads);
    for (int i = tid; i < C; i += kThreads) gca_part[i] = 0.f;
    __syncthreads();
    for (int w = tw; w < nw; w += kThreads) {
      float acc[EPL];
#pragma unroll
      for (int e = 0; e < EPL; ++e) acc[e] = 0.f;
      if (pg < groups)
        for (int p = pg; p < np; p += groups) {
          float v[EPL], gv[EPL];
          Vec<T, VW>::load(xc + (size_t)p * C + w * EPL, v);
          Vec<T, VW>::load(gc + (size_t)p * C + w * EPL, gv);
          const int am = __float_as_int(smap[3 * pc + p]);
#pragma unroll
          for (int e = 0; e < EPL; ++e) {
            const float gx1 = gv[e] * sas[p] + sas[pc + p] + ((w * EPL + e) == am ? sas[2 * pc + p] : 0.f);
            acc[e] += gx1 * v[e];
          }
        }
      if (pg < groups) {
#pragma unroll
        for (int e = 0; e < EPL; ++e) red[(size_t)pg * C + w * EPL + e] = acc[e];
      }
    }
    __syncthreads();
    {
      const int groups2 = min(nw >= kThreads ? 1 : kThreads / nw, max(1, kRedBytes / (C * 12)));
      for (int c = tid; c < C; c += kThreads) {
        float s = 0.f;
        for (int gi = 0; gi < groups2; ++gi) s += red[(size_t)gi * C + c];
        gca_part[c] = s;
      }
    }
  }
  cluster.sync();  // (3) g_ca partials + pooled partials visible
  // slice owner: reduce g_ca and pooled stats over ranks; hidden partials for the forward recompute
  float* pav = slc + C;      // [cper] pooled avg (slice)
  float* pmx = slc + 2 * C;  // [cper] pooled max (slice)
  float* gca_s = red;        // [cper] g_a = g_ca * ca * (1-ca) for the slice
  int* amx = reinterpret_cast<int*>(red) + cper;  // [cper] argmax_hw for the slice
  for (int c = cs0 + tid; c < cs1; c += kThreads) {
    float s = 0.f, m = -INFINITY, gsum = 0.f;
    int mi = 0;
    bool nan = false;
    for (int k2 = 0; k2 < CS; ++k2) {
      s += cluster.map_shared_rank(psum, k2)[c];
      const float v = cluster.map_shared_rank(pmax, k2)[c];
      const int vi = cluster.map_shared_rank(pidx, k2)[c];
      if (v != v) { m = v; mi = vi; nan = true; }          // ranks ascend in pixel order: last NaN wins
      else if (!nan && v > m) { m = v; mi = vi; }          // strict >: first occurrence
      else if (!nan && k2 == 0) { mi = vi; }
      if (P.mode == B200_CBAM_FULL) gsum += cluster.map_shared_rank(gca_part, k2)[c];
    }
    if (P.mode == B200_CBAM_CA) gsum = reinterpret_cast<const float*>(P.g)[(size_t)b * C + c];
    pav[c - cs0] = s / (float)HW;
    pmx[c - cs0] = m;
    amx[c - cs0] = mi;
    const float a = ca[c];
    gca_s[c - cs0] = gsum * a * (1.f - a);
  }
  __syncthreads();
  // partial hidden pre-activations (forward recompute) and partial W2^T g_a, both over the slice
  for (int j = warp; j < 3 * r; j += kThreads / 32) {
    float acc = 0.f;
    if (j < 2 * r) {
      const int jj = j < r ? j : j - r;
      const float* src = j < r ? pav : pmx;
      for (int c = cs0 + lane; c < cs1; c += 32) acc += P.w1[(size_t)jj * C + c] * src[c - cs0];
    } else {
      const int jj = j - 2 * r;
      for (int c = cs0 + lane; c < cs1; c += 32) acc += P.w2[(size_t)c * r + jj] * gca_s[c - cs0];
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      if (j < 2 * r) hpart[j] = acc;
      else hid[2 * r + (j - 2 * r)] = acc;
    }
  }
  cluster.sync();  // (4) hidden partials visible
  // h[2][r] pre-activations and u[r] = W2^T g_a, summed over ranks
  float* hfull = hid;          // [2r] pre-activations (hid[2r,3r) holds this rank's partial of u, read by peers)
  float* ufull = hid + 3 * r;  // [r]
  for (int j = tid; j < 3 * r; j += kThreads) {
    float acc = 0.f;
    for (int k2 = 0; k2 < CS; ++k2)
      acc += (j < 2 * r) ? cluster.map_shared_rank(hpart, k2)[j] : cluster.map_shared_rank(hid, k2)[2 * r + (j - 2 * r)];
    if (j < 2 * r) hfull[j] = acc; else ufull[j - 2 * r] = acc;
  }
  __syncthreads();
  // slice owner: weight-gradient partials for its channel slice and g_p_t = W1^T g_h_t
  for (int c = cs0 + tid; c < cs1; c += kThreads) {
    const float ga = gca_s[c - cs0];
    float gpa = 0.f, gpm = 0.f;
    for (int j = 0; j < r; ++j) {
      const float ha = hfull[j], hm = hfull[r + j], u = ufull[j];
      const float gha = ha > 0.f ? u : 0.f, ghm = hm > 0.f ? u : 0.f;
      gw_part[(size_t)r * C + (size_t)c * r + j] = ga * (fmaxf(ha, 0.f) + fmaxf(hm, 0.f));           // g_W2[c,j]
      gw_part[(size_t)j * C + c] = gha * pav[c - cs0] + ghm * pmx[c - cs0];                          // g_W1[j,c]
      const float w = P.w1[(size_t)j * C + c];
      gpa += w * gha;
      gpm += w * ghm;
    }
    ca[C + c] = gpa / (float)HW;
    ca[2 * C + c] = gpm;
    ca[3 * C + c] = __int_as_float(amx[c - cs0]);
  }
  cluster.sync();  // (5) slice results visible
  for (int c = tid; c < C; c += kThreads) {
    const int owner = min(c / cper, CS - 1);
    if (owner != rank) {
      const float* rc = cluster.map_shared_rank(ca, owner);
      ca[C + c] = rc[C + c]; ca[2 * C + c] = rc[2 * C + c]; ca[3 * C + c] = rc[3 * C + c];
    }
  }
  __syncthreads();
  // (6) g_x = g_x1 * ca + g_pavg/HW + [p == argmax_hw] g_pmax
  {
    T* og = reinterpret_cast<T*>(P.out) + ((size_t)b * HW + p0) * C;
    for (int p = warp; p < np; p += kThreads / 32) {
      const int am = use_sa && P.mode == B200_CBAM_FULL ? __float_as_int(smap[3 * pc + p]) : -1;
      const float sp = sas[p], g0 = sas[pc + p], g1 = sas[2 * pc + p];
      for (int w = lane; w < nw; w += 32) {
        float gv[EPL], o[EPL];
        if (gc) Vec<T, VW>::load(gc + (size_t)p * C + w * EPL, gv);
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          const int c = w * EPL + e;
          float gx1 = 0.f;
          if (P.mode == B200_CBAM_FULL) gx1 = gv[e] * sp + g0 + (c == am ? g1 : 0.f);
          o[e] = gx1 * ca[c] + ca[C + c] + ((p0 + p) == __float_as_int(ca[3 * C + c]) ? ca[2 * C + c] : 0.f);
        }
        Vec<T, VW>::store(og + (size_t)p * C + w * EPL, o);
      }
    }
  }
  cluster.sync();
}

// fold per-image weight-gradient partials [B][n] -> [n] in a fixed order (deterministic)
__global__ void fold_partials_kernel(const float* __restrict__ part, float* gw1, float* gw2, float* gwsa, int B,
                                     int n1, int n2, int n3) {
  const int n = n1 + n2 + n3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += part[(size_t)b * n + i];
    if (i < n1) { if (gw1) gw1[i] = acc; }
    else if (i < n1 + n2) { if (gw2) gw2[i - n1] = acc; }
    else if (gwsa) gwsa[i - n1 - n2] = acc;
  }
}

template <typename K>
int launch_cluster(K kern, int grid, int cs, size_t smem, cudaStream_t st, CbamParams P) {
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (cs > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, P);
  if (e != cudaSuccess) {
    set_error("cbam: cluster launch failed (grid=%d cs=%d smem=%zu): %s", grid, cs, smem, cudaGetErrorString(e));
    cudaGetLastError();
    return B200_ERR_LAUNCH;
  }
  return check_launch("cbam");
}

int check_common(const void* x, int B, int C, int H, int W, int r, int ksa, int dtype, int mode) {
  B200_REQUIRE(x, B200_ERR_SHAPE, "cbam: null input");
  B200_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, B200_ERR_SHAPE, "cbam: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
  B200_REQUIRE(mode >= 0 && mode <= 2, B200_ERR_SHAPE, "cbam: bad mode %d", mode);
  if (mode != B200_CBAM_SA) B200_REQUIRE(r > 0, B200_ERR_SHAPE, "cbam: hidden width r must be > 0");
  if (mode != B200_CBAM_CA) B200_REQUIRE(ksa == 3 || ksa == 7, B200_ERR_SHAPE, "cbam: kernel size must be 3 or 7 (cbam.py:43), got %d", ksa);
  if (dtype != B200_F32) B200_REQUIRE(C % 2 == 0, B200_ERR_ALIGN, "cbam: C must be even for 16-bit dtypes");
  B200_REQUIRE(((uintptr_t)x & 3) == 0, B200_ERR_ALIGN, "cbam: input must be 4-byte aligned");
  return B200_OK;
}

// choose cluster size + residency.  Returns pchunk; sets cs/resident/smem.
void plan(int C, int r, int H, int W, int ksa, size_t esize, bool bwd, int* cs, bool* resident, size_t* smem, int* pchunk) {
  const int HW = H * W;
  const size_t lim = (size_t)max_smem_optin();
  for (int c : {8, 16}) {
    int pc = (HW + c - 1) / c;
    SmemLayout L(C, r, W, ksa, pc, esize, true, bwd);
    // 16-CTA clusters are non-portable: only take them when 2 CTAs still fit per SM
    if (L.total <= (c == 8 ? lim : (size_t)100 * 1024)) { *cs = c; *resident = true; *smem = L.total; *pchunk = pc; return; }
  }
  int pc = (HW + 7) / 8;
  SmemLayout L(C, r, W, ksa, pc, esize, false, bwd);
  *cs = 8; *resident = false; *smem = L.total; *pchunk = pc;
}

}  // namespace
}  // namespace b200

extern "C" B200_API int b200_cbam_fwd(const void* x, const float* w1, const float* w2, const float* wsa, void* out,
                             float* ca_out, float* sa_out, int32_t B, int32_t C, int32_t H, int32_t W, int32_t r,
                             int32_t ksa, int32_t dtype, int32_t mode, void* stream) {
  using namespace b200;
  if (mode == B200_CBAM_CA) ksa = 3;
  if (mode == B200_CBAM_SA) r = 1;
  if (int rc = check_common(x, B, C, H, W, r, ksa, dtype, mode)) return rc;
  if (mode != B200_CBAM_SA) B200_REQUIRE(w1 && w2, B200_ERR_SHAPE, "cbam_fwd: null MLP weights");
  if (mode != B200_CBAM_CA) B200_REQUIRE(wsa, B200_ERR_SHAPE, "cbam_fwd: null conv weight");
  if (mode == B200_CBAM_FULL) B200_REQUIRE(out, B200_ERR_SHAPE, "cbam_fwd: null output");
  if (mode == B200_CBAM_CA) B200_REQUIRE(ca_out, B200_ERR_SHAPE, "cbam_fwd: null ca output");
  if (mode == B200_CBAM_SA) B200_REQUIRE(sa_out, B200_ERR_SHAPE, "cbam_fwd: null sa output");
  const size_t esize = dtype == B200_F32 ? 4 : 2;
  int cs; bool res; size_t smem; int pc;
  plan(C, r, H, W, ksa, esize, false, &cs, &res, &smem, &pc);
  B200_REQUIRE(smem <= (size_t)max_smem_optin(), B200_ERR_UNSUPPORTED, "cbam_fwd: shape needs %zu B of shared memory", smem);
  CbamParams P{x, nullptr, out, w1, w2, wsa, ca_out, sa_out, nullptr, B, C, H, W, r, ksa, mode, pc};
  cudaStream_t st = (cudaStream_t)stream;
  return B200_DISPATCH_DTYPE(dtype, [&]() -> int {
    constexpr int VE = Words<T>::VE, EW = Words<T>::EPL;
    if (C % VE == 0 && ((uintptr_t)x & 15) == 0 && (mode != B200_CBAM_FULL || ((uintptr_t)out & 15) == 0))
      return res ? launch_cluster(cbam_fwd_kernel<T, true, VE>, B * cs, cs, smem, st, P)
                 : launch_cluster(cbam_fwd_kernel<T, false, VE>, B * cs, cs, smem, st, P);
    return res ? launch_cluster(cbam_fwd_kernel<T, true, EW>, B * cs, cs, smem, st, P)
               : launch_cluster(cbam_fwd_kernel<T, false, EW>, B * cs, cs, smem, st, P);
  });
}

extern "C" B200_API size_t b200_cbam_bwd_workspace_bytes(int32_t B, int32_t C, int32_t H, int32_t W, int32_t r, int32_t ksa) {
  (void)H; (void)W;
  return (size_t)B * (2 * (size_t)r * C + 2 * (size_t)ksa * ksa) * sizeof(float);
}

extern "C" B200_API int b200_cbam_bwd(const void* g, const void* x, const float* w1, const float* w2, const float* wsa,
                             const float* ca, const float* sa, void* gx, float* gw1, float* gw2, float* gwsa,
                             void* workspace, size_t workspace_bytes, int32_t B, int32_t C, int32_t H, int32_t W,
                             int32_t r, int32_t ksa, int32_t dtype, int32_t mode, void* stream) {
  using namespace b200;
  if (mode == B200_CBAM_CA) ksa = 3;
  if (mode == B200_CBAM_SA) r = 1;
  if (int rc = check_common(x, B, C, H, W, r, ksa, dtype, mode)) return rc;
  B200_REQUIRE(g && gx, B200_ERR_SHAPE, "cbam_bwd: null gradient pointer");
  if (mode != B200_CBAM_SA) B200_REQUIRE(w1 && w2 && ca, B200_ERR_SHAPE, "cbam_bwd: null MLP weights / ca map");
  if (mode != B200_CBAM_CA) B200_REQUIRE(wsa && sa, B200_ERR_SHAPE, "cbam_bwd: null conv weight / sa map");
  const size_t need = b200_cbam_bwd_workspace_bytes(B, C, H, W, r, ksa);
  B200_REQUIRE(workspace && workspace_bytes >= need, B200_ERR_WORKSPACE, "cbam_bwd: workspace %zu < %zu bytes", workspace_bytes, need);
  const size_t esize = dtype == B200_F32 ? 4 : 2;
  int cs; bool res; size_t smem; int pc;
  plan(C, r, H, W, ksa, esize, true, &cs, &res, &smem, &pc);
  B200_REQUIRE(smem <= (size_t)max_smem_optin(), B200_ERR_UNSUPPORTED, "cbam_bwd: shape needs %zu B of shared memory", smem);
  CbamParams P{x, g, gx, w1, w2, wsa, const_cast<float*>(ca), const_cast<float*>(sa), (float*)workspace,
               B, C, H, W, r, ksa, mode, pc};
  cudaStream_t st = (cudaStream_t)stream;
  // partial slots that a mode never writes must read as zero
  cudaMemsetAsync(workspace, 0, need, st);
  int rc = B200_DISPATCH_DTYPE(dtype, [&]() -> int {
    constexpr int VE = Words<T>::VE, EW = Words<T>::EPL;
    if (C % VE == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)gx & 15) == 0 && (mode != B200_CBAM_FULL || ((uintptr_t)g & 15) == 0))
      return res ? launch_cluster(cbam_bwd_kernel<T, true, VE>, B * cs, cs, smem, st, P)
                 : launch_cluster(cbam_bwd_kernel<T, false, VE>, B * cs, cs, smem, st, P);
    return res ? launch_cluster(cbam_bwd_kernel<T, true, EW>, B * cs, cs, smem, st, P)
               : launch_cluster(cbam_bwd_kernel<T, false, EW>, B * cs, cs, smem, st, P);
  });
  if (rc) return rc;
  const int n1 = r * C, n2 = C * r, n3 = 2 * ksa * ksa;
  const int n = n1 + n2 + n3;
  fold_partials_kernel<<<(n + 255) / 256, 256, 0, st>>>((const float*)workspace, mode != B200_CBAM_SA ? gw1 : nullptr,
                                                        mode != B200_CBAM_SA ? gw2 : nullptr,
                                                        mode != B200_CBAM_CA ? gwsa : nullptr, B, n1, n2, n3);
  return check_launch("cbam_bwd_fold");
}
