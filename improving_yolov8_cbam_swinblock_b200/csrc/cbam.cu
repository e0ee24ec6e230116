// CBAM (channel attention -> spatial attention -> gate), forward and backward, one launch each.
// Replaces cbam.py:29-38 / :48-53 / :62-71 of the reference (and their autograd backward).
//
// One thread-block CLUSTER per image (grid = CS x B, cluster = CS x 1).  The image (NHWC, so a pixel range is one
// contiguous byte range) is split into CS contiguous pixel chunks, one per CTA of the cluster, and each chunk is
// pulled into shared memory ONCE by a 1-D bulk TMA copy (cp.async.bulk -> UBLKCP).  Everything else happens on chip:
//   A  per-channel sum / max over the chunk (thread = one 16-byte channel vector, looping over pixels)
//      [cluster.sync]  every rank reduces the partials of all ranks over DSMEM and runs the tiny shared MLP + sigmoid
//                      redundantly (32 KB of weights from L2 per CTA is cheaper than two more cluster barriers)
//   B  per-pixel mean_c / max_c of x*ca (sub-warp per pixel, 16-byte vectors)  -> 2-channel map chunk in smem
//      [cluster.sync]  gather a zero-padded 2-D tile (rows of the chunk +-3, columns -3..W+3) of the map from the
//                      neighbouring ranks (DSMEM): the 7x7 taps then need no bounds checks
//   C  7x7 conv + sigmoid -> sa (4 lanes per pixel); out = x*ca*sa with packed 16-bit multiplies, 16-byte stores.
// HBM traffic = the algorithmic 1 read + 1 write of the feature map.  If a chunk does not fit in shared memory the
// same kernel runs "non-resident": phases A/B/C re-read the chunk from global memory (L2 at these sizes).
// A 3x3 spatial-attention kernel (cbam.py:43) runs as a 7x7 kernel embedded in zeros.  All shape-dependent
// constants (chunking, lane mapping, shared-memory offsets) are computed once on the host (`Plan`): at the model's
// shapes a CTA owns ~50 pixels, so scalar set-up code is what the kernel's latency is made of.
//
// NaN note: the channel/pixel maxima use the hardware max (which drops NaN) because every output they can reach is
// already NaN through the sum/mean computed over the same elements (a NaN in x makes the pooled average, hence the
// whole ca vector, NaN; a NaN in x*ca makes the channel mean at that pixel NaN, and the mean and max maps feed the
// same conv window) -- the result is NaN in exactly the positions where the reference's is.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace b200 {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int KS = 7, PADK = 3, KROW = 8;   // conv taps are stored [2][7][8] (rows padded to 8 floats)
constexpr int NTAPS = 2 * KS * KROW;        // 112 slots, 98 used
constexpr int NT7 = 2 * KS * KS;            // 98
constexpr int kMaxCS = 16;
constexpr int kRedBytes = 12288 + 1024;

// Host-computed launch plan (passed by value in the kernel parameters).
struct Plan {
  int B, C, H, W, HW, r, ksa, mode;
  int cs, pchunk, nch, lpp, groups, th, tw, resident;
  float invC, invHW;
  // shared-memory byte offsets
  int xs, gs, psum, pmax, pidx, gca, pav, pmx, hid, ca, vec, smap, tile, pix, poff, wsa, red, bar, total;
};

struct CbamParams {
  const void* x;
  const void* g;  // bwd only
  void* out;      // fwd: out; bwd: gx
  const float* w1;
  const float* w2;
  const float* wsa;
  float* ca;  // fwd: out (nullable); bwd: in
  float* sa;
  float* part;   // bwd: per-image MLP weight-grad partials [B][2rC]
  float* cpart;  // bwd: per-CTA conv weight-grad partials [B][kMaxCS][98]
  long long* prof;  // debug: per-CTA phase timestamps (b200_debug_cbam_prof), normally null
  Plan pl;
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
template <typename T> struct Pair2 { using type = float2; };
template <> struct Pair2<__nv_bfloat16> { using type = __nv_bfloat162; };
template <> struct Pair2<__half> { using type = __half2; };
__device__ __forceinline__ __nv_bfloat162 make_pair(__nv_bfloat16, float a, float b) { return __floats2bfloat162_rn(a, b); }
__device__ __forceinline__ __half2 make_pair(__half, float a, float b) { return __floats2half2_rn(a, b); }

__device__ __forceinline__ void prof_mark(const CbamParams& P, int slot) {
  if (P.prof && threadIdx.x == 0) {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    P.prof[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + slot] = t;
  }
}
long long* g_cbam_prof = nullptr;

inline int align16i(size_t v) { return (int)((v + 15) & ~(size_t)15); }

// conv weights -> smem [2][7][8], a 3x3 kernel centred in zeros
__device__ __forceinline__ void load_taps(float* wsas, const float* wsa, int ks) {
  for (int i = threadIdx.x; i < NTAPS; i += kThreads) {
    const int ch = i / (KS * KROW), u = (i / KROW) % KS, v = i % KROW;
    float w = 0.f;
    if (wsa && v < KS) {
      const int o = (KS - ks) / 2, uu = u - o, vv = v - o;
      if (uu >= 0 && uu < ks && vv >= 0 && vv < ks) w = wsa[(ch * ks + uu) * ks + vv];
    }
    wsas[i] = w;
  }
}

// bulk-TMA staging of one or two contiguous chunks; returns whether the caller has to wait on the mbarrier
template <typename T>
__device__ __forceinline__ bool stage_chunks(T* xs, const T* xg, T* gs, const T* gg, int n, uint64_t* bar) {
  const size_t bytes = (size_t)n * sizeof(T);
  const bool aligned = ((bytes & 15) == 0) && ((reinterpret_cast<uintptr_t>(xg) & 15) == 0) &&
                       (!gg || (reinterpret_cast<uintptr_t>(gg) & 15) == 0);
  if (aligned) {
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      fence_mbar_init();
      if (bytes) {
        mbar_expect_tx(bar, (uint32_t)(bytes * (gg ? 2 : 1)));
        for (size_t off = 0; off < bytes; off += 32768) {
          const uint32_t k = (uint32_t)((bytes - off) > 32768 ? 32768 : (bytes - off));
          bulk_g2s(reinterpret_cast<char*>(xs) + off, reinterpret_cast<const char*>(xg) + off, k, bar);
          if (gg) bulk_g2s(reinterpret_cast<char*>(gs) + off, reinterpret_cast<const char*>(gg) + off, k, bar);
        }
      }
    }
    return bytes != 0;
  }
  for (int i = threadIdx.x; i < n; i += kThreads) { xs[i] = xg[i]; if (gg) gs[i] = gg[i]; }
  return false;
}

// ---- phase A: per-channel sum / max (+ first argmax pixel when IDX) over this CTA's pixels ---------------------
template <typename T, int VW, bool IDX>
__device__ __forceinline__ void channel_partials(const T* xc, int np, int p0, int C, int nw, int groups, float* psum,
                                                 float* pmax, int* pidx, float* red) {
  const int tpg = nw < kThreads ? nw : kThreads;   // threads per pixel group
  const int pg = threadIdx.x / tpg, tw = threadIdx.x - pg * tpg;
  for (int w = tw; w < nw; w += kThreads) {
    float s[VW], m[VW];
    int mi[VW];
#pragma unroll
    for (int e = 0; e < VW; ++e) { s[e] = 0.f; m[e] = -INFINITY; mi[e] = p0; }
    if (pg < groups) {
      const T* src = xc + w * VW;
#pragma unroll 2
      for (int p = pg; p < np; p += groups) {
        float v[VW];
        Vec<T, VW>::load(src + (size_t)p * C, v);
#pragma unroll
        for (int e = 0; e < VW; ++e) {
          s[e] += v[e];
          if (IDX) {
            if (v[e] > m[e]) { m[e] = v[e]; mi[e] = p0 + p; }
          } else {
            m[e] = fmaxf(m[e], v[e]);
          }
        }
      }
    }
    if (groups == 1) {
      if (pg == 0) {
#pragma unroll
        for (int e = 0; e < VW; ++e) {
          psum[w * VW + e] = s[e]; pmax[w * VW + e] = m[e];
          if (IDX) pidx[w * VW + e] = mi[e];
        }
      }
    } else if (pg < groups) {
      float* rs = red + (size_t)pg * C * 3;
#pragma unroll
      for (int e = 0; e < VW; ++e) {
        rs[w * VW + e] = s[e];
        rs[C + w * VW + e] = m[e];
        if (IDX) reinterpret_cast<int*>(rs)[2 * C + w * VW + e] = mi[e];
      }
    }
  }
  if (groups > 1) {
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kThreads) {
      float s = 0.f, m = -INFINITY;
      int mi = p0;
      for (int gi = 0; gi < groups; ++gi) {  // groups interleave pixels: combine by (value, then smaller index)
        const float* rs = red + (size_t)gi * C * 3;
        s += rs[c];
        const float v = rs[C + c];
        if (IDX) {
          const int vi = reinterpret_cast<const int*>(rs)[2 * C + c];
          if (v > m || (v == m && vi < mi && v != -INFINITY)) { m = v; mi = vi; }
        } else {
          m = fmaxf(m, v);
        }
      }
      psum[c] = s; pmax[c] = m;
      if (IDX) pidx[c] = mi;
    }
  }
}

// gather NM zero-padded [nrow][tw] tiles (rows ya-3 .., columns -3 .. W+3; tile stride th*tw) of per-pixel maps
// that live, chunked by pchunk, in the `smap` buffers of the cluster's ranks.  Map j comes from smap + moff[j].
template <int NM>
__device__ __forceinline__ void gather_tiles(cg::cluster_group& cluster, float* tile, float* smap, int m0, int m1, int m2,
                                             int ya, int nrow, const Plan& L) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int row = warp; row < NM * nrow; row += kWarps) {
    const int j = row / nrow, ty = row - j * nrow, y = ya - PADK + ty;
    const int mo = j == 0 ? m0 : (j == 1 ? m1 : m2);
    float* dst = tile + (size_t)(j * L.th + ty) * L.tw;
    for (int tx = lane; tx < L.tw; tx += 32) {
      const int x = tx - PADK;
      float v = 0.f;
      if (y >= 0 && y < L.H && x >= 0 && x < L.W) {
        const int q = y * L.W + x;
        const int owner = min(q / L.pchunk, L.cs - 1);
        v = cluster.map_shared_rank(smap, owner)[mo + (q - owner * L.pchunk)];
      }
      dst[tx] = v;
    }
  }
}

__device__ __forceinline__ float relu_nan(float a) { return (a != a) ? a : fmaxf(a, 0.f); }  // torch.relu keeps NaN

template <typename T, bool RES, int VW>
__global__ void __launch_bounds__(kThreads) cbam_fwd_kernel(const __grid_constant__ CbamParams P) {
  cg::cluster_group cluster = cg::this_cluster();
  const Plan& L = P.pl;
  const int CS = L.cs, rank = blockIdx.x, b = blockIdx.y;
  const int C = L.C, HW = L.HW, W = L.W, r = L.r, nch = L.nch;
  const int p0 = min(rank * L.pchunk, HW), p1 = min(p0 + L.pchunk, HW), np = p1 - p0;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* xs = reinterpret_cast<T*>(smem_raw + L.xs);
  float* psum = reinterpret_cast<float*>(smem_raw + L.psum);
  float* pmax = reinterpret_cast<float*>(smem_raw + L.pmax);
  float* pav = reinterpret_cast<float*>(smem_raw + L.pav);
  float* pmx = reinterpret_cast<float*>(smem_raw + L.pmx);
  float* hid = reinterpret_cast<float*>(smem_raw + L.hid);
  float* ca = reinterpret_cast<float*>(smem_raw + L.ca);
  float* smap = reinterpret_cast<float*>(smem_raw + L.smap);
  float* tile = reinterpret_cast<float*>(smem_raw + L.tile);
  float* sas = reinterpret_cast<float*>(smem_raw + L.pix);
  float* wsas = reinterpret_cast<float*>(smem_raw + L.wsa);
  float* red = reinterpret_cast<float*>(smem_raw + L.red);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.bar);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  prof_mark(P, 0);

  const T* xg = reinterpret_cast<const T*>(P.x) + ((size_t)b * HW + p0) * C;
  const T* xc = xg;
  bool wait_tma = false;
  if (RES) {
    wait_tma = stage_chunks<T>(xs, xg, nullptr, nullptr, np * C, bar);
    xc = xs;
  }
  load_taps(wsas, P.wsa, L.ksa);   // overlaps the bulk copy
  __syncthreads();                 // mbarrier init + taps visible
  if (wait_tma) mbar_wait(bar, 0);
  prof_mark(P, 1);

  if (L.mode != B200_CBAM_SA) {
    channel_partials<T, VW, false>(xc, np, p0, C, nch, L.groups, psum, pmax, nullptr, red);
    prof_mark(P, 2);
    cluster.sync();  // (1) partials visible cluster-wide
    prof_mark(P, 3);
    for (int c = tid; c < C; c += kThreads) {   // pooled avg / max of every channel (cbam.py:8-9)
      float s = 0.f, m = -INFINITY;
      for (int k2 = 0; k2 < CS; ++k2) {
        s += cluster.map_shared_rank(psum, k2)[c];
        m = fmaxf(m, cluster.map_shared_rank(pmax, k2)[c]);
      }
      pav[c] = s * L.invHW;
      pmx[c] = (s != s) ? s : m;   // a NaN anywhere in the channel makes the pooled max NaN too
    }
    __syncthreads();
    for (int j = warp; j < 2 * r; j += kWarps) {  // hidden = relu(W1 pooled); j < r: avg branch, j >= r: max branch
      const float* wrow = P.w1 + (size_t)(j < r ? j : j - r) * C;
      const float* src = j < r ? pav : pmx;
      float acc = 0.f;
#pragma unroll 4
      for (int c = lane; c < C; c += 32) acc += wrow[c] * src[c];
      acc = warp_sum(acc);
      if (lane == 0) hid[j] = relu_nan(acc);
    }
    __syncthreads();
    for (int c = tid; c < C; c += kThreads) {
      const float* wrow = P.w2 + (size_t)c * r;
      float z = 0.f;
#pragma unroll 4
      for (int j = 0; j < r; ++j) z += wrow[j] * (hid[j] + hid[r + j]);
      const float a = sigmoidf_(z);
      ca[c] = a;
      if (P.ca && rank == 0) P.ca[(size_t)b * C + c] = a;
    }
    __syncthreads();
    prof_mark(P, 4);
    if (L.mode == B200_CBAM_CA) { cluster.sync(); return; }
  } else {
    for (int c = tid; c < C; c += kThreads) ca[c] = 1.f;
    __syncthreads();
  }

  // ---- phase B: per-pixel channel mean / max of x*ca (sub-warp of LPP lanes per pixel) ---------------------------
  const int LPP = L.lpp, PPW = 32 / LPP, sub = lane / LPP, sl = lane & (LPP - 1);
  const bool one = nch <= LPP;   // one chunk per lane: its ca values stay in registers
  float car[VW];
#pragma unroll
  for (int e = 0; e < VW; ++e) car[e] = (one && sl < nch) ? ca[sl * VW + e] : 0.f;
  for (int pb = warp * PPW; pb < np; pb += kWarps * PPW) {
    const int p = pb + sub;
    float s = 0.f, m = -INFINITY;
    if (p < np) {
      for (int w = sl; w < nch; w += LPP) {
        float v[VW];
        Vec<T, VW>::load(xc + (size_t)p * C + w * VW, v);
        if (!one) {
#pragma unroll
          for (int e = 0; e < VW; ++e) car[e] = ca[w * VW + e];
        }
#pragma unroll
        for (int e = 0; e < VW; ++e) {
          const float t = v[e] * car[e];
          s += t;
          m = fmaxf(m, t);
        }
      }
    }
    for (int o = LPP >> 1; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    if (sl == 0 && p < np) { smap[p] = s * L.invC; smap[L.pchunk + p] = m; }
  }
  prof_mark(P, 5);
  cluster.sync();  // (2) map chunks visible
  prof_mark(P, 6);
  const int ya = p0 / W, tw = L.tw;
  const int nrow = (np > 0 ? (p1 - 1) / W - ya + 1 : 0) + 2 * PADK;
  gather_tiles<2>(cluster, tile, smap, 0, L.pchunk, 0, ya, nrow, L);
  __syncthreads();
  prof_mark(P, 7);
  // ---- phase C: 7x7 conv + sigmoid; 4 lanes per pixel split the 14 tap rows, taps broadcast as float4 ------------
  for (int base = warp * 32; base < np * 4; base += kThreads) {
    const int i = base + lane, p = i >> 2, part = i & 3;
    float z = 0.f;
    if (p < np) {
      const int q = p0 + p, y = q / W, x = q - y * W;
      const float* t0 = tile + (size_t)(y - ya) * tw + x;
      for (int rr = part; rr < 2 * KS; rr += 4) {       // rr = ch*7 + u
        const int ch = rr >= KS ? 1 : 0, u = rr - ch * KS;
        const float4 wa = *reinterpret_cast<const float4*>(wsas + rr * KROW);
        const float4 wb = *reinterpret_cast<const float4*>(wsas + rr * KROW + 4);
        const float* tr = t0 + (size_t)(ch * L.th + u) * tw;
        z += wa.x * tr[0] + wa.y * tr[1] + wa.z * tr[2] + wa.w * tr[3] + wb.x * tr[4] + wb.y * tr[5] + wb.z * tr[6];
      }
    }
    z += __shfl_xor_sync(0xffffffffu, z, 1);
    z += __shfl_xor_sync(0xffffffffu, z, 2);
    if (part == 0 && p < np) {
      const float a = sigmoidf_(z);
      sas[p] = a;
      if (P.sa) P.sa[(size_t)b * HW + p0 + p] = a;
    }
  }
  __syncthreads();
  prof_mark(P, 8);
  if (L.mode == B200_CBAM_FULL) {
    T* og = reinterpret_cast<T*>(P.out) + ((size_t)b * HW + p0) * C;
    if constexpr (sizeof(T) == 2 && VW == 8) {
      // packed 16-bit gate: out = (x*ca)*sa, rounded after each product exactly like the 16-bit reference ops
      using P2 = typename Pair2<T>::type;
      P2 ca2[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) ca2[i] = make_pair(T(), car[2 * i], car[2 * i + 1]);
      for (int pb = warp * PPW; pb < np; pb += kWarps * PPW) {
        const int p = pb + sub;
        if (p >= np) continue;
        const float sp = sas[p];
        const P2 sp2 = make_pair(T(), sp, sp);
        for (int w = sl; w < nch; w += LPP) {
          uint4 raw = *reinterpret_cast<const uint4*>(xc + (size_t)p * C + w * VW);
          P2* rp = reinterpret_cast<P2*>(&raw);
          if (!one) {
#pragma unroll
            for (int i = 0; i < 4; ++i) ca2[i] = make_pair(T(), ca[w * VW + 2 * i], ca[w * VW + 2 * i + 1]);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) rp[i] = __hmul2(__hmul2(rp[i], ca2[i]), sp2);
          stg_stream16(og + (size_t)p * C + w * VW, raw);
        }
      }
    } else {
      for (int pb = warp * PPW; pb < np; pb += kWarps * PPW) {
        const int p = pb + sub;
        if (p >= np) continue;
        const float sp = sas[p];
        for (int w = sl; w < nch; w += LPP) {
          float v[VW];
          Vec<T, VW>::load(xc + (size_t)p * C + w * VW, v);
#pragma unroll
          for (int e = 0; e < VW; ++e) v[e] = v[e] * (one ? car[e] : ca[w * VW + e]) * sp;
          Vec<T, VW>::store(og + (size_t)p * C + w * VW, v);
        }
      }
    }
  }
  prof_mark(P, 9);
  cluster.sync();  // keep smem alive until every peer finished its DSMEM reads
  prof_mark(P, 10);
}

// =====================================================================================================
// backward.  Same cluster decomposition; x and g chunks both staged once.
// =====================================================================================================
template <typename T, bool RES, int VW>
__global__ void __launch_bounds__(kThreads) cbam_bwd_kernel(const __grid_constant__ CbamParams P) {
  cg::cluster_group cluster = cg::this_cluster();
  const Plan& L = P.pl;
  const int CS = L.cs, rank = blockIdx.x, b = blockIdx.y;
  const int C = L.C, HW = L.HW, W = L.W, r = L.r, nch = L.nch, pc = L.pchunk;
  const int p0 = min(rank * pc, HW), p1 = min(p0 + pc, HW), np = p1 - p0;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* xs = reinterpret_cast<T*>(smem_raw + L.xs);
  T* gs = reinterpret_cast<T*>(smem_raw + L.gs);
  float* psum = reinterpret_cast<float*>(smem_raw + L.psum);
  float* pmax = reinterpret_cast<float*>(smem_raw + L.pmax);
  int* pidx = reinterpret_cast<int*>(smem_raw + L.pidx);
  float* gca_part = reinterpret_cast<float*>(smem_raw + L.gca);  // [C] this CTA's partial of g_ca (peers read it)
  float* pav = reinterpret_cast<float*>(smem_raw + L.pav);
  float* pmx = reinterpret_cast<float*>(smem_raw + L.pmx);
  float* hid = reinterpret_cast<float*>(smem_raw + L.hid);       // [0,2r): pre-activations, [2r,3r): u = W2^T g_a
  float* ca = reinterpret_cast<float*>(smem_raw + L.ca);
  float* vec = reinterpret_cast<float*>(smem_raw + L.vec);       // [0]=g_pavg/HW, [1]=g_pmax, [2]=argmax_hw (int bits)
  float* smap = reinterpret_cast<float*>(smem_raw + L.smap);     // [0..1]: s map chunk, [2]: g_z chunk
  float* tile = reinterpret_cast<float*>(smem_raw + L.tile);     // [0]: g_z tile, [1..2]: s tiles
  float4* pix = reinterpret_cast<float4*>(smem_raw + L.pix);     // {sa, g_s0/C, g_s1, argmax_c (int bits)}
  int* poff = reinterpret_cast<int*>(smem_raw + L.poff);
  float* wsas = reinterpret_cast<float*>(smem_raw + L.wsa);
  float* red = reinterpret_cast<float*>(smem_raw + L.red);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + L.bar);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int mode = L.mode;

  const T* xg = reinterpret_cast<const T*>(P.x) + ((size_t)b * HW + p0) * C;
  const T* gg = mode == B200_CBAM_FULL ? reinterpret_cast<const T*>(P.g) + ((size_t)b * HW + p0) * C : nullptr;
  const T* xc = xg;
  const T* gc = gg;
  bool wait_tma = false;
  if (RES) {
    wait_tma = stage_chunks<T>(xs, xg, gs, gg, np * C, bar);
    xc = xs;
    gc = gg ? gs : nullptr;
  }
  load_taps(wsas, P.wsa, L.ksa);
  const bool use_ca = mode != B200_CBAM_SA;
  const bool use_sa = mode != B200_CBAM_CA;
  for (int c = tid; c < C; c += kThreads) ca[c] = use_ca ? P.ca[(size_t)b * C + c] : 1.f;
  const int ya = p0 / W, tw = L.tw;
  const int nrow = (np > 0 ? (p1 - 1) / W - ya + 1 : 0) + 2 * PADK;
  for (int p = tid; p < np; p += kThreads) {
    const int q = p0 + p, y = q / W;
    pix[p] = make_float4(use_sa ? P.sa[(size_t)b * HW + q] : 1.f, 0.f, 0.f, __int_as_float(-1));
    poff[p] = (y - ya) * tw + (q - y * W);
  }
  __syncthreads();
  if (wait_tma) mbar_wait(bar, 0);

  const int LPP = L.lpp, PPW = 32 / LPP, sub = lane / LPP, sl = lane & (LPP - 1);
  const bool one = nch <= LPP;

  if (use_sa) {
    // (1) recompute s map + channel argmax; g_sa[p] = sum_c g * x * ca   (SA mode: g_sa is the input itself)
    float car[VW];
#pragma unroll
    for (int e = 0; e < VW; ++e) car[e] = (one && sl < nch) ? ca[sl * VW + e] : 0.f;
    for (int pb = warp * PPW; pb < np; pb += kWarps * PPW) {
      const int p = pb + sub;
      float s = 0.f, m = -INFINITY, gsa = 0.f;
      int mi = 0x7fffffff;
      if (p < np) {
        for (int w = sl; w < nch; w += LPP) {
          float v[VW], gv[VW];
          Vec<T, VW>::load(xc + (size_t)p * C + w * VW, v);
          if (gc) Vec<T, VW>::load(gc + (size_t)p * C + w * VW, gv);
          if (!one) {
#pragma unroll
            for (int e = 0; e < VW; ++e) car[e] = ca[w * VW + e];
          }
#pragma unroll
          for (int e = 0; e < VW; ++e) {
            const float t = v[e] * car[e];
            s += t;
            if (t > m) { m = t; mi = w * VW + e; }
            if (gc) gsa += gv[e] * t;
          }
        }
      }
      for (int o = LPP >> 1; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        gsa += __shfl_xor_sync(0xffffffffu, gsa, o);
        const float tm = __shfl_xor_sync(0xffffffffu, m, o);
        const int ti = __shfl_xor_sync(0xffffffffu, mi, o);
        if (tm > m || (tm == m && ti < mi)) { m = tm; mi = ti; }   // larger value, then first channel (max.dim rule)
      }
      if (sl == 0 && p < np) {
        smap[p] = s * L.invC;
        smap[pc + p] = m;
        const float a = pix[p].x;
        const float gsa_in = mode == B200_CBAM_SA ? reinterpret_cast<const float*>(P.g)[(size_t)b * HW + p0 + p] : gsa;
        smap[2 * pc + p] = gsa_in * a * (1.f - a);  // g_z
        pix[p].w = __int_as_float(mi == 0x7fffffff ? 0 : mi);
      }
    }
    cluster.sync();  // (1) s map + g_z chunks visible
    gather_tiles<3>(cluster, tile, smap, 2 * pc, 0, pc, ya, nrow, L);
    __syncthreads();
    // g_s = conv^T(g_z): g_s[j][y][x] = sum_{u,v} w[j][u][v] * g_z[y-(u-3)][x-(v-3)]; 4 lanes per pixel split the rows u
    for (int base = warp * 32; base < np * 4; base += kThreads) {
      const int i = base + lane, p = i >> 2, part = i & 3;
      float g0 = 0.f, g1 = 0.f;
      if (p < np) {
        const float* t0 = tile + poff[p];
        for (int u = part; u < KS; u += 4) {
          const float* tr = t0 + (KS - 1 - u) * tw;
          const float* w0 = wsas + u * KROW;
          const float* w1 = wsas + (KS + u) * KROW;
#pragma unroll
          for (int v = 0; v < KS; ++v) {
            const float gz = tr[KS - 1 - v];
            g0 += w0[v] * gz;
            g1 += w1[v] * gz;
          }
        }
      }
      g0 += __shfl_xor_sync(0xffffffffu, g0, 1);
      g1 += __shfl_xor_sync(0xffffffffu, g1, 1);
      g0 += __shfl_xor_sync(0xffffffffu, g0, 2);
      g1 += __shfl_xor_sync(0xffffffffu, g1, 2);
      if (part == 0 && p < np) {
        pix[p].y = g0 * L.invC;  // broadcast share of the channel mean
        pix[p].z = g1;           // routed to the argmax channel
      }
    }
    // g_Wsa[j,u,v] = sum_p g_z[p] * s[j, p + (u-3, v-3)]: warp per tap over this CTA's pixels -> per-CTA partial in the
    // workspace (folded over images and ranks in a fixed order by fold_partials_kernel: deterministic, no atomics)
    float* cp = P.cpart + ((size_t)b * kMaxCS + rank) * NT7;
    for (int t = warp; t < NT7; t += kWarps) {
      const int j = t / (KS * KS), u = (t / KS) % KS, v = t % KS;
      const float* tj = tile + (size_t)((1 + j) * L.th + u) * tw + v;
      float acc = 0.f;
      for (int p = lane; p < np; p += 32) acc += smap[2 * pc + p] * tj[poff[p]];
      acc = warp_sum(acc);
      if (lane == 0) cp[t] = acc;
    }
    __syncthreads();  // pix complete
  }

  T* og = reinterpret_cast<T*>(P.out) + ((size_t)b * HW + p0) * C;
  if (mode == B200_CBAM_SA) {
    // gx = (g_s0 + [c == argmax] g_s1), no channel attention involved
    for (int i = tid; i < np * nch; i += kThreads) {
      const int p = i / nch, w = i - p * nch;
      const float4 pp = pix[p];
      const int am = __float_as_int(pp.w);
      float o[VW];
#pragma unroll
      for (int e = 0; e < VW; ++e) o[e] = pp.y + ((w * VW + e) == am ? pp.z : 0.f);
      Vec<T, VW>::store(og + (size_t)p * C + w * VW, o);
    }
    cluster.sync();
    return;
  }

  // (3)+(4): g_x1 = g*sa + g_s0 + [c==argmax_c] g_s1 ;  g_ca[c] = sum_p g_x1 * x  (per-CTA partial)
  // also the pooled statistics again (avg / max / argmax_hw) for the MLP backward.
  channel_partials<T, VW, true>(xc, np, p0, C, nch, L.groups, psum, pmax, pidx, red);
  __syncthreads();
  if (mode == B200_CBAM_FULL) {
    const int groups = L.groups;
    const int tpg = nch < kThreads ? nch : kThreads;
    const int pg = tid / tpg, twi = tid - pg * tpg;
    for (int w = twi; w < nch; w += kThreads) {
      float acc[VW];
#pragma unroll
      for (int e = 0; e < VW; ++e) acc[e] = 0.f;
      if (pg < groups) {
#pragma unroll 2
        for (int p = pg; p < np; p += groups) {
          float v[VW], gv[VW];
          Vec<T, VW>::load(xc + (size_t)p * C + w * VW, v);
          Vec<T, VW>::load(gc + (size_t)p * C + w * VW, gv);
          const float4 pp = pix[p];
          const unsigned d = (unsigned)(__float_as_int(pp.w) - w * VW);
#pragma unroll
          for (int e = 0; e < VW; ++e) acc[e] += (gv[e] * pp.x + pp.y) * v[e];
          if (d < (unsigned)VW) {   // the argmax channel of this pixel lies in this thread's vector (1 lane per pixel)
#pragma unroll
            for (int e = 0; e < VW; ++e) acc[e] += (d == (unsigned)e) ? pp.z * v[e] : 0.f;
          }
        }
#pragma unroll
        for (int e = 0; e < VW; ++e) red[(size_t)pg * C + w * VW + e] = acc[e];
      }
    }
    __syncthreads();
    for (int c = tid; c < C; c += kThreads) {
      float s = 0.f;
      for (int gi = 0; gi < groups; ++gi) s += red[(size_t)gi * C + c];
      gca_part[c] = s;
    }
  }
  cluster.sync();  // (2) g_ca partials + pooled partials visible
  // every rank: totals over the ranks for all channels, then the (tiny) MLP backward redundantly
  float* ga = red;  // [C] g_a = g_ca * ca * (1-ca)
  for (int c = tid; c < C; c += kThreads) {
    float s = 0.f, m = -INFINITY, gsum = 0.f;
    int mi = 0;
    for (int k2 = 0; k2 < CS; ++k2) {
      s += cluster.map_shared_rank(psum, k2)[c];
      const float v = cluster.map_shared_rank(pmax, k2)[c];
      const int vi = cluster.map_shared_rank(pidx, k2)[c];
      if (v > m) { m = v; mi = vi; }          // ranks ascend in pixel order, strict >: first occurrence
      else if (k2 == 0) { mi = vi; }
      if (mode == B200_CBAM_FULL) gsum += cluster.map_shared_rank(gca_part, k2)[c];
    }
    if (mode == B200_CBAM_CA) gsum = reinterpret_cast<const float*>(P.g)[(size_t)b * C + c];
    pav[c] = s * L.invHW;
    pmx[c] = (s != s) ? s : m;
    vec[2 * C + c] = __int_as_float(mi);
    const float a = ca[c];
    ga[c] = gsum * a * (1.f - a);
  }
  __syncthreads();
  // hidden pre-activations h[2][r] (forward recompute) and u[r] = W2^T g_a
  for (int j = warp; j < 3 * r; j += kWarps) {
    float acc = 0.f;
    if (j < 2 * r) {
      const float* wrow = P.w1 + (size_t)(j < r ? j : j - r) * C;
      const float* src = j < r ? pav : pmx;
#pragma unroll 4
      for (int c = lane; c < C; c += 32) acc += wrow[c] * src[c];
    } else {
      const float* wcol = P.w2 + (j - 2 * r);
#pragma unroll 4
      for (int c = lane; c < C; c += 32) acc += wcol[(size_t)c * r] * ga[c];
    }
    acc = warp_sum(acc);
    if (lane == 0) hid[j] = acc;
  }
  __syncthreads();
  // per channel: g_p_t = W1^T g_h_t; the rank's own channel slice also emits the weight-gradient partials
  {
    float* gw_part = P.part + (size_t)b * (2 * (size_t)r * C);  // per-image partials [r*C | C*r]
    const int cper = (C + CS - 1) / CS, cs0 = rank * cper, cs1 = min(cs0 + cper, C);
    for (int c = tid; c < C; c += kThreads) {
      const float gac = ga[c];
      const bool mine = c >= cs0 && c < cs1;
      float gpa = 0.f, gpm = 0.f;
      for (int j = 0; j < r; ++j) {
        const float ha = hid[j], hm = hid[r + j], u = hid[2 * r + j];
        const float gha = ha > 0.f ? u : 0.f, ghm = hm > 0.f ? u : 0.f;
        if (mine) {
          gw_part[(size_t)r * C + (size_t)c * r + j] = gac * (fmaxf(ha, 0.f) + fmaxf(hm, 0.f));  // g_W2[c,j]
          gw_part[(size_t)j * C + c] = gha * pav[c] + ghm * pmx[c];                              // g_W1[j,c]
        }
        const float w = P.w1[(size_t)j * C + c];
        gpa += w * gha;
        gpm += w * ghm;
      }
      vec[c] = gpa * L.invHW;
      vec[C + c] = gpm;
    }
  }
  __syncthreads();
  // (6) g_x = g_x1 * ca + g_pavg/HW (+ g_pmax at the pooled-max pixel, patched in below)
  {
    float car[VW], gav[VW];
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      car[e] = (one && sl < nch) ? ca[sl * VW + e] : 0.f;
      gav[e] = (one && sl < nch) ? vec[sl * VW + e] : 0.f;
    }
    const bool full = mode == B200_CBAM_FULL;
    for (int pb = warp * PPW; pb < np; pb += kWarps * PPW) {
      const int p = pb + sub;
      if (p >= np) continue;
      const float4 pp = pix[p];
      const int am = __float_as_int(pp.w);
      for (int w = sl; w < nch; w += LPP) {
        float gv[VW], o[VW];
        if (full) Vec<T, VW>::load(gc + (size_t)p * C + w * VW, gv);
        if (!one) {
#pragma unroll
          for (int e = 0; e < VW; ++e) { car[e] = ca[w * VW + e]; gav[e] = vec[w * VW + e]; }
        }
        const unsigned d = (unsigned)(am - w * VW);
#pragma unroll
        for (int e = 0; e < VW; ++e) {
          float gx1 = 0.f;
          if (full) gx1 = gv[e] * pp.x + pp.y;
          o[e] = gx1 * car[e] + gav[e];
        }
        if (full && d < (unsigned)VW) {
#pragma unroll
          for (int e = 0; e < VW; ++e) o[e] += (d == (unsigned)e) ? pp.z * car[e] : 0.f;
        }
        Vec<T, VW>::store(og + (size_t)p * C + w * VW, o);
      }
    }
  }
  __syncthreads();
  // the pooled-max pixel of each channel receives g_pmax on top (adaptive_max_pool2d backward, first occurrence)
  for (int c = tid; c < C; c += kThreads) {
    const int pm = __float_as_int(vec[2 * C + c]) - p0;
    if (pm >= 0 && pm < np) {
      T* dst = og + (size_t)pm * C + c;
      *dst = DT<T>::from_f(DT<T>::to_f(*dst) + vec[C + c]);
    }
  }
  cluster.sync();  // keep smem alive until every peer finished its DSMEM reads
}

// fold the weight-gradient partials in a fixed order (deterministic): MLP weights [B][n12] -> [n12] (thread per
// element), conv taps [B][kMaxCS][98] -> [2*ks*ks] (warp per tap; blocks past the MLP part)
__global__ void fold_partials_kernel(const float* __restrict__ part, const float* __restrict__ cpart, float* gw1,
                                     float* gw2, float* gwsa, int B, int CS, int n1, int n2, int ks, int mlp_blocks) {
  if ((int)blockIdx.x < mlp_blocks) {
    const int n = n1 + n2, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int b = 0;
    for (; b + 4 <= B; b += 4) {
      a0 += part[(size_t)b * n + i];
      a1 += part[(size_t)(b + 1) * n + i];
      a2 += part[(size_t)(b + 2) * n + i];
      a3 += part[(size_t)(b + 3) * n + i];
    }
    for (; b < B; ++b) a0 += part[(size_t)b * n + i];
    const float acc = (a0 + a1) + (a2 + a3);
    if (i < n1) gw1[i] = acc; else gw2[i - n1] = acc;
    return;
  }
  const int warps = blockDim.x / 32, lane = threadIdx.x & 31;
  const int t = ((int)blockIdx.x - mlp_blocks) * warps + (threadIdx.x >> 5);
  if (t >= 2 * ks * ks) return;
  const int j = t / (ks * ks), uu = (t / ks) % ks, vv = t % ks, o = (KS - ks) / 2;
  const int slot = (j * KS + uu + o) * KS + vv + o;
  float acc = 0.f;
  for (int i = lane; i < B * CS; i += 32) {
    const int b = i / CS, k2 = i - b * CS;
    acc += cpart[((size_t)b * kMaxCS + k2) * NT7 + slot];
  }
  acc = warp_sum(acc);
  if (lane == 0) gwsa[t] = acc;
}

template <typename K>
int launch_cluster(K kern, const CbamParams& P, cudaStream_t st) {
  const Plan& L = P.pl;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
  if (L.cs > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(L.cs, L.B);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = L.total;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = L.cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, P);
  if (e != cudaSuccess) {
    set_error("cbam: cluster launch failed (grid=%dx%d cs=%d smem=%d): %s", L.cs, L.B, L.cs, L.total, cudaGetErrorString(e));
    cudaGetLastError();
    return B200_ERR_LAUNCH;
  }
  return check_launch("cbam");
}

int check_common(const void* x, int B, int C, int H, int W, int r, int ksa, int dtype, int mode) {
  B200_REQUIRE(x, B200_ERR_SHAPE, "cbam: null input");
  B200_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, B200_ERR_SHAPE, "cbam: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
  B200_REQUIRE(B <= 65535, B200_ERR_SHAPE, "cbam: batch %d > 65535", B);
  B200_REQUIRE(mode >= 0 && mode <= 2, B200_ERR_SHAPE, "cbam: bad mode %d", mode);
  if (mode != B200_CBAM_SA) B200_REQUIRE(r > 0, B200_ERR_SHAPE, "cbam: hidden width r must be > 0");
  if (mode != B200_CBAM_CA) B200_REQUIRE(ksa == 3 || ksa == 7, B200_ERR_SHAPE, "cbam: kernel size must be 3 or 7 (cbam.py:43), got %d", ksa);
  if (dtype != B200_F32) B200_REQUIRE(C % 2 == 0, B200_ERR_ALIGN, "cbam: C must be even for 16-bit dtypes");
  B200_REQUIRE(((uintptr_t)x & 3) == 0, B200_ERR_ALIGN, "cbam: input must be 4-byte aligned");
  return B200_OK;
}

// shared-memory carve-up for a given chunking
void layout(Plan& L, size_t esize, bool bwd) {
  const int C = L.C, r = L.r, pc = L.pchunk;
  L.tw = L.W + 2 * PADK;
  L.th = (pc - 1) / L.W + 2 + 2 * PADK;
  const size_t tsz = (size_t)L.th * L.tw;
  size_t o = 0;
  auto take = [&](size_t bytes) { const int at = (int)o; o += (bytes + 15) & ~(size_t)15; return at; };
  L.xs = take(L.resident ? (size_t)pc * C * esize : 0);
  L.gs = take((L.resident && bwd) ? (size_t)pc * C * esize : 0);
  L.psum = take((size_t)C * 4);
  L.pmax = take((size_t)C * 4);
  L.pidx = take(bwd ? (size_t)C * 4 : 0);
  L.gca = take(bwd ? (size_t)C * 4 : 0);
  L.pav = take((size_t)C * 4);
  L.pmx = take((size_t)C * 4);
  L.hid = take((size_t)3 * r * 4);
  L.ca = take((size_t)C * 4);
  L.vec = take(bwd ? (size_t)3 * C * 4 : 0);
  L.smap = take((size_t)pc * 4 * (bwd ? 3 : 2));
  L.tile = take(tsz * 4 * (bwd ? 3 : 2));
  L.pix = take((size_t)pc * (bwd ? 16 : 4));
  L.poff = take(bwd ? (size_t)pc * 4 : 0);
  L.wsa = take((size_t)NTAPS * 4);
  L.red = take(L.groups > 1 ? (size_t)L.groups * C * 12 : (size_t)C * 4);
  L.bar = take(16);
  L.total = (int)o;
}

// choose cluster size + residency and fill the plan
void make_plan(Plan& L, int B, int C, int H, int W, int r, int ksa, int mode, int vw, size_t esize, bool bwd) {
  L.B = B; L.C = C; L.H = H; L.W = W; L.HW = H * W; L.r = r; L.ksa = ksa; L.mode = mode;
  L.nch = C / vw;
  L.lpp = 1;
  while (L.lpp < L.nch && L.lpp < 32) L.lpp <<= 1;
  L.groups = L.nch >= kThreads ? 1 : std::min(kThreads / L.nch, std::max(1, kRedBytes / (C * 12)));
  L.invC = 1.f / (float)C;
  L.invHW = 1.f / (float)L.HW;
  const size_t lim = (size_t)max_smem_optin();
  for (int c : {8, 16}) {
    L.cs = c; L.pchunk = (L.HW + c - 1) / c; L.resident = 1;
    layout(L, esize, bwd);
    // 16-CTA clusters are non-portable: only take them when 2 CTAs still fit per SM
    if ((size_t)L.total <= (c == 8 ? lim : (size_t)100 * 1024)) return;
  }
  L.cs = 8; L.pchunk = (L.HW + 7) / 8; L.resident = 0;
  layout(L, esize, bwd);
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
  const int ve = 16 / (int)esize, ew = 4 / (int)esize;
  const bool vec16 = C % ve == 0 && ((uintptr_t)x & 15) == 0 && (mode != B200_CBAM_FULL || ((uintptr_t)out & 15) == 0);
  CbamParams P{x, nullptr, out, w1, w2, wsa, ca_out, sa_out, nullptr, nullptr, g_cbam_prof, {}};
  make_plan(P.pl, B, C, H, W, r, ksa, mode, vec16 ? ve : ew, esize, false);
  B200_REQUIRE((size_t)P.pl.total <= (size_t)max_smem_optin(), B200_ERR_UNSUPPORTED, "cbam_fwd: shape needs %d B of shared memory", P.pl.total);
  cudaStream_t st = (cudaStream_t)stream;
  const bool res = P.pl.resident != 0;
  return B200_DISPATCH_DTYPE(dtype, [&]() -> int {
    constexpr int VE = Words<T>::VE, EW = Words<T>::EPL;
    if (vec16)
      return res ? launch_cluster(cbam_fwd_kernel<T, true, VE>, P, st) : launch_cluster(cbam_fwd_kernel<T, false, VE>, P, st);
    return res ? launch_cluster(cbam_fwd_kernel<T, true, EW>, P, st) : launch_cluster(cbam_fwd_kernel<T, false, EW>, P, st);
  });
}

extern "C" B200_API size_t b200_cbam_bwd_workspace_bytes(int32_t B, int32_t C, int32_t H, int32_t W, int32_t r, int32_t ksa) {
  (void)H; (void)W; (void)ksa;
  return (size_t)B * (2 * (size_t)r * C + (size_t)b200::kMaxCS * b200::NT7) * sizeof(float);
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
  if (mode != B200_CBAM_SA) B200_REQUIRE(w1 && w2 && ca && gw1 && gw2, B200_ERR_SHAPE, "cbam_bwd: null MLP weights / ca map / gradients");
  if (mode != B200_CBAM_CA) B200_REQUIRE(wsa && sa && gwsa, B200_ERR_SHAPE, "cbam_bwd: null conv weight / sa map / gradient");
  const size_t need = b200_cbam_bwd_workspace_bytes(B, C, H, W, r, ksa);
  B200_REQUIRE(workspace && workspace_bytes >= need, B200_ERR_WORKSPACE, "cbam_bwd: workspace %zu < %zu bytes", workspace_bytes, need);
  const size_t esize = dtype == B200_F32 ? 4 : 2;
  const int ve = 16 / (int)esize, ew = 4 / (int)esize;
  const bool vec16 = C % ve == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)gx & 15) == 0 &&
                     (mode != B200_CBAM_FULL || ((uintptr_t)g & 15) == 0);
  float* part = (float*)workspace;
  float* cpart = part + (size_t)B * 2 * r * C;
  CbamParams P{x, g, gx, w1, w2, wsa, const_cast<float*>(ca), const_cast<float*>(sa), part, cpart, nullptr, {}};
  make_plan(P.pl, B, C, H, W, r, ksa, mode, vec16 ? ve : ew, esize, true);
  B200_REQUIRE((size_t)P.pl.total <= (size_t)max_smem_optin(), B200_ERR_UNSUPPORTED, "cbam_bwd: shape needs %d B of shared memory", P.pl.total);
  cudaStream_t st = (cudaStream_t)stream;
  const bool res = P.pl.resident != 0;
  int rc = B200_DISPATCH_DTYPE(dtype, [&]() -> int {
    constexpr int VE = Words<T>::VE, EW = Words<T>::EPL;
    if (vec16)
      return res ? launch_cluster(cbam_bwd_kernel<T, true, VE>, P, st) : launch_cluster(cbam_bwd_kernel<T, false, VE>, P, st);
    return res ? launch_cluster(cbam_bwd_kernel<T, true, EW>, P, st) : launch_cluster(cbam_bwd_kernel<T, false, EW>, P, st);
  });
  if (rc) return rc;
  const int n1 = mode != B200_CBAM_SA ? r * C : 0, n2 = n1;
  const int mlp_blocks = (n1 + n2 + 127) / 128;
  const int tap_blocks = mode != B200_CBAM_CA ? (2 * ksa * ksa + 3) / 4 : 0;
  fold_partials_kernel<<<mlp_blocks + tap_blocks, 128, 0, st>>>(part, cpart, gw1, gw2, gwsa, B, P.pl.cs, n1, n2, ksa, mlp_blocks);
  return check_launch("cbam_bwd_fold");
}

/* debug hook (not part of the drop-in ABI): device buffer of [grid][16] int64 that receives %globaltimer stamps of the
 * forward kernel's phases for CTA-level latency analysis (profiles/cbam_phases.py); NULL switches it off. */
extern "C" B200_API void b200_debug_cbam_prof(void* dev_buffer) { b200::g_cbam_prof = (long long*)dev_buffer; }
