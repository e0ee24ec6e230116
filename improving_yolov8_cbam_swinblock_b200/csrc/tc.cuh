// tcgen05 / TMEM / TMA (tensor-map) primitives for sm_100a, as inline PTX.  No CUTLASS dependency.
#pragma once
#include <cuda.h>  // CUtensorMap (types only; the encode entry point is fetched at run time, no libcuda link)

#include "common.cuh"

namespace b200 {
namespace tc {

// ---- host: tensor maps -------------------------------------------------------------------------------------
// 2-D row-major [rows, cols] tensor of 16-bit elements, box = [box_rows, box_cols], 128-byte swizzle
// (box_cols * 2 bytes must be <= 128).  Cached per (ptr, shape, box).  Returns nullptr + sets error on failure.
const CUtensorMap* tensor_map_2d(const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                                 uint32_t box_rows, uint32_t box_cols, int dtype);

// NHWC map [B, H, W, C] of 16-bit elements as a 4-D tensor, box = [1, box_h, box_w, 64 channels], 128-byte swizzle: one box
// is a window of box_h x box_w pixels landing as box_h*box_w consecutive 128-byte rows (a K-major UMMA operand block).
// Out-of-bounds pixels read as zero and are skipped on store -- the zero padding / crop of swin_block.py:41-45,58.
const CUtensorMap* tensor_map_nhwc(const void* base, uint64_t B, uint64_t H, uint64_t W, uint64_t C, uint32_t box_h, uint32_t box_w,
                                   int dtype);

// ---- device ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

// TMA: global (tensor map) -> shared, 2-D tile, completion on mbarrier.  c0 = inner (column) coord, c1 = row coord.
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
// TMA: shared -> global (tensor map), 2-D tile, bulk async-group completion.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src_smem, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(smem_u32(src_smem))
               : "memory");
}

// 4-D variants (coordinates innermost first: channel, x, y, image)
__device__ __forceinline__ void tma_load_4d(void* dst_smem, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src_smem, int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(src_smem))
               : "memory");
}

// ---- TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // whole warp, ncols pow2 >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread (thread = lane of the warp's TMEM quadrant)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors ----
// K-major operand tile in shared memory, 128-byte swizzle: rows of 64 16-bit elements (128 B), 8-row atoms of
// 1024 B (SBO), exactly what a SWIZZLE_128B TMA box [rows][64] produces.  Tile base must be 1024-byte aligned;
// advancing K by 16 elements inside the 64-wide block = +32 bytes on the start address.
__device__ __forceinline__ uint64_t smem_desc_k_sw128(const void* smem_ptr) {
  const uint32_t addr = smem_u32(smem_ptr);
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);   // start address, bits [0,14)
  d |= (uint64_t)0 << 16;                  // leading byte offset: unused for swizzled K-major
  d |= (uint64_t)(1024 >> 4) << 32;        // stride byte offset between 8-row atoms, bits [32,46)
  d |= (uint64_t)1 << 46;                  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                  // layout type SWIZZLE_128B
  return d;
}
// MN-major operand tile, 128-byte swizzle: the tile is stored as [K rows][64 MN-elements = 128 B] boxes (again what a
// SWIZZLE_128B TMA box produces when the MN dimension is the contiguous one in global memory).  8 K-rows form a
// 1024-byte atom (SBO); consecutive 64-element MN blocks are `lbo_bytes` apart.  Advancing K by 16 = +16 rows = +2048 B.
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(const void* smem_ptr, uint32_t lbo_bytes) {
  const uint32_t addr = smem_u32(smem_ptr);
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor, kind::f16: D=f32, A/B = bf16 (fmt 1) or f16 (fmt 0), both K-major, M x N tile
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int fmt, int a_mn = 0, int b_mn = 0) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues for the CTA
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// byte offset of 16-byte chunk `chunk` (0..7) of row `row` in a SWIZZLE_128B [rows][128 B] tile
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk) { return row * 128u + ((chunk ^ (row & 7u)) << 4); }

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace tc
}  // namespace b200
