// satd_tc.cuh - 8x8 Hadamard SATD on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// SATD(o - p) = sum | (H8 (x) H8) (o - p) |.  The 64x64 matrix H8 (x) H8 has entries +-1, the pixels are
// 8-bit, so  D[r][j] = sum_k o_r[k] * B[j][k] - sum_k p_r[k] * B[j][k]  is an exact int8 x int8 -> int32
// tensor-core product: A = 128 tiles x 64 pixels (u8, K-major, one tile per TMEM lane), B = H8 (x) H8
// (s8), accumulated in TMEM, and the epilogue only has to sum |D[r][0..63]| per lane.
// north_star adopts this path only where ncu shows it beats the integer-ALU butterflies.
//
// Shared-memory operand layout (UMMA K-major, no swizzle, see cute/atom/mma_traits_sm100.hpp
// make_umma_desc<Major::K>): 16-byte chunks; chunk c of row r lives at
//     base + c * LBO + (r / 8) * SBO + (r % 8) * 16      with SBO = 128, LBO = rows * 16
// i.e. [chunk][row-group][8 rows][16 B]; one kind::i8 MMA consumes K = 32 bytes = 2 chunks.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace cucd {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE descriptor (version 1 = Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t smemAddr, uint32_t lboBytes, uint32_t sboBytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smemAddr >> 4) & 0x3fffu);
  d |= (uint64_t)((lboBytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sboBytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// kind::i8 instruction descriptor: D s32, A u8 (aSigned = 0) or s8, B s8, both K-major, M x N
__device__ __forceinline__ uint32_t make_idesc_i8(int M, int N, int aSigned) {
  uint32_t d = 0;
  d |= 2u << 4;                       // c_format = S32
  d |= (uint32_t)(aSigned ? 1 : 0) << 7;
  d |= 1u << 10;                      // b_format = signed 8 bit
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
__device__ __forceinline__ void mma_i8(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      :: "r"(tmemD), "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n"
      :: "r"(smem_u32(mbar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slotInSmem, uint32_t cols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(slotInSmem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {        // the same full warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(addr), "r"(cols) : "memory");
}
// lane = TMEM lane of this thread's warp quarter, 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(addr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// B = H8 (x) H8 as s8 in the canonical K-major layout for N = 64 rows (coefficients), K = 64 pixels;
// `sign` = +1 or -1.  Called cooperatively by `nthreads` threads.  4 KB.
__device__ __forceinline__ void fill_hadamard64(int8_t* dst, int sign, int tid, int nthreads) {
  for (int i = tid; i < 64 * 64; i += nthreads) {
    const int j = i >> 6, k = i & 63;                 // coefficient j = (u, v), pixel k = (y, x)
    const int u = j >> 3, v = j & 7, y = k >> 3, x = k & 7;
    const int s = (__popc(u & y) + __popc(v & x)) & 1; // natural-ordered Hadamard entry (-1)^(<u,y> + <v,x>)
    const int c = k >> 4;                              // 16-byte chunk along K
    dst[c * (64 * 16) + (j >> 3) * 128 + (j & 7) * 16 + (k & 15)] = (int8_t)((s ? -1 : 1) * sign);
  }
}

}  // namespace tc
}  // namespace cucd
