// Pieces of the TMA-streamed CSR SpMV shared by solve.cu (k_spmv_tma) and dist.cu (k_dist_spmv_tma): chunk
// descriptors, stage geometry, mbarrier / bulk-copy wrappers.  Not part of the public ABI.
#pragma once
#include "common.cuh"

namespace efb {

// Product i of a chunk is parked at slot i + i/16: the row sums read the products with one lane per row, i.e. at a
// stride of one row length (~16 entries = 256 B), which without the skew lands every lane of a 16-byte access phase
// in the same bank group (measured: 12 M bank conflicts, 60 % of all L1 data-pipe wavefronts of the SpMV).
__device__ __forceinline__ int spmv_slot(int i) { return i + (i >> 4); }
constexpr int SPMV_SLOTS = SPMV_STREAM_W + SPMV_STREAM_W / 16;  // 272

// chunk descriptors and stage geometry of the TMA-streamed SpMV below
constexpr int SPMV_STAGE_COLS = SPMV_STREAM_W + 4;  // aligned superset of the column slice
constexpr int SPMV_VAL_BYTES = SPMV_SLOTS * 16;  // value area of a stage, sized for the skewed products written in place
struct ChunkDesc {
  int r0, nrow, k0, k1;
};
__device__ __forceinline__ ChunkDesc load_chunk_desc(const int32_t *__restrict__ sp_chunk, const int32_t *__restrict__ rowptr, int ch,
                                                      int n_chunks) {
  ChunkDesc d{0, 0, 0, 0};
  if (ch < n_chunks) {
    d.r0 = __ldg(&sp_chunk[ch]);
    const int r1 = __ldg(&sp_chunk[ch + 1]);
    d.nrow = r1 - d.r0;
    d.k0 = __ldg(&rowptr[d.r0]);
    d.k1 = __ldg(&rowptr[r1]);
  }
  return d;
}

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int SPMV_TMA_STAGE = SPMV_VAL_BYTES + SPMV_STAGE_COLS * 4;  // 5392 B (16-byte multiple)


}  // namespace efb
