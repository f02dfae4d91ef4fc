// Helpers shared by the skinning kernels (lbs.cu: real vertex tiles; joints.cu: virtual joint tiles).
#pragma once

#include "common.cuh"

namespace b200smpl {

constexpr int CHUNK_WORDS = TILE_V * 3 * 32;     // 3072 floats = 12 KB: one (tile, group) block of the blend output
constexpr int AG_WORDS = NJ * AELEMS * 32;       // 9216 floats = 36 KB: one group's skinning transforms
constexpr int TPITCH = 33;                       // transposition tile [96 columns][33]: (33 c + r) % 32 = (c + r) % 32,
constexpr int TTILE_WORDS = TILE_V * 3 * TPITCH; // conflict-free both for lane = body (fixed c) and lane = column (fixed r)

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > (1u << 26)) __trap();   // a lost completion must fail loudly, not hang the GPU
  }
}
// global -> shared bulk-async copy (TMA unit), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}

// one thread: initialise the barrier and pull the group's A[24][12][32] (36 KB, contiguous) into smem
__device__ __forceinline__ void fetch_group_transforms(float* A_s, const float* A_blk, int g, uint64_t* bar) {
  mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  mbar_expect_tx(bar, AG_WORDS * 4);
  bulk_g2s(A_s, A_blk + (size_t)g * AG_WORDS, AG_WORDS * 4, bar);
}

// streaming global accesses: read-once inputs bypass L1, written-once outputs are not kept hot
__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(float* p, float v) {
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void st_stream_u32(void* p, uint32_t v) {
  asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void load_slot(float (&a)[AELEMS], const float* A_s, int joint, int lane) {
#pragma unroll
  for (int e = 0; e < AELEMS; ++e) a[e] = A_s[(joint * AELEMS + e) * 32 + lane];
}
__device__ __forceinline__ void load_rot(float (&a)[9], const float* A_s, int joint, int lane) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) a[r * 3 + c] = A_s[(joint * AELEMS + r * 4 + c) * 32 + lane];
}
__device__ __forceinline__ void flush_slot(float (&d)[AELEMS], float* dA_s, int joint, int lane) {
#pragma unroll
  for (int e = 0; e < AELEMS; ++e) {
    atomicAdd(&dA_s[(joint * AELEMS + e) * 32 + lane], d[e]);
    d[e] = 0.f;
  }
}

// add a CTA's shared accumulators [rows][32] into the slab-wide buffer (fp32 RED, rows that stayed 0 skipped)
__device__ __forceinline__ void accumulate_rows(float* __restrict__ dst, const float* src_s, int rows, int warp,
                                                int nwarps, int lane) {
  for (int r = warp; r < rows; r += nwarps) {
    const float v = src_s[r * 32 + lane];
    if (__any_sync(0xffffffffu, v != 0.f)) atomicAdd(dst + r * 32 + lane, v);
  }
}

// x -> bf16 hi | bf16 lo << 16   (x ~ hi + lo to 16 mantissa bits)
__device__ __forceinline__ uint32_t pack_hi_lo(float x) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(x);
  const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
  return (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
}

// flush a transposition tile of packed (hi | lo<<16) words into the K-major bf16 operand rows
// dvp_{hi,lo}[row0 + r][col_base .. col_base + 96): even lanes store two columns per 32-bit word
__device__ __forceinline__ void flush_dvp_tile(const uint32_t* tile_u, __nv_bfloat16* dvp_hi, __nv_bfloat16* dvp_lo,
                                               size_t row0, int n_pad, size_t col_base, int lane) {
#pragma unroll 4
  for (int r = 0; r < 32; ++r) {
    const size_t o = (row0 + r) * (size_t)n_pad + col_base + lane;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const uint32_t u0 = tile_u[(lane + 32 * k) * TPITCH + r];
      const uint32_t u1 = __shfl_down_sync(0xffffffffu, u0, 1);
      if (!(lane & 1)) {
        st_stream_u32(dvp_hi + o + 32 * k, (u0 & 0xFFFFu) | (u1 << 16));
        if (dvp_lo != nullptr) st_stream_u32(dvp_lo + o + 32 * k, (u0 >> 16) | (u1 & 0xFFFF0000u));
      }
    }
  }
}

}  // namespace b200smpl
