// Fused backward (opt-in, B200_FUSED_BWD=1): the skinning backward (lbs.cu::lbs_bwd_kernel) produces the A operand of
// the gradient GEMM (blend_umma.cu::umma_gemm2_kernel) straight into TENSOR MEMORY, so dv_posed -- 82.7 KB per mesh
// written and read back as bf16 hi / lo by the two-kernel path -- never exists in HBM (DRAM traffic of the backward
// 1.42 GB -> 0.73 GB at B = 4096).  Parity-green, but NOT faster today: 311 us + 23 us (virtual-row GEMM) against
// 209 + 93 us.  The producers execute the same arithmetic as lbs_bwd 8 % more instructions at a 17 % lower issue rate
// (one dV staging tile instead of two, v_posed through registers instead of TMA, two warps per body group instead
// of eight on one) and lose ~20 % to pipeline fill / drain, so the GEMM that now hides behind them does not pay
// for it (profiles/r02_experiments.md).  Kept as the starting point for that tuning; the default path is unchanged.
//
//   dfeat[b][f] = sum_n dvp[b][n] * Wb[f][n]      M = bodies (TMEM lane = producer thread), N = 224 features, K = rows
//
// CTA pair (tcgen05 cta_group::2, M = 256): each CTA owns 128 bodies = 4 groups.  12 warps:
//   warp 0      TMA: streams this CTA's half (112 features) of the Wb rows of one 16-vertex item (48 rows, hi + lo:
//               two 10.5 KB bulk copies, pre-blocked at pack time into the no-swizzle core-matrix layout)
//   warp 1      leader: MMA issuer (A from tensor memory -- "TS" form, probed in scripts/exp/ts_mma_probe.cu -- B
//               from shared memory; three products hi.hi, lo.hi, hi.lo per K step); peer: relays "stage landed"
//   warps 4-11  producers, two per TMEM lane quadrant (= body group): lane = body.  A warp walks a contiguous run of
//               items of its group exactly like lbs_bwd (dV rows by 8-byte cp.async through a transposing tile,
//               v_posed by coalesced 16-byte loads one half item ahead, four cached rotations as packed pairs, dA
//               and dtransl by fp32 REDs), splits the gradient rows to bf16 hi / lo pairs and writes them with
//               tcgen05.st into one of six 48-column stages of the TMEM ring next to the 224-column accumulator.
//               They run on the registers warpgroup 0 gives up (setmaxnreg: 56 / 224 / 224).
// The two warps of a quadrant take the first and the second half of the cluster's item range (slots persist along a
// run); their stages interleave, which only permutes the K order of the sum.  The rotations of the CTA's 128 bodies
// stay in shared memory for the whole kernel in a compacted 9-float form (110 KB; the 12-float transforms would
// not fit next to the Wb ring).  The work is the flat (body pair, item) list cut into one range per TPC
// (build_bs_work); every piece writes its own split-K partial, summed by pose_bwd; the virtual (joint) rows keep
// joints_bwd and a small GEMM of their own.
#include <algorithm>

#include "lbs_tiles.cuh"
#include "umma_common.cuh"

namespace b200smpl {

constexpr int FB_BN = 224;                            // gradient features (nf_pad)
constexpr int FB_BNH = FB_BN / 2;                     // rows of Wb each CTA of the pair stages
constexpr int FB_ITEM_V = 16;                         // vertices per item = one TMEM stage
constexpr int FB_ITEM_ROWS = 3 * FB_ITEM_V;           // 48 K rows = 3 MMA K steps
constexpr int FB_NSTAGE_A = 6;                        // TMEM ring: 6 x (24 hi + 24 lo columns) behind the accumulator
constexpr int FB_A_COL0 = FB_BN;                      // first ring column
constexpr int FB_A_STAGE_COLS = FB_ITEM_ROWS;         // 48
constexpr int FB_NSTAGE_B = 3;                        // Wb stages in shared memory
constexpr int FB_BCHUNK = FB_BNH * 16;                // 1792 B: 8 K rows x 112 features (LBO of the B descriptor)
constexpr int FB_BTILE = (FB_ITEM_ROWS / 8) * FB_BCHUNK;   // 10752 B: hi (or lo) rows of one item
constexpr int FB_BSTAGE = 2 * FB_BTILE;
constexpr int FB_RG_WORDS = NJ * 32 * 9;              // compacted rotations of one body group
constexpr int FB_PROD_WARPS = 8;
constexpr int FB_PROD_WARP0 = 4;
constexpr int FB_THREADS = (FB_PROD_WARP0 + FB_PROD_WARPS) * 32;
constexpr int FB_REGS_WG0 = 56;                       // setmaxnreg: warpgroup 0 (TMA, MMA, two idle warps) gives registers up
constexpr int FB_REGS_PROD = 224;
constexpr int FB_HV = 8;                              // vertices per dV staging tile (half an item)
constexpr int FB_TILE_WORDS = ItemShape<FB_HV>::TILE_WORDS;
constexpr int FB_PLAN_WORDS = 40;                     // plan record of 8 vertices
constexpr int FB_NBARS = 2 * FB_NSTAGE_A + 3 * FB_NSTAGE_B + 1 + 2 * FB_PROD_WARPS;
constexpr size_t FB_SMEM = (size_t)4 * FB_RG_WORDS * 4 + (size_t)FB_NSTAGE_B * FB_BSTAGE +
                           (size_t)FB_PROD_WARPS * (FB_TILE_WORDS + 2 * FB_PLAN_WORDS) * 4 + FB_NBARS * 8 + 16 + 1024;
static_assert(FB_SMEM <= 232448, "fused backward: shared memory budget");
static_assert(FB_A_COL0 + FB_NSTAGE_A * FB_A_STAGE_COLS <= 512, "fused backward: tensor memory budget");
static_assert(FB_REGS_WG0 + 2 * FB_REGS_PROD <= 512, "register file");

struct FbWork {
  uint16_t bp[BS_MAX_WORK], i0[BS_MAX_WORK], i1[BS_MAX_WORK], slot[BS_MAX_WORK];
};

// D[tmem] (+)= A[tmem] . B[smem]^T on the CTA pair
__device__ __forceinline__ void umma_bf16_2cta_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x4u(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// compacted rotations of one group in shared memory: [24][32] float4 (r00 r10 r01 r11) | [24][32] float4
// (r02 r12 r20 r21) | [24][32] float r22
template <bool HI>
__device__ __forceinline__ void load_rot_half_c(SlotPair& P, float& r22, const float* R_g, int joint, int lane) {
  const float4 q0 = reinterpret_cast<const float4*>(R_g)[joint * 32 + lane];
  const float4 q1 = reinterpret_cast<const float4*>(R_g + NJ * 32 * 4)[joint * 32 + lane];
  r22 = R_g[NJ * 32 * 8 + joint * 32 + lane];
  P.R[0] = set_half<HI>(P.R[0], q0.x); P.R[3] = set_half<HI>(P.R[3], q0.y);
  P.R[1] = set_half<HI>(P.R[1], q0.z); P.R[4] = set_half<HI>(P.R[4], q0.w);
  P.R[2] = set_half<HI>(P.R[2], q1.x); P.R[5] = set_half<HI>(P.R[5], q1.y);
  P.R[6] = set_half<HI>(P.R[6], q1.z); P.R[7] = set_half<HI>(P.R[7], q1.w);
}

// 4 vertices (the arithmetic of lbs.cu::skin_bwd4 on registers): P = v_posed, G = dV in / dv_posed out
// z22: r22 of the four slots as plain scalars.  (Kept out of the packed pairs: with r22 as the ninth pair of SlotPair,
// updated one half at a time from a scalar shared-memory load, the z rows of dv_posed came out wrong after a reload
// of one slot of a pair -- nvcc 12.9 used the pair's registers as the 64-bit address of the flush REDs.)
__device__ __forceinline__ void skin_bwd4_r(BwdState& s, float (&z22)[4], const float* R_g, float* __restrict__ dA_g, int lane,
                                            const uint32_t* meta_s, const float4* wts_s, uint32_t force,
                                            const float (&P)[12], float (&G)[12]) {
  const uint4 m4 = *reinterpret_cast<const uint4*>(meta_s);
  const uint32_t mts[4] = {m4.x | force, m4.y, m4.z, m4.w};
  float4 w = wts_s[0];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t mt = mts[i];
    const float4 wn = wts_s[i < 3 ? i + 1 : 3];            // next vertex's weights, one vertex ahead
    if (mt & (0xFu << 20)) {
      const uint32_t pv = s.prev;
      if (mt & (1u << 20)) { flush_half<false>(s.A, dA_g, pv & 31, lane); load_rot_half_c<false>(s.A, z22[0], R_g, mt & 31, lane); }
      if (mt & (1u << 21)) { flush_half<true>(s.A, dA_g, (pv >> 5) & 31, lane); load_rot_half_c<true>(s.A, z22[1], R_g, (mt >> 5) & 31, lane); }
      if (mt & (1u << 22)) { flush_half<false>(s.B, dA_g, (pv >> 10) & 31, lane); load_rot_half_c<false>(s.B, z22[2], R_g, (mt >> 10) & 31, lane); }
      if (mt & (1u << 23)) { flush_half<true>(s.B, dA_g, (pv >> 15) & 31, lane); load_rot_half_c<true>(s.B, z22[3], R_g, (mt >> 15) & 31, lane); }
    }
    s.prev = mt;
    const f2 px = bc2(P[i * 3]), py = bc2(P[i * 3 + 1]), pz = bc2(P[i * 3 + 2]);
    const float gx = G[i * 3], gy = G[i * 3 + 1], gz = G[i * 3 + 2];
    s.sx += gx; s.sy += gy; s.sz += gz;
    const f2 wA = mk2(w.x, w.y), wB = mk2(w.z, w.w);
    const f2 hxA = mul2(wA, bc2(gx)), hyA = mul2(wA, bc2(gy)), hzA = mul2(wA, bc2(gz));
    const f2 hxB = mul2(wB, bc2(gx)), hyB = mul2(wB, bc2(gy)), hzB = mul2(wB, bc2(gz));
    f2 qx = fma2(s.A.R[0], hxA, fma2(s.A.R[3], hyA, mul2(s.A.R[6], hzA)));
    f2 qy = fma2(s.A.R[1], hxA, fma2(s.A.R[4], hyA, mul2(s.A.R[7], hzA)));
    f2 qz = fma2(s.A.R[2], hxA, fma2(s.A.R[5], hyA, mul2(mk2(z22[0], z22[1]), hzA)));
    qx = fma2(s.B.R[0], hxB, fma2(s.B.R[3], hyB, fma2(s.B.R[6], hzB, qx)));
    qy = fma2(s.B.R[1], hxB, fma2(s.B.R[4], hyB, fma2(s.B.R[7], hzB, qy)));
    qz = fma2(s.B.R[2], hxB, fma2(s.B.R[5], hyB, fma2(mk2(z22[2], z22[3]), hzB, qz)));
    G[i * 3] = lo2(qx) + hi2(qx); G[i * 3 + 1] = lo2(qy) + hi2(qy); G[i * 3 + 2] = lo2(qz) + hi2(qz);
#define B200_ACC(P_, hx_, hy_, hz_)                                                                  \
  P_.D[0] = fma2(hx_, px, P_.D[0]); P_.D[1] = fma2(hx_, py, P_.D[1]); P_.D[2] = fma2(hx_, pz, P_.D[2]);   \
  P_.D[3] = add2(P_.D[3], hx_);                                                                      \
  P_.D[4] = fma2(hy_, px, P_.D[4]); P_.D[5] = fma2(hy_, py, P_.D[5]); P_.D[6] = fma2(hy_, pz, P_.D[6]);   \
  P_.D[7] = add2(P_.D[7], hy_);                                                                      \
  P_.D[8] = fma2(hz_, px, P_.D[8]); P_.D[9] = fma2(hz_, py, P_.D[9]); P_.D[10] = fma2(hz_, pz, P_.D[10]); \
  P_.D[11] = add2(P_.D[11], hz_);
    B200_ACC(s.A, hxA, hyA, hzA)
    B200_ACC(s.B, hxB, hyB, hzB)
#undef B200_ACC
    w = wn;
  }
}

// stage k of a piece of n items -> item: the first half of the range (producer warps h = 0) on even stages, the
// second half (h = 1) on odd stages
__device__ __forceinline__ int fb_item_of_stage(int k, int i0, int n) {
  return i0 + ((k & 1) ? (n + 1) / 2 + (k >> 1) : (k >> 1));
}

template <bool SPLIT3>
__global__ void __launch_bounds__(FB_THREADS, 1)
lbs_bwd_gemm_kernel(const float4* __restrict__ vpB, int nc4, const float4* __restrict__ A_blk, int b0, int nb,
                    int ntiles_n, const float* __restrict__ grad_verts, int V, const uint32_t* __restrict__ vplan,
                    const __nv_bfloat16* __restrict__ Wbi_hi, const __nv_bfloat16* __restrict__ Wbi_lo, int nitems_all,
                    const __grid_constant__ FbWork work, float* __restrict__ dA_acc, float* __restrict__ dtr_acc,
                    float* __restrict__ D, long long part_stride) {
  constexpr uint32_t TMEM_COLS = 512;
  using SH = ItemShape<FB_HV>;
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);   // stays a shared-space pointer
  unsigned char* b_s = smem;                                           // [FB_NSTAGE_B][hi | lo][6][112][8] bf16
  float* R_s = reinterpret_cast<float*>(b_s + FB_NSTAGE_B * FB_BSTAGE);   // [4 groups] compacted rotations
  float* tiles = R_s + 4 * FB_RG_WORDS;                                // [8][32][HROW] dV staging
  uint32_t* stashes = reinterpret_cast<uint32_t*>(tiles + FB_PROD_WARPS * FB_TILE_WORDS);   // [8][2][40] plan records
  uint64_t* bars = reinterpret_cast<uint64_t*>(stashes + FB_PROD_WARPS * 2 * FB_PLAN_WORDS);
  uint64_t* a_full = bars;                                    // [6]  (the leader's: 8 producer warps of the pair)
  uint64_t* a_empty = a_full + FB_NSTAGE_A;                   // [6]
  uint64_t* b_full = a_empty + FB_NSTAGE_A;                   // [3]
  uint64_t* peer_b_full = b_full + FB_NSTAGE_B;               // [3]
  uint64_t* b_empty = peer_b_full + FB_NSTAGE_B;              // [3]
  uint64_t* tmem_full = b_empty + FB_NSTAGE_B;
  uint64_t* wbars = tmem_full + 1;                            // [8][2] plan record landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbars + 2 * FB_PROD_WARPS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int piece = (int)(blockIdx.x >> 1);
  const int btile = (int)work.bp[piece] * 2 + (int)rank;              // this CTA's 128 bodies
  const int i0 = work.i0[piece], nit = (int)work.i1[piece] - i0;      // items of this cluster

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < FB_NSTAGE_A; ++i) {
      mbar_init(&a_full[i], 8);                     // one arrival per producer warp pair-wide and stage
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < FB_NSTAGE_B; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&peer_b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    mbar_init(tmem_full, 1);
    for (int i = 0; i < 2 * FB_PROD_WARPS; ++i) mbar_init(&wbars[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < FB_PROD_WARP0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FB_REGS_WG0));
    if (warp == 0) {
      // ===== TMA producer (both CTAs): this CTA's half of the Wb rows of every item, in stage order =====
      const uint64_t pol_keep = l2_policy_evict_last();     // re-read by every body pair
      int sb = 0;
      uint32_t phase = 0;
      for (int k = 0; k < nit; ++k) {
        const int item = fb_item_of_stage(k, i0, nit);
        mbar_wait(&b_empty[sb], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&b_full[sb], SPLIT3 ? FB_BSTAGE : FB_BTILE);
          const size_t off = ((size_t)rank * nitems_all + item) * (FB_BTILE / 2);
          bulk_g2s_hint(b_s + sb * FB_BSTAGE, Wbi_hi + off, FB_BTILE, &b_full[sb], pol_keep);
          if (SPLIT3) bulk_g2s_hint(b_s + sb * FB_BSTAGE + FB_BTILE, Wbi_lo + off, FB_BTILE, &b_full[sb], pol_keep);
        }
        __syncwarp();
        if (++sb == FB_NSTAGE_B) { sb = 0; phase ^= 1; }
      }
    } else if (warp == 1) {
      int sb = 0;
      uint32_t bphase = 0;
      if (rank == 0) {
        // ===== leader: MMA issuer of the pair =====
        constexpr uint32_t idesc = make_idesc(2 * BM, FB_BN);
        const uint64_t db_base = make_nosw_desc(smem_u32(b_s), FB_BCHUNK, 128);
        for (int k = 0; k < nit; ++k) {
          const int sa = k % FB_NSTAGE_A;
          mbar_wait(&a_full[sa], (uint32_t)(k / FB_NSTAGE_A) & 1u);    // both CTAs' producers have written the stage
          mbar_wait(&b_full[sb], bphase);
          mbar_wait(&peer_b_full[sb], bphase);
          tc_fence_after();
          const uint32_t a_hi = tmem_base + (uint32_t)(FB_A_COL0 + sa * FB_A_STAGE_COLS);
          const uint32_t a_lo = a_hi + FB_A_STAGE_COLS / 2;
          const uint64_t b_hi = db_base + (uint64_t)((sb * FB_BSTAGE) >> 4), b_lo = b_hi + (uint64_t)(FB_BTILE >> 4);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < FB_ITEM_ROWS / UMMA_K; ++ks) {
              // one K step = 16 rows: 8 TMEM columns of packed pairs, two 8-row chunks of the B tile
              const uint64_t bo = (uint64_t)((ks * 2 * FB_BCHUNK) >> 4);
              umma_bf16_2cta_ts(tmem_base, a_hi + ks * 8, b_hi + bo, idesc, (k | ks) != 0);
              if (SPLIT3) {
                umma_bf16_2cta_ts(tmem_base, a_lo + ks * 8, b_hi + bo, idesc, 1u);
                umma_bf16_2cta_ts(tmem_base, a_hi + ks * 8, b_lo + bo, idesc, 1u);
              }
            }
            umma_commit_2cta(&a_empty[sa]);
            umma_commit_2cta(&b_empty[sb]);
            if (k == nit - 1) umma_commit_2cta(tmem_full);
          }
          __syncwarp();
          if (++sb == FB_NSTAGE_B) { sb = 0; bphase ^= 1; }
        }
      } else {
        // ===== peer: tell the leader when this CTA's Wb stage has landed =====
        for (int k = 0; k < nit; ++k) {
          mbar_wait(&b_full[sb], bphase);
          if (elect_one()) remote_arrive(&peer_b_full[sb], 0);
          __syncwarp();
          if (++sb == FB_NSTAGE_B) { sb = 0; bphase ^= 1; }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FB_REGS_PROD));
    // ===== producers: warp (quadrant q = body group, half h of the item range) =====
    // the thread's coordinates are pinned in registers: left to itself the compiler re-derives them from S2R
    // (threadIdx.x) all over the loop body to save registers, 12 % of the producers' stall samples
    int lane_p = lane, q_p = warp & 3, h_p = (warp - FB_PROD_WARP0) >> 2;
    asm volatile("" : "+r"(lane_p), "+r"(q_p), "+r"(h_p));
    const int lane = lane_p, q = q_p, h = h_p, e = h * 4 + q;
    const bool live = btile < ntiles_n;
    const int g = btile * 4 + q;                                       // slab-local body group
    float* R_g = R_s + q * FB_RG_WORDS;
    float* tile = tiles + e * FB_TILE_WORDS;
    float* my_row = tile + lane * SH::HROW;
    uint32_t* stash = stashes + e * 2 * FB_PLAN_WORDS;
    uint64_t* wbar = wbars + e * 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    // ---- the group's rotations, compacted, -> shared memory (the two warps of the quadrant take 12 joints each) ----
    if (live) {
      const float4* ap = A_blk + (size_t)g * (NJ * 3 * 32) + lane;
#pragma unroll 1
      for (int j = h * (NJ / 2); j < (h + 1) * (NJ / 2); ++j) {
        const float4 x0 = ap[(j * 3) * 32], x1 = ap[(j * 3 + 1) * 32], x2 = ap[(j * 3 + 2) * 32];
        reinterpret_cast<float4*>(R_g)[j * 32 + lane] = x0;
        reinterpret_cast<float4*>(R_g + NJ * 32 * 4)[j * 32 + lane] = make_float4(x1.x, x1.y, x2.x, x2.y);
        R_g[NJ * 32 * 8 + j * 32 + lane] = x2.z;
      }
    }
    asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(64) : "memory");     // both warps of the quadrant have staged their joints
    const int n_first = (nit + 1) / 2;
    const int my_n = h == 0 ? n_first : nit - n_first;                 // items of this warp
    const int it_first = i0 + (h == 0 ? 0 : n_first);
    const int nrows = max(0, min(32, nb - g * 32));
    const size_t row_stride = (size_t)V * 3;
    if (live && my_n > 0) {
      float* dA_g = dA_acc + (size_t)g * AG_WORDS;
      const float4* vp_g = vpB + (size_t)g * nc4 * 32 + lane;
      const float* dv_g = grad_verts + (size_t)(b0 + g * 32) * V * 3;
      auto issue_half = [&](uint32_t b, int t8) {           // plan record (lane 0, bulk copy) of half item t8
        if (lane == 0) {
          mbar_expect_tx(&wbar[b], FB_PLAN_WORDS * 4);
          bulk_g2s(stash + b * FB_PLAN_WORDS, vplan + (size_t)t8 * FB_PLAN_WORDS, FB_PLAN_WORDS * 4, &wbar[b]);
        }
      };
      auto issue_dv = [&](int t8) {                         // the 8 vertices' dV rows of the 32 bodies -> staging tile
        issue_rows<FB_HV>(tile, dv_g + (size_t)t8 * FB_HV * 3, row_stride, nrows, max(0, min(FB_HV, V - t8 * FB_HV)) * 3, lane);
        cp_async_commit();
      };
      auto load_vp = [&](float (&Pn)[24], int t8) {         // v_posed of half item t8: 6 coalesced float4
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const float4 x = ld_stream4(vp_g + (size_t)(t8 * 6 + i) * 32);
          Pn[4 * i] = x.x; Pn[4 * i + 1] = x.y; Pn[4 * i + 2] = x.z; Pn[4 * i + 3] = x.w;
        }
      };
      BwdState st;
#pragma unroll
      for (int i = 0; i < AELEMS; ++i) st.A.D[i] = st.B.D[i] = mk2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 9; ++i) st.A.R[i] = st.B.R[i] = mk2(0.f, 0.f);
      st.prev = 0u;
      st.sx = st.sy = st.sz = 0.f;
      float z22[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t n_half = 0;                                   // half items pulled so far (stash parity)
      float Pn[24];
      const int t8_first = it_first * 2, t8_end = (it_first + my_n) * 2;
      issue_half(0u, t8_first);
      issue_dv(t8_first);
      load_vp(Pn, t8_first);
#pragma unroll 1
      for (int m = 0; m < my_n; ++m) {
        const int k = 2 * m + h, sa = k % FB_NSTAGE_A;
        mbar_wait(&a_empty[sa], ((uint32_t)(k / FB_NSTAGE_A) & 1u) ^ 1u);   // the MMAs that read this stage have retired
        tc_fence_after();
        const uint32_t a_col = lane_base + (uint32_t)(FB_A_COL0 + sa * FB_A_STAGE_COLS);
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh, ++n_half) {
          const int t8 = (it_first + m) * 2 + hh;
          const uint32_t b = n_half & 1;
          if (t8 + 1 < t8_end) issue_half(b ^ 1, t8 + 1);
          float P[24], G[24];
#pragma unroll
          for (int i = 0; i < 24; ++i) P[i] = Pn[i];
          if (t8 + 1 < t8_end) load_vp(Pn, t8 + 1);
          cp_async_wait<0>();
          __syncwarp();                                                // the tile's rows were copied by other lanes
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const float4 x = *reinterpret_cast<const float4*>(my_row + i * 4);
            G[4 * i] = x.x; G[4 * i + 1] = x.y; G[4 * i + 2] = x.z; G[4 * i + 3] = x.w;
          }
          __syncwarp();                                                // every lane has its row: the tile is free
          if (t8 + 1 < t8_end) issue_dv(t8 + 1);
          mbar_wait(&wbar[b], (n_half >> 1) & 1);
          const uint32_t* ps = stash + b * FB_PLAN_WORDS;
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            float Pu[12], Gu[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) { Pu[i] = P[u * 12 + i]; Gu[i] = G[u * 12 + i]; }
            skin_bwd4_r(st, z22, R_g, dA_g, lane, plan_meta(ps, u), plan_wts(ps, u), (m == 0 && hh == 0 && u == 0) ? (0xFu << 20) : 0u,
                        Pu, Gu);
#pragma unroll
            for (int i = 0; i < 12; ++i) G[u * 12 + i] = Gu[i];
          }
          // 24 gradient rows of this body -> bf16 hi / lo pairs -> this half's 12 + 12 columns of the stage
          uint32_t hi[12], lo[12];
#pragma unroll
          for (int c = 0; c < 12; ++c) split2(G[2 * c], G[2 * c + 1], hi[c], lo[c]);
#pragma unroll
          for (int c = 0; c < 12; c += 4) {
            tmem_st_x4u(a_col + hh * 12 + c, hi[c], hi[c + 1], hi[c + 2], hi[c + 3]);
            if (SPLIT3) tmem_st_x4u(a_col + FB_A_STAGE_COLS / 2 + hh * 12 + c, lo[c], lo[c + 1], lo[c + 2], lo[c + 3]);
          }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&a_full[sa])) : "memory");
          else remote_arrive(&a_full[sa], 0);
        }
      }
      bwd_close_group(st, dA_g, dtr_acc + (size_t)g * 96, lane);
    } else {
      // a CTA without bodies (odd number of body tiles) still owes the pair its arrivals
      for (int m = 0; m < my_n; ++m) {
        const int k = 2 * m + h, sa = k % FB_NSTAGE_A;
        mbar_wait(&a_empty[sa], ((uint32_t)(k / FB_NSTAGE_A) & 1u) ^ 1u);
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&a_full[sa])) : "memory");
          else remote_arrive(&a_full[sa], 0);
        }
      }
    }
    // ===== epilogue: this warp stores its half of the lane quadrant's accumulator rows (112 features) =====
    if (nit > 0) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
      if (live) {
        float* drow = D + (long long)work.slot[piece] * part_stride + (long long)(btile * BM + q * 32 + lane) * FB_BN + h * FB_BNH;
#pragma unroll 1
        for (int c0 = 0; c0 < FB_BNH; c0 += 16) {
          uint32_t v[16];
          tmem_ld_x16(lane_base + (uint32_t)(h * FB_BNH + c0), v);
#pragma unroll
          for (int i = 0; i < 16; i += 4) *reinterpret_cast<uint4*>(drow + c0 + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- host side ------------------------------------------------------------------------------
static int fb_build_work(const DevModel& m, int Sw, int num_sms, FbWork& work, int& npieces) {
  BsWork bs;
  const int bp_total = (Sw / BM + 1) / 2;
  npieces = build_bs_work(bp_total, m.ntiles * (TILE_V / FB_ITEM_V), std::max(1, num_sms / 2), bs);
  if (npieces < 0) return -1;
  int nslots = 0;
  std::vector<int> used(bp_total, 0);
  memset(&work, 0, sizeof(work));
  for (int c = 0; c < npieces; ++c) {
    work.bp[c] = bs.bp[c]; work.i0[c] = bs.w0[c]; work.i1[c] = bs.w1[c];
    work.slot[c] = (uint16_t)used[bs.bp[c]]++;
    nslots = std::max(nslots, (int)work.slot[c] + 1);
  }
  return nslots;
}

bool fused_bwd_usable(const DevModel& m, int mode, const float* grad_verts) {
  // opt-in: correct (parity tests run it in a child process) but slower than the two kernels today, see the header
  static const bool on = getenv("B200_FUSED_BWD") != nullptr && atoi(getenv("B200_FUSED_BWD")) != 0;
  return on && mode != B200SMPL_MODE_FP32_SIMT && m.fl.nf_pad == FB_BN && (m.V & 1) == 0 &&
         (reinterpret_cast<uintptr_t>(grad_verts) & 7) == 0 && m.n_virt0 == m.ntiles * 96;
}

int fused_bwd_parts(const DevModel& m, int Sw, int num_sms) {
  FbWork work;
  int npieces = 0;
  return fb_build_work(m, Sw, num_sms, work, npieces);
}

// dfeat_part: [fused_bwd_parts][S][nf_pad], zeroed by the caller (a body pair uses as many parts as pieces cut it)
int launch_lbs_bwd_gemm(const DevModel& m, int mode, const float* vpB, int S, int Sw, const float* A_blk, int b0, int nb,
                        const float* grad_verts, float* dA_acc, float* dtr_acc, float* dfeat_part, int num_sms,
                        cudaStream_t st) {
  FbWork work;
  int npieces = 0;
  const int nslots = fb_build_work(m, Sw, num_sms, work, npieces);
  if (nslots < 0) return fail(B200SMPL_ERR_INVALID, "slab too wide for the fused backward work list");
  const bool split3 = mode != B200SMPL_MODE_BF16;
  auto kern = split3 ? lbs_bwd_gemm_kernel<true> : lbs_bwd_gemm_kernel<false>;
  B200_SMEM_ATTR_ONCE(lbs_bwd_gemm_kernel<true>, FB_SMEM);
  B200_SMEM_ATTR_ONCE(lbs_bwd_gemm_kernel<false>, FB_SMEM);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * npieces, 1, 1);
  cfg.blockDim = dim3(FB_THREADS, 1, 1);
  cfg.dynamicSmemBytes = FB_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LaunchTimer _timer("lbs_bwd_gemm", st);
  B200_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, reinterpret_cast<const float4*>(vpB), m.n_pad / 4,
                                   reinterpret_cast<const float4*>(A_blk), b0, nb, Sw / BM, grad_verts, m.V, m.vplan, m.Wbi_hi,
                                   m.Wbi_lo, m.ntiles * (TILE_V / FB_ITEM_V), work, dA_acc, dtr_acc, dfeat_part,
                                   (long long)S * FB_BN));
  B200_LAUNCH_CHECK("lbs_bwd_gemm");
  return 0;
}

}  // namespace b200smpl

