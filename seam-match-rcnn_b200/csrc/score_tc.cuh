// K2/K3: pair scorer as a tcgen05 GEMM with a fused per-query top-k candidate filter.
//
// Reference arithmetic (models/match_head.py:160-162, evaluate_movingfashion.py:263-268):
//   x5 = last((q - g)^2),  score = softmax(x5)[1] = sigmoid(l1 - l0).
// Ranking needs only d_ij = l1 - l0 = dw.(q_i - g_j)^2 + db with dw = w1 - w0, which expands to
//   d_ij = rq_i + [ a_i . g_j + cg_j ] + db,   a_i = -2 dw (.) q_i,  rq_i = dw.q_i^2,  cg_j = dw.g_j^2.
// The bracket is what this kernel evaluates: a_i . g_j on the tensor cores (fp16 operands,
// fp32 accumulation in TMEM), + cg_j in the epilogue.  The (Q,G) matrix never leaves the SM:
// each epilogue thread owns one query row, compares its accumulator columns against the
// row's running threshold and appends survivors {value, gallery row} to a per-row buffer in
// shared memory.  When a buffer runs full every thread of the warp prunes ITS OWN row in
// parallel (sampled-pivot partition, no cross-lane traffic), keeping the best 32..44 entries
// and raising the threshold.  Thresholds are shared between CTAs working on the same rows
// through global memory (atomicMax on an order-preserving integer image of the float).
//
// Work decomposition: the (query tile, gallery tile) grid is linearised query-major and cut
// into one contiguous, equally long range per CTA; a range is processed as at most a few
// "segments" (one query tile x a run of gallery tiles).  At the end of a segment each thread
// appends its row's surviving candidates to that row's list in global memory.
//
// Roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one lane), warp 2 = TMEM
// allocator, warps 4..7 = epilogue (warp%4 selects the TMEM lane quarter).
// Tile: 128 queries x 256 gallery rows, K = 256 as 4 k-blocks of 64 fp16 (128-byte swizzle).
// The A (query) tile stays resident in shared memory for a whole segment; B (gallery)
// k-blocks stream through a 3-stage ring; two 256-column TMEM accumulators alternate so the
// epilogue of tile n overlaps the MMAs of tile n+1.
#pragma once
#include <cstdint>
#include <cuda.h>
#include "sm100_ptx.cuh"

namespace seam {
namespace score {

constexpr int BM = 128, BN = 256, BK = 64, NKB = 4, NSTAGE = 3;
constexpr int CAP = 60;           // per-row buffer slots ({fp32 value, int32 gallery row} = 8 B)
constexpr int CHUNK = 16;         // accumulator columns per tcgen05.ld
constexpr int KEEP_LO = 32;       // a prune keeps between KEEP_LO ...
constexpr int KEEP_HI = CAP - CHUNK;   // ... and KEEP_HI entries (room for one more chunk)
constexpr int THREADS = 256;
constexpr uint32_t A_KB_BYTES = BM * BK * 2;
constexpr uint32_t B_ST_BYTES = BN * BK * 2;
constexpr uint32_t SLOT_STRIDE = BM * 8;   // bytes between consecutive slots of one row

constexpr uint32_t OFF_A = 0;
constexpr uint32_t OFF_B = OFF_A + NKB * A_KB_BYTES;
constexpr uint32_t OFF_CAND = OFF_B + NSTAGE * B_ST_BYTES;
constexpr uint32_t OFF_CG = OFF_CAND + CAP * SLOT_STRIDE;
constexpr uint32_t OFF_BAR = OFF_CG + 2 * BN * 4;
constexpr uint32_t NUM_BARS = 2 * NSTAGE + 2 + 4;
constexpr uint32_t OFF_TMEM = OFF_BAR + NUM_BARS * 8;
constexpr uint32_t SMEM_BYTES = OFF_TMEM + 16 + 1024;   // + slack for manual 1024-byte alignment
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
static_assert(OFF_CAND % 16 == 0 && OFF_CG % 16 == 0 && OFF_BAR % 8 == 0, "alignment");
static_assert(KEEP_HI >= KEEP_LO + 8, "prune window too narrow");

struct Params {
  int Q, G, num_mtiles, ntiles_n;
  long long total_tiles;
  int RB;                     // capacity of a row's global candidate list
  int debug_mode;             // 0 = normal; 1 = drain only; 2 = load TMEM, no filter; 3 = filter, never append
  const float* cg;            // (G)
  uint32_t* thr_global;       // (Q) ordered-uint image of the per-row lower bound
  uint32_t* rowcnt;           // (Q) entries appended to rowbuf so far
  uint32_t* rowflag;          // (Q) nonzero: the row lost candidates, must be ranked exhaustively
  uint2* rowbuf;              // (Q, RB) {approximate a.g + cg, shard-local gallery row}
};

// contiguous tile range of CTA b out of nb
__device__ __forceinline__ void cta_range(long long total, int nb, int b, long long& t0, long long& t1) {
  t0 = total * b / nb;
  t1 = total * (b + 1) / nb;
}

__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

#define SEAM_CE(a, b)               \
  {                                 \
    const float hi_ = fmaxf(a, b);  \
    const float lo_ = fminf(a, b);  \
    a = hi_;                        \
    b = lo_;                        \
  }

// Thread-local prune of one row's buffer (slots [0,cnt) at base + s*SLOT_STRIDE).
// Picks a pivot from 8 sorted samples such that between KEEP_LO and KEEP_HI entries exceed it,
// compacts those to the front, returns the new count and raises thr to the pivot.
// Returns false when no pivot works (massive ties): the row is then compacted lossily and must
// be flagged for the exhaustive path by the caller.
__device__ __forceinline__ bool prune_row_local(uint32_t base, int& cnt, float& thr) {
  float s0, s1, s2, s3, s4, s5, s6, s7;
  s0 = lds32(base + ((0 * cnt) >> 3) * SLOT_STRIDE);
  s1 = lds32(base + ((1 * cnt) >> 3) * SLOT_STRIDE);
  s2 = lds32(base + ((2 * cnt) >> 3) * SLOT_STRIDE);
  s3 = lds32(base + ((3 * cnt) >> 3) * SLOT_STRIDE);
  s4 = lds32(base + ((4 * cnt) >> 3) * SLOT_STRIDE);
  s5 = lds32(base + ((5 * cnt) >> 3) * SLOT_STRIDE);
  s6 = lds32(base + ((6 * cnt) >> 3) * SLOT_STRIDE);
  s7 = lds32(base + ((7 * cnt) >> 3) * SLOT_STRIDE);
  // 19-comparator sorting network, descending
  SEAM_CE(s0, s1) SEAM_CE(s2, s3) SEAM_CE(s4, s5) SEAM_CE(s6, s7)
  SEAM_CE(s0, s2) SEAM_CE(s1, s3) SEAM_CE(s4, s6) SEAM_CE(s5, s7)
  SEAM_CE(s1, s2) SEAM_CE(s5, s6) SEAM_CE(s0, s4) SEAM_CE(s3, s7)
  SEAM_CE(s1, s5) SEAM_CE(s2, s6)
  SEAM_CE(s1, s4) SEAM_CE(s3, s6)
  SEAM_CE(s2, s4) SEAM_CE(s3, s5)
  SEAM_CE(s3, s4)
  // counts above four candidate pivots, and the row maximum, in one pass
  int c4 = 0, c5 = 0, c6 = 0, c7 = 0;
  float vmax = -INFINITY;
  for (int s = 0; s < cnt; ++s) {
    const float v = lds32(base + s * SLOT_STRIDE);
    c4 += v > s4;
    c5 += v > s5;
    c6 += v > s6;
    c7 += v > s7;
    vmax = fmaxf(vmax, v);
  }
  // bracket the window [KEEP_LO, KEEP_HI]: pl has too many entries above it, ph too few
  float pl = -INFINITY, ph = vmax, pivot = vmax;
  bool found = false;
#define SEAM_TRY(P, C)                                   \
  if (!found) {                                          \
    if ((C) > KEEP_HI) pl = fmaxf(pl, (P));              \
    else if ((C) >= KEEP_LO) { pivot = (P); found = true; } \
    else ph = fminf(ph, (P));                            \
  }
  SEAM_TRY(s7, c7) SEAM_TRY(s6, c6) SEAM_TRY(s5, c5) SEAM_TRY(s4, c4)
#undef SEAM_TRY
  // no sample landed in the window: bisect between the bracketing values
  for (int iter = 0; iter < 12 && !found; ++iter) {
    const float mid = pl == -INFINITY ? ph - fmaxf(1e-3f, fabsf(ph) * 1e-3f) * (float)(1 << iter) : 0.5f * (pl + ph);
    if (!(mid > pl) || !(mid < ph)) break;               // no representable value in between (ties)
    int c = 0;
    for (int s = 0; s < cnt; ++s) c += lds32(base + s * SLOT_STRIDE) > mid;
    if (c > KEEP_HI) pl = mid;
    else if (c >= KEEP_LO) { pivot = mid; found = true; }
    else ph = mid;
  }
  bool ok = found;
  if (!found) pivot = ph;                                // ties: keep fewer than KEEP_LO, row becomes lossy
  int w = 0;
  for (int s = 0; s < cnt; ++s) {
    const uint2 e = lds64(base + s * SLOT_STRIDE);
    if (__uint_as_float(e.x) > pivot) {
      if (w < KEEP_HI) sts64(base + w * SLOT_STRIDE, e.x, e.y);
      ++w;
    }
  }
  if (w > KEEP_HI) {   // could not make room without dropping entries above the pivot
    w = KEEP_HI;
    ok = false;
  }
  cnt = w;
  thr = fmaxf(thr, pivot);
  return ok;
}

__global__ void __launch_bounds__(THREADS, 1)
score_topk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint8_t* sA = smem + OFF_A;
  uint8_t* sB = smem + OFF_B;
  float* cg_s = reinterpret_cast<float*>(smem + OFF_CG);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* full = bars;                    // [NSTAGE]  TMA -> MMA
  uint64_t* empty = bars + NSTAGE;          // [NSTAGE]  MMA -> TMA
  uint64_t* a_full = bars + 2 * NSTAGE;     // A tile landed
  uint64_t* a_empty = a_full + 1;           // all MMAs of the segment retired
  uint64_t* t_full = a_empty + 1;           // [2] accumulator ready
  uint64_t* t_empty = t_full + 2;           // [2] accumulator drained
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    for (int i = 0; i < NSTAGE; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    ptx::mbar_init(a_full, 1);
    ptx::mbar_init(a_empty, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&t_full[i], 1);
      ptx::mbar_init(&t_empty[i], 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr_s, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  long long t_begin, t_end;
  cta_range(p.total_tiles, gridDim.x, blockIdx.x, t_begin, t_end);

  if (warp == 0) {
    // ================================================================= TMA producer
    if (lane == 0) {
      uint32_t stage = 0, sphase = 0, iphase = 0;
      long long t = t_begin;
      while (t < t_end) {
        const int m = (int)(t / p.ntiles_n);
        const int nt0 = (int)(t - (long long)m * p.ntiles_n);
        const long long seg_end = min(t_end, (long long)(m + 1) * p.ntiles_n);
        const int nt1 = nt0 + (int)(seg_end - t);
        ptx::mbar_wait(a_empty, iphase ^ 1);
        ptx::mbar_arrive_expect_tx(a_full, NKB * A_KB_BYTES);
        for (int kb = 0; kb < NKB; ++kb) ptx::tma_load_2d(sA + kb * A_KB_BYTES, &tmA, a_full, kb * BK, m * BM);
        for (int nt = nt0; nt < nt1; ++nt) {
          for (int kb = 0; kb < NKB; ++kb) {
            ptx::mbar_wait(&empty[stage], sphase ^ 1);
            ptx::mbar_arrive_expect_tx(&full[stage], B_ST_BYTES);
            ptx::tma_load_2d(sB + stage * B_ST_BYTES, &tmB, &full[stage], kb * BK, nt * BN);
            if (++stage == NSTAGE) {
              stage = 0;
              sphase ^= 1;
            }
          }
        }
        iphase ^= 1;
        t = seg_end;
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc(0 /*fp16*/, BM, BN);
      const uint32_t a_addr = ptx::smem_u32(sA), b_addr = ptx::smem_u32(sB);
      uint32_t stage = 0, sphase = 0, iphase = 0, acc = 0, aphase = 0;
      long long t = t_begin;
      while (t < t_end) {
        const int m = (int)(t / p.ntiles_n);
        const long long seg_end = min(t_end, (long long)(m + 1) * p.ntiles_n);
        const int ntiles = (int)(seg_end - t);
        ptx::mbar_wait(a_full, iphase);
        for (int it = 0; it < ntiles; ++it) {
          ptx::mbar_wait(&t_empty[acc], aphase ^ 1);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          for (int kb = 0; kb < NKB; ++kb) {
            ptx::mbar_wait(&full[stage], sphase);
            ptx::tc_fence_after();
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t ad = ptx::umma_desc_k_sw128(a_addr + kb * A_KB_BYTES + k * 32);
              const uint64_t bd = ptx::umma_desc_k_sw128(b_addr + stage * B_ST_BYTES + k * 32);
              ptx::umma_f16(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            ptx::umma_commit(&empty[stage]);
            if (++stage == NSTAGE) {
              stage = 0;
              sphase ^= 1;
            }
          }
          ptx::umma_commit(&t_full[acc]);
          if (++acc == 2) {
            acc = 0;
            aphase ^= 1;
          }
        }
        ptx::umma_commit(a_empty);
        iphase ^= 1;
        t = seg_end;
      }
    }
  } else if (warp >= 4) {
    // ================================================================= epilogue
    const int ew = warp - 4;                 // TMEM lane quarter
    const int R = ew * 32 + lane;            // row within the CTA tile
    const int etid = tid - 128;
    const uint32_t cand_base = ptx::smem_u32(smem + OFF_CAND) + R * 8;
    uint32_t acc = 0, aphase = 0;
    long long t = t_begin;
    while (t < t_end) {
      const int m = (int)(t / p.ntiles_n);
      const int nt0 = (int)(t - (long long)m * p.ntiles_n);
      const long long seg_end = min(t_end, (long long)(m + 1) * p.ntiles_n);
      const int nt1 = nt0 + (int)(seg_end - t);
      const int grow = m * BM + R;
      const bool row_ok = grow < p.Q;
      float thr = (row_ok && p.debug_mode != 3) ? -INFINITY : INFINITY;
      uint32_t wp = cand_base;               // address of the next free slot
      bool lossy = false;
      for (int nt = nt0; nt < nt1; ++nt) {
        // stage cg for this tile (-inf beyond G so padded columns never qualify)
        {
          float* dst = cg_s + acc * BN;
          const int j0 = nt * BN + etid, j1 = j0 + 128;
          dst[etid] = j0 < p.G ? __ldg(p.cg + j0) : -INFINITY;
          dst[etid + 128] = j1 < p.G ? __ldg(p.cg + j1) : -INFINITY;
        }
        if (row_ok) thr = fmaxf(thr, ptx::ordered_to_float(__ldcg(p.thr_global + grow)));
        ptx::named_bar_sync(1, 128);
        ptx::mbar_wait(&t_full[acc], aphase);
        ptx::tc_fence_after();
        if (p.debug_mode == 1 || p.debug_mode == 2) {
          if (p.debug_mode == 2) {
            const uint32_t ta = tmem_base + (uint32_t(ew * 32) << 16) + acc * BN;
            uint32_t rr[CHUNK];
            uint32_t accum = 0;
            for (int ch = 0; ch < BN / CHUNK; ++ch) {
              ptx::tmem_ld_x16(ta + ch * CHUNK, rr);
              ptx::tmem_ld_wait_x16(rr);
#pragma unroll
              for (int e = 0; e < CHUNK; ++e) accum ^= rr[e];
            }
            if (accum == 0x12345678u) p.rowflag[0] = 1;
          }
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&t_empty[acc]);
          if (++acc == 2) {
            acc = 0;
            aphase ^= 1;
          }
          continue;
        }
        const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + acc * BN;
        const float* cgt = cg_s + acc * BN;
        const int col0 = nt * BN;
        uint32_t r[CHUNK];
        ptx::tmem_ld_x16(taddr, r);
#pragma unroll 1
        for (int ch = 0; ch < BN / CHUNK; ++ch) {
          ptx::tmem_ld_wait_x16(r);
          float x[CHUNK];
#pragma unroll
          for (int e = 0; e < CHUNK; ++e) x[e] = __uint_as_float(r[e]);
          if (ch + 1 < BN / CHUNK) ptx::tmem_ld_x16(taddr + (ch + 1) * CHUNK, r);   // prefetch next chunk
#pragma unroll
          for (int c4 = 0; c4 < CHUNK / 4; ++c4) {
            const float4 g4 = *reinterpret_cast<const float4*>(cgt + ch * CHUNK + c4 * 4);
            const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float v = x[c4 * 4 + e] + gg[e];
              if (v > thr) {
                sts64(wp, __float_as_uint(v), (uint32_t)(col0 + ch * CHUNK + c4 * 4 + e));
                wp += SLOT_STRIDE;
              }
            }
          }
          const bool need = (wp - cand_base) > (uint32_t)KEEP_HI * SLOT_STRIDE;
          if (__any_sync(ptx::FULL_MASK, need)) {
            int cnt = (int)((wp - cand_base) / SLOT_STRIDE);
            if (cnt > KEEP_HI - 4) {          // rows close to the limit prune together
              const float before = thr;
              if (!prune_row_local(cand_base, cnt, thr)) lossy = true;
              wp = cand_base + cnt * SLOT_STRIDE;
              if (thr > before && !lossy) atomicMax(p.thr_global + grow, ptx::float_to_ordered(thr));
            }
            __syncwarp();
          }
        }
        // accumulator drained: hand it back to the MMA warp
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&t_empty[acc]);
        if (++acc == 2) {
          acc = 0;
          aphase ^= 1;
        }
      }
      // ---- flush: append this row's surviving candidates to its global list
      if (row_ok) {
        const int cnt = (int)((wp - cand_base) / SLOT_STRIDE);
        const float tg = ptx::ordered_to_float(__ldcg(p.thr_global + grow));
        int npass = 0;
        for (int s = 0; s < cnt; ++s) npass += lds32(cand_base + s * SLOT_STRIDE) >= tg;
        if (lossy) atomicOr(p.rowflag + grow, 1u);
        if (npass > 0) {
          uint32_t slot = atomicAdd(p.rowcnt + grow, (uint32_t)npass);
          uint2* dst = p.rowbuf + (size_t)grow * p.RB;
          for (int s = 0; s < cnt; ++s) {
            const uint2 e = lds64(cand_base + s * SLOT_STRIDE);
            if (__uint_as_float(e.x) >= tg) {
              if (slot < (uint32_t)p.RB) dst[slot] = e;
              ++slot;
            }
          }
          if (slot > (uint32_t)p.RB) atomicOr(p.rowflag + grow, 2u);
        }
      }
      __syncwarp();
      t = seg_end;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace score
}  // namespace seam
