// K2/K3: pair scorer as a tcgen05 GEMM with a fused per-query top-k candidate filter.
//
// Reference arithmetic (models/match_head.py:160-162, evaluate_movingfashion.py:263-268):
//   x5 = last((q - g)^2),  score = softmax(x5)[1] = sigmoid(l1 - l0).
// Ranking needs only d_ij = l1 - l0 = dw.(q_i - g_j)^2 + db with dw = w1 - w0, which expands to
//   d_ij = rq_i + [ a_i . g_j + cg_j ] + db,   a_i = -2 dw (.) q_i,  rq_i = dw.q_i^2,  cg_j = dw.g_j^2.
// The bracket is what this kernel evaluates: a_i . g_j on the tensor cores (fp16 operands,
// fp32 accumulation in TMEM), + cg_j in the epilogue.  The (Q,G) matrix never leaves the SM:
// each epilogue thread owns one query row, compares its accumulator columns against the
// row's running threshold and appends survivors to a 64-slot per-row buffer in shared
// memory; a warp-cooperative bitonic pass prunes a full buffer to the best KP=32 and raises
// the threshold.  Thresholds are shared between CTAs working on the same rows through
// global memory (atomicMax on an order-preserving integer image of the float).
//
// Roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one lane), warp 2 = TMEM
// allocator, warps 4..7 = epilogue (warp%4 selects the TMEM lane quarter).
// Tile: 128 queries x 256 gallery rows, K = 256 as 4 k-blocks of 64 fp16 (128-byte swizzle).
// The A (query) tile stays resident in shared memory for a whole work item; B (gallery)
// k-blocks stream through a 3-stage ring; two 256-column TMEM accumulators alternate so the
// epilogue of tile n overlaps the MMAs of tile n+1.
#pragma once
#include <cstdint>
#include <cuda.h>
#include "sm100_ptx.cuh"
#include "warp_sort.cuh"

namespace seam {
namespace score {

constexpr int BM = 128, BN = 256, BK = 64, NKB = 4, NSTAGE = 3;
constexpr int KP = 32;            // candidates kept per (row, part)
constexpr int CAP = 64;           // per-row buffer slots
constexpr int CHUNK = 16;         // accumulator columns per tcgen05.ld
constexpr int PITCH_V = 129;      // floats; odd pitch -> row-wise and slot-wise access conflict-free
constexpr int PITCH_I = 130;      // uint16
constexpr int THREADS = 256;
constexpr uint32_t A_KB_BYTES = BM * BK * 2;
constexpr uint32_t B_ST_BYTES = BN * BK * 2;

constexpr uint32_t OFF_A = 0;
constexpr uint32_t OFF_B = OFF_A + NKB * A_KB_BYTES;
constexpr uint32_t OFF_CV = OFF_B + NSTAGE * B_ST_BYTES;
constexpr uint32_t OFF_CI = OFF_CV + CAP * PITCH_V * 4;
constexpr uint32_t OFF_CG = OFF_CI + CAP * PITCH_I * 2;
constexpr uint32_t OFF_BAR = OFF_CG + 2 * BN * 4;
constexpr uint32_t NUM_BARS = 2 * NSTAGE + 2 + 4;
constexpr uint32_t OFF_TMEM = OFF_BAR + NUM_BARS * 8;
constexpr uint32_t SMEM_BYTES = OFF_TMEM + 16 + 1024;   // + slack for manual 1024-byte alignment
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
static_assert(OFF_CV % 16 == 0 && OFF_CI % 16 == 0 && OFF_CG % 16 == 0 && OFF_BAR % 8 == 0, "alignment");

struct Params {
  int Q, G, P, tiles_per_part, num_mtiles, ntiles_n, num_items;
  const float* cg;            // (G)
  uint32_t* thr_global;       // (Q) ordered-uint image of the per-row lower bound
  float* cand_v;              // (Q, P, KP) approximate a.g + cg, best first, -inf padded
  int32_t* cand_i;            // (Q, P, KP) shard-local gallery row, -1 padded
};

// Prune one row's buffer to its best KP entries (sorted, best first), raise the row's
// threshold, optionally write the list out.  Executed by the whole warp for row `rl`.
__device__ __forceinline__ void prune_row(float* cv, uint16_t* ci, int R, int n, int lane, float& keep_v,
                                          uint32_t& keep_i) {
  float v0 = -INFINITY, v1 = -INFINITY;
  uint32_t i0 = 0xffffu, i1 = 0xffffu;
  if (lane < n) {
    v0 = cv[lane * PITCH_V + R];
    i0 = ci[lane * PITCH_I + R];
  }
  if (lane + 32 < n) {
    v1 = cv[(lane + 32) * PITCH_V + R];
    i1 = ci[(lane + 32) * PITCH_I + R];
  }
  wsort::sort32<true>(v0, i0, lane);
  if (n > 32) {   // warp-uniform
    wsort::sort32<false>(v1, i1, lane);
    if (v1 > v0) {
      v0 = v1;
      i0 = i1;
    }
    wsort::merge32<true>(v0, i0, lane);
  }
  keep_v = v0;
  keep_i = i0;
}

__global__ void __launch_bounds__(THREADS, 1)
score_topk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint8_t* sA = smem + OFF_A;
  uint8_t* sB = smem + OFF_B;
  float* cv = reinterpret_cast<float*>(smem + OFF_CV);
  uint16_t* ci = reinterpret_cast<uint16_t*>(smem + OFF_CI);
  float* cg_s = reinterpret_cast<float*>(smem + OFF_CG);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* full = bars;                    // [NSTAGE]  TMA -> MMA
  uint64_t* empty = bars + NSTAGE;          // [NSTAGE]  MMA -> TMA
  uint64_t* a_full = bars + 2 * NSTAGE;     // A tile landed
  uint64_t* a_empty = a_full + 1;           // all MMAs of the item retired
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

  if (warp == 0) {
    // ================================================================= TMA producer
    if (lane == 0) {
      uint32_t stage = 0, sphase = 0, iphase = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int m = item % p.num_mtiles, part = item / p.num_mtiles;
        const int nt0 = part * p.tiles_per_part;
        const int nt1 = min(nt0 + p.tiles_per_part, p.ntiles_n);
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
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc(0 /*fp16*/, BM, BN);
      const uint32_t a_addr = ptx::smem_u32(sA), b_addr = ptx::smem_u32(sB);
      uint32_t stage = 0, sphase = 0, iphase = 0, acc = 0, aphase = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int part = item / p.num_mtiles;
        const int nt0 = part * p.tiles_per_part;
        const int nt1 = min(nt0 + p.tiles_per_part, p.ntiles_n);
        ptx::mbar_wait(a_full, iphase);
        for (int nt = nt0; nt < nt1; ++nt) {
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
      }
    }
  } else if (warp >= 4) {
    // ================================================================= epilogue
    const int ew = warp - 4;                 // TMEM lane quarter
    const int R = ew * 32 + lane;            // row within the CTA tile
    const int etid = tid - 128;
    uint32_t acc = 0, aphase = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int m = item % p.num_mtiles, part = item / p.num_mtiles;
      const int nt0 = part * p.tiles_per_part;
      const int nt1 = min(nt0 + p.tiles_per_part, p.ntiles_n);
      const int grow = m * BM + R;
      const bool row_ok = grow < p.Q;
      float thr = row_ok ? -INFINITY : INFINITY;
      int cnt = 0;
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
        const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + acc * BN;
        const float* cgt = cg_s + acc * BN;
        const int colbase = (nt - nt0) * BN;
#pragma unroll 1
        for (int ch = 0; ch < BN / CHUNK; ++ch) {
          uint32_t r[CHUNK];
          ptx::tmem_ld_x16(taddr + ch * CHUNK, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c4 = 0; c4 < CHUNK / 4; ++c4) {
            const float4 g4 = *reinterpret_cast<const float4*>(cgt + ch * CHUNK + c4 * 4);
            const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float x = __uint_as_float(r[c4 * 4 + e]) + gg[e];
              if (x > thr) {
                cv[cnt * PITCH_V + R] = x;
                ci[cnt * PITCH_I + R] = (uint16_t)(colbase + ch * CHUNK + c4 * 4 + e);
                ++cnt;
              }
            }
          }
          uint32_t need = __ballot_sync(ptx::FULL_MASK, cnt > CAP - CHUNK);
          if (need) {
            __syncwarp();
            while (need) {
              const int rl = __ffs(need) - 1;
              need &= need - 1;
              const int n = __shfl_sync(ptx::FULL_MASK, cnt, rl);
              const int RR = ew * 32 + rl;
              float kv;
              uint32_t ki;
              prune_row(cv, ci, RR, n, lane, kv, ki);
              cv[lane * PITCH_V + RR] = kv;
              ci[lane * PITCH_I + RR] = (uint16_t)ki;
              const float t32 = __shfl_sync(ptx::FULL_MASK, kv, KP - 1);
              if (lane == rl) {
                cnt = KP;
                if (t32 > thr) {
                  thr = t32;
                  atomicMax(p.thr_global + grow, ptx::float_to_ordered(t32));
                }
              }
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
      // ---- flush: every row's best KP of this part, sorted, to global memory
      __syncwarp();
      for (int rl = 0; rl < 32; ++rl) {
        const int n = __shfl_sync(ptx::FULL_MASK, cnt, rl);
        const int RR = ew * 32 + rl;
        const int gr = m * BM + RR;
        if (gr >= p.Q) break;                  // warp-uniform
        float kv;
        uint32_t ki;
        prune_row(cv, ci, RR, n, lane, kv, ki);
        const size_t o = ((size_t)gr * p.P + part) * KP + lane;
        const bool ok = lane < n;
        p.cand_v[o] = ok ? kv : -INFINITY;
        p.cand_i[o] = ok ? (int32_t)(nt0 * BN + (int)ki) : -1;
        if (n >= KP) {
          const float t32 = __shfl_sync(ptx::FULL_MASK, kv, KP - 1);
          if (lane == 0) atomicMax(p.thr_global + gr, ptx::float_to_ordered(t32));
        }
      }
      __syncwarp();
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace score
}  // namespace seam
