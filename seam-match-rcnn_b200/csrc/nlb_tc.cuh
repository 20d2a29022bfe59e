// K1b: the non-local block's g / W projections applied to the pooled frame mix, batched over
// tracks on the tensor cores:   out[i] = pooled'[i] + M r[i],   M = W_W W_g (256 x 256).
// This is models/nlb.py:74-75 (g), :95 (f @ g_x), :98-99 (W(y) + x) after the attention pooling
// of models/match_head.py:149-151 has been pushed through the (linear) projections; pooled'
// already carries the bias terms (aggregate_warp.cuh).
//
// fp32-grade product from tf32 tensor-core passes: r = r_hi + r_lo and M = M_hi + M_lo are split
// into tf32-exact halves (the low 13 mantissa bits of every *_hi word are zero), and
//   M r ~= r_hi M_hi + r_hi M_lo + r_lo M_hi      (the dropped r_lo M_lo term is ~2^-22 relative)
// is accumulated in fp32 in TMEM by tcgen05.mma kind::tf32.
//
// One CTA per 128 tracks, N = 256 output channels, K = 256 in 16 k-blocks of 16 fp32 (64-byte
// swizzle) through a 4-stage ring of 48 KB.  Measured dead ends: fetching the CTA's pooled' tile into
// shared memory by TMA (leaves room for 2 stages only: 30 us instead of 25), a 3-deep register
// prefetch of pooled' started during the main loop (spills, competes with the operand loads: 33 us).  Roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one lane),
// warp 2 = TMEM allocator, warps 4..7 = epilogue (thread = track row).
#pragma once
#include <cstdint>
#include <cuda.h>
#include "sm100_ptx.cuh"

namespace seam {
namespace nlbtc {

constexpr int BM = 128, BN = 256, BK = 16, NKB = 16, NSTAGE = 4;   // k-blocks of 16 fp32 = 64-byte swizzle rows
constexpr int THREADS = 256;
constexpr uint32_t A_BYTES = BM * BK * 4;   // 8 KB
constexpr uint32_t B_BYTES = BN * BK * 4;   // 16 KB
constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // r_hi, r_lo, M_hi, M_lo
constexpr uint32_t OFF_BAR = NSTAGE * STAGE_BYTES;
constexpr uint32_t OFF_TMEM = OFF_BAR + (2 * NSTAGE + 1) * 8;
constexpr uint32_t SMEM_BYTES = OFF_TMEM + 16 + 1024;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");

struct Params {
  const float* pooled;   // (rows,256) pooled'
  float* out;            // (rows,256)
  int rows;
};

__global__ void __launch_bounds__(THREADS, 1)
nlb_tc_kernel(const __grid_constant__ CUtensorMap tmRh, const __grid_constant__ CUtensorMap tmRl,
              const __grid_constant__ CUtensorMap tmMh, const __grid_constant__ CUtensorMap tmMl, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* full = bars;                  // [NSTAGE]
  uint64_t* empty = bars + NSTAGE;        // [NSTAGE]
  uint64_t* d_full = bars + 2 * NSTAGE;   // accumulator complete
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(ptx::FULL_MASK, tid >> 5, 0);
  const int row0 = blockIdx.x * BM;

  if (tid == 0) {
    ptx::prefetch_tensormap(&tmRh);
    ptx::prefetch_tensormap(&tmRl);
    ptx::prefetch_tensormap(&tmMh);
    ptx::prefetch_tensormap(&tmMl);
    for (int i = 0; i < NSTAGE; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    ptx::mbar_init(d_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr_s, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < NKB; ++kb) {
        const int stage = kb % NSTAGE;
        const uint32_t ph = (uint32_t)(kb / NSTAGE) & 1u;
        ptx::mbar_wait(&empty[stage], ph ^ 1);
        uint8_t* st = smem + stage * STAGE_BYTES;
        ptx::mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
        ptx::tma_load_2d(st, &tmRh, &full[stage], kb * BK, row0);
        ptx::tma_load_2d(st + A_BYTES, &tmRl, &full[stage], kb * BK, row0);
        ptx::tma_load_2d(st + 2 * A_BYTES, &tmMh, &full[stage], kb * BK, 0);
        ptx::tma_load_2d(st + 2 * A_BYTES + B_BYTES, &tmMl, &full[stage], kb * BK, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc(2 /*tf32*/, BM, BN);
      for (int kb = 0; kb < NKB; ++kb) {
        const int stage = kb % NSTAGE;
        const uint32_t ph = (uint32_t)(kb / NSTAGE) & 1u;
        ptx::mbar_wait(&full[stage], ph);
        ptx::tc_fence_after();
        const uint32_t st = ptx::smem_u32(smem + stage * STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          const uint64_t ah = ptx::umma_desc_k_sw64(st + k * 32);
          const uint64_t al = ptx::umma_desc_k_sw64(st + A_BYTES + k * 32);
          const uint64_t bh = ptx::umma_desc_k_sw64(st + 2 * A_BYTES + k * 32);
          const uint64_t bl = ptx::umma_desc_k_sw64(st + 2 * A_BYTES + B_BYTES + k * 32);
          ptx::umma_tf32(tmem_base, al, bh, idesc, (kb | k) != 0 ? 1u : 0u);   // small terms first
          ptx::umma_tf32(tmem_base, ah, bl, idesc, 1u);
          ptx::umma_tf32(tmem_base, ah, bh, idesc, 1u);
        }
        ptx::umma_commit(&empty[stage]);
      }
      ptx::umma_commit(d_full);
    }
  } else if (warp >= 4) {
    // Epilogue.  A thread owns one track row of the accumulator, but a row-per-thread walk over
    // global memory would touch 32 different 1 KB rows per instruction; each warp therefore
    // transposes 32x32 blocks through a padded tile in the (now idle) operand stages, so that
    // every global access is one 128-byte line.
    const int ew = warp - 4;
    ptx::mbar_wait(d_full, 0);
    ptx::tc_fence_after();
    float* tile = reinterpret_cast<float*>(smem) + ew * (32 * 33);
    const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16);
    const int rbase = row0 + ew * 32;
    // pooled' for chunk ch+1 is fetched (32 independent 128-byte lines per warp) while chunk ch is
    // transposed and written
    float pl[32], pn[32];
    auto fetch = [&](int ch, float (&dst)[32]) {
#pragma unroll
      for (int rr = 0; rr < 32; ++rr) {
        const int row = rbase + rr;
        dst[rr] = row < p.rows ? __ldg(p.pooled + (size_t)row * 256 + ch * 32 + lane) : 0.f;
      }
    };
    fetch(0, pl);
#pragma unroll 1
    for (int ch = 0; ch < BN / 32; ++ch) {
      if (ch + 1 < BN / 32) fetch(ch + 1, pn);
      uint32_t r[32];
      ptx::tmem_ld_x32(taddr + ch * 32, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 32; ++c) tile[lane * 33 + c] = __uint_as_float(r[c]);
      __syncwarp();
#pragma unroll
      for (int rr = 0; rr < 32; ++rr) {
        const int row = rbase + rr;
        if (row < p.rows) p.out[(size_t)row * 256 + ch * 32 + lane] = pl[rr] + tile[rr * 33 + lane];
      }
      __syncwarp();
#pragma unroll
      for (int rr = 0; rr < 32; ++rr) pl[rr] = pn[rr];
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 256);
}

}  // namespace nlbtc
}  // namespace seam
