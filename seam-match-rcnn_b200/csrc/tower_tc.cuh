// f3: the match head's conv tower -> 256-d embedding on the tensor cores.
//
// Reference (models/match_head.py:50-62, forward :67-69 / :93-95):
//   conv_seq = 3 x [Conv2d(256,256,3) + ReLU] + Conv2d(256,1024,3) + ReLU     (valid convolutions: 14 -> 12 -> 10 -> 8 -> 6)
//   pool     = AvgPool2d(6) + ReLU ;  linear = Linear(1024,256) + BatchNorm1d(256)      x: (K,256,14,14) -> x3: (K,256)
// ~0.53 GFLOP per ROI, up to 100 ROIs per image (detections_per_img): this is where the eval's device time goes once
// the aggregation / scoring path is fused.
//
// Convolution as SHIFTED GEMMs.  Activations are kept position-major ("NHWC"): a matrix [K*H*W rows, C columns] in fp16.
// On the input grid, the 3x3 valid convolution at flat position p is
//     out[p, :] = sum over taps (dy,dx) of  in[p + dy*W + dx, :] @ Wt[tap]          (256 x C_out per tap),
// which is right for every p whose (y,x) has y < H-2 and x < W-2 -- for those, p + dy*W + dx stays inside the same ROI --
// and garbage elsewhere.  So one output tile of 128 consecutive flat positions is 9 taps x 4 k-blocks = 36 plain
// tcgen05 MMAs-steps whose A operand is a CONTIGUOUS block of 128 activation rows shifted by the tap offset: a plain 2-D
// TMA box (out-of-range rows read as zeros), no im2col, no gather.  The epilogue (bias, ReLU, fp16) keeps the valid
// positions and writes them COMPACTED to the (H-2)x(W-2) grid the next layer runs on, so the waste stays at the border
// ring of each layer (73 / 69 / 64 / 56 % of the rows of a tile are useful) instead of compounding.
// Operands fp16, accumulation fp32 in tensor memory: the precision class of cuDNN's default TF32 path the reference runs.
//
// Kernel: persistent CTAs over (row tile, N tile of 256 output channels); warp 0 = TMA producer, warp 1 = MMA issuer
// (converged warp, elected lane), warp 2 = TMEM allocator, warps 4..7 = epilogue (thread = position row); 4-stage ring of
// 48 KB (A 128x64 + B 256x64 fp16, 128-byte swizzle); two 256-column accumulators alternate so the epilogue of tile i
// overlaps the MMAs of tile i+1.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_fp16.h>
#include "sm100_ptx.cuh"

namespace seam {
namespace tower {

constexpr int BM = 128, BN = 256, BK = 64, NSTAGE = 4, CIN = 256, NTAP = 9;
constexpr int THREADS = 256;
constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr uint32_t OFF_BIAS = NSTAGE * STAGE_BYTES;            // 1024 floats (all output channels of the layer)
constexpr uint32_t OFF_BAR = OFF_BIAS + 1024 * 4;
constexpr uint32_t OFF_TMEM = OFF_BAR + (2 * NSTAGE + 4) * 8;
constexpr uint32_t SMEM_BYTES = OFF_TMEM + 16 + 1024;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");

struct ConvParams {
  int H, W;                 // input grid of the layer
  int K;                    // ROIs
  int Cout;                 // 256 or 1024
  long long rows_in;        // K*H*W
  int m_tiles, n_tiles;
  const float* bias;        // (Cout)
  __half* out;              // (K*(H-2)*(W-2), Cout) compact
};

__global__ void __launch_bounds__(THREADS, 1)
conv3x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  float* bias_s = reinterpret_cast<float*>(smem + OFF_BIAS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* full = bars;                      // [NSTAGE] TMA -> MMA
  uint64_t* empty = bars + NSTAGE;            // [NSTAGE] MMA -> TMA
  uint64_t* t_full = bars + 2 * NSTAGE;       // [2] accumulator ready
  uint64_t* t_empty = t_full + 2;             // [2] accumulator drained
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(ptx::FULL_MASK, tid >> 5, 0);
  if (tid == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    for (int i = 0; i < NSTAGE; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&t_full[i], 1);
      ptx::mbar_init(&t_empty[i], 4);
    }
    ptx::fence_mbar_init();
  }
  for (int i = tid; i < p.Cout; i += THREADS) bias_s[i] = p.bias[i];
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr_s, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  const int total = p.m_tiles * p.n_tiles;

  if (warp == 0) {
    // ================================================================= TMA producer
    uint32_t stage = 0, sphase = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      const int mt = t / p.n_tiles, nt = t % p.n_tiles;      // N tiles of one row tile are adjacent: A stays in L2
      for (int tap = 0; tap < NTAP; ++tap) {
        const int row0 = mt * BM + (tap / 3) * p.W + (tap % 3);
        for (int kb = 0; kb < CIN / BK; ++kb) {
          ptx::mbar_wait(&empty[stage], sphase ^ 1, 301);
          if (ptx::elect_one()) {
            uint8_t* st = smem + stage * STAGE_BYTES;
            ptx::mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
            ptx::tma_load_2d(st, &tmA, &full[stage], kb * BK, row0);
            ptx::tma_load_2d(st + A_BYTES, &tmB, &full[stage], tap * CIN + kb * BK, nt * BN);
          }
          __syncwarp();
          if (++stage == NSTAGE) {
            stage = 0;
            sphase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    constexpr uint32_t idesc = ptx::umma_idesc(0 /*fp16*/, BM, BN);
    uint32_t stage = 0, sphase = 0, acc = 0, aphase = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      ptx::mbar_wait(&t_empty[acc], aphase ^ 1, 302);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int it = 0; it < NTAP * (CIN / BK); ++it) {
        ptx::mbar_wait(&full[stage], sphase, 303);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t a_addr = ptx::smem_u32(smem + stage * STAGE_BYTES), b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            ptx::umma_f16(d_tmem, ptx::umma_desc_k_sw128(a_addr + k * 32), ptx::umma_desc_k_sw128(b_addr + k * 32), idesc,
                          (it | k) != 0 ? 1u : 0u);
          ptx::umma_commit(&empty[stage]);
          if (it == NTAP * (CIN / BK) - 1) ptx::umma_commit(&t_full[acc]);
        }
        __syncwarp();
        if (++stage == NSTAGE) {
          stage = 0;
          sphase ^= 1;
        }
      }
      if (++acc == 2) {
        acc = 0;
        aphase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ================================================================= epilogue: thread = position row
    const int ew = warp - 4;
    const int HW = p.H * p.W, Ho = p.H - 2, Wo = p.W - 2;
    uint32_t acc = 0, aphase = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      const int mt = t / p.n_tiles, nt = t % p.n_tiles;
      const long long pos = (long long)mt * BM + ew * 32 + lane;
      const int roi = (int)(pos / HW), rem = (int)(pos % HW), y = rem / p.W, x = rem % p.W;
      const bool valid = pos < p.rows_in && y < Ho && x < Wo;
      __half* orow = p.out + ((size_t)roi * Ho * Wo + (size_t)y * Wo + x) * p.Cout + nt * BN;
      ptx::mbar_wait(&t_full[acc], aphase, 304);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld_x32(taddr + c0, r);
        ptx::tmem_ld_wait_x32(r);
        if (c0 == BN - 32) {                       // the accumulator is drained: hand it back before the last stores
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&t_empty[acc]);
        }
        if (valid) {
          const float* b = bias_s + nt * BN + c0;
          uint32_t h[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float v0 = fmaxf(__uint_as_float(r[2 * i]) + b[2 * i], 0.f);
            const float v1 = fmaxf(__uint_as_float(r[2 * i + 1]) + b[2 * i + 1], 0.f);
            const __half2 hh = __floats2half2_rn(v0, v1);
            h[i] = *reinterpret_cast<const uint32_t*>(&hh);
          }
          uint4* dst = reinterpret_cast<uint4*>(orow + c0);
#pragma unroll
          for (int i = 0; i < 4; ++i) dst[i] = make_uint4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
        }
      }
      if (++acc == 2) {
        acc = 0;
        aphase ^= 1;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------- layout / weight preparation
// x (K,256,14,14) fp32 NCHW -> (K*196, 256) fp16 position-major.  One CTA per (ROI, 32-channel group): coalesced reads
// of 32 x 196 floats, transposed through shared memory, 64-byte row segments out.
__global__ void __launch_bounds__(256) nchw_to_rows_kernel(const float* __restrict__ x, __half* __restrict__ out, int K) {
  __shared__ float tile[32][197];
  const int roi = blockIdx.x, cg = blockIdx.y;
  const float* src = x + ((size_t)roi * 256 + cg * 32) * 196;
  for (int e = threadIdx.x; e < 32 * 196; e += 256) tile[e / 196][e % 196] = src[e];
  __syncthreads();
  for (int e = threadIdx.x; e < 196 * 16; e += 256) {         // (position, channel pair)
    const int pos = e >> 4, c2 = (e & 15) * 2;
    const __half2 h = __floats2half2_rn(tile[c2][pos], tile[c2 + 1][pos]);
    *reinterpret_cast<__half2*>(out + ((size_t)roi * 196 + pos) * 256 + cg * 32 + c2) = h;
  }
}

// conv weight (Cout, 256, 3, 3) fp32 -> (Cout, 9*256) fp16 with k = tap*256 + ci (the K order the kernel sweeps)
__global__ void conv_weight_rows_kernel(const float* __restrict__ w, __half* __restrict__ out, int Cout) {
  const int co = blockIdx.x;
  for (int e = threadIdx.x; e < 9 * 256; e += blockDim.x) {
    const int tap = e / 256, ci = e % 256;
    out[(size_t)co * 2304 + e] = __float2half_rn(w[((size_t)co * 256 + ci) * 9 + tap]);
  }
}

// ---------------------------------------------------------------------------------------- pool + linear + BatchNorm (eval)
// a4 (K*36, 1024) fp16 (post-ReLU) -> AvgPool2d(6) (+ ReLU: a no-op on non-negative values) -> Linear(1024,256) ->
// BatchNorm1d(256) with running statistics -> fp32 row dst_row[roi] (or roi) of `out` (row stride 256 floats): the rows
// can be slots of the time-major x3_1_seq the aggregation kernel reads (models/match_head.py:101-111).
// One CTA per 8 ROIs: pooled vectors in shared memory, thread = output channel, the 1024 x 256 weight streamed once per
// CTA (transposed copy: consecutive threads read consecutive addresses).
constexpr int PL_ROIS = 8;
struct PoolLinearParams {
  const __half* a4;
  const float* wt;          // (1024, 256) = linear.weight transposed
  const float* lin_b;
  const float* bn_scale;    // gamma / sqrt(var + eps)
  const float* bn_shift;    // beta - mean * scale
  const long long* dst_row; // (K) or null
  float* out;
  int K;
};
__global__ void __launch_bounds__(256) pool_linear_bn_kernel(const PoolLinearParams p) {
  __shared__ float pooled[PL_ROIS][1024];
  const int r0 = blockIdx.x * PL_ROIS, nr = min(PL_ROIS, p.K - r0);
  for (int e = threadIdx.x; e < nr * 512; e += 256) {          // (roi, channel pair)
    const int r = e >> 9, c2 = (e & 511) * 2;
    const __half* src = p.a4 + ((size_t)(r0 + r) * 36) * 1024 + c2;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
    for (int pos = 0; pos < 36; ++pos) {
      const float2 v = __half22float2(*reinterpret_cast<const __half2*>(src + (size_t)pos * 1024));
      s0 += v.x;
      s1 += v.y;
    }
    pooled[r][c2] = fmaxf(s0 * (1.f / 36.f), 0.f);
    pooled[r][c2 + 1] = fmaxf(s1 * (1.f / 36.f), 0.f);
  }
  __syncthreads();
  const int o = threadIdx.x;
  float acc[PL_ROIS];
#pragma unroll
  for (int r = 0; r < PL_ROIS; ++r) acc[r] = 0.f;
  for (int c = 0; c < 1024; ++c) {
    const float w = __ldg(p.wt + (size_t)c * 256 + o);
#pragma unroll
    for (int r = 0; r < PL_ROIS; ++r) acc[r] = fmaf(pooled[r][c], w, acc[r]);
  }
  const float sc = p.bn_scale[o], sh = p.bn_shift[o], lb = p.lin_b[o];
  for (int r = 0; r < nr; ++r) {
    const long long row = p.dst_row ? p.dst_row[r0 + r] : (long long)(r0 + r);
    p.out[row * 256 + o] = fmaf(acc[r] + lb, sc, sh);
  }
}

// fold BatchNorm1d (eval) into scale / shift; transpose linear.weight (256,1024) -> (1024,256)
__global__ void tower_fold_kernel(const float* __restrict__ lin_w, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, const float* __restrict__ mean,
                                  const float* __restrict__ var, float eps, float* __restrict__ wt,
                                  float* __restrict__ scale, float* __restrict__ shift) {
  const int o = blockIdx.x;                                      // 256 blocks
  for (int c = threadIdx.x; c < 1024; c += blockDim.x) wt[(size_t)c * 256 + o] = lin_w[(size_t)o * 1024 + c];
  if (threadIdx.x == 0) {
    const float s = gamma[o] / sqrtf(var[o] + eps);
    scale[o] = s;
    shift[o] = beta[o] - mean[o] * s;
  }
}

}  // namespace tower
}  // namespace seam
