// K1b: the non-local block's g / W projections applied to the pooled frame mix, batched
// over tracks:   out[i] = pooled[i] + M r[i] + s_i * (W_W b_g) + active_i * b_W
// with M = W_W W_g (256x256).  This is models/nlb.py:74-75 (g), :95 (f @ g_x), :98-99
// (W(y) + x) after the attention pooling of models/match_head.py:149-151 has been pushed
// through the (linear) projections.
//
// nlb_gemm_simt_kernel: fp32 CUDA-core tile GEMM (exact fp32 products; reference-grade).
// It also serves seam_nlb_forward with rows = (batch, frame) and a channel-major store.
#pragma once
#include <cstdint>
#include "fold.cuh"

namespace seam {
namespace nlbgemm {

constexpr int TM = 64, TN = 64, TK = 16;

struct Params {
  const float* pooled;  // (rows,256) added to the product
  const float* R;       // (rows,256)
  const float* sv;      // (rows,2)
  const float* fold;
  float* out;
  int rows;
  int T;                // 0: out is (rows,256) row-major; >0: rows=(b,t), out is (B,256,T)
};

__global__ void __launch_bounds__(256) nlb_gemm_simt_kernel(const Params p) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN];
  const int t = threadIdx.x;
  const int row0 = blockIdx.x * TM, col0 = blockIdx.y * TN;
  const int ty = t >> 4, tx = t & 15;
  const float* Mt = p.fold + Fold::MT;
  float acc[4][4] = {};
  const int lr = t >> 2, lk = (t & 3) * 4;     // A loader: row, k offset
  const int bk = t >> 4, bo = (t & 15) * 4;    // B loader
  for (int k0 = 0; k0 < 256; k0 += TK) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + lr < p.rows) a = *reinterpret_cast<const float4*>(p.R + (size_t)(row0 + lr) * 256 + k0 + lk);
    As[lk + 0][lr] = a.x;
    As[lk + 1][lr] = a.y;
    As[lk + 2][lr] = a.z;
    As[lk + 3][lr] = a.w;
    *reinterpret_cast<float4*>(&Bs[bk][bo]) = *reinterpret_cast<const float4*>(Mt + (size_t)(k0 + bk) * 256 + col0 + bo);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w};
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float4 wbg = *reinterpret_cast<const float4*>(p.fold + Fold::WBG + col0 + tx * 4);
  const float4 bW = *reinterpret_cast<const float4*>(p.fold + Fold::BW + col0 + tx * 4);
  const float wb[4] = {wbg.x, wbg.y, wbg.z, wbg.w};
  const float bw[4] = {bW.x, bW.y, bW.z, bW.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = row0 + ty * 4 + i;
    if (row >= p.rows) continue;
    const float s = p.sv[2 * row], act = p.sv[2 * row + 1];
    const float4 pl = *reinterpret_cast<const float4*>(p.pooled + (size_t)row * 256 + col0 + tx * 4);
    const float pp[4] = {pl.x, pl.y, pl.z, pl.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = pp[j] + (acc[i][j] + fmaf(s, wb[j], act * bw[j]));
    if (p.T == 0) {
      *reinterpret_cast<float4*>(p.out + (size_t)row * 256 + col0 + tx * 4) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
      const int b = row / p.T, tt = row - b * p.T;
#pragma unroll
      for (int j = 0; j < 4; ++j) p.out[((size_t)b * 256 + col0 + tx * 4 + j) * p.T + tt] = o[j];
    }
  }
}

// Front half of the full (un-pooled) block for seam_nlb_forward: x (B,256,T) channel-major.
// One CTA per batch element; writes row-major Xt[(b,t)][c] = x, R[(b,t)][c] = sum_j f_tj x_j,
// sv[(b,t)] = {sum_j f_tj, 1}.  models/nlb.py:78-95.
__global__ void __launch_bounds__(256) nlb_full_front_kernel(const float* __restrict__ x, int T,
                                                             const float* __restrict__ fold, float* __restrict__ Xt,
                                                             float* __restrict__ R, float* __restrict__ sv) {
  extern __shared__ float sm[];
  float* xs = sm;                  // [T][257]
  float* a_s = xs + T * 257;       // [T]
  float* b_s = a_s + T;            // [T]
  const int b = blockIdx.x, c = threadIdx.x;
  const float* xb = x + (size_t)b * 256 * T;
  for (int i = threadIdx.x; i < 256 * T; i += 256) {
    const int cc = i / T, tt = i - cc * T;
    xs[tt * 257 + cc] = xb[i];
  }
  __syncthreads();
  const int warp = c >> 5, lane = c & 31;
  for (int tt = warp; tt < T; tt += 8) {
    float sa = 0.f, sb = 0.f;
    for (int k = lane; k < 256; k += 32) {
      const float xv = xs[tt * 257 + k];
      sa = fmaf(xv, fold[Fold::U_THETA + k], sa);
      sb = fmaf(xv, fold[Fold::U_PHI + k], sb);
    }
    sa = ptx::warp_sum(sa);
    sb = ptx::warp_sum(sb);
    if (lane == 0) {
      a_s[tt] = sa + fold[Fold::CONSTS + 0];
      b_s[tt] = sb + fold[Fold::CONSTS + 1];
    }
  }
  __syncthreads();
  const float invT = 1.f / (float)T;
  for (int tt = 0; tt < T; ++tt) {
    float r = 0.f, fs = 0.f;
    const float at = a_s[tt];
    for (int j = 0; j < T; ++j) {
      const float f = fmaxf(at + b_s[j], 0.f) * invT;
      r = fmaf(f, xs[j * 257 + c], r);
      fs += f;
    }
    const size_t row = (size_t)b * T + tt;
    Xt[row * 256 + c] = xs[tt * 257 + c];
    R[row * 256 + c] = r;
    if (c == 0) {
      sv[2 * row] = fs;
      sv[2 * row + 1] = 1.f;
    }
  }
}

}  // namespace nlbgemm
}  // namespace seam
