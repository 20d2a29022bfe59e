// K0: fold the reference's aggregator weights into the collapsed form used by the kernels.
//
// Reference parameters (models/nlb.py:34-60, models/match_head.py:64,86):
//   theta, phi, g : Conv1d(256->128, k=1)+bias     W : Conv1d(128->256, k=1)+bias
//   concat_project: Conv2d(256->1, k=1, no bias)   attention_scorer: Linear(256,1)
//   last          : Linear(256,2)
// Every projection is linear and only the pooled descriptor leaves the block, so per frame
// only four scalars are needed (DESIGN.md "K1 algebra"):
//   a_t = x_t.u_theta + c_theta     b_t = x_t.u_phi + c_phi
//   c_t = x_t.u_g + c_g             d_t = x_t.w_a
// with u_theta = W_theta^T wc[:128], u_phi = W_phi^T wc[128:], v = W_W^T w_a, u_g = W_g^T v,
// c_s = b_W.w_a + b_a, and per track one 256->256 map M = W_W W_g plus wbg = W_W b_g.
// Sums are accumulated in fp64 and rounded once to fp32.
#pragma once
#include <cstdint>
#include <cuda_fp16.h>

namespace seam {

// offsets (in floats) into the folded-weight buffer owned by the handle
struct Fold {
  static constexpr int U_THETA = 0;
  static constexpr int U_PHI = 256;
  static constexpr int U_G = 512;
  static constexpr int W_A = 768;
  static constexpr int CONSTS = 1024;   // c_theta, c_phi, c_g, c_s, db, b0, b1, -
  static constexpr int WBG = 1280;
  static constexpr int BW = 1536;
  static constexpr int DW = 1792;
  static constexpr int LAST_W = 2048;   // (2,256)
  static constexpr int MT = 2560;       // Mt[k*256 + o] = M[o][k], fp32
  // M as two fp16 terms, M * S = M1 + M2 (S = the power of two that brings max|M| into [1,2); CONSTS+8..10),
  // laid out as the fused aggregation kernel wants them (aggregate_fused.cuh): images of the tensor-memory
  // resident A operand -- word (h*COLS + c)*128 + lane = {M[128h+lane][2c], M[128h+lane][2c+1]} -- for all
  // of M1 (COLS = 128) and the first M2_KT k's of M2 (COLS = M2_KT/2), and the remaining k's of M2 as the
  // 64-byte-swizzled K-major shared-memory tile (128 rows x 64 B per half).
  static constexpr int M2_KT = 224;
  static constexpr int M1_IMG = MT + 65536;                  // 2*128*128 words
  static constexpr int M2_IMG = M1_IMG + 2 * 128 * 128;      // 2*(M2_KT/2)*128 words
  static constexpr int M2_TAIL = M2_IMG + M2_KT * 128;       // 256 rows x (256-M2_KT) fp16
  static constexpr int TOTAL = M2_TAIL + 256 * (256 - M2_KT) / 2;
};

struct FoldIn {
  const float *theta_w, *theta_b, *phi_w, *phi_b, *g_w, *g_b, *W_w, *W_b, *concat_w, *att_w, *att_b, *last_w,
      *last_b;
};

// one block of 256 threads: all the 256-vectors and scalars
__global__ void fold_vectors_kernel(FoldIn in, float* __restrict__ fold) {
  __shared__ double v_s[128];
  __shared__ double red[256];
  const int k = threadIdx.x;
  // v[c] = sum_o W_W[o][c] * w_a[o]
  if (k < 128) {
    double s = 0.0;
    for (int o = 0; o < 256; ++o) s += (double)in.W_w[o * 128 + k] * (double)in.att_w[o];
    v_s[k] = s;
  }
  __syncthreads();
  double ut = 0.0, up = 0.0, ug = 0.0;
  for (int c = 0; c < 128; ++c) {
    ut += (double)in.theta_w[c * 256 + k] * (double)in.concat_w[c];
    up += (double)in.phi_w[c * 256 + k] * (double)in.concat_w[128 + c];
    ug += (double)in.g_w[c * 256 + k] * v_s[c];
  }
  fold[Fold::U_THETA + k] = (float)ut;
  fold[Fold::U_PHI + k] = (float)up;
  fold[Fold::U_G + k] = (float)ug;
  fold[Fold::W_A + k] = in.att_w[k];
  double wb = 0.0;
  for (int c = 0; c < 128; ++c) wb += (double)in.W_w[k * 128 + c] * (double)in.g_b[c];
  fold[Fold::WBG + k] = (float)wb;
  fold[Fold::BW + k] = in.W_b[k];
  fold[Fold::DW + k] = in.last_w[256 + k] - in.last_w[k];
  fold[Fold::LAST_W + k] = in.last_w[k];
  fold[Fold::LAST_W + 256 + k] = in.last_w[256 + k];
  // scalars: four block reductions
  double part[4];
  part[0] = k < 128 ? (double)in.theta_b[k] * (double)in.concat_w[k] : 0.0;
  part[1] = k < 128 ? (double)in.phi_b[k] * (double)in.concat_w[128 + k] : 0.0;
  part[2] = k < 128 ? (double)in.g_b[k] * v_s[k] : 0.0;
  part[3] = (double)in.W_b[k] * (double)in.att_w[k];
  for (int s = 0; s < 4; ++s) {
    __syncthreads();
    red[k] = part[s];
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
      if (k < off) red[k] += red[k + off];
      __syncthreads();
    }
    if (k == 0) {
      double r = red[0];
      if (s == 3) r += (double)in.att_b[0];
      fold[Fold::CONSTS + s] = (float)r;
    }
  }
  if (k == 0) {
    fold[Fold::CONSTS + 4] = in.last_b[1] - in.last_b[0];
    fold[Fold::CONSTS + 5] = in.last_b[0];
    fold[Fold::CONSTS + 6] = in.last_b[1];
    fold[Fold::CONSTS + 7] = 0.f;
    fold[Fold::CONSTS + 8] = 0.f;   // max |M| (fold_matrix_kernel, atomicMax on the bit image)
  }
}

// scorer-only variant (seam_load_scorer): just `last`
__global__ void fold_scorer_kernel(const float* __restrict__ last_w, const float* __restrict__ last_b,
                                   float* __restrict__ fold) {
  const int k = threadIdx.x;
  fold[Fold::DW + k] = last_w[256 + k] - last_w[k];
  fold[Fold::LAST_W + k] = last_w[k];
  fold[Fold::LAST_W + 256 + k] = last_w[256 + k];
  if (k == 0) {
    fold[Fold::CONSTS + 4] = last_b[1] - last_b[0];
    fold[Fold::CONSTS + 5] = last_b[0];
    fold[Fold::CONSTS + 6] = last_b[1];
  }
}

// M[o][k] = sum_c W_W[o][c] * W_g[c][k]; grid = 256 blocks (o), 256 threads (k)
__global__ void fold_matrix_kernel(const float* __restrict__ W_w, const float* __restrict__ g_w,
                                   float* __restrict__ fold) {
  __shared__ float wrow[128];
  const int o = blockIdx.x, k = threadIdx.x;
  if (k < 128) wrow[k] = W_w[o * 128 + k];
  __syncthreads();
  double s = 0.0;
  for (int c = 0; c < 128; ++c) s += (double)wrow[c] * (double)g_w[c * 256 + k];
  const float m = (float)s;
  fold[Fold::MT + k * 256 + o] = m;
  float amax = fabsf(m);
  for (int off = 16; off > 0; off >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
  if ((k & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(fold + Fold::CONSTS + 8), __float_as_uint(amax));
}

// power of two that brings a non-negative finite x into [1,2) (1 for x == 0); *inv = its reciprocal
__device__ __forceinline__ float pow2_normaliser(float x, float* inv) {
  unsigned eb = (__float_as_uint(x) >> 23) & 0xffu;
  if (x == 0.f) eb = 127u;
  eb = eb < 1u ? 1u : (eb > 253u ? 253u : eb);
  *inv = __uint_as_float(eb << 23);
  return __uint_as_float((254u - eb) << 23);
}

// after fold_matrix_kernel (same stream): the fp16 two-term images of M * S.  grid = 256 blocks (o), 256 threads (k)
__global__ void fold_m16_kernel(float* __restrict__ fold) {
  const int o = blockIdx.x, k = threadIdx.x;
  float inv;
  const float S = pow2_normaliser(fold[Fold::CONSTS + 8], &inv);
  if (o == 0 && k == 0) {
    fold[Fold::CONSTS + 9] = inv;   // 1 / S
    fold[Fold::CONSTS + 10] = S;
  }
  const float m = fold[Fold::MT + k * 256 + o] * S;
  const __half h1 = __float2half_rn(m);
  const __half h2 = __float2half_rn(m - __half2float(h1));
  const unsigned u1 = __half_as_ushort(h1), u2 = __half_as_ushort(h2);
  const unsigned p1 = __shfl_down_sync(0xffffffffu, u1, 1), p2 = __shfl_down_sync(0xffffffffu, u2, 1);
  if (k & 1) return;
  unsigned* f = reinterpret_cast<unsigned*>(fold);
  const int h = o >> 7, row = o & 127, c = k >> 1;
  // images: [column / 4][row 0..127][column % 4] words -- a lane fetches four columns of its row with one 16-byte load
  f[Fold::M1_IMG + ((h * 128 + c) >> 2) * 512 + row * 4 + (c & 3)] = u1 | (p1 << 16);
  if (k < Fold::M2_KT) {
    f[Fold::M2_IMG + ((h * (Fold::M2_KT / 2) + c) >> 2) * 512 + row * 4 + (c & 3)] = u2 | (p2 << 16);
  } else {
    constexpr int KS = 256 - Fold::M2_KT;            // 32 fp16 = 64-byte rows, SWIZZLE_64B: chunk ^= (row >> 1) & 3
    const int kk = k - Fold::M2_KT;
    const int byte = h * (128 * KS * 2) + row * (KS * 2) + (((kk >> 3) ^ ((row >> 1) & 3)) << 4) + (kk & 7) * 2;
    f[Fold::M2_TAIL + (byte >> 2)] = u2 | (p2 << 16);
  }
}

}  // namespace seam
