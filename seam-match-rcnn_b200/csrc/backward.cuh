// f4: backward of the hot path for training.
//
// The reference trains through the aggregator with autograd (models/match_head.py:339, 429: the losses consume
// x5 = last((x3_1b - x3_2)^2) of TemporalAggregationNLB.forward's x-branch; stuffs/engine.py:158-185).  The forward
// kernels work on FOLDED weights (DESIGN.md "K1 algebra"), which has no use for a backward pass: the gradients
// with respect to theta / phi / g / W / concat_project / attention_scorer / last need the un-folded block
// (models/nlb.py:66-101).  So the backward kernels recompute a track's forward in the reference's own formulation
// (shared memory, fp32) and differentiate that:
//
//   theta_t = Wt x_t + bt   phi_j = Wp x_j + bp   g_j = Wg x_j + bg                        nlb.py:72-79
//   pre_tj = wc[:128].theta_t + wc[128:].phi_j ;  f_tj = relu(pre_tj) / T                   nlb.py:84-93
//   y_t = sum_j f_tj g_j ;  z_t = WW y_t + bW + x_t                                         nlb.py:95-99
//   s_t = wa.z_t + ba ;  p = softmax_t(s) ;  out = sum_t p_t z_t                            match_head.py:149-151
//   (T == 1: the block is skipped, out = x_0: match_head.py:145-147)
//
// agg_backward_kernel: one CTA per track (tracks of up to BWD_MAX_T = 16 frames: training uses ~10,
// train_movingfashion.py:165), weight gradients accumulated with atomicAdd (training batches are tens of tracks).
// scorer_backward_{q,g}_kernel: gradients of x5 = last((q - g)^2) (match_head.py:160-162) for the dense (Q,G,2)
// logits training uses.
#pragma once
#include <cstdint>
#include "sm100_ptx.cuh"

namespace seam {
namespace bwd {

constexpr int D = 256, DI = 128, BWD_MAX_T = 16;

struct AggWeights {     // the reference's parameters, fp32, device pointers
  const float *theta_w, *theta_b, *phi_w, *phi_b, *g_w, *g_b, *W_w, *W_b, *concat_w, *att_w, *att_b;
};
struct AggGrads {       // same shapes, accumulated into (caller zero-fills)
  float *theta_w, *theta_b, *phi_w, *phi_b, *g_w, *g_b, *W_w, *W_b, *concat_w, *att_w, *att_b;
};
struct AggBwdParams {
  const float* seq;
  const uint8_t* mask;
  const int32_t* lens;
  int Tmax, Q;
  long long frame_stride, track_stride;
  const float* dout;      // (Q,256)
  float* dseq;            // (1+Tmax, Q, 256) contiguous; row 0 and padded frames are zero-filled by the caller
  AggWeights w;
  AggGrads g;
};

constexpr size_t agg_bwd_smem_bytes(int T) {
  // x, z, dz: T*256 each; theta, phi, gg, y, dy, dth, dph, dgg: T*128 each; f, pre: T*T each; small vectors
  return (size_t)(3 * T * D + 8 * T * DI + 2 * T * T + 8 * T + 64) * sizeof(float);
}

// out[i] = add + sum_k A[i*lda + k] * B[k], i < n; one warp per row
__device__ __forceinline__ void rows_dot(float* out, const float* A, int lda, const float* B, int n, int K, float add) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < n; i += 8) {
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(A[i * lda + k], B[k], acc);
    acc = ptx::warp_sum(acc);
    if (lane == 0) out[i] = acc + add;
  }
}

__global__ void __launch_bounds__(256) agg_backward_kernel(const AggBwdParams p) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  const int track = blockIdx.x;
  // ---- track length (models/match_head.py:136-139)
  int len;
  if (p.lens) {
    len = p.lens[track];
  } else if (p.mask) {
    len = p.Tmax;
    for (int t = 0; t <= p.Tmax; ++t)
      if (p.mask[(size_t)track * (1 + p.Tmax) + t]) {
        len = t - 1;
        break;
      }
  } else {
    len = p.Tmax;
  }
  len = max(0, min(len, p.Tmax));
  if (len == 0) return;
  const float dout_o = p.dout[(size_t)track * D + tid];
  if (len == 1) {                       // block skipped, softmax of one logit: out = x_0
    p.dseq[((size_t)1 * p.Q + track) * D + tid] = dout_o;
    return;
  }
  const int T = len;
  float* x = sm;                        // [T][256]
  float* z = x + T * D;                 // [T][256]
  float* dz = z + T * D;                // [T][256]
  float* th = dz + T * D;               // [T][128]
  float* ph = th + T * DI;
  float* gg = ph + T * DI;
  float* y = gg + T * DI;
  float* dy = y + T * DI;
  float* dth = dy + T * DI;
  float* dph = dth + T * DI;
  float* dgg = dph + T * DI;
  float* f = dgg + T * DI;              // [T][T]
  float* du = f + T * T;                // [T][T]  (first pre, then du)
  float* va = du + T * T;               // [T] a_t = wc_theta . theta_t
  float* vb = va + T;                   // [T] b_j
  float* vs = vb + T;                   // [T] s_t -> p_t
  float* vd = vs + T;                   // [T] dout . z_t -> ds_t
  float* vda = vd + T;                  // [T] sum_j du_tj
  float* vdb = vda + T;                 // [T] sum_t du_tj
  const float invT = 1.f / (float)T;

  for (int t = 0; t < T; ++t)
    x[t * D + tid] = p.seq[(long long)(t + 1) * p.frame_stride + (long long)track * p.track_stride + tid];
  __syncthreads();

  // ---- forward: projections (thread c < 128: theta and g; thread 128 + c: phi)
  {
    const int c = tid & 127;
    const bool second = tid >= 128;
    const float* W1 = (second ? p.w.phi_w : p.w.theta_w) + (size_t)c * D;
    float acc[BWD_MAX_T];
#pragma unroll
    for (int t = 0; t < BWD_MAX_T; ++t) acc[t] = 0.f;
    for (int k = 0; k < D; ++k) {
      const float w = W1[k];
#pragma unroll
      for (int t = 0; t < BWD_MAX_T; ++t)
        if (t < T) acc[t] = fmaf(w, x[t * D + k], acc[t]);
    }
    const float b1 = second ? p.w.phi_b[c] : p.w.theta_b[c];
    float* dst = second ? ph : th;
#pragma unroll
    for (int t = 0; t < BWD_MAX_T; ++t)
      if (t < T) dst[t * DI + c] = acc[t] + b1;
    // g: threads split the frames (even / odd) so that all 256 threads work
    const float* Wg = p.w.g_w + (size_t)c * D;
    const float bg = p.w.g_b[c];
    for (int t = second ? 1 : 0; t < T; t += 2) {
      float a = 0.f;
      for (int k = 0; k < D; ++k) a = fmaf(Wg[k], x[t * D + k], a);
      gg[t * DI + c] = a + bg;
    }
  }
  __syncthreads();
  rows_dot(va, th, DI, p.w.concat_w, T, DI, 0.f);
  rows_dot(vb, ph, DI, p.w.concat_w + DI, T, DI, 0.f);
  __syncthreads();
  for (int e = tid; e < T * T; e += 256) {
    const float pre = va[e / T] + vb[e % T];
    du[e] = pre;                                             // kept for the ReLU mask
    f[e] = fmaxf(pre, 0.f) * invT;
  }
  __syncthreads();
  if (tid < DI) {
    for (int t = 0; t < T; ++t) {
      float a = 0.f;
      for (int j = 0; j < T; ++j) a = fmaf(f[t * T + j], gg[j * DI + tid], a);
      y[t * DI + tid] = a;
    }
  }
  __syncthreads();
  {
    const float* Wrow = p.w.W_w + (size_t)tid * DI;          // thread = output channel o
    float acc[BWD_MAX_T];
#pragma unroll
    for (int t = 0; t < BWD_MAX_T; ++t) acc[t] = 0.f;
    for (int c = 0; c < DI; ++c) {
      const float w = Wrow[c];
#pragma unroll
      for (int t = 0; t < BWD_MAX_T; ++t)
        if (t < T) acc[t] = fmaf(w, y[t * DI + c], acc[t]);
    }
    const float bW = p.w.W_b[tid];
#pragma unroll
    for (int t = 0; t < BWD_MAX_T; ++t)
      if (t < T) z[t * D + tid] = acc[t] + bW + x[t * D + tid];
  }
  __syncthreads();
  rows_dot(vs, z, D, p.w.att_w, T, D, p.w.att_b[0]);
  // dout . z_t
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int t = warp; t < T; t += 8) {
      float a = 0.f;
      for (int k = lane; k < D; k += 32) a = fmaf(p.dout[(size_t)track * D + k], z[t * D + k], a);
      a = ptx::warp_sum(a);
      if (lane == 0) vd[t] = a;
    }
  }
  __syncthreads();
  if (tid == 0) {                                            // softmax over T <= 16 frames and its backward
    float m = -INFINITY;
    for (int t = 0; t < T; ++t) m = fmaxf(m, vs[t]);
    float zsum = 0.f;
    for (int t = 0; t < T; ++t) {
      vs[t] = expf(vs[t] - m);
      zsum += vs[t];
    }
    float dsum = 0.f;
    for (int t = 0; t < T; ++t) {
      vs[t] /= zsum;
      dsum = fmaf(vs[t], vd[t], dsum);
    }
    float dba = 0.f;
    for (int t = 0; t < T; ++t) {
      vd[t] = vs[t] * (vd[t] - dsum);                        // ds_t
      dba += vd[t];
    }
    atomicAdd(p.g.att_b, dba);
  }
  __syncthreads();

  // ---- backward
  {
    const float wa_o = p.w.att_w[tid];
    float dwa = 0.f, dbW = 0.f;
    for (int t = 0; t < T; ++t) {
      const float v = fmaf(vs[t], dout_o, vd[t] * wa_o);     // dz_t[o] = p_t dout[o] + ds_t wa[o]
      dz[t * D + tid] = v;
      dwa = fmaf(vd[t], z[t * D + tid], dwa);
      dbW += v;
    }
    atomicAdd(p.g.att_w + tid, dwa);
    atomicAdd(p.g.W_b + tid, dbW);
  }
  __syncthreads();
  // dW_W[o][c] += sum_t dz_t[o] y_t[c]   (thread = o)
  for (int c = 0; c < DI; ++c) {
    float a = 0.f;
    for (int t = 0; t < T; ++t) a = fmaf(dz[t * D + tid], y[t * DI + c], a);
    atomicAdd(p.g.W_w + (size_t)tid * DI + c, a);
  }
  // dy_t[c] = sum_o W_W[o][c] dz_t[o]   (thread = c, two frame halves)
  {
    const int c = tid & 127;
    for (int t = tid >> 7; t < T; t += 2) {
      float a = 0.f;
      for (int o = 0; o < D; ++o) a = fmaf(p.w.W_w[(size_t)o * DI + c], dz[t * D + o], a);
      dy[t * DI + c] = a;
    }
  }
  __syncthreads();
  // df_tj = dy_t . g_j ; du_tj = df_tj [pre_tj > 0] / T
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int e = warp; e < T * T; e += 8) {
      const int t = e / T, j = e % T;
      float a = 0.f;
      for (int c = lane; c < DI; c += 32) a = fmaf(dy[t * DI + c], gg[j * DI + c], a);
      a = ptx::warp_sum(a);
      if (lane == 0) du[e] = du[e] > 0.f ? a * invT : 0.f;
    }
  }
  __syncthreads();
  if (tid < T) {
    float a = 0.f, b = 0.f;
    for (int j = 0; j < T; ++j) {
      a += du[tid * T + j];
      b += du[j * T + tid];
    }
    vda[tid] = a;
    vdb[tid] = b;
  }
  __syncthreads();
  if (tid < DI) {
    const int c = tid;
    const float wt = p.w.concat_w[c], wp = p.w.concat_w[DI + c];
    float dwt = 0.f, dwp = 0.f, dbt = 0.f, dbp = 0.f, dbg = 0.f;
    for (int t = 0; t < T; ++t) {
      dwt = fmaf(vda[t], th[t * DI + c], dwt);
      dwp = fmaf(vdb[t], ph[t * DI + c], dwp);
      const float a = vda[t] * wt, b = vdb[t] * wp;
      dth[t * DI + c] = a;
      dph[t * DI + c] = b;
      dbt += a;
      dbp += b;
      float gsum = 0.f;                                      // dg_j[c] = sum_t f_tj dy_t[c]   (here j = t)
      for (int u = 0; u < T; ++u) gsum = fmaf(f[u * T + t], dy[u * DI + c], gsum);
      dgg[t * DI + c] = gsum;
      dbg += gsum;
    }
    atomicAdd(p.g.concat_w + c, dwt);
    atomicAdd(p.g.concat_w + DI + c, dwp);
    atomicAdd(p.g.theta_b + c, dbt);
    atomicAdd(p.g.phi_b + c, dbp);
    atomicAdd(p.g.g_b + c, dbg);
  }
  __syncthreads();
  // projection weight gradients and the input gradient (thread = input channel k)
  {
    const int k = tid;
    float dx[BWD_MAX_T];
#pragma unroll
    for (int t = 0; t < BWD_MAX_T; ++t) dx[t] = t < T ? dz[t * D + k] : 0.f;     // residual path
    for (int c = 0; c < DI; ++c) {
      const float wt = p.w.theta_w[(size_t)c * D + k], wp = p.w.phi_w[(size_t)c * D + k], wg = p.w.g_w[(size_t)c * D + k];
      float gt = 0.f, gp = 0.f, ggr = 0.f;
#pragma unroll
      for (int t = 0; t < BWD_MAX_T; ++t) {
        if (t < T) {
          const float xv = x[t * D + k];
          const float a = dth[t * DI + c], b = dph[t * DI + c], g3 = dgg[t * DI + c];
          gt = fmaf(a, xv, gt);
          gp = fmaf(b, xv, gp);
          ggr = fmaf(g3, xv, ggr);
          dx[t] = fmaf(wt, a, fmaf(wp, b, fmaf(wg, g3, dx[t])));
        }
      }
      atomicAdd(p.g.theta_w + (size_t)c * D + k, gt);
      atomicAdd(p.g.phi_w + (size_t)c * D + k, gp);
      atomicAdd(p.g.g_w + (size_t)c * D + k, ggr);
    }
#pragma unroll
    for (int t = 0; t < BWD_MAX_T; ++t)
      if (t < T) p.dseq[((size_t)(t + 1) * p.Q + track) * D + k] = dx[t];
  }
}

// ---------------------------------------------------------------------------------------- pair scorer backward
// x5[i,j,c] = sum_k W[c,k] (q_ik - g_jk)^2 + b_c.  One CTA per query (thread = channel k): dq, and this query's
// contribution to dW / db.
__global__ void __launch_bounds__(256) scorer_backward_q_kernel(const float* __restrict__ q, int Q,
                                                                const float* __restrict__ g, int G,
                                                                const float* __restrict__ last_w,
                                                                const float* __restrict__ dx5, float* __restrict__ dq,
                                                                float* __restrict__ dlast_w, float* __restrict__ dlast_b) {
  const int i = blockIdx.x, k = threadIdx.x;
  const float qv = q[(size_t)i * 256 + k], w0 = last_w[k], w1 = last_w[256 + k];
  float dqv = 0.f, dw0 = 0.f, dw1 = 0.f, db0 = 0.f, db1 = 0.f;
  for (int j = 0; j < G; ++j) {
    const float2 e = *reinterpret_cast<const float2*>(dx5 + ((size_t)i * G + j) * 2);
    const float diff = qv - g[(size_t)j * 256 + k];
    dqv = fmaf(2.f * fmaf(e.x, w0, e.y * w1), diff, dqv);
    const float d2 = diff * diff;
    dw0 = fmaf(e.x, d2, dw0);
    dw1 = fmaf(e.y, d2, dw1);
    db0 += e.x;
    db1 += e.y;
  }
  dq[(size_t)i * 256 + k] = dqv;
  atomicAdd(dlast_w + k, dw0);
  atomicAdd(dlast_w + 256 + k, dw1);
  if (k == 0) {
    atomicAdd(dlast_b, db0);
    atomicAdd(dlast_b + 1, db1);
  }
}
// one CTA per gallery row: dg (no atomics)
__global__ void __launch_bounds__(256) scorer_backward_g_kernel(const float* __restrict__ q, int Q,
                                                                const float* __restrict__ g, int G,
                                                                const float* __restrict__ last_w,
                                                                const float* __restrict__ dx5, float* __restrict__ dg) {
  const int j = blockIdx.x, k = threadIdx.x;
  const float gv = g[(size_t)j * 256 + k], w0 = last_w[k], w1 = last_w[256 + k];
  float d = 0.f;
  for (int i = 0; i < Q; ++i) {
    const float2 e = *reinterpret_cast<const float2*>(dx5 + ((size_t)i * G + j) * 2);
    d = fmaf(-2.f * fmaf(e.x, w0, e.y * w1), q[(size_t)i * 256 + k] - gv, d);
  }
  dg[(size_t)j * 256 + k] = d;
}

}  // namespace bwd
}  // namespace seam
