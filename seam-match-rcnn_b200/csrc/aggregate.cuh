// K1a: streaming temporal aggregation (non-local block + frame-attention pooling).
//
// Replaces the per-track Python loop of TemporalAggregationNLB.forward's seq-branch,
// models/match_head.py:133-154, and the block it calls, models/nlb.py:66-101.
//
// One persistent CTA per SM.  A producer warp streams the frames of NT consecutive tracks
// (<= 64 rows of 1 KB) per pipeline stage into shared memory with 1-D bulk async copies
// (cp.async.bulk + mbarrier complete_tx); padded frames of ragged tracks are never read.
// Eight consumer warps then compute, per tile,
//   A  four length-256 dots per frame (a,b,c,d of DESIGN.md "K1 algebra"), reduced with a
//      transposing butterfly (31 shuffles for 32 values instead of 160),
//   B  s_t = d_t + c_s + (1/T) sum_j relu(a_t+b_j) c_j ; p = softmax_t(s) ;
//      q_j = (1/T) sum_t p_t relu(a_t+b_j),
//   C  pooled = sum_t p_t x_t and r = sum_j q_j x_j (one thread per channel).
// Each frame is read from HBM exactly once.  The per-track 256->256 map M r (the NLB's
// g/W projections applied to r) is batched over tracks by K1b (nlb_tc.cuh).  This kernel serves
// tracks longer than 16 frames; shorter ones take the warp-per-track kernel (aggregate_warp.cuh).
#pragma once
#include <cstdint>
#include "fold.cuh"
#include "sm100_ptx.cuh"

namespace seam {
namespace agg {

constexpr int D = 256;
constexpr int ROWS_MAX = 64;     // frames per pipeline stage
constexpr int STAGES = 3;
constexpr int CONS_WARPS = 8;
constexpr int CONS_THREADS = CONS_WARPS * 32;
constexpr int THREADS = 32 + CONS_THREADS;
constexpr int MAX_NT = 32;       // tracks per tile (one producer lane per track)

struct Params {
  const float* seq;
  const uint8_t* mask;     // (Q, 1+Tmax) or null
  const int32_t* lens;     // (Q) or null
  int Tmax, Q, NT, rows, num_tiles;
  long long frame_stride, track_stride;   // floats
  const float* fold;
  float* pooled;   // (Q,256)  pooled' = sum_t p_t x_t + (sum_j q_j) W_W b_g + [T>1] b_W
  float* r_hi;     // (Q,256)  r = sum_j q_j x_j split into tf32-exact halves (nlb_tc.cuh)
  float* r_lo;     // (Q,256)
  float* att;      // (Q,Tmax) or null
};

struct Smem {
  float x[STAGES][ROWS_MAX][D];        // 192 KB
  float scal[ROWS_MAX][4];             // a, b, c, d per frame
  float sbuf[ROWS_MAX];
  float pbuf[ROWS_MAX];
  float qbuf[ROWS_MAX];
  int lens[STAGES][MAX_NT];
  uint8_t row_n[ROWS_MAX];             // row -> track within tile
  uint8_t row_t[ROWS_MAX];             // row -> frame
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
};

// sum over the 32 lanes of 32 per-lane values; lane l returns the total of v[l]
__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? v[i] : v[i + half];
      const float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(ptx::FULL_MASK, send, half);
    }
  }
  return v[0];
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

__global__ void __launch_bounds__(THREADS, 1) aggregate_kernel(const Params p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Tmax = p.Tmax, NT = p.NT, rows = p.rows;

  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&s.full[i], 1);
      ptx::mbar_init(&s.empty[i], CONS_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (tid < ROWS_MAX) {
    s.row_n[tid] = (uint8_t)(tid / Tmax);
    s.row_t[tid] = (uint8_t)(tid % Tmax);
  }
  __syncthreads();

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int stage = it % STAGES;
      const uint32_t phase = (it / STAGES) & 1;
      ptx::mbar_wait(&s.empty[stage], phase ^ 1);
      const int track = tile * NT + lane;
      int len = 0;
      if (lane < NT && track < p.Q) {
        if (p.lens) {
          len = p.lens[track];
        } else if (p.mask) {
          // first nonzero of the mask row ends the track; row 0 is the dummy
          const uint8_t* m = p.mask + (size_t)track * (1 + Tmax);
          int end = 1 + Tmax;
          for (int j = 0; j <= Tmax; ++j)
            if (m[j]) { end = j; break; }
          len = end - 1;
        } else {
          len = Tmax;
        }
        len = max(0, min(len, Tmax));
      }
      if (lane < NT) s.lens[stage][lane] = len;
      int total = len;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(ptx::FULL_MASK, total, o);
      __syncwarp();
      if (lane == 0) {
        if (total > 0) ptx::mbar_arrive_expect_tx(&s.full[stage], (uint32_t)total * (D * 4));
        else ptx::mbar_arrive(&s.full[stage]);
      }
      __syncwarp();
      for (int r0 = 0; r0 < rows; r0 += 32) {
        const int r = r0 + lane;
        const int n = r < rows ? s.row_n[r] : 0;
        const int t = r < rows ? s.row_t[r] : 0;
        const int ln = __shfl_sync(ptx::FULL_MASK, len, n);
        if (r < rows && t < ln) {
          const float* src = p.seq + (long long)(t + 1) * p.frame_stride + (long long)(tile * NT + n) * p.track_stride;
          ptx::bulk_load_1d(&s.x[stage][r][0], src, D * 4, &s.full[stage]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ consumers
    const int cw = warp - 1;
    const int ctid = tid - 32;
    const float* fold = p.fold;
    // folded vectors at this lane's two float4 positions
    float4 ut0 = *reinterpret_cast<const float4*>(fold + Fold::U_THETA + 4 * lane);
    float4 ut1 = *reinterpret_cast<const float4*>(fold + Fold::U_THETA + 128 + 4 * lane);
    float4 up0 = *reinterpret_cast<const float4*>(fold + Fold::U_PHI + 4 * lane);
    float4 up1 = *reinterpret_cast<const float4*>(fold + Fold::U_PHI + 128 + 4 * lane);
    float4 ug0 = *reinterpret_cast<const float4*>(fold + Fold::U_G + 4 * lane);
    float4 ug1 = *reinterpret_cast<const float4*>(fold + Fold::U_G + 128 + 4 * lane);
    float4 wa0 = *reinterpret_cast<const float4*>(fold + Fold::W_A + 4 * lane);
    float4 wa1 = *reinterpret_cast<const float4*>(fold + Fold::W_A + 128 + 4 * lane);
    const float c_s = fold[Fold::CONSTS + 3];
    const float my_const = (lane & 3) < 3 ? fold[Fold::CONSTS + (lane & 3)] : 0.f;

    const int item = ctid >> 2, sub = ctid & 3;
    const int item_n = s.row_n[item], item_t = s.row_t[item];
    const int item_base = item_n * Tmax;

    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int stage = it % STAGES;
      const uint32_t phase = (it / STAGES) & 1;
      ptx::mbar_wait(&s.full[stage], phase);
      const float* xs = &s.x[stage][0][0];

      // ---- A: four dots per frame
      {
        float acc[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = cw * 8 + i;
          float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
          if (row < rows && s.row_t[row] < s.lens[stage][s.row_n[row]]) {
            x0 = *reinterpret_cast<const float4*>(xs + row * D + 4 * lane);
            x1 = *reinterpret_cast<const float4*>(xs + row * D + 128 + 4 * lane);
          }
          acc[4 * i + 0] = dot4(x0, ut0) + dot4(x1, ut1);
          acc[4 * i + 1] = dot4(x0, up0) + dot4(x1, up1);
          acc[4 * i + 2] = dot4(x0, ug0) + dot4(x1, ug1);
          acc[4 * i + 3] = dot4(x0, wa0) + dot4(x1, wa1);
        }
        const float tot = transpose_reduce32(acc, lane);
        s.scal[cw * 8 + (lane >> 2)][lane & 3] = tot + my_const;
      }
      ptx::named_bar_sync(1, CONS_THREADS);

      // ---- B1: attention logits s_t
      const int len = (item < rows) ? s.lens[stage][item_n] : 0;
      const bool valid = item < rows && item_t < len;
      const float inv_len = len > 0 ? 1.f / (float)len : 0.f;
      {
        float sum = 0.f;
        if (valid && len > 1) {
          const float a_t = s.scal[item][0];
          for (int j = sub; j < len; j += 4) {
            const float f = fmaxf(a_t + s.scal[item_base + j][1], 0.f) * inv_len;
            sum = fmaf(f, s.scal[item_base + j][2], sum);
          }
        }
        sum += __shfl_xor_sync(ptx::FULL_MASK, sum, 1);
        sum += __shfl_xor_sync(ptx::FULL_MASK, sum, 2);
        if (valid && sub == 0) s.sbuf[item] = s.scal[item][3] + sum + c_s;
      }
      ptx::named_bar_sync(1, CONS_THREADS);

      // ---- B2: softmax over the track's frames
      {
        float m = -INFINITY;
        if (valid)
          for (int j = sub; j < len; j += 4) m = fmaxf(m, s.sbuf[item_base + j]);
        m = fmaxf(m, __shfl_xor_sync(ptx::FULL_MASK, m, 1));
        m = fmaxf(m, __shfl_xor_sync(ptx::FULL_MASK, m, 2));
        float z = 0.f;
        if (valid)
          for (int j = sub; j < len; j += 4) z += expf(s.sbuf[item_base + j] - m);
        z += __shfl_xor_sync(ptx::FULL_MASK, z, 1);
        z += __shfl_xor_sync(ptx::FULL_MASK, z, 2);
        const float pt = valid ? expf(s.sbuf[item] - m) / z : 0.f;
        if (item < rows && sub == 0) {
          s.pbuf[item] = pt;
          const int track = tile * NT + item_n;
          if (p.att && track < p.Q) p.att[(size_t)track * Tmax + item_t] = pt;
        }
      }
      ptx::named_bar_sync(1, CONS_THREADS);

      // ---- B3: q_j = (1/T) sum_t p_t relu(a_t + b_j)
      {
        float sum = 0.f;
        if (valid && len > 1) {
          const float b_j = s.scal[item][1];
          for (int t = sub; t < len; t += 4)
            sum = fmaf(s.pbuf[item_base + t], fmaxf(s.scal[item_base + t][0] + b_j, 0.f) * inv_len, sum);
        }
        sum += __shfl_xor_sync(ptx::FULL_MASK, sum, 1);
        sum += __shfl_xor_sync(ptx::FULL_MASK, sum, 2);
        if (item < rows && sub == 0) s.qbuf[item] = sum;
      }
      ptx::named_bar_sync(1, CONS_THREADS);

      // ---- C: weighted sums, one thread per channel
      {
        const int c = ctid;
        const float wbg = fold[Fold::WBG + c], bW = fold[Fold::BW + c];
        for (int n = 0; n < NT; ++n) {
          const int track = tile * NT + n;
          if (track >= p.Q) break;
          const int ln = s.lens[stage][n];
          const float* xr = xs + (n * Tmax) * D + c;
          float pooled = 0.f, r = 0.f, qs = 0.f;
          for (int t = 0; t < ln; ++t) {
            const float xv = xr[t * D];
            const float pt = s.pbuf[n * Tmax + t], qt = s.qbuf[n * Tmax + t];
            pooled = fmaf(pt, xv, pooled);
            r = fmaf(qt, xv, r);
            qs += qt;
          }
          if (ln > 1) pooled += fmaf(qs, wbg, bW);
          const float hi = __uint_as_float(__float_as_uint(r) & 0xffffe000u);   // tf32-exact split
          p.pooled[(size_t)track * D + c] = pooled;
          p.r_hi[(size_t)track * D + c] = hi;
          p.r_lo[(size_t)track * D + c] = r - hi;
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&s.empty[stage]);
    }
  }
}

}  // namespace agg
}  // namespace seam
