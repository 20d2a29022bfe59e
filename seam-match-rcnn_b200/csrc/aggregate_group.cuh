// K1a (long tracks, 17..64 frames): streaming temporal aggregation, one GROUP OF FOUR WARPS per
// track.
//
// Same algebra and outputs as aggregate_warp.cuh (models/match_head.py:133-154, models/nlb.py:66-101
// collapsed as in DESIGN.md "K1 algebra"); a track is cut into four blocks of 16 frames and warp w of
// the group keeps frames [16w, 16w+16) in registers, exactly like the warp-per-track kernel does for a
// whole short track: private 16 KB buffer filled by 1 KB bulk async copies (lane t fetches frame t of
// the block), re-armed for the group's next track as soon as the frames sit in registers.  What
// crosses warps goes through a few hundred bytes of shared memory and six 128-thread named barriers
// per track: the per-frame scalars (a, d, b, c), the softmax maximum and denominator, and the partial
// weighted sums.  Eight warps per CTA (two groups of four, or four groups of two for tracks of up
// to 32 frames), one CTA per SM.
#pragma once
#include <cstdint>
#include "aggregate_warp.cuh"
#include "fold.cuh"
#include "sm100_ptx.cuh"

namespace seam {
namespace aggg {

constexpr int D = 256;
constexpr int FB = 16;                 // frames per warp
constexpr int WARPS = 8;               // per CTA (register-limited: 255 registers per thread)
constexpr int THREADS = WARPS * 32;

// GW = warps per group: 2 for tracks of up to 32 frames, 4 for up to 64
template <int GW>
struct alignas(16) GroupSmem {
  float scal[FB * GW][4];              // per frame: a, d -> p, b, c -> q
  float part[GW][2 * D];               // per warp: partial pooled | partial r
  float red_max[GW];
  float red_sum[GW];
  float red_q[GW];
  float pad[4];
};
template <int GW>
struct Smem {
  float x[WARPS][FB][D];               // 8 x 16 KB
  GroupSmem<GW> g[WARPS / GW];
  uint64_t bar[WARPS];
};

using aggw::Params;
using aggw::Vec8;
using aggw::dot8;
using aggw::fma8;
using aggw::load_vec8;
using aggw::zero_vec8;
using aggw::unpack_vec8;
using aggw::pk;
using ptx::treduce;

template <int GW>
__global__ void __launch_bounds__(THREADS, 1) aggregate_group_kernel(const Params p) {
  constexpr int GROUPS_PER_CTA = WARPS / GW;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem<GW>& s = *reinterpret_cast<Smem<GW>*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = warp / GW, wg = warp % GW;
  GroupSmem<GW>& gs = s.g[grp];
  float* xs = &s.x[warp][0][0];
  uint64_t* bar = &s.bar[warp];
  const int Tmax = p.Tmax;
  const uint32_t bar_id = 1 + grp;

  if (lane == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_mbar_init();
  }
  __syncwarp();

  const long long stride = (long long)gridDim.x * GROUPS_PER_CTA;
  const long long first = (long long)blockIdx.x * GROUPS_PER_CTA + grp;

  // length of a track (every warp of the group derives it on its own)
  auto track_len = [&](long long track) -> int {
    if (track >= p.Q) return 0;
    int len;
    if (p.lens) {
      len = p.lens[track];
    } else if (p.mask) {
      // first nonzero of the mask row ends the track; row 0 is the dummy (models/match_head.py:136-139)
      const uint8_t* m = p.mask + (size_t)track * (1 + Tmax);
      const uint32_t b0 = __ballot_sync(ptx::FULL_MASK, lane <= Tmax && m[lane] != 0);
      const uint32_t b1 = __ballot_sync(ptx::FULL_MASK, 32 + lane <= Tmax && m[min(32 + lane, Tmax)] != 0);
      const uint32_t b2 = __ballot_sync(ptx::FULL_MASK, lane == 0 && Tmax >= 64 && m[min(64, Tmax)] != 0);
      const int end = b0 ? __ffs(b0) - 1 : b1 ? 32 + __ffs(b1) - 1 : b2 ? 64 : 1 + Tmax;
      len = end - 1;
    } else {
      len = Tmax;
    }
    return max(0, min(len, Tmax));
  };
  // start the copies of this warp's frame block of one track
  auto issue = [&](long long track, int len) {
    const int nw = max(0, min(len - FB * wg, FB));
    if (lane == 0) {
      if (nw > 0) ptx::mbar_arrive_expect_tx(bar, (uint32_t)nw * (D * 4));
      else ptx::mbar_arrive(bar);
    }
    __syncwarp();
    if (lane < nw) {
      const float* src = p.seq + (long long)(FB * wg + lane + 1) * p.frame_stride + track * p.track_stride;
      ptx::bulk_load_1d(xs + (size_t)lane * D, src, D * 4, bar);
    }
  };

  const float* fold = p.fold;
  const Vec8 ut = load_vec8(fold + Fold::U_THETA, lane);
  const Vec8 up = load_vec8(fold + Fold::U_PHI, lane);
  const Vec8 ug = load_vec8(fold + Fold::U_G, lane);
  const Vec8 wa = load_vec8(fold + Fold::W_A, lane);
  const float c_s = fold[Fold::CONSTS + 3];
  // scalar layout per frame: [a, d, b, c]; constants c_theta, 0, c_phi, c_g
  const int comp = lane & 3;
  const float my_const = comp == 0 ? fold[Fold::CONSTS + 0] : comp == 2 ? fold[Fold::CONSTS + 1]
                       : comp == 3 ? fold[Fold::CONSTS + 2] : 0.f;

  int len = track_len(first);
  issue(first, len);
  int it = 0;
#pragma unroll 1
  for (long long track = first; track < p.Q; track += stride, ++it) {
    ptx::mbar_wait(bar, (uint32_t)it & 1u);
    const int nw = max(0, min(len - FB * wg, FB));

    // ---- my 16 frames -> registers, four dots per frame
    Vec8 x[FB];
    float acc[32], acc2[32];
#pragma unroll
    for (int t = 0; t < FB; ++t) {
      if (t < nw) x[t] = load_vec8(xs + t * D, lane);
      else x[t] = zero_vec8();
      const float va = dot8(x[t], ut);
      const float vd = dot8(x[t], wa);
      const float vb = dot8(x[t], up);
      const float vc = dot8(x[t], ug);
      if (t < 8) {
        acc[4 * t + 0] = va;
        acc[4 * t + 1] = vd;
        acc[4 * t + 2] = vb;
        acc[4 * t + 3] = vc;
      } else {
        acc2[4 * (t - 8) + 0] = va;
        acc2[4 * (t - 8) + 1] = vd;
        acc2[4 * (t - 8) + 2] = vb;
        acc2[4 * (t - 8) + 3] = vc;
      }
    }
    // the buffer is free again: fetch this warp's block of the group's next track
    const int len_next = track_len(track + stride);
    __syncwarp();
    issue(track + stride, len_next);

    {
      const float tot = treduce<32>(acc, lane);
      const float tot2 = treduce<32>(acc2, lane);
      gs.scal[FB * wg + (lane >> 2)][comp] = tot + my_const;
      gs.scal[FB * wg + 8 + (lane >> 2)][comp] = tot2 + my_const;
    }
    ptx::named_bar_sync(bar_id, GW * 32);                       // #1 all scalars of the track are visible

    // ---- attention over the track's frames: lane = (frame f of my block, half of the j / t range)
    const int f = lane & 15, half = lane >> 4;
    const int F = FB * wg + f;
    const bool valid = F < len;
    const float inv_len = len > 0 ? 1.f / (float)len : 0.f;
    float4 sc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) sc = *reinterpret_cast<const float4*>(&gs.scal[F][0]);   // a, d, b, c of my frame
    float sum = 0.f;
    if (len > 1) {
      for (int j = half; j < len; j += 2) {
        const float2 bc = *reinterpret_cast<const float2*>(&gs.scal[j][2]);
        sum = fmaf(fmaxf(sc.x + bc.x, 0.f) * inv_len, bc.y, sum);
      }
    }
    sum += __shfl_xor_sync(ptx::FULL_MASK, sum, 16);
    const float s_t = valid ? sc.y + sum + c_s : -INFINITY;
    const float m_w = ptx::warp_max(s_t);
    if (lane == 0) gs.red_max[wg] = m_w;
    ptx::named_bar_sync(bar_id, GW * 32);                       // #2 (also: nobody reads b, c, d any more)
    float m = gs.red_max[0];
#pragma unroll
    for (int w = 1; w < GW; ++w) m = fmaxf(m, gs.red_max[w]);
    const float e_t = valid ? expf(s_t - m) : 0.f;
    const float z_w = ptx::warp_sum(half == 0 ? e_t : 0.f);
    if (lane == 0) gs.red_sum[wg] = z_w;
    ptx::named_bar_sync(bar_id, GW * 32);                       // #3
    float z = gs.red_sum[0];
#pragma unroll
    for (int w = 1; w < GW; ++w) z += gs.red_sum[w];
    const float p_t = valid ? e_t / z : 0.f;
    if (half == 0) gs.scal[F][1] = p_t;
    if (p.att && half == 0 && F < Tmax) p.att[(size_t)track * Tmax + F] = p_t;
    ptx::named_bar_sync(bar_id, GW * 32);                       // #4 all p_t are visible
    float q_j = 0.f;
    if (len > 1 && valid) {
      for (int t = half; t < len; t += 2) {
        const float2 ap = *reinterpret_cast<const float2*>(&gs.scal[t][0]);
        q_j = fmaf(ap.y, fmaxf(ap.x + sc.z, 0.f) * inv_len, q_j);
      }
    }
    q_j += __shfl_xor_sync(ptx::FULL_MASK, q_j, 16);
    const float qsum_w = ptx::warp_sum(half == 0 ? q_j : 0.f);

    // ---- partial weighted sums over my frames, 8 channels per lane
    Vec8 pov = zero_vec8(), rv = zero_vec8();
#pragma unroll
    for (int t = 0; t < FB; ++t) {
      const float pt = __shfl_sync(ptx::FULL_MASK, p_t, t);
      const float qt = __shfl_sync(ptx::FULL_MASK, q_j, t);
      if (t < nw) {
        fma8(pov, pk(pt, pt), x[t]);
        fma8(rv, pk(qt, qt), x[t]);
      }
    }
    float4 po0, po1, r0, r1;
    unpack_vec8(pov, po0, po1);
    unpack_vec8(rv, r0, r1);
    {
      float* pw = &gs.part[wg][0];
      *reinterpret_cast<float4*>(pw + 4 * lane) = po0;
      *reinterpret_cast<float4*>(pw + 128 + 4 * lane) = po1;
      *reinterpret_cast<float4*>(pw + D + 4 * lane) = r0;
      *reinterpret_cast<float4*>(pw + D + 128 + 4 * lane) = r1;
      if (lane == 0) gs.red_q[wg] = qsum_w;
    }
    ptx::named_bar_sync(bar_id, GW * 32);                       // #5 partial sums are visible

    // ---- warp w finishes channels [CW w, CW (w+1)), CW = 256 / GW: CW / 32 per lane in pairs
    {
      constexpr int CW = D / GW;
#pragma unroll
      for (int c0 = 0; c0 < CW; c0 += 64) {
        const int c = CW * wg + c0 + 2 * lane;
        float2 po = make_float2(0.f, 0.f), rr = po;
#pragma unroll
        for (int w = 0; w < GW; ++w) {
          const float2 a = *reinterpret_cast<const float2*>(&gs.part[w][c]);
          const float2 b = *reinterpret_cast<const float2*>(&gs.part[w][D + c]);
          po.x += a.x;
          po.y += a.y;
          rr.x += b.x;
          rr.y += b.y;
        }
        if (len > 1) {
          float qsum = gs.red_q[0];
#pragma unroll
          for (int w = 1; w < GW; ++w) qsum += gs.red_q[w];
          const float2 wbg = *reinterpret_cast<const float2*>(fold + Fold::WBG + c);
          const float2 bw = *reinterpret_cast<const float2*>(fold + Fold::BW + c);
          po.x += fmaf(qsum, wbg.x, bw.x);
          po.y += fmaf(qsum, wbg.y, bw.y);
        }
        float2 hi, lo;
        hi.x = __uint_as_float(__float_as_uint(rr.x) & 0xffffe000u);   // tf32-exact split
        hi.y = __uint_as_float(__float_as_uint(rr.y) & 0xffffe000u);
        lo.x = rr.x - hi.x;
        lo.y = rr.y - hi.y;
        const size_t o = (size_t)track * D + c;
        *reinterpret_cast<float2*>(p.pooled + o) = po;
        *reinterpret_cast<float2*>(p.r_hi + o) = hi;
        *reinterpret_cast<float2*>(p.r_lo + o) = lo;
      }
    }
    len = len_next;
  }
}

}  // namespace aggg
}  // namespace seam
