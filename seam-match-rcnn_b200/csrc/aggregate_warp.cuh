// K1a (short tracks): streaming temporal aggregation, one WARP per track.
//
// Replaces the per-track Python loop of TemporalAggregationNLB.forward's seq-branch,
// models/match_head.py:133-154, and the block it calls, models/nlb.py:66-101, for tracks of up
// to TR frames (TR = 4, 10 or 16; longer tracks use the CTA-per-tile kernel in aggregate.cuh).
//
// Every warp owns a private ring of SLOTS track buffers in shared memory and fills it itself:
// lane t issues one 1 KB bulk async copy (cp.async.bulk + mbarrier complete_tx) for frame t of
// the track SLOTS-1 iterations ahead; padded frames of ragged tracks are never read.  There is
// no block-wide synchronisation anywhere: warps drift apart freely, which keeps ~100 KB of
// loads in flight per SM.  A track's frames are pulled into registers once (8 floats per lane
// and frame) and used for both passes:
//   A  four length-256 dots per frame (a, d, b, c of DESIGN.md "K1 algebra"), reduced with
//      transposing butterflies (40 shuffles for 40 values at T = 10),
//   B  lane t: s_t = d_t + c_s + (1/T) sum_j relu(a_t+b_j) c_j ; p = softmax_t(s) ;
//      q_j = (1/T) sum_t p_t relu(a_t+b_j),
//   C  pooled = sum_t p_t x_t and r = sum_j q_j x_j (8 channels per lane).
// Outputs per track: pooled' = pooled + (sum_j q_j) W_W b_g + [T>1] b_W, and r split into
// tf32-exact halves r_hi + r_lo for the tensor-core product M r of K1b (nlb_tc.cuh).
#pragma once
#include <cstdint>
#include <type_traits>
#include "fold.cuh"
#include "sm100_ptx.cuh"

namespace seam {
namespace aggw {

constexpr int D = 256;

struct Params {
  const float* seq;
  const uint8_t* mask;     // (Q, 1+Tmax) or null
  const int32_t* lens;     // (Q) or null
  int Tmax, Q;
  long long frame_stride, track_stride;   // floats
  const float* fold;
  float* pooled;   // (Q,256)  pooled'
  float* r_hi;     // (Q,256)
  float* r_lo;     // (Q,256)
  float* att;      // (Q,Tmax) or null
};

template <int TR>
struct Cfg;
template <>
struct Cfg<4> { static constexpr int NW = 16, SLOTS = 3; };
template <>
struct Cfg<10> { static constexpr int NW = 12, SLOTS = 1; };
template <>
struct Cfg<16> { static constexpr int NW = 8, SLOTS = 1; };

template <int TR>
constexpr size_t smem_bytes() {
  return (size_t)Cfg<TR>::NW * Cfg<TR>::SLOTS * TR * D * 4      // track buffers
         + (size_t)Cfg<TR>::NW * 64 * 4                          // per-warp scalars (a,p,p/b,q/c per frame)
         + (size_t)Cfg<TR>::NW * Cfg<TR>::SLOTS * 8              // mbarriers
         + (size_t)Cfg<TR>::NW * Cfg<TR>::SLOTS * 4;             // track lengths
}

using ptx::treduce;

// Packed fp32 pairs (FFMA2 / FMUL2 on sm_100a): the kernels here are bound by instruction issue, not by
// the fp32 pipe, and a pair instruction does the work of two in one issue slot.  A lane's 8 channels
// of a frame are two 16-byte shared-memory loads = four pairs.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
struct Vec8 {   // this lane's channels [4l,4l+4) and [128+4l,128+4l+4) as four pairs
  u64 a, b, c, d;
};
__device__ __forceinline__ Vec8 load_vec8(const float* row, int lane) {
  const ulonglong2 lo = *reinterpret_cast<const ulonglong2*>(row + 4 * lane);
  const ulonglong2 hi = *reinterpret_cast<const ulonglong2*>(row + 128 + 4 * lane);
  return Vec8{lo.x, lo.y, hi.x, hi.y};
}
__device__ __forceinline__ Vec8 zero_vec8() { return Vec8{0ull, 0ull, 0ull, 0ull}; }
// x . u over the lane's 8 channels
__device__ __forceinline__ float dot8(const Vec8& x, const Vec8& u) {
  u64 acc = mul2(x.a, u.a);
  acc = fma2(x.b, u.b, acc);
  acc = fma2(x.c, u.c, acc);
  acc = fma2(x.d, u.d, acc);
  float lo, hi;
  upk(acc, lo, hi);
  return lo + hi;
}
// acc += s * x   (ss = {s, s})
__device__ __forceinline__ void fma8(Vec8& acc, u64 ss, const Vec8& x) {
  acc.a = fma2(ss, x.a, acc.a);
  acc.b = fma2(ss, x.b, acc.b);
  acc.c = fma2(ss, x.c, acc.c);
  acc.d = fma2(ss, x.d, acc.d);
}
__device__ __forceinline__ void unpack_vec8(const Vec8& v, float4& lo, float4& hi) {
  upk(v.a, lo.x, lo.y);
  upk(v.b, lo.z, lo.w);
  upk(v.c, hi.x, hi.y);
  upk(v.d, hi.z, hi.w);
}
// tf32-exact split: hi keeps the 10 explicit mantissa bits the tensor core reads, lo the rest
__device__ __forceinline__ void split_tf32(const float4& x, float4& hi, float4& lo) {
  hi.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
  hi.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
  hi.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
  hi.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
  lo = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
}

template <int TR>
__global__ void __launch_bounds__(Cfg<TR>::NW * 32, 1) aggregate_warp_kernel(const Params p) {
  constexpr int NW = Cfg<TR>::NW, SLOTS = Cfg<TR>::SLOTS;
  constexpr int NV = TR * 4;                       // scalars per track
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* xbuf = reinterpret_cast<float*>(smem_raw) + (size_t)warp * SLOTS * TR * D;
  float* scal = reinterpret_cast<float*>(smem_raw) + (size_t)NW * SLOTS * TR * D + warp * 64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NW * SLOTS * TR * D * 4 + (size_t)NW * 64 * 4) +
                   warp * SLOTS;
  int* slot_len = reinterpret_cast<int*>(smem_raw + (size_t)NW * SLOTS * TR * D * 4 + (size_t)NW * 64 * 4 +
                                         (size_t)NW * SLOTS * 8) + warp * SLOTS;
  const int Tmax = p.Tmax;

  if (lane == 0) {
    for (int i = 0; i < SLOTS; ++i) ptx::mbar_init(&bars[i], 1);
    ptx::fence_mbar_init();
  }
  __syncwarp();

  const long long stride = (long long)gridDim.x * NW;
  const long long first = (long long)blockIdx.x * NW + warp;

  // A track's length comes from lens[] or from its mask row.  peek() only issues that (dependent, ~1 us)
  // load -- one track ahead of its use, so that its latency hides behind the current track's arithmetic --
  // and issue() turns the loaded word into the length and starts the frame copies into a slot.
  auto peek = [&](long long track) -> int {
    if (track >= p.Q) return 0;
    if (p.lens) return p.lens[track];
    if (p.mask) return lane <= Tmax ? (int)p.mask[(size_t)track * (1 + Tmax) + lane] : 0;
    return 0;
  };
  auto issue = [&](long long track, int slot, int raw) {
    int len = 0;
    if (track < p.Q) {
      if (p.lens) {
        len = raw;
      } else if (p.mask) {
        // first nonzero of the mask row ends the track; row 0 is the dummy (models/match_head.py:136-139)
        const uint32_t b = __ballot_sync(ptx::FULL_MASK, raw != 0);
        const int end = b ? __ffs(b) - 1 : 1 + Tmax;
        len = end - 1;
      } else {
        len = Tmax;
      }
      len = max(0, min(len, Tmax));
    }
    if (lane == 0) {
      slot_len[slot] = len;
      if (len > 0) ptx::mbar_arrive_expect_tx(&bars[slot], (uint32_t)len * (D * 4));
      else ptx::mbar_arrive(&bars[slot]);
    }
    __syncwarp();
    if (lane < len) {
      const float* src = p.seq + (long long)(lane + 1) * p.frame_stride + track * p.track_stride;
      ptx::bulk_load_1d(xbuf + ((size_t)slot * TR + lane) * D, src, D * 4, &bars[slot]);
    }
  };

  // SLOTS == 1: the single buffer is re-armed for the next track as soon as the current one's
  // frames sit in registers, so its copy overlaps the rest of the computation.  The first copies start
  // before the folded vectors are fetched.
  constexpr int AHEAD = SLOTS == 1 ? 1 : SLOTS - 1;       // tracks in flight ahead of the one being processed
#pragma unroll 1
  for (int i = 0; i < AHEAD; ++i) issue(first + i * stride, i, peek(first + i * stride));
  int raw_next = peek(first + (long long)AHEAD * stride);  // for the next issue() below

  // folded vectors at this lane's two float4 positions
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

  int it = 0;
#pragma unroll 1
  for (long long track = first; track < p.Q; track += stride, ++it) {
    const int slot = it % SLOTS;
    const uint32_t phase = (uint32_t)(it / SLOTS) & 1u;
    if constexpr (SLOTS > 1) {
      __syncwarp();                                  // every lane is done with the slot being refilled
      issue(track + (long long)(SLOTS - 1) * stride, (it + SLOTS - 1) % SLOTS, raw_next);
      raw_next = peek(track + (long long)SLOTS * stride);
    }
    ptx::mbar_wait(&bars[slot], phase);
    const int len = slot_len[slot];
    const float* xs = xbuf + (size_t)slot * TR * D;

    // One body, two instantiations: tracks with all TR frames (the common case) run without
    // per-frame guards and with unrolled T x T loops.
    auto process = [&](auto full_tag) {
      constexpr bool FULL = decltype(full_tag)::value;
      // ---- frames -> registers, four dots per frame
      Vec8 x[TR];
      constexpr int N1 = NV <= 16 ? 16 : 32;                       // first butterfly: frames 0..7
      constexpr int N2 = NV <= 32 ? 1 : (NV - 32 <= 8 ? 8 : 32);    // second butterfly: frames 8..
      float acc[N1], acc2[N2];
  #pragma unroll
      for (int i = 0; i < N1; ++i) acc[i] = 0.f;
  #pragma unroll
      for (int i = 0; i < N2; ++i) acc2[i] = 0.f;
  #pragma unroll
      for (int t = 0; t < TR; ++t) {
        if (FULL || t < len) x[t] = load_vec8(xs + t * D, lane);
        else x[t] = zero_vec8();
        const float va = dot8(x[t], ut);
        const float vd = dot8(x[t], wa);
        const float vb = dot8(x[t], up);
        const float vc = dot8(x[t], ug);
        if (4 * t < 32) {
          acc[4 * t + 0] = va;
          acc[4 * t + 1] = vd;
          acc[4 * t + 2] = vb;
          acc[4 * t + 3] = vc;
        } else {
          acc2[4 * t - 32 + 0] = va;
          acc2[4 * t - 32 + 1] = vd;
          acc2[4 * t - 32 + 2] = vb;
          acc2[4 * t - 32 + 3] = vc;
        }
      }
      if constexpr (SLOTS == 1) {
        __syncwarp();                                // all lanes hold their frames in registers
        issue(track + stride, 0, raw_next);
        raw_next = peek(track + 2 * stride);
      }
      if constexpr (NV <= 16) {
        const float tot = treduce<16>(acc, lane);
        if (lane < 16) scal[lane] = tot + my_const;
      } else {
        const float tot = treduce<32>(acc, lane);
        scal[lane] = tot + my_const;
        if constexpr (NV > 32) {
          if constexpr (N2 == 8) {
            const float tot2 = treduce<8>(acc2, lane);
            if (lane < 8) scal[32 + lane] = tot2 + my_const;
          } else {
            const float tot2 = treduce<32>(acc2, lane);
            scal[32 + lane] = tot2 + my_const;
          }
        }
      }
      __syncwarp();

      // ---- attention over the track's frames (lane = frame)
      const int L = FULL ? TR : len;                   // compile-time trip counts for full tracks
      const bool valid = lane < L;
      const float inv_len = FULL ? 1.f / (float)TR : (len > 0 ? 1.f / (float)len : 0.f);
      float4 sc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) sc = *reinterpret_cast<const float4*>(scal + 4 * lane);   // a, d, b, c of my frame
      float sum = 0.f;
      if (FULL ? TR > 1 : len > 1) {
  #pragma unroll
        for (int j = 0; j < (FULL ? TR : len); ++j) {
          const float2 bc = *reinterpret_cast<const float2*>(scal + 4 * j + 2);
          sum = fmaf(fmaxf(sc.x + bc.x, 0.f) * inv_len, bc.y, sum);
        }
      }
      const float s_t = valid ? sc.y + sum + c_s : -INFINITY;
      const float m = ptx::warp_max(s_t);
      const float e_t = valid ? expf(s_t - m) : 0.f;
      const float z = ptx::warp_sum(e_t);
      const float p_t = valid ? e_t / z : 0.f;
      __syncwarp();                                    // all lanes have read b, c, d
      if (lane < TR) {
        scal[4 * lane + 1] = p_t;
        scal[4 * lane + 2] = p_t;
      }
      __syncwarp();
      float q_j = 0.f;
      if ((FULL ? TR > 1 : len > 1) && valid) {
  #pragma unroll
        for (int t = 0; t < (FULL ? TR : len); ++t) {
          const float2 ap = *reinterpret_cast<const float2*>(scal + 4 * t);
          q_j = fmaf(ap.y, fmaxf(ap.x + sc.z, 0.f) * inv_len, q_j);
        }
      }
      const float qsum = ptx::warp_sum(q_j);
      __syncwarp();                                    // all lanes have read the a_t
      if (lane < TR) *reinterpret_cast<float4*>(scal + 4 * lane) = make_float4(p_t, p_t, q_j, q_j);
      if (p.att && lane < Tmax) p.att[(size_t)track * Tmax + lane] = p_t;
      __syncwarp();

      // ---- weighted sums over frames, 8 channels per lane
      Vec8 pov = zero_vec8(), rv = zero_vec8();
  #pragma unroll
      for (int t = 0; t < TR; ++t) {
        if (FULL || t < len) {
          const ulonglong2 pq = *reinterpret_cast<const ulonglong2*>(scal + 4 * t);   // {p,p}, {q,q}
          fma8(pov, pq.x, x[t]);
          fma8(rv, pq.y, x[t]);
        }
      }
      float4 po0, po1, r0, r1;
      unpack_vec8(pov, po0, po1);
      unpack_vec8(rv, r0, r1);
      if (FULL ? TR > 1 : len > 1) {
        const float4 wbg0 = *reinterpret_cast<const float4*>(fold + Fold::WBG + 4 * lane);
        const float4 wbg1 = *reinterpret_cast<const float4*>(fold + Fold::WBG + 128 + 4 * lane);
        const float4 bw0 = *reinterpret_cast<const float4*>(fold + Fold::BW + 4 * lane);
        const float4 bw1 = *reinterpret_cast<const float4*>(fold + Fold::BW + 128 + 4 * lane);
        po0.x += fmaf(qsum, wbg0.x, bw0.x);
        po0.y += fmaf(qsum, wbg0.y, bw0.y);
        po0.z += fmaf(qsum, wbg0.z, bw0.z);
        po0.w += fmaf(qsum, wbg0.w, bw0.w);
        po1.x += fmaf(qsum, wbg1.x, bw1.x);
        po1.y += fmaf(qsum, wbg1.y, bw1.y);
        po1.z += fmaf(qsum, wbg1.z, bw1.z);
        po1.w += fmaf(qsum, wbg1.w, bw1.w);
      }
      float4 h0, l0, h1, l1;
      split_tf32(r0, h0, l0);
      split_tf32(r1, h1, l1);
      const size_t o = (size_t)track * D + 4 * lane;
      *reinterpret_cast<float4*>(p.pooled + o) = po0;
      *reinterpret_cast<float4*>(p.pooled + o + 128) = po1;
      *reinterpret_cast<float4*>(p.r_hi + o) = h0;
      *reinterpret_cast<float4*>(p.r_hi + o + 128) = h1;
      *reinterpret_cast<float4*>(p.r_lo + o) = l0;
      *reinterpret_cast<float4*>(p.r_lo + o + 128) = l1;
    };
    if (len == TR) process(std::true_type{});
    else process(std::false_type{});
  }
}

}  // namespace aggw
}  // namespace seam
