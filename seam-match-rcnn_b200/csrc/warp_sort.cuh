// Warp-level bitonic sorting networks (one element per lane) used by the top-k stages.
#pragma once
#include <cstdint>
#include "sm100_ptx.cuh"

namespace seam {
namespace wsort {

// ---- plain float key + 32-bit payload ------------------------------------------------
// One compare-exchange = 2 shuffles, one min/max whose direction is a predicate (FMNMX), one compare
// and one select for the payload.  Equal keys keep their own payloads on both sides (no duplicates);
// keys are never NaN here.
template <bool DESC>
__device__ __forceinline__ void cmpx(float& v, uint32_t& p, int lane, int j, int k) {
  const float ov = __shfl_xor_sync(ptx::FULL_MASK, v, j);
  const uint32_t op = __shfl_xor_sync(ptx::FULL_MASK, p, j);
  const bool keep_min = ((((lane & k) == 0) == ((lane & j) == 0)) != DESC);
  const float nv = keep_min ? fminf(v, ov) : fmaxf(v, ov);
  if (nv != v) p = op;
  v = nv;
}
template <bool DESC>
__device__ __forceinline__ void sort32(float& v, uint32_t& p, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) cmpx<DESC>(v, p, lane, j, k);
}
// input: a bitonic sequence across the lanes
template <bool DESC>
__device__ __forceinline__ void merge32(float& v, uint32_t& p, int lane) {
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) cmpx<DESC>(v, p, lane, j, 32);
}

// ---- ranking order: margin descending, ties by lowest index; float payload -------------
// "a ranks before b"
__device__ __forceinline__ bool ranks_before(float da, int ia, float db, int ib) {
  return da > db || (da == db && ia < ib);
}
__device__ __forceinline__ void cmpx_rank(float& d, int& i, float& s, int lane, int j, int k) {
  const float od = __shfl_xor_sync(ptx::FULL_MASK, d, j);
  const int oi = __shfl_xor_sync(ptx::FULL_MASK, i, j);
  const float os = __shfl_xor_sync(ptx::FULL_MASK, s, j);
  const bool up = (lane & k) == 0;
  const bool lower = (lane & j) == 0;
  const bool keep_better = (lower == up);
  const bool take = keep_better ? ranks_before(od, oi, d, i) : ranks_before(d, i, od, oi);
  if (take) {
    d = od;
    i = oi;
    s = os;
  }
}
// same order, no payload
__device__ __forceinline__ void cmpx_rank2(float& d, int& i, int lane, int j, int k) {
  const float od = __shfl_xor_sync(ptx::FULL_MASK, d, j);
  const int oi = __shfl_xor_sync(ptx::FULL_MASK, i, j);
  const bool up = (lane & k) == 0;
  const bool lower = (lane & j) == 0;
  const bool keep_better = (lower == up);
  const bool take = keep_better ? ranks_before(od, oi, d, i) : ranks_before(d, i, od, oi);
  if (take) {
    d = od;
    i = oi;
  }
}
// The ranking order as one 64-bit key: order-preserving image of d in the high word, ~i in the low
// word (lower index = larger key); "ranks before" = larger key.  Keys are unique except for padding
// entries (d = -inf, i = INT_MAX), for which taking the partner's identical key is harmless.
__device__ __forceinline__ void sort32_rank2(float& d, int& i, int lane) {
  uint32_t hi = ptx::float_to_ordered(d), lo = ~(uint32_t)i;
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const uint32_t ohi = __shfl_xor_sync(ptx::FULL_MASK, hi, j);
      const uint32_t olo = __shfl_xor_sync(ptx::FULL_MASK, lo, j);
      const bool keep_better = ((lane & k) == 0) == ((lane & j) == 0);
      const bool other_better = ohi > hi || (ohi == hi && olo > lo);
      if (other_better == keep_better) {
        hi = ohi;
        lo = olo;
      }
    }
  d = ptx::ordered_to_float(hi);
  i = (int)~lo;
}
// best first
__device__ __forceinline__ void sort32_rank(float& d, int& i, float& s, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) cmpx_rank(d, i, s, lane, j, k);
}
__device__ __forceinline__ void merge32_rank(float& d, int& i, float& s, int lane) {
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) cmpx_rank(d, i, s, lane, j, 32);
}

}  // namespace wsort
}  // namespace seam
