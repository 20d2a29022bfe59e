// Gallery-sharded search across the GPUs of one NVSwitch box: what the kernels need to exchange data
// THEMSELVES.  Each rank scores every query against its gallery shard; two things have to cross GPUs per
// step -- the aggregated descriptors (every rank needs all Q of them) and the per-shard top-k lists (they
// meet at the rank that OWNS the query, which merges them) -- plus, when every rank wants the whole
// result, the merged rows.  All three are written by the producing kernel straight into the consumers'
// memory (peer-mapped symmetric buffers: plain stores over NVLink), and ordered by flag words instead of
// barriers or collectives:
//   producer grid:  stores ... ; every CTA: __threadfence (device scope) + atomic count; the LAST CTA, having
//                   observed every other CTA's count, issues ONE system-scope fence and writes
//                   flags[kind][my rank] = step on every rank (relaxed stores after it).  Causality is transitive across
//                   the two scopes, so the peers' acquire of the flag covers every CTA's stores -- a
//                   system-scope fence in each of a few hundred CTAs cost ~15 us per kernel
//   consumer grid:  first thing, CTA-wide: spin (ld.acquire.sys) until flags[kind][r] >= step for all r
// The step number lives in device memory and is advanced by the last kernel of a step, so a whole
// step replays as one CUDA graph.  Descriptor and list buffers are double-buffered by step parity: a
// rank can only get one step ahead of its peers (it needs their lists to finish a step), so writing
// buffer (s+2)&1 can never hit data a peer still reads for step s -- no barrier at the top of a step.
#pragma once
#include <cstdint>
#include "sm100_ptx.cuh"

namespace seam {
namespace xchg {

constexpr int MAX_WORLD = 8;
enum { KIND_Q = 0, KIND_L = 1, KIND_F = 2, NKIND = 3 };

struct Exchange {                        // device-side image of seam_exchange (include/seam_b200.h)
  int world, rank, Q, k, own_max;
  int q_lo[MAX_WORLD + 1];               // rank r owns queries [q_lo[r], q_lo[r+1])
  float* q_all[MAX_WORLD];               // (2, Q, 256) on each rank
  float* list_margin[MAX_WORLD];         // (2, world, own_max, k) on each rank: lists for the queries it owns
  int32_t* list_idx[MAX_WORLD];
  float* final_score[MAX_WORLD];         // (Q, k) on each rank: merged result (null: owners keep their rows)
  float* final_margin[MAX_WORLD];
  int32_t* final_idx[MAX_WORLD];
  uint32_t* flags[MAX_WORLD];            // (NKIND, MAX_WORLD) words on each rank
  uint32_t* step;                        // local: number of the step in progress (starts at 1)
  uint32_t* done;                        // local: NKIND CTA-completion counters (zero between kernels)
  float* q_all_mc;                       // NVSwitch multicast mapping of q_all (null: none): one store reaches every rank
  float* final_score_mc;                 // ... of the merged rows (all three or none)
  float* final_margin_mc;
  int32_t* final_idx_mc;
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t current_step(const Exchange& x) { return *reinterpret_cast<volatile uint32_t*>(x.step); }
__device__ __forceinline__ int owner_of(const Exchange& x, int qi) {
  int o = 0;
#pragma unroll
  for (int r = 1; r < MAX_WORLD; ++r)
    if (r < x.world && qi >= x.q_lo[r]) o = r;
  return o;
}

// Call from EVERY CTA of a producing grid, after a __syncthreads that follows its last store: returns true in
// thread 0 of the grid's last CTA once every rank has been told (callers that must do something "after the
// whole grid" -- advancing the step -- hang it on that).
__device__ __forceinline__ bool signal_all(const Exchange& x, int kind, uint32_t step, bool stored = true) {
  if (threadIdx.x != 0) return false;
  if (stored) __threadfence();                             // this CTA's stores (ordered by the CTA barrier) before its count;
                                                           // a CTA that stored nothing only needs to be counted
  const unsigned n = atomicAdd(x.done + kind, 1u);
  if (n != gridDim.x - 1) return false;
  x.done[kind] = 0u;                                       // ready for the next grid that uses this kind
  // ONE system-scope fence orders everything observed so far before the flag stores (a release store per rank
  // would repeat it world times: ~2 us each over NVLink)
  __threadfence_system();
  for (int r = 0; r < x.world; ++r) st_relaxed_sys(x.flags[r] + kind * MAX_WORLD + x.rank, step);
  return true;
}
// Spin until every rank's data of `kind` for `step` has landed here (watchdog: a peer that never shows up
// traps the launch after 10 s instead of hanging the GPU).  Thread r < world watches rank r.
__device__ __forceinline__ void wait_ranks(const Exchange& x, int kind, uint32_t step) {
  if ((int)threadIdx.x < x.world) {
    const uint32_t* f = x.flags[x.rank] + kind * MAX_WORLD + threadIdx.x;
    uint64_t t0 = 0;
    while ((int32_t)(ld_acquire_sys(f) - step) < 0) {
      __nanosleep(100);
      const uint64_t now = ptx::globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 10000000000ull) ptx::watchdog_trap(200u + (uint32_t)kind, (uint32_t)threadIdx.x, step);
    }
  }
}
__device__ __forceinline__ void wait_all(const Exchange& x, int kind, uint32_t step) {
  wait_ranks(x, kind, step);
  __syncthreads();
}

}  // namespace xchg
}  // namespace seam
