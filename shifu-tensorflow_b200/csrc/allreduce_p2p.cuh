// Two-shot all-reduce of the flat fp32 gradient over NVLink peer memory (CUDA IPC), replacing the parameter-server
// accumulator of the reference (ConditionalAccumulator mean over gRPC, res/ssgd_monitor.py:136-141) and, on the hot
// path, NCCL: one kernel per step, captured in the step graph.
//
//   every rank r owns slice r of the flat vector.
//   phase A  "my gradient is complete"  : store epoch into arrive[r] of every peer (st.release.sys), wait until all
//                                          peers' arrive slots in MY flag block carry the epoch (ld.acquire.sys)
//   phase B  reduce-scatter + all-gather : for my slice, read the slice from every rank (P2P loads through NVLink,
//                                          16 B per thread, fixed rank order -> bit-identical on every rank), store the
//                                          sum into the slice of EVERY rank's buffer (P2P stores)
//   phase C  "my slice is written"       : last block to finish publishes done[r] = epoch to every peer, then waits
//                                          until every peer has done the same -> all slices of my buffer are final
//
// A rank only ever reads other ranks' copies of ITS slice and only ever writes ITS slice of other ranks' copies, so
// there is no write/read overlap between ranks; the next step's clearing of the gradient happens after phase C, i.e.
// after every peer has finished reading it.
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace sb {

#define SB_MAX_RANKS 16

struct P2PFlags {                      // lives right behind the gradient in the IPC-exported allocation
  unsigned int arrive[SB_MAX_RANKS];   // arrive[q] written by rank q
  unsigned int done[SB_MAX_RANKS];     // done[q]   written by rank q
  unsigned int blocks_done;            // local: grid-wide completion counter of phase B
  unsigned int pad[31];
};

struct P2PPeers {                      // device-resident table, same order on every rank
  float* grad[SB_MAX_RANKS];
  P2PFlags* flags[SB_MAX_RANKS];
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void wait_flags(const unsigned int* slots, int world, unsigned int epoch) {
  // one lane per rank polls; a watchdog turns a lost peer into a trap instead of a hung GPU
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    if (lane < world) {
      unsigned long long t0 = 0;
      unsigned int spins = 0;
      while (static_cast<int>(ld_acquire_sys(slots + lane) - epoch) < 0) {
        if ((++spins & 0xFFFu) == 0) {
          const unsigned long long now = globaltimer_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > 20000000000ull) __trap();  // 20 s
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();
}

// n4 = number of float4 in the (padded) vector, a multiple of world.  W = compile-time upper bound of `world`
// (2 / 4 / 8 / 16): ALL W x U remote loads of a thread are issued before the first add, so one NVLink round trip
// (~2-3 us) is paid per pass instead of one per rank.
template <int W>
static __global__ void __launch_bounds__(512)
allreduce_p2p_kernel(const P2PPeers* __restrict__ peers, const BatchDesc* __restrict__ desc, int rank, int world, long long n4) {
  constexpr int U = (W <= 2) ? 8 : (W <= 4 ? 4 : (W <= 8 ? 2 : 1));   // W * U = 16 float4 (256 B) in flight per thread
  pdl_launch_dependents();   // the optimizer behind it may be scheduled now; it waits (griddepcontrol.wait) for this grid
  const unsigned int epoch = desc->epoch;
  P2PFlags* mine = peers->flags[rank];
  // ---- phase A ----
  if (blockIdx.x == 0 && threadIdx.x < world) st_release_sys(&peers->flags[threadIdx.x]->arrive[rank], epoch);
  wait_flags(mine->arrive, world, epoch);
  // ---- phase B ----
  float* gp[W];
#pragma unroll
  for (int q = 0; q < W; ++q) gp[q] = peers->grad[q < world ? q : rank];
  const long long slice = n4 / world;
  const long long base = slice * rank;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i0 < slice; i0 += stride * U) {
    float4 v[W][U];
#pragma unroll
    for (int q = 0; q < W; ++q)
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = i0 + u * stride;
        v[q][u] = (q < world && i < slice) ? ld_peer_f4(gp[q] + (base + i) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float4 acc = v[0][u];                 // fixed rank order 0..world-1 -> every rank computes bit-identical sums
#pragma unroll
      for (int q = 1; q < W; ++q) { acc.x += v[q][u].x; acc.y += v[q][u].y; acc.z += v[q][u].z; acc.w += v[q][u].w; }
      const long long i = i0 + u * stride;
      if (i < slice) {
#pragma unroll
        for (int q = 0; q < W; ++q)
          if (q < world) *reinterpret_cast<float4*>(gp[q] + (base + i) * 4) = acc;
      }
    }
  }
  // ---- phase C ----
  __threadfence_system();
  __syncthreads();
  __shared__ unsigned int last;
  if (threadIdx.x == 0) last = (atomicAdd(&mine->blocks_done, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) mine->blocks_done = 0;
    __threadfence_system();
    if (threadIdx.x < world) st_release_sys(&peers->flags[threadIdx.x]->done[rank], epoch);
    wait_flags(mine->done, world, epoch);
  }
}

}  // namespace sb
