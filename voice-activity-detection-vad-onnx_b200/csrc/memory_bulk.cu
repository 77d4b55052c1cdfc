// memory_bulk.cu -- FSMN/DFSMN memory block (a5-a7) for whole chunks that fit in shared memory.
//
//   out[s][t][c] = p[s][t][c] + sum_k wl[c][k] p[s][t-(N1-1)+k][c] + sum_k wr[c][k] p[s][t+1+k][c] (+ res[s][t][c])
//
// The block is pure streaming (1.5 KB per frame row in and out, 80 FLOP per output), so the only thing
// that matters is how many bytes each SM keeps in flight.  The register-window kernel
// (fsmn_memory_stream_kernel) tops out at ~57 KB per SM = 0.65 of the HBM copy peak.  Here a persistent
// CTA owns one stream at a time: a producer thread pulls the stream's whole p tile ([T][C] fp32,
// contiguous because activations are time-major) and its residual tile into a two-slot shared-memory
// ring with cp.async.bulk (mbarrier complete_tx; thread 0 issues the next stream's copies before it
// starts computing the current one), so ~100 KB of loads per SM are in flight with no registers
// involved, while the 8 warps run the FIR out of shared memory (thread = channel PAIR x
// quarter of the time axis, sliding register window, taps in shared memory) and write 256-byte lines.
// The FIR itself is as much an FMA-issue problem as a bandwidth one (80 FLOP per 12 bytes moved: 3.9k
// issue cycles per [98][128] tile against 7.5k cycles of HBM time), so the two channels of a pair go
// through the packed fp32x2 FMA of sm_100 (fma.rn.f32x2 -> FFMA2): same IEEE result per lane, half the
// instructions.
#include "tc_ptx.cuh"

namespace vadx {

constexpr int kMbC = 128;           // channels (consumer threads = 2 * kMbC)
// outputs per register group R and groups per thread G are template parameters: a thread owns <= G * R frames; R independent
// FFMA2 chains, the taps are re-read once per group
constexpr int kMbMaxOwn = 26;        // frames per thread every (R, G) variant can hold (T <= 104)
constexpr int kMbConsumers = 2 * kMbC;
constexpr int kMbThreads = kMbConsumers;   // no separate producer warp: 9 warps would be allocated as 12 (register file)

// packed fp32x2 arithmetic (sm_100): two independent IEEE fp32 operations per instruction
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

struct MemBulkArgs {
  const float* p;
  const float* res;   // may be null
  const float* wl;    // [C][N1]
  const float* wr;    // [C][N2]
  float* out;
  int64_t n_streams;
  int T;
  int pf;   // L2 prefetch distance, in streams of this CTA beyond the one being copied to shared memory (0 = off)
};

template <int N1, int N2, int kMbR, int kMbGroups>
__global__ void __launch_bounds__(kMbThreads, 1) fsmn_memory_bulk_kernel(const MemBulkArgs g) {
  constexpr int HL = N1 - 1, HR = N2;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int T = g.T;
  const uint32_t tile_bytes = (uint32_t)T * kMbC * 4u;
  const int n_tiles_per_slot = g.res ? 2 : 1;
  float* taps = reinterpret_cast<float*>(smem_raw);                       // [N1 + N2][C]
  uint8_t* ring = smem_raw + (size_t)(N1 + N2) * kMbC * 4;               // 2 slots x (p [, res])
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)2 * n_tiles_per_slot * tile_bytes);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (2 + s); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), kMbConsumers / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < kMbConsumers) {
    // taps[k][c]: channels (2j, 2j+1) of one tap are adjacent, i.e. one 64-bit pair per consumer thread
    const int c = threadIdx.x & (kMbC - 1);
    for (int k = threadIdx.x >> 7; k < N1; k += 2) taps[k * kMbC + c] = g.wl[c * N1 + k];
    for (int k = threadIdx.x >> 7; k < N2; k += 2) taps[(N1 + k) * kMbC + c] = g.wr[c * N2 + k];
  }
  __syncthreads();

  // producer duty (thread 0): stream `it`'s tiles go to slot it & 1 once every consumer warp has released it
  auto produce = [&](int it, int64_t s) {
    const int slot = it & 1;
    mbar_wait(empty_bar(slot), (((uint32_t)it >> 1) & 1u) ^ 1u);
    mbar_expect_tx(full_bar(slot), tile_bytes * n_tiles_per_slot);
    uint8_t* dst = ring + (size_t)slot * n_tiles_per_slot * tile_bytes;
    const uint8_t* src_p = reinterpret_cast<const uint8_t*>(g.p) + (size_t)s * tile_bytes;
    // pieces of <= 16 KB keep every copy's size field comfortably inside the instruction's range
    for (uint32_t off = 0; off < tile_bytes; off += 16384u) {
      const uint32_t nb = tile_bytes - off < 16384u ? tile_bytes - off : 16384u;
      bulk_g2s(smem_u32(dst) + off, src_p + off, nb, full_bar(slot));
    }
    // one stream's tiles in flight do not cover the HBM latency: ask L2 for a later stream's now
    const int64_t s_pf = s + (int64_t)g.pf * gridDim.x;
    if (g.pf && s_pf < g.n_streams) {
      for (uint32_t off = 0; off < tile_bytes; off += 16384u) {
        const uint32_t nb = tile_bytes - off < 16384u ? tile_bytes - off : 16384u;
        l2_prefetch(reinterpret_cast<const uint8_t*>(g.p) + (size_t)s_pf * tile_bytes + off, nb);
        if (g.res) l2_prefetch(reinterpret_cast<const uint8_t*>(g.res) + (size_t)s_pf * tile_bytes + off, nb);
      }
    }
    if (g.res) {
      const uint8_t* src_r = reinterpret_cast<const uint8_t*>(g.res) + (size_t)s * tile_bytes;
      for (uint32_t off = 0; off < tile_bytes; off += 16384u) {
        const uint32_t nb = tile_bytes - off < 16384u ? tile_bytes - off : 16384u;
        bulk_g2s(smem_u32(dst) + tile_bytes + off, src_r + off, nb, full_bar(slot));
      }
    }
  };
  if (threadIdx.x == 0 && (int64_t)blockIdx.x < g.n_streams) produce(0, blockIdx.x);
  {
    // ===================== consumers: thread = (channel pair, quarter of the time axis) =====================
    typedef unsigned long long u64;
    const int cp = threadIdx.x & (kMbC / 2 - 1);
    const int qd = threadIdx.x >> 6;                    // 0..3
    const int base = T >> 2, rem = T & 3;
    const int ta = qd * base + (qd < rem ? qd : rem);
    const int tb = ta + base + (qd < rem ? 1 : 0);
    const u64* tp = reinterpret_cast<const u64*>(taps) + cp;   // [N1 + N2][C/2] pairs
    constexpr int LD2 = kMbC / 2;
    int it = 0;
    for (int64_t s = blockIdx.x; s < g.n_streams; s += gridDim.x, ++it) {
      const int slot = it & 1;
      if (threadIdx.x == 0 && s + gridDim.x < g.n_streams) produce(it + 1, s + gridDim.x);   // next stream's loads fly while this one computes
      mbar_wait(full_bar(slot), ((uint32_t)it >> 1) & 1u, 32);
      const u64* sp = reinterpret_cast<const u64*>(ring + (size_t)slot * n_tiles_per_slot * tile_bytes) + cp;
      const u64* sr = sp + (size_t)T * LD2;
      u64* o = reinterpret_cast<u64*>(g.out + ((size_t)s * T) * kMbC) + cp;
      auto P = [&](int t) -> u64 { return (t >= 0 && t < T) ? sp[t * LD2] : 0ull; };
      // the whole quarter's window with static indices (the group loop is fully unrolled): no register
      // shifting between groups -- those moves would run on the FMA pipe
      u64 win[HL + HR + kMbGroups * kMbR];
#pragma unroll
      for (int j = 0; j < HL + HR; ++j) win[j] = P(ta - HL + j);
#pragma unroll
      for (int gi = 0; gi < kMbGroups; ++gi) {
        const int t0 = ta + gi * kMbR;
        if (t0 < tb) {
#pragma unroll
          for (int r = 0; r < kMbR; ++r) win[HL + HR + gi * kMbR + r] = P(t0 + HR + r);
          u64 acc[kMbR];
#pragma unroll
          for (int r = 0; r < kMbR; ++r) acc[r] = win[gi * kMbR + r + HL];
#pragma unroll
          for (int k = 0; k < N1; ++k) {
            const u64 ck = tp[k * LD2];
#pragma unroll
            for (int r = 0; r < kMbR; ++r) acc[r] = ffma2(ck, win[gi * kMbR + r + k], acc[r]);
          }
#pragma unroll
          for (int k = 0; k < N2; ++k) {
            const u64 ck = tp[(N1 + k) * LD2];
#pragma unroll
            for (int r = 0; r < kMbR; ++r) acc[r] = ffma2(ck, win[gi * kMbR + r + N1 + k], acc[r]);
          }
#pragma unroll
          for (int r = 0; r < kMbR; ++r) {
            const int t = t0 + r;
            if (t < tb) o[(size_t)t * LD2] = g.res ? fadd2(acc[r], sr[t * LD2]) : acc[r];
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar(slot));
    }
  }
}

// host side: returns false when the shape does not fit this kernel (the caller falls back)
bool memory_bulk_fits(int n_back, int stride_back, int n_ahead, int stride_ahead, int n_frames, int n_channels,
                      int64_t ldp, int64_t ldr, int64_t ldo, const float* p, const float* res, const float* out,
                      const float* cache_in, const float* cache_out, size_t* smem_out) {
  if (cache_in || cache_out) return false;
  if (n_channels != kMbC || ldp != kMbC || ldo != kMbC || (res && ldr != kMbC)) return false;
  if (stride_back != 1 || n_back != 20 || !(n_ahead == 0 || (n_ahead == 20 && stride_ahead == 1))) return false;
  if (n_frames < 2 || (n_frames + 3) / 4 > kMbMaxOwn) return false;
  if (!aligned16(p) || !aligned16(out) || (res && !aligned16(res))) return false;
  const size_t tile = (size_t)n_frames * kMbC * 4;
  const size_t smem = (size_t)(n_back + n_ahead) * kMbC * 4 + 2 * (res ? 2 : 1) * tile + 4 * 8 + 16;
  if (smem > (size_t)kTcSmemBudget) return false;
  *smem_out = smem;
  return true;
}

int memory_bulk_launch(const float* p, const float* wl, const float* wr, int n_ahead, const float* res, float* out,
                       int64_t n_streams, int n_frames, size_t smem, cudaStream_t st) {
  static PerDevice per_device;
  int n_sm = 148;
  VADX_TRY(per_device.ensure(&n_sm, [] {
    cudaError_t e = cudaSuccess;
    auto opt_in = [&](auto kern) {
      if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBudget);
    };
    opt_in(fsmn_memory_bulk_kernel<20, 20, 7, 4>);  opt_in(fsmn_memory_bulk_kernel<20, 0, 7, 4>);
    opt_in(fsmn_memory_bulk_kernel<20, 20, 13, 2>); opt_in(fsmn_memory_bulk_kernel<20, 0, 13, 2>);
    opt_in(fsmn_memory_bulk_kernel<20, 20, 26, 1>); opt_in(fsmn_memory_bulk_kernel<20, 0, 26, 1>);
    return e;
  }));
  static const int pf = ab_env("VADX_MEM_PF", 0);   // measured slower on B200 (2.17 -> 2.5 ms per step): off
  MemBulkArgs g{p, res, wl, wr, out, n_streams, n_frames, pf};
  const int grid = (int)std::min<int64_t>(n_streams, n_sm);
  static const int rg = ab_env("VADX_MEM_RG", 13);
  if (rg == 7) {
    if (n_ahead == 20) fsmn_memory_bulk_kernel<20, 20, 7, 4><<<grid, kMbThreads, smem, st>>>(g);
    else fsmn_memory_bulk_kernel<20, 0, 7, 4><<<grid, kMbThreads, smem, st>>>(g);
  } else if (rg == 26) {
    if (n_ahead == 20) fsmn_memory_bulk_kernel<20, 20, 26, 1><<<grid, kMbThreads, smem, st>>>(g);
    else fsmn_memory_bulk_kernel<20, 0, 26, 1><<<grid, kMbThreads, smem, st>>>(g);
  } else {
    if (n_ahead == 20) fsmn_memory_bulk_kernel<20, 20, 13, 2><<<grid, kMbThreads, smem, st>>>(g);
    else fsmn_memory_bulk_kernel<20, 0, 13, 2><<<grid, kMbThreads, smem, st>>>(g);
  }
  return after_launch("vadx_fsmn_memory_f32(bulk)");
}

}  // namespace vadx
