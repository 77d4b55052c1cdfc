// block_tc.cu -- the second half of a DFSMN block in ONE kernel: dense layer (tcgen05) + memory block (FIR) + residual.
//
//   p   = act(X[rows][K] W^T + b)                       (FireRedVAD/Export_FireRedVAD.py:253-263: fc2; :290-296: fc2 + ReLU)
//   out = p + sum_k wl[c][k] p[t-(N1-1)+k] + sum_k wr[c][k] p[t+1+k]  (+ res)      (FSMN.forward, :213-236)
//
// Unfused, p makes a round trip through HBM (0.5 KB per frame row written by the dense layer, read back by the
// memory kernel) and the memory block is its own 1.5 KB-per-row streaming pass.  Here a row tile is ONE stream's
// chunk (T <= 128 frames; activations are time-major, so the chunk is contiguous): the accumulator tile leaves TMEM
// into a shared-memory tile, the four epilogue warps then own one channel each (128 channels = N) and run the FIR
// over the chunk's time axis out of that tile, add the residual and write the block output.  The FIR needs no halo
// because a chunk is zero-padded on both sides by definition.  Loaders / MMA issue are those of gemm_tc.cu with the
// row mapping tile -> stream; the activation ring is one 32 KB stage (the p tile takes the room of the second one).
//
// STATUS: correct (tests/test_gpu_tc.py::test_fused_dense_plus_memory_block) and wired in behind the scalar
// "engine.fuse_block", but NOT the default: shared memory (131 KB weight image + stage + 50 KB p tile) and the
// register file (loaders hold two stages of loads in registers, the FIR wants its 46-deep window) leave the FIR only
// four warps, and measured on B200 the fused block is slower than the two streaming kernels it replaces
// (0.75 ms vs 0.27 + 0.27 ms per block with a dedicated MMA warp; this 12-warp layout is slower still).  Kept as the
// starting point for the version that moves the activation loads to bulk copies and frees the loaders' registers:
// that version is block_stages.cu (default since the end of round 1).
#include "tc_ptx.cuh"

namespace vadx {

constexpr int kBkLoaderWarps = 8;
constexpr int kBkEpiWarp0 = kBkLoaderWarps;
// 12 warps: 13 would be allocated as 16 (128 registers per thread); the FIR phase wants more, so the MMA issuer is
// thread 0 of the first loader warp -- with a single activation stage loader and issuer alternate anyway
constexpr int kBkThreads = (kBkLoaderWarps + 4) * 32;
constexpr int kBkC = 128;      // N of the dense layer = channels of the memory block
constexpr int kBkR = 7;        // FIR outputs per register group
constexpr int kBkGroups = 17;  // unrolled groups: chunks of up to 119 frames
constexpr int kBkStageBytes = 2 * kTcTileBytes;

struct BlockArgs {
  const float* X;
  int64_t ldx;
  const uint8_t* Wimg;
  const float* bias;   // [128] or null
  const float* wl;     // [128][N1]
  const float* wr;     // [128][N2] or null
  const float* res;    // [rows][128] or null
  float* out;          // [rows][128]
  int64_t n_streams;
  int T, K, kc, n_k16, vec_x;
};

template <int ACT, int N1, int N2>
__global__ void __launch_bounds__(kBkThreads, 1) fc2_memory_tc_kernel(const BlockArgs g) {
  constexpr int HL = N1 - 1, HR = N2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* w_smem = smem_raw;
  const int w_bytes = g.kc * 2 * kBkC * 128;
  uint8_t* a_smem = w_smem + w_bytes;                                     // one stage: hi tile + lo tile
  float* p_tile = reinterpret_cast<float*>(a_smem + kBkStageBytes);       // [T][128], 16-byte chunks XOR-swizzled by t & 7
  float* bias_s = p_tile + (size_t)g.T * kBkC;
  float* taps_l = bias_s + kBkC;                                          // [N1][128] look-back taps
  uint64_t* bars = reinterpret_cast<uint64_t*>(taps_l + (size_t)N1 * kBkC);
  // bars: full, empty, tmem_full[2], tmem_empty[2], wbar
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t full_bar = bar0, empty_bar = bar0 + 8u;
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (4 + b); };
  const uint32_t wbar = bar0 + 8u * 6;
  constexpr uint32_t kTmemCols = 256;   // two 128-column accumulators

  if (threadIdx.x == 0) {
    mbar_init(full_bar, kBkLoaderWarps);
    mbar_init(empty_bar, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 4);
    }
    mbar_init(wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < kBkC; i += blockDim.x) bias_s[i] = g.bias ? g.bias[i] : 0.f;
  for (int i = threadIdx.x; i < N1 * kBkC; i += blockDim.x) {
    const int k = i / kBkC, c = i - k * kBkC;
    taps_l[i] = g.wl[c * N1 + k];
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kBkLoaderWarps) {
    // ===================== loaders: rows of one stream, fp32 -> (hi, lo) bf16, swizzled =====================
    constexpr int kPasses = kTcBM / (kBkLoaderWarps * 4);   // 4
    const int t = threadIdx.x;
    const int kq = t & 7, r_in = t >> 3;
    struct Seq { int64_t tile; int c; };
    auto valid = [&](const Seq& q) { return q.tile < g.n_streams; };
    auto advance = [&](Seq& q) { if (++q.c == g.kc) { q.c = 0; q.tile += gridDim.x; } };
    auto issue = [&](const Seq& q, float4 (*ld)[2]) {
      const int64_t row0 = q.tile * g.T;
      const int k = q.c * kTcBK + kq * 8;
#pragma unroll
      for (int pass = 0; pass < kPasses; ++pass) {
        const int r = min(pass * 32 + r_in, g.T - 1);
        const float* src = g.X + (row0 + r) * g.ldx + k;
        if (g.vec_x && k + 7 < g.K) {
          ld[pass][0] = __ldg(reinterpret_cast<const float4*>(src));
          ld[pass][1] = __ldg(reinterpret_cast<const float4*>(src) + 1);
        } else {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = (k + j < g.K) ? __ldg(src + j) : 0.f;
          ld[pass][0] = make_float4(v[0], v[1], v[2], v[3]);
          ld[pass][1] = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
    };
    uint32_t phase = 0;
    // MMA issue (thread 0): weights first, then one step per activation chunk
    const uint32_t img_bytes = (uint32_t)kBkC * 128u;
    const uint32_t idesc = umma_idesc_bf16(kBkC);
    uint32_t mma_phase = 0;
    int mma_it = 0;
    if (t == 0) {
      mbar_expect_tx(wbar, (uint32_t)w_bytes);
      for (int i = 0; i < g.kc * 2; ++i) bulk_g2s(smem_u32(w_smem) + i * img_bytes, g.Wimg + (size_t)i * img_bytes, img_bytes, wbar);
      mbar_wait(wbar, 0);
    }
    auto mma_step = [&](int c) {
      const int b = mma_it & 1;
      if (c == 0) {
        mbar_wait(tempty_bar(b), (((uint32_t)(mma_it >> 1)) & 1u) ^ 1u);
        tc_fence_after();
      }
      mbar_wait(full_bar, mma_phase);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(b * kBkC);
      const uint32_t a_hi = smem_u32(a_smem), a_lo = a_hi + kTcTileBytes;
      const uint32_t w_hi = smem_u32(w_smem) + (uint32_t)(c * 2) * img_bytes, w_lo = w_hi + img_bytes;
      const int nk = min(4, g.n_k16 - c * 4);
      for (int j = 0; j < nk; ++j)
        umma_bf16(d_tmem, umma_desc_sw128(a_hi + 32u * j), umma_desc_sw128(w_hi + 32u * j), idesc, (c | j) ? 1u : 0u);
      for (int j = 0; j < nk; ++j) umma_bf16(d_tmem, umma_desc_sw128(a_lo + 32u * j), umma_desc_sw128(w_hi + 32u * j), idesc, 1u);
      for (int j = 0; j < nk; ++j) umma_bf16(d_tmem, umma_desc_sw128(a_hi + 32u * j), umma_desc_sw128(w_lo + 32u * j), idesc, 1u);
      umma_commit(empty_bar);
      mma_phase ^= 1u;
      if (c == g.kc - 1) {
        umma_commit(tfull_bar(b));
        ++mma_it;
      }
    };
    auto consume = [&](const Seq& sq, float4 (*ld)[2]) {
      mbar_wait(empty_bar, phase ^ 1u, 64);
      uint8_t* st_hi = a_smem;
      uint8_t* st_lo = st_hi + kTcTileBytes;
#pragma unroll
      for (int pass = 0; pass < kPasses; ++pass) {
        const int r = pass * 32 + r_in;
        float4 p0 = ld[pass][0], p1 = ld[pass][1];
        if (r >= g.T) p0 = p1 = make_float4(0.f, 0.f, 0.f, 0.f);
        uint4 hi, lo;
        split2(p0.x, p0.y, hi.x, lo.x);
        split2(p0.z, p0.w, hi.y, lo.y);
        split2(p1.x, p1.y, hi.z, lo.z);
        split2(p1.z, p1.w, hi.w, lo.w);
        const int off = r * 128 + ((kq ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(st_hi + off) = hi;
        *reinterpret_cast<uint4*>(st_lo + off) = lo;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar);
      phase ^= 1u;
      if (t == 0) mma_step(sq.c);
    };
    float4 b0[kPasses][2], b1[kPasses][2], b2[kPasses][2];
    Seq nxt{(int64_t)blockIdx.x, 0}, cur{(int64_t)blockIdx.x, 0};
    if (valid(nxt)) { issue(nxt, b0); advance(nxt); }
    if (valid(nxt)) { issue(nxt, b1); advance(nxt); }
    while (valid(cur)) {
      if (valid(nxt)) { issue(nxt, b2); advance(nxt); }
      consume(cur, b0); advance(cur);
      if (!valid(cur)) break;
      if (valid(nxt)) { issue(nxt, b0); advance(nxt); }
      consume(cur, b1); advance(cur);
      if (!valid(cur)) break;
      if (valid(nxt)) { issue(nxt, b1); advance(nxt); }
      consume(cur, b2); advance(cur);
    }
  } else {
    // ===================== epilogue warps: TMEM -> p tile, then one channel each through the FIR =====================
    const int q = warp - kBkEpiWarp0;
    const int c = q * 32 + lane;                 // this thread's channel in the FIR phase, its ROW in the TMEM phase
    const int T = g.T;
    float cr[N2 > 0 ? N2 : 1];     // look-ahead taps in registers, look-back taps in shared memory (register budget)
#pragma unroll
    for (int k = 0; k < N2; ++k) cr[k] = g.wr[c * N2 + k];
    const float* tl = taps_l + c;
    // element (t, ch) of the p tile: 16-byte chunk index (ch / 4) ^ (t & 7) keeps both access patterns conflict-free
    auto p_at = [&](int t, int ch) -> float* { return p_tile + t * kBkC + ((((ch >> 2) ^ (t & 7)) << 2) | (ch & 3)); };
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < g.n_streams; tile += gridDim.x, ++it) {
      const int b = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait(tfull_bar(b), use & 1u, 64);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * kBkC);
      const int r = c;                            // TMEM lane = row of the tile = frame t
#pragma unroll
      for (int c0 = 0; c0 < kBkC; c0 += 32) {
        float v[32];
        tmem_ld32(taddr + (uint32_t)c0, v);
        if (r < T) {
          const float4* b4 = reinterpret_cast<const float4*>(bias_s + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = b4[j];
            *reinterpret_cast<float4*>(p_at(r, c0 + 4 * j)) =
                make_float4(apply_act(v[4 * j] + bb.x, ACT), apply_act(v[4 * j + 1] + bb.y, ACT),
                            apply_act(v[4 * j + 2] + bb.z, ACT), apply_act(v[4 * j + 3] + bb.w, ACT));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(b));          // the accumulator is free for the tile after next
      asm volatile("bar.sync 1, 128;" ::: "memory");      // p tile complete (the four epilogue warps)
      // ---- FIR over the chunk's time axis for channel c ----
      const int64_t row0 = tile * T;
      const float* rs = g.res ? g.res + row0 * kBkC + c : nullptr;
      float* os = g.out + row0 * kBkC + c;
      auto P = [&](int t) -> float { return (t >= 0 && t < T) ? *p_at(t, c) : 0.f; };
      // the chunk's whole window with static indices (groups fully unrolled, guarded by T): no register shifting,
      // taps in registers -- per group 7 shared-memory reads, 7 residual loads, 280 FMAs, 7 stores
      float win[HL + HR + kBkGroups * kBkR];
#pragma unroll
      for (int j = 0; j < HL + HR; ++j) win[j] = P(j - HL);
#pragma unroll
      for (int gi = 0; gi < kBkGroups; ++gi) {
        const int t0 = gi * kBkR;
        if (t0 < T) {
          float rv[kBkR];
#pragma unroll
          for (int rr = 0; rr < kBkR; ++rr) rv[rr] = (rs && t0 + rr < T) ? __ldg(rs + (int64_t)(t0 + rr) * kBkC) : 0.f;
#pragma unroll
          for (int rr = 0; rr < kBkR; ++rr) win[HL + HR + t0 + rr] = P(t0 + HR + rr);
          float acc[kBkR];
#pragma unroll
          for (int rr = 0; rr < kBkR; ++rr) acc[rr] = win[t0 + rr + HL];
          float cl[N1];
#pragma unroll
          for (int k = 0; k < N1; ++k) cl[k] = tl[k * kBkC];      // all reads issued before the first FMA needs one
#pragma unroll
          for (int k = 0; k < N1; ++k) {
#pragma unroll
            for (int rr = 0; rr < kBkR; ++rr) acc[rr] = fmaf(cl[k], win[t0 + rr + k], acc[rr]);
          }
#pragma unroll
          for (int k = 0; k < N2; ++k) {
#pragma unroll
            for (int rr = 0; rr < kBkR; ++rr) acc[rr] = fmaf(cr[k], win[t0 + rr + N1 + k], acc[rr]);
          }
#pragma unroll
          for (int rr = 0; rr < kBkR; ++rr)
            if (t0 + rr < T) os[(int64_t)(t0 + rr) * kBkC] = rs ? acc[rr] + rv[rr] : acc[rr];
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");      // everyone is done reading the p tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

static size_t block_smem_bytes(int kc, int T, int n1) {
  return (size_t)kc * 2 * kBkC * 128 + kBkStageBytes + (size_t)T * kBkC * 4 + kBkC * 4 + (size_t)n1 * kBkC * 4 + 7 * 8 + 16;
}

}  // namespace vadx

using namespace vadx;

extern "C" int vadx_fc2_memory_tc_supported(int n_in, int n_out, int n_frames, int n_back, int stride_back, int n_ahead,
                                            int stride_ahead) {
  if (n_out != kBkC || n_in < 1 || n_frames < 2 || n_frames > kTcBM || n_frames > kBkGroups * kBkR) return 0;
  if (n_back != 20 || stride_back != 1 || !(n_ahead == 0 || (n_ahead == 20 && stride_ahead == 1))) return 0;
  const int kc = (int)ceil_div(n_in, kTcBK);
  return block_smem_bytes(kc, n_frames, n_back) <= (size_t)kTcSmemBudget ? 1 : 0;
}

extern "C" int vadx_fc2_memory_tc_f32(const float* d_x, int64_t ldx, const void* d_wimg, const float* d_bias, int act,
                                      const float* d_wl, int n_back, const float* d_wr, int n_ahead,
                                      const float* d_residual, float* d_out, int64_t n_streams, int n_frames, int n_in,
                                      void* stream) {
  StageTimer _timer(VADX_STAGE_LINEAR, (cudaStream_t)stream);
  VADX_REQUIRE(d_x && d_wimg && d_wl && d_out && (n_ahead == 0 || d_wr), "vadx_fc2_memory_tc_f32: null pointer");
  VADX_REQUIRE(vadx_fc2_memory_tc_supported(n_in, kBkC, n_frames, n_back, 1, n_ahead, 1),
               "vadx_fc2_memory_tc_f32: shape (K=%d, T=%d, taps %d+%d) is not supported", n_in, n_frames, n_back, n_ahead);
  VADX_REQUIRE(ldx >= n_in && n_streams >= 0 && aligned16(d_wimg), "vadx_fc2_memory_tc_f32: bad argument");
  VADX_REQUIRE((act & 15) == VADX_ACT_NONE || (act & 15) == VADX_ACT_RELU, "vadx_fc2_memory_tc_f32: activation %d", act);
  if (n_streams == 0) return VADX_OK;
  static int n_sm = 0;
  static bool configured = false;
  if (!configured) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(fc2_memory_tc_kernel<VADX_ACT_NONE, 20, 20>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBudget);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fc2_memory_tc_kernel<VADX_ACT_RELU, 20, 20>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBudget);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fc2_memory_tc_kernel<VADX_ACT_NONE, 20, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBudget);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fc2_memory_tc_kernel<VADX_ACT_RELU, 20, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBudget);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(fc2_memory_tc_kernel)");
    configured = true;
  }
  BlockArgs g{};
  g.X = d_x; g.ldx = ldx; g.Wimg = static_cast<const uint8_t*>(d_wimg); g.bias = d_bias; g.wl = d_wl; g.wr = d_wr;
  g.res = d_residual; g.out = d_out; g.n_streams = n_streams; g.T = n_frames; g.K = n_in;
  g.kc = (int)ceil_div(n_in, kTcBK); g.n_k16 = (int)ceil_div(n_in, 16);
  g.vec_x = ((ldx & 3) == 0) && aligned16(d_x);
  const size_t smem = block_smem_bytes(g.kc, n_frames, n_back);
  const int grid = (int)std::min<int64_t>(n_streams, n_sm > 0 ? n_sm : 148);
  const bool relu = (act & 15) == VADX_ACT_RELU;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_ahead == 20) {
    if (relu) fc2_memory_tc_kernel<VADX_ACT_RELU, 20, 20><<<grid, kBkThreads, smem, st>>>(g);
    else fc2_memory_tc_kernel<VADX_ACT_NONE, 20, 20><<<grid, kBkThreads, smem, st>>>(g);
  } else {
    if (relu) fc2_memory_tc_kernel<VADX_ACT_RELU, 20, 0><<<grid, kBkThreads, smem, st>>>(g);
    else fc2_memory_tc_kernel<VADX_ACT_NONE, 20, 0><<<grid, kBkThreads, smem, st>>>(g);
  }
  return after_launch("vadx_fc2_memory_tc_f32");
}
