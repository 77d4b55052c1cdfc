// block_stages.cu -- the second half of a DFSMN block in ONE kernel, second version: dense layer (tcgen05) + memory block
// (FIR) + residual, fed with ready-made operand stages and with the FIR running straight out of tensor memory.
//
//   p   = act(h W^T + b)                                                    (FireRedVAD/Export_FireRedVAD.py:253-263, :290-296)
//   out = p + sum_k wl[c][k] p[t-(N1-1)+k] + sum_k wr[c][k] p[t+1+k]  (+ res)      (FSMN.forward, :213-236)
//
// What the first version (block_tc.cu, removed in round 2; slower than the unfused pair) ran out of was shared memory and registers: a
// 131 KB weight image, a 32 KB activation stage, a 50 KB p tile, and loader warps holding two stages of loads.  Here
//   * h arrives as per-stream operand stages ([stream][K chunk][hi | lo] swizzled bf16 images written by fc1's epilogue,
//     gemm_tc.cu y_split == 2): one thread streams them in with cp.async.bulk, there are no loader warps;
//   * the MMA is issued TRANSPOSED: A = the weight image (M = 128 output channels), B = the stream's h stage
//     (N = round_up(T, 16) frames), so the accumulator holds p^T -- TMEM lane = channel, TMEM column = frame.  A thread
//     reads ITS channel's whole time series with tcgen05.ld and runs the FIR along it in registers: no p tile at all,
//     and since a chunk is zero-padded on both sides by definition there is no halo either;
//   * T is a template parameter: every window index and every edge test is a compile-time constant.
//   * MIRROR (default): the second half of the frames is processed in reversed time, so both halves run one shared
//     unrolled body (half the code of a kernel that was instruction-fetch bound).
// p never exists in HBM (the unfused pair writes and re-reads 0.5 KB per frame row), and the eight FIR warps overlap the
// next stream's MMAs through two accumulators.  Callers: model.cu (FireRed, engine.fuse_stages); shapes it does not
// cover (other chunk lengths or tap counts, streaming caches) take the two-kernel tail.
#include <type_traits>

#include "tc_ptx.cuh"

namespace vadx {

constexpr int kBsC = 128;                            // channels: N of the dense layer = M of the transposed MMA
// warp 0: bulk-copy producer, warp 1: MMA issuer, then 4 * NP FIR warps = (TMEM lane quarter) x (NP parts of the time axis)
constexpr int kBsStages = 2;   // three stages measured no faster

struct BsArgs {
  const uint8_t* Himg;   // [S][kc][2][img bytes]
  const uint8_t* Wimg;   // [kc][2][128 x 128 B]
  const float* bias;     // [128] or null
  const float* wl;       // [128][N1]
  const float* wr;       // [128][N2] or null
  const float* res;      // [S*T][128] or null
  float* out;            // [S*T][128]
  int64_t n_streams;
  int kc, n_k16;
  int pf;   // L2 prefetch distance in streams of this CTA (0 = off)
  int mirror;   // 1: both halves run one shared body, the second in reversed time (needs NP == 2, T even, N1 == N2)
};

template <int ACT, int T, int N1, int N2, int NP, int kBsR, bool MIRROR>   // NP time parts, kBsR FIR outputs per register group
__global__ void __launch_bounds__(NP == 2 ? 384 : 512, 1) fc2_memory_stages_kernel(const BsArgs g) {
  constexpr int kBsFirWarps = 4 * NP;
  constexpr int TP = (T + 15) / 16 * 16;     // accumulator columns = N of the MMA
  constexpr int HL = N1 - 1;
  constexpr int kImg = TP * 128, kStage = 2 * kImg, kCopy = T * 128;
  constexpr int kWImg = kBsC * 128;
  static_assert(kImg % 1024 == 0, "stage images must keep the 1024-byte swizzle atom alignment");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* w_smem = smem_raw;
  const int w_bytes = g.kc * 2 * kWImg;
  uint8_t* a_smem = w_smem + w_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_smem + kBsStages * kStage);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (4 + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (8 + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (10 + b); };
  const uint32_t wbar = bar0 + 8u * 12;
  constexpr uint32_t kTmemCols = 256;   // two accumulators, 128 columns apart

  if (threadIdx.x == 0) {
    for (int s = 0; s < kBsStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), kBsFirWarps);
    }
    mbar_init(wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== producer: one stream's K chunk = two bulk copies (the T real rows of each image) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t s = blockIdx.x; s < g.n_streams; s += gridDim.x) {
        // two 25 KB stages in flight do not cover the HBM latency: ask L2 for this CTA's next stream (operand stages
        // and residual rows) while the current one is processed
        const int64_t sn = s + (int64_t)g.pf * gridDim.x;
        if (g.pf && sn < g.n_streams) {
          const uint8_t* hp = g.Himg + (size_t)sn * g.kc * 2 * kImg;
          const uint32_t hb = (uint32_t)(g.kc * 2 * kImg);
          for (uint32_t off = 0; off < hb; off += 16384u) l2_prefetch(hp + off, hb - off < 16384u ? hb - off : 16384u);
          if (g.res) {
            const uint8_t* rp = reinterpret_cast<const uint8_t*>(g.res + (size_t)sn * T * kBsC);
            const uint32_t rb = (uint32_t)(T * kBsC * 4);
            for (uint32_t off = 0; off < rb; off += 16384u) l2_prefetch(rp + off, rb - off < 16384u ? rb - off : 16384u);
          }
        }
        for (int c = 0; c < g.kc; ++c) {
          mbar_wait(empty_bar(stage), phase ^ 1u, 32);
          mbar_expect_tx(full_bar(stage), 2u * kCopy);
          const uint8_t* src = g.Himg + ((size_t)(s * g.kc + c) * 2) * kImg;
          const uint32_t dst = smem_u32(a_smem) + (uint32_t)stage * kStage;
          bulk_g2s(dst, src, kCopy, full_bar(stage));
          bulk_g2s(dst + kImg, src + kImg, kCopy, full_bar(stage));
          if (++stage == kBsStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: p^T = W h^T, three products of the two-term split =====================
    if (lane == 0) {
      mbar_expect_tx(wbar, (uint32_t)w_bytes);
      for (int i = 0; i < g.kc * 2; ++i) bulk_g2s(smem_u32(w_smem) + i * kWImg, g.Wimg + (size_t)i * kWImg, kWImg, wbar);
      mbar_wait(wbar, 0);
      const uint32_t idesc = umma_idesc_bf16(TP);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int64_t s = blockIdx.x; s < g.n_streams; s += gridDim.x, ++it) {
        const int b = it & 1;
        mbar_wait(tempty_bar(b), (((uint32_t)(it >> 1)) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(b * 128);
        for (int c = 0; c < g.kc; ++c) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t h_hi = smem_u32(a_smem) + (uint32_t)stage * kStage, h_lo = h_hi + kImg;
          const uint32_t w_hi = smem_u32(w_smem) + (uint32_t)(c * 2) * kWImg, w_lo = w_hi + kWImg;
          const int nk = min(4, g.n_k16 - c * 4);
          const uint64_t dh_hi = umma_desc_sw128(h_hi), dh_lo = umma_desc_sw128(h_lo);
          const uint64_t dw_hi = umma_desc_sw128(w_hi), dw_lo = umma_desc_sw128(w_lo);
          umma_k64(d_tmem, dw_hi, dh_hi, idesc, c ? 1u : 0u, nk);
          umma_k64(d_tmem, dw_hi, dh_lo, idesc, 1u, nk);
          umma_k64(d_tmem, dw_lo, dh_hi, idesc, 1u, nk);
          umma_commit(empty_bar(stage));
          if (++stage == kBsStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(b));
      }
    }
  } else {
    // ===================== FIR warps: thread = (channel, half of the time axis), the window comes from TMEM =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may touch
    const int part = (warp - 2) >> 2;       // warps 2..5: first part of the frames, 6..9: second, ...
    const int c = q * 32 + lane;
    const float bias_c = g.bias ? __ldg(g.bias + c) : 0.f;
    // the channel's taps are loop invariants: in registers for the whole kernel (a shared-memory read per tap and group
    // put ~30 cycles of latency in front of every seven FMAs)
    float tw[N1 + N2];
    if (!MIRROR) {
#pragma unroll
      for (int k = 0; k < N1; ++k) tw[k] = __ldg(g.wl + c * N1 + k);
#pragma unroll
      for (int k = 0; k < N2; ++k) tw[N1 + k] = __ldg(g.wr + c * N2 + k);
    }
    // Mirrored form (g.mirror, T even, two parts, N2 == N1): the second half of the frames is processed in REVERSED time, which
    // turns its right edge into a left edge -- both halves then run the SAME unrolled code (one 49-output body instead of
    // two: half the instruction footprint of a kernel that is instruction-fetch bound).  Canonical time u = t (first half)
    // or T-1-t (second half); canonical taps over offsets d = -N2 .. +N2: forward c_d = wl[N1-1+d] (d <= 0) | wr[d-1] (d > 0),
    // mirrored c'_d = c_{-d}; the one offset a half does not have (d = -N2 forward, +N2 mirrored) is a zero tap.
    constexpr int kD = N2;                       // canonical offsets -kD .. +kD
    float tc[2 * kD + 1];
    if (MIRROR) {
#pragma unroll
      for (int i = 0; i < 2 * kD + 1; ++i) {
        const int d = part == 0 ? i - kD : kD - i;                 // the physical offset this canonical tap applies to
        tc[i] = d == -kD ? 0.f : (d <= 0 ? __ldg(g.wl + c * N1 + (N1 - 1 + d)) : __ldg(g.wr + c * N2 + (d - 1)));
      }
    }
    int it = 0;
    for (int64_t s = blockIdx.x; s < g.n_streams; s += gridDim.x, ++it) {
      const int b = it & 1;
      mbar_wait(tfull_bar(b), ((uint32_t)(it >> 1)) & 1u, 64);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 128);
      const float* rs = g.res ? g.res + (size_t)s * T * kBsC + c : nullptr;
      float* os = g.out + (size_t)s * T * kBsC + c;
      if (MIRROR) {
        constexpr int kOut = T / 2;                      // outputs per half
        constexpr int kWin = kOut + kD;                  // canonical window: P[0 .. kOut-1+kD], P[u] = p[u] or p[T-1-u]
        constexpr int kLd = (kWin + 7) / 8 * 8;
        float win[kLd];
        if (part == 0) {
          // all loads issued, then collected: one TMEM latency per stream instead of nine
          uint32_t raw[kLd];
#pragma unroll
          for (int i = 0; i < kLd / 8; ++i) tmem_ld8_issue(taddr + (uint32_t)(8 * i), raw + 8 * i);
#pragma unroll
          for (int i = 0; i < kLd / 8; ++i) tmem_ld8_finish(raw + 8 * i);
#pragma unroll
          for (int j = 0; j < kLd; ++j) win[j] = __uint_as_float(raw[j]);
        } else {
          // columns T-1-u for u = 0 .. kWin-1, i.e. [T-kWin, T): loaded ascending from an 8-column boundary, renamed reversed
          constexpr int c_lo = (T - kWin) / 8 * 8;
          constexpr int n_ld = (T - c_lo + 7) / 8;
          static_assert(c_lo + 8 * n_ld <= TP, "mirrored window load runs past the accumulator");
          uint32_t raw[8 * n_ld];
#pragma unroll
          for (int i = 0; i < n_ld; ++i) tmem_ld8_issue(taddr + (uint32_t)(c_lo + 8 * i), raw + 8 * i);
#pragma unroll
          for (int i = 0; i < n_ld; ++i) tmem_ld8_finish(raw + 8 * i);
#pragma unroll
          for (int i = 0; i < 8 * n_ld; ++i) {
            const int u = T - 1 - (c_lo + i);
            if (u >= 0 && u < kLd) win[u] = __uint_as_float(raw[i]);
          }
#pragma unroll
          for (int u = T - c_lo; u < kLd; ++u) win[u] = 0.f;    // (not reached by the loads; never used either)
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(b));      // the accumulator is free for the stream after next
#pragma unroll
        for (int j = 0; j < kLd; ++j) win[j] = apply_act(win[j] + bias_c, ACT);
        constexpr int kGroups = (kOut + kBsR - 1) / kBsR;
#pragma unroll
        for (int gi = 0; gi < kGroups; ++gi) {
          const int u0 = gi * kBsR;
          float rv[kBsR], acc[kBsR];
#pragma unroll
          for (int r = 0; r < kBsR; ++r) {
            const int t = part == 0 ? u0 + r : T - 1 - (u0 + r);
            rv[r] = (rs && u0 + r < kOut) ? __ldg(rs + (size_t)t * kBsC) : 0.f;
          }
#pragma unroll
          for (int r = 0; r < kBsR; ++r) acc[r] = u0 + r < kOut ? win[u0 + r] : 0.f;
#pragma unroll
          for (int i = 0; i < 2 * kD + 1; ++i) {
#pragma unroll
            for (int r = 0; r < kBsR; ++r) {
              const int col = u0 + r + i - kD;
              if (u0 + r < kOut && col >= 0) acc[r] = fmaf(tc[i], win[col], acc[r]);
            }
          }
#pragma unroll
          for (int r = 0; r < kBsR; ++r) {
            const int t = part == 0 ? u0 + r : T - 1 - (u0 + r);
            if (u0 + r < kOut) os[(size_t)t * kBsC] = acc[r] + rv[r];
          }
        }
        continue;
      }
      // Both parts fully unrolled with compile-time window indices and edge tests.  (A looped variant -- one 8-output
      // group body, the window shifted by 8 registers and refilled from TMEM between groups -- is 15 times less code
      // but measured 60 % slower; this one is instruction-fetch bound, ncu: 60 % of stall samples "no instruction".)
      auto fir = [&](auto part_tag) {
        constexpr int H = decltype(part_tag)::value;
        constexpr int ha = T * H / NP, hb = T * (H + 1) / NP;             // this thread's outputs
        constexpr int lo_col = ha - HL > 0 ? ha - HL : 0;                 // first / last frame the window needs
        constexpr int hi_col = hb - 1 + N2 < T - 1 ? hb - 1 + N2 : T - 1;
        constexpr int cb = lo_col / 8 * 8;                                // loads start on an 8-column boundary
        constexpr int NL = (hi_col - cb + 1 + 7) / 8 * 8;
        static_assert(cb + NL <= TP, "window load runs past the accumulator");
        float win[NL];
#pragma unroll
        for (int i = 0; i < NL / 8; ++i) tmem_ld8(taddr + (uint32_t)(cb + 8 * i), win + 8 * i);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(b));      // the accumulator is free for the stream after next
#pragma unroll
        for (int j = 0; j < NL; ++j) win[j] = apply_act(win[j] + bias_c, ACT);
        constexpr int kGroups = (hb - ha + kBsR - 1) / kBsR;
#pragma unroll
        for (int gi = 0; gi < kGroups; ++gi) {
          const int t0 = ha + gi * kBsR;
          float rv[kBsR], acc[kBsR];
#pragma unroll
          for (int r = 0; r < kBsR; ++r) rv[r] = (rs && t0 + r < hb) ? __ldg(rs + (size_t)(t0 + r) * kBsC) : 0.f;
#pragma unroll
          for (int r = 0; r < kBsR; ++r) acc[r] = t0 + r < hb ? win[t0 + r - cb] : 0.f;
#pragma unroll
          for (int k = 0; k < N1; ++k) {
#pragma unroll
            for (int r = 0; r < kBsR; ++r) {
              const int col = t0 + r - HL + k;
              if (t0 + r < hb && col >= 0 && col < T) acc[r] = fmaf(tw[k], win[col - cb], acc[r]);
            }
          }
#pragma unroll
          for (int k = 0; k < N2; ++k) {
#pragma unroll
            for (int r = 0; r < kBsR; ++r) {
              const int col = t0 + r + 1 + k;
              if (t0 + r < hb && col < T) acc[r] = fmaf(tw[N1 + k], win[col - cb], acc[r]);
            }
          }
#pragma unroll
          for (int r = 0; r < kBsR; ++r)
            if (t0 + r < hb) os[(size_t)(t0 + r) * kBsC] = acc[r] + rv[r];
        }
      };
      if (MIRROR) {
      } else if (part == 0) fir(std::integral_constant<int, 0>{});
      else if (part == 1) fir(std::integral_constant<int, 1>{});
      else if (NP > 2) fir(std::integral_constant<int, (NP > 2 ? 2 : 0)>{});
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

static size_t bs_smem_bytes(int kc, int T, int n1, int n2) {
  const int tp = (T + 15) / 16 * 16;
  return (size_t)kc * 2 * kBsC * 128 + (size_t)kBsStages * 2 * tp * 128 + 13 * 8 + 16;
}

}  // namespace vadx

using namespace vadx;

// bytes of one stream's operand stages for a K-wide hidden layer: [K/64][hi | lo][round_up(T, 16) rows x 128 B]
size_t fc2_memory_stages_stream_bytes(int n_in, int n_frames) {
  return (size_t)(n_in / kTcBK) * 2 * ((n_frames + 15) / 16 * 16) * 128;
}

bool fc2_memory_stages_supported(int n_in, int n_out, int n_frames, int n_back, int stride_back, int n_ahead, int stride_ahead) {
  if (n_out != kBsC || n_in <= 0 || n_in % kTcBK != 0 || n_in > 256) return false;
  if (n_frames != 98 || n_back != 20 || stride_back != 1) return false;
  if (!(n_ahead == 20 && stride_ahead == 1)) return false;
  return bs_smem_bytes(n_in / kTcBK, n_frames, n_back, n_ahead) <= (size_t)kTcSmemBudget;
}

static int parts_env() {
  static const int parts = ab_env("VADX_BS_PARTS", 2) == 3 ? 3 : 2;
  return parts;
}

int fc2_memory_stages_f32(const void* d_himg, int n_in, const void* d_wimg, const float* d_bias, int act, const float* d_wl,
                          int n_back, const float* d_wr, int n_ahead, const float* d_res, float* d_out, int64_t n_streams,
                          int n_frames, void* stream) {
  // h stages (two bf16 terms = 4 B per value) + residual rows in, out rows
  StageTimer _timer(VADX_STAGE_MEMORY, (cudaStream_t)stream, "fc2_memory_stages_kernel",
                    4.0 * n_streams * n_frames * (n_in + kBsC * (d_res ? 2 : 1)),
                    2.0 * n_streams * n_frames * kBsC * (double)(n_in + n_back + n_ahead));
  VADX_REQUIRE(d_himg && d_wimg && d_wl && d_out, "fc2_memory_stages_f32: null pointer");
  VADX_REQUIRE(fc2_memory_stages_supported(n_in, kBsC, n_frames, n_back, 1, n_ahead, 1), "fc2_memory_stages_f32: shape not supported");
  VADX_REQUIRE(act == VADX_ACT_NONE || act == VADX_ACT_RELU, "fc2_memory_stages_f32: activation %d", act);
  VADX_REQUIRE(aligned16(d_himg) && aligned16(d_wimg), "fc2_memory_stages_f32: operand images must be 16-byte aligned");
  if (n_streams == 0) return VADX_OK;
  static PerDevice per_device;
  int n_sm = 148;
  VADX_TRY(per_device.ensure(&n_sm, [] {
    cudaError_t e = cudaSuccess;
    auto opt_in = [&](auto kern) {
      if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBudget);
    };
    opt_in(fc2_memory_stages_kernel<VADX_ACT_NONE, 98, 20, 20, 2, 7, false>);  opt_in(fc2_memory_stages_kernel<VADX_ACT_RELU, 98, 20, 20, 2, 7, false>);
    opt_in(fc2_memory_stages_kernel<VADX_ACT_NONE, 98, 20, 20, 2, 7, true>);   opt_in(fc2_memory_stages_kernel<VADX_ACT_RELU, 98, 20, 20, 2, 7, true>);
    opt_in(fc2_memory_stages_kernel<VADX_ACT_NONE, 98, 20, 20, 3, 7, false>); opt_in(fc2_memory_stages_kernel<VADX_ACT_RELU, 98, 20, 20, 3, 7, false>);
    return e;
  }));
  BsArgs g{};
  g.Himg = static_cast<const uint8_t*>(d_himg); g.Wimg = static_cast<const uint8_t*>(d_wimg); g.bias = d_bias;
  g.wl = d_wl; g.wr = d_wr; g.res = d_res; g.out = d_out; g.n_streams = n_streams;
  g.kc = n_in / kTcBK; g.n_k16 = n_in / 16;
  static const int pf = ab_env("VADX_BS_PF", 0);   // measured slower on B200: off
  g.pf = pf;
  static const int mirror = ab_env("VADX_BS_MIRROR", 1);
  g.mirror = (mirror && parts_env() == 2) ? 1 : 0;
  const size_t smem = bs_smem_bytes(g.kc, n_frames, n_back, n_ahead);
  const int grid = (int)std::min<int64_t>(n_streams, n_sm);
  const int parts = parts_env();
  cudaStream_t cs = (cudaStream_t)stream;
  if (parts == 2) {
    if (g.mirror) {
      if (act == VADX_ACT_RELU) fc2_memory_stages_kernel<VADX_ACT_RELU, 98, 20, 20, 2, 7, true><<<grid, 10 * 32, smem, cs>>>(g);
      else fc2_memory_stages_kernel<VADX_ACT_NONE, 98, 20, 20, 2, 7, true><<<grid, 10 * 32, smem, cs>>>(g);
    } else if (act == VADX_ACT_RELU) fc2_memory_stages_kernel<VADX_ACT_RELU, 98, 20, 20, 2, 7, false><<<grid, 10 * 32, smem, cs>>>(g);
    else fc2_memory_stages_kernel<VADX_ACT_NONE, 98, 20, 20, 2, 7, false><<<grid, 10 * 32, smem, cs>>>(g);
  } else {
    if (act == VADX_ACT_RELU) fc2_memory_stages_kernel<VADX_ACT_RELU, 98, 20, 20, 3, 7, false><<<grid, 14 * 32, smem, cs>>>(g);
    else fc2_memory_stages_kernel<VADX_ACT_NONE, 98, 20, 20, 3, 7, false><<<grid, 14 * 32, smem, cs>>>(g);
  }
  return after_launch("fc2_memory_stages_f32");
}

extern "C" size_t vadx_fc2_memory_stages_stream_bytes(int n_in, int n_frames) { return fc2_memory_stages_stream_bytes(n_in, n_frames); }
extern "C" int vadx_fc2_memory_stages_supported(int n_in, int n_out, int n_frames, int n_back, int stride_back, int n_ahead,
                                                int stride_ahead) {
  return fc2_memory_stages_supported(n_in, n_out, n_frames, n_back, stride_back, n_ahead, stride_ahead) ? 1 : 0;
}
extern "C" int vadx_fc2_memory_stages_f32(const void* d_himg, int n_in, const void* d_wimg, const float* d_bias, int act,
                                          const float* d_wl, int n_back, const float* d_wr, int n_ahead, const float* d_residual,
                                          float* d_out, int64_t n_streams, int n_frames, void* stream) {
  VADX_REQUIRE(n_streams >= 0, "vadx_fc2_memory_stages_f32: negative stream count");
  return fc2_memory_stages_f32(d_himg, n_in, d_wimg, d_bias, act, d_wl, n_back, d_wr, n_ahead, d_residual, d_out, n_streams,
                               n_frames, stream);
}
