// block_stages.cu -- the second half of a DFSMN block in ONE kernel: dense layer (tcgen05) + memory block (FIR) + residual,
// fed with ready-made operand stages, the FIR running straight out of tensor memory.
//
//   p   = act(h W^T + b)                                                    (FireRedVAD/Export_FireRedVAD.py:253-263, :290-296)
//   out = p + sum_k wl[c][k] p[t-(N1-1)+k] + sum_k wr[c][k] p[t+1+k]  (+ res)      (FSMN.forward, :213-236)
//
//   * h arrives as per-stream operand stages ([stream][K chunk][hi | lo] swizzled bf16 images written by fc1's epilogue,
//     gemm_tc.cu y_split == 2): one thread streams them in with cp.async.bulk, there are no loader warps;
//   * the MMA is issued TRANSPOSED: A = the weight image (M = 128 output channels), B = the stream's h stage
//     (N = round_up(T, 16) frames), so the accumulator holds p^T -- TMEM lane = channel, TMEM column = frame.  A thread
//     reads ITS channel's whole time series with tcgen05.ld and runs the FIR along it in registers: no p tile at all,
//     and since a chunk is zero-padded on both sides by definition there is no halo either;
//   * T is a template parameter: every window index and every edge test is a compile-time constant;
//   * the second half of the frames is processed in reversed time, so both halves run one shared unrolled body.
// p never exists in HBM (the unfused pair writes and re-reads 0.5 KB per frame row), and the eight FIR warps overlap the
// next stream's MMAs through two accumulators.  Callers: model.cu (FireRed, engine.fuse_stages); shapes it does not
// cover (other chunk lengths or tap counts, streaming caches) take the two-kernel tail.
//
// History: block_tc.cu (shared-memory p tile; slower than the unfused pair), then the first
// kernel of this file (scalar FIR, residual rows through per-thread loads: 0.38 ms per launch at 8192 streams), both removed.
#include <type_traits>

#include "tc_ptx.cuh"

namespace vadx {

constexpr int kBsC = 128;                            // channels: N of the dense layer = M of the transposed MMA
struct BsArgs {
  const uint8_t* Himg;   // [S][kc][2][img bytes]
  const uint8_t* Wimg;   // [kc][2][128 x 128 B]
  const float* bias;     // [128] or null
  const float* wl;       // [128][N1]
  const float* wr;       // [128][N2]
  const float* res;      // [S*T][128] or null
  float* out;            // [S*T][128]
  int64_t n_streams;
  int kc, n_k16;
  int dbg;      // measurement builds only (make AB=1): 1 no tap arithmetic, 2 no residual, 4 no stores, 8 no operand copies
};

// ------------------------------------------------------------------------------------------------------------------------------
// Every HBM read a bulk copy, FIR on the packed fp32x2 FMA.  What the previous version (scalar FIR, residual rows through
// per-thread loads) was bound by, measured by switching its parts off one at a time (tools/block_microbench.py,
// `make AB=1`, 8192 streams; whole kernel 0.38 ms against 0.25 ms of HBM time):
//   * the residual rows: per-thread loads issued one register group (~0.2 us of arithmetic) before their use -- the FIR warps
//     sat out an HBM latency per group.  Without them 0.28 ms.  Requesting them a whole stream ahead through registers does
//     not help (0.43 ms): a warp has six scoreboards, so a wait on last stream's rows also waits for the rows requested a
//     moment ago.  Only an mbarrier-tracked bulk copy into shared memory gets them off the warps' critical path;
//   * its size: 46 KB of unrolled FIR code.  The packed form below (1799 FFMA -> 968 FFMA2 per stream and thread) moved the
//     kernel by 6 % on its own, and every measurement variant whose loop body grew past ~30 KB lost 10-15 % again
//     (an `AB=1` build of THIS kernel, with its run-time debug branches, runs at 0.40 ms; the production build at 0.295).
// So: shared memory holds the resident weight image, a ring of NS operand stages and a ring of NR residual PIECES (the rows
// one register group of both halves needs: 2 x 7 rows = 7 KB; a whole 50 KB tile does not fit beside the 131 KB of W), filled
// by two producer threads with cp.async.bulk.  The FIR warps touch global memory only to store.
// Tried and dropped: W resident in tensor memory as the A operand (`tcgen05.mma [d], [a_tmem], b_desc`; frees 131 KB of
// shared memory for whole residual tiles and four operand stages; results identical) -- no faster in a measurement build.
//
// The FIR: a thread owns one TMEM lane (= one channel), so the two halves of a packed FMA cannot be two channels as in
// memory_bulk.cu.  They are two TAPS of the same output instead:
//   * the window is held as aligned pairs WP[k] = (P[2k], P[2k+1]) -- exactly how tcgen05.ld delivers two columns;
//   * an even output u runs over the tap pairs E[m] = (c[2m], c[2m+1]) (window pair (u + 2m - D) / 2), an odd output over
//     O[m] = (c[2m-1], c[2m]) (window pair (u + 2m - D - 1) / 2): 21 FFMA2 instead of 41 FFMA per output, the two half sums
//     are added once at the end.  The taps are held twice (E and O: 84 registers), the window once;
//   * the mirrored half (P[u] = p[T-1-u]) gets its pairs from TMEM in swapped order (columns T-2-2k, T-1-2k), which is
//     absorbed by swapping the halves of its tap pairs when they are loaded: both halves of the frames still run ONE body;
//   * `out = res + p + FIR(p)` keeps p as its own exact term: the chain starts as WP[u/2] * (1, 0) (or (0, 1)) + (res, 0).
// 84 + 70 window + accumulators do not fit the 168 registers a 12-warp allocation leaves per thread, so the CTA is three
// warpgroups: producer + MMA issuer (+ two idle warps) give registers back (setmaxnreg.dec), the eight FIR warps take them
// (setmaxnreg.inc).
typedef unsigned long long bs_u64;
__device__ __forceinline__ bs_u64 bs_pack(float x, float y) {
  bs_u64 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(x), "f"(y));
  return d;
}
__device__ __forceinline__ void bs_unpack(bs_u64 v, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ bs_u64 bs_ffma2(bs_u64 a, bs_u64 b, bs_u64 c) {
  bs_u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ bs_u64 bs_fadd2(bs_u64 a, bs_u64 b) {
  bs_u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// two consecutive fp32 columns of this thread's TMEM lane; the registers are valid after a wait took them
__device__ __forceinline__ void bs_tmem_ld2_issue(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr));
}
__device__ __forceinline__ void bs_tmem_wait2(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]) : : "memory");
}
__device__ __forceinline__ bs_u64 bs_pack_u32(uint32_t x, uint32_t y) {
  bs_u64 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(x), "r"(y));
  return d;
}
// mbar_wait with the polling loop out of line: the FIR body waits at eight places per stream and is instruction-cache sensitive
__device__ __noinline__ void bs_wait_slow(uint32_t bar, uint32_t parity, unsigned backoff_ns) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(backoff_ns);
    if (clock64() - t0 > 20000000000LL) __trap();
  }
}
__device__ __forceinline__ void bs_wait(uint32_t bar, uint32_t parity, unsigned backoff_ns) {
  if (!mbar_try_wait(bar, parity)) bs_wait_slow(bar, parity, backoff_ns);
}

constexpr int kBpRegsLow = 40, kBpRegsHigh = 232;   // 128 x 40 + 256 x 232 = 384 x 168
constexpr int kBpBars = 32;                         // full[4] empty[4] tfull[2] tempty[2] rfull[8] rempty[8] wbar

template <int ACT, int T, int ND, int R, int NS, int NR>   // ND taps on each side (N1 == N2), R outputs per register group, NS operand stages, NR residual pieces
__global__ void __launch_bounds__(384, 1) fc2_memory_stages_kernel(const BsArgs g) {
  static_assert(T % 2 == 0 && ND % 2 == 0, "the pair form needs an even frame count and an even look-back / look-ahead");
  static_assert(NS >= 2 && NS <= 4 && NR >= 2 && NR <= 8, "operand stages / residual pieces");
  static_assert((T / 2) % R == 0, "a half of the frames must be whole register groups (the residual pieces are cut along them)");
  constexpr int TP = (T + 15) / 16 * 16;
  constexpr int kImg = TP * 128, kStage = 2 * kImg, kCopy = T * 128;
  constexpr int kWImg = kBsC * 128;
  constexpr int kResHalf = R * kBsC * 4, kResPiece = 2 * kResHalf;   // one register group's rows of both halves
  constexpr int kFirWarps = 8;
  static_assert(kImg % 1024 == 0, "stage images must keep the 1024-byte swizzle atom alignment");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* w_smem = smem_raw;                                   // W: [K chunk][hi | lo] images, resident
  const int w_bytes = g.kc * 2 * kWImg;
  uint8_t* a_smem = w_smem + w_bytes;                           // NS stages of [hi | lo] images
  uint8_t* r_smem = a_smem + NS * kStage;                       // NR residual pieces: [front R rows | back R rows][128] fp32
  uint64_t* bars = reinterpret_cast<uint64_t*>(r_smem + NR * kResPiece);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kBpBars);
#ifdef VADX_AB_SWITCHES
  const int dbg = g.dbg;
#else
  constexpr int dbg = 0;
#endif
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (4 + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (8 + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (10 + b); };
  auto rfull_bar = [&](int b) { return bar0 + 8u * (12 + b); };
  auto rempty_bar = [&](int b) { return bar0 + 8u * (20 + b); };
  const uint32_t wbar = bar0 + 8u * 28;
  constexpr uint32_t kTmemCols = 256;   // two accumulators, 128 columns apart
  constexpr uint32_t kAccCol = 0;
  const bool use_res = g.res != nullptr && !(dbg & 2);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), kFirWarps);
    }
    for (int b = 0; b < NR; ++b) {
      mbar_init(rfull_bar(b), 1);
      mbar_init(rempty_bar(b), kFirWarps);
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

  if (warp < 4) {
    // warpgroup 0: producer, MMA issuer, two idle warps -- all four hand their registers to the FIR warpgroups
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kBpRegsLow));
    if (warp == 0 && lane == 0) {
      // ===================== operand producer: one stream's K chunk = two bulk copies (the T real rows of each image) =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t s = blockIdx.x; s < g.n_streams; s += gridDim.x) {
        for (int c = 0; c < g.kc; ++c) {
          mbar_wait(empty_bar(stage), phase ^ 1u, 32);
          if (dbg & 8) {
            mbar_arrive(full_bar(stage));
          } else {
            mbar_expect_tx(full_bar(stage), 2u * kCopy);
            const uint8_t* src = g.Himg + ((size_t)(s * g.kc + c) * 2) * kImg;
            const uint32_t dst = smem_u32(a_smem) + (uint32_t)stage * kStage;
            bulk_g2s(dst, src, kCopy, full_bar(stage));
            bulk_g2s(dst + kImg, src + kImg, kCopy, full_bar(stage));
          }
          if (++stage == NS) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ===================== MMA issuer: p^T = W h^T, three products of the two-term split =====================
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
        const uint32_t d_tmem = tmem_base + kAccCol + (uint32_t)(b * 128);
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
          if (++stage == NS) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(b));
      }
    } else if (warp == 2 && lane == 0 && use_res) {
      // ===================== residual producer: per stream and register group the R rows of each half =====================
      int slot = 0;
      uint32_t phase = 0;
      for (int64_t s = blockIdx.x; s < g.n_streams; s += gridDim.x) {
        const uint8_t* rows = reinterpret_cast<const uint8_t*>(g.res + (size_t)s * T * kBsC);
        for (int gi = 0; gi < (T / 2) / R; ++gi) {
          mbar_wait(rempty_bar(slot), phase ^ 1u, 32);
          mbar_expect_tx(rfull_bar(slot), (uint32_t)kResPiece);
          const uint32_t dst = smem_u32(r_smem) + (uint32_t)slot * kResPiece;
          bulk_g2s(dst, rows + (size_t)gi * kResHalf, kResHalf, rfull_bar(slot));                          // frames gi R .. gi R + R - 1
          // the mirrored half's rows in ITS order (frames T-1 - gi R, descending): row by row, so that both halves read
          // their piece with the same ascending code
          for (int r = 0; r < R; ++r)
            bulk_g2s(dst + kResHalf + r * (kBsC * 4), rows + (size_t)(T - 1 - gi * R - r) * kBsC * 4, kBsC * 4, rfull_bar(slot));
          if (++slot == NR) { slot = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ===================== FIR warps (warpgroups 1, 2): thread = (channel, half of the time axis) =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kBpRegsHigh));
    const int q = warp & 3;                 // TMEM lane quarter this warp may touch
    const int part = (warp - 4) >> 2;       // warps 4..7: first half of the frames, 8..11: second half in reversed time
    const int c = q * 32 + lane;
    constexpr int kOut = T / 2;                               // outputs per half, canonical time u = t or T-1-t
    constexpr int kPairs = (kOut - 1 + ND) / 2 + 1;           // window pairs: P[0 .. kOut-1+ND] (+ one zero-tap neighbour)
    static_assert(2 * kPairs <= T, "the window of one half runs past the chunk");
    // canonical taps over offsets d = i - ND, i = 0 .. 2 ND: forward c_d = wl[ND-1+d] (d <= 0) | wr[d-1] (d > 0), zero at
    // d = -ND; the mirrored half reads them with d -> -d.  Outside [0, 2 ND]: zero.
    auto ctap = [&](int i) -> float {
      if (i < 0 || i > 2 * ND) return 0.f;
      const int d = part == 0 ? i - ND : ND - i;
      return d == -ND ? 0.f : (d <= 0 ? __ldg(g.wl + c * ND + (ND - 1 + d)) : __ldg(g.wr + c * ND + (d - 1)));
    };
    // pair halves follow the window pairs: (P[2k], P[2k+1]) forward, (P[2k+1], P[2k]) mirrored
    bs_u64 te[ND + 1], to[ND + 1];
#pragma unroll
    for (int m = 0; m <= ND; ++m) {
      const float e0 = ctap(2 * m), e1 = ctap(2 * m + 1), o0 = ctap(2 * m - 1);
      te[m] = part == 0 ? bs_pack(e0, e1) : bs_pack(e1, e0);
      to[m] = part == 0 ? bs_pack(o0, e0) : bs_pack(e0, o0);
    }
    const bs_u64 sel_even = part == 0 ? bs_pack(1.f, 0.f) : bs_pack(0.f, 1.f);   // picks P[u] out of its pair, u even
    const bs_u64 sel_odd = part == 0 ? bs_pack(0.f, 1.f) : bs_pack(1.f, 0.f);
    const bs_u64 zero2 = bs_pack(0.f, 0.f);
    const float bias_c = g.bias ? __ldg(g.bias + c) : 0.f;
    const bs_u64 bias2 = bs_pack(bias_c, bias_c);
    const uint32_t col0 = part == 0 ? 0u : (uint32_t)(T - 2);     // TMEM column of window pair k: col0 +- 2 k
    const int cstep = part == 0 ? 2 : -2;
    const int64_t t0 = part == 0 ? 0 : T - 1;                      // frame of canonical time u: t0 + u (forward) | t0 - u (mirrored)
    constexpr int kGroups = (kOut + R - 1) / R;
    const float* rpiece = reinterpret_cast<const float*>(r_smem) + (part == 0 ? 0 : R * kBsC) + c;   // this half's rows of the current piece
    uint32_t rfull = rfull_bar(0);                                 // its barrier (rempty is eight barriers on)
    int rslot = 0;
    uint32_t rphase = 0;
    int it = 0;
    for (int64_t s = blockIdx.x; s < g.n_streams; s += gridDim.x, ++it) {
      const int b = it & 1;
      const uint32_t par = ((uint32_t)(it >> 1)) & 1u;
      bs_wait(tfull_bar(b), par, 64);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + kAccCol + (uint32_t)(b * 128) + col0;
      // row u of a half: the offset is an immediate of the load / store in each direction (a run-time stride costs a 64-bit
      // address computation per access)
      float* os = g.out + ((size_t)s * T + t0) * kBsC + c;
      // all loads issued, then collected (one TMEM latency per stream); every register is a read-write operand of a wait,
      // so no use can be scheduled above it
      uint32_t raw[2 * kPairs];
#pragma unroll
      for (int k = 0; k < kPairs; ++k) bs_tmem_ld2_issue(taddr + (uint32_t)(k * cstep), raw + 2 * k);
#pragma unroll
      for (int k = 0; k < kPairs; k += 4) {
        if (k + 3 < kPairs) tmem_ld8_finish(raw + 2 * k);
        else
#pragma unroll
          for (int kk = k; kk < kPairs; ++kk) bs_tmem_wait2(raw + 2 * kk);
      }
      bs_u64 wp[kPairs];
#pragma unroll
      for (int k = 0; k < kPairs; ++k) wp[k] = bs_pack_u32(raw[2 * k], raw[2 * k + 1]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(b));      // the accumulator is free for the stream after next
#pragma unroll
      for (int k = 0; k < kPairs; ++k) {
        wp[k] = bs_fadd2(wp[k], bias2);
        if (ACT == VADX_ACT_RELU) {
          float x, y;
          bs_unpack(wp[k], x, y);
          wp[k] = bs_pack(fmaxf(x, 0.f), fmaxf(y, 0.f));
        }
      }
#pragma unroll
      for (int gi = 0; gi < kGroups; ++gi) {
        const int u0 = gi * R;
        float rv[R];
#pragma unroll
        for (int r = 0; r < R; ++r) rv[r] = 0.f;
        if (use_res) {
          // the group's piece: front rows ascending for the forward half, back rows (ascending frames) read backwards for the mirrored one
          bs_wait(rfull, rphase, 32);
#pragma unroll
          for (int r = 0; r < R; ++r) rv[r] = rpiece[r * kBsC];
        }
        // chain start: (p[u], 0) or (0, p[u]) -- p[u] picked out of its window pair by a (1, 0) / (0, 1) factor
        bs_u64 acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int u = u0 + r;
          acc[r] = u < kOut ? bs_ffma2(wp[u >> 1], (u & 1) ? sel_odd : sel_even, zero2) : zero2;
        }
#pragma unroll
        for (int m = 0; m <= ND; ++m) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int u = u0 + r;
            const int j = u + 2 * m - ND - (u & 1);             // even: the window pair (P[j], P[j+1])
            if (u < kOut && j >= 0 && !((dbg & 1) && m != ND / 2)) acc[r] = bs_ffma2((u & 1) ? to[m] : te[m], wp[j >> 1], acc[r]);
          }
        }
        float ov[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float x, y;
          bs_unpack(acc[r], x, y);
          ov[r] = (x + y) + rv[r];
        }
        if (!((dbg & 4) && ov[0] != 1234.5f)) {
          if (part == 0) {
#pragma unroll
            for (int r = 0; r < R; ++r)
              if (u0 + r < kOut) os[(u0 + r) * kBsC] = ov[r];
          } else {
#pragma unroll
            for (int r = 0; r < R; ++r)
              if (u0 + r < kOut) os[-(u0 + r) * kBsC] = ov[r];
          }
        }
        if (use_res) {
          // The piece is released only HERE, behind the stores that depend on its values.  An arrive right after the
          // shared-memory loads were ISSUED let the refill (a bulk copy that hits L2) land before the slowest warp's loads
          // had executed: the first group of a CTA then saw rows of the piece six groups ahead (caught by
          // test_block_pair_from_stages_against_float64 at 300 streams, K = 128, where the FIR warps are the bottleneck and
          // the producer refills at once).  A release-arrive cannot move above the stores, the stores need the loaded values.
          __syncwarp();
          if (lane == 0) mbar_arrive(rfull + 8u * 8u);          // rempty of the same slot
          rfull += 8u;
          rpiece += kResPiece / 4;
          if (++rslot == NR) {
            rslot = 0; rphase ^= 1u;
            rfull -= 8u * NR;
            rpiece -= NR * (kResPiece / 4);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

constexpr int kBpR = 7;   // outputs per register group = rows of one half in a residual piece
static size_t bs_smem_bytes(int kc, int T, int n_stages, int n_res) {
  const int tp = (T + 15) / 16 * 16;
  return (size_t)kc * 2 * kBsC * 128 + (size_t)n_stages * 2 * tp * 128 + (size_t)n_res * 2 * kBpR * kBsC * 4 + kBpBars * 8 + 16;
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
  return bs_smem_bytes(n_in / kTcBK, n_frames, 2, 4) <= (size_t)kTcSmemBudget;
}

int fc2_memory_stages_f32(const void* d_himg, int n_in, const void* d_wimg, const float* d_bias, int act, const float* d_wl,
                          int n_back, const float* d_wr, int n_ahead, const float* d_res, float* d_out, int64_t n_streams,
                          int n_frames, void* stream) {
  // h stages (two bf16 terms = 4 B per value) + residual rows in, out rows
  StageTimer _timer(VADX_STAGE_MEMORY, (cudaStream_t)stream, "fc2_memory_stages_kernel",
                    4.0 * n_streams * n_frames * (n_in + kBsC * (d_res ? 2 : 1)),
                    2.0 * n_streams * n_frames * kBsC * (double)(n_in + n_back + n_ahead));
  VADX_REQUIRE(d_himg && d_wimg && d_wl && d_wr && d_out, "fc2_memory_stages_f32: null pointer");
  VADX_REQUIRE(fc2_memory_stages_supported(n_in, kBsC, n_frames, n_back, 1, n_ahead, 1), "fc2_memory_stages_f32: shape not supported");
  VADX_REQUIRE(act == VADX_ACT_NONE || act == VADX_ACT_RELU, "fc2_memory_stages_f32: activation %d", act);
  VADX_REQUIRE(aligned16(d_himg) && aligned16(d_wimg) && aligned16(d_res), "fc2_memory_stages_f32: operand images and residual rows must be 16-byte aligned");
  if (n_streams == 0) return VADX_OK;
  static PerDevice per_device;
  int n_sm = 148;
  VADX_TRY(per_device.ensure(&n_sm, [] {
    cudaError_t e = cudaSuccess;
    auto opt_in = [&](auto kern) {
      if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBudget);
    };
    opt_in(fc2_memory_stages_kernel<VADX_ACT_NONE, 98, 20, kBpR, 2, 6>); opt_in(fc2_memory_stages_kernel<VADX_ACT_RELU, 98, 20, kBpR, 2, 6>);
    opt_in(fc2_memory_stages_kernel<VADX_ACT_NONE, 98, 20, kBpR, 2, 4>); opt_in(fc2_memory_stages_kernel<VADX_ACT_RELU, 98, 20, kBpR, 2, 4>);
    return e;
  }));
  BsArgs g{};
  g.Himg = static_cast<const uint8_t*>(d_himg); g.Wimg = static_cast<const uint8_t*>(d_wimg); g.bias = d_bias;
  g.wl = d_wl; g.wr = d_wr; g.res = d_res; g.out = d_out; g.n_streams = n_streams;
  g.kc = n_in / kTcBK; g.n_k16 = n_in / 16;
  g.dbg = ab_env("VADX_BS_DBG", 0);
  // six residual pieces (43 KB ahead of the FIR warps) when they fit beside the weight image, else four
  static const int want_res = ab_env("VADX_BS_RES", 6) == 4 ? 4 : 6;
  const int nr = bs_smem_bytes(g.kc, n_frames, 2, want_res) <= (size_t)kTcSmemBudget ? want_res : 4;
  const size_t smem = bs_smem_bytes(g.kc, n_frames, 2, nr);
  const int grid = (int)std::min<int64_t>(n_streams, n_sm);
  cudaStream_t cs = (cudaStream_t)stream;
  auto launch = [&](auto kern) { kern<<<grid, 12 * 32, smem, cs>>>(g); };
  const bool relu = act == VADX_ACT_RELU;
  if (nr == 6) {
    if (relu) launch(fc2_memory_stages_kernel<VADX_ACT_RELU, 98, 20, kBpR, 2, 6>);
    else launch(fc2_memory_stages_kernel<VADX_ACT_NONE, 98, 20, kBpR, 2, 6>);
  } else {
    if (relu) launch(fc2_memory_stages_kernel<VADX_ACT_RELU, 98, 20, kBpR, 2, 4>);
    else launch(fc2_memory_stages_kernel<VADX_ACT_NONE, 98, 20, kBpR, 2, 4>);
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
