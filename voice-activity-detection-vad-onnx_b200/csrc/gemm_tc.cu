// gemm_tc.cu -- dense layers (a8) on the 5th-generation tensor cores: tcgen05.mma with TMEM
// accumulators, weights-stationary persistent CTAs, fp32-grade accuracy through a two-term bf16
// split of both operands.
//
//   Y[rows][N] = act(X[rows][K] * W^T + bias) (+ residual)       X, Y fp32 row-major in HBM
//
// Numerics: x = xh + xl, w = wh + wl with xh = bf16(x), xl = bf16(x - xh) (same for w).  The tile
// product is accumulated in fp32 TMEM as  xh*wh + xl*wh + xh*wl ; the dropped xl*wl term and the
// rounding of the low parts are ~2^-17 relative per product, i.e. fp32-grade for the <=1e-3
// frame-probability budget (measured: see tests/test_gpu_tc.py).
//
// Layout: operands live in shared memory as K-major, 128-byte-swizzled bf16 tiles of 64 k-columns
// (one swizzle atom wide): element (row r, k) of a tile sits at byte
//     (r/8)*1024 + (r%8)*128 + (((k/8) ^ (r%8))*16) + (k%8)*2 .
// The weights are packed into exactly that image ONCE on the host (vadx_pack_weight_tc), so a CTA
// pulls its whole stationary operand with a handful of 1-D bulk async copies (cp.async.bulk ->
// mbarrier complete_tx); activations are converted fp32 -> (hi, lo) bf16 by the loader warps.
//
// CTA = 21 warps: 0-15 loaders (global fp32 -> split -> swizzled smem; 8 in the LW = 8 measurement variant), 16-19
// epilogue (TMEM -> regs -> bias/act/residual -> global), 20 = TMEM allocator + single-thread MMA issuer.  Two TMEM
// accumulators ping-pong so the epilogue of tile i overlaps the MMAs of tile i+1; a 2-4 deep
// mbarrier ring decouples the loaders from the MMA issuer.  The loaders keep TWO 32 KB stages of
// global loads in flight per SM (three register buffers per thread): with one stage in flight the
// kernel sat at Little's-law bandwidth (~32 KB / ~1.3 us per SM = 0.57 of the HBM peak).
//
// Epilogues (template parameter YB; DESIGN.md section 3 "Bulk-store epilogues"): the wide layers were bound by their STORE
// INSTRUCTIONS, so the two hot output forms leave through the bulk-copy engine -- YB = 1: per-stream operand stages for the
// fused block tail (block_stages.cu), assembled in shared memory as they lie in the images and written with 1-D
// cp.async.bulk shared -> global; YB = 2: plain fp32 rows through a 2-D tensor map (cp.async.bulk.tensor, 32 x 32 boxes,
// edges clipped by the copy engine).  YB = 0 keeps the per-lane paths: residual input, fused head, K-split accumulation,
// unaligned outputs, and the operand-stage forms the bulk path does not cover.
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, nothing links against libcuda)

#include "tc_ptx.cuh"

namespace vadx {

constexpr int kTcStageBytes = 2 * kTcTileBytes;  // hi + lo
// loader warps LW (8 or 16, a template parameter): epilogue warps LW..LW+3 ((warp & 3) = TMEM lane quarter), MMA warp LW+4
constexpr int kTcOutLd = 36;  // floats per staged row: 32 columns + 4 pad (144 B: conflict-free 16-byte accesses)

struct TcArgs {
  const float* X;
  int64_t ldx;
  const uint8_t* Wimg;
  const float* bias;
  const float* res;
  int64_t ldr;
  float* Y;
  int64_t ldy;
  int64_t M;
  int K, N, n_pad, kc, n_k16, act, n_stages, n_tiles, tmem_cols, vec_x, vec_y;
  unsigned backoff_ld, backoff_epi;   // nanoseconds between mbarrier probes of the loader / epilogue warps
  int pf_tiles;   // L2 prefetch distance in row tiles of this CTA (0 = off)
  int k_live;     // columns < k_live are read by the MMA (n_k16 * 16)
  int x_split;    // 1: X is a sequence of ready-made operand stages [tile][k chunk][hi | lo] (16 KB swizzled bf16 images
                  //    written by a y_split producer): one thread streams them in with cp.async.bulk, no loader warps
  int y_split;    // 1: Y is written as such stages for a consumer with K = N (N % 64 == 0), rows padded to whole tiles
                  // 2: the same, but one image set per STREAM of ys_T rows ([stream][N/64][hi | lo], images of ys_img
                  //    bytes = round_up(ys_T, 16) rows): the fused fc2 + memory kernel's operand (block_stages.cu)
  int ys_T, ys_img;
  int y_bulk;     // y_split == 2 only: the epilogue assembles each warp's 32 rows of a K chunk in shared memory exactly as they lie in
                  // the images and one lane writes them with bulk copies (shared -> global), 4 KB per copy
  int pf_spread, pf_at;   // pf_at: the K step of the current tile at which the prefetch is issued
  const float* head_w;   // optional fused 1-output head: out = sigmoid(sum_n act(y[n]) * head_w[n] + head_b)
  float head_b;
  float* head_out;       // [rows]; when set the N-wide output itself is not written
  int debug;  // bit0: skip global stores, bit1: skip global loads (VADX_TC_DEBUG, perf experiments only)
};

// bias + activation of one accumulator value; LOG_CLAMP reads the bias slot as the floor
template <int ACT>
__device__ __forceinline__ float bias_act(float v, float b) {
  // __logf = lg2.approx * ln 2: relative error 2^-22 of the log2, i.e. < 1e-5 absolute on log-mel values (|ln| < 40) --
  // below the operand split's own error; the library logf costs ~20 dependent instructions per element and made the
  // four epilogue warps the bottleneck of the mel layer
  if (ACT == VADX_ACT_LOG_CLAMP) return __logf(fmaxf(v, b));
  if (ACT == VADX_ACT_LOG) return __logf(v + b);
  return apply_act(v + b, ACT);
}

static int lin_loader_warps() {
  static const int lw = [] { const int v = ab_env("VADX_LIN_LOADERS", 16); return v == 8 || v == 164 ? v : 16; }();
  return lw;
}
static bool lw_is16() { return lin_loader_warps() != 8; }

// ------------------------------------------------------------------------------------------ kernel
// YB: the epilogue is ONE bulk-store writer and nothing else (its own instantiations, so that the register allocation of the
// general kernel does not change): 1 = per-stream operand stages (1-D bulk copies), 2 = fp32 rows through a 2-D tensor map
// (cp.async.bulk.tensor: 32 x 32 boxes, rows / columns past the end clipped by the copy engine).
template <int ACT, int LW, int NB = 3, int YB = 0>  // activation code is a compile-time constant: the per-element epilogue must not carry the sigmoid path around
__global__ void __launch_bounds__((LW + 5) * 32, 1) linear_tc_kernel(const TcArgs g, const __grid_constant__ CUtensorMap ymap) {
  constexpr int kTcLoaderWarps = LW, kTcEpiWarp0 = LW, kTcMmaWarp = LW + 4;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve (base is 1024-aligned by the launch: dynamic smem starts at a 1024-aligned offset
  // because the kernel has no static shared memory)
  uint8_t* w_smem = smem_raw;
  const int w_bytes = g.kc * 2 * g.n_pad * 128;
  uint8_t* a_smem = w_smem + w_bytes;
  // YB: the staging tiles come first (1024-byte aligned: the tensor-map copies read them 128-byte swizzled)
  float* stage_out = reinterpret_cast<float*>(a_smem + (size_t)g.n_stages * kTcStageBytes);  // 4 warps x (32 rows x 36 floats | 8 KB)
  float* bias_s = stage_out + (YB ? 4 * 8192 / 4 : 4 * 32 * kTcOutLd);
  float* head_s = bias_s + g.n_pad;  // [n_pad] fused-head weights (zeros when unused)
  uint64_t* bars = reinterpret_cast<uint64_t*>(head_s + g.n_pad);
  // bars: full[4], empty[4], tmem_full[2], tmem_empty[2], wbar
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (4 + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (8 + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (10 + b); };
  const uint32_t wbar = bar0 + 8u * 12;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(full_bar(s), g.x_split ? 1 : kTcLoaderWarps);     // one elected lane per loader warp / the bulk-copy thread
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 4);   // one elected lane per epilogue warp
    }
    mbar_init(wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < g.n_pad; i += blockDim.x) {
    bias_s[i] = (g.bias && i < g.N) ? g.bias[i] : 0.f;
    head_s[i] = (g.head_w && i < g.N) ? g.head_w[i] : 0.f;
  }
  if (warp == kTcMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)g.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kTcMmaWarp) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      // stationary operand: the packed weight image, kc*2 tiles of n_pad*128 bytes
      mbar_expect_tx(wbar, (uint32_t)w_bytes);
      const uint32_t img_bytes = (uint32_t)g.n_pad * 128u;
      for (int t = 0; t < g.kc * 2; ++t)
        bulk_g2s(smem_u32(w_smem) + t * img_bytes, g.Wimg + (size_t)t * img_bytes, img_bytes, wbar);
      mbar_wait(wbar, 0);
      const uint32_t idesc = umma_idesc_bf16(g.n_pad);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait(tempty_bar(b), (use & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(b * g.n_pad);
        for (int c = 0; c < g.kc; ++c) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(a_smem) + (uint32_t)stage * kTcStageBytes;
          const uint32_t a_lo = a_hi + kTcTileBytes;
          const uint32_t w_hi = smem_u32(w_smem) + (uint32_t)(c * 2) * img_bytes;
          const uint32_t w_lo = w_hi + img_bytes;
          const int nk = min(4, g.n_k16 - c * 4);
          const uint64_t da_hi = umma_desc_sw128(a_hi), da_lo = umma_desc_sw128(a_lo);
          const uint64_t dw_hi = umma_desc_sw128(w_hi), dw_lo = umma_desc_sw128(w_lo);
          umma_k64(d_tmem, da_hi, dw_hi, idesc, c ? 1u : 0u, nk);
          umma_k64(d_tmem, da_lo, dw_hi, idesc, 1u, nk);
          umma_k64(d_tmem, da_hi, dw_lo, idesc, 1u, nk);
          umma_commit(empty_bar(stage));  // frees the activation stage once these MMAs retire
          if (++stage == g.n_stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(b));        // accumulator complete -> epilogue
      }
    }
  } else if (warp < kTcLoaderWarps && g.x_split) {
    // ===================== ready-made operand stages: one thread, one 32 KB bulk copy pair per stage =====================
    if (threadIdx.x == 0) {
      const uint8_t* img = reinterpret_cast<const uint8_t*>(g.X);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
        for (int c = 0; c < g.kc; ++c) {
          mbar_wait(empty_bar(stage), phase ^ 1u, g.backoff_ld);
          mbar_expect_tx(full_bar(stage), (uint32_t)kTcStageBytes);
          const uint8_t* src = img + ((size_t)tile * g.kc + c) * kTcStageBytes;
          const uint32_t dst = smem_u32(a_smem) + (uint32_t)stage * kTcStageBytes;
          bulk_g2s(dst, src, kTcTileBytes, full_bar(stage));
          bulk_g2s(dst + kTcTileBytes, src + kTcTileBytes, kTcTileBytes, full_bar(stage));
          if (++stage == g.n_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp < kTcLoaderWarps) {
    // ===================== loaders: fp32 rows -> (hi, lo) bf16, swizzled =====================
    // 256 threads: thread = (16-byte k chunk, row mod 32), 4 rows per stage.  Three register buffers
    // rotate (issue for step i+2, convert step i), so two stages of loads are always in flight;
    // rows past the end are clamped to a valid row and zeroed afterwards.
    constexpr int kPasses = kTcBM / (kTcLoaderWarps * 4);   // 4
    const int t = threadIdx.x;   // 0..255
    const int kq = t & 7;        // 16-byte chunk (8 bf16) within the 64-k atom row
    const int r_in = t >> 3;     // 0..31
    struct Seq { int tile, c; };
    auto valid = [&](const Seq& q) { return q.tile < g.n_tiles; };
    auto advance = [&](Seq& q) { if (++q.c == g.kc) { q.c = 0; q.tile += gridDim.x; } };
    const int k_live = g.k_live;
    auto issue = [&](const Seq& q, float4 (*ld)[2]) {
      const int64_t row0 = (int64_t)q.tile * kTcBM;
      const int k = q.c * kTcBK + kq * 8;
      if (g.pf_tiles && lane == 0 && (g.pf_spread || q.c == g.pf_at)) {
        // ask L2 for the rows of a later tile of this CTA: the register buffers only keep two stages of loads in
        // flight, and at HBM latency that is not enough bytes per SM; a prefetched row is an L2 hit when its turn comes.
        // pf_spread: each K step of the current tile asks for 1/kc of this warp's rows instead of all of them at once
        constexpr int kRows = kTcBM / kTcLoaderWarps;
        const int per = g.pf_spread ? kRows / g.kc : kRows;
        const int64_t pr0 = ((int64_t)q.tile + (int64_t)g.pf_tiles * gridDim.x) * kTcBM + warp * kRows + (g.pf_spread ? q.c * per : 0);
        const int64_t n = min_i64((int64_t)per, g.M - pr0);
        if (n > 0) {
          if (g.ldx == g.K) {
            l2_prefetch(g.X + pr0 * g.ldx, (uint32_t)(n * g.K * 4));
          } else {
            for (int i = 0; i < (int)n; ++i) l2_prefetch(g.X + (pr0 + i) * g.ldx, (uint32_t)(g.K * 4));
          }
        }
      }
      if (k >= k_live) {
        // columns beyond the last K16 step the MMA issues (K = 80 or 200 in a 64-wide chunk): never read, so neither
        // loaded nor stored
      } else if (g.debug & 2) {
#pragma unroll
        for (int pass = 0; pass < kPasses; ++pass) ld[pass][0] = ld[pass][1] = make_float4(1.f, 2.f, 3.f, 4.f);
      } else if (g.vec_x == 2 && (k + 7 < g.K)) {
        // one 256-bit load per row chunk: the eight lanes of a row read 256 contiguous bytes, every 32-byte sector
        // of the request is fully used (two 128-bit loads at a 32-byte lane stride touch each sector twice and cost
        // twice the L1 data-stage wavefronts -- the stage this kernel saturates)
#pragma unroll
        for (int pass = 0; pass < kPasses; ++pass) {
          const int64_t row = min_i64(row0 + pass * (kTcLoaderWarps * 4) + r_in, g.M - 1);
          ldg256_nc(g.X + row * g.ldx + k, ld[pass][0], ld[pass][1]);
        }
      } else if (g.vec_x && (k + 7 < g.K)) {
#pragma unroll
        for (int pass = 0; pass < kPasses; ++pass) {
          const int64_t row = min_i64(row0 + pass * (kTcLoaderWarps * 4) + r_in, g.M - 1);
          const float4* src = reinterpret_cast<const float4*>(g.X + row * g.ldx + k);
          ld[pass][0] = __ldg(src);
          ld[pass][1] = __ldg(src + 1);
        }
      } else {
#pragma unroll
        for (int pass = 0; pass < kPasses; ++pass) {
          const int64_t row = min_i64(row0 + pass * (kTcLoaderWarps * 4) + r_in, g.M - 1);
          const float* src = g.X + row * g.ldx + k;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = (k + j < g.K) ? __ldg(src + j) : 0.f;
          ld[pass][0] = make_float4(v[0], v[1], v[2], v[3]);
          ld[pass][1] = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
    };
    int stage = 0;
    uint32_t phase = 0;
    auto consume = [&](const Seq& q, float4 (*ld)[2]) {
      mbar_wait(empty_bar(stage), phase ^ 1u, g.backoff_ld);
      uint8_t* st_hi = a_smem + (size_t)stage * kTcStageBytes;
      uint8_t* st_lo = st_hi + kTcTileBytes;
      const int64_t row0 = (int64_t)q.tile * kTcBM;
      if (q.c * kTcBK + kq * 8 < k_live)
#pragma unroll
      for (int pass = 0; pass < kPasses; ++pass) {
        const int r = pass * (kTcLoaderWarps * 4) + r_in;
        float4 p0 = ld[pass][0], p1 = ld[pass][1];
        if (row0 + r >= g.M) p0 = p1 = make_float4(0.f, 0.f, 0.f, 0.f);
        uint4 hi, lo;
        split2(p0.x, p0.y, hi.x, lo.x);
        split2(p0.z, p0.w, hi.y, lo.y);
        split2(p1.x, p1.y, hi.z, lo.z);
        split2(p1.z, p1.w, hi.w, lo.w);
        const int off = r * 128 + ((kq ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(st_hi + off) = hi;
        *reinterpret_cast<uint4*>(st_lo + off) = lo;
      }
      fence_proxy_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(stage));
      if (++stage == g.n_stages) { stage = 0; phase ^= 1u; }
    };
    // kBufs register buffers rotate: the loads of kBufs - 1 stages are in flight while one is converted
    constexpr int kBufs = NB;
    float4 bufs[kBufs][kPasses][2];
    Seq nxt{(int)blockIdx.x, 0}, cur{(int)blockIdx.x, 0};
#pragma unroll
    for (int i = 0; i < kBufs - 1; ++i)
      if (valid(nxt)) { issue(nxt, bufs[i]); advance(nxt); }
    while (valid(cur)) {
#pragma unroll
      for (int i = 0; i < kBufs; ++i) {
        if (valid(nxt)) { issue(nxt, bufs[(i + kBufs - 1) % kBufs]); advance(nxt); }
        consume(cur, bufs[i]); advance(cur);
        if (!valid(cur)) break;
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int q = warp - kTcEpiWarp0;  // TMEM lane quarter this warp may touch (== warp % 4)
    int buf = 0;                        // YB == 2: staging tile in use
    int it = 0;
    for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++it) {
      const int b = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait(tfull_bar(b), use & 1u, g.backoff_epi);
      tc_fence_after();
      const int64_t row = (int64_t)tile * kTcBM + q * 32 + lane;
      const bool row_ok = row < g.M;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * g.n_pad);
      if (YB == 2) {
        // fp32 rows through the tensor map: the warp stages 32 rows x 32 columns (128-byte rows, 16-byte chunks XOR-swizzled
        // with the row -- conflict-free for row-per-thread stores and exactly the tensor map's SWIZZLE_128B), lane 0 issues
        // one 2-D bulk store per tile; two tiles alternate so that a store is in flight while the next is assembled.
        uint8_t* my = reinterpret_cast<uint8_t*>(stage_out) + (size_t)q * 8192;
        const uint32_t my_s = smem_u32(my);
        const int row0 = tile * kTcBM + q * 32;
        for (int c0 = 0; c0 < g.N; c0 += 32, buf ^= 1) {   // (buf runs on across tiles: the wait below allows ONE copy in flight)
          float v[32];
          tmem_ld16(taddr + (uint32_t)c0, v);
          if (c0 + 16 < g.n_pad) tmem_ld16(taddr + (uint32_t)(c0 + 16), v + 16);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = c0 + j < g.n_pad ? bias_act<ACT>(v[j], bias_s[c0 + j]) : 0.f;
          // the copy issued two rounds ago must have READ this tile
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
          uint8_t* dst = my + buf * 4096 + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(dst + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          fence_proxy_async();   // generic-proxy stores -> visible to the bulk copy (async proxy)
          __syncwarp();
          if (lane == 0 && row0 < g.M && !(g.debug & 1)) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&ymap), "r"(c0), "r"(row0),
                         "r"(my_s + (uint32_t)buf * 4096u)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      } else if (YB == 1) {
        // Per-stream operand stages through BULK stores.  Measured on the 128 -> 256 layer (8192 x 98 rows, tools/block_microbench.py,
        // VADX_TC_DEBUG): 0.29 ms with the per-lane 16-byte stores of the y_split branch below (each instruction: eight 64-byte
        // segments in eight different images), 0.175 ms with the stores removed, 0.286 ms with the LOADS removed -- the layer
        // was bound by its store instructions, not by bytes.  Here a warp lays its 32 rows of one K chunk (hi image rows |
        // lo image rows, 2 x 4 KB, the chunk swizzle of the IMAGE row) out in shared memory and lane 0 hands each run of
        // rows that is contiguous in a stream's image to cp.async.bulk (at most two runs: a stream has >= 32 rows).
        uint8_t* my = reinterpret_cast<uint8_t*>(stage_out) + (size_t)q * 8192;
        const uint32_t my_s = smem_u32(my);
        const int64_t grow0 = (int64_t)tile * kTcBM + q * 32;
        const int64_t s0 = grow0 / g.ys_T;
        const int t0 = (int)(grow0 - s0 * g.ys_T);
        const int n_valid = (int)min_i64(32, g.M - grow0);
        const int n1 = min(n_valid, g.ys_T - t0), n2 = n_valid - n1;     // rows in stream s0 / in stream s0 + 1
        const int t = lane < n1 ? t0 + lane : lane - n1;                 // this lane's row in its stream's images
        const size_t stream_bytes = (size_t)(g.N / kTcBK) * 2 * g.ys_img;
        uint8_t* const img0 = reinterpret_cast<uint8_t*>(g.Y) + (size_t)s0 * stream_bytes + (size_t)t0 * 128;
        uint8_t* const img1 = reinterpret_cast<uint8_t*>(g.Y) + (size_t)(s0 + 1) * stream_bytes;
        uint8_t* const row_hi = my + lane * 128;
        for (int c0 = 0; c0 < g.N && n_valid > 0; c0 += kTcBK) {
          // the previous chunk's copies must have READ the tile before it is overwritten
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            float v[32];
            tmem_ld32(taddr + (uint32_t)(c0 + 32 * hh), v);
            const float4* b4 = reinterpret_cast<const float4*>(bias_s + c0 + 32 * hh);
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = b4[j];
              split2(bias_act<ACT>(v[4 * j], b.x), bias_act<ACT>(v[4 * j + 1], b.y), hi[2 * j], lo[2 * j]);
              split2(bias_act<ACT>(v[4 * j + 2], b.z), bias_act<ACT>(v[4 * j + 3], b.w), hi[2 * j + 1], lo[2 * j + 1]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int off = ((4 * hh + j) ^ (t & 7)) << 4;
              *reinterpret_cast<uint4*>(row_hi + off) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
              *reinterpret_cast<uint4*>(row_hi + 4096 + off) = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
            }
          }
          fence_proxy_async();   // generic-proxy stores -> visible to the bulk copy (async proxy)
          __syncwarp();
          if (lane == 0 && !(g.debug & 1)) {
            const size_t ci = (size_t)(c0 / kTcBK) * 2 * g.ys_img;
            bulk_s2g(img0 + ci, my_s, (uint32_t)n1 * 128u);
            bulk_s2g(img0 + ci + g.ys_img, my_s + 4096u, (uint32_t)n1 * 128u);
            if (n2 > 0) {
              bulk_s2g(img1 + ci, my_s + (uint32_t)n1 * 128u, (uint32_t)n2 * 128u);
              bulk_s2g(img1 + ci + g.ys_img, my_s + 4096u + (uint32_t)n1 * 128u, (uint32_t)n2 * 128u);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      } else if (g.head_out) {
        // fused narrow head: the row's N activations never leave the SM
        float acc = 0.f;
        for (int c0 = 0; c0 < g.n_pad; c0 += 16) {
          float v[16];
          tmem_ld16(taddr + (uint32_t)c0, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc = fmaf(bias_act<ACT>(v[j], bias_s[c0 + j]), head_s[c0 + j], acc);
        }
        if (row_ok) g.head_out[row] = 1.0f / (1.0f + expf(-(acc + g.head_b)));
      } else if (g.y_split) {
        // output as the next layer's operand stages: bias/activation, the two-term bf16 split, and the swizzled 16 KB
        // images [tile][N/64][hi | lo] -- the consumer streams them with bulk copies and converts nothing.
        // Staged per warp as 32 rows x (16 hi words | 16 lo words); eight lanes write one row's four hi and four lo chunks.
        uint32_t* my = reinterpret_cast<uint32_t*>(stage_out + (q * 32) * kTcOutLd);
        uint8_t* yimg = reinterpret_cast<uint8_t*>(g.Y) + (g.y_split == 1 ? (size_t)tile * (g.N / kTcBK) * kTcStageBytes : 0);
        // per-stream images: byte offset of each of this lane's eight rows (stream base + row in the image), ~0 = past M
        uint32_t roff[8];
        if (g.y_split == 2) {
          // one division per tile and lane; the lane's rows are 4 apart and a stream has more than 4 rows, so the
          // (stream, row in stream) pair of the next row is an add and at most one wrap (eight 64-bit divisions here
          // made the N = 256 epilogue, already the busiest warps of the kernel, 40 % slower)
          const int64_t grow0 = (int64_t)tile * kTcBM + q * 32 + (lane >> 3);
          int64_t sidx = grow0 / g.ys_T;
          int t = (int)(grow0 - sidx * g.ys_T);
          const uint32_t stream_bytes = (uint32_t)((g.N / kTcBK) * 2 * g.ys_img);
          uint32_t sbase = (uint32_t)sidx * stream_bytes;
#pragma unroll
          for (int itr = 0; itr < 8; ++itr) {
            roff[itr] = grow0 + itr * 4 < g.M ? sbase + (uint32_t)t * 128u : 0xffffffffu;
            t += 4;
            if (t >= g.ys_T) { t -= g.ys_T; sbase += stream_bytes; }
          }
        }
        for (int c0 = 0; c0 < g.N; c0 += 32) {
          float v[32];
          tmem_ld32(taddr + (uint32_t)c0, v);
          const float4* b4 = reinterpret_cast<const float4*>(bias_s + c0);
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = b4[j];
            split2(bias_act<ACT>(v[4 * j], b.x), bias_act<ACT>(v[4 * j + 1], b.y), hi[2 * j], lo[2 * j]);
            split2(bias_act<ACT>(v[4 * j + 2], b.z), bias_act<ACT>(v[4 * j + 3], b.w), hi[2 * j + 1], lo[2 * j + 1]);
          }
          uint4* d4 = reinterpret_cast<uint4*>(my + lane * kTcOutLd);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            d4[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
            d4[4 + j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
          }
          __syncwarp();
          const int piece = lane & 7, half = piece >> 2, ch = piece & 3;
          const int jj = ((c0 % kTcBK) >> 3) + ch;   // 16-byte chunk of the 128-byte image row
          uint4 val[8];
#pragma unroll
          for (int itr = 0; itr < 8; ++itr)
            val[itr] = *reinterpret_cast<const uint4*>(my + (itr * 4 + (lane >> 3)) * kTcOutLd + half * 16 + ch * 4);
          if (g.debug & 1) {
          } else if (g.y_split == 1) {
            uint8_t* base = yimg + (size_t)(c0 / kTcBK) * kTcStageBytes + (size_t)half * kTcTileBytes;
#pragma unroll
            for (int itr = 0; itr < 8; ++itr) {
              const int r = q * 32 + itr * 4 + (lane >> 3);
              *reinterpret_cast<uint4*>(base + r * 128 + ((jj ^ (r & 7)) << 4)) = val[itr];
            }
          } else {
            uint8_t* base = yimg + (size_t)((c0 / kTcBK) * 2 + half) * g.ys_img;
#pragma unroll
            for (int itr = 0; itr < 8; ++itr)
              if (roff[itr] != 0xffffffffu)
                *reinterpret_cast<uint4*>(base + roff[itr] + ((jj ^ ((roff[itr] >> 7) & 7)) << 4)) = val[itr];
          }
          __syncwarp();
        }
      } else if (!g.res && g.vec_y && !(g.debug & 1) && !(g.debug & 8)) {
        // coalesced path: each thread stages 32 columns of its row in shared memory, then the warp
        // writes them out 4 rows x 128 B per instruction (row-per-thread stores would touch 32
        // different 128-byte lines per instruction and saturate the LSU tag stage)
        float* my = stage_out + (q * 32) * kTcOutLd;
        const int64_t tile_row0 = (int64_t)tile * kTcBM + q * 32;
        const bool rows_full = tile_row0 + 32 <= g.M && !g.debug;
        float4* const st_dst = reinterpret_cast<float4*>(my + lane * kTcOutLd);
        const float* const ld_src = my + (lane >> 3) * kTcOutLd + (lane & 7) * 4;
        float* const out0 = g.Y + (tile_row0 + (lane >> 3)) * g.ldy + (lane & 7) * 4;
        const int64_t ld4 = 4 * g.ldy;
        for (int c0 = 0; c0 < g.n_pad; c0 += 32) {
          if (rows_full && c0 + 32 <= g.N) {
            // interior round (all 32 rows and 32 columns valid): no per-element guards, bias as float4,
            // the eight staged rows-of-four are read back before the first store is issued
            float v[32];
            tmem_ld32(taddr + (uint32_t)c0, v);
            const float4* b4 = reinterpret_cast<const float4*>(bias_s + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = b4[j];
              st_dst[j] = make_float4(bias_act<ACT>(v[4 * j], b.x), bias_act<ACT>(v[4 * j + 1], b.y),
                                      bias_act<ACT>(v[4 * j + 2], b.z), bias_act<ACT>(v[4 * j + 3], b.w));
            }
            __syncwarp();
            float4 val[8];
#pragma unroll
            for (int itr = 0; itr < 8; ++itr) val[itr] = *reinterpret_cast<const float4*>(ld_src + itr * 4 * kTcOutLd);
            float* o = out0 + c0;
#pragma unroll
            for (int itr = 0; itr < 8; ++itr) *reinterpret_cast<float4*>(o + itr * ld4) = val[itr];
            __syncwarp();
            continue;
          }
          float v[32];
          tmem_ld16(taddr + (uint32_t)c0, v);
          if (c0 + 16 < g.n_pad) tmem_ld16(taddr + (uint32_t)(c0 + 16), v + 16);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = c0 + j;
            v[j] = col < g.n_pad ? bias_act<ACT>(v[j], bias_s[col < g.n_pad ? col : 0]) : 0.f;
          }
          float4* dst = reinterpret_cast<float4*>(my + lane * kTcOutLd);
          if (!(g.debug & 16))
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          __syncwarp();
          if (!(g.debug & 32))
#pragma unroll
          for (int itr = 0; itr < 8; ++itr) {
            const int r = itr * 4 + (lane >> 3);
            const int cc = c0 + (lane & 7) * 4;
            const int64_t grow = tile_row0 + r;
            if (grow < g.M && cc < g.N && !(g.debug & 4)) {
              const float4 val = *reinterpret_cast<const float4*>(my + r * kTcOutLd + (lane & 7) * 4);
              float* out = g.Y + grow * g.ldy + cc;
              if (cc + 3 < g.N) {
                *reinterpret_cast<float4*>(out) = val;
              } else {
                out[0] = val.x;
                if (cc + 1 < g.N) out[1] = val.y;
                if (cc + 2 < g.N) out[2] = val.z;
              }
            }
          }
          __syncwarp();
        }
      } else {
      for (int c0 = 0; c0 < g.n_pad; c0 += 16) {
          float v[16];
          tmem_ld16(taddr + (uint32_t)c0, v);
          if (!row_ok || c0 >= g.N || (g.debug & 1)) continue;
          const bool res_first = (g.act & VADX_ACT_RES_FIRST) != 0;
          if (ACT == VADX_ACT_LOG || ACT == VADX_ACT_LOG_CLAMP) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = bias_act<ACT>(v[j], bias_s[c0 + j]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += bias_s[c0 + j];
          }
          if (!res_first && ACT != VADX_ACT_LOG && ACT != VADX_ACT_LOG_CLAMP) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = apply_act(v[j], ACT);
          }
          float* out = g.Y + row * g.ldy + c0;
          const float* rs = g.res ? g.res + row * g.ldr + c0 : nullptr;
          const bool vec = g.vec_y && c0 + 15 < g.N;
          if (rs) {
            if (vec) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float4 r4 = __ldg(reinterpret_cast<const float4*>(rs) + j);
                v[4 * j] += r4.x; v[4 * j + 1] += r4.y; v[4 * j + 2] += r4.z; v[4 * j + 3] += r4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < g.N) v[j] += rs[j];
            }
          }
          if (res_first) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = apply_act(v[j], ACT);
          }
          if (vec) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              reinterpret_cast<float4*>(out)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < g.N) out[j] = v[j];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(b));
    }
  }
  if (YB && warp >= kTcEpiWarp0 && warp < kTcMmaWarp && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (warp == kTcMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host
static inline uint16_t bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)(u >> 16);  // inf / nan
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float bf16_to_f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

struct TcShape {
  int n_pad, kc, n_k16, n_stages, tmem_cols;
  size_t w_bytes, smem_bytes;
  bool ok;
};
TcShape tc_shape(int n_in, int n_out, bool bulk_out = false);
TcShape tc_shape(int n_in, int n_out, bool bulk_out) {
  TcShape s{};
  s.n_pad = (int)round_up(n_out, 16);
  s.kc = (int)ceil_div(n_in, kTcBK);
  s.n_k16 = (int)ceil_div(n_in, 16);
  s.w_bytes = (size_t)s.kc * 2 * s.n_pad * 128;
  // bias + head vectors, the four epilogue warps' staging tiles (padded 32 x 32 transposes, or 2 x 4 KB image rows), barriers
  const size_t misc = (size_t)s.n_pad * 8 + (bulk_out ? (size_t)4 * 8192 : (size_t)4 * 32 * kTcOutLd * 4) + 13 * 8 + 16;
  s.ok = s.n_pad <= 256 && n_out > 8;
  if (s.ok) {
    size_t left = kTcSmemBudget > s.w_bytes + misc ? kTcSmemBudget - s.w_bytes - misc : 0;
    s.n_stages = (int)std::min<size_t>(4, left / kTcStageBytes);
    s.ok = s.n_stages >= 2;
    s.smem_bytes = s.w_bytes + (size_t)s.n_stages * kTcStageBytes + misc;
    int cols = 32;
    while (cols < 2 * s.n_pad) cols *= 2;
    s.tmem_cols = cols;
  }
  return s;
}

}  // namespace vadx

using namespace vadx;

// fc1 with its output as per-stream operand stages (rows_per_stream rows per image set)
int linear_tc_stream_stages_f32(const float* d_x, const void* d_wimg, const float* d_bias, float* d_y, int64_t n_rows,
                                int rows_per_stream, int n_in, int n_out, int act, void* stream);

extern "C" int vadx_tc_supported(int n_in, int n_out) { return tc_shape(n_in, n_out).ok ? 1 : 0; }

extern "C" int vadx_pack_weight_tc(const float* h_w, int n_out, int n_in, void* h_img, size_t img_capacity,
                                   size_t* img_bytes) {
  VADX_REQUIRE(h_w && img_bytes && n_out > 0 && n_in > 0, "vadx_pack_weight_tc: bad argument");
  TcShape s = tc_shape(n_in, n_out);
  VADX_REQUIRE(s.ok, "vadx_pack_weight_tc: shape %d -> %d is not supported by the tensor-core path", n_in, n_out);
  *img_bytes = s.w_bytes;
  if (!h_img) return VADX_OK;  // size query
  VADX_REQUIRE(img_capacity >= s.w_bytes, "vadx_pack_weight_tc: image buffer too small");
  uint8_t* img = static_cast<uint8_t*>(h_img);
  memset(img, 0, s.w_bytes);
  const size_t tile = (size_t)s.n_pad * 128;
  for (int n = 0; n < n_out; ++n)
    for (int k = 0; k < n_in; ++k) {
      float w = h_w[(size_t)n * n_in + k];
      uint16_t hi = bf16_rn(w);
      uint16_t lo = bf16_rn(w - bf16_to_f(hi));
      int c = k / kTcBK, kk = k % kTcBK;
      size_t off = (size_t)(n / 8) * 1024 + (size_t)(n % 8) * 128 + (size_t)(((kk / 8) ^ (n % 8)) * 16) + (kk % 8) * 2;
      memcpy(img + (size_t)(c * 2) * tile + off, &hi, 2);
      memcpy(img + (size_t)(c * 2 + 1) * tile + off, &lo, 2);
    }
  return VADX_OK;
}

static int linear_tc_launch(const float* d_x, int64_t ldx, const void* d_wimg, const float* d_bias,
                            const float* d_residual, int64_t ldr, float* d_y, int64_t ldy, int64_t n_rows, int n_in,
                            int n_out, int act, const float* d_head_w, float head_b, float* d_head_out, void* stream,
                            int x_split = 0, int y_split = 0, int rows_per_stream = 0);

// Internal (model.cu): the same layer with its input and/or output in the operand-stage format (TcArgs::x_split / y_split).

int linear_tc_stages_f32(const float* d_x, const void* d_wimg, const float* d_bias, float* d_y, int64_t n_rows, int n_in,
                         int n_out, int act, int x_split, int y_split, void* stream) {
  VADX_REQUIRE(d_y, "linear_tc_stages_f32: null pointer");
  VADX_REQUIRE(!x_split || n_in % kTcBK == 0, "linear_tc_stages_f32: a staged input needs K %% 64 == 0 (got %d)", n_in);
  VADX_REQUIRE(!y_split || n_out % kTcBK == 0, "linear_tc_stages_f32: a staged output needs N %% 64 == 0 (got %d)", n_out);
  return linear_tc_launch(d_x, n_in, d_wimg, d_bias, nullptr, 0, d_y, n_out, n_rows, n_in, n_out, act, nullptr, 0.f, nullptr,
                          stream, x_split, y_split);
}

extern "C" int vadx_linear_tc_f32(const float* d_x, int64_t ldx, const void* d_wimg, const float* d_bias,
                                  const float* d_residual, int64_t ldr, float* d_y, int64_t ldy, int64_t n_rows,
                                  int n_in, int n_out, int act, void* stream) {
  VADX_REQUIRE(d_y, "vadx_linear_tc_f32: null pointer");
  return linear_tc_launch(d_x, ldx, d_wimg, d_bias, d_residual, ldr, d_y, ldy, n_rows, n_in, n_out, act, nullptr, 0.f,
                          nullptr, stream);
}

extern "C" int vadx_linear_head_tc_f32(const float* d_x, int64_t ldx, const void* d_wimg, const float* d_bias,
                                       int64_t n_rows, int n_in, int n_out, int act, const float* d_head_w,
                                       float head_bias, float* d_head_out, void* stream) {
  VADX_REQUIRE(d_head_w && d_head_out, "vadx_linear_head_tc_f32: null pointer");
  return linear_tc_launch(d_x, ldx, d_wimg, d_bias, nullptr, 0, nullptr, n_out, n_rows, n_in, n_out, act, d_head_w,
                          head_bias, d_head_out, stream);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (null when the driver does not export it)
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder() {
  static const TensorMapEncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    (void)cudaGetLastError();
    return reinterpret_cast<TensorMapEncodeFn>(p);
  }();
  return fn;
}
// fp32 [rows][n_out] with a row pitch of ldy floats as a 2-D tensor: 32 x 32 boxes, 128-byte swizzled in shared memory
static bool encode_rows_map(CUtensorMap* m, float* d_y, int64_t ldy, int64_t n_rows, int n_out) {
  const TensorMapEncodeFn enc = tensor_map_encoder();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)n_out, (cuuint64_t)n_rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ldy * 4u};
  const cuuint32_t box[2] = {32u, 32u}, estr[2] = {1u, 1u};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int linear_tc_launch(const float* d_x, int64_t ldx, const void* d_wimg, const float* d_bias,
                            const float* d_residual, int64_t ldr, float* d_y, int64_t ldy, int64_t n_rows, int n_in,
                            int n_out, int act, const float* d_head_w, float head_b, float* d_head_out, void* stream,
                            int x_split, int y_split, int rows_per_stream) {
  const int g_ys_T = y_split == 2 ? rows_per_stream : 0;
  // rows in + rows out at 4 B per value (the operand-stage hand-overs carry two bf16 terms = the same 4 B), + residual rows
  StageTimer _timer(VADX_STAGE_LINEAR, (cudaStream_t)stream, d_head_out ? "linear_tc_kernel+head" : "linear_tc_kernel",
                    4.0 * n_rows * (n_in + (d_head_out ? 1 : n_out) + (d_residual ? n_out : 0)), 2.0 * n_rows * n_in * n_out);
  VADX_REQUIRE(d_x && d_wimg, "vadx_linear_tc_f32: null pointer");
  // ldx < n_in is allowed: the input rows then OVERLAP (hop-strided frames of a signal: the framed DFT as a dense layer);
  // the input is only read, and the caller guarantees that the last row's n_in values are in bounds
  VADX_REQUIRE(n_rows >= 0 && n_in > 0 && n_out > 0 && ldx >= 1 && ldy >= n_out, "vadx_linear_tc_f32: bad shape");
  TcShape s = tc_shape(n_in, n_out);
  VADX_REQUIRE(s.ok, "vadx_linear_tc_f32: shape %d -> %d is not supported by the tensor-core path", n_in, n_out);
  VADX_REQUIRE(aligned16(d_wimg), "vadx_linear_tc_f32: weight image must be 16-byte aligned");
  if (n_rows == 0) return VADX_OK;
  // per-stream stages leave through bulk stores when the wider staging tiles cost no operand stage
  bool y_bulk = false;
  if (y_split == 2 && rows_per_stream >= 32 && !x_split && !d_residual && !d_head_out && lin_loader_warps() == 16 &&
      ((act & 15) == VADX_ACT_NONE || (act & 15) == VADX_ACT_RELU)) {
    static const int want = ab_env("VADX_LIN_YBULK", 1);
    const TcShape sb = tc_shape(n_in, n_out, true);
    if (want && sb.ok && sb.n_stages == s.n_stages) { s = sb; y_bulk = true; }
  }
  // plain fp32 rows leave through 2-D tensor-map stores under the same condition
  alignas(64) CUtensorMap ymap;
  memset(&ymap, 0, sizeof(ymap));
  bool y_tma = false;
  if (!y_split && d_y && !d_residual && !d_head_out && lin_loader_warps() == 16 && (ldy & 3) == 0 && aligned16(d_y) &&
      n_rows < (1LL << 31) && n_rows > kSkinnyMaxRows) {
    static const int want = ab_env("VADX_LIN_YTMA", 1);
    const TcShape sb = tc_shape(n_in, n_out, true);
    // (the wider staging tiles may cost the fourth operand stage, never the second or third)
    if (want && sb.ok && sb.n_stages >= std::min(s.n_stages, 3) && encode_rows_map(&ymap, d_y, ldy, n_rows, n_out)) { s = sb; y_tma = true; }
  }
  static PerDevice per_device;
  int n_sm = 148;
  VADX_TRY(per_device.ensure(&n_sm, [] {
    cudaError_t e = cudaSuccess;
    auto opt_in = [&](auto kern) {
      if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBudget);
    };
    opt_in(linear_tc_kernel<VADX_ACT_NONE, 16, 3, 1>); opt_in(linear_tc_kernel<VADX_ACT_RELU, 16, 3, 1>);
    opt_in(linear_tc_kernel<VADX_ACT_NONE, 16, 3, 2>); opt_in(linear_tc_kernel<VADX_ACT_RELU, 16, 3, 2>);
    opt_in(linear_tc_kernel<VADX_ACT_SIGMOID, 16, 3, 2>); opt_in(linear_tc_kernel<VADX_ACT_LOG, 16, 3, 2>);
    opt_in(linear_tc_kernel<VADX_ACT_LOG_CLAMP, 16, 3, 2>);
    opt_in(linear_tc_kernel<VADX_ACT_NONE, 8>);       opt_in(linear_tc_kernel<VADX_ACT_NONE, 16>); opt_in(linear_tc_kernel<VADX_ACT_NONE, 16, 4>);
    opt_in(linear_tc_kernel<VADX_ACT_RELU, 8>);       opt_in(linear_tc_kernel<VADX_ACT_RELU, 16>); opt_in(linear_tc_kernel<VADX_ACT_RELU, 16, 4>);
    opt_in(linear_tc_kernel<VADX_ACT_SIGMOID, 8>);    opt_in(linear_tc_kernel<VADX_ACT_SIGMOID, 16>); opt_in(linear_tc_kernel<VADX_ACT_SIGMOID, 16, 4>);
    opt_in(linear_tc_kernel<VADX_ACT_LOG, 8>);        opt_in(linear_tc_kernel<VADX_ACT_LOG, 16>); opt_in(linear_tc_kernel<VADX_ACT_LOG, 16, 4>);
    opt_in(linear_tc_kernel<VADX_ACT_LOG_CLAMP, 8>);  opt_in(linear_tc_kernel<VADX_ACT_LOG_CLAMP, 16>); opt_in(linear_tc_kernel<VADX_ACT_LOG_CLAMP, 16, 4>);
    return e;
  }));
  TcArgs g{};
  g.X = d_x; g.ldx = ldx; g.Wimg = static_cast<const uint8_t*>(d_wimg); g.bias = d_bias; g.res = d_residual;
  g.ldr = ldr; g.Y = d_y; g.ldy = ldy; g.M = n_rows; g.K = n_in; g.N = n_out; g.n_pad = s.n_pad; g.kc = s.kc;
  g.n_k16 = s.n_k16; g.act = act; g.n_stages = s.n_stages; g.tmem_cols = s.tmem_cols;
  g.head_w = d_head_w; g.head_b = head_b; g.head_out = d_head_out;
  int64_t tiles = ceil_div(n_rows, kTcBM);
  VADX_REQUIRE(tiles <= 0x7fffffffLL, "vadx_linear_tc_f32: too many rows");
  g.n_tiles = (int)tiles;
  {
    static const int dbg = ab_env("VADX_TC_DEBUG", 0);
    g.debug = dbg;
  }
  g.vec_x = ((ldx & 3) == 0) && aligned16(d_x);
  {
    static const int bl = ab_env("VADX_TC_BACKOFF_LD", 64);
    static const int be = ab_env("VADX_TC_BACKOFF_EPI", 64);
    g.backoff_ld = (unsigned)bl; g.backoff_epi = (unsigned)be;
    static const bool all_cols = ab_env("VADX_TC_ALLCOLS", 0) != 0;
    g.k_live = all_cols ? s.kc * kTcBK : s.n_k16 * 16;
    static const int pf = ab_env("VADX_LIN_PF", 1);
    g.pf_tiles = (g.vec_x && (n_in & 3) == 0) ? pf : 0;   // bulk prefetch wants 16-byte aligned addresses and sizes
    static const int spread = ab_env("VADX_LIN_PF_SPREAD", 0);
    static const int at = ab_env("VADX_LIN_PF_AT", 0);
    g.pf_at = at == 1 ? s.kc / 2 : (at == 2 ? s.kc - 1 : 0);
    g.x_split = x_split; g.y_split = y_split; g.y_bulk = y_bulk ? 1 : 0;
    g.ys_T = g_ys_T; g.ys_img = (int)round_up(g_ys_T, 16) * 128;
    if (y_split == 2) VADX_REQUIRE(ceil_div(n_rows, (int64_t)std::max(1, g_ys_T)) * (int64_t)(n_out / kTcBK) * 2 * (round_up(g_ys_T, 16) * 128) < (1LL << 32),
                                   "linear_tc: per-stream stages of %lld rows exceed the 32-bit image offsets", (long long)n_rows);
    if (y_split == 2) VADX_REQUIRE(g_ys_T > 4 && (round_up(g_ys_T, 16) * 128) % 1024 == 0, "linear_tc: per-stream stages need round_up(T, 16) %% 8 == 0");
    if (x_split) { g.pf_tiles = 0; VADX_REQUIRE(aligned16(d_x), "linear_tc: staged input must be 16-byte aligned"); }
    if (y_split) VADX_REQUIRE(aligned16(d_y), "linear_tc: staged output must be 16-byte aligned");
    g.pf_spread = (spread && s.kc <= 8 && ((kTcBM / 16) % s.kc) == 0 && lw_is16()) ? 1 : 0;
  }
  {
    static const bool no256 = ab_env("VADX_LIN_NO_LDG256", 0) != 0;
    if (g.vec_x && !no256 && (ldx & 7) == 0 && (reinterpret_cast<uintptr_t>(d_x) & 31u) == 0) g.vec_x = 2;
  }
  g.vec_y = ((ldy & 3) == 0) && aligned16(d_y) && (!d_residual || (((ldr & 3) == 0) && aligned16(d_residual)));
  int grid = (int)std::min<int64_t>(tiles, n_sm);
  const int lw = lin_loader_warps();
#define VADX_LIN_LAUNCH(A)                                                                                      \
  do {                                                                                                          \
    if (y_tma) linear_tc_kernel<A, 16, 3, 2><<<grid, 21 * 32, s.smem_bytes, (cudaStream_t)stream>>>(g, ymap);   \
    else if (lw == 8) linear_tc_kernel<A, 8><<<grid, 13 * 32, s.smem_bytes, (cudaStream_t)stream>>>(g, ymap);   \
    else if (lw == 16) linear_tc_kernel<A, 16><<<grid, 21 * 32, s.smem_bytes, (cudaStream_t)stream>>>(g, ymap); \
    else linear_tc_kernel<A, 16, 4><<<grid, 21 * 32, s.smem_bytes, (cudaStream_t)stream>>>(g, ymap);            \
  } while (0)
  if (y_bulk) {
    if ((act & 15) == VADX_ACT_RELU) linear_tc_kernel<VADX_ACT_RELU, 16, 3, 1><<<grid, 21 * 32, s.smem_bytes, (cudaStream_t)stream>>>(g, ymap);
    else linear_tc_kernel<VADX_ACT_NONE, 16, 3, 1><<<grid, 21 * 32, s.smem_bytes, (cudaStream_t)stream>>>(g, ymap);
    return after_launch("vadx_linear_tc_f32");
  }
  switch (act & 15) {
    case VADX_ACT_NONE: VADX_LIN_LAUNCH(VADX_ACT_NONE); break;
    case VADX_ACT_RELU: VADX_LIN_LAUNCH(VADX_ACT_RELU); break;
    case VADX_ACT_SIGMOID: VADX_LIN_LAUNCH(VADX_ACT_SIGMOID); break;
    case VADX_ACT_LOG:
    case VADX_ACT_LOG_CLAMP:
      VADX_REQUIRE(!d_residual && !d_head_out && d_bias, "vadx_linear_tc_f32: the log epilogues need a bias/floor vector and take no residual or head");
      if ((act & 15) == VADX_ACT_LOG) VADX_LIN_LAUNCH(VADX_ACT_LOG);
      else VADX_LIN_LAUNCH(VADX_ACT_LOG_CLAMP);
      break;
    default: set_error("vadx_linear_tc_f32: activation %d is not supported", act); return VADX_EINVAL;
  }
#undef VADX_LIN_LAUNCH
  return after_launch("vadx_linear_tc_f32");
}

int linear_tc_stream_stages_f32(const float* d_x, const void* d_wimg, const float* d_bias, float* d_y, int64_t n_rows,
                                int rows_per_stream, int n_in, int n_out, int act, void* stream) {
  VADX_REQUIRE(d_y && rows_per_stream > 0, "linear_tc_stream_stages_f32: bad argument");
  VADX_REQUIRE(n_out % kTcBK == 0, "linear_tc_stream_stages_f32: a staged output needs N %% 64 == 0 (got %d)", n_out);
  return linear_tc_launch(d_x, n_in, d_wimg, d_bias, nullptr, 0, d_y, n_out, n_rows, n_in, n_out, act, nullptr, 0.f, nullptr,
                          stream, 0, 2, rows_per_stream);
}

extern "C" int vadx_linear_tc_stream_stages_f32(const float* d_x, const void* d_wimg, const float* d_bias, void* d_himg,
                                                int64_t n_rows, int rows_per_stream, int n_in, int n_out, int act, void* stream) {
  VADX_REQUIRE(rows_per_stream > 0 && n_rows % rows_per_stream == 0,
               "vadx_linear_tc_stream_stages_f32: %lld rows are not whole streams of %d", (long long)n_rows, rows_per_stream);
  VADX_REQUIRE((n_rows / rows_per_stream) * (int64_t)((n_out / kTcBK) * 2 * ((rows_per_stream + 15) / 16 * 16) * 128) < (1LL << 32),
               "vadx_linear_tc_stream_stages_f32: the stage images of one launch must stay below 4 GiB");
  return linear_tc_stream_stages_f32(d_x, d_wimg, d_bias, static_cast<float*>(d_himg), n_rows, rows_per_stream, n_in, n_out, act,
                                     stream);
}
