// gemm_simt.cu -- exact-fp32 SIMT contractions: dense layers (a8) and the framed DFT + power (a1).
//
// One register-blocked, shared-memory double-buffered GEMM template with two A-loaders
// (row-major activations / implicit framing of a padded signal) and two epilogues
// (bias+activation+residual / re^2+im^2).  This is the full-precision path: every product and
// sum is an fp32 FFMA, so it doubles as the on-device reference the tensor-core kernels are
// checked against.
#include "common.cuh"

namespace vadx {

constexpr int kBM = 128;  // rows (frames) per CTA
constexpr int kBK = 16;   // K slab
constexpr int kThreads = 256;
constexpr int kTM = 8;    // rows per thread

struct GemmArgs {
  const float* A;
  int64_t lda;        // row stride (linear) / stream stride of the padded signal (frames)
  int n_frames;       // frames loader: frames per stream
  int hop;            // frames loader: hop
  const float* B;     // [K][ldb]
  int ldb;
  int64_t M;
  int N;
  int K;
  const float* bias;
  const float* res;
  int64_t ldr;
  float* C;
  int64_t ldc;
  int act;
  int vec_store;      // C/res rows are 16B aligned -> float4 epilogue
};

template <int ALOAD>
__device__ __forceinline__ const float* a_row_ptr(const GemmArgs& g, int64_t row) {
  if (ALOAD == 0) return g.A + row * g.lda;
  int64_t s = row / g.n_frames;
  int t = (int)(row - s * g.n_frames);
  return g.A + s * g.lda + (int64_t)t * g.hop;
}

// BN in {64,128}; EPI 0 = bias/act/residual, 1 = power (re/im interleaved columns)
template <int BN, int ALOAD, int EPI>
__global__ void __launch_bounds__(kThreads, 2) gemm_f32_kernel(const GemmArgs g) {
  constexpr int TN = BN / 16;  // 4 or 8 columns per thread
  constexpr int LDA_S = kBM + 4;
  __shared__ __align__(16) float As[2][kBK][LDA_S];
  __shared__ __align__(16) float Bs[2][kBK][BN];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * kBM;
  const int n0 = blockIdx.y * BN;

  // A tile loader: 128 rows x 16 k = 512 float4, two per thread
  const int a_row = tid >> 2;          // 0..63 (+64)
  const int a_k = (tid & 3) * 4;       // 0,4,8,12
  const float* a_ptr[2];
  bool a_ok[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    int64_t row = m0 + a_row + h * 64;
    a_ok[h] = row < g.M;
    a_ptr[h] = a_row_ptr<ALOAD>(g, a_ok[h] ? row : 0);
  }
  const bool a_vec = ((g.lda & 3) == 0) && ((g.hop & 3) == 0 || ALOAD == 0) && aligned16_dev(g.A);
  // B tile loader: 16 k x BN = 4*BN float4
  constexpr int B_F4 = kBK * BN / 4;            // 512 or 256
  constexpr int B_PER_THREAD = B_F4 / kThreads; // 2 or 1
  constexpr int B_COLS4 = BN / 4;               // float4 per k-row

  float4 ra[2];
  float4 rb[B_PER_THREAD];

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int k = k0 + a_k;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a_ok[h]) {
        const float* p = a_ptr[h] + k;
        if (a_vec && k + 3 < g.K) {
          v = *reinterpret_cast<const float4*>(p);
        } else {
          if (k + 0 < g.K) v.x = p[0];
          if (k + 1 < g.K) v.y = p[1];
          if (k + 2 < g.K) v.z = p[2];
          if (k + 3 < g.K) v.w = p[3];
        }
      }
      ra[h] = v;
    }
#pragma unroll
    for (int j = 0; j < B_PER_THREAD; ++j) {
      int idx = tid + j * kThreads;
      int k = k0 + idx / B_COLS4;
      int n = n0 + (idx % B_COLS4) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < g.K && n < g.ldb) v = *reinterpret_cast<const float4*>(g.B + (int64_t)k * g.ldb + n);
      rb[j] = v;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int r = a_row + h * 64;
      As[buf][a_k + 0][r] = ra[h].x;
      As[buf][a_k + 1][r] = ra[h].y;
      As[buf][a_k + 2][r] = ra[h].z;
      As[buf][a_k + 3][r] = ra[h].w;
    }
#pragma unroll
    for (int j = 0; j < B_PER_THREAD; ++j) {
      int idx = tid + j * kThreads;
      *reinterpret_cast<float4*>(&Bs[buf][idx / B_COLS4][(idx % B_COLS4) * 4]) = rb[j];
    }
  };

  float acc[kTM][TN];
#pragma unroll
  for (int i = 0; i < kTM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int n_slabs = (g.K + kBK - 1) / kBK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int s = 0; s < n_slabs; ++s) {
    const int buf = s & 1;
    if (s + 1 < n_slabs) load_tiles((s + 1) * kBK);
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      float a[kTM], b[TN];
      *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      if (TN == 8) *reinterpret_cast<float4*>(&b[4]) = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
#pragma unroll
      for (int i = 0; i < kTM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (s + 1 < n_slabs) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < kTM; ++i) {
    const int64_t row = m0 + ty * 8 + i;
    if (row >= g.M) continue;
#pragma unroll
    for (int half = 0; half < TN / 4; ++half) {
      const int col = n0 + half * 64 + tx * 4;
      float v[4] = {acc[i][half * 4 + 0], acc[i][half * 4 + 1], acc[i][half * 4 + 2], acc[i][half * 4 + 3]};
      if (EPI == 1) {
        // columns (2f, 2f+1) are (re, im) of bin f
        const int f = col >> 1;
        float p0 = v[0] * v[0] + v[1] * v[1];
        float p1 = v[2] * v[2] + v[3] * v[3];
        float* out = g.C + row * g.ldc + f;
        if (f + 1 < g.N) {
          if (g.vec_store) {
            *reinterpret_cast<float2*>(out) = make_float2(p0, p1);
          } else {
            out[0] = p0;
            out[1] = p1;
          }
        } else if (f < g.N) {
          out[0] = p0;
        }
      } else {
        if (col >= g.N) continue;
        const bool full = col + 3 < g.N;
        if (g.bias) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (col + j < g.N) v[j] += g.bias[col + j];
        }
        const bool res_first = (g.act & VADX_ACT_RES_FIRST) != 0;
        if (!res_first) {
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = apply_act(v[j], g.act);
        }
        if (g.res) {
          const float* r = g.res + row * g.ldr + col;
          if (full && g.vec_store) {
            float4 rv = *reinterpret_cast<const float4*>(r);
            v[0] += rv.x; v[1] += rv.y; v[2] += rv.z; v[3] += rv.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (col + j < g.N) v[j] += r[j];
          }
        }
        if (res_first) {
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = apply_act(v[j], g.act);
        }
        float* out = g.C + row * g.ldc + col;
        if (full && g.vec_store) {
          *reinterpret_cast<float4*>(out) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (col + j < g.N) out[j] = v[j];
        }
      }
    }
  }
}

// Narrow heads (n_out <= 8): one warp per row, lanes split K, shuffle reduction.
// out index = (row / rows_per_group) * group_stride + o * out_stride + (row % rows_per_group)
// so that a [S*T][odim] head can be written directly as the reference's [S][odim][T].
__global__ void __launch_bounds__(256) linear_narrow_kernel(const float* __restrict__ X, int64_t ldx,
                                                            const float* __restrict__ Wt, int ldw,
                                                            const float* __restrict__ bias, float* __restrict__ Y,
                                                            int64_t M, int K, int N, int act, int rows_per_group,
                                                            int64_t group_stride, int64_t out_stride) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp; row < M; row += n_warps) {
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = 0.f;
    const float* x = X + row * ldx;
    for (int k = lane; k < K; k += 32) {
      float xv = x[k];
      const float* w = Wt + (int64_t)k * ldw;
#pragma unroll
      for (int o = 0; o < 8; ++o)
        if (o < N) acc[o] = fmaf(xv, w[o], acc[o]);
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      if (o < N) {
        float v = acc[o];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        acc[o] = v;
      }
    }
    if (lane == 0) {
      int64_t grp = row / rows_per_group;
      int64_t within = row - grp * rows_per_group;
      if (act == VADX_ACT_SOFTMAX) {
        float mx = -INFINITY, sum = 0.f;
        for (int o = 0; o < N; ++o) {
          acc[o] += bias ? bias[o] : 0.f;
          mx = fmaxf(mx, acc[o]);
        }
        for (int o = 0; o < N; ++o) {
          acc[o] = expf(acc[o] - mx);
          sum += acc[o];
        }
        for (int o = 0; o < N; ++o) Y[grp * group_stride + o * out_stride + within] = acc[o] / sum;
      } else {
        for (int o = 0; o < N; ++o) {
          float v = acc[o] + (bias ? bias[o] : 0.f);
          Y[grp * group_stride + o * out_stride + within] = apply_act(v, act);
        }
      }
    }
  }
}

// Skinny problems (M <= 16 rows: one stream, one short window): the 128-row tile above would run the
// whole K loop in ONE CTA (~25 us of serial latency).  Here the rows sit in shared memory, a CTA owns 32
// output columns (lane = column, so every weight row is one coalesced 128 B read), its 16 warps split K,
// and the partial sums meet in shared memory.  Still exact fp32 FFMA.
constexpr int kSkinnyRows = kSkinnyMaxRows;
constexpr int kSkinnyWarps = 16;
constexpr int kSkinnySmem = 48 * 1024;

template <int ROWS, int ALOAD, int EPI>
__global__ void __launch_bounds__(kSkinnyWarps * 32) gemm_skinny_kernel(const GemmArgs g) {
  extern __shared__ __align__(16) float sm[];   // phase 1: xs[K][ROWS]; phase 2: red[warp][ROWS][32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = (int)g.M, K = g.K;
  for (int idx = tid; idx < ROWS * K; idx += kSkinnyWarps * 32) {
    const int r = idx / K, k = idx - r * K;
    sm[k * ROWS + r] = r < M ? a_row_ptr<ALOAD>(g, r)[k] : 0.f;
  }
  __syncthreads();
  const int col = blockIdx.x * 32 + lane;
  const bool col_ok = col < g.ldb;
  const float* b = g.B + (col_ok ? col : 0);
  float acc[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) acc[r] = 0.f;
  constexpr int U = 4;
  int k = warp;
  for (; k + (U - 1) * kSkinnyWarps < K; k += U * kSkinnyWarps) {
    float w[U];
#pragma unroll
    for (int u = 0; u < U; ++u) w[u] = __ldg(b + (int64_t)(k + u * kSkinnyWarps) * g.ldb);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float4* x4 = reinterpret_cast<const float4*>(sm + (k + u * kSkinnyWarps) * ROWS);
#pragma unroll
      for (int q = 0; q < ROWS / 4; ++q) {
        const float4 x = x4[q];
        acc[4 * q + 0] = fmaf(x.x, w[u], acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(x.y, w[u], acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(x.z, w[u], acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(x.w, w[u], acc[4 * q + 3]);
      }
    }
  }
  for (; k < K; k += kSkinnyWarps) {
    const float w = __ldg(b + (int64_t)k * g.ldb);
    const float4* x4 = reinterpret_cast<const float4*>(sm + k * ROWS);
#pragma unroll
    for (int q = 0; q < ROWS / 4; ++q) {
      const float4 x = x4[q];
      acc[4 * q + 0] = fmaf(x.x, w, acc[4 * q + 0]);
      acc[4 * q + 1] = fmaf(x.y, w, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(x.z, w, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(x.w, w, acc[4 * q + 3]);
    }
  }
  __syncthreads();   // xs is dead from here on
#pragma unroll
  for (int r = 0; r < ROWS; ++r) sm[(warp * ROWS + r) * 32 + lane] = acc[r];
  __syncthreads();
  // warp r finishes row r (ROWS <= kSkinnyWarps)
  if (warp >= ROWS || warp >= M) return;
  const int row = warp;
  float v = 0.f;
#pragma unroll
  for (int w = 0; w < kSkinnyWarps; ++w) v += sm[(w * ROWS + row) * 32 + lane];
  if (EPI == 1) {
    // lanes (2j, 2j+1) hold (re, im) of bin col/2
    const float sq = v * v;
    const float p = sq + __shfl_xor_sync(0xffffffffu, sq, 1);
    const int f = col >> 1;
    if ((lane & 1) == 0 && f < g.N) g.C[(int64_t)row * g.ldc + f] = p;
  } else {
    if (col >= g.N) return;
    if (g.bias) v += g.bias[col];
    const bool res_first = (g.act & VADX_ACT_RES_FIRST) != 0;
    if (!res_first) v = apply_act(v, g.act);
    if (g.res) v += g.res[(int64_t)row * g.ldr + col];
    if (res_first) v = apply_act(v, g.act);
    g.C[(int64_t)row * g.ldc + col] = v;
  }
}

template <int ALOAD, int EPI>
static bool skinny_fits(const GemmArgs& g) {
  if (g.M > kSkinnyRows) return false;
  const int rows = g.M <= 4 ? 4 : (g.M <= 8 ? 8 : 16);
  return (size_t)rows * g.K * 4 <= (size_t)kSkinnySmem;
}

template <int ALOAD, int EPI>
static int launch_skinny(const GemmArgs& g, int n_cols, cudaStream_t st, const char* what) {
  const unsigned grid = (unsigned)ceil_div(n_cols, 32);
  const int rows = g.M <= 4 ? 4 : (g.M <= 8 ? 8 : 16);
  const size_t smem = std::max((size_t)rows * g.K * 4, (size_t)kSkinnyWarps * rows * 32 * 4);
  if (rows == 4) gemm_skinny_kernel<4, ALOAD, EPI><<<grid, kSkinnyWarps * 32, smem, st>>>(g);
  else if (rows == 8) gemm_skinny_kernel<8, ALOAD, EPI><<<grid, kSkinnyWarps * 32, smem, st>>>(g);
  else gemm_skinny_kernel<16, ALOAD, EPI><<<grid, kSkinnyWarps * 32, smem, st>>>(g);
  return after_launch(what);
}

template <int ALOAD, int EPI>
static int launch_gemm(const GemmArgs& g, int n_cols, cudaStream_t st, const char* what) {
  if (g.M == 0) return VADX_OK;
  if (skinny_fits<ALOAD, EPI>(g)) return launch_skinny<ALOAD, EPI>(g, n_cols, st, what);
  // pick the column tile that wastes fewer padded columns
  int waste128 = (int)(round_up(n_cols, 128) - n_cols), waste64 = (int)(round_up(n_cols, 64) - n_cols);
  bool use64 = waste64 < waste128;
  int64_t gx = ceil_div(g.M, kBM);
  VADX_REQUIRE(gx <= 0x7fffffffLL, "%s: too many rows (%lld)", what, (long long)g.M);
  if (use64) {
    dim3 grid((unsigned)gx, (unsigned)ceil_div(n_cols, 64));
    gemm_f32_kernel<64, ALOAD, EPI><<<grid, kThreads, 0, st>>>(g);
  } else {
    dim3 grid((unsigned)gx, (unsigned)ceil_div(n_cols, 128));
    gemm_f32_kernel<128, ALOAD, EPI><<<grid, kThreads, 0, st>>>(g);
  }
  return after_launch(what);
}

int linear_narrow(const float* d_x, int64_t ldx, const float* d_wt, int ldw, const float* d_bias, float* d_y,
                  int64_t n_rows, int n_in, int n_out, int act, int rows_per_group, int64_t group_stride,
                  int64_t out_stride, cudaStream_t st) {
  StageTimer _timer(VADX_STAGE_HEAD, st, "linear_narrow_kernel", 4.0 * n_rows * (n_in + n_out), 2.0 * n_rows * n_in * n_out);
  VADX_REQUIRE(n_out >= 1 && n_out <= 8, "linear_narrow: n_out=%d not in [1,8]", n_out);
  if (n_rows == 0) return VADX_OK;
  int64_t blocks = ceil_div(n_rows, 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  linear_narrow_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_x, ldx, d_wt, ldw, d_bias, d_y, n_rows, n_in, n_out, act,
                                                         rows_per_group, group_stride, out_stride);
  return after_launch("linear_narrow");
}

}  // namespace vadx

using namespace vadx;

extern "C" int vadx_linear_f32(const float* d_x, int64_t ldx, const float* d_wt, int ldw, const float* d_bias,
                               const float* d_residual, int64_t ldr, float* d_y, int64_t ldy, int64_t n_rows,
                               int n_in, int n_out, int act, void* stream) {
  StageTimer _timer(VADX_STAGE_LINEAR, (cudaStream_t)stream, n_rows <= kSkinnyMaxRows ? "gemm_skinny_kernel" : "gemm_f32_kernel(linear)",
                    4.0 * n_rows * (n_in + n_out + (d_residual ? n_out : 0)), 2.0 * n_rows * n_in * n_out);
  VADX_REQUIRE(d_x && d_wt && d_y, "vadx_linear_f32: null pointer");
  VADX_REQUIRE(n_rows >= 0 && n_in > 0 && n_out > 0, "vadx_linear_f32: bad shape rows=%lld in=%d out=%d",
               (long long)n_rows, n_in, n_out);
  VADX_REQUIRE(ldw >= n_out && (ldw & 3) == 0 && aligned16(d_wt),
               "vadx_linear_f32: ldw=%d must be >= n_out=%d, a multiple of 4, and d_wt 16B aligned", ldw, n_out);
  VADX_REQUIRE(ldx >= n_in && ldy >= n_out, "vadx_linear_f32: row strides smaller than the row");
  VADX_REQUIRE((act & 15) <= VADX_ACT_SOFTMAX, "vadx_linear_f32: activation %d exists on the tensor-core path only", act);
  cudaStream_t st = (cudaStream_t)stream;
  if (n_out <= 8 && !d_residual)
    return linear_narrow(d_x, ldx, d_wt, ldw, d_bias, d_y, n_rows, n_in, n_out, act, 1, ldy, 1, st);
  GemmArgs g{};
  g.A = d_x; g.lda = ldx; g.n_frames = 1; g.hop = 0;
  g.B = d_wt; g.ldb = ldw; g.M = n_rows; g.N = n_out; g.K = n_in;
  g.bias = d_bias; g.res = d_residual; g.ldr = ldr; g.C = d_y; g.ldc = ldy; g.act = act;
  g.vec_store = ((ldy & 3) == 0) && aligned16(d_y) && (!d_residual || (((ldr & 3) == 0) && aligned16(d_residual)));
  return launch_gemm<0, 0>(g, n_out, st, "vadx_linear_f32");
}

extern "C" int vadx_stft_power_f32(const float* d_sig, int64_t sig_stride, int64_t n_streams, int n_frames, int hop,
                                   int n_taps, const float* d_basis, int ld_basis, int n_bins, float* d_power,
                                   int64_t ld_power, void* stream) {
  StageTimer _timer(VADX_STAGE_STFT, (cudaStream_t)stream, "gemm_f32_kernel(stft_power)",
                    4.0 * n_streams * ((n_frames - 1) * (double)hop + n_taps) + 4.0 * n_streams * n_frames * n_bins,
                    2.0 * n_streams * n_frames * n_taps * 2.0 * n_bins);
  VADX_REQUIRE(d_sig && d_basis && d_power, "vadx_stft_power_f32: null pointer");
  VADX_REQUIRE(n_streams >= 0 && n_frames > 0 && hop > 0 && n_taps > 0 && n_bins > 0,
               "vadx_stft_power_f32: bad shape");
  VADX_REQUIRE(ld_basis >= 2 * n_bins && (ld_basis & 3) == 0 && aligned16(d_basis),
               "vadx_stft_power_f32: ld_basis=%d must be >= 2*n_bins=%d and a multiple of 4", ld_basis, 2 * n_bins);
  VADX_REQUIRE(ld_power >= n_bins, "vadx_stft_power_f32: ld_power < n_bins");
  VADX_REQUIRE(sig_stride >= (int64_t)(n_frames - 1) * hop + n_taps,
               "vadx_stft_power_f32: signal stride %lld shorter than the last frame (%lld)", (long long)sig_stride,
               (long long)((int64_t)(n_frames - 1) * hop + n_taps));
  GemmArgs g{};
  g.A = d_sig; g.lda = sig_stride; g.n_frames = n_frames; g.hop = hop;
  g.B = d_basis; g.ldb = ld_basis; g.M = n_streams * n_frames; g.N = n_bins; g.K = n_taps;
  g.C = d_power; g.ldc = ld_power; g.act = 0;
  g.vec_store = ((ld_power & 1) == 0) && ((reinterpret_cast<uintptr_t>(d_power) & 7u) == 0);
  return launch_gemm<1, 1>(g, 2 * n_bins, (cudaStream_t)stream, "vadx_stft_power_f32");
}

extern "C" int vadx_stft_complex_f32(const float* d_sig, int64_t sig_stride, int64_t n_streams, int n_frames, int hop,
                                     int n_taps, const float* d_basis, int ld_basis, int n_bins, float* d_out,
                                     int64_t ld_out, void* stream) {
  StageTimer _timer(VADX_STAGE_STFT, (cudaStream_t)stream, "gemm_f32_kernel(stft_complex)",
                    4.0 * n_streams * ((n_frames - 1) * (double)hop + n_taps) + 8.0 * n_streams * n_frames * n_bins,
                    2.0 * n_streams * n_frames * n_taps * 2.0 * n_bins);
  VADX_REQUIRE(d_sig && d_basis && d_out, "vadx_stft_complex_f32: null pointer");
  VADX_REQUIRE(n_streams >= 0 && n_frames > 0 && hop > 0 && n_taps > 0 && n_bins > 0, "vadx_stft_complex_f32: bad shape");
  VADX_REQUIRE(ld_basis >= 2 * n_bins && (ld_basis & 3) == 0 && aligned16(d_basis) && ld_out >= 2 * n_bins,
               "vadx_stft_complex_f32: ld_basis must be >= 2*n_bins and a multiple of 4, ld_out >= 2*n_bins");
  VADX_REQUIRE(sig_stride >= (int64_t)(n_frames - 1) * hop + n_taps, "vadx_stft_complex_f32: signal stride too short");
  GemmArgs g{};
  g.A = d_sig; g.lda = sig_stride; g.n_frames = n_frames; g.hop = hop;
  g.B = d_basis; g.ldb = ld_basis; g.M = n_streams * n_frames; g.N = 2 * n_bins; g.K = n_taps;
  g.C = d_out; g.ldc = ld_out; g.act = 0;
  g.vec_store = ((ld_out & 3) == 0) && aligned16(d_out);
  return launch_gemm<1, 0>(g, 2 * n_bins, (cudaStream_t)stream, "vadx_stft_complex_f32");
}
