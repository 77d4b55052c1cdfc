// elementwise.cu -- the HBM-bound stages: audio preparation (a2), filterbank + log (a3) and the
// FSMN / DFSMN memory block (a5-a7).
#include "common.cuh"

namespace vadx {

// ------------------------------------------------------------------------------------------ a2
template <typename T>
__device__ __forceinline__ float load_sample(const T* p, int64_t i) {
  return (float)p[i];
}

// grid (chunks, S); when remove_dc the launch uses one chunk per stream and the block first
// reduces the stream's mean.
template <typename T>
__global__ void __launch_bounds__(256) prep_audio_kernel(const T* __restrict__ audio, int64_t n_samples,
                                                         int64_t in_stride, float scale, int remove_dc,
                                                         int preemph_mode, float c, int64_t pad_left,
                                                         float* __restrict__ out, int64_t out_stride) {
  const int64_t s = blockIdx.y;
  const T* x = audio + s * in_stride;
  float* y = out + s * out_stride;
  __shared__ double red[256];
  float mean = 0.f;
  if (remove_dc) {
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < n_samples; i += blockDim.x) acc += (double)(load_sample(x, i) * scale);
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
      if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
      __syncthreads();
    }
    mean = (float)(red[0] / (double)n_samples);
  }
  const int64_t per_block = ceil_div_dev(out_stride, (int64_t)gridDim.x);
  const int64_t lo = (int64_t)blockIdx.x * per_block;
  const int64_t hi = min_i64(lo + per_block, out_stride);
  for (int64_t j = lo + threadIdx.x; j < hi; j += blockDim.x) {
    const int64_t n = j - pad_left;
    float v = 0.f;
    if (n >= 0 && n < n_samples) {
      float cur = load_sample(x, n) * scale - mean;
      if (preemph_mode == VADX_PREEMPH_NONE) {
        v = cur;
      } else if (n == 0) {
        v = cur;  // ZERO_HISTORY: x[-1] = 0; KEEP_FIRST: y[0] = x[0]
      } else {
        float prev = load_sample(x, n - 1) * scale - mean;
        v = cur - c * prev;
      }
    }
    y[j] = v;
  }
}

// ------------------------------------------------------------------------------------------ a3
// Persistent CTAs: the sparse filterbank (weights, start, length per mel) is staged in shared memory ONCE
// per CTA, then 32-row tiles of the power spectrum stream through: a tile is copied as one contiguous
// block of float4 (rows are ld_power apart, so 32 rows are 32*ld_power consecutive floats), each warp
// owns 4 rows and its lanes walk the mels (lane, lane+32, lane+64), so the log-mel row leaves as
// coalesced 128-byte stores.  No integer division anywhere on the per-element path.
constexpr int kMelRows = 32;
__global__ void __launch_bounds__(256) mel_log_kernel(const float* __restrict__ power, int64_t ld_power,
                                                      int64_t n_rows, int n_bins, int n_mels,
                                                      const int32_t* __restrict__ start,
                                                      const int32_t* __restrict__ len, const float* __restrict__ w,
                                                      int max_len, int floor_mode, float floor_value,
                                                      float* __restrict__ out, int64_t ld_out, int vec_ok) {
  extern __shared__ __align__(16) float tile[];  // [kMelRows][ldt] | w [n_mels][wl] | start, len [n_mels]
  const int ldt = (int)ld_power;                 // tile rows keep the global pitch (one block copy)
  const int wl = max_len | 1;                    // odd pitch: lanes (consecutive mels) hit different banks
  float* ws = tile + kMelRows * ldt;
  int* ss = reinterpret_cast<int*>(ws + n_mels * wl);
  int* ls = ss + n_mels;
  for (int i = threadIdx.x; i < n_mels * max_len; i += blockDim.x) {
    const int m = i / max_len;
    ws[m * wl + (i - m * max_len)] = w[i];
  }
  for (int i = threadIdx.x; i < n_mels; i += blockDim.x) {
    ss[i] = start[i];
    ls[i] = len[i];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_tiles = (n_rows + kMelRows - 1) / kMelRows;
  for (int64_t tl = blockIdx.x; tl < n_tiles; tl += gridDim.x) {
    const int64_t r0 = tl * kMelRows;
    const int rows = (int)min_i64((int64_t)kMelRows, n_rows - r0);
    __syncthreads();  // previous tile fully consumed (and the tables visible on the first pass)
    const float* src = power + r0 * ld_power;
    if (vec_ok) {
      const int n4 = rows * ldt / 4;
      const float4* s4 = reinterpret_cast<const float4*>(src);
      float4* t4 = reinterpret_cast<float4*>(tile);
      for (int i = threadIdx.x; i < n4; i += blockDim.x) t4[i] = __ldg(s4 + i);
    } else {
      for (int i = threadIdx.x; i < rows * ldt; i += blockDim.x) tile[i] = src[i];
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < kMelRows / 8; ++rr) {
      const int r = warp * (kMelRows / 8) + rr;
      if (r >= rows) break;
      const float* prow = tile + r * ldt;
      float* orow = out + (r0 + r) * ld_out;
      for (int m = lane; m < n_mels; m += 32) {
        const float* p = prow + ss[m];
        const float* wm = ws + m * wl;
        const int L = ls[m];
        float acc = 0.f;
        for (int j = 0; j < L; ++j) acc = fmaf(wm[j], p[j], acc);
        acc = (floor_mode == VADX_FLOOR_CLAMP) ? fmaxf(acc, floor_value) : acc + floor_value;
        orow[m] = logf(acc);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ a5-a7
// One CTA = one stream x a tile of kMemT output frames x all channels.  The tile plus its halo and
// the transposed taps are staged in shared memory; each thread then produces outputs for one
// channel (consecutive threads = consecutive channels, so both the global and the shared accesses
// are unit-stride across a warp).
constexpr int kMemT = 64;
__global__ void __launch_bounds__(256) fsmn_memory_kernel(const float* __restrict__ p, int64_t ldp,
                                                          const float* __restrict__ wl, int n_back, int s_back,
                                                          const float* __restrict__ wr, int n_ahead, int s_ahead,
                                                          const float* __restrict__ res, int64_t ldr,
                                                          float* __restrict__ out, int64_t ldo, int n_frames,
                                                          int C, const float* __restrict__ cache_in) {
  extern __shared__ float sm[];
  const int halo_l = (n_back - 1) * s_back;
  const int halo_r = (n_frames > 1) ? n_ahead * s_ahead : 0;
  const int use_ahead = (n_frames > 1) ? n_ahead : 0;
  const int rows_in = kMemT + halo_l + halo_r;
  float* tile = sm;                          // [rows_in][C]
  float* wls = tile + (size_t)rows_in * C;   // [n_back][C]
  float* wrs = wls + (size_t)n_back * C;     // [use_ahead][C]
  const int64_t s = blockIdx.y;
  const int t0 = blockIdx.x * kMemT;
  const float* ps = p + s * (int64_t)n_frames * ldp;

  for (int i = threadIdx.x; i < rows_in * C; i += blockDim.x) {
    int r = i / C, c = i - r * C;
    int t = t0 - halo_l + r;
    float v = 0.f;
    if (t >= 0 && t < n_frames) v = ps[(int64_t)t * ldp + c];
    else if (t < 0 && cache_in) v = cache_in[(s * C + c) * (int64_t)halo_l + (halo_l + t)];
    tile[i] = v;
  }
  for (int i = threadIdx.x; i < n_back * C; i += blockDim.x) {
    int k = i / C, c = i - k * C;
    wls[i] = wl[c * n_back + k];
  }
  for (int i = threadIdx.x; i < use_ahead * C; i += blockDim.x) {
    int k = i / C, c = i - k * C;
    wrs[i] = wr[c * n_ahead + k];
  }
  __syncthreads();

  const int t_end = min(kMemT, n_frames - t0);
  for (int i = threadIdx.x; i < t_end * C; i += blockDim.x) {
    int tt = i / C, c = i - tt * C;
    const float* col = tile + (size_t)(tt + halo_l) * C + c;  // p[t]
    float acc = col[0];
    // look-back: sum_k wl[k] * p[t - (n_back-1-k)*s_back]
    const float* src = col - (size_t)halo_l * C;
    for (int k = 0; k < n_back; ++k) acc = fmaf(wls[k * C + c], src[(size_t)k * s_back * C], acc);
    // look-ahead: sum_k wr[k] * p[t + (k+1)*s_ahead]
    for (int k = 0; k < use_ahead; ++k) acc = fmaf(wrs[k * C + c], col[(size_t)(k + 1) * s_ahead * C], acc);
    const int64_t row = s * (int64_t)n_frames + t0 + tt;
    if (res) acc += res[row * ldr + c];
    out[row * ldo + c] = acc;
  }
}


// Streaming variant for unit strides: thread = (stream, time tile, channel).  The taps and a
// (R + N1 - 1 + N2)-deep delay line live in registers; every input is read from global memory once
// per tile (a warp reads 32 consecutive channels of one frame = one 128-byte line), the next R
// inputs are prefetched while the current R outputs are computed, and no shared memory or barrier
// is involved, so many independent warps keep loads in flight (HBM-bound by design).
constexpr int kMemR = 14;
template <int N1, int N2>
__global__ void __launch_bounds__(128, 4) fsmn_memory_stream_kernel(const float* __restrict__ p, int64_t ldp,
                                                                 const float* __restrict__ wl,
                                                                 const float* __restrict__ wr,
                                                                 const float* __restrict__ res, int64_t ldr,
                                                                 float* __restrict__ out, int64_t ldo, int n_frames,
                                                                 int C, const float* __restrict__ cache_in,
                                                                 float* __restrict__ cache_out, int tile_t) {
  constexpr int HL = N1 - 1, HR = N2, W = kMemR + HL + HR;
  const int c = blockIdx.z * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int64_t s = blockIdx.y;
  const int t0 = blockIdx.x * tile_t;
  const int t1 = min(t0 + tile_t, n_frames);
  const float* ps = p + s * (int64_t)n_frames * ldp + c;
  const float* rs = res ? res + s * (int64_t)n_frames * ldr + c : nullptr;
  float* os = out + s * (int64_t)n_frames * ldo + c;
  auto load = [&](int t) -> float {
    if (t >= 0 && t < n_frames) return __ldg(ps + (int64_t)t * ldp);
    if (t < 0 && cache_in) return cache_in[(s * C + c) * (int64_t)HL + (HL + t)];
    return 0.f;
  };
  // taps in shared memory, one private column per thread (written and read by the same thread, so no
  // barrier): frees 40 registers -> 4 CTAs per SM instead of 3, i.e. a third more loads in flight
  __shared__ float taps[N1 + N2][128];
#pragma unroll
  for (int k = 0; k < N1; ++k) taps[k][threadIdx.x] = __ldg(wl + c * N1 + k);
#pragma unroll
  for (int k = 0; k < N2; ++k) taps[N1 + k][threadIdx.x] = __ldg(wr + c * N2 + k);
  float w[W];
#pragma unroll
  for (int j = 0; j < HL + HR; ++j) w[j] = load(t0 - HL + j);
  float nxt[kMemR], rv[kMemR];
#pragma unroll
  for (int r = 0; r < kMemR; ++r) nxt[r] = load(t0 + HR + r);
  for (int tb = t0; tb < t1; tb += kMemR) {
#pragma unroll
    for (int r = 0; r < kMemR; ++r) w[HL + HR + r] = nxt[r];
    // prefetch the next group and this group's residual; groups whose reads cannot leave
    // [0, n_frames) (all but the edges of the stream) take the unguarded path
    const int tn = tb + kMemR + HR;  // first frame of the next group's new inputs
    if (tb + kMemR < t1) {
      if (tn + kMemR <= n_frames) {
        const float* q = ps + (int64_t)tn * ldp;
#pragma unroll
        for (int r = 0; r < kMemR; ++r) nxt[r] = __ldg(q + (int64_t)r * ldp);
      } else {
#pragma unroll
        for (int r = 0; r < kMemR; ++r) nxt[r] = load(tn + r);
      }
    }
    const bool full = tb + kMemR <= t1;
    if (rs) {
      const float* q = rs + (int64_t)tb * ldr;
      if (full) {
#pragma unroll
        for (int r = 0; r < kMemR; ++r) rv[r] = __ldg(q + (int64_t)r * ldr);
      } else {
#pragma unroll
        for (int r = 0; r < kMemR; ++r) rv[r] = (tb + r < t1) ? __ldg(q + (int64_t)r * ldr) : 0.f;
      }
    }
    float acc[kMemR];
#pragma unroll
    for (int r = 0; r < kMemR; ++r) acc[r] = w[r + HL];
#pragma unroll
    for (int k = 0; k < N1; ++k) {
      const float ck = taps[k][threadIdx.x];
#pragma unroll
      for (int r = 0; r < kMemR; ++r) acc[r] = fmaf(ck, w[r + k], acc[r]);
    }
#pragma unroll
    for (int k = 0; k < N2; ++k) {
      const float ck = taps[N1 + k][threadIdx.x];
#pragma unroll
      for (int r = 0; r < kMemR; ++r) acc[r] = fmaf(ck, w[r + N1 + k], acc[r]);
    }
    float* o = os + (int64_t)tb * ldo;
    if (full) {
#pragma unroll
      for (int r = 0; r < kMemR; ++r) o[(int64_t)r * ldo] = rs ? acc[r] + rv[r] : acc[r];
    } else {
#pragma unroll
      for (int r = 0; r < kMemR; ++r)
        if (tb + r < t1) o[(int64_t)r * ldo] = rs ? acc[r] + rv[r] : acc[r];
    }
#pragma unroll
    for (int j = 0; j < HL + HR; ++j) w[j] = w[j + kMemR];
  }
  // streaming hand-over: the last HL frames of cat(cache_in, p), written by the stream's last tile
  if (cache_out && t1 == n_frames) {
    float* co = cache_out + (s * C + c) * (int64_t)HL;
#pragma unroll
    for (int j = 0; j < HL; ++j) co[j] = load(n_frames - HL + j);
  }
}

// cache_out[s][c][j] = cat(cache_in, p)[T + j] for j in [0, halo): the streaming state hand-over
__global__ void __launch_bounds__(256) fsmn_cache_out_kernel(const float* __restrict__ p, int64_t ldp,
                                                             const float* __restrict__ cache_in,
                                                             float* __restrict__ cache_out, int64_t n_streams,
                                                             int n_frames, int C, int halo) {
  const int64_t total = n_streams * C * halo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int j = (int)(i % halo);
    int64_t sc = i / halo;
    int c = (int)(sc % C);
    int64_t s = sc / C;
    int t = n_frames - halo + j;
    float v;
    if (t >= 0) v = p[(s * n_frames + t) * ldp + c];
    else v = cache_in ? cache_in[sc * halo + (halo + t)] : 0.f;
    cache_out[i] = v;
  }
}

}  // namespace vadx

namespace vadx {
bool memory_bulk_fits(int n_back, int stride_back, int n_ahead, int stride_ahead, int n_frames, int n_channels,
                      int64_t ldp, int64_t ldr, int64_t ldo, const float* p, const float* res, const float* out,
                      const float* cache_in, const float* cache_out, size_t* smem_out);
int memory_bulk_launch(const float* p, const float* wl, const float* wr, int n_ahead, const float* res, float* out,
                       int64_t n_streams, int n_frames, size_t smem, cudaStream_t st);
}  // namespace vadx

using namespace vadx;

extern "C" int vadx_prep_audio(const void* d_audio, int in_dtype, int64_t n_streams, int64_t n_samples,
                               int64_t in_stride, float scale, int remove_dc, int preemph_mode, float preemph,
                               int64_t pad_left, float* d_out, int64_t out_stride, void* stream) {
  StageTimer _timer(VADX_STAGE_PREP, (cudaStream_t)stream, "prep_audio_kernel", n_streams * (double)n_samples * (in_dtype == VADX_DT_I16 ? 2.0 : 4.0) + 4.0 * n_streams * out_stride);
  VADX_REQUIRE(d_audio && d_out, "vadx_prep_audio: null pointer");
  VADX_REQUIRE(in_dtype == VADX_DT_I16 || in_dtype == VADX_DT_F32, "vadx_prep_audio: dtype %d", in_dtype);
  // in_stride < n_samples is allowed: the rows are then OVERLAPPING windows of one recording (the input is only read)
  VADX_REQUIRE(n_streams >= 0 && n_samples > 0 && in_stride >= 1 && pad_left >= 0 &&
                   out_stride >= pad_left + n_samples,
               "vadx_prep_audio: bad shape S=%lld L=%lld in_stride=%lld pad_left=%lld out_stride=%lld",
               (long long)n_streams, (long long)n_samples, (long long)in_stride, (long long)pad_left,
               (long long)out_stride);
  VADX_REQUIRE(preemph_mode >= 0 && preemph_mode <= 2, "vadx_prep_audio: preemph_mode %d", preemph_mode);
  if (n_streams == 0) return VADX_OK;
  int chunks = remove_dc ? 1 : (int)std::min<int64_t>(ceil_div(out_stride, 256 * 16), 1024);
  cudaStream_t st = (cudaStream_t)stream;
  for (int64_t s0 = 0; s0 < n_streams; s0 += 65535) {            // grid.y carries the stream: 65535 rows per launch
    dim3 grid((unsigned)chunks, (unsigned)std::min<int64_t>(65535, n_streams - s0));
    if (in_dtype == VADX_DT_I16)
      prep_audio_kernel<int16_t><<<grid, 256, 0, st>>>((const int16_t*)d_audio + s0 * in_stride, n_samples, in_stride, scale,
                                                       remove_dc, preemph_mode, preemph, pad_left, d_out + s0 * out_stride, out_stride);
    else
      prep_audio_kernel<float><<<grid, 256, 0, st>>>((const float*)d_audio + s0 * in_stride, n_samples, in_stride, scale, remove_dc,
                                                     preemph_mode, preemph, pad_left, d_out + s0 * out_stride, out_stride);
  }
  return after_launch("vadx_prep_audio");
}

extern "C" int vadx_mel_log_f32(const float* d_power, int64_t ld_power, int64_t n_rows, int n_bins, int n_mels,
                                const int32_t* d_start, const int32_t* d_len, const float* d_w, int max_len,
                                int floor_mode, float floor_value, float* d_out, int64_t ld_out, void* stream) {
  StageTimer _timer(VADX_STAGE_MEL, (cudaStream_t)stream, "mel_log_kernel", 4.0 * n_rows * (n_bins + n_mels), 2.0 * n_rows * n_mels * (double)max_len);
  VADX_REQUIRE(d_power && d_start && d_len && d_w && d_out, "vadx_mel_log_f32: null pointer");
  VADX_REQUIRE(n_rows >= 0 && n_bins > 0 && n_mels > 0 && max_len > 0 && ld_power >= n_bins && ld_out >= n_mels,
               "vadx_mel_log_f32: bad shape");
  VADX_REQUIRE(floor_mode == VADX_FLOOR_CLAMP || floor_mode == VADX_FLOOR_ADD, "vadx_mel_log_f32: floor_mode");
  if (n_rows == 0) return VADX_OK;
  size_t smem = ((size_t)kMelRows * ld_power + (size_t)n_mels * (max_len | 1) + 2 * (size_t)n_mels) * sizeof(float);
  VADX_REQUIRE(smem <= 200 * 1024, "vadx_mel_log_f32: ld_power=%lld too large", (long long)ld_power);
  static PerDevice per_device;
  int n_sm = 148;
  VADX_TRY(per_device.ensure(&n_sm, [] {
    return cudaFuncSetAttribute(mel_log_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }));
  const int64_t tiles = ceil_div(n_rows, kMelRows);
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / smem));   // resident CTAs per SM
  const int64_t blocks = std::min<int64_t>(tiles, (int64_t)n_sm * per_sm);
  const int vec_ok = ((ld_power & 3) == 0) && aligned16(d_power);
  mel_log_kernel<<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(d_power, ld_power, n_rows, n_bins, n_mels,
                                                                        d_start, d_len, d_w, max_len, floor_mode,
                                                                        floor_value, d_out, ld_out, vec_ok);
  return after_launch("vadx_mel_log_f32");
}

extern "C" int vadx_fsmn_memory_f32(const float* d_p, int64_t ldp, const float* d_wl, int n_back, int stride_back,
                                    const float* d_wr, int n_ahead, int stride_ahead, const float* d_residual,
                                    int64_t ldr, float* d_out, int64_t ldo, int64_t n_streams, int n_frames,
                                    int n_channels, const float* d_cache_in, float* d_cache_out, void* stream) {
  StageTimer _timer(VADX_STAGE_MEMORY, (cudaStream_t)stream, "fsmn_memory_kernel",
                    4.0 * n_streams * n_frames * n_channels * (2 + (d_residual ? 1 : 0)),      // p (+ residual) in, out
                    2.0 * n_streams * n_frames * n_channels * (double)(n_back + n_ahead));
  VADX_REQUIRE(d_p && d_wl && d_out, "vadx_fsmn_memory_f32: null pointer");
  VADX_REQUIRE(n_back >= 1 && stride_back >= 1 && n_ahead >= 0 && (n_ahead == 0 || (d_wr && stride_ahead >= 1)),
               "vadx_fsmn_memory_f32: bad taps back=%d/%d ahead=%d/%d", n_back, stride_back, n_ahead, stride_ahead);
  VADX_REQUIRE(n_streams >= 0 && n_frames >= 1 && n_channels >= 1 && ldp >= n_channels && ldo >= n_channels,
               "vadx_fsmn_memory_f32: bad shape");
  VADX_REQUIRE(n_streams <= 65535, "vadx_fsmn_memory_f32: at most 65535 streams per call");
  VADX_REQUIRE(d_p != d_out, "vadx_fsmn_memory_f32: in-place operation is not supported");
  VADX_REQUIRE(!d_cache_out || d_cache_out != d_cache_in, "vadx_fsmn_memory_f32: cache_out must not alias cache_in");
  if (n_streams == 0) return VADX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int halo_l = (n_back - 1) * stride_back;
  const int halo_r = n_frames > 1 ? n_ahead * stride_ahead : 0;
  size_t smem = ((size_t)(kMemT + halo_l + halo_r) + n_back + n_ahead) * n_channels * sizeof(float);
  VADX_REQUIRE(smem <= 200 * 1024, "vadx_fsmn_memory_f32: tile of %zu bytes exceeds shared memory", smem);
  static PerDevice per_device;
  VADX_TRY(per_device.ensure(nullptr, [] {
    return cudaFuncSetAttribute(fsmn_memory_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }));
  // whole-chunk shared-memory kernel (memory_bulk.cu) when the stream's tile fits: FireRed's [98][128]
  {
    size_t smem_b = 0;
    static const int no_bulk = ab_env("VADX_MEM_NO_BULK", 0);
    if (!no_bulk && n_streams >= 32 &&
        memory_bulk_fits(n_back, stride_back, n_ahead, stride_ahead, n_frames, n_channels, ldp, ldr, ldo, d_p, d_residual,
                         d_out, d_cache_in, d_cache_out, &smem_b))
      return memory_bulk_launch(d_p, d_wl, d_wr, n_ahead, d_residual, d_out, n_streams, n_frames, smem_b, st);
  }
  // fast path: unit strides, 20(+20) taps
  const bool fast = stride_back == 1 && (n_ahead == 0 || stride_ahead == 1) && n_back == 20 &&
                    (n_ahead == 20 || n_ahead == 0) && n_frames > 1;
  if (fast) {
    const int tile_t = (int)std::min<int64_t>(round_up(n_frames, kMemR), 8 * kMemR);
    dim3 grid2((unsigned)ceil_div(n_frames, tile_t), (unsigned)n_streams, (unsigned)ceil_div(n_channels, 128));
    if (n_ahead == 20)
      fsmn_memory_stream_kernel<20, 20><<<grid2, 128, 0, st>>>(d_p, ldp, d_wl, d_wr, d_residual, ldr, d_out, ldo,
                                                               n_frames, n_channels, d_cache_in, d_cache_out, tile_t);
    else
      fsmn_memory_stream_kernel<20, 0><<<grid2, 128, 0, st>>>(d_p, ldp, d_wl, d_wr, d_residual, ldr, d_out, ldo,
                                                              n_frames, n_channels, d_cache_in, d_cache_out, tile_t);
  } else {
    dim3 grid((unsigned)ceil_div(n_frames, kMemT), (unsigned)n_streams);
    fsmn_memory_kernel<<<grid, 256, smem, st>>>(d_p, ldp, d_wl, n_back, stride_back, d_wr, n_ahead, stride_ahead,
                                                d_residual, ldr, d_out, ldo, n_frames, n_channels, d_cache_in);
  }
  VADX_TRY(after_launch("vadx_fsmn_memory_f32"));
  if (d_cache_out && halo_l > 0 && !fast) {
    int64_t total = n_streams * n_channels * halo_l;
    int64_t blocks = std::min<int64_t>(ceil_div(total, 256), 148 * 8);
    fsmn_cache_out_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_p, ldp, d_cache_in, d_cache_out, n_streams, n_frames,
                                                            n_channels, halo_l);
    VADX_TRY(after_launch("fsmn_cache_out"));
  }
  return VADX_OK;
}

// ------------------------------------------------------------------------------------------ a12
// Depthwise conv1d over time (Jasper separable blocks), time-major.  One CTA = one stream x a tile of
// kDwT output frames x all channels; input rows incl. halo and the transposed taps are staged in
// shared memory, consecutive threads own consecutive channels.
namespace vadx {
constexpr int kDwT = 32;
__global__ void __launch_bounds__(256) depthwise_conv1d_kernel(const float* __restrict__ x, int64_t ldx,
                                                               const float* __restrict__ w, int K, int stride,
                                                               int dil, int pad, float* __restrict__ y, int64_t ldy,
                                                               int t_in, int t_out, int C) {
  extern __shared__ float sm[];
  const int rows_in = (kDwT - 1) * stride + (K - 1) * dil + 1;
  float* tile = sm;                        // [rows_in][C]
  float* ws = tile + (size_t)rows_in * C;  // [K][C]
  const int64_t s = blockIdx.y;
  const int o0 = blockIdx.x * kDwT;
  const int i0 = o0 * stride - pad;
  const float* xs = x + s * (int64_t)t_in * ldx;
  for (int i = threadIdx.x; i < rows_in * C; i += blockDim.x) {
    int r = i / C, c = i - r * C;
    int t = i0 + r;
    tile[i] = (t >= 0 && t < t_in) ? xs[(int64_t)t * ldx + c] : 0.f;
  }
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) {
    int k = i / C, c = i - k * C;
    ws[i] = w[c * K + k];
  }
  __syncthreads();
  const int n_out = min(kDwT, t_out - o0);
  for (int i = threadIdx.x; i < n_out * C; i += blockDim.x) {
    int tt = i / C, c = i - tt * C;
    const float* src = tile + (size_t)(tt * stride) * C + c;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(ws[k * C + c], src[(size_t)k * dil * C], acc);
    y[(s * (int64_t)t_out + o0 + tt) * ldy + c] = acc;
  }
}
}  // namespace vadx

namespace vadx {
// Register-window depthwise conv for the shapes the Jasper blocks use (compile-time taps / stride / dilation, channels a
// multiple of 4): a CTA stages TS = groups * 8 output frames (+ halo) of one stream as float4 rows, thread = (channel quad,
// 8 consecutive outputs) walks its input rows ONCE -- every float4 it reads feeds all the outputs and taps it belongs to,
// the taps live in registers -- and a CTA runs kDwTiles consecutive tiles so the tap loads are amortised.  The generic
// kernel above reads one shared-memory scalar per FMA and re-loads a 32-frame tile's halo (up to 175 % at 29 taps, dilation 2).
constexpr int kDwR = 8;       // outputs per thread
constexpr int kDwTiles = 4;   // consecutive tiles per CTA
template <int K, int STRIDE, int DIL, int MINB>
__global__ void __launch_bounds__(256, MINB) depthwise_conv1d_reg_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                        int pad, float* __restrict__ y, int t_in, int t_out, int C) {
  constexpr int NI = (kDwR - 1) * STRIDE + (K - 1) * DIL + 1;      // input rows one thread reads per tile
  extern __shared__ float4 dw_tile[];                               // [rows_in][C / 4]
  const int C4 = C >> 2;
  const int groups = 256 / C4;
  const int ts = groups * kDwR;                                     // outputs per tile
  const int rows_in = (ts - 1) * STRIDE + (K - 1) * DIL + 1;
  const int q = threadIdx.x % C4, g = threadIdx.x / C4;
  const int64_t s = blockIdx.y;
  const float4* xs = reinterpret_cast<const float4*>(x + s * (int64_t)t_in * C);
  float4* ys = reinterpret_cast<float4*>(y + s * (int64_t)t_out * C);
  float4 wq[K];
  if (g < groups) {
#pragma unroll
    for (int k = 0; k < K; ++k)
      wq[k] = make_float4(__ldg(w + (4 * q + 0) * K + k), __ldg(w + (4 * q + 1) * K + k), __ldg(w + (4 * q + 2) * K + k),
                          __ldg(w + (4 * q + 3) * K + k));
  }
  for (int tile = 0; tile < kDwTiles; ++tile) {
    const int o0 = (blockIdx.x * kDwTiles + tile) * ts;
    if (o0 >= t_out) break;
    const int i0 = o0 * STRIDE - pad;
    __syncthreads();                                                // the previous tile's readers are done
    for (int i = threadIdx.x; i < rows_in * C4; i += 256) {
      const int r = i / C4, t = i0 + r;
      dw_tile[i] = (t >= 0 && t < t_in) ? __ldg(xs + (int64_t)t * C4 + (i - r * C4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    if (g < groups) {
      float4 acc[kDwR];
#pragma unroll
      for (int j = 0; j < kDwR; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* src = dw_tile + (size_t)(g * kDwR * STRIDE) * C4 + q;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const float4 v = src[(size_t)i * C4];
#pragma unroll
        for (int j = 0; j < kDwR; ++j) {
          const int d = i - j * STRIDE;                             // compile-time after unrolling
          if (d >= 0 && d % DIL == 0 && d / DIL < K) {
            const float4 t = wq[d / DIL];
            acc[j].x = fmaf(t.x, v.x, acc[j].x); acc[j].y = fmaf(t.y, v.y, acc[j].y);
            acc[j].z = fmaf(t.z, v.z, acc[j].z); acc[j].w = fmaf(t.w, v.w, acc[j].w);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < kDwR; ++j) {
        const int o = o0 + g * kDwR + j;
        if (o < t_out) ys[(int64_t)o * C4 + q] = acc[j];
      }
    }
  }
}

template <int K, int STRIDE, int DIL, int MINB>
static int launch_depthwise_reg(const float* x, const float* w, int pad, float* y, int64_t S, int t_in, int t_out, int C,
                                cudaStream_t st) {
  const int C4 = C / 4, groups = 256 / C4, ts = groups * kDwR;
  const int rows_in = (ts - 1) * STRIDE + (K - 1) * DIL + 1;
  const size_t smem = (size_t)rows_in * C * sizeof(float);
  static PerDevice per_device;
  VADX_TRY(per_device.ensure(nullptr, [] {
    return cudaFuncSetAttribute(depthwise_conv1d_reg_kernel<K, STRIDE, DIL, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                100 * 1024);
  }));
  dim3 grid((unsigned)ceil_div(t_out, (int64_t)ts * kDwTiles), (unsigned)S);
  depthwise_conv1d_reg_kernel<K, STRIDE, DIL, MINB><<<grid, 256, smem, st>>>(x, w, pad, y, t_in, t_out, C);
  return after_launch("vadx_depthwise_conv1d_f32");
}
}  // namespace vadx

extern "C" int vadx_depthwise_conv1d_f32(const float* d_x, int64_t ldx, const float* d_w, int kernel, int stride,
                                         int dilation, int pad, float* d_y, int64_t ldy, int64_t n_streams, int t_in,
                                         int t_out, int n_channels, void* stream) {
  StageTimer _timer(VADX_STAGE_MEMORY, (cudaStream_t)stream, "depthwise_conv1d_kernel",
                    4.0 * n_streams * n_channels * ((double)t_in + t_out), 2.0 * n_streams * t_out * n_channels * (double)kernel);
  VADX_REQUIRE(d_x && d_w && d_y && d_x != d_y, "vadx_depthwise_conv1d_f32: null or aliased pointer");
  VADX_REQUIRE(kernel >= 1 && stride >= 1 && dilation >= 1 && pad >= 0 && n_streams >= 0 && t_in >= 1 && t_out >= 1 &&
                   n_channels >= 1 && ldx >= n_channels && ldy >= n_channels,
               "vadx_depthwise_conv1d_f32: bad shape");
  VADX_REQUIRE((int64_t)(t_out - 1) * stride - pad + (int64_t)(kernel - 1) * dilation < t_in + pad + dilation * kernel,
               "vadx_depthwise_conv1d_f32: t_out inconsistent with t_in");
  VADX_REQUIRE(n_streams <= 65535, "vadx_depthwise_conv1d_f32: at most 65535 streams per call");
  if (n_streams == 0) return VADX_OK;
  // the Jasper shapes on the register-window kernel: dense rows of 16 .. 256 channels (a multiple of 4), 16-byte aligned
  if (ldx == n_channels && ldy == n_channels && (n_channels & 3) == 0 && n_channels >= 16 && n_channels <= 256 &&
      aligned16(d_x) && aligned16(d_y) && (int64_t)t_in * n_channels < (1LL << 40)) {
    cudaStream_t cs = (cudaStream_t)stream;
#define VADX_DW_CASE(KK, SS, DD, MB)                                                                                   \
    if (kernel == KK && stride == SS && dilation == DD)                                                              \
      return launch_depthwise_reg<KK, SS, DD, MB>(d_x, d_w, pad, d_y, n_streams, t_in, t_out, n_channels, cs);
    VADX_DW_CASE(11, 2, 1, 2) VADX_DW_CASE(13, 1, 1, 2) VADX_DW_CASE(15, 1, 1, 2) VADX_DW_CASE(17, 1, 1, 2)
    VADX_DW_CASE(29, 1, 2, 1) VADX_DW_CASE(1, 1, 1, 2)
#undef VADX_DW_CASE
  }
  const int rows_in = (kDwT - 1) * stride + (kernel - 1) * dilation + 1;
  size_t smem = ((size_t)rows_in + kernel) * n_channels * sizeof(float);
  VADX_REQUIRE(smem <= 200 * 1024, "vadx_depthwise_conv1d_f32: tile of %zu bytes exceeds shared memory", smem);
  static PerDevice per_device;
  VADX_TRY(per_device.ensure(nullptr, [] {
    return cudaFuncSetAttribute(depthwise_conv1d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }));
  dim3 grid((unsigned)ceil_div(t_out, kDwT), (unsigned)n_streams);
  depthwise_conv1d_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(d_x, ldx, d_w, kernel, stride, dilation, pad, d_y,
                                                                     ldy, t_in, t_out, n_channels);
  return after_launch("vadx_depthwise_conv1d_f32");
}
