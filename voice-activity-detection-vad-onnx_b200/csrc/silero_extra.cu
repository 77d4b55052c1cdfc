// silero_extra.cu -- Silero-specific stages (a11, a16): reflect-padded window assembly, spectral
// magnitude, LSTM cell update, and the get_speech_timestamps trigger machine on the device.
#include "common.cuh"

namespace vadx {

// out[w][s][j] = x[s][w*step + j] (j < n_in), out[w][s][n_in + j] = x[s][w*step + n_in - 2 - j] (j < pad): the right
// reflect pad of window w of stream s; all W windows of a multi-window call in one launch
__global__ void __launch_bounds__(256) reflect_window_kernel(const float* __restrict__ x, int64_t in_stride,
                                                             int64_t n_streams, int n_windows, int64_t window_step, int n_in,
                                                             int pad, float* __restrict__ out) {
  const int n_out = n_in + pad;
  const int64_t total = n_streams * n_windows * n_out;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / n_out;                 // row = w * n_streams + s
    const int j = (int)(i - row * n_out);
    const int64_t w = row / n_streams, s = row - w * n_streams;
    const int src = j < n_in ? j : n_in - 2 - (j - n_in);
    out[i] = x[s * in_stride + w * window_step + src];
  }
}

__global__ void __launch_bounds__(256) sqrt_kernel(float* __restrict__ p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = sqrtf(p[i]);
}

__global__ void __launch_bounds__(256) affine_kernel(const float* __restrict__ x, float a, float b, float* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = fmaf(a, x[i], b);
}

__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

// PyTorch LSTMCell gate order (i, f, g, o); state = [h ; c] planes of [S][H]
__global__ void __launch_bounds__(256) lstm_cell_kernel(const float* __restrict__ gates, const float* __restrict__ c_in,
                                                        float* __restrict__ h_out, float* __restrict__ c_out,
                                                        float* __restrict__ h_relu, int64_t n_streams, int H) {
  const int64_t total = n_streams * H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t s = i / H;
    int j = (int)(i - s * H);
    const float* g = gates + s * 4 * H;
    float ig = sigmoidf_(g[j]), fg = sigmoidf_(g[H + j]), gg = tanhf(g[2 * H + j]), og = sigmoidf_(g[3 * H + j]);
    float c = fg * c_in[i] + ig * gg;
    float h = og * tanhf(c);
    c_out[i] = c;
    h_out[i] = h;
    if (h_relu) h_relu[i] = fmaxf(h, 0.f);
  }
}

// The whole recurrence of a multi-window call in ONE launch: per window only gates = g_in[w] + h W_hh^T, the cell and the
// 1-output head depend on the previous window, and streams are independent -- so a persistent CTA owns SB streams and
// walks the W windows with h and c resident in shared memory (state in -> W windows -> state out), instead of three
// launches per 32 ms window over a grid that 4096 streams cannot fill.  thread = gate column n (4H = 512 threads): the
// recurrent product runs in exact fp32 with W_hh^T read k-row by k-row (coalesced, L2-resident: 256 KB) and h broadcast
// from shared memory as float4; then (stream, unit) items update the cell, and each warp reduces the heads of its streams.
template <int H, int SB>
__global__ void __launch_bounds__(4 * H, 1) silero_lstm_windows_kernel(
    const float* __restrict__ g_in, const float* __restrict__ wt_hh, int ldw, const float* __restrict__ state_in,
    float* __restrict__ state_out, const float* __restrict__ head_w, float head_b, float* __restrict__ probs, int64_t S, int W) {
  extern __shared__ float lw_sm[];
  float* h_s = lw_sm;                 // [SB][H]
  float* c_s = h_s + SB * H;          // [SB][H]
  float* gt = c_s + SB * H;           // [SB][4H]
  const int n = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t s0 = (int64_t)blockIdx.x * SB;
  const int nb = (int)min((int64_t)SB, S - s0);
  for (int i = n; i < SB * H; i += 4 * H) {
    const int s = i / H, j = i - s * H;
    h_s[i] = s < nb ? state_in[(s0 + s) * H + j] : 0.f;
    c_s[i] = s < nb ? state_in[(S + s0 + s) * H + j] : 0.f;
  }
  float hw[H / 32];
#pragma unroll
  for (int i = 0; i < H / 32; ++i) hw[i] = __ldg(head_w + lane + 32 * i);
  const float4* h4 = reinterpret_cast<const float4*>(h_s);
  for (int w = 0; w < W; ++w) {
    __syncthreads();
    float acc[SB];
#pragma unroll
    for (int s = 0; s < SB; ++s) acc[s] = 0.f;
#pragma unroll 2
    for (int k4 = 0; k4 < H / 4; ++k4) {
      const float w0 = __ldg(wt_hh + (size_t)(4 * k4 + 0) * ldw + n), w1 = __ldg(wt_hh + (size_t)(4 * k4 + 1) * ldw + n);
      const float w2 = __ldg(wt_hh + (size_t)(4 * k4 + 2) * ldw + n), w3 = __ldg(wt_hh + (size_t)(4 * k4 + 3) * ldw + n);
#pragma unroll
      for (int s = 0; s < SB; ++s) {
        const float4 hv = h4[s * (H / 4) + k4];
        acc[s] = fmaf(hv.w, w3, fmaf(hv.z, w2, fmaf(hv.y, w1, fmaf(hv.x, w0, acc[s]))));
      }
    }
    const float* gi = g_in + ((int64_t)w * S + s0) * (4 * H) + n;
#pragma unroll
    for (int s = 0; s < SB; ++s)
      if (s < nb) gt[s * 4 * H + n] = acc[s] + gi[(int64_t)s * 4 * H];
    __syncthreads();
    for (int i = n; i < nb * H; i += 4 * H) {
      const int s = i / H, j = i - s * H;
      const float* g = gt + s * 4 * H;
      const float ig = sigmoidf_(g[j]), fg = sigmoidf_(g[H + j]), gg = tanhf(g[2 * H + j]), og = sigmoidf_(g[3 * H + j]);
      const float c = fg * c_s[i] + ig * gg;
      c_s[i] = c;
      h_s[i] = og * tanhf(c);
    }
    __syncthreads();
    for (int s = warp; s < nb; s += 4 * H / 32) {
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < H / 32; ++i) part = fmaf(fmaxf(h_s[s * H + lane + 32 * i], 0.f), hw[i], part);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
      if (lane == 0) probs[(int64_t)w * S + s0 + s] = sigmoidf_(part + head_b);
    }
  }
  __syncthreads();
  for (int i = n; i < nb * H; i += 4 * H) {
    const int s = i / H, j = i - s * H;
    state_out[(s0 + s) * H + j] = h_s[i];
    state_out[(S + s0 + s) * H + j] = c_s[i];
  }
}

// a16: the trigger / release / max-speech machine of get_speech_timestamps
// (Silero/modeling_modified/utils_vad.py:374-462), one stream per lane.  Emits raw (start, end)
// sample pairs; padding and rounding (:464-482) are per-segment host work.
struct SileroTsCfg {
  double threshold, neg_threshold, min_speech_samples, max_speech_samples, min_silence_samples,
      min_silence_samples_at_max_speech;
  int window, use_max_poss_sil;
};
__global__ void __launch_bounds__(128) silero_timestamps_kernel(const float* __restrict__ probs, int64_t ld,
                                                                const int32_t* __restrict__ n_windows,
                                                                const int64_t* __restrict__ n_samples,
                                                                int64_t n_streams, SileroTsCfg cfg,
                                                                int32_t* __restrict__ seg_count,
                                                                int64_t* __restrict__ segments, int max_segments) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_streams) return;
  const float* p = probs + s * ld;
  const int n = n_windows[s];
  int64_t* seg = segments + s * (int64_t)max_segments * 2;
  int count = 0;
  auto emit = [&](int64_t a, int64_t b) {
    if (count < max_segments) { seg[2 * count] = a; seg[2 * count + 1] = b; }
    ++count;
  };
  bool triggered = false, have_cur = false;
  int64_t cur_start = 0, temp_end = 0, prev_end = 0, next_start = 0;
  bool have_best = false;            // possible_ends: only its first-longest entry is ever used
  int64_t best_end = 0, best_dur = 0;
  for (int i = 0; i < n; ++i) {
    const double sp = (double)p[i];
    const int64_t cur = (int64_t)cfg.window * i;
    if (sp >= cfg.threshold && temp_end) {
      int64_t sil = cur - temp_end;
      if ((double)sil > cfg.min_silence_samples_at_max_speech) {
        if (!have_best || sil > best_dur) { have_best = true; best_end = temp_end; best_dur = sil; }
      }
      temp_end = 0;
      if (next_start < prev_end) next_start = cur;
    }
    if (sp >= cfg.threshold && !triggered) {
      triggered = true;
      cur_start = cur;
      have_cur = true;
      continue;
    }
    if (triggered && (double)(cur - cur_start) > cfg.max_speech_samples) {
      if (cfg.use_max_poss_sil && have_best) {
        prev_end = best_end;
        emit(cur_start, prev_end);
        have_cur = false;
        next_start = prev_end + best_dur;
        if (next_start < prev_end + cur) { cur_start = next_start; have_cur = true; }
        else triggered = false;
        prev_end = next_start = temp_end = 0;
        have_best = false;
      } else {
        if (prev_end) {
          emit(cur_start, prev_end);
          have_cur = false;
          if (next_start < prev_end) triggered = false;
          else { cur_start = next_start; have_cur = true; }
          prev_end = next_start = temp_end = 0;
          have_best = false;
        } else {
          emit(cur_start, cur);
          have_cur = false;
          prev_end = next_start = temp_end = 0;
          triggered = false;
          have_best = false;
          continue;
        }
      }
    }
    if (sp < cfg.neg_threshold && triggered) {
      if (!temp_end) temp_end = cur;
      int64_t sil_now = cur - temp_end;
      if (!cfg.use_max_poss_sil && (double)sil_now > cfg.min_silence_samples_at_max_speech) prev_end = temp_end;
      if ((double)sil_now < cfg.min_silence_samples) continue;
      if ((double)(temp_end - cur_start) > cfg.min_speech_samples) emit(cur_start, temp_end);
      have_cur = false;
      prev_end = next_start = temp_end = 0;
      triggered = false;
      have_best = false;
      continue;
    }
  }
  const int64_t len = n_samples[s];
  if (have_cur && (double)(len - cur_start) > cfg.min_speech_samples) emit(cur_start, len);
  seg_count[s] = count;
}

}  // namespace vadx

using namespace vadx;

static inline unsigned grid1d(int64_t items) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(items, 256), 148 * 16));
}

extern "C" int vadx_reflect_window_f32(const float* d_x, int64_t in_stride, int64_t n_streams, int n_in, int pad,
                                       float* d_out, void* stream) {
  StageTimer _timer(VADX_STAGE_PREP, (cudaStream_t)stream, "reflect_window_kernel", 4.0 * n_streams * (2.0 * n_in + pad));
  VADX_REQUIRE(d_x && d_out && n_streams >= 0 && n_in >= 2 && pad >= 0 && pad <= n_in - 1 && in_stride >= 1,
               "vadx_reflect_window_f32: bad argument");
  if (n_streams == 0) return VADX_OK;
  reflect_window_kernel<<<grid1d(n_streams * (n_in + pad)), 256, 0, (cudaStream_t)stream>>>(d_x, in_stride, n_streams, 1, 0,
                                                                                          n_in, pad, d_out);
  return after_launch("vadx_reflect_window_f32");
}

extern "C" int vadx_reflect_windows_f32(const float* d_x, int64_t in_stride, int64_t n_streams, int n_windows,
                                        int64_t window_step, int n_in, int pad, float* d_out, void* stream) {
  StageTimer _timer(VADX_STAGE_PREP, (cudaStream_t)stream, "reflect_window_kernel",
                    4.0 * n_streams * n_windows * (2.0 * n_in + pad));
  VADX_REQUIRE(d_x && d_out && n_streams >= 0 && n_windows >= 1 && window_step >= 0 && n_in >= 2 && pad >= 0 &&
                   pad <= n_in - 1 && in_stride >= 1,
               "vadx_reflect_windows_f32: bad argument");
  if (n_streams == 0) return VADX_OK;
  reflect_window_kernel<<<grid1d(n_streams * n_windows * (n_in + pad)), 256, 0, (cudaStream_t)stream>>>(
      d_x, in_stride, n_streams, n_windows, window_step, n_in, pad, d_out);
  return after_launch("vadx_reflect_windows_f32");
}

extern "C" int vadx_affine_f32(const float* d_x, float a, float b, float* d_y, int64_t n, void* stream) {
  StageTimer _timer(VADX_STAGE_HEAD, (cudaStream_t)stream, "affine_kernel", 8.0 * n);
  VADX_REQUIRE(d_x && d_y && n >= 0, "vadx_affine_f32: bad argument");
  if (n == 0) return VADX_OK;
  affine_kernel<<<grid1d(n), 256, 0, (cudaStream_t)stream>>>(d_x, a, b, d_y, n);
  return after_launch("vadx_affine_f32");
}

extern "C" int vadx_sqrt_inplace_f32(float* d_p, int64_t n, void* stream) {
  StageTimer _timer(VADX_STAGE_MEL, (cudaStream_t)stream, "sqrt_inplace_kernel", 8.0 * n);
  VADX_REQUIRE(d_p && n >= 0, "vadx_sqrt_inplace_f32: bad argument");
  if (n == 0) return VADX_OK;
  sqrt_kernel<<<grid1d(n), 256, 0, (cudaStream_t)stream>>>(d_p, n);
  return after_launch("vadx_sqrt_inplace_f32");
}

// interleaved (re, im) rows of the framed DFT -> magnitudes, dropping the junk rows between windows: the DFT ran as a
// dense layer over hop-strided rows, `per` rows per window of which the first T are frames (model_silero.cu)
namespace vadx {
__global__ void __launch_bounds__(256) stft_mag_compact_kernel(const float* __restrict__ y, int64_t ldy, int64_t n_windows,
                                                               int per, int T, int F, float* __restrict__ mag) {
  const int64_t total = n_windows * T * F;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ((int64_t)T * F);
    const int rem = (int)(i - r * (int64_t)T * F);
    const int t = rem / F, f = rem - t * F;
    const float2 v = *reinterpret_cast<const float2*>(y + (r * per + t) * ldy + 2 * f);
    mag[i] = sqrtf(fmaf(v.x, v.x, v.y * v.y));
  }
}
}  // namespace vadx

extern "C" int vadx_stft_mag_compact_f32(const float* d_y, int64_t ldy, int64_t n_windows, int rows_per_window, int n_frames,
                                         int n_bins, float* d_mag, void* stream) {
  StageTimer _timer(VADX_STAGE_MEL, (cudaStream_t)stream, "stft_mag_compact_kernel", 12.0 * n_windows * n_frames * n_bins);
  VADX_REQUIRE(d_y && d_mag && n_windows >= 0 && rows_per_window >= n_frames && n_frames >= 1 && n_bins >= 1 &&
                   ldy >= 2 * n_bins && (ldy & 1) == 0 && (reinterpret_cast<uintptr_t>(d_y) & 7u) == 0,
               "vadx_stft_mag_compact_f32: bad argument");
  if (n_windows == 0) return VADX_OK;
  stft_mag_compact_kernel<<<grid1d(n_windows * n_frames * n_bins), 256, 0, (cudaStream_t)stream>>>(d_y, ldy, n_windows,
                                                                                                 rows_per_window, n_frames, n_bins, d_mag);
  return after_launch("vadx_stft_mag_compact_f32");
}

extern "C" int vadx_lstm_cell_f32(const float* d_gates, const float* d_c_in, float* d_h_out, float* d_c_out,
                                  float* d_h_relu, int64_t n_streams, int hidden, void* stream) {
  StageTimer _timer(VADX_STAGE_MEMORY, (cudaStream_t)stream, "lstm_cell_kernel", 4.0 * n_streams * hidden * (4 + 1 + 2 + (d_h_relu ? 1 : 0)));
  VADX_REQUIRE(d_gates && d_c_in && d_h_out && d_c_out && n_streams >= 0 && hidden >= 1, "vadx_lstm_cell_f32: bad argument");
  if (n_streams == 0) return VADX_OK;
  lstm_cell_kernel<<<grid1d(n_streams * hidden), 256, 0, (cudaStream_t)stream>>>(d_gates, d_c_in, d_h_out, d_c_out,
                                                                                d_h_relu, n_streams, hidden);
  return after_launch("vadx_lstm_cell_f32");
}

extern "C" int vadx_silero_lstm_windows_f32(const float* d_gates_in, const float* d_wt_hh, int ldw, const float* d_state_in,
                                            float* d_state_out, const float* d_head_w, float head_bias, float* d_probs,
                                            int64_t n_streams, int n_windows, int hidden, void* stream) {
  StageTimer _timer(VADX_STAGE_MEMORY, (cudaStream_t)stream, "silero_lstm_windows_kernel",
                    4.0 * n_streams * n_windows * (4.0 * hidden + 1.0) + 16.0 * n_streams * hidden,
                    2.0 * n_streams * n_windows * 4.0 * hidden * hidden);
  VADX_REQUIRE(d_gates_in && d_wt_hh && d_state_in && d_state_out && d_head_w && d_probs && d_state_in != d_state_out,
               "vadx_silero_lstm_windows_f32: null or aliased pointer");
  VADX_REQUIRE(n_streams >= 0 && n_windows >= 1 && hidden == 128 && ldw >= 4 * hidden,
               "vadx_silero_lstm_windows_f32: hidden must be 128 (got %d) and ldw >= 4*hidden", hidden);
  if (n_streams == 0) return VADX_OK;
  constexpr int H = 128, SB = 28;            // 4096 streams -> 147 CTAs of 28 streams: one wave on 148 SMs
  const size_t smem = (size_t)SB * H * 6 * sizeof(float);
  static PerDevice per_device;
  VADX_TRY(per_device.ensure(nullptr, [] {
    return cudaFuncSetAttribute(silero_lstm_windows_kernel<H, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  }));
  silero_lstm_windows_kernel<H, SB><<<(unsigned)ceil_div(n_streams, SB), 4 * H, smem, (cudaStream_t)stream>>>(
      d_gates_in, d_wt_hh, ldw, d_state_in, d_state_out, d_head_w, head_bias, d_probs, n_streams, n_windows);
  return after_launch("vadx_silero_lstm_windows_f32");
}

extern "C" int vadx_silero_timestamps(const float* d_probs, int64_t ld, const int32_t* d_n_windows,
                                      const int64_t* d_n_samples, int64_t n_streams, double threshold,
                                      double neg_threshold, double min_speech_samples, double max_speech_samples,
                                      double min_silence_samples, double min_silence_samples_at_max_speech,
                                      int window, int use_max_poss_sil, int32_t* d_seg_count, int64_t* d_segments,
                                      int max_segments, void* stream) {
  StageTimer _timer(VADX_STAGE_POSTPROC, (cudaStream_t)stream, "silero_timestamps_kernel");
  VADX_REQUIRE(d_probs && d_n_windows && d_n_samples && d_seg_count && d_segments && max_segments >= 1 && window >= 1,
               "vadx_silero_timestamps: bad argument");
  if (n_streams == 0) return VADX_OK;
  SileroTsCfg c{threshold, neg_threshold, min_speech_samples, max_speech_samples, min_silence_samples,
                min_silence_samples_at_max_speech, window, use_max_poss_sil};
  silero_timestamps_kernel<<<(unsigned)ceil_div(n_streams, 128), 128, 0, (cudaStream_t)stream>>>(
      d_probs, ld, d_n_windows, d_n_samples, n_streams, c, d_seg_count, d_segments, max_segments);
  return after_launch("vadx_silero_timestamps");
}
