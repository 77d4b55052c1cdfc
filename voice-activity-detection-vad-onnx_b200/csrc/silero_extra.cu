// silero_extra.cu -- Silero-specific stages (a11, a16): reflect-padded window assembly, spectral
// magnitude, LSTM cell update, and the get_speech_timestamps trigger machine on the device.
#include "common.cuh"

namespace vadx {

// out[s][j] = x[s][j] (j < n_in), out[s][n_in + j] = x[s][n_in - 2 - j] (j < pad): right reflect pad
__global__ void __launch_bounds__(256) reflect_window_kernel(const float* __restrict__ x, int64_t in_stride,
                                                             int64_t n_streams, int n_in, int pad,
                                                             float* __restrict__ out) {
  const int n_out = n_in + pad;
  const int64_t total = n_streams * n_out;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t s = i / n_out;
    int j = (int)(i - s * n_out);
    int src = j < n_in ? j : n_in - 2 - (j - n_in);
    out[i] = x[s * in_stride + src];
  }
}

__global__ void __launch_bounds__(256) sqrt_kernel(float* __restrict__ p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = sqrtf(p[i]);
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
  reflect_window_kernel<<<grid1d(n_streams * (n_in + pad)), 256, 0, (cudaStream_t)stream>>>(d_x, in_stride, n_streams,
                                                                                          n_in, pad, d_out);
  return after_launch("vadx_reflect_window_f32");
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
