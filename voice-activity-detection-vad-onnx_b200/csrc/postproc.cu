// postproc.cu -- on-device post-processing state machines (a15).
//
// One stream per lane, strictly sequential in time: the reference smooths with a float32
// np.cumsum (a sequential running sum) and thresholds the result, so a tree scan would flip
// decisions at ties.  Every arithmetic step below reproduces one numpy float32 operation of
// FireRedVAD/Inference_FireRed_ONNX.py:181-304 (MarbleNet copy:
// NVIDIA_*/Inference_NVIDIA_MarbleNet_VAD_ONNX.py:160-353).
#include "common.cuh"

namespace vadx {

constexpr int kMaxSmooth = 64;

__global__ void __launch_bounds__(128) postprocess_frames_kernel(const float* __restrict__ probs, int64_t ld_probs,
                                                                 const int32_t* __restrict__ n_frames_per_stream,
                                                                 int64_t n_streams, int n_frames_max,
                                                                 const vadx_post_cfg cfg, int8_t* __restrict__ dec_all,
                                                                 int32_t* __restrict__ seg_count,
                                                                 int32_t* __restrict__ segments, int max_segments) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_streams) return;
  int n = n_frames_per_stream ? n_frames_per_stream[s] : n_frames_max;
  if (n > n_frames_max) n = n_frames_max;
  if (n < 0) n = 0;
  const float* p = probs + s * ld_probs;
  int8_t* dec = dec_all + s * (int64_t)n_frames_max;
  const int ws = cfg.smooth_window < 1 ? 1 : cfg.smooth_window;
  const float thr = cfg.threshold;
  const int min_sp = cfg.min_speech_frame, min_si = cfg.min_silence_frame;
  const float inv_ws = (float)(1.0 / (double)ws);

  // ---- smoothing + threshold + 4-state machine (:181-233) ----
  float ring[kMaxSmooth + 1];  // running sums cumsum[i+1-ws .. i+1]
  float run = 0.f;
  ring[0] = 0.f;
  const bool plain = (min_sp <= 0 && min_si <= 0);
  int state = 0, speech_start = 0, silence_start = 0;  // 0 SIL, 1 POSSIBLE_SPEECH, 2 SPEECH, 3 POSSIBLE_SILENCE
  for (int t = 0; t < n; ++t) {
    float sm;
    if (ws > 1) {
      run = __fadd_rn(run, p[t]);           // cumsum[t+1]
      ring[(t + 1) % (ws + 1)] = run;
      if (t < ws - 1) sm = __fdiv_rn(run, (float)(t + 1));
      else sm = __fmul_rn(__fsub_rn(run, ring[(t + 1 - ws) % (ws + 1)]), inv_ws);
    } else {
      sm = p[t];
    }
    const bool hot = sm >= thr;
    if (plain) {
      dec[t] = hot ? 1 : 0;
      continue;
    }
    if (state == 0) {
      if (hot) { state = 1; speech_start = t; }
    } else if (state == 1) {
      if (hot) {
        if (t - speech_start >= min_sp) {
          state = 2;
          for (int j = speech_start; j < t; ++j) dec[j] = 1;
        }
      } else {
        state = 0;
      }
    } else if (state == 2) {
      if (!hot) { state = 3; silence_start = t; }
    } else {
      if (!hot) {
        if (t - silence_start >= min_si) state = 0;
      } else {
        state = 2;
      }
    }
    dec[t] = state >= 2 ? 1 : 0;
  }

  // ---- rising edges move left by ws (:235-243) ----
  if (ws > 1) {
    for (int t = 1; t < n; ++t) {
      if (dec[t] == 1 && dec[t - 1] == 0) {
        int start = t >= ws ? t - ws : 0;
        for (int j = start; j < t; ++j) dec[j] = 1;
      }
    }
  }
  // ---- short gaps are filled (:245-257) ----
  if (cfg.merge_silence_frame > 0) {
    int gap = -1;
    for (int t = 1; t < n; ++t) {
      int a = dec[t - 1], b = dec[t];
      if (a == 1 && b == 0 && gap < 0) {
        gap = t;
      } else if (a == 0 && b == 1 && gap >= 0) {
        if (t - gap < cfg.merge_silence_frame)
          for (int j = gap; j < t; ++j) dec[j] = 1;
        gap = -1;
      }
    }
  }
  // ---- dilation (:259-277) ----
  if (cfg.extend_speech_frame > 0) {
    const int ext = cfg.extend_speech_frame;
    int dist = ext + 1;
    for (int t = 0; t < n; ++t) {
      if (dec[t]) dist = 0;
      else if (++dist <= ext) dec[t] = 1;
    }
    dist = ext + 1;
    for (int t = n - 1; t >= 0; --t) {
      if (dec[t]) dist = 0;
      else if (++dist <= ext) dec[t] = 1;
    }
  }
  // ---- split over-long runs at the least likely frame of the back half (:279-304) ----
  {
    const int max_sf = cfg.max_speech_frame, half = cfg.max_speech_frame >> 1;
    int t = 0;
    while (t < n) {
      if (!dec[t]) { ++t; continue; }
      int seg_start = t;
      while (t < n && dec[t]) ++t;
      if (t - seg_start > max_sf) {
        int pos = seg_start;
        const int seg_end = t;
        while (pos + max_sf < seg_end) {
          int a = pos + half, b = pos + max_sf;
          if (b > seg_end) b = seg_end;
          if (a >= b) break;
          int arg = a;
          float best = p[a];
          for (int j = a + 1; j < b; ++j)
            if (p[j] < best) { best = p[j]; arg = j; }
          dec[arg] = 0;
          pos = arg + 1;
        }
      }
    }
  }
  // ---- edges -> (start, end) frame pairs (decision_to_segment :146-166) ----
  int count = 0;
  int32_t* seg = segments ? segments + s * (int64_t)max_segments * 2 : nullptr;
  int prev = 0, start = 0;
  for (int t = 0; t <= n; ++t) {
    int cur = t < n ? dec[t] : 0;
    if (cur && !prev) start = t;
    if (!cur && prev) {
      if (seg && count < max_segments) { seg[2 * count] = start; seg[2 * count + 1] = t; }
      ++count;
    }
    prev = cur;
  }
  if (seg_count) seg_count[s] = count;
}

}  // namespace vadx

using namespace vadx;

extern "C" int vadx_postprocess_frames(const float* d_probs, int64_t ld_probs, const int32_t* d_n_frames,
                                       int64_t n_streams, int n_frames, const vadx_post_cfg* cfg,
                                       int8_t* d_decisions, int32_t* d_seg_count, int32_t* d_segments,
                                       int max_segments, void* stream) {
  StageTimer _timer(VADX_STAGE_POSTPROC, (cudaStream_t)stream);
  VADX_REQUIRE(d_probs && cfg && d_decisions, "vadx_postprocess_frames: null pointer (d_decisions is required)");
  VADX_REQUIRE(n_streams >= 0 && n_frames >= 0 && ld_probs >= n_frames, "vadx_postprocess_frames: bad shape");
  VADX_REQUIRE(cfg->smooth_window <= kMaxSmooth, "vadx_postprocess_frames: smooth_window %d > %d", cfg->smooth_window,
               kMaxSmooth);
  VADX_REQUIRE(max_segments >= 0 && (max_segments == 0 || d_segments), "vadx_postprocess_frames: segments buffer");
  if (n_streams == 0) return VADX_OK;
  int64_t blocks = ceil_div(n_streams, 128);
  postprocess_frames_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(
      d_probs, ld_probs, d_n_frames, n_streams, n_frames, *cfg, d_decisions, d_seg_count, d_segments, max_segments);
  return after_launch("vadx_postprocess_frames");
}
