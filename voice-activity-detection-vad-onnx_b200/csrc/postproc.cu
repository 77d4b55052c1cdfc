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

// One warp = 32 streams.  The decision arrays of the warp's streams live in shared memory
// (byte-interleaved by lane: dec(t) of lane l at sdec[t*32 + l], conflict-free), so the back-fill,
// merge, dilation and split passes never touch global memory; probabilities are pulled through the
// read-only path, decisions are written back once at the end.
__global__ void __launch_bounds__(32) postprocess_frames_kernel(const float* __restrict__ probs, int64_t ld_probs,
                                                                const int32_t* __restrict__ n_frames_per_stream,
                                                                int64_t n_streams, int n_frames_max,
                                                                const vadx_post_cfg cfg, int8_t* __restrict__ dec_all,
                                                                int32_t* __restrict__ seg_count,
                                                                int32_t* __restrict__ segments, int max_segments,
                                                                int use_smem) {
  extern __shared__ int8_t sdec[];
  const int lane = threadIdx.x;
  const int64_t s = (int64_t)blockIdx.x * 32 + lane;
  if (s >= n_streams) return;
  int n = n_frames_per_stream ? n_frames_per_stream[s] : n_frames_max;
  if (n > n_frames_max) n = n_frames_max;
  if (n < 0) n = 0;
  const float* p = probs + s * ld_probs;
  int8_t* gdec = dec_all + s * (int64_t)n_frames_max;
  // dec(t): shared (stride 32, offset lane) or global (stride 1)
  int8_t* dec = use_smem ? sdec + lane : gdec;
  const int ds = use_smem ? 32 : 1;
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
    const float pt = __ldg(p + t);
    if (ws > 1) {
      run = __fadd_rn(run, pt);             // cumsum[t+1]
      ring[(t + 1) % (ws + 1)] = run;
      if (t < ws - 1) sm = __fdiv_rn(run, (float)(t + 1));
      else sm = __fmul_rn(__fsub_rn(run, ring[(t + 1 - ws) % (ws + 1)]), inv_ws);
    } else {
      sm = pt;
    }
    const bool hot = sm >= thr;
    if (plain) {
      dec[t * ds] = hot ? 1 : 0;
      continue;
    }
    if (state == 0) {
      if (hot) { state = 1; speech_start = t; }
    } else if (state == 1) {
      if (hot) {
        if (t - speech_start >= min_sp) {
          state = 2;
          for (int j = speech_start; j < t; ++j) dec[j * ds] = 1;
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
    dec[t * ds] = state >= 2 ? 1 : 0;
  }

  // ---- rising edges move left by ws (:235-243) ----
  if (ws > 1) {
    for (int t = 1; t < n; ++t) {
      if (dec[t * ds] == 1 && dec[(t - 1) * ds] == 0) {
        int start = t >= ws ? t - ws : 0;
        for (int j = start; j < t; ++j) dec[j * ds] = 1;
      }
    }
  }
  // ---- short gaps are filled (:245-257) ----
  if (cfg.merge_silence_frame > 0) {
    int gap = -1;
    for (int t = 1; t < n; ++t) {
      int a = dec[(t - 1) * ds], b = dec[t * ds];
      if (a == 1 && b == 0 && gap < 0) {
        gap = t;
      } else if (a == 0 && b == 1 && gap >= 0) {
        if (t - gap < cfg.merge_silence_frame)
          for (int j = gap; j < t; ++j) dec[j * ds] = 1;
        gap = -1;
      }
    }
  }
  // ---- dilation (:259-277) ----
  if (cfg.extend_speech_frame > 0) {
    const int ext = cfg.extend_speech_frame;
    int dist = ext + 1;
    for (int t = 0; t < n; ++t) {
      if (dec[t * ds]) dist = 0;
      else if (++dist <= ext) dec[t * ds] = 1;
    }
    dist = ext + 1;
    for (int t = n - 1; t >= 0; --t) {
      if (dec[t * ds]) dist = 0;
      else if (++dist <= ext) dec[t * ds] = 1;
    }
  }
  // ---- split over-long runs at the least likely frame of the back half (:279-304) ----
  {
    const int max_sf = cfg.max_speech_frame, half = cfg.max_speech_frame >> 1;
    int t = 0;
    while (t < n) {
      if (!dec[t * ds]) { ++t; continue; }
      int seg_start = t;
      while (t < n && dec[t * ds]) ++t;
      if (t - seg_start > max_sf) {
        int pos = seg_start;
        const int seg_end = t;
        while (pos + max_sf < seg_end) {
          int a = pos + half, b = pos + max_sf;
          if (b > seg_end) b = seg_end;
          if (a >= b) break;
          int arg = a;
          float best = __ldg(p + a);
          for (int j = a + 1; j < b; ++j) {
            float v = __ldg(p + j);
            if (v < best) { best = v; arg = j; }
          }
          dec[arg * ds] = 0;
          pos = arg + 1;
        }
      }
    }
  }
  // ---- edges -> (start, end) frame pairs (decision_to_segment :146-166) + decisions out ----
  int count = 0;
  int32_t* seg = segments ? segments + s * (int64_t)max_segments * 2 : nullptr;
  int prev = 0, start = 0;
  for (int t = 0; t <= n; ++t) {
    int cur = t < n ? dec[t * ds] : 0;
    if (use_smem && t < n) gdec[t] = (int8_t)cur;
    if (cur && !prev) start = t;
    if (!cur && prev) {
      if (seg && count < max_segments) { seg[2 * count] = start; seg[2 * count + 1] = t; }
      ++count;
    }
    prev = cur;
  }
  if (seg_count) seg_count[s] = count;
}


// Warp-per-stream variant (the default whenever a stream's frames fit in shared memory): lanes load
// the probabilities coalesced, lane 0 forms the float32 running sum in order (the only truly
// sequential float arithmetic), ALL lanes evaluate the smoothed value and the threshold test for
// their frames in parallel from that running sum (each is one subtraction + one multiply of
// already-rounded operands, so the result is bit-identical to numpy's), lane 0 then walks the
// integer state machines over shared memory, and all lanes write the decisions back coalesced.
// One stream per warp means no lane divergence in the sequential part.
constexpr int kPpWarps = 4;
__global__ void __launch_bounds__(kPpWarps * 32) postprocess_frames_warp_kernel(
    const float* __restrict__ probs, int64_t ld_probs, const int32_t* __restrict__ n_frames_per_stream,
    int64_t n_streams, int n_frames_max, const vadx_post_cfg cfg, int8_t* __restrict__ dec_all,
    int32_t* __restrict__ seg_count, int32_t* __restrict__ segments, int max_segments, int per_warp_floats) {
  extern __shared__ float smem_f[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t s = (int64_t)blockIdx.x * kPpWarps + warp;
  if (s >= n_streams) return;
  int n = n_frames_per_stream ? n_frames_per_stream[s] : n_frames_max;
  if (n > n_frames_max) n = n_frames_max;
  if (n < 0) n = 0;
  float* sp = smem_f + (size_t)warp * per_warp_floats;   // probs [n_frames_max]
  float* cs = sp + n_frames_max;                          // running sums [n_frames_max + 1]
  int8_t* dec = reinterpret_cast<int8_t*>(cs + n_frames_max + 1);
  const float* p = probs + s * ld_probs;
  int8_t* gdec = dec_all + s * (int64_t)n_frames_max;
  const int ws = cfg.smooth_window < 1 ? 1 : cfg.smooth_window;
  const float thr = cfg.threshold;
  const int min_sp = cfg.min_speech_frame, min_si = cfg.min_silence_frame;
  const float inv_ws = (float)(1.0 / (double)ws);

  for (int t = lane; t < n; t += 32) sp[t] = __ldg(p + t);
  __syncwarp();
  if (ws > 1) {
    if (lane == 0) {
      float run = 0.f;
      cs[0] = 0.f;
      for (int t = 0; t < n; ++t) {
        run = __fadd_rn(run, sp[t]);
        cs[t + 1] = run;
      }
    }
    __syncwarp();
  }
  for (int t = lane; t < n; t += 32) {
    float sm;
    if (ws > 1) {
      if (t < ws - 1) sm = __fdiv_rn(cs[t + 1], (float)(t + 1));
      else sm = __fmul_rn(__fsub_rn(cs[t + 1], cs[t + 1 - ws]), inv_ws);
    } else {
      sm = sp[t];
    }
    dec[t] = sm >= thr ? 1 : 0;   // "hot" flags; turned into decisions in place below
  }
  __syncwarp();
  int count = 0;
  if (lane == 0) {
    if (!(min_sp <= 0 && min_si <= 0)) {
      int state = 0, speech_start = 0, silence_start = 0;
      for (int t = 0; t < n; ++t) {
        const bool hot = dec[t] != 0;
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
    }
    if (ws > 1) {
      for (int t = 1; t < n; ++t) {
        if (dec[t] == 1 && dec[t - 1] == 0) {
          int start = t >= ws ? t - ws : 0;
          for (int j = start; j < t; ++j) dec[j] = 1;
        }
      }
    }
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
            float best = sp[a];
            for (int j = a + 1; j < b; ++j)
              if (sp[j] < best) { best = sp[j]; arg = j; }
            dec[arg] = 0;
            pos = arg + 1;
          }
        }
      }
    }
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
  __syncwarp();
  for (int t = lane; t < n; t += 32) gdec[t] = dec[t];
}

// ---- run-based variant of the warp-per-stream kernel ----
// Same results, different cost model: after the (inherently sequential) float32 running sum, the frame
// flags become a bit mask (__ballot_sync) and the state machines advance from run to run with
// find-first-set instead of frame by frame.  The 4-state machine, the window extension of rising edges,
// the gap merge and the dilation are closed-form on (start, end) pairs:
//   * a speech segment opens at the first frame a of a run of flags that stays set through frame
//     a + max(min_speech, 1) (the reference back-fills to a, :196-204);
//   * it closes at e = b + max(min_silence, 1), b the first cleared frame after which the flags stay
//     clear through e (possible-silence frames count as speech, :209-221), or at the end of the stream;
//   * every segment that does not start at frame 0 grows left by the smoothing window (:235-243),
//     gaps shorter than merge_silence close (:245-257), extend_speech dilates both ways (:259-277);
//   * max_speech splitting at the first minimum of the raw probabilities in the second half of the
//     window (:279-304) is done while the final pairs are emitted, the arg-min by all lanes.
// Every lane executes the same control flow on the same shared data (no divergence), lane 0 writes.
__device__ __forceinline__ int pp_next_set(const uint32_t* w, int t, int n) {
  if (t >= n) return n;
  const int nw = (n + 31) >> 5;
  int wi = t >> 5;
  uint32_t cur = w[wi] & (0xffffffffu << (t & 31));
  while (!cur) {
    if (++wi >= nw) return n;
    cur = w[wi];
  }
  const int r = (wi << 5) + __ffs(cur) - 1;
  return r < n ? r : n;
}
__device__ __forceinline__ int pp_next_clear(const uint32_t* w, int t, int n) {
  if (t >= n) return n;
  const int nw = (n + 31) >> 5;
  int wi = t >> 5;
  uint32_t cur = ~w[wi] & (0xffffffffu << (t & 31));
  while (!cur) {
    if (++wi >= nw) return n;
    cur = ~w[wi];
  }
  const int r = (wi << 5) + __ffs(cur) - 1;
  return r < n ? r : n;
}

__global__ void __launch_bounds__(kPpWarps * 32) postprocess_frames_runs_kernel(
    const float* __restrict__ probs, int64_t ld_probs, const int32_t* __restrict__ n_frames_per_stream,
    int64_t n_streams, int n_frames_max, const vadx_post_cfg cfg, int8_t* __restrict__ dec_all,
    int32_t* __restrict__ seg_count, int32_t* __restrict__ segments, int max_segments, int per_warp_floats) {
  extern __shared__ float smem_f[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t s = (int64_t)blockIdx.x * kPpWarps + warp;
  if (s >= n_streams) return;
  int n = n_frames_per_stream ? n_frames_per_stream[s] : n_frames_max;
  if (n > n_frames_max) n = n_frames_max;
  if (n < 0) n = 0;
  float* sp = smem_f + (size_t)warp * per_warp_floats;     // probs [n_frames_max]
  float* cs = sp + n_frames_max;                            // running sums [n_frames_max + 1], then the pairs
  int* pairs = reinterpret_cast<int*>(cs);                  // (start, end) x up to (n + 1) / 2
  int8_t* dec = reinterpret_cast<int8_t*>(cs + n_frames_max + 1);
  uint32_t* hot = reinterpret_cast<uint32_t*>(dec + ((n_frames_max + 3) & ~3));
  const float* p = probs + s * ld_probs;
  int8_t* gdec = dec_all + s * (int64_t)n_frames_max;
  const int ws = cfg.smooth_window < 1 ? 1 : cfg.smooth_window;
  const float thr = cfg.threshold;
  const int min_sp = cfg.min_speech_frame, min_si = cfg.min_silence_frame;
  const float inv_ws = (float)(1.0 / (double)ws);

  for (int t = lane; t < n; t += 32) sp[t] = __ldg(p + t);
  __syncwarp();
  if (ws > 1) {
    if (lane == 0) {
      float run = 0.f;
      cs[0] = 0.f;
      int t = 0;
      for (; t + 4 <= n; t += 4) {   // four independent loads per trip; the additions stay in order
        const float a = sp[t], b = sp[t + 1], c = sp[t + 2], d = sp[t + 3];
        run = __fadd_rn(run, a); cs[t + 1] = run;
        run = __fadd_rn(run, b); cs[t + 2] = run;
        run = __fadd_rn(run, c); cs[t + 3] = run;
        run = __fadd_rn(run, d); cs[t + 4] = run;
      }
      for (; t < n; ++t) { run = __fadd_rn(run, sp[t]); cs[t + 1] = run; }
    }
    __syncwarp();
  }
  for (int t0 = 0; t0 < n; t0 += 32) {
    const int t = t0 + lane;
    bool h = false;
    if (t < n) {
      float sm;
      if (ws > 1) {
        if (t < ws - 1) sm = __fdiv_rn(cs[t + 1], (float)(t + 1));
        else sm = __fmul_rn(__fsub_rn(cs[t + 1], cs[t + 1 - ws]), inv_ws);
      } else {
        sm = sp[t];
      }
      h = sm >= thr;
    }
    const uint32_t word = __ballot_sync(0xffffffffu, h);
    if (lane == 0) hot[t0 >> 5] = word;
    if (t < n) dec[t] = 0;
  }
  __syncwarp();   // cs is dead from here: its storage becomes the pair list

  // ---- pass A: flags -> speech segments (all lanes, same control flow) ----
  int ns = 0;
  auto push = [&](int a, int e) {
    if (lane == 0) { pairs[2 * ns] = a; pairs[2 * ns + 1] = e; }
    ++ns;
  };
  if (min_sp <= 0 && min_si <= 0) {
    int t = 0;
    while (t < n) {
      const int a = pp_next_set(hot, t, n);
      if (a >= n) break;
      const int e = pp_next_clear(hot, a + 1, n);
      push(a, e);
      t = e;
    }
  } else {
    const int need_sp = min_sp > 1 ? min_sp : 1, need_si = min_si > 1 ? min_si : 1;
    int t = 0;
    while (t < n) {
      const int a = pp_next_set(hot, t, n);
      if (a >= n) break;
      const int b = pp_next_clear(hot, a + 1, n);
      if (a + need_sp >= b) { t = b; continue; }          // the run ends before speech is confirmed
      int cur = b, e;
      while (true) {
        if (cur >= n) { e = n; break; }
        const int c = pp_next_set(hot, cur + 1, n);
        const int lim = cur + need_si;
        if (lim < c && lim < n) { e = lim; break; }        // silence confirmed at frame lim
        if (c >= n) { e = n; break; }                      // the stream ends inside possible silence
        cur = pp_next_clear(hot, c + 1, n);
      }
      push(a, e);
      t = e;
    }
  }
  __syncwarp();
  // ---- passes B-D on the pair list, in place (each pass only ever shrinks the list) ----
  auto sweep = [&](int grow_left, int grow_left_skip0, int grow_right, int merge_below) {
    // grow every pair, then fuse pairs that touch/overlap or whose gap is < merge_below
    int m = 0, pa = 0, pe = 0;
    for (int k = 0; k < ns; ++k) {
      int a = pairs[2 * k], e = pairs[2 * k + 1];
      if (!(grow_left_skip0 && a == 0)) a = a - grow_left < 0 ? 0 : a - grow_left;
      e = e + grow_right > n ? n : e + grow_right;
      if (m > 0 && (a <= pe || a - pe < merge_below)) {
        if (e > pe) pe = e;
      } else {
        if (m > 0) {
          __syncwarp();
          if (lane == 0) { pairs[2 * (m - 1)] = pa; pairs[2 * (m - 1) + 1] = pe; }
        }
        pa = a; pe = e; ++m;
      }
    }
    __syncwarp();
    if (m > 0 && lane == 0) { pairs[2 * (m - 1)] = pa; pairs[2 * (m - 1) + 1] = pe; }
    __syncwarp();
    ns = m;
  };
  if (ws > 1) sweep(ws, 1, 0, 0);
  if (cfg.merge_silence_frame > 0) sweep(0, 0, 0, cfg.merge_silence_frame);
  if (cfg.extend_speech_frame > 0) sweep(cfg.extend_speech_frame, 0, cfg.extend_speech_frame, 0);

  // ---- pass E + emission: split long segments, write pairs, rasterise the decisions ----
  int count = 0;
  int32_t* seg = segments ? segments + s * (int64_t)max_segments * 2 : nullptr;
  auto emit = [&](int a, int e) {
    if (e <= a) return;
    if (lane == 0 && seg && count < max_segments) { seg[2 * count] = a; seg[2 * count + 1] = e; }
    ++count;
    for (int t = a + lane; t < e; t += 32) dec[t] = 1;
  };
  const int max_sf = cfg.max_speech_frame, half = cfg.max_speech_frame >> 1;
  for (int k = 0; k < ns; ++k) {
    const int a0 = pairs[2 * k], e0 = pairs[2 * k + 1];
    int pos = a0;
    if (e0 - a0 > max_sf) {
      while (pos + max_sf < e0) {
        const int lo = pos + half, hi = pos + max_sf;     // hi < e0
        if (lo >= hi) break;
        float best = INFINITY;
        int arg = 0x7fffffff;
        for (int j = lo + lane; j < hi; j += 32) {
          const float v = sp[j];
          if (v < best) { best = v; arg = j; }            // strided scan keeps the lowest index per lane
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, best, off);
          const int oa = __shfl_xor_sync(0xffffffffu, arg, off);
          if (ov < best || (ov == best && oa < arg)) { best = ov; arg = oa; }
        }
        emit(pos, arg);
        pos = arg + 1;
      }
    }
    emit(pos, e0);
  }
  if (lane == 0 && seg_count) seg_count[s] = count;
  __syncwarp();
  for (int t = lane; t < n; t += 32) gdec[t] = dec[t];
}

// ---- Stream-VAD segmenter (FireRedVAD/Inference_FireRed_ONNX.py:307-490) ----
// state words per stream: 0 head, 1 fill, 2 frame (1-based, last consumed), 3 mode, 4 speech run, 5 silence run,
// 6 re-arm flag, 7 first frame of the open segment (0 = none), 8 last frame of the latest closed segment
// (0 = none), 9 ring sum (float bits), 10..15 spare, 16.. ring of smooth_window floats.
constexpr int kSpHeader = 16;

__global__ void __launch_bounds__(128) stream_post_kernel(const float* __restrict__ probs, int64_t ld_probs,
                                                          const int32_t* __restrict__ n_frames_per_stream,
                                                          int64_t n_streams, int n_frames_max,
                                                          const vadx_stream_post_cfg cfg, int32_t* __restrict__ state,
                                                          int32_t* __restrict__ seg_count,
                                                          int32_t* __restrict__ segments, int max_segments,
                                                          int32_t* __restrict__ open_seg) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_streams) return;
  int n = n_frames_per_stream ? n_frames_per_stream[s] : n_frames_max;
  n = n < 0 ? 0 : (n > n_frames_max ? n_frames_max : n);
  const int ws = cfg.smooth_window < 1 ? 1 : cfg.smooth_window;
  const int pad = cfg.pad_start_frame > ws ? cfg.pad_start_frame : ws;
  int32_t* st = state + s * (int64_t)(kSpHeader + ws);
  float ring[kMaxSmooth];
  for (int i = 0; i < ws; ++i) ring[i] = __int_as_float(st[kSpHeader + i]);
  int head = st[0], fill = st[1], frame = st[2], mode = st[3], n_sp = st[4], n_si = st[5];
  bool rearm = st[6] != 0;
  int open_from = st[7], closed_at = st[8];
  float acc = __int_as_float(st[9]);
  int count = seg_count ? seg_count[s] : 0;
  int32_t* seg = segments ? segments + s * (int64_t)max_segments * 2 : nullptr;
  const float* p = probs + s * ld_probs;
  for (int t = 0; t < n; ++t) {
    const float pt = __ldg(p + t);
    ++frame;
    float sm = pt;
    if (ws > 1) {
      const float gone = ring[head];
      ring[head] = pt;
      acc = __fadd_rn(acc, __fsub_rn(pt, gone));
      head = head + 1 == ws ? 0 : head + 1;
      if (fill < ws) ++fill;
      sm = __fdiv_rn(acc, (float)fill);
    }
    const bool speech = sm >= cfg.threshold;
    int begin = 0, end = 0;  // 1-based, 0 = not set
    bool force_close = false;
    if (rearm) { begin = open_from = frame; rearm = false; }
    if (mode == 0) {
      if (speech) { mode = 1; n_sp = 1; }
      else { ++n_si; n_sp = 0; }
    } else if (mode == 1) {
      if (speech) {
        if (++n_sp >= cfg.min_speech_frame) {
          mode = 2;
          int b = frame - n_sp + 1 - pad;
          if (b < 1) b = 1;
          if (b < closed_at + 1) b = closed_at + 1;
          begin = open_from = b;
          n_si = 0;
        }
      } else { mode = 0; n_si = 1; n_sp = 0; }
    } else if (mode == 2) {
      ++n_sp;
      if (speech) { n_si = 0; force_close = n_sp >= cfg.max_speech_frame; }
      else { mode = 3; n_si = 1; }
    } else {
      ++n_sp;
      if (speech) { mode = 2; n_si = 0; force_close = n_sp >= cfg.max_speech_frame; }
      else if (++n_si >= cfg.min_silence_frame) {
        mode = 0;
        begin = open_from; end = frame;
        open_from = 0; closed_at = frame; n_sp = 0;
      }
    }
    if (force_close) {
      rearm = true;
      n_sp = 0;
      begin = open_from; end = frame;
      open_from = 0; closed_at = frame;
    }
    if (begin > 0 && end > 0) {
      if (seg && count < max_segments) { seg[2 * count] = begin - 1; seg[2 * count + 1] = end - 1; }
      ++count;
    }
  }
  st[0] = head; st[1] = fill; st[2] = frame; st[3] = mode; st[4] = n_sp; st[5] = n_si; st[6] = rearm ? 1 : 0;
  st[7] = open_from; st[8] = closed_at; st[9] = __float_as_int(acc);
  for (int i = 0; i < ws; ++i) st[kSpHeader + i] = __float_as_int(ring[i]);
  if (seg_count) seg_count[s] = count;
  if (open_seg) {
    open_seg[2 * s] = open_from > 0 ? open_from - 1 : -1;
    open_seg[2 * s + 1] = open_from > 0 ? frame - 1 : -1;
  }
}

}  // namespace vadx

using namespace vadx;

extern "C" int vadx_stream_post_state_words(int smooth_window) {
  return kSpHeader + (smooth_window < 1 ? 1 : smooth_window);
}

extern "C" int vadx_stream_postprocess(const float* d_probs, int64_t ld_probs, const int32_t* d_n_frames,
                                       int64_t n_streams, int n_frames, const vadx_stream_post_cfg* cfg,
                                       int32_t* d_state, int32_t* d_seg_count, int32_t* d_segments, int max_segments,
                                       int32_t* d_open, void* stream) {
  StageTimer _timer(VADX_STAGE_POSTPROC, (cudaStream_t)stream, "stream_postprocess_kernel", 4.0 * n_streams * n_frames);
  VADX_REQUIRE(d_probs && cfg && d_state, "vadx_stream_postprocess: null pointer");
  VADX_REQUIRE(n_streams >= 0 && n_frames >= 0 && ld_probs >= n_frames, "vadx_stream_postprocess: bad shape");
  VADX_REQUIRE(cfg->smooth_window <= kMaxSmooth, "vadx_stream_postprocess: smooth_window %d > %d", cfg->smooth_window,
               kMaxSmooth);
  VADX_REQUIRE(max_segments >= 0 && (max_segments == 0 || (d_segments && d_seg_count)),
               "vadx_stream_postprocess: segments buffer without counts");
  if (n_streams == 0) return VADX_OK;
  stream_post_kernel<<<(unsigned)ceil_div(n_streams, 128), 128, 0, (cudaStream_t)stream>>>(
      d_probs, ld_probs, d_n_frames, n_streams, n_frames, *cfg, d_state, d_seg_count, d_segments, max_segments, d_open);
  return after_launch("vadx_stream_postprocess");
}

extern "C" int vadx_postprocess_frames(const float* d_probs, int64_t ld_probs, const int32_t* d_n_frames,
                                       int64_t n_streams, int n_frames, const vadx_post_cfg* cfg,
                                       int8_t* d_decisions, int32_t* d_seg_count, int32_t* d_segments,
                                       int max_segments, void* stream) {
  StageTimer _timer(VADX_STAGE_POSTPROC, (cudaStream_t)stream, "postprocess_frames_runs_kernel", 5.0 * n_streams * n_frames);
  VADX_REQUIRE(d_probs && cfg && d_decisions, "vadx_postprocess_frames: null pointer (d_decisions is required)");
  VADX_REQUIRE(n_streams >= 0 && n_frames >= 0 && ld_probs >= n_frames, "vadx_postprocess_frames: bad shape");
  VADX_REQUIRE(cfg->smooth_window <= kMaxSmooth, "vadx_postprocess_frames: smooth_window %d > %d", cfg->smooth_window,
               kMaxSmooth);
  VADX_REQUIRE(max_segments >= 0 && (max_segments == 0 || d_segments), "vadx_postprocess_frames: segments buffer");
  if (n_streams == 0) return VADX_OK;
  // warp-per-stream kernel when one stream (probs + running sums + decisions) fits in shared memory
  const int per_warp_floats = (int)round_up(2 * (int64_t)std::max(n_frames, 1) + 1 + (std::max(n_frames, 1) + 3) / 4 +
                                               (std::max(n_frames, 1) + 31) / 32 + 1, 4);
  const size_t smem_w = (size_t)per_warp_floats * 4 * kPpWarps;
  if (smem_w <= 200 * 1024) {
    static PerDevice per_device_w;
    VADX_TRY(per_device_w.ensure(nullptr, [] {
      cudaError_t e = cudaFuncSetAttribute(postprocess_frames_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(postprocess_frames_runs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      return e;
    }));
    // VADX_PP_LEGACY=1 (AB build only): the frame-by-frame kernel, kept as the cross-check of the run-based one
    static const int legacy = ab_env("VADX_PP_LEGACY", 0);
    if (legacy)
      postprocess_frames_warp_kernel<<<(unsigned)ceil_div(n_streams, kPpWarps), kPpWarps * 32, smem_w, (cudaStream_t)stream>>>(
          d_probs, ld_probs, d_n_frames, n_streams, n_frames, *cfg, d_decisions, d_seg_count, d_segments, max_segments,
          per_warp_floats);
    else
      postprocess_frames_runs_kernel<<<(unsigned)ceil_div(n_streams, kPpWarps), kPpWarps * 32, smem_w, (cudaStream_t)stream>>>(
          d_probs, ld_probs, d_n_frames, n_streams, n_frames, *cfg, d_decisions, d_seg_count, d_segments, max_segments,
          per_warp_floats);
    return after_launch("vadx_postprocess_frames");
  }
  const int64_t blocks = ceil_div(n_streams, 32);
  const size_t smem = (size_t)std::max(n_frames, 1) * 32;
  const int use_smem = smem <= 200 * 1024;
  static PerDevice per_device;
  VADX_TRY(per_device.ensure(nullptr, [] {
    return cudaFuncSetAttribute(postprocess_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }));
  postprocess_frames_kernel<<<(unsigned)blocks, 32, use_smem ? smem : 0, (cudaStream_t)stream>>>(
      d_probs, ld_probs, d_n_frames, n_streams, n_frames, *cfg, d_decisions, d_seg_count, d_segments, max_segments,
      use_smem);
  return after_launch("vadx_postprocess_frames");
}
