// ingest.cu -- real-audio ingest (next-3): stereo -> mono and rate conversion of int16 PCM on the device,
// bit-identical to audioop.tomono + audioop.ratecv (what pydub's set_channels(1).set_frame_rate(sr) runs,
// the reference loader: FSMN/Inference_FSMN_VAD_ONNX.py:68, FireRedVAD/Inference_FireRed_ONNX.py:535).
// ratecv is written as a sequential recurrence, but with its default weights every output depends on two
// neighbouring input frames and an integer phase that has a closed form, so outputs are independent:
// one thread per output sample, coalesced stores, inputs through the read-only path.
#include "common.cuh"

namespace vadx {

static int64_t gcd_i64(int64_t a, int64_t b) {
  while (b) { int64_t t = a % b; a = b; b = t; }
  return a;
}

__device__ __forceinline__ int64_t floor_div(int64_t n, int64_t d) {  // d > 0
  int64_t q = n / d;
  return (n % d != 0 && n < 0) ? q - 1 : q;
}

__global__ void __launch_bounds__(256) ingest_pcm16_kernel(const int16_t* __restrict__ pcm, int64_t in_stride,
                                                           const int64_t* __restrict__ n_in_per_stream,
                                                           int64_t n_frames_in, int n_channels, int64_t a, int64_t b,
                                                           int16_t* __restrict__ out, int64_t out_stride,
                                                           int64_t n_out_max, int64_t* __restrict__ n_out_per_stream) {
  const int64_t s = blockIdx.y;
  int64_t n_in = n_in_per_stream ? n_in_per_stream[s] : n_frames_in;
  n_in = n_in < 0 ? 0 : (n_in > n_frames_in ? n_frames_in : n_in);
  const int64_t n_out = n_in == 0 ? 0 : ((n_in - 1) * b) / a + 1;
  const int16_t* x = pcm + s * in_stride;
  auto mono = [&](int64_t i) -> int64_t {
    if (n_channels == 1) return (int64_t)__ldg(x + i);
    const int64_t l = __ldg(x + 2 * i), r = __ldg(x + 2 * i + 1);
    return (l + r) >> 1;                               // floor((l + r) / 2)
  };
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k == 0 && n_out_per_stream) n_out_per_stream[s] = n_out;
  if (k >= n_out_max) return;
  int16_t v = 0;
  if (k < n_out) {
    if (a == b) {
      v = (int16_t)mono(k);
    } else {
      const int64_t n = (k * a + b - 1) / b + 1;       // inputs consumed when output k is produced
      const int64_t d = (n - 1) * b - k * a;           // phase in [0, b)
      const int64_t cur = mono(n - 1);
      const int64_t prev = n >= 2 ? mono(n - 2) : 0;
      v = (int16_t)floor_div(prev * d + cur * (b - d), b);
    }
  }
  out[s * out_stride + k] = v;
}

// F.interpolate(mode='linear', align_corners=False) with a given scale_factor (ATen UpSampleLinear1d: the
// reciprocal scale is formed in double and cast to float, coordinates and weights are float)
__global__ void __launch_bounds__(256) resample_linear_kernel(const float* __restrict__ in, int64_t in_stride, int64_t n_in,
                                                              float rscale, float* __restrict__ out, int64_t out_stride,
                                                              int64_t out_offset, int64_t n_out) {
  const int64_t s = blockIdx.y;
  const float* x = in + s * in_stride;
  float* y = out + s * out_stride + out_offset;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (int64_t)gridDim.x * blockDim.x) {
    float src = __fsub_rn(__fmul_rn(rscale, __fadd_rn((float)i, 0.5f)), 0.5f);
    if (src < 0.f) src = 0.f;
    const int64_t i0 = (int64_t)src;
    const int64_t i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
    const float w1 = __fsub_rn(src, (float)i0), w0 = __fsub_rn(1.0f, w1);
    y[i] = __fadd_rn(__fmul_rn(w0, __ldg(x + i0)), __fmul_rn(w1, __ldg(x + i1)));
  }
}

}  // namespace vadx

using namespace vadx;

extern "C" int64_t vadx_resample_out_len(int64_t n_in, double scale) {
  if (n_in <= 0 || !(scale > 0.0)) return 0;
  return (int64_t)floor((double)n_in * scale);
}

extern "C" int vadx_resample_linear_f32(const float* d_in, int64_t in_stride, int64_t n_in, int64_t n_streams, double scale,
                                        float* d_out, int64_t out_stride, int64_t out_offset, void* stream) {
  StageTimer _timer(VADX_STAGE_PREP, (cudaStream_t)stream, "resample_linear_kernel", 4.0 * n_streams * n_in * (1.0 + scale));
  VADX_REQUIRE(d_in && d_out && scale > 0.0, "vadx_resample_linear_f32: bad argument");
  const int64_t n_out = vadx_resample_out_len(n_in, scale);
  VADX_REQUIRE(n_streams >= 0 && n_streams <= 65535 && n_in >= 1 && in_stride >= n_in && out_offset >= 0 &&
                   out_stride >= out_offset + n_out,
               "vadx_resample_linear_f32: bad shape");
  if (n_streams == 0 || n_out == 0) return VADX_OK;
  dim3 grid((unsigned)std::min<int64_t>(ceil_div(n_out, 256), 4096), (unsigned)n_streams);
  resample_linear_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_in, in_stride, n_in, (float)(1.0 / scale), d_out,
                                                                 out_stride, out_offset, n_out);
  return after_launch("vadx_resample_linear_f32");
}

extern "C" int64_t vadx_ingest_out_frames(int64_t n_frames_in, int in_rate, int out_rate) {
  if (n_frames_in <= 0 || in_rate <= 0 || out_rate <= 0) return 0;
  const int64_t g = gcd_i64(in_rate, out_rate);
  const int64_t a = in_rate / g, b = out_rate / g;
  return ((n_frames_in - 1) * b) / a + 1;
}

extern "C" int vadx_ingest_pcm16(const int16_t* d_pcm, int64_t in_stride, const int64_t* d_n_in, int64_t n_streams,
                                 int64_t n_frames_in, int n_channels, int in_rate, int out_rate, int16_t* d_out,
                                 int64_t out_stride, int64_t* d_n_out, void* stream) {
  StageTimer _timer(VADX_STAGE_PREP, (cudaStream_t)stream, "ingest_pcm16_kernel", 2.0 * n_streams * n_frames_in * n_channels * (1.0 + (double)out_rate / (in_rate * n_channels)));
  VADX_REQUIRE(d_pcm && d_out, "vadx_ingest_pcm16: null pointer");
  VADX_REQUIRE(n_channels == 1 || n_channels == 2, "vadx_ingest_pcm16: %d channels are not supported", n_channels);
  VADX_REQUIRE(in_rate > 0 && out_rate > 0 && out_rate <= 65536 * (int)gcd_i64(in_rate, out_rate),
               "vadx_ingest_pcm16: bad rates %d -> %d", in_rate, out_rate);
  VADX_REQUIRE(n_streams >= 0 && n_streams <= 65535 && n_frames_in >= 0 && in_stride >= n_frames_in * n_channels,
               "vadx_ingest_pcm16: bad shape S=%lld n=%lld stride=%lld", (long long)n_streams, (long long)n_frames_in,
               (long long)in_stride);
  const int64_t n_out_max = vadx_ingest_out_frames(n_frames_in, in_rate, out_rate);
  VADX_REQUIRE(out_stride >= n_out_max, "vadx_ingest_pcm16: out_stride %lld < %lld output frames", (long long)out_stride,
               (long long)n_out_max);
  if (n_streams == 0 || (n_out_max == 0 && !d_n_out)) return VADX_OK;
  const int64_t g = gcd_i64(in_rate, out_rate);
  dim3 grid((unsigned)std::max<int64_t>(1, ceil_div(n_out_max, 256)), (unsigned)n_streams);
  ingest_pcm16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_pcm, in_stride, d_n_in, n_frames_in, n_channels,
                                                              in_rate / g, out_rate / g, d_out, out_stride, n_out_max,
                                                              d_n_out);
  return after_launch("vadx_ingest_pcm16");
}
