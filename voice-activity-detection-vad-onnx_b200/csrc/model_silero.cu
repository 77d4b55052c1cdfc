// model_silero.cu -- Silero VAD v5 (16 kHz) step: the graph behind OnnxWrapper.__call__
// (Silero/modeling_modified/utils_vad.py:114-123) for S streams at once.
//   hparams  {window(512), context(64), reflect_pad(64), n_fft(256), hop(128), hidden(128), n_enc_layers,
//             [out_dim, in_dim] x n_enc_layers}   (dense re-expression of the k=3 conv stack)
//   inputs   [0] x fp32 [S][row stride]  (context + window = 576 valid samples per row; the row stride is
//                the scalar "input.row_stride", default 576, so windows can be strided views of one
//                long zero-prefixed signal)
//   outputs  [0] out fp32 [S][1]
//   state    [0] state in fp32 [2][S][128] (h ; c)   [1] state out (distinct buffer)
#include "model.hpp"

extern "C" {
int vadx_reflect_window_f32(const float*, int64_t, int64_t, int, int, float*, void*);
int vadx_reflect_windows_f32(const float*, int64_t, int64_t, int, int64_t, int, int, float*, void*);
int vadx_sqrt_inplace_f32(float*, int64_t, void*);
int vadx_lstm_cell_f32(const float*, const float*, float*, float*, float*, int64_t, int, void*);
int vadx_stft_mag_compact_f32(const float*, int64_t, int64_t, int, int, int, float*, void*);
int vadx_silero_lstm_windows_f32(const float*, const float*, int, const float*, float*, const float*, float, float*, int64_t, int,
                                 int, void*);
}

namespace {
struct SileroHP {
  int window, context, reflect, n_fft, hop, hidden, n_layers;
  std::vector<std::pair<int, int>> dims;  // (out, in)
  int n_in() const { return window + context; }
  int n_bins() const { return n_fft / 2 + 1; }
  int n_frames() const { return (n_in() + reflect - n_fft) / hop + 1; }
  int ld_basis() const { return (int)round_up(2 * n_bins(), 4); }
};
int silero_hp(const vadx_model* m, SileroHP* h) {
  const auto& v = m->hp;
  VADX_REQUIRE(v.size() >= 7 && (int)v.size() == 7 + 2 * v[6], "silero: malformed hyper-parameter list (%zu)", v.size());
  *h = SileroHP{v[0], v[1], v[2], v[3], v[4], v[5], v[6], {}};
  for (int i = 0; i < h->n_layers; ++i) h->dims.push_back({v[7 + 2 * i], v[8 + 2 * i]});
  VADX_REQUIRE(h->n_layers >= 1 && h->dims[0].second == h->n_frames() * h->n_bins(),
               "silero: first dense layer must take n_frames*n_bins = %d inputs", h->n_frames() * h->n_bins());
  for (int i = 1; i < h->n_layers; ++i)
    VADX_REQUIRE(h->dims[i].second == h->dims[i - 1].first, "silero: dense layer %d input mismatch", i);
  VADX_REQUIRE(h->dims.back().first == h->hidden || true, "silero: encoder output");
  return VADX_OK;
}
}  // namespace

int silero_check(const vadx_model* m) {
  SileroHP h;
  return silero_hp(m, &h);
}
int silero_frames(const vadx_model* m, int64_t, int32_t* out) {
  (void)m;
  *out = 1;
  return VADX_OK;
}
int silero_finalize(vadx_model* m) {
  SileroHP h;
  VADX_TRY(silero_hp(m, &h));
  VADX_TRY(m->upload_raw("frontend.basis", (int64_t)h.n_fft * h.ld_basis(), VADX_DT_F32));
  {
    // the framed DFT as a dense tensor-core layer over the hop-strided frames of the padded window: W[2f | 2f+1][k] =
    // basis[k][re f | im f] (column blocks, 2F = 258 > 256)
    const float* hb = m->find("frontend.basis")->f32();
    HostTensor t;
    t.dtype = VADX_DT_F32;
    t.dims = {2 * h.n_bins(), h.n_fft};
    t.bytes.resize((size_t)2 * h.n_bins() * h.n_fft * sizeof(float));
    float* w = reinterpret_cast<float*>(t.bytes.data());
    for (int j = 0; j < 2 * h.n_bins(); ++j)
      for (int k = 0; k < h.n_fft; ++k) w[(size_t)j * h.n_fft + k] = hb[(size_t)k * h.ld_basis() + j];
    m->host["frontend.stft.weight"] = std::move(t);
    VADX_TRY(m->upload_linear("frontend.stft.weight", 2 * h.n_bins(), h.n_fft));
  }
  for (int i = 0; i < h.n_layers; ++i) {
    std::string p = "enc." + std::to_string(i) + ".";
    VADX_TRY(m->upload_linear(p + "weight", h.dims[i].first, h.dims[i].second));
    VADX_TRY(m->upload_raw(p + "bias", h.dims[i].first, VADX_DT_F32));
  }
  const int enc_out = h.dims.back().first;
  VADX_TRY(m->upload_linear("rnn.weight_ih", 4 * h.hidden, enc_out));
  VADX_TRY(m->upload_linear("rnn.weight_hh", 4 * h.hidden, h.hidden));
  VADX_TRY(m->upload_raw("rnn.bias", 4 * h.hidden, VADX_DT_F32));
  VADX_TRY(m->upload_linear("head.weight", 1, h.hidden));
  VADX_TRY(m->upload_raw("head.weight", h.hidden, VADX_DT_F32));
  VADX_TRY(m->upload_raw("head.bias", 1, VADX_DT_F32));
  return VADX_OK;
}

// "input.n_windows" = W > 1 (offline / batched use): in[0] rows are strided views of a long signal and the call
// covers W consecutive windows of every stream.  Everything that does not depend on the LSTM state (reflect pad,
// STFT, encoder, the input half of the gates) runs ONCE over S*W rows; only the recurrent half of the gates, the
// cell and the head run per window (3 launches instead of 10).  out[0] is then [W][S][1], state in -> state out
// spans all W windows.
int silero_run(vadx_model* m, bool dry, const void* const* in, void* const* out, void* const* state, int64_t S,
               int64_t L, void* ws_ptr, size_t ws_bytes, size_t* need, cudaStream_t st) {
  SileroHP h;
  VADX_TRY(silero_hp(m, &h));
  const int W = std::max(1, (int)m->scalar("input.n_windows", 1.0));
  const int64_t rows = S * W;
  const int nwin = h.n_in() + h.reflect;
  const int T = h.n_frames(), F = h.n_bins();
  int wide = 4 * h.hidden;
  for (auto& d : h.dims) wide = std::max(wide, d.first);  // only layer OUTPUTS live in the ping-pong buffers
  const int ldw = (int)round_up(wide, 4);
  Workspace ws(ws_ptr, ws_bytes, dry);
  float* win = ws.take<float>(rows * nwin);
  float* mag = ws.take<float>(rows * T * F + 4);
  // tensor-core DFT: `per` hop-strided rows per window (the first T are its frames), (re, im) interleaved
  const int per = (nwin % h.hop) == 0 ? nwin / h.hop : 0;
  const int ld_ri = (int)round_up(2 * F, 4);
  const bool stft_tc = per >= T && (h.hop % 8) == 0 && rows * (int64_t)T > kSkinnyMaxRows;
  float* ri = stft_tc ? ws.take<float>(((rows - 1) * per + T) * (int64_t)ld_ri) : nullptr;
  float* bufA = ws.take<float>(rows * ldw);
  float* bufB = ws.take<float>(rows * ldw);
  float* gates = ws.take<float>(S * ldw);
  float* hrelu = ws.take<float>(S * h.hidden);
  float* tmp_state[2] = {W > 1 ? ws.take<float>(2 * S * h.hidden) : nullptr, W > 2 ? ws.take<float>(2 * S * h.hidden) : nullptr};
  if (need) *need = ws.off;
  if (dry) return VADX_OK;
  if (ws.off > ws_bytes) {
    set_error("silero: workspace of %zu bytes is smaller than the %zu needed", ws_bytes, ws.off);
    return VADX_ENOMEM;
  }
  VADX_REQUIRE(L == h.n_in(), "silero: input rows must hold context + window = %d samples, got %lld", h.n_in(),
               (long long)L);
  VADX_REQUIRE(state && state[0] && state[1] && state[0] != state[1], "silero: state in/out buffers are required");
  const int64_t in_stride = (int64_t)m->scalar("input.row_stride", (double)h.n_in());
  VADX_REQUIRE(in_stride >= 1 && (W == 1 || in_stride >= (int64_t)(W - 1) * h.window + h.n_in()),
               "silero: input.row_stride %lld does not hold %d windows", (long long)in_stride, W);
  const float* st_in = static_cast<const float*>(state[0]);
  float* st_out = static_cast<float*>(state[1]);
  const bool tc_ok = m->scalar("engine.use_tc", 1.0) != 0.0;
  auto lin = [&](int64_t n_rows, const float* x, int64_t ldx, int n_in, const std::string& w, const char* b, const float* res,
                 int64_t ldr, float* y, int64_t ldy, int n_out, int act) -> int {
    return m->linear(w, x, ldx, b ? m->d<float>(b) : nullptr, res, ldr, y, ldy, n_rows, n_in, n_out, act, tc_ok, st);
  };
  // rows are ordered [window][stream] so that one window's gates are contiguous for the recurrence
  VADX_TRY(vadx_reflect_windows_f32(static_cast<const float*>(in[0]), in_stride, S, W, h.window, h.n_in(), h.reflect, win, st));
  if (stft_tc && tc_ok && m->scalar("derived.nsplit.frontend.stft.weight", 0.0) > 0) {
    // frame (r, t) = row r*per + t of the [.][n_fft] matrix with row stride hop over `win`; the rows in between are junk
    VADX_TRY(m->linear("frontend.stft.weight", win, h.hop, nullptr, nullptr, 0, ri, ld_ri, (rows - 1) * per + T, h.n_fft, 2 * F,
                       VADX_ACT_NONE, true, st));
    VADX_TRY(vadx_stft_mag_compact_f32(ri, ld_ri, rows, per, T, F, mag, st));
  } else {
    VADX_TRY(vadx_stft_power_f32(win, nwin, rows, T, h.hop, h.n_fft, m->d<float>("frontend.basis"), h.ld_basis(), F, mag, F,
                                 st));
    VADX_TRY(vadx_sqrt_inplace_f32(mag, rows * (int64_t)T * F, st));
  }
  const float* cur = mag;
  int64_t ldc = (int64_t)T * F;
  float* pp[2] = {bufA, bufB};
  for (int i = 0; i < h.n_layers; ++i) {
    std::string p = "enc." + std::to_string(i) + ".";
    std::string b = p + "bias";
    float* dst = pp[i & 1];
    VADX_TRY(lin(rows, cur, ldc, h.dims[i].second, p + "weight", b.c_str(), nullptr, 0, dst, ldw, h.dims[i].first,
                 VADX_ACT_RELU));
    cur = dst;
    ldc = ldw;
  }
  float* g1 = pp[h.n_layers & 1];
  VADX_TRY(lin(rows, cur, ldc, h.dims.back().first, "rnn.weight_ih", "rnn.bias", nullptr, 0, g1, ldw, 4 * h.hidden,
               VADX_ACT_NONE));
  // gates rows have stride ldw; the cell kernel wants them dense [S][4H]: ldw == 4H whenever 4H is the widest layer
  VADX_REQUIRE(ldw == 4 * h.hidden, "silero: 4*hidden must be the widest layer (got ld %d)", ldw);
  float* probs = static_cast<float*>(out[0]);
  if (W > 1 && h.hidden == 128 && S > kSkinnyMaxRows && m->scalar("engine.fuse_recurrence", 1.0) != 0.0) {
    // the whole recurrence in one launch (silero_extra.cu): state in -> W windows -> state out
    const HostTensor* hb = m->find("head.bias");
    return vadx_silero_lstm_windows_f32(g1, m->d<float>("rnn.weight_hh#T"), ldw, st_in, st_out, m->d<float>("head.weight"),
                                        hb->f32()[0], probs, S, W, h.hidden, st);
  }
  for (int w = 0; w < W; ++w) {
    const float* src = w == 0 ? st_in : tmp_state[(w - 1) & 1];
    float* dst = w == W - 1 ? st_out : tmp_state[w & 1];
    VADX_TRY(lin(S, src, h.hidden, h.hidden, "rnn.weight_hh", nullptr, g1 + (int64_t)w * S * ldw, ldw, gates, ldw, 4 * h.hidden,
                 VADX_ACT_NONE));
    VADX_TRY(vadx_lstm_cell_f32(gates, src + S * h.hidden, dst, dst + S * h.hidden, hrelu, S, h.hidden, st));
    VADX_TRY(linear_narrow(hrelu, h.hidden, m->d<float>("head.weight#T"), 4, m->d<float>("head.bias"), probs + (int64_t)w * S, S,
                           h.hidden, 1, VADX_ACT_SIGMOID, 1, 1, 1, st));
  }
  return VADX_OK;
}
