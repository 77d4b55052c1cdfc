// iccrn.cu -- building blocks of the SDAEC / ICCRN echo estimator (a10) on [stream][frame][bin][channel]
// activations: 4-D permute, (C,F)-LayerNorm with unbiased std, batched LSTM sequences, gated
// element-wise ops, 3-tap frequency im2col, the AlphaPredictor scaling and the ISTFT overlap-add.
// Reference: DFSMN/near_and_far_end_audio/Export_DFSMN_VAD.py:65-354.
#include "tc_ptx.cuh"

namespace vadx {

// ---------------------------------------------------------------------------------- permute
// out dims = (n[p0], n[p1], n[p2], n[p3]); out[i0][i1][i2][i3] = in[j0][j1][j2][j3] with j[p_k] = i_k
__global__ void __launch_bounds__(256) permute4_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                       int64_t n0, int64_t n1, int64_t n2, int64_t n3, int p0, int p1,
                                                       int p2, int p3) {
  const int64_t n[4] = {n0, n1, n2, n3};
  const int64_t st[4] = {n1 * n2 * n3, n2 * n3, n3, 1};
  const int p[4] = {p0, p1, p2, p3};
  const int64_t o1 = n[p[1]], o2 = n[p[2]], o3 = n[p[3]];
  const int64_t total = n0 * n1 * n2 * n3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int64_t i3 = r % o3; r /= o3;
    const int64_t i2 = r % o2; r /= o2;
    const int64_t i1 = r % o1; r /= o1;
    const int64_t i0 = r;
    out[i] = in[i0 * st[p[0]] + i1 * st[p[1]] + i2 * st[p[2]] + i3 * st[p[3]]];
  }
}

// ---------------------------------------------------------------------------------- LayerNorm (unbiased std)
// one CTA per row of D contiguous elements: (x - mean) / (std + eps) * w[d] + b[d]
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int64_t n_rows, int D,
                                                        const float* __restrict__ w, const float* __restrict__ b,
                                                        float eps, float* __restrict__ out) {
  __shared__ float red[256];
  __shared__ float s_mean, s_inv;
  const int64_t row = blockIdx.x;
  const float* xr = x + row * D;
  float acc = 0.f;
  for (int i = threadIdx.x; i < D; i += blockDim.x) acc += xr[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) s_mean = red[0] / (float)D;
  __syncthreads();
  const float mean = s_mean;
  acc = 0.f;
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    float d = xr[i] - mean;
    acc = fmaf(d, d, acc);
  }
  __syncthreads();
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) s_inv = 1.0f / (sqrtf(red[0] / (float)(D - 1)) + eps);
  __syncthreads();
  const float inv = s_inv;
  float* o = out + row * D;
  for (int i = threadIdx.x; i < D; i += blockDim.x) o[i] = (xr[i] - mean) * inv * w[i] + b[i];
}

// LayerNorm with an in-row permutation folded in: the row is a [n1][n2][n3] block, the output is written in the order
// (q0, q1, q2) of those axes (identity = plain LayerNorm); the affine tables are indexed in input order or, with
// w_out_order, in output order.  The row is staged in shared memory once (the plain kernel reads it three times), so a
// transposing permute4 launch before or after a LayerNorm costs nothing.  D = n1*n2*n3 <= 12288.
__global__ void __launch_bounds__(256) layernorm_perm_kernel(const float* __restrict__ x, int n1, int n2, int n3, int q0, int q1,
                                                             int q2, const float* __restrict__ w, const float* __restrict__ b,
                                                             int w_out_order, float eps, float* __restrict__ out, int64_t out_stride,
                                                             int pad) {
  extern __shared__ float ln_row[];
  __shared__ float red[256];
  __shared__ float s_mean, s_inv;
  const int D = n1 * n2 * n3;
  const int64_t row = blockIdx.x;
  const float* xr = x + row * D;
  float acc = 0.f;
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    const float v = xr[i];
    ln_row[i] = v;
    acc += v;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) s_mean = red[0] / (float)D;
  __syncthreads();
  const float mean = s_mean;
  acc = 0.f;
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    const float d = ln_row[i] - mean;
    acc = fmaf(d, d, acc);
  }
  __syncthreads();
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) s_inv = 1.0f / (sqrtf(red[0] / (float)(D - 1)) + eps);
  __syncthreads();
  const float inv = s_inv;
  const int n[3] = {n1, n2, n3};
  const int st[3] = {n2 * n3, n3, 1};
  const int o1 = n[q1], o2 = n[q2];
  const int s0 = st[q0], s1 = st[q1], s2 = st[q2];
  // out_stride >= D + 2*pad floats per row: `pad` zeros are written before and after the row (the zero padding a following
  // 3-tap frequency conv reads when it runs as a dense layer over overlapping rows)
  float* o = out + row * out_stride + pad;
  for (int i = threadIdx.x; i < pad; i += blockDim.x) {
    o[i - pad] = 0.f;
    o[D + i] = 0.f;
  }
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    const int i2 = i % o2, r = i / o2;
    const int i1 = r % o1, i0 = r / o1;
    const int src = i0 * s0 + i1 * s1 + i2 * s2;
    const int wi = w_out_order ? i : src;
    o[i] = (ln_row[src] - mean) * inv * w[wi] + b[wi];
  }
}

// ---------------------------------------------------------------------------------- gated-block front, fused
// The front of a gated conv block (CFB, Export_DFSMN_VAD.py:87-93,133-154) in ONE kernel, one CTA per (stream, frame) block
// b of F bins:   g = sigmoid(W_g LN0(x) + b_g),  xi = W_i x + b_i,  gx = g * xi,  d = xi - gx,
//                col = LN1(gx) as padded rows [F + 2][C] (the overlapping-row input of the (3,1) frequency conv),
//                z   = LN2(d) transposed to [C][F] (the input of the cepstral DFT layer).
// The unfused sequence is six launches (LN0, two dense layers, the gating, LN1, LN2) that move ~1300 bytes per (bin) row; here
// the block is read once (CIN floats per row) and two C-wide rows are written.  thread = bin f: its x row and LN0 row live in
// registers, the two C x CIN weights are broadcast from shared memory, gx / d go to odd-stride shared-memory rows for the two
// block statistics.  LayerNorm = (v - mean) / (unbiased std + eps) * w + b over the whole [F][.] block, as layernorm_kernel.
template <int CIN, int C>
__global__ void __launch_bounds__(160, 4) cfb_front_kernel(const float* __restrict__ x, int F, const float* __restrict__ ln0_w,
                                                          const float* __restrict__ ln0_b, const float* __restrict__ wg,
                                                          const float* __restrict__ bg, const float* __restrict__ wi,
                                                          const float* __restrict__ bi, const float* __restrict__ ln1_w,
                                                          const float* __restrict__ ln1_b, const float* __restrict__ ln2_w,
                                                          const float* __restrict__ ln2_b, float eps, float* __restrict__ col,
                                                          float* __restrict__ z) {
  constexpr int kT = 160;                 // threads = max bins per block
  constexpr int LD = C + 1;               // odd row stride: a thread's private row is conflict-free
  extern __shared__ __align__(16) float cf_sm[];
  float* wgs = cf_sm;                     // [C][CIN]
  float* wis = wgs + C * CIN;             // [C][CIN]
  float* gxs = wis + C * CIN;             // [F][LD]
  float* ds = gxs + kT * LD;              // [F][LD]
  __shared__ float red[4 * 16];
  const int tid = threadIdx.x;
  const int64_t b = blockIdx.x;
  for (int i = tid; i < C * CIN; i += kT) {
    wgs[i] = wg[i];
    wis[i] = wi[i];
  }
  // block sums of up to two values at once: warp shuffles, one partial per warp in a slot of its own (no slot is reused, so
  // one barrier per reduction)
  const int lane = tid & 31, warp = tid >> 5;
  auto block_sum2 = [&](float va, float vb, int slot, float& ra, float& rb) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      va += __shfl_xor_sync(0xffffffffu, va, off);
      vb += __shfl_xor_sync(0xffffffffu, vb, off);
    }
    if (lane == 0) {
      red[slot * 16 + warp] = va;
      red[slot * 16 + 8 + warp] = vb;
    }
    __syncthreads();
    ra = rb = 0.f;
#pragma unroll
    for (int wq = 0; wq < kT / 32; ++wq) {
      ra += red[slot * 16 + wq];
      rb += red[slot * 16 + 8 + wq];
    }
  };
  // ---- LN0 statistics over the [F][CIN] block; the thread keeps ITS row
  const bool live = tid < F;
  float xr[CIN];
  const float4* xp = reinterpret_cast<const float4*>(x + (b * F + (live ? tid : 0)) * CIN);
  float part = 0.f;
#pragma unroll
  for (int k4 = 0; k4 < CIN / 4; ++k4) {
    const float4 v = live ? __ldg(xp + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
    xr[4 * k4] = v.x; xr[4 * k4 + 1] = v.y; xr[4 * k4 + 2] = v.z; xr[4 * k4 + 3] = v.w;
    part += (v.x + v.y) + (v.z + v.w);
  }
  const float D0 = (float)(F * CIN);
  float r0, r1;
  block_sum2(part, 0.f, 0, r0, r1);
  const float mean0 = r0 / D0;
  part = 0.f;
  if (live) {
#pragma unroll
    for (int k = 0; k < CIN; ++k) {
      const float dlt = xr[k] - mean0;
      part = fmaf(dlt, dlt, part);
    }
  }
  block_sum2(part, 0.f, 1, r0, r1);
  const float inv0 = 1.0f / (sqrtf(r0 / (D0 - 1.0f)) + eps);
  // ---- the two C x CIN products of this bin, gating.  Two passes over ONE register copy of the row (the input product on
  // the raw row first, its results parked in the thread's shared-memory row; then the row is normalised in place for the gate
  // product): 40 instead of 80 row registers, one more resident CTA per SM
  float sum_gx = 0.f, sum_d = 0.f;
  if (live) {
#pragma unroll 2
    for (int o = 0; o < C; ++o) {
      const float4* i4 = reinterpret_cast<const float4*>(wis + o * CIN);
      float ai0 = __ldg(bi + o), ai1 = 0.f;
#pragma unroll
      for (int k4 = 0; k4 < CIN / 4; ++k4) {
        const float4 c4 = i4[k4];
        ai0 = fmaf(c4.y, xr[4 * k4 + 1], fmaf(c4.x, xr[4 * k4], ai0));
        ai1 = fmaf(c4.w, xr[4 * k4 + 3], fmaf(c4.z, xr[4 * k4 + 2], ai1));
      }
      ds[tid * LD + o] = ai0 + ai1;
    }
    const float4* w0 = reinterpret_cast<const float4*>(ln0_w + (size_t)tid * CIN);
    const float4* b0 = reinterpret_cast<const float4*>(ln0_b + (size_t)tid * CIN);
#pragma unroll
    for (int k4 = 0; k4 < CIN / 4; ++k4) {
      const float4 wv = __ldg(w0 + k4), bv = __ldg(b0 + k4);
      xr[4 * k4] = (xr[4 * k4] - mean0) * inv0 * wv.x + bv.x;
      xr[4 * k4 + 1] = (xr[4 * k4 + 1] - mean0) * inv0 * wv.y + bv.y;
      xr[4 * k4 + 2] = (xr[4 * k4 + 2] - mean0) * inv0 * wv.z + bv.z;
      xr[4 * k4 + 3] = (xr[4 * k4 + 3] - mean0) * inv0 * wv.w + bv.w;
    }
#pragma unroll 2
    for (int o = 0; o < C; ++o) {
      const float4* g4 = reinterpret_cast<const float4*>(wgs + o * CIN);
      float ag0 = __ldg(bg + o), ag1 = 0.f;
#pragma unroll
      for (int k4 = 0; k4 < CIN / 4; ++k4) {
        const float4 a = g4[k4];
        ag0 = fmaf(a.y, xr[4 * k4 + 1], fmaf(a.x, xr[4 * k4], ag0));
        ag1 = fmaf(a.w, xr[4 * k4 + 3], fmaf(a.z, xr[4 * k4 + 2], ag1));
      }
      const float g = 1.0f / (1.0f + expf(-(ag0 + ag1)));
      const float xi = ds[tid * LD + o];
      const float gx = g * xi, dd = xi - gx;
      gxs[tid * LD + o] = gx;
      ds[tid * LD + o] = dd;
      sum_gx += gx;
      sum_d += dd;
    }
  }
  // ---- LN1 / LN2 statistics over the [F][C] blocks
  const float D1 = (float)(F * C);
  block_sum2(sum_gx, sum_d, 2, r0, r1);
  const float mean1 = r0 / D1, mean2 = r1 / D1;
  float v1 = 0.f, v2 = 0.f;
  if (live) {
#pragma unroll
    for (int o = 0; o < C; ++o) {
      const float a = gxs[tid * LD + o] - mean1, c2 = ds[tid * LD + o] - mean2;
      v1 = fmaf(a, a, v1);
      v2 = fmaf(c2, c2, v2);
    }
  }
  block_sum2(v1, v2, 3, r0, r1);
  const float inv1 = 1.0f / (sqrtf(r0 / (D1 - 1.0f)) + eps);
  const float inv2 = 1.0f / (sqrtf(r1 / (D1 - 1.0f)) + eps);
  // ---- col: padded rows [F + 2][C]; z: [C][F]
  float* cb_ = col + b * (int64_t)(F + 2) * C;
  for (int i = tid; i < C; i += kT) {
    cb_[i] = 0.f;
    cb_[(int64_t)(F + 1) * C + i] = 0.f;
  }
  for (int i = tid; i < F * C; i += kT) {
    const int f = i / C, o = i - f * C;
    cb_[C + i] = (gxs[f * LD + o] - mean1) * inv1 * __ldg(ln1_w + i) + __ldg(ln1_b + i);
  }
  float* zb = z + b * (int64_t)F * C;
  for (int i = tid; i < F * C; i += kT) {
    const int o = i / F, f = i - o * F;
    const int src = f * C + o;
    zb[i] = (ds[f * LD + o] - mean2) * inv2 * __ldg(ln2_w + src) + __ldg(ln2_b + src);
  }
}

// ---------------------------------------------------------------------------------- LSTM sequences
// Thread = one sequence; weights, the thread's x / h / c / gate columns live in shared memory
// ([k][thread]: conflict-free), gate order i,f,g,o (PyTorch).  Sequence q = o*n_inner + i starts at
// x + o*x_outer + i*x_inner, steps are x_step apart (same decomposition for the output).
struct LstmArgs {
  const float* x;
  int64_t x_outer, x_inner, x_step;
  float* y;
  int64_t y_outer, y_inner, y_step;
  const float* w_ih;
  const float* w_hh;
  const float* b_ih;
  const float* b_hh;
  int64_t n_seq;
  int n_inner, L, n_in, H, reverse;
};
constexpr int kLstmThreads = 64;
__global__ void __launch_bounds__(kLstmThreads) lstm_seq_kernel(const LstmArgs a) {
  extern __shared__ float sm[];
  const int IN = a.n_in, H = a.H, G = 4 * a.H;
  float* wih = sm;                       // [G][IN]
  float* whh = wih + G * IN;             // [G][H]
  float* bias = whh + G * H;             // [G]
  float* xs = bias + G;                  // [IN][T]
  float* hs = xs + IN * kLstmThreads;    // [H][T]
  float* cs = hs + H * kLstmThreads;     // [H][T]
  float* gs = cs + H * kLstmThreads;     // [G][T]
  const int tid = threadIdx.x;
  for (int i = tid; i < G * IN; i += kLstmThreads) wih[i] = a.w_ih[i];
  for (int i = tid; i < G * H; i += kLstmThreads) whh[i] = a.w_hh[i];
  for (int i = tid; i < G; i += kLstmThreads) bias[i] = a.b_ih[i] + a.b_hh[i];
  for (int k = 0; k < H; ++k) {
    hs[k * kLstmThreads + tid] = 0.f;
    cs[k * kLstmThreads + tid] = 0.f;
  }
  __syncthreads();
  const int64_t q = (int64_t)blockIdx.x * kLstmThreads + tid;
  if (q >= a.n_seq) return;
  const int64_t qo = q / a.n_inner, qi = q - qo * a.n_inner;
  const float* xq = a.x + qo * a.x_outer + qi * a.x_inner;
  float* yq = a.y + qo * a.y_outer + qi * a.y_inner;
  for (int step = 0; step < a.L; ++step) {
    const int t = a.reverse ? a.L - 1 - step : step;
    const float* xt = xq + (int64_t)t * a.x_step;
    for (int k = 0; k < IN; ++k) xs[k * kLstmThreads + tid] = xt[k];
    for (int j = 0; j < H; ++j) {
      float gi = bias[j], gf = bias[H + j], gg = bias[2 * H + j], go = bias[3 * H + j];
      const float* wi0 = wih + j * IN;
      const float* wi1 = wih + (H + j) * IN;
      const float* wi2 = wih + (2 * H + j) * IN;
      const float* wi3 = wih + (3 * H + j) * IN;
      for (int k = 0; k < IN; ++k) {
        const float v = xs[k * kLstmThreads + tid];
        gi = fmaf(wi0[k], v, gi); gf = fmaf(wi1[k], v, gf); gg = fmaf(wi2[k], v, gg); go = fmaf(wi3[k], v, go);
      }
      const float* wh0 = whh + j * H;
      const float* wh1 = whh + (H + j) * H;
      const float* wh2 = whh + (2 * H + j) * H;
      const float* wh3 = whh + (3 * H + j) * H;
      for (int k = 0; k < H; ++k) {
        const float v = hs[k * kLstmThreads + tid];
        gi = fmaf(wh0[k], v, gi); gf = fmaf(wh1[k], v, gf); gg = fmaf(wh2[k], v, gg); go = fmaf(wh3[k], v, go);
      }
      gs[j * kLstmThreads + tid] = gi;
      gs[(H + j) * kLstmThreads + tid] = gf;
      gs[(2 * H + j) * kLstmThreads + tid] = gg;
      gs[(3 * H + j) * kLstmThreads + tid] = go;
    }
    float* yt = yq + (int64_t)t * a.y_step;
    for (int j = 0; j < H; ++j) {
      const float ig = 1.0f / (1.0f + expf(-gs[j * kLstmThreads + tid]));
      const float fg = 1.0f / (1.0f + expf(-gs[(H + j) * kLstmThreads + tid]));
      const float gg = tanhf(gs[(2 * H + j) * kLstmThreads + tid]);
      const float og = 1.0f / (1.0f + expf(-gs[(3 * H + j) * kLstmThreads + tid]));
      const float c = fg * cs[j * kLstmThreads + tid] + ig * gg;
      const float h = og * tanhf(c);
      cs[j * kLstmThreads + tid] = c;
      hs[j * kLstmThreads + tid] = h;
      yt[j] = h;
    }
  }
}

// Gate-parallel variant: one thread per (sequence, gate row).  The thread keeps its row of
// [W_ih | W_hh] in registers for the whole sequence (compile-time IN, H), x_t and h_{t-1} of the
// sequence sit in shared memory (broadcast reads), the cell state lives in a register of the
// thread that owns hidden unit j.  Two barriers per time step.
template <int IN, int H, int NSEQ>
__global__ void __launch_bounds__(4 * H * NSEQ) lstm_seq_gates_kernel(const LstmArgs a) {
  constexpr int G = 4 * H;
  __shared__ float xs[NSEQ][IN];
  __shared__ float hs[NSEQ][H];
  __shared__ float gs[NSEQ][G];
  const int g = threadIdx.x % G;       // gate row (i: 0..H-1, f: H.., g: 2H.., o: 3H..)
  const int ql = threadIdx.x / G;      // local sequence
  const int64_t q = (int64_t)blockIdx.x * NSEQ + ql;
  const bool live = q < a.n_seq;
  const int64_t qc = live ? q : 0;
  const int64_t qo = qc / a.n_inner, qi = qc - qo * a.n_inner;
  const float* xq = a.x + qo * a.x_outer + qi * a.x_inner;
  float* yq = a.y + qo * a.y_outer + qi * a.y_inner;
  float wi[IN], wh[H];
#pragma unroll
  for (int k = 0; k < IN; ++k) wi[k] = a.w_ih[g * IN + k];
#pragma unroll
  for (int k = 0; k < H; ++k) wh[k] = a.w_hh[g * H + k];
  const float bias = a.b_ih[g] + a.b_hh[g];
  float c = 0.f;
  if (g < H) hs[ql][g] = 0.f;
  {
    const int t0 = a.reverse ? a.L - 1 : 0;
    if (g < IN) xs[ql][g] = live ? xq[(int64_t)t0 * a.x_step + g] : 0.f;
  }
  __syncthreads();
  for (int step = 0; step < a.L; ++step) {
    const int t = a.reverse ? a.L - 1 - step : step;
    // four independent partial sums: the (IN + H)-long dependent FMA chain is the latency of a step
    float a4[4] = {bias, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < IN; ++k) a4[k & 3] = fmaf(wi[k], xs[ql][k], a4[k & 3]);
#pragma unroll
    for (int k = 0; k < H; ++k) a4[(IN + k) & 3] = fmaf(wh[k], hs[ql][k], a4[(IN + k) & 3]);
    const float acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
    gs[ql][g] = acc;
    __syncthreads();
    if (g < H) {
      const float ig = 1.0f / (1.0f + expf(-gs[ql][g]));
      const float fg = 1.0f / (1.0f + expf(-gs[ql][H + g]));
      const float gg = tanhf(gs[ql][2 * H + g]);
      const float og = 1.0f / (1.0f + expf(-gs[ql][3 * H + g]));
      c = fg * c + ig * gg;
      const float h = og * tanhf(c);
      hs[ql][g] = h;
      if (live) yq[(int64_t)t * a.y_step + g] = h;
    } else if (g - H < IN && step + 1 < a.L) {
      // the other threads fetch the next step's input meanwhile
      const int tn = a.reverse ? a.L - 2 - step : step + 1;
      xs[ql][g - H] = live ? xq[(int64_t)tn * a.x_step + (g - H)] : 0.f;
    }
    __syncthreads();
  }
}

template <int IN, int H, int NSEQ>
static int launch_lstm_gates(const LstmArgs& a, cudaStream_t st) {
  static_assert(IN <= 3 * H, "the threads of gate rows H..4H-1 prefetch x: need IN <= 3H");
  lstm_seq_gates_kernel<IN, H, NSEQ><<<(unsigned)ceil_div(a.n_seq, NSEQ), 4 * H * NSEQ, 0, st>>>(a);
  return after_launch("vadx_lstm_seq_f32");
}

// Recurrence-only LSTM (the input half of the gates, x W_ih^T + b_ih + b_hh, is computed beforehand as ONE dense layer over
// all steps of all sequences -- no recurrence in it, so it runs on the tensor cores): one THREAD per sequence, h in
// registers, c and the next h in conflict-free shared-memory columns, W_hh broadcast from shared memory as float4.  The gate
// rows are PERMUTED so that the four gates of hidden unit j are adjacent (row' = 4j + gate): one float4 of pre-activations
// per unit and step.  No barrier inside the time loop (the gate-parallel kernel above needs two per step and one
// shared-memory read per FMA).
struct LstmRecArgs {
  const float* g;      // permuted gate pre-activations; sequence q = (qo, qi) at g + qo*g_outer + qi*g_inner, step t at + t*g_step
  int64_t g_outer, g_inner, g_step;
  float* y;
  int64_t y_outer, y_inner, y_step;
  const float* w_hh_p;  // [H][4][H]: unit j, gate (i, f, g, o), k
  int64_t n_seq;
  int n_inner, L, reverse;
};
constexpr int kRecThreads = 128;
template <int H>
__global__ void __launch_bounds__(kRecThreads) lstm_rec_kernel(const LstmRecArgs a) {
  extern __shared__ float rec_sm[];
  float* w = rec_sm;                        // [H][4][H]
  float* cs = w + 4 * H * H;                // [H][kRecThreads]
  float* hn = cs + H * kRecThreads;         // [H][kRecThreads]
  const int tid = threadIdx.x;
  for (int i = tid; i < 4 * H * H; i += kRecThreads) w[i] = a.w_hh_p[i];
#pragma unroll
  for (int j = 0; j < H; ++j) cs[j * kRecThreads + tid] = 0.f;
  __syncthreads();
  const int64_t q = (int64_t)blockIdx.x * kRecThreads + tid;
  if (q >= a.n_seq) return;
  const int64_t qo = q / a.n_inner, qi = q - qo * a.n_inner;
  const float* gq = a.g + qo * a.g_outer + qi * a.g_inner;
  float* yq = a.y + qo * a.y_outer + qi * a.y_inner;
  float h[H];
#pragma unroll
  for (int k = 0; k < H; ++k) h[k] = 0.f;
  for (int step = 0; step < a.L; ++step) {
    const int t = a.reverse ? a.L - 1 - step : step;
    const float4* gt = reinterpret_cast<const float4*>(gq + (int64_t)t * a.g_step);
    // two units per 256-bit load: a lane's row is 52 KB away from its neighbour's, so every load touches 32 different
    // sectors; with 128-bit loads each 32-byte sector was requested twice and L1 (10 % hit rate under ncu) did not keep it
    float4 nxt[2];
    ldg256_nc(reinterpret_cast<const float*>(gt), nxt[0], nxt[1]);
    for (int j = 0; j < H; j += 2) {
      float4 acc2[2] = {nxt[0], nxt[1]};
      // the next pair's pre-activations are requested before this pair's arithmetic (the last iteration re-reads pair 0 of
      // the row: harmless, in bounds)
      ldg256_nc(reinterpret_cast<const float*>(gt + (j + 2 < H ? j + 2 : 0)), nxt[0], nxt[1]);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float4 acc = acc2[u];
        const float4* wj = reinterpret_cast<const float4*>(w + (j + u) * 4 * H);
#pragma unroll
        for (int k4 = 0; k4 < H / 4; ++k4) {
          const float4 wi = wj[k4], wf = wj[H / 4 + k4], wg = wj[2 * (H / 4) + k4], wo = wj[3 * (H / 4) + k4];
          acc.x = fmaf(wi.w, h[4 * k4 + 3], fmaf(wi.z, h[4 * k4 + 2], fmaf(wi.y, h[4 * k4 + 1], fmaf(wi.x, h[4 * k4], acc.x))));
          acc.y = fmaf(wf.w, h[4 * k4 + 3], fmaf(wf.z, h[4 * k4 + 2], fmaf(wf.y, h[4 * k4 + 1], fmaf(wf.x, h[4 * k4], acc.y))));
          acc.z = fmaf(wg.w, h[4 * k4 + 3], fmaf(wg.z, h[4 * k4 + 2], fmaf(wg.y, h[4 * k4 + 1], fmaf(wg.x, h[4 * k4], acc.z))));
          acc.w = fmaf(wo.w, h[4 * k4 + 3], fmaf(wo.z, h[4 * k4 + 2], fmaf(wo.y, h[4 * k4 + 1], fmaf(wo.x, h[4 * k4], acc.w))));
        }
        const float ig = 1.0f / (1.0f + expf(-acc.x));
        const float fg = 1.0f / (1.0f + expf(-acc.y));
        const float gg = tanhf(acc.z);
        const float og = 1.0f / (1.0f + expf(-acc.w));
        const float c = fg * cs[(j + u) * kRecThreads + tid] + ig * gg;
        cs[(j + u) * kRecThreads + tid] = c;
        hn[(j + u) * kRecThreads + tid] = og * tanhf(c);
      }
    }
    float4* yt = reinterpret_cast<float4*>(yq + (int64_t)t * a.y_step);
#pragma unroll
    for (int k4 = 0; k4 < H / 4; ++k4) {
      h[4 * k4] = hn[(4 * k4) * kRecThreads + tid];
      h[4 * k4 + 1] = hn[(4 * k4 + 1) * kRecThreads + tid];
      h[4 * k4 + 2] = hn[(4 * k4 + 2) * kRecThreads + tid];
      h[4 * k4 + 3] = hn[(4 * k4 + 3) * kRecThreads + tid];
      yt[k4] = make_float4(h[4 * k4], h[4 * k4 + 1], h[4 * k4 + 2], h[4 * k4 + 3]);
    }
  }
}

template <int H>
static int launch_lstm_rec(const LstmRecArgs& a, cudaStream_t st) {
  const size_t smem = ((size_t)4 * H * H + 2 * (size_t)H * kRecThreads) * sizeof(float);
  static PerDevice per_device;
  VADX_TRY(per_device.ensure(nullptr, [] {
    return cudaFuncSetAttribute(lstm_rec_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  }));
  lstm_rec_kernel<H><<<(unsigned)ceil_div(a.n_seq, kRecThreads), kRecThreads, smem, st>>>(a);
  return after_launch("vadx_lstm_recurrence_f32");
}

// ---------------------------------------------------------------------------------- element-wise
// op 0: out = a + b; 1: out = a * b; 2: out = a - s*b; 3: out = a (copy); 4: gate: out = a*b, out2 = b - a*b
__global__ void __launch_bounds__(256) ew2_kernel(int op, const float* __restrict__ a, int64_t lda,
                                                  const float* __restrict__ b, int64_t ldb, float* __restrict__ out,
                                                  int64_t ldo, float* __restrict__ out2, int64_t ldo2, int64_t rows,
                                                  int cols, float s) {
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    const int c = (int)(i - r * cols);
    const float av = a[r * lda + c];
    const float bv = b ? b[r * ldb + c] : 0.f;
    float v;
    if (op == 0) v = av + bv;
    else if (op == 1) v = av * bv;
    else if (op == 2) v = av - s * bv;
    else if (op == 3) v = av;
    else {
      v = av * bv;
      out2[r * ldo2 + c] = bv - v;
    }
    out[r * ldo + c] = v;
  }
}

// CepsUnit complex product on [rows = (blk, bin)][2C]: q = LSTM output (re | im), p = spectrum (re | im)
__global__ void __launch_bounds__(256) cmul_kernel(const float* __restrict__ q, const float* __restrict__ p,
                                                   float* __restrict__ out, int64_t rows, int C) {
  const int64_t total = rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    const float qr = q[r * 2 * C + c], qi = q[r * 2 * C + C + c];
    const float pr = p[r * 2 * C + c], pi = p[r * 2 * C + C + c];
    out[r * 2 * C + c] = qr * pr - qi * pi;
    out[r * 2 * C + C + c] = qr * pi + qi * pr;
  }
}

// CepsUnit complex gain with the layout change folded in (Export_DFSMN_VAD.py:145-146): q = the bi-LSTM head's output in
// cepstral-major order [B][cb][re(C) | im(C)], spec = the cepstral DFT's native output [B][C][re(cb) | im(cb)]; the product is
// written in spec's layout, which is what the inverse DFT layer reads -- neither permute4 launch of the unfused sequence
// (spec -> cepstral-major copy, product -> back) is needed.
__global__ void __launch_bounds__(256) cmul_t_kernel(const float* __restrict__ q, const float* __restrict__ spec,
                                                     float* __restrict__ out, int64_t B, int C, int cb) {
  const int64_t total = B * C * cb;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % cb);
    const int64_t r = i / cb;
    const int c = (int)(r % C);
    const int64_t b = r / C;
    const float* qp = q + (b * cb + k) * 2 * C;
    const float qr = qp[c], qi = qp[C + c];
    const int64_t sp = (b * C + c) * 2 * cb + k;
    const float pr = spec[sp], pi = spec[sp + cb];
    out[sp] = qr * pr - qi * pi;
    out[sp + cb] = qr * pi + qi * pr;
  }
}

// out[b][f][c] = a[b][f][c] + t[b][c][f]: the block's last add with the [C][F] -> [F][C] transpose of the inverse DFT folded in
__global__ void __launch_bounds__(256) add_transposed_kernel(const float* __restrict__ a, int64_t a_block, const float* __restrict__ t,
                                                             float* __restrict__ out, int64_t B, int F, int C) {
  const int64_t total = B * F * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t r = i / C;
    const int f = (int)(r % F);
    const int64_t b = r / F;
    out[i] = a[b * a_block + (int64_t)f * C + c] + t[(b * C + c) * F + f];
  }
}

// 3-tap frequency im2col: out[(blk, f)][j*C + c] = x[(blk, f + j - 1)][c], zero outside [0, F)
__global__ void __launch_bounds__(256) im2col_f3_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                        int64_t n_blk, int F, int C) {
  const int64_t total = n_blk * F * 3 * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int j = (int)(r % 3);
    r /= 3;
    const int f = (int)(r % F);
    const int64_t blk = r / F;
    const int fs = f + j - 1;
    out[i] = (fs >= 0 && fs < F) ? x[(blk * F + fs) * C + c] : 0.f;
  }
}

// AlphaPredictor + input assembly: near/far complex spectra [S*T][2F] (re, im interleaved) ->
// x4 [S*T*F][4] = (near_re, near_im, far_re*|alpha|, far_im*|alpha|), alpha per (t, f) from the
// last k frames' powers (zero history): Export_DFSMN_VAD.py:326-336.
__global__ void __launch_bounds__(256) alpha_x4_kernel(const float* __restrict__ near_ri, const float* __restrict__ far_ri,
                                                       int64_t n_streams, int T, int F, int k, float w1_far, float w1_mix,
                                                       float b1, const float* __restrict__ w2, float b2,
                                                       float* __restrict__ x4, float* __restrict__ alpha_out) {
  const int64_t total = n_streams * T * F;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    const int64_t st = i / F;
    const int t = (int)(st % T);
    float acc = b2;
    for (int j = 0; j < k; ++j) {
      const int tt = t - (k - 1) + j;
      float pm = 0.f, pf = 0.f;
      if (tt >= 0) {
        const int64_t row = st - t + tt;
        const float nr = near_ri[row * 2 * F + 2 * f], ni = near_ri[row * 2 * F + 2 * f + 1];
        const float fr = far_ri[row * 2 * F + 2 * f], fi = far_ri[row * 2 * F + 2 * f + 1];
        pm = nr * nr + ni * ni;
        pf = fr * fr + fi * fi;
      }
      acc = fmaf(w2[j], fmaf(w1_far, pf, fmaf(w1_mix, pm, b1)), acc);
    }
    const float aa = fabsf(acc);
    x4[i * 4 + 0] = near_ri[st * 2 * F + 2 * f];
    x4[i * 4 + 1] = near_ri[st * 2 * F + 2 * f + 1];
    x4[i * 4 + 2] = far_ri[st * 2 * F + 2 * f] * aa;
    x4[i * 4 + 3] = far_ri[st * 2 * F + 2 * f + 1] * aa;
    if (alpha_out) alpha_out[i] = acc;
  }
}

// Near-end-only variant (DFSMN/only_near_end_audio/Export_DFSMN_VAD.py:319-335): the far end's k-frame power
// window and its spectrum are constants of the graph, pow_far [F][Tmax][k] and far_comp [2][F][Tmax].
__global__ void __launch_bounds__(256) alpha_x4_const_kernel(const float* __restrict__ near_ri,
                                                             const float* __restrict__ pow_far,
                                                             const float* __restrict__ far_comp, int t_max,
                                                             int64_t n_streams, int T, int F, int k, float w1_far,
                                                             float w1_mix, float b1, const float* __restrict__ w2, float b2,
                                                             float* __restrict__ x4, float* __restrict__ alpha_out) {
  const int64_t total = n_streams * T * F;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    const int64_t st = i / F;
    const int t = (int)(st % T);
    float acc = b2;
    for (int j = 0; j < k; ++j) {
      const int tt = t - (k - 1) + j;
      float pm = 0.f;
      if (tt >= 0) {
        const int64_t row = st - t + tt;
        const float nr = near_ri[row * 2 * F + 2 * f], ni = near_ri[row * 2 * F + 2 * f + 1];
        pm = nr * nr + ni * ni;
      }
      const float pf = pow_far[((int64_t)f * t_max + t) * k + j];
      acc = fmaf(w2[j], fmaf(w1_far, pf, fmaf(w1_mix, pm, b1)), acc);
    }
    const float aa = fabsf(acc);
    x4[i * 4 + 0] = near_ri[st * 2 * F + 2 * f];
    x4[i * 4 + 1] = near_ri[st * 2 * F + 2 * f + 1];
    x4[i * 4 + 2] = far_comp[(int64_t)f * t_max + t] * aa;
    x4[i * 4 + 3] = far_comp[((int64_t)F + f) * t_max + t] * aa;
    if (alpha_out) alpha_out[i] = acc;
  }
}

// ISTFT overlap-add (NET.istft :226-230): frames [S*T][ld] (n_fft valid) -> y[s][i] = window_sum_inv[i + half] *
// sum_t frames[t][i + half - t*hop]
__global__ void __launch_bounds__(256) istft_ola_kernel(const float* __restrict__ frames, int64_t ld, int64_t n_streams,
                                                        int T, int n_fft, int hop, const float* __restrict__ wsum_inv,
                                                        int n_out, float* __restrict__ y, int64_t ldy) {
  const int half = n_fft / 2;
  const int64_t total = n_streams * n_out;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / n_out;
    const int n = (int)(i - s * n_out) + half;
    float acc = 0.f;
    int t_hi = n / hop;
    if (t_hi > T - 1) t_hi = T - 1;
    for (int t = t_hi; t >= 0 && n - t * hop < n_fft; --t) acc += frames[(s * T + t) * ld + (n - t * hop)];
    y[s * ldy + (n - half)] = acc * wsum_inv[n];
  }
}

}  // namespace vadx

using namespace vadx;

static inline unsigned g1(int64_t items) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(items, 256), 148 * 32));
}

extern "C" int vadx_permute4_f32(const float* d_in, float* d_out, int64_t n0, int64_t n1, int64_t n2, int64_t n3, int p0,
                                 int p1, int p2, int p3, void* stream) {
  StageTimer _timer(VADX_STAGE_PREP, (cudaStream_t)stream, "permute4_kernel", 8.0 * n0 * n1 * n2 * n3);
  VADX_REQUIRE(d_in && d_out && d_in != d_out && n0 >= 0 && n1 >= 1 && n2 >= 1 && n3 >= 1, "vadx_permute4_f32: bad argument");
  int seen = 0;
  for (int p : {p0, p1, p2, p3}) {
    VADX_REQUIRE(p >= 0 && p < 4, "vadx_permute4_f32: bad permutation");
    seen |= 1 << p;
  }
  VADX_REQUIRE(seen == 15, "vadx_permute4_f32: not a permutation");
  if (n0 == 0) return VADX_OK;
  permute4_kernel<<<g1(n0 * n1 * n2 * n3), 256, 0, (cudaStream_t)stream>>>(d_in, d_out, n0, n1, n2, n3, p0, p1, p2, p3);
  return after_launch("vadx_permute4_f32");
}

extern "C" int vadx_layernorm_f32(const float* d_x, int64_t n_rows, int row_len, const float* d_w, const float* d_b,
                                  float eps, float* d_out, void* stream) {
  StageTimer _timer(VADX_STAGE_MEL, (cudaStream_t)stream, "layernorm_kernel", 8.0 * n_rows * row_len);
  VADX_REQUIRE(d_x && d_w && d_b && d_out && n_rows >= 0 && row_len >= 2, "vadx_layernorm_f32: bad argument");
  VADX_REQUIRE(n_rows <= 0x7fffffffLL, "vadx_layernorm_f32: too many rows");
  if (n_rows == 0) return VADX_OK;
  layernorm_kernel<<<(unsigned)n_rows, 256, 0, (cudaStream_t)stream>>>(d_x, n_rows, row_len, d_w, d_b, eps, d_out);
  return after_launch("vadx_layernorm_f32");
}

extern "C" int vadx_layernorm_perm_f32(const float* d_x, int64_t n_rows, int n1, int n2, int n3, int q0, int q1, int q2,
                                       const float* d_w, const float* d_b, int w_out_order, float eps, float* d_out,
                                       int64_t out_stride, int pad, void* stream) {
  StageTimer _timer(VADX_STAGE_MEL, (cudaStream_t)stream, "layernorm_perm_kernel", 8.0 * n_rows * n1 * n2 * n3);
  VADX_REQUIRE(d_x && d_w && d_b && d_out && d_x != d_out && n_rows >= 0 && n1 >= 1 && n2 >= 1 && n3 >= 1,
               "vadx_layernorm_perm_f32: bad argument");
  VADX_REQUIRE((int64_t)n1 * n2 * n3 >= 2 && (int64_t)n1 * n2 * n3 <= 12288, "vadx_layernorm_perm_f32: row of %lld elements",
               (long long)n1 * n2 * n3);
  VADX_REQUIRE(((1 << q0) | (1 << q1) | (1 << q2)) == 7 && q0 >= 0 && q1 >= 0 && q2 >= 0, "vadx_layernorm_perm_f32: not a permutation");
  VADX_REQUIRE(n_rows <= 0x7fffffffLL, "vadx_layernorm_perm_f32: too many rows");
  if (out_stride == 0) out_stride = (int64_t)n1 * n2 * n3 + 2 * pad;
  VADX_REQUIRE(pad >= 0 && out_stride >= (int64_t)n1 * n2 * n3 + 2 * pad, "vadx_layernorm_perm_f32: out_stride %lld too small",
               (long long)out_stride);
  if (n_rows == 0) return VADX_OK;
  static PerDevice per_device;
  VADX_TRY(per_device.ensure(nullptr, [] {
    return cudaFuncSetAttribute(layernorm_perm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 12288 * 4);
  }));
  layernorm_perm_kernel<<<(unsigned)n_rows, 256, (size_t)n1 * n2 * n3 * sizeof(float), (cudaStream_t)stream>>>(
      d_x, n1, n2, n3, q0, q1, q2, d_w, d_b, w_out_order, eps, d_out, out_stride, pad);
  return after_launch("vadx_layernorm_perm_f32");
}

extern "C" int vadx_ceps_cmul_t_f32(const float* d_q, const float* d_spec, float* d_out, int64_t n_blocks, int n_channels,
                                    int n_ceps, void* stream) {
  StageTimer _timer(VADX_STAGE_MEL, (cudaStream_t)stream, "cmul_t_kernel", 4.0 * n_blocks * n_channels * n_ceps * 6.0);
  VADX_REQUIRE(d_q && d_spec && d_out && n_blocks >= 0 && n_channels >= 1 && n_ceps >= 1, "vadx_ceps_cmul_t_f32: bad argument");
  if (n_blocks == 0) return VADX_OK;
  cmul_t_kernel<<<g1(n_blocks * n_channels * n_ceps), 256, 0, (cudaStream_t)stream>>>(d_q, d_spec, d_out, n_blocks, n_channels, n_ceps);
  return after_launch("vadx_ceps_cmul_t_f32");
}

extern "C" int vadx_add_transposed_f32(const float* d_a, int64_t a_block_stride, const float* d_t, float* d_out, int64_t n_blocks,
                                       int n_bins, int n_channels, void* stream) {
  StageTimer _timer(VADX_STAGE_MEL, (cudaStream_t)stream, "add_transposed_kernel", 12.0 * n_blocks * n_bins * n_channels);
  VADX_REQUIRE(d_a && d_t && d_out && n_blocks >= 0 && n_bins >= 1 && n_channels >= 1, "vadx_add_transposed_f32: bad argument");
  if (n_blocks == 0) return VADX_OK;
  if (a_block_stride == 0) a_block_stride = (int64_t)n_bins * n_channels;
  VADX_REQUIRE(a_block_stride >= (int64_t)n_bins * n_channels, "vadx_add_transposed_f32: a_block_stride too small");
  add_transposed_kernel<<<g1(n_blocks * n_bins * n_channels), 256, 0, (cudaStream_t)stream>>>(d_a, a_block_stride, d_t, d_out,
                                                                                          n_blocks, n_bins, n_channels);
  return after_launch("vadx_add_transposed_f32");
}

extern "C" int vadx_cfb_front_supported(int n_in, int n_channels, int n_bins) {
  return (n_channels == 20 && (n_in == 20 || n_in == 40) && n_bins >= 2 && n_bins <= 160) ? 1 : 0;
}

extern "C" int vadx_cfb_front_f32(const float* d_x, int64_t n_blocks, int n_bins, int n_in, int n_channels, const float* d_ln0_w,
                                  const float* d_ln0_b, const float* d_wg, const float* d_bg, const float* d_wi, const float* d_bi,
                                  const float* d_ln1_w, const float* d_ln1_b, const float* d_ln2_w, const float* d_ln2_b, float eps,
                                  float* d_col, float* d_z, void* stream) {
  StageTimer _timer(VADX_STAGE_MEL, (cudaStream_t)stream, "cfb_front_kernel",
                    4.0 * n_blocks * n_bins * ((double)n_in + 2.0 * n_channels), 4.0 * n_blocks * n_bins * n_channels * (double)n_in);
  VADX_REQUIRE(d_x && d_ln0_w && d_ln0_b && d_wg && d_bg && d_wi && d_bi && d_ln1_w && d_ln1_b && d_ln2_w && d_ln2_b && d_col && d_z,
               "vadx_cfb_front_f32: null pointer");
  VADX_REQUIRE(vadx_cfb_front_supported(n_in, n_channels, n_bins), "vadx_cfb_front_f32: shape (%d -> %d channels, %d bins) not instantiated",
               n_in, n_channels, n_bins);
  VADX_REQUIRE(n_blocks >= 0 && n_blocks <= 0x7fffffffLL && aligned16(d_x) && aligned16(d_ln0_w) && aligned16(d_ln0_b),
               "vadx_cfb_front_f32: bad argument");
  if (n_blocks == 0) return VADX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  constexpr int C = 20;
  if (n_in == 20) {
    const size_t smem = (size_t)(2 * C * 20 + 2 * 160 * (C + 1)) * sizeof(float);
    cfb_front_kernel<20, C><<<(unsigned)n_blocks, 160, smem, st>>>(d_x, n_bins, d_ln0_w, d_ln0_b, d_wg, d_bg, d_wi, d_bi, d_ln1_w, d_ln1_b,
                                                                d_ln2_w, d_ln2_b, eps, d_col, d_z);
  } else {
    const size_t smem = (size_t)(2 * C * 40 + 2 * 160 * (C + 1)) * sizeof(float);
    cfb_front_kernel<40, C><<<(unsigned)n_blocks, 160, smem, st>>>(d_x, n_bins, d_ln0_w, d_ln0_b, d_wg, d_bg, d_wi, d_bi, d_ln1_w, d_ln1_b,
                                                                d_ln2_w, d_ln2_b, eps, d_col, d_z);
  }
  return after_launch("vadx_cfb_front_f32");
}

extern "C" int vadx_lstm_seq_f32(const float* d_x, int64_t x_outer, int64_t x_inner, int64_t x_step, float* d_y,
                                 int64_t y_outer, int64_t y_inner, int64_t y_step, const float* d_w_ih,
                                 const float* d_w_hh, const float* d_b_ih, const float* d_b_hh, int64_t n_seq,
                                 int n_inner, int seq_len, int n_in, int hidden, int reverse, void* stream) {
  StageTimer _timer(VADX_STAGE_MEMORY, (cudaStream_t)stream, "lstm_seq_kernel",
                    4.0 * n_seq * seq_len * ((double)n_in + hidden), 2.0 * n_seq * seq_len * 4.0 * hidden * ((double)n_in + hidden));
  VADX_REQUIRE(d_x && d_y && d_w_ih && d_w_hh && d_b_ih && d_b_hh, "vadx_lstm_seq_f32: null pointer");
  VADX_REQUIRE(n_seq >= 0 && n_inner >= 1 && seq_len >= 1 && n_in >= 1 && hidden >= 1, "vadx_lstm_seq_f32: bad shape");
  if (n_seq == 0) return VADX_OK;
  {
    LstmArgs a{d_x, x_outer, x_inner, x_step, d_y, y_outer, y_inner, y_step, d_w_ih, d_w_hh, d_b_ih, d_b_hh,
               n_seq, n_inner, seq_len, n_in, hidden, reverse};
    cudaStream_t st = (cudaStream_t)stream;
    if (n_in == 4 && hidden == 20) return launch_lstm_gates<4, 20, 4>(a, st);
    if (n_in == 40 && hidden == 20) return launch_lstm_gates<40, 20, 4>(a, st);
    if (n_in == 20 && hidden == 40) return launch_lstm_gates<20, 40, 2>(a, st);
    if (n_in == 40 && hidden == 40) return launch_lstm_gates<40, 40, 2>(a, st);
  }
  const size_t G = 4 * (size_t)hidden;
  const size_t smem = (G * n_in + G * hidden + G + (size_t)kLstmThreads * (n_in + 2 * hidden + G)) * sizeof(float);
  VADX_REQUIRE(smem <= 200 * 1024, "vadx_lstm_seq_f32: in=%d hidden=%d needs %zu bytes of shared memory", n_in, hidden, smem);
  static PerDevice per_device;
  VADX_TRY(per_device.ensure(nullptr, [] {
    return cudaFuncSetAttribute(lstm_seq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }));
  LstmArgs a{d_x, x_outer, x_inner, x_step, d_y, y_outer, y_inner, y_step, d_w_ih, d_w_hh, d_b_ih, d_b_hh,
             n_seq, n_inner, seq_len, n_in, hidden, reverse};
  lstm_seq_kernel<<<(unsigned)ceil_div(n_seq, kLstmThreads), kLstmThreads, smem, (cudaStream_t)stream>>>(a);
  return after_launch("vadx_lstm_seq_f32");
}

extern "C" int vadx_lstm_recurrence_supported(int hidden) { return (hidden == 20 || hidden == 40) ? 1 : 0; }

extern "C" int vadx_lstm_recurrence_f32(const float* d_gates_in, int64_t g_outer, int64_t g_inner, int64_t g_step, float* d_y,
                                        int64_t y_outer, int64_t y_inner, int64_t y_step, const float* d_w_hh_perm,
                                        int64_t n_seq, int n_inner, int seq_len, int hidden, int reverse, void* stream) {
  StageTimer _timer(VADX_STAGE_MEMORY, (cudaStream_t)stream, "lstm_rec_kernel", 4.0 * n_seq * seq_len * 5.0 * hidden,
                    2.0 * n_seq * seq_len * 4.0 * hidden * hidden);
  VADX_REQUIRE(d_gates_in && d_y && d_w_hh_perm, "vadx_lstm_recurrence_f32: null pointer");
  VADX_REQUIRE(n_seq >= 0 && n_inner >= 1 && seq_len >= 1, "vadx_lstm_recurrence_f32: bad shape");
  VADX_REQUIRE(vadx_lstm_recurrence_supported(hidden), "vadx_lstm_recurrence_f32: hidden %d is not instantiated (20, 40)", hidden);
  VADX_REQUIRE(((y_outer | y_inner | y_step) & 3) == 0 && aligned16(d_y) && ((g_outer | g_inner | g_step) & 7) == 0 &&
                   (reinterpret_cast<uintptr_t>(d_gates_in) & 31u) == 0,
               "vadx_lstm_recurrence_f32: output strides must be multiples of 4 floats (16-byte aligned base), gate strides multiples "
               "of 8 floats (32-byte aligned base)");
  if (n_seq == 0) return VADX_OK;
  LstmRecArgs a{d_gates_in, g_outer, g_inner, g_step, d_y, y_outer, y_inner, y_step, d_w_hh_perm, n_seq, n_inner, seq_len, reverse};
  if (hidden == 20) return launch_lstm_rec<20>(a, (cudaStream_t)stream);
  return launch_lstm_rec<40>(a, (cudaStream_t)stream);
}

extern "C" int vadx_ew2_f32(int op, const float* d_a, int64_t lda, const float* d_b, int64_t ldb, float* d_out,
                            int64_t ldo, float* d_out2, int64_t ldo2, int64_t n_rows, int n_cols, float scalar,
                            void* stream) {
  StageTimer _timer(VADX_STAGE_MEL, (cudaStream_t)stream, "ew2_kernel", 4.0 * n_rows * n_cols * (1 + (d_b ? 1 : 0) + 1 + (d_out2 ? 1 : 0)));
  VADX_REQUIRE(d_a && d_out && op >= 0 && op <= 4 && (op == 3 || d_b) && (op != 4 || d_out2) && n_rows >= 0 && n_cols >= 1,
               "vadx_ew2_f32: bad argument");
  if (n_rows == 0) return VADX_OK;
  ew2_kernel<<<g1(n_rows * n_cols), 256, 0, (cudaStream_t)stream>>>(op, d_a, lda, d_b, ldb, d_out, ldo, d_out2, ldo2,
                                                                    n_rows, n_cols, scalar);
  return after_launch("vadx_ew2_f32");
}

extern "C" int vadx_ceps_cmul_f32(const float* d_q, const float* d_p, float* d_out, int64_t n_rows, int n_channels,
                                  void* stream) {
  StageTimer _timer(VADX_STAGE_MEL, (cudaStream_t)stream, "cmul_kernel", 4.0 * n_rows * n_channels * 6.0);
  VADX_REQUIRE(d_q && d_p && d_out && n_rows >= 0 && n_channels >= 1, "vadx_ceps_cmul_f32: bad argument");
  if (n_rows == 0) return VADX_OK;
  cmul_kernel<<<g1(n_rows * n_channels), 256, 0, (cudaStream_t)stream>>>(d_q, d_p, d_out, n_rows, n_channels);
  return after_launch("vadx_ceps_cmul_f32");
}

extern "C" int vadx_im2col_f3_f32(const float* d_x, float* d_out, int64_t n_blocks, int n_bins, int n_channels,
                                  void* stream) {
  StageTimer _timer(VADX_STAGE_PREP, (cudaStream_t)stream, "im2col_f3_kernel", 4.0 * n_blocks * n_bins * n_channels * 4.0);
  VADX_REQUIRE(d_x && d_out && n_blocks >= 0 && n_bins >= 1 && n_channels >= 1, "vadx_im2col_f3_f32: bad argument");
  if (n_blocks == 0) return VADX_OK;
  im2col_f3_kernel<<<g1(n_blocks * n_bins * 3 * n_channels), 256, 0, (cudaStream_t)stream>>>(d_x, d_out, n_blocks, n_bins,
                                                                                         n_channels);
  return after_launch("vadx_im2col_f3_f32");
}

extern "C" int vadx_alpha_x4_f32(const float* d_near_ri, const float* d_far_ri, int64_t n_streams, int n_frames,
                                 int n_bins, int k, float w1_far, float w1_mix, float b1, const float* d_w2, float b2,
                                 float* d_x4, float* d_alpha, void* stream) {
  StageTimer _timer(VADX_STAGE_PREP, (cudaStream_t)stream, "alpha_x4_kernel", 4.0 * n_streams * n_frames * n_bins * 8.0);
  VADX_REQUIRE(d_near_ri && d_far_ri && d_w2 && d_x4 && n_streams >= 0 && n_frames >= 1 && n_bins >= 1 && k >= 1,
               "vadx_alpha_x4_f32: bad argument");
  if (n_streams == 0) return VADX_OK;
  alpha_x4_kernel<<<g1(n_streams * n_frames * n_bins), 256, 0, (cudaStream_t)stream>>>(
      d_near_ri, d_far_ri, n_streams, n_frames, n_bins, k, w1_far, w1_mix, b1, d_w2, b2, d_x4, d_alpha);
  return after_launch("vadx_alpha_x4_f32");
}

extern "C" int vadx_alpha_x4_const_f32(const float* d_near_ri, const float* d_pow_far, const float* d_far_comp,
                                       int t_max, int64_t n_streams, int n_frames, int n_bins, int k, float w1_far,
                                       float w1_mix, float b1, const float* d_w2, float b2, float* d_x4, float* d_alpha,
                                       void* stream) {
  StageTimer _timer(VADX_STAGE_PREP, (cudaStream_t)stream, "alpha_x4_const_kernel", 4.0 * n_streams * n_frames * n_bins * 6.0);
  VADX_REQUIRE(d_near_ri && d_pow_far && d_far_comp && d_w2 && d_x4 && n_streams >= 0 && n_frames >= 1 && n_bins >= 1 &&
                   k >= 1 && t_max >= n_frames,
               "vadx_alpha_x4_const_f32: bad argument (t_max %d must cover n_frames %d)", t_max, n_frames);
  if (n_streams == 0) return VADX_OK;
  alpha_x4_const_kernel<<<g1(n_streams * n_frames * n_bins), 256, 0, (cudaStream_t)stream>>>(
      d_near_ri, d_pow_far, d_far_comp, t_max, n_streams, n_frames, n_bins, k, w1_far, w1_mix, b1, d_w2, b2, d_x4, d_alpha);
  return after_launch("vadx_alpha_x4_const_f32");
}

extern "C" int vadx_istft_ola_f32(const float* d_frames, int64_t ld, int64_t n_streams, int n_frames, int n_fft, int hop,
                                  const float* d_wsum_inv, int n_out, float* d_y, int64_t ldy, void* stream) {
  StageTimer _timer(VADX_STAGE_PREP, (cudaStream_t)stream, "istft_ola_kernel", 4.0 * n_streams * ((double)n_frames * n_fft + n_out));
  VADX_REQUIRE(d_frames && d_wsum_inv && d_y && n_streams >= 0 && n_frames >= 1 && n_fft >= 2 && hop >= 1 && ld >= n_fft &&
                   n_out >= 1 && n_out <= (n_frames - 1) * hop + n_fft - 2 * (n_fft / 2) && ldy >= n_out,
               "vadx_istft_ola_f32: bad argument");
  if (n_streams == 0) return VADX_OK;
  istft_ola_kernel<<<g1(n_streams * n_out), 256, 0, (cudaStream_t)stream>>>(d_frames, ld, n_streams, n_frames, n_fft, hop,
                                                                            d_wsum_inv, n_out, d_y, ldy);
  return after_launch("vadx_istft_ola_f32");
}
