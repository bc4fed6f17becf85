// engine.cu - the whole HRNet forward behind three C-ABI calls (calib_b200.h, "engine"):
//   cal_hrnet_create   architecture walk of HighResolutionNet.__init__ (src/models/hrnet/hrnet.py:255-330,
//                      _make_transition_layer :357-391, _make_stage :412-435, HighResolutionModule
//                      :102-220) from a CalHrnetConfig, eval-mode BatchNorm folded into the convs in
//                      fp64 (y = gamma (conv - mu) / sqrt(var + 1e-5) + beta), weights rounded once to
//                      fp16 and packed into the kernels' layouts, all on the device;
//   cal_hrnet_forward  the layer schedule of HighResolutionNet.forward (hrnet.py:437-511, line/hrnet.py:
//                      185-249) and HighResolutionModule.forward (:222-246): one kernel launch per fused
//                      conv + BN (+ residual) (+ ReLU), one per multi-resolution sum, one for the head;
//                      intermediate tensors come from the stream-ordered allocator and are returned to
//                      it as soon as their last consumer is enqueued;
//   cal_hrnet_destroy
// A host written in any language runs the network through these; the Python mirror (hrnet.py) walks the
// same schedule op by op through the single-op entry points and is kept as the readable twin - the two are
// required to agree bit for bit (tests/test_engine_gpu.py).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#define CAL_TU "engine.cu"
#include "common.cuh"

namespace cal {
namespace engine {

constexpr double BN_EPS = 1e-5;

static inline int pad_to(int c, int m = 64) { return (c + m - 1) / m * m; }

struct Conv {                    // one conv (+BN) of the reference, in state_dict order
  int cin, cout, k, s;
  bool has_bias, has_bn;
  size_t w_off, b_off, bn_off;   // offsets (floats) into the weight blob
  // packed, device
  __half* w = nullptr;           // slice-major for 3x3 stride-1, else K-major
  __half* w_k = nullptr;         // K-major copy (the generic kernel's layout) when w is slice-major
  float* b = nullptr;
  int rows = 0, cin_pad = 0, cout_pad = 0;
  bool slices = false;
};

struct Block { std::vector<int> convs; int ds = -1; };
struct Module {
  std::vector<std::vector<Block>> branches;
  std::vector<std::vector<std::vector<int>>> fuse;   // [i][j] = chain of conv ids (empty when j == i)
};

struct Tensor { __half* p = nullptr; int B = 0, H = 0, W = 0, C = 0; size_t bytes() const { return (size_t)B * H * W * C * 2; } };

struct Net {
  CalHrnetConfig cfg;
  std::vector<Conv> convs;
  int conv1 = -1, conv2 = -1, head1 = -1, head2 = -1;
  std::vector<Block> layer1;
  std::vector<std::vector<int>> transition[3];
  std::vector<Module> stage[3];
  std::vector<int> branch_channels;
  int last_in = 0, upscale = 1;
  size_t blob_floats = 0;
  // stem + head, device
  float* stem_w = nullptr; float* stem_b = nullptr;
  std::vector<Conv> head1_src;      // the first head conv split by source (stem, b0..b3), Cout padded to cpad
  float* head_b1 = nullptr; int cpad = 0;
  __half* head2_w64 = nullptr;
  std::vector<void*> device_allocs;
  long launches = 0;
};

static int expansion(int block_type) { return block_type == 1 ? 4 : 1; }

static int new_conv(Net& n, int cin, int cout, int k, int s, bool bias, bool bn) {
  Conv c{};
  c.cin = cin; c.cout = cout; c.k = k; c.s = s; c.has_bias = bias; c.has_bn = bn;
  c.w_off = n.blob_floats;
  n.blob_floats += (size_t)cout * cin * k * k;
  if (bias) { c.b_off = n.blob_floats; n.blob_floats += cout; }
  if (bn) { c.bn_off = n.blob_floats; n.blob_floats += 4 * (size_t)cout; }     // weight, bias, running_mean, running_var
  n.convs.push_back(c);
  return (int)n.convs.size() - 1;
}

static std::vector<Block> block_list(Net& n, int block_type, int cin, int planes, int count) {
  std::vector<Block> out;
  const int exp = expansion(block_type);
  for (int i = 0; i < count; ++i) {
    const int ci = i == 0 ? cin : planes * exp;
    Block b;
    if (block_type == 1) {
      b.convs.push_back(new_conv(n, ci, planes, 1, 1, false, true));
      b.convs.push_back(new_conv(n, planes, planes, 3, 1, false, true));
      b.convs.push_back(new_conv(n, planes, planes * 4, 1, 1, false, true));
    } else {
      b.convs.push_back(new_conv(n, ci, planes, 3, 1, false, true));
      b.convs.push_back(new_conv(n, planes, planes, 3, 1, false, true));
    }
    if (i == 0 && ci != planes * exp) b.ds = new_conv(n, ci, planes * exp, 1, 1, false, true);
    out.push_back(b);
  }
  return out;
}

// the architecture, conv ids in the reference's state_dict order
static int walk(Net& n) {
  const CalHrnetConfig& c = n.cfg;
  const int sw = c.stem_width;
  n.conv1 = new_conv(n, 3, sw, 3, 2, false, true);
  n.conv2 = new_conv(n, sw, sw, 3, 2, false, true);
  const CalHrnetStage& s1 = c.stage[0];
  n.layer1 = block_list(n, s1.block_type, 64, s1.num_channels[0], s1.num_blocks[0]);
  std::vector<int> pre = {expansion(s1.block_type) * s1.num_channels[0]};
  for (int idx = 1; idx < 4; ++idx) {
    const CalHrnetStage& sc = c.stage[idx];
    const int exp = expansion(sc.block_type);
    std::vector<int> ch;
    for (int i = 0; i < sc.num_branches; ++i) ch.push_back(sc.num_channels[i] * exp);
    std::vector<std::vector<int>>& tr = n.transition[idx - 1];
    for (int i = 0; i < (int)ch.size(); ++i) {
      std::vector<int> chain;
      if (i < (int)pre.size()) {
        if (ch[i] != pre[i]) chain.push_back(new_conv(n, pre[i], ch[i], 3, 1, false, true));
      } else {
        for (int j = 0; j < i + 1 - (int)pre.size(); ++j) {
          const int co = (j == i - (int)pre.size()) ? ch[i] : pre.back();
          chain.push_back(new_conv(n, pre.back(), co, 3, 2, false, true));
        }
      }
      tr.push_back(chain);
    }
    for (int m = 0; m < sc.num_modules; ++m) {
      Module mod;
      for (int b = 0; b < (int)ch.size(); ++b) mod.branches.push_back(block_list(n, sc.block_type, ch[b], ch[b] / exp, sc.num_blocks[b]));
      for (int i = 0; i < (int)ch.size(); ++i) {
        std::vector<std::vector<int>> row;
        for (int j = 0; j < (int)ch.size(); ++j) {
          std::vector<int> chain;
          if (j > i) {
            chain.push_back(new_conv(n, ch[j], ch[i], 1, 1, false, true));
          } else if (j < i) {
            for (int k = 0; k < i - j; ++k) chain.push_back(new_conv(n, ch[j], (k == i - j - 1) ? ch[i] : ch[j], 3, 2, false, true));
          }
          row.push_back(chain);
        }
        mod.fuse.push_back(row);
      }
      n.stage[idx - 1].push_back(mod);
    }
    pre = ch;
  }
  n.upscale = (c.kind == 0 && c.upscale > 1) ? c.upscale : 1;
  n.last_in = (n.upscale > 1 ? sw : 0);
  for (int v : pre) n.last_in += v;
  n.branch_channels = pre;
  n.head1 = new_conv(n, n.last_in, n.last_in, 1, 1, true, true);
  n.head2 = new_conv(n, n.last_in, c.num_classes, 1, 1, true, false);
  return CAL_OK;
}

// ---------------------------------------------------------------- weight preparation (host)
struct Folded { std::vector<double> w, b; };       // w: (cout, cin, k, k)

static Folded fold(const Conv& c, const float* blob) {
  Folded f;
  const size_t per = (size_t)c.cin * c.k * c.k;
  f.w.resize((size_t)c.cout * per);
  f.b.assign(c.cout, 0.0);
  for (size_t i = 0; i < f.w.size(); ++i) f.w[i] = blob[c.w_off + i];
  if (c.has_bias) for (int o = 0; o < c.cout; ++o) f.b[o] = blob[c.b_off + o];
  if (c.has_bn) {
    const float* g = blob + c.bn_off, *beta = g + c.cout, *mu = beta + c.cout, *var = mu + c.cout;
    for (int o = 0; o < c.cout; ++o) {
      const double s = (double)g[o] / sqrt((double)var[o] + BN_EPS);
      for (size_t i = 0; i < per; ++i) f.w[o * per + i] *= s;
      f.b[o] = (f.b[o] - (double)mu[o]) * s + (double)beta[o];
    }
  }
  return f;
}

template <typename T>
static int to_device(Net& n, const std::vector<T>& h, T** d) {
  void* p = nullptr;
  CAL_CHECK_CUDA(cudaMalloc(&p, h.size() * sizeof(T)));
  n.device_allocs.push_back(p);
  CAL_CHECK_CUDA(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  *d = static_cast<T*>(p);
  return CAL_OK;
}

// (cout, cin_sel, k, k) fp64 -> fp16 (rows, k*k, cin_pad) K-major [+ slice-major], fp32 bias (cout_pad)
static int pack(Net& n, Conv& c, const double* w, int cout, int cin, int cin_stride, int cin_off, const double* b, int cout_pad_force,
                bool want_slices) {
  const int kk = c.k * c.k;
  c.cin_pad = pad_to(cin);
  c.cout_pad = cout_pad_force > 0 ? cout_pad_force : pad_to(cout);
  c.rows = pad_to(cout, 16);
  std::vector<__half> wk((size_t)c.rows * kk * c.cin_pad, __float2half(0.0f));
  for (int o = 0; o < cout; ++o)
    for (int ci = 0; ci < cin; ++ci)
      for (int t = 0; t < kk; ++t)
        // double -> float -> half, the two roundings torch's .to(torch.float16) performs on a float64 tensor (the
        // Python packer, packing.pack_conv): the engine and its twin must produce the same bits
        wk[((size_t)o * kk + t) * c.cin_pad + ci] = __float2half((float)w[((size_t)o * cin_stride + cin_off + ci) * kk + t]);
  std::vector<float> bp(c.cout_pad, 0.0f);
  for (int o = 0; o < cout; ++o) bp[o] = (float)b[o];
  int rc = to_device(n, bp, &c.b);
  if (rc != CAL_OK) return rc;
  c.slices = want_slices;
  if (want_slices) {
    // (rows, kk*cin_pad/64, 64) -> (kk*cin_pad/64, rows, 64)
    const int ns = kk * c.cin_pad / 64;
    std::vector<__half> ws(wk.size());
    for (int o = 0; o < c.rows; ++o)
      for (int s = 0; s < ns; ++s)
        memcpy(&ws[((size_t)s * c.rows + o) * 64], &wk[((size_t)o * ns + s) * 64], 64 * sizeof(__half));
    rc = to_device(n, ws, &c.w);
    if (rc != CAL_OK) return rc;
    return to_device(n, wk, &c.w_k);
  }
  return to_device(n, wk, &c.w);
}

static int prepare(Net& n, const float* blob) {
  for (int id = 0; id < (int)n.convs.size(); ++id) {
    if (id == n.conv1 || id == n.head1 || id == n.head2) continue;
    Conv& c = n.convs[id];
    const Folded f = fold(c, blob);
    const int rc = pack(n, c, f.w.data(), c.cout, c.cin, c.cin, 0, f.b.data(), 0, c.k == 3);   // 3x3: slice-major (+ a K-major copy for the generic kernel)
    if (rc != CAL_OK) return rc;
  }
  {  // stem conv1: fp32 (64, 27) [co][ci*9 + ky*3 + kx]
    const Conv& c = n.convs[n.conv1];
    CAL_REQUIRE(c.cout == 64, CAL_E_UNSUPPORTED, "cal_hrnet_create: stem_width %d (the stem kernel is built for 64)", c.cout);
    const Folded f = fold(c, blob);
    std::vector<float> w(f.w.size()), b(f.b.size());
    for (size_t i = 0; i < w.size(); ++i) w[i] = (float)f.w[i];
    for (size_t i = 0; i < b.size(); ++i) b[i] = (float)f.b[i];
    int rc = to_device(n, w, &n.stem_w);
    if (rc != CAL_OK) return rc;
    rc = to_device(n, b, &n.stem_b);
    if (rc != CAL_OK) return rc;
  }
  {  // head: first 1x1 conv split by source in the concat order (hrnet.py:509: stem, branch 0..3)
    const Conv& h1 = n.convs[n.head1];
    const Folded f = fold(h1, blob);
    n.cpad = pad_to(n.last_in);
    std::vector<int> src_c;
    if (n.upscale > 1) src_c.push_back(n.cfg.stem_width);
    for (int v : n.branch_channels) src_c.push_back(v);
    std::vector<double> zero(h1.cout, 0.0);
    int off = 0;
    for (int c : src_c) {
      Conv part{};
      part.cin = c; part.cout = h1.cout; part.k = 1; part.s = 1;
      const int rc = pack(n, part, f.w.data(), h1.cout, c, h1.cin, off, zero.data(), n.cpad, false);
      if (rc != CAL_OK) return rc;
      n.head1_src.push_back(part);
      off += c;
    }
    std::vector<float> b1(n.cpad, 0.0f);
    for (int o = 0; o < h1.cout; ++o) b1[o] = (float)f.b[o];
    int rc = to_device(n, b1, &n.head_b1);
    if (rc != CAL_OK) return rc;
    Conv& h2 = n.convs[n.head2];
    const Folded f2 = fold(h2, blob);
    rc = pack(n, h2, f2.w.data(), h2.cout, h2.cin, h2.cin, 0, f2.b.data(), 64, false);
    if (rc != CAL_OK) return rc;
    CAL_REQUIRE(h2.rows <= 64, CAL_E_UNSUPPORTED, "cal_hrnet_create: num_classes %d > 64", h2.cout);
    // the chained head kernel wants the final conv as a full 64-row K-major tile
    std::vector<__half> w64((size_t)64 * h2.cin_pad, __float2half(0.0f));
    std::vector<__half> tmp((size_t)h2.rows * h2.cin_pad);
    CAL_CHECK_CUDA(cudaMemcpy(tmp.data(), h2.w, tmp.size() * sizeof(__half), cudaMemcpyDeviceToHost));
    memcpy(w64.data(), tmp.data(), tmp.size() * sizeof(__half));
    rc = to_device(n, w64, &n.head2_w64);
    if (rc != CAL_OK) return rc;
  }
  return CAL_OK;
}

// ---------------------------------------------------------------- forward
struct Run {
  Net& n;
  cudaStream_t st;
  int rc = CAL_OK;
  std::vector<void*> live;
  Run(Net& net, cudaStream_t s) : n(net), st(s) {}

  Tensor alloc(int B, int H, int W, int C) {
    Tensor t; t.B = B; t.H = H; t.W = W; t.C = C;
    if (rc != CAL_OK) return t;
    void* p = nullptr;
    const cudaError_t e = cudaMallocAsync(&p, t.bytes(), st);
    if (e != cudaSuccess) { set_error("cudaMallocAsync(%zu bytes) failed: %s", t.bytes(), cudaGetErrorString(e)); rc = CAL_E_CUDA; return t; }
    t.p = static_cast<__half*>(p);
    live.push_back(p);
    return t;
  }
  void release(Tensor& t) {
    if (!t.p) return;
    for (size_t i = 0; i < live.size(); ++i)
      if (live[i] == t.p) { live.erase(live.begin() + i); break; }
    cudaFreeAsync(t.p, st);
    t.p = nullptr;
  }
  void release_all() {
    for (void* p : live) cudaFreeAsync(p, st);
    live.clear();
  }

  Tensor conv(const Tensor& x, Conv& c, bool relu, const Tensor* res) {
    const int pad = c.k / 2;
    const int Ho = (x.H + 2 * pad - c.k) / c.s + 1, Wo = (x.W + 2 * pad - c.k) / c.s + 1;
    Tensor y = alloc(x.B, Ho, Wo, c.cout_pad);
    if (rc != CAL_OK) return y;
    CalConvArgs a{};
    a.x = x.p; a.bias = c.b; a.res = res ? res->p : nullptr; a.y = y.p;
    a.B = x.B; a.Hin = x.H; a.Win = x.W; a.Cin_pad = x.C; a.Hout = Ho; a.Wout = Wo; a.Cout_pad = c.cout_pad; a.Cout_rows = c.rows;
    a.ksize = c.k; a.stride = c.s; a.relu = relu ? 1 : 0; a.mode = 0; a.n_classes = 0; a.Cin = c.cin;
    if (x.C != c.cin_pad) { set_error("cal_hrnet_forward: channel mismatch (%d vs %d)", x.C, c.cin_pad); rc = CAL_E_INVALID; return y; }
    ++n.launches;
    if (c.slices) {
      a.w = c.w; a.w_slices = 1;
      const int r = cal_conv2d(&a, st);
      if (r == CAL_OK) return y;
      if (r != CAL_E_UNSUPPORTED) { rc = r; return y; }
      c.slices = false;                          // this shape is served by the generic kernel: K-major from now on
      c.w = c.w_k;
    }
    a.w = c.w; a.w_slices = 0;
    const int r = cal_conv2d(&a, st);
    if (r != CAL_OK) rc = r;
    return y;
  }

  // Both convs of a BasicBlock of the full-resolution branch in one kernel (basicblock.cu); false when the shape
  // is served by the two launches.
  bool basicblock(const Tensor& x, Block& b, Tensor* out) {
    static const bool enabled = [] { const char* e = getenv("CAL_BASICBLOCK"); return !(e && e[0] == '0'); }();
    if (!enabled || b.ds >= 0 || b.convs.size() != 2) return false;
    Conv& c1 = n.convs[b.convs[0]];
    Conv& c2 = n.convs[b.convs[1]];
    for (const Conv* c : {&c1, &c2})
      if (!(c->slices && c->k == 3 && c->s == 1 && c->cin_pad == 64 && c->cout_pad == 64 && c->rows <= 48)) return false;
    if (x.C != 64 || c1.rows != c2.rows || c1.cin != c2.cin) return false;
    Tensor y = alloc(x.B, x.H, x.W, 64);
    if (rc != CAL_OK) return false;
    CalBasicBlockArgs a{};
    a.x = x.p; a.w1 = c1.w; a.bias1 = c1.b; a.w2 = c2.w; a.bias2 = c2.b; a.y = y.p;
    a.B = x.B; a.H = x.H; a.W = x.W; a.C_pad = 64; a.rows = c1.rows; a.C = c1.cin;
    const int r = cal_basicblock(&a, st);
    if (r == CAL_OK) { ++n.launches; *out = y; return true; }
    release(y);
    if (r != CAL_E_UNSUPPORTED) rc = r;
    return false;
  }

  Tensor blocks(Tensor x, std::vector<Block>& bl, bool own_x) {
    for (Block& b : bl) {
      {
        Tensor y;
        if (basicblock(x, b, &y)) {
          if (own_x) release(x);
          x = y; own_x = true;
          continue;
        }
        if (rc != CAL_OK) break;
      }
      Tensor r = x;
      bool own_r = false;
      if (b.ds >= 0) { r = conv(x, n.convs[b.ds], false, nullptr); own_r = true; }
      Tensor t = x;
      bool own_t = false;
      for (size_t i = 0; i + 1 < b.convs.size(); ++i) {
        Tensor u = conv(t, n.convs[b.convs[i]], true, nullptr);
        if (own_t) release(t);
        t = u; own_t = true;
      }
      Tensor y = conv(t, n.convs[b.convs.back()], true, &r);
      if (own_t) release(t);
      if (own_r) release(r);
      if (own_x) release(x);
      x = y; own_x = true;
      if (rc != CAL_OK) break;
    }
    return x;
  }

  int combine(Tensor& y, const std::vector<Tensor>& srcs, const float* bias, bool relu, int c_real = 0) {
    CalCombineArgs a{};
    a.C = c_real;
    a.y = y.p; a.B = y.B; a.H = y.H; a.W = y.W; a.C_pad = y.C; a.n_src = (int)srcs.size();
    for (size_t i = 0; i < srcs.size(); ++i) { a.src[i] = srcs[i].p; a.src_h[i] = srcs[i].H; a.src_w[i] = srcs[i].W; }
    a.bias = bias; a.relu = relu ? 1 : 0;
    ++n.launches;
    return cal_fuse_combine(&a, st);
  }

  // HighResolutionModule.forward (hrnet.py:222-246); takes ownership of xs
  std::vector<Tensor> module(std::vector<Tensor> xs, Module& m) {
    for (size_t b = 0; b < xs.size(); ++b) xs[b] = blocks(xs[b], m.branches[b], true);
    const int nb = (int)xs.size();
    std::vector<Tensor> outs;
    for (int i = 0; i < nb && rc == CAL_OK; ++i) {
      // same-resolution terms chained through conv epilogues: x_i + sum_{j<i} down_ij(x_j)
      const bool has_up = i < nb - 1;
      Tensor r = xs[i];
      bool own_r = false;
      for (int j = 0; j < i; ++j) {
        Tensor t = xs[j];
        bool own_t = false;
        std::vector<int>& chain = m.fuse[i][j];
        for (size_t k = 0; k + 1 < chain.size(); ++k) {
          Tensor u = conv(t, n.convs[chain[k]], true, nullptr);
          if (own_t) release(t);
          t = u; own_t = true;
        }
        const bool last_term = (j == i - 1);
        Tensor u = conv(t, n.convs[chain.back()], last_term && !has_up, &r);
        if (own_t) release(t);
        if (own_r) release(r);
        r = u; own_r = true;
      }
      if (has_up) {
        std::vector<Tensor> srcs = {r};
        for (int j = i + 1; j < nb; ++j) srcs.push_back(conv(xs[j], n.convs[m.fuse[i][j][0]], false, nullptr));
        Tensor y = alloc(xs[i].B, xs[i].H, xs[i].W, xs[i].C);
        if (rc == CAL_OK) { const int r2 = combine(y, srcs, nullptr, true, n.convs[m.fuse[i][i + 1][0]].cout); if (r2 != CAL_OK) rc = r2; }
        for (size_t q = 1; q < srcs.size(); ++q) release(srcs[q]);
        if (own_r) release(r);
        r = y; own_r = true;
      }
      if (!own_r) {               // nb == 1 never happens in a fused module; keep the input alive as the output
        outs.push_back(r);
        xs[i].p = nullptr;
      } else {
        outs.push_back(r);
      }
    }
    for (Tensor& t : xs) release(t);
    return outs;
  }

  int head(Tensor& stem, std::vector<Tensor>& ys, float* heat) {
    const int up = n.upscale;
    const int h = ys[0].H * up, w = ys[0].W * up, B = ys[0].B;
    Tensor full;
    bool own_full = false;
    size_t first_low = 0;
    const Conv* pf;
    if (up > 1) {
      full = stem; pf = &n.head1_src[0];
      if (stem.H != h || stem.W != w) {
        // the reference resamples x_stem to the head size (hrnet.py:494-498)
        full = alloc(B, h, w, stem.C);
        own_full = true;
        if (rc != CAL_OK) return rc;
        const int r = combine(full, {stem}, nullptr, false);
        if (r != CAL_OK) return r;
      }
    } else {
      full = ys[0]; pf = &n.head1_src[0]; first_low = 1;
    }
    std::vector<Tensor> proj;
    for (size_t i = first_low; i < ys.size(); ++i) {
      Conv& c = n.head1_src[(up > 1 ? 1 : 0) + i];
      proj.push_back(conv(ys[i], c, false, nullptr));
      if (rc != CAL_OK) return rc;
    }
    Conv& h2 = n.convs[n.head2];
    const int mode = n.cfg.kind == 0 ? 1 : 2;
    bool done = false;
    if (full.C == 64 && proj.size() >= 1 && proj.size() <= CAL_MAX_LOW) {
      CalHeadArgs a{};
      a.full = full.p; a.w_full = pf->w;
      for (size_t i = 0; i < proj.size(); ++i) { a.low[i] = proj[i].p; a.low_h[i] = proj[i].H; a.low_w[i] = proj[i].W; }
      a.n_low = (int)proj.size(); a.bias = n.head_b1; a.z = nullptr;
      a.B = B; a.H = h; a.W = w; a.Cf_pad = full.C; a.Cout_pad = n.cpad; a.Cout_rows = pf->rows;
      a.w2 = n.head2_w64; a.bias2 = h2.b; a.heat = heat; a.n_classes = n.cfg.num_classes; a.mode = mode;
      ++n.launches;
      int r = cal_head_fused(&a, st);
      if (r == CAL_OK) done = true;
      else if (r != CAL_E_UNSUPPORTED) return r;
      else {
        --n.launches;
        // unchained: z = relu(W1_full full + sum up(p_i) + b) to HBM, then the final conv
        Tensor z = alloc(B, h, w, n.cpad);
        if (rc != CAL_OK) return rc;
        a.w2 = nullptr; a.bias2 = nullptr; a.heat = nullptr; a.z = z.p; a.n_classes = 0; a.mode = 0;
        ++n.launches;
        r = cal_head_fused(&a, st);
        if (r == CAL_OK) {
          r = final_conv(z, h2, heat, mode);
          release(z);
          if (r != CAL_OK) return r;
          done = true;
        } else {
          --n.launches;
          release(z);
          if (r != CAL_E_UNSUPPORTED) return r;
        }
      }
    }
    if (!done) {
      // generic path: u = sum up(p_i) + b1; z = relu(W1_full full + u); heat = softmax(W2 z + b2)
      Tensor u = alloc(B, h, w, n.cpad);
      if (rc != CAL_OK) return rc;
      int r = combine(u, proj, n.head_b1, false);
      if (r != CAL_OK) return r;
      Conv cf = *pf;                               // (bias-free part of the first head conv on the full-resolution source)
      Tensor z = conv(full, cf, true, &u);
      release(u);
      if (rc != CAL_OK) return rc;
      r = final_conv(z, h2, heat, mode);
      release(z);
      if (r != CAL_OK) return r;
    }
    for (Tensor& t : proj) release(t);
    if (own_full) release(full);
    return CAL_OK;
  }

  int final_conv(const Tensor& z, Conv& h2, float* heat, int mode) {
    CalConvArgs a{};
    a.x = z.p; a.w = h2.w; a.bias = h2.b; a.res = nullptr; a.y = heat;
    a.B = z.B; a.Hin = z.H; a.Win = z.W; a.Cin_pad = z.C; a.Hout = z.H; a.Wout = z.W; a.Cout_pad = 64; a.Cout_rows = h2.rows;
    a.ksize = 1; a.stride = 1; a.relu = 0; a.mode = mode; a.n_classes = n.cfg.num_classes; a.Cin = h2.cin; a.w_slices = 0;
    ++n.launches;
    return cal_conv2d(&a, st);
  }

  int forward(const void* x, int x_u8, int B, int H, int W, float* heat) {
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    Tensor stem = alloc(B, Ho, Wo, 64);
    if (rc != CAL_OK) return rc;
    ++n.launches;
    int r = x_u8 ? cal_stem_conv_u8(static_cast<const uint8_t*>(x), n.stem_w, n.stem_b, stem.p, B, H, W, Ho, Wo, st)
                 : cal_stem_conv(static_cast<const float*>(x), n.stem_w, n.stem_b, stem.p, B, H, W, Ho, Wo, st);
    if (r != CAL_OK) return r;
    Tensor t = conv(stem, n.convs[n.conv2], true, nullptr);
    t = blocks(t, n.layer1, true);
    std::vector<Tensor> ys = {t};
    for (int idx = 0; idx < 3 && rc == CAL_OK; ++idx) {
      std::vector<std::vector<int>>& tr = n.transition[idx];
      const size_t nprev = ys.size();
      std::vector<Tensor> xs;
      std::vector<bool> moved(nprev, false);
      for (size_t i = 0; i < tr.size(); ++i) {
        if (tr[i].empty()) { xs.push_back(ys[i]); moved[i] = true; continue; }
        Tensor v = i < nprev ? ys[i] : ys[nprev - 1];
        bool own = false;
        for (int cid : tr[i]) {
          Tensor u = conv(v, n.convs[cid], true, nullptr);
          if (own) release(v);
          v = u; own = true;
        }
        xs.push_back(v);
      }
      for (size_t i = 0; i < nprev; ++i)
        if (!moved[i]) release(ys[i]);
      for (Module& m : n.stage[idx]) xs = module(xs, m);
      ys = xs;
    }
    if (rc != CAL_OK) return rc;
    r = head(stem, ys, heat);
    if (r != CAL_OK) return r;
    release(stem);
    for (Tensor& y : ys) release(y);
    return rc;
  }
};

}  // namespace engine
}  // namespace cal

extern "C" int cal_hrnet_weight_count(const CalHrnetConfig* cfg, size_t* n_floats) {
  using namespace cal;
  CAL_REQUIRE(cfg && n_floats, CAL_E_INVALID, "cal_hrnet_weight_count: null pointer");
  engine::Net n;
  n.cfg = *cfg;
  engine::walk(n);
  *n_floats = n.blob_floats;
  return CAL_OK;
}

extern "C" int cal_hrnet_create(const CalHrnetConfig* cfg, const float* h_weights, size_t n_floats, void** handle) {
  using namespace cal;
  CAL_REQUIRE(cfg && h_weights && handle, CAL_E_INVALID, "cal_hrnet_create: null pointer");
  CAL_REQUIRE(cfg->kind == 0 || cfg->kind == 1, CAL_E_INVALID, "cal_hrnet_create: kind %d", cfg->kind);
  CAL_REQUIRE(cfg->num_classes >= 1 && cfg->num_classes <= 64 && cfg->stem_width == 64, CAL_E_UNSUPPORTED,
              "cal_hrnet_create: num_classes %d (<= 64) / stem_width %d (64)", cfg->num_classes, cfg->stem_width);
  for (int s = 0; s < 4; ++s) {
    const CalHrnetStage& st = cfg->stage[s];
    CAL_REQUIRE(st.num_modules >= 1 && st.num_branches == s + 1 && (st.block_type == 0 || st.block_type == 1), CAL_E_INVALID,
                "cal_hrnet_create: stage %d: modules %d branches %d block %d", s + 1, st.num_modules, st.num_branches, st.block_type);
    for (int b = 0; b < st.num_branches; ++b)
      CAL_REQUIRE(st.num_blocks[b] >= 1 && st.num_channels[b] >= 1, CAL_E_INVALID, "cal_hrnet_create: stage %d branch %d", s + 1, b);
  }
  engine::Net* n = new engine::Net();
  n->cfg = *cfg;
  engine::walk(*n);
  if (n->blob_floats != n_floats) {
    set_error("cal_hrnet_create: the weight blob has %zu floats, this architecture needs %zu", n_floats, n->blob_floats);
    delete n;
    return CAL_E_INVALID;
  }
  {
    // keep freed activation memory in the stream-ordered pool between forwards
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  const int rc = engine::prepare(*n, h_weights);
  if (rc != CAL_OK) {
    for (void* p : n->device_allocs) cudaFree(p);
    delete n;
    return rc;
  }
  *handle = n;
  return CAL_OK;
}

extern "C" int cal_hrnet_forward(void* handle, const void* x, int x_is_u8, int B, int H, int W, float* heat, void* stream) {
  using namespace cal;
  CAL_REQUIRE(handle && x && heat, CAL_E_INVALID, "cal_hrnet_forward: null pointer");
  CAL_REQUIRE(B >= 1 && H >= 8 && W >= 8, CAL_E_INVALID, "cal_hrnet_forward: bad shape %d x %d x %d", B, H, W);
  engine::Net& n = *static_cast<engine::Net*>(handle);
  engine::Run run(n, static_cast<cudaStream_t>(stream));
  int rc = run.forward(x, x_is_u8, B, H, W, heat);
  if (rc == CAL_OK) rc = run.rc;
  run.release_all();
  return rc;
}

extern "C" int cal_hrnet_output_shape(void* handle, int H, int W, int* n_classes, int* h, int* w) {
  using namespace cal;
  CAL_REQUIRE(handle && n_classes && h && w, CAL_E_INVALID, "cal_hrnet_output_shape: null pointer");
  const engine::Net& n = *static_cast<engine::Net*>(handle);
  const int h1 = (H - 1) / 2 + 1, w1 = (W - 1) / 2 + 1;
  *n_classes = n.cfg.num_classes;
  *h = ((h1 - 1) / 2 + 1) * n.upscale;
  *w = ((w1 - 1) / 2 + 1) * n.upscale;
  return CAL_OK;
}

extern "C" long cal_hrnet_launches(void* handle) { return handle ? static_cast<cal::engine::Net*>(handle)->launches : -1; }

extern "C" int cal_hrnet_destroy(void* handle) {
  if (!handle) return CAL_OK;
  cal::engine::Net* n = static_cast<cal::engine::Net*>(handle);
  for (void* p : n->device_allocs) cudaFree(p);
  delete n;
  return CAL_OK;
}
