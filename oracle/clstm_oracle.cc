// clstm_oracle.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see clstm_oracle.h header comment).
//
// Plain C++17, zero dependencies.  A restatement -- not a copy -- of the reference math:
// each routine names the reference lines it follows (paths relative to /root/reference).
// Where the reference has an Eigen::Tensor branch and an Eigen::Matrix "#else" branch the
// Matrix branch is followed, because it states the arithmetic unambiguously.
//
// Parity pin: CTC pinned by test-ctc.cc KATs; LSTM/Softmax "parity unpinned" (no golden
// vectors exist upstream; Eigen absent so the reference cannot be built here).
#include "clstm_oracle.h"

#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <atomic>
#include <thread>

namespace {

// ------------------------------------------------------------------------------------------
// storage: Tensor2 / Batch / Sequence   (tensor.h:176-335, batches.h:12-148)
// ------------------------------------------------------------------------------------------
template <class F>
struct Mat {  // column-major, element (i,j) at a[i + j*r]  (tensor.h:252,288)
  int r = 0, c = 0;
  std::vector<F> a;
  void resize(int n, int m) {  // setZero(n,m): batches.h:37-41
    r = n; c = m;
    a.assign((size_t)n * m, F(0));
  }
  void zero() { std::fill(a.begin(), a.end(), F(0)); }
  F& operator()(int i, int j) { return a[(size_t)i + (size_t)j * r]; }
  F operator()(int i, int j) const { return a[(size_t)i + (size_t)j * r]; }
  F* col(int j) { return a.data() + (size_t)j * r; }
  const F* col(int j) const { return a.data() + (size_t)j * r; }
};

template <class F>
struct Batch {  // batches.h:12-25
  Mat<F> v, d;
  int rows() const { return v.r; }
  int cols() const { return v.c; }
  void resize(int n, int m) { v.resize(n, m); d.resize(n, m); }
  void zeroGrad() { d.resize(v.r, v.c); }
};

template <class F>
struct Seq {  // batches.h:45-148; resize() always zero-fills (batches.h:127)
  std::vector<Batch<F>> steps;
  int n = 0, m = 0;
  int size() const { return (int)steps.size(); }
  int rows() const { return n; }
  int cols() const { return m; }
  void resize(int N, int rows, int cols) {
    n = rows; m = cols;
    steps.resize(N);
    for (auto& s : steps) s.resize(rows, cols);
  }
  void like(const Seq& o) { resize(o.size(), o.n, o.m); }
  void copy(const Seq& o) { steps = o.steps; n = o.n; m = o.m; }  // deep copy v and d (batches.h:133-139)
  Batch<F>& operator[](int t) { return steps[t]; }
  void zeroGrad() { for (auto& s : steps) s.zeroGrad(); }
};

// ------------------------------------------------------------------------------------------
// scalar helpers  (tensor.h:70-91, 337-374)
// ------------------------------------------------------------------------------------------
template <class F> inline F limexp(F x) {  // tensor.h:78-82, MAXEXP=30
  if (x < F(-30)) return std::exp(F(-30));
  if (x > F(30)) return std::exp(F(30));
  return std::exp(x);
}
template <class F> inline F log_add(F x, F y) {  // tensor.h:86-89
  if (std::fabs(x - y) > F(10)) return std::fmax(x, y);
  return std::log(std::exp(x - y) + F(1)) + y;
}
template <class F> inline int argmax1(const F* m, int n) {  // tensor.h:357-366 (ties -> last)
  int mi = -1;
  F mv = m[0];
  for (int i = 0; i < n; i++) {
    if (m[i] < mv) continue;
    mi = i;
    mv = m[i];
  }
  return mi;
}

enum { LIN = 0, SIG = 1, TANH = 2, RELU = 3 };   // clstm_compute.h nonlinearity codes used here

// ------------------------------------------------------------------------------------------
// compute ops  (clstm_compute.cc)
// ------------------------------------------------------------------------------------------
// forward_lin1: y = W[:,1:] * x + W[:,0]   (clstm_compute.cc:275-293, Matrix branch :291)
template <class F>
void forward_lin1(Batch<F>& y, const Batch<F>& W, const Batch<F>& x) {
  const int n = W.v.r, m = W.v.c, bs = x.v.c;
  assert(y.v.r == n && y.v.c == bs && x.v.r == m - 1);
  for (int b = 0; b < bs; b++) {
    F* yb = y.v.col(b);
    const F* w0 = W.v.col(0);
    for (int i = 0; i < n; i++) yb[i] = F(0);
    const F* xb = x.v.col(b);
    for (int j = 0; j < m - 1; j++) {
      const F* wj = W.v.col(j + 1);
      const F xv = xb[j];
      for (int i = 0; i < n; i++) yb[i] += wj[i] * xv;
    }
    for (int i = 0; i < n; i++) yb[i] += w0[i];
  }
}
// backward_lin1 (clstm_compute.cc:294-304, Matrix branch :300-302)
template <class F>
void backward_lin1(Batch<F>& y, Batch<F>& W, Batch<F>& x) {
  const int n = W.v.r, m = W.v.c, bs = x.v.c;
  for (int b = 0; b < bs; b++) {  // x.d += W[:,1:]^T * y.d
    const F* yd = y.d.col(b);
    F* xd = x.d.col(b);
    for (int j = 0; j < m - 1; j++) {
      const F* wj = W.v.col(j + 1);
      F acc = F(0);
      for (int i = 0; i < n; i++) acc += wj[i] * yd[i];
      xd[j] += acc;
    }
  }
  for (int b = 0; b < bs; b++) {  // W.d[:,1:] += y.d * x.v^T ; W.d[:,0] += rowsum(y.d)
    const F* yd = y.d.col(b);
    const F* xv = x.v.col(b);
    for (int j = 0; j < m - 1; j++) {
      F* wdj = W.d.col(j + 1);
      const F xj = xv[j];
      for (int i = 0; i < n; i++) wdj[i] += yd[i] * xj;
    }
    F* wd0 = W.d.col(0);
    for (int i = 0; i < n; i++) wd0[i] += yd[i];
  }
}
// forward_nonlin0 (clstm_compute.cc:209-229): sigmoid = 1/(1+exp(-x)) unclamped (:116-118)
template <class F>
void forward_nonlin0(Batch<F>& y, int nl) {
  F* p = y.v.a.data();
  const size_t N = y.v.a.size();
  if (nl == SIG) for (size_t i = 0; i < N; i++) p[i] = F(1) / (F(1) + std::exp(-p[i]));
  else if (nl == TANH) for (size_t i = 0; i < N; i++) p[i] = std::tanh(p[i]);
  else if (nl == RELU) for (size_t i = 0; i < N; i++) p[i] = std::max(p[i], F(0));   // forward_relu :123-125
}
// backward_nonlin0 (clstm_compute.cc:231-267): in place on d
template <class F>
void backward_nonlin0(Batch<F>& y, int nl) {
  F* d = y.d.a.data();
  const F* v = y.v.a.data();
  const size_t N = y.v.a.size();
  if (nl == SIG) for (size_t i = 0; i < N; i++) d[i] = v[i] * (-v[i] + F(1)) * d[i];
  else if (nl == TANH) for (size_t i = 0; i < N; i++) d[i] = (-v[i] * v[i] + F(1)) * d[i];
  else if (nl == RELU) for (size_t i = 0; i < N; i++) d[i] = d[i] * (v[i] > F(0) ? F(1) : F(0));   // backward_relu0 :241-244
}
template <class F> void forward_full1(Batch<F>& y, const Batch<F>& W, const Batch<F>& x, int nl) {  // :308-314
  forward_lin1(y, W, x);
  forward_nonlin0(y, nl);
}
template <class F> void backward_full1(Batch<F>& y, Batch<F>& W, Batch<F>& x, int nl) {  // :316-320
  backward_nonlin0(y, nl);
  backward_lin1(y, W, x);
}
// forward_softmax (clstm_compute.cc:324-345, Matrix branch :340-343): limexp, no max-subtraction
template <class F>
void forward_softmax(Batch<F>& z, const Batch<F>& W, const Batch<F>& x) {
  forward_lin1(z, W, x);
  const int n = z.v.r, bs = z.v.c;
  for (int b = 0; b < bs; b++) {
    F* zb = z.v.col(b);
    for (int i = 0; i < n; i++) zb[i] = limexp(zb[i]);
    F sum = F(0);
    for (int i = 0; i < n; i++) sum += zb[i];
    for (int i = 0; i < n; i++) zb[i] /= sum;
  }
}
// backward_softmax (clstm_compute.cc:346-356): x.d ASSIGNED, no Jacobian
template <class F>
void backward_softmax(Batch<F>& z, Batch<F>& W, Batch<F>& x) {
  x.d.zero();
  backward_lin1(z, W, x);
}
// forward_stack / backward_stack (clstm_compute.cc:360-373)
template <class F>
void forward_stack(Batch<F>& z, const Batch<F>& x, const Batch<F>& y) {
  const int nx = x.v.r, ny = y.v.r, bs = x.v.c;
  for (int b = 0; b < bs; b++) {
    for (int i = 0; i < nx; i++) z.v(i, b) = x.v(i, b);
    for (int i = 0; i < ny; i++) z.v(nx + i, b) = y.v(i, b);
  }
}
template <class F>
void backward_stack(const Batch<F>& z, Batch<F>& x, Batch<F>& y) {
  const int nx = x.v.r, ny = y.v.r, bs = x.v.c;
  for (int b = 0; b < bs; b++) {
    for (int i = 0; i < nx; i++) x.d(i, b) += z.d(i, b);
    for (int i = 0; i < ny; i++) y.d(i, b) += z.d(nx + i, b);
  }
}
// forward_stack_delay / backward_stack_delay (clstm_compute.cc:377-410)
template <class F>
void forward_stack_delay(Batch<F>& z, const Batch<F>& x, Seq<F>& y, int last) {
  const int nx = x.v.r, ny = y.rows(), bs = x.v.c;
  for (int b = 0; b < bs; b++) {
    for (int i = 0; i < nx; i++) z.v(i, b) = x.v(i, b);
    for (int i = 0; i < ny; i++) z.v(nx + i, b) = last >= 0 ? y[last].v(i, b) : F(0);
  }
}
template <class F>
void backward_stack_delay(const Batch<F>& z, Batch<F>& x, Seq<F>& y, int last) {
  const int nx = x.v.r, ny = y.rows(), bs = x.v.c;
  for (int b = 0; b < bs; b++) {
    for (int i = 0; i < nx; i++) x.d(i, b) += z.d(i, b);
    if (last >= 0)
      for (int i = 0; i < ny; i++) y[last].d(i, b) += z.d(nx + i, b);
  }
}
// forward_statemem / backward_statemem (clstm_compute.cc:504-515)
template <class F>
void forward_statemem(Batch<F>& state, const Batch<F>& ci, const Batch<F>& gi, Seq<F>& states, int last,
                      const Batch<F>& gf) {
  const size_t N = state.v.a.size();
  for (size_t i = 0; i < N; i++) state.v.a[i] = ci.v.a[i] * gi.v.a[i];
  if (last >= 0)
    for (size_t i = 0; i < N; i++) state.v.a[i] += gf.v.a[i] * states[last].v.a[i];
}
template <class F>
void backward_statemem(Batch<F>& state, Batch<F>& ci, Batch<F>& gi, Seq<F>& states, int last, Batch<F>& gf) {
  const size_t N = state.v.a.size();
  if (last >= 0)
    for (size_t i = 0; i < N; i++) states[last].d.a[i] += state.d.a[i] * gf.v.a[i];
  if (last >= 0)
    for (size_t i = 0; i < N; i++) gf.d.a[i] += state.d.a[i] * states[last].v.a[i];
  for (size_t i = 0; i < N; i++) gi.d.a[i] += state.d.a[i] * ci.v.a[i];
  for (size_t i = 0; i < N; i++) ci.d.a[i] += state.d.a[i] * gi.v.a[i];
}
// forward_nonlingate / backward_nonlingate (clstm_compute.cc:519-547), H = nl (TANH for NPLSTM; LIN / RELU variants)
template <class F> inline F nonlin_h(F x, int nl) { return nl == TANH ? std::tanh(x) : (nl == RELU ? std::max(x, F(0)) : x); }
template <class F>
void forward_nonlingate(Batch<F>& out, const Batch<F>& state, const Batch<F>& go, int nl = TANH) {
  const size_t N = out.v.a.size();
  for (size_t i = 0; i < N; i++) out.v.a[i] = nonlin_h(state.v.a[i], nl) * go.v.a[i];
}
template <class F>
void backward_nonlingate(const Batch<F>& out, Batch<F>& state, Batch<F>& go, int nl = TANH) {
  const size_t N = out.v.a.size();
  for (size_t i = 0; i < N; i++) {
    const F th = nonlin_h(state.v.a[i], nl);      // temp.v  (recomputed, :543)
    go.d.a[i] += th * out.d.a[i];                 // backward_gate :524
    const F td = go.v.a[i] * out.d.a[i];          // temp.d  :525 (temp.d starts at 0)
    const F dv = nl == TANH ? (-th * th + F(1)) : (nl == RELU ? (th > F(0) ? F(1) : F(0)) : F(1));
    state.d.a[i] += dv * td;                      // backward_nonlin :172 / relu / identity (additive form)
  }
}
// clip_gradient + sgd_update(Params) (clstm_compute.cc:553-563)
template <class F>
void clip_and_update(Batch<F>& p, F lr, F mom, F gc) {
  if (gc < F(1e6))
    for (auto& d : p.d.a) { d = std::min(d, gc); d = std::max(d, -gc); }
  const size_t N = p.v.a.size();
  for (size_t i = 0; i < N; i++) p.v.a[i] += p.d.a[i] * lr;
  for (size_t i = 0; i < N; i++) p.d.a[i] = p.d.a[i] * mom;
}

// ------------------------------------------------------------------------------------------
// deterministic init  (batches.cc:11-52)
// ------------------------------------------------------------------------------------------
double g_state = 0.1;
inline double randu() {
  g_state = 189843.9384938 * g_state + 0.328340981343;
  g_state -= std::floor(g_state);
  return g_state;
}
inline double randn() {  // batches.cc:19-26 (sic: no sqrt)
  double u1 = randu(), u2 = randu();
  double r = -2 * std::log(u1);
  double theta = 2 * M_PI * u2;
  return r * std::cos(theta);
}
template <class F>
void rinit_mat(Mat<F>& a, F s, const std::string& mode, F offset) {
  const int n = a.r, m = a.c;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < m; j++) {
      if (mode == "unif") a(i, j) = 2 * s * randu() - s + offset;
      else if (mode == "negbiased") a(i, j) = 3 * s * randu() - 2 * s + offset;
      else if (mode == "pos") a(i, j) = s * randu() + offset;
      else if (mode == "neg") a(i, j) = -s * randu() + offset;
      else if (mode == "normal") a(i, j) = s * randn() + offset;
    }
}

// ------------------------------------------------------------------------------------------
// layers  (clstm.cc:354-668)
// ------------------------------------------------------------------------------------------
template <class F>
struct Layer {
  virtual ~Layer() {}
  Seq<F> inputs, outputs;
  std::map<std::string, Batch<F>*> parameters;  // std::map => alphabetical (clstm.h:109)
  std::vector<std::shared_ptr<Layer<F>>> sub;
  virtual void forward() = 0;
  virtual void backward() = 0;
  virtual int noutput() = 0;
};
template <class F, class Fn>
void walk_params(Layer<F>* net, Fn f) {  // clstm.cc:59-62
  for (auto& it : net->parameters) f(it.second);
  for (auto& s : net->sub) walk_params(s.get(), f);
}
template <class F>
void rinit_params(Batch<F>& p, int r, int c) {  // clstm.cc:30-36 defaults: negbiased, 0.01, 0
  p.resize(r, c);
  rinit_mat(p.v, F(0.01f), "negbiased", F(0));
}

template <class F>
struct NPLSTM : Layer<F> {  // GenericNPLSTM<F=SIG, G, H>  clstm.cc:546-668 (G = g_nl, H = h_nl; NPLSTM: TANH, TANH)
  using Layer<F>::inputs; using Layer<F>::outputs;
  Seq<F> source, gi, gf, go, ci, state;
  Batch<F> WGI, WGF, WGO, WCI;
  int ni, no, nf;
  int g_nl = TANH, h_nl = TANH;
  NPLSTM(int ni_, int no_, int g_nl_ = TANH, int h_nl_ = TANH) : ni(ni_), no(no_), nf(ni_ + no_), g_nl(g_nl_), h_nl(h_nl_) {
    this->parameters["WGI"] = &WGI; this->parameters["WGF"] = &WGF;
    this->parameters["WGO"] = &WGO; this->parameters["WCI"] = &WCI;
    rinit_params(WGI, no, nf + 1);  // draw order clstm.cc:588-591
    rinit_params(WGF, no, nf + 1);
    rinit_params(WGO, no, nf + 1);
    rinit_params(WCI, no, nf + 1);
  }
  int noutput() override { return no; }
  void forward() override {  // clstm.cc:600-621
    const int N = inputs.size(), bs = inputs.cols();
    source.resize(N, nf, bs); state.resize(N, no, bs);
    gi.resize(N, no, bs); go.resize(N, no, bs); gf.resize(N, no, bs); ci.resize(N, no, bs);
    outputs.resize(N, no, bs);
    for (int t = 0; t < N; t++) {
      forward_stack_delay(source[t], inputs[t], outputs, t - 1);
      forward_full1(gi[t], WGI, source[t], SIG);
      forward_full1(gf[t], WGF, source[t], SIG);
      forward_full1(go[t], WGO, source[t], SIG);
      forward_full1(ci[t], WCI, source[t], g_nl);
      forward_statemem(state[t], ci[t], gi[t], state, t - 1, gf[t]);
      forward_nonlingate(outputs[t], state[t], go[t], h_nl);
    }
  }
  void backward() override {  // clstm.cc:622-653 (the O(T^2) anynan asserts are omitted)
    inputs.zeroGrad();  // clearStateDerivs clstm.cc:188-195
    gi.zeroGrad(); gf.zeroGrad(); go.zeroGrad(); ci.zeroGrad(); state.zeroGrad(); source.zeroGrad();
    const int N = inputs.size();
    Seq<F> out;
    out.copy(outputs);
    for (int t = N - 1; t >= 0; t--) {
      backward_nonlingate(out[t], state[t], go[t], h_nl);
      backward_statemem(state[t], ci[t], gi[t], state, t - 1, gf[t]);
      backward_full1(ci[t], WCI, source[t], g_nl);
      backward_full1(go[t], WGO, source[t], SIG);
      backward_full1(gf[t], WGF, source[t], SIG);
      backward_full1(gi[t], WGI, source[t], SIG);
      backward_stack_delay(source[t], inputs[t], out, t - 1);
    }
  }
};

template <class F>
struct Reversed : Layer<F> {  // clstm.cc:458-479 ; forward_reverse/backward_reverse clstm_compute.cc:414-421
  using Layer<F>::inputs; using Layer<F>::outputs; using Layer<F>::sub;
  int noutput() override { return sub[0]->noutput(); }
  void forward() override {
    auto& net = *sub[0];
    const int N = inputs.size();
    net.inputs.like(inputs);
    for (int i = 0; i < N; i++) net.inputs[N - i - 1] = inputs[i];  // copies v and d
    net.forward();
    outputs.like(net.outputs);
    for (int i = 0; i < N; i++) outputs[N - i - 1] = net.outputs[i];
  }
  void backward() override {
    auto& net = *sub[0];
    const int N = outputs.size();
    net.outputs.zeroGrad();
    for (int i = 0; i < N; i++) {  // backward_reverse(outputs, net->outputs): x=net.outputs, y=outputs
      auto& xd = net.outputs[N - i - 1].d.a; auto& yd = outputs[i].d.a;
      for (size_t k = 0; k < xd.size(); k++) xd[k] += yd[k];
    }
    net.backward();
    outputs.zeroGrad();
    for (int i = 0; i < N; i++) {  // backward_reverse(net->inputs, inputs): x=inputs, y=net.inputs
      auto& xd = inputs[N - i - 1].d.a; auto& yd = net.inputs[i].d.a;
      for (size_t k = 0; k < xd.size(); k++) xd[k] += yd[k];
    }
  }
};

template <class F>
struct Parallel : Layer<F> {  // clstm.cc:506-544
  using Layer<F>::inputs; using Layer<F>::outputs; using Layer<F>::sub;
  int noutput() override { return sub[0]->noutput() + sub[1]->noutput(); }
  void forward() override {
    const int N = inputs.size();
    sub[0]->inputs.copy(inputs); sub[0]->forward();
    sub[1]->inputs.copy(inputs); sub[1]->forward();
    outputs.resize(N, noutput(), inputs.cols());
    for (int t = 0; t < N; t++) forward_stack(outputs[t], sub[0]->outputs[t], sub[1]->outputs[t]);
  }
  void backward() override {
    const int N = outputs.size();
    sub[0]->outputs.zeroGrad(); sub[1]->outputs.zeroGrad();
    for (int t = N - 1; t >= 0; t--) backward_stack(outputs[t], sub[0]->outputs[t], sub[1]->outputs[t]);
    sub[0]->backward(); sub[1]->backward();
    for (int t = 0; t < N; t++) {
      inputs[t].d = sub[0]->inputs[t].d;
      auto& a = inputs[t].d.a; auto& b = sub[1]->inputs[t].d.a;
      for (size_t k = 0; k < a.size(); k++) a[k] += b[k];
    }
  }
};

template <class F>
struct Softmax : Layer<F> {  // SoftmaxLayer clstm.cc:391-419
  using Layer<F>::inputs; using Layer<F>::outputs;
  Batch<F> W1;
  Softmax(int ni, int no) {
    this->parameters["W1"] = &W1;
    rinit_params(W1, no, ni + 1);
  }
  int noutput() override { return W1.rows(); }
  void forward() override {
    outputs.resize(inputs.size(), W1.rows(), inputs.cols());
    for (int t = 0; t < inputs.size(); t++) forward_softmax(outputs[t], W1, inputs[t]);
  }
  void backward() override {
    for (int t = outputs.size() - 1; t >= 0; t--) backward_softmax(outputs[t], W1, inputs[t]);
  }
};

template <class F>
struct Full : Layer<F> {  // Full<NONLIN> clstm.cc:354-389: LinearLayer, SigmoidLayer, TanhLayer, ReluLayer
  using Layer<F>::inputs; using Layer<F>::outputs;
  Batch<F> W1;
  int nl;
  Full(int ni, int no, int nl_) : nl(nl_) {
    this->parameters["W1"] = &W1;
    rinit_params(W1, no, ni + 1);
  }
  int noutput() override { return W1.rows(); }
  void forward() override {
    outputs.resize(inputs.size(), W1.rows(), inputs.cols());
    for (int t = 0; t < inputs.size(); t++) forward_full1(outputs[t], W1, inputs[t], nl);
  }
  void backward() override {
    for (int t = outputs.size() - 1; t >= 0; t--) backward_full1(outputs[t], W1, inputs[t], nl);
  }
};

template <class F>
struct Stacked : Layer<F> {  // clstm.cc:421-456
  using Layer<F>::inputs; using Layer<F>::outputs; using Layer<F>::sub;
  int noutput() override { return sub.back()->noutput(); }
  void forward() override {
    for (size_t n = 0; n < sub.size(); n++) {
      if (n == 0) sub[n]->inputs.copy(inputs);
      else sub[n]->inputs.copy(sub[n - 1]->outputs);
      sub[n]->forward();
    }
    outputs.copy(sub.back()->outputs);
  }
  void backward() override {
    for (int n = (int)sub.size() - 1; n >= 0; n--) {
      if (n + 1 == (int)sub.size())
        for (int t = 0; t < outputs.size(); t++) sub[n]->outputs[t].d = outputs[t].d;
      else
        for (int t = 0; t < sub[n + 1]->inputs.size(); t++) sub[n]->outputs[t].d = sub[n + 1]->inputs[t].d;
      sub[n]->backward();
    }
    for (int t = 0; t < sub[0]->inputs.size(); t++) inputs[t].d = sub[0]->inputs[t].d;
  }
};

template <class F>
std::shared_ptr<Layer<F>> make_bidi(int ni, int nh, int no) {  // clstm_prefab.cc:52-68
  auto fwd = std::make_shared<NPLSTM<F>>(ni, nh);   // RNG draw order: fwd LSTM, rev LSTM, softmax
  auto revl = std::make_shared<NPLSTM<F>>(ni, nh);
  auto rev = std::make_shared<Reversed<F>>();
  rev->sub.push_back(revl);
  auto par = std::make_shared<Parallel<F>>();
  par->sub.push_back(fwd); par->sub.push_back(rev);
  auto sm = std::make_shared<Softmax<F>>(2 * nh, no);
  auto st = std::make_shared<Stacked<F>>();
  st->sub.push_back(par); st->sub.push_back(sm);
  return st;
}

// the 1-D prefabs of clstm_prefab.cc:22-129.  cell: 0 NPLSTM, 1 LINNPLSTM, 2 RELUTANHNPLSTM, 3 RELUNPLSTM, 4 RELU2NPLSTM
// (clstm.cc:655-668); output: 0 SoftmaxLayer, 1 SigmoidLayer, 2 LinearLayer, 3 TanhLayer, 4 ReluLayer, -1 none (bidi0)
template <class F>
std::shared_ptr<Layer<F>> make_prefab(const std::string& kind, int ni, int nh, int nh2, int no, int cell, int output) {
  const int g_nl = cell >= 2 ? RELU : TANH;
  const int h_nl = (cell == 0 || cell == 2) ? TANH : (cell == 4 ? RELU : LIN);
  auto lstm = [&](int i, int o) { return std::make_shared<NPLSTM<F>>(i, o, g_nl, h_nl); };
  auto bidi_block = [&](int i, int o) {
    auto fwd = lstm(i, o);
    auto rev = std::make_shared<Reversed<F>>();
    rev->sub.push_back(lstm(i, o));
    auto par = std::make_shared<Parallel<F>>();
    par->sub.push_back(fwd); par->sub.push_back(rev);
    return par;
  };
  auto out_layer = [&](int i) -> std::shared_ptr<Layer<F>> {
    if (output == 0) return std::make_shared<Softmax<F>>(i, no);
    static const int nls[5] = {0, SIG, LIN, TANH, RELU};
    return std::make_shared<Full<F>>(i, no, nls[output]);
  };
  if (kind == "bidi0") return bidi_block(ni, nh);
  auto st = std::make_shared<Stacked<F>>();
  if (kind == "lstm1") { st->sub.push_back(lstm(ni, nh)); st->sub.push_back(out_layer(nh)); }
  else if (kind == "revlstm1") {
    auto rev = std::make_shared<Reversed<F>>();
    rev->sub.push_back(lstm(ni, nh));
    st->sub.push_back(rev); st->sub.push_back(out_layer(nh));
  } else if (kind == "bidi") { st->sub.push_back(bidi_block(ni, nh)); st->sub.push_back(out_layer(2 * nh)); }
  else if (kind == "bidi2") {
    st->sub.push_back(bidi_block(ni, nh)); st->sub.push_back(bidi_block(2 * nh, nh2)); st->sub.push_back(out_layer(2 * nh2));
  } else return nullptr;
  return st;
}

template <class F>
void sgd_update_net(Layer<F>* net, F lr, F mom, F gc) {  // clstm.cc:201-217 (state-clip of dead d's skipped: no effect on results)
  for (auto& it : net->parameters) clip_and_update(*it.second, lr, mom, gc);
  for (auto& s : net->sub) sgd_update_net(s.get(), lr, mom, gc);
}

// ------------------------------------------------------------------------------------------
// CTC  (ctc.cc)
// ------------------------------------------------------------------------------------------
template <class F> struct RM {  // row-major helper for the CTC lattices (EigenTensor2 used only as a container)
  int r, c; std::vector<F> a;
  RM(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, F(0)) {}
  F& operator()(int i, int j) { return a[(size_t)i * c + j]; }
  F operator()(int i, int j) const { return a[(size_t)i * c + j]; }
};
// forward_algorithm ctc.cc:24-40
template <class F>
void forward_algorithm(RM<F>& lr, const RM<F>& lmatch, double skip = -5) {
  const int n = lmatch.r, m = lmatch.c;
  std::vector<F> v(m), w(m);
  for (int j = 0; j < m; j++) v[j] = skip * j;
  for (int i = 0; i < n; i++) {
    w[0] = skip * i;
    for (int j = 1; j < m; j++) w[j] = v[j - 1];
    for (int j = 0; j < m; j++) {
      F same = v[j] + lmatch(i, j);
      F next = w[j] + lmatch(i, j);
      v[j] = log_add(same, next);
    }
    for (int j = 0; j < m; j++) lr(i, j) = v[j];
  }
}
// forwardbackward ctc.cc:42-55
template <class F>
void forwardbackward(RM<F>& both, const RM<F>& lmatch) {
  const int n = lmatch.r, m = lmatch.c;
  RM<F> lr(n, m);
  forward_algorithm(lr, lmatch);
  RM<F> rlmatch(n, m);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < m; j++) rlmatch(i, j) = lmatch(n - i - 1, m - j - 1);
  RM<F> rrl(n, m);
  forward_algorithm(rrl, rlmatch);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < m; j++) both(i, j) = lr(i, j) + rrl(n - i - 1, m - j - 1);
}
// ctc_align_targets ctc.cc:57-112 ; all row-major here
template <class F>
void ctc_align_targets(F* posteriors, const F* outputs, const F* targets, int n1, int n2, int nc) {
  const double lo = 1e-5;
  RM<F> lmatch(n1, n2);
  std::vector<F> out(nc);
  for (int t1 = 0; t1 < n1; t1++) {
    for (int i = 0; i < nc; i++) out[i] = std::fmax(lo, (double)outputs[(size_t)t1 * nc + i]);
    F s = F(0);  // asum1 tensor.h:337-342 (Float accumulator)
    for (int i = 0; i < nc; i++) s += out[i];
    for (int i = 0; i < nc; i++) out[i] = out[i] / s;
    for (int t2 = 0; t2 < n2; t2++) {
      double total = 0.0;
      for (int k = 0; k < nc; k++) total += out[k] * targets[(size_t)t2 * nc + k];
      lmatch(t1, t2) = std::log(total);
    }
  }
  RM<F> both(n1, n2);
  forwardbackward(both, lmatch);
  F mx = both(0, 0);  // amax2 tensor.h:367-374
  for (auto v : both.a) mx = std::fmax(mx, v);
  RM<F> epath(n1, n2);
  for (size_t i = 0; i < both.a.size(); i++) epath.a[i] = limexp(F(both.a[i] - mx));
  for (int j = 0; j < n2; j++) {
    double total = 0.0;
    for (int i = 0; i < n1; i++) total += epath(i, j);
    total = std::fmax(1e-9, total);
    for (int i = 0; i < n1; i++) epath(i, j) /= total;
  }
  for (int i = 0; i < n1; i++) {
    for (int j = 0; j < nc; j++) {
      double total = 0.0;
      for (int k = 0; k < n2; k++) {
        double value = epath(i, k) * targets[(size_t)k * nc + j];
        total += value;
      }
      posteriors[(size_t)i * nc + j] = total;
    }
  }
  for (int i = 0; i < n1; i++) {
    double total = 0.0;
    for (int j = 0; j < nc; j++) total += posteriors[(size_t)i * nc + j];
    total = std::fmax(total, 1e-9);
    for (int j = 0; j < nc; j++) posteriors[(size_t)i * nc + j] /= total;
  }
}
// mktargets ctc.cc:148-157 (row-major S x nc)
template <class F>
std::vector<F> mktargets(const int* labels, int L, int nc) {
  const int S = 2 * L + 1;
  std::vector<F> tg((size_t)S * nc, F(0));
  for (int t = 0; t < S; t++) {
    if (t % 2 == 1) tg[(size_t)t * nc + labels[(t - 1) / 2]] = 1;
    else tg[(size_t)t * nc + 0] = 1;
  }
  return tg;
}
// trivial_decode ctc.cc:159-194
int trivial_decode(const float* outputs, int N, int nc, int* cs, int* locs) {
  int count = 0, t = 0;
  float mv = 0; int mc = -1, mt = -1;
  while (t < N) {
    const float* row = outputs + (size_t)t * nc;
    int index = argmax1(row, nc);
    float v = row[index];
    if (index == 0) {
      if (mc != -1 && mc != 0) { cs[count] = mc; if (locs) locs[count] = mt; count++; }
      mv = 0; mc = -1; mt = -1; t++;
      continue;
    }
    if (v > mv) { mv = v; mc = index; mt = t; }
    t++;
  }
  return count;
}

// ------------------------------------------------------------------------------------------
// flat parameter (un)packing  (clstm.cc:838-918)
// ------------------------------------------------------------------------------------------
template <class F> size_t nparams(Layer<F>* net) {
  size_t n = 0; walk_params(net, [&](Batch<F>* p) { n += p->v.a.size(); }); return n;
}
template <class F, class G> void get_flat(Layer<F>* net, G* flat, bool deriv) {
  size_t k = 0;
  walk_params(net, [&](Batch<F>* p) { auto& a = deriv ? p->d.a : p->v.a; for (auto x : a) flat[k++] = (G)x; });
}
template <class F, class G> void set_flat(Layer<F>* net, const G* flat, bool deriv) {
  size_t k = 0;
  walk_params(net, [&](Batch<F>* p) { auto& a = deriv ? p->d.a : p->v.a; for (auto& x : a) x = (F)flat[k++]; });
}
template <class F> void set_inputs_image(Layer<F>* net, const F* image, int T, int ni) {  // clstm.cc:684-690
  net->inputs.resize(T, ni, 1);
  for (int t = 0; t < T; t++)
    for (int i = 0; i < ni; i++) net->inputs[t].v(i, 0) = image[(size_t)t * ni + i];
}

}  // namespace

// ==========================================================================================
// C API
// ==========================================================================================
struct oracle_net {
  std::shared_ptr<Layer<float>> net;
  int ni, nh, nc;
};
struct oracle_net64 {
  std::shared_ptr<Layer<double>> net;
  int ni, no;
};

namespace {
// CLSTMOCR::fwdbwd minus normalizer/codec (clstmhl.h:201-217)
void fwdbwd_line(Layer<float>* net, int ni, int nc, const float* image, int T, const int* labels, int L,
                 float* out, float* aligned_out) {
  set_inputs_image(net, image, T, ni);
  net->forward();
  std::vector<float> outs((size_t)T * nc), aligned((size_t)T * nc);
  for (int t = 0; t < T; t++)
    for (int k = 0; k < nc; k++) outs[(size_t)t * nc + k] = net->outputs[t].v(k, 0);
  std::vector<float> tg = mktargets<float>(labels, L, nc);
  ctc_align_targets<float>(aligned.data(), outs.data(), tg.data(), T, 2 * L + 1, nc);
  for (int t = 0; t < T; t++)
    for (int k = 0; k < nc; k++)
      net->outputs[t].d(k, 0) = aligned[(size_t)t * nc + k] - net->outputs[t].v(k, 0);
  net->backward();
  if (out) std::memcpy(out, outs.data(), outs.size() * sizeof(float));
  if (aligned_out) std::memcpy(aligned_out, aligned.data(), aligned.size() * sizeof(float));
}
}  // namespace

extern "C" {

void oracle_seed(double s) { g_state = s; }
double oracle_randu(void) { return randu(); }
void oracle_rinit(float* a, int rows, int cols, float s, const char* mode, float offset) {
  Mat<float> m;
  m.resize(rows, cols);
  rinit_mat(m, s, std::string(mode), offset);
  std::memcpy(a, m.a.data(), m.a.size() * sizeof(float));
}

oracle_net* oracle_bidi_create(int ninput, int nhidden, int noutput) {
  auto* o = new oracle_net;
  o->ni = ninput; o->nh = nhidden; o->nc = noutput;
  o->net = make_bidi<float>(ninput, nhidden, noutput);
  return o;
}
oracle_net* oracle_prefab_create(const char* prefab, int ninput, int nhidden, int nhidden2, int noutput, int cell, int output) {
  auto net = make_prefab<float>(prefab, ninput, nhidden, nhidden2, noutput, cell, output);
  if (!net) return nullptr;
  auto* o = new oracle_net;
  o->ni = ninput; o->nh = nhidden; o->nc = net->noutput();
  o->net = net;
  return o;
}
void oracle_destroy(oracle_net* o) { delete o; }
size_t oracle_nparams(oracle_net* o) { return nparams(o->net.get()); }
void oracle_get_params(oracle_net* o, float* flat) { get_flat(o->net.get(), flat, false); }
void oracle_set_params(oracle_net* o, const float* flat) { set_flat(o->net.get(), flat, false); }
void oracle_get_derivs(oracle_net* o, float* flat) { get_flat(o->net.get(), flat, true); }
void oracle_set_derivs(oracle_net* o, const float* flat) { set_flat(o->net.get(), flat, true); }
void oracle_clear_derivs(oracle_net* o) {
  walk_params(o->net.get(), [](Batch<float>* p) { p->d.zero(); });
}

void oracle_forward(oracle_net* o, const float* image, int T, float* out) {
  set_inputs_image(o->net.get(), image, T, o->ni);
  o->net->forward();
  for (int t = 0; t < T; t++)
    for (int k = 0; k < o->nc; k++) out[(size_t)t * o->nc + k] = o->net->outputs[t].v(k, 0);
}
void oracle_backward(oracle_net* o, const float* deltas, float* din) {
  const int T = o->net->outputs.size();
  for (int t = 0; t < T; t++)
    for (int k = 0; k < o->nc; k++) o->net->outputs[t].d(k, 0) = deltas[(size_t)t * o->nc + k];
  o->net->backward();
  if (din)
    for (int t = 0; t < T; t++)
      for (int i = 0; i < o->ni; i++) din[(size_t)t * o->ni + i] = o->net->inputs[t].d(i, 0);
}
void oracle_sgd_update(oracle_net* o, float lr, float momentum, float gc) {
  sgd_update_net<float>(o->net.get(), lr, momentum, gc);
}
void oracle_fwdbwd(oracle_net* o, const float* image, int T, const int* labels, int L, float* out,
                   float* aligned) {
  fwdbwd_line(o->net.get(), o->ni, o->nc, image, T, labels, L, out, aligned);
}

void oracle_forward_batched(oracle_net* o, const float* x, int T, int bs, float* out) {
  auto* net = o->net.get();
  net->inputs.resize(T, o->ni, bs);
  for (int t = 0; t < T; t++)
    for (int b = 0; b < bs; b++)
      for (int i = 0; i < o->ni; i++) net->inputs[t].v(i, b) = x[((size_t)t * bs + b) * o->ni + i];
  net->forward();
  for (int t = 0; t < T; t++)
    for (int b = 0; b < bs; b++)
      for (int k = 0; k < o->nc; k++) out[((size_t)t * bs + b) * o->nc + k] = net->outputs[t].v(k, b);
}
void oracle_backward_batched(oracle_net* o, const float* dout, float* din) {
  auto* net = o->net.get();
  const int T = net->outputs.size(), bs = net->outputs.cols();
  for (int t = 0; t < T; t++)
    for (int b = 0; b < bs; b++)
      for (int k = 0; k < o->nc; k++) net->outputs[t].d(k, b) = dout[((size_t)t * bs + b) * o->nc + k];
  net->backward();
  if (din)
    for (int t = 0; t < T; t++)
      for (int b = 0; b < bs; b++)
        for (int i = 0; i < o->ni; i++) din[((size_t)t * bs + b) * o->ni + i] = net->inputs[t].d(i, b);
}

void oracle_ctc_align_dense(float* posteriors, const float* outputs, const float* targets, int n1, int n2,
                            int nc) {
  ctc_align_targets<float>(posteriors, outputs, targets, n1, n2, nc);
}
void oracle_ctc_align_dense_f64(double* posteriors, const double* outputs, const double* targets, int n1,
                                int n2, int nc) {
  ctc_align_targets<double>(posteriors, outputs, targets, n1, n2, nc);
}
void oracle_ctc_align_labels(float* posteriors, const float* outputs, int T, int nc, const int* labels,
                             int L) {
  std::vector<float> tg = mktargets<float>(labels, L, nc);
  ctc_align_targets<float>(posteriors, outputs, tg.data(), T, 2 * L + 1, nc);
}
int oracle_trivial_decode(const float* outputs, int T, int nc, int* classes, int* locs) {
  return trivial_decode(outputs, T, nc, classes, locs);
}
void oracle_argmax_rows(const float* m, int T, int nc, int* idx) {
  for (int t = 0; t < T; t++) idx[t] = argmax1(m + (size_t)t * nc, nc);
}

// ---- double nets for gradient checks ------------------------------------------------------
oracle_net64* oracle64_create(int kind, int ni, int nh, int no) {
  auto* o = new oracle_net64;
  o->ni = ni;
  if (kind == 0) {
    o->net = std::make_shared<NPLSTM<double>>(ni, nh);
    o->no = nh;
  } else if (kind == 1) {
    auto r = std::make_shared<Reversed<double>>();
    r->sub.push_back(std::make_shared<NPLSTM<double>>(ni, nh));
    o->net = r;
    o->no = nh;
  } else {
    o->net = make_bidi<double>(ni, nh, no);
    o->no = no;
  }
  return o;
}
void oracle64_destroy(oracle_net64* o) { delete o; }
size_t oracle64_nparams(oracle_net64* o) { return nparams(o->net.get()); }
int oracle64_noutput(oracle_net64* o) { return o->no; }
void oracle64_get_params(oracle_net64* o, double* flat) { get_flat(o->net.get(), flat, false); }
void oracle64_set_params(oracle_net64* o, const double* flat) { set_flat(o->net.get(), flat, false); }
void oracle64_forward(oracle_net64* o, const double* x, int T, int bs, double* out) {
  auto* net = o->net.get();
  net->inputs.resize(T, o->ni, bs);
  for (int t = 0; t < T; t++)
    for (int b = 0; b < bs; b++)
      for (int i = 0; i < o->ni; i++) net->inputs[t].v(i, b) = x[((size_t)t * bs + b) * o->ni + i];
  net->forward();
  for (int t = 0; t < T; t++)
    for (int b = 0; b < bs; b++)
      for (int k = 0; k < o->no; k++) out[((size_t)t * bs + b) * o->no + k] = net->outputs[t].v(k, b);
}
void oracle64_backward(oracle_net64* o, const double* dout, double* din, double* dparams) {
  auto* net = o->net.get();
  const int T = net->outputs.size(), bs = net->outputs.cols();
  walk_params(net, [](Batch<double>* p) { p->d.zero(); });
  for (int t = 0; t < T; t++)
    for (int b = 0; b < bs; b++)
      for (int k = 0; k < o->no; k++) net->outputs[t].d(k, b) = dout[((size_t)t * bs + b) * o->no + k];
  net->backward();
  if (din)
    for (int t = 0; t < T; t++)
      for (int b = 0; b < bs; b++)
        for (int i = 0; i < o->ni; i++) din[((size_t)t * bs + b) * o->ni + i] = net->inputs[t].d(i, b);
  if (dparams) get_flat(net, dparams, true);
}

// ---- CPU baseline driver ------------------------------------------------------------------
double oracle_train_lines(oracle_net* o, const float* x, const int* T, int B, const int* labels,
                          const int* L, float lr, float momentum, int threads, int reps) {
  std::vector<size_t> xoff(B + 1, 0), loff(B + 1, 0);
  for (int b = 0; b < B; b++) { xoff[b + 1] = xoff[b] + (size_t)T[b] * o->ni; loff[b + 1] = loff[b] + L[b]; }
  if (threads < 1) threads = 1;
  std::vector<std::shared_ptr<Layer<float>>> reps_nets;  // replicas 1..threads-1 (replica 0 = o->net)
  const size_t P = nparams(o->net.get());
  std::vector<float> flat(P);
  for (int i = 1; i < threads; i++) reps_nets.push_back(make_bidi<float>(o->ni, o->nh, o->nc));
  auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < reps; r++) {
    if (threads == 1) {
      for (int b = 0; b < B; b++)
        fwdbwd_line(o->net.get(), o->ni, o->nc, x + xoff[b], T[b], labels + loff[b], L[b], nullptr, nullptr);
    } else {
      get_flat(o->net.get(), flat.data(), false);  // distribute_weights clstm.cc:718-729
      for (auto& n : reps_nets) {
        set_flat(n.get(), flat.data(), false);
        walk_params(n.get(), [](Batch<float>* p) { p->d.zero(); });
      }
      std::atomic<int> next(0);  // lines handed out dynamically, one replica net per host thread
      auto work = [&](int tid) {
        Layer<float>* net = tid == 0 ? o->net.get() : reps_nets[tid - 1].get();
        for (;;) {
          int b = next.fetch_add(1);
          if (b >= B) break;
          fwdbwd_line(net, o->ni, o->nc, x + xoff[b], T[b], labels + loff[b], L[b], nullptr, nullptr);
        }
      };
      std::vector<std::thread> pool;
      for (int tid = 1; tid < threads; tid++) pool.emplace_back(work, tid);
      work(0);
      for (auto& th : pool) th.join();
      // share_deltas clstm.cc:731-744: sum every replica's Params.d into replica 0
      std::vector<float> acc(P), tmp(P);
      get_flat(o->net.get(), acc.data(), true);
      for (auto& n : reps_nets) {
        get_flat(n.get(), tmp.data(), true);
        for (size_t k = 0; k < P; k++) acc[k] += tmp[k];
      }
      set_flat(o->net.get(), acc.data(), true);
    }
    sgd_update_net<float>(o->net.get(), lr, momentum, 100.0f);
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
