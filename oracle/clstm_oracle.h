// clstm_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the one hot path of tmbdev/clstm (bidirectional NPLSTM over image
// columns + SoftmaxLayer + OCRopus-style CTC alignment + clip/SGD-momentum update).
// It exists only to check the CUDA path in clstm_b200/ and to serve as the timed CPU
// baseline ("cpu_baseline.kind" = "port") in bench.py.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may call into it.
//
// Parity pin status:
//   * CTC (ctc_align_targets)   : PINNED by the reference's two known-answer tests
//                                 (/root/reference/test-ctc.cc:47-74, 76-109), see
//                                 tests/golden/ctc_kat.json + tests/test_oracle_ctc.py.
//   * LSTM / Softmax arithmetic : "parity unpinned" -- the reference ships no golden vectors
//                                 for these (only numeric gradient checks < 0.1 relative,
//                                 test-deriv.cc:134-171) and the reference itself cannot be
//                                 compiled here (Eigen is absent).  The restatement is checked
//                                 with the reference's own gradient-check recipe instead.
//
// Every function cites the reference file:line it restates (paths relative to /root/reference).
#pragma once
#include <cstddef>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_net oracle_net;

// ---- deterministic "random" init (batches.cc:11-17, 31-52) -----------------------------
void   oracle_seed(double s);                 // batches.cc:11  (env "seed", default 0.1)
double oracle_randu(void);                    // batches.cc:13-17
// rinit(TensorMap2, s, mode, offset) batches.cc:31-52; a is col-major rows x cols,
// draw order i outer / j inner.
void   oracle_rinit(float* a, int rows, int cols, float s, const char* mode, float offset);

// ---- bidi network (clstm_prefab.cc:52-68) in float --------------------------------------
// Stacked{ Parallel{ NPLSTM, Reversed{NPLSTM} }, SoftmaxLayer }
oracle_net* oracle_bidi_create(int ninput, int nhidden, int noutput);   // weights via rinit, reference draw order
void   oracle_destroy(oracle_net*);
size_t oracle_nparams(oracle_net*);
// flat order = walk_params order (clstm.cc:59-62): fwd.WCI,WGF,WGI,WGO, rev.WCI,WGF,WGI,WGO, softmax.W1,
// each matrix col-major rows x cols with column 0 = bias.
void   oracle_get_params(oracle_net*, float* flat);
void   oracle_set_params(oracle_net*, const float* flat);
void   oracle_get_derivs(oracle_net*, float* flat);
void   oracle_set_derivs(oracle_net*, const float* flat);
void   oracle_clear_derivs(oracle_net*);

// set_inputs(net, image) clstm.cc:684-690 ; image is row-major T x ninput (image(t,i)); batch 1
// followed by net->forward() (Stacked::forward clstm.cc:424-439).  out: row-major T x noutput.
void   oracle_forward(oracle_net*, const float* image, int T, float* out);
// outputs[t].d = deltas (row-major T x noutput), then net->backward() (clstm.cc:440-454).
// din (nullable): inputs[t].d, row-major T x ninput.
void   oracle_backward(oracle_net*, const float* deltas, float* din);
// sgd_update(Network) clstm.cc:201-217 with gradient_clip gc (100 default)
void   oracle_sgd_update(oracle_net*, float lr, float momentum, float gc);

// CLSTMOCR::fwdbwd without the normalizer (clstmhl.h:201-217): forward, mktargets, ctc_align,
// delta = aligned - outputs, backward.  labels: L class indices.  out/aligned nullable, T x nc.
void   oracle_fwdbwd(oracle_net*, const float* image, int T, const int* labels, int L,
                     float* out, float* aligned);

// ---- CTC (ctc.cc) -----------------------------------------------------------------------
// ctc_align_targets(EigenTensor2&, EigenTensor2&, EigenTensor2&) ctc.cc:57-112
// outputs: row-major n1 x nc, targets row-major n2 x nc, posteriors row-major n1 x nc
void   oracle_ctc_align_dense(float* posteriors, const float* outputs, const float* targets,
                              int n1, int n2, int nc);
void   oracle_ctc_align_dense_f64(double* posteriors, const double* outputs, const double* targets,
                              int n1, int n2, int nc);
// mktargets (ctc.cc:148-157) + ctc_align_targets on a transcript
void   oracle_ctc_align_labels(float* posteriors, const float* outputs, int T, int nc,
                               const int* labels, int L);
// trivial_decode (ctc.cc:159-194): returns count; classes/locs must hold >= T entries
int    oracle_trivial_decode(const float* outputs, int T, int nc, int* classes, int* locs);
// argmax with the reference's tie rule (tensor.h:357-366), per row
void   oracle_argmax_rows(const float* m, int T, int nc, int* idx);

// ---- generic small nets for the reference's gradient-check recipe (test-deriv.cc) in double
// kind: 0 = NPLSTM(ni->no), 1 = Reversed{NPLSTM}, 2 = bidi with softmax output (ni,nh,no)
// Computes forward on inputs (T x ni x bs, layout [t][b][i]), sets outputs.d = dout ([t][b][o]),
// runs backward; returns outputs, input deltas and parameter derivs (flat walk_params order).
typedef struct oracle_net64 oracle_net64;
oracle_net64* oracle64_create(int kind, int ni, int nh, int no);
void   oracle64_destroy(oracle_net64*);
size_t oracle64_nparams(oracle_net64*);
int    oracle64_noutput(oracle_net64*);
void   oracle64_get_params(oracle_net64*, double* flat);
void   oracle64_set_params(oracle_net64*, const double* flat);
void   oracle64_forward(oracle_net64*, const double* x, int T, int bs, double* out);
void   oracle64_backward(oracle_net64*, const double* dout, double* din, double* dparams);

// float batched variant of the same (bs > 1 columns per Batch; test-batchlstm.cc semantics)
void   oracle_forward_batched(oracle_net*, const float* x, int T, int bs, float* out);
void   oracle_backward_batched(oracle_net*, const float* dout, float* din);

// ---- CPU baseline driver -----------------------------------------------------------------
// Runs the reference training step (fwdbwd per line, batch 1, then one sgd_update) over B lines.
// x packed [sum T][ninput]; labels packed [sum L].  threads<=1: faithful single thread;
// threads>1: std::thread pool over lines on `threads` replicas with share_deltas semantics
// (clstm.cc:731-744: Params.d summed over replicas) before the update.
// Returns seconds of wall time spent (steady_clock) for `reps` repetitions.
double oracle_train_lines(oracle_net*, const float* x, const int* T, int B,
                          const int* labels, const int* L, float lr, float momentum,
                          int threads, int reps);

#ifdef __cplusplus
}
#endif
