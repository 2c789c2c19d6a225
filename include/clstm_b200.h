/* clstm_b200.h -- C ABI of the B200-native replacement for clstm's hot path.
 *
 * Scope: the `bidi` network of tmbdev/clstm
 *     Stacked{ Parallel{ NPLSTM, Reversed{NPLSTM} }, SoftmaxLayer }        (clstm_prefab.cc:52-68)
 * forward, OCRopus-style CTC alignment, backward, clip + SGD-momentum update, and the data-parallel
 * gradient sum -- executed as hand-written sm_100a CUDA kernels.  There is NO CPU fallback: every entry
 * point fails (non-zero status + clstm_b200_last_error()) when no CUDA device / kernel image is usable.
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success, non-zero on failure and never throws.
 *     The host-side INetwork mirror (clstm_b200/host/) turns non-zero into THROW(clstm_b200_last_error()),
 *     matching the reference's error style (clstm.h THROW / utils.h:260-267 throwf).
 *   - all host buffers are caller-owned and only read/written during the call.
 *   - a handle owns all of its device memory and one CUDA stream; a handle is single-caller, different
 *     handles may be driven from different host threads (no process-global mutable state besides the
 *     thread-local error string).
 *   - "column" = one timestep of one text line (a 1 x ninput pixel column).  A batch of B lines with lengths
 *     T[b] is passed PACKED: x[sum_b T[b]][ninput], line b occupying rows off_b .. off_b+T[b]-1, row-major,
 *     x[(off_b+t)*ninput + i] == inputs[t].v(i, 0) of line b in the reference (clstm.cc:684-690).
 *     Outputs / aligned / deltas are packed the same way with nclasses entries per column
 *     (== outputs[t].v(c,0), clstmhl.h:262-271).
 *   - flat parameter vectors use the reference's walk_params order (clstm.cc:59-62, std::map => alphabetical):
 *       fwd.WCI, fwd.WGF, fwd.WGI, fwd.WGO, rev.WCI, rev.WGF, rev.WGI, rev.WGO, softmax.W1
 *     each matrix column-major rows x cols exactly as Params.v.ptr (tensor.h:252), column 0 = bias,
 *     LSTM matrices nhidden x (1+ninput+nhidden), W1 nclasses x (1+2*nhidden).
 *     This is what get_params/set_params/get_derivs (clstm.cc:859-918) exchange.
 */
#ifndef CLSTM_B200_H_
#define CLSTM_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct clstm_b200_net clstm_b200_net;

typedef struct clstm_b200_cfg {
  int ninput;    /* 48 for OCR lines (CLSTMOCR::target_height, clstmhl.h:149)        */
  int nhidden;   /* per direction                                                     */
  int nclasses;  /* codec size incl. class 0 = blank (clstm.cc:247-267)              */
  int device;    /* CUDA ordinal; replaces the layer attribute "gpu" (clstm.cc:571)   */
} clstm_b200_cfg;

/* replaces make_net("bidi", ...) + initialize() for the device-resident net (clstmhl.h:191-200).
 * Weights are zero until clstm_b200_set_params(); the host mirror fills them with the reference LCG init. */
int clstm_b200_create(const clstm_b200_cfg* cfg, clstm_b200_net** out);

/* The other prefab topologies of clstm_prefab.cc:22-129 and the layer variants of clstm.cc:382-389, 655-668:
 * nblocks stacked recurrent blocks (each one LSTM, one Reversed{LSTM}, or Parallel{LSTM, Reversed{LSTM}}) followed by
 * an output layer.   lstm1: {1, direction 0}   revlstm1: {1, direction 1}   bidi: {1, direction 2}
 * bidi2: {2 blocks, direction 2, 2}   bidi0: bidi with output -1.   The flat parameter order is walk_params of the
 * corresponding reference tree (clstm.cc:59-62).  The CTC entry points need a Softmax output; the register / cluster
 * recurrent kernels serve cell 0 (NPLSTM), the other cells run on the generic kernels. */
typedef struct {
  int ninput;        /* features per column                                                              */
  int noutput;       /* outputs of the output layer (ignored for output -1)                              */
  int device;
  int nblocks;       /* 1 or 2                                                                           */
  int nhidden[2];    /* hidden units per LSTM of block k ("nhidden", "nhidden2")                         */
  int direction[2];  /* 0 forward only, 1 reversed only, 2 both (Parallel, outputs concatenated)         */
  int cell;          /* 0 NPLSTM, 1 LINNPLSTM, 2 RELUTANHNPLSTM, 3 RELUNPLSTM, 4 RELU2NPLSTM              */
  int output;        /* 0 SoftmaxLayer, 1 SigmoidLayer, 2 LinearLayer, 3 TanhLayer, 4 ReluLayer, -1 none */
} clstm_b200_cfg_ex;
int clstm_b200_create_ex(const clstm_b200_cfg_ex* cfg, clstm_b200_net** out);
void clstm_b200_destroy(clstm_b200_net* net);

/* n_params / set_params / get_params / get_derivs / set_derivs / clear_derivs (clstm.cc:838-918) */
size_t clstm_b200_nparams(const clstm_b200_net* net);
int clstm_b200_set_params(clstm_b200_net* net, const float* flat, size_t n);
int clstm_b200_get_params(clstm_b200_net* net, float* flat, size_t n);
int clstm_b200_get_derivs(clstm_b200_net* net, float* flat, size_t n);
int clstm_b200_set_derivs(clstm_b200_net* net, const float* flat, size_t n);
int clstm_b200_clear_derivs(clstm_b200_net* net);

/* set_inputs(...) + net->forward() (clstm.cc:684-690, 424-439) for B independent lines.
 * x: packed host columns, T[b]: line lengths, out (nullable): packed host outputs [sumT][nclasses]. */
int clstm_b200_forward(clstm_b200_net* net, const float* x, const int* T, int B, float* out);

/* mktargets + ctc_align_targets (ctc.cc:148-157, 57-134) on the outputs of the last forward, and the
 * delta injection outputs[t].d = aligned[t].v - outputs[t].v (clstmhl.h:211-212) kept on the device.
 * labels: packed transcripts (class indices, no blanks), L[b] entries for line b.
 * aligned (nullable): packed host [sumT][nclasses]. */
int clstm_b200_ctc_align(clstm_b200_net* net, const int* labels, const int* L, float* aligned);

/* ctc_align_targets(Sequence& posteriors, Sequence& outputs, Classes& targets) (ctc.cc:136-146): free-standing
 * alignment of caller-supplied outputs (packed host [sumT][nclasses]) against explicit target STATE classes
 * (states: packed, S[b] entries for line b, no blank interleaving is added).  This is the form the reference's
 * known-answer tests use (test-ctc.cc:47-109).  aligned: packed host [sumT][nclasses]. */
int clstm_b200_ctc_align_states(clstm_b200_net* net, const float* outputs, const int* T, int B, const int* states,
                                const int* S, float* aligned);

/* net->backward() (clstm.cc:440-454).  deltas (nullable): packed host outputs[t].d; NULL = use the deltas
 * left on the device by clstm_b200_ctc_align.  din (nullable): packed host inputs[t].d [sumT][ninput].
 * Parameter derivatives ACCUMULATE across calls like Params.d does (clstm_compute.cc:296-298). */
int clstm_b200_backward(clstm_b200_net* net, const float* deltas, float* din);

/* trivial_decode (ctc.cc:159-194) per line on the device.  which: 0 = outputs (prediction),
 * 1 = aligned (CLSTMOCR::aligned_utf8, clstmhl.h:224-229).  classes/locs: host [B][max_per_line],
 * counts: host [B].  A line that decodes to more than max_per_line symbols is an error. */
int clstm_b200_decode(clstm_b200_net* net, int which, int* classes, int* locs, int* counts, int max_per_line);

/* per-column argmax with the reference tie rule (tensor.h:357-366: last maximal index).
 * which as above; idx: host [sumT]. */
int clstm_b200_argmax(clstm_b200_net* net, int which, int* idx);

/* sgd_update(Network) (clstm.cc:201-217): d = clamp(d, +-clip) (skipped when clip >= 1e6);
 * v += lr * d; d *= momentum. */
int clstm_b200_sgd_update(clstm_b200_net* net, float lr, float momentum, float clip);

/* share_deltas (clstm.cc:731-744) across processes: sums THIS STEP's parameter derivatives over all ranks
 * (one NCCL all-reduce of one flat fp32 buffer over NVLink/NVSwitch) before they are folded into Params.d,
 * so momentum is not multiplied by the world size.  unique_id: the 128-byte ncclUniqueId created by
 * clstm_b200_comm_unique_id on rank 0 and distributed by the launcher (torch.distributed / MPI / files). */
int clstm_b200_comm_unique_id(void* id128);
int clstm_b200_comm_init(clstm_b200_net* net, const void* id128, int rank, int world);
int clstm_b200_allreduce_derivs(clstm_b200_net* net);

/* NVLink peer-memory variant of the same exchange, fused with the update: every rank exports its derivative buffer
 * (64-byte cudaIpcMemHandle_t), the launcher gathers the handles of all ranks, and after clstm_b200_p2p_connect the
 * fused training step (clstm_b200_train_step / clstm_b200_step_resident) replaces "NCCL all-reduce + sgd_update" by ONE
 * kernel that reads all ranks' derivatives directly over NVLink/NVSwitch, sums them in rank order and applies
 * clip + update (at most 8 ranks of one node). */
int clstm_b200_p2p_handle(clstm_b200_net* net, void* handle64);
int clstm_b200_p2p_connect(clstm_b200_net* net, const void* handles /* world x 64 bytes */, int rank, int world);

/* CLSTMOCR::train for a minibatch (clstmhl.h:201-223): forward, CTC align, backward, [all-reduce when a
 * communicator is attached], sgd_update, decode of the outputs.  Host buffers as above; out/aligned/classes/
 * locs/counts nullable.  One stream, no host synchronisation between the kernels. */
int clstm_b200_train_step(clstm_b200_net* net, const float* x, const int* T, int B, const int* labels,
                          const int* L, float lr, float momentum, float clip, float* out, float* aligned,
                          int* classes, int* locs, int* counts, int max_per_line);

/* The same step split for measurement with inputs already resident in HBM: upload once, then run the step
 * (no host<->device traffic, no host sync), then fetch results. */
int clstm_b200_upload_batch(clstm_b200_net* net, const float* x, const int* T, int B, const int* labels,
                            const int* L);
int clstm_b200_step_resident(clstm_b200_net* net, float lr, float momentum, float clip);
int clstm_b200_fetch_decoded(clstm_b200_net* net, int which, int* classes, int* locs, int* counts,
                             int max_per_line);
int clstm_b200_synchronize(clstm_b200_net* net);

/* Input pipeline for a training loop (the data-loader side of clstmocrtrain.cc:172-177): stage the NEXT batch into a
 * second input set on a copy stream while the step on the current batch is still running, then run the step on it.
 *   prefetch_batch(i+1) may be called as soon as step i has been launched (step_resident / step_prefetched return
 *   without waiting); the copy is asynchronous, so the host buffers must stay valid (and unmodified) until the
 *   first clstm_b200_fetch_decoded / clstm_b200_synchronize AFTER the step_prefetched that consumes the batch, and
 *   should be pinned (clstm_b200_alloc_pinned) for the copy to overlap.  step_prefetched makes the prefetched batch
 *   current and runs clstm_b200_step_resident on it.  fetch_decoded afterwards returns the results of that step.
 * If the prefetched batch needs larger device buffers than any batch before, the call waits for the running step and
 * re-allocates (device-resident activations of the current batch are lost; decoded results are kept). */
int clstm_b200_prefetch_batch(clstm_b200_net* net, const float* x, const int* T, int B, const int* labels, const int* L);
/* the same for RAW line images (the readSample -> normalizer->measure/normalize -> set_inputs prefix of
 * clstmocrtrain.cc:173-176 / clstmhl.h:202-205): clstm_b200_normalize_batch of batch i+1 on the copy stream into the spare input set
 * (its kernels run on the SMs the recurrent kernels of step i leave idle); blocks the calling thread only for the
 * normaliser's own round trip (the line widths), never for the running step.  Follow with step_prefetched. */
int clstm_b200_prefetch_raw_batch(clstm_b200_net* net, const float* raw, const int* W, const int* H, int B, int kind,
                                  const float* params, const int* labels, const int* L, int* T_out);
int clstm_b200_step_prefetched(clstm_b200_net* net, float lr, float momentum, float clip);

/* Text-line normalizers on the device (extras.h:31-47, extras.cc:146-301): measure() + normalize() of B raw line
 * images, written straight into the resident input batch (the place clstm_b200_upload_batch fills), so that
 * CLSTMOCR::fwdbwd / predict (clstmhl.h:201-205, 233-237) never materialise the normalised image on the host.
 *   raw     concatenated images, line b is W[b] x H[b] floats with pixel (i = column, j = row) at i + j*W[b]
 *           (the reference's Tensor2 image(i, j), already inverted: ink = 1, as clstmocrtrain.cc:74-75 does)
 *   kind    0 "none" (NoNormalizer; H[b] must equal ninput), 1 "mean" (MeanNormalizer), 2 "center" (CenterNormalizer)
 *   params  {range, smooth2d, smooth1d, vscale} or NULL for the reference defaults (extras.h:36-39; mean: range 1)
 *   labels/L nullable transcripts staged with the batch (as in upload_batch);  T_out[b] = normalised width
 * target_height is the net's ninput (clstmhl.h:164).  Results are bit-identical to the reference's float/double
 * arithmetic (same operations in the same order; Gaussian masks from the host libm). */
int clstm_b200_normalize_batch(clstm_b200_net* net, const float* raw, const int* W, const int* H, int B, int kind,
                               const float* params, const int* labels, const int* L, int* T_out);
/* CenterNormalizer state of the last normalize_batch: `center` (sum of W floats, extras.cc:229) and `r` (B floats, :230) */
int clstm_b200_normalizer_state(clstm_b200_net* net, float* center, float* r);
/* the resident input batch, [N][ninput] (what set_inputs would hold, clstm.cc:684-690) */
int clstm_b200_get_inputs(clstm_b200_net* net, float* x);
/* INetwork::forward on the resident batch (after upload_batch / normalize_batch); out nullable */
int clstm_b200_forward_resident(clstm_b200_net* net, float* out);

/* Measurement hooks (bench.py): per-phase CUDA-event timing on the handle's own stream.
 * phases: see clstm_b200_phase_name(i), i < clstm_b200_num_phases().  ms[i] = accumulated milliseconds and
 * launches[i] = kernel launches since the last reset. */
int clstm_b200_profile(clstm_b200_net* net, int enable);
int clstm_b200_num_phases(void);
const char* clstm_b200_phase_name(int i);
int clstm_b200_phase_stats(clstm_b200_net* net, float* ms, long long* launches, int n);
/* cudaStream_t of the handle (for external event timing) */
void* clstm_b200_stream(clstm_b200_net* net);
/* which recurrent kernel variant the last forward pass used: "regs" (weights register-resident, one CTA per line and
 * direction), "cluster" (thread-block cluster per line and direction), "tc" (batched tcgen05 recurrence: all lines of a
 * 128-slot tile in lock step, clstm.cc:612-620 / 629-650 as one GEMM per step) or "generic".
 * Environment: CLSTM_B200_LSTM=tc|simt forces / forbids the tensor-core recurrence (default: by size and batch). */
const char* clstm_b200_lstm_variant(const clstm_b200_net* net);

/* device self-test: tcgen05 (3xTF32) dense products vs the fp32 SIMT tiles on random data with the shapes of
 * this net; err[i] = max|difference| / max|reference| per case; returns the number of cases (<0 on error). */
int clstm_b200_selftest_gemm(clstm_b200_net* net, float* err, int max_cases);

/* Measurement hook of the fused NVLink all-reduce + update kernel (share_deltas + sgd_update, clstm.cc:731-744, 201-217):
 * the first call arms device counters, later calls return out4 = {launches, mean us a launch waited for the slowest rank,
 * mean us of the peer-read + update phase, bytes read over NVLink per launch = (world-1) * 4 * nparams}. */
int clstm_b200_peer_stats(clstm_b200_net* net, double* out4, int reset);

/* device self-test of the batched tensor-core recurrence (no net needed): random bidirectional problem with `nlines`
 * lines of tmin..tmax columns; out9[0..3] = max |difference| of gates, cell states, outputs and previous outputs against
 * the fp32 SIMT kernels, out9[4] = max |difference| of the backward deltas relative to their maximum,
 * out9[5..8] = milliseconds of tensor-core forward / backward and SIMT forward / backward.
 * Replaces nothing in the reference; it checks GenericNPLSTM::forward/backward (clstm.cc:600-653) computed two ways. */
int clstm_b200_selftest_lstm(int device, int nhidden, int nlines, int tmin, int tmax, unsigned seed, float wscale, float* out9);
/* the same A/B for the cluster-resident tensor-core recurrence (lstm_tcx.cu: 16 lines per thread-block cluster, h exchanged
 * through distributed shared memory); nhidden <= 256 */
int clstm_b200_selftest_lstm_x(int device, int nhidden, int nlines, int tmin, int tmax, unsigned seed, float wscale, float* out9);

/* page-locked host memory for batches that are copied every step (cudaHostAlloc) */
void* clstm_b200_alloc_pinned(size_t bytes);
void clstm_b200_free_pinned(void* p);

/* last error message of the calling thread ("" if none) */
const char* clstm_b200_last_error(void);
/* library version string */
const char* clstm_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CLSTM_B200_H_ */
