// kernels.h -- internal launcher declarations shared by the .cu files of libclstm_b200.so.
// Everything here is device-side plumbing for the one hot path; the public
// surface is include/clstm_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace cb200 {

// Per-batch line metadata, all device pointers (filled by capi.cu::upload_meta).
struct Lines {
  int B;                 // number of text lines in the batch
  int N;                 // total columns = sum_b T[b]
  int Tmax;              // longest line
  const int* T;          // [B] columns per line
  const int* off;        // [B] first column of line b in the packed arrays
  const int* L;          // [B] transcript length
  const int* lab_off;    // [B] first label of line b
  const int* labels;     // packed transcripts
  const long long* lat_off;  // [B] first element of line b's T x S lattice scratch
  const int* order;      // [B] line indices sorted by decreasing T (longest lines start first)
  const int* st_off;     // [B] first CTC state of line b in the packed per-state arrays
  int ntiles;            // 32-column tiles over all lines (a tile never crosses a line boundary)
  const int* tile_line;  // [ntiles] line of the tile
  const int* tile_t0;    // [ntiles] first column (within the line) of the tile
};

// ---------------------------------------------------------------- gemm.cu
// C[M x N] (row-major, ldc) = beta*C + sum_k A(m,k)*B(k,n) (+ bias[n]); A(m,k)=A[m*sam+k*sak], B(k,n)=B[k*sbk+n*sbn].
// fp32 SIMT tiles; `ws` is split-K workspace of at least splits*M*N floats (may be null => no split).
// Both return the number of kernels launched.
int gemm_f32(cudaStream_t st, int M, int N, int K, const float* A, long long sam, long long sak, const float* B,
              long long sbk, long long sbn, float* C, long long ldc, const float* bias, float beta, float* ws,
              size_t ws_floats, int num_sms);
// out[n] = beta*out[n] + sum_m A[m*lda + n]   (column sums of a row-major M x N matrix)
int colsum_f32(cudaStream_t st, int M, int N, const float* A, long long lda, float* out, float beta, float* ws,
                size_t ws_floats, int num_sms);

// ---------------------------------------------------------------- gemm_tc.cu (tcgen05, 3xTF32)
struct TcSeg { const float* p; long long ld; int len; };
struct TcOut {            // destinations of a split-K product: up to 3 row-major column blocks + a per-row vector
  float* p[3]; long long ld[3]; int len[3]; int nseg; float* bias;
};
struct TcArgs {
  int M, N;
  int a_mode, b_mode;     // 0: K-contiguous (elem (r,k) = p[r*ld + k]) ; 1: MN-contiguous (elem (r,k) = p[k*ld + r])
  TcSeg a_k[2], b_k[2];   // mode 0: one source per K segment (segments are padded to multiples of 32 in the k loop)
  TcSeg a_mn[1], b_mn[3]; // mode 1: A single source; B up to 3 column blocks (len = columns of the block)
  int k_nseg, k_len[2];   // K segments (mode 0) or the single reduction length (mode 1 x mode 1)
  int a_vec, b_vec;       // 16-byte loads allowed (ld % 4 == 0 and 16-byte aligned base)
  int b_nseg, b_ones;     // mode-1 B: number of blocks, index of the all-ones column (-1: none)
  float* C; long long ldc; const float* bias; float beta;   // direct epilogue: C = beta*C + D + bias[col]
  float* ws; size_t ws_floats;                              // split-K partials (used when a scatter target is given)
  int BN, nkb, kb_per_split, stages;               // filled by the launcher (stages: smem stages, 1 or 2)
};
// returns kernels launched.  scatter != null: split-K over the reduction, partials reduced deterministically into
// the scatter targets with `beta`.
int gemm_tc(cudaStream_t st, TcArgs g, const TcOut* scatter, int num_sms);
int gemm_tc_configure();
// out(seg)[m][c] = beta*out + sum_z ws[z][m][c] (fixed order), scattered into the column blocks of `o`
void tc_reduce_scatter(cudaStream_t st, int M, int N, int splits, const float* ws, const TcOut& o, float beta, int num_sms);

// ---------------------------------------------------------------- gemm_x.cu (persistent TMA-fed tcgen05 GEMM on fp16 hi/lo planes)
struct GxPlan;                                        // operand-plane scratch of one stream
GxPlan* gemm_x_create(int num_sms);                   // nullptr: driver entry point missing
void gemm_x_destroy(GxPlan* p);
const char* gemm_x_error(const GxPlan* p);
// C = A[M x K] * B[N x K]^T + bias (fp32 row-major operands; values are multiplied by scale_a / scale_b before the fp16 split:
// |value * scale| must stay below 65504).  Returns kernels launched, < 0 on error.
int gemm_x_nt(GxPlan* p, cudaStream_t st, int M, int N, int K, const float* A, long long lda, const float* B, long long ldb, float* C,
              long long ldc, const float* bias, float scale_a, float scale_b);
// partials of A^T [B0 | B1 | 1] over the Kc rows (A [Kc x M], B blocks [Kc x n0 / n1]) into ws as [splits][M][n0 + n1 + 1]
int gemm_x_tn(GxPlan* p, cudaStream_t st, int M, int Kc, const float* A, long long lda, const float* B0, int n0, const float* B1, int n1,
              float scale_a, float scale_b, float* ws, size_t ws_floats, int* splits_out);

// ---------------------------------------------------------------- lstm.cu
// One recurrent block = up to two LSTMs over the same input: slot index d = 0 runs forward in time, d = 1 reversed
// (Reversed{}, clstm.cc:458-479).  A launch covers the `ndir` directions d0 .. d0+ndir-1 (bidirectional: d0=0, ndir=2;
// lstm1: d0=0, ndir=1; revlstm1: d0=1, ndir=1).  H / dH rows hold the directions side by side: direction d occupies
// columns [hoff[d], hoff[d]+no) of rows with `hstride` floats (Parallel's concat, clstm.cc:525-527).
// cell: 0 NPLSTM (SIG,TANH,TANH), 1 LINNPLSTM (SIG,TANH,LIN), 2 RELUTANHNPLSTM (SIG,RELU,TANH), 3 RELUNPLSTM (SIG,RELU,LIN),
// 4 RELU2NPLSTM (SIG,RELU,RELU)  (clstm.cc:655-668); the register / cluster kernels implement cell 0 only.
struct LstmFwdArgs {
  int no;                 // hidden units per direction
  int d0 = 0, ndir = 2, hstride = 0, hoff[2] = {0, 0}, cell = 0;
  const float* XP[2];     // [N][4no] input projection + bias, gate-interleaved rows r = 4*j + g
  const float* R[2];      // [4no][no] recurrent weights, row-major, rows gate-interleaved
  const float* Rt[2];     // [no][4no] transposed copy
  float* G[2];            // [N][4no] gate activations (gi,gf,go,ci)
  float* C[2];            // [N][no] cell states
  float* H;               // [N][hstride] outputs of the block's directions (Parallel concat)
  float* Hprev[2];        // [N][no] output of the previous step in the direction's own time order
};
struct LstmBwdArgs {
  int no;
  int d0 = 0, ndir = 2, hstride = 0, hoff[2] = {0, 0}, cell = 0;
  const float* R[2];
  const float* G[2];
  const float* C[2];
  const float* dH;        // [N][hstride] d(loss)/d(H) from the layer above
  float* DG[2];           // [N][4no] deltas of the gate pre-activations
};
// returns the variant name actually used
const char* lstm_forward(cudaStream_t st, const Lines& ln, const LstmFwdArgs& a);
const char* lstm_backward(cudaStream_t st, const Lines& ln, const LstmBwdArgs& a);
const char* lstm_variant_for(int no);
int lstm_configure();   // opt in to large dynamic smem etc.; returns cudaError_t as int

// the L2-streaming kernels for any size / cell variant (also the comparison baseline of the recurrent self-test)
void lstm_forward_generic(cudaStream_t st, const Lines& ln, const LstmFwdArgs& a);
void lstm_backward_generic(cudaStream_t st, const Lines& ln, const LstmBwdArgs& a);

// ---------------------------------------------------------------- lstm_tc.cu (batched recurrence on tcgen05 + TMEM + TMA)
struct LstmTcPlan;                                   // split weight copies, exchange buffers, tensor maps of one block
bool lstm_tc_supported(int no);
int lstm_tc_configure();
LstmTcPlan* lstm_tc_create(int no, int num_sms);     // nullptr: size not supported / driver entry point missing
void lstm_tc_destroy(LstmTcPlan* p);
void lstm_tc_mark_stale(LstmTcPlan* p);              // the weights changed: the fp16 hi/lo copies are rebuilt on next use
const char* lstm_tc_error(const LstmTcPlan* p);
// 0: launched, -1: not applicable to this batch (use another variant), > 0: CUDA error (see lstm_tc_error)
int lstm_tc_forward(LstmTcPlan* p, cudaStream_t st, const Lines& ln, const LstmFwdArgs& a);
int lstm_tc_backward(LstmTcPlan* p, cudaStream_t st, const Lines& ln, const LstmBwdArgs& a);
// device-side A/B of the tensor-core recurrence against the generic kernels on random data; out[0..4] = max abs
// difference of gates, cell, h, h_prev and (relative) deltas, out[5..8] = ms of tc fwd, tc bwd, generic fwd, generic bwd
// variant 0: lstm_tc.cu (lock-step tiles, L2 exchange), 1: lstm_tcx.cu (cluster-resident, DSMEM exchange)
int lstm_tc_selftest(int no, int B, int Tmin, int Tmax, unsigned seed, float wscale, float* out, char* msg, int msglen, int variant = 0);

// ---------------------------------------------------------------- lstm_tcx.cu (cluster-resident tcgen05 recurrence, DSMEM exchange)
struct LstmTcxPlan;
bool lstm_tcx_supported(int no);                     // nhidden 33..480: clusters of ceil(nhidden / 32) <= 15 CTAs
LstmTcxPlan* lstm_tcx_create(int no, int num_sms);
void lstm_tcx_destroy(LstmTcxPlan* p);
void lstm_tcx_mark_stale(LstmTcxPlan* p);
const char* lstm_tcx_error(const LstmTcxPlan* p);
const long long* lstm_tcx_debug(LstmTcxPlan* p);       // clock stamps of one step of CTA 0 (CLSTM_B200_TC_DBG), else nullptr
int lstm_tcx_forward(LstmTcxPlan* p, cudaStream_t st, const Lines& ln, const LstmFwdArgs& a);    // 0 / -1 n.a. / > 0 error
int lstm_tcx_backward(LstmTcxPlan* p, cudaStream_t st, const Lines& ln, const LstmBwdArgs& a);

// ---------------------------------------------------------------- lstm_cluster.cu (thread-block clusters + DSMEM)
bool lstm_cluster_supported(int no);
int lstm_cluster_configure();
int lstm_cluster_forward(cudaStream_t st, const Lines& ln, const LstmFwdArgs& a);    // cudaError_t as int, -1: size n/a
int lstm_cluster_backward(cudaStream_t st, const Lines& ln, const LstmBwdArgs& a);

// ---------------------------------------------------------------- normalize.cu (text-line normalizers, extras.cc)
constexpr int kNormMaxHeight = 1024;   // rows of a raw line image
struct NormLines {
  const int *W, *H;          // raw width (columns) / height (rows) per line
  const int *poff, *coff;    // offset of the line's pixels / columns inside the concatenated buffers
  const float* masks;        // Gaussian masks, built on the host exactly like gauss1d does
  const int *moff, *mrange;  // [3*b + k]: offset / half width of mask k (0: along y, 1: along x, 2: centre line)
  float range;               // CenterNormalizer::range
};
int norm_configure();
int norm_center_measure(cudaStream_t st, const NormLines& nl, int B, int maxw, int maxh, int maxrange, const float* raw,
                        float* tmp, float* smooth, float* a, float* center, float* r_out);
constexpr int kNormMaxRange = 1 + 3 * 3 * kNormMaxHeight / 2;   // Gaussian masks up to sigma = 1.5 * kNormMaxHeight
int norm_mean_measure(cudaStream_t st, const NormLines& nl, int B, int maxh, const float* raw, double* ymean, double* ymad);
int norm_resample(cudaStream_t st, const NormLines& nl, int B, int maxT, const float* raw, const float* center,
                  const float* scale, const double* ymean, const int* T, const int* off, float* x, int ni, int kind);

// ---------------------------------------------------------------- ctc.cu
struct CtcArgs {
  int nc;
  const float* out;       // [N][nc] softmax outputs
  float* aligned;         // [N][nc]
  float* delta;           // [N][nc] = aligned - out   (clstmhl.h:211-212)
  float* lmatch;          // lattice scratch, per line T x S
  float* lr;              // forward lattice
  float* rl;              // backward lattice
  double* tot;            // per-state partial totals over time: [st_off*8 + slice*S + s] (8 time slices per line)
  float* mx_part;         // [B][8] slice maxima of lr + rl
  int* amax;              // [N] argmax of aligned per column (tensor.h:357-366 tie rule)
  float* amaxv;           // [N] its value
  int* status;            // device int (reserved)
  int raw;                // 1: Lines::L holds the state count S and labels hold one class per state (ctc.cc:136-146)
};
int ctc_align(cudaStream_t st, const Lines& ln, const CtcArgs& a);   // returns kernels launched
int ctc_configure();
constexpr int kLatPad = 8192;        // floats of padding before and after the lattice scratch (prefetch overrun)
constexpr int kCtcMaxStates = 1024;  // S = 2L+1 must not exceed this
constexpr int kCtcMaxClasses = 512;  // nclasses limit of the per-warp class accumulators

// ---------------------------------------------------------------- misc.cu
// out[n][:] = limexp(z[n][:]) / sum  in place (clstm_compute.cc:324-345)
void softmax_rows(cudaStream_t st, float* z, int N, int nc, int* amax, float* amaxv);
// Full<F> output layers (kind 1 SIG, 2 LIN, 3 TANH, 4 RELU): activation in place (+ per-column argmax) and delta <- f'(y) delta
void full_rows(cudaStream_t st, float* z, int N, int nc, int kind, int* amax, float* amaxv);
void full_backward(cudaStream_t st, float* delta, const float* y, size_t n, int kind);
// d += g; g = 0; d = clamp(d); v += lr*d; d *= mom      (clstm_compute.cc:553-563)
void sgd_update(cudaStream_t st, float* v, float* d, float* g, size_t n, float lr, float mom, float clip,
                int fold_only);
// fused share_deltas + sgd_update over NVLink peer memory (misc.cu)
constexpr int kPeerHeaderFloats = 256;            // 1 KiB header in front of g inside the comm buffer
constexpr int kPeerArrive = 0, kPeerDepart = 16, kPeerCounter = 32;   // header slots (32-bit words)
constexpr int kMaxPeers = 8;
struct PeerArgs {
  float* comm[kMaxPeers];   // comm buffer of every rank (own one included), peer-mapped
  int rank, world;
  unsigned epoch;
  float *v, *d;
  size_t n;
  float lr, mom, clip;
  unsigned long long* stats;   // optional device counters {ns spent waiting for the peers, ns in the data phase, launches} (block 0)
};
void peer_allreduce_update(cudaStream_t st, const PeerArgs& a);   // 2 launches

// dst[c][r] = src[r][c] for up to 10 small matrices in one launch (derived weight layouts)
struct TransposeJob { const float* src; float* dst; int rows, cols; };
struct TransposeJobs { TransposeJob job[10]; int n; };
void transpose_batch(cudaStream_t st, const TransposeJobs& jobs);
// trivial_decode per line (ctc.cc:159-194) from the per-column argmax arrays written by softmax_rows / ctc_posterior
void decode_lines(cudaStream_t st, const Lines& ln, const int* argmax_idx, const float* argmax_val, int* classes,
                  int* locs, int* counts, int max_per_line);

}  // namespace cb200
