// capi.cu -- the extern "C" boundary (include/clstm_b200.h): device-resident bidi net, batch staging,
// kernel orchestration on one stream, parameter layout conversion, NCCL gradient sum, measurement hooks.
// No CPU compute path exists here: every entry point needs a CUDA device.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <map>
#include <vector>

#include "../../include/clstm_b200.h"
#include "kernels.h"

using namespace cb200;

namespace {

thread_local std::string g_err;
int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}
#define CU(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define TRY(expr)            \
  do {                       \
    int r_ = (expr);         \
    if (r_) return r_;       \
  } while (0)

enum Phase { PH_H2D, PH_NORMALIZE, PH_XPROJ, PH_LSTM_FWD, PH_SOFTMAX_FWD, PH_CTC, PH_SOFTMAX_BWD, PH_LSTM_BWD, PH_WGRAD, PH_DX,
             PH_ALLREDUCE, PH_UPDATE, PH_DECODE, PH_D2H, PH_COUNT };
const char* kPhaseNames[PH_COUNT] = {"h2d", "normalize", "xproj_gemm", "lstm_fwd", "softmax_fwd", "ctc_align", "softmax_bwd",
                                     "lstm_bwd", "wgrad_gemm", "dx_gemm", "allreduce", "sgd_update", "decode", "d2h"};

// ---- minimal NCCL surface, resolved at run time (the library under torch is reused when already loaded) ----
struct Id128 { char internal[128]; };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id128, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
int load_nccl() {
  if (g_nccl.lib) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return fail("NCCL not found: %s", dlerror());
  g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(h, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
    return fail("NCCL symbols missing");
  g_nccl.lib = h;
  return 0;
}

template <class T>
int dev_alloc(T** p, size_t n) {
  *p = nullptr;
  if (n == 0) n = 1;
  CU(cudaMalloc((void**)p, n * sizeof(T)));
  return 0;
}
template <class T>
void dev_free(T*& p) {
  if (p) cudaFree(p);
  p = nullptr;
}

}  // namespace

struct clstm_b200_net {
  clstm_b200_cfg cfg;
  int ni, no, nc, nf;
  int num_sms = 148;
  cudaStream_t st = nullptr;
  cudaStream_t st2 = nullptr;              // side stream: the W1 derivative product overlaps the backward recurrence
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_fork2 = nullptr, ev_dx = nullptr;
  bool dx_pending = false;                 // the input-delta product is still running on the side stream
  float* ws2 = nullptr;                    // its private split-K workspace
  size_t ws2_floats = 0;

  // ---- topology: nblk stacked recurrent blocks (each 1 or 2 LSTMs over the same input), then an output layer.
  //   bidi   : 1 block {fwd, reversed}          lstm1 : 1 block {fwd}        revlstm1 : 1 block {reversed}
  //   bidi2  : 2 blocks {fwd, reversed}         bidi0 : 1 block, no output layer          (clstm_prefab.cc:22-129)
  struct Block {
    int ni = 0, no = 0, ndir = 2, d0 = 0;  // directions d0 .. d0+ndir-1 (index 0: forward in time, 1: reversed)
    int hoff[2] = {0, 0};                  // column offset of direction d inside H / dH rows
    size_t oWx[2] = {}, oB[2] = {}, oR[2] = {};
    float *Rt[2] = {}, *WxT[2] = {};       // derived layouts, refreshed by prepare_weights(): R^T, Wx^T [ni][4no]
    float *XP[2] = {}, *G[2] = {}, *C[2] = {}, *Hprev[2] = {}, *DG[2] = {};
    float *H = nullptr, *dH = nullptr;     // [N][ndir*no]
    LstmTcPlan* tc = nullptr;              // batched tensor-core recurrence (lstm_tc.cu), nullptr if the size is not covered
    bool tc_used = false;                  // the forward pass of the current batch ran on it (backward follows suit)
    LstmTcxPlan* tcx = nullptr;            // cluster-resident tensor-core recurrence (lstm_tcx.cu), nullptr if the size is not covered
    bool tcx_used = false;
    int nout() const { return ndir * no; }
  } blk[2];
  int nblk = 1;
  int cell = 0;                            // LSTM variant, see kernels.h
  int out_kind = 0;                        // 0 SoftmaxLayer, 1 SigmoidLayer, 2 LinearLayer, 3 TanhLayer, 4 ReluLayer, -1 none
  int nfeat = 0;                           // inputs of the output layer = outputs of the last block
  // ---- parameters, DEVICE layout per block and direction [Wx(4no x ni) | bias(4no) | R(4no x no)], then [W1(nc x nfeat) | b1(nc)]
  // rows of the LSTM matrices are gate-interleaved: r = 4*j + g, g: 0=gi(WGI) 1=gf(WGF) 2=go(WGO) 3=ci(WCI)
  size_t P = 0;
  size_t oW1 = 0, oB1 = 0;
  float *v = nullptr, *d = nullptr, *g = nullptr;   // weights, Params.d (derivative+momentum), this step's derivatives
  float* W1T = nullptr;                    // W1^T [nfeat][nc]
  bool g_pending = false;
  bool use_tc = true;           // dense products on tcgen05 (3xTF32); false: fp32 SIMT tiles (A/B testing)
  int gx_mode = 0;              // persistent TMA-fed GEMM (gemm_x.cu): 0 by size, 1 always, 2 never  (CLSTM_B200_GEMM=x / tc)
  GxPlan* gx[2] = {nullptr, nullptr};   // operand-plane scratch of the main / side stream
  bool gx_small_k = true;       // also products with K < 128 (CLSTM_B200_GX_SMALLK=0 keeps them on the 3xTF32 kernel, A/B runs)
  int lstm_mode = 0;            // recurrence: 0 auto (by size and batch), 1 always the lock-step tensor-core kernels, 2 never, 3 always the cluster-resident ones

  // ---- batch capacity and buffers.  The INPUT SET (x, metadata, tiles, their pinned staging, host geometry, Lines view)
  // exists twice: the members below are the current set, `spare` holds the other one, swap_sets() exchanges them.
  // clstm_b200_prefetch_batch fills the spare set on the copy stream while a step that was launched on the current
  // set is still running (kernels captured its pointers by value).
  int capN = 0, capB = 0, capLab = 0;
  size_t capX = 0;                         // floats in x (per set; capN sizes the compute buffers shared by both sets)
  int capB2 = 0, capLab2 = 0;              // capacity of the shared per-line compute buffers (tot, mx_part, dcnt, h_small)
  cudaStream_t stc = nullptr;              // copy stream of the input pipeline
  cudaEvent_t ev_ready = nullptr, ev_consumed = nullptr;   // per set: staging finished / last step on the set finished
  bool consumed_recorded = false, prefetched = false;
  int dec_B = 0;                           // lines of the batch the decode buffers refer to
  long long capLat = 0;
  float *x = nullptr, *out = nullptr, *aligned = nullptr, *delta = nullptr, *dx = nullptr;
  float *lm = nullptr, *lr = nullptr, *rl = nullptr;
  int* meta = nullptr;          // device: T | off | L | lab_off | order | st_off | labels
  int* tiles = nullptr;         // device: tile_line | tile_t0
  int* h_tiles = nullptr;
  int capTiles = 0;
  double* tot = nullptr;        // CTC per-state totals (8 time-slice partials)
  float* mx_part = nullptr;     // CTC slice maxima
  int capStates = 0;
  long long* lat_off = nullptr;
  int* status = nullptr;
  int* amax[2] = {};
  float* amaxv[2] = {};
  int *dcls[2] = {}, *dlocs[2] = {}, *dcnt[2] = {};
  int capDec = 0;               // max_per_line capacity of the decode buffers
  int decB = 0;                 // lines the decode buffers were sized for (they outlive a regrow of the per-line scratch)
  float* ws = nullptr;
  size_t ws_floats = 0;
  // pinned host staging for metadata and small results
  int* h_meta = nullptr;
  long long* h_lat = nullptr;
  int* h_small = nullptr;       // status + counts
  cudaEvent_t meta_done = nullptr;   // guards re-use of the pinned staging buffers

  std::vector<int> hT, hOff, hL, hLabOff, hOrder, hStOff;
  Lines ln{};
  bool have_batch = false, have_forward = false, have_labels = false, have_ctc = false, raw_targets = false;
  struct InputSet {                        // storage of the set that is not current (same meaning as the members above)
    float* x = nullptr; size_t capX = 0;
    int* meta = nullptr; long long* lat_off = nullptr; int* h_meta = nullptr; long long* h_lat = nullptr;
    int capB = 0, capLab = 0;
    int* tiles = nullptr; int* h_tiles = nullptr; int capTiles = 0;
    cudaEvent_t meta_done = nullptr, ev_ready = nullptr, ev_consumed = nullptr;
    bool consumed_recorded = false;
    std::vector<int> hT, hOff, hL, hLabOff, hOrder, hStOff;
    Lines ln{};
    bool have_batch = false, have_labels = false, raw_targets = false;
  } spare;
  const char* variant = "generic";

  // ---- line normalizer scratch (normalize.cu)
  float *n_raw = nullptr, *n_tmp = nullptr, *n_smooth = nullptr, *n_a = nullptr, *n_center = nullptr, *n_r = nullptr,
        *n_scale = nullptr, *n_masks = nullptr;
  double *n_ym = nullptr, *n_yd = nullptr;
  int* n_meta = nullptr;        // W | H | poff | coff | moff[3B] | mrange[3B]
  size_t n_capPix = 0, n_capCols = 0, n_capMask = 0;
  int n_capB = 0, n_B = 0, n_cols = 0;
  std::vector<float> n_hr;      // r (centre) of the last normalised batch

  // ---- measurement
  bool prof = false;
  struct Rec { int ph; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  float ph_ms[PH_COUNT] = {};
  long long ph_launch[PH_COUNT] = {};

  // ---- communicator
  void* comm = nullptr;
  int rank = 0, world = 1;
  // ---- NVLink peer path: g lives inside a peer-mappable comm buffer [1 KiB header | g]
  float* comm_buf = nullptr;
  float* peer_buf[kMaxPeers] = {};
  bool p2p = false;
  unsigned epoch = 0;
  unsigned long long* peer_stats = nullptr;   // device: {ns waiting for peers, ns data phase, launches} of the fused update kernel
};

namespace {

struct Scope {  // brackets one phase with events when profiling is on
  clstm_b200_net* n;
  int ph;
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t cs;
  Scope(clstm_b200_net* net, int phase, cudaStream_t stream = nullptr) : n(net), ph(phase), cs(stream ? stream : net->st) {
    if (!n->prof) return;
    a = get();
    b = get();
    cudaEventRecord(a, cs);
  }
  cudaEvent_t get() {
    cudaEvent_t e;
    if (!n->pool.empty()) { e = n->pool.back(); n->pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
  }
  void launches(int k) { n->ph_launch[ph] += k; }
  ~Scope() {
    if (!n->prof) return;
    cudaEventRecord(b, cs);
    n->recs.push_back({ph, a, b});
  }
};

void harvest(clstm_b200_net* n) {  // stream must be idle
  for (auto& r : n->recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) n->ph_ms[r.ph] += ms;
    n->pool.push_back(r.a);
    n->pool.push_back(r.b);
  }
  n->recs.clear();
}

int gate_ref_index(int g) {  // position of gate g's matrix inside a direction's reference block [WCI,WGF,WGI,WGO]
  static const int m[4] = {2, 1, 3, 0};
  return m[g];
}

// reference flat (walk_params order: blocks in stacking order, forward LSTM before the reversed one, matrices sorted by
// name WCI,WGF,WGI,WGO, each col-major with the bias column first; then W1) <-> device layout
template <bool TO_DEV, class A, class B>
void convert_params(const clstm_b200_net* n, A* ref, B* dev) {
  size_t pos = 0;
  for (int k = 0; k < n->nblk; k++) {
    const auto& bk = n->blk[k];
    const int ni = bk.ni, no = bk.no, nf = ni + no;
    const size_t msz = (size_t)no * (1 + nf);
    for (int d = bk.d0; d < bk.d0 + bk.ndir; d++) {
      for (int g = 0; g < 4; g++) {
        A* W = ref + pos + (size_t)gate_ref_index(g) * msz;
        for (int j = 0; j < no; j++) {
          const int r = 4 * j + g;
          if (TO_DEV) {
            const_cast<float&>(dev[bk.oB[d] + r]) = W[j];
            for (int i = 0; i < ni; i++) const_cast<float&>(dev[bk.oWx[d] + (size_t)r * ni + i]) = W[j + (size_t)(1 + i) * no];
            for (int q = 0; q < no; q++) const_cast<float&>(dev[bk.oR[d] + (size_t)r * no + q]) = W[j + (size_t)(1 + ni + q) * no];
          } else {
            const_cast<float&>(W[j]) = dev[bk.oB[d] + r];
            for (int i = 0; i < ni; i++) const_cast<float&>(W[j + (size_t)(1 + i) * no]) = dev[bk.oWx[d] + (size_t)r * ni + i];
            for (int q = 0; q < no; q++) const_cast<float&>(W[j + (size_t)(1 + ni + q) * no]) = dev[bk.oR[d] + (size_t)r * no + q];
          }
        }
      }
      pos += 4 * msz;
    }
  }
  if (n->out_kind >= 0) {
    const int nc = n->nc, nfeat = n->nfeat;
    A* W1 = ref + pos;
    for (int c = 0; c < nc; c++) {
      if (TO_DEV) {
        const_cast<float&>(dev[n->oB1 + c]) = W1[c];
        for (int q = 0; q < nfeat; q++) const_cast<float&>(dev[n->oW1 + (size_t)c * nfeat + q]) = W1[c + (size_t)(1 + q) * nc];
      } else {
        const_cast<float&>(W1[c]) = dev[n->oB1 + c];
        for (int q = 0; q < nfeat; q++) const_cast<float&>(W1[c + (size_t)(1 + q) * nc]) = dev[n->oW1 + (size_t)c * nfeat + q];
      }
    }
  }
}
void ref_to_dev(const clstm_b200_net* n, const float* ref, float* dev) { convert_params<true>(n, ref, dev); }
void dev_to_ref(const clstm_b200_net* n, const float* dev, float* ref) { convert_params<false>(n, ref, dev); }

void swap_sets(clstm_b200_net* n) {
  auto& o = n->spare;
  std::swap(n->x, o.x); std::swap(n->capX, o.capX);
  std::swap(n->meta, o.meta); std::swap(n->lat_off, o.lat_off); std::swap(n->h_meta, o.h_meta); std::swap(n->h_lat, o.h_lat);
  std::swap(n->capB, o.capB); std::swap(n->capLab, o.capLab);
  std::swap(n->tiles, o.tiles); std::swap(n->h_tiles, o.h_tiles); std::swap(n->capTiles, o.capTiles);
  std::swap(n->meta_done, o.meta_done); std::swap(n->ev_ready, o.ev_ready); std::swap(n->ev_consumed, o.ev_consumed);
  std::swap(n->consumed_recorded, o.consumed_recorded);
  n->hT.swap(o.hT); n->hOff.swap(o.hOff); n->hL.swap(o.hL); n->hLabOff.swap(o.hLabOff); n->hOrder.swap(o.hOrder);
  n->hStOff.swap(o.hStOff);
  std::swap(n->ln, o.ln);
  std::swap(n->have_batch, o.have_batch); std::swap(n->have_labels, o.have_labels); std::swap(n->raw_targets, o.raw_targets);
}
void free_batch(clstm_b200_net* n) {
  dev_free(n->out); dev_free(n->aligned); dev_free(n->delta); dev_free(n->dx);
  for (int k = 0; k < n->nblk; k++) {
    auto& bk = n->blk[k];
    dev_free(bk.H); dev_free(bk.dH);
    for (int d = 0; d < 2; d++) { dev_free(bk.XP[d]); dev_free(bk.G[d]); dev_free(bk.C[d]); dev_free(bk.Hprev[d]); dev_free(bk.DG[d]); }
  }
  for (int w = 0; w < 2; w++) { dev_free(n->amax[w]); dev_free(n->amaxv[w]); }
}

int ensure_columns(clstm_b200_net* n, int N) {
  if (N <= n->capN) return 0;
  CU(cudaStreamSynchronize(n->st));
  free_batch(n);
  const size_t cap = (size_t)N + N / 8 + 64;
  const int ni = n->ni, nc = n->nc;
  TRY(dev_alloc(&n->out, cap * nc));
  TRY(dev_alloc(&n->aligned, cap * nc));
  TRY(dev_alloc(&n->delta, cap * nc));
  TRY(dev_alloc(&n->dx, cap * ni));
  for (int k = 0; k < n->nblk; k++) {
    auto& bk = n->blk[k];
    TRY(dev_alloc(&bk.H, cap * bk.nout()));
    TRY(dev_alloc(&bk.dH, cap * bk.nout()));
    for (int d = bk.d0; d < bk.d0 + bk.ndir; d++) {
      TRY(dev_alloc(&bk.XP[d], cap * 4 * bk.no));
      TRY(dev_alloc(&bk.G[d], cap * 4 * bk.no));
      TRY(dev_alloc(&bk.C[d], cap * bk.no));
      TRY(dev_alloc(&bk.Hprev[d], cap * bk.no));
      TRY(dev_alloc(&bk.DG[d], cap * 4 * bk.no));
    }
  }
  for (int w = 0; w < 2; w++) {
    TRY(dev_alloc(&n->amax[w], cap));
    TRY(dev_alloc(&n->amaxv[w], cap));
  }
  n->capN = (int)cap;
  return 0;
}
int ensure_x(clstm_b200_net* n, size_t floats) {      // per input set
  if (floats <= n->capX) return 0;
  CU(cudaStreamSynchronize(n->st));
  CU(cudaStreamSynchronize(n->stc));
  dev_free(n->x);
  n->capX = floats + floats / 8 + 64 * (size_t)n->ni;
  TRY(dev_alloc(&n->x, n->capX));
  return 0;
}
int ensure_lines(clstm_b200_net* n, int B, int nlab) {
  if (B > n->capB || nlab > n->capLab) {              // per input set: metadata and its pinned staging
    CU(cudaStreamSynchronize(n->st));
    CU(cudaStreamSynchronize(n->stc));
    const int cb = std::max(B + B / 4 + 8, n->capB), cl = std::max(nlab + nlab / 4 + 64, n->capLab);
    dev_free(n->meta); dev_free(n->lat_off);
    if (n->h_meta) cudaFreeHost(n->h_meta);
    if (n->h_lat) cudaFreeHost(n->h_lat);
    TRY(dev_alloc(&n->meta, (size_t)6 * cb + cl));
    TRY(dev_alloc(&n->lat_off, (size_t)cb));
    CU(cudaHostAlloc((void**)&n->h_meta, ((size_t)6 * cb + cl) * sizeof(int), cudaHostAllocDefault));
    CU(cudaHostAlloc((void**)&n->h_lat, (size_t)cb * sizeof(long long), cudaHostAllocDefault));
    n->capB = cb;
    n->capLab = cl;
  }
  if (B > n->capB2 || nlab > n->capLab2) {            // shared by both sets: per-line compute scratch
    CU(cudaStreamSynchronize(n->st));
    const int cb = std::max(B + B / 4 + 8, n->capB2), cl = std::max(nlab + nlab / 4 + 64, n->capLab2);
    if (n->h_small) cudaFreeHost(n->h_small);
    dev_free(n->tot);
    n->capStates = 2 * cl + cb;
    TRY(dev_alloc(&n->tot, (size_t)n->capStates * 8));
    dev_free(n->mx_part);
    TRY(dev_alloc(&n->mx_part, (size_t)cb * 8));
    CU(cudaHostAlloc((void**)&n->h_small, ((size_t)2 * cb + 8) * sizeof(int), cudaHostAllocDefault));
    // the decode buffers (dcnt / dcls / dlocs) are NOT touched here: in the pipelined order step(i), prefetch(i+1),
    // fetch_decoded(i) they still hold step i's result; ensure_decode() regrows them when the next decode needs more lines
    n->capB2 = cb;
    n->capLab2 = cl;
  }
  return 0;
}
int ensure_tiles(clstm_b200_net* n, int ntiles) {
  if (ntiles <= n->capTiles) return 0;
  CU(cudaStreamSynchronize(n->st));
  CU(cudaStreamSynchronize(n->stc));
  dev_free(n->tiles);
  if (n->h_tiles) cudaFreeHost(n->h_tiles);
  const int cap = ntiles + ntiles / 4 + 64;
  TRY(dev_alloc(&n->tiles, (size_t)2 * cap));
  CU(cudaHostAlloc((void**)&n->h_tiles, (size_t)2 * cap * sizeof(int), cudaHostAllocDefault));
  n->capTiles = cap;
  return 0;
}
int ensure_decode(clstm_b200_net* n, int max_per_line) {
  if (max_per_line <= n->capDec && n->capB2 <= n->decB) return 0;
  CU(cudaStreamSynchronize(n->st));
  max_per_line = std::max(max_per_line, n->capDec);
  for (int w = 0; w < 2; w++) {
    dev_free(n->dcls[w]); dev_free(n->dlocs[w]); dev_free(n->dcnt[w]);
    TRY(dev_alloc(&n->dcls[w], (size_t)n->capB2 * max_per_line));
    TRY(dev_alloc(&n->dlocs[w], (size_t)n->capB2 * max_per_line));
    TRY(dev_alloc(&n->dcnt[w], (size_t)n->capB2));
  }
  n->capDec = max_per_line;
  n->decB = n->capB2;
  return 0;
}
int ensure_lattice(clstm_b200_net* n, long long elems) {
  if (elems <= n->capLat) return 0;
  CU(cudaStreamSynchronize(n->st));
  dev_free(n->lm); dev_free(n->lr); dev_free(n->rl);
  const long long cap = elems + elems / 8 + 1024;
  const size_t alloc = (size_t)cap + 2 * (size_t)kLatPad;      // the lattice warps prefetch past the line's rows
  TRY(dev_alloc(&n->lm, alloc));
  TRY(dev_alloc(&n->lr, alloc));
  TRY(dev_alloc(&n->rl, alloc));
  CU(cudaMemsetAsync(n->lm, 0, alloc * sizeof(float), n->st));
  n->capLat = cap;
  return 0;
}

// Stage line lengths (and optionally transcripts) and publish the Lines view.
int stage_lines(clstm_b200_net* n, const int* T, int B, const int* labels, const int* L, bool raw = false,
                cudaStream_t cs = nullptr) {
  if (!cs) cs = n->st;
  if (B <= 0) return fail("batch must contain at least one line");
  long long N = 0;
  for (int b = 0; b < B; b++) {
    if (T[b] <= 0) return fail("line %d has non-positive length %d", b, T[b]);
    N += T[b];
  }
  const int widest = std::max(std::max(n->blk[0].no, n->nblk > 1 ? n->blk[1].no : 0), 16);
  if (N > 0x7fffffff / (4LL * widest)) return fail("batch too large: %lld columns", N);
  int nlab = 0;
  if (L) for (int b = 0; b < B; b++) {
    if (L[b] < 0) return fail("line %d has negative transcript length", b);
    if ((raw ? L[b] : 2 * L[b] + 1) > kCtcMaxStates) return fail("transcript of line %d too long (%d labels, max %d)", b, L[b], (kCtcMaxStates - 1) / 2);
    if (raw && L[b] < 1) return fail("line %d needs at least one target state", b);
    nlab += L[b];
  }
  TRY(ensure_columns(n, (int)N));
  TRY(ensure_x(n, (size_t)N * n->ni));
  TRY(ensure_lines(n, B, nlab));
  if (n->meta_done) CU(cudaEventSynchronize(n->meta_done));   // pinned staging buffers free again?
  n->hT.assign(T, T + B);
  n->hOff.resize(B); n->hL.assign(B, 0); n->hLabOff.assign(B, 0); n->hOrder.resize(B);
  int off = 0, tmax = 0;
  for (int b = 0; b < B; b++) { n->hOff[b] = off; off += T[b]; tmax = std::max(tmax, T[b]); }
  std::iota(n->hOrder.begin(), n->hOrder.end(), 0);
  std::stable_sort(n->hOrder.begin(), n->hOrder.end(), [&](int a, int c) { return T[a] > T[c]; });
  long long lat = 0;
  int lo = 0, so = 0, ntiles = 0;
  for (int b = 0; b < B; b++) ntiles += (T[b] + 31) / 32;
  TRY(ensure_tiles(n, ntiles));
  if (n->meta_done) CU(cudaEventSynchronize(n->meta_done));
  n->hStOff.assign(B, 0);
  ntiles = 0;
  for (int b = 0; b < B; b++) {
    if (L) { n->hL[b] = L[b]; n->hLabOff[b] = lo; lo += L[b]; }
    const int Sb = raw ? n->hL[b] : 2 * n->hL[b] + 1;
    n->hStOff[b] = so;
    so += Sb;
    n->h_lat[b] = lat + kLatPad;
    lat += (long long)T[b] * Sb;
    for (int t0 = 0; t0 < T[b]; t0 += 32) {
      n->h_tiles[ntiles] = b;
      n->h_tiles[n->capTiles + ntiles] = t0;
      ntiles++;
    }
  }
  if (L) {
    for (int i = 0; i < nlab; i++)
      if (labels[i] < 0 || labels[i] >= n->nc) return fail("label %d out of range [0,%d)", labels[i], n->nc);
    TRY(ensure_lattice(n, lat));
  }
  int* hm = n->h_meta;
  const int cb = n->capB;
  memcpy(hm + 0 * cb, n->hT.data(), B * sizeof(int));
  memcpy(hm + 1 * cb, n->hOff.data(), B * sizeof(int));
  memcpy(hm + 2 * cb, n->hL.data(), B * sizeof(int));
  memcpy(hm + 3 * cb, n->hLabOff.data(), B * sizeof(int));
  memcpy(hm + 4 * cb, n->hOrder.data(), B * sizeof(int));
  memcpy(hm + 5 * cb, n->hStOff.data(), B * sizeof(int));
  if (L && nlab) memcpy(hm + 6 * cb, labels, nlab * sizeof(int));
  CU(cudaMemcpyAsync(n->meta, hm, ((size_t)6 * cb + nlab) * sizeof(int), cudaMemcpyHostToDevice, cs));
  CU(cudaMemcpyAsync(n->tiles, n->h_tiles, ((size_t)n->capTiles + ntiles) * sizeof(int), cudaMemcpyHostToDevice, cs));
  CU(cudaMemcpyAsync(n->lat_off, n->h_lat, (size_t)B * sizeof(long long), cudaMemcpyHostToDevice, cs));
  if (!n->meta_done) CU(cudaEventCreateWithFlags(&n->meta_done, cudaEventDisableTiming));
  CU(cudaEventRecord(n->meta_done, cs));
  Lines& ln = n->ln;
  ln.B = B; ln.N = (int)N; ln.Tmax = tmax;
  ln.T = n->meta; ln.off = n->meta + cb; ln.L = n->meta + 2 * cb; ln.lab_off = n->meta + 3 * cb;
  ln.order = n->meta + 4 * cb; ln.st_off = n->meta + 5 * cb; ln.labels = n->meta + 6 * cb; ln.lat_off = n->lat_off;
  ln.ntiles = ntiles; ln.tile_line = n->tiles; ln.tile_t0 = n->tiles + n->capTiles;
  n->have_batch = true;
  n->have_labels = (L != nullptr);
  n->raw_targets = raw;
  n->have_forward = false;
  n->have_ctc = false;
  return 0;
}

// ---- normalizers -------------------------------------------------------------------------------------------------
// gauss1d's mask (extras.cc:60-69), built with the host libm so that the device pass reproduces the reference bit for bit
void host_gauss_mask(float sigma, std::vector<float>& mask, int& range) {
  range = 1 + int(3.0 * sigma);
  mask.assign(2 * range + 1, 0.f);
  for (int i = 0; i <= range; i++) {
    double y = exp(-i * i / 2.0 / sigma / sigma);
    mask[range + i] = mask[range - i] = (float)y;
  }
  float total = 0.0;
  for (size_t i = 0; i < mask.size(); i++) total += mask[i];
  for (size_t i = 0; i < mask.size(); i++) mask[i] /= total;
}
int ensure_norm(clstm_b200_net* n, size_t pix, size_t cols, int B, size_t nmask) {
  if (pix > n->n_capPix || cols > n->n_capCols || B > n->n_capB || nmask > n->n_capMask) {
    CU(cudaStreamSynchronize(n->st));
    CU(cudaStreamSynchronize(n->stc));
    dev_free(n->n_raw); dev_free(n->n_tmp); dev_free(n->n_smooth); dev_free(n->n_a); dev_free(n->n_center);
    dev_free(n->n_r); dev_free(n->n_scale); dev_free(n->n_masks); dev_free(n->n_ym); dev_free(n->n_yd); dev_free(n->n_meta);
    n->n_capPix = std::max(pix + pix / 8 + 1024, n->n_capPix);
    n->n_capCols = std::max(cols + cols / 8 + 256, n->n_capCols);
    n->n_capB = std::max(B + B / 4 + 8, n->n_capB);
    n->n_capMask = std::max(nmask + nmask / 4 + 1024, n->n_capMask);
    TRY(dev_alloc(&n->n_raw, n->n_capPix)); TRY(dev_alloc(&n->n_tmp, n->n_capPix)); TRY(dev_alloc(&n->n_smooth, n->n_capPix));
    TRY(dev_alloc(&n->n_a, n->n_capCols)); TRY(dev_alloc(&n->n_center, n->n_capCols));
    TRY(dev_alloc(&n->n_r, (size_t)n->n_capB)); TRY(dev_alloc(&n->n_scale, (size_t)n->n_capB));
    TRY(dev_alloc(&n->n_ym, (size_t)n->n_capB)); TRY(dev_alloc(&n->n_yd, (size_t)n->n_capB));
    TRY(dev_alloc(&n->n_masks, n->n_capMask));
    TRY(dev_alloc(&n->n_meta, (size_t)10 * n->n_capB));
  }
  return 0;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("%s launch failed: %s", what, cudaGetErrorString(e));
  return 0;
}

void prepare_weights(clstm_b200_net* n) {   // after every change of v
  TransposeJobs j{};
  for (int k = 0; k < n->nblk; k++) {
    auto& bk = n->blk[k];
    lstm_tc_mark_stale(bk.tc);
    lstm_tcx_mark_stale(bk.tcx);
    for (int d = bk.d0; d < bk.d0 + bk.ndir; d++) {
      j.job[j.n++] = {n->v + bk.oR[d], bk.Rt[d], 4 * bk.no, bk.no};
      j.job[j.n++] = {n->v + bk.oWx[d], bk.WxT[d], 4 * bk.no, bk.ni};
    }
  }
  if (n->out_kind >= 0) j.job[j.n++] = {n->v + n->oW1, n->W1T, n->nc, n->nfeat};
  transpose_batch(n->st, j);
}
// ------------------------------------------------------------------------------------------------ dense products
bool vec_ok(const float* p, long long ld) { return (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0); }

// C[M x N] = beta*C + A[M x K] * B^T (+bias) with A K-contiguous; B either K-contiguous ([N][K], b_mn=false) or
// MN-contiguous ([K][N], b_mn=true)
// The persistent TMA-fed GEMM pays two conversion launches per product: it takes over where the products are real GEMMs
// (measured crossover on B200: a few GFLOP per product; BASELINE config 2 stays on the on-the-fly 3xTF32 kernels).
bool want_gx(const clstm_b200_net* n, int side, double flops) {
  if (!n->use_tc || !n->gx[side] || n->gx_mode == 2) return false;
  if (n->cell != 0 && n->gx_mode != 1) return false;   // LIN / RELU cells: outputs are unbounded, the fp16 hi/lo planes saturate at 4096
  return n->gx_mode == 1 || flops >= 8e9;
}
int dense_nt(clstm_b200_net* n, int M, int N, int K, const float* A, long long lda, const float* B, long long ldb,
             bool b_mn, float* C, long long ldc, const float* bias, float beta, float scale_a = 16.f, float scale_b = 16.f) {
  if (!n->use_tc) {
    return gemm_f32(n->st, M, N, K, A, lda, 1, B, b_mn ? ldb : 1, b_mn ? 1 : ldb, C, ldc, bias, beta, nullptr, 0,
                    n->num_sms);
  }
  // (products with a short reduction -- the input projection, K = 48 -- are bound by writing their output; with the bias tile
  // staged in shared memory the TMA-fed kernel wins there too: 1.57 against 1.97 ms at nhidden 400 x 256 lines)
  if (!b_mn && beta == 0.f && (K >= 128 || n->gx_small_k) && want_gx(n, 0, 2.0 * M * N * K)) {
    const int r = gemm_x_nt(n->gx[0], n->st, M, N, K, A, lda, B, ldb, C, ldc, bias, scale_a, scale_b);
    if (r >= 0) return r;
    cudaGetLastError();                                         // (out of scratch memory: the on-the-fly kernels below need none)
  }
  TcArgs g{};
  g.M = M; g.N = N;
  g.a_mode = 0; g.a_k[0] = {A, lda, K}; g.a_vec = vec_ok(A, lda);
  g.k_nseg = 1; g.k_len[0] = K;
  if (b_mn) { g.b_mode = 1; g.b_mn[0] = {B, ldb, N}; g.b_nseg = 1; g.b_ones = -1; g.b_vec = vec_ok(B, ldb); }
  else { g.b_mode = 0; g.b_k[0] = {B, ldb, K}; g.b_vec = vec_ok(B, ldb); }
  g.C = C; g.ldc = ldc; g.bias = bias; g.beta = beta;
  return gemm_tc(n->st, g, nullptr, n->num_sms);
}

// derivative product reduced over all columns: out += A^T [B0 | B1 | 1], A [K x M] and B blocks [K x len] row-major
int dense_tn(clstm_b200_net* n, int M, int K, const float* A, long long lda, const float* B0, int n0, const float* B1,
             int n1, float* out0, float* out1, float* out_bias, bool side = false) {
  cudaStream_t stream = side ? n->st2 : n->st;
  float* ws = side ? n->ws2 : n->ws;
  const size_t wsf = side ? n->ws2_floats : n->ws_floats;
  if (want_gx(n, side ? 1 : 0, 2.0 * M * (n0 + n1 + 1) * K)) {
    int splits = 1;
    const int r = gemm_x_tn(n->gx[side ? 1 : 0], stream, M, K, A, lda, B0, n0, B1, n1, 256.f, 16.f, ws, wsf, &splits);
    if (r >= 0) {
      TcOut o{};
      o.p[0] = out0; o.ld[0] = n0; o.len[0] = n0; o.nseg = 1;
      if (B1) { o.p[1] = out1; o.ld[1] = n1; o.len[1] = n1; o.nseg = 2; }
      o.bias = out_bias;
      tc_reduce_scatter(stream, M, n0 + n1 + 1, splits, ws, o, 1.f, n->num_sms);
      return r + 1;
    }
    cudaGetLastError();
  }
  if (!n->use_tc) {
    int k = gemm_f32(stream, M, n0, K, A, 1, lda, B0, n0, 1, out0, n0, nullptr, 1.f, ws, wsf, n->num_sms);
    if (B1) k += gemm_f32(stream, M, n1, K, A, 1, lda, B1, n1, 1, out1, n1, nullptr, 1.f, ws, wsf, n->num_sms);
    k += colsum_f32(stream, K, M, A, lda, out_bias, 1.f, ws, wsf, n->num_sms);
    return k;
  }
  TcArgs g{};
  g.M = M; g.N = n0 + n1 + 1;
  g.a_mode = 1; g.a_mn[0] = {A, lda, M}; g.a_vec = vec_ok(A, lda);
  g.b_mode = 1; g.b_mn[0] = {B0, n0, n0}; g.b_nseg = 1;
  g.b_vec = vec_ok(B0, n0) && (n0 % 4 == 0);
  if (B1) { g.b_mn[1] = {B1, n1, n1}; g.b_nseg = 2; g.b_vec = g.b_vec && vec_ok(B1, n1); }
  g.b_ones = n0 + n1;
  g.k_nseg = 1; g.k_len[0] = K;
  g.beta = 1.f; g.ws = ws; g.ws_floats = wsf;
  TcOut o{};
  o.p[0] = out0; o.ld[0] = n0; o.len[0] = n0; o.nseg = 1;
  if (B1) { o.p[1] = out1; o.ld[1] = n1; o.len[1] = n1; o.nseg = 2; }
  o.bias = out_bias;
  return gemm_tc(stream, g, &o, n->num_sms);
}

// ------------------------------------------------------------------------------------------------ device passes
// The batched tensor-core recurrence advances all lines of a 128-slot tile in lock step: its time per pass is
// (longest line) x (step latency), nearly independent of the number of lines, while the register / cluster kernels run one
// chain per (line, direction) at 0.3-1 us per step but only as many chains at a time as fit the SMs.  Crossovers measured
// on B200 (T = 200..2000): nhidden 400: cluster 0.56 ms per line vs tensor-core 37..46 ms per batch => 64 lines;
// nhidden 200: cluster 0.125 ms per line vs 31 ms per batch => ~256 lines; nhidden <= 100: the register kernels win
// at every batch size; beyond 400 only the L2-streaming generic kernels remain, so the tensor-core path always wins.
bool want_lstm_tc(const clstm_b200_net* n, const clstm_b200_net::Block& bk, int B) {
  if (!bk.tc || n->cell != 0 || n->lstm_mode == 2 || n->lstm_mode == 3) return false;
  if (n->lstm_mode == 1) return true;
  const int no = bk.no;
  if (no > 400) return B >= 4;
  if (no == 400) return B >= 64;
  if (no > 200) return B >= 48;                   // no cluster instance for these widths: generic kernels otherwise
  if (no == 200) return B >= 224;
  if (no > 100) return B >= 64;                   // generic kernels otherwise
  return false;
}

// The cluster-resident tensor-core recurrence (lstm_tcx.cu: 16 lines per cluster, weights in shared / tensor memory, DSMEM
// exchange): nhidden 33..480.  Measured on B200 (forward + backward): nhidden 200, 128 lines, T 200..2000: 9.1 ms against
// 15.2 ms for the 4-CTA cluster kernels; 32 lines x 512 steps: 2 ms against 7.4 ms; nhidden 400, 256 lines: 39.5 ms against
// 44 ms for the lock-step kernels of lstm_tc.cu, 32 lines x 512 steps: 4 ms against 11.7 ms.  At nhidden <= 100 the register
// kernels (one SM per chain, 0.6-0.8 us per step) stay ahead at every batch size (32 lines x 500 steps: 0.69 ms vs 1.8 ms).
bool want_lstm_tcx(const clstm_b200_net* n, const clstm_b200_net::Block& bk, int B) {
  if (!bk.tcx || n->cell != 0 || n->lstm_mode == 1 || n->lstm_mode == 2) return false;
  if (n->lstm_mode == 3) return true;
  return bk.no > 100;
}

const float* block_input(const clstm_b200_net* n, int k) { return k == 0 ? n->x : n->blk[k - 1].H; }
int run_forward(clstm_b200_net* n) {
  const Lines& ln = n->ln;
  const int N = ln.N, nc = n->nc;
  for (int k = 0; k < n->nblk; k++) {
    auto& bk = n->blk[k];
    const float* in = block_input(n, k);
    {
      Scope s(n, PH_XPROJ);
      for (int d = bk.d0; d < bk.d0 + bk.ndir; d++)
        s.launches(dense_nt(n, N, 4 * bk.no, bk.ni, in, bk.ni, n->v + bk.oWx[d], bk.ni, false, bk.XP[d], 4 * bk.no, n->v + bk.oB[d], 0.f));
    }
    {
      Scope s(n, PH_LSTM_FWD);
      LstmFwdArgs a;
      a.no = bk.no; a.d0 = bk.d0; a.ndir = bk.ndir; a.hstride = bk.nout(); a.cell = n->cell;
      for (int d = 0; d < 2; d++) {
        a.hoff[d] = bk.hoff[d];
        a.XP[d] = bk.XP[d]; a.R[d] = n->v + bk.oR[d]; a.Rt[d] = bk.Rt[d];
        a.G[d] = bk.G[d]; a.C[d] = bk.C[d]; a.Hprev[d] = bk.Hprev[d];
      }
      a.H = bk.H;
      const char* var = nullptr;
      bk.tc_used = bk.tcx_used = false;
      if (want_lstm_tcx(n, bk, ln.B)) {
        const int r = lstm_tcx_forward(bk.tcx, n->st, ln, a);
        if (r > 0) return fail("%s", lstm_tcx_error(bk.tcx));
        if (r == 0) { var = "tcx"; bk.tcx_used = true; s.launches(1); }
      }
      if (!var && want_lstm_tc(n, bk, ln.B)) {
        const int r = lstm_tc_forward(bk.tc, n->st, ln, a);
        if (r > 0) return fail("%s", lstm_tc_error(bk.tc));
        if (r == 0) { var = "tc"; bk.tc_used = true; s.launches(1); }
      }
      if (!var) var = lstm_forward(n->st, ln, a);
      if (k == 0) n->variant = var;
      s.launches(1);
    }
  }
  {
    Scope s(n, PH_SOFTMAX_FWD);
    const auto& last = n->blk[n->nblk - 1];
    if (n->out_kind < 0) {          // bidi0: the block's outputs are the network's outputs
      CU(cudaMemcpyAsync(n->out, last.H, (size_t)N * nc * sizeof(float), cudaMemcpyDeviceToDevice, n->st));
      full_rows(n->st, n->out, N, nc, 2, n->amax[0], n->amaxv[0]);
      s.launches(1);
    } else {
      s.launches(dense_nt(n, N, nc, n->nfeat, last.H, n->nfeat, n->v + n->oW1, n->nfeat, false, n->out, nc, n->v + n->oB1, 0.f));
      if (n->out_kind == 0) softmax_rows(n->st, n->out, N, nc, n->amax[0], n->amaxv[0]);
      else full_rows(n->st, n->out, N, nc, n->out_kind, n->amax[0], n->amaxv[0]);
      s.launches(1);
    }
  }
  TRY(check_launch("forward"));
  n->have_forward = true;
  n->have_ctc = false;
  return 0;
}

int run_ctc(clstm_b200_net* n) {
  if (n->out_kind != 0) return fail("CTC alignment needs a SoftmaxLayer output (this net has output kind %d)", n->out_kind);
  Scope s(n, PH_CTC);
  CtcArgs a;
  a.nc = n->nc; a.out = n->out; a.aligned = n->aligned; a.delta = n->delta;
  a.lmatch = n->lm; a.lr = n->lr; a.rl = n->rl; a.tot = n->tot; a.mx_part = n->mx_part; a.amax = n->amax[1]; a.amaxv = n->amaxv[1]; a.status = n->status; a.raw = n->raw_targets ? 1 : 0;
  s.launches(ctc_align(n->st, n->ln, a));
  TRY(check_launch("ctc_align"));
  n->have_ctc = true;
  return 0;
}

// defer_dx: the input-delta product (nobody downstream of the step needs it) goes to the side stream and is joined by
// join_dx() at the end of the step, so it overlaps the weight-derivative tail, the gradient exchange and the update.
void join_dx(clstm_b200_net* n) {
  if (!n->dx_pending) return;
  cudaStreamWaitEvent(n->st, n->ev_dx, 0);
  n->dx_pending = false;
}
int run_backward(clstm_b200_net* n, bool defer_dx = false) {
  const Lines& ln = n->ln;
  const int N = ln.N, nc = n->nc, nfeat = n->nfeat;
  auto& last = n->blk[n->nblk - 1];
  {
    Scope s(n, PH_SOFTMAX_BWD);   // backward_softmax clstm_compute.cc:346-356 / backward_full :316-320
    if (n->out_kind < 0) {
      CU(cudaMemcpyAsync(last.dH, n->delta, (size_t)N * nc * sizeof(float), cudaMemcpyDeviceToDevice, n->st));
    } else {
      if (n->out_kind > 0) { full_backward(n->st, n->delta, n->out, (size_t)N * nc, n->out_kind); s.launches(1); }
      if (n->use_tc) s.launches(dense_nt(n, N, nfeat, nc, n->delta, nc, n->W1T, nc, false, last.dH, nfeat, nullptr, 0.f, 256.f, 16.f));
      else s.launches(dense_nt(n, N, nfeat, nc, n->delta, nc, n->v + n->oW1, nfeat, true, last.dH, nfeat, nullptr, 0.f));
      // the W1 derivative product needs only delta and H: run it on the side stream, concurrently with the backward
      // recurrence (which occupies 2B of the 148 SMs), and join before anything consumes g
      cudaEventRecord(n->ev_fork, n->st);
      cudaStreamWaitEvent(n->st2, n->ev_fork, 0);
      s.launches(dense_tn(n, nc, N, n->delta, nc, last.H, nfeat, nullptr, 0, n->g + n->oW1, nullptr, n->g + n->oB1, true));
      cudaEventRecord(n->ev_join, n->st2);
    }
  }
  for (int k = n->nblk - 1; k >= 0; k--) {
    auto& bk = n->blk[k];
    const int ni = bk.ni, no = bk.no;
    {
      Scope s(n, PH_LSTM_BWD);
      LstmBwdArgs a;
      a.no = no; a.d0 = bk.d0; a.ndir = bk.ndir; a.hstride = bk.nout(); a.cell = n->cell; a.dH = bk.dH;
      for (int d = 0; d < 2; d++) {
        a.hoff[d] = bk.hoff[d];
        a.R[d] = n->v + bk.oR[d]; a.G[d] = bk.G[d]; a.C[d] = bk.C[d]; a.DG[d] = bk.DG[d];
      }
      bool done = false;
      if (bk.tcx_used) {
        const int r = lstm_tcx_backward(bk.tcx, n->st, ln, a);
        if (r > 0) return fail("%s", lstm_tcx_error(bk.tcx));
        done = (r == 0);
        if (done) s.launches(1);
      }
      if (!done && bk.tc_used) {
        const int r = lstm_tc_backward(bk.tc, n->st, ln, a);
        if (r > 0) return fail("%s", lstm_tc_error(bk.tc));
        done = (r == 0);
        if (done) s.launches(1);
      }
      if (!done) lstm_backward(n->st, ln, a);
      s.launches(1);
    }
    cudaStream_t dxs = n->st;
    if (k == 0 && defer_dx && n->use_tc && !n->prof) {   // (phase profiling keeps everything on one stream)
      cudaEventRecord(n->ev_fork2, n->st);
      cudaStreamWaitEvent(n->st2, n->ev_fork2, 0);
      dxs = n->st2;
    }
    {
      Scope s(n, PH_WGRAD);   // W.d += delta * src^T over all columns (backward_lin1 clstm_compute.cc:297-298)
      for (int d = bk.d0; d < bk.d0 + bk.ndir; d++)
        s.launches(dense_tn(n, 4 * no, N, bk.DG[d], 4 * no, block_input(n, k), ni, bk.Hprev[d], no, n->g + bk.oWx[d],
                            n->g + bk.oR[d], n->g + bk.oB[d]));
    }
    {
      Scope s(n, PH_DX);      // deltas of the block's input = sum over its directions of Wx^T delta (clstm.cc:537-541)
      float* din = (k == 0) ? n->dx : n->blk[k - 1].dH;
      if (n->use_tc) {   // one product over the directions: K = [DG(d0) | DG(d0+1)], B = [Wx ; Wx']
        TcArgs g{};
        g.M = N; g.N = ni;
        g.a_mode = 0; g.b_mode = 0; g.k_nseg = bk.ndir;
        g.a_vec = 1; g.b_vec = 1;
        for (int q = 0; q < bk.ndir; q++) {
          const int d = bk.d0 + q;
          g.a_k[q] = {bk.DG[d], 4 * no, 4 * no};
          g.b_k[q] = {bk.WxT[d], 4 * no, 4 * no};       // Wx^T [ni][4no]: K-contiguous B operand
          g.k_len[q] = 4 * no;
          g.a_vec = g.a_vec && vec_ok(bk.DG[d], 4 * no);
          g.b_vec = g.b_vec && vec_ok(bk.WxT[d], 4 * no);
        }
        g.C = din; g.ldc = ni; g.bias = nullptr; g.beta = 0.f;
        s.launches(gemm_tc(dxs, g, nullptr, n->num_sms));
        if (dxs != n->st) {
          cudaEventRecord(n->ev_dx, n->st2);
          n->dx_pending = true;
        }
      } else {
        for (int q = 0; q < bk.ndir; q++) {
          const int d = bk.d0 + q;
          s.launches(dense_nt(n, N, ni, 4 * no, bk.DG[d], 4 * no, n->v + bk.oWx[d], ni, true, din, ni, nullptr, q ? 1.f : 0.f));
        }
      }
    }
  }
  if (n->out_kind >= 0) cudaStreamWaitEvent(n->st, n->ev_join, 0);
  TRY(check_launch("backward"));
  n->g_pending = true;
  return 0;
}

int fold_pending(clstm_b200_net* n) {
  if (!n->g_pending) return 0;
  sgd_update(n->st, n->v, n->d, n->g, n->P, 0.f, 0.f, 0.f, 1);
  n->g_pending = false;
  return check_launch("fold");
}

int run_allreduce(clstm_b200_net* n) {
  if (!n->comm || n->world <= 1) return 0;
  Scope s(n, PH_ALLREDUCE);
  int r = g_nccl.AllReduce(n->g, n->g, n->P, /*ncclFloat*/ 7, /*ncclSum*/ 0, n->comm, n->st);
  s.launches(1);
  if (r != 0) return fail("ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  return 0;
}

int run_update(clstm_b200_net* n, float lr, float mom, float clip) {
  Scope s(n, PH_UPDATE);
  sgd_update(n->st, n->v, n->d, n->g, n->P, lr, mom, clip, 0);
  n->g_pending = false;
  join_dx(n);              // the deferred input-delta product reads Wx^T, which prepare_weights is about to rewrite
  prepare_weights(n);
  s.launches(2);
  return check_launch("sgd_update");
}

// share_deltas + sgd_update fused over NVLink peer memory (all ranks must call this the same number of times)
int run_peer_update(clstm_b200_net* n, float lr, float mom, float clip) {
  Scope s(n, PH_ALLREDUCE);
  PeerArgs a{};
  for (int r = 0; r < n->world; r++) a.comm[r] = n->peer_buf[r];
  a.rank = n->rank; a.world = n->world; a.epoch = ++n->epoch;
  a.v = n->v; a.d = n->d; a.n = n->P; a.lr = lr; a.mom = mom; a.clip = clip;
  a.stats = n->peer_stats;
  peer_allreduce_update(n->st, a);
  n->g_pending = false;
  join_dx(n);              // see run_update
  prepare_weights(n);
  s.launches(3);
  return check_launch("peer_allreduce_update");
}

int run_decode(clstm_b200_net* n, int which, int max_per_line) {
  TRY(ensure_decode(n, max_per_line));
  Scope s(n, PH_DECODE);
  decode_lines(n->st, n->ln, n->amax[which], n->amaxv[which], n->dcls[which], n->dlocs[which], n->dcnt[which],
               n->capDec);
  n->dec_B = n->ln.B;
  s.launches(1);
  return check_launch("decode");
}

int fetch_decode(clstm_b200_net* n, int which, int* classes, int* locs, int* counts, int max_per_line) {
  const int B = n->dec_B;                 // the batch that was decoded (a prefetched batch may be current by now)
  {
    Scope s(n, PH_D2H);
    CU(cudaMemcpyAsync(counts, n->dcnt[which], (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, n->st));
    if (n->capDec == max_per_line) {
      CU(cudaMemcpyAsync(classes, n->dcls[which], (size_t)B * max_per_line * sizeof(int), cudaMemcpyDeviceToHost, n->st));
      CU(cudaMemcpyAsync(locs, n->dlocs[which], (size_t)B * max_per_line * sizeof(int), cudaMemcpyDeviceToHost, n->st));
    } else {
      CU(cudaMemcpy2DAsync(classes, (size_t)max_per_line * sizeof(int), n->dcls[which], (size_t)n->capDec * sizeof(int),
                           (size_t)max_per_line * sizeof(int), B, cudaMemcpyDeviceToHost, n->st));
      CU(cudaMemcpy2DAsync(locs, (size_t)max_per_line * sizeof(int), n->dlocs[which], (size_t)n->capDec * sizeof(int),
                           (size_t)max_per_line * sizeof(int), B, cudaMemcpyDeviceToHost, n->st));
    }
  }
  CU(cudaStreamSynchronize(n->st));
  for (int b = 0; b < B; b++)
    if (counts[b] > max_per_line) return fail("line %d decodes to %d symbols > max_per_line %d", b, counts[b], max_per_line);
  return 0;
}

int check_status(clstm_b200_net* n) {  // stream must be synchronised
  if (n->h_small[0] != 0) return fail("ctc_align: transcript too long for the device lattice (status %d)", n->h_small[0]);
  return 0;
}

}  // namespace

// =================================================================================================== C ABI
extern "C" {

const char* clstm_b200_last_error(void) { return g_err.c_str(); }
const char* clstm_b200_version(void) { return "clstm_b200 0.1 (sm_100a)"; }
int clstm_b200_num_phases(void) { return PH_COUNT; }
const char* clstm_b200_phase_name(int i) { return (i >= 0 && i < PH_COUNT) ? kPhaseNames[i] : ""; }

int clstm_b200_create(const clstm_b200_cfg* cfg, clstm_b200_net** out) {
  if (!cfg || !out) return fail("null argument");
  clstm_b200_cfg_ex ex;
  memset(&ex, 0, sizeof ex);
  ex.ninput = cfg->ninput; ex.noutput = cfg->nclasses; ex.device = cfg->device;
  ex.nblocks = 1; ex.nhidden[0] = cfg->nhidden; ex.direction[0] = 2; ex.cell = 0; ex.output = 0;
  return clstm_b200_create_ex(&ex, out);
}

int clstm_b200_create_ex(const clstm_b200_cfg_ex* cfg, clstm_b200_net** out) {
  if (!cfg || !out) return fail("null argument");
  *out = nullptr;
  if (cfg->nblocks < 1 || cfg->nblocks > 2) return fail("nblocks must be 1 or 2 (got %d)", cfg->nblocks);
  if (cfg->cell < 0 || cfg->cell > 4) return fail("unknown LSTM cell variant %d", cfg->cell);
  if (cfg->output < -1 || cfg->output > 4) return fail("unknown output layer kind %d", cfg->output);
  for (int k = 0; k < cfg->nblocks; k++) {
    if (cfg->nhidden[k] <= 0) return fail("bad dimensions (nhidden[%d] = %d)", k, cfg->nhidden[k]);
    if (cfg->direction[k] < 0 || cfg->direction[k] > 2) return fail("direction[%d] must be 0 (forward), 1 (reversed) or 2 (both)", k);
  }
  const int last_out = cfg->nhidden[cfg->nblocks - 1] * (cfg->direction[cfg->nblocks - 1] == 2 ? 2 : 1);
  const int nclasses = cfg->output < 0 ? last_out : cfg->noutput;
  if (cfg->ninput <= 0 || nclasses < 1) return fail("bad dimensions");
  if (cfg->output == 0 && nclasses < 2) return fail("bad dimensions (Softmax requires nclasses>=2, clstm.cc:400)");
  if (nclasses > kCtcMaxClasses) return fail("nclasses %d > %d not supported", nclasses, kCtcMaxClasses);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail("no CUDA device available (%s); clstm_b200 has no CPU fallback", cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev) return fail("device %d out of range (%d devices)", cfg->device, ndev);
  CU(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) return fail("device %d is sm_%d%d; this library contains sm_100a code only", cfg->device, prop.major, prop.minor);
  auto* n = new clstm_b200_net;
  n->cfg.ninput = cfg->ninput; n->cfg.nhidden = cfg->nhidden[0]; n->cfg.nclasses = nclasses; n->cfg.device = cfg->device;
  n->ni = cfg->ninput; n->no = cfg->nhidden[0]; n->nc = nclasses; n->nf = n->ni + n->no;
  n->nblk = cfg->nblocks; n->cell = cfg->cell; n->out_kind = cfg->output;
  n->num_sms = prop.multiProcessorCount;
  const int nc = n->nc;
  size_t o = 0;
  int nin = cfg->ninput;
  for (int k = 0; k < n->nblk; k++) {
    auto& bk = n->blk[k];
    bk.ni = nin; bk.no = cfg->nhidden[k];
    bk.ndir = cfg->direction[k] == 2 ? 2 : 1;
    bk.d0 = cfg->direction[k] == 1 ? 1 : 0;
    bk.hoff[0] = 0; bk.hoff[1] = (bk.ndir == 2) ? bk.no : 0;
    for (int d = bk.d0; d < bk.d0 + bk.ndir; d++) {
      bk.oWx[d] = o; o += (size_t)4 * bk.no * bk.ni;
      bk.oB[d] = o;  o += (size_t)4 * bk.no;
      bk.oR[d] = o;  o += (size_t)4 * bk.no * bk.no;
    }
    nin = bk.nout();
  }
  n->nfeat = nin;
  if (n->out_kind >= 0) {
    n->oW1 = o; o += (size_t)nc * n->nfeat;
    n->oB1 = o; o += nc;
  }
  n->P = o;
  if (cudaStreamCreateWithFlags(&n->st, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&n->st2, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&n->stc, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&n->ev_ready, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&n->ev_consumed, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&n->spare.ev_ready, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&n->spare.ev_consumed, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&n->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&n->ev_join, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&n->ev_fork2, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&n->ev_dx, cudaEventDisableTiming) != cudaSuccess) {
    clstm_b200_destroy(n);
    return fail("cudaStreamCreate failed");
  }
  int rc = 0;
  rc |= dev_alloc(&n->v, n->P); rc |= dev_alloc(&n->d, n->P);
  rc |= dev_alloc(&n->comm_buf, n->P + kPeerHeaderFloats);
  n->g = n->comm_buf ? n->comm_buf + kPeerHeaderFloats : nullptr;
  size_t wmax = 0;                       // largest per-direction LSTM parameter block (split-K workspace sizing)
  for (int k = 0; k < n->nblk; k++) {
    auto& bk = n->blk[k];
    wmax = std::max(wmax, (size_t)4 * bk.no * (1 + bk.ni + bk.no));
    for (int d = bk.d0; d < bk.d0 + bk.ndir; d++) {
      rc |= dev_alloc(&bk.Rt[d], (size_t)4 * bk.no * bk.no);
      rc |= dev_alloc(&bk.WxT[d], (size_t)4 * bk.no * bk.ni);
      if (!rc) {
        cudaMemsetAsync(bk.Rt[d], 0, (size_t)4 * bk.no * bk.no * sizeof(float), n->st);
        cudaMemsetAsync(bk.WxT[d], 0, (size_t)4 * bk.no * bk.ni * sizeof(float), n->st);
      }
    }
  }
  rc |= dev_alloc(&n->W1T, (size_t)std::max(1, n->nfeat) * nc);
  rc |= dev_alloc(&n->status, 1);
  // split-K workspace: enough for ~2 waves of 64x64 tiles plus the largest derivative matrix a few times over
  n->ws_floats = std::max<size_t>((size_t)4 * n->num_sms * 64 * 64, (size_t)40 * wmax);
  rc |= dev_alloc(&n->ws, n->ws_floats);
  n->ws2_floats = (size_t)64 * nc * (n->nfeat + 1);
  rc |= dev_alloc(&n->ws2, n->ws2_floats);
  if (rc) { clstm_b200_destroy(n); return 1; }
  cudaMemsetAsync(n->v, 0, n->P * sizeof(float), n->st);
  cudaMemsetAsync(n->d, 0, n->P * sizeof(float), n->st);
  cudaMemsetAsync(n->comm_buf, 0, (n->P + kPeerHeaderFloats) * sizeof(float), n->st);
  cudaMemsetAsync(n->W1T, 0, (size_t)std::max(1, n->nfeat) * nc * sizeof(float), n->st);
  cudaMemsetAsync(n->status, 0, sizeof(int), n->st);
  {
    const char* e = getenv("CLSTM_B200_GEMM");   // "simt" selects the fp32 SIMT tiles (A/B testing against tcgen05)
    n->use_tc = !(e && strcmp(e, "simt") == 0);
    n->gx_mode = (e && strcmp(e, "x") == 0) ? 1 : ((e && strcmp(e, "tc") == 0) ? 2 : 0);   // "x": always the TMA-fed GEMM, "tc": never
    const char* e2 = getenv("CLSTM_B200_GX_SMALLK");
    n->gx_small_k = !(e2 && atoi(e2) == 0) || n->gx_mode == 1;
  }
  {
    const char* e = getenv("CLSTM_B200_LSTM");   // "tc": always the batched tensor-core recurrence; "simt": never
    n->lstm_mode = (e && strcmp(e, "tc") == 0) ? 1 : ((e && strcmp(e, "simt") == 0) ? 2 : ((e && strcmp(e, "tcx") == 0) ? 3 : 0));
  }
  if (n->use_tc && n->gx_mode != 2) {
    n->gx[0] = gemm_x_create(n->num_sms); n->gx[1] = gemm_x_create(n->num_sms);
    if (!n->gx[0] || !n->gx[1]) { clstm_b200_destroy(n); return fail("the TMA-fed GEMM (gemm_x.cu) could not be set up on this device"); }
  }
  if (lstm_tc_configure() != 0) { clstm_b200_destroy(n); return fail("kernel image for sm_100a not usable on this device"); }
  if (n->cell == 0 && n->lstm_mode != 2)
    for (int k = 0; k < n->nblk; k++)
    {
      if (lstm_tc_supported(n->blk[k].no)) n->blk[k].tc = lstm_tc_create(n->blk[k].no, n->num_sms);
      if (lstm_tcx_supported(n->blk[k].no)) n->blk[k].tcx = lstm_tcx_create(n->blk[k].no, n->num_sms);
    }
  if (lstm_configure() != 0 || ctc_configure() != 0 || gemm_tc_configure() != 0 || norm_configure() != 0) { clstm_b200_destroy(n); return fail("kernel image for sm_100a not usable on this device"); }
  n->variant = n->cell == 0 ? lstm_variant_for(n->no) : "generic";
  if (cudaStreamSynchronize(n->st) != cudaSuccess) { clstm_b200_destroy(n); return fail("device initialisation failed"); }
  *out = n;
  return 0;
}

void clstm_b200_destroy(clstm_b200_net* n) {
  if (!n) return;
  cudaSetDevice(n->cfg.device);
  if (n->st) cudaStreamSynchronize(n->st);
  if (n->st2) cudaStreamSynchronize(n->st2);
  if (n->stc) cudaStreamSynchronize(n->stc);
  if (n->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(n->comm);
  free_batch(n);
  for (int k = 0; k < 2; k++) {            // both input sets
    dev_free(n->x); dev_free(n->meta); dev_free(n->lat_off); dev_free(n->tiles);
    if (n->h_tiles) cudaFreeHost(n->h_tiles);
    if (n->h_meta) cudaFreeHost(n->h_meta);
    if (n->h_lat) cudaFreeHost(n->h_lat);
    n->h_tiles = nullptr; n->h_meta = nullptr; n->h_lat = nullptr;
    if (n->meta_done) cudaEventDestroy(n->meta_done);
    if (n->ev_ready) cudaEventDestroy(n->ev_ready);
    if (n->ev_consumed) cudaEventDestroy(n->ev_consumed);
    n->meta_done = nullptr; n->ev_ready = nullptr; n->ev_consumed = nullptr;
    swap_sets(n);
  }
  if (n->stc) cudaStreamDestroy(n->stc);
  for (int r = 0; r < kMaxPeers; r++)
    if (n->peer_buf[r] && r != n->rank) cudaIpcCloseMemHandle(n->peer_buf[r]);
  dev_free(n->v); dev_free(n->d); dev_free(n->comm_buf); n->g = nullptr; dev_free(n->W1T);
  gemm_x_destroy(n->gx[0]); gemm_x_destroy(n->gx[1]); n->gx[0] = n->gx[1] = nullptr;
  for (int k = 0; k < 2; k++) {
    lstm_tc_destroy(n->blk[k].tc);
    n->blk[k].tc = nullptr;
    lstm_tcx_destroy(n->blk[k].tcx);
    n->blk[k].tcx = nullptr;
    for (int d = 0; d < 2; d++) { dev_free(n->blk[k].Rt[d]); dev_free(n->blk[k].WxT[d]); }
  }
  dev_free(n->lm); dev_free(n->lr); dev_free(n->rl); dev_free(n->status); dev_free(n->peer_stats);
  dev_free(n->ws); dev_free(n->tot); dev_free(n->mx_part); dev_free(n->ws2);
  dev_free(n->n_raw); dev_free(n->n_tmp); dev_free(n->n_smooth); dev_free(n->n_a); dev_free(n->n_center);
  dev_free(n->n_r); dev_free(n->n_scale); dev_free(n->n_masks); dev_free(n->n_ym); dev_free(n->n_yd); dev_free(n->n_meta);
  if (n->st2) cudaStreamDestroy(n->st2);
  if (n->ev_fork) cudaEventDestroy(n->ev_fork);
  if (n->ev_join) cudaEventDestroy(n->ev_join);
  if (n->ev_fork2) cudaEventDestroy(n->ev_fork2);
  if (n->ev_dx) cudaEventDestroy(n->ev_dx);
  for (int w = 0; w < 2; w++) { dev_free(n->dcls[w]); dev_free(n->dlocs[w]); dev_free(n->dcnt[w]); }
  if (n->h_small) cudaFreeHost(n->h_small);
  for (auto& r : n->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto e : n->pool) cudaEventDestroy(e);
  if (n->st) cudaStreamDestroy(n->st);
  delete n;
}

size_t clstm_b200_nparams(const clstm_b200_net* n) { return n ? n->P : 0; }

int clstm_b200_set_params(clstm_b200_net* n, const float* flat, size_t cnt) {
  if (!n || !flat) return fail("null argument");
  if (cnt != n->P) return fail("set_params size mismatch: got %zu, net has %zu", cnt, n->P);   // clstm.cc:871
  CU(cudaSetDevice(n->cfg.device));
  std::vector<float> dev(n->P);
  ref_to_dev(n, flat, dev.data());
  CU(cudaMemcpyAsync(n->v, dev.data(), n->P * sizeof(float), cudaMemcpyHostToDevice, n->st));
  prepare_weights(n);
  CU(cudaStreamSynchronize(n->st));
  return check_launch("set_params");
}
int clstm_b200_get_params(clstm_b200_net* n, float* flat, size_t cnt) {
  if (!n || !flat) return fail("null argument");
  if (cnt != n->P) return fail("get_params size mismatch: got %zu, net has %zu", cnt, n->P);
  CU(cudaSetDevice(n->cfg.device));
  std::vector<float> dev(n->P);
  CU(cudaMemcpyAsync(dev.data(), n->v, n->P * sizeof(float), cudaMemcpyDeviceToHost, n->st));
  CU(cudaStreamSynchronize(n->st));
  dev_to_ref(n, dev.data(), flat);
  return 0;
}
int clstm_b200_get_derivs(clstm_b200_net* n, float* flat, size_t cnt) {
  if (!n || !flat) return fail("null argument");
  if (cnt != n->P) return fail("get_derivs size mismatch");
  CU(cudaSetDevice(n->cfg.device));
  TRY(fold_pending(n));
  std::vector<float> dev(n->P);
  CU(cudaMemcpyAsync(dev.data(), n->d, n->P * sizeof(float), cudaMemcpyDeviceToHost, n->st));
  CU(cudaStreamSynchronize(n->st));
  dev_to_ref(n, dev.data(), flat);
  return 0;
}
int clstm_b200_set_derivs(clstm_b200_net* n, const float* flat, size_t cnt) {
  if (!n || !flat) return fail("null argument");
  if (cnt != n->P) return fail("set_derivs size mismatch");
  CU(cudaSetDevice(n->cfg.device));
  std::vector<float> dev(n->P);
  ref_to_dev(n, flat, dev.data());
  CU(cudaMemsetAsync(n->g, 0, n->P * sizeof(float), n->st));
  n->g_pending = false;
  CU(cudaMemcpyAsync(n->d, dev.data(), n->P * sizeof(float), cudaMemcpyHostToDevice, n->st));
  CU(cudaStreamSynchronize(n->st));
  return 0;
}
int clstm_b200_clear_derivs(clstm_b200_net* n) {
  if (!n) return fail("null argument");
  CU(cudaSetDevice(n->cfg.device));
  CU(cudaMemsetAsync(n->g, 0, n->P * sizeof(float), n->st));
  CU(cudaMemsetAsync(n->d, 0, n->P * sizeof(float), n->st));
  n->g_pending = false;
  return 0;
}

int clstm_b200_upload_batch(clstm_b200_net* n, const float* x, const int* T, int B, const int* labels, const int* L) {
  if (!n || !x || !T) return fail("null argument");
  if (L != nullptr && labels == nullptr) return fail("labels missing");
  CU(cudaSetDevice(n->cfg.device));
  Scope s(n, PH_H2D);
  TRY(stage_lines(n, T, B, labels, L));
  CU(cudaMemcpyAsync(n->x, x, (size_t)n->ln.N * n->ni * sizeof(float), cudaMemcpyHostToDevice, n->st));
  return 0;
}

int clstm_b200_forward(clstm_b200_net* n, const float* x, const int* T, int B, float* out) {
  TRY(clstm_b200_upload_batch(n, x, T, B, nullptr, nullptr));
  TRY(run_forward(n));
  if (out) {
    Scope s(n, PH_D2H);
    CU(cudaMemcpyAsync(out, n->out, (size_t)n->ln.N * n->nc * sizeof(float), cudaMemcpyDeviceToHost, n->st));
  }
  CU(cudaStreamSynchronize(n->st));
  return 0;
}

// measure + normalize B raw lines into the CURRENT input set on stream `cs` (the step stream, or the copy stream when
// called for the spare set by the input pipeline)
static int normalize_into(clstm_b200_net* n, const float* raw, const int* W, const int* H, int B, int kind,
                          const float* params, const int* labels, const int* L, int* T_out, cudaStream_t cs) {
  if (!n || !raw || !W || !H) return fail("null argument");
  if (B <= 0) return fail("batch must contain at least one line");
  if (kind < 0 || kind > 2) return fail("unknown normalizer name");      // extras.cc:299
  if (L != nullptr && labels == nullptr) return fail("labels missing");
  CU(cudaSetDevice(n->cfg.device));
  const int th = n->ni;                                                  // target_height = ninput (clstmhl.h:164)
  const float range = params ? params[0] : (kind == 1 ? 1.0f : 4.0f);    // extras.h:36, extras.cc:159,233
  const float smooth2d = params ? params[1] : 1.0f, smooth1d = params ? params[2] : 0.3f;
  const float vscale = params ? params[3] : 1.0f;
  std::vector<int> meta((size_t)10 * B);
  int* mW = meta.data(); int* mH = mW + B; int* mP = mH + B; int* mC = mP + B; int* mO = mC + B; int* mR = mO + 3 * B;
  size_t pix = 0, cols = 0;
  int maxpix = 0, maxh = 0, maxw = 0, maxrange = 0;
  for (int b = 0; b < B; b++) {
    if (W[b] <= 0 || H[b] <= 0) return fail("line %d has an empty image (%d x %d)", b, W[b], H[b]);
    if (H[b] > kNormMaxHeight) return fail("line %d is %d rows high (max %d)", b, H[b], kNormMaxHeight);
    if (kind == 0 && H[b] != th) return fail("line %d: height %d != target_height %d (NoNormalizer, extras.cc:149)", b, H[b], th);
    mW[b] = W[b]; mH[b] = H[b]; mP[b] = (int)pix; mC[b] = (int)cols;
    pix += (size_t)W[b] * H[b]; cols += W[b];
    if (pix > 0x7fffffff) return fail("raw batch too large");
    maxpix = std::max(maxpix, W[b] * H[b]); maxh = std::max(maxh, H[b]); maxw = std::max(maxw, W[b]);
  }
  std::vector<float> masks;
  if (kind == 2) {
    std::map<float, std::pair<int, int>> cache;   // sigma -> (offset, range)
    std::vector<float> m;
    for (int b = 0; b < B; b++) {
      const int h = H[b];
      const float sig[3] = {(float)(h * 0.5), h * smooth2d, h * smooth1d};   // gauss2d(h*smooth2d, h*0.5); gauss1d(h*smooth1d)
      for (int k = 0; k < 3; k++) {
        auto it = cache.find(sig[k]);
        if (it == cache.end()) {
          int r;
          host_gauss_mask(sig[k], m, r);
          it = cache.emplace(sig[k], std::make_pair((int)masks.size(), r)).first;
          masks.insert(masks.end(), m.begin(), m.end());
        }
        mO[3 * b + k] = it->second.first; mR[3 * b + k] = it->second.second;
        maxrange = std::max(maxrange, it->second.second);
      }
    }
  }
  TRY(ensure_norm(n, pix, cols, B, masks.size()));
  NormLines nl;
  nl.W = n->n_meta; nl.H = nl.W + B; nl.poff = nl.H + B; nl.coff = nl.poff + B; nl.moff = nl.coff + B; nl.mrange = nl.moff + 3 * B;
  nl.masks = n->n_masks; nl.range = range;
  std::vector<int> tw(B);
  std::vector<float> scale(B, 0.f);
  std::vector<double> ym(B, 0.0), yd(B, 0.0);
  {
    Scope s(n, PH_H2D, cs);
    CU(cudaMemcpyAsync(n->n_raw, raw, pix * sizeof(float), cudaMemcpyHostToDevice, cs));
    CU(cudaMemcpyAsync(n->n_meta, meta.data(), meta.size() * sizeof(int), cudaMemcpyHostToDevice, cs));
    if (!masks.empty()) CU(cudaMemcpyAsync(n->n_masks, masks.data(), masks.size() * sizeof(float), cudaMemcpyHostToDevice, cs));
  }
  n->n_hr.assign(B, 0.f);
  if (kind == 2) {
    {
      Scope s(n, PH_NORMALIZE, cs);
      if (maxrange > kNormMaxRange) return fail("normalizer smoothing width %d exceeds the supported %d", maxrange, kNormMaxRange);
      s.launches(norm_center_measure(cs, nl, B, maxw, maxh, maxrange, n->n_raw, n->n_tmp, n->n_smooth, n->n_a, n->n_center, n->n_r));
      TRY(check_launch("normalizer measure"));
    }
    CU(cudaMemcpyAsync(n->n_hr.data(), n->n_r, B * sizeof(float), cudaMemcpyDeviceToHost, cs));
    CU(cudaStreamSynchronize(cs));
    for (int b = 0; b < B; b++) {                         // CenterNormalizer::normalize, extras.cc:275-276
      const float r = n->n_hr[b];
      scale[b] = (2.0 * r) / th;
      // a blank (all-zero) line makes r NaN in the reference and int(W / NaN) is undefined behaviour there; fail loudly
      if (!(scale[b] > 0.f) || !std::isfinite(scale[b]))
        return fail("line %d: the centre normalizer measured no ink (r = %g); blank line images cannot be normalised", b, (double)r);
      tw[b] = std::max(int(W[b] / scale[b]), 1);
    }
  } else if (kind == 1) {
    {
      Scope s(n, PH_NORMALIZE, cs);
      s.launches(norm_mean_measure(cs, nl, B, maxh, n->n_raw, n->n_ym, n->n_yd));
      TRY(check_launch("normalizer measure"));
    }
    CU(cudaMemcpyAsync(ym.data(), n->n_ym, B * sizeof(double), cudaMemcpyDeviceToHost, cs));
    CU(cudaMemcpyAsync(yd.data(), n->n_yd, B * sizeof(double), cudaMemcpyDeviceToHost, cs));
    CU(cudaStreamSynchronize(cs));
    for (int b = 0; b < B; b++) {                         // MeanNormalizer::normalize, extras.cc:185-188
      float actual = vscale * 2 * range * yd[b];
      scale[b] = actual / th;
      if (!(scale[b] > 0.f) || !std::isfinite(scale[b]))
        return fail("line %d: the mean normalizer measured no ink (y_mad = %g); blank line images cannot be normalised", b, yd[b]);
      tw[b] = int(W[b] / scale[b]);
      if (tw[b] <= 0) return fail("line %d: normalised width %d", b, tw[b]);
    }
  } else {
    for (int b = 0; b < B; b++) tw[b] = W[b];
  }
  for (int b = 0; b < B; b++)
    if ((long long)tw[b] > 64LL * 1024 * 1024) return fail("line %d: normalised width %d is not plausible", b, tw[b]);
  TRY(stage_lines(n, tw.data(), B, labels, L, false, cs));
  {
    Scope s(n, PH_NORMALIZE, cs);
    CU(cudaMemcpyAsync(n->n_scale, scale.data(), B * sizeof(float), cudaMemcpyHostToDevice, cs));
    s.launches(norm_resample(cs, nl, B, n->ln.Tmax, n->n_raw, n->n_center, n->n_scale, n->n_ym, n->ln.T, n->ln.off, n->x, th, kind));
    TRY(check_launch("normalizer resample"));
  }
  n->n_B = B; n->n_cols = (int)cols;
  if (T_out) memcpy(T_out, tw.data(), B * sizeof(int));
  CU(cudaStreamSynchronize(cs));     // `scale` and `meta` are stack/heap temporaries of this call
  return 0;
}


int clstm_b200_normalize_batch(clstm_b200_net* n, const float* raw, const int* W, const int* H, int B, int kind,
                               const float* params, const int* labels, const int* L, int* T_out) {
  if (!n) return fail("null argument");
  return normalize_into(n, raw, W, H, B, kind, params, labels, L, T_out, n->st);
}

// input pipeline for RAW lines: normalise batch i+1 into the spare input set on the copy stream while step i runs (the
// recurrent kernels leave more than half of the SMs idle at 32 lines per GPU, so the normaliser is hidden behind them)
int clstm_b200_prefetch_raw_batch(clstm_b200_net* n, const float* raw, const int* W, const int* H, int B, int kind,
                                  const float* params, const int* labels, const int* L, int* T_out) {
  if (!n || !labels || !L) return fail("null argument");
  CU(cudaSetDevice(n->cfg.device));
  const bool keep_fwd = n->have_forward, keep_ctc = n->have_ctc;
  swap_sets(n);
  int rc = 0;
  do {
    if (n->consumed_recorded && cudaStreamWaitEvent(n->stc, n->ev_consumed, 0) != cudaSuccess) { rc = fail("cudaStreamWaitEvent failed"); break; }
    if ((rc = normalize_into(n, raw, W, H, B, kind, params, labels, L, T_out, n->stc)) != 0) break;
    if (cudaEventRecord(n->ev_ready, n->stc) != cudaSuccess) { rc = fail("cudaEventRecord failed"); break; }
  } while (0);
  swap_sets(n);
  n->have_forward = keep_fwd; n->have_ctc = keep_ctc;
  n->prefetched = (rc == 0);
  return rc;
}

int clstm_b200_normalizer_state(clstm_b200_net* n, float* center, float* r) {
  if (!n) return fail("null argument");
  if (n->n_B <= 0) return fail("normalizer_state called before normalize_batch");
  CU(cudaSetDevice(n->cfg.device));
  if (center) CU(cudaMemcpyAsync(center, n->n_center, (size_t)n->n_cols * sizeof(float), cudaMemcpyDeviceToHost, n->st));
  CU(cudaStreamSynchronize(n->st));
  if (r) memcpy(r, n->n_hr.data(), (size_t)n->n_B * sizeof(float));
  return 0;
}

int clstm_b200_get_inputs(clstm_b200_net* n, float* x) {
  if (!n || !x) return fail("null argument");
  if (!n->have_batch) return fail("get_inputs called without a resident batch");
  CU(cudaSetDevice(n->cfg.device));
  CU(cudaMemcpyAsync(x, n->x, (size_t)n->ln.N * n->ni * sizeof(float), cudaMemcpyDeviceToHost, n->st));
  CU(cudaStreamSynchronize(n->st));
  return 0;
}

int clstm_b200_forward_resident(clstm_b200_net* n, float* out) {
  if (!n) return fail("null argument");
  if (!n->have_batch) return fail("forward_resident called without a resident batch");
  CU(cudaSetDevice(n->cfg.device));
  TRY(run_forward(n));
  if (out) {
    Scope s(n, PH_D2H);
    CU(cudaMemcpyAsync(out, n->out, (size_t)n->ln.N * n->nc * sizeof(float), cudaMemcpyDeviceToHost, n->st));
  }
  CU(cudaStreamSynchronize(n->st));
  return 0;
}

int clstm_b200_ctc_align(clstm_b200_net* n, const int* labels, const int* L, float* aligned) {
  if (!n || !L) return fail("null argument");
  if (!n->have_forward) return fail("ctc_align called before forward");
  CU(cudaSetDevice(n->cfg.device));
  {
    // re-stage the metadata with the transcripts; line geometry is unchanged
    std::vector<int> T = n->hT;
    if (!labels) {
      for (size_t b = 0; b < T.size(); b++)
        if (L[b] != 0) return fail("labels missing");
      labels = L;   // never dereferenced: every transcript is empty
    }
    TRY(stage_lines(n, T.data(), (int)T.size(), labels, L));
    n->have_forward = true;
  }
  CU(cudaMemsetAsync(n->status, 0, sizeof(int), n->st));
  TRY(run_ctc(n));
  CU(cudaMemcpyAsync(n->h_small, n->status, sizeof(int), cudaMemcpyDeviceToHost, n->st));
  if (aligned) {
    Scope s(n, PH_D2H);
    CU(cudaMemcpyAsync(aligned, n->aligned, (size_t)n->ln.N * n->nc * sizeof(float), cudaMemcpyDeviceToHost, n->st));
  }
  CU(cudaStreamSynchronize(n->st));
  return check_status(n);
}

int clstm_b200_ctc_align_states(clstm_b200_net* n, const float* outputs, const int* T, int B, const int* states,
                                const int* S, float* aligned) {
  if (!n || !outputs || !T || !states || !S || !aligned) return fail("null argument");
  CU(cudaSetDevice(n->cfg.device));
  // a free-standing alignment on the geometry of the last forward keeps that forward usable for backward()
  const bool same_geometry = n->have_forward && (int)n->hT.size() == B && std::equal(T, T + B, n->hT.begin());
  TRY(stage_lines(n, T, B, states, S, /*raw=*/true));
  n->have_forward = same_geometry;
  CU(cudaMemcpyAsync(n->out, outputs, (size_t)n->ln.N * n->nc * sizeof(float), cudaMemcpyHostToDevice, n->st));
  CU(cudaMemsetAsync(n->status, 0, sizeof(int), n->st));
  TRY(run_ctc(n));
  n->have_ctc = false;   // deltas of a free-standing alignment must not feed backward()
  CU(cudaMemcpyAsync(n->h_small, n->status, sizeof(int), cudaMemcpyDeviceToHost, n->st));
  CU(cudaMemcpyAsync(aligned, n->aligned, (size_t)n->ln.N * n->nc * sizeof(float), cudaMemcpyDeviceToHost, n->st));
  CU(cudaStreamSynchronize(n->st));
  return check_status(n);
}

int clstm_b200_backward(clstm_b200_net* n, const float* deltas, float* din) {
  if (!n) return fail("null argument");
  if (!n->have_forward) return fail("backward called before forward");
  if (!deltas && !n->have_ctc) return fail("backward: no deltas given and no ctc_align result on the device");
  CU(cudaSetDevice(n->cfg.device));
  if (deltas) {
    Scope s(n, PH_H2D);
    CU(cudaMemcpyAsync(n->delta, deltas, (size_t)n->ln.N * n->nc * sizeof(float), cudaMemcpyHostToDevice, n->st));
  }
  TRY(run_backward(n));
  if (din) {
    Scope s(n, PH_D2H);
    CU(cudaMemcpyAsync(din, n->dx, (size_t)n->ln.N * n->ni * sizeof(float), cudaMemcpyDeviceToHost, n->st));
  }
  CU(cudaStreamSynchronize(n->st));
  return 0;
}

int clstm_b200_decode(clstm_b200_net* n, int which, int* classes, int* locs, int* counts, int max_per_line) {
  if (!n || !classes || !locs || !counts) return fail("null argument");
  if (which < 0 || which > 1 || max_per_line <= 0) return fail("bad argument");
  if (!n->have_forward || (which == 1 && !n->have_ctc)) return fail("decode: nothing to decode yet");
  CU(cudaSetDevice(n->cfg.device));
  TRY(run_decode(n, which, max_per_line));
  return fetch_decode(n, which, classes, locs, counts, max_per_line);
}

int clstm_b200_argmax(clstm_b200_net* n, int which, int* idx) {
  if (!n || !idx) return fail("null argument");
  if (which < 0 || which > 1) return fail("bad argument");
  if (!n->have_forward || (which == 1 && !n->have_ctc)) return fail("argmax: nothing to decode yet");
  CU(cudaSetDevice(n->cfg.device));
  TRY(run_decode(n, which, std::max(n->capDec, 1)));
  CU(cudaMemcpyAsync(idx, n->amax[which], (size_t)n->ln.N * sizeof(int), cudaMemcpyDeviceToHost, n->st));
  CU(cudaStreamSynchronize(n->st));
  return 0;
}

int clstm_b200_sgd_update(clstm_b200_net* n, float lr, float momentum, float clip) {
  if (!n) return fail("null argument");
  CU(cudaSetDevice(n->cfg.device));
  TRY(run_update(n, lr, momentum, clip));
  CU(cudaStreamSynchronize(n->st));
  return 0;
}

int clstm_b200_comm_unique_id(void* id128) {
  if (!id128) return fail("null argument");
  TRY(load_nccl());
  int r = g_nccl.GetUniqueId(id128);
  if (r != 0) return fail("ncclGetUniqueId failed (%d)", r);
  return 0;
}
int clstm_b200_comm_init(clstm_b200_net* n, const void* id128, int rank, int world) {
  if (!n || !id128) return fail("null argument");
  if (world < 1 || rank < 0 || rank >= world) return fail("bad rank/world");
  TRY(load_nccl());
  CU(cudaSetDevice(n->cfg.device));
  Id128 id;
  memcpy(&id, id128, sizeof id);
  int r = g_nccl.CommInitRank(&n->comm, world, id, rank);
  if (r != 0) return fail("ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  n->rank = rank;
  n->world = world;
  return 0;
}
// Peer-memory path: export this rank's comm buffer, import everybody's.  handle64: 64-byte cudaIpcMemHandle_t.
int clstm_b200_p2p_handle(clstm_b200_net* n, void* handle64) {
  if (!n || !handle64) return fail("null argument");
  CU(cudaSetDevice(n->cfg.device));
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, n->comm_buf));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, sizeof h);
  return 0;
}
int clstm_b200_p2p_connect(clstm_b200_net* n, const void* handles, int rank, int world) {
  if (!n || !handles) return fail("null argument");
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) return fail("bad rank/world (at most %d peers)", kMaxPeers);
  if (n->P % 4 != 0 && false) return fail("unreachable");
  CU(cudaSetDevice(n->cfg.device));
  CU(cudaStreamSynchronize(n->st));
  for (int r = 0; r < world; r++) {
    if (r == rank) { n->peer_buf[r] = n->comm_buf; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)r * 64, sizeof h);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    n->peer_buf[r] = (float*)p;
  }
  n->rank = rank; n->world = world; n->p2p = true; n->epoch = 0;
  return 0;
}

int clstm_b200_allreduce_derivs(clstm_b200_net* n) {
  if (!n) return fail("null argument");
  if (!n->comm) return fail("no communicator attached (clstm_b200_comm_init)");
  CU(cudaSetDevice(n->cfg.device));
  TRY(run_allreduce(n));
  CU(cudaStreamSynchronize(n->st));
  return 0;
}

int clstm_b200_step_resident(clstm_b200_net* n, float lr, float momentum, float clip) {
  if (!n) return fail("null argument");
  if (!n->have_batch || !n->have_labels) return fail("step_resident: upload a batch with transcripts first");
  CU(cudaSetDevice(n->cfg.device));
  TRY(run_forward(n));
  TRY(run_ctc(n));
  TRY(run_backward(n, /*defer_dx=*/true));
  if (n->p2p && n->world > 1) {
    TRY(run_peer_update(n, lr, momentum, clip));
  } else {
    TRY(run_allreduce(n));
    TRY(run_update(n, lr, momentum, clip));
  }
  TRY(run_decode(n, 0, std::max(n->capDec, n->ln.Tmax / 2 + 1)));
  join_dx(n);
  return 0;
}

// ---- input pipeline: stage batch i+1 on the copy stream into the spare input set while step i runs
int clstm_b200_prefetch_batch(clstm_b200_net* n, const float* x, const int* T, int B, const int* labels, const int* L) {
  if (!n || !x || !T || !labels || !L) return fail("null argument");
  CU(cudaSetDevice(n->cfg.device));
  const bool keep_fwd = n->have_forward, keep_ctc = n->have_ctc;   // flags of the CURRENT batch (stage_lines resets them)
  swap_sets(n);                            // work on the spare set under the usual member names
  int rc = 0;
  do {
    if (n->consumed_recorded) {            // the step that last ran on this set must be done before it is overwritten
      if (cudaStreamWaitEvent(n->stc, n->ev_consumed, 0) != cudaSuccess) { rc = fail("cudaStreamWaitEvent failed"); break; }
    }
    if ((rc = stage_lines(n, T, B, labels, L, false, n->stc)) != 0) break;
    if (cudaMemcpyAsync(n->x, x, (size_t)n->ln.N * n->ni * sizeof(float), cudaMemcpyHostToDevice, n->stc) != cudaSuccess ||
        cudaEventRecord(n->ev_ready, n->stc) != cudaSuccess) { rc = fail("prefetch copy failed"); break; }
  } while (0);
  swap_sets(n);
  n->have_forward = keep_fwd; n->have_ctc = keep_ctc;
  n->prefetched = (rc == 0);
  return rc;
}

int clstm_b200_step_prefetched(clstm_b200_net* n, float lr, float momentum, float clip) {
  if (!n) return fail("null argument");
  if (!n->prefetched) return fail("step_prefetched: no prefetched batch");
  CU(cudaSetDevice(n->cfg.device));
  CU(cudaEventRecord(n->ev_consumed, n->st));   // everything enqueued so far on the current set
  n->consumed_recorded = true;
  swap_sets(n);
  n->prefetched = false;
  n->have_forward = false; n->have_ctc = false;
  CU(cudaStreamWaitEvent(n->st, n->ev_ready, 0));
  return clstm_b200_step_resident(n, lr, momentum, clip);
}

int clstm_b200_fetch_decoded(clstm_b200_net* n, int which, int* classes, int* locs, int* counts, int max_per_line) {
  if (!n || !classes || !locs || !counts) return fail("null argument");
  if (which != 0) TRY(run_decode(n, which, std::max(n->capDec, max_per_line)));
  if (max_per_line > n->capDec) return fail("fetch_decoded: max_per_line %d exceeds the decoded capacity %d", max_per_line, n->capDec);
  return fetch_decode(n, which, classes, locs, counts, max_per_line);
}

int clstm_b200_synchronize(clstm_b200_net* n) {
  if (!n) return fail("null argument");
  CU(cudaSetDevice(n->cfg.device));
  CU(cudaStreamSynchronize(n->st));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("device error: %s", cudaGetErrorString(e));
  return 0;
}

int clstm_b200_train_step(clstm_b200_net* n, const float* x, const int* T, int B, const int* labels, const int* L,
                          float lr, float momentum, float clip, float* out, float* aligned, int* classes, int* locs,
                          int* counts, int max_per_line) {
  if (!labels || !L) return fail("train_step needs transcripts");
  TRY(clstm_b200_upload_batch(n, x, T, B, labels, L));
  CU(cudaMemsetAsync(n->status, 0, sizeof(int), n->st));
  if (classes && max_per_line > 0) TRY(ensure_decode(n, max_per_line));
  TRY(clstm_b200_step_resident(n, lr, momentum, clip));
  CU(cudaMemcpyAsync(n->h_small, n->status, sizeof(int), cudaMemcpyDeviceToHost, n->st));
  {
    Scope s(n, PH_D2H);
    if (out) CU(cudaMemcpyAsync(out, n->out, (size_t)n->ln.N * n->nc * sizeof(float), cudaMemcpyDeviceToHost, n->st));
    if (aligned) CU(cudaMemcpyAsync(aligned, n->aligned, (size_t)n->ln.N * n->nc * sizeof(float), cudaMemcpyDeviceToHost, n->st));
  }
  if (classes && locs && counts && max_per_line > 0) {
    TRY(fetch_decode(n, 0, classes, locs, counts, max_per_line));   // synchronises
  } else {
    CU(cudaStreamSynchronize(n->st));
  }
  return check_status(n);
}

int clstm_b200_profile(clstm_b200_net* n, int enable) {
  if (!n) return fail("null argument");
  CU(cudaStreamSynchronize(n->st));
  harvest(n);
  n->prof = enable != 0;
  for (int i = 0; i < PH_COUNT; i++) { n->ph_ms[i] = 0.f; n->ph_launch[i] = 0; }
  return 0;
}
int clstm_b200_phase_stats(clstm_b200_net* n, float* ms, long long* launches, int cnt) {
  if (!n) return fail("null argument");
  CU(cudaStreamSynchronize(n->st));
  harvest(n);
  for (int i = 0; i < cnt && i < PH_COUNT; i++) {
    if (ms) ms[i] = n->ph_ms[i];
    if (launches) launches[i] = n->ph_launch[i];
  }
  return 0;
}
void* clstm_b200_stream(clstm_b200_net* n) { return n ? (void*)n->st : nullptr; }
const char* clstm_b200_lstm_variant(const clstm_b200_net* n) { return n ? n->variant : ""; }

// Self-test of the tcgen05 (3xTF32) dense products against the fp32 SIMT tiles on random data, same shapes/strides
// as the products of the path.  err[i] = max |tc - simt| / max|simt| for case i.  Returns the number of cases.
int clstm_b200_peer_stats(clstm_b200_net* n, double* out4, int reset) {
  if (!n || !out4) return fail("null argument");
  CU(cudaSetDevice(n->cfg.device));
  if (!n->peer_stats) {
    TRY(dev_alloc(&n->peer_stats, 3));
    CU(cudaMemsetAsync(n->peer_stats, 0, 3 * sizeof(unsigned long long), n->st));
  }
  unsigned long long h[3] = {0, 0, 0};
  CU(cudaMemcpyAsync(h, n->peer_stats, sizeof h, cudaMemcpyDeviceToHost, n->st));
  CU(cudaStreamSynchronize(n->st));
  const double k = h[2] ? 1.0 / (double)h[2] : 0.0;
  out4[0] = (double)h[2];
  out4[1] = 1e-3 * (double)h[0] * k;                                   // us per launch waiting for the slowest rank
  out4[2] = 1e-3 * (double)h[1] * k;                                   // us per launch reading the peers + updating
  out4[3] = (double)(n->world > 1 ? n->world - 1 : 0) * 4.0 * (double)n->P;   // bytes read over NVLink per launch
  if (reset) CU(cudaMemsetAsync(n->peer_stats, 0, 3 * sizeof(unsigned long long), n->st));
  return 0;
}

int clstm_b200_selftest_lstm_x(int device, int nhidden, int nlines, int tmin, int tmax, unsigned seed, float wscale, float* out9) {
  if (!out9) return fail("null argument");
  if (nhidden <= 0 || nlines <= 0 || tmin <= 0 || tmax < tmin) return fail("bad self-test geometry");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fail("no CUDA device %d; clstm_b200 has no CPU fallback", device);
  CU(cudaSetDevice(device));
  char msg[256];
  const int rc = lstm_tc_selftest(nhidden, nlines, tmin, tmax, seed, wscale, out9, msg, (int)sizeof msg, 1);
  if (rc != 0) return fail("selftest_lstm_x (nhidden %d, %d lines): %s", nhidden, nlines, msg);
  return 0;
}

int clstm_b200_selftest_lstm(int device, int nhidden, int nlines, int tmin, int tmax, unsigned seed, float wscale, float* out9) {
  if (!out9) return fail("null argument");
  if (nhidden <= 0 || nlines <= 0 || tmin <= 0 || tmax < tmin) return fail("bad self-test geometry");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fail("no CUDA device %d; clstm_b200 has no CPU fallback", device);
  CU(cudaSetDevice(device));
  char msg[256];
  const int rc = lstm_tc_selftest(nhidden, nlines, tmin, tmax, seed, wscale, out9, msg, (int)sizeof msg);
  if (rc != 0) return fail("selftest_lstm (nhidden %d, %d lines): %s", nhidden, nlines, msg);
  return 0;
}

int clstm_b200_selftest_gemm(clstm_b200_net* n, float* err, int max_cases) {
  if (!n || !err) return -1;
  if (cudaSetDevice(n->cfg.device) != cudaSuccess) return -1;
  const int ni = n->ni, no = n->no, nc = n->nc;
  const int N = 1000 + 37;   // columns, deliberately not a multiple of the tile sizes
  struct Case { int kind, M, Nn, K; };   // kind 0: nt (B K-contig), 1: nt (B MN-contig), 2: tn with 2 blocks + ones
  const Case cases[] = {{0, N, 4 * no, ni}, {0, N, nc, 2 * no}, {1, N, 2 * no, nc}, {1, N, ni, 4 * no},
                        {2, 4 * no, ni + no + 1, N}, {2, nc, 2 * no + 1, N}};
  const int ncases = (int)(sizeof(cases) / sizeof(cases[0]));
  const bool saved = n->use_tc;
  int done = 0;
  for (int ci = 0; ci < ncases && ci < max_cases; ci++) {
    const Case& c = cases[ci];
    size_t na, nb, ncout;
    if (c.kind == 2) { na = (size_t)c.K * c.M; nb = (size_t)c.K * (c.Nn - 1); ncout = (size_t)c.M * c.Nn; }
    else { na = (size_t)c.M * c.K; nb = (size_t)c.Nn * c.K; ncout = (size_t)c.M * c.Nn; }
    std::vector<float> ha(na), hb(nb), hbias(c.Nn);
    unsigned lcg = 12345u + ci;
    auto rnd = [&]() { lcg = lcg * 1664525u + 1013904223u; return ((lcg >> 8) & 0xFFFF) / 32768.0f - 1.0f; };
    for (auto& v : ha) v = rnd();
    for (auto& v : hb) v = rnd();
    for (auto& v : hbias) v = rnd();
    float *dA = nullptr, *dB = nullptr, *dbias = nullptr, *dC[2] = {nullptr, nullptr};
    if (dev_alloc(&dA, na) || dev_alloc(&dB, nb) || dev_alloc(&dbias, (size_t)c.Nn) || dev_alloc(&dC[0], ncout) ||
        dev_alloc(&dC[1], ncout))
      return -1;
    cudaMemcpyAsync(dA, ha.data(), na * 4, cudaMemcpyHostToDevice, n->st);
    cudaMemcpyAsync(dB, hb.data(), nb * 4, cudaMemcpyHostToDevice, n->st);
    cudaMemcpyAsync(dbias, hbias.data(), (size_t)c.Nn * 4, cudaMemcpyHostToDevice, n->st);
    std::vector<float> res[2];
    for (int pass = 0; pass < 2; pass++) {
      n->use_tc = (pass == 1);
      cudaMemsetAsync(dC[pass], 0, ncout * 4, n->st);
      if (c.kind == 0) dense_nt(n, c.M, c.Nn, c.K, dA, c.K, dB, c.K, false, dC[pass], c.Nn, dbias, 0.f);
      else if (c.kind == 1) dense_nt(n, c.M, c.Nn, c.K, dA, c.K, dB, c.Nn, true, dC[pass], c.Nn, nullptr, 0.f);
      else {
        const int n0 = (c.Nn - 1) / 3, n1 = c.Nn - 1 - n0;      // two column blocks + bias column
        // B0 = hb[0 .. K*n0), B1 = following K*n1 ; outputs packed [M*n0 | M*n1 | M]
        dense_tn(n, c.M, c.K, dA, c.M, dB, n0, dB + (size_t)c.K * n0, n1, dC[pass], dC[pass] + (size_t)c.M * n0,
                 dC[pass] + (size_t)c.M * (n0 + n1));
      }
      res[pass].resize(ncout);
      cudaMemcpyAsync(res[pass].data(), dC[pass], ncout * 4, cudaMemcpyDeviceToHost, n->st);
      if (cudaStreamSynchronize(n->st) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
        n->use_tc = saved;
        fail("selftest_gemm case %d pass %d: device error", ci, pass);
        return -1;
      }
    }
    float mx = 0.f, md = 0.f;
    for (size_t i = 0; i < ncout; i++) { mx = std::max(mx, std::fabs(res[0][i])); md = std::max(md, std::fabs(res[0][i] - res[1][i])); }
    err[ci] = md / std::max(mx, 1e-20f);
    cudaFree(dA); cudaFree(dB); cudaFree(dbias); cudaFree(dC[0]); cudaFree(dC[1]);
    done++;
  }
  n->use_tc = saved;
  return done;
}

void* clstm_b200_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { fail("cudaHostAlloc(%zu) failed", bytes); return nullptr; }
  return p;
}
void clstm_b200_free_pinned(void* p) {
  if (p) cudaFreeHost(p);
}

}  // extern "C"
