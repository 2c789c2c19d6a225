// lstm_tc.cu -- the recurrence of the NPLSTM as a BATCHED tensor-core problem (tcgen05 + TMEM + TMA), for batches with
// many lines per GPU and for hidden sizes whose recurrent matrix no longer fits one SM's registers.
//
// Reference semantics (paths relative to /root/reference), identical to lstm.cu:
//   GenericNPLSTM<SIG,TANH,TANH>::forward   clstm.cc:600-621   loop body  :612-620
//   GenericNPLSTM::backward                 clstm.cc:622-653   loop body  :629-650
//   the two contractions that stay inside the time loop: forward_lin1 (clstm_compute.cc:286, the R h_{t-1} half) and
//   backward_lin1 (clstm_compute.cc:296, the R^T delta half); Reversed / Parallel wiring clstm.cc:458-479, 506-544.
//
// Formulation.  Lines are sorted by decreasing length (Lines::order) and cut into TILES of 128 line slots; all lines of a
// tile advance in lock step, so one step of one direction is a real GEMM
//        pre[128 lines x 4no gate rows] = h_{s-1}[128 x no] * R^T[no x 4no]            (UMMA M = 128 lines)
// The 4no gate rows (gate-interleaved: row 4j+g) are cut into NT slices of NR rows; CTA (slice m, tile, direction) keeps
// its slice of R resident in shared memory for the whole sequence, as fp16 hi/lo parts (tc_common.cuh::split_f16):
// three kind::f16 MMAs (lo*hi, hi*lo, hi*hi) per K = 16 give fp32-grade products at half the bytes of 3xTF32.
// Accumulators live in TMEM, lane = line, column = gate row, so the four gates of a hidden unit arrive in four adjacent
// registers of ONE thread and the cell update needs no shuffles; the cell state stays in registers across the sequence.
//
// Forward step s of a (tile, direction), per CTA (warp-specialised):
//   producer warp : waits until every slice has published h_{s-1} (global step counter), then streams the h tile
//                   [active lines x no] (fp16 hi/lo, K-major) from L2 into a shared-memory ring with TMA (SWIZZLE_128B)
//   MMA warp      : one thread issues the 3 x ceil(no/16) tcgen05.mma of the step; tcgen05.commit frees ring stages
//                   and signals the accumulator
//   8 epilogue warps (thread = line, two warps per TMEM lane quadrant split the columns): tcgen05.ld, + input
//                   projection, sigma / tanh, cell update, publish h_s as fp16 hi/lo (the next step's TMA source),
//                   bump the step counter, then write the stash (gates, cell, h, h_prev) for the backward pass.
// Backward step: each CTA owns the same NR = 64 gate rows (16 units).  It sums the partial products that every slice
// wrote for ITS units in the previous step (fixed order => deterministic), does the pointwise delta math, stores the
// deltas (DG) and puts them -- fp16 hi/lo -- into shared memory as the A operand [128 lines x 64]; the MMA warp multiplies
// with the resident (or TMA-streamed) R slice [no x 64] in chunks of 256 outputs, double buffered in TMEM, and the
// epilogue warps drain the chunks into the partial-sum exchange buffer (L2).
// The recurrence never leaves the GPU; the per-step exchange between the CTAs of a (tile, direction) goes through L2 and
// one release/acquire counter.  All CTAs of a launch must be co-resident: the launch is cooperative.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "kernels.h"
#include "tc_common.cuh"

namespace cb200 {
namespace {
using namespace tc;

constexpr int kTcLines = 128;          // line slots per tile = UMMA M = TMEM lanes
constexpr int kTcThreads = 320;        // warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue
constexpr int kEpiThreads = 256;
constexpr int kMaxStages = 6;
constexpr float kScaleH = 16.f;        // h in (-1,1)  -> fp16 hi/lo of 16 h
constexpr float kScaleR = 16.f;        // weights      -> fp16 hi/lo of 16 R   (|R| < 4096)
constexpr float kScaleD = 256.f;       // deltas       -> fp16 hi/lo of 256 d  (|d| < 256)
constexpr unsigned kStageBytes = 2 * kTcLines * 128;   // one ring stage: hi tile + lo tile of [128 lines x 64 k] fp16
constexpr int kBwdNR = 64;             // gate rows per CTA in the backward kernel (one 128-byte swizzle row of K)
constexpr int kDbgCtas = 160;          // debug counter slots (one per CTA)
constexpr int kDbgStep = 64;           // the step whose event timestamps the debug counters record
constexpr int kBwdChunk = 256;         // outputs per MMA chunk in the backward kernel (UMMA N)
constexpr unsigned kBwdStageBytes = 2 * kBwdChunk * 128;
constexpr unsigned kBwdScratch = 8 * 32 * 16 * 4;   // row-transfer scratch of the 8 epilogue warps (rows go in 64-byte pieces)

struct TcFwd {
  int no, no4;            // hidden units, gate rows
  int KC, nks;            // 64-wide k chunks / 16-wide k slices of h
  int NT, ntiles, nst;    // row slices per direction, line tiles in the batch, ring stages
  int nsig;               // counter increments per step of a (tile, direction): 8 epilogue warps of every slice
  int d0, hstride, hoff[2];
  int rows_pad;           // rows per direction of the split weight copy (NT * NR)
  int KP;                 // row pitch (halves) of the h exchange buffer
  const float* XP[2];
  float* G[2];
  float* C[2];
  float* Hprev[2];
  float* H;
  __half* hx_hi;          // [direction slot][parity][tile][128][KP]
  __half* hx_lo;
  unsigned* flags;        // [direction slot][tile] steps published
  long long* dbg;         // optional per-phase cycle counters of CTA (0,0,0) (self-test / tuning), nullptr otherwise
  int opt;                // tuning switches (CLSTM_B200_TC_OPT): 2 writer-side proxy fence,
                          // 4 no stash stores (timing experiment, wrong results), 8 no input-projection loads (ditto)
};

struct TcBwd {
  int no, no4;
  int NT, ntiles, nst, resident;   // resident: every B chunk has its own stage and is loaded once
  int nchunk, nop16;      // output chunks of 256, outputs padded to 16
  int d0, hstride, hoff[2];
  int kp_rows;            // rows per direction of the transposed split copy
  const float* G[2];
  const float* C[2];
  const float* dH;
  float* DG[2];
  float* part;            // [direction slot][tile][parity][dest slice][src slice][128][16]
  unsigned* flags;
  long long* dbg;
  int opt;
};

// vectorised per-thread stores of N consecutive floats (N % VW == 0, address VW-float aligned)
template <int N, int VW>
__device__ __forceinline__ void store_run(float* dst, const float* v) {
#pragma unroll
  for (int i = 0; i < N; i += VW) {
    if (VW == 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    else if (VW == 2) *reinterpret_cast<float2*>(dst + i) = make_float2(v[i], v[i + 1]);
    else dst[i] = v[i];
  }
}
template <int N, int VW>
__device__ __forceinline__ void load_run(float* v, const float* src) {
#pragma unroll
  for (int i = 0; i < N; i += VW) {
    if (VW == 4) {
      const float4 t = *reinterpret_cast<const float4*>(src + i);
      v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    } else if (VW == 2) {
      const float2 t = *reinterpret_cast<const float2*>(src + i);
      v[i] = t.x; v[i + 1] = t.y;
    } else v[i] = src[i];
  }
}

// ---------------------------------------------------------------------------------------------- warp-cooperative rows
// Every epilogue thread owns one text line, so its loads / stores of the line's current column go to a row of NF floats at
// an address of its own: a lane-private vector access touches 32 different 128-byte lines per instruction (32 LSU
// wavefronts for 512 bytes).  These helpers move the 32 rows of a warp through a per-warp shared-memory scratch instead,
// so that consecutive lanes cover consecutive 16-byte chunks of the SAME row: an instruction then touches 32*16/(4 NF)
// rows in full lines -- 8x fewer wavefronts for NF = 32.  Chunk c of row r sits at position c ^ (r mod CH) of the
// row (conflict-free in both directions).  Rows of lanes without work are passed as nullptr.
template <int NF> struct RowXfer {
  static constexpr int CH = NF / 4;
  static constexpr bool coop = (NF % 4 == 0) && CH >= 2 && (CH & (CH - 1)) == 0;
};
// issue: the global loads of a gather, results stay in registers (nothing waits for them here)
template <int NF>
__device__ __forceinline__ void gather_issue(const float* row, float4* t, int lane) {
  constexpr int CH = RowXfer<NF>::CH;
  if constexpr (!RowXfer<NF>::coop) {
#pragma unroll
    for (int i = 0; i < NF / 4; i++) t[i] = row ? __ldg(reinterpret_cast<const float4*>(row) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
#pragma unroll
    for (int i = 0; i < CH; i++) {
      const int f = i * 32 + lane, src = f / CH, c = f % CH;
      const float* sp = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(row), src));
      t[i] = sp ? __ldg(reinterpret_cast<const float4*>(sp) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}
// commit: through the scratch, every lane ends up with its own row in v
template <int NF>
__device__ __forceinline__ void gather_commit(unsigned scr, const float4* t, float* v, int lane) {
  constexpr int CH = RowXfer<NF>::CH;
  if constexpr (!RowXfer<NF>::coop) {
#pragma unroll
    for (int i = 0; i < NF / 4; i++) { v[4 * i] = t[i].x; v[4 * i + 1] = t[i].y; v[4 * i + 2] = t[i].z; v[4 * i + 3] = t[i].w; }
  } else {
#pragma unroll
    for (int i = 0; i < CH; i++) {
      const int f = i * 32 + lane, src = f / CH, c = f % CH;
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(scr + (unsigned)((src * CH + (c ^ (src % CH))) << 4)), "f"(t[i].x),
                   "f"(t[i].y), "f"(t[i].z), "f"(t[i].w)
                   : "memory");
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < CH; c++)
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(v[4 * c]), "=f"(v[4 * c + 1]), "=f"(v[4 * c + 2]), "=f"(v[4 * c + 3])
                   : "r"(scr + (unsigned)((lane * CH + (c ^ (lane % CH))) << 4))
                   : "memory");
    __syncwarp();
  }
}
template <int NF>
__device__ __forceinline__ void scatter_rows(unsigned scr, float* row, const float* v, int lane) {
  constexpr int CH = RowXfer<NF>::CH;
  if constexpr (!RowXfer<NF>::coop) {
    if (row) store_run<NF, (NF % 4 == 0) ? 4 : 2>(row, v);
  } else {
#pragma unroll
    for (int c = 0; c < CH; c++)
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(scr + (unsigned)((lane * CH + (c ^ (lane % CH))) << 4)), "f"(v[4 * c]),
                   "f"(v[4 * c + 1]), "f"(v[4 * c + 2]), "f"(v[4 * c + 3])
                   : "memory");
    __syncwarp();
#pragma unroll
    for (int i = 0; i < CH; i++) {
      const int f = i * 32 + lane, src = f / CH, c = f % CH;
      float* dp = reinterpret_cast<float*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(row), src));
      float4 t;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                   : "r"(scr + (unsigned)((src * CH + (c ^ (src % CH))) << 4))
                   : "memory");
      if (dp) reinterpret_cast<float4*>(dp)[c] = t;
    }
    __syncwarp();
  }
}
// Rows wider than 16 floats go in 64-byte pieces, so the scratch is 2 KB per warp (16 KB per CTA) whatever the width.
constexpr unsigned kWarpScratch = 32 * 16 * 4;
template <int NF>
__device__ __forceinline__ void gather_issue_w(const float* row, float4* t, int lane) {
  if constexpr (NF > 16 && NF % 16 == 0) {
#pragma unroll
    for (int h = 0; h < NF / 16; h++) gather_issue<16>(row ? row + 16 * h : nullptr, t + 4 * h, lane);
  } else gather_issue<NF>(row, t, lane);
}
template <int NF>
__device__ __forceinline__ void gather_commit_w(unsigned scr, const float4* t, float* v, int lane) {
  if constexpr (NF > 16 && NF % 16 == 0) {
#pragma unroll
    for (int h = 0; h < NF / 16; h++) gather_commit<16>(scr, t + 4 * h, v + 16 * h, lane);
  } else gather_commit<NF>(scr, t, v, lane);
}
template <int NF>
__device__ __forceinline__ void scatter_rows_w(unsigned scr, float* row, const float* v, int lane) {
  if constexpr (NF > 16 && NF % 16 == 0) {
#pragma unroll
    for (int h = 0; h < NF / 16; h++) scatter_rows<16>(scr, row ? row + 16 * h : nullptr, v + 16 * h, lane);
  } else scatter_rows<NF>(scr, row, v, lane);
}
template <int NR> constexpr unsigned fwd_scratch_bytes() { return RowXfer<NR / 2>::coop ? 8u * kWarpScratch : 0u; }

// ================================================================================================ forward
template <int NR>
__global__ void __launch_bounds__(kTcThreads, 1)
lstm_tc_fwd(const __grid_constant__ CUtensorMap tmR_hi, const __grid_constant__ CUtensorMap tmR_lo,
            const __grid_constant__ CUtensorMap tmH32, const __grid_constant__ CUtensorMap tmH64,
            const __grid_constant__ CUtensorMap tmH128, const __grid_constant__ CUtensorMap tmHm_hi,
            const __grid_constant__ CUtensorMap tmHm_lo, Lines ln, TcFwd p) {
  constexpr int NC = NR / 2;                      // accumulator columns (gate rows) per epilogue thread
  constexpr int NU = NR / 8;                      // hidden units per epilogue thread
  constexpr int VW = (NU % 4 == 0) ? 4 : 2;       // vector width of the per-unit runs
  constexpr unsigned r_bytes = NR * 128;          // hi (or lo) part of one k chunk of the weight slice
  constexpr int TMEM_COLS = (2 * NR <= 64) ? 64 : 128;
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  __shared__ __align__(8) unsigned long long bars[2 * kMaxStages + 2];
  __shared__ unsigned tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m = blockIdx.x, q = blockIdx.z, d = p.d0 + q;
  const unsigned smem0 = (smem_u32(tc_smem) + 1023u) & ~1023u;
  const unsigned rs0 = smem0;                                   // weight slice: chunk kc at rs0 + kc*2*r_bytes (hi | lo)
  const unsigned ring0 = rs0 + (unsigned)p.KC * 2 * r_bytes;    // stage st at ring0 + st*kStageBytes (hi | lo)
  const unsigned scr0 = ring0 + (unsigned)p.nst * kStageBytes;  // per-warp row-transfer scratch of the epilogue warps
  const unsigned bar0 = smem_u32(&bars[0]);
  auto full = [&](unsigned st) { return bar0 + 8 * st; };
  auto empty = [&](unsigned st) { return bar0 + 8 * (kMaxStages + st); };
  const unsigned rfull = bar0 + 8 * (2 * kMaxStages), accfull = rfull + 8;

  // Thread-block cluster along the row slices (all of them read the SAME h tile): with CS > 1 every CTA fetches a share of
  // the rows and TMA-multicasts it into all CS shared memories -- one L2 read instead of CS, and CS request streams in
  // parallel.  A ring stage is then free only when ALL CS tensor cores are done with it: the MMA warps commit to the
  // `empty` barrier of every CTA of the cluster.
  const unsigned CS = cluster_nctarank(), crank = cluster_ctarank();
  const unsigned short cmask = (unsigned short)((1u << CS) - 1u);
  if (tid == 0) {
    for (int i = 0; i < 2 * kMaxStages + 2; i++)
      mbar_init(bar0 + 8 * i, (i >= kMaxStages && i < 2 * kMaxStages) ? CS : 1u);
    mbar_init_fence();
    tma_prefetch_desc(&tmR_hi); tma_prefetch_desc(&tmR_lo); tma_prefetch_desc(&tmH32); tma_prefetch_desc(&tmH64); tma_prefetch_desc(&tmH128);
    tma_prefetch_desc(&tmHm_hi); tma_prefetch_desc(&tmHm_lo);
  }
  __syncthreads();
  if (CS > 1) cluster_sync_all();      // no remote arrive / multicast write may meet an uninitialised barrier
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_d = tmem_base_s;
  unsigned* const flag_q = p.flags + (size_t)q * p.ntiles;
  const bool dbg = p.dbg != nullptr;
  long long* const mydbg = p.dbg + (size_t)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 32;
  long long dt[4] = {0, 0, 0, 0};
#define TC_T0 const long long t0_ = dbg ? clock64() : 0
#define TC_T(i) if (dbg) { const long long t1_ = clock64(); dt[i] += t1_ - tlast_; tlast_ = t1_; }

  // Roles 0 and 1 run their loops WARP-UNIFORMLY (all 32 lanes wait on the barriers / counters) and only the
  // instruction that must come from one thread (TMA, tcgen05.mma, tcgen05.commit) sits under elect_one(): the loop
  // bookkeeping then lives in uniform registers instead of being converted for every tensor-core instruction.
  if (warp == 0) {
    // ------------------------------------------------------------------------------------------ TMA producer
    long long tlast_ = dbg ? clock64() : 0;
    if (elect_one()) {
      mbar_expect_tx(rfull, (unsigned)p.KC * 2 * r_bytes);
      for (int kc = 0; kc < p.KC; kc++) {
        tma_load_2d(rs0 + kc * 2 * r_bytes, &tmR_hi, kc * 64, d * p.rows_pad + m * NR, rfull);
        tma_load_2d(rs0 + kc * 2 * r_bytes + r_bytes, &tmR_lo, kc * 64, d * p.rows_pad + m * NR, rfull);
      }
    }
    __syncwarp();
    unsigned it = 0;
    for (int tile = blockIdx.y; tile < p.ntiles; tile += gridDim.y) {
      const int l0 = tile * kTcLines;
      const int nl = min(kTcLines, ln.B - l0);
      const int Tt = ln.T[ln.order[l0]];
      int act = nl;
      for (int s = 1; s < Tt; s++) {
        while (act > 0 && ln.T[ln.order[l0 + act - 1]] <= s) act--;      // lines still running at step s: a prefix
        // rows to fetch: the smallest box (32 / 64 / 128 rows) that holds the running lines; ONE 3-D TMA per k chunk
        // brings the hi and the lo plane ([2][brows][64 halves], lo tile right behind the hi tile)
        const int brows = act <= 32 ? 32 : (act <= 64 ? 64 : 128);
        const CUtensorMap* mh = brows == 32 ? &tmH32 : (brows == 64 ? &tmH64 : &tmH128);
        TC_T(3);
        wait_counter<false>(flag_q + tile, (unsigned)p.nsig * (unsigned)s);   // h_{s-1} of every row slice is in L2
        TC_T(0);
        if (dbg && s == kDbgStep + 1) mydbg[16 + 3] = clock64();
        const int row0 = ((q * 2 + ((s - 1) & 1)) * p.ntiles + tile) * kTcLines;
        for (int kc = 0; kc < p.KC; kc++, it++) {
          const unsigned st = it % (unsigned)p.nst, use = it / (unsigned)p.nst;
          if (use > 0 && p.nst < p.KC) mbar_wait(empty(st), (use - 1) & 1);   // nst >= KC: the step counter already implies it
          TC_T(1);
          if (elect_one()) {
            if (kc == 0) fence_proxy_async_global();                      // generic-proxy writes (other SMs) -> async-proxy reads
            if (CS == 1) {
              mbar_expect_tx(full(st), (unsigned)brows * 256);
              tma_load_3d(ring0 + st * kStageBytes, mh, kc * 64, row0, 0, full(st));
            } else {             // 32-row boxes dealt round-robin to the CTAs of the cluster, each multicast to all of them
              const int nb32 = (act + 31) >> 5;
              mbar_expect_tx(full(st), (unsigned)nb32 * 2 * 4096);
              for (int j = (int)crank; j < nb32; j += (int)CS) {
                tma_load_2d_mc(ring0 + st * kStageBytes + j * 4096, &tmHm_hi, kc * 64, row0 + 32 * j, full(st), cmask);
                tma_load_2d_mc(ring0 + st * kStageBytes + kTcLines * 128 + j * 4096, &tmHm_lo, kc * 64, row0 + 32 * j, full(st), cmask);
              }
            }
          }
          __syncwarp();
          TC_T(2);
          if (dbg && s == kDbgStep + 1 && kc == 0) mydbg[16 + 4] = clock64();
        }
        if (dbg && s == kDbgStep + 1) mydbg[16 + 5] = clock64();
      }
    }
    if (dbg && lane == 0) for (int i = 0; i < 4; i++) mydbg[i] = dt[i];     // flag wait | ring wait | TMA issue | other
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    // Per 16 k: TWO instructions instead of three.  The hi and lo tiles of the weight slice are adjacent in shared memory,
    // so one MMA with N = 2 NR multiplies h_hi with [R_hi ; R_lo] (columns [0,NR) collect hi*hi, [NR,2NR) hi*lo), a second
    // one adds h_lo * R_hi to columns [0,NR); the epilogue adds the two column ranges.
    const unsigned idesc2 = make_idesc_f16(kTcLines, 2 * NR), idesc1 = make_idesc_f16(kTcLines, NR);
    const unsigned long long dbase = make_desc(0);
    auto desc_of = [&](unsigned addr) { return dbase | (unsigned long long)((addr & 0x3FFFF) >> 4); };
    mbar_wait(rfull, 0);
    unsigned it = 0;
    long long tlast_ = dbg ? clock64() : 0;
    for (int tile = blockIdx.y; tile < p.ntiles; tile += gridDim.y) {
      const int l0 = tile * kTcLines;
      const int Tt = ln.T[ln.order[l0]];
      int act = min(kTcLines, ln.B - l0);
      for (int s = 1; s < Tt; s++) {
        while (act > 0 && ln.T[ln.order[l0 + act - 1]] <= s) act--;      // same row count as the producer: where the lo tile starts
        const unsigned lo_off = (CS > 1) ? kTcLines * 128u : (act <= 32 ? 32u : (act <= 64 ? 64u : 128u)) * 128u;
        for (int kc = 0; kc < p.KC; kc++, it++) {
          const unsigned st = it % (unsigned)p.nst, use = it / (unsigned)p.nst;
          TC_T(2);
          mbar_wait(full(st), use & 1);
          tc_fence_after();
          TC_T(0);
          if (dbg && s == kDbgStep + 1 && kc == 0) mydbg[16 + 6] = clock64();
          if (dbg && s == kDbgStep + 1 && kc == p.KC - 1) mydbg[16 + 7] = clock64();
          const unsigned a_hi = ring0 + st * kStageBytes;
          const unsigned b_hi = rs0 + kc * 2 * r_bytes;
          // 16 k = 32 bytes along the swizzled row = +2 in the descriptor's address field
          const unsigned long long ah = desc_of(a_hi), al = desc_of(a_hi + lo_off);
          const unsigned long long bh = desc_of(b_hi);
          const int nk = min(4, p.nks - kc * 4);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
              if (ks < nk) {
                mma_f16(tmem_d, ah + 2 * ks, bh + 2 * ks, idesc2, (kc > 0 || ks > 0) ? 1u : 0u);
                mma_f16(tmem_d, al + 2 * ks, bh + 2 * ks, idesc1, 1u);
              }
            }
            if (CS == 1) mma_commit(empty(st)); else mma_commit_mc(empty(st), cmask);
            if (kc == p.KC - 1) mma_commit(accfull);
          }
          __syncwarp();
          TC_T(1);
        }
        if (dbg && s == kDbgStep + 1) mydbg[16 + 8] = clock64();
      }
    }
    if (dbg && lane == 0) for (int i = 0; i < 3; i++) mydbg[4 + i] = dt[i];   // full wait | MMA issue | other
  } else {
    // ------------------------------------------------------------------------------------------ epilogue warps
    const int ew = warp - 2;
    const int lq = warp & 3, ch = ew >> 2;                 // TMEM lane quadrant (fixed by the warp id), column half
    long long tlast_ = dbg ? clock64() : 0;
    long long de[6] = {0, 0, 0, 0, 0, 0};
#define TC_E(i) if (dbg) { const long long t1_ = clock64(); de[i] += t1_ - tlast_; tlast_ = t1_; }
    const int pl = 32 * lq + lane;                         // line slot of this thread
    const unsigned taddr = tmem_d + ((unsigned)(32 * lq) << 16) + (unsigned)(ch * NC);
    const int ub = m * (NR / 4) + ch * NU;                 // first hidden unit of this thread
    const int no = p.no, no4 = p.no4;
    const unsigned scr = scr0 + (unsigned)ew * kWarpScratch;
    const bool mine = ub + NU <= no;                       // all of this thread's units are real (no % 8 == 0)
    const float* __restrict__ XPd = p.XP[d];
    float* __restrict__ Gd = p.G[d];
    float* __restrict__ Cd = p.C[d];
    float* __restrict__ Hpd = p.Hprev[d];
    float* __restrict__ Hd = p.H + p.hoff[d];
    constexpr float inv_scale = 1.0f / (kScaleH * kScaleR);
    unsigned accph = 0;
    for (int tile = blockIdx.y; tile < p.ntiles; tile += gridDim.y) {
      const int l0 = tile * kTcLines;
      const int nl = min(kTcLines, ln.B - l0);
      const int Tt = ln.T[ln.order[l0]];
      const int li = (pl < nl) ? ln.order[l0 + pl] : -1;
      const int Tp = (li >= 0) ? ln.T[li] : 0;
      const int off = (li >= 0) ? ln.off[li] : 0;
      float c[NU];
#pragma unroll
      for (int u = 0; u < NU; u++) c[u] = 0.f;
      for (int s = 0; s < Tt; s++) {
        const bool active = s < Tp;
        const int t = d ? Tp - 1 - s : s;
        const size_t col = (size_t)off + t;
        float xp[NC];
        float4 xq[NC / 4];
        if (RowXfer<NC>::coop && NU % 4 == 0) {
          gather_issue_w<NC>((active && mine && !(p.opt & 8)) ? XPd + col * no4 + 4 * ub : nullptr, xq, lane);
        } else if (active) {
#pragma unroll
          for (int u = 0; u < NU; u++) {
            if (ub + u < no) load_run<4, 4>(xp + 4 * u, XPd + col * no4 + 4 * (ub + u));
            else { xp[4 * u] = xp[4 * u + 1] = xp[4 * u + 2] = xp[4 * u + 3] = 0.f; }
          }
        }
        float acc[NC];
        TC_E(5);
        if (s > 0) {
          mbar_wait(accfull, accph);
          accph ^= 1;
          tc_fence_after();
          TC_E(0);
          if (dbg && s == kDbgStep) mydbg[16 + 0] = clock64();
          if (dbg && s == kDbgStep + 1) mydbg[16 + 9] = clock64();
          float acc2[NC];
          tmem_ld<NC>(taddr, acc);
          tmem_ld<NC>(taddr + NR, acc2);
#pragma unroll
          for (int i = 0; i < NC; i++) acc[i] += acc2[i];
          TC_E(1);
        } else {
#pragma unroll
          for (int i = 0; i < NC; i++) acc[i] = 0.f;
        }
        if (RowXfer<NC>::coop && NU % 4 == 0) gather_commit_w<NC>(scr, xq, xp, lane);
        float gv[NC], hh[NU];
        if (active) {
#pragma unroll
          for (int u = 0; u < NU; u++) {
            const float gi = sigmoid_fast(fmaf(acc[4 * u + 0], inv_scale, xp[4 * u + 0]));   // forward_full1 clstm.cc:614-617
            const float gf = sigmoid_fast(fmaf(acc[4 * u + 1], inv_scale, xp[4 * u + 1]));
            const float go = sigmoid_fast(fmaf(acc[4 * u + 2], inv_scale, xp[4 * u + 2]));
            const float ci = tanh_fast(fmaf(acc[4 * u + 3], inv_scale, xp[4 * u + 3]));
            c[u] = fmaf(gf, c[u], ci * gi);                // forward_statemem clstm_compute.cc:504-508 (c = 0 before step 0)
            hh[u] = tanh_fast(c[u]) * go;                  // forward_nonlingate :530-537
            gv[4 * u + 0] = gi; gv[4 * u + 1] = gf; gv[4 * u + 2] = go; gv[4 * u + 3] = ci;
          }
          // publish h_s for the next step's TMA loads: fp16 hi/lo of 16 h, row = line slot, K-major
          const size_t hrow = ((size_t)((q * 2 + (s & 1)) * p.ntiles + tile) * kTcLines + pl) * p.KP + ub;
          unsigned short hi[NU], lo[NU];
#pragma unroll
          for (int u = 0; u < NU; u++) split_f16((ub + u < no) ? hh[u] * kScaleH : 0.f, hi[u], lo[u]);
          if (NU % 8 == 0) {
#pragma unroll
            for (int u = 0; u < NU; u += 8) {
              *reinterpret_cast<uint4*>(p.hx_hi + hrow + u) = make_uint4(pack_h2(hi[u], hi[u + 1]), pack_h2(hi[u + 2], hi[u + 3]),
                                                                         pack_h2(hi[u + 4], hi[u + 5]), pack_h2(hi[u + 6], hi[u + 7]));
              *reinterpret_cast<uint4*>(p.hx_lo + hrow + u) = make_uint4(pack_h2(lo[u], lo[u + 1]), pack_h2(lo[u + 2], lo[u + 3]),
                                                                         pack_h2(lo[u + 4], lo[u + 5]), pack_h2(lo[u + 6], lo[u + 7]));
            }
          } else {
#pragma unroll
            for (int u = 0; u < NU; u += 2) {
              *reinterpret_cast<unsigned*>(p.hx_hi + hrow + u) = pack_h2(hi[u], hi[u + 1]);
              *reinterpret_cast<unsigned*>(p.hx_lo + hrow + u) = pack_h2(lo[u], lo[u + 1]);
            }
          }
        }
        // this warp has read its accumulators and written its h: publish (release-increment per warp, no CTA barrier;
        // the producer of every slice waits for all 8 warps of all slices, which also protects the accumulator)
        TC_E(2);
        if (dbg && s == kDbgStep) mydbg[16 + 1] = clock64();
        tc_fence_before();
        if (p.opt & 2) fence_proxy_async();
        __syncwarp();
        if (lane == 0) signal_counter(flag_q + tile);
        TC_E(3);
        if (dbg && s == kDbgStep) mydbg[16 + 2] = clock64();
        // stash for the backward pass and the dense products, off the critical path
        if (RowXfer<NC>::coop && NU % 4 == 0) {
          const bool st = active && mine && !(p.opt & 4);
          scatter_rows_w<NC>(scr, st ? Gd + col * no4 + 4 * ub : nullptr, gv, lane);
          scatter_rows<NU>(scr, st ? Cd + col * no + ub : nullptr, c, lane);
          scatter_rows<NU>(scr, st ? Hd + col * p.hstride + ub : nullptr, hh, lane);
          scatter_rows<NU>(scr, (st && s + 1 < Tp) ? Hpd + (col + (d ? -1 : 1)) * (size_t)no + ub : nullptr, hh, lane);
          if (st && s == 0) {
            float z[NU];
#pragma unroll
            for (int u = 0; u < NU; u++) z[u] = 0.f;
            store_run<NU, VW>(Hpd + col * no + ub, z);
          }
        } else if (active && !(p.opt & 4)) {
#pragma unroll
          for (int u = 0; u < NU; u++)
            if (ub + u < no) store_run<4, 4>(Gd + col * no4 + 4 * (ub + u), gv + 4 * u);
          if (ub + NU <= no) {
            store_run<NU, VW>(Cd + col * no + ub, c);
            store_run<NU, VW>(Hd + col * p.hstride + ub, hh);
            if (s + 1 < Tp) store_run<NU, VW>(Hpd + (col + (d ? -1 : 1)) * (size_t)no + ub, hh);
            if (s == 0) {
              float z[NU];
#pragma unroll
              for (int u = 0; u < NU; u++) z[u] = 0.f;
              store_run<NU, VW>(Hpd + col * no + ub, z);
            }
          } else {
#pragma unroll
            for (int u = 0; u < NU; u++) {
              if (ub + u < no) {
                Cd[col * no + ub + u] = c[u];
                Hd[col * p.hstride + ub + u] = hh[u];
                if (s + 1 < Tp) Hpd[(col + (d ? -1 : 1)) * (size_t)no + ub + u] = hh[u];
                if (s == 0) Hpd[col * no + ub + u] = 0.f;
              }
            }
          }
        }
        TC_E(4);
      }
    }
    if (dbg && ew == 0 && lane == 0) for (int i = 0; i < 6; i++) mydbg[8 + i] = de[i];   // acc wait | tmem ld | math+h | publish | stash | xp issue
  }
  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();      // the last commits of the other CTAs still arrive on this CTA's barriers
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TMEM_COLS) : "memory");
  }
}

// ================================================================================================ backward
__global__ void __launch_bounds__(kTcThreads, 1)
lstm_tc_bwd(const __grid_constant__ CUtensorMap tmT_hi, const __grid_constant__ CUtensorMap tmT_lo, Lines ln, TcBwd p) {
  constexpr int NU = 8;                    // hidden units per epilogue thread (two warps share the CTA's 16 units)
  constexpr int TMEM_COLS = 512;           // two accumulator buffers of 256 columns
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  __shared__ __align__(8) unsigned long long bars[2 * kMaxStages + 5];
  __shared__ unsigned tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m = blockIdx.x, q = blockIdx.z, d = p.d0 + q;
  const unsigned smem0 = (smem_u32(tc_smem) + 1023u) & ~1023u;
  const unsigned a_hi = smem0, a_lo = smem0 + kTcLines * 128;          // A operand: deltas of this CTA's 64 gate rows
  const unsigned scr0 = smem0 + 2 * kTcLines * 128;                    // per-warp row-transfer scratch (8 x 4 KB)
  const unsigned ring0 = scr0 + kBwdScratch;                           // B stages (hi | lo), kBwdStageBytes each
  const unsigned bar0 = smem_u32(&bars[0]);
  auto bfull = [&](unsigned st) { return bar0 + 8 * st; };
  auto bempty = [&](unsigned st) { return bar0 + 8 * (kMaxStages + st); };
  const unsigned afull = bar0 + 8 * (2 * kMaxStages);
  auto accfull = [&](unsigned b) { return afull + 8 + 8 * b; };
  auto accempty = [&](unsigned b) { return afull + 24 + 8 * b; };

  if (tid == 0) {
    for (int i = 0; i < 2 * kMaxStages; i++) mbar_init(bar0 + 8 * i, 1);
    mbar_init(afull, kEpiThreads);
    mbar_init(accfull(0), 1); mbar_init(accfull(1), 1);
    mbar_init(accempty(0), kEpiThreads); mbar_init(accempty(1), kEpiThreads);
    mbar_init_fence();
    tma_prefetch_desc(&tmT_hi); tma_prefetch_desc(&tmT_lo);
  }
  // the A tile must never hold NaN patterns (rows of line slots that are not running are written as zeros below, but only
  // from the first step on; the MMA of a step multiplies whatever the tile holds)
  for (unsigned i = tid; i < 2 * kTcLines * 128 / 16; i += blockDim.x)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(smem0 + 16 * i), "r"(0u) : "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_d = tmem_base_s;
  unsigned* const flag_q = p.flags + (size_t)q * p.ntiles;
  const int row_t = d * p.kp_rows;                     // first row of this direction in the transposed split copy
  const bool dbg = p.dbg != nullptr;
  long long* const mydbg = p.dbg + (size_t)((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 32;

  auto load_chunk = [&](unsigned st, int chunk) {      // B chunk: outputs [256*chunk, +256) x this CTA's 64 gate rows
    mbar_expect_tx(bfull(st), kBwdStageBytes);
    const unsigned dst = ring0 + st * kBwdStageBytes;
    for (int j = 0; j < kBwdChunk / 64; j++) {
      tma_load_2d(dst + j * 8192, &tmT_hi, m * kBwdNR, row_t + chunk * kBwdChunk + 64 * j, bfull(st));
      tma_load_2d(dst + kBwdChunk * 128 + j * 8192, &tmT_lo, m * kBwdNR, row_t + chunk * kBwdChunk + 64 * j, bfull(st));
    }
  };

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------ TMA producer (weights)
    if (lane == 0) {
      if (p.resident) {
        for (int i = 0; i < p.nchunk; i++) load_chunk(i, i);
      } else {
        unsigned it = 0;
        for (int tile = blockIdx.y; tile < p.ntiles; tile += gridDim.y) {
          const int Tt = ln.T[ln.order[tile * kTcLines]];
          for (int fs = Tt - 1; fs >= 1; fs--)
            for (int i = 0; i < p.nchunk; i++, it++) {
              const unsigned st = it % (unsigned)p.nst, use = it / (unsigned)p.nst;
              if (use > 0) mbar_wait(bempty(st), (use - 1) & 1);
              load_chunk(st, i);
            }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------ MMA issuer (warp-uniform loop)
    {
      unsigned it = 0, cnt = 0, aph = 0;
      long long tlast_ = dbg ? clock64() : 0;
      long long dt[4] = {0, 0, 0, 0};
      const unsigned long long dbase = make_desc(0);
      auto desc_of = [&](unsigned addr) { return dbase | (unsigned long long)((addr & 0x3FFFF) >> 4); };
      const unsigned long long ah = desc_of(a_hi), al = desc_of(a_lo);
      for (int tile = blockIdx.y; tile < p.ntiles; tile += gridDim.y) {
        const int Tt = ln.T[ln.order[tile * kTcLines]];
        for (int fs = Tt - 1; fs >= 1; fs--) {
          TC_T(3);
          mbar_wait(afull, aph);                      // the deltas of this step are in shared memory
          aph ^= 1;
          tc_fence_after();
          TC_T(0);
          for (int i = 0; i < p.nchunk; i++, it++, cnt++) {
            const unsigned buf = cnt & 1;
            if (cnt >= 2) mbar_wait(accempty(buf), ((cnt >> 1) - 1) & 1);   // the chunk that used this buffer is drained
            unsigned st;
            if (p.resident) { st = i; if (it < (unsigned)p.nchunk) mbar_wait(bfull(st), 0); }
            else { st = it % (unsigned)p.nst; mbar_wait(bfull(st), (it / (unsigned)p.nst) & 1); }
            tc_fence_after();
            TC_T(1);
            const int nw = min(kBwdChunk, p.nop16 - i * kBwdChunk);
            const unsigned idesc = make_idesc_f16(kTcLines, nw);
            const unsigned b_hi = ring0 + st * kBwdStageBytes;
            const unsigned long long bh = desc_of(b_hi), bl = desc_of(b_hi + kBwdChunk * 128);
            const unsigned dcol = tmem_d + buf * kBwdChunk;
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < kBwdNR / 16; ks++) {
                mma_f16(dcol, al + 2 * ks, bh + 2 * ks, idesc, ks > 0 ? 1u : 0u);
                mma_f16(dcol, ah + 2 * ks, bl + 2 * ks, idesc, 1u);
                mma_f16(dcol, ah + 2 * ks, bh + 2 * ks, idesc, 1u);
              }
              if (!p.resident) mma_commit(bempty(st));
              mma_commit(accfull(buf));
            }
            __syncwarp();
            TC_T(2);
          }
        }
      }
      if (dbg && lane == 0) for (int i = 0; i < 4; i++) mydbg[i] = dt[i];   // delta wait | buffer / weight wait | MMA issue | other
    }
  } else {
    // ------------------------------------------------------------------------------------------ epilogue warps
    const int ew = warp - 2;
    const int lq = warp & 3, ch = ew >> 2;
    const int pl = 32 * lq + lane;
    const int ub = m * (kBwdNR / 4) + ch * NU;             // first hidden unit of this thread
    const int no = p.no, no4 = p.no4, NT = p.NT;
    const float* __restrict__ Gd = p.G[d];
    const float* __restrict__ Cd = p.C[d];
    const float* __restrict__ dHd = p.dH + p.hoff[d];
    float* __restrict__ DGd = p.DG[d];
    const bool mine = ub + NU <= no;                       // no % 8 == 0: a thread's units are all real or all padding
    constexpr float inv_scale = 1.0f / (kScaleD * kScaleR);
    const size_t slab = (size_t)NT * NT * (kTcLines * 16);     // floats of one [dest][src][128][16] exchange buffer
    unsigned cnt = 0;
    long long tlast_ = dbg ? clock64() : 0;
    long long de[6] = {0, 0, 0, 0, 0, 0};
    for (int tile = blockIdx.y; tile < p.ntiles; tile += gridDim.y) {
      const int l0 = tile * kTcLines;
      const int nl = min(kTcLines, ln.B - l0);
      const int Tt = ln.T[ln.order[l0]];
      const int li = (pl < nl) ? ln.order[l0 + pl] : -1;
      const int Tp = (li >= 0) ? ln.T[li] : 0;
      const int off = (li >= 0) ? ln.off[li] : 0;
      float* const part_t = p.part + (size_t)(q * p.ntiles + tile) * 2 * slab;
      float dcc[NU];
#pragma unroll
      for (int u = 0; u < NU; u++) dcc[u] = 0.f;
      for (int it = 0; it < Tt; it++) {
        const int fs = Tt - 1 - it;                        // forward step index handled now
        const bool active = mine && fs < Tp;
        const int t = d ? Tp - 1 - fs : fs;
        const size_t col = (size_t)off + t;
        float g[4 * NU], cc[NU], cp[NU], dh[NU];
        float4 gq4[NU], cq4[NU / 4], pq4[NU / 4], dq4[NU / 4];
        // operands that do not depend on the exchange: their loads fly while the step counter is polled
        gather_issue_w<4 * NU>(active ? Gd + col * no4 + 4 * ub : nullptr, gq4, lane);
        gather_issue<NU>(active ? Cd + col * no + ub : nullptr, cq4, lane);
        gather_issue<NU>((active && fs > 0) ? Cd + (col + (d ? 1 : -1)) * (size_t)no + ub : nullptr, pq4, lane);
        gather_issue<NU>(active ? dHd + col * p.hstride + ub : nullptr, dq4, lane);
        TC_E(0);
        float r[NU];
        bool have_r = false;
        if (it > 0) {                                      // partial products of the previous step, all slices
          wait_counter(flag_q + tile, (unsigned)(NT * (kEpiThreads / 32)) * (unsigned)it);
          TC_E(1);
          if (active && fs < Tp - 1) {
            // exchange layout [dest slice][src slice][4 unit quads][128 lines][4 units]: a warp reads / writes 512 contiguous bytes
            const float* src = part_t + (size_t)((it - 1) & 1) * slab + ((size_t)m * NT * 4 + 2 * ch) * (kTcLines * 4) + pl * 4;
#pragma unroll
            for (int u = 0; u < NU; u++) r[u] = 0.f;
#pragma unroll 8
            for (int sm = 0; sm < NT; sm++) {             // fixed order => deterministic
              const float4 x0 = __ldcg(reinterpret_cast<const float4*>(src + (size_t)sm * (kTcLines * 16)));
              const float4 x1 = __ldcg(reinterpret_cast<const float4*>(src + (size_t)sm * (kTcLines * 16) + kTcLines * 4));
              r[0] += x0.x; r[1] += x0.y; r[2] += x0.z; r[3] += x0.w;
              r[4] += x1.x; r[5] += x1.y; r[6] += x1.z; r[7] += x1.w;
            }
            have_r = true;
          }
        }
        {
          const unsigned scr = scr0 + (unsigned)ew * kWarpScratch;
          gather_commit_w<4 * NU>(scr, gq4, g, lane);
          gather_commit<NU>(scr, cq4, cc, lane);
          gather_commit<NU>(scr, pq4, cp, lane);
          gather_commit<NU>(scr, dq4, dh, lane);
        }
        if (have_r) {
#pragma unroll
          for (int u = 0; u < NU; u++) dh[u] += r[u];
        }
        TC_E(2);
        unsigned hi[2 * NU], lo[2 * NU];                    // packed half2: 4 gate rows of a unit = 2 words
        if (active) {
          float dl[4 * NU];
#pragma unroll
          for (int u = 0; u < NU; u++) {
            const float gi = g[4 * u], gf = g[4 * u + 1], go = g[4 * u + 2], ci = g[4 * u + 3];
            const float th = tanh_fast(cc[u]);                            // backward_nonlingate clstm_compute.cc:539-547
            const float dgo = th * dh[u];
            const float dc = fmaf(1.f - th * th, go * dh[u], dcc[u]);
            float dgf = 0.f, carry = 0.f;
            if (fs > 0) { dgf = dc * cp[u]; carry = dc * gf; }            // backward_statemem :509-515
            dcc[u] = carry;
            const float dgi = dc * ci, dci = dc * gi;
            dl[4 * u + 0] = gi * (1.f - gi) * dgi;                        // backward_nonlin0 :231-267
            dl[4 * u + 1] = gf * (1.f - gf) * dgf;
            dl[4 * u + 2] = go * (1.f - go) * dgo;
            dl[4 * u + 3] = (1.f - ci * ci) * dci;
          }
#pragma unroll
          for (int i = 0; i < 4 * NU; i++) g[i] = dl[i];   // (the gate values are dead: reuse their registers for the DG rows)
#pragma unroll
          for (int i = 0; i < 2 * NU; i++) {
            unsigned short h0, l0_, h1, l1_;
            split_f16(dl[2 * i] * kScaleD, h0, l0_);
            split_f16(dl[2 * i + 1] * kScaleD, h1, l1_);
            hi[i] = pack_h2(h0, h1);
            lo[i] = pack_h2(l0_, l1_);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 2 * NU; i++) { hi[i] = 0u; lo[i] = 0u; }
        }
        scatter_rows_w<4 * NU>(scr0 + (unsigned)ew * kWarpScratch, active ? DGd + col * no4 + 4 * ub : nullptr, g, lane);
        if (fs == 0) break;                                // the first forward step has no predecessor: nothing to propagate
        // A operand row `pl`, k = 32*ch .. 32*ch+31: four 16-byte chunks of the 128-byte swizzled row
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const unsigned chunk = (unsigned)(ch * 4 + i) ^ (unsigned)(pl & 7);
          const unsigned o = (unsigned)pl * 128 + (chunk << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + o), "r"(hi[4 * i]), "r"(hi[4 * i + 1]),
                       "r"(hi[4 * i + 2]), "r"(hi[4 * i + 3])
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lo + o), "r"(lo[4 * i]), "r"(lo[4 * i + 1]),
                       "r"(lo[4 * i + 2]), "r"(lo[4 * i + 3])
                       : "memory");
        }
        fence_proxy_async_smem();
        mbar_arrive(afull);
        TC_E(3);
        // drain the output chunks into the exchange buffer: group gq of 16 outputs belongs to slice (16 chunk + gq)
        float* const dst_par = part_t + (size_t)(it & 1) * slab + (size_t)m * (kTcLines * 16) + pl * 4;
        const bool wr = (li >= 0) && fs < Tp;              // rows of lines that are not running are never read
        for (int i = 0; i < p.nchunk; i++, cnt++) {
          const unsigned buf = cnt & 1;
          mbar_wait(accfull(buf), (cnt >> 1) & 1);
          tc_fence_after();
          const int nw = min(kBwdChunk, p.nop16 - i * kBwdChunk);
          const unsigned tbase = tmem_d + ((unsigned)(32 * lq) << 16) + buf * kBwdChunk;
          for (int gq = ch; gq < nw / 16; gq += 2) {
            float v[16];
            tmem_ld<16>(tbase + 16 * gq, v);
            if (wr) {
              float* o = dst_par + (size_t)(i * 16 + gq) * NT * (kTcLines * 16);
#pragma unroll
              for (int e = 0; e < 16; e += 4)
                __stcg(reinterpret_cast<float4*>(o + e * kTcLines),
                       make_float4(v[e] * inv_scale, v[e + 1] * inv_scale, v[e + 2] * inv_scale, v[e + 3] * inv_scale));
            }
          }
          tc_fence_before();
          mbar_arrive(accempty(buf));
        }
        TC_E(4);
        __syncwarp();
        if (lane == 0) signal_counter(flag_q + tile);      // per warp: 8 increments per slice and step
        TC_E(5);
      }
    }
    if (dbg && ew == 0 && lane == 0) for (int i = 0; i < 6; i++) mydbg[8 + i] = de[i];   // loads | flag wait | reduce | pointwise | drain | publish
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TMEM_COLS) : "memory");
  }
}

// ================================================================================================ weight split
// R [4no x no] fp32 (row-major, gate-interleaved rows) -> fp16 hi/lo of 16 R, once per update:
//   Rs [rows_pad x KP]     rows = gate rows, K = k   : B operand of the forward kernel (box NR x 64)
//   Rt [kp_rows x RP]      rows = k, K = gate rows   : B operand of the backward kernel (box 64 x 64)
// Padding (rows / columns beyond the matrix) stays zero from the allocation.
__global__ void lstm_tc_split_kernel(const float* __restrict__ R, int no, __half* __restrict__ rs_hi, __half* __restrict__ rs_lo,
                                     int KP, __half* __restrict__ rt_hi, __half* __restrict__ rt_lo, int RP) {
  const size_t total = (size_t)4 * no * no;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / no), k = (int)(i % no);
    unsigned short hi, lo;
    split_f16(R[i] * kScaleR, hi, lo);
    reinterpret_cast<unsigned short*>(rs_hi)[(size_t)r * KP + k] = hi;
    reinterpret_cast<unsigned short*>(rs_lo)[(size_t)r * KP + k] = lo;
    reinterpret_cast<unsigned short*>(rt_hi)[(size_t)k * RP + r] = hi;
    reinterpret_cast<unsigned short*>(rt_lo)[(size_t)k * RP + r] = lo;
  }
}

// ================================================================================================ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int load_encode() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) return 1;
  g_encode = (EncodeTiledFn)fn;
  return 0;
}
// 2-D fp16 tensor [rows][cols] (cols contiguous), box [box_rows][64 halves = 128 bytes], SWIZZLE_128B
int make_map(CUtensorMap* m, void* base, size_t rows, size_t cols, int box_rows) {
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
  const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  return g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? 0 : 1;
}
// 3-D view of two stacked [rows][cols] fp16 planes (hi, lo): box [2][box_rows][64 halves]
int make_map3(CUtensorMap* m, void* base, size_t rows, size_t cols, int box_rows) {
  const cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 2};
  const cuuint64_t gstride[2] = {(cuuint64_t)cols * 2, (cuuint64_t)rows * cols * 2};
  const cuuint32_t box[3] = {64u, (cuuint32_t)box_rows, 2u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  return g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? 0 : 1;
}
constexpr size_t kSmemLimit = 232448;     // 227 KB per CTA on sm_100
constexpr size_t kSmemSlack = 1024 + 256; // alignment of the dynamic part + static barriers

template <int NR> size_t fwd_smem(int KC, int nst) { return (size_t)KC * 2 * NR * 128 + (size_t)nst * kStageBytes + fwd_scratch_bytes<NR>() + 1024; }

}  // namespace

struct LstmTcPlan {
  int no = 0, num_sms = 148;
  int KC = 0, KP = 0, nks = 0;
  int nr_max_rows = 0;            // rows per direction of the forward split copy (4no padded to a multiple of 192)
  int kp_rows = 0, RP = 0;        // backward split copy: rows per direction (outputs padded to 256), row pitch (gate rows padded to 64)
  __half *rs_hi = nullptr, *rs_lo = nullptr, *rt_hi = nullptr, *rt_lo = nullptr;
  bool stale[2] = {true, true};
  // exchange buffers, sized for cap_tiles line tiles
  int cap_tiles = 0, cap_nt_b = 0;
  __half *hx_hi = nullptr, *hx_lo = nullptr;
  float* part = nullptr;
  unsigned* flags = nullptr;
  CUtensorMap tmR_hi[3], tmR_lo[3];   // box rows 32 / 48 / 64
  CUtensorMap tmT_hi, tmT_lo, tmH[3];   // h exchange maps (3-D: k, row, hi/lo plane): box rows 32 / 64 / 128
  CUtensorMap tmHm[2];                  // 2-D 32-row maps of the hi / lo plane (cluster multicast path)
  int cluster = -1;                     // CLSTM_B200_TC_CLUSTER: -1 auto (largest divisor <= 8 of the slice count), 0/1 off
  bool coop = true;
  int opt = 0;                    // CLSTM_B200_TC_OPT tuning switches
  int force_nr = 0;               // CLSTM_B200_TC_NR: forward row-slice width (tuning)
  int last_ctas[2] = {0, 0};      // grid size of the last forward / backward launch (debug counters: one slot per CTA)
  int last_gx[2] = {1, 1};
  int last_cluster = 1;
  long long* dbg = nullptr;       // 16 cycle counters + 16 event timestamps of step kDbgStep (CLSTM_B200_TC_DBG=1 or the self-test)
  long long dbg_host[2][kDbgCtas * 32] = {};
  char err[256] = {0};
};

namespace {
int fwd_nr_options[3] = {32, 48, 64};      // tensor-map order
int fwd_nr_try[3] = {32, 64, 48};          // preference: the widths with cooperative row transfers first

// choose the forward row-slice width: the smallest NR that lets every tile run concurrently and fits shared memory; if no
// width does, the smallest that fits at all (tiles then run in groups)
bool pick_fwd(const LstmTcPlan* pl, int ndir, int ntiles, int* NR, int* NT, int* tg, int* nst) {
  for (int pass = 0; pass < 2; pass++) {
    for (int o = 0; o < 3; o++) {
      const int nr = fwd_nr_try[o];
      if (pl->force_nr && nr != pl->force_nr) continue;
      const int nt = (4 * pl->no + nr - 1) / nr;
      if (nt * ndir > pl->num_sms) continue;
      const size_t base = (size_t)pl->KC * 2 * nr * 128 + kSmemSlack + (nr == 32 ? fwd_scratch_bytes<32>() : (nr == 48 ? fwd_scratch_bytes<48>() : fwd_scratch_bytes<64>()));
      if (base + 2 * kStageBytes > kSmemLimit) continue;
      int st = (int)((kSmemLimit - base) / kStageBytes);
      st = std::min(st, std::min(kMaxStages, std::max(2, pl->KC)));
      const int groups = std::min(ntiles, pl->num_sms / (nt * ndir));
      if (pass == 0 && groups < ntiles) continue;
      *NR = nr; *NT = nt; *tg = groups; *nst = st;
      return true;
    }
  }
  return false;
}
}  // namespace

bool lstm_tc_supported(int no) {
  if (no < 32 || no % 8 != 0 || no > 1024) return false;
  return true;
}

void lstm_tc_destroy(LstmTcPlan* p) {
  if (!p) return;
  cudaFree(p->rs_hi); cudaFree(p->rs_lo); cudaFree(p->rt_hi); cudaFree(p->rt_lo);
  cudaFree(p->hx_hi); cudaFree(p->part); cudaFree(p->flags); cudaFree(p->dbg);
  delete p;
}

const char* lstm_tc_error(const LstmTcPlan* p) { return p ? p->err : "no plan"; }

LstmTcPlan* lstm_tc_create(int no, int num_sms) {
  if (!lstm_tc_supported(no) || load_encode() != 0) return nullptr;
  auto* p = new LstmTcPlan;
  p->no = no; p->num_sms = num_sms;
  p->KC = (no + 63) / 64; p->KP = p->KC * 64; p->nks = (no + 15) / 16;
  p->nr_max_rows = ((4 * no + 191) / 192) * 192;                 // multiple of 32, 48 and 64
  p->kp_rows = ((no + kBwdChunk - 1) / kBwdChunk) * kBwdChunk;
  p->RP = ((4 * no + 63) / 64) * 64;
  const size_t rs = (size_t)2 * p->nr_max_rows * p->KP, rt = (size_t)2 * p->kp_rows * p->RP;
  bool ok = cudaMalloc((void**)&p->rs_hi, rs * 2) == cudaSuccess && cudaMalloc((void**)&p->rs_lo, rs * 2) == cudaSuccess &&
            cudaMalloc((void**)&p->rt_hi, rt * 2) == cudaSuccess && cudaMalloc((void**)&p->rt_lo, rt * 2) == cudaSuccess;
  if (ok) {
    cudaMemset(p->rs_hi, 0, rs * 2); cudaMemset(p->rs_lo, 0, rs * 2);
    cudaMemset(p->rt_hi, 0, rt * 2); cudaMemset(p->rt_lo, 0, rt * 2);
    for (int o = 0; o < 3 && ok; o++)
      ok = make_map(&p->tmR_hi[o], p->rs_hi, (size_t)2 * p->nr_max_rows, p->KP, fwd_nr_options[o]) == 0 &&
           make_map(&p->tmR_lo[o], p->rs_lo, (size_t)2 * p->nr_max_rows, p->KP, fwd_nr_options[o]) == 0;
    ok = ok && make_map(&p->tmT_hi, p->rt_hi, (size_t)2 * p->kp_rows, p->RP, 64) == 0 &&
         make_map(&p->tmT_lo, p->rt_lo, (size_t)2 * p->kp_rows, p->RP, 64) == 0;
  }
  if (const char* e = getenv("CLSTM_B200_TC_COOP")) p->coop = atoi(e) != 0;
  if (const char* e = getenv("CLSTM_B200_TC_OPT")) p->opt = atoi(e);
  if (const char* e = getenv("CLSTM_B200_TC_NR")) p->force_nr = atoi(e);
  if (const char* e = getenv("CLSTM_B200_TC_CLUSTER")) p->cluster = atoi(e);
  if (const char* e = getenv("CLSTM_B200_TC_DBG")) {
    if (atoi(e) != 0 && cudaMalloc((void**)&p->dbg, kDbgCtas * 32 * sizeof(long long)) == cudaSuccess) cudaMemset(p->dbg, 0, kDbgCtas * 32 * sizeof(long long));
  }
  if (!ok) { lstm_tc_destroy(p); return nullptr; }
  return p;
}

// per-phase cycle counters of the last forward (which = 0) / backward (1) launch; the stream must be idle
const long long* lstm_tc_debug_counters(LstmTcPlan* p, int which) {
  if (!p || !p->dbg) return nullptr;
  cudaMemcpy(p->dbg_host[which & 1], p->dbg, kDbgCtas * 32 * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaMemset(p->dbg, 0, kDbgCtas * 32 * sizeof(long long));
  return p->dbg_host[which & 1];
}

void lstm_tc_mark_stale(LstmTcPlan* p) { if (p) p->stale[0] = p->stale[1] = true; }

namespace {
int ensure_split(LstmTcPlan* p, cudaStream_t st, const float* const R[2], int d0, int ndir) {
  for (int d = d0; d < d0 + ndir; d++) {
    if (!p->stale[d]) continue;
    const size_t so = (size_t)d * p->nr_max_rows * p->KP, to = (size_t)d * p->kp_rows * p->RP;
    const size_t total = (size_t)4 * p->no * p->no;
    const int nb = (int)std::min<size_t>((total + 255) / 256, (size_t)p->num_sms * 8);
    lstm_tc_split_kernel<<<nb, 256, 0, st>>>(R[d], p->no, p->rs_hi + so, p->rs_lo + so, p->KP, p->rt_hi + to, p->rt_lo + to, p->RP);
    p->stale[d] = false;
  }
  return (int)cudaGetLastError();
}
int ensure_exchange(LstmTcPlan* p, cudaStream_t st, int ntiles, int nt_b) {
  if (ntiles <= p->cap_tiles && nt_b <= p->cap_nt_b) return 0;
  cudaStreamSynchronize(st);
  cudaFree(p->hx_hi); cudaFree(p->part); cudaFree(p->flags);
  p->hx_hi = p->hx_lo = nullptr; p->part = nullptr; p->flags = nullptr;
  const int ct = std::max(ntiles, p->cap_tiles), cn = std::max(nt_b, p->cap_nt_b);
  const size_t hrows = (size_t)2 * 2 * ct * kTcLines;
  const size_t pf = (size_t)2 * ct * 2 * cn * cn * (kTcLines * 16);
  if (cudaMalloc((void**)&p->hx_hi, 2 * hrows * p->KP * 2) != cudaSuccess ||
      cudaMalloc((void**)&p->part, pf * sizeof(float)) != cudaSuccess || cudaMalloc((void**)&p->flags, (size_t)2 * ct * sizeof(unsigned)) != cudaSuccess) {
    snprintf(p->err, sizeof p->err, "out of memory for the exchange buffers (%d tiles)", ct);
    p->cap_tiles = 0; p->cap_nt_b = 0;
    return 1;
  }
  p->hx_lo = p->hx_hi + hrows * p->KP;      // lo plane right behind the hi plane
  cudaMemsetAsync(p->hx_hi, 0, 2 * hrows * p->KP * 2, st);
  for (int o = 0; o < 3; o++)
    if (make_map3(&p->tmH[o], p->hx_hi, hrows, p->KP, 32 << o) != 0) {
      snprintf(p->err, sizeof p->err, "cuTensorMapEncodeTiled failed for the h exchange buffer");
      return 1;
    }
  if (make_map(&p->tmHm[0], p->hx_hi, hrows, p->KP, 32) != 0 || make_map(&p->tmHm[1], p->hx_lo, hrows, p->KP, 32) != 0) {
    snprintf(p->err, sizeof p->err, "cuTensorMapEncodeTiled failed for the h exchange buffer (multicast maps)");
    return 1;
  }
  p->cap_tiles = ct; p->cap_nt_b = cn;
  return 0;
}

template <class K, class... Args>
cudaError_t launch_coop(K kernel, dim3 grid, size_t smem, cudaStream_t st, bool coop, int cluster_x, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (coop) { at[na].id = cudaLaunchAttributeCooperative; at[na].val.cooperative = 1; na++; }
  if (cluster_x > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = (unsigned)cluster_x; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
    na++;
  }
  cfg.attrs = at; cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
int pick_cluster(const LstmTcPlan* p, int NT) {
  if (p->cluster == 0 || p->cluster == 1) return 1;
  if (p->cluster > 1) return (NT % p->cluster == 0) ? p->cluster : 1;
  for (int c = 8; c > 1; c--)
    if (NT % c == 0) return c;
  return 1;
}
}  // namespace

int lstm_tc_configure() {
  cudaError_t e = cudaFuncSetAttribute(lstm_tc_fwd<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemLimit - 1024));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_tc_fwd<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemLimit - 1024));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_tc_fwd<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemLimit - 1024));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_tc_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemLimit - 1024));
  return (int)e;
}

// 0: launched; -1: this batch / size is not handled by the tensor-core recurrence (caller uses another variant);
// > 0: CUDA error (message in lstm_tc_error)
int lstm_tc_forward(LstmTcPlan* p, cudaStream_t st, const Lines& ln, const LstmFwdArgs& a) {
  if (!p || a.cell != 0 || a.no != p->no) return -1;
  const int ntiles = (ln.B + kTcLines - 1) / kTcLines;
  int NR, NT, tg, nst;
  if (!pick_fwd(p, a.ndir, ntiles, &NR, &NT, &tg, &nst)) return -1;
  const int nt_b = (4 * p->no + kBwdNR - 1) / kBwdNR;
  if (ensure_exchange(p, st, ntiles, nt_b) != 0) return 1;
  if (ensure_split(p, st, a.R, a.d0, a.ndir) != 0) { snprintf(p->err, sizeof p->err, "weight split launch failed"); return 1; }
  cudaMemsetAsync(p->flags, 0, (size_t)2 * p->cap_tiles * sizeof(unsigned), st);
  TcFwd f{};
  f.no = p->no; f.no4 = 4 * p->no; f.KC = p->KC; f.nks = p->nks; f.NT = NT; f.ntiles = ntiles; f.nst = nst;
  f.nsig = NT * (kEpiThreads / 32);
  f.d0 = a.d0; f.hstride = a.hstride; f.hoff[0] = a.hoff[0]; f.hoff[1] = a.hoff[1];
  f.rows_pad = p->nr_max_rows; f.KP = p->KP;
  for (int d = 0; d < 2; d++) { f.XP[d] = a.XP[d]; f.G[d] = a.G[d]; f.C[d] = a.C[d]; f.Hprev[d] = a.Hprev[d]; }
  f.H = a.H; f.hx_hi = p->hx_hi; f.hx_lo = p->hx_lo; f.flags = p->flags; f.dbg = p->dbg; f.opt = p->opt;
  const dim3 grid(NT, tg, a.ndir);
  p->last_ctas[0] = NT * tg * a.ndir; p->last_gx[0] = NT;
  cudaError_t e = cudaSuccess;
  const int o = NR == 32 ? 0 : (NR == 48 ? 1 : 2);
  int cs = pick_cluster(p, NT);
  for (int attempt = 0; attempt < 2; attempt++) {        // a cluster shape the device cannot place falls back to unicast
    if (NR == 32) e = launch_coop(lstm_tc_fwd<32>, grid, fwd_smem<32>(p->KC, nst), st, p->coop, cs, p->tmR_hi[o], p->tmR_lo[o], p->tmH[0], p->tmH[1], p->tmH[2], p->tmHm[0], p->tmHm[1], ln, f);
    else if (NR == 48) e = launch_coop(lstm_tc_fwd<48>, grid, fwd_smem<48>(p->KC, nst), st, p->coop, cs, p->tmR_hi[o], p->tmR_lo[o], p->tmH[0], p->tmH[1], p->tmH[2], p->tmHm[0], p->tmHm[1], ln, f);
    else e = launch_coop(lstm_tc_fwd<64>, grid, fwd_smem<64>(p->KC, nst), st, p->coop, cs, p->tmR_hi[o], p->tmR_lo[o], p->tmH[0], p->tmH[1], p->tmH[2], p->tmHm[0], p->tmHm[1], ln, f);
    if (e == cudaSuccess || cs == 1) break;
    cudaGetLastError();
    cs = 1;
  }
  p->last_cluster = cs;
  if (e != cudaSuccess) {
    snprintf(p->err, sizeof p->err, "lstm_tc_fwd<%d> launch (grid %d x %d x %d, %d stages): %s", NR, NT, tg, a.ndir, nst, cudaGetErrorString(e));
    cudaGetLastError();
    return (int)e;
  }
  return 0;
}

int lstm_tc_backward(LstmTcPlan* p, cudaStream_t st, const Lines& ln, const LstmBwdArgs& a) {
  if (!p || a.cell != 0 || a.no != p->no) return -1;
  const int ntiles = (ln.B + kTcLines - 1) / kTcLines;
  const int NT = (4 * p->no + kBwdNR - 1) / kBwdNR;
  if (NT * a.ndir > p->num_sms) return -1;
  const int tg = std::min(ntiles, p->num_sms / (NT * a.ndir));
  const int nop16 = ((p->no + 15) / 16) * 16;
  const int nchunk = (nop16 + kBwdChunk - 1) / kBwdChunk;
  const size_t a_bytes = 2 * kTcLines * 128 + kBwdScratch;
  const int max_st = (int)((kSmemLimit - kSmemSlack - a_bytes) / kBwdStageBytes);   // 3
  const bool resident = nchunk <= std::min(max_st, kMaxStages);
  const int nst = resident ? nchunk : std::min(max_st, kMaxStages);
  if (nst < 2 && !resident) return -1;
  if (ensure_exchange(p, st, ntiles, NT) != 0) return 1;
  if (ensure_split(p, st, a.R, a.d0, a.ndir) != 0) { snprintf(p->err, sizeof p->err, "weight split launch failed"); return 1; }
  cudaMemsetAsync(p->flags, 0, (size_t)2 * p->cap_tiles * sizeof(unsigned), st);
  TcBwd b{};
  b.no = p->no; b.no4 = 4 * p->no; b.NT = NT; b.ntiles = ntiles; b.nst = nst; b.resident = resident ? 1 : 0;
  b.nchunk = nchunk; b.nop16 = nop16;
  b.d0 = a.d0; b.hstride = a.hstride; b.hoff[0] = a.hoff[0]; b.hoff[1] = a.hoff[1];
  b.kp_rows = p->kp_rows;
  for (int d = 0; d < 2; d++) { b.G[d] = a.G[d]; b.C[d] = a.C[d]; b.DG[d] = a.DG[d]; }
  b.dH = a.dH; b.part = p->part; b.flags = p->flags; b.dbg = p->dbg; b.opt = p->opt;
  const size_t smem = a_bytes + (size_t)nst * kBwdStageBytes + 1024;
  const dim3 grid(NT, tg, a.ndir);
  p->last_ctas[1] = NT * tg * a.ndir; p->last_gx[1] = NT;
  cudaError_t e = launch_coop(lstm_tc_bwd, grid, smem, st, p->coop, 1, p->tmT_hi, p->tmT_lo, ln, b);
  if (e != cudaSuccess) {
    snprintf(p->err, sizeof p->err, "lstm_tc_bwd launch (grid %d x %d x %d, %d stages%s): %s", NT, tg, a.ndir, nst,
             resident ? ", resident" : "", cudaGetErrorString(e));
    cudaGetLastError();
    return (int)e;
  }
  return 0;
}


// ================================================================================================ self-test (device A/B)
namespace {
// min / median / max over the CTAs of one debug counter, and which CTA holds the minimum
void spread(const char* name, const long long* c, int nctas, int idx, int gx) {
  std::vector<std::pair<long long, int>> v;
  for (int i = 0; i < nctas && i < kDbgCtas; i++) v.push_back({c[(size_t)i * 32 + idx], i});
  if (v.empty()) return;
  std::sort(v.begin(), v.end());
  fprintf(stderr, "    %-10s min %lld (cta %d = slice %d) median %lld max %lld (cta %d = slice %d)\n", name, v.front().first, v.front().second,
          v.front().second % gx, v[v.size() / 2].first, v.back().first, v.back().second, v.back().second % gx);
}
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { cudaFree(p); }
  template <class T> T* as() { return reinterpret_cast<T*>(p); }
  bool alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16) == cudaSuccess; }
};
float max_abs_diff(const char* what, const std::vector<float>& a, const std::vector<float>& b, float* maxref = nullptr) {
  float m = 0.f, r = 0.f;
  size_t nan_a = 0, nan_b = 0, first = (size_t)-1, worst = 0, firstbad = (size_t)-1;
  for (size_t i = 0; i < a.size(); i++) {
    if (std::isnan(a[i])) nan_a++;
    if (std::isnan(b[i])) { nan_b++; if (first == (size_t)-1) first = i; }
    const float d = std::fabs(a[i] - b[i]);
    if (d > m) { m = d; worst = i; }
    if (d > 1e-3f && firstbad == (size_t)-1) firstbad = i;
    if (std::fabs(a[i]) > r) r = std::fabs(a[i]);
  }
  if (nan_a || nan_b) {
    fprintf(stderr, "selftest_lstm %s: %zu NaN in the SIMT result, %zu NaN in the tensor-core result (first at %zu of %zu)\n", what,
            nan_a, nan_b, first, a.size());
    m = std::nanf("");
  }
  if (firstbad != (size_t)-1)
    fprintf(stderr, "selftest_lstm %s: first |diff| > 1e-3 at %zu (simt %g, tc %g), worst at %zu (simt %g, tc %g) of %zu\n", what, firstbad,
            a[firstbad], b[firstbad], worst, a[worst], b[worst], a.size());
  if (maxref) *maxref = r;
  return m;
}
}  // namespace

int lstm_tc_selftest(int no, int B, int Tmin, int Tmax, unsigned seed, float wscale, float* out, char* msg, int msglen, int variant) {
  auto say = [&](const char* m) { if (msg && msglen > 0) snprintf(msg, msglen, "%s", m); };
  say("");
  for (int i = 0; i < 9; i++) out[i] = -1.f;
  if (variant == 1 ? !lstm_tcx_supported(no) : !lstm_tc_supported(no)) { say("size not supported"); return 1; }
  cudaDeviceProp prop;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaGetDeviceProperties(&prop, dev);
  if (lstm_configure() != 0 || lstm_tc_configure() != 0) { say("configure failed"); return 1; }
  LstmTcPlan* plan = (variant == 1) ? nullptr : lstm_tc_create(no, prop.multiProcessorCount);
  if (variant != 1 && !plan) { say("lstm_tc_create failed"); return 1; }
  LstmTcxPlan* xplan = (variant == 1) ? lstm_tcx_create(no, prop.multiProcessorCount) : nullptr;
  if (variant == 1 && !xplan) { say("lstm_tcx_create failed"); lstm_tc_destroy(plan); return 1; }
  if (plan && !plan->dbg && cudaMalloc((void**)&plan->dbg, kDbgCtas * 32 * sizeof(long long)) == cudaSuccess) cudaMemset(plan->dbg, 0, kDbgCtas * 32 * sizeof(long long));
  unsigned long long rng = 0x9E3779B97F4A7C15ull ^ seed;
  auto uni = [&]() { rng = rng * 6364136223846793005ull + 1442695040888963407ull; return (float)((rng >> 40) & 0xFFFFFF) / 16777216.f; };
  auto nrm = [&]() { float s = 0.f; for (int i = 0; i < 4; i++) s += uni(); return (s - 2.f) * 1.7320508f; };
  std::vector<int> T(B), off(B), order(B), zero(B, 0);
  int N = 0, tmax = 0;
  for (int b = 0; b < B; b++) { T[b] = Tmin + (int)(uni() * (Tmax - Tmin + 1)); if (T[b] > Tmax) T[b] = Tmax; off[b] = N; N += T[b]; tmax = std::max(tmax, T[b]); }
  for (int b = 0; b < B; b++) order[b] = b;
  std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return T[a] > T[c]; });
  const size_t n4 = (size_t)N * 4 * no, n1 = (size_t)N * no, n2 = (size_t)N * 2 * no, nr = (size_t)4 * no * no;
  std::vector<float> hR(2 * nr), hXP(2 * n4), hdH(n2);
  for (auto& v : hR) v = nrm() * wscale;
  for (auto& v : hXP) v = nrm();
  for (auto& v : hdH) v = nrm() * 0.01f;
  DevBuf dT, dOff, dOrd, dR, dRt, dXP, dG[2], dC[2], dH[2], dHp[2], ddH, dDG[2];
  bool ok = dT.alloc(B * 4) && dOff.alloc(B * 4) && dOrd.alloc(B * 4) && dR.alloc(2 * nr * 4) && dRt.alloc(2 * nr * 4) &&
            dXP.alloc(2 * n4 * 4) && ddH.alloc(n2 * 4);
  for (int v = 0; v < 2; v++)
    ok = ok && dG[v].alloc(2 * n4 * 4) && dC[v].alloc(2 * n1 * 4) && dH[v].alloc(n2 * 4) && dHp[v].alloc(2 * n1 * 4) && dDG[v].alloc(2 * n4 * 4);
  if (!ok) { say("out of memory"); lstm_tc_destroy(plan); return 1; }
  cudaMemcpy(dT.p, T.data(), B * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dOff.p, off.data(), B * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dOrd.p, order.data(), B * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dR.p, hR.data(), 2 * nr * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dXP.p, hXP.data(), 2 * n4 * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(ddH.p, hdH.data(), n2 * 4, cudaMemcpyHostToDevice);
  {  // transposed copy for the generic forward kernel
    std::vector<float> hRt(2 * nr);
    for (int d = 0; d < 2; d++)
      for (int r = 0; r < 4 * no; r++)
        for (int k = 0; k < no; k++) hRt[d * nr + (size_t)k * 4 * no + r] = hR[d * nr + (size_t)r * no + k];
    cudaMemcpy(dRt.p, hRt.data(), 2 * nr * 4, cudaMemcpyHostToDevice);
  }
  for (int v = 0; v < 2; v++) {
    cudaMemset(dG[v].p, 0xff, 2 * n4 * 4); cudaMemset(dC[v].p, 0xff, 2 * n1 * 4); cudaMemset(dH[v].p, 0xff, n2 * 4);
    cudaMemset(dHp[v].p, 0xff, 2 * n1 * 4); cudaMemset(dDG[v].p, 0xff, 2 * n4 * 4);
  }
  cudaDeviceSynchronize();   // the memsets above run on the legacy stream, the kernels below on a non-blocking stream
  Lines ln{};
  ln.B = B; ln.N = N; ln.Tmax = tmax; ln.T = dT.as<int>(); ln.off = dOff.as<int>(); ln.order = dOrd.as<int>();
  cudaStream_t st;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  cudaEvent_t ev[5];
  for (auto& e : ev) cudaEventCreate(&e);
  auto fargs = [&](int v) {
    LstmFwdArgs a;
    a.no = no; a.d0 = 0; a.ndir = 2; a.hstride = 2 * no; a.hoff[0] = 0; a.hoff[1] = no; a.cell = 0;
    for (int d = 0; d < 2; d++) {
      a.XP[d] = dXP.as<float>() + d * n4; a.R[d] = dR.as<float>() + d * nr; a.Rt[d] = dRt.as<float>() + d * nr;
      a.G[d] = dG[v].as<float>() + d * n4; a.C[d] = dC[v].as<float>() + d * n1; a.Hprev[d] = dHp[v].as<float>() + d * n1;
    }
    a.H = dH[v].as<float>();
    return a;
  };
  auto bargs = [&](int v) {   // both variants differentiate the SAME stash (the generic forward's)
    LstmBwdArgs a;
    a.no = no; a.d0 = 0; a.ndir = 2; a.hstride = 2 * no; a.hoff[0] = 0; a.hoff[1] = no; a.cell = 0; a.dH = ddH.as<float>();
    for (int d = 0; d < 2; d++) {
      a.R[d] = dR.as<float>() + d * nr; a.G[d] = dG[0].as<float>() + d * n4; a.C[d] = dC[0].as<float>() + d * n1;
      a.DG[d] = dDG[v].as<float>() + d * n4;
    }
    return a;
  };
  int rc = 0;
  float ms[4] = {-1.f, -1.f, -1.f, -1.f};
  cudaEventRecord(ev[0], st);
  lstm_forward_generic(st, ln, fargs(0));
  cudaEventRecord(ev[1], st);
  lstm_backward_generic(st, ln, bargs(0));
  cudaEventRecord(ev[2], st);
  if (cudaStreamSynchronize(st) != cudaSuccess) { say("generic kernels failed"); rc = 2; }
  if (!rc) { cudaEventElapsedTime(&ms[2], ev[0], ev[1]); cudaEventElapsedTime(&ms[3], ev[1], ev[2]); }
  const char* only = getenv("CLSTM_B200_SELFTEST_ONLY");   // "fwd" / "bwd": run one pass only (a trap in one does not hide the other)
  const bool run_fwd = !(only && only[0] == 'b'), run_bwd = !(only && only[0] == 'f');
  if (!rc && run_fwd) {
    for (int rep = 0; rep < 2 && !rc; rep++) {    // second repetition = warm timing
      cudaEventRecord(ev[0], st);
      int r = xplan ? lstm_tcx_forward(xplan, st, ln, fargs(1)) : lstm_tc_forward(plan, st, ln, fargs(1));
      cudaEventRecord(ev[1], st);
      if (r != 0) { say(r < 0 ? "tensor-core forward: not applicable" : (xplan ? lstm_tcx_error(xplan) : lstm_tc_error(plan))); rc = 3; break; }
      cudaError_t e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) { char b2[200]; snprintf(b2, sizeof b2, "lstm_tc_fwd failed: %s", cudaGetErrorString(e)); say(b2); rc = 4; break; }
      cudaEventElapsedTime(&ms[0], ev[0], ev[1]);
      if (const long long* c = xplan ? nullptr : lstm_tc_debug_counters(plan, 0))
        if (rep == 1)
          fprintf(stderr, "selftest_lstm fwd cluster=%d no=%d B=%d Tmax=%d %.3f ms | producer: flag %lld ring %lld issue %lld other %lld | mma: full %lld issue %lld "
                  "other %lld | epilogue: accwait %lld tmemld %lld math %lld publish %lld stash %lld xp %lld (cycles, CTA 0)\n",
                  plan->last_cluster, no, B, tmax, ms[0], c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[8], c[9], c[10], c[11], c[12], c[13]);
      if (rep == 1 && !xplan && plan->dbg) {
        const long long* c = plan->dbg_host[0];
        const char* names[] = {"flagwait", "ringwait", "tmaissue", "p.other", "fullwait", "mmaissue", "m.other", "", "accwait", "tmemld", "math", "publish", "stash", "xp"};
        for (int i : {0, 4, 5, 8, 10, 11, 12, 13}) spread(names[i], c, plan->last_ctas[0], i, plan->last_gx[0]);
      }
      if (const long long* c = xplan ? nullptr : plan->dbg_host[0])
        if (rep == 1 && c[16])
          fprintf(stderr, "selftest_lstm fwd timeline of step %d -> %d (cycles after the accumulator of step %d was ready): math done %lld, published %lld | "
                  "producer: flag seen %lld, first TMA out %lld, last TMA out %lld | mma: first tile in %lld, last tile in %lld, last MMA issued %lld | "
                  "next accumulator ready %lld\n", kDbgStep, kDbgStep + 1, kDbgStep, c[17] - c[16], c[18] - c[16], c[19] - c[16], c[20] - c[16],
                  c[21] - c[16], c[22] - c[16], c[23] - c[16], c[24] - c[16], c[25] - c[16]);
    }
  }
  if (!rc && run_bwd) {
    for (int rep = 0; rep < 2 && !rc; rep++) {
      cudaEventRecord(ev[0], st);
      int r = xplan ? lstm_tcx_backward(xplan, st, ln, bargs(1)) : lstm_tc_backward(plan, st, ln, bargs(1));
      cudaEventRecord(ev[1], st);
      if (r != 0) { say(r < 0 ? "tensor-core backward: not applicable" : (xplan ? lstm_tcx_error(xplan) : lstm_tc_error(plan))); rc = 5; break; }
      cudaError_t e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) { char b2[200]; snprintf(b2, sizeof b2, "lstm_tc_bwd failed: %s", cudaGetErrorString(e)); say(b2); rc = 6; break; }
      cudaEventElapsedTime(&ms[1], ev[0], ev[1]);
      if (const long long* c = xplan ? nullptr : lstm_tc_debug_counters(plan, 1))
        if (rep == 1)
          fprintf(stderr, "selftest_lstm bwd no=%d B=%d Tmax=%d %.3f ms | mma: deltawait %lld bufwait %lld issue %lld other %lld | epilogue: loads %lld "
                  "flagwait %lld reduce %lld pointwise %lld drain %lld publish %lld (cycles, CTA 0)\n",
                  no, B, tmax, ms[1], c[0], c[1], c[2], c[3], c[8], c[9], c[10], c[11], c[12], c[13]);
      if (rep == 1 && !xplan && plan->dbg) {
        const long long* c = plan->dbg_host[1];
        const char* names[] = {"deltawait", "bufwait", "mmaissue", "m.other", "", "", "", "", "loads", "flagwait", "reduce", "pointwise", "drain", "publish"};
        for (int i : {0, 2, 8, 9, 10, 11, 12, 13}) spread(names[i], c, plan->last_ctas[1], i, plan->last_gx[1]);
      }
    }
  }
  if (rc == 0 && xplan)
    if (const long long* t = lstm_tcx_debug(xplan)) {
      if (run_fwd && t[0])
        fprintf(stderr, "selftest_lstm_x fwd no=%d timeline of one step (cycles after the accumulator was ready): tmem read %lld, packed %lld, staged + epilogue barrier %lld, "
                "copy issued %lld, stash done %lld | next step: own chunk in + MMAs issued %lld, last but one %lld, last MMA issued %lld\n", no, t[1] - t[0], t[2] - t[0],
                t[3] - t[0], t[10] - t[0], t[4] - t[0], t[8] - t[0], t[12] - t[0], t[9] - t[0]);
      if (run_bwd && t[32])
        fprintf(stderr, "selftest_lstm_x bwd no=%d timeline of one step (cycles after the partial sums arrived): deltas in smem %lld, accumulator ready %lld, staged + epilogue "
                "barrier %lld, copy issued %lld | mma warp: MMAs issued %lld\n", no, t[33] - t[32], t[34] - t[32], t[35] - t[32], t[42] - t[32], t[41] - t[32]);
    }
  if ((rc == 0 || rc >= 5) && !run_fwd) { out[0] = out[1] = out[2] = out[3] = 0.f; }
  if ((rc == 0 || rc >= 5) && run_fwd) {
    auto fetch = [&](DevBuf& b, size_t n) { std::vector<float> h(n); cudaMemcpy(h.data(), b.p, n * 4, cudaMemcpyDeviceToHost); return h; };
    out[0] = max_abs_diff("gates", fetch(dG[0], 2 * n4), fetch(dG[1], 2 * n4));
    out[1] = max_abs_diff("cell", fetch(dC[0], 2 * n1), fetch(dC[1], 2 * n1));
    out[2] = max_abs_diff("h", fetch(dH[0], n2), fetch(dH[1], n2));
    out[3] = max_abs_diff("hprev", fetch(dHp[0], 2 * n1), fetch(dHp[1], 2 * n1));
  }
  if (rc == 0) {
    auto fetch = [&](DevBuf& b, size_t n) { std::vector<float> h(n); cudaMemcpy(h.data(), b.p, n * 4, cudaMemcpyDeviceToHost); return h; };
    if (!run_bwd) out[4] = 0.f;
    else {
      float ref = 0.f;
      const float dd = max_abs_diff("deltas", fetch(dDG[0], 2 * n4), fetch(dDG[1], 2 * n4), &ref);
      out[4] = dd / std::max(ref, 1e-30f);
    }
  }
  out[5] = ms[0]; out[6] = ms[1]; out[7] = ms[2]; out[8] = ms[3];
  for (auto& e : ev) cudaEventDestroy(e);
  cudaStreamDestroy(st);
  lstm_tc_destroy(plan);
  lstm_tcx_destroy(xplan);
  return rc;
}

}  // namespace cb200
