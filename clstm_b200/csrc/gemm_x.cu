// gemm_x.cu -- the dense products of the path (see gemm.cu for the list) as a PERSISTENT, warp-specialised tcgen05 GEMM fed by TMA,
// for the large-batch configurations where the products are real GEMMs (BASELINE configs 3-5):
//
//   C[M x N] (fp32) = A[M x K] * B[N x K]^T (+ bias)        both operands K-major
//
// fp32-grade accuracy comes from fp16 hi/lo operand planes (tc_common.cuh::split_f16: x*s = hi + lo) and three kind::f16 MMAs per
// 16 k (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM): the same ~2^-21 relative error as the 3xTF32 kernels of gemm_tc.cu at
// twice the MMA rate and K = 16 per instruction.  The planes are written by one conversion pass per operand
// (split_rows_kernel; split_transpose_kernel for the derivative products, whose reduction runs over the columns of the batch:
// the pass transposes while it converts, so EVERY product sees plain K-major SWIZZLE_128B tiles) -- 4 B read + 4 B written per
// element, against operands that are re-read once per output tile.
// Kernel: 6 warps.  Warp 0 = TMA producer (4 loads per 64-wide k block: A hi/lo [128 x 64], B hi/lo [BN x 64], ring of 2-3
// stages, mbarrier expect_tx), warp 1 = MMA issuer (12 tcgen05.mma per k block from one elected thread, tcgen05.commit frees the
// stage), warps 2-5 = epilogue (tcgen05.ld 32 columns at a time, scale, bias from a shared-memory tile, 128-bit row stores).  Two TMEM
// accumulators of 256 columns: the epilogue of tile i overlaps the main loop of tile i+1.  Persistent: CTA b walks tiles
// b, b + grid, ... (n fastest, so neighbouring CTAs share the A tile in L2).  Split-K (derivative products) writes fp32 partials
// to a workspace that gemm_tc.cu's fixed-order reduce-scatter kernel folds into the derivative blocks.
// Replaces for those sizes: forward_lin1 / backward_lin1 / softmax products hoisted over all columns
// (/root/reference/clstm_compute.cc:275-304, 331-354).
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "tc_common.cuh"

namespace cb200 {
namespace {
using namespace tc;

constexpr int GX_BM = 128, GX_BK = 64, GX_THREADS = 192, GX_MAXST = 4;

struct GxArgs {
  int M, N, nkb;               // nkb: 64-wide k blocks (the planes are zero padded to a multiple of 64)
  int BN, mtiles, ntiles, splits, kb_per_split, stages;
  float* C; long long ldc;     // direct output (splits == 1 and ws == nullptr)
  const float* bias;
  float* ws;                   // partials [splits][M][N]
  float out_scale;             // 1 / (scale_a * scale_b)
};

// ---------------------------------------------------------------------------------------------------------------- conversion
// src [R][C] fp32 (row pitch ld) -> hi / lo fp16 planes [R][Cp], Cp = C rounded up to 64, zero padded; values scaled by `scale`
__global__ void split_rows_kernel(const float* __restrict__ src, long long ld, int R, int C, int Cp, float scale, __half* __restrict__ hi,
                                  __half* __restrict__ lo) {
  const int c8n = Cp >> 3;
  const size_t total = (size_t)R * c8n;
  const bool vec = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / c8n), c0 = (int)(i % c8n) * 8;
    float v[8];
    const float* p = src + (size_t)r * ld + c0;
    if (vec && c0 + 7 < C) {
      const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; e++) v[e] = (c0 + e < C) ? p[e] : 0.f;
    }
    unsigned h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
      unsigned short h0, l0, h1, l1;
      split_f16(v[2 * e] * scale, h0, l0);
      split_f16(v[2 * e + 1] * scale, h1, l1);
      h[e] = pack_h2(h0, h1); l[e] = pack_h2(l0, l1);
    }
    *reinterpret_cast<uint4*>(hi + (size_t)r * Cp + c0) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo + (size_t)r * Cp + c0) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}
// src [R][C] fp32 -> TRANSPOSED hi / lo planes: out row (row0 + c), column r; row pitch Rp = R rounded up to 64, zero padded.
// 64 x 64 tiles through shared memory: coalesced 256-byte reads along c, 32-byte writes along r.
__global__ void __launch_bounds__(256) split_transpose_kernel(const float* __restrict__ src, long long ld, int R, int C, int row0, int Rp,
                                                              float scale, __half* __restrict__ hi, __half* __restrict__ lo) {
  __shared__ float tile[64][65];
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
#pragma unroll 4
  for (int rr = ty; rr < 64; rr += 4) {
    const int r = r0 + rr, c = c0 + tx;
    tile[rr][tx] = (r < R && c < C) ? src[(size_t)r * ld + c] * scale : 0.f;
  }
  __syncthreads();
  const int c = threadIdx.x >> 2, rg = threadIdx.x & 3;
  if (c0 + c < C) {
    unsigned h[8], l[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
      unsigned short h0, l0, h1, l1;
      split_f16(tile[rg * 16 + 2 * e][c], h0, l0);
      split_f16(tile[rg * 16 + 2 * e + 1][c], h1, l1);
      h[e] = pack_h2(h0, h1); l[e] = pack_h2(l0, l1);
    }
    const size_t o = (size_t)(row0 + c0 + c) * Rp + r0 + rg * 16;
    *reinterpret_cast<uint4*>(hi + o) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(hi + o + 8) = make_uint4(h[4], h[5], h[6], h[7]);
    *reinterpret_cast<uint4*>(lo + o) = make_uint4(l[0], l[1], l[2], l[3]);
    *reinterpret_cast<uint4*>(lo + o + 8) = make_uint4(l[4], l[5], l[6], l[7]);
  }
}
// one row of the transposed planes = the constant `value` for the first R columns (the all-ones column of the derivative product)
__global__ void fill_row_kernel(__half* __restrict__ hi, __half* __restrict__ lo, int R, int Rp, float value) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < Rp; r += gridDim.x * blockDim.x) {
    unsigned short h = 0, l = 0;
    if (r < R) split_f16(value, h, l);
    reinterpret_cast<unsigned short*>(hi)[r] = h;
    reinterpret_cast<unsigned short*>(lo)[r] = l;
  }
}

// ---------------------------------------------------------------------------------------------------------------- the GEMM
__global__ void __launch_bounds__(GX_THREADS, 1)
gemm_x_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl, const __grid_constant__ CUtensorMap tmBh,
              const __grid_constant__ CUtensorMap tmBl, GxArgs g) {
  extern __shared__ __align__(1024) unsigned char gx_smem[];
  __shared__ __align__(8) unsigned long long bars[2 * GX_MAXST + 4];   // full[S], empty[S], tfull[2], tempty[2]
  __shared__ unsigned tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned smem0 = (smem_u32(gx_smem) + 1023u) & ~1023u;
  const unsigned a_bytes = GX_BM * 128u, b_bytes = (unsigned)g.BN * 128u;
  const unsigned stage_bytes = 2u * a_bytes + 2u * b_bytes;
  const unsigned epi0 = smem0 + (unsigned)g.stages * stage_bytes;      // per-warp bias tiles of the epilogue (4 x 1 KB)
  const unsigned bar0 = smem_u32(&bars[0]);
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (GX_MAXST + s); };
  auto tfull = [&](int a) { return bar0 + 8u * (2 * GX_MAXST + a); };
  auto tempty = [&](int a) { return bar0 + 8u * (2 * GX_MAXST + 2 + a); };
  if (tid == 0) {
    for (int s = 0; s < GX_MAXST; s++) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int a = 0; a < 2; a++) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 4); }
    mbar_init_fence();
    tma_prefetch_desc(&tmAh); tma_prefetch_desc(&tmAl); tma_prefetch_desc(&tmBh); tma_prefetch_desc(&tmBl);
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_d = tmem_base_s;
  const int ntotal = g.mtiles * g.ntiles * g.splits;

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------------ TMA producer
    int st = 0;
    unsigned eph = 0;                      // bit s: parity to wait for on empty[s]; a fresh barrier passes a wait on parity 1
    for (int t = blockIdx.x; t < ntotal; t += gridDim.x) {
      const int nt = t % g.ntiles, mt = (t / g.ntiles) % g.mtiles, z = t / (g.ntiles * g.mtiles);
      const int kb0 = z * g.kb_per_split, kb1 = min(g.nkb, kb0 + g.kb_per_split);
      for (int kb = kb0; kb < kb1; kb++) {
        mbar_wait(empty(st), ((eph >> st) & 1u) ^ 1u);
        eph ^= 1u << st;
        if (elect_one()) {
          const unsigned sa = smem0 + (unsigned)st * stage_bytes;
          mbar_expect_tx(full(st), stage_bytes);
          tma_load_2d(sa, &tmAh, kb * GX_BK, mt * GX_BM, full(st));
          tma_load_2d(sa + a_bytes, &tmAl, kb * GX_BK, mt * GX_BM, full(st));
          tma_load_2d(sa + 2u * a_bytes, &tmBh, kb * GX_BK, nt * g.BN, full(st));
          tma_load_2d(sa + 2u * a_bytes + b_bytes, &tmBl, kb * GX_BK, nt * g.BN, full(st));
        }
        __syncwarp();
        if (++st == g.stages) st = 0;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------------ MMA issuer
    const unsigned idesc = make_idesc_f16(GX_BM, g.BN);
    const unsigned long long dbase = make_desc(0);
    auto desc_of = [&](unsigned addr) { return dbase | (unsigned long long)((addr & 0x3FFFF) >> 4); };
    int st = 0, acc = 0;
    unsigned fph = 0, teph = 0;
    for (int t = blockIdx.x; t < ntotal; t += gridDim.x) {
      const int z = t / (g.ntiles * g.mtiles);
      const int kb0 = z * g.kb_per_split, kb1 = min(g.nkb, kb0 + g.kb_per_split);
      mbar_wait(tempty(acc), ((teph >> acc) & 1u) ^ 1u);          // the epilogue has drained this accumulator
      teph ^= 1u << acc;
      tc_fence_after();
      const unsigned d = tmem_d + 256u * (unsigned)acc;
      for (int kb = kb0; kb < kb1; kb++) {
        mbar_wait(full(st), (fph >> st) & 1u);
        fph ^= 1u << st;
        tc_fence_after();
        if (elect_one()) {
          const unsigned sa = smem0 + (unsigned)st * stage_bytes;
          const unsigned long long ah = desc_of(sa), al = desc_of(sa + a_bytes), bh = desc_of(sa + 2u * a_bytes), bl = desc_of(sa + 2u * a_bytes + b_bytes);
#pragma unroll
          for (int ks = 0; ks < 4; ks++) {
            mma_f16(d, al + 2 * ks, bh + 2 * ks, idesc, (kb > kb0 || ks > 0) ? 1u : 0u);   // small terms first
            mma_f16(d, ah + 2 * ks, bl + 2 * ks, idesc, 1u);
            mma_f16(d, ah + 2 * ks, bh + 2 * ks, idesc, 1u);
          }
          mma_commit(empty(st));                                   // frees the stage once these MMAs have read it
          if (kb == kb1 - 1) mma_commit(tfull(acc));
        }
        __syncwarp();
        if (++st == g.stages) st = 0;
      }
      acc ^= 1;
    }
  } else {
    // ------------------------------------------------------------------------------------------------ epilogue warps
    const int lq = warp & 3;                                       // TMEM lane quadrant this warp may read
    int acc = 0;
    unsigned tfph = 0;
    for (int t = blockIdx.x; t < ntotal; t += gridDim.x) {
      const int nt = t % g.ntiles, mt = (t / g.ntiles) % g.mtiles, z = t / (g.ntiles * g.mtiles);
      const int col0 = nt * g.BN;
      mbar_wait(tfull(acc), (tfph >> acc) & 1u);
      tfph ^= 1u << acc;
      tc_fence_after();
      const unsigned taddr = tmem_d + ((unsigned)(32 * lq) << 16) + 256u * (unsigned)acc;
      // every thread owns one ROW of the tile (TMEM lane = row) and writes 32 consecutive floats of it per round as 128-bit
      // stores.  The bias tile is staged per warp in shared memory first: bias loads between the stores (possible aliasing keeps
      // the compiler from hoisting them) cost ~100 cycles each, 25 000 per tile -- the input projection ran 1.6x slower than on
      // the 3xTF32 kernel until they went.  (A transposed epilogue -- full 128-byte lines per store -- measured 2x slower.)
      const unsigned bias_s = epi0 + (unsigned)(warp - 2) * 1024u;
      const bool has_bias = g.bias && !g.ws;
      if (has_bias) {
        __syncwarp();
        for (int i = lane; i < g.BN; i += 32) {
          const float bv = (col0 + i < g.N) ? g.bias[col0 + i] : 0.f;
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_s + 4u * i), "f"(bv) : "memory");
        }
        __syncwarp();
      }
      const int row = mt * GX_BM + 32 * lq + lane;
      float* __restrict__ dst = g.ws ? g.ws + ((size_t)z * g.M + row) * g.N : g.C + (size_t)row * g.ldc;
      const bool vec = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
      for (int cb = 0; cb < g.BN; cb += 32) {
        float v[32];
        tmem_ld<32>(taddr + (unsigned)cb, v);
        if (row < g.M) {
#pragma unroll
          for (int q = 0; q < 8; q++) {
            const int col = col0 + cb + 4 * q;
            float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_bias && cb + 4 * q < g.BN)
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bb.x), "=f"(bb.y), "=f"(bb.z), "=f"(bb.w) : "r"(bias_s + 4u * (cb + 4 * q)));
            const float o0 = fmaf(v[4 * q], g.out_scale, bb.x), o1 = fmaf(v[4 * q + 1], g.out_scale, bb.y);
            const float o2 = fmaf(v[4 * q + 2], g.out_scale, bb.z), o3 = fmaf(v[4 * q + 3], g.out_scale, bb.w);
            if (cb + 4 * q < g.BN) {
              if (vec && col + 3 < g.N) *reinterpret_cast<float4*>(dst + col) = make_float4(o0, o1, o2, o3);
              else {
                if (col < g.N) dst[col] = o0;
                if (col + 1 < g.N) dst[col + 1] = o1;
                if (col + 2 < g.N) dst[col + 2] = o2;
                if (col + 3 < g.N) dst[col + 3] = o3;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(acc));
      acc ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(512) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_gx = nullptr;
int load_encode_gx() {
  if (g_encode_gx) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) return 1;
  g_encode_gx = (EncodeTiledFn)fn;
  return 0;
}
// planes [rows][pitch] halfs, K extent kp (multiple of 64): box = 64 k x box_rows rows, SWIZZLE_128B, rows beyond `rows` read as zero
int make_map_gx(CUtensorMap* m, const void* base, size_t rows, size_t kp, size_t pitch, int box_rows) {
  const cuuint64_t gdim[2] = {(cuuint64_t)kp, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)pitch * 2};
  const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  return g_encode_gx(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? 0 : 1;
}
constexpr size_t kGxSmemMax = 226 * 1024;   // dynamic part; the 227 KB limit of a CTA includes the kernel's static shared memory
size_t gx_smem_bytes(int BN, int stages) { return (size_t)stages * (2 * GX_BM * 128 + 2 * (size_t)BN * 128) + 4 * 1024 + 1024; }
}  // namespace

struct GxPlan {
  int num_sms = 148;
  __half* a[2] = {nullptr, nullptr};   // hi / lo planes of the A operand
  __half* b[2] = {nullptr, nullptr};
  size_t cap_a = 0, cap_b = 0;         // halfs per plane
  char err[200] = {0};
};

GxPlan* gemm_x_create(int num_sms) {
  if (load_encode_gx() != 0) return nullptr;
  if (cudaFuncSetAttribute(gemm_x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGxSmemMax) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  auto* p = new GxPlan;
  p->num_sms = num_sms;
  return p;
}
void gemm_x_destroy(GxPlan* p) {
  if (!p) return;
  cudaFree(p->a[0]); cudaFree(p->a[1]); cudaFree(p->b[0]); cudaFree(p->b[1]);
  delete p;
}
const char* gemm_x_error(const GxPlan* p) { return p ? p->err : "no plan"; }

namespace {
int gx_reserve(GxPlan* p, cudaStream_t st, size_t need_a, size_t need_b) {
  for (int w = 0; w < 2; w++) {
    __half** buf = w ? p->b : p->a;
    size_t& cap = w ? p->cap_b : p->cap_a;
    const size_t need = w ? need_b : need_a;
    if (need <= cap) continue;
    cudaStreamSynchronize(st);                      // (kernels in flight still read the old planes)
    cudaFree(buf[0]); cudaFree(buf[1]);
    buf[0] = buf[1] = nullptr; cap = 0;
    const size_t want = need + need / 8;
    if (cudaMalloc((void**)&buf[0], want * 2) != cudaSuccess || cudaMalloc((void**)&buf[1], want * 2) != cudaSuccess) {
      cudaGetLastError();
      snprintf(p->err, sizeof p->err, "gemm_x: out of memory for %zu operand halfs", want);
      return 1;
    }
    cap = want;
  }
  return 0;
}
int gx_blocks(size_t work, int num_sms) { return (int)std::min<size_t>((work + 255) / 256, (size_t)num_sms * 16); }
int pick_bn(int N) {                                // as few N tiles as possible, tile a multiple of 16, <= 256
  const int nt = (N + 255) / 256;
  int bn = (((N + nt - 1) / nt) + 15) & ~15;
  return std::max(bn, 16);
}
// launches the GEMM on planes already in p->a / p->b.  ws != nullptr: split-K partials [splits][M][N]; returns the split count
int gx_launch(GxPlan* p, cudaStream_t st, int M, int N, int Kp, size_t pitch_a, size_t pitch_b, float* C, long long ldc, const float* bias,
              float out_scale, float* ws, size_t ws_floats, int* splits_out) {
  GxArgs g{};
  g.M = M; g.N = N; g.nkb = Kp / GX_BK;
  g.BN = pick_bn(N);
  g.mtiles = (M + GX_BM - 1) / GX_BM; g.ntiles = (N + g.BN - 1) / g.BN;
  int splits = 1;
  if (ws) {                                          // about one wave, at least 8 k blocks per CTA, fits the workspace
    splits = std::max(1, p->num_sms / (g.mtiles * g.ntiles));
    splits = std::min(splits, std::max(1, g.nkb / 8));
    splits = std::min(splits, 64);
    while (splits > 1 && (size_t)splits * M * N > ws_floats) splits--;
    if ((size_t)splits * M * N > ws_floats) { snprintf(p->err, sizeof p->err, "gemm_x: split-K workspace too small"); return 1; }
  }
  g.kb_per_split = (g.nkb + splits - 1) / splits;
  splits = (g.nkb + g.kb_per_split - 1) / g.kb_per_split;
  g.splits = splits;
  g.stages = (gx_smem_bytes(g.BN, 4) <= kGxSmemMax) ? 4 : (gx_smem_bytes(g.BN, 3) <= kGxSmemMax ? 3 : 2);
  g.C = C; g.ldc = ldc; g.bias = bias; g.ws = ws; g.out_scale = out_scale;
  CUtensorMap mAh, mAl, mBh, mBl;
  if (make_map_gx(&mAh, p->a[0], M, Kp, pitch_a, GX_BM) || make_map_gx(&mAl, p->a[1], M, Kp, pitch_a, GX_BM) ||
      make_map_gx(&mBh, p->b[0], N, Kp, pitch_b, g.BN) || make_map_gx(&mBl, p->b[1], N, Kp, pitch_b, g.BN)) {
    snprintf(p->err, sizeof p->err, "gemm_x: cuTensorMapEncodeTiled failed (M %d N %d Kp %d)", M, N, Kp);
    return 1;
  }
  const int ntotal = g.mtiles * g.ntiles * splits;
  const int grid = std::min(ntotal, p->num_sms);
  gemm_x_kernel<<<grid, GX_THREADS, gx_smem_bytes(g.BN, g.stages), st>>>(mAh, mAl, mBh, mBl, g);
  if (splits_out) *splits_out = splits;
  return 0;
}
}  // namespace

// C = A[M x K] * B[N x K]^T + bias, fp32 row-major operands (lda, ldb).  Returns kernels launched, < 0 on error (gemm_x_error).
int gemm_x_nt(GxPlan* p, cudaStream_t st, int M, int N, int K, const float* A, long long lda, const float* B, long long ldb, float* C,
              long long ldc, const float* bias, float scale_a, float scale_b) {
  const int Kp = (K + 63) & ~63;
  if (gx_reserve(p, st, (size_t)M * Kp, (size_t)N * Kp)) return -1;
  split_rows_kernel<<<gx_blocks((size_t)M * (Kp / 8), p->num_sms), 256, 0, st>>>(A, lda, M, K, Kp, scale_a, p->a[0], p->a[1]);
  split_rows_kernel<<<gx_blocks((size_t)N * (Kp / 8), p->num_sms), 256, 0, st>>>(B, ldb, N, K, Kp, scale_b, p->b[0], p->b[1]);
  if (gx_launch(p, st, M, N, Kp, Kp, Kp, C, ldc, bias, 1.f / (scale_a * scale_b), nullptr, 0, nullptr)) return -1;
  return 3;
}

// Derivative product reduced over the Kc columns of the batch: partials of A^T [B0 | B1 | 1] (A [Kc x M], B blocks [Kc x n0 / n1],
// all row-major) into ws as [splits][M][n0 + n1 + 1]; the caller folds them with tc_reduce_scatter.  Returns kernels launched.
int gemm_x_tn(GxPlan* p, cudaStream_t st, int M, int Kc, const float* A, long long lda, const float* B0, int n0, const float* B1, int n1,
              float scale_a, float scale_b, float* ws, size_t ws_floats, int* splits_out) {
  const int Kp = (Kc + 63) & ~63;
  const int N = n0 + n1 + 1;
  if (gx_reserve(p, st, (size_t)M * Kp, (size_t)N * Kp)) return -1;
  int launches = 0;
  split_transpose_kernel<<<dim3(Kp / 64, (M + 63) / 64), 256, 0, st>>>(A, lda, Kc, M, 0, Kp, scale_a, p->a[0], p->a[1]); launches++;
  split_transpose_kernel<<<dim3(Kp / 64, (n0 + 63) / 64), 256, 0, st>>>(B0, n0, Kc, n0, 0, Kp, scale_b, p->b[0], p->b[1]); launches++;
  if (B1) { split_transpose_kernel<<<dim3(Kp / 64, (n1 + 63) / 64), 256, 0, st>>>(B1, n1, Kc, n1, n0, Kp, scale_b, p->b[0], p->b[1]); launches++; }
  fill_row_kernel<<<gx_blocks((size_t)Kp, p->num_sms), 256, 0, st>>>(p->b[0] + (size_t)(n0 + n1) * Kp, p->b[1] + (size_t)(n0 + n1) * Kp, Kc, Kp, scale_b);
  launches++;
  if (gx_launch(p, st, M, N, Kp, Kp, Kp, nullptr, 0, nullptr, 1.f / (scale_a * scale_b), ws, ws_floats, splits_out)) return -1;
  return launches + 1;
}

}  // namespace cb200
