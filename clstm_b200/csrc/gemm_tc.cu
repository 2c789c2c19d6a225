// gemm_tc.cu -- tcgen05 (5th-gen tensor core) GEMM with fp32-grade accuracy for the batched dense products of the
// path (see gemm.cu for the list).  fp32 operands are split on the fly into TF32 hi + TF32 lo parts and each
// K=8 step issues three MMAs (hi*hi + hi*lo + lo*hi, "3xTF32"): ~2^-21 relative error per product, inside the 1e-4
// parity bar that plain TF32 (2^-11) would break.
//
//   D[128 x BN] (fp32, TMEM) += A[128 x 32] * B[BN x 32]^T     per k-block, UMMA M=128, N=BN, K=8 (kind::tf32)
//
// Operand tiles live in shared memory in the canonical K-major SWIZZLE_128B layout (rows of 128 bytes = 32 tf32,
// 16-byte chunk c of row r stored at chunk c ^ (r & 7), 8-row groups 1024 bytes apart), written by the CTA's own
// threads because the hi/lo split has to happen between global memory and shared memory anyway.  Either operand may
// be K-contiguous in global memory (activations x weights^T products) or MN-contiguous (the weight-derivative
// products that reduce over all columns); the loader transposes into the same K-major tile, so one descriptor
// format serves every product.  Two smem stages; MMA completion is tracked with tcgen05.commit -> mbarrier.
#include "kernels.h"

namespace cb200 {
namespace {

constexpr int TCT = 256;                 // threads per CTA
constexpr int BM = 128, BK = 32;         // UMMA M, k-block (32 tf32 = one 128-byte swizzle row)
constexpr int TMEM_COLS = 256;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned done = 0;
  for (unsigned spin = 0; !done; spin++) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 26)) __trap();     // a lost arrival must surface as an error, never as a hang
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (SBO), version 1 (sm_100)
__device__ __forceinline__ unsigned long long make_desc(unsigned saddr) {
  unsigned long long d = 0;
  d |= (unsigned long long)((saddr & 0x3FFFF) >> 4);          // start address, 16-byte units
  d |= (unsigned long long)(1024 >> 4) << 32;                 // stride byte offset
  d |= 1ull << 46;                                            // descriptor version
  d |= 2ull << 61;                                            // SWIZZLE_128B
  return d;
}
// instruction descriptor, kind::tf32: D=f32, A=B=tf32, both K-major, N>>3 at bit 17, M>>4 at bit 24
__device__ __forceinline__ unsigned make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(BM >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc,
                                         unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(unsigned bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(unsigned taddr, float* v) {
  unsigned r0, r1, r2, r3, r4, r5, r6, r7;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
  v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}

// fp32 -> (tf32 hi, tf32 lo)
__device__ __forceinline__ void split_tf32(float x, unsigned& hi, unsigned& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float rem = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(rem));
}
__device__ __forceinline__ void sts16(unsigned addr, unsigned a, unsigned b, unsigned c, unsigned d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts4(unsigned addr, unsigned a) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}

// Tile of R rows x 32 k from a K-contiguous source: element (r, k) = p[(r0 + r) * ld + k0 + k]
__device__ __forceinline__ void load_kcontig(unsigned hi_base, unsigned lo_base, const float* __restrict__ p,
                                             long long ld, int R, int rows_valid, int k_valid, bool vec) {
  const int tid = threadIdx.x;
  const int chunk = tid & 7;
  for (int row = tid >> 3; row < R; row += TCT / 8) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (row < rows_valid) {
      const float* src = p + row * ld + chunk * 4;
      if (vec && chunk * 4 + 3 < k_valid) {
        const float4 f = *reinterpret_cast<const float4*>(src);
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; e++)
          if (chunk * 4 + e < k_valid) v[e] = src[e];
      }
    }
    unsigned h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; e++) split_tf32(v[e], h[e], l[e]);
    const unsigned off = row * 128 + ((chunk ^ (row & 7)) << 4);
    sts16(hi_base + off, h[0], h[1], h[2], h[3]);
    sts16(lo_base + off, l[0], l[1], l[2], l[3]);
  }
}

// Tile of R rows x 32 k from MN-contiguous sources: element (r, k) = seg.p[(k0 + k) * seg.ld + (c - seg.c0)], where
// c = r0 + r is the global row/column index, looked up in up to 3 segments; c == ones_col yields 1.0.
__device__ __forceinline__ void load_mncontig(unsigned hi_base, unsigned lo_base, const TcSeg* seg, int nseg,
                                              int ones_col, int r0, int R, int c_total, int k0, int k_valid) {
  for (int idx = threadIdx.x; idx < R * BK; idx += TCT) {
    const int r = idx % R, k = idx / R;          // consecutive threads -> consecutive rows: coalesced global reads
    const int c = r0 + r;
    float v = 0.f;
    if (k < k_valid && c < c_total) {
      if (c == ones_col) v = 1.f;
      else {
        int cl = c;
#pragma unroll
        for (int s = 0; s < 3; s++) {
          if (s < nseg) {
            if (cl >= 0 && cl < seg[s].len) v = seg[s].p[(long long)(k0 + k) * seg[s].ld + cl];
            cl -= seg[s].len;
          }
        }
      }
    }
    unsigned h, l;
    split_tf32(v, h, l);
    const unsigned off = r * 128 + (((k >> 2) ^ (r & 7)) << 4) + ((k & 3) << 2);
    sts4(hi_base + off, h);
    sts4(lo_base + off, l);
  }
}

__global__ void __launch_bounds__(TCT, 1) gemm_tc_kernel(TcArgs g) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  __shared__ __align__(8) unsigned long long mbar_s[2];
  __shared__ unsigned tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = g.BN;
  const unsigned a_bytes = BM * 128, b_bytes = BN * 128;
  const unsigned stage_bytes = 2 * a_bytes + 2 * b_bytes;
  const unsigned smem0 = (smem_u32(tc_smem) + 1023u) & ~1023u;      // swizzle atoms need 1024-byte alignment
  const unsigned bar0 = smem_u32(&mbar_s[0]), bar1 = smem_u32(&mbar_s[1]);

  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_d = tmem_base_s;

  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  // k-block range of this CTA (split-K over blockIdx.z)
  const int nkb_total = g.nkb;
  const int kb0 = blockIdx.z * g.kb_per_split;
  const int kb1 = min(nkb_total, kb0 + g.kb_per_split);
  const unsigned idesc = make_idesc(BN);

  int it = 0;
  for (int kb = kb0; kb < kb1; kb++, it++) {
    const int s = it & 1;
    const unsigned bar = s ? bar1 : bar0;
    if (it >= 2) mbar_wait(bar, ((it >> 1) - 1) & 1);        // the MMAs that read this stage two iterations ago are done
    const unsigned a_hi = smem0 + s * stage_bytes, a_lo = a_hi + a_bytes;
    const unsigned b_hi = a_lo + a_bytes, b_lo = b_hi + b_bytes;
    // locate the k-block inside the K segments (each K-contiguous segment is padded to a multiple of 32)
    int seg = 0, kloc = kb * BK;
    if (g.a_mode == 0 || g.b_mode == 0) {
      while (seg + 1 < g.k_nseg && kloc >= ((g.k_len[seg] + BK - 1) / BK) * BK) {
        kloc -= ((g.k_len[seg] + BK - 1) / BK) * BK;
        seg++;
      }
    }
    const int k_valid = (g.a_mode == 0 || g.b_mode == 0) ? (g.k_len[seg] - kloc) : (g.k_len[0] - kb * BK);
    if (g.a_mode == 0)
      load_kcontig(a_hi, a_lo, g.a_k[seg].p + (long long)m0 * g.a_k[seg].ld + kloc, g.a_k[seg].ld, BM, g.M - m0, k_valid,
                   g.a_vec != 0);
    else
      load_mncontig(a_hi, a_lo, g.a_mn, 1, -1, m0, BM, g.M, kb * BK, k_valid);
    if (g.b_mode == 0)
      load_kcontig(b_hi, b_lo, g.b_k[seg].p + (long long)n0 * g.b_k[seg].ld + kloc, g.b_k[seg].ld, BN, g.N - n0, k_valid,
                   g.b_vec != 0);
    else
      load_mncontig(b_hi, b_lo, g.b_mn, g.b_nseg, g.b_ones, n0, BN, g.N, (g.a_mode == 0) ? kloc : kb * BK, k_valid);
    fence_proxy_async();                                     // generic-proxy smem writes -> visible to the tensor core
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < BK / 8; ks++) {
        const unsigned long long ah = make_desc(a_hi + ks * 32), al = make_desc(a_lo + ks * 32);
        const unsigned long long bh = make_desc(b_hi + ks * 32), bl = make_desc(b_lo + ks * 32);
        mma_tf32(tmem_d, al, bh, idesc, (it > 0 || ks > 0) ? 1u : 0u);   // small terms first
        mma_tf32(tmem_d, ah, bl, idesc, 1u);
        mma_tf32(tmem_d, ah, bh, idesc, 1u);
      }
      mma_commit(bar);
    }
  }
  // all MMAs retire in order: the last commit covers everything
  if (it > 0) {
    const int last = it - 1;
    mbar_wait((last & 1) ? bar1 : bar0, (last >> 1) & 1);
  }
  tc_fence_after();

  // ---- epilogue: TMEM -> registers -> global.  warp w reads lanes 32*(w&3).., column half (w>>2)
  const int lq = warp & 3, ch = warp >> 2;
  const int row = m0 + 32 * lq + lane;
  const int half = BN / 2;
  for (int c = 0; c < half; c += 8) {
    float v[8];
    if (it > 0) tmem_ld8(tmem_d + ((unsigned)(32 * lq) << 16) + (unsigned)(ch * half + c), v);
    else {
#pragma unroll
      for (int e = 0; e < 8; e++) v[e] = 0.f;
    }
    if (row < g.M) {
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const int col = n0 + ch * half + c + e;
        if (col < g.N) {
          if (g.ws) {
            g.ws[((size_t)blockIdx.z * g.M + row) * g.N + col] = v[e];
          } else {
            float o = v[e];
            if (g.bias) o += g.bias[col];
            float* dst = g.C + (long long)row * g.ldc + col;
            *dst = (g.beta != 0.f) ? fmaf(g.beta, *dst, o) : o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TMEM_COLS) : "memory");
  }
}

// out(seg)[m][c_local] = beta*out + sum_z ws[z][m][c]; column c of the partial tile is scattered into up to 3
// row-major destinations plus a vector for the bias column.  Fixed summation order => deterministic.
__global__ void tc_reduce_scatter_kernel(int M, int N, int splits, const float* __restrict__ ws, TcOut o, float beta) {
  const size_t total = (size_t)M * N;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; z++) s += ws[(size_t)z * total + i];
    const int m = (int)(i / N);
    int c = (int)(i % N);
    float* dst = nullptr;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      if (k < o.nseg) {
        if (c >= 0 && c < o.len[k]) dst = o.p[k] + (long long)m * o.ld[k] + c;
        c -= o.len[k];
      }
    }
    if (!dst && c == 0 && o.bias) dst = o.bias + m;
    if (dst) *dst = (beta != 0.f) ? fmaf(beta, *dst, s) : s;
  }
}

size_t tc_smem_bytes(int BN) { return (size_t)2 * (2 * BM * 128 + 2 * BN * 128) + 1024; }

}  // namespace

int gemm_tc_configure() {
  return (int)cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(256));
}

int gemm_tc(cudaStream_t st, TcArgs g, const TcOut* scatter, int num_sms) {
  if (g.M <= 0 || g.N <= 0) return 0;
  // N tile: multiple of 16, <= 256, as few tiles as possible
  const int ntiles = (g.N + 255) / 256;
  int BN = (((g.N + ntiles - 1) / ntiles) + 15) & ~15;
  if (BN < 16) BN = 16;
  g.BN = BN;
  const int mtiles = (g.M + BM - 1) / BM;
  int nkb = 0;
  if (g.a_mode == 0 || g.b_mode == 0)
    for (int s = 0; s < g.k_nseg; s++) nkb += (g.k_len[s] + BK - 1) / BK;
  else
    nkb = (g.k_len[0] + BK - 1) / BK;
  g.nkb = nkb;
  int splits = 1;
  if (scatter) {  // split-K until ~1 wave, at least 4 k-blocks per CTA
    splits = num_sms / (mtiles * ntiles);
    if (splits > nkb / 4) splits = nkb / 4;
    if (splits > 64) splits = 64;
    if (splits < 1) splits = 1;
    while (splits > 1 && (size_t)splits * g.M * g.N > g.ws_floats) splits--;
  }
  g.kb_per_split = (nkb + splits - 1) / splits;
  splits = (nkb + g.kb_per_split - 1) / g.kb_per_split;
  if (!scatter) g.ws = nullptr;
  dim3 grid(ntiles, mtiles, splits);
  gemm_tc_kernel<<<grid, TCT, tc_smem_bytes(BN), st>>>(g);
  if (scatter) {
    const size_t total = (size_t)g.M * g.N;
    size_t nb = (total + 255) / 256;
    if (nb > (size_t)num_sms * 8) nb = (size_t)num_sms * 8;
    tc_reduce_scatter_kernel<<<(int)nb, 256, 0, st>>>(g.M, g.N, splits, g.ws, *scatter, g.beta);
    return 2;
  }
  return 1;
}

}  // namespace cb200
