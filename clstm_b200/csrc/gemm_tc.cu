// gemm_tc.cu -- tcgen05 (5th-gen tensor core) GEMM with fp32-grade accuracy for the batched dense products of the
// path (see gemm.cu for the list).  fp32 operands are split on the fly into TF32 hi + TF32 lo parts and each
// K=8 step issues three MMAs (hi*hi + hi*lo + lo*hi, "3xTF32"): ~2^-21 relative error per product, inside the 1e-4
// parity bar that plain TF32 (2^-11) would break.
//
//   D[128 x BN] (fp32, TMEM) += A[128 x 32] * B[BN x 32]^T     per k-block, UMMA M=128, N=BN, K=8 (kind::tf32)
//
// Operand tiles live in shared memory in the canonical SWIZZLE_128B layouts, written by the CTA's own threads because
// the hi/lo split has to happen between global memory and shared memory anyway:
//   K-major tile: rows of 128 bytes = 32 k, chunk c of row r at c ^ (r & 7), 8-row groups 1024 bytes apart.
//   K-contiguous sources (activations x weights^T) move as 16-byte chunks; MN-contiguous sources (the derivative
//   products that reduce over all columns; W1 / Wx used untransposed) are transposed on the way in, lane = k, which
//   makes the scattered 4-byte stores bank-conflict free.
// Two smem stages; MMA completion is tracked with tcgen05.commit -> mbarrier.
#include <cstdlib>

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

// ---- K-major tile (mode 0): R rows x 32 k from a K-contiguous source, element (r, k) = p[(r0 + r) * ld + k0 + k].
// All global loads of the tile are issued before the first conversion (R <= 256 => at most 8 passes).
__device__ __forceinline__ void load_kmajor(unsigned hi_base, unsigned lo_base, const float* __restrict__ p,
                                            long long ld, int R, int rows_valid, int k_valid, bool vec) {
  const int tid = threadIdx.x;
  const int chunk = tid & 7, rsub = tid >> 3;
  float v[8][4];
#pragma unroll
  for (int ps = 0; ps < 8; ps++) {
    const int row = ps * 32 + rsub;
    v[ps][0] = v[ps][1] = v[ps][2] = v[ps][3] = 0.f;
    if (row < R && row < rows_valid) {
      const float* src = p + row * ld + chunk * 4;
      if (vec && chunk * 4 + 3 < k_valid) {
        const float4 f = *reinterpret_cast<const float4*>(src);
        v[ps][0] = f.x; v[ps][1] = f.y; v[ps][2] = f.z; v[ps][3] = f.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; e++)
          if (chunk * 4 + e < k_valid) v[ps][e] = src[e];
      }
    }
  }
#pragma unroll
  for (int ps = 0; ps < 8; ps++) {
    const int row = ps * 32 + rsub;
    if (row < R) {
      unsigned h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; e++) split_tf32(v[ps][e], h[e], l[e]);
      const unsigned off = row * 128 + ((chunk ^ (row & 7)) << 4);
      sts16(hi_base + off, h[0], h[1], h[2], h[3]);
      sts16(lo_base + off, l[0], l[1], l[2], l[3]);
    }
  }
}

// ---- MN-contiguous sources (mode 1): element (k, c) = seg.p[(k0 + k) * seg.ld + c_local], c = r0 + r looked up in up
// to 3 column blocks; c == ones_col yields 1.0.
__device__ __forceinline__ float mn_elem(const TcSeg* seg, int nseg, int ones_col, int c, int c_total, long long krow) {
  if (c >= c_total) return 0.f;
  if (c == ones_col) return 1.f;
  float v = 0.f;
  int cl = c;
#pragma unroll
  for (int s = 0; s < 3; s++) {
    if (s < nseg) {
      if (cl >= 0 && cl < seg[s].len) v = seg[s].p[krow * seg[s].ld + cl];
      cl -= seg[s].len;
    }
  }
  return v;
}
// ---- MN-contiguous source into a K-major tile (in-kernel transpose): a warp takes 4 consecutive mn x 32 k, lane = k.
// Each lane reads 16 bytes (4 mn) of its k row and scatters them to 4 tile rows; for a fixed component all 32 lanes hit
// 32 distinct banks (bank = 4*((k>>2)^(mn&7)) + (k&3)), so the transposing stores are conflict-free.
__device__ __forceinline__ void load_mn_to_kmajor(unsigned hi_base, unsigned lo_base, const TcSeg* seg, int nseg,
                                                  int ones_col, int r0, int R, int c_total, int k0, int k_valid,
                                                  bool vec) {
  const int warp = threadIdx.x >> 5, k = threadIdx.x & 31;
  const int nquad = (R + 3) >> 2;
  const long long krow = k0 + k;
  for (int q0 = warp; q0 < nquad; q0 += 4 * (TCT / 32)) {
    float v[4][4];
#pragma unroll
    for (int u = 0; u < 4; u++) {             // 4 quads in flight per thread
      const int q = q0 + u * (TCT / 32);
      v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0.f;
      if (q < nquad && k < k_valid) {
        const int c = r0 + q * 4;
        bool done = false;
        if (vec) {
          int cl = c;
#pragma unroll
          for (int s = 0; s < 3; s++) {
            if (s < nseg) {
              if (!done && cl >= 0 && cl + 3 < seg[s].len) {
                const float4 f = *reinterpret_cast<const float4*>(seg[s].p + krow * seg[s].ld + cl);
                v[u][0] = f.x; v[u][1] = f.y; v[u][2] = f.z; v[u][3] = f.w;
                done = true;
              }
              cl -= seg[s].len;
            }
          }
        }
        if (!done) {
#pragma unroll
          for (int e = 0; e < 4; e++) v[u][e] = mn_elem(seg, nseg, ones_col, c + e, c_total, krow);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int q = q0 + u * (TCT / 32);
      if (q < nquad) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int r = q * 4 + e;
          if (r < R) {
            unsigned h, l;
            split_tf32(v[u][e], h, l);
            const unsigned off = r * 128 + (((k >> 2) ^ (r & 7)) << 4) + ((k & 3) << 2);
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(hi_base + off), "r"(h) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(lo_base + off), "r"(l) : "memory");
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(TCT, 1) gemm_tc_kernel(TcArgs g) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  __shared__ __align__(8) unsigned long long mbar_s[2];
  __shared__ unsigned tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = g.BN;
  const unsigned a_bytes = BM * 128, b_bytes = (unsigned)((BN + 31) & ~31) * 128;
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
  const int kb0 = blockIdx.z * g.kb_per_split;
  const int kb1 = min(g.nkb, kb0 + g.kb_per_split);
  // instruction descriptor: bit 15 / 16 = A / B is MN-major
  // Both smem tiles are always K-major.  (kind::tf32 with MN-major descriptors -- idesc bits 15/16 -- was tried for
  // the MN-contiguous sources and returned all-zero accumulators on B200, for either LBO/SBO role assignment, so
  // those sources are transposed by the loader instead.)
  const unsigned idesc = make_idesc(BN);
  const bool kseg = (g.a_mode == 0 || g.b_mode == 0);

  int it = 0;
  for (int kb = kb0; kb < kb1; kb++, it++) {
    const int s = it & 1;
    const unsigned bar = s ? bar1 : bar0;
    if (it >= 2) mbar_wait(bar, ((it >> 1) - 1) & 1);        // the MMAs that read this stage two iterations ago are done
    const unsigned a_hi = smem0 + s * stage_bytes, a_lo = a_hi + a_bytes;
    const unsigned b_hi = a_lo + a_bytes, b_lo = b_hi + b_bytes;
    // locate the k-block inside the K segments (each K-contiguous segment is padded to a multiple of 32)
    int seg = 0, kloc = kb * BK;
    if (kseg) {
      while (seg + 1 < g.k_nseg && kloc >= ((g.k_len[seg] + BK - 1) / BK) * BK) {
        kloc -= ((g.k_len[seg] + BK - 1) / BK) * BK;
        seg++;
      }
    }
    const int k_valid = g.k_len[seg] - kloc;
    if (g.a_mode == 0)
      load_kmajor(a_hi, a_lo, g.a_k[seg].p + (long long)m0 * g.a_k[seg].ld + kloc, g.a_k[seg].ld, BM, g.M - m0, k_valid,
                  g.a_vec != 0);
    else
      load_mn_to_kmajor(a_hi, a_lo, g.a_mn, 1, -1, m0, BM, g.M, kloc, k_valid, g.a_vec != 0);
    if (g.b_mode == 0)
      load_kmajor(b_hi, b_lo, g.b_k[seg].p + (long long)n0 * g.b_k[seg].ld + kloc, g.b_k[seg].ld, BN, g.N - n0, k_valid,
                  g.b_vec != 0);
    else if (g.k_nseg > 1)   // one MN-contiguous source per K segment (both directions of the input-delta product)
      load_mn_to_kmajor(b_hi, b_lo, &g.b_mn[seg], 1, -1, n0, BN, g.N, kloc, k_valid, g.b_vec != 0);
    else
      load_mn_to_kmajor(b_hi, b_lo, g.b_mn, g.b_nseg, g.b_ones, n0, BN, g.N, kloc, k_valid, g.b_vec != 0);
    fence_proxy_async();                                     // generic-proxy smem writes -> visible to the tensor core
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < BK / 8; ks++) {
        const unsigned long long ah = make_desc(a_hi + ks * 32), al = make_desc(a_lo + ks * 32);   // 32 bytes = 8 tf32 along
        const unsigned long long bh = make_desc(b_hi + ks * 32), bl = make_desc(b_lo + ks * 32);   // the swizzled row
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
  const bool st_vec = !g.ws && g.beta == 0.f && (g.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) &&
                      ((n0 + ch * half) % 4 == 0) && (!g.bias || (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0);
  for (int c = 0; c < half; c += 8) {
    float v[8];
    if (it > 0) tmem_ld8(tmem_d + ((unsigned)(32 * lq) << 16) + (unsigned)(ch * half + c), v);
    else {
#pragma unroll
      for (int e = 0; e < 8; e++) v[e] = 0.f;
    }
    if (row < g.M) {
      const int col0 = n0 + ch * half + c;
      if (st_vec && col0 + 7 < g.N) {
        if (g.bias) {
          const float4 b0 = *reinterpret_cast<const float4*>(g.bias + col0);
          const float4 b1 = *reinterpret_cast<const float4*>(g.bias + col0 + 4);
          v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
          v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
        }
        float4* dst = reinterpret_cast<float4*>(g.C + (long long)row * g.ldc + col0);
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
      } else {
#pragma unroll
        for (int e = 0; e < 8; e++) {
          const int col = col0 + e;
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

size_t tc_smem_bytes(int BN) { return (size_t)2 * (2 * BM * 128 + 2 * ((BN + 31) & ~31) * 128) + 1024; }

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
