// gemm_tc.cu -- tcgen05 (5th-gen tensor core) GEMM with fp32-grade accuracy for the batched dense products of the
// path (see gemm.cu for the list).  fp32 operands are split on the fly into TF32 hi + TF32 lo parts (truncation split, see split_tf32) and each
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
// Direct products: one smem stage per CTA and two CTAs per SM (the CTAs overlap each other's load / MMA / epilogue
// phases; the next k-block is prefetched into registers during the MMAs).  Split-K derivative products: two stages, one CTA.
// MMA completion is tracked with tcgen05.commit -> mbarrier.
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

// four 32x32b.x8 loads in flight under ONE wait: 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(unsigned taddr, float* v) {
  unsigned r[32];
#pragma unroll
  for (int q = 0; q < 4; q++)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[8 * q + 0]), "=r"(r[8 * q + 1]), "=r"(r[8 * q + 2]), "=r"(r[8 * q + 3]), "=r"(r[8 * q + 4]),
                   "=r"(r[8 * q + 5]), "=r"(r[8 * q + 6]), "=r"(r[8 * q + 7])
                 : "r"(taddr + 8 * q));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// fp32 -> (tf32 hi, tf32 lo).  hi = x with the 13 low mantissa bits cleared (exactly representable in TF32), lo = x - hi
// (exact in fp32; the tensor core reads only the TF32 bits of it).  |x - hi - tf32(lo)| <= 2^-20 |x|.  `cvt.rna.tf32`
// would round instead of truncate (2^-22) but is emulated with ~8 instructions on sm_100a and this split sits in the
// innermost load path of every product.
__device__ __forceinline__ void split_tf32(float x, unsigned& hi, unsigned& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void sts16(unsigned addr, unsigned a, unsigned b, unsigned c, unsigned d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---- operand staging.  A k-block of an operand is moved in two phases so that every global load of the k-block (and
// of the NEXT k-block while the tensor core works on this one) is in flight before the first conversion:
//   issue_*  : global -> registers (float4 slots, zero filled outside the matrix)
//   commit_* : registers -> TF32 hi/lo -> shared memory tile (K-major SWIZZLE_128B)
// Slot maps (TCT = 256 threads):
//   K-contiguous source (mode 0): slot ps <-> row ps*32 + tid/8, 16-byte chunk tid%8 (4 consecutive k)
//   MN-contiguous source (mode 1): slot u <-> warp task (tid/32) + 8u = (32-mn group, 4-k group), see issue_mncontig.
template <int NS>
__device__ __forceinline__ void issue_kcontig(float4 (&v)[NS], const float* __restrict__ p, long long ld, int R,
                                              int rows_valid, int k_valid, bool vec) {
  const int tid = threadIdx.x;
  const int chunk = tid & 7, rsub = tid >> 3;
#pragma unroll
  for (int ps = 0; ps < NS; ps++) {
    const int row = ps * 32 + rsub;
    v[ps] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < R && row < rows_valid) {
      const float* src = p + row * ld + chunk * 4;
      if (vec && chunk * 4 + 3 < k_valid) v[ps] = *reinterpret_cast<const float4*>(src);
      else {
        if (chunk * 4 + 0 < k_valid) v[ps].x = src[0];
        if (chunk * 4 + 1 < k_valid) v[ps].y = src[1];
        if (chunk * 4 + 2 < k_valid) v[ps].z = src[2];
        if (chunk * 4 + 3 < k_valid) v[ps].w = src[3];
      }
    }
  }
}
template <int NS>
__device__ __forceinline__ void commit_kcontig(const float4 (&v)[NS], unsigned hi_base, unsigned lo_base, int R) {
  const int tid = threadIdx.x;
  const int chunk = tid & 7, rsub = tid >> 3;
#pragma unroll
  for (int ps = 0; ps < NS; ps++) {
    const int row = ps * 32 + rsub;
    if (row < R) {
      unsigned h[4], l[4];
      split_tf32(v[ps].x, h[0], l[0]); split_tf32(v[ps].y, h[1], l[1]);
      split_tf32(v[ps].z, h[2], l[2]); split_tf32(v[ps].w, h[3], l[3]);
      const unsigned off = row * 128 + ((chunk ^ (row & 7)) << 4);
      sts16(hi_base + off, h[0], h[1], h[2], h[3]);
      sts16(lo_base + off, l[0], l[1], l[2], l[3]);
    }
  }
}
// element (k, c) of MN-contiguous sources: c looked up in up to 3 column blocks; c == ones_col yields 1.0
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
// MN-contiguous source: a warp task is (4 consecutive k) x (32 consecutive mn): lane = (kk = lane>>3, ml = lane&7), each
// lane reads 16 bytes (mn 4*ml .. 4*ml+3 of row k) => every 8 lanes read one contiguous 128-byte line.  To make the
// transposing 4-byte stores conflict free the tile rows of every 32-row block are PERMUTED: matrix row 4*ml + e lives
// in tile row 8*e + ml, so for a fixed component e the 32 lanes hit banks 4*(kg ^ ml) + kk = 32 distinct banks.
// The epilogue undoes the permutation (rows for A, columns for B).
template <int NS>
__device__ __forceinline__ void issue_mncontig(float4 (&v)[NS], const TcSeg* seg, int nseg, int ones_col, int r0, int R,
                                               int c_total, int k0, int k_valid, bool vec) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ml = lane & 7, kk = lane >> 3;
  const int G = (R + 31) >> 5;                 // 32-mn groups
  const int ntask = 8 * G;                     // (mg, kg), kg = 0..7
#pragma unroll
  for (int u = 0; u < NS; u++) {
    const int t = warp + u * (TCT / 32);
    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < ntask) {
      const int mg = t % G, kg = t / G;
      const int k = kg * 4 + kk;
      const int c = r0 + mg * 32 + 4 * ml;
      if (k < k_valid && mg * 32 + 4 * ml < R) {
        const long long krow = k0 + k;
        bool done = false;
        if (vec) {                            // whole quad inside one column block?
          int cl = c;
#pragma unroll
          for (int s2 = 0; s2 < 3; s2++) {
            if (s2 < nseg) {
              if (!done && cl >= 0 && cl + 3 < seg[s2].len) {
                v[u] = *reinterpret_cast<const float4*>(seg[s2].p + krow * seg[s2].ld + cl);
                done = true;
              }
              cl -= seg[s2].len;
            }
          }
        }
        if (!done) {
          v[u].x = mn_elem(seg, nseg, ones_col, c + 0, c_total, krow);
          v[u].y = mn_elem(seg, nseg, ones_col, c + 1, c_total, krow);
          v[u].z = mn_elem(seg, nseg, ones_col, c + 2, c_total, krow);
          v[u].w = mn_elem(seg, nseg, ones_col, c + 3, c_total, krow);
        }
      }
    }
  }
}
template <int NS>
__device__ __forceinline__ void commit_mncontig(const float4 (&v)[NS], unsigned hi_base, unsigned lo_base, int R) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ml = lane & 7, kk = lane >> 3;
  const int G = (R + 31) >> 5;
  const int ntask = 8 * G;
#pragma unroll
  for (int u = 0; u < NS; u++) {
    const int t = warp + u * (TCT / 32);
    if (t < ntask) {
      const int mg = t % G, kg = t / G;
      const float e4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int rho = mg * 32 + 8 * e + ml;                  // permuted tile row of matrix row mg*32 + 4*ml + e
        unsigned h, l;
        split_tf32(e4[e], h, l);
        const unsigned off = rho * 128 + ((kg ^ ml) << 4) + (kk << 2);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(hi_base + off), "r"(h) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(lo_base + off), "r"(l) : "memory");
      }
    }
  }
}
// inverse of the row permutation inside a 32-block: tile row (or TMEM lane / column) rho -> matrix index
__device__ __forceinline__ int unpermute32(int rho) { return (rho & ~31) + 4 * (rho & 7) + ((rho & 31) >> 3); }

// AMODE / BMODE: operand modes fixed at compile time (0: K-contiguous), -1: taken from g at run time.  The <0,0> instance
// (every forward product, dH, dx) drops the transposing loaders and column blocks and is a fraction of the code size --
// the generic instance spent ~30% of its issue slots waiting for instruction fetch.
template <int AMODE, int BMODE>
__global__ void __launch_bounds__(TCT, 2) gemm_tc_kernel(TcArgs g) {
  const int a_mode = (AMODE >= 0) ? AMODE : g.a_mode, b_mode = (BMODE >= 0) ? BMODE : g.b_mode;
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  __shared__ __align__(8) unsigned long long mbar_s[2];
  __shared__ unsigned tmem_base_s;
  __shared__ __align__(16) float bias_s[256];
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
  const bool two_stage = g.stages == 2;
  const bool kseg = (a_mode == 0 || b_mode == 0);

  // k-block -> (K segment, offset inside the segment); K-contiguous segments are padded to multiples of 32
  auto locate = [&](int kb, int& seg, int& kloc) {
    seg = 0; kloc = kb * BK;
    if (kseg) {
      while (seg + 1 < g.k_nseg && kloc >= ((g.k_len[seg] + BK - 1) / BK) * BK) {
        kloc -= ((g.k_len[seg] + BK - 1) / BK) * BK;
        seg++;
      }
    }
  };
  float4 ra[4], rb[8];                                       // register staging of one k-block of A (128 rows) and B (<=256)
  auto issue = [&](int kb) {
    int seg, kloc;
    locate(kb, seg, kloc);
    const int k_valid = g.k_len[seg] - kloc;
    if (a_mode == 0)
      issue_kcontig(ra, g.a_k[seg].p + (long long)m0 * g.a_k[seg].ld + kloc, g.a_k[seg].ld, BM, g.M - m0, k_valid, g.a_vec != 0);
    else
      issue_mncontig(ra, g.a_mn, 1, -1, m0, BM, g.M, kloc, k_valid, g.a_vec != 0);
    if (b_mode == 0)
      issue_kcontig(rb, g.b_k[seg].p + (long long)n0 * g.b_k[seg].ld + kloc, g.b_k[seg].ld, BN, g.N - n0, k_valid, g.b_vec != 0);
    else if (g.k_nseg > 1)   // one MN-contiguous source per K segment (both directions of the input-delta product)
      issue_mncontig(rb, &g.b_mn[seg], 1, -1, n0, BN, g.N, kloc, k_valid, g.b_vec != 0);
    else
      issue_mncontig(rb, g.b_mn, g.b_nseg, g.b_ones, n0, BN, g.N, kloc, k_valid, g.b_vec != 0);
  };

  // bias tile -> shared memory once (the epilogue must not wait on global loads)
  for (int c = tid; c < BN; c += TCT) bias_s[c] = (g.bias && n0 + c < g.N) ? g.bias[n0 + c] : 0.f;
  __syncthreads();

  int it = 0;
  if (kb0 < kb1) issue(kb0);
  for (int kb = kb0; kb < kb1; kb++, it++) {
    const int s = two_stage ? (it & 1) : 0;
    const unsigned bar = s ? bar1 : bar0;
    if (two_stage) { if (it >= 2) mbar_wait(bar, ((it >> 1) - 1) & 1); }   // the MMAs that last read this stage are done
    else if (it >= 1) mbar_wait(bar, (it - 1) & 1);
    const unsigned a_hi = smem0 + s * stage_bytes, a_lo = a_hi + a_bytes;
    const unsigned b_hi = a_lo + a_bytes, b_lo = b_hi + b_bytes;
    if (a_mode == 0) commit_kcontig(ra, a_hi, a_lo, BM); else commit_mncontig(ra, a_hi, a_lo, BM);
    if (b_mode == 0) commit_kcontig(rb, b_hi, b_lo, BN); else commit_mncontig(rb, b_hi, b_lo, BN);
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
    if (kb + 1 < kb1) issue(kb + 1);                         // next k-block's loads fly while the tensor core works
  }
  // all MMAs retire in order: the last commit covers everything
  if (it > 0) {
    const int last = it - 1;
    if (two_stage) mbar_wait((last & 1) ? bar1 : bar0, (last >> 1) & 1);
    else mbar_wait(bar0, last & 1);
  }
  tc_fence_after();

  // ---- epilogue: TMEM -> registers -> global.  warp w reads lanes 32*(w&3).., column half (w>>2).
  // Operands staged from MN-contiguous sources carry the 32-block permutation: undo it here.
  const int lq = warp & 3, ch = warp >> 2;
  const int trow = 32 * lq + lane;                                   // TMEM lane = tile row
  const int row = m0 + (a_mode ? unpermute32(trow) : trow);
  const int half = BN / 2;
  const bool st_vec = !g.ws && b_mode == 0 && g.beta == 0.f && (g.ldc % 4 == 0) &&
                      ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) && ((n0 + ch * half) % 4 == 0);
  // one group of 8 accumulator columns -> global
  auto emit8 = [&](int tc0, const float* v) {                       // tc0: first tile column of the group
    if (row >= g.M) return;
    if (st_vec && n0 + tc0 + 7 < g.N) {
      const float4 b0 = *reinterpret_cast<const float4*>(&bias_s[tc0]);
      const float4 b1 = *reinterpret_cast<const float4*>(&bias_s[tc0 + 4]);
      float4* dst = reinterpret_cast<float4*>(g.C + (long long)row * g.ldc + n0 + tc0);
      dst[0] = make_float4(v[0] + b0.x, v[1] + b0.y, v[2] + b0.z, v[3] + b0.w);
      dst[1] = make_float4(v[4] + b1.x, v[5] + b1.y, v[6] + b1.z, v[7] + b1.w);
    } else {
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const int lc = b_mode ? unpermute32(tc0 + e) : tc0 + e;  // column inside the N tile
        const int col = n0 + lc;
        if (col < g.N) {
          if (g.ws) {
            g.ws[((size_t)blockIdx.z * g.M + row) * g.N + col] = v[e];
          } else {
            const float o = v[e] + bias_s[lc];
            float* dst = g.C + (long long)row * g.ldc + col;
            *dst = (g.beta != 0.f) ? fmaf(g.beta, *dst, o) : o;
          }
        }
      }
    }
  };
  const unsigned tbase = tmem_d + ((unsigned)(32 * lq) << 16) + (unsigned)(ch * half);
  int c = 0;
  for (; c + 32 <= half; c += 32) {                                  // 32 columns per TMEM round trip
    float v[32];
    if (it > 0) tmem_ld32(tbase + (unsigned)c, v);
    else {
#pragma unroll
      for (int e = 0; e < 32; e++) v[e] = 0.f;
    }
#pragma unroll
    for (int q = 0; q < 4; q++) emit8(ch * half + c + 8 * q, v + 8 * q);
  }
  for (; c < half; c += 8) {
    float v[8];
    if (it > 0) tmem_ld8(tbase + (unsigned)c, v);
    else {
#pragma unroll
      for (int e = 0; e < 8; e++) v[e] = 0.f;
    }
    emit8(ch * half + c, v);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Dedicated kernel for the derivative products  P[M x N] = A^T [B0 | B1 | 1]  (A: [K x M], B blocks: [K x len], all
// row-major, reduction over the K = N_columns rows), split over K.  Same tiles / MMAs / permutation as above, but every
// per-slot quantity that does not depend on k (source pointer, leading dimension, smem offset, special-quad flag) is
// computed once before the k loop, so a k-block costs one LDG.128, 12 conversion and 8 store instructions per slot.
template <int NSB>
__global__ void __launch_bounds__(TCT, 2) gemm_tn_kernel(TcArgs g) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  __shared__ __align__(8) unsigned long long mbar_s[2];
  __shared__ unsigned tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ml = lane & 7, kk = lane >> 3;
  constexpr int BN = NSB * 32;
  constexpr unsigned a_bytes = BM * 128, b_bytes = BN * 128, stage_bytes = 2 * a_bytes + 2 * b_bytes;
  const unsigned smem0 = (smem_u32(tc_smem) + 1023u) & ~1023u;
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
  const int K = g.k_len[0];
  const unsigned idesc = make_idesc(BN);
  const bool two_stage = g.stages == 2;

  // ---- per-slot invariants.  A: task t = warp + 8u -> (mg = warp & 3, kg = (warp >> 2) + 2u)
  const float* pA;            // row (kg0*4 + kk), column m0 + mg*32 + 4*ml ; slot u adds 8*u rows
  bool a_fast;                // whole quad inside the matrix and 16-byte loadable
  unsigned offA[4];
  {
    const int mg = warp & 3, c = m0 + mg * 32 + 4 * ml;
    a_fast = g.a_vec && (c + 3 < g.M);
    pA = g.a_mn[0].p + (long long)(((warp >> 2)) * 4 + kk) * g.a_mn[0].ld + c;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int kg = (warp >> 2) + 2 * u;
      offA[u] = (unsigned)((mg * 32 + ml) * 128 + ((kg ^ ml) << 4) + (kk << 2));
    }
  }
  // B: task t = warp + 8u -> (mg = t % NSB, kg = t / NSB); the quad's 4 columns usually sit inside one column block
  const float* pB[NSB];
  int ldB[NSB], cB[NSB];
  unsigned offB[NSB];
  bool b_fast[NSB];
#pragma unroll
  for (int u = 0; u < NSB; u++) {
    const int t = warp + 8 * u, mg = t % NSB, kg = t / NSB;
    const int c = n0 + mg * 32 + 4 * ml;
    cB[u] = c;
    offB[u] = (unsigned)((mg * 32 + ml) * 128 + ((kg ^ ml) << 4) + (kk << 2));
    pB[u] = nullptr; ldB[u] = 0; b_fast[u] = false;
    int cl = c;
#pragma unroll
    for (int s2 = 0; s2 < 3; s2++) {
      if (s2 < g.b_nseg) {
        if (g.b_vec && !b_fast[u] && cl >= 0 && cl + 3 < g.b_mn[s2].len) {
          pB[u] = g.b_mn[s2].p + (long long)(kg * 4 + kk) * g.b_mn[s2].ld + cl;
          ldB[u] = (int)g.b_mn[s2].ld;
          b_fast[u] = true;
        }
        cl -= g.b_mn[s2].len;
      }
    }
  }

  float4 ra[4], rb[NSB];
  auto issue = [&](int kb) {
    const int k0 = kb * BK;
    const bool full = (k0 + BK <= K);
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int k = ((warp >> 2) + 2 * u) * 4 + kk;
      ra[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (full || k0 + k < K) {
        if (a_fast) ra[u] = *reinterpret_cast<const float4*>(pA + (long long)(k0 + 8 * u) * g.a_mn[0].ld);
        else {
          const int c = m0 + (warp & 3) * 32 + 4 * ml;
          const long long krow = k0 + k;
          ra[u].x = mn_elem(g.a_mn, 1, -1, c + 0, g.M, krow); ra[u].y = mn_elem(g.a_mn, 1, -1, c + 1, g.M, krow);
          ra[u].z = mn_elem(g.a_mn, 1, -1, c + 2, g.M, krow); ra[u].w = mn_elem(g.a_mn, 1, -1, c + 3, g.M, krow);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < NSB; u++) {
      const int k = ((warp + 8 * u) / NSB) * 4 + kk;
      rb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (full || k0 + k < K) {
        if (b_fast[u]) rb[u] = *reinterpret_cast<const float4*>(pB[u] + (long long)k0 * ldB[u]);
        else {
          const long long krow = k0 + k;
          rb[u].x = mn_elem(g.b_mn, g.b_nseg, g.b_ones, cB[u] + 0, g.N, krow);
          rb[u].y = mn_elem(g.b_mn, g.b_nseg, g.b_ones, cB[u] + 1, g.N, krow);
          rb[u].z = mn_elem(g.b_mn, g.b_nseg, g.b_ones, cB[u] + 2, g.N, krow);
          rb[u].w = mn_elem(g.b_mn, g.b_nseg, g.b_ones, cB[u] + 3, g.N, krow);
        }
      }
    }
  };
  auto put4 = [&](const float4& v, unsigned hi_base, unsigned lo_base, unsigned off) {
    const float e4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; e++) {             // matrix row 4*ml + e -> tile row 8*e + ml: +1024 bytes per e
      unsigned h, l;
      split_tf32(e4[e], h, l);
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(hi_base + off + e * 1024), "r"(h) : "memory");
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(lo_base + off + e * 1024), "r"(l) : "memory");
    }
  };

  int it = 0;
  if (kb0 < kb1) issue(kb0);
  for (int kb = kb0; kb < kb1; kb++, it++) {
    const int s = two_stage ? (it & 1) : 0;
    const unsigned bar = s ? bar1 : bar0;
    if (two_stage) { if (it >= 2) mbar_wait(bar, ((it >> 1) - 1) & 1); }
    else if (it >= 1) mbar_wait(bar, (it - 1) & 1);
    const unsigned a_hi = smem0 + s * stage_bytes, a_lo = a_hi + a_bytes;
    const unsigned b_hi = a_lo + a_bytes, b_lo = b_hi + b_bytes;
#pragma unroll
    for (int u = 0; u < 4; u++) put4(ra[u], a_hi, a_lo, offA[u]);
#pragma unroll
    for (int u = 0; u < NSB; u++) put4(rb[u], b_hi, b_lo, offB[u]);
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < BK / 8; ks++) {
        const unsigned long long ah = make_desc(a_hi + ks * 32), al = make_desc(a_lo + ks * 32);
        const unsigned long long bh = make_desc(b_hi + ks * 32), bl = make_desc(b_lo + ks * 32);
        mma_tf32(tmem_d, al, bh, idesc, (it > 0 || ks > 0) ? 1u : 0u);
        mma_tf32(tmem_d, ah, bl, idesc, 1u);
        mma_tf32(tmem_d, ah, bh, idesc, 1u);
      }
      mma_commit(bar);
    }
    if (kb + 1 < kb1) issue(kb + 1);
  }
  if (it > 0) {
    const int last = it - 1;
    if (two_stage) mbar_wait((last & 1) ? bar1 : bar0, (last >> 1) & 1);
    else mbar_wait(bar0, last & 1);
  }
  tc_fence_after();
  // ---- epilogue: partial tile -> workspace slice of this split, undoing the row and column permutations
  const int lq = warp & 3, ch = warp >> 2;
  const int row = m0 + unpermute32(32 * lq + lane);
  constexpr int half = BN / 2;
  float* __restrict__ wsz = g.ws + (size_t)blockIdx.z * g.M * g.N;
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
        const int col = n0 + unpermute32(ch * half + c + e);
        if (col < g.N) wsz[(size_t)row * g.N + col] = v[e];
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

size_t tc_smem_bytes(int BN, int stages = 2) { return (size_t)stages * (2 * BM * 128 + 2 * ((BN + 31) & ~31) * 128) + 1024; }
// smem stages of the direct products / of the split-K derivative products.  1: one stage, two CTAs per SM overlap each
// other's load / MMA / epilogue phases (measured better for the few-k-block direct products: xproj 52.6 -> 44.3 us);
// 2: two stages, one CTA per SM (better for the long split-K loops: wgrad 109 vs 132 us).
int g_tc_stages = 1, g_tn_stages = 2;

}  // namespace

int gemm_tc_configure() {
  if (const char* e = getenv("CLSTM_B200_TC_STAGES")) g_tc_stages = (atoi(e) == 2) ? 2 : 1;
  if (const char* e = getenv("CLSTM_B200_TN_STAGES")) g_tn_stages = (atoi(e) == 2) ? 2 : 1;
  cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<-1, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(256));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(gemm_tc_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(256));
#define CB200_TN_ATTR(N_)                                                                                          \
  if (e == cudaSuccess)                                                                                            \
    e = cudaFuncSetAttribute(gemm_tn_kernel<N_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(32 * N_));
  CB200_TN_ATTR(1) CB200_TN_ATTR(2) CB200_TN_ATTR(3) CB200_TN_ATTR(4) CB200_TN_ATTR(5) CB200_TN_ATTR(6) CB200_TN_ATTR(7)
  CB200_TN_ATTR(8)
#undef CB200_TN_ATTR
  return (int)e;
}

void tc_reduce_scatter(cudaStream_t st, int M, int N, int splits, const float* ws, const TcOut& o, float beta, int num_sms) {
  const size_t total = (size_t)M * N;
  size_t nb = (total + 255) / 256;
  if (nb > (size_t)num_sms * 8) nb = (size_t)num_sms * 8;
  tc_reduce_scatter_kernel<<<(int)nb, 256, 0, st>>>(M, N, splits, ws, o, beta);
}

int gemm_tc(cudaStream_t st, TcArgs g, const TcOut* scatter, int num_sms) {
  if (g.M <= 0 || g.N <= 0) return 0;
  // N tile: multiple of 16, <= 256, as few tiles as possible
  const int ntiles = (g.N + 255) / 256;
  int BN = (((g.N + ntiles - 1) / ntiles) + 15) & ~15;
  if (g.b_mode == 1) BN = (BN + 31) & ~31;    // permuted 32-row blocks must be complete (see issue_mncontig)
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
    splits = (num_sms * (g_tn_stages == 1 ? 2 : 1)) / (mtiles * ntiles);
    if (splits > nkb / 4) splits = nkb / 4;
    if (splits > 64) splits = 64;
    if (splits < 1) splits = 1;
    while (splits > 1 && (size_t)splits * g.M * g.N > g.ws_floats) splits--;
  }
  g.kb_per_split = (nkb + splits - 1) / splits;
  splits = (nkb + g.kb_per_split - 1) / g.kb_per_split;
  if (!scatter) g.ws = nullptr;
  g.stages = scatter ? g_tn_stages : g_tc_stages;
  dim3 grid(ntiles, mtiles, splits);
  if (g.a_mode == 1 && g.b_mode == 1 && scatter) {
    switch (BN / 32) {
      case 1: gemm_tn_kernel<1><<<grid, TCT, tc_smem_bytes(BN, g.stages), st>>>(g); break;
      case 2: gemm_tn_kernel<2><<<grid, TCT, tc_smem_bytes(BN, g.stages), st>>>(g); break;
      case 3: gemm_tn_kernel<3><<<grid, TCT, tc_smem_bytes(BN, g.stages), st>>>(g); break;
      case 4: gemm_tn_kernel<4><<<grid, TCT, tc_smem_bytes(BN, g.stages), st>>>(g); break;
      case 5: gemm_tn_kernel<5><<<grid, TCT, tc_smem_bytes(BN, g.stages), st>>>(g); break;
      case 6: gemm_tn_kernel<6><<<grid, TCT, tc_smem_bytes(BN, g.stages), st>>>(g); break;
      case 7: gemm_tn_kernel<7><<<grid, TCT, tc_smem_bytes(BN, g.stages), st>>>(g); break;
      default: gemm_tn_kernel<8><<<grid, TCT, tc_smem_bytes(BN, g.stages), st>>>(g); break;
    }
  } else {
    if (g.a_mode == 0 && g.b_mode == 0) gemm_tc_kernel<0, 0><<<grid, TCT, tc_smem_bytes(BN, g.stages), st>>>(g);
    else gemm_tc_kernel<-1, -1><<<grid, TCT, tc_smem_bytes(BN, g.stages), st>>>(g);
  }
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
