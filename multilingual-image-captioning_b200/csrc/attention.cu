// Attention for the short sequences of the captioning path (50 / 197 visual tokens, 64 text tokens, head_dim 64):
// the Q/K/V tiles of one (batch, head) fit in shared memory.  S=QK^T, softmax and PV run on warp-level bf16 MMA
// (m16n8k16) with fp32 accumulation; these problems are latency- and issue-bound, far too small to amortise a TMEM
// round trip, so the legacy tensor path is the right tool here (~1.3% of the CLIP-mBART step's FLOPs; SURVEY.md
// §8a E3/D2/D3).  Two kernel families: the row-tiled kernels further down (default, any length <= 256) and the
// original one-CTA-per-head kernels (<= 64 tokens, kept behind mic_attention_impl(1) as the A/B reference;
// tools/attn_bench.py, profiles/r02_attention_ab.txt).
//
// Semantics = flax dot_product_attention_weights as used by FlaxCLIPAttention / FlaxMBartAttention:
// scores = (q/sqrt(64)) . k ; additive mask 0/-inf from (causal AND key padding) ; softmax ; . v
#include "common.cuh"
#include "decode_device.cuh"

#include "../../include/mic_b200.h"

namespace {

using micdec::HD;          // head dim 64
using micdec::DecAttnArgs;
constexpr int TMAX = 64;   // max queries / keys per (batch, head) tile
constexpr int LDS = 72;    // smem row pitch in bf16 (144 B: conflict-free fragment loads)

__device__ __forceinline__ void mma_bf16_16816(float* d, const uint32_t* a, const uint32_t* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// A fragment (16x16) from row-major smem tile S[m][k]
__device__ __forceinline__ void frag_a(const bf16* s, int r0, int k0, int lane, uint32_t* a) {
  const int g = lane >> 2, t = lane & 3;
  a[0] = *reinterpret_cast<const uint32_t*>(s + (r0 + g) * LDS + k0 + 2 * t);
  a[1] = *reinterpret_cast<const uint32_t*>(s + (r0 + g + 8) * LDS + k0 + 2 * t);
  a[2] = *reinterpret_cast<const uint32_t*>(s + (r0 + g) * LDS + k0 + 2 * t + 8);
  a[3] = *reinterpret_cast<const uint32_t*>(s + (r0 + g + 8) * LDS + k0 + 2 * t + 8);
}
// B fragment (16x8, "col") from smem tile stored as Bs[n][k] (k contiguous)
__device__ __forceinline__ void frag_b(const bf16* s, int n0, int k0, int lane, uint32_t* b) {
  const int g = lane >> 2, t = lane & 3;
  b[0] = *reinterpret_cast<const uint32_t*>(s + (n0 + g) * LDS + k0 + 2 * t);
  b[1] = *reinterpret_cast<const uint32_t*>(s + (n0 + g) * LDS + k0 + 2 * t + 8);
}

// B fragment (k16 x n8) from a ROW-MAJOR tile M[k][n] (n contiguous) via ldmatrix.trans: no transposed copy needed
__device__ __forceinline__ void frag_b_trans(const bf16* s, int k0, int n0, int lane, uint32_t* b) {
  const bf16* row = s + (k0 + (lane & 15)) * LDS + n0;
  const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(row));
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(b[0]), "=r"(b[1]) : "r"(addr));
}
// A fragment (m16 x k16) of the TRANSPOSE of a row-major tile M[k][m]: A[m][k] = M[k][m]
__device__ __forceinline__ void frag_a_trans(const bf16* s, int m0, int k0, int lane, uint32_t* a) {
  const int i = lane & 7, sel = lane >> 3;     // matrices: 0:(k0,m0) 1:(k0,m0+8) 2:(k0+8,m0) 3:(k0+8,m0+8)
  const bf16* row = s + (k0 + i + (sel >> 1) * 8) * LDS + m0 + (sel & 1) * 8;
  const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(row));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
               : "r"(addr));
}

struct AttnArgs {
  const bf16 *Q, *K, *V;
  long long ldq, ldk, ldv;
  bf16* O;
  long long ldo;
  float* lse;            // [B, H, Tq]
  const int* key_mask;   // [B, Tk] (1 = keep) or null
  int causal, B, H, Tq, Tk;
  float scale;
  // backward only
  const bf16* dO;
  long long lddo;
  bf16 *dQ, *dK, *dV;
  long long lddq, lddk, lddv;
};

// load a [rows x 64] head slice into smem, row-major, zero-filled beyond `rows` (16-byte vectors, conflict-free)
__device__ __forceinline__ void load_tile(const bf16* g, long long ld, int rows, bf16* s, int tid, int nthreads) {
  for (int i = tid; i < TMAX * (HD / 8); i += nthreads) {
    const int r = i >> 3, c = (i & 7) * 8;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (r < rows) u = *reinterpret_cast<const uint4*>(g + (long long)r * ld + c);
    *reinterpret_cast<uint4*>(s + r * LDS + c) = u;
  }
}

__device__ __forceinline__ bool key_allowed(int row, int col, int Tk, int causal, const int* km) {
  if (col >= Tk) return false;
  if (causal && col > row) return false;
  if (km && km[col] == 0) return false;
  return true;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attention_fwd_kernel(const AttnArgs a) {
  __shared__ __align__(16) bf16 sQ[TMAX * LDS];
  __shared__ __align__(16) bf16 sK[TMAX * LDS];
  __shared__ __align__(16) bf16 sV[TMAX * LDS];
  __shared__ int sMask[TMAX];
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  load_tile(a.Q + ((long long)b * a.Tq) * a.ldq + h * HD, a.ldq, a.Tq, sQ, tid, 128);
  load_tile(a.K + ((long long)b * a.Tk) * a.ldk + h * HD, a.ldk, a.Tk, sK, tid, 128);
  load_tile(a.V + ((long long)b * a.Tk) * a.ldv + h * HD, a.ldv, a.Tk, sV, tid, 128);
  if (tid < TMAX) sMask[tid] = (a.key_mask && tid < a.Tk) ? a.key_mask[(long long)b * a.Tk + tid] : 1;
  __syncthreads();
  const int r0 = warp * 16;
  if (r0 >= a.Tq) return;
  const int g = lane >> 2, t = lane & 3;
  float s[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) s[nt][j] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t af[4];
    frag_a(sQ, r0, kk * 16, lane, af);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      uint32_t bfr[2];
      frag_b(sK, nt * 8, kk * 16, lane, bfr);
      mma_bf16_16816(s[nt], af, bfr);
    }
  }
  float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = r0 + g + (j >> 1) * 8, col = nt * 8 + 2 * t + (j & 1);
      const bool ok = key_allowed(row, col, a.Tk, a.causal, a.key_mask ? sMask : nullptr);
      s[nt][j] = ok ? s[nt][j] * a.scale : -INFINITY;
      mx[j >> 1] = fmaxf(mx[j >> 1], s[nt][j]);
    }
  float sum[2] = {0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
    mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
    if (mx[i] == -INFINITY) mx[i] = 0.f;
  }
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float p = __expf(s[nt][j] - mx[j >> 1]);
      s[nt][j] = p;
      sum[j >> 1] += p;
    }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], 1);
    sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], 2);
  }
  float o[8][4];
#pragma unroll
  for (int nd = 0; nd < 8; ++nd)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[nd][j] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t af[4];
    af[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
    af[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
    af[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    af[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
      uint32_t bfr[2];
      frag_b_trans(sV, kk * 16, nd * 8, lane, bfr);
      mma_bf16_16816(o[nd], af, bfr);
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int row = r0 + g + i * 8;
    if (row < a.Tq) {
      const float inv = sum[i] > 0.f ? 1.0f / sum[i] : 0.f;
      bf16* orow = a.O + ((long long)b * a.Tq + row) * a.ldo + h * HD;
#pragma unroll
      for (int nd = 0; nd < 8; ++nd)
        *reinterpret_cast<uint32_t*>(orow + nd * 8 + 2 * t) = pack_bf16(o[nd][2 * i] * inv, o[nd][2 * i + 1] * inv);
      if (a.lse && t == 0) a.lse[((long long)b * a.H + h) * a.Tq + row] = mx[i] + logf(sum[i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward: recompute P from Q,K and the saved log-sum-exp; dV = P^T dO ; dP = dO V^T ;
// dS = scale * P o (dP - rowsum(dO o O)) ; dQ = dS K ; dK = dS^T Q
// ---------------------------------------------------------------------------------------------
constexpr int BWD_SMEM = (6 * TMAX * LDS) * 2 + TMAX * 4 * 2;

__global__ void __launch_bounds__(128) attention_bwd_kernel(const AttnArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sK = sQ + TMAX * LDS;
  bf16* sV = sK + TMAX * LDS;
  bf16* sdO = sV + TMAX * LDS;
  bf16* sP = sdO + TMAX * LDS;       // [q][key]
  bf16* sdS = sP + TMAX * LDS;       // [q][key]
  float* sD = reinterpret_cast<float*>(sdS + TMAX * LDS);
  int* sMask = reinterpret_cast<int*>(sD + TMAX);

  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bf16* gdO = a.dO + ((long long)b * a.Tq) * a.lddo + h * HD;
  const bf16* gO = a.O + ((long long)b * a.Tq) * a.ldo + h * HD;
  load_tile(a.Q + ((long long)b * a.Tq) * a.ldq + h * HD, a.ldq, a.Tq, sQ, tid, 128);
  load_tile(a.K + ((long long)b * a.Tk) * a.ldk + h * HD, a.ldk, a.Tk, sK, tid, 128);
  load_tile(a.V + ((long long)b * a.Tk) * a.ldv + h * HD, a.ldv, a.Tk, sV, tid, 128);
  load_tile(gdO, a.lddo, a.Tq, sdO, tid, 128);
  if (tid < TMAX) sMask[tid] = (a.key_mask && tid < a.Tk) ? a.key_mask[(long long)b * a.Tk + tid] : 1;
  {
    // D[q] = sum_d dO[q][d] * O[q][d] : two threads per row
    const int r = tid >> 1, hf = tid & 1;
    float d = 0.f;
    if (r < a.Tq) {
#pragma unroll
      for (int c = hf * 32; c < hf * 32 + 32; c += 8) {
        const uint4 u = *reinterpret_cast<const uint4*>(gdO + (long long)r * a.lddo + c);
        const uint4 w = *reinterpret_cast<const uint4*>(gO + (long long)r * a.ldo + c);
        const uint32_t* uu = reinterpret_cast<const uint32_t*>(&u);
        const uint32_t* ww = reinterpret_cast<const uint32_t*>(&w);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 x = unpack_bf16(uu[j]), y = unpack_bf16(ww[j]);
          d += x.x * y.x + x.y * y.y;
        }
      }
    }
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    if (hf == 0) sD[r] = d;
  }
  __syncthreads();

  const int r0 = warp * 16;
  const int g = lane >> 2, t = lane & 3;
  {
    float s[8][4], dp[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[nt][j] = 0.f;
        dp[nt][j] = 0.f;
      }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t aq[4], ado[4];
      frag_a(sQ, r0, kk * 16, lane, aq);
      frag_a(sdO, r0, kk * 16, lane, ado);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        uint32_t bk[2], bv[2];
        frag_b(sK, nt * 8, kk * 16, lane, bk);
        frag_b(sV, nt * 8, kk * 16, lane, bv);
        mma_bf16_16816(s[nt], aq, bk);
        mma_bf16_16816(dp[nt], ado, bv);
      }
    }
    float lse[2], dd[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int row = r0 + g + i * 8;
      lse[i] = row < a.Tq ? a.lse[((long long)b * a.H + h) * a.Tq + row] : 0.f;
      dd[i] = sD[row];
    }
    // P and dS (in place: s <- P, dp <- dS); stash both row-major for the key-side contractions
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int row = r0 + g + i * 8;
        float pv[2], dsv[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = i * 2 + e, col = nt * 8 + 2 * t + e;
          const bool ok = row < a.Tq && key_allowed(row, col, a.Tk, a.causal, a.key_mask ? sMask : nullptr);
          const float p = ok ? __expf(s[nt][j] * a.scale - lse[i]) : 0.f;
          const float ds = p * (dp[nt][j] - dd[i]) * a.scale;
          s[nt][j] = p;
          dp[nt][j] = ds;
          pv[e] = p;
          dsv[e] = ds;
        }
        *reinterpret_cast<uint32_t*>(sP + row * LDS + nt * 8 + 2 * t) = pack_bf16(pv[0], pv[1]);
        *reinterpret_cast<uint32_t*>(sdS + row * LDS + nt * 8 + 2 * t) = pack_bf16(dsv[0], dsv[1]);
      }
    }
    // dQ = dS K   (B: n = d, k = key -> row-major sK through ldmatrix.trans)
    float dq[8][4];
#pragma unroll
    for (int nd = 0; nd < 8; ++nd)
#pragma unroll
      for (int j = 0; j < 4; ++j) dq[nd][j] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t af[4];
      af[0] = pack_bf16(dp[2 * kk][0], dp[2 * kk][1]);
      af[1] = pack_bf16(dp[2 * kk][2], dp[2 * kk][3]);
      af[2] = pack_bf16(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
      af[3] = pack_bf16(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
      for (int nd = 0; nd < 8; ++nd) {
        uint32_t bfr[2];
        frag_b_trans(sK, kk * 16, nd * 8, lane, bfr);
        mma_bf16_16816(dq[nd], af, bfr);
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int row = r0 + g + i * 8;
      if (row < a.Tq) {
        bf16* drow = a.dQ + ((long long)b * a.Tq + row) * a.lddq + h * HD;
#pragma unroll
        for (int nd = 0; nd < 8; ++nd)
          *reinterpret_cast<uint32_t*>(drow + nd * 8 + 2 * t) = pack_bf16(dq[nd][2 * i], dq[nd][2 * i + 1]);
      }
    }
  }
  __syncthreads();
  // key side: this warp owns keys [r0, r0+16):  dV = P^T dO,  dK = dS^T Q
  if (r0 < a.Tk) {
    float dv[8][4], dk[8][4];
#pragma unroll
    for (int nd = 0; nd < 8; ++nd)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        dv[nd][j] = 0.f;
        dk[nd][j] = 0.f;
      }
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
      uint32_t ap[4], ads[4];
      frag_a_trans(sP, r0, qq * 16, lane, ap);
      frag_a_trans(sdS, r0, qq * 16, lane, ads);
#pragma unroll
      for (int nd = 0; nd < 8; ++nd) {
        uint32_t bdo[2], bq[2];
        frag_b_trans(sdO, qq * 16, nd * 8, lane, bdo);
        frag_b_trans(sQ, qq * 16, nd * 8, lane, bq);
        mma_bf16_16816(dv[nd], ap, bdo);
        mma_bf16_16816(dk[nd], ads, bq);
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int key = r0 + g + i * 8;
      if (key < a.Tk) {
        bf16* krow = a.dK + ((long long)b * a.Tk + key) * a.lddk + h * HD;
        bf16* vrow = a.dV + ((long long)b * a.Tk + key) * a.lddv + h * HD;
#pragma unroll
        for (int nd = 0; nd < 8; ++nd) {
          *reinterpret_cast<uint32_t*>(krow + nd * 8 + 2 * t) = pack_bf16(dk[nd][2 * i], dk[nd][2 * i + 1]);
          *reinterpret_cast<uint32_t*>(vrow + nd * 8 + 2 * t) = pack_bf16(dv[nd][2 * i], dv[nd][2 * i + 1]);
        }
      }
    }
  }
}

constexpr int GEN_MAX_T = 256;   // longest sequence of the row-tiled kernels (ViT-B/16: 197 tokens)

// ---------------------------------------------------------------------------------------------
// Row-tiled kernels (any Tq, Tk <= 256).  One WARP owns a 16-row tile of queries (forward, dQ) or of keys
// (dK/dV) and walks the other sequence in register-sized blocks, so nothing but the operands lives in shared
// memory and there is no CTA-wide barrier after the load:
//   * a CTA = up to 8 warps = up to 8 row tiles of one (batch, head); ceil(tiles/8) CTAs share a head and sit
//     next to each other in the grid (their common K/V or Q/dO tile is served by L2 the second time);
//   * operands arrive by cp.async (every 16-byte chunk of the CTA in flight at once: these kernels used to be
//     bound by the latency of dependent global loads, not by the tensor pipe or by bandwidth);
//   * all fragments come from ldmatrix.x4 (one instruction feeds two MMAs); P / dS go from accumulator
//     registers straight into the A operand of the next MMA; the transposed products of the dK/dV kernel are
//     formed by computing S^T = K Q^T directly, so no tile is ever staged through shared memory;
//   * loops run to the sequence length rounded up to 16 (197 -> 208), with a 16- or 32-wide tail block.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(void* s, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void ldsm_x4(const bf16* p, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_addr(p)));
}
__device__ __forceinline__ void ldsm_x4_t(const bf16* p, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_addr(p)));
}
// A fragment (m16 x k16) of a row-major tile T[m][k]
__device__ __forceinline__ void lda(const bf16* s, int r0, int k0, int lane, uint32_t* a) {
  ldsm_x4(s + (r0 + (lane & 15)) * LDS + k0 + (lane >> 4) * 8, a);
}
// B fragments of TWO n-tiles (n0..n0+15) x k16 from a tile stored [n][k]: b[0..1] = tile n0, b[2..3] = tile n0+8
__device__ __forceinline__ void ldb2(const bf16* s, int n0, int k0, int lane, uint32_t* b) {
  const int m = lane >> 3;
  ldsm_x4(s + (n0 + (lane & 7) + (m >> 1) * 8) * LDS + k0 + (m & 1) * 8, b);
}
// B fragments of TWO n-tiles (n0..n0+15) x k16 from a tile stored [k][n] (row-major in k)
__device__ __forceinline__ void ldb2_t(const bf16* s, int k0, int n0, int lane, uint32_t* b) {
  const int m = lane >> 3;
  ldsm_x4_t(s + (k0 + (lane & 7) + (m & 1) * 8) * LDS + n0 + (m >> 1) * 8, b);
}
// rows [row0, row0+nrows) of a head slice -> smem rows [0, nrows) by cp.async; rows >= rows_valid are zeroed.
// nthreads is a multiple of 8: a thread keeps its 16-byte column and steps nthreads/8 rows (pointer increments only)
__device__ __forceinline__ void load_rows_async(const bf16* g, long long ld, int row0, int rows_valid, int nrows, bf16* s,
                                                int tid, int nthreads) {
  const int c = (tid & 7) * 8, rstep = nthreads >> 3;
  int r = tid >> 3;
  const bf16* gp = g + (long long)(row0 + r) * ld + c;
  bf16* sp = s + r * LDS + c;
  const long long gstep = (long long)rstep * ld;
  const int nv = min(nrows, rows_valid - row0);
  for (; r < nv; r += rstep, gp += gstep, sp += rstep * LDS) cp_async16(sp, gp);
  for (; r < nrows; r += rstep, sp += rstep * LDS) *reinterpret_cast<uint4*>(sp) = make_uint4(0, 0, 0, 0);
}
// 16 rows x 64 columns of bf16 held as MMA accumulators -> global, through the warp's own (dead) smem tile so
// that the stores are whole 128-byte rows
__device__ __forceinline__ void store_tile_rows(float (*acc)[4], const float* rs, bf16* stile, bf16* gbase, long long ld,
                                                int row0, int rows_valid, int lane) {
  const int g = lane >> 2, t = lane & 3;
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int nd = 0; nd < 8; ++nd)
      *reinterpret_cast<uint32_t*>(stile + (g + i * 8) * LDS + nd * 8 + 2 * t) =
          pack_bf16(acc[nd][2 * i] * rs[i], acc[nd][2 * i + 1] * rs[i]);
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = it * 4 + (lane >> 3), c = (lane & 7) * 8;
    if (row0 + r < rows_valid)
      *reinterpret_cast<uint4*>(gbase + (long long)(row0 + r) * ld + c) = *reinterpret_cast<const uint4*>(stile + r * LDS + c);
  }
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

struct TileGeom {
  int nsplit;    // CTAs per (batch, head)
  int per_cta;   // 16-row tiles (= warps) per CTA
  int Tqp, Tkp;  // sequence lengths rounded up to 16
};

// one block of NT*8 keys of the forward pass for a 16-query tile (online softmax; the running maximum m is kept in
// raw score units and the scale is folded into the exponent: p = 2^(s*sc2 - m*sc2), one FFMA + one MUFU per score).
// MASKED = false is the interior block: every key exists and is visible to every row, no predicate per score.
template <int NT, bool MASKED>
__device__ __forceinline__ void fwd_block(const AttnArgs& a, const uint32_t (*qf)[4], const bf16* sK, const bf16* sV,
                                          const int* sMask, int key0, int row_base, int lane, float sc2, float* m,
                                          float* l, float (*o)[4]) {
  const int g = lane >> 2, t = lane & 3;
  float s[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) s[nt][j] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      uint32_t b[4];
      ldb2(sK, key0 + np * 16, kk * 16, lane, b);
      mma_bf16_16816(s[2 * np], qf[kk], b);
      mma_bf16_16816(s[2 * np + 1], qf[kk], b + 2);
    }
  float bm[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (MASKED) {
        const int row = row_base + g + (j >> 1) * 8, col = key0 + nt * 8 + 2 * t + (j & 1);
        if (!key_allowed(row, col, a.Tk, a.causal, sMask)) s[nt][j] = -INFINITY;
      }
      bm[j >> 1] = fmaxf(bm[j >> 1], s[nt][j]);
    }
  float msc[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    bm[i] = fmaxf(bm[i], __shfl_xor_sync(0xffffffffu, bm[i], 1));
    bm[i] = fmaxf(bm[i], __shfl_xor_sync(0xffffffffu, bm[i], 2));
    const float mn = fmaxf(m[i], bm[i]);
    msc[i] = mn == -INFINITY ? 0.f : mn * sc2;
    const float corr = fast_exp2(m[i] * sc2 - msc[i]);   // m = -inf (nothing seen yet) -> 0
    m[i] = mn;
    l[i] *= corr;
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
      o[nd][2 * i] *= corr;
      o[nd][2 * i + 1] *= corr;
    }
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float p = fast_exp2(fmaf(s[nt][j], sc2, -msc[j >> 1]));
      s[nt][j] = p;
      l[j >> 1] += p;   // per-thread partial; the quad is summed once at the end
    }
#pragma unroll
  for (int kk = 0; kk < NT / 2; ++kk) {
    uint32_t af[4];
    af[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
    af[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
    af[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    af[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
    for (int ndp = 0; ndp < 4; ++ndp) {
      uint32_t b[4];
      ldb2_t(sV, key0 + kk * 16, ndp * 16, lane, b);
      mma_bf16_16816(o[2 * ndp], af, b);
      mma_bf16_16816(o[2 * ndp + 1], af, b + 2);
    }
  }
}
// picks the predicate-free instantiation when the whole block is visible to the whole row tile
template <int NT>
__device__ __forceinline__ void fwd_block_any(const AttnArgs& a, const uint32_t (*qf)[4], const bf16* sK, const bf16* sV,
                                              const int* sMask, int key0, int row_base, int lane, float sc2, float* m,
                                              float* l, float (*o)[4]) {
  const bool masked = sMask != nullptr || key0 + NT * 8 > a.Tk || (a.causal && key0 + NT * 8 - 1 > row_base);
  if (masked) fwd_block<NT, true>(a, qf, sK, sV, sMask, key0, row_base, lane, sc2, m, l, o);
  else fwd_block<NT, false>(a, qf, sK, sV, sMask, key0, row_base, lane, sc2, m, l, o);
}

__global__ void __launch_bounds__(256, 2) attention_fwd_tiled_kernel(const AttnArgs a, const TileGeom gm) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sK = sQ + gm.per_cta * 16 * LDS;
  bf16* sV = sK + gm.Tkp * LDS;
  int* sMask = reinterpret_cast<int*>(sV + gm.Tkp * LDS);
  const int bh = blockIdx.x / gm.nsplit, split = blockIdx.x % gm.nsplit;
  const int b = bh / a.H, h = bh % a.H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
  const int q0 = split * gm.per_cta * 16;
  load_rows_async(a.Q + ((long long)b * a.Tq) * a.ldq + h * HD, a.ldq, q0, a.Tq, gm.per_cta * 16, sQ, tid, nthr);
  load_rows_async(a.K + ((long long)b * a.Tk) * a.ldk + h * HD, a.ldk, 0, a.Tk, gm.Tkp, sK, tid, nthr);
  load_rows_async(a.V + ((long long)b * a.Tk) * a.ldv + h * HD, a.ldv, 0, a.Tk, gm.Tkp, sV, tid, nthr);
  if (a.key_mask)
    for (int j = tid; j < gm.Tkp; j += nthr) sMask[j] = j < a.Tk ? a.key_mask[(long long)b * a.Tk + j] : 0;
  cp_async_wait_all();
  __syncthreads();
  const int row_base = q0 + warp * 16;
  if (row_base >= a.Tq) return;
  const int g = lane >> 2, t = lane & 3;
  bf16* qtile = sQ + warp * 16 * LDS;
  uint32_t qf[4][4];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) lda(qtile, 0, kk * 16, lane, qf[kk]);
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  float o[8][4];
#pragma unroll
  for (int nd = 0; nd < 8; ++nd)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[nd][j] = 0.f;
  const float sc2 = a.scale * LOG2E;
  const int* km = a.key_mask ? sMask : nullptr;
  int kend = gm.Tkp;
  if (a.causal) kend = min(kend, row_base + 16);   // keys beyond the tile's last query are all masked
  int key0 = 0;
  for (; key0 + 64 <= kend; key0 += 64) fwd_block_any<8>(a, qf, sK, sV, km, key0, row_base, lane, sc2, m, l, o);
  if (key0 + 32 <= kend) { fwd_block_any<4>(a, qf, sK, sV, km, key0, row_base, lane, sc2, m, l, o); key0 += 32; }
  if (key0 + 16 <= kend) { fwd_block_any<2>(a, qf, sK, sV, km, key0, row_base, lane, sc2, m, l, o); key0 += 16; }
  float inv[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    l[i] += __shfl_xor_sync(0xffffffffu, l[i], 1);
    l[i] += __shfl_xor_sync(0xffffffffu, l[i], 2);
    inv[i] = l[i] > 0.f ? 1.0f / l[i] : 0.f;
    const int row = row_base + g + i * 8;
    if (a.lse && t == 0 && row < a.Tq)
      a.lse[((long long)b * a.H + h) * a.Tq + row] = ((m[i] == -INFINITY ? 0.f : m[i] * sc2) + log2f(l[i])) * LN2;
  }
  store_tile_rows(o, inv, qtile, a.O + ((long long)b * a.Tq) * a.ldo + h * HD, a.ldo, row_base, a.Tq, lane);
}

// one block of NT*8 keys of the dQ pass for a 16-query tile.  dS is formed without its scale factor (applied once to
// dQ at the end).  Padding needs no predicate here: a padded key has K = V = 0 (its dS multiplies a zero K row), a
// padded query has Q = dO = 0, lse = D = 0 (dS = 1 * (0 - 0)); MASKED is only for the causal diagonal / key padding masks.
template <int NT, bool MASKED>
__device__ __forceinline__ void dq_block(const AttnArgs& a, const uint32_t (*qf)[4], const uint32_t (*dof)[4], const bf16* sK,
                                         const bf16* sV, const int* sMask, int key0, int row_base, int lane, float sc2,
                                         const float* lse2, const float* dd, float (*dq)[4]) {
  const int g = lane >> 2, t = lane & 3;
  float s[NT][4], dp[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s[nt][j] = 0.f;
      dp[nt][j] = 0.f;
    }
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      uint32_t bk[4], bv[4];
      ldb2(sK, key0 + np * 16, kk * 16, lane, bk);
      ldb2(sV, key0 + np * 16, kk * 16, lane, bv);
      mma_bf16_16816(s[2 * np], qf[kk], bk);
      mma_bf16_16816(s[2 * np + 1], qf[kk], bk + 2);
      mma_bf16_16816(dp[2 * np], dof[kk], bv);
      mma_bf16_16816(dp[2 * np + 1], dof[kk], bv + 2);
    }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = j >> 1;
      float p = fast_exp2(fmaf(s[nt][j], sc2, -lse2[i]));
      if (MASKED) {
        const int row = row_base + g + i * 8, col = key0 + nt * 8 + 2 * t + (j & 1);
        if (!key_allowed(row, col, a.Tk, a.causal, sMask)) p = 0.f;
      }
      dp[nt][j] = p * (dp[nt][j] - dd[i]);
    }
#pragma unroll
  for (int kk = 0; kk < NT / 2; ++kk) {
    uint32_t af[4];
    af[0] = pack_bf16(dp[2 * kk][0], dp[2 * kk][1]);
    af[1] = pack_bf16(dp[2 * kk][2], dp[2 * kk][3]);
    af[2] = pack_bf16(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
    af[3] = pack_bf16(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
    for (int ndp = 0; ndp < 4; ++ndp) {
      uint32_t b[4];
      ldb2_t(sK, key0 + kk * 16, ndp * 16, lane, b);
      mma_bf16_16816(dq[2 * ndp], af, b);
      mma_bf16_16816(dq[2 * ndp + 1], af, b + 2);
    }
  }
}
template <int NT>
__device__ __forceinline__ void dq_block_any(const AttnArgs& a, const uint32_t (*qf)[4], const uint32_t (*dof)[4], const bf16* sK,
                                             const bf16* sV, const int* sMask, int key0, int row_base, int lane, float sc2,
                                             const float* lse2, const float* dd, float (*dq)[4]) {
  const bool masked = sMask != nullptr || (a.causal && key0 + NT * 8 - 1 > row_base);
  if (masked) dq_block<NT, true>(a, qf, dof, sK, sV, sMask, key0, row_base, lane, sc2, lse2, dd, dq);
  else dq_block<NT, false>(a, qf, dof, sK, sV, sMask, key0, row_base, lane, sc2, lse2, dd, dq);
}

// dQ of one 16-query tile (one warp): Q / dO / O fragments of the tile from shared memory, D = rowsum(dO o O) formed from
// the fragments themselves (returned in dd for rows g, g+8), all keys walked in 32-wide blocks; the result leaves through
// `stage` (a 16-row smem tile this warp owns and no longer needs)
__device__ __forceinline__ void dq_tile_pass(const AttnArgs& a, int b, int h, int row_base, const bf16* qtile, const bf16* dotile,
                                             const bf16* otile, const bf16* sK, const bf16* sV, const int* km, int Tkp,
                                             int lane, bf16* stage, float* dd) {
  const int g = lane >> 2;
  uint32_t qf[4][4], dof[4][4];
  dd[0] = dd[1] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t of[4];
    lda(qtile, 0, kk * 16, lane, qf[kk]);
    lda(dotile, 0, kk * 16, lane, dof[kk]);
    lda(otile, 0, kk * 16, lane, of);
#pragma unroll
    for (int r = 0; r < 4; ++r) {   // fragment registers 0,2 belong to row g, 1,3 to row g+8
      const float2 x = unpack_bf16(dof[kk][r]), y = unpack_bf16(of[r]);
      dd[r & 1] += x.x * y.x + x.y * y.y;
    }
  }
  float lse2[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    dd[i] += __shfl_xor_sync(0xffffffffu, dd[i], 1);
    dd[i] += __shfl_xor_sync(0xffffffffu, dd[i], 2);
    const int row = row_base + g + i * 8;
    lse2[i] = row < a.Tq ? a.lse[((long long)b * a.H + h) * a.Tq + row] * LOG2E : 0.f;
  }
  float dq[8][4];
#pragma unroll
  for (int nd = 0; nd < 8; ++nd)
#pragma unroll
    for (int j = 0; j < 4; ++j) dq[nd][j] = 0.f;
  const float sc2 = a.scale * LOG2E;
  int kend = Tkp;
  if (a.causal) kend = min(kend, row_base + 16);
  int key0 = 0;
  for (; key0 + 32 <= kend; key0 += 32) dq_block_any<4>(a, qf, dof, sK, sV, km, key0, row_base, lane, sc2, lse2, dd, dq);
  if (key0 + 16 <= kend) dq_block_any<2>(a, qf, dof, sK, sV, km, key0, row_base, lane, sc2, lse2, dd, dq);
  const float sc[2] = {a.scale, a.scale};
  store_tile_rows(dq, sc, stage, a.dQ + ((long long)b * a.Tq) * a.lddq + h * HD, a.lddq, row_base, a.Tq, lane);
}

__global__ void __launch_bounds__(256, 2) attention_bwd_dq_tiled_kernel(const AttnArgs a, const TileGeom gm) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int qrows = gm.per_cta * 16;
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sdO = sQ + qrows * LDS;
  bf16* sO = sdO + qrows * LDS;
  bf16* sK = sO + qrows * LDS;
  bf16* sV = sK + gm.Tkp * LDS;
  int* sMask = reinterpret_cast<int*>(sV + gm.Tkp * LDS);
  const int bh = blockIdx.x / gm.nsplit, split = blockIdx.x % gm.nsplit;
  const int b = bh / a.H, h = bh % a.H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
  const int q0 = split * qrows;
  load_rows_async(a.Q + ((long long)b * a.Tq) * a.ldq + h * HD, a.ldq, q0, a.Tq, qrows, sQ, tid, nthr);
  load_rows_async(a.dO + ((long long)b * a.Tq) * a.lddo + h * HD, a.lddo, q0, a.Tq, qrows, sdO, tid, nthr);
  load_rows_async(a.O + ((long long)b * a.Tq) * a.ldo + h * HD, a.ldo, q0, a.Tq, qrows, sO, tid, nthr);
  load_rows_async(a.K + ((long long)b * a.Tk) * a.ldk + h * HD, a.ldk, 0, a.Tk, gm.Tkp, sK, tid, nthr);
  load_rows_async(a.V + ((long long)b * a.Tk) * a.ldv + h * HD, a.ldv, 0, a.Tk, gm.Tkp, sV, tid, nthr);
  if (a.key_mask)
    for (int j = tid; j < gm.Tkp; j += nthr) sMask[j] = j < a.Tk ? a.key_mask[(long long)b * a.Tk + j] : 0;
  cp_async_wait_all();
  __syncthreads();
  const int row_base = q0 + warp * 16;
  if (row_base >= a.Tq) return;
  float dd[2];
  dq_tile_pass(a, b, h, row_base, sQ + warp * 16 * LDS, sdO + warp * 16 * LDS, sO + warp * 16 * LDS, sK, sV,
               a.key_mask ? sMask : nullptr, gm.Tkp, lane, sQ + warp * 16 * LDS, dd);
}

// one block of NT*8 queries of the dK/dV pass for a 16-key tile: S^T = K Q^T, dP^T = V dO^T (keys are the rows).
// Same conventions as dq_block: dS unscaled (dK scaled once at the end), padding needs no predicate.
template <int NT, bool MASKED>
__device__ __forceinline__ void dkv_block(const AttnArgs& a, const bf16* ktile, const bf16* vtile, const bf16* sQ,
                                          const bf16* sdO, const float* sLse2, const float* sD, const int* sMaskTile,
                                          int q0, int key_base, int lane, float sc2, float (*dk)[4], float (*dv)[4]) {
  const int g = lane >> 2, t = lane & 3;
  float s[NT][4], dp[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s[nt][j] = 0.f;
      dp[nt][j] = 0.f;
    }
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    // the K / V tile fragments are re-read per block rather than held (32 registers: the difference between one
    // and two resident CTAs per SM)
    uint32_t kf[4], vf[4];
    lda(ktile, 0, kk * 16, lane, kf);
    lda(vtile, 0, kk * 16, lane, vf);
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      uint32_t bq[4], bdo[4];
      ldb2(sQ, q0 + np * 16, kk * 16, lane, bq);
      ldb2(sdO, q0 + np * 16, kk * 16, lane, bdo);
      mma_bf16_16816(s[2 * np], kf, bq);
      mma_bf16_16816(s[2 * np + 1], kf, bq + 2);
      mma_bf16_16816(dp[2 * np], vf, bdo);
      mma_bf16_16816(dp[2 * np + 1], vf, bdo + 2);
    }
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int qc = q0 + nt * 8 + 2 * t;
    const float2 ls = *reinterpret_cast<const float2*>(sLse2 + qc);
    const float2 d2 = *reinterpret_cast<const float2*>(sD + qc);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float p = fast_exp2(fmaf(s[nt][j], sc2, -((j & 1) ? ls.y : ls.x)));
      if (MASKED) {
        const int q = qc + (j & 1), kl = g + (j >> 1) * 8;
        if ((sMaskTile && sMaskTile[kl] == 0) || (a.causal && key_base + kl > q)) p = 0.f;
      }
      s[nt][j] = p;
      dp[nt][j] = p * (dp[nt][j] - ((j & 1) ? d2.y : d2.x));
    }
  }
#pragma unroll
  for (int kk = 0; kk < NT / 2; ++kk) {
    uint32_t ap[4], ads[4];
    ap[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
    ap[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
    ap[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    ap[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
    ads[0] = pack_bf16(dp[2 * kk][0], dp[2 * kk][1]);
    ads[1] = pack_bf16(dp[2 * kk][2], dp[2 * kk][3]);
    ads[2] = pack_bf16(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
    ads[3] = pack_bf16(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
    for (int ndp = 0; ndp < 4; ++ndp) {
      uint32_t bdo[4], bq[4];
      ldb2_t(sdO, q0 + kk * 16, ndp * 16, lane, bdo);
      ldb2_t(sQ, q0 + kk * 16, ndp * 16, lane, bq);
      mma_bf16_16816(dv[2 * ndp], ap, bdo);
      mma_bf16_16816(dv[2 * ndp + 1], ap, bdo + 2);
      mma_bf16_16816(dk[2 * ndp], ads, bq);
      mma_bf16_16816(dk[2 * ndp + 1], ads, bq + 2);
    }
  }
}

template <int NT>
__device__ __forceinline__ void dkv_block_any(const AttnArgs& a, const bf16* ktile, const bf16* vtile, const bf16* sQ,
                                              const bf16* sdO, const float* sLse2, const float* sD, const int* sMaskTile,
                                              int q0, int key_base, int lane, float sc2, float (*dk)[4], float (*dv)[4]) {
  const bool masked = sMaskTile != nullptr || (a.causal && key_base + 15 > q0);
  if (masked) dkv_block<NT, true>(a, ktile, vtile, sQ, sdO, sLse2, sD, sMaskTile, q0, key_base, lane, sc2, dk, dv);
  else dkv_block<NT, false>(a, ktile, vtile, sQ, sdO, sLse2, sD, sMaskTile, q0, key_base, lane, sc2, dk, dv);
}

// dK / dV of one 16-key tile (one warp): all queries walked in 32-wide blocks; results leave through the warp's own
// K / V smem tiles
__device__ __forceinline__ void dkv_tile_pass(const AttnArgs& a, int b, int h, int key_base, bf16* ktile, bf16* vtile,
                                              const bf16* sQ, const bf16* sdO, const float* sLse2, const float* sD,
                                              const int* km, int Tqp, int lane) {
  float dk[8][4], dv[8][4];
#pragma unroll
  for (int nd = 0; nd < 8; ++nd)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      dk[nd][j] = 0.f;
      dv[nd][j] = 0.f;
    }
  const float sc2 = a.scale * LOG2E;
  int q0 = 0;
  if (a.causal) q0 = key_base & ~31;   // queries before the tile's first key see none of its keys
  for (; q0 + 32 <= Tqp; q0 += 32) dkv_block_any<4>(a, ktile, vtile, sQ, sdO, sLse2, sD, km, q0, key_base, lane, sc2, dk, dv);
  if (q0 + 16 <= Tqp) dkv_block_any<2>(a, ktile, vtile, sQ, sdO, sLse2, sD, km, q0, key_base, lane, sc2, dk, dv);
  const float one[2] = {1.f, 1.f}, sc[2] = {a.scale, a.scale};
  store_tile_rows(dk, sc, ktile, a.dK + ((long long)b * a.Tk) * a.lddk + h * HD, a.lddk, key_base, a.Tk, lane);
  store_tile_rows(dv, one, vtile, a.dV + ((long long)b * a.Tk) * a.lddv + h * HD, a.lddv, key_base, a.Tk, lane);
}

__global__ void __launch_bounds__(256, 2) attention_bwd_dkv_tiled_kernel(const AttnArgs a, const TileGeom gm) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int krows = gm.per_cta * 16;
  bf16* sK = reinterpret_cast<bf16*>(smem_raw);
  bf16* sV = sK + krows * LDS;
  bf16* sQ = sV + krows * LDS;
  bf16* sdO = sQ + gm.Tqp * LDS;
  float* sLse2 = reinterpret_cast<float*>(sdO + gm.Tqp * LDS);
  float* sD = sLse2 + gm.Tqp;
  int* sMask = reinterpret_cast<int*>(sD + gm.Tqp);
  const int bh = blockIdx.x / gm.nsplit, split = blockIdx.x % gm.nsplit;
  const int b = bh / a.H, h = bh % a.H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
  const int k0 = split * krows;
  const bf16* gdO = a.dO + ((long long)b * a.Tq) * a.lddo + h * HD;
  const bf16* gO = a.O + ((long long)b * a.Tq) * a.ldo + h * HD;
  load_rows_async(a.K + ((long long)b * a.Tk) * a.ldk + h * HD, a.ldk, k0, a.Tk, krows, sK, tid, nthr);
  load_rows_async(a.V + ((long long)b * a.Tk) * a.ldv + h * HD, a.ldv, k0, a.Tk, krows, sV, tid, nthr);
  load_rows_async(a.Q + ((long long)b * a.Tq) * a.ldq + h * HD, a.ldq, 0, a.Tq, gm.Tqp, sQ, tid, nthr);
  load_rows_async(gdO, a.lddo, 0, a.Tq, gm.Tqp, sdO, tid, nthr);
  // D[q] = sum_d dO[q][d] O[q][d]: 8 lanes per row, one 16-byte chunk each, straight from global while the copies
  // above are in flight; four rows per thread are loaded before any is reduced (one round trip, not four).
  {
    const int c = (tid & 7) * 8, rstep = nthr >> 3;
    for (int rb = 0; rb < gm.Tqp; rb += 4 * rstep) {   // trip count uniform over the CTA (full-mask shuffles below)
      const int r0 = rb + (tid >> 3);
      uint4 u[4], w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = r0 + k * rstep;
        u[k] = w[k] = make_uint4(0, 0, 0, 0);
        if (r < a.Tq) {
          u[k] = *reinterpret_cast<const uint4*>(gdO + (long long)r * a.lddo + c);
          w[k] = *reinterpret_cast<const uint4*>(gO + (long long)r * a.ldo + c);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = r0 + k * rstep;
        const uint32_t* uu = reinterpret_cast<const uint32_t*>(&u[k]);
        const uint32_t* ww = reinterpret_cast<const uint32_t*>(&w[k]);
        float d = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 x = unpack_bf16(uu[j]), y = unpack_bf16(ww[j]);
          d += x.x * y.x + x.y * y.y;
        }
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        if ((tid & 7) == 0 && r < gm.Tqp) sD[r] = d;
      }
    }
  }
  for (int r = tid; r < gm.Tqp; r += nthr) sLse2[r] = r < a.Tq ? a.lse[((long long)b * a.H + h) * a.Tq + r] * LOG2E : 0.f;
  if (a.key_mask)
    for (int j = tid; j < krows; j += nthr) sMask[j] = k0 + j < a.Tk ? a.key_mask[(long long)b * a.Tk + k0 + j] : 0;
  cp_async_wait_all();
  __syncthreads();
  const int key_base = k0 + warp * 16;
  if (key_base >= a.Tk) return;
  dkv_tile_pass(a, b, h, key_base, sK + warp * 16 * LDS, sV + warp * 16 * LDS, sQ, sdO, sLse2, sD,
                a.key_mask ? sMask + warp * 16 : nullptr, gm.Tqp, lane);
}

// Single-kernel backward (one CTA per (batch, head) covers every tile of both passes; measured against the two-kernel
// form: 119 -> 82 us at 64 tokens, 509 -> 499 us at 197 tokens, profiles/r02_attention_ab.txt): warp w first
// forms dQ of query tile w, then dK / dV of key tile w, from ONE shared-memory copy of Q, K, V, dO, O - the two-kernel
// form reads those five tensors twice.  D = rowsum(dO o O) comes out of the dQ pass and crosses to the dK/dV pass
// through shared memory (the only CTA barrier after the load).
__device__ __forceinline__ void attention_bwd_fused_body(const AttnArgs& a, const TileGeom& gm) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sdO = sQ + gm.Tqp * LDS;
  bf16* sO = sdO + gm.Tqp * LDS;
  bf16* sK = sO + gm.Tqp * LDS;
  bf16* sV = sK + gm.Tkp * LDS;
  float* sLse2 = reinterpret_cast<float*>(sV + gm.Tkp * LDS);
  float* sD = sLse2 + gm.Tqp;
  int* sMask = reinterpret_cast<int*>(sD + gm.Tqp);
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
  load_rows_async(a.Q + ((long long)b * a.Tq) * a.ldq + h * HD, a.ldq, 0, a.Tq, gm.Tqp, sQ, tid, nthr);
  load_rows_async(a.dO + ((long long)b * a.Tq) * a.lddo + h * HD, a.lddo, 0, a.Tq, gm.Tqp, sdO, tid, nthr);
  load_rows_async(a.O + ((long long)b * a.Tq) * a.ldo + h * HD, a.ldo, 0, a.Tq, gm.Tqp, sO, tid, nthr);
  load_rows_async(a.K + ((long long)b * a.Tk) * a.ldk + h * HD, a.ldk, 0, a.Tk, gm.Tkp, sK, tid, nthr);
  load_rows_async(a.V + ((long long)b * a.Tk) * a.ldv + h * HD, a.ldv, 0, a.Tk, gm.Tkp, sV, tid, nthr);
  for (int r = tid; r < gm.Tqp; r += nthr) sLse2[r] = r < a.Tq ? a.lse[((long long)b * a.H + h) * a.Tq + r] * LOG2E : 0.f;
  if (a.key_mask)
    for (int j = tid; j < gm.Tkp; j += nthr) sMask[j] = j < a.Tk ? a.key_mask[(long long)b * a.Tk + j] : 0;
  cp_async_wait_all();
  __syncthreads();
  const int base = warp * 16;
  if (base < a.Tq) {
    float dd[2];
    dq_tile_pass(a, b, h, base, sQ + base * LDS, sdO + base * LDS, sO + base * LDS, sK, sV, a.key_mask ? sMask : nullptr,
                 gm.Tkp, lane, sO + base * LDS, dd);      // staged through the warp's O rows: Q stays for the second pass
    if ((lane & 3) == 0) {
      sD[base + (lane >> 2)] = dd[0];
      sD[base + (lane >> 2) + 8] = dd[1];
    }
  }
  __syncthreads();      // D complete; every warp is done reading K / V as dQ operands
  if (base < a.Tk)
    dkv_tile_pass(a, b, h, base, sK + base * LDS, sV + base * LDS, sQ, sdO, sLse2, sD, a.key_mask ? sMask + base : nullptr,
                  gm.Tqp, lane);
}
__global__ void __launch_bounds__(256, 2) attention_bwd_fused_tiled_kernel(const AttnArgs a, const TileGeom gm) {
  attention_bwd_fused_body(a, gm);
}
// up to 16 tiles per sequence (<= 256 tokens: ViT-B/16's 197 -> 13 warps), one CTA per SM
__global__ void __launch_bounds__(512, 1) attention_bwd_fused_tiled_wide_kernel(const AttnArgs a, const TileGeom gm) {
  attention_bwd_fused_body(a, gm);
}

// Cached decode attention: one warp per (row, head); body shared with the persistent decoder-step kernel
// (decode_device.cuh).
__global__ void __launch_bounds__(128) decode_attention_kernel(const DecAttnArgs a) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_p[4][micdec::DEC_MAX_KEYS];
  __shared__ int s_row[4][micdec::DEC_MAX_KEYS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int wid = blockIdx.x * 4 + w;
  if (wid >= a.R * a.H) return;
  micdec::decode_attn_item(a, wid / a.H, wid % a.H, s_p[w], s_row[w], lane);
}

}  // namespace

#define STREAM reinterpret_cast<cudaStream_t>(stream)

static int check_attn(int head_dim, int Tq, int Tk) {
  MIC_CHECK_ARG(head_dim == HD, "attention: head_dim %d != 64", head_dim);
  MIC_CHECK_ARG(Tq >= 1 && Tq <= GEN_MAX_T && Tk >= 1 && Tk <= GEN_MAX_T, "attention: Tq=%d Tk=%d must be in [1,256]", Tq,
                Tk);
  return MIC_OK;
}
static int g_attn_impl = 0;   // 0: row-tiled kernels, single-kernel backward; 1: one-CTA-per-head kernels where they apply
                              // (<= 64 tokens); 2: row-tiled with the backward as a dQ and a dK/dV kernel
extern "C" int mic_attention_impl(int impl) {
  MIC_CHECK_ARG(impl >= 0 && impl <= 2, "attention impl %d not in [0,2]", impl);
  g_attn_impl = impl;
  return MIC_OK;
}
static bool use_tiled(int Tq, int Tk) { return !(g_attn_impl == 1 && Tq <= TMAX && Tk <= TMAX); }
// warps tile `rows` in 16-row tiles, at most 8 per CTA, spread evenly over the CTAs of a (batch, head)
static TileGeom tile_geom(int Tq, int Tk, int rows) {
  TileGeom gm;
  const int tiles = (rows + 15) / 16;
  gm.nsplit = (tiles + 7) / 8;
  gm.per_cta = (tiles + gm.nsplit - 1) / gm.nsplit;
  gm.Tqp = (Tq + 15) / 16 * 16;
  gm.Tkp = (Tk + 15) / 16 * 16;
  return gm;
}
template <typename K>
static int ensure_smem(K kern, int bytes, int* have) {
  if (bytes > *have) {
    MIC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    // two CTAs per SM need the full shared-memory carve-out; the driver's default heuristic leaves room for one
    MIC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    *have = bytes;
  }
  return MIC_OK;
}
extern "C" int mic_attention_fwd(void* stream, const void* Q, long long ldq, const void* K, long long ldk,
                                 const void* V, long long ldv, void* O, long long ldo, float* lse,
                                 const int* key_mask, int causal, int B, int H, int Tq, int Tk, int head_dim,
                                 float scale) {
  int rc = check_attn(head_dim, Tq, Tk);
  if (rc) return rc;
  AttnArgs a = {};
  a.Q = (const bf16*)Q; a.K = (const bf16*)K; a.V = (const bf16*)V;
  a.ldq = ldq; a.ldk = ldk; a.ldv = ldv;
  a.O = (bf16*)O; a.ldo = ldo; a.lse = lse; a.key_mask = key_mask;
  a.causal = causal; a.B = B; a.H = H; a.Tq = Tq; a.Tk = Tk; a.scale = scale;
  if (use_tiled(Tq, Tk)) {
    const TileGeom gm = tile_geom(Tq, Tk, Tq);
    const int smem = (gm.per_cta * 16 + 2 * gm.Tkp) * LDS * 2 + gm.Tkp * 4;
    static int have = 0;
    rc = ensure_smem(attention_fwd_tiled_kernel, smem, &have);
    if (rc) return rc;
    attention_fwd_tiled_kernel<<<B * H * gm.nsplit, gm.per_cta * 32, smem, STREAM>>>(a, gm);
  } else {
    attention_fwd_kernel<<<B * H, 128, 0, STREAM>>>(a);
  }
  MIC_CHECK_LAUNCH();
  return MIC_OK;
}

extern "C" int mic_attention_bwd(void* stream, const void* Q, long long ldq, const void* K, long long ldk,
                                 const void* V, long long ldv, const void* O, long long ldo, const void* dO,
                                 long long lddo, const float* lse, const int* key_mask, int causal, void* dQ,
                                 long long lddq, void* dK, long long lddk, void* dV, long long lddv, int B, int H,
                                 int Tq, int Tk, int head_dim, float scale) {
  int rc = check_attn(head_dim, Tq, Tk);
  if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    MIC_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    attr = true;
  }
  AttnArgs a = {};
  a.Q = (const bf16*)Q; a.K = (const bf16*)K; a.V = (const bf16*)V;
  a.ldq = ldq; a.ldk = ldk; a.ldv = ldv;
  a.O = (bf16*)const_cast<void*>(O); a.ldo = ldo; a.lse = const_cast<float*>(lse); a.key_mask = key_mask;
  a.causal = causal; a.B = B; a.H = H; a.Tq = Tq; a.Tk = Tk; a.scale = scale;
  a.dO = (const bf16*)dO; a.lddo = lddo;
  a.dQ = (bf16*)dQ; a.dK = (bf16*)dK; a.dV = (bf16*)dV;
  a.lddq = lddq; a.lddk = lddk; a.lddv = lddv;
  if (use_tiled(Tq, Tk) && g_attn_impl != 2) {
    TileGeom gm = tile_geom(Tq, Tk, Tq > Tk ? Tq : Tk);      // one CTA per (batch, head): warps = max(query, key tiles)
    gm.nsplit = 1;
    gm.per_cta = ((Tq > Tk ? Tq : Tk) + 15) / 16;
    const int smem = (3 * gm.Tqp + 2 * gm.Tkp) * LDS * 2 + 2 * gm.Tqp * 4 + gm.Tkp * 4;
    if (gm.per_cta <= 8) {
      static int have = 0;
      rc = ensure_smem(attention_bwd_fused_tiled_kernel, smem, &have);
      if (rc) return rc;
      attention_bwd_fused_tiled_kernel<<<B * H, gm.per_cta * 32, smem, STREAM>>>(a, gm);
    } else {
      static int have_w = 0;
      rc = ensure_smem(attention_bwd_fused_tiled_wide_kernel, smem, &have_w);
      if (rc) return rc;
      attention_bwd_fused_tiled_wide_kernel<<<B * H, gm.per_cta * 32, smem, STREAM>>>(a, gm);
    }
  } else if (use_tiled(Tq, Tk)) {
    const TileGeom gq = tile_geom(Tq, Tk, Tq), gk = tile_geom(Tq, Tk, Tk);
    const int smem_dq = (3 * gq.per_cta * 16 + 2 * gq.Tkp) * LDS * 2 + gq.Tkp * 4;
    const int smem_dkv = (2 * gk.per_cta * 16 + 2 * gk.Tqp) * LDS * 2 + 2 * gk.Tqp * 4 + gk.per_cta * 16 * 4;
    static int have_dq = 0, have_dkv = 0;
    rc = ensure_smem(attention_bwd_dq_tiled_kernel, smem_dq, &have_dq);
    if (rc) return rc;
    rc = ensure_smem(attention_bwd_dkv_tiled_kernel, smem_dkv, &have_dkv);
    if (rc) return rc;
    attention_bwd_dq_tiled_kernel<<<B * H * gq.nsplit, gq.per_cta * 32, smem_dq, STREAM>>>(a, gq);
    MIC_CHECK_LAUNCH();
    attention_bwd_dkv_tiled_kernel<<<B * H * gk.nsplit, gk.per_cta * 32, smem_dkv, STREAM>>>(a, gk);
  } else {
    attention_bwd_kernel<<<B * H, 128, BWD_SMEM, STREAM>>>(a);
  }
  MIC_CHECK_LAUNCH();
  return MIC_OK;
}

extern "C" int mic_decode_attention(void* stream, const void* q, long long ldq, const void* k_cache,
                                    const void* v_cache, long long ldkv, const int* ancestors, int cache_len,
                                    int n_keys, int rows_per_kv, void* o, long long ldo, int R, int H, int head_dim,
                                    float scale) {
  MIC_CHECK_ARG(head_dim == HD, "decode attention: head_dim %d != 64", head_dim);
  MIC_CHECK_ARG(n_keys >= 1 && n_keys <= micdec::DEC_MAX_KEYS && n_keys <= cache_len, "decode attention: n_keys=%d (max 256) cache_len=%d",
                n_keys, cache_len);
  DecAttnArgs a;
  a.q = (const bf16*)q; a.ldq = ldq; a.kc = (const bf16*)k_cache; a.vc = (const bf16*)v_cache; a.ldkv = ldkv;
  a.anc = ancestors; a.T = cache_len; a.n_keys = n_keys; a.rows_per_kv = rows_per_kv < 1 ? 1 : rows_per_kv;
  a.o = (bf16*)o; a.ldo = ldo; a.R = R; a.H = H; a.scale = scale; a.q_acc = nullptr; a.q_bias = nullptr; a.o_tiled_kb = 0;
  MIC_CHECK_CUDA(mic_launch(decode_attention_kernel, dim3((R * H + 3) / 4), dim3(128), 0, STREAM, a));
  return MIC_OK;
}
