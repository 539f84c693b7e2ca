// Attention for the tiny sequence lengths of the captioning path (50 visual tokens, 64 text tokens,
// head_dim 64): one CTA per (batch, head) keeps the whole Q/K/V tile of that head in shared memory.
// S=QK^T, softmax and PV run on warp-level bf16 MMA (m16n8k16) with fp32 accumulation; these
// 64x64x64 problems are latency-bound, far too small to amortise a TMEM round trip, so the legacy
// tensor path is the right tool here (they are ~1.3% of the step's FLOPs).
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

// ---------------------------------------------------------------------------------------------
// General path for longer sequences (ViT-B/16: 197 tokens; BART cross-attention over 197 keys):
//   forward : grid (B*H, ceil(Tq/64)); a CTA owns 64 queries and ALL keys (<= 256) in shared memory
//   backward: dQ kernel  grid (B*H, query blocks)  loops over 64-key blocks (no online softmax needed: the
//             forward saved the log-sum-exp), dK/dV kernel grid (B*H, key blocks) loops over query blocks.
// ---------------------------------------------------------------------------------------------
constexpr int GEN_MAX_T = 256;

__device__ __forceinline__ void load_rows(const bf16* g, long long ld, int row0, int rows_valid, int nrows, bf16* s,
                                          int tid, int nthreads) {
  // rows [row0, row0+nrows) of a head slice -> smem rows [0, nrows); rows >= rows_valid are zero
  for (int i = tid; i < nrows * (HD / 8); i += nthreads) {
    const int r = i >> 3, c = (i & 7) * 8;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (row0 + r < rows_valid) u = *reinterpret_cast<const uint4*>(g + (long long)(row0 + r) * ld + c);
    *reinterpret_cast<uint4*>(s + r * LDS + c) = u;
  }
}

template <int NKB>   // number of 64-key blocks held in registers per query row block
__global__ void __launch_bounds__(128) attention_fwd_gen_kernel(const AttnArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sK = sQ + 64 * LDS;
  bf16* sV = sK + NKB * 64 * LDS;
  int* sMask = reinterpret_cast<int*>(sV + NKB * 64 * LDS);
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H, qb = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  load_rows(a.Q + ((long long)b * a.Tq) * a.ldq + h * HD, a.ldq, qb * 64, a.Tq, 64, sQ, tid, 128);
  load_rows(a.K + ((long long)b * a.Tk) * a.ldk + h * HD, a.ldk, 0, a.Tk, NKB * 64, sK, tid, 128);
  load_rows(a.V + ((long long)b * a.Tk) * a.ldv + h * HD, a.ldv, 0, a.Tk, NKB * 64, sV, tid, 128);
  for (int j = tid; j < NKB * 64; j += 128) sMask[j] = (a.key_mask && j < a.Tk) ? a.key_mask[(long long)b * a.Tk + j] : 1;
  __syncthreads();
  const int r0 = warp * 16;
  if (qb * 64 + r0 >= a.Tq) return;
  const int g = lane >> 2, t = lane & 3;
  float s[NKB * 8][4];
#pragma unroll
  for (int nt = 0; nt < NKB * 8; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) s[nt][j] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t af[4];
    frag_a(sQ, r0, kk * 16, lane, af);
#pragma unroll
    for (int nt = 0; nt < NKB * 8; ++nt) {
      uint32_t bfr[2];
      frag_b(sK, nt * 8, kk * 16, lane, bfr);
      mma_bf16_16816(s[nt], af, bfr);
    }
  }
  float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int nt = 0; nt < NKB * 8; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = qb * 64 + r0 + g + (j >> 1) * 8, col = nt * 8 + 2 * t + (j & 1);
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
  for (int nt = 0; nt < NKB * 8; ++nt)
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
  for (int kk = 0; kk < NKB * 4; ++kk) {
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
    const int row = qb * 64 + r0 + g + i * 8;
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

// D[q] = sum_d dO[q][d] * O[q][d] for rows [row0, row0+nrows) -> sD[0..nrows)
__device__ __forceinline__ void rowdot_dO_O(const AttnArgs& a, int b, int h, int row0, int nrows, float* sD, int tid,
                                            int nthreads) {
  const bf16* gdO = a.dO + ((long long)b * a.Tq) * a.lddo + h * HD;
  const bf16* gO = a.O + ((long long)b * a.Tq) * a.ldo + h * HD;
  for (int r = tid; r < nrows; r += nthreads) {
    float d = 0.f;
    if (row0 + r < a.Tq) {
#pragma unroll
      for (int c = 0; c < HD; c += 8) {
        const uint4 u = *reinterpret_cast<const uint4*>(gdO + (long long)(row0 + r) * a.lddo + c);
        const uint4 w = *reinterpret_cast<const uint4*>(gO + (long long)(row0 + r) * a.ldo + c);
        const uint32_t* uu = reinterpret_cast<const uint32_t*>(&u);
        const uint32_t* ww = reinterpret_cast<const uint32_t*>(&w);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 x = unpack_bf16(uu[j]), y = unpack_bf16(ww[j]);
          d += x.x * y.x + x.y * y.y;
        }
      }
    }
    sD[r] = d;
  }
}

// P and dS for a 16-query x 64-key block: s <- P (fp32), dp <- dS (fp32)
__device__ __forceinline__ void p_ds_block(const AttnArgs& a, float (*s)[4], float (*dp)[4], int q_row0, int key0, int g,
                                           int t, const float* lse, const float* dd, const int* sMask) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = j >> 1;
      const int row = q_row0 + g + i * 8, col = key0 + nt * 8 + 2 * t + (j & 1);
      const bool ok = row < a.Tq && key_allowed(row, col, a.Tk, a.causal, sMask);
      const float p = ok ? __expf(s[nt][j] * a.scale - lse[i]) : 0.f;
      dp[nt][j] = p * (dp[nt][j] - dd[i]) * a.scale;
      s[nt][j] = p;
    }
}

__global__ void __launch_bounds__(128) attention_bwd_dq_kernel(const AttnArgs a, int nkb) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_raw);
  bf16* sdO = sQ + 64 * LDS;
  bf16* sK = sdO + 64 * LDS;
  bf16* sV = sK + nkb * 64 * LDS;
  float* sD = reinterpret_cast<float*>(sV + nkb * 64 * LDS);
  int* sMask = reinterpret_cast<int*>(sD + 64);
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H, qb = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  load_rows(a.Q + ((long long)b * a.Tq) * a.ldq + h * HD, a.ldq, qb * 64, a.Tq, 64, sQ, tid, 128);
  load_rows(a.dO + ((long long)b * a.Tq) * a.lddo + h * HD, a.lddo, qb * 64, a.Tq, 64, sdO, tid, 128);
  load_rows(a.K + ((long long)b * a.Tk) * a.ldk + h * HD, a.ldk, 0, a.Tk, nkb * 64, sK, tid, 128);
  load_rows(a.V + ((long long)b * a.Tk) * a.ldv + h * HD, a.ldv, 0, a.Tk, nkb * 64, sV, tid, 128);
  for (int j = tid; j < nkb * 64; j += 128) sMask[j] = (a.key_mask && j < a.Tk) ? a.key_mask[(long long)b * a.Tk + j] : 1;
  rowdot_dO_O(a, b, h, qb * 64, 64, sD, tid, 128);
  __syncthreads();
  const int r0 = warp * 16;
  if (qb * 64 + r0 >= a.Tq) return;
  const int g = lane >> 2, t = lane & 3;
  float lse[2], dd[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int row = qb * 64 + r0 + g + i * 8;
    lse[i] = row < a.Tq ? a.lse[((long long)b * a.H + h) * a.Tq + row] : 0.f;
    dd[i] = sD[r0 + g + i * 8];
  }
  float dq[8][4];
#pragma unroll
  for (int nd = 0; nd < 8; ++nd)
#pragma unroll
    for (int j = 0; j < 4; ++j) dq[nd][j] = 0.f;
  for (int kb = 0; kb < nkb; ++kb) {
    const bf16* kblk = sK + kb * 64 * LDS;
    const bf16* vblk = sV + kb * 64 * LDS;
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
        frag_b(kblk, nt * 8, kk * 16, lane, bk);
        frag_b(vblk, nt * 8, kk * 16, lane, bv);
        mma_bf16_16816(s[nt], aq, bk);
        mma_bf16_16816(dp[nt], ado, bv);
      }
    }
    p_ds_block(a, s, dp, qb * 64 + r0, kb * 64, g, t, lse, dd, a.key_mask ? sMask : nullptr);
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
        frag_b_trans(kblk, kk * 16, nd * 8, lane, bfr);
        mma_bf16_16816(dq[nd], af, bfr);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int row = qb * 64 + r0 + g + i * 8;
    if (row < a.Tq) {
      bf16* drow = a.dQ + ((long long)b * a.Tq + row) * a.lddq + h * HD;
#pragma unroll
      for (int nd = 0; nd < 8; ++nd)
        *reinterpret_cast<uint32_t*>(drow + nd * 8 + 2 * t) = pack_bf16(dq[nd][2 * i], dq[nd][2 * i + 1]);
    }
  }
}

__global__ void __launch_bounds__(128) attention_bwd_dkv_kernel(const AttnArgs a, int nqb) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sK = reinterpret_cast<bf16*>(smem_raw);
  bf16* sV = sK + 64 * LDS;
  bf16* sP = sV + 64 * LDS;
  bf16* sdS = sP + 64 * LDS;
  bf16* sQ = sdS + 64 * LDS;
  bf16* sdO = sQ + nqb * 64 * LDS;
  float* sD = reinterpret_cast<float*>(sdO + nqb * 64 * LDS);
  float* sLse = sD + nqb * 64;
  int* sMask = reinterpret_cast<int*>(sLse + nqb * 64);
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H, kb = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  load_rows(a.K + ((long long)b * a.Tk) * a.ldk + h * HD, a.ldk, kb * 64, a.Tk, 64, sK, tid, 128);
  load_rows(a.V + ((long long)b * a.Tk) * a.ldv + h * HD, a.ldv, kb * 64, a.Tk, 64, sV, tid, 128);
  load_rows(a.Q + ((long long)b * a.Tq) * a.ldq + h * HD, a.ldq, 0, a.Tq, nqb * 64, sQ, tid, 128);
  load_rows(a.dO + ((long long)b * a.Tq) * a.lddo + h * HD, a.lddo, 0, a.Tq, nqb * 64, sdO, tid, 128);
  rowdot_dO_O(a, b, h, 0, nqb * 64, sD, tid, 128);
  for (int r = tid; r < nqb * 64; r += 128) sLse[r] = r < a.Tq ? a.lse[((long long)b * a.H + h) * a.Tq + r] : 0.f;
  if (tid < 64) sMask[tid] = (a.key_mask && kb * 64 + tid < a.Tk) ? a.key_mask[(long long)b * a.Tk + kb * 64 + tid] : 1;
  __syncthreads();
  const int r0 = warp * 16;
  const int g = lane >> 2, t = lane & 3;
  float dv[8][4], dk[8][4];
#pragma unroll
  for (int nd = 0; nd < 8; ++nd)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      dv[nd][j] = 0.f;
      dk[nd][j] = 0.f;
    }
  for (int qb = 0; qb < nqb; ++qb) {
    const bf16* qblk = sQ + qb * 64 * LDS;
    const bf16* doblk = sdO + qb * 64 * LDS;
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
        frag_a(qblk, r0, kk * 16, lane, aq);
        frag_a(doblk, r0, kk * 16, lane, ado);
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
        lse[i] = sLse[qb * 64 + r0 + g + i * 8];
        dd[i] = sD[qb * 64 + r0 + g + i * 8];
      }
      // key index inside sMask is block-local: shift the mask pointer so that key_allowed(col) indexes col - kb*64
      p_ds_block(a, s, dp, qb * 64 + r0, kb * 64, g, t, lse, dd, a.key_mask ? (sMask - kb * 64) : nullptr);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int row = r0 + g + i * 8;
          *reinterpret_cast<uint32_t*>(sP + row * LDS + nt * 8 + 2 * t) = pack_bf16(s[nt][2 * i], s[nt][2 * i + 1]);
          *reinterpret_cast<uint32_t*>(sdS + row * LDS + nt * 8 + 2 * t) = pack_bf16(dp[nt][2 * i], dp[nt][2 * i + 1]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
      uint32_t ap[4], ads[4];
      frag_a_trans(sP, r0, qq * 16, lane, ap);
      frag_a_trans(sdS, r0, qq * 16, lane, ads);
#pragma unroll
      for (int nd = 0; nd < 8; ++nd) {
        uint32_t bdo[2], bq[2];
        frag_b_trans(doblk, qq * 16, nd * 8, lane, bdo);
        frag_b_trans(qblk, qq * 16, nd * 8, lane, bq);
        mma_bf16_16816(dv[nd], ap, bdo);
        mma_bf16_16816(dk[nd], ads, bq);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int key = kb * 64 + r0 + g + i * 8;
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
template <typename K>
static int set_smem(K kern, int bytes) {
  MIC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
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
  if (Tq <= TMAX && Tk <= TMAX) {
    attention_fwd_kernel<<<B * H, 128, 0, STREAM>>>(a);
  } else {
    const int nkb = (Tk + 63) / 64;
    dim3 grid(B * H, (Tq + 63) / 64);
    if (nkb <= 2) {
      const int smem = (64 + 2 * 2 * 64) * LDS * 2 + 2 * 64 * 4;
      static bool attr2 = false;
      if (!attr2) { rc = set_smem(attention_fwd_gen_kernel<2>, smem); if (rc) return rc; attr2 = true; }
      attention_fwd_gen_kernel<2><<<grid, 128, smem, STREAM>>>(a);
    } else {
      const int smem = (64 + 2 * 4 * 64) * LDS * 2 + 4 * 64 * 4;
      static bool attr4 = false;
      if (!attr4) { rc = set_smem(attention_fwd_gen_kernel<4>, smem); if (rc) return rc; attr4 = true; }
      attention_fwd_gen_kernel<4><<<grid, 128, smem, STREAM>>>(a);
    }
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
  if (Tq <= TMAX && Tk <= TMAX) {
    attention_bwd_kernel<<<B * H, 128, BWD_SMEM, STREAM>>>(a);
  } else {
    const int nkb = (Tk + 63) / 64, nqb = (Tq + 63) / 64;
    const int smem_dq = (2 * 64 + 2 * nkb * 64) * LDS * 2 + 64 * 4 + nkb * 64 * 4;
    const int smem_dkv = (4 * 64 + 2 * nqb * 64) * LDS * 2 + 2 * nqb * 64 * 4 + 64 * 4;
    static int set_dq = 0, set_dkv = 0;
    if (smem_dq > set_dq) { rc = set_smem(attention_bwd_dq_kernel, smem_dq); if (rc) return rc; set_dq = smem_dq; }
    if (smem_dkv > set_dkv) { rc = set_smem(attention_bwd_dkv_kernel, smem_dkv); if (rc) return rc; set_dkv = smem_dkv; }
    attention_bwd_dq_kernel<<<dim3(B * H, nqb), 128, smem_dq, STREAM>>>(a, nkb);
    MIC_CHECK_LAUNCH();
    attention_bwd_dkv_kernel<<<dim3(B * H, nkb), 128, smem_dkv, STREAM>>>(a, nqb);
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
