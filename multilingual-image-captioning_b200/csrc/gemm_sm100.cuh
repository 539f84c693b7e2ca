// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma
// (cta_group::1, M=128, N=BN) -> fp32 accumulators in TMEM (double buffered) -> epilogue warps.
//
//   D[M,N] = epilogue( A[M,K] * B[N,K]^T )
//
// Either operand may be K-major (stored [rows, K], K contiguous) or MN-major (stored [K, rows], rows
// contiguous); that covers forward (A K-major, Flax `kernel (in,out)` = MN-major B), dgrad (K,K) and
// wgrad (MN,MN) without any transposed copies.  The epilogue is a policy class so the same mainloop
// serves the plain store (+bias/activation/residual), the fused lm_head + log-softmax/CE statistics,
// the CE backward (dlogits) and the beam-search candidate selection.
//
// Warp roles (384 threads): w0 TMA producer, w1 MMA issuer, w2 TMEM allocator, w3 spare,
// w4..w11 epilogue (two warps per TMEM lane quarter, each takes half of the tile's columns).
#pragma once

#include "common.cuh"

namespace micgemm {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + NUM_EPI_WARPS * 32;
constexpr int TMEM_COLS = 512;

template <int BN>
struct Cfg {
  static_assert(BN == 128 || BN == 192 || BN == 256, "unsupported BLOCK_N");
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BN * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BN == 256 ? 4 : (BN == 192 ? 5 : 6);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int COLS_PER_WARP = BN / 2;
};

struct Shape {
  int M, N, K;
  int num_m_blocks, num_n_blocks, group_m;
};

struct TileCoord {
  int m_blk, n_blk;
};

__device__ __forceinline__ TileCoord tile_coord(const Shape& s, int t) {
  const int per_group = s.group_m * s.num_n_blocks;
  const int g = t / per_group;
  const int first_m = g * s.group_m;
  const int gsz = min(s.group_m, s.num_m_blocks - first_m);
  const int r = t - g * per_group;
  TileCoord c;
  c.n_blk = r / gsz;
  c.m_blk = first_m + r % gsz;
  return c;
}

// --------------------------------------------------------------------------------------------
// Epilogue policy 0: store with optional bias / activation / residual / accumulate
// --------------------------------------------------------------------------------------------
struct EpiStoreParams {
  void* D;              // bf16 or fp32 [M, ldd]
  long long ldd;
  int d_f32;            // 1: fp32 output
  int accumulate;       // fp32 only: D += result
  int vec_ok;           // pointers/strides allow 16-byte vector access
  const float* bias;    // [N] fp32 or null
  int act;              // MIC_ACT_*
  bf16* D2;             // optional pre-activation copy (bf16, same ld as D) or null
  const bf16* residual; // optional bf16 [M, ldr] added after activation
  long long ldr;
  float out_scale;      // applied to (acc + bias) before activation (1.0 normally)
};

struct EpiStore {
  typedef EpiStoreParams Params;
  struct State {};
  __device__ static void tile_begin(const Params&, State&, const Shape&, int, int, int) {}
  __device__ static void tile_end(const Params&, State&, const Shape&, int, int, int, int) {}
  // v: 32 consecutive columns [col0, col0+32) of row `row`
  __device__ static void chunk(const Params& p, State&, const Shape& s, int row, int col0, float* v) {
    if (row >= s.M || col0 >= s.N) return;
    const bool full = (col0 + 32 <= s.N) && p.vec_ok;
    if (full) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int c = col0 + g * 8;
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = v[g * 8 + j] * p.out_scale;
        if (p.bias) {
          const float4 b0 = *reinterpret_cast<const float4*>(p.bias + c);
          const float4 b1 = *reinterpret_cast<const float4*>(p.bias + c + 4);
          x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
          x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
        }
        if (p.D2) {
          uint4 u = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]),
                               pack_bf16(x[6], x[7]));
          *reinterpret_cast<uint4*>(p.D2 + (long long)row * p.ldd + c) = u;
        }
        if (p.act != MIC_ACT_NONE) {
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = act_fwd(x[j], p.act);
        }
        if (p.residual) {
          const uint4 r = *reinterpret_cast<const uint4*>(p.residual + (long long)row * p.ldr + c);
          float2 f;
          f = unpack_bf16(r.x); x[0] += f.x; x[1] += f.y;
          f = unpack_bf16(r.y); x[2] += f.x; x[3] += f.y;
          f = unpack_bf16(r.z); x[4] += f.x; x[5] += f.y;
          f = unpack_bf16(r.w); x[6] += f.x; x[7] += f.y;
        }
        if (p.d_f32) {
          float* d = reinterpret_cast<float*>(p.D) + (long long)row * p.ldd + c;
          float4 o0 = make_float4(x[0], x[1], x[2], x[3]);
          float4 o1 = make_float4(x[4], x[5], x[6], x[7]);
          if (p.accumulate) {
            const float4 a0 = *reinterpret_cast<const float4*>(d);
            const float4 a1 = *reinterpret_cast<const float4*>(d + 4);
            o0.x += a0.x; o0.y += a0.y; o0.z += a0.z; o0.w += a0.w;
            o1.x += a1.x; o1.y += a1.y; o1.z += a1.z; o1.w += a1.w;
          }
          *reinterpret_cast<float4*>(d) = o0;
          *reinterpret_cast<float4*>(d + 4) = o1;
        } else {
          uint4 u = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]),
                               pack_bf16(x[6], x[7]));
          *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.D) + (long long)row * p.ldd + c) = u;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int c = col0 + j;
        if (c >= s.N) continue;
        float x = v[j] * p.out_scale;
        if (p.bias) x += p.bias[c];
        if (p.D2) p.D2[(long long)row * p.ldd + c] = __float2bfloat16_rn(x);
        x = act_fwd(x, p.act);
        if (p.residual) x += __bfloat162float(p.residual[(long long)row * p.ldr + c]);
        if (p.d_f32) {
          float* d = reinterpret_cast<float*>(p.D) + (long long)row * p.ldd + c;
          *d = p.accumulate ? (*d + x) : x;
        } else {
          reinterpret_cast<bf16*>(p.D)[(long long)row * p.ldd + c] = __float2bfloat16_rn(x);
        }
      }
    }
  }
};

// --------------------------------------------------------------------------------------------
// Epilogue policy 1: lm_head + log-softmax / label-smoothed CE statistics (no logits written)
//   per (row, column-half of an N tile): max, sum exp(z - max), sum z ; plus z[label]
// --------------------------------------------------------------------------------------------
struct EpiCEStatsParams {
  const float* bias;     // final_logits_bias [N] or null
  const int* labels;     // [M]
  float* pmax;           // [2*num_n_blocks, M]
  float* psum;           // [2*num_n_blocks, M]
  float* psumz;          // [2*num_n_blocks, M]
  float* zlabel;         // [M]
};

struct EpiCEStats {
  typedef EpiCEStatsParams Params;
  struct State {
    float mx, sm, sz;
    int label;
  };
  __device__ static void tile_begin(const Params& p, State& st, const Shape& s, int row, int, int) {
    st.mx = -INFINITY;
    st.sm = 0.f;
    st.sz = 0.f;
    st.label = (row < s.M) ? p.labels[row] : -1;
  }
  __device__ static void chunk(const Params& p, State& st, const Shape& s, int row, int col0, float* v) {
    if (row >= s.M || col0 >= s.N) return;
    float cmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int c = col0 + j;
      float z = v[j];
      if (c < s.N) {
        if (p.bias) z += p.bias[c];
        st.sz += z;
        if (c == st.label) p.zlabel[row] = z;
      } else {
        z = -INFINITY;
      }
      v[j] = z;
      cmax = fmaxf(cmax, z);
    }
    const float nm = fmaxf(st.mx, cmax);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) acc += __expf(v[j] - nm);
    st.sm = st.sm * __expf(st.mx - nm) + acc;
    st.mx = nm;
  }
  __device__ static void tile_end(const Params& p, State& st, const Shape& s, int row, int, int n_blk, int half) {
    if (row >= s.M) return;
    const long long o = (long long)(n_blk * 2 + half) * s.M + row;
    p.pmax[o] = st.mx;
    p.psum[o] = st.sm;
    p.psumz[o] = st.sz;
  }
};

// --------------------------------------------------------------------------------------------
// Epilogue policy 2: CE backward — recompute the logits tile and emit
//   dlogits = (softmax - soft_labels) * row_weight   as bf16 into [M, ldd] (ldd = padded vocab)
// --------------------------------------------------------------------------------------------
struct EpiCEGradParams {
  const float* bias;     // [N] or null
  const int* labels;     // [M]
  const float* lse;      // [M]
  const float* row_w;    // [M]  mask / sum(mask) (0 for padded targets)
  float conf, low;       // soft-label values
  bf16* dlogits;         // [M, ldd]
  long long ldd;         // multiple of 8; columns [N, ldd) are written as zero
};

struct EpiCEGrad {
  typedef EpiCEGradParams Params;
  struct State {
    float lse, w;
    int label;
  };
  __device__ static void tile_begin(const Params& p, State& st, const Shape& s, int row, int, int) {
    const bool ok = row < s.M;
    st.lse = ok ? p.lse[row] : 0.f;
    st.w = ok ? p.row_w[row] : 0.f;
    st.label = ok ? p.labels[row] : -1;
  }
  __device__ static void chunk(const Params& p, State& st, const Shape& s, int row, int col0, float* v) {
    if (row >= s.M || col0 >= p.ldd) return;
    bf16* out = p.dlogits + (long long)row * p.ldd + col0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = col0 + g * 8 + j;
        float z = v[g * 8 + j];
        float d = 0.f;
        if (c < s.N) {
          if (p.bias) z += p.bias[c];
          const float pr = __expf(z - st.lse);
          d = (pr - ((c == st.label) ? p.conf : p.low)) * st.w;
        }
        x[j] = d;
      }
      if (col0 + g * 8 + 8 <= p.ldd) {
        *reinterpret_cast<uint4*>(out + g * 8) = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]),
                                                            pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
      }
    }
  }
  __device__ static void tile_end(const Params&, State&, const Shape&, int, int, int, int) {}
};

// --------------------------------------------------------------------------------------------
// Epilogue policy 3: decode-time lm_head for search: per (row, column-half) partial log-softmax
// statistics + the TOPK best (value, index) of the half tile.  Used by greedy (TOPK=1 semantics via
// k=2*beams>=2) and beam search; a follow-up kernel merges the partial lists (beam.cu).
// --------------------------------------------------------------------------------------------
constexpr int SEARCH_TOPK = 8;
struct EpiSearchParams {
  const float* bias;   // [N] or null
  int mask_token;      // FlaxMinLengthLogitsProcessor: this token id scores -inf (-1 = none)
  float* pmax;         // [2*num_n_blocks, M]
  float* psum;         // [2*num_n_blocks, M]
  float* cand_val;     // [2*num_n_blocks, M, SEARCH_TOPK] raw logits (descending)
  int* cand_idx;       // [2*num_n_blocks, M, SEARCH_TOPK] vocab ids
};

struct EpiSearch {
  typedef EpiSearchParams Params;
  struct State {
    float mx, sm;
    float tv[SEARCH_TOPK];
    int ti[SEARCH_TOPK];
  };
  __device__ static void tile_begin(const Params&, State& st, const Shape&, int, int, int) {
    st.mx = -INFINITY;
    st.sm = 0.f;
#pragma unroll
    for (int i = 0; i < SEARCH_TOPK; ++i) {
      st.tv[i] = -INFINITY;
      st.ti[i] = 0x7fffffff;
    }
  }
  __device__ static void chunk(const Params& p, State& st, const Shape& s, int row, int col0, float* v) {
    if (row >= s.M || col0 >= s.N) return;
    float cmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int c = col0 + j;
      float z = v[j];
      if (c < s.N && c != p.mask_token) {
        if (p.bias) z += p.bias[c];
      } else {
        z = -INFINITY;
      }
      v[j] = z;
      cmax = fmaxf(cmax, z);
      // sorted insert (descending; on ties the earlier = lower index stays ahead: strict >)
      if (z > st.tv[SEARCH_TOPK - 1]) {
        float cv = z;
        int ci = c;
#pragma unroll
        for (int i = 0; i < SEARCH_TOPK; ++i) {
          if (cv > st.tv[i]) {
            const float tvv = st.tv[i];
            const int tii = st.ti[i];
            st.tv[i] = cv;
            st.ti[i] = ci;
            cv = tvv;
            ci = tii;
          }
        }
      }
    }
    const float nm = fmaxf(st.mx, cmax);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) acc += __expf(v[j] - nm);
    st.sm = st.sm * __expf(st.mx - nm) + acc;
    st.mx = nm;
  }
  __device__ static void tile_end(const Params& p, State& st, const Shape& s, int row, int, int n_blk, int half) {
    if (row >= s.M) return;
    const long long o = (long long)(n_blk * 2 + half) * s.M + row;
    p.pmax[o] = st.mx;
    p.psum[o] = st.sm;
#pragma unroll
    for (int i = 0; i < SEARCH_TOPK; ++i) {
      p.cand_val[o * SEARCH_TOPK + i] = st.tv[i];
      p.cand_idx[o * SEARCH_TOPK + i] = st.ti[i];
    }
  }
};

// --------------------------------------------------------------------------------------------
// The kernel
// --------------------------------------------------------------------------------------------
template <int A_MN, int B_MN, int BN, class Epi>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const Shape shape, const typename Epi::Params ep) {
  typedef Cfg<BN> C;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);   // 1024B alignment for SWIZZLE_128B
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::STAGES * C::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::STAGES;
  uint64_t* tmem_full = bars + 2 * C::STAGES;
  uint64_t* tmem_empty = bars + 2 * C::STAGES + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = shape.num_m_blocks * shape.num_n_blocks;
  const int num_k_blocks = (shape.K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const TileCoord tc = tile_coord(shape, t);
        const int m0 = tc.m_blk * BLOCK_M, n0 = tc.n_blk * BN;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          uint8_t* sa = smem_a + stage * C::A_BYTES;
          uint8_t* sb = smem_b + stage * C::B_BYTES;
          const int k0 = kb * BLOCK_K;
          if (A_MN == 0) {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int i = 0; i < BLOCK_M / 64; ++i)
              tma_load_2d(sa + i * (BLOCK_K * 128), &tmap_a, &full_bar[stage], m0 + i * 64, k0);
          }
          if (B_MN == 0) {
            tma_load_2d(sb, &tmap_b, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tma_load_2d(sb + i * (BLOCK_K * 128), &tmap_b, &full_bar[stage], n0 + i * 64, k0);
          }
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BN, A_MN, B_MN);
      // K-major:  8-row groups 1024B apart (SBO), one swizzle atom along K (LBO unused);
      //           advancing K by 16 elements = +32 bytes inside the 128B row.
      // MN-major: 64-element MN atoms BLOCK_K*128B apart (LBO), 8-k-row groups 1024B apart (SBO);
      //           advancing K by 16 rows = +2048 bytes.
      constexpr uint32_t a_lbo = A_MN ? BLOCK_K * 128 : 0, a_sbo = 1024, a_kstep = A_MN ? UMMA_K * 128 : UMMA_K * 2;
      constexpr uint32_t b_lbo = B_MN ? BLOCK_K * 128 : 0, b_sbo = 1024, b_kstep = B_MN ? UMMA_K * 128 : UMMA_K * 2;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const uint32_t as = it & 1, aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * C::A_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * C::B_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = umma_smem_desc(a_addr + k * a_kstep, a_lbo, a_sbo);
            const uint64_t db = umma_smem_desc(b_addr + k * b_kstep, b_lbo, b_sbo);
            umma_bf16(tmem_d, da, db, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);   // frees the smem slot once these MMAs retire
          if (kb == num_k_blocks - 1) umma_commit(&tmem_full[as]);
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp - 4;
    const int quarter = warp & 3;       // TMEM lane quarter this warp may read
    const int half = ew >> 2;           // which half of the tile's columns
    uint32_t it = 0;
    typename Epi::State st;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const TileCoord tc = tile_coord(shape, t);
      const uint32_t as = it & 1, aphase = (it >> 1) & 1;
      const int row = tc.m_blk * BLOCK_M + quarter * 32 + lane;
      Epi::tile_begin(ep, st, shape, row, tc.m_blk, tc.n_blk);
      mbar_wait(&tmem_full[as], aphase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN + half * C::COLS_PER_WARP;
#pragma unroll 1
      for (int c = 0; c < C::COLS_PER_WARP / 32; ++c) {
        float v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        tmem_ld_wait();
        Epi::chunk(ep, st, shape, row, tc.n_blk * BN + half * C::COLS_PER_WARP + c * 32, v);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      Epi::tile_end(ep, st, shape, row, tc.m_blk, tc.n_blk, half);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace micgemm
