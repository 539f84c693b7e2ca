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
// w4..w11 epilogue.  An epilogue warp owns one TMEM lane quarter (32 rows) and a share of the tile's
// 64-column groups.  Outputs leave through a warp-private 4 KB swizzled staging buffer and TMA bulk
// stores (coalesced, asynchronous, tails clipped by the tensor map).  Register-heavy policies (store,
// search) take a group as two 32-column halves (HALF_GROUPS: 32 live accumulators instead of 64 - the
// 64-column form spilled to local memory in the hot path); residual rows are L2-prefetched one tile
// ahead and loaded into registers before the accumulator wait (tile_prefetch / group_pre).
#pragma once

#include "common.cuh"

namespace micgemm {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;           // default; activation-heavy epilogues use 16 (Epi::EW)
constexpr int NUM_THREADS = 128 + NUM_EPI_WARPS * 32;
constexpr int TMEM_COLS = 512;
constexpr int GROUP_COLS = 64;          // epilogue granularity: 64 accumulator columns
constexpr int STG_BYTES = 4096;         // per-warp staging: 32 rows x 128 B

template <int BN, int NBUF = 1, int EW = NUM_EPI_WARPS>
struct Cfg {
  static_assert(BN == 64 || BN == 128 || BN == 192 || BN == 256, "unsupported BLOCK_N");
  static_assert(NBUF == 1 || NBUF == 2, "one or two staging buffers per epilogue warp");
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BN * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // smem budget 227 KB: mainloop stages + NBUF x 32 KB of epilogue staging
  static constexpr int EPI_BYTES = EW * STG_BYTES * NBUF;
  static constexpr int STAGES_WANTED = BN == 256 ? 4 : (BN == 192 ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int STAGES_FIT = (232448 - 1280 - EPI_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_WANTED < STAGES_FIT ? STAGES_WANTED : STAGES_FIT;
  static constexpr int THREADS = 128 + EW * 32;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int NUM_GROUPS = BN / GROUP_COLS;
  static constexpr int GROUPS_HALF0 = (NUM_GROUPS + 1) / 2;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");
};

struct Shape {
  int M, N, K;
  int num_m_blocks, num_n_blocks, group_m;
  int split_k;        // >1: K is cut into `split_k` slices, one work unit each, combined by TMA reduce-add
  int kb_per_split;   // k-blocks per slice
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

// per-call context handed to the epilogue policies
struct EpiCtx {
  int row;       // global row owned by this thread
  int row0;      // first global row of the warp's 32-row slab
  int lane;
  uint8_t* stg;  // warp-private staging buffer (1024-byte aligned)
  int next_col;  // first column of the group this warp handles next in the same tile, or -1 (set by the kernel loop)
  const CUtensorMap* tmap_d;
  const CUtensorMap* tmap_d2;
};

// swizzled (SWIZZLE_128B) address of 16-byte unit `u` of row `r` in a [32 x 128 B] staging tile
__device__ __forceinline__ uint8_t* stg_addr(uint8_t* stg, int r, int u) { return stg + r * 128 + ((u ^ (r & 7)) << 4); }

// make the staging buffer writable again: the issuing lane waits until earlier TMA stores have read it
template <int PENDING = 0>
__device__ __forceinline__ void stg_acquire(int lane) {
  if (lane == 0) {
    if (PENDING == 0)
      tma_store_wait_read();
    else
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the newest store may still be reading
  }
  __syncwarp();
}
// publish the staging buffer (written with st.shared by all lanes) through a TMA store
__device__ __forceinline__ void stg_store(const CUtensorMap* map, uint8_t* stg, int lane, int c0, int r0, bool reduce_add) {
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) {
    if (reduce_add)
      tma_reduce_add_2d(map, stg, c0, r0);
    else
      tma_store_2d(map, stg, c0, r0);
    tma_store_commit();
  }
}

// --------------------------------------------------------------------------------------------
// Epilogue policy 0: store with optional bias / activation / residual / accumulate
// --------------------------------------------------------------------------------------------
struct EpiStoreParams {
  void* D;              // bf16 or fp32 [M, ldd]
  long long ldd;
  int d_f32;            // 1: fp32 output
  int accumulate;       // fp32 only: D += result
  int tma_ok;           // outputs reachable by TMA (16-byte aligned pitch / base) and N % 8 == 0
  const float* bias;    // [N] fp32 or null
  int act;              // MIC_ACT_*
  bf16* D2;             // optional pre-activation copy (bf16, same ld as D) or null
  const bf16* residual; // optional bf16 [M, ldr] added after activation
  long long ldr;
  float out_scale;      // applied to acc before the bias (1.0 normally)
  DropoutParams drop;   // dropout on the activation output, before the residual add (flax: x + dropout(f(x)))
};

// NBUF = 2: a second staging buffer per warp carries the pre-activation copy (D2), so the two outputs of a
// group alternate buffers and never wait on each other's TMA read (used for the fc1 GEMMs of training).
// ACT_BWD = true (EpiStoreActBwd16): the `residual` tile is the saved pre-activation U of the PREVIOUS layer and
// the result is multiplied by act'(U) instead of added to it - the fc2 dgrad GEMM then emits d(fc1 pre-activation)
// directly and the separate activation-backward pass over [M, ffn] (read dG, read U, write dU) disappears.
template <int NBUF_, int EW_ = NUM_EPI_WARPS, bool ACT_BWD = false>
struct EpiStoreT {
  static constexpr int NBUF = NBUF_;
  static constexpr int EW = EW_;
  typedef EpiStoreParams Params;
  // The single-buffer store policies take every 64-column group as two 32-column halves: with 64 live accumulators the
  // 16-warp form (96 registers per thread) spilled 592 B per thread in the GELU epilogue of the fc1 forward GEMMs (the GEMM
  // class furthest from the roofline in profiles/r02_gemm_table.txt) and the 8-warp form 76 B next to its residual registers.
  // The 16-warp form carries no residual path (the launcher routes residual epilogues to the 8-warp policy).
  static constexpr bool HALF_GROUPS = (!ACT_BWD && NBUF_ == 1);
  struct State {
    uint4 res[8];        // this thread's 64 residual values of the group about to be processed (loaded ahead by group_pre)
    uint32_t keep[16];   // HALF_GROUPS: the packed bf16 outputs of the first half, until the second half completes the row
  };
  __device__ static void kernel_begin(const Params&, State&) {}
  __device__ static void kernel_end(const Params&, State&, const Shape&, int, int) {}
  __device__ static void tile_begin(const Params&, State&, const Shape&, int, int, int) {}
  __device__ static void tile_end(const Params&, State&, const Shape&, int, int, int, int) {}
  // Called one tile ahead by every epilogue warp: pull the residual rows this warp will add in its NEXT tile towards
  // L2.  The residual read otherwise sits on the epilogue's critical path (HBM latency per 64-column group: +17 us
  // on a 35 us [16384,1024,1024] GEMM, profiles/r02_gemm_epilogue_costs.txt) while the tensor pipe waits for TMEM.
  __device__ static void tile_prefetch(const Params& p, const Shape& s, int row, int col_begin, int col_end) {
    if (!p.residual || ACT_BWD || row >= s.M) return;
    const bf16* r = p.residual + (long long)row * p.ldr;
    for (int c = col_begin; c < col_end && c < s.N; c += GROUP_COLS)   // 64 bf16 = one 128-byte line per group
      asm volatile("prefetch.global.L2 [%0];" ::"l"(r + c));
  }

  // The residual (or saved pre-activation) values of one 64-column group, straight into registers: each lane reads
  // the 128 contiguous bytes of its own row.  Issued BEFORE the wait on the accumulator (first group of a tile) or at
  // the start of the previous group's math, so the load latency is off the TMEM-drain critical path; the lines were
  // pulled into L2 one tile earlier by tile_prefetch.  Safe for in-place use (residual == D): a thread reads exactly
  // the elements it later overwrites.
  __device__ static void group_pre(const Params& p, State& st, const Shape& s, const EpiCtx& ctx, int col0) {
    if (EW_ == 16 || !p.residual || !p.tma_ok) return;
    const bool row_ok = ctx.row < s.M;
    const bf16* r = p.residual + (long long)ctx.row * p.ldr + col0;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      st.res[u] = make_uint4(0, 0, 0, 0);
      if (row_ok && col0 + u * 8 < s.N) st.res[u] = *reinterpret_cast<const uint4*>(r + u * 8);
    }
  }

  // scalar fallback (unaligned outputs / odd N): v = 32 consecutive columns [col0, col0+32) of row `row`
  __device__ static void chunk_scalar(const Params& p, const Shape& s, int row, int col0, const float* v) {
    if (row >= s.M || col0 >= s.N) return;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int c = col0 + j;
      if (c >= s.N) continue;
      float x = v[j] * p.out_scale;
      if (p.bias) x += p.bias[c];
      if (p.D2) p.D2[(long long)row * p.ldd + c] = __float2bfloat16_rn(x);
      x = act_fwd(x, p.act);
      if (p.drop.seed_ptr) {
        const uint32_t e = (uint32_t)row * (uint32_t)s.N + (uint32_t)c;
        bool k0, k1;
        drop_keep2(e >> 1, *p.drop.seed_ptr + p.drop.site, p.drop.thr16, &k0, &k1);
        x = ((e & 1) ? k1 : k0) ? x * p.drop.scale : 0.f;
      }
      if (p.residual) x += __bfloat162float(p.residual[(long long)row * p.ldr + c]);
      if (p.d_f32) {
        float* d = reinterpret_cast<float*>(p.D) + (long long)row * p.ldd + c;
        *d = p.accumulate ? (*d + x) : x;
      } else {
        reinterpret_cast<bf16*>(p.D)[(long long)row * p.ldd + c] = __float2bfloat16_rn(x);
      }
    }
  }

  // HALF_GROUPS form: v = 32 consecutive columns [col0, col0+32); the two halves of a 64-column group arrive back to
  // back (first the even, then the odd multiple of 32) and share the 128-byte-per-row staging tile: the pre-activation
  // copy is staged half by half and stored after the second, the activated halves meet in the staging tile through
  // st.keep.  The residual registers of the whole group were loaded ahead by group_pre; half hf consumes res[4 hf .. 4 hf + 3].
  template <int NC>
  __device__ static void group_n(const Params& p, State& st, const Shape& s, const EpiCtx& ctx, int col0, float* v) {
    static_assert(NC == 32, "half groups are 32 columns");
    const int hf = (col0 >> 5) & 1;
    const int gcol0 = col0 - hf * 32;
    if (gcol0 >= s.N) return;                     // warp-uniform: the whole 64-column group is outside the matrix
    if (!p.tma_ok) {
      chunk_scalar(p, s, ctx.row, col0, v);
      return;
    }
    const int lane = ctx.lane;
    uint8_t* stg = ctx.stg;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = col0 + u * 8;
      float4 b0 = make_float4(0, 0, 0, 0), b1 = b0;
      if (p.bias && c < s.N) {
        b0 = *reinterpret_cast<const float4*>(p.bias + c);
        b1 = *reinterpret_cast<const float4*>(p.bias + c + 4);
      }
      float* x = v + u * 8;
      x[0] = x[0] * p.out_scale + b0.x; x[1] = x[1] * p.out_scale + b0.y;
      x[2] = x[2] * p.out_scale + b0.z; x[3] = x[3] * p.out_scale + b0.w;
      x[4] = x[4] * p.out_scale + b1.x; x[5] = x[5] * p.out_scale + b1.y;
      x[6] = x[6] * p.out_scale + b1.z; x[7] = x[7] * p.out_scale + b1.w;
    }
    if (p.D2) {
      if (hf == 0) stg_acquire<0>(lane);          // the previous store out of this tile has been read
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* x = v + u * 8;
        *reinterpret_cast<uint4*>(stg_addr(stg, lane, hf * 4 + u)) =
            make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
      }
      if (hf == 1) stg_store(ctx.tmap_d2, stg, lane, gcol0, ctx.row0, false);
    }
    if (p.act != MIC_ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = act_fwd(v[j], p.act);
    }
    if (p.drop.seed_ptr) {
      const uint32_t seed = *p.drop.seed_ptr + p.drop.site;
      const uint32_t base = ((uint32_t)ctx.row * (uint32_t)s.N + (uint32_t)col0) >> 1;   // N, col0 even on this path
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        bool k0, k1;
        drop_keep2(base + j, seed, p.drop.thr16, &k0, &k1);
        v[2 * j] = k0 ? v[2 * j] * p.drop.scale : 0.f;
        v[2 * j + 1] = k1 ? v[2 * j + 1] * p.drop.scale : 0.f;
      }
    }
    if (EW_ != 16 && p.residual) {      // (the 96-register 16-warp form has no room for the residual registers)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float* x = v + u * 8;
        const uint4 r = hf ? st.res[4 + u] : st.res[u];
        float2 f;
        f = unpack_bf16(r.x); x[0] += f.x; x[1] += f.y;
        f = unpack_bf16(r.y); x[2] += f.x; x[3] += f.y;
        f = unpack_bf16(r.z); x[4] += f.x; x[5] += f.y;
        f = unpack_bf16(r.w); x[6] += f.x; x[7] += f.y;
      }
      // the next group's residual loads go out as soon as this group's registers are free (overlaps the staging below)
      if (hf == 1 && ctx.next_col >= 0) group_pre(p, st, s, ctx, ctx.next_col);
    }
    if (p.d_f32) {
      // fp32 output: 32 columns are exactly one 128-byte staging row
      if (col0 >= s.N) return;                     // warp-uniform: this half lies wholly outside the matrix
      stg_acquire<0>(lane);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float* x = v + u * 4;
        *reinterpret_cast<float4*>(stg_addr(stg, lane, u)) = make_float4(x[0], x[1], x[2], x[3]);
      }
      stg_store(ctx.tmap_d, stg, lane, col0, ctx.row0, p.accumulate != 0);
      return;
    }
    if (hf == 0) {
#pragma unroll
      for (int j = 0; j < 16; ++j) st.keep[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
      return;
    }
    stg_acquire<0>(lane);                          // (the pre-activation store, if any, has been read)
#pragma unroll
    for (int u = 0; u < 4; ++u)
      *reinterpret_cast<uint4*>(stg_addr(stg, lane, u)) =
          make_uint4(st.keep[4 * u], st.keep[4 * u + 1], st.keep[4 * u + 2], st.keep[4 * u + 3]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float* x = v + u * 8;
      *reinterpret_cast<uint4*>(stg_addr(stg, lane, 4 + u)) =
          make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
    }
    stg_store(ctx.tmap_d, stg, lane, gcol0, ctx.row0, false);
  }

  // v: 64 consecutive columns [col0, col0+64) of row ctx.row (fp32 accumulators)
  __device__ static void group(const Params& p, State& st, const Shape& s, const EpiCtx& ctx, int col0, float* v) {
    if (col0 >= s.N) return;                      // warp-uniform
    if (!p.tma_ok) {
      chunk_scalar(p, s, ctx.row, col0, v);
      chunk_scalar(p, s, ctx.row, col0 + 32, v + 32);
      return;
    }
    const int lane = ctx.lane;
    uint8_t* stg = ctx.stg;
    // ---- residual values: loaded ahead by group_pre (the next group's loads are issued further down, once these
    // registers have been consumed: live registers stay at accumulators + one residual set) ----
    uint4 (&res)[8] = st.res;
    // ---- bias (vector loads; N % 8 == 0 on this path so a unit is all-in or all-out) ----
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c = col0 + u * 8;
      float4 b0 = make_float4(0, 0, 0, 0), b1 = b0;
      if (p.bias && c < s.N) {
        b0 = *reinterpret_cast<const float4*>(p.bias + c);
        b1 = *reinterpret_cast<const float4*>(p.bias + c + 4);
      }
      float* x = v + u * 8;
      x[0] = x[0] * p.out_scale + b0.x; x[1] = x[1] * p.out_scale + b0.y;
      x[2] = x[2] * p.out_scale + b0.z; x[3] = x[3] * p.out_scale + b0.w;
      x[4] = x[4] * p.out_scale + b1.x; x[5] = x[5] * p.out_scale + b1.y;
      x[6] = x[6] * p.out_scale + b1.z; x[7] = x[7] * p.out_scale + b1.w;
    }
    // ---- optional pre-activation copy ----
    if (p.D2 && (NBUF == 2 || EW == 16)) {
      // staged TMA store of the copy: into the second buffer (NBUF == 2), or — with 16 epilogue warps, one
      // group per warp and tile — into the single buffer, whose read finishes during the activation math
      uint8_t* stg2 = NBUF == 2 ? stg + STG_BYTES : stg;
      if (NBUF == 2)
        stg_acquire<1>(lane);
      else
        stg_acquire<0>(lane);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float* x = v + u * 8;
        *reinterpret_cast<uint4*>(stg_addr(stg2, lane, u)) =
            make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
      }
      stg_store(ctx.tmap_d2, stg2, lane, col0, ctx.row0, false);
    } else if (p.D2) {
      // single staging buffer: direct row-per-thread stores for the copy
      if (ctx.row < s.M) {
        bf16* d2 = p.D2 + (long long)ctx.row * p.ldd + col0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float* x = v + u * 8;
          if (col0 + u * 8 < s.N)
            *reinterpret_cast<uint4*>(d2 + u * 8) =
                make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
        }
      }
    }
    if constexpr (ACT_BWD) {
      if (p.residual) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float* x = v + u * 8;
          float2 f;
          f = unpack_bf16(res[u].x); x[0] *= act_bwd(f.x, p.act); x[1] *= act_bwd(f.y, p.act);
          f = unpack_bf16(res[u].y); x[2] *= act_bwd(f.x, p.act); x[3] *= act_bwd(f.y, p.act);
          f = unpack_bf16(res[u].z); x[4] *= act_bwd(f.x, p.act); x[5] *= act_bwd(f.y, p.act);
          f = unpack_bf16(res[u].w); x[6] *= act_bwd(f.x, p.act); x[7] *= act_bwd(f.y, p.act);
        }
      }
    } else if (p.act != MIC_ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 64; ++j) v[j] = act_fwd(v[j], p.act);
    }
    if (p.drop.seed_ptr) {
      const uint32_t seed = *p.drop.seed_ptr + p.drop.site;
      const uint32_t base = ((uint32_t)ctx.row * (uint32_t)s.N + (uint32_t)col0) >> 1;   // N, col0 even on this path
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        bool k0, k1;
        drop_keep2(base + j, seed, p.drop.thr16, &k0, &k1);
        v[2 * j] = k0 ? v[2 * j] * p.drop.scale : 0.f;
        v[2 * j + 1] = k1 ? v[2 * j + 1] * p.drop.scale : 0.f;
      }
    }
    if (!ACT_BWD && p.residual) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float* x = v + u * 8;
        float2 f;
        f = unpack_bf16(res[u].x); x[0] += f.x; x[1] += f.y;
        f = unpack_bf16(res[u].y); x[2] += f.x; x[3] += f.y;
        f = unpack_bf16(res[u].z); x[4] += f.x; x[5] += f.y;
        f = unpack_bf16(res[u].w); x[6] += f.x; x[7] += f.y;
      }
    }
    if (ctx.next_col >= 0) group_pre(p, st, s, ctx, ctx.next_col);   // overlaps the packing / staging / TMA store below
    if (!p.d_f32) {
      if (NBUF == 2 && p.D2)
        stg_acquire<1>(lane);
      else
        stg_acquire<0>(lane);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float* x = v + u * 8;
        *reinterpret_cast<uint4*>(stg_addr(stg, lane, u)) =
            make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
      }
      stg_store(ctx.tmap_d, stg, lane, col0, ctx.row0, false);
    } else {
      // fp32: two 32-column (128-byte) halves through the same buffer
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        if (col0 + hh * 32 >= s.N) break;        // warp-uniform
        stg_acquire<0>(lane);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float* x = v + hh * 32 + u * 4;
          *reinterpret_cast<float4*>(stg_addr(stg, lane, u)) = make_float4(x[0], x[1], x[2], x[3]);
        }
        stg_store(ctx.tmap_d, stg, lane, col0 + hh * 32, ctx.row0, p.accumulate != 0);
      }
    }
  }
};
typedef EpiStoreT<1> EpiStore;
typedef EpiStoreT<2> EpiStoreDual;
typedef EpiStoreT<1, 16> EpiStoreAct16;      // 16 epilogue warps: GELU / quick-GELU epilogues are issue bound
typedef EpiStoreT<1, 8, true> EpiStoreActBwd16;    // dgrad GEMM fused with the activation backward of its consumer (8 warps: the 16-warp form is register-starved)

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
  int store_logits;      // 1: also emit the (bias-added) logits as bf16 through tmap_d (backward reuses them)
};

#define MIC_LOG2E 1.4426950408889634f

struct EpiCEStats {
  typedef EpiCEStatsParams Params;
  static constexpr int NBUF = 1;
  static constexpr int EW = NUM_EPI_WARPS;
  struct State {
    float mx, sm, sz;
    int label;
  };
  __device__ static void kernel_begin(const Params&, State&) {}
  __device__ static void kernel_end(const Params&, State&, const Shape&, int, int) {}
  __device__ static void tile_begin(const Params& p, State& st, const Shape& s, int row, int, int) {
    st.mx = -INFINITY;
    st.sm = 0.f;
    st.sz = 0.f;
    st.label = (row < s.M) ? p.labels[row] : -1;
  }
  __device__ static void group(const Params& p, State& st, const Shape& s, const EpiCtx& ctx, int col0, float* v) {
    if (col0 >= s.N) return;                      // warp-uniform
    const bool full = col0 + 64 <= s.N;
    const bool bias_vec = p.bias && ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
    if (full && (bias_vec || !p.bias)) {
      if (p.bias) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const float4 b = *reinterpret_cast<const float4*>(p.bias + col0 + u * 4);
          v[u * 4 + 0] += b.x; v[u * 4 + 1] += b.y; v[u * 4 + 2] += b.z; v[u * 4 + 3] += b.w;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        const int c = col0 + j;
        v[j] = (c < s.N) ? (p.bias ? v[j] + p.bias[c] : v[j]) : -INFINITY;
      }
    }
    const unsigned rel = static_cast<unsigned>(st.label - col0);
    if (rel < 64u && ctx.row < s.M) {              // the label's logit lives in this group (rare)
      float zy = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) zy = (rel == static_cast<unsigned>(j)) ? v[j] : zy;
      p.zlabel[ctx.row] = zy;
    }
    float cmax = v[0], csum = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
      cmax = fmaxf(cmax, v[j]);
      csum += full ? v[j] : ((col0 + j < s.N) ? v[j] : 0.f);
    }
    st.sz += csum;
    const float nm = fmaxf(st.mx, cmax);
    const float nm2 = nm * MIC_LOG2E;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) acc += ex2_approx(fmaf(v[j], MIC_LOG2E, -nm2));
    st.sm = st.sm * ex2_approx((st.mx - nm) * MIC_LOG2E) + acc;
    st.mx = nm;
    if (p.store_logits) {
      stg_acquire<0>(ctx.lane);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float* x = v + u * 8;
        *reinterpret_cast<uint4*>(stg_addr(ctx.stg, ctx.lane, u)) =
            make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
      }
      stg_store(ctx.tmap_d, ctx.stg, ctx.lane, col0, ctx.row0, false);
    }
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
  long long ldd;         // multiple of 256; columns [N, ldd) are written as zero
};

struct EpiCEGrad {
  typedef EpiCEGradParams Params;
  static constexpr int NBUF = 1;
  static constexpr int EW = NUM_EPI_WARPS;
  struct State {
    float lse2, w;
    int label;
  };
  __device__ static void kernel_begin(const Params&, State&) {}
  __device__ static void kernel_end(const Params&, State&, const Shape&, int, int) {}
  __device__ static void tile_begin(const Params& p, State& st, const Shape& s, int row, int, int) {
    const bool ok = row < s.M;
    st.lse2 = ok ? p.lse[row] * MIC_LOG2E : 0.f;
    st.w = ok ? p.row_w[row] : 0.f;
    st.label = ok ? p.labels[row] : -1;
  }
  __device__ static void group(const Params& p, State& st, const Shape& s, const EpiCtx& ctx, int col0, float* v) {
    if (col0 >= p.ldd) return;                    // warp-uniform
    const bool full = col0 + 64 <= s.N;
    const bool bias_vec = p.bias && ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
    if (full && (bias_vec || !p.bias)) {
      if (p.bias) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const float4 b = *reinterpret_cast<const float4*>(p.bias + col0 + u * 4);
          v[u * 4 + 0] += b.x; v[u * 4 + 1] += b.y; v[u * 4 + 2] += b.z; v[u * 4 + 3] += b.w;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        const int c = col0 + j;
        v[j] = (c < s.N) ? (p.bias ? v[j] + p.bias[c] : v[j]) : -INFINITY;   // exp2(-inf) = 0
      }
    }
    const float lw = p.low * st.w;
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = fmaf(ex2_approx(fmaf(v[j], MIC_LOG2E, -st.lse2)), st.w, -lw);
    if (!full) {
#pragma unroll
      for (int j = 0; j < 64; ++j) v[j] = (col0 + j < s.N) ? v[j] : 0.f;     // padded vocab columns: exactly zero
    }
    const unsigned rel = static_cast<unsigned>(st.label - col0);
    if (rel < 64u) {
      const float fix = (p.low - p.conf) * st.w;
#pragma unroll
      for (int j = 0; j < 64; ++j) v[j] += (rel == static_cast<unsigned>(j)) ? fix : 0.f;
    }
    stg_acquire<0>(ctx.lane);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float* x = v + u * 8;
      *reinterpret_cast<uint4*>(stg_addr(ctx.stg, ctx.lane, u)) =
          make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
    }
    stg_store(ctx.tmap_d, ctx.stg, ctx.lane, col0, ctx.row0, false);
  }
  __device__ static void tile_end(const Params&, State&, const Shape&, int, int, int, int) {}
};

// --------------------------------------------------------------------------------------------
// Epilogue policy 3: decode-time lm_head for search: per (row, column-half) partial log-softmax
// statistics + the TOPK best (value, index) of the half tile; a follow-up kernel merges the partial
// lists (beam.cu).
// --------------------------------------------------------------------------------------------
constexpr int SEARCH_TOPK = 8;
struct EpiSearchParams {
  const float* bias;   // [N] or null
  int mask_token;      // FlaxMinLengthLogitsProcessor: this token id scores -inf (-1 = none)
  float* pmax;         // [num_partials, M]
  float* psum;         // [num_partials, M]
  float* cand_val;     // [num_partials, M, SEARCH_TOPK] raw logits (descending)
  int* cand_idx;       // [num_partials, M, SEARCH_TOPK] vocab ids
  // packed-operand variant (EpiSearchPacked): both operands pre-arranged as contiguous SWIZZLE_128B tile images
  // [row tile][k block], fetched with one bulk copy per operand per stage instead of 384 TMA box rows
  const bf16* a_tiles;
  const bf16* b_tiles;
  // second pass for 5..8 beams (2K = 10..16 candidates per row): only (logit, token) pairs that rank strictly after
  // the row's pair (upper_val, upper_idx) - the 8th best of the first pass - are considered; null = first pass
  const float* upper_val;
  const int* upper_idx;
  // device flag of the search loop's while_loop condition (null = always run): 0 -> the kernel returns at once, the
  // (stale) partial lists are ignored by the equally gated bookkeeping kernels
  const int* active;
  // `_sample` (generation_clip_vision_utils.py:537-663): next = jax.random.categorical(key, raw logits) =
  // argmax(logits + Gumbel noise).  gumbel_on: every value gets -log(-log(u)) with u from the threefry2x32 stream of
  // (gumbel_k0, gumbel_k1) at the element's flat index in the [M, N] logits array, so the row's top-1 IS the sample.
  int gumbel_on;
  unsigned int gumbel_k0, gumbel_k1;
};

// jax._src.random (0.2.16, pinned by requirements.txt:13) restated [MEMORY, see oracle/reference_generate.py]:
// threefry2x32 with 20 rounds; random_bits(key, 32, shape) evaluates it on counts iota(size) split into halves
// (x0 = counts[:half], x1 = counts[half:], odd sizes padded with one 0), output = concat(y0, y1);
// uniform(minval = tiny, maxval = 1): f = bitcast((bits >> 9) | 0x3f800000) - 1, u = max(tiny, f * (1 - tiny) + tiny);
// gumbel = -log(-log(u)).
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
__device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  x0 += ks[0];
  x1 += ks[1];
#pragma unroll
  for (int g = 0; g < 5; ++g) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      x0 += x1;
      x1 = rotl32(x1, R[g & 1][r]);
      x1 ^= x0;
    }
    x0 += ks[(g + 1) % 3];
    x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
  }
}
__device__ __forceinline__ float gumbel_at(uint32_t k0, uint32_t k1, unsigned long long i, unsigned long long total) {
  const unsigned long long half = (total + 1) >> 1;
  uint32_t x0, x1;
  if (i < half) {
    x0 = (uint32_t)i;
    x1 = (i + half < total) ? (uint32_t)(i + half) : 0u;      // the pad element of an odd-sized count array
  } else {
    x0 = (uint32_t)(i - half);
    x1 = (uint32_t)i;
  }
  threefry2x32(k0, k1, x0, x1);
  const uint32_t bits = (i < half) ? x0 : x1;
  const float f = __uint_as_float((bits >> 9) | 0x3f800000u) - 1.0f;
  const float tiny = 1.17549435e-38f;
  const float u = fmaxf(tiny, f * (1.0f - tiny) + tiny);
  return -logf(-logf(u));
}

// The launcher sizes the grid as a multiple of num_m_blocks with group_m == num_m_blocks, so every CTA
// keeps the SAME m-block for all of its tiles: the running (max, sum) and the running top-8 of a row live
// in registers across the whole kernel and are written once (partial slot = CTA rank within the m-block).
template <bool kGumbel>
struct EpiSearchT {
  typedef EpiSearchParams Params;
  static constexpr int NBUF = 1;
  static constexpr int EW = NUM_EPI_WARPS;
  struct State {
    float mx, sm;
    float tv[SEARCH_TOPK];
    int ti[SEARCH_TOPK];
    float uv;
    int ui;
  };
  __device__ static void kernel_begin(const Params&, State& st) {
    st.mx = -INFINITY;
    st.sm = 0.f;
#pragma unroll
    for (int i = 0; i < SEARCH_TOPK; ++i) {
      st.tv[i] = -INFINITY;
      st.ti[i] = 0x7fffffff;
    }
  }
  __device__ static void tile_begin(const Params& p, State& st, const Shape& s, int row, int, int) {
    if (p.upper_val) {
      st.uv = row < s.M ? p.upper_val[row] : -INFINITY;
      st.ui = row < s.M ? p.upper_idx[row] : 0;
    }
  }
  __device__ static void tile_end(const Params&, State&, const Shape&, int, int, int, int) {}
  // The kernel loop hands this policy HALF groups (32 accumulator columns at a time, HALF_GROUPS below): with 64 live
  // accumulators next to the top-8 state the epilogue spilled ~70 values per group to local memory in its hot path
  // (536 B of spill stores, LDL between the MUFU.EX2 of the exp-sum; the search GEMM ran 200 us against 137 us for the
  // same operand stream with a plain store epilogue).
  static constexpr bool HALF_GROUPS = true;
  __device__ static void group(const Params& p, State& st, const Shape& s, const EpiCtx& ctx, int col0, float* v) {
    group_n<64>(p, st, s, ctx, col0, v);
  }
  template <int NC>
  __device__ static void group_n(const Params& p, State& st, const Shape& s, const EpiCtx& ctx, int col0, float* v) {
    if (col0 >= s.N) return;
    const bool full = col0 + NC <= s.N;
    const bool bias_vec = p.bias && ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
    if (full && (bias_vec || !p.bias)) {
      if (p.bias) {
#pragma unroll
        for (int u = 0; u < NC / 4; ++u) {
          const float4 b = *reinterpret_cast<const float4*>(p.bias + col0 + u * 4);
          v[u * 4 + 0] += b.x; v[u * 4 + 1] += b.y; v[u * 4 + 2] += b.z; v[u * 4 + 3] += b.w;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const int c = col0 + j;
        v[j] = (c < s.N) ? (p.bias ? v[j] + p.bias[c] : v[j]) : -INFINITY;
      }
    }
    const unsigned mrel = static_cast<unsigned>(p.mask_token - col0);
    if (mrel < static_cast<unsigned>(NC)) {
#pragma unroll
      for (int j = 0; j < NC; ++j) v[j] = (mrel == static_cast<unsigned>(j)) ? -INFINITY : v[j];
    }
    if (p.upper_val) {
#pragma unroll
      for (int j = 0; j < NC; ++j)
        if (!(v[j] < st.uv || (v[j] == st.uv && col0 + j > st.ui))) v[j] = -INFINITY;
    }
    if constexpr (kGumbel) {            // `_sample`: a separate instantiation, the search kernels carry none of this
      const unsigned long long total = (unsigned long long)s.M * (unsigned long long)s.N;
      const unsigned long long base = (unsigned long long)ctx.row * (unsigned long long)s.N + (unsigned long long)col0;
      if (ctx.row < s.M) {
#pragma unroll 4
        for (int j = 0; j < NC; ++j)
          if (col0 + j < s.N) v[j] += gumbel_at(p.gumbel_k0, p.gumbel_k1, base + j, total);
      }
    }
    float cmax = v[0];
#pragma unroll
    for (int j = 1; j < NC; ++j) cmax = fmaxf(cmax, v[j]);
    if (__any_sync(0xffffffffu, cmax > st.tv[SEARCH_TOPK - 1])) {
      // Each lane owns a row, so a per-value `if (v[j] > threshold) insert` makes the WARP run the ~40-instruction
      // sorted insert whenever ANY of its 32 rows has a hit: ~60 % of the columns although a lane itself inserts
      // ~2 of 64 (measured: 107 of the kernel's 229 us).  Instead: (1) predicated compaction of every lane's
      // candidates (value above the threshold at group start) into a per-lane queue in the warp's staging
      // shared memory, (2) max-over-lanes(count) rounds of one insert per lane.  Order within a lane stays
      // ascending in the column index, so ties resolve exactly as in the sequential scan.
      constexpr int QCAP = 16;
      const float thr = st.tv[SEARCH_TOPK - 1];
      float2* q = reinterpret_cast<float2*>(ctx.stg);          // [QCAP][32 lanes] (value, column bits)
      int cnt = 0;
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        if (v[j] > thr) {
          if (cnt < QCAP) q[cnt * 32 + ctx.lane] = make_float2(v[j], __int_as_float(col0 + j));
          ++cnt;
        }
      }
      if (__any_sync(0xffffffffu, cnt > QCAP)) {
        // queue overflow (only while the running threshold is still low, i.e. the first group or two of a CTA):
        // sequential scan, FULLY unrolled - a partially unrolled scan indexes v[] dynamically, which puts the
        // whole 64-float group into local memory
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          if (v[j] > st.tv[SEARCH_TOPK - 1]) {
            float cv = v[j];
            int ci = col0 + j;
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
      } else {
        int rounds = cnt;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rounds = max(rounds, __shfl_xor_sync(0xffffffffu, rounds, o));
        for (int r = 0; r < rounds; ++r) {
          if (r < cnt) {
            const float2 e = q[r * 32 + ctx.lane];
            float cv = e.x;
            int ci = __float_as_int(e.y);
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
      }
      __syncwarp();                                            // the queue is reused by the warp's next group
    }
    const float nm = fmaxf(st.mx, cmax);
    if (nm == -INFINITY) return;                   // every column masked so far
    const float nm2 = nm * MIC_LOG2E;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < NC; ++j) acc += ex2_approx(fmaf(v[j], MIC_LOG2E, -nm2));
    st.sm = st.sm * ex2_approx((st.mx - nm) * MIC_LOG2E) + acc;
    st.mx = nm;
  }
  __device__ static void kernel_end(const Params& p, State& st, const Shape& s, int row, int slot) {
    if (row >= s.M) return;
    const long long o = (long long)slot * s.M + row;
    p.pmax[o] = st.mx;
    p.psum[o] = st.sm;
#pragma unroll
    for (int i = 0; i < SEARCH_TOPK; ++i) {
      p.cand_val[o * SEARCH_TOPK + i] = st.tv[i];
      p.cand_idx[o * SEARCH_TOPK + i] = st.ti[i];
    }
  }
};

typedef EpiSearchT<false> EpiSearch;
struct EpiSearchPacked : EpiSearchT<false> {};
typedef EpiSearchT<true> EpiSample;                  // + Gumbel noise of jax.random.categorical: row top-1 = the sample
struct EpiSamplePacked : EpiSearchT<true> {};
template <class Epi> struct PackedOperands { static constexpr bool value = false; };
template <class Epi> struct IsSearchEpi { static constexpr bool value = false; };
template <> struct IsSearchEpi<EpiSearch> { static constexpr bool value = true; };
template <> struct IsSearchEpi<EpiSearchPacked> { static constexpr bool value = true; };
template <> struct IsSearchEpi<EpiSample> { static constexpr bool value = true; };
template <> struct IsSearchEpi<EpiSamplePacked> { static constexpr bool value = true; };
template <> struct PackedOperands<EpiSearchPacked> { static constexpr bool value = true; };
template <> struct PackedOperands<EpiSamplePacked> { static constexpr bool value = true; };

// --------------------------------------------------------------------------------------------
// The kernel
// --------------------------------------------------------------------------------------------
template <class E, class = void>
struct EpiHasPrefetch { static constexpr bool value = false; };
template <class E>
struct EpiHasPrefetch<E, decltype((void)&E::tile_prefetch)> { static constexpr bool value = true; };

template <class E, class = void>
struct EpiHalfGroups { static constexpr bool value = false; };
template <class E>
struct EpiHalfGroups<E, decltype((void)E::HALF_GROUPS)> { static constexpr bool value = E::HALF_GROUPS; };

template <int A_MN, int B_MN, int BN, class Epi>
__global__ void __launch_bounds__(128 + Epi::EW * 32, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_d2,
            const Shape shape, const typename Epi::Params ep) {
  typedef Cfg<BN, Epi::NBUF, Epi::EW> C;
  if constexpr (IsSearchEpi<Epi>::value) {
    if (ep.active != nullptr && *ep.active == 0) return;     // uniform over the grid: written by an earlier kernel
  }
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);   // 1024B alignment for SWIZZLE_128B
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::STAGES * C::A_BYTES;
  uint8_t* smem_epi = smem + C::STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + C::EPI_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::STAGES;
  uint64_t* tmem_full = bars + 2 * C::STAGES;
  uint64_t* tmem_empty = bars + 2 * C::STAGES + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = shape.num_m_blocks * shape.num_n_blocks;
  const int num_units = num_tiles * shape.split_k;
  const int num_k_blocks = (shape.K + BLOCK_K - 1) / BLOCK_K;

  if (!PackedOperands<Epi>::value && warp == 0 && lane == 0) {
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
      mbar_init(&tmem_empty[i], Epi::EW);
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
      for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
        const TileCoord tc = tile_coord(shape, u % num_tiles);
        const int m0 = tc.m_blk * BLOCK_M, n0 = tc.n_blk * BN;
        const int kb0 = (u / num_tiles) * shape.kb_per_split;
        const int kb1 = min(kb0 + shape.kb_per_split, num_k_blocks);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          uint8_t* sa = smem_a + stage * C::A_BYTES;
          uint8_t* sb = smem_b + stage * C::B_BYTES;
          const int k0 = kb * BLOCK_K;
          if constexpr (PackedOperands<Epi>::value) {
            bulk_load(sa, ep.a_tiles + ((long long)tc.m_blk * num_k_blocks + kb) * (C::A_BYTES / 2), C::A_BYTES,
                      &full_bar[stage]);
            bulk_load(sb, ep.b_tiles + ((long long)tc.n_blk * num_k_blocks + kb) * (C::B_BYTES / 2), C::B_BYTES,
                      &full_bar[stage]);
          } else if (A_MN == 0) {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int i = 0; i < BLOCK_M / 64; ++i)
              tma_load_2d(sa + i * (BLOCK_K * 128), &tmap_a, &full_bar[stage], m0 + i * 64, k0);
          }
          if constexpr (PackedOperands<Epi>::value) {
          } else if (B_MN == 0) {
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
      for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++it) {
        const uint32_t as = it & 1, aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        const int kb0 = (u / num_tiles) * shape.kb_per_split;
        const int kb1 = min(kb0 + shape.kb_per_split, num_k_blocks);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * C::A_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * C::B_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = umma_smem_desc(a_addr + k * a_kstep, a_lbo, a_sbo);
            const uint64_t db = umma_smem_desc(b_addr + k * b_kstep, b_lbo, b_sbo);
            umma_bf16(tmem_d, da, db, idesc, (kb > kb0) || (k > 0));
          }
          umma_commit(&empty_bar[stage]);   // frees the smem slot once these MMAs retire
          if (kb == kb1 - 1) umma_commit(&tmem_full[as]);
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
    const int half = ew >> 2;           // which share of the tile's 64-column groups (Epi::EW / 4 shares)
    constexpr int kShares = Epi::EW / 4;
    const int g_begin = kShares == 2 ? (half == 0 ? 0 : C::GROUPS_HALF0) : (half * C::NUM_GROUPS) / kShares;
    const int g_end = kShares == 2 ? (half == 0 ? C::GROUPS_HALF0 : C::NUM_GROUPS) : ((half + 1) * C::NUM_GROUPS) / kShares;
    EpiCtx ctx;
    ctx.lane = lane;
    ctx.stg = smem_epi + ew * STG_BYTES * Epi::NBUF;
    ctx.tmap_d = &tmap_d;
    ctx.tmap_d2 = &tmap_d2;
    uint32_t it = 0;
    typename Epi::State st;
    Epi::kernel_begin(ep, st);
    if constexpr (EpiHasPrefetch<Epi>::value) {
      if (blockIdx.x < num_units) {
        const TileCoord t0 = tile_coord(shape, blockIdx.x % num_tiles);
        Epi::tile_prefetch(ep, shape, t0.m_blk * BLOCK_M + quarter * 32 + lane, t0.n_blk * BN + g_begin * GROUP_COLS,
                           t0.n_blk * BN + g_end * GROUP_COLS);
      }
    }
    for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++it) {
      const TileCoord tc = tile_coord(shape, u % num_tiles);
      const uint32_t as = it & 1, aphase = (it >> 1) & 1;
      ctx.row0 = tc.m_blk * BLOCK_M + quarter * 32;
      ctx.row = ctx.row0 + lane;
      if constexpr (EpiHasPrefetch<Epi>::value) {
        if (u + (int)gridDim.x < num_units) {       // one tile ahead: a whole epilogue of lead time
          const TileCoord tn = tile_coord(shape, (u + gridDim.x) % num_tiles);
          Epi::tile_prefetch(ep, shape, tn.m_blk * BLOCK_M + quarter * 32 + lane, tn.n_blk * BN + g_begin * GROUP_COLS,
                             tn.n_blk * BN + g_end * GROUP_COLS);
        }
      }
      Epi::tile_begin(ep, st, shape, ctx.row, tc.m_blk, tc.n_blk);
      if constexpr (EpiHasPrefetch<Epi>::value) {
        if (g_begin < g_end) Epi::group_pre(ep, st, shape, ctx, tc.n_blk * BN + g_begin * GROUP_COLS);
      }
      mbar_wait(&tmem_full[as], aphase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * BN;
#pragma unroll 1
      for (int g = g_begin; g < g_end; ++g) {
        if constexpr (EpiHalfGroups<Epi>::value) {
          // register-heavy epilogues take the group as two 32-column halves (32 live accumulators each)
          ctx.next_col = g + 1 < g_end ? tc.n_blk * BN + (g + 1) * GROUP_COLS : -1;
#pragma unroll 1
          for (int hf = 0; hf < 2; ++hf) {
            float v[GROUP_COLS / 2];
            tmem_ld_32x32(taddr + g * GROUP_COLS + hf * 32, v);
            tmem_ld_wait();
            if (g == g_end - 1 && hf == 1) {         // accumulator drained: hand the TMEM buffer back early
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty[as]);
            }
            Epi::template group_n<GROUP_COLS / 2>(ep, st, shape, ctx, tc.n_blk * BN + g * GROUP_COLS + hf * 32, v);
          }
          continue;
        }
        float v[GROUP_COLS];
        tmem_ld_32x32(taddr + g * GROUP_COLS, v);
        tmem_ld_32x32(taddr + g * GROUP_COLS + 32, v + 32);
        tmem_ld_wait();
        if (g == g_end - 1) {                      // accumulator drained: hand the TMEM buffer back early
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[as]);
        }
        ctx.next_col = g + 1 < g_end ? tc.n_blk * BN + (g + 1) * GROUP_COLS : -1;
        Epi::group(ep, st, shape, ctx, tc.n_blk * BN + g * GROUP_COLS, v);
      }
      if (g_begin == g_end) {                        // narrow tiles: this warp owns no column group
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[as]);
      }
      Epi::tile_end(ep, st, shape, ctx.row, tc.m_blk, tc.n_blk, half);
    }
    // persistent-state policies flush once: this CTA's fixed m-block is blockIdx.x % num_m_blocks
    Epi::kernel_end(ep, st, shape, (int)(blockIdx.x % shape.num_m_blocks) * BLOCK_M + quarter * 32 + lane,
                    (int)(blockIdx.x / shape.num_m_blocks) * 2 + half);
    if (lane == 0) tma_store_wait_all();            // staged tiles fully written before the CTA retires
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace micgemm
