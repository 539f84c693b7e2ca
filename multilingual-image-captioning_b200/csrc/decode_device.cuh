// Device-side building blocks of the cached decoder step, shared by the stand-alone kernels
// (attention.cu / elementwise.cu) and the persistent decoder-step kernel (decoder_step.cu).
#pragma once
#include "common.cuh"

namespace micdec {

constexpr int HD = 64;            // head dim
constexpr int DEC_MAX_KEYS = 256; // per-op cached attention: keys per row (model default max_length is 200)
constexpr int LN_MAX_ITERS = 4;   // features <= 1024 (32 lanes * 8 elements * 4)

__device__ __forceinline__ void load8(const bf16* p, float* x) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 f;
  f = unpack_bf16(u.x); x[0] = f.x; x[1] = f.y;
  f = unpack_bf16(u.y); x[2] = f.x; x[3] = f.y;
  f = unpack_bf16(u.z); x[4] = f.x; x[5] = f.y;
  f = unpack_bf16(u.w); x[6] = f.x; x[7] = f.y;
}
__device__ __forceinline__ void store8(bf16* p, const float* x) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]),
                                            pack_bf16(x[6], x[7]));
}
__device__ __forceinline__ void load8f(const float* p, float* x) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
// L2-coherent variants (ld.global.cg): data another SM wrote earlier in the SAME launch must not be served
// from this SM's L1
__device__ __forceinline__ void load8_cg(const bf16* p, float* x) {
  const uint4 u = __ldcg(reinterpret_cast<const uint4*>(p));
  float2 f;
  f = unpack_bf16(u.x); x[0] = f.x; x[1] = f.y;
  f = unpack_bf16(u.y); x[2] = f.x; x[3] = f.y;
  f = unpack_bf16(u.z); x[4] = f.x; x[5] = f.y;
  f = unpack_bf16(u.w); x[6] = f.x; x[7] = f.y;
}
__device__ __forceinline__ void load8f_cg(const float* p, float* x) {
  const float4 a = __ldcg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldcg(reinterpret_cast<const float4*>(p + 4));
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}

// "UMMA tile image" layout of a K-major bf16 activation matrix [rows, cols]: 128-row x 64-column tiles, each a
// contiguous 16 KB block that is exactly the SWIZZLE_128B shared-memory image tcgen05.mma reads (row r of the tile at
// r*128 B, its 16-byte chunk c stored at chunk c ^ (r & 7)); tiles ordered [row tile][k block].  A consumer fetches a
// whole operand stage with ONE bulk copy instead of a 128-row TMA box.  Returns the element offset of (row, col),
// col % 8 == 0; tiled_kb = cols / 64.
__device__ __forceinline__ long long tiled_off(int row, int col, int tiled_kb) {
  const int r = row & 127, c = (col & 63) >> 3;
  return ((long long)(row >> 7) * tiled_kb + (col >> 6)) * 8192 + r * 64 + ((c ^ (r & 7)) << 3);
}

// flax.linen.LayerNorm statistics: mean, E[x^2] - mean^2 (clamped at 0), biased
__device__ __forceinline__ void ln_stats(const float (*x)[8], int d, int lane, float* mean, float* rstd, float eps) {
  float s = 0.f, s2 = 0.f;
#pragma unroll
  for (int it = 0; it < LN_MAX_ITERS; ++it) {
    if (lane * 8 + it * 256 < d) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s += x[it][j];
        s2 += x[it][j] * x[it][j];
      }
    }
  }
  s = warp_sum(s);
  s2 = warp_sum(s2);
  const float m = s / d;
  const float var = fmaxf(s2 / d - m * m, 0.f);
  *mean = m;
  *rstd = rsqrtf(var + eps);
}

// x[row] += acc[row] + bias ; y[row] = LayerNorm(x[row]) ; acc[row] = 0.      One warp per row.
// kCoherent: acc/x were produced by other SMs during this launch (persistent kernel) -> L2 loads.
template <bool kCoherent>
__device__ __forceinline__ void residual_ln_row(float* __restrict__ acc, const float* __restrict__ bias,
                                                bf16* __restrict__ x, const float* __restrict__ gamma,
                                                const float* __restrict__ beta, float eps, bf16* __restrict__ y,
                                                int row, int d, int lane, int y_tiled_kb = 0) {
  float v[LN_MAX_ITERS][8];
#pragma unroll
  for (int it = 0; it < LN_MAX_ITERS; ++it) {
    const int c = lane * 8 + it * 256;
    if (c < d) {
      float a[8], b[8];
      if (kCoherent) {
        load8_cg(x + (long long)row * d + c, v[it]);
        load8f_cg(acc + (long long)row * d + c, a);
      } else {
        load8(x + (long long)row * d + c, v[it]);
        load8f(acc + (long long)row * d + c, a);
      }
      load8f(bias + c, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[it][j] = bf16_round(v[it][j] + a[j] + b[j]);
      store8(x + (long long)row * d + c, v[it]);
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(acc + (long long)row * d + c) = z;
      *reinterpret_cast<float4*>(acc + (long long)row * d + c + 4) = z;
    }
  }
  float mean, rstd;
  ln_stats(v, d, lane, &mean, &rstd, eps);
#pragma unroll
  for (int it = 0; it < LN_MAX_ITERS; ++it) {
    const int c = lane * 8 + it * 256;
    if (c < d) {
      float g[8], b[8], o[8];
      load8f(gamma + c, g);
      load8f(beta + c, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[it][j] - mean) * rstd * g[j] + b[j];
      store8(y + (y_tiled_kb ? tiled_off(row, c, y_tiled_kb) : (long long)row * d + c), o);
    }
  }
}

// Same op, written for CODE SIZE (persistent kernel: every phase starts on a cold instruction cache, so straight-
// line unrolled code costs an L2 round trip per 128 bytes of SASS): two rolled passes, the second re-reads the
// row it just wrote instead of keeping it in registers.
__device__ __forceinline__ void residual_ln_row_lean(float* __restrict__ acc, const float* __restrict__ bias,
                                                     bf16* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float eps, bf16* __restrict__ y,
                                                     int row, int d, int lane, int y_tiled_kb,
                                                     unsigned long long* trace = nullptr) {
  float s = 0.f, s2 = 0.f;
#pragma unroll 1
  for (int c = lane * 8; c < d; c += 256) {
    float v[8], a[8], b[8];
    load8_cg(x + (long long)row * d + c, v);
    if (trace && c == lane * 8) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(trace[0]) : "f"(v[0])); }
    load8f_cg(acc + (long long)row * d + c, a);
    if (trace && c == lane * 8) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(trace[1]) : "f"(a[0])); }
    load8f(bias + c, b);
    if (trace && c == lane * 8) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(trace[2]) : "f"(b[0])); }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] = bf16_round(v[j] + a[j] + b[j]);
      s += v[j];
      s2 += v[j] * v[j];
    }
    store8(x + (long long)row * d + c, v);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(acc + (long long)row * d + c) = z;
    *reinterpret_cast<float4*>(acc + (long long)row * d + c + 4) = z;
  }
  if (trace) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(trace[3]) : "f"(s)); }
  s = warp_sum(s);
  s2 = warp_sum(s2);
  const float mean = s / d;
  const float rstd = rsqrtf(fmaxf(s2 / d - mean * mean, 0.f) + eps);
  if (trace) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(trace[4]) : "f"(rstd)); }
#pragma unroll 1
  for (int c = lane * 8; c < d; c += 256) {
    float v[8], g[8], b[8], o[8];
    load8_cg(x + (long long)row * d + c, v);        // this thread's own store, read back through L2
    load8f(gamma + c, g);
    load8f(beta + c, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (v[j] - mean) * rstd * g[j] + b[j];
    store8(y + (y_tiled_kb ? tiled_off(row, c, y_tiled_kb) : (long long)row * d + c), o);
  }
}

// ---------------------------------------------------------------------------------------------
// Cached decode attention (1 query token per row), SURVEY.md A.3.
// Self-attention reads the K/V history of a beam through an ancestor table instead of physically
// reordering the cache (generation_clip_vision_utils.py:945-953 gathers 24 arrays every step):
//   slot(r, j) = anc[r*T + j]  = cache row that holds position j of beam-row r's history
// Cross-attention: K/V of the S visual tokens, shared by the beams of an image (row r -> r / beams).
// One warp per (row, head); lanes split keys for the scores and head-dim for the output.
// ---------------------------------------------------------------------------------------------
struct DecAttnArgs {
  const bf16* q;         // [R, ldq] (head h at h*64)
  long long ldq;
  const bf16 *kc, *vc;   // cache base: element (row, pos, h, d) at ((row*T + pos) * ldkv + h*64 + d)
  long long ldkv;
  const int* anc;        // [R, T] ancestor rows or null (identity)
  int T;                 // cache length (positions per row)
  int n_keys;            // keys to attend (cur position + 1) or S for cross (<= DEC_MAX_KEYS)
  int rows_per_kv;       // cross-attention: beams per image (kv row = r / rows_per_kv); 1 otherwise
  bf16* o;               // [R, ldo]
  long long ldo;
  int R, H;
  float scale;
  // persistent-kernel variant: q = bf16(q_acc + q_bias) taken from an fp32 split-K accumulator that is zeroed
  // again once read (q == nullptr then)
  float* q_acc;
  const float* q_bias;
  int o_tiled_kb;        // > 0: o is written in the UMMA tile image layout (tiled_off), = H
  // persistent-kernel cross-attention: K|V of every (kv row, head) pre-arranged as ONE contiguous 16 KB stage image
  // (K rows 0..63 then V rows 0..63, 128 B each, chunk c of row j at c ^ (j & 7), rows >= n_keys zero):
  // element offset = kv_row * kv_tiles_stride + head * 8192.  Null = gather from kc / vc with cp.async.
  const bf16* kv_tiles;
  long long kv_tiles_stride;
};

// s_p: DEC_MAX_KEYS floats, s_row: DEC_MAX_KEYS ints of per-warp shared scratch.  q / k / v go through L2 (ld.global.cg): in the
// persistent kernel the newest position was written by another SM within the same launch.
__device__ __forceinline__ void decode_attn_item(const DecAttnArgs& a, int r, int h, float* s_p, int* s_row,
                                                 int lane) {
  const int nk = a.n_keys;
  const int kvrow_default = r / a.rows_per_kv;
  for (int j = lane; j < nk; j += 32) s_row[j] = a.anc ? a.anc[(long long)r * a.T + j] : kvrow_default;
  // q in registers: every lane holds the full 64-d query, pre-scaled
  float qv[HD];
  if (a.q != nullptr) {
    const bf16* qp = a.q + (long long)r * a.ldq + h * HD;
#pragma unroll
    for (int c = 0; c < HD; c += 8) {
      float f[8];
      load8_cg(qp + c, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) qv[c + j] = f[j] * a.scale;
    }
  } else {
    float* qa = a.q_acc + (long long)r * a.ldq + h * HD;
    const float* qb = a.q_bias + h * HD;
#pragma unroll
    for (int c = 0; c < HD; c += 8) {
      float f[8], b[8];
      load8f_cg(qa + c, f);
      load8f(qb + c, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) qv[c + j] = bf16_round(f[j] + b[j]) * a.scale;
    }
    __syncwarp();
    // hand the accumulator back zeroed (each lane clears 2 of the 64 floats)
    *reinterpret_cast<float2*>(qa + lane * 2) = make_float2(0.f, 0.f);
  }
  __syncwarp();
  // scores: lane handles keys lane, lane+32, ... (8 independent 16-byte loads per key)
  float mx = -INFINITY;
  float sc[DEC_MAX_KEYS / 32];
#pragma unroll
  for (int i = 0; i < DEC_MAX_KEYS / 32; ++i) {
    const int j = lane + i * 32;
    float sdot = -INFINITY;
    if (j < nk) {
      const bf16* kp = a.kc + ((long long)s_row[j] * a.T + j) * a.ldkv + h * HD;
      uint4 kk[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) kk[c] = __ldcg(reinterpret_cast<const uint4*>(kp + c * 8));
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t* uu = reinterpret_cast<const uint32_t*>(&kk[c]);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const float2 f = unpack_bf16(uu[jj]);
          acc = fmaf(qv[c * 8 + 2 * jj], f.x, acc);
          acc = fmaf(qv[c * 8 + 2 * jj + 1], f.y, acc);
        }
      }
      sdot = acc;
    }
    sc[i] = sdot;
    mx = fmaxf(mx, sdot);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < DEC_MAX_KEYS / 32; ++i) {
    const int j = lane + i * 32;
    const float p = (j < nk) ? __expf(sc[i] - mx) : 0.f;
    if (j < nk) s_p[j] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  __syncwarp();
  // P.V: 4 key groups x 8 dim groups; lane (kg, dl) accumulates dims dl*8..dl*8+7 over keys kg, kg+4, ...
  const int kg = lane >> 3, dl = lane & 7;
  float o[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = 0.f;
  for (int j = kg; j < nk; j += 4) {
    const bf16* vp = a.vc + ((long long)s_row[j] * a.T + j) * a.ldkv + h * HD + dl * 8;
    const uint4 u = __ldcg(reinterpret_cast<const uint4*>(vp));
    const float p = s_p[j];
    const uint32_t* uu = reinterpret_cast<const uint32_t*>(&u);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const float2 f = unpack_bf16(uu[jj]);
      o[2 * jj] = fmaf(p, f.x, o[2 * jj]);
      o[2 * jj + 1] = fmaf(p, f.y, o[2 * jj + 1]);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    o[e] += __shfl_xor_sync(0xffffffffu, o[e], 8);
    o[e] += __shfl_xor_sync(0xffffffffu, o[e], 16);
  }
  if (kg == 0) {
    bf16* dst = a.o + (a.o_tiled_kb ? tiled_off(r, h * HD + dl * 8, a.o_tiled_kb) : (long long)r * a.ldo + h * HD + dl * 8);
    *reinterpret_cast<uint4*>(dst) =
        make_uint4(pack_bf16(o[0] * inv, o[1] * inv), pack_bf16(o[2] * inv, o[3] * inv),
                   pack_bf16(o[4] * inv, o[5] * inv), pack_bf16(o[6] * inv, o[7] * inv));
  }
  __syncwarp();   // scratch is reused by the warp's next item
}

// ---------------------------------------------------------------------------------------------
// Staged variant for the persistent kernel, where only 8 warps per SM are available to hide latency: ALL K and V
// rows of an item are pulled into a per-warp 16 KB shared-memory stage with 16-byte cp.async copies issued at once
// (one round trip instead of a dependent chain), and an item covers the `rows_per_kv` query rows that share the
// same K/V (cross-attention: the beams of an image), so shared K/V are read once.  n_keys <= 64.
//   kv_smem: K row j at j*128, V row j at 8192 + j*128; 16-byte chunk c of row j is stored at chunk c ^ (j & 7)
//   q_smem : [rows_per_kv (<= 4), 64] pre-scaled queries;   p_smem: [64] softmax numerators
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void decode_attn_group_staged(const DecAttnArgs& a, int kv_item, int h, uint8_t* kv_smem,
                                                         float* q_smem, float* p_smem, int lane) {
  const int nk = a.n_keys;
  const int rpk = a.rows_per_kv;
  const int r0 = kv_item * rpk;
  int row0 = kv_item, row1 = kv_item;          // cache rows of keys lane, lane + 32
  if (a.anc) {
    row0 = lane < nk ? a.anc[(long long)r0 * a.T + lane] : 0;
    row1 = lane + 32 < nk ? a.anc[(long long)r0 * a.T + 32 + lane] : 0;
  }
  const uint32_t kv_base = smem_u32(kv_smem);
  const int total = nk * 16;
  for (int base = 0; base < total; base += 32) {
    const int idx = base + lane;
    const int j = idx >> 4, c = idx & 15;
    const int rj0 = __shfl_sync(0xffffffffu, row0, j & 31), rj1 = __shfl_sync(0xffffffffu, row1, j & 31);
    if (idx < total) {
      const int rj = j < 32 ? rj0 : rj1;
      const int cc = c & 7;
      const bf16* src = (c < 8 ? a.kc : a.vc) + ((long long)rj * a.T + j) * a.ldkv + h * HD + cc * 8;
      cp_async_16(kv_base + (c < 8 ? 0 : 8192) + j * 128 + ((cc ^ (j & 7)) << 4), src);
    }
  }
  cp_async_commit();
  // queries of the item's rows -> smem, pre-scaled (8 floats per lane, lanes >= 8*rpk idle)
  if (lane < 8 * rpk && r0 + (lane >> 3) < a.R) {
    const int rr = lane >> 3, c8 = (lane & 7) * 8;
    const int r = r0 + rr;
    float f[8];
    if (a.q != nullptr) {
      load8_cg(a.q + (long long)r * a.ldq + h * HD + c8, f);
    } else {
      float* qa = a.q_acc + (long long)r * a.ldq + h * HD + c8;
      float b[8];
      load8f_cg(qa, f);
      load8f(a.q_bias + h * HD + c8, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = bf16_round(f[j] + b[j]);
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(qa) = z;                  // hand the split-K accumulator back zeroed
      *reinterpret_cast<float4*>(qa + 4) = z;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) q_smem[rr * HD + c8 + j] = f[j] * a.scale;
  }
  cp_async_wait_all();
  __syncwarp();
  const int kg = lane >> 3, dl = lane & 7;
  for (int rr = 0; rr < rpk; ++rr) {
    const int r = r0 + rr;
    if (r >= a.R) break;
    const float4* q4 = reinterpret_cast<const float4*>(q_smem + rr * HD);
    float sc[2];
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int j = lane + t * 32;
      float acc = -INFINITY;
      if (j < nk) {
        acc = 0.f;
        const uint8_t* krow = kv_smem + j * 128;
#pragma unroll 2
        for (int c = 0; c < 8; ++c) {
          const uint4 kk = *reinterpret_cast<const uint4*>(krow + ((c ^ (j & 7)) << 4));
          const float4 qa = q4[2 * c], qb = q4[2 * c + 1];
          float2 f;
          f = unpack_bf16(kk.x); acc = fmaf(qa.x, f.x, acc); acc = fmaf(qa.y, f.y, acc);
          f = unpack_bf16(kk.y); acc = fmaf(qa.z, f.x, acc); acc = fmaf(qa.w, f.y, acc);
          f = unpack_bf16(kk.z); acc = fmaf(qb.x, f.x, acc); acc = fmaf(qb.y, f.y, acc);
          f = unpack_bf16(kk.w); acc = fmaf(qb.z, f.x, acc); acc = fmaf(qb.w, f.y, acc);
        }
      }
      sc[t] = acc;
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int j = lane + t * 32;
      const float p = (j < nk) ? __expf(sc[t] - mx) : 0.f;
      if (j < nk) p_smem[j] = p;
      sum += p;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    __syncwarp();
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = 0.f;
    for (int j = kg; j < nk; j += 4) {
      const uint4 u = *reinterpret_cast<const uint4*>(kv_smem + 8192 + j * 128 + ((dl ^ (j & 7)) << 4));
      const float p = p_smem[j];
      float2 f;
      f = unpack_bf16(u.x); o[0] = fmaf(p, f.x, o[0]); o[1] = fmaf(p, f.y, o[1]);
      f = unpack_bf16(u.y); o[2] = fmaf(p, f.x, o[2]); o[3] = fmaf(p, f.y, o[3]);
      f = unpack_bf16(u.z); o[4] = fmaf(p, f.x, o[4]); o[5] = fmaf(p, f.y, o[5]);
      f = unpack_bf16(u.w); o[6] = fmaf(p, f.x, o[6]); o[7] = fmaf(p, f.y, o[7]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      o[e] += __shfl_xor_sync(0xffffffffu, o[e], 8);
      o[e] += __shfl_xor_sync(0xffffffffu, o[e], 16);
    }
    if (kg == 0) {
      bf16* dst = a.o + (a.o_tiled_kb ? tiled_off(r, h * HD + dl * 8, a.o_tiled_kb) : (long long)r * a.ldo + h * HD + dl * 8);
      *reinterpret_cast<uint4*>(dst) =
          make_uint4(pack_bf16(o[0] * inv, o[1] * inv), pack_bf16(o[2] * inv, o[3] * inv),
                     pack_bf16(o[4] * inv, o[5] * inv), pack_bf16(o[6] * inv, o[7] * inv));
    }
    __syncwarp();   // p_smem is rewritten by the next row
  }
  __syncwarp();     // the stage is refilled by the warp's next item
}

// ---------------------------------------------------------------------------------------------
// Tensor-core variant of the staged item (persistent kernel).  A lone warp per item is instruction-latency bound
// (~5 cycles per dependent instruction, nothing to overlap with), so the ~2400 scalar instructions of the SIMT
// version cost more than its memory traffic; here S = Q K^T and O = P V are warp-level m16n8k16 bf16 MMAs fed by
// ldmatrix straight from the swizzled K / V stage (~250 instructions per item).  The item's <= 4 query rows sit
// in rows 0..3 of the 16-row MMA tile; rows 8..15 are hard zeros (their A registers), rows 4..7 are zero-filled.
//   q_stage: bf16 [8][72] (144-byte pitch: conflict-free fragment loads)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_16816(float* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

#define MIC_TRACE(slot, dep)                                                                          \
  if (trace) {                                                                                        \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(trace[slot]) : "r"(dep));                         \
  }
// S = Q K^T and the row softmax of the item's <= 8 query rows (stage row g = lane >> 2) over nk <= 64 keys.
// Returns the un-normalised probabilities packed as the A fragments of the P V MMAs and 1 / row sum.
__device__ __forceinline__ void attn_qk_softmax(uint32_t kv_base, const bf16* q_stage, int nk, int lane, uint32_t* pa,
                                                float* inv_out) {
  constexpr int QP = 72;
  const int g = lane >> 2, t4 = lane & 3;
  const int npairs = (nk + 15) >> 4;           // 16-key groups
  float sacc[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) sacc[n][0] = sacc[n][1] = sacc[n][2] = sacc[n][3] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const uint32_t a0 = *reinterpret_cast<const uint32_t*>(q_stage + g * QP + ks * 16 + 2 * t4);
    const uint32_t a2 = *reinterpret_cast<const uint32_t*>(q_stage + g * QP + ks * 16 + 8 + 2 * t4);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      if (np < npairs) {
        const int key = np * 16 + (lane >> 4) * 8 + (lane & 7);
        const int chunk = ks * 2 + ((lane >> 3) & 1);
        uint32_t b0, b1, b2, b3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                     : "r"(kv_base + key * 128 + ((chunk ^ (key & 7)) << 4)));
        mma_16816(sacc[2 * np], a0, 0u, a2, 0u, b0, b1);
        mma_16816(sacc[2 * np + 1], a0, 0u, a2, 0u, b2, b3);
      }
    }
  }
  float mx = -INFINITY;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int col = n * 8 + 2 * t4;
    if (col >= nk) sacc[n][0] = -INFINITY;
    if (col + 1 >= nk) sacc[n][1] = -INFINITY;
    mx = fmaxf(mx, fmaxf(sacc[n][0], sacc[n][1]));
  }
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  float sum = 0.f;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const float p0 = __expf(sacc[n][0] - mx), p1 = __expf(sacc[n][1] - mx);     // exp(-inf) = 0 for masked keys
    sum += p0 + p1;
    pa[n] = pack_bf16(p0, p1);
  }
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);
  *inv_out = 1.0f / sum;
}
// O = P V (the S accumulator layout is the A-fragment layout of the next MMA, k = keys) and the bf16 store of the
// rows_per_kv valid rows
__device__ __forceinline__ void attn_pv_store(const DecAttnArgs& a, uint32_t kv_base, int nk, const uint32_t* pa,
                                              float inv, int r0, int h, int lane) {
  const int g = lane >> 2, t4 = lane & 3;
  const int npairs = (nk + 15) >> 4;
  float oacc[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) oacc[n][0] = oacc[n][1] = oacc[n][2] = oacc[n][3] = 0.f;
#pragma unroll
  for (int kp = 0; kp < 4; ++kp) {
    if (kp < npairs) {
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        const int mi = lane >> 3;
        const int key = kp * 16 + (mi & 1) * 8 + (lane & 7);
        const int chunk = dp * 2 + (mi >> 1);
        uint32_t b0, b1, b2, b3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                     : "r"(kv_base + 8192 + key * 128 + ((chunk ^ (key & 7)) << 4)));
        mma_16816(oacc[2 * dp], pa[2 * kp], 0u, pa[2 * kp + 1], 0u, b0, b1);
        mma_16816(oacc[2 * dp + 1], pa[2 * kp], 0u, pa[2 * kp + 1], 0u, b2, b3);
      }
    }
  }
  if (g < a.rows_per_kv && r0 + g < a.R) {
    const int r = r0 + g;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      bf16* dst = a.o + (a.o_tiled_kb ? tiled_off(r, h * HD + n * 8, a.o_tiled_kb) : (long long)r * a.ldo + h * HD + n * 8) +
                  2 * t4;
      *reinterpret_cast<uint32_t*>(dst) = pack_bf16(oacc[n][0] * inv, oacc[n][1] * inv);
    }
  }
}

// cache rows of keys (lane, lane + 32) of an item: loaded one item ahead by the caller (hides the table latency)
__device__ __forceinline__ void decode_attn_rows(const DecAttnArgs& a, int kv_item, int lane, int* row0, int* row1) {
  *row0 = kv_item;
  *row1 = kv_item;
  if (a.anc) {
    const int r0 = kv_item * a.rows_per_kv;
    *row0 = lane < a.n_keys ? a.anc[(long long)r0 * a.T + lane] : 0;
    *row1 = lane + 32 < a.n_keys ? a.anc[(long long)r0 * a.T + 32 + lane] : 0;
  }
}

__device__ __forceinline__ void decode_attn_group_mma(const DecAttnArgs& a, int kv_item, int h, int row0, int row1,
                                                      uint8_t* kv_smem, bf16* q_stage, int lane,
                                                      unsigned long long* trace = nullptr) {
  constexpr int QP = 72;                       // q_stage row pitch in elements
  const int nk = a.n_keys;
  const int rpk = a.rows_per_kv;
  const int r0 = kv_item * rpk;
  MIC_TRACE(0, row0 + row1)
  // ---- query loads first (their latency hides behind the copy-issue loop): 64 16-byte chunks, 2 per lane
  float qf[2][8];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int id = lane + t * 32;
    const int rr = id >> 3, c8 = (id & 7) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) qf[t][j] = 0.f;
    if (rr < rpk && r0 + rr < a.R) {
      const int r = r0 + rr;
      if (a.q != nullptr) {
        load8_cg(a.q + (long long)r * a.ldq + h * HD + c8, qf[t]);
      } else {
        load8f_cg(a.q_acc + (long long)r * a.ldq + h * HD + c8, qf[t]);
      }
    }
  }
  // ---- K / V rows -> stage: lanes 0..15 take the 16 chunks (8 K + 8 V) of an even key, lanes 16..31 of the odd
  // key next to it; one shuffle and 32-bit offset arithmetic per copy (this loop is instruction-latency bound)
  const uint32_t kv_base = smem_u32(kv_smem);
  {
    const int c = lane & 15, jsub = lane >> 4, cc = c & 7;
    const bf16* src_base = (c < 8 ? a.kc : a.vc) + h * HD + cc * 8;
    const uint32_t dst_base = kv_base + (c < 8 ? 0 : 8192);
    const int ldkv = (int)a.ldkv, rowpitch = a.T * (int)a.ldkv;
#pragma unroll 4
    for (int j2 = 0; j2 < nk; j2 += 2) {
      const int j = j2 + jsub;
      const int rj = __shfl_sync(0xffffffffu, j2 < 32 ? row0 : row1, j & 31);
      if (j < nk) cp_async_16(dst_base + j * 128 + ((cc ^ (j & 7)) << 4), src_base + (rj * rowpitch + j * ldkv));
    }
  }
  cp_async_commit();
  MIC_TRACE(1, nk)
  // V rows nk .. next multiple of 16 take part in the P.V MMAs with P = 0: they must be finite
  {
    const int pad_rows = ((nk + 15) & ~15) - nk;
    for (int i = lane; i < pad_rows * 8; i += 32)
      *reinterpret_cast<uint4*>(kv_smem + 8192 + (nk + (i >> 3)) * 128 + ((i & 7) << 4)) = make_uint4(0, 0, 0, 0);
  }
  // queries -> bf16 stage rows 0..rpk-1 (pre-scaled: 1/8 is exact), zero rows rpk..7
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int id = lane + t * 32;
    const int rr = id >> 3, c8 = (id & 7) * 8;
    if (a.q == nullptr && rr < rpk && r0 + rr < a.R) {
      float* qa = a.q_acc + (long long)(r0 + rr) * a.ldq + h * HD + c8;
      float b[8];
      load8f(a.q_bias + h * HD + c8, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) qf[t][j] = bf16_round(qf[t][j] + b[j]);
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(qa) = z;                  // hand the split-K accumulator back zeroed
      *reinterpret_cast<float4*>(qa + 4) = z;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) qf[t][j] *= a.scale;
    store8(q_stage + rr * QP + c8, qf[t]);
  }
  const int total = nk;
  MIC_TRACE(2, total)
  cp_async_wait_all();
  __syncwarp();
  MIC_TRACE(3, total)
  uint32_t pa[8];
  float inv;
  attn_qk_softmax(kv_base, q_stage, nk, lane, pa, &inv);
  MIC_TRACE(4, __float_as_int(inv))
  attn_pv_store(a, kv_base, nk, pa, inv, r0, h, lane);
  MIC_TRACE(5, nk)
  __syncwarp();     // the stages are refilled by the warp's next item
}

// ---------------------------------------------------------------------------------------------
// Self-attention items of one warp (rows_per_kv == 1), software-pipelined on the ONE 16 KB stage: the next item's
// K rows are requested (cp.async) as soon as S = Q K^T has consumed the stage's K half, its V rows after P V, so an
// item's HBM round trip runs under the previous item's math instead of in front of its own.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void attn_issue_operand(const DecAttnArgs& a, const bf16* base, int h, int row0, int row1,
                                                   uint32_t dst_base, int lane) {
  // 4 keys per instruction: lane = (key offset 0..3, 16-byte chunk 0..7)
  const int nk = a.n_keys;
  const int cc = lane & 7, jsub = lane >> 3;
  const bf16* src_base = base + h * HD + cc * 8;
  const int ldkv = (int)a.ldkv, rowpitch = a.T * (int)a.ldkv;
#pragma unroll 4
  for (int j4 = 0; j4 < nk; j4 += 4) {
    const int j = j4 + jsub;
    const int rj = __shfl_sync(0xffffffffu, j4 < 32 ? row0 : row1, j & 31);
    if (j < nk) cp_async_16(dst_base + j * 128 + ((cc ^ (j & 7)) << 4), src_base + (rj * rowpitch + j * ldkv));
  }
}

__device__ __forceinline__ void decode_self_attn_pipelined(const DecAttnArgs& a, int first, int stride, int items,
                                                           uint8_t* kv_smem, bf16* q_stage, int lane) {
  constexpr int QP = 72;
  const int nk = a.n_keys;
  const uint32_t kv_base = smem_u32(kv_smem);
  int i = first;
  if (i >= items) return;
  {   // V rows nk .. next multiple of 16 take part in the P.V MMAs with P = 0: they must be finite (never overwritten)
    const int pad_rows = ((nk + 15) & ~15) - nk;
    for (int t = lane; t < pad_rows * 8; t += 32)
      *reinterpret_cast<uint4*>(kv_smem + 8192 + (nk + (t >> 3)) * 128 + ((t & 7) << 4)) = make_uint4(0, 0, 0, 0);
  }
  int row0, row1;
  decode_attn_rows(a, i / a.H, lane, &row0, &row1);
  attn_issue_operand(a, a.kc, i % a.H, row0, row1, kv_base, lane);
  attn_issue_operand(a, a.vc, i % a.H, row0, row1, kv_base + 8192, lane);
  cp_async_commit();
  while (true) {
    const int r = i / a.H, h = i % a.H;
    const int inext = i + stride;
    const bool has_next = inext < items;
    int nrow0 = 0, nrow1 = 0;
    if (has_next) decode_attn_rows(a, inext / a.H, lane, &nrow0, &nrow1);
    // query row -> bf16 stage row 0 (pre-scaled: 1/8 is exact), rows 1..7 zero: 64 chunks, 2 per lane
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int id = lane + t * 32;
      const int rr = id >> 3, c8 = (id & 7) * 8;
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = 0.f;
      if (rr == 0) {
        load8_cg(a.q + (long long)r * a.ldq + h * HD + c8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] *= a.scale;
      }
      store8(q_stage + rr * QP + c8, f);
    }
    cp_async_wait_all();
    __syncwarp();
    uint32_t pa[8];
    float inv;
    attn_qk_softmax(kv_base, q_stage, nk, lane, pa, &inv);
    __syncwarp();                                   // every lane is done with the K half (and the query stage)
    if (has_next) {
      attn_issue_operand(a, a.kc, inext % a.H, nrow0, nrow1, kv_base, lane);
      cp_async_commit();
    }
    attn_pv_store(a, kv_base, nk, pa, inv, r, h, lane);
    __syncwarp();                                   // ... and with the V half
    if (!has_next) break;
    attn_issue_operand(a, a.vc, inext % a.H, nrow0, nrow1, kv_base + 8192, lane);
    cp_async_commit();
    i = inext;
  }
}

// ---------------------------------------------------------------------------------------------
// Self-attention items of one warp on the HEAD-MAJOR cache of the persistent kernel (round 2).
//
// Cache layout per layer: [row][head][K plane | V plane][pos][64] bf16 (row pitch T*2*H*64 elements, as before), the
// eight 16-byte chunks of a 128-byte (pos) row stored at chunk c ^ (pos & 7) — i.e. exactly the stage image the MMA
// path reads.  The K (or V) history of ONE cache row and head is therefore contiguous over positions, and a beam's
// history is a handful of RUNS of positions that live in the same cache row (its ancestry changes rows only where the
// beam search re-parented it).  An item's operand is fetched with one bulk copy per run (typically 1-4) instead of
// 2 * n_keys * 8 = ~650 sixteen-byte cp.async whose issue alone cost 1.3 us per item (profiles/r01_decoder_step_phases
// .txt: 18 us per self-attention phase).  Worst case (ancestry alternating every position) degrades to one 128-byte
// copy per key: correct, only slower.
//   kc = layer base of the cache, T = cache length; rows of keys (lane, lane+32) come from decode_attn_rows().
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void attn_issue_runs(const bf16* plane, long long rowpitch, int nk, int row0, int row1,
                                                uint8_t* dst, uint64_t* bar, int lane) {
  const int prev0 = __shfl_up_sync(0xffffffffu, row0, 1);
  int prev1 = __shfl_up_sync(0xffffffffu, row1, 1);
  const int last0 = __shfl_sync(0xffffffffu, row0, 31);
  if (lane == 0) prev1 = last0;
  const bool s0 = lane < nk && (lane == 0 || row0 != prev0);
  const bool s1 = lane + 32 < nk && row1 != prev1;
  const unsigned long long m = (unsigned long long)__ballot_sync(0xffffffffu, s0) |
                               ((unsigned long long)__ballot_sync(0xffffffffu, s1) << 32);
  if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)nk * 128u);
  __syncwarp();
  if (s0) {
    const int k = lane;
    const unsigned long long rest = m >> (k + 1);
    const int len = rest ? __ffsll((long long)rest) : nk - k;
    bulk_load(dst + k * 128, plane + (long long)row0 * rowpitch + k * 64, (uint32_t)len * 128u, bar);
  }
  if (s1) {
    const int k = lane + 32;
    const unsigned long long rest = k + 1 < 64 ? m >> (k + 1) : 0ull;
    const int len = rest ? __ffsll((long long)rest) : nk - k;
    bulk_load(dst + k * 128, plane + (long long)row1 * rowpitch + k * 64, (uint32_t)len * 128u, bar);
  }
}

__device__ __forceinline__ void decode_self_attn_runs(const DecAttnArgs& a, int first, int stride, int items,
                                                      uint8_t* kv_smem, bf16* q_stage, uint64_t* bar_k, uint64_t* bar_v,
                                                      uint32_t* par_k, uint32_t* par_v, int lane) {
  constexpr int QP = 72;
  const int nk = a.n_keys;
  const uint32_t kv_base = smem_u32(kv_smem);
  const long long rowpitch = (long long)a.T * 2 * a.H * HD;
  const int plane = a.T * HD;                      // elements per (head, K or V) plane
  int i = first;
  if (i >= items) return;
  {   // V rows nk .. next multiple of 16 take part in the P.V MMAs with P = 0: they must be finite (never overwritten)
    const int pad_rows = ((nk + 15) & ~15) - nk;
    for (int t = lane; t < pad_rows * 8; t += 32)
      *reinterpret_cast<uint4*>(kv_smem + 8192 + (nk + (t >> 3)) * 128 + ((t & 7) << 4)) = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();                             // this warp's earlier generic use of the stage precedes the copies
  __syncwarp();
  int row0, row1;
  decode_attn_rows(a, i / a.H, lane, &row0, &row1);
  {
    const bf16* hb = a.kc + (long long)(i % a.H) * 2 * plane;
    attn_issue_runs(hb, rowpitch, nk, row0, row1, kv_smem, bar_k, lane);
    attn_issue_runs(hb + plane, rowpitch, nk, row0, row1, kv_smem + 8192, bar_v, lane);
  }
  // query stage: row 0 = the item's query (pre-scaled: 1/8 is exact), rows 1..7 stay zero for every item of the phase.
  // Lanes 0..7 own the 8 chunks of row 0; the NEXT item's query is loaded one item ahead (its L2 round trip runs
  // under this item's S = Q K^T instead of in front of the next one).
  float qn[8];
  {
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int id = lane + t * 32;
      if (id >= 8) *reinterpret_cast<uint4*>(q_stage + (id >> 3) * QP + (id & 7) * 8) = make_uint4(0, 0, 0, 0);
    }
    if (lane < 8) load8_cg(a.q + (long long)(i / a.H) * a.ldq + (i % a.H) * HD + lane * 8, qn);
  }
  while (true) {
    const int r = i / a.H, h = i % a.H;
    const int inext = i + stride;
    const bool has_next = inext < items;
    int nrow0 = 0, nrow1 = 0;
    if (lane < 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) qn[j] *= a.scale;
      store8(q_stage + lane * 8, qn);
    }
    if (has_next) {
      decode_attn_rows(a, inext / a.H, lane, &nrow0, &nrow1);
      if (lane < 8) load8_cg(a.q + (long long)(inext / a.H) * a.ldq + (inext % a.H) * HD + lane * 8, qn);
    }
    __syncwarp();
    mbar_wait(bar_k, *par_k);
    *par_k ^= 1;
    uint32_t pa[8];
    float inv;
    attn_qk_softmax(kv_base, q_stage, nk, lane, pa, &inv);
    fence_proxy_async();                            // generic reads of the K half precede the next item's bulk copies
    __syncwarp();
    const bf16* hb = a.kc + (long long)(inext % a.H) * 2 * plane;
    if (has_next) attn_issue_runs(hb, rowpitch, nk, nrow0, nrow1, kv_smem, bar_k, lane);
    mbar_wait(bar_v, *par_v);
    *par_v ^= 1;
    attn_pv_store(a, kv_base, nk, pa, inv, r, h, lane);
    fence_proxy_async();
    __syncwarp();
    if (!has_next) break;
    attn_issue_runs(hb + plane, rowpitch, nk, nrow0, nrow1, kv_smem + 8192, bar_v, lane);
    i = inext;
  }
}

// ---------------------------------------------------------------------------------------------
// Cross-attention item on pre-packed K|V stage images: the whole 16 KB stage arrives with ONE bulk copy (vs 800
// 16-byte cp.async), the query rows of the item are loaded while it is in flight.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void decode_cross_attn_packed(const DecAttnArgs& a, int kv_item, int h, uint8_t* kv_smem,
                                                         bf16* q_stage, uint64_t* bar, uint32_t* parity, int lane) {
  constexpr int QP = 72;
  const int nk = a.n_keys;
  const int rpk = a.rows_per_kv;
  const int r0 = kv_item * rpk;
  if (lane == 0) {
    mbar_arrive_expect_tx(bar, 16384);
    bulk_load(kv_smem, a.kv_tiles + (long long)kv_item * a.kv_tiles_stride + (long long)h * 8192, 16384, bar);
  }
  // queries -> bf16 stage rows 0..rpk-1 (pre-scaled: 1/8 is exact), zero rows rpk..7: 64 16-byte chunks, 2 per lane
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int id = lane + t * 32;
    const int rr = id >> 3, c8 = (id & 7) * 8;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = 0.f;
    if (rr < rpk && r0 + rr < a.R) {
      const int r = r0 + rr;
      if (a.q != nullptr) {
        load8_cg(a.q + (long long)r * a.ldq + h * HD + c8, f);
      } else {
        float* qa = a.q_acc + (long long)r * a.ldq + h * HD + c8;
        float b[8];
        load8f_cg(qa, f);
        load8f(a.q_bias + h * HD + c8, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = bf16_round(f[j] + b[j]);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(qa) = z;                  // hand the split-K accumulator back zeroed
        *reinterpret_cast<float4*>(qa + 4) = z;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= a.scale;
    }
    store8(q_stage + rr * QP + c8, f);
  }
  __syncwarp();
  mbar_wait(bar, *parity);
  *parity ^= 1;
  const uint32_t kv_base = smem_u32(kv_smem);
  uint32_t pa[8];
  float inv;
  attn_qk_softmax(kv_base, q_stage, nk, lane, pa, &inv);
  attn_pv_store(a, kv_base, nk, pa, inv, r0, h, lane);
  __syncwarp();
  fence_proxy_async();      // generic reads of the stage are ordered before the next item's bulk copy into it
}

// (A software-pipelined variant that fetched every 128-byte key row with its own cp.async.bulk - cache rows stored
//  pre-swizzled, next item's K requested right after S = Q K^T - was measured SLOWER: 24 us vs 20 us per
//  self-attention phase; ~650 small bulk copies per SM per round queue up in the single TMA unit.)

}  // namespace micdec
