// HBM-bound kernels of the captioning path: LayerNorm fwd/bwd, embedding gather/scatter, patchify,
// activation backward + bias-gradient column sums, CE finalisation, AdamW.  All use 16-byte vector
// access on the contiguous feature dimension and warp-shuffle reductions; fp32 statistics.
#include "common.cuh"
#include "decode_device.cuh"

#include "../../include/mic_b200.h"

namespace {

using namespace micdec;

// ---------------------------------------------------------------------------------------------
// LayerNorm forward.  One warp per row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps,
                                                            bf16* __restrict__ y, float* __restrict__ mean_out,
                                                            float* __restrict__ rstd_out, int M, int d) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  float v[LN_MAX_ITERS][8];
#pragma unroll
  for (int it = 0; it < LN_MAX_ITERS; ++it) {
    const int c = lane * 8 + it * 256;
    if (c < d) load8(x + (long long)row * d + c, v[it]);
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
      store8(y + (long long)row * d + c, o);
    }
  }
  if (lane == 0 && mean_out) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
}

// ---------------------------------------------------------------------------------------------
// Decode-time fusion: x += acc + bias ; y = LayerNorm(x) ; acc = 0.
// `acc` is the fp32 split-K accumulator a residual GEMM reduce-added into (TMA reduce-add); zeroing it
// here keeps it ready for the next GEMM without a memset launch.  One warp per row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
residual_ln_fwd_kernel(float* __restrict__ acc, const float* __restrict__ bias, bf16* __restrict__ x,
                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                       bf16* __restrict__ y, int M, int d) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  residual_ln_row<false>(acc, bias, x, gamma, beta, eps, y, row, d, lane);
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward, two kernels:
//  (1) dx = dres + rstd * (dy*g - mean(dy*g) - xhat * mean(dy*g*xhat))   one warp per row, lean registers
//  (2) dgamma = sum_rows dy*xhat, dbeta = sum_rows dy : column-parallel over row chunks; the last CTA of a
//      column block (atomic ticket) folds the per-chunk partials, so no follow-up reduction launch.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
layernorm_bwd_dx_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ gamma,
                        const float* __restrict__ mean, const float* __restrict__ rstd, const bf16* __restrict__ dres,
                        bf16* __restrict__ dx, int M, int d) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float mu = mean[row], rs = rstd[row];
  float dyg[LN_MAX_ITERS][8], xh[LN_MAX_ITERS][8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int it = 0; it < LN_MAX_ITERS; ++it) {
    const int c = lane * 8 + it * 256;
    if (c < d) {
      float g[8];
      load8(dy + (long long)row * d + c, dyg[it]);
      load8(x + (long long)row * d + c, xh[it]);
      load8f(gamma + c, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh[it][j] = (xh[it][j] - mu) * rs;
        dyg[it][j] *= g[j];
        s1 += dyg[it][j];
        s2 = fmaf(dyg[it][j], xh[it][j], s2);
      }
    }
  }
  s1 = warp_sum(s1) / d;
  s2 = warp_sum(s2) / d;
#pragma unroll
  for (int it = 0; it < LN_MAX_ITERS; ++it) {
    const int c = lane * 8 + it * 256;
    if (c < d) {
      float o[8], r[8];
      if (dres) load8(dres + (long long)row * d + c, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j] = rs * (dyg[it][j] - s1 - xh[it][j] * s2);
        if (dres) o[j] += r[j];
      }
      store8(dx + (long long)row * d + c, o);
    }
  }
}

// "last CTA reduces": after writing its partial row, a CTA takes a ticket; the CTA that draws the last
// ticket of its column block sums all partials of that block and resets the counter for the next launch.
__device__ __forceinline__ bool last_cta_of_column(unsigned int* counter, unsigned int total) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(counter, 1u);
    is_last = (t == total - 1);
    if (is_last) *counter = 0;
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

// grid = (ceil(d/256), chunks); block 256 = 32 column groups (8 cols) x 8 row lanes
__global__ void __launch_bounds__(256)
ln_param_grad_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ mean,
                     const float* __restrict__ rstd, float* __restrict__ part, unsigned int* __restrict__ counters,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, int M, int d, int rows_per_chunk) {
  __shared__ float red[8][2][256 + 8];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + cg * 8;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(M, r0 + rows_per_chunk);
  float ag[8], ab[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ag[j] = 0.f;
    ab[j] = 0.f;
  }
  if (c < d) {
    for (int rb = r0 + rl; rb < r1; rb += 16) {           // two rows per trip, all four loads issued up front
      const int rn = rb + 8;
      const bool two = rn < r1;
      float g[8], xv[8], g2[8], xv2[8];
      load8(dy + (long long)rb * d + c, g);
      load8(x + (long long)rb * d + c, xv);
      if (two) {
        load8(dy + (long long)rn * d + c, g2);
        load8(x + (long long)rn * d + c, xv2);
      }
      const float mu = mean[rb], rs = rstd[rb];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ag[j] = fmaf(g[j], (xv[j] - mu) * rs, ag[j]);
        ab[j] += g[j];
      }
      if (two) {
        const float mu2 = mean[rn], rs2 = rstd[rn];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          ag[j] = fmaf(g2[j], (xv2[j] - mu2) * rs2, ag[j]);
          ab[j] += g2[j];
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[rl][0][cg * 8 + j] = ag[j];
    red[rl][1][cg * 8 + j] = ab[j];
  }
  __syncthreads();
  const int t = threadIdx.x;
  const int col = blockIdx.x * 256 + t;
  const int nch = gridDim.y;
  if (col < d) {
    float sg = 0.f, sb = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      sg += red[w][0][t];
      sb += red[w][1][t];
    }
    part[((long long)blockIdx.y * 2 + 0) * d + col] = sg;
    part[((long long)blockIdx.y * 2 + 1) * d + col] = sb;
  }
  if (last_cta_of_column(counters + blockIdx.x, nch) && col < d) {
    // fixed-order fold with 4 + 4 independent accumulators (one L2 round trip per 4 partials)
    float sg[4] = {0.f, 0.f, 0.f, 0.f}, sb[4] = {0.f, 0.f, 0.f, 0.f};
    int p = 0;
    for (; p + 3 < nch; p += 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        sg[k] += part[((long long)(p + k) * 2 + 0) * d + col];
        sb[k] += part[((long long)(p + k) * 2 + 1) * d + col];
      }
    }
    for (; p < nch; ++p) {
      sg[p & 3] += part[((long long)p * 2 + 0) * d + col];
      sb[p & 3] += part[((long long)p * 2 + 1) * d + col];
    }
    dgamma[col] = (sg[0] + sg[1]) + (sg[2] + sg[3]);
    dbeta[col] = (sb[0] + sb[1]) + (sb[2] + sb[3]);
  }
}

// ---------------------------------------------------------------------------------------------
// dU = dY * act'(U)  (optional) and column sums of the result (bias gradient), same last-CTA scheme.
// block = 32 column-groups (8 cols each) x 8 row lanes; grid = (ceil(N/256), row_chunks)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
act_bwd_colsum_kernel(const bf16* __restrict__ dY, long long ldy, const bf16* __restrict__ U, long long ldu, int act,
                      bf16* __restrict__ dU, long long lddu, float* __restrict__ part,
                      unsigned int* __restrict__ counters, float* __restrict__ dbias, int accumulate, int M, int N,
                      int rows_per_chunk, DropoutParams drop) {
  __shared__ float red[8][256 + 8];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + cg * 8;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(M, r0 + rows_per_chunk);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (c < N) {
    // 4 rows per trip with every load issued before the first use: a thread otherwise has ONE 16-byte load (two with
    // U) in flight and the kernel sits at 30-45 % of the HBM rate (profiles/r01q_launches.csv: 120 us for 403 MB)
    constexpr int UNR = 4;
    for (int rb = r0 + rl; rb < r1; rb += 8 * UNR) {
      uint4 gv[UNR], uv[UNR];
#pragma unroll
      for (int k = 0; k < UNR; ++k) {
        const int r = rb + 8 * k;
        gv[k] = make_uint4(0, 0, 0, 0);
        uv[k] = make_uint4(0, 0, 0, 0);
        if (r < r1) {
          gv[k] = *reinterpret_cast<const uint4*>(dY + (long long)r * ldy + c);
          if (act != MIC_ACT_NONE) uv[k] = *reinterpret_cast<const uint4*>(U + (long long)r * ldu + c);
        }
      }
#pragma unroll
      for (int k = 0; k < UNR; ++k) {
        const int r = rb + 8 * k;
        if (r >= r1) continue;
        float g[8];
        {
          float2 f;
          f = unpack_bf16(gv[k].x); g[0] = f.x; g[1] = f.y;
          f = unpack_bf16(gv[k].y); g[2] = f.x; g[3] = f.y;
          f = unpack_bf16(gv[k].z); g[4] = f.x; g[5] = f.y;
          f = unpack_bf16(gv[k].w); g[6] = f.x; g[7] = f.y;
        }
        if (drop.seed_ptr) {          // backward of x + dropout(z): dz = dy * mask / (1-p), same mask as forward
          const uint32_t seed = *drop.seed_ptr + drop.site;
          const uint32_t base = ((uint32_t)r * (uint32_t)N + (uint32_t)c) >> 1;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            bool k0, k1;
            drop_keep2(base + j, seed, drop.thr16, &k0, &k1);
            g[2 * j] = k0 ? bf16_round(g[2 * j] * drop.scale) : 0.f;
            g[2 * j + 1] = k1 ? bf16_round(g[2 * j + 1] * drop.scale) : 0.f;
          }
        }
        if (act != MIC_ACT_NONE) {
          float u[8];
          float2 f;
          f = unpack_bf16(uv[k].x); u[0] = f.x; u[1] = f.y;
          f = unpack_bf16(uv[k].y); u[2] = f.x; u[3] = f.y;
          f = unpack_bf16(uv[k].z); u[4] = f.x; u[5] = f.y;
          f = unpack_bf16(uv[k].w); u[6] = f.x; u[7] = f.y;
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] = bf16_round(g[j] * act_bwd(u[j], act));
        }
        if (dU) store8(dU + (long long)r * lddu + c, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += g[j];
      }
    }
  }
  if (!dbias) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][cg * 8 + j] = acc[j];
  __syncthreads();
  const int t = threadIdx.x;
  const int col = blockIdx.x * 256 + t;
  const int nch = gridDim.y;
  if (col < N) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][t];
    part[(long long)blockIdx.y * N + col] = s;
  }
  if (last_cta_of_column(counters + blockIdx.x, nch) && col < N) {
    // fixed-order fold of the per-chunk partials with 8 independent accumulators (one L2 round trip per 8 partials)
    float sacc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) sacc[k] = 0.f;
    int p = 0;
    for (; p + 7 < nch; p += 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) sacc[k] += part[(long long)(p + k) * N + col];
    }
    for (; p < nch; ++p) sacc[p & 7] += part[(long long)p * N + col];
    const float tot = ((sacc[0] + sacc[1]) + (sacc[2] + sacc[3])) + ((sacc[4] + sacc[5]) + (sacc[6] + sacc[7]));
    dbias[col] = accumulate ? dbias[col] + tot : tot;
  }
}

// ---------------------------------------------------------------------------------------------
// CE backward without recomputing the lm_head: the forward kernel left the bf16 logits z in `zdz`
// [M, ld]; this pass rewrites them IN PLACE as dlogits = (softmax(z) - soft_labels) * row_w and, in the
// same sweep, accumulates the column sums (= gradient of final_logits_bias).  Columns >= V become 0.
// grid = (ld/256, row chunks); block 256 = 32 column groups (8 cols) x 8 row lanes.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ce_softmax_bwd_kernel(bf16* __restrict__ zdz, long long ld, const int* __restrict__ labels,
                      const float* __restrict__ lse, const float* __restrict__ row_w, float conf, float low, int M,
                      int V, float* __restrict__ part, unsigned int* __restrict__ counters,
                      float* __restrict__ dbias, int rows_per_chunk) {
  __shared__ float red[8][256 + 8];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + cg * 8;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(M, r0 + rows_per_chunk);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int r = r0 + rl; r < r1; r += 8) {
    bf16* p = zdz + (long long)r * ld + c;
    float z[8];
    load8(p, z);
    const float w = row_w[r];
    const float l2 = lse[r] * 1.4426950408889634f;
    const int rel = labels[r] - c;
    const float lw = low * w, fix = (low - conf) * w;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float d = fmaf(exp2f(fmaf(z[j], 1.4426950408889634f, -l2)), w, -lw);
      d += (rel == j) ? fix : 0.f;
      d = (c + j < V) ? d : 0.f;
      z[j] = bf16_round(d);
      acc[j] += z[j];
    }
    store8(p, z);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][cg * 8 + j] = acc[j];
  __syncthreads();
  const int t = threadIdx.x;
  const int col = blockIdx.x * 256 + t;
  const int nch = gridDim.y;
  {
    float sum = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) sum += red[w8][t];
    part[(long long)blockIdx.y * ld + col] = sum;
  }
  if (last_cta_of_column(counters + blockIdx.x, nch) && col < V) {
    float s0 = 0.f, s1 = 0.f;
    int pch = 0;
    for (; pch + 1 < nch; pch += 2) {
      s0 += part[(long long)pch * ld + col];
      s1 += part[(long long)(pch + 1) * ld + col];
    }
    if (pch < nch) s0 += part[(long long)pch * ld + col];
    dbias[col] = s0 + s1;
  }
}

// ---------------------------------------------------------------------------------------------
// Decoder embedding: shared[id]*scale + positions[pos+offset] -> emb (bf16) -> LayerNorm -> y
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_ln_fwd_kernel(const int* __restrict__ ids, const int* __restrict__ pos_ids, int pos_mod, int pos_offset,
                    const bf16* __restrict__ table, const bf16* __restrict__ pos_table, float scale,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float eps, bf16* __restrict__ emb,
                    bf16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int M, int d,
                    DropoutParams drop) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const long long id = ids[row];
  const long long pos = (pos_ids ? pos_ids[row] : (row % pos_mod)) + pos_offset;
  float v[LN_MAX_ITERS][8];
#pragma unroll
  for (int it = 0; it < LN_MAX_ITERS; ++it) {
    const int c = lane * 8 + it * 256;
    if (c < d) {
      float e[8], p[8];
      load8(table + id * d + c, e);
      load8(pos_table + pos * d + c, p);
      // Flax bf16 mode rounds (embedding*scale) and the sum to bf16; we keep one rounding at the store
#pragma unroll
      for (int j = 0; j < 8; ++j) v[it][j] = bf16_round(e[j] * scale + p[j]);
      if (emb) store8(emb + (long long)row * d + c, v[it]);
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
      if (drop.seed_ptr) {
        const uint32_t seed = *drop.seed_ptr + drop.site;
        const uint32_t base = ((uint32_t)row * (uint32_t)d + (uint32_t)c) >> 1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          bool k0, k1;
          drop_keep2(base + j, seed, drop.thr16, &k0, &k1);
          o[2 * j] = k0 ? o[2 * j] * drop.scale : 0.f;
          o[2 * j + 1] = k1 ? o[2 * j + 1] * drop.scale : 0.f;
        }
      }
      store8(y + (long long)row * d + c, o);
    }
  }
  if (lane == 0 && mean_out) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
}

// scatter-add of d_emb*scale into the (tied) embedding gradient, fp32 atomics.  One warp per row, EXCEPT
// rows whose id is `hot_id` (the pad token: ~40 % of decoder_input_ids, all hitting one table row): those
// are summed by hot_sum_kernel below and added with a single atomic per column.
__global__ void __launch_bounds__(256)
embed_scatter_bwd_kernel(const int* __restrict__ ids, const bf16* __restrict__ d_emb, float scale,
                         float* __restrict__ d_table, int M, int d, int hot_id) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const long long id = ids[row];
  if (id == hot_id) return;
  for (int c = lane * 8; c < d; c += 256) {
    float g[8];
    load8(d_emb + (long long)row * d + c, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(d_table + id * d + c + j, g[j] * scale);
  }
}

// grid = (ceil(d/256), chunks): column sums of the rows with id == hot_id, one atomic per (chunk, column)
__global__ void __launch_bounds__(256)
hot_sum_kernel(const int* __restrict__ ids, const bf16* __restrict__ d_emb, float scale, float* __restrict__ d_table,
               int M, int d, int hot_id, int rows_per_chunk) {
  __shared__ float red[8][256 + 8];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + cg * 8;
  const int r0 = blockIdx.y * rows_per_chunk, r1 = min(M, r0 + rows_per_chunk);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (c < d) {
    for (int r = r0 + rl; r < r1; r += 8) {
      if (ids[r] != hot_id) continue;
      float g[8];
      load8(d_emb + (long long)r * d + c, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += g[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][cg * 8 + j] = acc[j];
  __syncthreads();
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col < d) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w][threadIdx.x];
    if (sum != 0.f) atomicAdd(d_table + (long long)hot_id * d + col, sum * scale);
  }
}

// out[t][c] = sum_b x[(b*T + t)][c]  (position-embedding gradient, class/position grads of the ViT)
// grid = (T, ceil(d/256)); block 256 = 32 col groups... here 1 thread per column for simplicity
__global__ void __launch_bounds__(256)
batch_sum_kernel(const bf16* __restrict__ x, int B, int T, int d, float* __restrict__ out, long long out_ld) {
  const int t = blockIdx.x;
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c >= d) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += __bfloat162float(x[((long long)b * T + t) * d + c]);
  out[(long long)t * out_ld + c] = s;
}

// ---------------------------------------------------------------------------------------------
// Vision front end
// ---------------------------------------------------------------------------------------------
// pixels fp32 (NHWC or NCHW) -> bf16 patch matrix [B*g*g, p*p*3], k ordered (kh, kw, c) = HWIO kernel
// rows.  trunc_int reproduces encode()'s int32 cast (modeling_clip_vision_mbart.py:330).
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ px, bf16* __restrict__ out, int B, int img, int p, int nchw, int trunc_int) {
  const int g = img / p;
  const int K = p * p * 3;
  const long long total = (long long)B * g * g * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const long long m = i / K;
    const int c = k % 3, kw = (k / 3) % p, kh = k / (3 * p);
    const int pw = (int)(m % g), ph = (int)((m / g) % g);
    const int b = (int)(m / (g * g));
    const int yy = ph * p + kh, xx = pw * p + kw;
    float v = nchw ? px[(((long long)b * 3 + c) * img + yy) * img + xx]
                   : px[(((long long)b * img + yy) * img + xx) * 3 + c];
    if (trunc_int) v = truncf(v);
    out[i] = __float2bfloat16_rn(v);
  }
}

// Input hand-off: the data pipeline's uint8 image (after Resize / CenterCrop, main.py:171-172) goes
// straight to the GPU — a quarter of the fp32 bytes over PCIe — and `ConvertImageDtype(torch.float)` (x / 255) and
// `Normalize(mean, std)` (main.py:173-174) happen here, fused with the patch gather and the bf16 conversion.
struct PixelNorm {
  float mean[3], inv_std[3];
};
__global__ void __launch_bounds__(256)
patchify_u8_kernel(const uint8_t* __restrict__ px, bf16* __restrict__ out, int B, int img, int p, int nchw,
                   int trunc_int, const PixelNorm nrm) {
  const int g = img / p;
  const int K = p * p * 3;
  const long long total = (long long)B * g * g * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const long long m = i / K;
    const int c = k % 3, kw = (k / 3) % p, kh = k / (3 * p);
    const int pw = (int)(m % g), ph = (int)((m / g) % g);
    const int b = (int)(m / (g * g));
    const int yy = ph * p + kh, xx = pw * p + kw;
    const uint8_t u = nchw ? px[(((long long)b * 3 + c) * img + yy) * img + xx]
                           : px[(((long long)b * img + yy) * img + xx) * 3 + c];
    // torchvision: convert_image_dtype = x / 255 (fp32 division), normalize = (x - mean) / std (fp32 division)
    float v = (__fdiv_rn((float)u, 255.0f) - nrm.mean[c]) * nrm.inv_std[c];
    if (trunc_int) v = truncf(v);
    out[i] = __float2bfloat16_rn(v);
  }
}

// tokens: [cls ; patch_out(+bias)] + pos -> emb (bf16) -> optional LayerNorm -> y.  One warp per token row.
__global__ void __launch_bounds__(256)
vit_embed_ln_fwd_kernel(const bf16* __restrict__ patch_out, const float* __restrict__ patch_bias,
                        const bf16* __restrict__ cls, const bf16* __restrict__ pos, const float* __restrict__ gamma,
                        const float* __restrict__ beta, float eps, int use_ln, bf16* __restrict__ emb,
                        bf16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int B, int S,
                        int d) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B * S) return;
  const int b = row / S, t = row % S;
  float v[LN_MAX_ITERS][8];
#pragma unroll
  for (int it = 0; it < LN_MAX_ITERS; ++it) {
    const int c = lane * 8 + it * 256;
    if (c < d) {
      float e[8], p[8];
      if (t == 0) {
        load8(cls + c, e);
      } else {
        load8(patch_out + ((long long)b * (S - 1) + (t - 1)) * d + c, e);
        if (patch_bias) {
          float pb[8];
          load8f(patch_bias + c, pb);
#pragma unroll
          for (int j = 0; j < 8; ++j) e[j] = bf16_round(e[j] + pb[j]);
        }
      }
      load8(pos + (long long)t * d + c, p);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[it][j] = bf16_round(e[j] + p[j]);
      if (emb) store8(emb + (long long)row * d + c, v[it]);
    }
  }
  if (!use_ln) {
#pragma unroll
    for (int it = 0; it < LN_MAX_ITERS; ++it) {
      const int c = lane * 8 + it * 256;
      if (c < d) store8(y + (long long)row * d + c, v[it]);
    }
    return;
  }
  float mean, rstd;
  ln_stats(v, d, lane, &mean, &rstd, eps);
#pragma unroll
  for (int it = 0; it < LN_MAX_ITERS; ++it) {
    const int c = lane * 8 + it * 256;
    if (c < d) {
      float g[8], bb[8], o[8];
      load8f(gamma + c, g);
      load8f(beta + c, bb);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[it][j] - mean) * rstd * g[j] + bb[j];
      store8(y + (long long)row * d + c, o);
    }
  }
  if (lane == 0 && mean_out) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
}

// d_emb [B,S,d] -> compact d_patch_out [B*(S-1), d] (drop the class token row)
__global__ void __launch_bounds__(256)
drop_cls_rows_kernel(const bf16* __restrict__ d_emb, bf16* __restrict__ out, int B, int S, int d) {
  const long long n8 = (long long)B * (S - 1) * d / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 8;
    const long long r = e / d;
    const int c = (int)(e % d);
    const long long b = r / (S - 1), t = r % (S - 1) + 1;
    *reinterpret_cast<uint4*>(out + e) = *reinterpret_cast<const uint4*>(d_emb + ((b * S + t) * d + c));
  }
}

// ---------------------------------------------------------------------------------------------
// CE finalisation: combine the per-slab partials -> lse, per-row loss, weights, scalar loss
// one warp per row for the combine; then a single-block reduction for the scalar
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ce_rows_kernel(const float* __restrict__ pmax, const float* __restrict__ psum, const float* __restrict__ psumz,
               const float* __restrict__ zlabel, int nparts, int M, int V, float eps_ls, float* __restrict__ lse_out,
               float* __restrict__ row_loss) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  float mx = -INFINITY;
  for (int p = lane; p < nparts; p += 32) mx = fmaxf(mx, pmax[(long long)p * M + row]);
  mx = warp_max(mx);
  float s = 0.f, sz = 0.f;
  for (int p = lane; p < nparts; p += 32) {
    const float m = pmax[(long long)p * M + row];
    if (m > -INFINITY) s += psum[(long long)p * M + row] * __expf(m - mx);
    if (psumz) sz += psumz[(long long)p * M + row];
  }
  s = warp_sum(s);
  sz = warp_sum(sz);
  if (lane == 0) {
    const float lse = mx + logf(s);
    lse_out[row] = lse;
    if (row_loss) {
      // main.py:666-675 closed form: lse - conf*z_y - low*(sum z - z_y) - const
      const float conf = 1.0f - eps_ls;
      const float low = eps_ls / (float)(V - 1);
      float cst = 0.f;
      if (eps_ls > 0.f) cst = -(conf * logf(conf) + (float)(V - 1) * low * logf(low + 1e-20f));
      const float zy = zlabel[row];
      row_loss[row] = lse - conf * zy - low * (sz - zy) - cst;
    }
  }
}

__global__ void __launch_bounds__(1024)
ce_reduce_kernel(const float* __restrict__ row_loss, const int* __restrict__ mask, int M, float* __restrict__ row_w,
                 float* __restrict__ out) {
  __shared__ float sl[32], sc[32];
  float l = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const float m = mask ? (float)mask[i] : 1.f;
    l += row_loss[i] * m;
    c += m;
  }
  l = warp_sum(l);
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) {
    sl[threadIdx.x >> 5] = l;
    sc[threadIdx.x >> 5] = c;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    l = sl[threadIdx.x];
    c = sc[threadIdx.x];
    l = warp_sum(l);
    c = warp_sum(c);
    if (threadIdx.x == 0) {
      sl[0] = l;
      sc[0] = c;
      out[0] = l / c;
      out[1] = c;
    }
  }
  __syncthreads();
  const float inv = 1.0f / sc[0];
  if (row_w)
    for (int i = threadIdx.x; i < M; i += blockDim.x) row_w[i] = (mask ? (float)mask[i] : 1.f) * inv;
}

// ---------------------------------------------------------------------------------------------
// AdamW (optax 0.0.9 chain: scale_by_adam -> add_decayed_weights -> scale(-lr)), flat fp32 state,
// bf16 shadow refresh.  The step's scalars are kernel ARGUMENTS (by value): the update runs outside the captured
// forward/backward graph, and a pinned staging buffer would be overwritten by a host that runs steps ahead.
//   c1 = 1/(1-b1^t), c2 = 1/(1-b2^t), gs = gradient scale (1/world)
// ---------------------------------------------------------------------------------------------
struct AdamHyper {
  float lr, b1, b2, eps, wd, c1, c2, gs;
};
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g,
             bf16* __restrict__ shadow, const AdamHyper hp, long long n) {
  const float lr = hp.lr, b1 = hp.b1, b2 = hp.b2, eps = hp.eps, wd = hp.wd, c1 = hp.c1, c2 = hp.c2, gs = hp.gs;
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float* pa = reinterpret_cast<float*>(&pp);
    float* ma = reinterpret_cast<float*>(&mm);
    float* va = reinterpret_cast<float*>(&vv);
    const float* ga = reinterpret_cast<const float*>(&gg);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = ga[j] * gs;
      ma[j] = b1 * ma[j] + (1.f - b1) * gr;
      va[j] = b2 * va[j] + (1.f - b2) * gr * gr;
      const float upd = (ma[j] * c1) / (sqrtf(va[j] * c2) + eps) + wd * pa[j];
      pa[j] -= lr * upd;
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (shadow)
      reinterpret_cast<uint2*>(shadow)[i] = make_uint2(pack_bf16(pa[0], pa[1]), pack_bf16(pa[2], pa[3]));
  }
  // tail (n not a multiple of 4)
  const long long tail0 = n4 << 2;
  const long long i = tail0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float gr = g[i] * gs;
    const float mj = b1 * m[i] + (1.f - b1) * gr;
    const float vj = b2 * v[i] + (1.f - b2) * gr * gr;
    const float upd = (mj * c1) / (sqrtf(vj * c2) + eps) + wd * p[i];
    const float pj = p[i] - lr * upd;
    p[i] = pj;
    m[i] = mj;
    v[i] = vj;
    if (shadow) shadow[i] = __float2bfloat16_rn(pj);
  }
}

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}

inline int grid_for(long long work_items, int per_block, int cap_mult = 8) {
  long long b = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)mic_num_sms() * cap_mult;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" int mic_layernorm_fwd(void* stream, const void* x, const float* gamma, const float* beta, float eps,
                                 void* y, float* mean, float* rstd, int M, int d) {
  MIC_CHECK_ARG(d % 8 == 0 && d <= 1024 && M > 0, "layernorm: d=%d must be a multiple of 8 and <= 1024", d);
  MIC_CHECK_CUDA(mic_launch(layernorm_fwd_kernel, dim3((M + 7) / 8), dim3(256), 0, STREAM, (const bf16*)x, gamma, beta, eps, (bf16*)y, mean, rstd, M, d));
  return MIC_OK;
}

extern "C" int mic_residual_ln_fwd(void* stream, float* acc, const float* bias, void* x, const float* gamma,
                                   const float* beta, float eps, void* y, int M, int d) {
  MIC_CHECK_ARG(d % 8 == 0 && d <= 1024 && M > 0, "residual_ln: d=%d must be a multiple of 8 and <= 1024", d);
  MIC_CHECK_CUDA(mic_launch(residual_ln_fwd_kernel, dim3((M + 7) / 8), dim3(256), 0, STREAM, acc, bias, (bf16*)x, gamma, beta, eps, (bf16*)y, M, d));
  return MIC_OK;
}

static int pick_chunks(int M, int col_blocks) {
  // 2 CTAs of 256 threads per SM.  (6 per SM was measured in round 2: the wide passes gained 10-40 %, the many narrow
  // ones lost as much to the longer partial fold - act_bwd_colsum 5.35 vs 5.45 ms, ln_param_grad 1.94 vs 1.57 ms per
  // training step - so the round-1 choice stays.)
  int chunks = (2 * mic_num_sms() + col_blocks - 1) / col_blocks;
  const int by_rows = (M + 511) / 512;                    // at most 512 rows (64 iterations per thread) per chunk
  if (chunks < by_rows) chunks = by_rows;
  const int max_chunks = (M + 31) / 32;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks > 256) chunks = 256;
  return chunks < 1 ? 1 : chunks;
}

// workspace sizes in floats (partials).  `counters`: >= 1024 uint32, zero-initialised ONCE by the caller;
// every kernel leaves them zero again.
extern "C" long long mic_layernorm_bwd_workspace_floats(int M, int d) {
  return 2ll * pick_chunks(M, (d + 255) / 256) * d;
}
extern "C" long long mic_colsum_workspace_floats(int M, int N) { return (long long)pick_chunks(M, (N + 255) / 256) * N; }

extern "C" int mic_layernorm_bwd(void* stream, const void* dy, const void* x, const float* gamma, const float* mean,
                                 const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta,
                                 float* workspace, unsigned int* counters, int M, int d) {
  MIC_CHECK_ARG(d % 8 == 0 && d <= 1024 && M > 0, "layernorm_bwd: d=%d must be a multiple of 8 and <= 1024", d);
  // (a fused single-pass variant - dx + per-CTA partial column sums, then a fold kernel - was built and measured in
  //  round 2: 71 us + 9 us per LayerNorm against 29 + 25 us for these two kernels; its 173 registers per thread leave one
  //  CTA per SM and the pass becomes latency-bound, so the two-kernel form stays)
  layernorm_bwd_dx_kernel<<<(M + 7) / 8, 256, 0, STREAM>>>((const bf16*)dy, (const bf16*)x, gamma, mean, rstd,
                                                           (const bf16*)dres, (bf16*)dx, M, d);
  MIC_CHECK_LAUNCH();
  if (dgamma) {
    const int cb = (d + 255) / 256;
    const int chunks = pick_chunks(M, cb);
    dim3 grid(cb, chunks);
    ln_param_grad_kernel<<<grid, 256, 0, STREAM>>>((const bf16*)dy, (const bf16*)x, mean, rstd, workspace, counters,
                                                   dgamma, dbeta, M, d, (M + chunks - 1) / chunks);
    MIC_CHECK_LAUNCH();
  }
  return MIC_OK;
}

extern "C" int mic_act_bwd_colsum(void* stream, const void* dY, long long ldy, const void* U, long long ldu, int act,
                                  void* dU, long long lddu, float* dbias, int accumulate, float* workspace,
                                  unsigned int* counters, int M, int N, const unsigned int* drop_seed,
                                  unsigned int drop_site, float drop_p) {
  MIC_CHECK_ARG(N % 8 == 0 && ldy % 8 == 0, "act_bwd_colsum: N and ld must be multiples of 8");
  MIC_CHECK_ARG(act == MIC_ACT_NONE || (U && dU), "act_bwd_colsum: activation backward needs U and dU");
  const int cb = (N + 255) / 256;
  MIC_CHECK_ARG(cb <= 1024, "act_bwd_colsum: N=%d too wide for the counter array", N);
  const int chunks = pick_chunks(M, cb);
  dim3 grid(cb, chunks);
  DropoutParams drop;
  drop.seed_ptr = (drop_seed && drop_p > 0.f) ? drop_seed : nullptr;
  drop.site = drop_site;
  drop.thr16 = (uint32_t)(drop_p * 65536.0f + 0.5f);
  drop.scale = 65536.0f / (65536.0f - (float)drop.thr16);
  MIC_CHECK_ARG(!drop.seed_ptr || dU, "act_bwd_colsum: dropout backward needs a dU output");
  act_bwd_colsum_kernel<<<grid, 256, 0, STREAM>>>((const bf16*)dY, ldy, (const bf16*)U, ldu, act, (bf16*)dU, lddu,
                                                  workspace, counters, dbias, accumulate, M, N,
                                                  (M + chunks - 1) / chunks, drop);
  MIC_CHECK_LAUNCH();
  return MIC_OK;
}

static int ce_bwd_chunks(int M) {
  int ch = (M + 511) / 512;
  return ch < 1 ? 1 : ch;
}
extern "C" long long mic_ce_softmax_bwd_workspace_floats(int M, long long ld) { return (long long)ce_bwd_chunks(M) * ld; }

extern "C" int mic_ce_softmax_bwd(void* stream, void* logits_inout, long long ld, const int* labels, const float* lse,
                                  const float* row_w, float conf, float low, int M, int V, float* dbias,
                                  float* workspace, unsigned int* counters) {
  MIC_CHECK_ARG(ld % 256 == 0 && ld >= V && ld / 256 <= 1024, "ce_softmax_bwd: ld=%lld must be a multiple of 256, >= V", ld);
  const int chunks = ce_bwd_chunks(M);
  dim3 grid((unsigned)(ld / 256), chunks);
  ce_softmax_bwd_kernel<<<grid, 256, 0, STREAM>>>((bf16*)logits_inout, ld, labels, lse, row_w, conf, low, M, V,
                                                  workspace, counters, dbias, (M + chunks - 1) / chunks);
  MIC_CHECK_LAUNCH();
  return MIC_OK;
}

extern "C" int mic_embed_ln_fwd(void* stream, const int* ids, const int* pos_ids, int pos_mod, int pos_offset,
                                const void* table, const void* pos_table, float scale, const float* gamma,
                                const float* beta, float eps, void* emb, void* y, float* mean, float* rstd, int M,
                                int d, const unsigned int* drop_seed, unsigned int drop_site, float drop_p) {
  MIC_CHECK_ARG(d % 8 == 0 && d <= 1024, "embed: d=%d must be a multiple of 8 and <= 1024", d);
  DropoutParams drop;
  drop.seed_ptr = (drop_seed && drop_p > 0.f) ? drop_seed : nullptr;
  drop.site = drop_site;
  drop.thr16 = (uint32_t)(drop_p * 65536.0f + 0.5f);
  drop.scale = 65536.0f / (65536.0f - (float)drop.thr16);
  MIC_CHECK_CUDA(mic_launch(embed_ln_fwd_kernel, dim3((M + 7) / 8), dim3(256), 0, STREAM, ids, pos_ids, pos_mod, pos_offset, (const bf16*)table,
                                                       (const bf16*)pos_table, scale, gamma, beta, eps, (bf16*)emb,
                                                       (bf16*)y, mean, rstd, M, d, drop));
  return MIC_OK;
}

extern "C" int mic_embed_bwd(void* stream, const int* ids, const void* d_emb, float scale, float* d_table,
                             float* d_pos_rows, int B, int T, int d, int hot_id) {
  const int M = B * T;
  MIC_CHECK_ARG(d % 8 == 0, "embed_bwd: d must be a multiple of 8");
  embed_scatter_bwd_kernel<<<(M + 7) / 8, 256, 0, STREAM>>>(ids, (const bf16*)d_emb, scale, d_table, M, d, hot_id);
  MIC_CHECK_LAUNCH();
  if (hot_id >= 0) {
    const int chunks = (M + 255) / 256;
    dim3 grid((d + 255) / 256, chunks);
    hot_sum_kernel<<<grid, 256, 0, STREAM>>>(ids, (const bf16*)d_emb, scale, d_table, M, d, hot_id, 256);
    MIC_CHECK_LAUNCH();
  }
  if (d_pos_rows) {
    dim3 grid(T, (d + 255) / 256);
    batch_sum_kernel<<<grid, 256, 0, STREAM>>>((const bf16*)d_emb, B, T, d, d_pos_rows, d);
    MIC_CHECK_LAUNCH();
  }
  return MIC_OK;
}

extern "C" int mic_batch_sum(void* stream, const void* x, int B, int T, int d, float* out, long long out_ld) {
  dim3 grid(T, (d + 255) / 256);
  batch_sum_kernel<<<grid, 256, 0, STREAM>>>((const bf16*)x, B, T, d, out, out_ld);
  MIC_CHECK_LAUNCH();
  return MIC_OK;
}

extern "C" int mic_patchify(void* stream, const float* pixels, void* out, int B, int image_size, int patch,
                            int channel_first, int trunc_int) {
  MIC_CHECK_ARG(image_size % patch == 0, "image size %d not divisible by patch %d", image_size, patch);
  const long long total = (long long)B * image_size * image_size * 3;
  patchify_kernel<<<grid_for(total, 256 * 4), 256, 0, STREAM>>>(pixels, (bf16*)out, B, image_size, patch,
                                                                channel_first, trunc_int);
  MIC_CHECK_LAUNCH();
  return MIC_OK;
}

extern "C" int mic_patchify_u8(void* stream, const unsigned char* pixels, void* out, int B, int image_size, int patch,
                               int channel_first, int trunc_int, const float* mean3, const float* std3) {
  MIC_CHECK_ARG(image_size % patch == 0, "image size %d not divisible by patch %d", image_size, patch);
  MIC_CHECK_ARG(pixels && out && mean3 && std3, "patchify_u8: null argument");
  PixelNorm nrm;
  for (int c = 0; c < 3; ++c) {
    MIC_CHECK_ARG(std3[c] > 0.f, "patchify_u8: std[%d] must be positive", c);
    nrm.mean[c] = mean3[c];
    nrm.inv_std[c] = 1.0f / std3[c];
  }
  const long long total = (long long)B * image_size * image_size * 3;
  patchify_u8_kernel<<<grid_for(total, 256 * 4), 256, 0, STREAM>>>(pixels, (bf16*)out, B, image_size, patch,
                                                                   channel_first, trunc_int, nrm);
  MIC_CHECK_LAUNCH();
  return MIC_OK;
}

extern "C" int mic_vit_embed_ln_fwd(void* stream, const void* patch_out, const float* patch_bias, const void* cls,
                                    const void* pos, const float* gamma, const float* beta, float eps, int use_ln,
                                    void* emb, void* y, float* mean, float* rstd, int B, int S, int d) {
  MIC_CHECK_ARG(d % 8 == 0 && d <= 1024, "vit_embed: d=%d must be a multiple of 8 and <= 1024", d);
  const int M = B * S;
  vit_embed_ln_fwd_kernel<<<(M + 7) / 8, 256, 0, STREAM>>>((const bf16*)patch_out, patch_bias, (const bf16*)cls,
                                                           (const bf16*)pos, gamma, beta, eps, use_ln, (bf16*)emb,
                                                           (bf16*)y, mean, rstd, B, S, d);
  MIC_CHECK_LAUNCH();
  return MIC_OK;
}

extern "C" int mic_drop_cls_rows(void* stream, const void* d_emb, void* out, int B, int S, int d) {
  MIC_CHECK_ARG(d % 8 == 0, "drop_cls_rows: d must be a multiple of 8");
  drop_cls_rows_kernel<<<grid_for((long long)B * (S - 1) * d / 8, 256), 256, 0, STREAM>>>((const bf16*)d_emb,
                                                                                          (bf16*)out, B, S, d);
  MIC_CHECK_LAUNCH();
  return MIC_OK;
}

extern "C" int mic_ce_finalize(void* stream, const float* pmax, const float* psum, const float* psumz,
                               const float* zlabel, const int* mask, int num_partials, int M, int V,
                               float label_smoothing, float* lse, float* row_loss, float* row_w, float* out) {
  ce_rows_kernel<<<(M + 7) / 8, 256, 0, STREAM>>>(pmax, psum, psumz, zlabel, num_partials, M, V, label_smoothing, lse,
                                                  row_loss);
  MIC_CHECK_LAUNCH();
  if (row_loss && out) {
    ce_reduce_kernel<<<1, 1024, 0, STREAM>>>(row_loss, mask, M, row_w, out);
    MIC_CHECK_LAUNCH();
  }
  return MIC_OK;
}

extern "C" int mic_adamw(void* stream, float* p, float* m, float* v, const float* g, void* shadow_bf16, long long n,
                         float lr, float b1, float b2, float eps, float weight_decay, float bias_corr1,
                         float bias_corr2, float grad_scale) {
  const AdamHyper hyper = {lr, b1, b2, eps, weight_decay, bias_corr1, bias_corr2, grad_scale};
  MIC_CHECK_ARG(((uintptr_t)p & 15) == 0 && ((uintptr_t)m & 15) == 0 && ((uintptr_t)v & 15) == 0 &&
                    ((uintptr_t)g & 15) == 0,
                "adamw: state pointers must be 16-byte aligned");
  // grid-stride CTAs live for the whole kernel and 8 of them (2,048 threads) fill an SM: with a communication margin
  // set (data-parallel step: the next bucket's all-reduce runs concurrently) only 5 CTAs per SM are launched, so an
  // NCCL CTA always finds room
  int grid = grid_for(n / 4 + 1, 256, 16);
  if (g_mic_launch.sm_margin > 0 && grid > mic_num_sms() * 5) grid = mic_num_sms() * 5;
  adamw_kernel<<<grid, 256, 0, STREAM>>>(p, m, v, g, (bf16*)shadow_bf16, hyper, n);
  MIC_CHECK_LAUNCH();
  return MIC_OK;
}

extern "C" int mic_cast_f32_to_bf16(void* stream, const float* in, void* out, long long n) {
  cast_f32_bf16_kernel<<<grid_for(n, 256 * 4), 256, 0, STREAM>>>(in, (bf16*)out, n);
  MIC_CHECK_LAUNCH();
  return MIC_OK;
}
