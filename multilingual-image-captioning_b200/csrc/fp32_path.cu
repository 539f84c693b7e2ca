// fp32 verification path: forward + loss of the captioning model with fp32 storage and fp32 SIMT arithmetic
// (no tensor cores), for the parity bar of BASELINE configs[0] ("B = 8, fp32": logits within 1e-3 relative, loss
// within 1e-4 of the reference restatement).  Small shapes only (8 x 64 tokens): the kernels are written for
// clarity and correct rounding, not for throughput - the product path is the bf16 tcgen05 one.
//   mic_f32_gemm        D = act(A . B + bias) + residual      (flax.linen.Dense / tied lm_head)
//   mic_f32_layernorm   flax.linen.LayerNorm (var = E[x^2] - E[x]^2)
//   mic_f32_attention   softmax((q / sqrt(64)) k^T + mask) v per (batch, head)
//   mic_f32_embed / mic_f32_patchify / mic_f32_vit_embed / mic_f32_ce_rows
#include "common.cuh"

#include "../../include/mic_b200.h"

namespace {

__device__ __forceinline__ float act_exact(float x, int act) {
  if (act == MIC_ACT_GELU) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
  if (act == MIC_ACT_QUICK_GELU) return x / (1.0f + expf(-1.702f * x));
  return x;
}

// 64 x 64 output tile per CTA, 16 x 16 threads, 4 x 4 outputs per thread, K step 16.
// B is [K, N] (b_nk = 0, Flax kernel) or [N, K] (b_nk = 1, embedding table as lm_head).
__global__ void __launch_bounds__(256) f32_gemm_kernel(const float* __restrict__ A, long long lda,
                                                       const float* __restrict__ B, long long ldb, int b_nk, int M,
                                                       int N, int K, const float* __restrict__ bias, int act,
                                                       const float* __restrict__ residual, long long ldr,
                                                       float* __restrict__ D, long long ldd) {
  __shared__ float As[16][64 + 1];
  __shared__ float Bs[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, kk = i & 15;                 // A tile: 64 rows x 16 k (k contiguous in memory)
      const int gm = m0 + r, gk = k0 + kk;
      As[kk][r] = (gm < M && gk < K) ? A[(long long)gm * lda + gk] : 0.f;
    }
    if (b_nk) {
      for (int i = threadIdx.x; i < 64 * 16; i += 256) {
        const int c = i >> 4, kk = i & 15;               // B[N, K]: k contiguous
        const int gn = n0 + c, gk = k0 + kk;
        Bs[kk][c] = (gn < N && gk < K) ? B[(long long)gn * ldb + gk] : 0.f;
      }
    } else {
      for (int i = threadIdx.x; i < 64 * 16; i += 256) {
        const int kk = i >> 6, c = i & 63;               // B[K, N]: n contiguous
        const int gn = n0 + c, gk = k0 + kk;
        Bs[kk][c] = (gn < N && gk < K) ? B[(long long)gk * ldb + gn] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[gn];
      v = act_exact(v, act);
      if (residual) v += residual[(long long)gm * ldr + gn];
      D[(long long)gm * ldd + gn] = v;
    }
  }
}

// one warp per row
__global__ void __launch_bounds__(256) f32_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps,
                                                            float* __restrict__ y, int M, int d) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const float* xr = x + (long long)row * d;
  float s = 0.f, s2 = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float v = xr[c];
    s += v;
    s2 += v * v;
  }
  s = warp_sum(s);
  s2 = warp_sum(s2);
  const float mean = s / d;
  const float var = fmaxf(s2 / d - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  for (int c = lane; c < d; c += 32) y[(long long)row * d + c] = (xr[c] - mean) * rstd * gamma[c] + beta[c];
}

// one CTA (128 threads) per (batch, head); Tq, Tk <= 256; head_dim 64.  Each warp owns query rows q = w, w+4, ...
__global__ void __launch_bounds__(128) f32_attention_kernel(const float* __restrict__ Q, long long ldq,
                                                            const float* __restrict__ Kp, long long ldk,
                                                            const float* __restrict__ Vp, long long ldv,
                                                            float* __restrict__ O, long long ldo,
                                                            const int* __restrict__ key_mask, int causal, int H,
                                                            int Tq, int Tk, float scale) {
  __shared__ float p[4][256];
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int q = w; q < Tq; q += 4) {
    const float* qr = Q + ((long long)b * Tq + q) * ldq + h * 64;
    const float q0 = qr[lane] * scale, q1 = qr[lane + 32] * scale;
    float mx = -INFINITY;
    for (int j = 0; j < Tk; ++j) {
      const float* kr = Kp + ((long long)b * Tk + j) * ldk + h * 64;
      float sdot = warp_sum(q0 * kr[lane] + q1 * kr[lane + 32]);
      const bool ok = !(causal && j > q) && !(key_mask && key_mask[b * Tk + j] == 0);
      sdot = ok ? sdot : -INFINITY;
      if (lane == 0) p[w][j] = sdot;
      mx = fmaxf(mx, sdot);
    }
    __syncwarp();
    float sum = 0.f;
    for (int j = lane; j < Tk; j += 32) {
      const float e = (p[w][j] == -INFINITY) ? 0.f : expf(p[w][j] - mx);
      p[w][j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < Tk; ++j) {
      const float* vr = Vp + ((long long)b * Tk + j) * ldv + h * 64;
      const float pj = p[w][j];
      o0 = fmaf(pj, vr[lane], o0);
      o1 = fmaf(pj, vr[lane + 32], o1);
    }
    float* orow = O + ((long long)b * Tq + q) * ldo + h * 64;
    orow[lane] = o0 / sum;
    orow[lane + 32] = o1 / sum;
    __syncwarp();
  }
}

// out[m] = table[ids[m]] * scale + pos_table[(m % T) + pos_offset]
__global__ void __launch_bounds__(256) f32_embed_kernel(const int* __restrict__ ids, const float* __restrict__ table,
                                                        float scale, const float* __restrict__ pos_table,
                                                        int pos_offset, int T, float* __restrict__ out, int M, int d) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= (long long)M * d) return;
  const int m = (int)(i / d), c = (int)(i % d);
  out[i] = table[(long long)ids[m] * d + c] * scale + pos_table[(long long)((m % T) + pos_offset) * d + c];
}

// pixels NHWC (or NCHW) -> [B * np, patch*patch*3] with the kernel's (kh, kw, cin) order; optional int32 truncation
__global__ void __launch_bounds__(256) f32_patchify_kernel(const float* __restrict__ px, float* __restrict__ out, int B,
                                                           int image, int patch, int channel_first, int trunc_int) {
  const int g = image / patch, P = patch * patch * 3;
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= (long long)B * g * g * P) return;
  const int e = (int)(i % P);
  const long long pi = i / P;
  const int pxl = (int)(pi % (g * g)), b = (int)(pi / (g * g));
  const int ph = pxl / g, pw = pxl % g;
  const int kh = e / (patch * 3), kw = (e / 3) % patch, c = e % 3;
  const int y = ph * patch + kh, x = pw * patch + kw;
  float v = channel_first ? px[(((long long)b * 3 + c) * image + y) * image + x]
                          : px[(((long long)b * image + y) * image + x) * 3 + c];
  if (trunc_int) v = (float)(int)v;
  out[i] = v;
}

// out[b, 0] = cls + pos[0]; out[b, 1 + p] = patch_out[b, p] (+ patch_bias) + pos[1 + p]
__global__ void __launch_bounds__(256) f32_vit_embed_kernel(const float* __restrict__ patch_out,
                                                            const float* __restrict__ patch_bias,
                                                            const float* __restrict__ cls,
                                                            const float* __restrict__ pos, float* __restrict__ out,
                                                            int B, int S, int d) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= (long long)B * S * d) return;
  const int c = (int)(i % d);
  const int s = (int)((i / d) % S), b = (int)(i / ((long long)d * S));
  float v;
  if (s == 0) {
    v = cls[c];
  } else {
    v = patch_out[((long long)b * (S - 1) + (s - 1)) * d + c];
    if (patch_bias) v += patch_bias[c];
  }
  out[i] = v + pos[(long long)s * d + c];
}

// per row: lse = logsumexp(z) and the label-smoothed cross entropy of main.py:658-675
__global__ void __launch_bounds__(256) f32_ce_rows_kernel(const float* __restrict__ z, long long ld,
                                                          const int* __restrict__ labels, int M, int V, float eps,
                                                          float* __restrict__ row_loss, float* __restrict__ lse_out) {
  __shared__ float red[8];
  const int row = blockIdx.x;
  const float* zr = z + (long long)row * ld;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < V; j += 256) mx = fmaxf(mx, zr[j]);
  mx = warp_max(mx);
  if (lane == 0) red[w] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float s = 0.f, sz = 0.f;
  for (int j = threadIdx.x; j < V; j += 256) {
    s += expf(zr[j] - mx);
    sz += zr[j];
  }
  s = warp_sum(s);
  if (lane == 0) red[w] = s;
  __syncthreads();
  s = 0.f;
  for (int i = 0; i < 8; ++i) s += red[i];
  __syncthreads();
  sz = warp_sum(sz);
  if (lane == 0) red[w] = sz;
  __syncthreads();
  sz = 0.f;
  for (int i = 0; i < 8; ++i) sz += red[i];
  if (threadIdx.x == 0) {
    const float lse = mx + logf(s);
    // main.py:658-675 in closed form (sum of the soft labels is 1): lse - conf*z_y - low*(sum z - z_y) - const
    const float conf = 1.0f - eps, low = eps / (float)(V - 1);
    const float zl = zr[labels[row]];
    float cst = 0.f;
    if (eps > 0.f) cst = -(conf * logf(conf) + (float)(V - 1) * low * logf(low + 1e-20f));
    const float loss = lse - conf * zl - low * (sz - zl) - cst;
    row_loss[row] = loss;
    lse_out[row] = lse;
  }
}

}  // namespace

#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" int mic_f32_gemm(void* stream, const float* A, long long lda, const float* B, long long ldb, int b_is_nk,
                            int M, int N, int K, const float* bias, int act, const float* residual, long long ldr,
                            float* D, long long ldd) {
  MIC_CHECK_ARG(A && B && D && M > 0 && N > 0 && K > 0, "f32_gemm: bad argument");
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  f32_gemm_kernel<<<grid, 256, 0, STREAM>>>(A, lda, B, ldb, b_is_nk, M, N, K, bias, act, residual, ldr, D, ldd);
  MIC_CHECK_LAUNCH();
  return MIC_OK;
}
extern "C" int mic_f32_layernorm(void* stream, const float* x, const float* gamma, const float* beta, float eps,
                                 float* y, int M, int d) {
  f32_layernorm_kernel<<<(M + 7) / 8, 256, 0, STREAM>>>(x, gamma, beta, eps, y, M, d);
  MIC_CHECK_LAUNCH();
  return MIC_OK;
}
extern "C" int mic_f32_attention(void* stream, const float* Q, long long ldq, const float* K, long long ldk,
                                 const float* V, long long ldv, float* O, long long ldo, const int* key_mask,
                                 int causal, int B, int H, int Tq, int Tk, int head_dim, float scale) {
  MIC_CHECK_ARG(head_dim == 64 && Tk <= 256 && Tq >= 1 && Tk >= 1, "f32_attention: head_dim %d / Tk %d unsupported",
                head_dim, Tk);
  f32_attention_kernel<<<B * H, 128, 0, STREAM>>>(Q, ldq, K, ldk, V, ldv, O, ldo, key_mask, causal, H, Tq, Tk, scale);
  MIC_CHECK_LAUNCH();
  return MIC_OK;
}
extern "C" int mic_f32_embed(void* stream, const int* ids, const float* table, float scale, const float* pos_table,
                             int pos_offset, int T, float* out, int M, int d) {
  const long long n = (long long)M * d;
  f32_embed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, STREAM>>>(ids, table, scale, pos_table, pos_offset, T, out, M, d);
  MIC_CHECK_LAUNCH();
  return MIC_OK;
}
extern "C" int mic_f32_patchify(void* stream, const float* pixels, float* out, int B, int image_size, int patch,
                                int channel_first, int trunc_int) {
  const int g = image_size / patch;
  const long long n = (long long)B * g * g * patch * patch * 3;
  f32_patchify_kernel<<<(unsigned)((n + 255) / 256), 256, 0, STREAM>>>(pixels, out, B, image_size, patch, channel_first,
                                                                       trunc_int);
  MIC_CHECK_LAUNCH();
  return MIC_OK;
}
extern "C" int mic_f32_vit_embed(void* stream, const float* patch_out, const float* patch_bias, const float* cls,
                                 const float* pos, float* out, int B, int S, int d) {
  const long long n = (long long)B * S * d;
  f32_vit_embed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, STREAM>>>(patch_out, patch_bias, cls, pos, out, B, S, d);
  MIC_CHECK_LAUNCH();
  return MIC_OK;
}
extern "C" int mic_f32_ce_rows(void* stream, const float* logits, long long ld, const int* labels, int M, int V,
                               float label_smoothing, float* row_loss, float* lse) {
  f32_ce_rows_kernel<<<M, 256, 0, STREAM>>>(logits, ld, labels, M, V, label_smoothing, row_loss, lse);
  MIC_CHECK_LAUNCH();
  return MIC_OK;
}
