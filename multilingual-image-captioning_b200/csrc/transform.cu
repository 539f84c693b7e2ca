// Input pipeline hand-off: the reference's image Transform on the GPU.
//   main.py:165-179 / evaluation.py:33-50:  Resize([S], BICUBIC) -> CenterCrop(S) -> ConvertImageDtype -> Normalize
// This file is the first half (Resize + CenterCrop on the raw uint8 CHW image of torchvision.io.read_image, main.py:224-226);
// the second half (x/255, Normalize, NHWC gather into patches) is fused into mic_patchify_u8 (elementwise.cu).
//
// Arithmetic = torchvision's tensor path of the release the reference ran (0.10: no antialias for tensors):
// uint8 -> fp32, ATen upsample_bicubic2d(align_corners=False), clamp [0,255], round half to even -> uint8.
//   source x = scale * (dst + 0.5) - 0.5, scale = in / out (fp32);  taps floor(x)-1 .. floor(x)+2 clamped to the image;
//   Keys cubic convolution weights, A = -0.75;  x axis first, then the four rows;  multiply-adds contracted to FMAs
//   exactly where oracle/reference_transform.py contracts them (the -mfma build of ATen does the same).
// Only the S x S pixels that survive the centre crop are computed.  One launch handles a batch of images of different
// sizes (descriptor table); HBM-bound gather: every source byte is read about once from DRAM (neighbouring outputs share
// taps through L1/L2), 3 bytes written per output pixel.
#include "common.cuh"

#include "../../include/mic_b200.h"

namespace {

struct Taps {
  int idx[4];
  float w[4];
};

__device__ __forceinline__ float cubic1(float x) {   // ((A + 2) x - (A + 3)) x x + 1
  const float p = fmaf(1.25f, x, -2.25f);
  return fmaf(__fmul_rn(p, x), x, 1.0f);
}
__device__ __forceinline__ float cubic2(float x) {   // ((A x - 5A) x + 8A) x - 4A
  float p = fmaf(-0.75f, x, 3.75f);
  p = fmaf(p, x, -6.0f);
  return fmaf(p, x, 3.0f);
}
__device__ __forceinline__ Taps bicubic_taps(int in_size, int out_size, int dst) {
  Taps t;
  const float scale = __fdiv_rn((float)in_size, (float)out_size);
  const float real = fmaf(scale, (float)dst + 0.5f, -0.5f);
  const float fl = floorf(real);
  const float lam = __fsub_rn(real, fl);
  const int base = (int)fl;
#pragma unroll
  for (int k = 0; k < 4; ++k) t.idx[k] = min(max(base - 1 + k, 0), in_size - 1);
  const float lam2 = __fsub_rn(1.0f, lam);
  t.w[0] = cubic2(__fadd_rn(lam, 1.0f));
  t.w[1] = cubic1(lam);
  t.w[2] = cubic1(lam2);
  t.w[3] = cubic2(__fadd_rn(lam2, 1.0f));
  return t;
}

template <bool INTERIOR>
__device__ __forceinline__ void interpolate3(const uint8_t* __restrict__ img, int H, int W, const Taps& ty, const Taps& tx,
                                          uint8_t* res) {
  const int plane_sz = H * W;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const uint8_t* plane = img + (long long)c * plane_sz;
    float val = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint8_t* row = plane + ty.idx[i] * W;
      float s0, s1, s2, s3;
      if (INTERIOR) {
        const uint8_t* q = row + tx.idx[0];
        s0 = (float)q[0]; s1 = (float)q[1]; s2 = (float)q[2]; s3 = (float)q[3];
      } else {
        s0 = (float)row[tx.idx[0]]; s1 = (float)row[tx.idx[1]]; s2 = (float)row[tx.idx[2]]; s3 = (float)row[tx.idx[3]];
      }
      float acc = __fmul_rn(s0, tx.w[0]);
      acc = fmaf(s1, tx.w[1], acc);
      acc = fmaf(s2, tx.w[2], acc);
      acc = fmaf(s3, tx.w[3], acc);
      val = i == 0 ? __fmul_rn(acc, ty.w[0]) : fmaf(acc, ty.w[i], val);
    }
    val = fminf(fmaxf(val, 0.f), 255.f);
    res[c] = (uint8_t)__float2int_rn(val);   // round half to even, as torch.round
  }
}

// desc[i] = {byte offset of image i in blob, H, W, resized H', resized W', crop top, crop left, unused}
__global__ void __launch_bounds__(256) resize_crop_u8_kernel(const uint8_t* __restrict__ blob,
                                                             const long long* __restrict__ desc, int S,
                                                             int channel_first, uint8_t* __restrict__ out) {
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, n = blockIdx.z;
  if (x >= S || y >= S) return;
  const long long* d = desc + (long long)n * 8;
  const int H = (int)d[1], W = (int)d[2], nh = (int)d[3], nw = (int)d[4], top = (int)d[5], left = (int)d[6];
  const uint8_t* img = blob + d[0];
  const Taps ty = bicubic_taps(H, nh, top + y);
  const Taps tx = bicubic_taps(W, nw, left + x);
  uint8_t res[3];
  // away from the left / right border the four x taps are consecutive bytes: one address per (row, channel) and
  // immediate offsets instead of four clamped index computations (warp-uniform branch: only the warps touching a border take the clamped path)
  if (__all_sync(0xffffffffu, tx.idx[0] + 3 == tx.idx[3]))      // a warp = 32 consecutive x of one row
    interpolate3<true>(img, H, W, ty, tx, res);
  else
    interpolate3<false>(img, H, W, ty, tx, res);
  if (channel_first) {
#pragma unroll
    for (int c = 0; c < 3; ++c) out[(((long long)n * 3 + c) * S + y) * S + x] = res[c];
  } else {
    uint8_t* o = out + (((long long)n * S + y) * S + x) * 3;
    o[0] = res[0];
    o[1] = res[1];
    o[2] = res[2];
  }
}

}  // namespace

extern "C" int mic_resize_crop_u8(void* stream, const unsigned char* blob, const long long* desc, int n, int S,
                                  int channel_first, unsigned char* out) {
  MIC_CHECK_ARG(n >= 1 && n <= 65535 && S >= 1 && S <= 4096, "resize_crop: n=%d S=%d out of range", n, S);
  dim3 grid((S + 31) / 32, (S + 7) / 8, n), block(32, 8);
  resize_crop_u8_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(blob, desc, S, channel_first, out);
  MIC_CHECK_LAUNCH();
  return MIC_OK;
}
