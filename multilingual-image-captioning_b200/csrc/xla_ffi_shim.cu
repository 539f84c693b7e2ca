// XLA FFI custom-call handlers over the C-ABI (include/mic_b200.h): what lets the reference's Flax module and the
// main.py training loop keep running under a current JAX while the arithmetic executes in libmic_b200.so.
//
// Compiled ONLY where jaxlib's headers are on the include path (`xla/ffi/api/ffi.h`, shipped in
// jaxlib/include): this image has no jax / jaxlib, so here the translation unit is empty and the handlers below have
// never been compiled or run — they are the binding a maintainer adds, written against the public FFI API
// (jax >= 0.4.31; the reference's pinned jax==0.2.16 predates XLA FFI).  Python side: INTEGRATION.md section 2.
//
//   nvcc ... -I$(python -c "import jaxlib, os; print(os.path.join(os.path.dirname(jaxlib.__file__), 'include'))") \
//        -c xla_ffi_shim.cu
//
// Conventions: every handler takes the platform stream from the call frame, forwards raw device pointers, and maps a
// non-zero status to ffi::Error::Internal(mic_last_error()).  All C-ABI entry points are enqueue-only and
// allocation-free, i.e. valid inside XLA command buffers (CUDA graphs).
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define MIC_HAVE_XLA_FFI 1
#endif
#endif

#ifdef MIC_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include "../../include/mic_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {
inline ffi::Error status(int rc) { return rc == 0 ? ffi::Error::Success() : ffi::Error::Internal(mic_last_error()); }
template <class B> inline long long ld_of(const B& b) { return (long long)b.dimensions().back(); }

// y = act(x @ kernel + bias) (+ residual)   -- flax.linen.Dense with the Flax (in, out) kernel layout [E3,E4,D2-D4]
ffi::Error DenseImpl(cudaStream_t s, ffi::Buffer<ffi::BF16> x, ffi::Buffer<ffi::BF16> kernel, ffi::Buffer<ffi::F32> bias,
                     ffi::ResultBuffer<ffi::BF16> y, int32_t act) {
  const int K = (int)x.dimensions().back(), N = (int)kernel.dimensions().back();
  const int M = (int)(x.element_count() / K);
  return status(mic_gemm_bf16(s, /*a_mn_major=*/0, /*b_mn_major=*/1, x.typed_data(), K, kernel.typed_data(), N, M, N, K,
                              y->typed_data(), N, /*d_is_f32=*/0, /*accumulate=*/0, bias.typed_data(), act, nullptr, nullptr,
                              0, 0, 0, 0, nullptr, 0, 0.0f));
}

// flax.linen.LayerNorm forward (fp32 statistics kept for the backward handler)
ffi::Error LayerNormImpl(cudaStream_t s, ffi::Buffer<ffi::BF16> x, ffi::Buffer<ffi::F32> scale, ffi::Buffer<ffi::F32> bias,
                         ffi::ResultBuffer<ffi::BF16> y, ffi::ResultBuffer<ffi::F32> mean, ffi::ResultBuffer<ffi::F32> rstd,
                         float eps) {
  const int d = (int)x.dimensions().back(), M = (int)(x.element_count() / d);
  return status(mic_layernorm_fwd(s, x.typed_data(), scale.typed_data(), bias.typed_data(), eps, y->typed_data(),
                                  mean->typed_data(), rstd->typed_data(), M, d));
}

// tied lm_head + log-softmax / label-smoothed CE statistics (modeling_clip_vision_mbart.py:170-178 + main.py:658-680);
// `logits` is the bf16 [M, ld] buffer the backward handler rewrites in place as dlogits
ffi::Error LmHeadCeStatsImpl(cudaStream_t s, ffi::Buffer<ffi::BF16> h, ffi::Buffer<ffi::BF16> emb, ffi::Buffer<ffi::F32> flb,
                             ffi::Buffer<ffi::S32> labels, ffi::ResultBuffer<ffi::F32> pmax, ffi::ResultBuffer<ffi::F32> psum,
                             ffi::ResultBuffer<ffi::F32> psumz, ffi::ResultBuffer<ffi::F32> zlabel,
                             ffi::ResultBuffer<ffi::BF16> logits) {
  const int M = (int)h.dimensions()[0], K = (int)h.dimensions()[1], V = (int)emb.dimensions()[0];
  return status(mic_lm_head_ce_stats(s, h.typed_data(), K, emb.typed_data(), K, flb.typed_data(), labels.typed_data(), M, V, K,
                                     pmax->typed_data(), psum->typed_data(), psumz->typed_data(), zlabel->typed_data(),
                                     logits->typed_data(), ld_of(*logits)));
}

// softmax((q / sqrt(64)) k^T + mask) v for [B, T, H*64] projections (flax dot_product_attention) [E3,D2,D3]
ffi::Error AttentionImpl(cudaStream_t s, ffi::Buffer<ffi::BF16> q, ffi::Buffer<ffi::BF16> k, ffi::Buffer<ffi::BF16> v,
                         ffi::Buffer<ffi::S32> key_mask, ffi::ResultBuffer<ffi::BF16> out, ffi::ResultBuffer<ffi::F32> lse,
                         int32_t causal, int32_t heads) {
  const int B = (int)q.dimensions()[0], Tq = (int)q.dimensions()[1], Tk = (int)k.dimensions()[1];
  const long long ld = (long long)heads * 64;
  return status(mic_attention_fwd(s, q.typed_data(), ld, k.typed_data(), ld, v.typed_data(), ld, out->typed_data(), ld,
                                  lse->typed_data(), key_mask.element_count() ? key_mask.typed_data() : nullptr, causal, B,
                                  heads, Tq, Tk, 64, 0.125f));
}

// optax.adamw on the flat fp32 state (main.py:629-635,701); the step's scalars travel as attributes (by value)
ffi::Error AdamWImpl(cudaStream_t s, ffi::Buffer<ffi::F32> g, ffi::ResultBuffer<ffi::F32> p, ffi::ResultBuffer<ffi::F32> m,
                     ffi::ResultBuffer<ffi::F32> v, ffi::ResultBuffer<ffi::BF16> shadow, float lr, float b1, float b2,
                     float eps, float wd, float bc1, float bc2, float grad_scale) {
  return status(mic_adamw(s, p->typed_data(), m->typed_data(), v->typed_data(), g.typed_data(), shadow->typed_data(),
                          (long long)g.element_count(), lr, b1, b2, eps, wd, bc1, bc2, grad_scale));
}
}  // namespace

#define MIC_STREAM Ctx<ffi::PlatformStream<cudaStream_t>>()
XLA_FFI_DEFINE_HANDLER_SYMBOL(mic_ffi_dense, DenseImpl,
                              ffi::Ffi::Bind().MIC_STREAM.Arg<ffi::Buffer<ffi::BF16>>().Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::BF16>>().Attr<int32_t>("act"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(mic_ffi_layernorm, LayerNormImpl,
                              ffi::Ffi::Bind().MIC_STREAM.Arg<ffi::Buffer<ffi::BF16>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::BF16>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Attr<float>("eps"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(mic_ffi_lm_head_ce_stats, LmHeadCeStatsImpl,
                              ffi::Ffi::Bind().MIC_STREAM.Arg<ffi::Buffer<ffi::BF16>>().Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(mic_ffi_attention, AttentionImpl,
                              ffi::Ffi::Bind().MIC_STREAM.Arg<ffi::Buffer<ffi::BF16>>().Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>().Arg<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::BF16>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Attr<int32_t>("causal").Attr<int32_t>("heads"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(mic_ffi_adamw, AdamWImpl,
                              ffi::Ffi::Bind().MIC_STREAM.Arg<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::BF16>>()
                                  .Attr<float>("lr").Attr<float>("b1").Attr<float>("b2").Attr<float>("eps").Attr<float>("wd")
                                  .Attr<float>("bc1").Attr<float>("bc2").Attr<float>("grad_scale"));
#undef MIC_STREAM
#endif  // MIC_HAVE_XLA_FFI
