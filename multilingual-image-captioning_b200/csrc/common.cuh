// Shared device/host helpers for the sm_100a kernels: error plumbing, mbarrier / TMA / tcgen05 PTX
// wrappers, bf16 packing, warp reductions.  Hand-written PTX only — no CUTLASS/CuTe dependency.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------------
// host-side status plumbing (C-ABI: int32 status + thread-local last error string)
// ------------------------------------------------------------------------------------------------
void mic_set_error(const char* fmt, ...);
#define MIC_OK 0
#define MIC_ERR_INVALID 1
#define MIC_ERR_CUDA 2
#define MIC_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      mic_set_error(__VA_ARGS__);                \
      return MIC_ERR_INVALID;                    \
    }                                            \
  } while (0)
#define MIC_CHECK_CUDA(expr)                                                             \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      mic_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MIC_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)
#define MIC_CHECK_LAUNCH() MIC_CHECK_CUDA(cudaGetLastError())

// ------------------------------------------------------------------------------------------------
// launch plumbing: every kernel goes through mic_launch so that programmatic dependent launch (PDL) can be
// switched on for all of them.  A kernel launched with the attribute may start while its predecessor in the
// stream is still running; it must execute pdl_wait() before it reads or writes anything a predecessor
// touches (everything before that point - barrier init, TMEM alloc, descriptor prefetch, loads of operands
// the caller declared static - overlaps the predecessor's tail).  pdl_trigger() lets the successor start.
// ------------------------------------------------------------------------------------------------
struct MicLaunchOptions {
  int pdl;        // launch with cudaLaunchAttributeProgrammaticStreamSerialization
  int static_b;   // GEMM B operands (weights) are not written by any kernel in flight: prefetch before pdl_wait
  int sm_margin;  // persistent tcgen05 GEMMs leave this many SMs free (their CTAs fill a whole SM: a concurrent NCCL
                  // all-reduce otherwise only runs in the gaps between GEMM kernels, see training.py)
};
extern thread_local MicLaunchOptions g_mic_launch;

template <typename... KArgs, typename... Args>
inline cudaError_t mic_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_mic_launch.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 %%rx;\n\t"
      ".reg .pred %%px;\n\t"
      "elect.sync %%rx|%%px, %1;\n\t"
      "@%%px mov.s32 %0, 1;\n\t"
      "}\n"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred != 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// activations (transformers ACT2FN): 1 = gelu (exact erf), 2 = quick_gelu
#define MIC_ACT_NONE 0
#define MIC_ACT_GELU 1
#define MIC_ACT_QUICK_GELU 2
// erf via Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below bf16 resolution): 2 MUFU + ~10 FMA
// MUFU-only reciprocal / exp2 (no IEEE slow paths, hence no branches: the compiler can interleave the independent
// elements of an epilogue; with __frcp_rn / exp2f one gelu cost ~190 cycles in a lone warp)
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float erf_fast(float x) {
  const float ax = fabsf(x);
  const float t = rcp_approx(fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  p *= t;
  const float e = ex2_approx(-ax * ax * 1.4426950408889634f);
  return copysignf(fmaf(-p, e, 1.0f), x);
}
__device__ __forceinline__ float act_fwd(float x, int act) {
  if (act == MIC_ACT_GELU) return 0.5f * x * (1.0f + erf_fast(x * 0.70710678118654752f));
  if (act == MIC_ACT_QUICK_GELU) return x / (1.0f + __expf(-1.702f * x));
  return x;
}
__device__ __forceinline__ float act_bwd(float x, int act) {  // d act / dx
  if (act == MIC_ACT_GELU) {
    // d/dx [x Phi(x)] = Phi(x) + x phi(x); erf and the normal pdf share exp(-x^2/2)
    const float ax = fabsf(x) * 0.70710678118654752f;
    const float t = rcp_approx(fmaf(0.3275911f, ax, 1.0f));
    float p = fmaf(t, 1.061405429f, -1.453152027f);
    p = fmaf(t, p, 1.421413741f);
    p = fmaf(t, p, -0.284496736f);
    p = fmaf(t, p, 0.254829592f);
    p *= t;
    const float e = ex2_approx(-ax * ax * 1.4426950408889634f);      // = exp(-x^2/2)
    const float erfv = copysignf(fmaf(-p, e, 1.0f), x);
    return fmaf(0.5f, erfv, 0.5f) + x * 0.3989422804014327f * e;
  }
  if (act == MIC_ACT_QUICK_GELU) {
    float s = 1.0f / (1.0f + __expf(-1.702f * x));
    return s * (1.0f + 1.702f * x * (1.0f - s));
  }
  return 1.0f;
}

// ------------------------------------------------------------------------------------------------
// dropout: counter-based mask, reproducible between forward and backward.  One 32-bit hash (lowbias32) of
// (element-pair index, seed) yields two 16-bit uniforms; an element is DROPPED when its uniform < thr16.
// flax.linen.Dropout semantics: kept values are scaled by 1/(1-p).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t drop_hash(uint32_t pair_idx, uint32_t seed) {
  uint32_t h = pair_idx * 0x9E3779B1u ^ seed;
  h ^= h >> 16;
  h *= 0x7FEB352Du;
  h ^= h >> 15;
  h *= 0x846CA68Bu;
  h ^= h >> 16;
  return h;
}
// keep flags for elements (2*pair_idx, 2*pair_idx+1)
__device__ __forceinline__ void drop_keep2(uint32_t pair_idx, uint32_t seed, uint32_t thr16, bool* k0, bool* k1) {
  const uint32_t h = drop_hash(pair_idx, seed);
  *k0 = (h & 0xFFFFu) >= thr16;
  *k1 = (h >> 16) >= thr16;
}
struct DropoutParams {
  const uint32_t* seed_ptr;   // device scalar (so a captured graph sees a new seed every replay); null = off
  uint32_t site;              // distinguishes the dropout sites of one step
  uint32_t thr16;             // round(p * 65536)
  float scale;                // 65536 / (65536 - thr16)
};

// ------------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(addr),
      "r"(parity)
      : "memory");
}
// one NON-blocking probe (test_wait): try_wait may park the thread for an implementation-defined time before it
// returns false, which serialises a polling loop that has other work to interleave
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {   // one potentially-blocking probe
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// generic-proxy writes of other SMs (observed through an acquire) ordered before this thread's TMA reads
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// polling pair: spin on the relaxed load (an acquire load invalidates the SM's L1 every time it executes, which
// starves every other warp of the SM), then ONE acquire fence once the awaited value has been seen
__device__ __forceinline__ unsigned int ld_relaxed_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acquire_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Flag-array grid barrier (one word per CTA, holding the number of phases that CTA has completed; a launch-epoch base
// makes the values monotonic across launches, so nothing is ever reset).  Arrival = one st.release to the CTA's own
// word: unlike N atomics on ONE word, the arrivals do not serialise in the L2 atomic unit.  A whole warp polls: each
// lane reads ceil(G/32) words.  Wrap-safe signed comparison.
__device__ __forceinline__ void flags_wait_warp(const unsigned int* flags, int G, unsigned int target, int lane) {
  for (;;) {
    bool ok = true;
    for (int i = lane; i < G; i += 32) {
      unsigned int v;
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
      ok = ok && (int)(v - target) >= 0;
    }
    if (__all_sync(0xffffffffu, ok)) break;
  }
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
__device__ __forceinline__ unsigned int atom_add_release_gpu(unsigned int* p, unsigned int v) {
  unsigned int old;
  asm volatile("atom.add.release.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const void* src) {   // L2 -> smem, bypasses L1
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void red_add_v4_f32(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

// ------------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// 1-D bulk copy global -> shared (no tensor map): `bytes` contiguous bytes, 16-byte aligned both sides
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// TMA bulk tensor STORE (smem -> global), bulk-group completion.  The smem source must be visible to the
// async proxy (fence_proxy_async after the st.shared writes) before one thread issues the copy.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
// fp32 reduce-add variant (D += tile)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed groups of this thread have finished READING their smem source (buffer reusable)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate; one thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets row (lane base + t), columns c..c+31
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
        "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32_x8(uint32_t taddr, float* v) {   // 32 lanes x 8 consecutive columns
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor (sm_100 format, SWIZZLE_128B):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// UMMA instruction descriptor, kind::f16, bf16 x bf16 -> fp32
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// host: TMA descriptor encode (driver entry point fetched through the runtime; no -lcuda needed)
// ------------------------------------------------------------------------------------------------
// 2-D bf16 tensor, `inner` contiguous elements, row pitch `ld` elements, 128B swizzle.
int mic_make_tmap_2d(CUtensorMap* out, const void* ptr, int elem_bytes, uint64_t inner, uint64_t outer, uint64_t ld,
                     uint32_t box_inner, uint32_t box_outer);
inline int mic_make_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                                 uint32_t box_inner, uint32_t box_outer) {
  return mic_make_tmap_2d(out, ptr, 2, inner, outer, ld, box_inner, box_outer);
}
int mic_num_sms();
