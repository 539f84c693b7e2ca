// Host launchers (C-ABI) for the tcgen05 GEMM family.  See include/mic_b200.h for the contract.
#include "gemm_sm100.cuh"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mic_b200.h"

// ------------------------------------------------------------------------------------------------
// error string / device info / tensor-map encode
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void mic_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* mic_last_error(void) { return g_err; }
extern "C" int mic_abi_version(void) { return MIC_B200_ABI_VERSION; }

thread_local MicLaunchOptions g_mic_launch = {0, 0, 0};   // measured on B200: PDL does not pay inside CUDA graphs (tools/microbench_pdl.py)
extern "C" int mic_launch_options(int programmatic_dependent_launch, int gemm_b_static, int gemm_sm_margin) {
  if (programmatic_dependent_launch >= 0) g_mic_launch.pdl = programmatic_dependent_launch != 0;
  if (gemm_b_static >= 0) g_mic_launch.static_b = gemm_b_static != 0;
  if (gemm_sm_margin >= 0) g_mic_launch.sm_margin = gemm_sm_margin < mic_num_sms() - 8 ? gemm_sm_margin : mic_num_sms() - 8;
  return MIC_OK;
}

int mic_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int mic_make_tmap_2d(CUtensorMap* out, const void* ptr, int elem_bytes, uint64_t inner, uint64_t outer, uint64_t ld,
                     uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  MIC_CHECK_ARG(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  MIC_CHECK_ARG((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA operand pointer %p not 16-byte aligned", ptr);
  MIC_CHECK_ARG((ld * elem_bytes) % 16 == 0, "TMA operand row pitch %llu elements x %d B not a multiple of 16 bytes",
                (unsigned long long)ld, elem_bytes);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MIC_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): inner=%llu outer=%llu ld=%llu box=%ux%u",
                (int)r, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, box_inner,
                box_outer);
  return MIC_OK;
}

// ------------------------------------------------------------------------------------------------
using namespace micgemm;

static int pick_block_n(int M, int N, int forced, bool splittable = false) {
  if (forced == 64 || forced == 128 || forced == 192 || forced == 256) return forced;
  // skinny-M (decode-time) GEMMs are weight streaming: narrow tiles so every weight matrix is pulled by
  // many SMs at once instead of N/256 of them
  if (M <= 2 * BLOCK_M && !splittable && N >= 512) return 64;
  // fp32 wgrad-style outputs can be split along K to fill the machine, so keep the efficient wide tile
  if (splittable && N > 128) return (N % 256 != 0 && N % 192 == 0) ? 192 : 256;
  const int sms = mic_num_sms();
  const int mb = (M + BLOCK_M - 1) / BLOCK_M;
  int best = 256;
  double best_cost = 1e30;
  const int cands[3] = {256, 192, 128};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    const long long tiles = (long long)mb * ((N + bn - 1) / bn);
    const long long waves = (tiles + sms - 1) / sms;
    // 128-wide tiles re-read A from smem twice as often per FLOP (UMMA operand bandwidth bound):
    // measured ~0.65x / 0.9x the 256-wide MMA rate on B200
    const double eff = bn == 256 ? 1.0 : (bn == 192 ? 0.9 : 0.65);
    const double cost = (double)waves * bn / eff;
    if (cost < best_cost) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

struct Operands {
  CUtensorMap ta, tb, td, td2;
  Shape shape;
  int bn;
};

static int setup_operands(Operands* o, int a_mn, int b_mn, const void* A, long long lda, const void* B, long long ldb,
                          int M, int N, int K, int block_n, int group_m, bool splittable = false) {
  MIC_CHECK_ARG(M > 0 && N > 0 && K > 0, "bad GEMM shape M=%d N=%d K=%d", M, N, K);
  memset(&o->td, 0, sizeof(CUtensorMap));
  memset(&o->td2, 0, sizeof(CUtensorMap));
  o->bn = pick_block_n(M, N, block_n, splittable);
  if (o->bn == 64 && !(a_mn == 0 && b_mn == 1)) o->bn = 128;
  int rc;
  if (!a_mn)
    rc = mic_make_tmap_bf16_2d(&o->ta, A, K, M, lda, BLOCK_K, BLOCK_M);
  else
    rc = mic_make_tmap_bf16_2d(&o->ta, A, M, K, lda, 64, BLOCK_K);
  if (rc) return rc;
  if (!b_mn)
    rc = mic_make_tmap_bf16_2d(&o->tb, B, K, N, ldb, BLOCK_K, o->bn);
  else
    rc = mic_make_tmap_bf16_2d(&o->tb, B, N, K, ldb, 64, BLOCK_K);
  if (rc) return rc;
  Shape& s = o->shape;
  s.M = M;
  s.N = N;
  s.K = K;
  s.num_m_blocks = (M + BLOCK_M - 1) / BLOCK_M;
  s.num_n_blocks = (N + o->bn - 1) / o->bn;
  if (group_m <= 0) group_m = ((long long)M * K * 2 <= (48ll << 20)) ? s.num_m_blocks : 16;
  s.group_m = group_m < s.num_m_blocks ? group_m : s.num_m_blocks;
  if (s.group_m < 1) s.group_m = 1;
  s.split_k = 1;
  s.kb_per_split = (K + BLOCK_K - 1) / BLOCK_K;
  return MIC_OK;
}

template <int A_MN, int B_MN, int BN, class Epi>
static int launch_one(cudaStream_t stream, const Operands& o, const typename Epi::Params& ep) {
  auto kern = gemm_kernel<A_MN, B_MN, BN, Epi>;
  static bool attr_set = false;
  if (!attr_set) {
    MIC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN, Epi::NBUF, Epi::EW>::SMEM_BYTES));
    attr_set = true;
  }
  const int tiles = o.shape.num_m_blocks * o.shape.num_n_blocks * o.shape.split_k;
  const int sms = mic_num_sms() - g_mic_launch.sm_margin;      // SMs left free for a concurrent collective
  const int grid = tiles < sms ? tiles : sms;
  // (plain launch: the tcgen05 GEMM does not take part in programmatic dependent launch - its producer loop is
  //  kept minimal; a PDL-aware variant with weight prefetch cost the training step 7 %)
  kern<<<grid, Cfg<BN, Epi::NBUF, Epi::EW>::THREADS, Cfg<BN, Epi::NBUF, Epi::EW>::SMEM_BYTES, stream>>>(o.ta, o.tb, o.td, o.td2, o.shape, ep);
  MIC_CHECK_LAUNCH();
  return MIC_OK;
}

template <int A_MN, int B_MN, class Epi>
static int launch_bn(cudaStream_t stream, const Operands& o, const typename Epi::Params& ep) {
  switch (o.bn) {
    case 256: return launch_one<A_MN, B_MN, 256, Epi>(stream, o, ep);
    case 192: return launch_one<A_MN, B_MN, 192, Epi>(stream, o, ep);
    case 64:
      if (A_MN == 0 && B_MN == 1) return launch_one<0, 1, 64, Epi>(stream, o, ep);   // forward layout only
      return launch_one<A_MN, B_MN, 128, Epi>(stream, o, ep);
    default: return launch_one<A_MN, B_MN, 128, Epi>(stream, o, ep);
  }
}

extern "C" int mic_gemm_bf16(void* stream, int a_mn_major, int b_mn_major, const void* A, long long lda,
                             const void* B, long long ldb, int M, int N, int K, void* D, long long ldd, int d_is_f32,
                             int accumulate, const float* bias, int act, void* D2, const void* residual,
                             long long ldr, int block_n, int group_m, int split_k, const unsigned int* drop_seed,
                             unsigned int drop_site, float drop_p) {
  Operands o;
  const bool splittable = d_is_f32 && !bias && act == MIC_ACT_NONE && !D2 && !residual && split_k != 1 &&
                          ((reinterpret_cast<uintptr_t>(D) & 15) == 0) && ((ldd * 4) % 16 == 0) && (N % 8 == 0) &&
                          K >= 2048;
  int rc = setup_operands(&o, a_mn_major, b_mn_major, A, lda, B, ldb, M, N, K, block_n, group_m, splittable);
  if (rc) return rc;
  MIC_CHECK_ARG(D != nullptr, "null output");
  MIC_CHECK_ARG(!(accumulate && !d_is_f32), "accumulate requires fp32 output");
  EpiStoreParams ep;
  ep.D = D;
  ep.ldd = ldd;
  ep.d_f32 = d_is_f32;
  ep.accumulate = accumulate;
  ep.bias = bias;
  ep.act = act;
  ep.D2 = reinterpret_cast<bf16*>(D2);
  ep.residual = reinterpret_cast<const bf16*>(residual);
  ep.ldr = ldr;
  ep.out_scale = 1.0f;
  ep.drop.seed_ptr = (drop_seed && drop_p > 0.f) ? drop_seed : nullptr;
  ep.drop.site = drop_site;
  ep.drop.thr16 = (uint32_t)(drop_p * 65536.0f + 0.5f);
  ep.drop.scale = 65536.0f / (65536.0f - (float)ep.drop.thr16);
  MIC_CHECK_ARG(!(ep.drop.seed_ptr && d_is_f32), "dropout epilogue is for bf16 outputs");
  const int esz = d_is_f32 ? 4 : 2;
  bool tma = ((reinterpret_cast<uintptr_t>(D) & 15) == 0) && ((ldd * esz) % 16 == 0) && (N % 8 == 0);
  if (bias) tma = tma && ((reinterpret_cast<uintptr_t>(bias) & 15) == 0);
  if (D2) tma = tma && ((reinterpret_cast<uintptr_t>(D2) & 15) == 0) && (ldd % 8 == 0);
  if (residual) tma = tma && ((reinterpret_cast<uintptr_t>(residual) & 15) == 0) && (ldr % 8 == 0);
  ep.tma_ok = tma ? 1 : 0;
  if (tma) {
    // output tiles leave through 32-row x 128-byte boxes (64 bf16 or 32 fp32 columns)
    rc = mic_make_tmap_2d(&o.td, D, esz, N, M, ldd, d_is_f32 ? 32 : 64, 32);
    if (rc) return rc;
    if (D2) {
      rc = mic_make_tmap_2d(&o.td2, D2, 2, N, M, ldd, 64, 32);
      if (rc) return rc;
    }
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // ---- split-K (fp32 outputs only): fills the machine when M*N yields less than a wave of tiles ----
  const int nkb = (K + BLOCK_K - 1) / BLOCK_K;
  const bool can_split = d_is_f32 && tma && !bias && act == MIC_ACT_NONE && !D2 && !residual;
  if (split_k == 0) {          // auto
    split_k = 1;
    if (can_split) {
      const int tiles = o.shape.num_m_blocks * o.shape.num_n_blocks, sms = mic_num_sms();
      double best = (double)((tiles + sms - 1) / sms);
      for (int sk = 2; sk <= 16; ++sk) {
        if (nkb / sk < 8) break;                     // keep >= 512 of K per slice
        const double cost = (double)((tiles * sk + sms - 1) / sms) / sk + 0.02 * sk;
        if (cost < best - 1e-9) {
          best = cost;
          split_k = sk;
        }
      }
    }
  }
  if (split_k > 1) {
    MIC_CHECK_ARG(can_split, "split_k needs a plain fp32 TMA-storable output (no bias/act/residual)");
    o.shape.kb_per_split = (nkb + split_k - 1) / split_k;
    o.shape.split_k = (nkb + o.shape.kb_per_split - 1) / o.shape.kb_per_split;
    if (o.shape.split_k > 1) {
      if (!accumulate) MIC_CHECK_CUDA(cudaMemset2DAsync(D, ldd * 4, 0, (size_t)N * 4, M, s));
      ep.accumulate = 1;
    }
  }
  if (act < 0) {
    // D = (A B^T) o act'(U): dgrad GEMM fused with the activation backward of the layer below; U enters as `residual`
    MIC_CHECK_ARG(!a_mn_major && !b_mn_major && tma && residual && !bias && !D2 && !d_is_f32 && o.bn == 256 &&
                      !ep.drop.seed_ptr,
                  "fused activation backward needs the (K,K) layout, bf16 TMA output, the pre-activation as `residual` "
                  "and 256-wide tiles");
    ep.act = -act;
    return launch_one<0, 0, 256, EpiStoreActBwd16>(s, o, ep);
  }
  // 16 epilogue warps (one 64-column group per warp and tile) for the activation epilogues (issue bound); tried for the
  // residual + dropout epilogues of the short-K GEMMs as well: no gain (profiles/r02_gemm_epilogue_costs.txt)
  if (!a_mn_major && b_mn_major && act != MIC_ACT_NONE && tma && o.bn == 256 && !residual)
    return launch_one<0, 1, 256, EpiStoreAct16>(s, o, ep);       // (its epilogue carries no residual path)
  if (!a_mn_major && b_mn_major) return launch_bn<0, 1, EpiStore>(s, o, ep);
  if (!a_mn_major && !b_mn_major) return launch_bn<0, 0, EpiStore>(s, o, ep);
  if (a_mn_major && b_mn_major) return launch_bn<1, 1, EpiStore>(s, o, ep);
  mic_set_error("GEMM layout (A MN-major, B K-major) is not instantiated");
  return MIC_ERR_INVALID;
}

// ------------------------------------------------------------------------------------------------
// fused lm_head + cross-entropy
// ------------------------------------------------------------------------------------------------
extern "C" int mic_lm_head_num_partials(int vocab) { return 2 * ((vocab + 255) / 256); }

extern "C" int mic_lm_head_ce_stats(void* stream, const void* H, long long ldh, const void* E, long long lde,
                                    const float* bias, const int* labels, int M, int V, int K, float* pmax,
                                    float* psum, float* psumz, float* zlabel, void* logits_out, long long ldl) {
  Operands o;
  int rc = setup_operands(&o, 0, 0, H, ldh, E, lde, M, V, K, 256, (M + BLOCK_M - 1) / BLOCK_M);
  if (rc) return rc;
  EpiCEStatsParams ep = {bias, labels, pmax, psum, psumz, zlabel, logits_out ? 1 : 0};
  if (logits_out) {
    MIC_CHECK_ARG(ldl % 256 == 0 && ldl >= V, "logits leading dimension %lld must be a multiple of 256 >= V", ldl);
    rc = mic_make_tmap_2d(&o.td, logits_out, 2, ldl, M, ldl, 64, 32);
    if (rc) return rc;
  }
  return launch_one<0, 0, 256, EpiCEStats>(reinterpret_cast<cudaStream_t>(stream), o, ep);
}

extern "C" int mic_lm_head_ce_grad(void* stream, const void* H, long long ldh, const void* E, long long lde,
                                   const float* bias, const int* labels, const float* lse, const float* row_w,
                                   float conf, float low, int M, int V, int K, void* dlogits, long long ldd) {
  MIC_CHECK_ARG(ldd % 256 == 0 && ldd >= V, "dlogits leading dimension %lld must be a multiple of 256 >= V", ldd);
  Operands o;
  int rc = setup_operands(&o, 0, 0, H, ldh, E, lde, M, V, K, 256, (M + BLOCK_M - 1) / BLOCK_M);
  if (rc) return rc;
  // cover the padded columns too so they are written as zeros
  o.shape.num_n_blocks = (int)(ldd / 256);
  rc = mic_make_tmap_2d(&o.td, dlogits, 2, ldd, M, ldd, 64, 32);
  if (rc) return rc;
  EpiCEGradParams ep = {bias, labels, lse, row_w, conf, low, reinterpret_cast<bf16*>(dlogits), ldd};
  return launch_one<0, 0, 256, EpiCEGrad>(reinterpret_cast<cudaStream_t>(stream), o, ep);
}

// partial lists per row produced by mic_lm_head_search (one per epilogue half of each CTA serving the row's m-block)
static int search_grid(int M) {
  const int mb = (M + BLOCK_M - 1) / BLOCK_M;
  int g = (mic_num_sms() / mb) * mb;
  return g < mb ? mb : g;
}
template <class Epi>
static int launch_search(void* stream, const Operands& o, const EpiSearchParams& ep, int M) {
  auto kern = gemm_kernel<0, 0, 256, Epi>;
  static bool attr_set = false;
  if (!attr_set) {
    MIC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<256>::SMEM_BYTES));
    attr_set = true;
  }
  kern<<<search_grid(M), NUM_THREADS, Cfg<256>::SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(o.ta, o.tb, o.td, o.td2,
                                                                                                   o.shape, ep);
  MIC_CHECK_LAUNCH();
  return MIC_OK;
}
extern "C" int mic_lm_head_search_num_partials(int M) { return 2 * (search_grid(M) / ((M + BLOCK_M - 1) / BLOCK_M)); }

extern "C" int mic_lm_head_search(void* stream, const void* H, long long ldh, const void* E, long long lde,
                                  const float* bias, int mask_token, int M, int V, int K, float* pmax, float* psum,
                                  float* cand_val, int* cand_idx, const float* upper_val, const int* upper_idx,
                                  const int* active, const unsigned int* gumbel_key) {
  const int mb = (M + BLOCK_M - 1) / BLOCK_M;
  MIC_CHECK_ARG(mb <= mic_num_sms(), "lm_head_search: M=%d rows exceed one m-block per SM", M);
  Operands o;
  int rc = setup_operands(&o, 0, 0, H, ldh, E, lde, M, V, K, 256, mb);
  if (rc) return rc;
  EpiSearchParams ep = {bias, mask_token, pmax, psum, cand_val, cand_idx, nullptr, nullptr, upper_val, upper_idx, active,
                        gumbel_key != nullptr, gumbel_key ? gumbel_key[0] : 0u, gumbel_key ? gumbel_key[1] : 0u};
  // fixed m-block per CTA: grid is a multiple of num_m_blocks and tiles are rasterised m-fastest
  // fixed m-block per CTA: grid is a multiple of num_m_blocks and tiles are rasterised m-fastest
  return gumbel_key ? launch_search<EpiSample>(stream, o, ep, M) : launch_search<EpiSearch>(stream, o, ep, M);
}

// ---- packed-operand lm_head search: K-major tile images + bulk copies ---------------------------------------
namespace {
// src [rows, K] row-major (pitch ld) -> tiles [row tile][k block] of tile_rows x 64 elements, each the SWIZZLE_128B
// image of a K-major tcgen05 operand (row r at r*128 B, 16-byte chunk c at c ^ (r & 7)); rows >= `rows` are zeros
__global__ void __launch_bounds__(256) pack_kmajor_tiles_kernel(const bf16* __restrict__ src, long long ld, int rows,
                                                                int num_kb, int tile_rows, bf16* __restrict__ out,
                                                                long long num_tiles) {
  for (long long t = blockIdx.x; t < num_tiles; t += gridDim.x) {
    const long long rt = t / num_kb;
    const int kb = (int)(t % num_kb);
    bf16* dst = out + t * tile_rows * 64;
    for (int i = threadIdx.x; i < tile_rows * 8; i += 256) {
      const int r = i >> 3, c = i & 7;
      const long long row = rt * tile_rows + r;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (row < rows) v = *reinterpret_cast<const uint4*>(src + row * ld + kb * 64 + c * 8);
      *reinterpret_cast<uint4*>(dst + r * 64 + ((c ^ (r & 7)) << 3)) = v;
    }
  }
}
}  // namespace

extern "C" long long mic_pack_kmajor_tiles_bytes(long long rows, int K, int tile_rows) {
  return ((rows + tile_rows - 1) / tile_rows) * tile_rows * (long long)K * 2;
}
extern "C" int mic_pack_kmajor_tiles(void* stream, const void* src, long long ld, long long rows, int K, int tile_rows,
                                     void* out) {
  MIC_CHECK_ARG(src && out && K % 64 == 0 && (tile_rows == 128 || tile_rows == 256) && rows > 0 && ld % 8 == 0,
                "pack_kmajor_tiles: K=%d must be a multiple of 64, tile_rows %d in {128, 256}", K, tile_rows);
  MIC_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 1023) == 0, "pack_kmajor_tiles: output must be 1024-byte aligned");
  const long long tiles = ((rows + tile_rows - 1) / tile_rows) * (K / 64);
  const int grid = (int)(tiles < 148ll * 16 ? tiles : 148ll * 16);
  pack_kmajor_tiles_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const bf16*)src, ld, (int)rows, K / 64, tile_rows, (bf16*)out, tiles);
  MIC_CHECK_LAUNCH();
  return MIC_OK;
}

extern "C" int mic_lm_head_search_packed(void* stream, const void* h_tiles, const void* e_tiles, const float* bias,
                                         int mask_token, int M, int V, int K, float* pmax, float* psum,
                                         float* cand_val, int* cand_idx, const float* upper_val,
                                         const int* upper_idx, const int* active, const unsigned int* gumbel_key) {
  const int mb = (M + BLOCK_M - 1) / BLOCK_M;
  MIC_CHECK_ARG(mb <= mic_num_sms(), "lm_head_search: M=%d rows exceed one m-block per SM", M);
  MIC_CHECK_ARG(K % 64 == 0 && h_tiles && e_tiles, "lm_head_search_packed: K=%d must be a multiple of 64", K);
  MIC_CHECK_ARG(((reinterpret_cast<uintptr_t>(h_tiles) | reinterpret_cast<uintptr_t>(e_tiles)) & 1023) == 0,
                "lm_head_search_packed: tile buffers must be 1024-byte aligned");
  Operands o;
  memset(&o, 0, sizeof(o));
  o.bn = 256;
  Shape& s = o.shape;
  s.M = M;
  s.N = V;
  s.K = K;
  s.num_m_blocks = mb;
  s.num_n_blocks = (V + 255) / 256;
  s.group_m = mb;
  s.split_k = 1;
  s.kb_per_split = K / BLOCK_K;
  EpiSearchParams ep = {bias, mask_token, pmax, psum, cand_val, cand_idx, (const bf16*)h_tiles, (const bf16*)e_tiles,
                        upper_val, upper_idx, active,
                        gumbel_key != nullptr, gumbel_key ? gumbel_key[0] : 0u, gumbel_key ? gumbel_key[1] : 0u};
  return gumbel_key ? launch_search<EpiSamplePacked>(stream, o, ep, M) : launch_search<EpiSearchPacked>(stream, o, ep, M);
}
