// Persistent decoder-step kernel: ONE launch runs every decoder layer of a cached decode step
// (FlaxMBartDecoderLayer x L with past_key_values, modeling_clip_vision_mbart.py:519-651 -> SURVEY.md A.3) for all
// R = batch*beams rows.
//
// Why: at 256 rows every op of the step is latency-bound (a 2 MB weight GEMM is 0.3 us of HBM traffic but 8 us as a
// kernel, and a dependent kernel boundary inside a CUDA graph costs 2.6 us; profiles/r01_pdl_microbench.txt), and a
// step is a chain of ~130 dependent ops.  Here one CTA per SM stays resident for the whole step, the ops become
// PHASES separated by a grid barrier (one L2 atomic + an acquire poll), and - the part a kernel boundary cannot do -
// the producer keeps streaming the NEXT phase's weight tiles into the shared-memory ring while the grid is still
// draining the current phase: weights never depend on activations, so HBM stays busy across barriers.
//
// Operand movement: a 128-row TMA tensor box costs the TMA unit ~2.7 ns per 128-byte row here (0.5 us per k-block,
// measured: profiles/r01_decoder_step_phases.txt), so neither operand uses tensor maps.  Weights are re-packed once
// per generate() call into 8 KB tiles that already are the SWIZZLE_128B shared-memory image of a [64 k, 64 n]
// MN-major B operand, and the activations the GEMM phases consume are WRITTEN by their producers (LayerNorm rows,
// attention outputs, the fc1 epilogue) in the same tile-image layout (decode_device.cuh: tiled_off).  One operand
// stage is then two contiguous bulk copies (16 KB + 8 KB).
//
// Roles per CTA (384 threads):  warp 0 lane 0 = copy producer, warp 1 lane 0 = tcgen05 MMA issuer, warp 2 = TMEM
// allocator, warps 4..11 = epilogue of GEMM phases AND the workers of the vector phases (attention, LayerNorm).
// GEMM work unit = (128-row block, 64-column block, K slice).
#include "common.cuh"
#include "decode_device.cuh"

#include "../../include/mic_b200.h"

#include <stdlib.h>
#include <string.h>

int mic_num_sms();

namespace {
using namespace micdec;

constexpr int BM = 128, BN = 64, BK = 64, UK = 16;
constexpr int STAGES = 8;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BK * BN * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int EW = 8;                         // epilogue / vector warps
constexpr int THREADS = 128 + EW * 32;
constexpr int NACC = 1;                       // (4 independent accumulators per unit were tried: no gain, the MMA
                                              // chain is not what bounds a unit - the L2 -> SM operand stream is)
constexpr int TMEM_COLS = 2 * NACC * BN;      // double-buffered
constexpr int VEC_SCRATCH = EW * 2048;        // per-warp attention scratch: queries [4, 64] f32, numerators [64], spare
constexpr int BAR_BYTES = 512;               // 2*STAGES + 4 + 2*EW mbarriers + the TMEM base word
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + VEC_SCRATCH + BAR_BYTES + 1024;

enum { PH_GEMM_STORE = 0, PH_GEMM_RED = 1, PH_LN = 2, PH_ATTN = 3 };
// barrier state (256 uint32, zero before the first launch): [0] = arrival counter.  A per-CTA flag-word barrier
// ([SYNC_FLAGS + cta] = phases completed, launch epoch in [SYNC_EPOCH]; st.release arrival, one polling warp per CTA) was
// built and measured in round 2: 1.49 us bare vs 1.35 us for the counter, and 1,389 vs 1,015 us for the whole step
// (148 polling warps x 5 words load the L2 far more than 148 single-word pollers) -> only the microbenchmark keeps it
// (barrier_bench_kernel variant 4, profiles/r02_barrier_microbench.txt).
constexpr int SYNC_EPOCH = 32, SYNC_FLAGS = 64, OPT_KEEP_WRITER_FENCE = 1;

struct Phase {
  int kind;
  // ---- GEMM: D[R, N] = A[R, K] . B[K, N]
  const bf16* a_tiles;            // activations in tile-image layout [row tile][k block][16 KB]
  const bf16* b_tiles;            // packed weight [n tile][k block][8 KB]
  int n_tiles, num_kb, split_k, kb_per_split;
  int act;                        // GEMM_STORE: activation after bias
  int col_split;                  // GEMM_STORE: columns >= col_split go to out2 (the K|V cache slot of this step)
  int out_tiled_kb;               // GEMM_STORE: > 0 -> `out` is a tile-image buffer with that many k blocks per row
  int cache_T;                    // GEMM_STORE with out2: cache length (positions per head plane)
  const float* bias;              // GEMM_STORE: bias[N];  LN: bias of the GEMM that produced acc
  void* out;                      // GEMM_STORE: bf16 [R, ldo];  GEMM_RED / LN: fp32 accumulator [R, ldo]
  long long ldo;
  bf16* out2;                     // cache layer base, head-major: element (row, head, K|V, pos, c) at
                                  // row*ldo2 + (head*2 + kv)*cache_T*64 + pos*64 + chunk-swizzled c (decode_device.cuh)
  long long ldo2;
  // ---- LN: x += acc + bias; y = LN(x); acc = 0
  bf16* x;
  bf16* y;
  int y_tiled_kb;                 // > 0: y in tile-image layout (next GEMM's A operand); 0: row-major (lm_head input)
  const float* gamma;
  const float* beta;
  float eps;
  int d;
  // ---- attention
  DecAttnArgs attn;
  int keys_from_pos;              // self-attention: n_keys = pos + 1
};

struct StepArgs {
  const Phase* phases;
  int num_phases;
  int R;
  int pos;                        // position being decoded (cache slot written, n_keys = pos + 1)
  unsigned int* sync;             // grid-barrier arrival counter (zero before the launch; the kernel re-zeroes it)
  unsigned long long* prof;       // optional [num_phases, gridDim] globaltimer stamps of each CTA's phase arrival
  const int* active;              // optional device flag of the search loop (beam_cond / greedy_cond): 0 = every
                                  // hypothesis finished -> the whole step is skipped (the while_loop has ended)
  int opts;                       // bit 0: ALSO fence generic->async on the writer side (readers always fence after their
                                  // acquire, which is what orders the proxies; the extra fence cost 45 us per step)
};

// ---- The role loops below are written for INSTRUCTION COUNT: each is a single thread (or a single warp per work
// ---- item), so every instruction costs its full pipeline latency (~5 cycles) with nothing to hide it behind; the
// ---- first version's general "cursor" walk cost 0.45 us per k-block in the producer thread alone.
struct Unit {
  const bf16* a_src;              // first activation tile of the unit's row block (tile kb at + kb * 8192 elements)
  const bf16* b_src;              // first weight tile of the unit's column block (tile kb at + kb * 4096 elements)
  int kb_lo, len, rot, m, n;
};
__device__ __forceinline__ Unit unit_of(const Phase& ph, int u, int m_tiles) {
  Unit t;
  t.m = u % m_tiles;
  const int rest = u / m_tiles;
  t.n = rest % ph.n_tiles;
  const int ks = rest / ph.n_tiles;
  t.a_src = ph.a_tiles + (long long)t.m * ph.num_kb * (A_BYTES / 2);
  t.b_src = ph.b_tiles + (long long)t.n * ph.num_kb * (B_BYTES / 2);
  t.kb_lo = ks * ph.kb_per_split;
  t.len = min(t.kb_lo + ph.kb_per_split, ph.num_kb) - t.kb_lo;
  t.rot = 0;      // every CTA walks its K slice from the start: CTAs that share an activation tile then request it
                  // at the same moment and L2 serves them together (a per-CTA rotation was measured 1 % slower)
  return t;
}

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

constexpr int TRACE_PHASE = 1 + 8 + 11 * 5, TRACE_CTA = 1;   // fine trace (prof != null): fc1 of layer 5 on one CTA

__global__ void __launch_bounds__(THREADS, 1) decoder_step_kernel(const StepArgs args) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint8_t* vec = smem + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(vec + VEC_SCRATCH);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint64_t* attn_bar = bars + 2 * STAGES + 4;       // per vector warp: K operand / whole cross-attention stage
  uint64_t* attn_bar_v = attn_bar + EW;             // per vector warp: V operand of the self-attention pipeline
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4 + 2 * EW);
  static_assert((2 * STAGES + 4 + 2 * EW) * 8 + 4 <= BAR_BYTES, "barrier block overflows its shared-memory slot");

  // the search loop has terminated (device-side while_loop condition): nothing to decode.  The flag was written by
  // an earlier kernel of the stream, so every thread of every CTA reads the same value.
  if (args.active != nullptr && *args.active == 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x, G = gridDim.x;
  const int P = args.num_phases;
  const Phase* phases = args.phases;
  const int m_tiles = (args.R + BM - 1) / BM;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], EW);
    }
    for (int i = 0; i < 2 * EW; ++i) mbar_init(&attn_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== weight producer: runs ahead through phase boundaries =====================
      // arms the stage barrier with the bytes of BOTH operands; the activation copy is issued by warp 3
      int slot = 0, round = 0;
      for (int p = 0; p < P; ++p) {
        const Phase& ph = phases[p];
        if (ph.kind > PH_GEMM_RED) continue;
        const int units = m_tiles * ph.n_tiles * ph.split_k;
        for (int u = cta; u < units; u += G) {
          const Unit t = unit_of(ph, u, m_tiles);
          int kb = t.kb_lo + t.rot;
          const int kb_end = t.kb_lo + t.len;
#pragma unroll 1
          for (int i = 0; i < t.len; ++i) {
            if (round > 0) mbar_wait(&empty_bar[slot], (round - 1) & 1);
            mbar_arrive_expect_tx(&full_bar[slot], STAGE_BYTES);
            bulk_load(smem_b + slot * B_BYTES, t.b_src + (long long)kb * (B_BYTES / 2), B_BYTES, &full_bar[slot]);
            if (args.prof && cta == TRACE_CTA && p == TRACE_PHASE) args.prof[(long long)P * G + 32 + i] = gtime();
            if (++kb == kb_end) kb = t.kb_lo;
            if (++slot == STAGES) {
              slot = 0;
              ++round;
            }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== activation producer: gated on the grid barrier of the previous phase ==========
    int slot = 0, round = 0;
    for (int p = 0; p < P; ++p) {
      const Phase& ph = phases[p];
      if (ph.kind > PH_GEMM_RED) continue;
      const int units = m_tiles * ph.n_tiles * ph.split_k;
      if (cta >= units) continue;
      Unit t = unit_of(ph, cta, m_tiles);              // address arithmetic of the first unit before the wait
      if (p > 0) {
        if (lane == 0) {
          while ((int)(ld_relaxed_gpu(args.sync) / (unsigned int)G) < p) __nanosleep(20);
          fence_acquire_gpu();
          fence_proxy_async_global();
        }
      }
      if (lane == 0) {
        const bool trc = args.prof && cta == TRACE_CTA && p == TRACE_PHASE;
        if (trc) args.prof[(long long)P * G + 129] = gtime();
        for (int u = cta; u < units; u += G) {
          if (u != cta) t = unit_of(ph, u, m_tiles);
          int kb = t.kb_lo + t.rot;
          const int kb_end = t.kb_lo + t.len;
#pragma unroll 1
          for (int i = 0; i < t.len; ++i) {
            if (round > 0) mbar_wait(&empty_bar[slot], (round - 1) & 1);
            bulk_load(smem_a + slot * A_BYTES, t.a_src + (long long)kb * (A_BYTES / 2), A_BYTES, &full_bar[slot]);
            if (trc) args.prof[(long long)P * G + i] = gtime();
            if (++kb == kb_end) kb = t.kb_lo;
            if (++slot == STAGES) {
              slot = 0;
              ++round;
            }
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 1);
      constexpr uint32_t a_kstep = UK * 2, b_kstep = UK * 128;
      const uint64_t da0 = umma_smem_desc(smem_u32(smem_a), 0, 1024);
      const uint64_t db0 = umma_smem_desc(smem_u32(smem_b), BK * 128, 1024);
      int slot = 0;
      uint32_t fpar = 0;
      uint32_t it = 0;
      for (int p = 0; p < P; ++p) {
        const Phase& ph = phases[p];
        if (ph.kind > PH_GEMM_RED) continue;
        const int units = m_tiles * ph.n_tiles * ph.split_k;
        const int nkb = ph.num_kb, kps = ph.kb_per_split, per_split = m_tiles * ph.n_tiles;
        for (int u = cta; u < units; u += G, ++it) {
          const int kb_lo = (u / per_split) * kps;
          const int len = min(kb_lo + kps, nkb) - kb_lo;
          const uint32_t as = it & 1, aphase = (it >> 1) & 1;
          mbar_wait(&tmem_empty[as], aphase ^ 1);
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + as * (NACC * BN);
#pragma unroll 1
          for (int i = 0; i < len; ++i) {
            mbar_wait(&full_bar[slot], fpar);
            tcgen05_fence_after();
            if (args.prof && cta == TRACE_CTA && p == TRACE_PHASE) args.prof[(long long)P * G + 64 + i] = gtime();
            const uint64_t da = da0 + (uint64_t)((slot * A_BYTES) >> 4);
            const uint64_t db = db0 + (uint64_t)((slot * B_BYTES) >> 4);
#pragma unroll
            for (int k = 0; k < BK / UK; ++k)
              umma_bf16(tmem_d, da + (uint64_t)((k * a_kstep) >> 4), db + (uint64_t)((k * b_kstep) >> 4), idesc,
                        (i > 0) || (k > 0));
            umma_commit(&empty_bar[slot]);          // frees the ring slot once these MMAs retire
            if (++slot == STAGES) {
              slot = 0;
              fpar ^= 1;
            }
          }
          umma_commit(&tmem_full[as]);
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue + vector phases =====================
    const int ew = warp - 4;
    const int quarter = warp & 3;              // TMEM lane quarter this warp may read
    const int half = ew >> 2;                  // which 32-column half of the 64-column tile
    // vector phases: the A half of the operand ring is idle then (activation loads of the next GEMM phase are gated
    // on this phase's barrier; only its weight tiles are prefetched, into the B half) -> 16 KB K/V stage per warp
    uint8_t* kv_stage = smem_a + ew * A_BYTES * (STAGES / EW);
    float* q_smem = reinterpret_cast<float*>(vec + ew * 2048);
    float* ln_part = reinterpret_cast<float*>(vec + EW * 2048 - 256);      // [2 parities][8 warps][2] LayerNorm partials
    const int vt = threadIdx.x - 128;          // 0..255 among the vector warps
    const bool leader = threadIdx.x == 128;
    uint32_t it = 0;
    uint32_t attn_parity = 0, attn_parity_v = 0;  // phase parities of this warp's attention-stage barriers
    for (int p = 0; p < P; ++p) {
      const Phase& ph = phases[p];
      if (ph.kind <= PH_GEMM_RED) {
        const int units = m_tiles * ph.n_tiles * ph.split_k;
        if (cta >= units && p > 0) {
          // no work here: still do not arrive for phase p before phase p-1 is complete everywhere, so that
          // (arrivals / G) counts whole phases (CTAs with work inherit this from their gated activation loads)
          if (leader) {
            while ((int)(ld_relaxed_gpu(args.sync) / (unsigned int)G) < p) __nanosleep(20);
            fence_acquire_gpu();
          }
        }
        for (int u = cta; u < units; u += G, ++it) {
          const int m = u % m_tiles, n = (u / m_tiles) % ph.n_tiles;
          const uint32_t as = it & 1, aphase = (it >> 1) & 1;
          const int row = m * BM + quarter * 32 + lane;
          const int col0 = n * BN + half * 32;
          float bias_r[32];                            // fetched while the MMAs are still running
          if (ph.kind == PH_GEMM_STORE) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) load8f(ph.bias + col0 + j, bias_r + j);
          }
          mbar_wait(&tmem_full[as], aphase);
          tcgen05_fence_after();
          if (args.prof && cta == TRACE_CTA && p == TRACE_PHASE && leader) args.prof[(long long)P * G + 128] = gtime();
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * (NACC * BN) + half * 32;
          const bool to_cache = col0 >= ph.col_split;
          bf16* dst = reinterpret_cast<bf16*>(ph.out) + (long long)row * ph.ldo + col0;
          int cache_chunk0 = 0;
          if (to_cache) {
            // this position's K|V row of (row, head): 32 of its 64 values, as 4 chunks swizzled by the position
            const int kvcol = col0 - ph.col_split, dm = ph.col_split;          // col_split == d_model
            const int kv = kvcol >= dm, head = (kvcol - kv * dm) >> 6;
            dst = ph.out2 + (long long)row * ph.ldo2 + (long long)(head * 2 + kv) * ph.cache_T * 64 + args.pos * 64;
            cache_chunk0 = (kvcol & 63) >> 3;
          }
          float* racc = reinterpret_cast<float*>(ph.out) + (long long)row * ph.ldo + col0;
          float v[32];
          tmem_ld_32x32(taddr, v);
          tmem_ld_wait();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[as]);      // accumulator drained: hand the TMEM buffer back
          if (args.prof && cta == TRACE_CTA && p == TRACE_PHASE && leader) args.prof[(long long)P * G + 132] = gtime();
          if (row < args.R) {
            if (ph.kind == PH_GEMM_RED) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) red_add_v4_f32(racc + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
              // fully unrolled on purpose: four independent 8-column chains give the lone warp some ILP
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                if (ph.act == MIC_ACT_GELU) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) v[j + e] = act_fwd(v[j + e] + bias_r[j + e], MIC_ACT_GELU);
                } else {
#pragma unroll
                  for (int e = 0; e < 8; ++e) v[j + e] += bias_r[j + e];
                }
                if (j == 24 && args.prof && cta == TRACE_CTA && p == TRACE_PHASE && leader) {
                  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(args.prof[(long long)P * G + 133]) : "f"(v[31]), "f"(v[0]), "f"(v[8]), "f"(v[16]));
                }
                if (!to_cache && ph.out_tiled_kb)
                  store8(reinterpret_cast<bf16*>(ph.out) + tiled_off(row, col0 + j, ph.out_tiled_kb), v + j);
                else if (to_cache)
                  store8(dst + (((cache_chunk0 + (j >> 3)) ^ (args.pos & 7)) << 3), v + j);
                else
                  store8(dst + j, v + j);
              }
            }
          }
        }
      } else {
        // vector phase: inputs come from earlier phases of other CTAs
        if (p > 0) {
          if (leader) {
            while ((int)(ld_relaxed_gpu(args.sync) / (unsigned int)G) < p) __nanosleep(20);
            fence_acquire_gpu();
          }
          asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
          // attention operands arrive through the async proxy (bulk copies) but were written with generic stores by
          // other CTAs (this position's K|V, the cross-attention queries' accumulator is read generically): order
          // the acquired writes before this thread's copies
          if (ph.kind == PH_ATTN) fence_proxy_async_global();
        }
        if (ph.kind == PH_LN) {
          // x += acc + bias; y = LN(x); acc = 0.  Two rows per CTA at a time, 128 vector threads (8 columns each)
          // per row: a lone warp per row is instruction-latency bound (5 us per row measured).
          float* acc = reinterpret_cast<float*>(ph.out);
          const int d = ph.d;
          const int sel = vt >> 7, c = (vt & 127) * 8;           // which row of the pair, first column
          int par = 0;
          for (int rowa = cta; rowa < args.R; rowa += 2 * G, par ^= 1) {
            const int row = rowa + sel * G;
            const bool on = row < args.R && c < d;
            float v[8], g[8], be[8];
            float s = 0.f, s2 = 0.f;
            if (on) {
              float ar[8], br[8];
              load8_cg(ph.x + (long long)row * d + c, v);
              load8f_cg(acc + (long long)row * d + c, ar);
              load8f(ph.bias + c, br);
              load8f(ph.gamma + c, g);
              load8f(ph.beta + c, be);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                v[j] = bf16_round(v[j] + ar[j] + br[j]);
                s += v[j];
                s2 += v[j] * v[j];
              }
              store8(ph.x + (long long)row * d + c, v);
              const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
              *reinterpret_cast<float4*>(acc + (long long)row * d + c) = z;
              *reinterpret_cast<float4*>(acc + (long long)row * d + c + 4) = z;
            }
            s = warp_sum(s);
            s2 = warp_sum(s2);
            if (lane == 0) {
              ln_part[par * 16 + ew * 2] = s;                    // warps 0..3 -> row a, 4..7 -> row b
              ln_part[par * 16 + ew * 2 + 1] = s2;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
            float ts = 0.f, ts2 = 0.f;
#pragma unroll
            for (int w = 0; w < EW / 2; ++w) {
              ts += ln_part[par * 16 + (sel * 4 + w) * 2];
              ts2 += ln_part[par * 16 + (sel * 4 + w) * 2 + 1];
            }
            const float mean = ts / d;
            const float rstd = rsqrtf(fmaxf(ts2 / d - mean * mean, 0.f) + ph.eps);
            if (on) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = (v[j] - mean) * rstd * g[j] + be[j];
              store8(ph.y + (ph.y_tiled_kb ? tiled_off(row, c, ph.y_tiled_kb) : (long long)row * d + c), v);
            }
          }
        } else {
          const int gw = ew * G + cta, nw = G * EW;      // consecutive items land on different SMs
          DecAttnArgs a = ph.attn;
          if (ph.keys_from_pos) a.n_keys = args.pos + 1;
          // host guarantees n_keys <= 64 and rows_per_kv <= 8 (mic_decoder_plan_init): the item's query rows sit in
          // rows 0..7 of the MMA tile
          const int items = ((a.R + a.rows_per_kv - 1) / a.rows_per_kv) * a.H;
#pragma unroll 1
          const bool tra = args.prof && cta == TRACE_CTA && ew == 0 && p == 1 + 1 + 11 * 5;
          if (tra && lane == 0) args.prof[(long long)P * G + 169] = gtime();
          if (ph.keys_from_pos && a.q != nullptr) {
            decode_self_attn_runs(a, gw, nw, items, kv_stage, reinterpret_cast<bf16*>(q_smem), &attn_bar[ew],
                                  &attn_bar_v[ew], &attn_parity, &attn_parity_v, lane);
          } else if (a.kv_tiles != nullptr) {
            for (int i = gw; i < items; i += nw)
              decode_cross_attn_packed(a, i / a.H, i % a.H, kv_stage, reinterpret_cast<bf16*>(q_smem), &attn_bar[ew],
                                       &attn_parity, lane);
          } else {
            for (int i = gw; i < items; i += nw) {
              int row0, row1;
              decode_attn_rows(a, i / a.H, lane, &row0, &row1);
              decode_attn_group_mma(a, i / a.H, i % a.H, row0, row1, kv_stage, reinterpret_cast<bf16*>(q_smem), lane,
                                    nullptr);
            }
          }
          if (tra && lane == 0) args.prof[(long long)P * G + 168] = gtime();
        }
      }
      // generic-proxy writes of this phase (tile-image activations in global memory, ring scratch in shared memory)
      // are ordered before the async-proxy bulk copies that follow the barrier
      if (args.prof && cta == TRACE_CTA && p == TRACE_PHASE && leader) args.prof[(long long)P * G + 130] = gtime();
      // only phases whose output is fetched by bulk copies (tile-image activations) or that used the ring as scratch
      // need the generic -> async proxy fence (0.6 us); split-K reductions and the q|k|v store feed generic loads
      if ((args.opts & OPT_KEEP_WRITER_FENCE) &&
          (ph.kind == PH_LN || ph.kind == PH_ATTN || (ph.kind == PH_GEMM_STORE && (ph.out_tiled_kb || ph.out2))))
        fence_proxy_async_global();
      if (args.prof && cta == TRACE_CTA && p == TRACE_PHASE && leader) args.prof[(long long)P * G + 131] = gtime();
      // ---- grid barrier arrival: this CTA's writes of phase p are done.  The leader's release (gpu scope) is
      // cumulative over the other warps' writes, which precede it through the CTA barrier.
      asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
      if (leader) {
        if (args.prof) args.prof[(long long)p * G + cta] = gtime();
        const unsigned int old = atom_add_release_gpu(args.sync, 1u);
        if (p == P - 1 && old == (unsigned int)P * (unsigned int)G - 1u) {
          __threadfence();
          *reinterpret_cast<volatile unsigned int*>(args.sync) = 0u;     // last arrival of the launch: re-arm
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- grid-barrier microbenchmark (tools/microbench_barrier.py): n barriers of the kind the step kernel uses,
// ---- variant 0 = as used (atom.add.release + relaxed polling with nanosleep(20) + one acquire fence),
// ---- 1 = polling without sleep, 2 = every thread fences before the CTA barrier, 3 = 256 pollers per CTA
__global__ void __launch_bounds__(THREADS, 1) barrier_bench_kernel(unsigned int* sync, int n, int variant) {
  const int G = gridDim.x;
  if (threadIdx.x < 128) return;
  const bool leader = threadIdx.x == 128;
  if (variant == 4) {          // per-CTA flag words + launch epoch (OPT_FLAG_BARRIER)
    const unsigned int epoch = ld_relaxed_gpu(sync + SYNC_EPOCH);
    const int w = (threadIdx.x - 128) >> 5, lane = threadIdx.x & 31;
    for (int p = 0; p < n; ++p) {
      if (p > 0) {
        if (w == 0) flags_wait_warp(sync + SYNC_FLAGS, G, epoch + (unsigned int)p, lane);
        asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
      if (leader) {
        st_release_gpu(sync + SYNC_FLAGS + blockIdx.x, epoch + (unsigned int)p + 1u);
        if (p == n - 1 && blockIdx.x == 0) *reinterpret_cast<volatile unsigned int*>(sync + SYNC_EPOCH) = epoch + (unsigned int)n;
      }
    }
    return;
  }
  for (int p = 0; p < n; ++p) {
    if (p > 0) {
      if (leader || variant == 3) {
        while ((int)(ld_relaxed_gpu(sync) / (unsigned int)G) < p) {
          if (variant != 1) __nanosleep(20);
        }
        fence_acquire_gpu();
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
    }
    if (variant == 2) __threadfence();
    asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
    if (leader) {
      const unsigned int old = atom_add_release_gpu(sync, 1u);
      if (p == n - 1 && old == (unsigned int)n * (unsigned int)G - 1u) {
        __threadfence();
        *reinterpret_cast<volatile unsigned int*>(sync) = 0u;
      }
    }
  }
}

// ---- weight re-pack: Flax kernel W[K, N] (row-major, pitch ld) -> tiles [n tile][k block] of 8 KB, each the
// ---- SWIZZLE_128B image of a [64 k rows x 64 n] MN-major operand (k row kk at kk*128 B, chunk c at c ^ (kk & 7))
struct PackJob {
  const bf16* w;
  long long ld;
  int K, N;
  bf16* out;
};
struct PackJobs {
  PackJob j[6];
};
__global__ void __launch_bounds__(128) pack_weight_tiles_kernel(const PackJobs jobs) {
  const PackJob& job = jobs.j[blockIdx.y];
  const int num_kb = job.K / BK, tiles = (job.N / BN) * num_kb;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int n = t / num_kb, kb = t % num_kb;
    bf16* dst = job.out + (long long)t * (B_BYTES / 2);
    for (int i = threadIdx.x; i < 512; i += 128) {
      const int kk = i >> 3, c = i & 7;
      const uint4 v = *reinterpret_cast<const uint4*>(job.w + (long long)(kb * BK + kk) * job.ld + n * BN + c * 8);
      *reinterpret_cast<uint4*>(dst + kk * 64 + ((c ^ (kk & 7)) << 3)) = v;
    }
  }
}

// ---- cross K|V re-pack: enc_kv [B*S, ld] (layer l: K at l*2d, V at l*2d + d) -> [image][layer][head] stage images
// ---- of 16 KB (K rows 0..63, V rows 0..63; 128 B rows, chunk c of row j at c ^ (j & 7); rows >= S zero)
__global__ void __launch_bounds__(256) pack_cross_kv_kernel(const bf16* __restrict__ enc_kv, long long ld, int B,
                                                            int S, int L, int H, int d, bf16* __restrict__ out) {
  const long long items = (long long)B * L * H;
  for (long long it = blockIdx.x; it < items; it += gridDim.x) {
    const int h = (int)(it % H), l = (int)((it / H) % L), b = (int)(it / ((long long)H * L));
    bf16* dst = out + it * 8192;
    for (int i = threadIdx.x; i < 2 * 64 * 8; i += 256) {
      const int c = i & 7, j = (i >> 3) & 63, kv = i >> 9;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (j < S) v = *reinterpret_cast<const uint4*>(enc_kv + ((long long)b * S + j) * ld + (long long)l * 2 * d + kv * d + h * 64 + c * 8);
      *reinterpret_cast<uint4*>(dst + kv * 4096 + j * 64 + ((c ^ (j & 7)) << 3)) = v;
    }
  }
}

}  // namespace

extern "C" long long mic_decoder_cross_kv_tiles_bytes(int B, int num_layers, int heads) {
  return (long long)B * num_layers * heads * 16384;
}
extern "C" int mic_decoder_pack_cross_kv(void* stream, const void* enc_kv, long long ld, int B, int S, int num_layers,
                                         int heads, int d_model, void* out) {
  MIC_CHECK_ARG(enc_kv && out && S >= 1 && S <= 64 && d_model == heads * 64, "pack_cross_kv: S=%d must be <= 64", S);
  MIC_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 1023) == 0, "pack_cross_kv: output must be 1024-byte aligned");
  const long long items = (long long)B * num_layers * heads;
  const int grid = (int)(items < 148ll * 8 ? items : 148ll * 8);
  pack_cross_kv_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>((const bf16*)enc_kv, ld, B, S, num_layers,
                                                                                 heads, d_model, (bf16*)out);
  MIC_CHECK_LAUNCH();
  return MIC_OK;
}

// ------------------------------------------------------------------------------------------------
// host side: the plan (phase table) lives in a caller-provided device buffer
// ------------------------------------------------------------------------------------------------
static const int kPhasesPerLayer = 11;
static long long layer_packed_elems(int d, int F) { return 6ll * d * d + 2ll * d * F; }   // qkv 3dd, 3 x dd, fc1, fc2

extern "C" long long mic_decoder_plan_bytes(int num_layers) {
  return (long long)(sizeof(Phase) * (size_t)(1 + kPhasesPerLayer * num_layers) + 256);
}
extern "C" long long mic_decoder_packed_bytes(int num_layers, int d_model, int ffn_dim) {
  return 2ll * num_layers * layer_packed_elems(d_model, ffn_dim);
}

extern "C" int mic_decoder_pack_weights(void* stream, const mic_decoder_layer_t* layers, int num_layers, int d_model,
                                        int ffn_dim, void* packed) {
  MIC_CHECK_ARG(layers && packed && num_layers > 0, "decoder pack: null argument");
  MIC_CHECK_ARG(d_model % 64 == 0 && ffn_dim % 64 == 0, "decoder pack: d_model/ffn_dim must be multiples of 64");
  MIC_CHECK_ARG((reinterpret_cast<uintptr_t>(packed) & 1023) == 0, "decoder pack: buffer must be 1024-byte aligned");
  const int d = d_model, F = ffn_dim;
  for (int l = 0; l < num_layers; ++l) {
    const mic_decoder_layer_t& w = layers[l];
    bf16* base = reinterpret_cast<bf16*>(packed) + (long long)l * layer_packed_elems(d, F);
    PackJobs jobs;
    jobs.j[0] = {(const bf16*)w.sa_qkv_w, 3ll * d, d, 3 * d, base};
    jobs.j[1] = {(const bf16*)w.sa_o_w, d, d, d, base + 3ll * d * d};
    jobs.j[2] = {(const bf16*)w.ca_q_w, d, d, d, base + 4ll * d * d};
    jobs.j[3] = {(const bf16*)w.ca_o_w, d, d, d, base + 5ll * d * d};
    jobs.j[4] = {(const bf16*)w.fc1_w, F, d, F, base + 6ll * d * d};
    jobs.j[5] = {(const bf16*)w.fc2_w, d, F, d, base + 6ll * d * d + (long long)d * F};
    MIC_CHECK_CUDA(mic_launch(pack_weight_tiles_kernel, dim3(256, 6), dim3(128), 0,
                              reinterpret_cast<cudaStream_t>(stream), jobs));
  }
  return MIC_OK;
}

extern "C" int mic_decoder_plan_init(void* stream, void* plan_dev, const mic_decoder_layer_t* layers, int num_layers,
                                     const mic_decoder_buffers_t* buf, const void* packed, int R, int d_model,
                                     int heads, int ffn_dim, int cache_len, int enc_tokens, int rows_per_image,
                                     long long ld_enc, int act, float eps) {
  MIC_CHECK_ARG(plan_dev && layers && buf && packed && num_layers > 0, "decoder plan: null argument");
  MIC_CHECK_ARG(d_model % 64 == 0 && d_model <= 1024 && ffn_dim % 64 == 0 && d_model == heads * HD,
                "decoder plan: d_model=%d heads=%d ffn=%d unsupported (head_dim 64, d_model <= 1024)", d_model, heads,
                ffn_dim);
  MIC_CHECK_ARG(cache_len <= 64 && enc_tokens <= 64 && rows_per_image >= 1 && rows_per_image <= 8 && R > 0,
                "decoder plan: cache_len=%d / enc_tokens=%d must be <= 64 and rows_per_image=%d in 1..8", cache_len,
                enc_tokens, rows_per_image);
  MIC_CHECK_ARG(act == MIC_ACT_GELU || act == MIC_ACT_NONE, "decoder plan: activation %d not supported (gelu only)", act);
  MIC_CHECK_ARG((reinterpret_cast<uintptr_t>(plan_dev) & 127) == 0, "decoder plan buffer must be 128-byte aligned");
  MIC_CHECK_ARG(((reinterpret_cast<uintptr_t>(buf->a_tiles) | reinterpret_cast<uintptr_t>(buf->o_tiles) |
                  reinterpret_cast<uintptr_t>(buf->g_tiles) | reinterpret_cast<uintptr_t>(packed)) & 1023) == 0,
                "decoder plan: tile-image buffers must be 1024-byte aligned");
  const int d = d_model, F = ffn_dim, L = num_layers;
  const size_t total = (size_t)mic_decoder_plan_bytes(L);
  static thread_local uint8_t* host = nullptr;
  static thread_local size_t host_cap = 0;
  if (host_cap < total) {
    free(host);
    host = (uint8_t*)aligned_alloc(128, (total + 127) / 128 * 128);
    host_cap = total;
  }
  MIC_CHECK_ARG(host != nullptr, "decoder plan: host staging allocation failed");
  memset(host, 0, total);
  Phase* all = reinterpret_cast<Phase*>(host);
  const float scale = 1.0f / sqrtf((float)HD);
  const bf16* a_t = (const bf16*)buf->a_tiles;
  const bf16* o_t = (const bf16*)buf->o_tiles;
  const bf16* g_t = (const bf16*)buf->g_tiles;
  auto gemm = [&](Phase& p, int kind, const bf16* a_tiles, const bf16* b_tiles, int N, int K, int split,
                  const float* bias, void* out, long long ldo) {
    p.kind = kind;
    p.a_tiles = a_tiles;
    p.b_tiles = b_tiles;
    p.n_tiles = N / BN;
    p.num_kb = K / BK;
    p.kb_per_split = (p.num_kb + split - 1) / split;
    p.split_k = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;
    p.act = MIC_ACT_NONE;
    p.col_split = 1 << 30;
    p.bias = bias;
    p.out = out;
    p.ldo = ldo;
  };
  auto ln = [&](Phase& p, float* acc, const float* bias, const float* gamma, const float* beta, void* y, int y_kb) {
    p.kind = PH_LN;
    p.out = acc;
    p.ldo = d;
    p.bias = bias;
    p.x = (bf16*)buf->x;
    p.y = (bf16*)y;
    p.y_tiled_kb = y_kb;
    p.gamma = gamma;
    p.beta = beta;
    p.eps = eps;
    p.d = d;
  };
  // phase 0: a = LN_0(x) (self_attn_layer_norm of layer 0) in tile-image layout.  Written as the residual form with
  // a zero accumulator and a zero "bias" (row 0 of the still-zero q_acc), which leaves x bit-identical.
  ln(all[0], buf->acc, buf->q_acc, layers[0].ln_sa_g, layers[0].ln_sa_b, buf->a_tiles, d / BK);
  for (int l = 0; l < L; ++l) {
    const mic_decoder_layer_t& w = layers[l];
    const bf16* pk = reinterpret_cast<const bf16*>(packed) + (long long)l * layer_packed_elems(d, F);
    Phase* p = all + 1 + kPhasesPerLayer * l;
    // 0: q|k|v projection of LN(x): q -> buf.q, k|v -> this position's cache slot
    gemm(p[0], PH_GEMM_STORE, a_t, pk, 3 * d, d, 1, w.sa_qkv_b, buf->q, d);
    p[0].col_split = d;
    p[0].out2 = (bf16*)w.self_kv;
    p[0].ldo2 = (long long)cache_len * 2 * d;
    p[0].cache_T = cache_len;
    // 1: cached self-attention through the ancestor table
    p[1].kind = PH_ATTN;
    p[1].keys_from_pos = 1;
    DecAttnArgs& sa = p[1].attn;
    sa.q = (const bf16*)buf->q; sa.ldq = d; sa.kc = (const bf16*)w.self_kv; sa.vc = nullptr;   // head-major cache
    sa.ldkv = 2 * d; sa.anc = buf->ancestors; sa.T = cache_len; sa.n_keys = 1; sa.rows_per_kv = 1;
    sa.o = (bf16*)buf->o_tiles; sa.ldo = d; sa.R = R; sa.H = heads; sa.scale = scale; sa.q_acc = nullptr;
    sa.q_bias = nullptr; sa.o_tiled_kb = d / BK; sa.kv_tiles = nullptr; sa.kv_tiles_stride = 0;
    // 2-3: out_proj (split-K into acc), residual + encoder_attn_layer_norm
    gemm(p[2], PH_GEMM_RED, o_t, pk + 3ll * d * d, d, d, 4, nullptr, buf->acc, d);
    ln(p[3], buf->acc, w.sa_o_b, w.ln_ca_g, w.ln_ca_b, buf->a_tiles, d / BK);
    // 4-5: cross-attention query (split-K into q_acc; bias + rounding in the attention phase), attention over the
    //      image's visual K/V
    gemm(p[4], PH_GEMM_RED, a_t, pk + 4ll * d * d, d, d, 4, nullptr, buf->q_acc, d);
    p[5].kind = PH_ATTN;
    DecAttnArgs& ca = p[5].attn;
    ca.q = nullptr; ca.ldq = d; ca.kc = (const bf16*)w.enc_k; ca.vc = (const bf16*)w.enc_v; ca.ldkv = ld_enc;
    ca.anc = nullptr; ca.T = enc_tokens; ca.n_keys = enc_tokens; ca.rows_per_kv = rows_per_image;
    ca.o = (bf16*)buf->o_tiles; ca.ldo = d; ca.R = R; ca.H = heads; ca.scale = scale; ca.q_acc = buf->q_acc;
    ca.q_bias = w.ca_q_b; ca.o_tiled_kb = d / BK;
    if (buf->cross_kv_tiles) {           // [image][layer][head][16 KB]
      ca.kv_tiles = (const bf16*)buf->cross_kv_tiles + (long long)l * heads * 8192;
      ca.kv_tiles_stride = (long long)L * heads * 8192;
    }
    // 6-7: cross out_proj, residual + final_layer_norm
    gemm(p[6], PH_GEMM_RED, o_t, pk + 5ll * d * d, d, d, 4, nullptr, buf->acc, d);
    ln(p[7], buf->acc, w.ca_o_b, w.ln_f_g, w.ln_f_b, buf->a_tiles, d / BK);
    // 8-10: fc1 + activation (tile-image output), fc2 (split-K), residual + the next block's LayerNorm
    gemm(p[8], PH_GEMM_STORE, a_t, pk + 6ll * d * d, F, d, 1, w.fc1_b, buf->g_tiles, F);
    p[8].act = act;
    p[8].out_tiled_kb = F / BK;
    gemm(p[9], PH_GEMM_RED, g_t, pk + 6ll * d * d + (long long)d * F, d, F, 4, nullptr, buf->acc, d);
    if (l + 1 < L)
      ln(p[10], buf->acc, w.fc2_b, layers[l + 1].ln_sa_g, layers[l + 1].ln_sa_b, buf->a_tiles, d / BK);
    else
      ln(p[10], buf->acc, w.fc2_b, buf->ln_out_g, buf->ln_out_b, buf->h_out_tiles ? buf->h_out_tiles : buf->h_out,
         buf->h_out_tiles ? d / BK : 0);                                                // decoder layer_norm
  }
  MIC_CHECK_CUDA(cudaMemcpyAsync(plan_dev, host, total, cudaMemcpyHostToDevice, reinterpret_cast<cudaStream_t>(stream)));
  return MIC_OK;
}

extern "C" int mic_decoder_step(void* stream, const void* plan_dev, int num_layers, int R, int pos,
                                unsigned int* sync_counter, unsigned long long* phase_times, const int* active,
                                int opts) {
  MIC_CHECK_ARG(plan_dev && sync_counter && num_layers > 0 && R > 0 && pos >= 0, "decoder step: bad argument");
  static bool attr_set = false;
  if (!attr_set) {
    MIC_CHECK_CUDA(cudaFuncSetAttribute(decoder_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  StepArgs a;
  a.phases = reinterpret_cast<const Phase*>(plan_dev);
  a.num_phases = 1 + kPhasesPerLayer * num_layers;
  a.R = R;
  a.pos = pos;
  a.sync = sync_counter;
  a.prof = phase_times;
  a.active = active;
  a.opts = opts;
  // every CTA must be resident at once (grid barrier): one CTA per SM (214 KB of shared memory each), grid = #SMs
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(mic_num_sms());
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  // cooperative launch: the driver guarantees (or refuses) co-residency of the whole grid, which the grid barrier
  // needs; works under stream capture (kernel node attribute)
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  MIC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, decoder_step_kernel, a));
  return MIC_OK;
}

extern "C" int mic_barrier_bench(void* stream, unsigned int* sync_counter, int n, int variant) {
  barrier_bench_kernel<<<mic_num_sms(), THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(sync_counter, n, variant);
  MIC_CHECK_LAUNCH();
  return MIC_OK;
}
