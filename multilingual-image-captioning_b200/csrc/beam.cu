// Search-step kernels for generate(): candidate merge after the fused lm_head, the beam-search
// bookkeeping of `beam_search_body_fn` (generation_clip_vision_utils.py:822-966) replicated in fp32
// statement by statement, the greedy step (:489-522) and the loop condition (:798-820).
// The KV cache is never gathered: beams carry an ancestor table (attention.cu) instead of :945-953.
#include "common.cuh"

#include "../../include/mic_b200.h"

namespace {

constexpr int TOPK = 8;          // == micgemm::SEARCH_TOPK
constexpr int MAX_BEAMS = 8;     // candidates kept per image = 2 * beams <= MAX_CAND
constexpr int MAX_CAND = 16;     // per-row candidate list length: 8 (beams <= 4) or 16 (two search passes, beams 5..8)
#define NEG_BIG (-1.0e7f)

__device__ __forceinline__ bool better(float v, int i, float v2, int i2) {   // (value desc, index asc)
  return v > v2 || (v == v2 && i < i2);
}

// ---------------------------------------------------------------------------------------------
// Per row: combine the slab partials into log-sum-exp and the row's top-8 (log-prob, token).
// One warp per row; each lane folds its share of the (already sorted) slab lists into a private
// top-8, then 8 rounds of warp arg-max pop the global order.
// log_softmax mirrors jax.nn.log_softmax: (z - max) - log(sum exp(z - max)).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
search_merge_kernel(const float* __restrict__ pmax, const float* __restrict__ psum, const float* __restrict__ cval,
                    const int* __restrict__ cidx, int nparts, int R, float* __restrict__ row_lp,
                    int* __restrict__ row_tok, float* __restrict__ row_max_lse, int ld_out, int col_off, int lse_given,
                    float* __restrict__ last_val, int* __restrict__ last_idx) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= R) return;
  float mx = -INFINITY;
  for (int p = lane; p < nparts; p += 32) mx = fmaxf(mx, pmax[(long long)p * R + r]);
  mx = warp_max(mx);
  float sm = 0.f;
  float tv[TOPK];
  int ti[TOPK];
#pragma unroll
  for (int i = 0; i < TOPK; ++i) {
    tv[i] = -INFINITY;
    ti[i] = 0x7fffffff;
  }
  for (int p = lane; p < nparts; p += 32) {
    const long long o = (long long)p * R + r;
    const float m = pmax[o];
    if (m > -INFINITY) sm += psum[o] * __expf(m - mx);
    // the whole sorted list of this slab in four 16-byte loads (one round trip instead of up to eight)
    float lv[TOPK];
    int li[TOPK];
    {
      const float4 v0 = *reinterpret_cast<const float4*>(cval + o * TOPK), v1 = *reinterpret_cast<const float4*>(cval + o * TOPK + 4);
      const int4 i0 = *reinterpret_cast<const int4*>(cidx + o * TOPK), i1 = *reinterpret_cast<const int4*>(cidx + o * TOPK + 4);
      lv[0] = v0.x; lv[1] = v0.y; lv[2] = v0.z; lv[3] = v0.w; lv[4] = v1.x; lv[5] = v1.y; lv[6] = v1.z; lv[7] = v1.w;
      li[0] = i0.x; li[1] = i0.y; li[2] = i0.z; li[3] = i0.w; li[4] = i1.x; li[5] = i1.y; li[6] = i1.z; li[7] = i1.w;
    }
#pragma unroll
    for (int e = 0; e < TOPK; ++e) {
      float cv = lv[e];
      int ci = li[e];
      if (!better(cv, ci, tv[TOPK - 1], ti[TOPK - 1])) break;   // list is sorted: nothing further can enter
#pragma unroll
      for (int i = 0; i < TOPK; ++i) {
        if (better(cv, ci, tv[i], ti[i])) {
          const float a = tv[i];
          const int b = ti[i];
          tv[i] = cv;
          ti[i] = ci;
          cv = a;
          ci = b;
        }
      }
    }
  }
  sm = warp_sum(sm);
  float logsum = logf(sm);
  if (lse_given) {            // second search pass (ranks 9..16): the softmax statistics are those of the first pass
    mx = row_max_lse[2 * r];
    logsum = row_max_lse[2 * r + 1];
  } else if (lane == 0 && row_max_lse) {
    row_max_lse[2 * r] = mx;
    row_max_lse[2 * r + 1] = logsum;
  }
  // pop the 8 best across lanes
  for (int k = 0; k < TOPK; ++k) {
    float bv = tv[0];
    int bi = ti[0];
    int bl = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
      if (better(ov, oi, bv, bi) || (ov == bv && oi == bi && ol < bl)) {
        bv = ov;
        bi = oi;
        bl = ol;
      }
    }
    if (lane == 0) {
      row_lp[(long long)r * ld_out + col_off + k] = (bv - mx) - logsum;
      row_tok[(long long)r * ld_out + col_off + k] = bi;
      if (k == TOPK - 1 && last_val) {      // raw (logit, token) of the 8th best: upper bound of the next pass
        last_val[r] = bv;
        last_idx[r] = bi;
      }
    }
    if (lane == bl) {
#pragma unroll
      for (int i = 0; i < TOPK - 1; ++i) {
        tv[i] = tv[i + 1];
        ti[i] = ti[i + 1];
      }
      tv[TOPK - 1] = -INFINITY;
      ti[TOPK - 1] = 0x7fffffff;
    }
  }
}

// lane id of the best (value desc, key asc) among the lanes with valid == true (all lanes get the same answer);
// -1 if none.  -inf candidates are legal entries (ForcedBOS fillers), hence the separate validity flag.
__device__ __forceinline__ int warp_best(float v, int key, bool valid, int lane) {
  float bv = v;
  int bk = key, bl = valid ? lane : -1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
    const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
    if (ol >= 0 && (bl < 0 || better(ov, ok, bv, bk))) {
      bv = ov;
      bk = ok;
      bl = ol;
    }
  }
  return bl;
}

// ---------------------------------------------------------------------------------------------
// Beam step: one warp per image.  State arrays follow the reference's BeamSearchState.
// ---------------------------------------------------------------------------------------------
struct BeamArgs {
  const float* row_lp;     // [B*K, cpr] log-probs of each beam-row's best tokens (unused when forced >= 0)
  const int* row_tok;      // [B*K, cpr]
  int forced_token;        // >= 0: ForcedBOS/ForcedEOS step (scores: -inf everywhere, 0 at this id)
  int cpr;                 // candidates per row in row_lp / row_tok (8 or 16), >= 2K
  int B, K, L, V;
  int cur_len;
  int eos, early_stopping;
  float length_penalty;
  int* running_seq;        // [B, K, L]
  float* running_scores;   // [B, K]
  int* sequences;          // [B, K, L]
  float* scores;           // [B, K]
  int* finished;           // [B, K]
  int* ancestors;          // [B*K, L] cache-row table (attention.cu) or null
  int* next_token;         // [B*K]
  int* active;             // device flag of the while_loop condition; body is skipped when 0
  int* all_flags;          // [B] per-image cond terms, reduced by cond kernel
};

__global__ void __launch_bounds__(32) beam_step_kernel(const BeamArgs a) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ int sh[];
  if (*a.active == 0) return;
  const int b = blockIdx.x, lane = threadIdx.x;
  const int K = a.K, L = a.L, K2 = 2 * K;
  int* old_run = sh;                    // [K, L]
  int* old_seq = sh + K * L;            // [K, L]
  int* old_anc = sh + 2 * K * L;        // [K, L]
  __shared__ float s_tlp[MAX_CAND], s_tlp2[MAX_BEAMS + MAX_CAND], s_newscore[MAX_BEAMS], s_newrun[MAX_BEAMS];
  __shared__ int s_beam[MAX_CAND], s_tok[MAX_CAND], s_fin[MAX_CAND], s_runsel[MAX_BEAMS], s_mergesel[MAX_BEAMS], s_newfin[MAX_BEAMS];
  for (int i = lane; i < K * L; i += 32) {
    old_run[i] = a.running_seq[(long long)b * K * L + i];
    old_seq[i] = a.sequences[(long long)b * K * L + i];
    if (a.ancestors) old_anc[i] = a.ancestors[(long long)b * K * L + i];
  }
  __syncwarp();
  // ---- 3..7 run warp-parallel: one candidate per lane, selections by warp arg-max rounds (the statement order and
  // every fp32 operation of the reference are kept; only WHO evaluates them changed - a single-lane version with
  // local-memory candidate arrays cost 36 us per step)
  {
    // ---- 3. top 2K over the K*V candidates (value desc, flat index asc) ----
    // up to K * cpr = 128 candidates: 4 slots per lane
    constexpr int SLOTS = MAX_BEAMS * MAX_CAND / 32;
    float cv[SLOTS];
    int cflat[SLOTS];
    bool valid[SLOTS];
#pragma unroll
    for (int q = 0; q < SLOTS; ++q) {
      cv[q] = -INFINITY;
      cflat[q] = 0x7fffffff;
      valid[q] = false;
      const int i = lane + 32 * q;
      if (a.forced_token >= 0) {
        if (i < K) {                                      // the forced id of every beam: log-prob 0
          cv[q] = 0.0f + a.running_scores[b * K + i];
          cflat[q] = i * a.V + a.forced_token;
          valid[q] = true;
        } else if (i < 3 * K) {                           // -inf fillers: lowest flat indices other than the forced id
          int f = i - K;
          if (f >= a.forced_token) ++f;
          if (f < a.V) {
            cflat[q] = f;
            valid[q] = true;
          }
        }
      } else if (i < K * a.cpr) {
        const int k = i / a.cpr;
        cv[q] = a.row_lp[(long long)b * K * a.cpr + i] + a.running_scores[b * K + k];
        cflat[q] = k * a.V + a.row_tok[(long long)b * K * a.cpr + i];
        valid[q] = true;
      }
    }
    // lane j (< 2K) ends up holding candidate j of the ordered top-2K list
    float my_tlp = 0.f;
    int my_flat = 0;
    for (int j = 0; j < K2; ++j) {
      // the lane's own best slot, then the warp's best lane
      float lv = cv[0];
      int lf = cflat[0], ls = 0;
      bool lvalid = valid[0];
#pragma unroll
      for (int q = 1; q < SLOTS; ++q) {
        if (valid[q] && (!lvalid || better(cv[q], cflat[q], lv, lf))) {
          lv = cv[q];
          lf = cflat[q];
          ls = q;
          lvalid = true;
        }
      }
      const int w = warp_best(lv, lf, lvalid, lane);
      const float wv = __shfl_sync(0xffffffffu, lv, w);
      const int wf = __shfl_sync(0xffffffffu, lf, w);
      if (lane == w) {
#pragma unroll
        for (int q = 0; q < SLOTS; ++q)
          if (q == ls) valid[q] = false;
      }
      if (lane == j) {
        my_tlp = wv;
        my_flat = wf;
      }
    }
    const int my_beam = my_flat / a.V, my_tok = my_flat % a.V;
    // ---- 4. did_topk_just_finished ; topk_log_probs += finished * -1e7 ----
    const int my_fin = (my_tok == a.eos) ? 1 : 0;
    my_tlp = my_tlp + (my_fin ? NEG_BIG : -0.0f);
    if (lane < K2) {
      s_tlp[lane] = my_tlp;
      s_beam[lane] = my_beam;
      s_tok[lane] = my_tok;
      s_fin[lane] = my_fin;
    }
    // ---- 5. next running = top K of the 2K (stable: strict > keeps the lower index), flipped to ascending ----
    bool v5 = lane < K2;
    for (int j = 0; j < K; ++j) {
      const int w = warp_best(my_tlp, lane, v5, lane);
      if (lane == w) v5 = false;
      if (lane == 0) s_runsel[K - 1 - j] = w;
    }
    __syncwarp();
    if (lane < K) s_newrun[lane] = s_tlp[s_runsel[lane]];
    // ---- 6. length penalty on the (already penalised) array, then the second penalty ----
    bool all_fin = true;
    for (int k = 0; k < K; ++k) all_fin = all_fin && (a.finished[b * K + k] != 0);
    const bool beams_full = all_fin && a.early_stopping;
    const float denom = powf((float)a.cur_len, a.length_penalty);
    float my_tlp2 = my_tlp / denom;
    my_tlp2 = my_tlp2 + (((!my_fin) || beams_full) ? NEG_BIG : -0.0f);
    // ---- 7. merge with the finished set: top K of (K old + 2K new), stable, flipped ----
    const float shifted = __shfl_sync(0xffffffffu, my_tlp2, (lane - K) & 31);      // lane l >= K takes candidate l - K
    float mv = -INFINITY;
    bool v7 = lane < K + K2;
    if (lane < K) mv = a.scores[b * K + lane];
    else if (v7) mv = shifted;
    for (int j = 0; j < K; ++j) {
      const int w = warp_best(mv, lane, v7, lane);
      if (lane == w) v7 = false;
      if (lane == 0) s_mergesel[K - 1 - j] = w;
    }
    __syncwarp();
    if (lane < K + K2) s_tlp2[lane] = mv;              // merged value list (reusing s_tlp2 as [K + 2K] needs 12 slots)
    __syncwarp();
    if (lane < K) {
      const int m = s_mergesel[lane];
      s_newscore[lane] = s_tlp2[m];
      s_newfin[lane] = m < K ? a.finished[b * K + m] : s_fin[m - K];
    }
  }
  __syncwarp();
  // ---- write back (all lanes) ----
  for (int k = 0; k < K; ++k) {
    const int sel = s_runsel[k];           // index into the 2K candidates
    const int parent = s_beam[sel];        // old beam it extends
    for (int i = lane; i < L; i += 32) {
      int tokv = old_run[parent * L + i];
      if (i == a.cur_len) tokv = s_tok[sel];
      a.running_seq[((long long)b * K + k) * L + i] = tokv;
      if (a.ancestors) {
        // history rows of the parent; position cur_len-1 (just written by the parent row itself)
        int anc = old_anc[parent * L + i];
        if (i == a.cur_len - 1) anc = b * K + parent;
        else if (i >= a.cur_len) anc = b * K + k;     // future positions: the row's own slots
        a.ancestors[((long long)b * K + k) * L + i] = anc;
      }
    }
    const int m = s_mergesel[k];
    for (int i = lane; i < L; i += 32) {
      int tokv;
      if (m < K) {
        tokv = old_seq[m * L + i];
      } else {
        const int c = m - K;
        tokv = old_run[s_beam[c] * L + i];
        if (i == a.cur_len) tokv = s_tok[c];
      }
      a.sequences[((long long)b * K + k) * L + i] = tokv;
    }
    if (lane == 0) {
      a.running_scores[b * K + k] = s_newrun[k];
      a.scores[b * K + k] = s_newscore[k];
      a.finished[b * K + k] = s_newfin[k];
      a.next_token[b * K + k] = s_tok[sel];
    }
  }
}

// while_loop condition (:798-820) evaluated for the NEXT iteration (cur_len already incremented)
__global__ void __launch_bounds__(256)
beam_cond_kernel(const float* __restrict__ running_scores, const float* __restrict__ scores,
                 const int* __restrict__ finished, int B, int K, int cur_len, int max_length, float length_penalty,
                 int early_stopping, int* __restrict__ active) {
  pdl_trigger();
  pdl_wait();
  __shared__ int s_improve, s_allfin;
  if (threadIdx.x == 0) {
    s_improve = 1;
    s_allfin = 1;
  }
  __syncthreads();
  if (*active == 0) return;
  const float denom = powf((float)max_length, length_penalty);
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float best_running = running_scores[b * K + K - 1] / denom;
    float mn = INFINITY;
    for (int k = 0; k < K; ++k) mn = fminf(mn, scores[b * K + k]);
    for (int k = 0; k < K; ++k) {
      const float worst = finished[b * K + k] ? mn : NEG_BIG;
      if (!(worst < best_running)) atomicAnd(&s_improve, 0);
      if (!finished[b * K + k]) atomicAnd(&s_allfin, 0);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int not_max = cur_len < max_length;
    const int still_open = !(s_allfin && early_stopping);
    *active = (not_max && still_open && s_improve) ? 1 : 0;
  }
}

// epilogue (:978-990): finished set if any finished else running set; take the last (best) beam
__global__ void beam_finalize_kernel(const int* __restrict__ sequences, const float* __restrict__ scores,
                                     const int* __restrict__ finished, const int* __restrict__ running_seq,
                                     const float* __restrict__ running_scores, int B, int K, int L,
                                     int* __restrict__ out_seq, float* __restrict__ out_scores) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  int any = 0;
  for (int k = 0; k < K; ++k) any |= finished[b * K + k];
  const int* src = (any ? sequences : running_seq) + ((long long)b * K + K - 1) * L;
  for (int i = threadIdx.x; i < L; i += blockDim.x) out_seq[(long long)b * L + i] = src[i];
  if (threadIdx.x == 0) out_scores[b] = (any ? scores : running_scores)[b * K + K - 1];
}

// ---------------------------------------------------------------------------------------------
// Greedy step (:489-522).  Processors act on RAW logits; argmax = candidate 0 of the merged list.
// ---------------------------------------------------------------------------------------------
__global__ void greedy_step_kernel(const int* __restrict__ row_tok, int forced_token, int R, int L, int cur_len,
                                   int eos, int pad, int* __restrict__ sequences, int* __restrict__ finished,
                                   int* __restrict__ next_token, int* __restrict__ active) {
  pdl_trigger();
  pdl_wait();
  if (*active == 0) return;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  int tok = forced_token >= 0 ? forced_token : row_tok[(long long)r * TOPK];
  const int fin = finished[r] | (tok == eos);
  tok = fin ? pad : tok;
  sequences[(long long)r * L + cur_len] = tok;
  finished[r] = fin;
  next_token[r] = tok;
}

__global__ void greedy_cond_kernel(const int* __restrict__ finished, int R, int cur_len, int max_length,
                                   int* __restrict__ active) {
  pdl_trigger();
  pdl_wait();
  __shared__ int s_all;
  if (threadIdx.x == 0) s_all = 1;
  __syncthreads();
  if (*active == 0) return;
  for (int r = threadIdx.x; r < R; r += blockDim.x)
    if (!finished[r]) atomicAnd(&s_all, 0);
  __syncthreads();
  if (threadIdx.x == 0) *active = (cur_len == max_length || s_all) ? 0 : 1;
}

}  // namespace

#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" int mic_search_merge(void* stream, const float* pmax, const float* psum, const float* cand_val,
                                const int* cand_idx, int num_partials, int R, float* row_lp, int* row_tok,
                                float* row_max_logsum, int ld_out, int col_off, int lse_given, float* last_val,
                                int* last_idx) {
  MIC_CHECK_ARG(ld_out >= TOPK && col_off >= 0 && col_off + TOPK <= ld_out, "search_merge: ld_out=%d col_off=%d", ld_out,
                col_off);
  MIC_CHECK_ARG(!lse_given || row_max_logsum, "search_merge: lse_given needs row_max_logsum");
  MIC_CHECK_CUDA(mic_launch(search_merge_kernel, dim3((R + 3) / 4), dim3(128), 0, STREAM, pmax, psum, cand_val, cand_idx,
                            num_partials, R, row_lp, row_tok, row_max_logsum, ld_out, col_off, lse_given, last_val,
                            last_idx));
  return MIC_OK;
}

extern "C" int mic_beam_step(void* stream, const float* row_lp, const int* row_tok, int forced_token, int B, int K,
                             int L, int V, int cur_len, int eos_token_id, int early_stopping, float length_penalty,
                             int* running_seq, float* running_scores, int* sequences, float* scores, int* finished,
                             int* ancestors, int* next_token, int* active, int cand_per_row) {
  MIC_CHECK_ARG(K >= 1 && K <= MAX_BEAMS, "beam_step: num_beams=%d must be in [1,%d]", K, MAX_BEAMS);
  MIC_CHECK_ARG((cand_per_row == TOPK || cand_per_row == MAX_CAND) && cand_per_row >= 2 * K,
                "beam_step: cand_per_row=%d must be 8 or 16 and >= 2*num_beams", cand_per_row);
  MIC_CHECK_ARG(cur_len >= 1 && cur_len < L, "beam_step: cur_len=%d out of range for max_length=%d", cur_len, L);
  BeamArgs a;
  a.row_lp = row_lp; a.row_tok = row_tok; a.forced_token = forced_token; a.cpr = cand_per_row;
  a.B = B; a.K = K; a.L = L; a.V = V; a.cur_len = cur_len; a.eos = eos_token_id;
  a.early_stopping = early_stopping; a.length_penalty = length_penalty;
  a.running_seq = running_seq; a.running_scores = running_scores; a.sequences = sequences; a.scores = scores;
  a.finished = finished; a.ancestors = ancestors; a.next_token = next_token; a.active = active; a.all_flags = nullptr;
  const size_t smem = (size_t)3 * K * L * sizeof(int);
  MIC_CHECK_ARG(smem <= 40 * 1024, "beam_step: max_length %d too large", L);
  MIC_CHECK_CUDA(mic_launch(beam_step_kernel, dim3(B), dim3(32), smem, STREAM, a));
  return MIC_OK;
}

extern "C" int mic_beam_cond(void* stream, const float* running_scores, const float* scores, const int* finished,
                             int B, int K, int cur_len, int max_length, float length_penalty, int early_stopping,
                             int* active) {
  MIC_CHECK_CUDA(mic_launch(beam_cond_kernel, dim3(1), dim3(256), 0, STREAM, running_scores, scores, finished, B, K, cur_len, max_length, length_penalty,
                                          early_stopping, active));
  return MIC_OK;
}

extern "C" int mic_beam_finalize(void* stream, const int* sequences, const float* scores, const int* finished,
                                 const int* running_seq, const float* running_scores, int B, int K, int L,
                                 int* out_seq, float* out_scores) {
  MIC_CHECK_CUDA(mic_launch(beam_finalize_kernel, dim3(B), dim3(64), 0, STREAM, sequences, scores, finished, running_seq, running_scores, B, K, L,
                                             out_seq, out_scores));
  return MIC_OK;
}

extern "C" int mic_greedy_step(void* stream, const int* row_tok, int forced_token, int R, int L, int cur_len, int eos,
                               int pad, int* sequences, int* finished, int* next_token, int* active) {
  MIC_CHECK_CUDA(mic_launch(greedy_step_kernel, dim3((R + 127) / 128), dim3(128), 0, STREAM, row_tok, forced_token, R, L, cur_len, eos, pad, sequences,
                                                          finished, next_token, active));
  return MIC_OK;
}

extern "C" int mic_greedy_cond(void* stream, const int* finished, int R, int cur_len, int max_length, int* active) {
  MIC_CHECK_CUDA(mic_launch(greedy_cond_kernel, dim3(1), dim3(256), 0, STREAM, finished, R, cur_len, max_length, active));
  return MIC_OK;
}
