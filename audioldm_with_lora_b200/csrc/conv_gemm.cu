// Implicit-GEMM convolution / linear kernel for sm_100a.
//
//   out[pixel, n] = epilogue( sum_k A[pixel, k] * Wp[n, k] )
//
// * A is never materialised: a K-block is one TMA box [BN_img x BH x W x 64ch] of an NHWC bf16
//   activation, fetched at a (dh, dw) tap offset; out-of-bounds rows/columns are zero-filled by
//   TMA, which IS the conv padding.  A linear layer is the degenerate case W=1, H=M, 1 tap.
// * Up to three A sources are chained along K ("segments"): segment 0 carries the 1 or 9 taps;
//   segments 1/2 are 1x1 taps over other tensors.  This is how the ResNet 1x1 `conv_shortcut`
//   over cat([h, skip]) is accumulated into conv2's TMEM tile (K11: the concat is never written)
//   and how the rank-r LoRA branch [x | x.A^T] . [W | s.B]^T rides in the base GEMM (K1).
// * tcgen05.mma (UMMA 128 x block_n x 16, bf16 -> fp32 in TMEM), 128B-swizzled K-major smem
//   tiles, mbarrier ring, persistent CTAs, double-buffered TMEM accumulator so the epilogue of
//   tile i overlaps the mainloop of tile i+1.
// * Epilogue: 8 warps (two per TMEM lane quarter, alternating 64-column chunks), one accumulator
//   row per thread: + bias[n] + per-image row vector (timestep/class embedding projection, K6)
//   + residual, optional GEGLU (K3).  bf16 results are staged in a per-warp 128B-swizzled smem
//   tile and written with TMA stores (coalesced, asynchronous, tails clipped by the tensor map);
//   fp32 / stride-2 outputs use direct 16-byte stores.
// * 2-CTA mode (cta_group::2, clusters of two CTAs on neighbouring SMs): the pair computes one
//   256 x block_n tile; each CTA stages its own 128 rows of A and HALF of the weight tile, the leader
//   issues tcgen05.mma.cta_group::2 reading both CTAs' shared memory.  Halves the weight traffic per
//   CTA and the smem per stage (6 stages instead of 4 at block_n = 256): the kernel is bound by
//   operand bytes in flight, not by the tensor pipe.
// * Programmatic dependent launch: the prologue (barrier init, TMEM alloc, descriptor prefetch)
//   overlaps the tail of the previous kernel in the stream / CUDA graph.
//
// Replaces (reference call sites -> diffusers/torch): F.conv2d / F.linear under
// UNet2DConditionModel.forward, /root/reference/script/train/train_audioldm_lora.py:539-546.
#include "host_util.h"
#include "ptx.cuh"

// Build-time variants (A/B-tested on B200; defaults are the faster ones):
#ifndef B200_GEMM_PROFILE
#define B200_GEMM_PROFILE 0      // 1: CTA-0 cycle timeline + MMA-thread cost breakdown (B200_GEMM_DEBUG & 4 / & 8)
#endif
#ifndef B200_GEMM_LB_THREADS
// The register budget is set through the thread count promised to ptxas (__maxnreg__ cannot be combined with
// __launch_bounds__): 352 (the real block size) -> 168 registers per thread = the whole register file for one CTA per SM;
// 512 -> 128 registers = 44 K of the SM's 64 K, which leaves room for the CTAs of a neighbouring norm / attention kernel
// in the stream, so programmatic dependent launch can actually overlap prologues and tails.
#define B200_GEMM_LB_THREADS 352
#endif
#ifndef B200_B_EARLY
#define B200_B_EARLY 1           // 1: the weight producer does not wait for the previous grid (see the PDL note in the kernel)
#endif

namespace b200 {

static constexpr int kBlockM = 128;
static constexpr int kBlockK = 64;              // 64 bf16 = one 128B swizzle row
static constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KB
static constexpr int kEpiWarps = 8;
static constexpr int kProducerBWarp = 2 + kEpiWarps;    // the weight (B) TMA stream has its own warp
static constexpr int kThreads = 64 + kEpiWarps * 32 + 32;   // warp0 TMA (A), warp1 MMA, warps2-9 epilogue, warp10 TMA (B)
static constexpr int kEpiStageBytes = 32 * 128;         // per-warp staging: 32 rows x 64 bf16
static constexpr int kEpiVecBytes = 256 * 4;            // per-warp scratch: (bias + per-image row vector) of the warp's output columns
static constexpr int kBarrierBytes = 512;
static constexpr int kSmemLimit = 227 * 1024;

struct ConvGemmParams {
  int seg_end0, seg_end1, num_kb;   // k-block boundaries of the segments
  int cb0;                          // 64-channel blocks per tap in segment 0
  int ntaps;                        // 1 or 9
  int H, W, NB;                     // activation geometry
  int BH, BNI;                      // TMA box rows / images (W * BH * BNI == 128)
  int tiles_h, num_m_tiles, num_n_tiles;
  int block_n, stages;
  int n_valid;
  const float* bias;
  const float* rowvec;
  int rowvec_ld;
  const __nv_bfloat16* residual;
  int res_ld;
  void* out;
  int out_ld;
  int out_fp32;
  int geglu;
  int tmem_cols;
  int a_stride;                     // 1, or 2 (Downsample2D: k3 s2 p1): the activation box is read with TMA element strides (2, 2),
                                    //   the tile's 128 rows are OUTPUT pixels and H / W / tiles_h describe the output
  int tma_out;                      // 1: bf16 stride-1 output through smem staging + TMA store
  int num_m_groups;                 // m-tiles (1-CTA) or pairs of m-tiles (2-CTA) the tile index runs over
  int ksplit;                       // > 1: split-K; work item = (split, m_tile, n_tile), fp32 partials to `out`
  int kb_per_split;
  size_t split_stride;              // elements between the partial-sum planes
  int a_rank2;                      // linear layers (W = 1, one image): A tensor maps are plain 2-D [rows, K]
  int debug;                        // profiling knobs (env B200_GEMM_DEBUG, -DB200_GEMM_PROFILE=1 builds only): 2 = no MMAs
                                    // (results are garbage), 4 = CTA-0 timeline, 8 = per-k-block time stamps
  // Fused LoRA down-projection (linear layers, 1-CTA mode): phase 0 of every tile computes T = x . A^T (N = lora_n) on the
  // tensor core into TMEM columns [2 block_n, 2 block_n + lora_n); the epilogue warps turn it into a bf16 K-major
  // shared-memory tile while the base k-blocks run; the LoRA k-block (segment 1) then takes that tile as its A operand.
  int lora_n;                       // 0: off; else 16 / 32 / 48 / 64
  int lora_kb;                      // k-blocks of phase 0 (= c0 / 64)
  int m_rows;                       // rows of the token matrix (bounds the optional global copy of T)
  __nv_bfloat16* t_out;             // optional [m_rows, 64] copy of T (the fine-tuning step keeps it for dB)
  // Fused LayerNorm (linear layers): the GEMM runs on the RAW rows x with gamma folded into the weights; the epilogue
  // finishes the normalisation per row:  LN(x) W^T = rstd (x (gamma o W)^T - mu g) + b',  g[n] = sum_c gamma_c W[n,c],
  // b'[n] = sum_c beta_c W[n,c] (+ bias).  (mu, rstd) of the tile's 128 rows are computed by the epilogue warps from
  // global memory while the main loop runs.  `bias` carries b'; ln_ga / ln_ba are g / b' of the LoRA down-projection.
  const float* ln_stats;            // [m_rows, ln_chunks, 2]: per 64-channel chunk (sum, sum of squares) of each row of x,
  int ln_chunks;                    //   written by the epilogue of the kernel that produced x (stat_out below)
  int ln_c;
  float ln_eps;
  const float* ln_g;                // [n_pad], packed (tile-interleaved for GEGLU) like bias
  const float* ln_ga;               // [64]
  const float* ln_ba;               // [64]
  // Producer side of the same fusion: the TMA-store epilogue also writes, per row and 64-column output chunk, the
  // (sum, sum of squares) of the bf16-rounded values it stores -- fixed slots, no atomics, deterministic.
  float* stat_out;                  // [m_total, stat_chunks, 2] or null
  int stat_chunks;
  // Producer side of the one-pass GroupNorm (b200_groupnorm_apply): the TMA-store epilogue also leaves, per image, per
  // 32-pixel slab and per 4-channel unit, the (sum, sum of squares) of the bf16 values it stores -- fixed slots, no
  // atomics, deterministic.  4 channels nest in every group width of the UNet, so the consumer can regroup them for
  // any GroupNorm, also one over a concatenation of two tensors.
  float* gn_stat;                   // [NB, gn_slabs, n_valid / 4, 2] or null
  int gn_slabs;                     // slabs per image = tiles_h * (W * BH / 32)
  // Tap walk of segment 0: tap t reads the input at (dh, dw) = (dh0 + (t / tw) * dh_step, dw0 + t % tw), tw = dw_end - dw0.
  // 3x3 conv: dh0 = dw0 = -1, dw_end = 2, dh_step = 1.  1-D conv over [nb, L, C] (W = 1; b200_conv1d): dw0 = 0, dw_end = 1,
  // dh0 = -(k - 1) / 2 * dilation, dh_step = dilation; one phase of a transposed 1-D conv: dh0 = b, dh_step = -1.
  int dh0, dw0, dw_end, dh_step;
  // kAct (b200_conv1d): out = act(acc + bias + unact(residual)).  act: LeakyReLU max(v, act_slope v) (slope 1 = identity) or,
  // act_tanh = 1, tanh(v), or, act_tanh = 2, the exact-erf GELU.  The residual tensor may itself be stored POST-activation (y = lrelu(x, s)): x = min(y, y / s) is
  // recovered on the fly with res_neg_gain = 1 / s (1 = the residual is stored as it is).
  float act_slope;
  float res_neg_gain;
  int act_tanh;
  // Epilogue inputs staged instead of fetched per thread (each costs a launch 2-8 us when read with per-row LDGs after the
  // accumulator is complete: gpurun_out/r02_res_ab.log): the residual tile comes by TMA into the warp's output staging
  // buffers (two per warp then: epi_bufs) while the main loop still runs.
  int res_tma;                      // 1: residual through tmRes
  int epi_bufs;                     // staging buffers per epilogue warp (1, or 2 with res_tma)
  int b_dynamic;                    // 1: the "weight" operand is an activation written by the previous kernel in the
                                    // stream (b200_gemm_nt): its producer warp must wait for that grid like everyone else
};

// exact-erf GELU to ~7e-7 absolute (bf16 output rounding is 4e-3 relative), ONE MUFU op per element.  With s = |g|:
//   gelu(g) = g/2 (1 + erf(g/sqrt2)) = relu(g) - s/2 erfc(s/sqrt2),
//   erfc(x) = (1 + a1 x + ... + a6 x^6)^-16 + eps, |eps| <= 3e-7          (Abramowitz-Stegun 7.1.28)
// the 1/sqrt2 is folded into the coefficients.  The form it replaced (A-S 7.1.26: a reciprocal AND an exponential, plus
// the non-ftz range fix-ups of __fdividef / __expf: ~25 instructions, 2 MUFU) made the GEGLU epilogue XU-pipe-bound
// (ncu: xu 64 % of peak, tensor pipe 4 %: profiles/r02_prof_geglu_l1.md); this one is 13 FP instructions + 1 MUFU.
__device__ __forceinline__ float gelu_erf(float g) {
  const float s = fabsf(g);
  constexpr float c1 = 0.0705230784f * 0.70710678118654752f;
  constexpr float c2 = 0.0422820123f * 0.5f;
  constexpr float c3 = 0.0092705272f * 0.35355339059327376f;
  constexpr float c4 = 0.0001520143f * 0.25f;
  constexpr float c5 = 0.0002765672f * 0.17677669529663688f;
  constexpr float c6 = 0.0000430638f * 0.125f;
  float p = fmaf(c6, s, c5);
  p = fmaf(p, s, c4);
  p = fmaf(p, s, c3);
  p = fmaf(p, s, c2);
  p = fmaf(p, s, c1);
  p = fmaf(p, s, 1.0f);
  p *= p; p *= p; p *= p; p *= p;                                // ^16 (inf for |g| > ~60: rcp -> 0, erfc -> 0)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p));
  return fmaf(-0.5f * s, r, fmaxf(g, 0.f));
}

// debug timeline (B200_GEMM_DEBUG & 4): SM cycle counter of CTA 0 at fixed points of the kernel
__device__ unsigned long long g_timeline[160];
#if B200_GEMM_PROFILE
#define TL(i)                                                        \
  do {                                                               \
    if ((p.debug & 4) && blockIdx.x == 0 && (threadIdx.x & 31) == 0) g_timeline[i] = clock64(); \
  } while (0)
// per-k-block time stamps of CTA 0's first tile (B200_GEMM_DEBUG & 8): slot base + k-block
#define PKB(base, k)                                                                               \
  do {                                                                                             \
    if ((p.debug & 8) && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && t == tile0 && (k) < 32) g_timeline[(base) + (k)] = clock64(); \
  } while (0)
#else
#define TL(i) do {} while (0)
#define PKB(base, k) do {} while (0)
#endif

__device__ __forceinline__ void add8(float (&v)[8], const float* src) {
  const float4 b0 = *reinterpret_cast<const float4*>(src);
  const float4 b1 = *reinterpret_cast<const float4*>(src + 4);
  v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
  v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
}

__device__ __forceinline__ void ln_affine8(float (&v)[8], const float* g, const float* b, float mu, float rs) {
  const float4 g0 = *reinterpret_cast<const float4*>(g), g1 = *reinterpret_cast<const float4*>(g + 4);
  const float4 b0 = *reinterpret_cast<const float4*>(b), b1 = *reinterpret_cast<const float4*>(b + 4);
  v[0] = fmaf(v[0] - mu * g0.x, rs, b0.x); v[1] = fmaf(v[1] - mu * g0.y, rs, b0.y);
  v[2] = fmaf(v[2] - mu * g0.z, rs, b0.z); v[3] = fmaf(v[3] - mu * g0.w, rs, b0.w);
  v[4] = fmaf(v[4] - mu * g1.x, rs, b1.x); v[5] = fmaf(v[5] - mu * g1.y, rs, b1.y);
  v[6] = fmaf(v[6] - mu * g1.z, rs, b1.z); v[7] = fmaf(v[7] - mu * g1.w, rs, b1.w);
}

// b200_conv1d epilogue pieces (HiFi-GAN): residual stored post-LeakyReLU -> pre-activation value, and the output activation
__device__ __forceinline__ void add_res8_unact(float (&v)[8], const uint4 rr, float neg_gain) {
  const float r[8] = {bf16_lo(rr.x), bf16_hi(rr.x), bf16_lo(rr.y), bf16_hi(rr.y),
                      bf16_lo(rr.z), bf16_hi(rr.z), bf16_lo(rr.w), bf16_hi(rr.w)};
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] += fminf(r[j], r[j] * neg_gain);      // neg_gain >= 1: only negative values are scaled
}
__device__ __forceinline__ void act8(float (&v)[8], float slope, int use_tanh) {
  if (use_tanh == 2) {                                                   // exact-erf GELU (b200_conv1d act_tanh = 2)
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
  } else if (use_tanh) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float e = __expf(2.0f * fminf(fmaxf(v[j], -15.f), 15.f));     // tanh(x) = 1 - 2 / (e^{2x} + 1)
      v[j] = 1.0f - __fdividef(2.0f, e + 1.0f);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], v[j] * slope);        // slope <= 1
  }
}

template <bool kCta2, bool kLora = false, bool kLn = false, bool kStat = false, bool kGn = false, bool kAct = false>
__global__ void __launch_bounds__(B200_GEMM_LB_THREADS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmLA,
                 const __grid_constant__ CUtensorMap tmRes, const ConvGemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024B alignment for the 128B swizzle atoms.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t cta_rank = kCta2 ? cluster_ctarank() : 0u;       // 0 = leader of the pair
  const int b_rows = kCta2 ? p.block_n / 2 : p.block_n;           // weight rows this CTA stages per k-block
  const int stage_bytes = kABytes + b_rows * kBlockK * 2;
  uint8_t* t_tile = smem + p.stages * stage_bytes;                 // [128 rows][64 bf16], 128B-swizzled (fused LoRA only)
  uint8_t* epi_smem = t_tile + (kLora ? kABytes : 0);
  float* epi_vec = reinterpret_cast<float*>(epi_smem + kEpiWarps * p.epi_bufs * kEpiStageBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + kEpiWarps * (p.epi_bufs * kEpiStageBytes + kEpiVecBytes));
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tfull_bar = empty_bar + p.stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* lt_tmem_bar = tempty_bar + 2;                          // phase-0 accumulator complete (MMA -> epilogue warps)
  uint64_t* lt_smem_bar = lt_tmem_bar + 1;                         // T tile written (4 epilogue warps -> MMA)
  uint64_t* res_bar = lt_smem_bar + 1;                             // [kEpiWarps][2]: residual tile landed in the warp's staging buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 2 * kEpiWarps);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) TL(0);
  const int num_tiles = p.num_m_groups * p.num_n_tiles * p.ksplit;  // work items of a CTA (1-CTA) / a pair (2-CTA)
  const int tile0 = kCta2 ? blockIdx.x >> 1 : blockIdx.x;
  const int tile_step = kCta2 ? gridDim.x >> 1 : gridDim.x;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB);
    if (p.tma_out) tma_prefetch_desc(&tmOut);
    if (p.res_tma) tma_prefetch_desc(&tmRes);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 2);    // one arrive.expect_tx per producer thread
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], kCta2 ? 2 * kEpiWarps : kEpiWarps);   // one arrival per epilogue warp (of both CTAs)
    }
    mbar_init(lt_tmem_bar, 1);
    mbar_init(lt_smem_bar, 4);
    for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(&res_bar[i], 1);
    if (kLora) tma_prefetch_desc(&tmLA);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (kCta2) {
      tmem_alloc2(tmem_slot, p.tmem_cols);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_slot, p.tmem_cols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (kCta2) cluster_sync_all();      // the peer's barriers are initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) TL(1);
  // PDL: everything above overlapped the previous kernel's tail; from here on we touch its outputs -- except the
  // weight (B) producer: weights are never written inside a step, so with split producers its TMA stream starts
  // before the previous grid has drained (hides the HBM latency of the first weight tiles; B200_B_EARLY=0 disables).
  pdl_launch_dependents();
  if (warp != 0 && !(B200_B_EARLY && warp == kProducerBWarp && !p.b_dynamic)) pdl_wait();   // warp 0 (A producer): after its index math

  if (warp == 0) {
    // ================================================================ activation (A) TMA producer
    // One warp's instruction stream is serial: every instruction costs its full latency, and at 128-wide tiles the
    // tensor pipe wants a k-block every 256 cycles.  So the loop body is kept minimal: the whole warp runs it with
    // warp-uniform control flow (coordinates and addresses in uniform registers, one elected lane issues: no
    // ELECT + R2UR.BROADCAST waterfall per TMA), 32-bit shared-window addresses advanced incrementally, the (segment,
    // tap, channel-block) walk as nested loops instead of a per-k-block state machine, the watchdog out of line.
    // (The lane is elected at every use: a loop-invariant predicate gets the loop unswitched into a one-lane copy,
    //  i.e. divergent code, and the waterfalls come back.)
    {
      const int stages = p.stages, cb0 = p.cb0;
      const int seg_end0 = p.seg_end0, seg_end1 = p.seg_end1;
      const uint32_t stage_b = stage_bytes;
      const uint32_t smem_a = smem_u32(smem), full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar);
      const uint32_t ring_end = smem_a + stages * stage_b;
      const uint32_t a_bytes = static_cast<uint32_t>(kABytes) * (kCta2 ? 2u : 1u);   // 2-CTA: the leader's barrier expects the pair's bytes
      const bool expect = !kCta2 || cta_rank == 0;
      uint32_t slot = smem_a, bar_off = 0, ph = 0;
      uint32_t d, fb;
      // next ring slot: wait until the MMAs that read it have retired (a fresh barrier passes at once)
      auto acquire = [&]() {
        mbar_wait(empty_a + bar_off, ph ^ 1);
        d = slot;
        fb = full_a + bar_off;
        slot += stage_b;
        bar_off += 8;
        if (slot == ring_end) { slot = smem_a; bar_off = 0; ph ^= 1; }
      };
      for (int t = tile0; t < num_tiles; t += tile_step) {
        // tile -> (n_tile, m_group, split); integer divisions cost ~150 cycles each in this lone thread: skip the trivial ones
        const int tn = p.num_n_tiles == 1 ? t : t / p.num_n_tiles;
        const int split = p.ksplit == 1 ? 0 : tn / p.num_m_groups;
        const int m_group = tn - split * p.num_m_groups;
        const int m_tile = kCta2 ? 2 * m_group + static_cast<int>(cta_rank) : m_group;   // may be a phantom tile: all OOB
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(p.num_kb, kb_begin + p.kb_per_split);
        const int n_img = p.tiles_h == 1 ? m_tile : m_tile / p.tiles_h;
        const int h0 = (m_tile - n_img * p.tiles_h) * p.BH;
        const int n0 = n_img * p.BNI;
        int kb = kb_begin;
        int cb = 0, dh = p.dh0, dw = p.dw0;
        const int dw0 = p.dw0, dw_end = p.dw_end, dh_step = p.dh_step;
        if (kb > 0 && kb < seg_end0) {
          const int tap = kb / cb0;
          cb = kb - tap * cb0;
          const int tw = dw_end - dw0;
          dh += (tap / tw) * dh_step;
          dw += tap % tw;
        }
        // PDL: the activations are the previous kernel's output (the index math above overlapped its tail)
        if (t == tile0) { pdl_wait(); TL(2); }
        if (kLora) {
          // ---- phase 0: the tile's activations once more, against the stacked lora_A rows (c0 / 64 k-blocks)
          for (int k0 = 0; k0 < p.lora_kb; ++k0) {
            acquire();
            if (elect_one()) {
              mbar_expect_tx(fb, static_cast<uint32_t>(kABytes));
              tma_load_4d(d, &tmA0, fb, k0 * kBlockK, 0, h0, n0);
            }
          }
        }
        // ---- segment 0: taps x channel blocks of the main input (three literal tensor-map operands: selecting the
        //      map through a pointer variable is slower)
        const int end0 = min(kb_end, seg_end0);
        {
          int left = cb0 - cb;              // k-blocks left in the current tap
          int c = cb * kBlockK, hh = h0 * p.a_stride + dh;
#pragma unroll 1
          for (; kb < end0; ++kb) {
            acquire();
            if (elect_one()) {
              if (expect) mbar_expect_tx(fb, a_bytes);
              if (kCta2) tma_load_4d_cta2(d, &tmA0, fb, c, dw, hh, n0);
              else tma_load_4d(d, &tmA0, fb, c, dw, hh, n0);
            }
            PKB(64, kb - kb_begin);
            c += kBlockK;
            if (--left == 0) {              // next tap: (dh, dw) walk the window row by row
              left = cb0;
              c = 0;
              if (++dw == dw_end) { dw = dw0; hh += dh_step; }
            }
          }
        }
        // ---- segment 1: the second input of a concatenation (or, fused LoRA: the T tile, already in shared memory)
        const int end1 = min(kb_end, seg_end1);
#pragma unroll 1
        for (int c = (kb - seg_end0) * kBlockK; kb < end1; ++kb, c += kBlockK) {
          acquire();
          if (!elect_one()) continue;
          if (kLora) {
            mbar_expect_tx(fb, 0u);
          } else {
            if (expect) mbar_expect_tx(fb, a_bytes);
            if (kCta2) tma_load_4d_cta2(d, &tmA1, fb, c, 0, h0, n0);
            else tma_load_4d(d, &tmA1, fb, c, 0, h0, n0);
          }
        }
        // ---- segment 2: the third input
#pragma unroll 1
        for (int c = (kb - seg_end1) * kBlockK; kb < kb_end; ++kb, c += kBlockK) {
          acquire();
          if (elect_one()) {
            if (expect) mbar_expect_tx(fb, a_bytes);
            if (kCta2) tma_load_4d_cta2(d, &tmA2, fb, c, 0, h0, n0);
            else tma_load_4d(d, &tmA2, fb, c, 0, h0, n0);
          }
        }
      }
      TL(5);
    }
  } else if (warp == kProducerBWarp) {
    // ================================================================ weight (B) TMA producer (same structure)
    {
      const int stages = p.stages;
      const uint32_t stage_b = stage_bytes;
      const uint32_t smem_b = smem_u32(smem) + kABytes, full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar);
      const uint32_t ring_end = smem_b + stages * stage_b;
      const uint32_t b_bytes = static_cast<uint32_t>(stage_bytes - kABytes) * (kCta2 ? 2u : 1u);
      const bool expect = !kCta2 || cta_rank == 0;
      uint32_t slot = smem_b, bar_off = 0, ph = 0;
      uint32_t d, fb;
      auto acquire = [&]() {
        mbar_wait(empty_a + bar_off, ph ^ 1);
        d = slot;
        fb = full_a + bar_off;
        slot += stage_b;
        bar_off += 8;
        if (slot == ring_end) { slot = smem_b; bar_off = 0; ph ^= 1; }
      };
      for (int t = tile0; t < num_tiles; t += tile_step) {
        const int tn = p.num_n_tiles == 1 ? t : t / p.num_n_tiles;
        const int n_tile = t - tn * p.num_n_tiles;
        const int split = p.ksplit == 1 ? 0 : tn / p.num_m_groups;
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(p.num_kb, kb_begin + p.kb_per_split);
        const int b_row0 = n_tile * p.block_n + static_cast<int>(cta_rank) * b_rows;
        if (kLora) {
          const uint32_t bytes0 = static_cast<uint32_t>(p.lora_n) * kBlockK * 2;
          for (int k0 = 0; k0 < p.lora_kb; ++k0) {
            acquire();
            if (elect_one()) {
              mbar_expect_tx(fb, bytes0);
              tma_load_2d(d, &tmLA, fb, k0 * kBlockK, 0);
            }
          }
        }
        int c = kb_begin * kBlockK;
#pragma unroll 1
        for (int kb = kb_begin; kb < kb_end; ++kb, c += kBlockK) {
          acquire();
          if (elect_one()) {
            if (expect) mbar_expect_tx(fb, b_bytes);
            if (kCta2) tma_load_2d_cta2(d, &tmB, fb, c, b_row0);
            else tma_load_2d(d, &tmB, fb, c, b_row0);
          }
          PKB(96, kb - kb_begin);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (one thread, leader CTA only)
    // The WHOLE warp runs the loop and waits on the barriers; one elected lane issues.  With warp-uniform control
    // flow the descriptors, TMEM and barrier addresses live in uniform registers, so a tcgen05.mma is one instruction
    // instead of an ELECT + five R2UR.BROADCAST waterfall in front of each (issuer work per k-block: 383 -> ~260 cycles).
    if (cta_rank == 0) {
      const bool lead = elect_one();
      const uint32_t idesc = make_idesc_bf16(kCta2 ? 2 * kBlockM : kBlockM, p.block_n, 0, 0);
      // K-major, 128B swizzle: 8-row atom = 1024 B -> SBO = 1024; LBO unused (1).  Stage s adds s * stage_bytes.
      const uint64_t a_desc0 = make_smem_desc(smem_u32(smem), 16, 1024, SWZ_128B);
      const uint64_t b_desc0 = make_smem_desc(smem_u32(smem) + kABytes, 16, 1024, SWZ_128B);
      const uint64_t desc_step = static_cast<uint64_t>(stage_bytes >> 4);
      uint64_t a_desc = a_desc0, b_desc = b_desc0;
      // (like the producers: ring state in registers, 32-bit barrier addresses, nothing re-read per k-block)
      const uint32_t full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar);
      const uint32_t bar_end = static_cast<uint32_t>(p.stages) * 8u;
      uint32_t bar_off = 0, ph = 0;
      auto advance = [&]() {
        a_desc += desc_step;
        b_desc += desc_step;
        bar_off += 8;
        if (bar_off == bar_end) { bar_off = 0; ph ^= 1; a_desc = a_desc0; b_desc = b_desc0; }
      };
      int it = 0;
      for (int t = tile0; t < num_tiles; t += tile_step, ++it) {
        const int buf = it & 1;
        const uint32_t use = it >> 1;
        if (kLora) {
          // ---- phase 0: T = x . A^T into its own TMEM columns (no dependence on the accumulator buffers)
          const uint32_t idesc_t = make_idesc_bf16(kBlockM, p.lora_n, 0, 0);
          const uint32_t t_tmem = tmem_base + 2 * p.block_n;
          for (int kb = 0; kb < p.lora_kb; ++kb) {
            mbar_wait(full_a + bar_off, ph);
            tc_fence_after();
            if (lead) {
              umma_bf16_ss(t_tmem, a_desc, b_desc, idesc_t, kb != 0);
              umma_bf16_ss(t_tmem, a_desc + 2, b_desc + 2, idesc_t, 1);
              umma_bf16_ss(t_tmem, a_desc + 4, b_desc + 4, idesc_t, 1);
              umma_bf16_ss(t_tmem, a_desc + 6, b_desc + 6, idesc_t, 1);
              umma_commit(empty_a + bar_off);
            }
            advance();
          }
          if (lead) umma_commit(lt_tmem_bar);
        }
        mbar_wait(&tempty_bar[buf], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * p.block_n;
        const int split = p.ksplit == 1 ? 0 : t / (p.num_n_tiles * p.num_m_groups);
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(p.num_kb, kb_begin + p.kb_per_split);
        const int kb_lora = kLora ? p.seg_end0 : -1;
        if (it == 0) TL(6);
        [[maybe_unused]] long long acc_wait = 0, acc_mma = 0, tq = 0;
#if B200_GEMM_PROFILE
        const bool prof = (p.debug & 8) && blockIdx.x == 0 && lead;
        const bool no_mma = p.debug & 2;
#else
        constexpr bool prof = false, no_mma = false;
#endif
        // (a non-blocking look-ahead probe of the next slot's barrier before issuing was measured: slower -- the
        //  loop is bound by the operand stream, the next slot is never ready at probe time)
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          if (prof) tq = clock64();
          mbar_wait(full_a + bar_off, ph);
          tc_fence_after();
          if (prof) { const long long now = clock64(); acc_wait += now - tq; tq = now; }
          PKB(32, kb - kb_begin);
          const uint32_t acc0 = kb != kb_begin;
          if (kLora && kb == kb_lora) {
            // the LoRA k-block: A = T (written by the epilogue warps during the base k-blocks), B = s.B from the ring
            mbar_wait(lt_smem_bar, it & 1);
            tc_fence_after();
            const uint64_t t_desc = make_smem_desc(smem_u32(t_tile), 16, 1024, SWZ_128B);
            if (lead) {
              umma_bf16_ss(d_tmem, t_desc, b_desc, idesc, acc0);
              umma_bf16_ss(d_tmem, t_desc + 2, b_desc + 2, idesc, 1);
              umma_bf16_ss(d_tmem, t_desc + 4, b_desc + 4, idesc, 1);
              umma_bf16_ss(d_tmem, t_desc + 6, b_desc + 6, idesc, 1);
              umma_commit(empty_a + bar_off);
            }
          } else if (kCta2) {
            // advance 16 bf16 = 32 B along K inside the swizzle atom: +2 in (addr >> 4) units
            if (lead) {
              umma_bf16_ss_cta2(d_tmem, a_desc, b_desc, idesc, acc0);
              umma_bf16_ss_cta2(d_tmem, a_desc + 2, b_desc + 2, idesc, 1);
              umma_bf16_ss_cta2(d_tmem, a_desc + 4, b_desc + 4, idesc, 1);
              umma_bf16_ss_cta2(d_tmem, a_desc + 6, b_desc + 6, idesc, 1);
              umma_commit_cta2(empty_a + bar_off, 3);          // frees the stage in both CTAs
            }
          } else if (lead) {
            if (!no_mma) {
              umma_bf16_ss(d_tmem, a_desc, b_desc, idesc, acc0);
              umma_bf16_ss(d_tmem, a_desc + 2, b_desc + 2, idesc, 1);
              umma_bf16_ss(d_tmem, a_desc + 4, b_desc + 4, idesc, 1);
              umma_bf16_ss(d_tmem, a_desc + 6, b_desc + 6, idesc, 1);
            }
            umma_commit(empty_a + bar_off);
          }
          advance();
          if (prof) { const long long now = clock64(); acc_mma += now - tq; tq = now; }
          PKB(128, kb - kb_begin);
        }
        if (prof && lead && it == 0) {
          g_timeline[25] = acc_wait; g_timeline[28] = acc_mma; g_timeline[31] = kb_end - kb_begin;
        }
        if (lead) {
          if (kCta2) umma_commit_cta2(&tfull_bar[buf], 3);     // both CTAs' epilogues own 128 rows each
          else umma_commit(&tfull_bar[buf]);
        }
        if (it == 0) TL(10);
      }
    }
  } else {
    // ================================================================ epilogue warps
    // (warp index through a lane-0 broadcast: the compiler then KNOWS that everything derived from it -- TMA coordinates,
    //  staging addresses, chunk loops -- is warp-uniform, and an elected lane can issue the epilogue's TMA loads / stores
    //  without an ELECT + R2UR.BROADCAST + BRA.U.ANY waterfall in front of each)
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const int q = warp_u & 3;               // TMEM lane quarter this warp may access
    const int hf = (warp_u - 2) >> 2;       // which of the two warps of the quarter: takes chunks cc % 2 == hf
    const int row = q * 32 + lane;          // accumulator row == tile pixel
    const int wl = row % p.W;
    const int hl = (row / p.W) % p.BH;
    const int nl = row / (p.W * p.BH);
    // TMA-store box of this warp's 32 rows inside the tile (w fastest, then h, then image)
    const int q_w = (q * 32) % p.W;
    const int q_h = ((q * 32) / p.W) % p.BH;
    const int q_n = (q * 32) / (p.W * p.BH);
    uint8_t* stage0 = epi_smem + (warp_u - 2) * p.epi_bufs * kEpiStageBytes;
    float* vec = epi_vec + (warp_u - 2) * (kEpiVecBytes / 4);
    uint64_t* my_res_bar = res_bar + 2 * (warp_u - 2);
    uint32_t res_ph0 = 0, res_ph1 = 0;
    // (bias + row vector) of the warp's columns go through `vec` when the warp's 32 rows lie in one image
    // TMA-store path (no folded LayerNorm): (bias + per-image row vector) of the warp's columns are staged in `vec` at the top of
    // every tile, unconditionally (zeros without a bias).  The row vector joins them when the warp's 32 rows lie in one image
    // (always, except for images of fewer than 32 pixels); otherwise it is fetched per thread.
    const bool use_vec = !kLn && p.tma_out;
    const bool rv_vec = p.rowvec != nullptr && (p.W * p.BH) % 32 == 0;
    const int half = p.block_n / 2;
    const int out_cols = p.geglu ? half : p.block_n;     // output columns per tile
    int it = 0;
    for (int t = tile0; t < num_tiles; t += tile_step, ++it) {
      const int buf = it & 1;
      const uint32_t use = it >> 1;
      const int n_tile = t % p.num_n_tiles;
      const int m_group = (t / p.num_n_tiles) % p.num_m_groups;
      const int m_tile = kCta2 ? 2 * m_group + static_cast<int>(cta_rank) : m_group;
      const int split = t / (p.num_n_tiles * p.num_m_groups);
      const int h_t = (m_tile % p.tiles_h) * p.BH;
      const int n_t = (m_tile / p.tiles_h) * p.BNI;
      const int h = h_t + hl;
      const int n = n_t + nl;
      const bool row_ok = (h < p.H) && (n < p.NB);
      const size_t pix = (static_cast<size_t>(n) * p.H + h) * p.W + wl;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * p.block_n;
      const int nchunks = p.tma_out ? (out_cols >> 6) : 0;
      [[maybe_unused]] float ln_mu = 0.f, ln_rs = 0.f;
      if (kLn) {
        // ---- LayerNorm statistics of this thread's row: the producer of x left per-chunk partial sums (fixed order)
        const int grow = m_tile * kBlockM + row;
        float sx = 0.f, sq = 0.f;
        if (grow < p.m_rows) {
          const float2* st = reinterpret_cast<const float2*>(p.ln_stats) + static_cast<size_t>(grow) * p.ln_chunks;
          for (int ch = 0; ch < p.ln_chunks; ++ch) {
            const float2 v = st[ch];
            sx += v.x;
            sq += v.y;
          }
        }
        ln_mu = sx / p.ln_c;
        ln_rs = grow < p.m_rows ? rsqrtf(fmaxf(sq / p.ln_c - ln_mu * ln_mu, 0.f) + p.ln_eps) : 0.f;
      }
      if (kLora && hf == 0) {
        // ---- T (fp32, TMEM) -> bf16 K-major 128B-swizzled shared-memory tile (+ optional global copy for the backward)
        mbar_wait(lt_tmem_bar, it & 1);
        tc_fence_after();
        const uint32_t t_src = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 2 * p.block_n;
        uint32_t pk[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) pk[j] = 0u;
        for (int c16 = 0; c16 < p.lora_n / 16; ++c16) {
          uint32_t r[16];
          tmem_ld_x16(t_src + c16 * 16, r);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float t0 = __uint_as_float(r[2 * j]), t1 = __uint_as_float(r[2 * j + 1]);
            if (kLn) {
              // T = LN(x) A^T = rstd (x (gamma o A)^T - mu ga) + ba.  The tile's accumulator is multiplied by rstd in the
              // epilogue as a whole, LoRA term included, so the tile holds T / rstd here.
              const int col = c16 * 16 + 2 * j;
              const float inv_rs = ln_rs > 0.f ? __fdividef(1.0f, ln_rs) : 0.f;
              t0 = fmaf(p.ln_ba[col], inv_rs, t0 - ln_mu * p.ln_ga[col]);
              t1 = fmaf(p.ln_ba[col + 1], inv_rs, t1 - ln_mu * p.ln_ga[col + 1]);
            }
            const uint32_t v = pack_bf16x2(t0, t1);
            // (c16 is not a compile-time constant: select the destination words without dynamic register indexing)
            if (c16 == 0) pk[j] = v; else if (c16 == 1) pk[8 + j] = v; else if (c16 == 2) pk[16 + j] = v; else pk[24 + j] = v;
          }
        }
        const uint32_t t_row_addr = smem_u32(t_tile) + row * 128;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const uint32_t addr = t_row_addr + ((g ^ (row & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[g * 4]), "r"(pk[g * 4 + 1]),
                       "r"(pk[g * 4 + 2]), "r"(pk[g * 4 + 3])
                       : "memory");
        }
        if (p.t_out != nullptr) {
          const int grow = m_tile * kBlockM + row;        // linear layers: one image, W = 1, 128 rows per tile
          if (n_tile == 0 && grow < p.m_rows) {
            uint4* dst = reinterpret_cast<uint4*>(p.t_out + static_cast<size_t>(grow) * 64);
#pragma unroll
            for (int g = 0; g < 8; ++g) dst[g] = make_uint4(pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(lt_smem_bar);
      }
      if (p.res_tma) {
        // ---- residual tiles of this warp's chunks: TMA into the staging buffers, in flight while the main loop runs
        // (bulk-store groups belong to the thread that committed them: elect.sync picks the same lane every time)
        if (elect_one()) tma_store_wait_read<0>();     // the previous tile's stores have finished reading the buffers
        for (int i = 0; hf + 2 * i < nchunks; ++i) {
          if (elect_one()) {
            mbar_expect_tx(&my_res_bar[i], static_cast<uint32_t>(kEpiStageBytes));
            tma_load_4d(stage0 + i * kEpiStageBytes, &tmRes, &my_res_bar[i], n_tile * out_cols + (hf + 2 * i) * 64, q_w, h_t + q_h,
                        n_t + q_n);
          }
        }
      }
      if (use_vec) {
        const int n_w = n_t + q_n;                     // image of this warp's rows
        for (int i = 0; hf + 2 * i < nchunks; ++i) {
          const int cc = hf + 2 * i;
          if (!p.geglu) {
            const int col = n_tile * out_cols + cc * 64 + 2 * lane;
            float2 b = p.bias ? *reinterpret_cast<const float2*>(p.bias + col) : make_float2(0.f, 0.f);
            if (rv_vec && n_w < p.NB && col < p.n_valid) {
              const float2 r = *reinterpret_cast<const float2*>(p.rowvec + static_cast<size_t>(n_w) * p.rowvec_ld + col);
              b.x += r.x; b.y += r.y;
            }
            *reinterpret_cast<float2*>(vec + i * 64 + 2 * lane) = b;
          } else {
            const int bcol = n_tile * p.block_n + cc * 64 + 2 * lane;
            const float2 z = make_float2(0.f, 0.f);
            *reinterpret_cast<float2*>(vec + i * 128 + 2 * lane) = p.bias ? *reinterpret_cast<const float2*>(p.bias + bcol) : z;
            *reinterpret_cast<float2*>(vec + i * 128 + 64 + 2 * lane) = p.bias ? *reinterpret_cast<const float2*>(p.bias + bcol + half) : z;
          }
        }
        __syncwarp();
      }
      mbar_wait(&tfull_bar[buf], use & 1);
      if (warp == 2 && lane == 0) TL(12);
      tc_fence_after();

      if (p.tma_out) {
        // -------- bf16, stride 1: 64-column chunks -> swizzled smem -> TMA store
        for (int cc = hf; cc < nchunks; cc += 2) {
          const int ci = cc >> 1;                          // this warp's ci-th chunk of the tile
          uint8_t* stage = stage0 + (p.epi_bufs == 2 ? ci * kEpiStageBytes : 0);
          const uint32_t stage_row = smem_u32(stage) + lane * 128;
          if (p.res_tma) {
            if (ci == 0) { mbar_wait(&my_res_bar[0], res_ph0); res_ph0 ^= 1; }
            else { mbar_wait(&my_res_bar[1], res_ph1); res_ph1 ^= 1; }
          }
          uint32_t pk[32];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int oc = cc * 64 + hh * 32;              // output column inside the tile
            const int gcol = n_tile * out_cols + oc;       // global output column
            uint32_t r[32];
            if (!p.geglu) {
              tmem_ld_x32(t_row + oc, r);
              tmem_wait_ld();
              if (warp == 2 && lane == 0 && it == 0 && cc == 0 && hh == 0) TL(17);
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
                const int cg = gcol + g * 8;
                if (kLn) ln_affine8(v, p.ln_g + cg, p.bias + cg, ln_mu, ln_rs);
                else add8(v, vec + ci * 64 + hh * 32 + g * 8);
                if (row_ok && cg < p.n_valid) {
                  if (p.rowvec && (kLn || !rv_vec)) add8(v, p.rowvec + static_cast<size_t>(n) * p.rowvec_ld + cg);
                  if (p.res_tma) {
                    uint4 rr;             // this thread's row of the residual tile the TMA left in the staging buffer
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rr.x), "=r"(rr.y), "=r"(rr.z), "=r"(rr.w)
                                 : "r"(stage_row + (((hh * 4 + g) ^ (lane & 7)) << 4)));
                    if (kAct) add_res8_unact(v, rr, p.res_neg_gain);
                    else {
                      v[0] += bf16_lo(rr.x); v[1] += bf16_hi(rr.x); v[2] += bf16_lo(rr.y); v[3] += bf16_hi(rr.y);
                      v[4] += bf16_lo(rr.z); v[5] += bf16_hi(rr.z); v[6] += bf16_lo(rr.w); v[7] += bf16_hi(rr.w);
                    }
                  }
                }
                if (kAct) act8(v, p.act_slope, p.act_tanh);
#pragma unroll
                for (int j = 0; j < 4; ++j) pk[hh * 16 + g * 4 + j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
              }
            } else {
              // GEGLU: tile columns [0, bn/2) are values, [bn/2, bn) the matching gates.
              uint32_t rg[32];
              tmem_ld_x32(t_row + oc, r);
              tmem_ld_x32(t_row + half + oc, rg);
              tmem_wait_ld();
              const int bcol = n_tile * p.block_n + oc;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float v[8], gt[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  v[j] = __uint_as_float(r[g * 8 + j]);
                  gt[j] = __uint_as_float(rg[g * 8 + j]);
                }
                if (kLn) {
                  ln_affine8(v, p.ln_g + bcol + g * 8, p.bias + bcol + g * 8, ln_mu, ln_rs);
                  ln_affine8(gt, p.ln_g + bcol + half + g * 8, p.bias + bcol + half + g * 8, ln_mu, ln_rs);
                } else {
                  add8(v, vec + ci * 128 + hh * 32 + g * 8);
                  add8(gt, vec + ci * 128 + 64 + hh * 32 + g * 8);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] *= gelu_erf(gt[j]);
#pragma unroll
                for (int j = 0; j < 4; ++j) pk[hh * 16 + g * 4 + j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
              }
            }
          }
          if (warp == 2 && lane == 0 && it == 0 && cc == 0) TL(16);
          if (kStat && row_ok && n_tile * out_cols + cc * 64 < p.n_valid) {
            // row statistics of the values being stored (after bf16 rounding): what the consuming LayerNorm-folded GEMM needs
            float sx = 0.f, sq = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float lo = bf16_lo(pk[j]), hi = bf16_hi(pk[j]);
              sx += lo + hi;
              sq = fmaf(lo, lo, sq);
              sq = fmaf(hi, hi, sq);
            }
            const int gchunk = (n_tile * out_cols + cc * 64) >> 6;
            reinterpret_cast<float2*>(p.stat_out)[pix * p.stat_chunks + gchunk] = make_float2(sx, sq);
          }
          // the previous TMA store of this warp must have finished reading the staging tile (res_tma: waited at the tile's top)
          if (!p.res_tma) {
            if (elect_one()) tma_store_wait_read<0>();
            __syncwarp();
          }
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const uint32_t addr = stage_row + ((g ^ (lane & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[g * 4]), "r"(pk[g * 4 + 1]),
                         "r"(pk[g * 4 + 2]), "r"(pk[g * 4 + 3])
                         : "memory");
          }
          if (warp == 2 && lane == 0 && it == 0 && cc == 0) TL(18);
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {
            tma_store_4d(&tmOut, stage, n_tile * out_cols + cc * 64, q_w, h_t + q_h, n_t + q_n);
            tma_store_commit();
          }
          if (warp == 2 && lane == 0 && it == 0 && cc == 0) TL(19);
          if (kGn) {
            // ---- GroupNorm partial statistics of this warp's 32 rows x 64 columns, read back TRANSPOSED from the staging
            //      tile (lane l owns columns 2l, 2l+1; the swizzle keeps the 32 lanes on 32 different banks), rows outside the
            //      image masked, lane pairs combined into 4-channel units
            const uint32_t valid = __ballot_sync(0xffffffffu, row_ok);
            const uint32_t tile_a = smem_u32(stage) + ((lane & 3) << 2);
            const uint32_t ch16 = lane >> 2;
            float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
            for (int r = 0; r < 32; ++r) {
              uint32_t v;
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(tile_a + r * 128 + ((ch16 ^ (r & 7)) << 4)));
              if ((valid >> r) & 1u) {
                const float lo = bf16_lo(v), hi = bf16_hi(v);
                s0 += lo; q0 = fmaf(lo, lo, q0);
                s1 += hi; q1 = fmaf(hi, hi, q1);
              }
            }
            s0 += s1; q0 += q1;
            s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
            q0 += __shfl_xor_sync(0xffffffffu, q0, 1);
            const int rpi = p.W * p.BH;                         // tile rows per image (a multiple of 32: host-checked)
            const int n_img = n_t + (q * 32) / rpi;
            const int col = n_tile * out_cols + cc * 64 + 2 * lane;
            if ((lane & 1) == 0 && col < p.n_valid && n_img < p.NB && m_tile < p.num_m_tiles) {
              const int spi = rpi >> 5;                         // slabs per (tile, image)
              const int slab = (m_tile % p.tiles_h) * spi + ((q * 32) % rpi >> 5);
              reinterpret_cast<float2*>(p.gn_stat)[(static_cast<size_t>(n_img) * p.gn_slabs + slab) * (p.n_valid >> 2) + (col >> 2)] =
                  make_float2(s0, q0);
            }
          }
        }
      } else {
        // -------- fp32 output and/or stride 2: direct 16-byte stores, 32-column groups
        const int col_base = n_tile * p.block_n;
        for (int c = hf; c < p.block_n / 32; c += 2) {
          uint32_t r[32];
          tmem_ld_x32(t_row + c * 32, r);
          tmem_wait_ld();
          const int col0 = col_base + c * 32;
          if (row_ok) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int cg = col0 + g * 8;
              if (cg < p.n_valid) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
                if (p.bias) add8(v, p.bias + cg);
                if (p.rowvec) add8(v, p.rowvec + static_cast<size_t>(n) * p.rowvec_ld + cg);
                if (p.residual) {
                  const uint4 rr = *reinterpret_cast<const uint4*>(p.residual + pix * p.res_ld + cg);
                  if (kAct) add_res8_unact(v, rr, p.res_neg_gain);
                  else {
                    v[0] += bf16_lo(rr.x); v[1] += bf16_hi(rr.x); v[2] += bf16_lo(rr.y); v[3] += bf16_hi(rr.y);
                    v[4] += bf16_lo(rr.z); v[5] += bf16_hi(rr.z); v[6] += bf16_lo(rr.w); v[7] += bf16_hi(rr.w);
                  }
                }
                if (kAct) act8(v, p.act_slope, p.act_tanh);
                if (p.out_fp32) {
                  float* o = reinterpret_cast<float*>(p.out) + split * p.split_stride + pix * p.out_ld + cg;
                  *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
                  *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
                } else {
                  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.out_ld + cg;
                  uint4 pk;
                  pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]);
                  pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
                  *reinterpret_cast<uint4*>(o) = pk;
                }
              }
            }
          }
        }
      }
      if (warp == 2 && lane == 0) TL(13);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kCta2) mbar_arrive_cluster(leader_addr(&tempty_bar[buf]));   // the MMA issuer lives in the leader CTA
        else mbar_arrive(&tempty_bar[buf]);
      }
    }
    // smem may not be released while a bulk store still reads it; global visibility comes with grid completion
    if (p.tma_out && elect_one()) tma_store_wait_read<0>();
    if (warp == 2 && lane == 0) TL(14);
  }
  if (warp == 0 && lane == 0) TL(5);

  tc_fence_before();
  if (kCta2) cluster_sync_all();      // nobody exits while the peer may still signal its barriers / read its smem
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (kCta2) tmem_dealloc2(tmem_base, p.tmem_cols);
    else tmem_dealloc(tmem_base, p.tmem_cols);
  }
  if (threadIdx.x == 32) TL(15);
}

// Split-K second stage: out[pix, n] = sum_s ws[s][pix][n] (fixed order) + bias + rowvec[image] + residual -> bf16.
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ ws, int ksplit, size_t split_stride, int M, int n_valid, int ld_ws,
                     const float* __restrict__ bias, const float* __restrict__ rowvec, int rowvec_ld, int hw,
                     const __nv_bfloat16* __restrict__ residual, int res_ld, __nv_bfloat16* __restrict__ out, int out_ld) {
  pdl_launch_dependents();
  pdl_wait();
  const int vec_per_row = n_valid >> 3;
  const size_t total = static_cast<size_t>(M) * vec_per_row;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t pix = i / vec_per_row;
    const int cg = static_cast<int>(i % vec_per_row) * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    for (int s = 0; s < ksplit; ++s) add8(v, ws + s * split_stride + pix * ld_ws + cg);
    if (bias) add8(v, bias + cg);
    if (rowvec) add8(v, rowvec + (pix / hw) * rowvec_ld + cg);
    if (residual) {
      const uint4 rr = *reinterpret_cast<const uint4*>(residual + pix * res_ld + cg);
      v[0] += bf16_lo(rr.x); v[1] += bf16_hi(rr.x); v[2] += bf16_lo(rr.y); v[3] += bf16_hi(rr.y);
      v[4] += bf16_lo(rr.z); v[5] += bf16_hi(rr.z); v[6] += bf16_lo(rr.w); v[7] += bf16_hi(rr.w);
    }
    uint4 pk;
    pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]);
    pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + pix * out_ld + cg) = pk;
  }
}

static int pick_box(int H, int W, int NB, int* BH, int* BNI) {
  if (W < 1 || W > 128 || (128 % W) != 0) return -1;
  long best = -1;
  for (int bh = 1; bh * W <= 128; bh *= 2) {
    const int bni = 128 / (W * bh);
    if (bni > 256) continue;
    const long tiles = static_cast<long>((H + bh - 1) / bh) * ((NB + bni - 1) / bni);
    if (best < 0 || tiles < best || (tiles == best && bni == 1)) {
      best = tiles;
      *BH = bh;
      *BNI = bni;
    }
  }
  return best > 0 ? 0 : -1;
}

}  // namespace b200

using namespace b200;

// Debug only (not part of the product API surface beyond the header note): copy the CTA-0 timeline out.
extern "C" int b200_debug_timeline(unsigned long long* host_out, int n) {
  if (n > 160) n = 160;
  cudaError_t e = cudaMemcpyFromSymbol(host_out, g_timeline, sizeof(unsigned long long) * n);
  return e == cudaSuccess ? B200_OK : fail(B200_ERR_CUDA, "debug_timeline: %s", cudaGetErrorString(e));
}

// Hits of the process-wide tensor-map cache (host_util.h) so far: the eager paths re-use encoded descriptors.
extern "C" long b200_tmap_cache_hits(void) {
  std::lock_guard<std::mutex> lock(tmap_mutex());
  return static_cast<long>(tmap_cache_hits());
}

static int conv_gemm_impl(const void* a0, int c0, const void* a1, int c1, const void* a2, int c2, int nb, int h,
                          int w, int ntaps, int stride, const void* wpacked, int n_pad, int n_valid,
                          const float* bias, const float* rowvec, int rowvec_ld, const void* residual, int res_ld,
                          void* out, int out_ld, int out_fp32, int geglu, int block_n, int max_ctas,
                          int ksplit, float* workspace, int cta_pair, const void* lora_down, int lora_rows, void* t_out,
                          const float* ln_g, const float* ln_ga, const float* ln_ba, float ln_eps, const float* ln_stats,
                          float* stat_out, void* stream_v, float* gn_stat = nullptr, int b_dynamic = 0,
                          const struct Conv1dOpts* c1d = nullptr, int s2_pad01 = 0);

// 1-D tap walk and output geometry of b200_conv1d (everything else is the 2-D kernel with W = 1)
struct Conv1dOpts {
  int dh0, dh_step;        // first tap's row offset and the step between taps
  int m_h;                 // output rows per image (>= or <= the input length h: transposed-conv phases)
  long out_batch_stride;   // elements between images of the output (0: m_h * out_ld)
  float act_slope;         // LeakyReLU slope applied after bias / residual (1: identity)
  float res_neg_gain;      // residual stored post-LeakyReLU(s): 1 / s; plain residual: 1
  int act_tanh;            // 1: tanh, 2: exact-erf GELU instead of LeakyReLU
};

// C-ABI: see include/b200ldm.h
extern "C" int b200_conv_gemm(const void* a0, int c0, const void* a1, int c1, const void* a2, int c2, int nb, int h,
                              int w, int ntaps, int stride, const void* wpacked, int n_pad, int n_valid,
                              const float* bias, const float* rowvec, int rowvec_ld, const void* residual, int res_ld,
                              void* out, int out_ld, int out_fp32, int geglu, int block_n, int max_ctas,
                              int ksplit, float* workspace, int cta_pair, void* stream_v) {
  return conv_gemm_impl(a0, c0, a1, c1, a2, c2, nb, h, w, ntaps, stride, wpacked, n_pad, n_valid, bias, rowvec, rowvec_ld,
                        residual, res_ld, out, out_ld, out_fp32, geglu, block_n, max_ctas, ksplit, workspace, cta_pair,
                        nullptr, 0, nullptr, nullptr, nullptr, nullptr, 0.f, nullptr, nullptr, stream_v);
}

// 3x3 stride-2 convolution with the VAE encoder's asymmetric padding (diffusers Downsample2D(padding=0):
// F.pad(x, (0, 1, 0, 1)) then Conv2d(k 3, s 2, p 0)): out[ho, wo] = sum in[2 ho + kh, 2 wo + kw] W[kh, kw], zeros past the
// bottom / right edge.  Same strided-TMA implicit GEMM as b200_conv_gemm(stride = 2); see include/b200ldm.h.
extern "C" int b200_conv3x3_s2_pad01(const void* x, int c, int nb, int h, int w, const void* wpacked, int n_pad, int n_valid,
                                     const float* bias, void* out, int out_ld, int block_n, int cta_pair, void* stream_v) {
  B200_CHECK_ARG(h >= 2 && w >= 2, "conv3x3_s2_pad01: needs at least 2 x 2 pixels");
  return conv_gemm_impl(x, c, nullptr, 0, nullptr, 0, nb, h, w, 9, 2, wpacked, n_pad, n_valid, bias, nullptr, 0, nullptr, 0, out,
                        out_ld, 0, 0, block_n, 0, 1, nullptr, cta_pair, nullptr, 0, nullptr, nullptr, nullptr, nullptr, 0.f, nullptr,
                        nullptr, stream_v, nullptr, 0, nullptr, 1);
}

// b200_conv_gemm whose epilogue also leaves the partial GroupNorm statistics of its (bf16, stride-1, un-split) output for
// b200_groupnorm_apply: gn_stat fp32 [nb, slabs, n_valid / 4, 2], slabs = b200_gn_stat_slabs(nb, h, w).
extern "C" int b200_conv_gemm_gnstat(const void* a0, int c0, const void* a1, int c1, const void* a2, int c2, int nb, int h,
                                     int w, int ntaps, const void* wpacked, int n_pad, int n_valid, const float* bias,
                                     const float* rowvec, int rowvec_ld, const void* residual, int res_ld, void* out,
                                     int out_ld, int block_n, int max_ctas, int cta_pair, float* gn_stat, void* stream_v) {
  B200_CHECK_ARG(gn_stat != nullptr, "conv_gemm_gnstat: null gn_stat");
  return conv_gemm_impl(a0, c0, a1, c1, a2, c2, nb, h, w, ntaps, 1, wpacked, n_pad, n_valid, bias, rowvec, rowvec_ld,
                        residual, res_ld, out, out_ld, 0, 0, block_n, max_ctas, 1, nullptr, cta_pair,
                        nullptr, 0, nullptr, nullptr, nullptr, nullptr, 0.f, nullptr, nullptr, stream_v, gn_stat);
}

// out[m, n] = sum_k a[m, k] * b[n, k] (+ residual) with BOTH operands activations (bf16, K-major, written earlier in the
// stream): the scores S = Q K^T and O = P V of the VAE decoder's head_dim-512 attention.  Same kernel as b200_conv_gemm;
// the only difference is that the "weight" producer warp honours the programmatic-dependent-launch wait (weights proper
// are never written inside a step, so b200_conv_gemm lets that warp start before the previous grid has drained).
// b must be readable for b_rows rows (a multiple of block_n >= n_valid; rows >= n_valid only feed discarded columns).
extern "C" int b200_gemm_nt(const void* a, int m, int k, const void* b, int b_rows, int n_valid, void* out, int out_ld,
                            int out_fp32, int block_n, int cta_pair, void* stream_v) {
  return conv_gemm_impl(a, k, nullptr, 0, nullptr, 0, 1, m, 1, 1, 1, b, b_rows, n_valid, nullptr, nullptr, 0, nullptr, 0, out,
                        out_ld, out_fp32, 0, block_n, 0, 1, nullptr, cta_pair, nullptr, 0, nullptr, nullptr, nullptr, nullptr,
                        0.f, nullptr, nullptr, stream_v, nullptr, 1);
}

// 1-D convolution over time-major bf16 activations x [nb, len, c]  (HiFi-GAN: Conv1d with dilation, and -- one launch per
// output phase -- ConvTranspose1d); see include/b200ldm.h::b200_conv1d:
//   out[n, q, :] = act( bias + sum_{t < ntaps} x[n, q + dh0 + t * dh_step, :] . W_t^T  + unact(residual[n, q, :]) ),   q < m_rows
// rows outside [0, len) read as zero (TMA fill = the convolution's zero padding).  wpacked bf16 [n_pad, ntaps * c], tap-major.
extern "C" int b200_conv1d(const void* x, int c, int nb, int len, int ntaps, int dh0, int dh_step, int m_rows,
                           const void* wpacked, int n_pad, int n_valid, const float* bias, const void* residual, int res_ld,
                           float res_neg_gain, void* out, int out_ld, long out_batch_stride, int out_fp32, float act_slope,
                           int act_tanh, int block_n, int cta_pair, void* stream_v) {
  Conv1dOpts o;
  o.dh0 = dh0; o.dh_step = dh_step; o.m_h = m_rows; o.out_batch_stride = out_batch_stride;
  o.act_slope = act_slope; o.res_neg_gain = res_neg_gain; o.act_tanh = act_tanh;
  B200_CHECK_ARG(!residual || m_rows == len, "conv1d: a residual needs m_rows == len");
  B200_CHECK_ARG(!out_fp32 || out_batch_stride == 0, "conv1d: fp32 output is contiguous");
  B200_CHECK_ARG(act_slope <= 1.0f && act_slope >= 0.f && res_neg_gain >= 1.0f, "conv1d: needs 0 <= act_slope <= 1 <= res_neg_gain");
  return conv_gemm_impl(x, c, nullptr, 0, nullptr, 0, nb, len, 1, ntaps, 1, wpacked, n_pad, n_valid, bias, nullptr, 0, residual,
                        res_ld, out, out_ld, out_fp32, 0, block_n, 0, 1, nullptr, cta_pair, nullptr, 0, nullptr, nullptr, nullptr,
                        nullptr, 0.f, nullptr, nullptr, stream_v, nullptr, 0, &o);
}

// Slabs per image of the statistics b200_conv_gemm_gnstat writes for an [nb, h, w, *] output; 0: this geometry is not
// supported (a 32-row warp slice of a tile would span two images).
extern "C" int b200_gn_stat_slabs(int nb, int h, int w) {
  int BH = 0, BNI = 0;
  if (pick_box(h, w, nb, &BH, &BNI) != 0) return 0;
  const int rpi = w * BH;
  if (rpi % 32 != 0) return 0;
  return ((h + BH - 1) / BH) * (rpi / 32);
}

// Linear layer with the rank-r LoRA branch computed INSIDE the kernel (peft lora.Linear, unmerged):
//   out = x . W^T + (x . A^T) . (s B)^T (+ bias + residual)
// x bf16 [m, c]; wpacked bf16 [n_pad, c + 64] = [W | s.B padded to 64 columns]; lora_down bf16 [64, c] = the stacked
// lora_A rows (lora_rows of them valid, the rest zero).  t_out (nullable): bf16 [m, 64] copy of T = x . A^T.
extern "C" int b200_linear_lora(const void* x, int c, int m, const void* wpacked, int n_pad, int n_valid, const float* bias,
                                const void* residual, int res_ld, void* out, int out_ld, int block_n, int max_ctas,
                                const void* lora_down, int lora_rows, void* t_out, void* stream_v) {
  B200_CHECK_ARG(lora_down && lora_rows > 0 && lora_rows <= 64, "linear_lora: lora_down / lora_rows (%d) invalid", lora_rows);
  return conv_gemm_impl(x, c, nullptr, 64, nullptr, 0, 1, m, 1, 1, 1, wpacked, n_pad, n_valid, bias, nullptr, 0, residual, res_ld,
                        out, out_ld, 0, 0, block_n, max_ctas, 1, nullptr, 0, lora_down, lora_rows, t_out, nullptr, nullptr,
                        nullptr, 0.f, nullptr, nullptr, stream_v);
}

// Any linear layer [m, c] -> [m, n] that ALSO leaves the row statistics of its (bf16) output for a following
// LayerNorm-folded GEMM (b200_linear_ln): stat_out fp32 [m, n_valid / 64, 2] = per row and 64-column chunk (sum, sum of
// squares).  a1 / c1: optional K segment (e.g. a LoRA T tensor); lora_down: optional in-kernel LoRA (then a1 = null).
extern "C" int b200_linear_stats(const void* x, int c, const void* a1, int c1, int m, const void* wpacked, int n_pad,
                                 int n_valid, const float* bias, const void* residual, int res_ld, void* out, int out_ld,
                                 int block_n, int max_ctas, const void* lora_down, int lora_rows, float* stat_out,
                                 void* stream_v) {
  B200_CHECK_ARG(stat_out && n_valid % 64 == 0, "linear_stats: stat_out required, n_valid %d must be a multiple of 64", n_valid);
  return conv_gemm_impl(x, c, a1, lora_down ? 64 : c1, nullptr, 0, 1, m, 1, 1, 1, wpacked, n_pad, n_valid, bias, nullptr, 0,
                        residual, res_ld, out, out_ld, 0, 0, block_n, max_ctas, 1, nullptr, 0, lora_down, lora_rows, nullptr,
                        nullptr, nullptr, nullptr, 0.f, nullptr, stat_out, stream_v);
}

// Linear layer consuming LayerNorm(x) with the LayerNorm folded in (and, optionally, the in-kernel LoRA branch):
//   out = LN(x) W^T (+ (LN(x) A^T)(s B)^T) (+ residual), optionally through GEGLU.
// The GEMM runs on the raw rows x with gamma folded into the packed weights (wpacked = [gamma o W | s.B]); the epilogue
// applies rstd, the rank-1 mean correction mu * ln_g[n] and b' = `bias` (= beta W^T + the layer's bias) per row -- the
// row statistics are computed by the epilogue warps while the main loop runs.  ln_g fp32 [n_pad] (packed order);
// with LoRA: lora_down = gamma o A stacked [64, c], ln_ga / ln_ba fp32 [64].  Removes the LayerNorm launch and the
// normalised activation's round trip through HBM.
extern "C" int b200_linear_ln(const void* x, int c, int m, const void* wpacked, int n_pad, int n_valid, const float* bias,
                              const float* ln_g, float ln_eps, const float* ln_stats, const void* residual, int res_ld,
                              void* out, int out_ld, int geglu, int block_n, int max_ctas, const void* lora_down,
                              int lora_rows, const float* ln_ga, const float* ln_ba, void* stream_v) {
  B200_CHECK_ARG(bias && ln_g && ln_stats && c % 64 == 0, "linear_ln: bias (b'), ln_g and ln_stats are required");
  B200_CHECK_ARG(!lora_down || (ln_ga && ln_ba && lora_rows > 0 && lora_rows <= 64), "linear_ln: LoRA arguments invalid");
  return conv_gemm_impl(x, c, nullptr, lora_down ? 64 : 0, nullptr, 0, 1, m, 1, 1, 1, wpacked, n_pad, n_valid, bias, nullptr, 0,
                        residual, res_ld, out, out_ld, 0, geglu, block_n, max_ctas, 1, nullptr, 0, lora_down, lora_rows, nullptr,
                        ln_g, ln_ga, ln_ba, ln_eps, ln_stats, nullptr, stream_v);
}

static int conv_gemm_impl(const void* a0, int c0, const void* a1, int c1, const void* a2, int c2, int nb, int h,
                          int w, int ntaps, int stride, const void* wpacked, int n_pad, int n_valid,
                          const float* bias, const float* rowvec, int rowvec_ld, const void* residual, int res_ld,
                          void* out, int out_ld, int out_fp32, int geglu, int block_n, int max_ctas,
                          int ksplit, float* workspace, int cta_pair, const void* lora_down, int lora_rows, void* t_out,
                          const float* ln_g, const float* ln_ga, const float* ln_ba, float ln_eps, const float* ln_stats,
                          float* stat_out, void* stream_v, float* gn_stat, int b_dynamic, const Conv1dOpts* c1d, int s2_pad01) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const bool s2 = stride == 2;                      // Downsample2D: computed at OUTPUT resolution (strided activation boxes)
  // (s2_pad01: the VAE encoder's Downsample2D(padding=0) = F.pad(x, (0, 1, 0, 1)) + conv k3 s2: taps at offsets 0..2, not -1..1)
  const int m_h = s2 ? (s2_pad01 ? (h - 2) / 2 + 1 : (h - 1) / 2 + 1) : (c1d ? c1d->m_h : h);     // output rows per image
  const int w_o = s2 ? (s2_pad01 ? (w - 2) / 2 + 1 : (w - 1) / 2 + 1) : w;                        // output width
  const bool fused_lora = lora_down != nullptr;
  const bool fused_ln = ln_g != nullptr;
  if (fused_ln) {
    B200_CHECK_ARG(ntaps == 1 && nb == 1 && w == 1 && stride == 1 && ksplit <= 1 && !out_fp32 && c2 == 0 && !rowvec &&
                   block_n % 64 == 0, "linear_ln: needs a plain bf16 linear layer");
    cta_pair = 0;
  }
  if (fused_lora) {
    B200_CHECK_ARG(ntaps == 1 && nb == 1 && w == 1 && stride == 1 && ksplit <= 1 && !geglu && !out_fp32 && c1 == 64 && !a1 &&
                   c2 == 0, "linear_lora: needs a plain linear layer with a 64-column LoRA segment");
    B200_CHECK_ARG(!fused_ln || !t_out, "linear_ln: no global copy of T in the LayerNorm-fused form");
    B200_CHECK_ARG(2 * block_n + 64 <= 512, "linear_lora: block_n %d leaves no TMEM columns for T", block_n);
    cta_pair = 0;
  }
  B200_CHECK_ARG(stride == 1 || (stride == 2 && ntaps == 9 && c1 == 0 && c2 == 0), "conv_gemm: stride %d unsupported", stride);
  B200_CHECK_ARG(a0 && wpacked && out, "conv_gemm: null pointer");
  B200_CHECK_ARG(ntaps == 1 || ntaps == 9 || (c1d && ntaps >= 1 && ntaps <= 16), "conv_gemm: ntaps must be 1 or 9 (got %d)", ntaps);
  B200_CHECK_ARG(!c1d || (w == 1 && stride == 1 && ksplit <= 1 && !geglu && !fused_lora && !fused_ln && !stat_out && !gn_stat &&
                          c1 == 0 && c2 == 0 && m_h > 0), "conv1d: needs a plain [nb, L, C] layer");
  B200_CHECK_ARG(c0 > 0 && c0 % 64 == 0 && c1 % 64 == 0 && c2 % 64 == 0, "conv_gemm: channels must be multiples of 64 (%d,%d,%d)", c0, c1, c2);
  B200_CHECK_ARG(((c1 == 0) == (a1 == nullptr) || fused_lora) && (c2 == 0) == (a2 == nullptr), "conv_gemm: segment pointer/channel mismatch");
  B200_CHECK_ARG(block_n >= 32 && block_n <= 256 && block_n % 32 == 0, "conv_gemm: block_n %d unsupported", block_n);
  B200_CHECK_ARG(n_pad % block_n == 0, "conv_gemm: n_pad %d not a multiple of block_n %d", n_pad, block_n);
  B200_CHECK_ARG(n_valid % 8 == 0 && n_valid <= (geglu ? n_pad / 2 : n_pad), "conv_gemm: n_valid %d invalid", n_valid);
  B200_CHECK_ARG(!geglu || (block_n % 128 == 0 && !out_fp32 && !residual && !rowvec && stride == 1), "conv_gemm: geglu constraints");
  B200_CHECK_ARG(nb > 0 && h > 0 && w > 0, "conv_gemm: empty activation");
  B200_CHECK_ARG(out_ld % 8 == 0, "conv_gemm: out_ld %d must be a multiple of 8", out_ld);
  if (ksplit < 1) ksplit = 1;
  B200_CHECK_ARG(ksplit == 1 || (workspace && stride == 1 && !geglu && !out_fp32),
                 "conv_gemm: split-K needs a workspace, stride 1, bf16 output, no GEGLU");

  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  if (pick_box(m_h, w_o, nb, &p.BH, &p.BNI) != 0) return fail(B200_ERR_UNSUPPORTED, "conv_gemm: width %d does not divide 128", w_o);
  if (c1d) {
    p.dh0 = c1d->dh0; p.dh_step = c1d->dh_step; p.dw0 = 0; p.dw_end = 1;
    p.act_slope = c1d->act_slope; p.res_neg_gain = c1d->res_neg_gain; p.act_tanh = c1d->act_tanh;
  } else if (ntaps == 9 && s2 && s2_pad01) {
    p.dh0 = 0; p.dw0 = 0; p.dw_end = 3; p.dh_step = 1;
  } else if (ntaps == 9) {
    p.dh0 = -1; p.dw0 = -1; p.dw_end = 2; p.dh_step = 1;
  } else {
    p.dh0 = 0; p.dw0 = 0; p.dw_end = 1; p.dh_step = 0;
  }
  p.cb0 = c0 / 64;
  p.ntaps = ntaps;
  p.seg_end0 = ntaps * p.cb0;
  p.seg_end1 = p.seg_end0 + c1 / 64;
  p.num_kb = p.seg_end1 + c2 / 64;
  p.H = m_h; p.W = w_o; p.NB = nb;        // the epilogue's geometry (== the input's except for stride 2 and b200_conv1d with m_h != h)
  p.a_stride = s2 ? 2 : 1;
  p.tiles_h = (m_h + p.BH - 1) / p.BH;
  p.num_m_tiles = p.tiles_h * ((nb + p.BNI - 1) / p.BNI);
  p.num_n_tiles = n_pad / block_n;
  p.block_n = block_n;
  p.n_valid = n_valid;
  p.bias = bias; p.rowvec = rowvec; p.rowvec_ld = rowvec_ld;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual); p.res_ld = res_ld;
  p.out = out; p.out_ld = out_ld; p.out_fp32 = out_fp32; p.geglu = geglu;
  p.tma_out = (!out_fp32 && block_n % 64 == 0) ? 1 : 0;
  p.ksplit = 1;
  p.kb_per_split = p.num_kb;
  const size_t m_total = static_cast<size_t>(nb) * m_h * w_o;
  if (ksplit > 1) {
    p.kb_per_split = (p.num_kb + ksplit - 1) / ksplit;
    p.ksplit = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;      // no empty splits
    p.split_stride = m_total * n_pad;
    // stage 1 writes raw fp32 partial sums [split][pixel][n_pad]; the epilogue terms move to the reduce kernel
    p.bias = nullptr; p.rowvec = nullptr; p.residual = nullptr;
    p.out = workspace; p.out_ld = n_pad; p.out_fp32 = 1; p.tma_out = 0;
  }
  if (fused_lora) {
    p.lora_n = (lora_rows + 15) / 16 * 16;
    p.lora_kb = c0 / 64;
    p.m_rows = h;
    p.t_out = reinterpret_cast<__nv_bfloat16*>(t_out);
  }
  if (fused_ln) {
    p.ln_stats = ln_stats;
    p.ln_chunks = c0 / 64;
    p.ln_c = c0;
    p.ln_eps = ln_eps;
    p.ln_g = ln_g; p.ln_ga = ln_ga; p.ln_ba = ln_ba;
    p.m_rows = h;
  }
  if (stat_out) {
    B200_CHECK_ARG(p.tma_out && ksplit <= 1 && !geglu && n_valid % 64 == 0 && !fused_ln,
                   "conv_gemm: stat_out needs the plain bf16 TMA-store epilogue");
    cta_pair = 0;
    p.stat_out = stat_out;
    p.stat_chunks = n_valid / 64;
  }
  p.b_dynamic = b_dynamic;
  if (gn_stat) {
    B200_CHECK_ARG(p.tma_out && ksplit <= 1 && !geglu && n_valid % 4 == 0 && !fused_ln && !fused_lora && !stat_out,
                   "conv_gemm: gn_stat needs the plain bf16 TMA-store epilogue");
    B200_CHECK_ARG((w * p.BH) % 32 == 0, "conv_gemm: gn_stat unsupported for %d x %d images (b200_gn_stat_slabs == 0)", h, w);
    p.gn_stat = gn_stat;
    p.gn_slabs = p.tiles_h * (w * p.BH / 32);
  }
  int tc = 32;
  while (tc < 2 * block_n + p.lora_n) tc *= 2;
  p.tmem_cols = tc;
  // 2-CTA pairs: needs two halves of >= 64 weight rows and more than one m-tile to pair up
  // (worth its longer prologue / cluster syncs only when the k-loop dominates)
  const bool cta2 = cta_pair && block_n % 128 == 0 && p.num_m_tiles >= 2 && (cta_pair > 1 || p.num_kb >= 16);
  p.num_m_groups = cta2 ? (p.num_m_tiles + 1) / 2 : p.num_m_tiles;
  const int b_rows = cta2 ? block_n / 2 : block_n;
  const int stage_bytes = kABytes + b_rows * kBlockK * 2;
  p.res_tma = (p.residual != nullptr && p.tma_out) ? 1 : 0;          // (the direct-store path keeps per-thread loads)
  p.epi_bufs = p.res_tma ? 2 : 1;
  const int fixed = kEpiWarps * (p.epi_bufs * kEpiStageBytes + kEpiVecBytes) + 1024 /*align slack*/ + kBarrierBytes +
                    (fused_lora ? kABytes : 0) /*T tile*/;
  // B200_GEMM_SMEM_KB: shared memory this kernel may take (default: all 227 KB).  Less leaves room for the CTAs of the
  // neighbouring kernels in the stream to become resident early (programmatic dependent launch).
  static const int smem_cap = getenv("B200_GEMM_SMEM_KB") ? atoi(getenv("B200_GEMM_SMEM_KB")) * 1024 : kSmemLimit;
  p.stages = ((smem_cap < kSmemLimit ? smem_cap : kSmemLimit) - fixed) / stage_bytes;
  if (p.stages < 2) p.stages = 2;
  if (p.stages > 8) p.stages = 8;
  static const int dbg_stages = getenv("B200_GEMM_STAGES") ? atoi(getenv("B200_GEMM_STAGES")) : 0;
  static const int dbg_flags = getenv("B200_GEMM_DEBUG") ? atoi(getenv("B200_GEMM_DEBUG")) : 0;
  if (dbg_stages > 0 && dbg_stages < p.stages) p.stages = dbg_stages;
  p.debug = dbg_flags;
  const int smem_bytes = p.stages * stage_bytes + fixed;

  CUtensorMap tA[3], tB, tO, tLA, tR;
  const void* srcs[3] = {a0, a1 ? a1 : a0, a2 ? a2 : a0};
  const int chans[3] = {c0, c1 ? c1 : c0, c2 ? c2 : c0};
  p.a_rank2 = 0;     // plain 2-D maps for linear layers were measured: no difference to the 4-D box
  for (int i = 0; i < 3; ++i) {
    const uint64_t C = chans[i];
    int rc;
    if (p.a_rank2) {
      uint64_t dims[2] = {C, (uint64_t)h};
      uint64_t strides[1] = {C};
      uint32_t box[2] = {64, 128};
      rc = make_tmap_bf16(&tA[i], srcs[i], 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    } else {
      uint64_t dims[4] = {C, (uint64_t)w, (uint64_t)h, (uint64_t)nb};
      uint64_t strides[3] = {C, C * w, C * w * h};
      if (s2) {
        // every second pixel in w and h: the box spans 2 w_o x 2 BH input pixels, of which w_o x BH are fetched
        uint32_t box[4] = {64, (uint32_t)(2 * w_o), (uint32_t)(2 * p.BH), (uint32_t)p.BNI};
        uint32_t es[4] = {1, 2, 2, 1};
        rc = make_tmap_bf16_strided(&tA[i], srcs[i], 4, dims, strides, box, es, CU_TENSOR_MAP_SWIZZLE_128B);
      } else {
        uint32_t box[4] = {64, (uint32_t)w, (uint32_t)p.BH, (uint32_t)p.BNI};
        rc = make_tmap_bf16(&tA[i], srcs[i], 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
      }
    }
    if (rc) return rc;
  }
  {
    const uint64_t K = static_cast<uint64_t>(p.num_kb) * 64;
    uint64_t dims[2] = {K, (uint64_t)n_pad};
    uint64_t strides[1] = {K};
    uint32_t box[2] = {64, (uint32_t)b_rows};
    int rc = make_tmap_bf16(&tB, wpacked, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  if (fused_lora) {
    uint64_t dims[2] = {(uint64_t)c0, 64};
    uint64_t strides[1] = {(uint64_t)c0};
    uint32_t box[2] = {64, (uint32_t)p.lora_n};
    int rc = make_tmap_bf16(&tLA, lora_down, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  } else {
    tLA = tB;
  }
  if (p.tma_out) {
    // each epilogue warp stores its 32 accumulator rows x 64 columns: box = (64, bw, bh, bn), w fastest
    const int bw = w_o < 32 ? w_o : 32;
    int bh = 32 / bw;
    if (bh > p.BH) bh = p.BH;
    const int bn = 32 / (bw * bh);
    const uint64_t L = out_ld;
    uint64_t dims[4] = {(uint64_t)n_valid, (uint64_t)w_o, (uint64_t)m_h, (uint64_t)nb};
    uint64_t strides[3] = {L, L * w_o, (c1d && c1d->out_batch_stride) ? (uint64_t)c1d->out_batch_stride : L * w_o * m_h};
    uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
    int rc = make_tmap_bf16(&tO, out, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  } else {
    tO = tB;
  }
  if (p.res_tma) {
    // same boxes as the output store, over the residual tensor
    const int bw = w_o < 32 ? w_o : 32;
    int bh = 32 / bw;
    if (bh > p.BH) bh = p.BH;
    const int bn = 32 / (bw * bh);
    const uint64_t L = res_ld;
    uint64_t dims[4] = {(uint64_t)n_valid, (uint64_t)w_o, (uint64_t)m_h, (uint64_t)nb};
    uint64_t strides[3] = {L, L * w_o, L * w_o * m_h};
    uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
    int rc = make_tmap_bf16(&tR, residual, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  } else {
    tR = tB;
  }

  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(conv_gemm_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    cudaFuncSetAttribute(conv_gemm_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    cudaFuncSetAttribute(conv_gemm_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    cudaFuncSetAttribute(conv_gemm_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    cudaFuncSetAttribute(conv_gemm_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    cudaFuncSetAttribute(conv_gemm_kernel<false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    cudaFuncSetAttribute(conv_gemm_kernel<false, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    cudaFuncSetAttribute(conv_gemm_kernel<false, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    cudaFuncSetAttribute(conv_gemm_kernel<true, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    cudaFuncSetAttribute(conv_gemm_kernel<false, false, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    cudaFuncSetAttribute(conv_gemm_kernel<true, false, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
  }
  int grid = p.num_m_groups * p.num_n_tiles * p.ksplit;
  int cap = max_ctas > 0 ? max_ctas : num_sms;
  if (cta2) {
    if (grid > cap / 2) grid = cap / 2;
    if (grid < 1) grid = 1;
    if (c1d)
      B200_CHECK_PDL("conv1d(2-CTA)", launch_pdl(conv_gemm_kernel<true, false, false, false, false, true>, dim3(2 * grid),
                                                 dim3(kThreads), (size_t)smem_bytes, stream, 2, tA[0], tA[1], tA[2], tB, tO, tLA, tR, p));
    else if (gn_stat)
      B200_CHECK_PDL("conv_gemm_gnstat(2-CTA)", launch_pdl(conv_gemm_kernel<true, false, false, false, true>, dim3(2 * grid),
                                                           dim3(kThreads), (size_t)smem_bytes, stream, 2, tA[0], tA[1], tA[2], tB,
                                                           tO, tLA, tR, p));
    else
      B200_CHECK_PDL("conv_gemm(2-CTA)", launch_pdl(conv_gemm_kernel<true, false>, dim3(2 * grid), dim3(kThreads), (size_t)smem_bytes,
                                                    stream, 2, tA[0], tA[1], tA[2], tB, tO, tLA, tR, p));
  } else {
    if (grid > cap) grid = cap;
    if (c1d)
      B200_CHECK_PDL("conv1d", launch_pdl(conv_gemm_kernel<false, false, false, false, false, true>, dim3(grid), dim3(kThreads),
                                               (size_t)smem_bytes, stream, 0, tA[0], tA[1], tA[2], tB, tO, tLA, tR, p));
    else if (gn_stat)
      B200_CHECK_PDL("conv_gemm_gnstat", launch_pdl(conv_gemm_kernel<false, false, false, false, true>, dim3(grid), dim3(kThreads),
                                                    (size_t)smem_bytes, stream, 0, tA[0], tA[1], tA[2], tB, tO, tLA, tR, p));
    else if (stat_out && !fused_ln && fused_lora)
      B200_CHECK_PDL("linear_stats(lora)", launch_pdl(conv_gemm_kernel<false, true, false, true>, dim3(grid), dim3(kThreads),
                                                      (size_t)smem_bytes, stream, 0, tA[0], tA[1], tA[2], tB, tO, tLA, tR, p));
    else if (stat_out && !fused_ln)
      B200_CHECK_PDL("linear_stats", launch_pdl(conv_gemm_kernel<false, false, false, true>, dim3(grid), dim3(kThreads),
                                                (size_t)smem_bytes, stream, 0, tA[0], tA[1], tA[2], tB, tO, tLA, tR, p));
    else if (fused_lora && fused_ln)
      B200_CHECK_PDL("linear_ln(lora)", launch_pdl(conv_gemm_kernel<false, true, true>, dim3(grid), dim3(kThreads),
                                                   (size_t)smem_bytes, stream, 0, tA[0], tA[1], tA[2], tB, tO, tLA, tR, p));
    else if (fused_ln)
      B200_CHECK_PDL("linear_ln", launch_pdl(conv_gemm_kernel<false, false, true>, dim3(grid), dim3(kThreads),
                                             (size_t)smem_bytes, stream, 0, tA[0], tA[1], tA[2], tB, tO, tLA, tR, p));
    else if (fused_lora)
      B200_CHECK_PDL("linear_lora", launch_pdl(conv_gemm_kernel<false, true>, dim3(grid), dim3(kThreads), (size_t)smem_bytes,
                                               stream, 0, tA[0], tA[1], tA[2], tB, tO, tLA, tR, p));
    else
      B200_CHECK_PDL("conv_gemm", launch_pdl(conv_gemm_kernel<false, false>, dim3(grid), dim3(kThreads), (size_t)smem_bytes,
                                             stream, 0, tA[0], tA[1], tA[2], tB, tO, tLA, tR, p));
  }
  if (p.ksplit > 1) {
    const size_t total = m_total * (n_valid / 8);
    int rgrid = static_cast<int>((total + 255) / 256);
    if (rgrid > num_sms * 8) rgrid = num_sms * 8;
    B200_CHECK_PDL("conv_gemm(split-K reduce)",
                   launch_pdl(splitk_reduce_kernel, dim3(rgrid), dim3(256), 0, stream, 0, workspace, p.ksplit,
                              p.split_stride, (int)m_total, n_valid, n_pad, bias, rowvec, rowvec_ld, m_h * w_o,
                              reinterpret_cast<const __nv_bfloat16*>(residual), res_ld,
                              reinterpret_cast<__nv_bfloat16*>(out), out_ld));
  }
  return B200_OK;
}
