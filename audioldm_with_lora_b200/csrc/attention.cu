// Fused multi-head self-attention over the latent sequence for sm_100a (K2).
//
//   out[b, s, h*d:(h+1)*d] = softmax(Q_h K_h^T * scale) V_h      (non-causal, no mask, no dropout)
//
// One CTA = 128 query rows of one (batch, head).  Per 128-key block:
//   S = Q K^T     tcgen05.mma 128x128x16 (x d/16), operands in smem, S in TMEM
//   softmax       4 warps, one query row per thread, ONE pass over S: p = exp2(s*c - m_ref*c) against a lazily
//                 updated reference maximum m_ref (exact: numerator and denominator share the reference; the
//                 reference is only moved, and O rescaled, when a row's maximum grows by more than 2^8), two
//                 probabilities per MUFU op (ex2.approx.ftz.bf16x2 -- P is needed in bf16 anyway), P written to
//                 smem in the UMMA K-major core-matrix layout
//   O += P [V|1]  tcgen05.mma 128 x (d+16) x 16 (x 8) accumulating in TMEM across key blocks; V is consumed
//                 straight from the [Q|K|V] GEMM output as an MN-major operand (no transposed copy), followed in
//                 smem by 16 columns of ones: O[:, d] is the softmax denominator, summed by the tensor core
//                 from exactly the bf16 probabilities the numerator uses.
// With d = 32 the kernel is bound by the softmax instruction stream (exp throughput), not by the tensor pipe:
// the structure above removes the second pass over S, the per-block O fold and the row-sum adds.
//
// Forms (template parameters; the host picks per head_dim / sequence, profiles/r02_attn_pipe.md):
//   PIPE  software-pipelined: two S buffers in TMEM (and two P buffers), QK^T issued two key blocks ahead of P V, half as
//         many keys per block (same TMEM columns, same CTAs per SM).  Default for head_dim 32 and >= 64.
//   TS    head_dim 32 / 64: Q, K AND V arrive as whole swizzled rows (V = the canonical MN-major swizzled operand), the
//         bf16 probabilities go back to TMEM (tcgen05.st) and P V takes its A operand from there, the denominator is an
//         fp32 register sum: no P tile and no ones columns in shared memory.
// Q/K/V tiles are fetched by TMA directly from the fused-QKV activation [B, S, 3C] with a 4-D tensor
// map (8 elems, rows, 16-byte channel chunks, batch): the box lands as 8x16B core matrices, i.e. the
// no-swizzle UMMA canonical layout, for any head_dim that is a multiple of 16 (32/48/80 for
// AudioLDM-S, 64/96/160 for -L).  Rows past the end of the sequence are zero-filled by TMA and
// masked to -inf in the softmax.
//
// Replaces F.scaled_dot_product_attention in diffusers AttnProcessor2_0.__call__ (attention-processor
// API; reached from /root/reference/script/train/train_audioldm_lora.py:539-546, app.py:14).
#include "host_util.h"
#include "ptx.cuh"

#ifndef B200_ATTN_NOEXP
#define B200_ATTN_NOEXP 0
#endif
#ifndef B200_ATTN_NOSTORE
#define B200_ATTN_NOSTORE 0
#endif
#ifndef B200_ATTN_NOPV
#define B200_ATTN_NOPV 0
#endif

namespace b200 {

static constexpr int kQ = 128;                 // query rows per CTA
static constexpr int kAttnThreads = 192;       // warps 0-3 softmax, warp 4 TMA, warp 5 MMA
// Keys per block: template parameter KV (128, or 64 for small head_dim: S (64 columns) + O (d + 16) then fit 128 TMEM
// columns and ~44 KB of shared memory, so FOUR CTAs share an SM instead of two and hide each other's
// MMA -> softmax -> MMA hand-offs; the kernel is bound by those hand-offs, not by any pipe).
// Behind every V tile: 16 columns of bf16 1.0 (MN-major: 2 chunks of KV x 16 B).
static constexpr float kRescaleThreshold = 8.0f;   // move the reference maximum when a row maximum exceeds it by 2^8

struct AttnParams {
  int seq, heads, d, batch;
  int nblk;                 // ceil(seq / 128)
  int stages;               // K/V ring depth (1 .. 4)
  int tmem_cols;            // 256 (d + 16 <= 128) or 512
  float scale_log2;         // scale * log2(e)
  __nv_bfloat16* out;
  int out_ld;               // heads * d
  int variant;              // bit0/bit1: descriptor-convention debug knobs; bit2: flips the packed / fp32 exp2 choice;
                            // bit3 / bit4 (host side): force the single-buffered / the pipelined form; bit5 / bit6: P through smem / TMEM
  float* lse;               // optional [batch, heads, seq] fp32: log2-domain log-sum-exp of the scaled scores (training)
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t ex2_bf16x2(uint32_t x) {
  uint32_t y;
  asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  uint32_t y;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(y) : "r"(a), "r"(b));
  return y;
}

// max over one KV-column S row in TMEM (two 32-column loads in flight at a time)
template <bool kMasked, int KV>
__device__ __forceinline__ float row_max(uint32_t t_row, int kvalid) {
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
  if constexpr (KV == 32) {
    uint32_t r0[32];
    tmem_ld_x32(t_row, r0);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float a[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = __uint_as_float(r0[i + u]);
        if (kMasked && i + u >= kvalid) a[u] = -INFINITY;
      }
      m0 = fmaxf(m0, a[0]); m1 = fmaxf(m1, a[1]); m2 = fmaxf(m2, a[2]); m3 = fmaxf(m3, a[3]);
    }
    return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
  } else {
#pragma unroll
  for (int c = 0; c < KV / 32; c += 2) {
    uint32_t r0[32], r1[32];
    tmem_ld_x32(t_row + c * 32, r0);
    tmem_ld_x32(t_row + (c + 1) * 32, r1);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      float a0 = __uint_as_float(r0[i]), a1 = __uint_as_float(r0[i + 1]);
      float b0 = __uint_as_float(r1[i]), b1 = __uint_as_float(r1[i + 1]);
      if (kMasked) {
        if (c * 32 + i >= kvalid) a0 = -INFINITY;
        if (c * 32 + i + 1 >= kvalid) a1 = -INFINITY;
        if ((c + 1) * 32 + i >= kvalid) b0 = -INFINITY;
        if ((c + 1) * 32 + i + 1 >= kvalid) b1 = -INFINITY;
      }
      m0 = fmaxf(m0, a0); m1 = fmaxf(m1, a1); m2 = fmaxf(m2, b0); m3 = fmaxf(m3, b1);
    }
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
  }
}

// One pass over a KV-column S row: p = exp2(s * scale - ref) as bf16 into the K-major core-matrix smem tile.
// Returns max_j (s_j * scale - ref) (bf16 precision: it only feeds the rescale decision).
// The next 32-column TMEM load is issued before the current one is processed.
template <bool kMasked, bool kPackedExp, int KV>
__device__ __forceinline__ float exp_store(uint32_t t_row, uint8_t* sp_row, float scale_log2, float neg_ref, int kvalid) {
  uint32_t mx = 0xFF80FF80u;                      // (-inf, -inf) in bf16x2
  uint32_t r[2][32];
  tmem_ld_x32(t_row, r[0]);
#pragma unroll
  for (int c = 0; c < KV / 32; ++c) {
    tmem_wait_ld();
    if (c + 1 < KV / 32) tmem_ld_x32(t_row + (c + 1) * 32, r[(c + 1) & 1]);
    const uint32_t(&cur)[32] = r[c & 1];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint32_t pk[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int col = c * 32 + g * 8 + 2 * i;
        float x0 = fmaf(__uint_as_float(cur[g * 8 + 2 * i]), scale_log2, neg_ref);
        float x1 = fmaf(__uint_as_float(cur[g * 8 + 2 * i + 1]), scale_log2, neg_ref);
        if (kMasked) {
          if (col >= kvalid) x0 = -INFINITY;
          if (col + 1 >= kvalid) x1 = -INFINITY;
        }
        const uint32_t x = pack_bf16x2(x0, x1);
        mx = max_bf16x2(mx, x);
#if B200_ATTN_NOEXP      // limiter experiment (tools/build_variant.sh; results are garbage): no MUFU work
        pk[i] = x;
#else
        pk[i] = kPackedExp ? ex2_bf16x2(x) : pack_bf16x2(ex2_approx(x0), ex2_approx(x1));
#endif
      }
#if B200_ATTN_NOSTORE    // limiter experiment: a quarter of the P stores (the P V product still reads the whole tile)
      if (g == 0)
#endif
      *reinterpret_cast<uint4*>(sp_row + (c * 4 + g) * 2048) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
  return fmaxf(bf16_lo(mx), bf16_hi(mx));
}

// TMEM form of the same pass (kernel template TS): the bf16 probabilities go back to tensor memory (two per 32-bit column:
// the layout of a K-major A operand read from TMEM) instead of shared memory, and the block's row sum is kept in fp32
// registers.  Per 64-key block this takes 16 KB of stores and 16 KB of tensor-core reads off the SM's shared-memory port,
// which -- not the MUFU pipe, not the copy engine -- is what the kernel ran into (profiles/r02_attn_pipe.md).
template <bool kMasked, int KV>
__device__ __forceinline__ float exp_store_tmem(uint32_t t_row, uint32_t t_prow, float scale_log2, float neg_ref, int kvalid,
                                                float& block_sum) {
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  uint32_t r[2][32];
  tmem_ld_x32(t_row, r[0]);
#pragma unroll
  for (int c = 0; c < KV / 32; ++c) {
    tmem_wait_ld();
    if (c + 1 < KV / 32) tmem_ld_x32(t_row + (c + 1) * 32, r[(c + 1) & 1]);
    const uint32_t(&cur)[32] = r[c & 1];
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int col = c * 32 + 2 * i;
      float x0 = fmaf(__uint_as_float(cur[2 * i]), scale_log2, neg_ref);
      float x1 = fmaf(__uint_as_float(cur[2 * i + 1]), scale_log2, neg_ref);
      if (kMasked) {
        if (col >= kvalid) x0 = -INFINITY;
        if (col + 1 >= kvalid) x1 = -INFINITY;
      }
      m0 = fmaxf(m0, x0);
      m1 = fmaxf(m1, x1);
      const float e0 = ex2_approx(x0), e1 = ex2_approx(x1);
      l0 += e0;
      l1 += e1;
      pk[i] = pack_bf16x2(e0, e1);
    }
    tmem_st_x16(t_prow + c * 16, pk);
  }
  block_sum = l0 + l1;
  return fmaxf(m0, m1);
}

// O (TMEM, ncols fp32 columns of this thread's row) *= f
__device__ __forceinline__ void scale_o(uint32_t t_o_row, int ncols, float f) {
  for (int c = 0; c < ncols / 16; ++c) {
    uint32_t r[16];
    tmem_ld_x16(t_o_row + c * 16, r);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * f);
    tmem_st_x16(t_o_row + c * 16, r);
  }
  tmem_wait_st();
}

// -DB200_ATTN_PROFILE=1 builds only (tools/build_variant.sh attnprof -DB200_ATTN_PROFILE=1; tools/attn_timeline.py): cycle
// stamps of CTA (0, 0), printed by the kernel itself
#if B200_ATTN_PROFILE
__device__ long long s_tl[48];
#define ATL(i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) s_tl[i] = clock64(); } while (0)
#else
#define ATL(i) do {} while (0)
#endif

template <int D, int KV, bool PIPE, bool TS = false>
__global__ void __launch_bounds__(kAttnThreads)
attention_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmKV,
                 const __grid_constant__ CUtensorMap tmKs, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // Q and K (the K-major operands of S = Q K^T) as whole rows of D * 2 bytes in the 64- / 128-byte swizzled layout when a
  // row IS a swizzle span (head_dim 32 / 64): one TMA piece per row instead of D / 8 (the copy engine moves about one
  // piece per clock per SM, and the key blocks were waiting for it: profiles/r02_attn_timeline.md).  V stays in the
  // 16-byte core-matrix layout (it is the MN-major operand of P V).
  constexpr bool kSwz = (D == 32 || D == 64);
  constexpr uint64_t kSwzMode = D == 64 ? SWZ_128B : SWZ_64B;
  constexpr uint32_t kSwzSbo = 8 * D * 2;       // 8 rows of D * 2 bytes
  constexpr int kKV = KV;
  constexpr int kTileBytes = 128 * D * 2;       // the Q tile
  constexpr int kKVTile = KV * D * 2;           // one K / V tile
  constexpr int kChunk = KV * 16;               // bytes between 8-element chunks of a K / V tile
  // Swizzled form (head_dim 32 / 64): V too arrives as whole swizzled rows -- a [keys][head_dim] tile with 64- / 128-byte
  // rows IS the canonical MN-major swizzled operand layout (8 keys x one swizzle span per atom) -- and the denominator is a
  // second, 16-column MMA per key step against ONE static block of ones.  (With V in 16-byte pieces the kernel could not
  // drop below 40 us at s1000 d32 even with the exponentials, the P stores and the P V products removed: the copy engine
  // delivers about one piece per clock per SM; gpurun_out/r02_attn_limiter.log.)
  constexpr int kOnesBytes = kSwz ? 0 : 2 * kChunk;          // per stage (no-swizzle form: the ones columns follow the V tile)
  constexpr int kOnesStatic = 512;                            // swizzled form: 16 keys x 16 columns of bf16 1.0
  constexpr int kPBytes = kQ * kKV * 2;
  constexpr int kStageBytes = 2 * kKVTile + kOnesBytes;         // K tile, V tile, ones columns
  constexpr int kOCols = D + 16;                // O accumulator columns: d outputs + the denominator (x16)
  // PIPE: software-pipelined form.  S (TMEM) and P (smem) are double-buffered, so the MMA warp issues QK^T of block
  // j + 2 as soon as the softmax has finished READING S of block j, and P V of block j while the softmax works on block
  // j + 1: the softmax warps never wait for a tensor-core hand-off (the per-CTA chain QK^T -> softmax -> P V -> QK^T was
  // what paced the single-buffered form: ~1000 of every 2400 cycles per 64 keys; profiles/r02_attn_timeline.md).  With
  // half the keys per block the two S buffers take the TMEM columns of the single one, so the CTAs per SM stay.
  constexpr bool kPipe = PIPE;
  constexpr int kBufs = kPipe ? 2 : 1;
  // TS (swizzled form only): P lives in TMEM (KV / 2 columns per buffer, behind the S buffers) and is the tensor core's A
  // operand from there; the denominator is summed in registers, so O is just the head_dim output columns.
  constexpr bool kTS = TS;
  static_assert(!TS || kSwz, "the TMEM-P form is built for the swizzled V layout (head_dim 32 / 64)");
  constexpr int kPCols = KV / 2;
  uint8_t* sQ = smem;
  uint8_t* sP = sQ + kTileBytes;                // kBufs x P tile (not in the TS form)
  uint8_t* sKV = sP + (kTS ? 0 : kBufs * kPBytes);          // stages x {K, V, ones}
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + p.stages * kStageBytes);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                 // [4]
  uint64_t* kv_empty = bars + 5;                // [4]
  uint64_t* s_full = bars + 9;                  // [2]: per S buffer
  uint64_t* p_full = bars + 11;                 // [2]: per P buffer
  uint64_t* o_full = bars + 13;                 // [2]: P V of the block that used P buffer i has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0) ATL(0);
  const int q0 = blockIdx.x * kQ;
  const int bh = blockIdx.y;
  const int b = bh / p.heads;
  const int h = bh % p.heads;
  const int C8 = p.heads * D / 8;               // 16-byte chunks per Q (or K, or V) section
  const int chunk_q = h * (D / 8);
  const int chunk_k = C8 + chunk_q;
  const int chunk_v = 2 * C8 + chunk_q;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmKV);
    if (kSwz) tma_prefetch_desc(&tmKs);
    mbar_init(q_full, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  uint8_t* sOnes = reinterpret_cast<uint8_t*>(bars) + 128;
  if constexpr (kTS) {
  } else if constexpr (kSwz) {
    for (int i = threadIdx.x; i < kOnesStatic / 16; i += kAttnThreads)
      *reinterpret_cast<uint4*>(sOnes + i * 16) = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  } else {
    // the ones columns behind every V tile (never overwritten: the TMA box covers the V tile only)
    for (int i = threadIdx.x; i < p.stages * (kOnesBytes / 16); i += kAttnThreads) {
      const int st = i / (kOnesBytes / 16), off = i % (kOnesBytes / 16);
      *reinterpret_cast<uint4*>(sKV + st * kStageBytes + 2 * kKVTile + off * 16) =
          make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    }
  }
  fence_proxy_async_smem();
  if (warp == 4) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) ATL(1);
  pdl_launch_dependents();
  pdl_wait();
  if (warp == 0) ATL(2);
  const uint32_t t_s = tmem_base;               // S: columns [0, KV) (pipelined form: two buffers, [0, 2 KV))
  const uint32_t t_p = tmem_base + kBufs * KV;  // TS form: P buffers (KV / 2 columns each)
  const uint32_t t_o = tmem_base + kBufs * KV + (kTS ? kBufs * kPCols : 0);  // O: the D + 16 (TS: D) columns behind S (and P)

  if (warp == 4) {
    // ============================================================ TMA producer
    // (the whole warp runs the loop with uniform control flow and one lane, elected at each use, issues: operands
    //  stay in uniform registers instead of an ELECT + R2UR.BROADCAST waterfall in front of every TMA / MMA)
    {
      if (elect_one()) {
        mbar_expect_tx(q_full, kTileBytes);
        if (kSwz) tma_load_3d(sQ, &tmQKV, q_full, h * D, q0, b);
        else tma_load_4d(sQ, &tmQKV, q_full, 0, q0, chunk_q, b);
      }
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < p.nblk; ++j) {
        mbar_wait(&kv_empty[s], ph ^ 1);
        uint8_t* k_dst = sKV + s * kStageBytes;
        if (elect_one()) {
          mbar_expect_tx(&kv_full[s], 2 * kKVTile);
          if (kSwz) {
            tma_load_3d(k_dst, &tmKs, &kv_full[s], p.heads * D + h * D, j * kKV, b);
            tma_load_3d(k_dst + kKVTile, &tmKs, &kv_full[s], 2 * p.heads * D + h * D, j * kKV, b);
          } else {
            tma_load_4d(k_dst, &tmKV, &kv_full[s], 0, j * kKV, chunk_k, b);
            tma_load_4d(k_dst + kKVTile, &tmKV, &kv_full[s], 0, j * kKV, chunk_v, b);
          }
        }
        if (++s == p.stages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 5) {
    // ============================================================ MMA issuer (whole warp, elected lane issues)
    {
      const uint32_t idesc_s = make_idesc_bf16(128, KV, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(128, kSwz ? D : kOCols, 0, 1);  // B = [V | 1] (swizzled form: V), MN-major
      const uint32_t idesc_1 = make_idesc_bf16(128, 16, 0, 1);      // swizzled form: B = the static ones block
      const uint32_t ones_addr = smem_u32(sOnes);
      // no-swizzle canonical layouts: core matrix = 8 rows x 16 B, contiguous (128 B).
      //  K-major tile [R rows][D]: next 8-row group +128 B (SBO), next 8-elem K chunk +R*16 B (LBO); R = 128 (Q, P) or KV (K)
      //  MN-major V   [KV keys][D]: next 8-key group +128 B (LBO), next 8-elem d chunk +KV*16 B (SBO)
      const uint32_t q_lbo = (p.variant & 1) ? 128 : 2048, q_sbo = (p.variant & 1) ? 2048 : 128;
      const uint32_t k_lbo = (p.variant & 1) ? 128 : kChunk, k_sbo = (p.variant & 1) ? kChunk : 128;
      const uint32_t v_lbo = (p.variant & 2) ? kChunk : 128, v_sbo = (p.variant & 2) ? 128 : kChunk;
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t p_addr = smem_u32(sP);
      mbar_wait(q_full, 0);
      // S_buf = Q K_blk^T for the key block in ring stage `st` (whose kv_full phase is `phs`)
      auto issue_qk = [&](int st, uint32_t phs, int buf) {
        const uint32_t k_addr = smem_u32(sKV + st * kStageBytes);
        mbar_wait(&kv_full[st], phs);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < D / 16; ++k) {
            // swizzled rows: a 16-element K step is +32 bytes inside the swizzle span (as in conv_gemm.cu)
            const uint64_t a_desc = kSwz ? make_smem_desc(q_addr + k * 32, 16, kSwzSbo, kSwzMode)
                                         : make_smem_desc(q_addr + k * 4096, q_lbo, q_sbo, SWZ_NONE);
            const uint64_t b_desc = kSwz ? make_smem_desc(k_addr + k * 32, 16, kSwzSbo, kSwzMode)
                                         : make_smem_desc(k_addr + k * 2 * kChunk, k_lbo, k_sbo, SWZ_NONE);
            umma_bf16_ss(t_s + buf * KV, a_desc, b_desc, idesc_s, k != 0);
          }
          umma_commit(&s_full[buf]);
        }
      };
      // O += P_buf [V | 1] with the V tile of ring stage `st`; frees the stage
      auto issue_pv = [&](int st, int buf, int j) {
        const uint32_t v_addr = smem_u32(sKV + st * kStageBytes) + kKVTile;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < (B200_ATTN_NOPV ? 1 : kKV / 16); ++k) {      // NOPV: limiter experiment (garbage results)
            const uint64_t a_desc = make_smem_desc(p_addr + buf * kPBytes + k * 4096, q_lbo, q_sbo, SWZ_NONE);
            if (kTS) {
              // A = P from TMEM: 16 keys = 8 columns of bf16 pairs
              const uint64_t b_desc = make_smem_desc(v_addr + k * 16 * D * 2, kSwzSbo, kSwzSbo, kSwzMode);
              umma_bf16_ts(t_o, t_p + buf * kPCols + k * 8, b_desc, idesc_o, (j | k) != 0);
            } else if (kSwz) {
              // MN-major swizzled V: 16 keys = two 8-row atoms of D * 2-byte rows (SBO apart); head_dim is one swizzle span
              const uint64_t b_desc = make_smem_desc(v_addr + k * 16 * D * 2, kSwzSbo, kSwzSbo, kSwzMode);
              umma_bf16_ss(t_o, a_desc, b_desc, idesc_o, (j | k) != 0);
              const uint64_t o_desc = make_smem_desc(ones_addr, 128, 256, SWZ_NONE);
              umma_bf16_ss(t_o + D, a_desc, o_desc, idesc_1, (j | k) != 0);
            } else {
              const uint64_t b_desc = make_smem_desc(v_addr + k * 256, v_lbo, v_sbo, SWZ_NONE);
              umma_bf16_ss(t_o, a_desc, b_desc, idesc_o, (j | k) != 0);
            }
          }
          umma_commit(&o_full[buf]);
          umma_commit(&kv_empty[st]);
        }
      };
      if (kPipe) {
        // ring cursors: `sq` walks the blocks whose QK^T is issued (two ahead), `sv` the blocks whose P V is issued
        int sq = 0, sv = 0;
        uint32_t phq = 0;
        auto adv = [&](int& st, uint32_t& phs) { if (++st == p.stages) { st = 0; phs ^= 1; } };
        uint32_t phv_unused = 0;
        for (int jj = 0; jj < 2 && jj < p.nblk; ++jj) {
          issue_qk(sq, phq, jj);
          if (jj < 4) ATL(8 + jj);
          adv(sq, phq);
        }
        for (int j = 0; j < p.nblk; ++j) {
          const int buf = j & 1;
          mbar_wait(&p_full[buf], (j >> 1) & 1);             // P_j written (and S_buf read, O rescaled if needed)
          tc_fence_after();
          issue_pv(sv, buf, j);
          adv(sv, phv_unused);
          if (j + 2 < p.nblk) {
            issue_qk(sq, phq, buf);                          // S_buf is free: the softmax of block j has read it
            if (j + 2 < 4) ATL(8 + j + 2);
            adv(sq, phq);
          }
        }
      } else {
        int st = 0;
        uint32_t ph = 0;
        for (int j = 0; j < p.nblk; ++j) {
          // S = Q K^T   (the previous block's softmax finished reading S before it released p_full)
          issue_qk(st, ph, 0);
          if (j < 4) ATL(8 + j);                  // K/V block j landed
          // O += P [V | 1]   (the softmax threads rescaled O, if needed, before they released p_full)
          mbar_wait(&p_full[0], j & 1);
          tc_fence_after();
          issue_pv(st, 0, j);
          if (++st == p.stages) {
            st = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else {
    // ============================================================ softmax (row = thread)
    const int row = warp * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    // Exponentials: two per MUFU op in bf16 (ex2.approx.ftz.bf16x2) or one per op in fp32.  Measured on B200 at
    // b16 s1000 d32: with 128-column tiles (2 CTAs/SM) the packed form wins; with 64-column tiles (4 CTAs/SM) the
    // fp32 form does (59.3 vs 63.4 us) and is more accurate (2.3e-3 vs 3.1e-3 rel-L2).  Moving part of the
    // exponentials to an FMA-pipe polynomial (FlashAttention-4) or replacing F2FP by integer packing made this kernel
    // SLOWER (80 / 66 us): ncu shows it bound by the tcgen05.ld -> exp -> st.shared dependency chain of each thread
    // (long-scoreboard + fixed-latency stalls, issue slots 39 % busy), not by the XU pipe.  variant bit 2 flips the default.
    const bool packed = ((p.variant & 4) == 0) != (KV <= 64);
    float ref = 0.f;                              // reference maximum, in exponent units: m_ref * scale * log2(e)
    float lsum = 0.f;                             // TS form: the row's denominator (fp32 sum of the probabilities)
    for (int j = 0; j < p.nblk; ++j) {
      const int kvalid = min(kKV, p.seq - j * kKV);
      const bool full = kvalid == kKV;
      // S / P buffer of this block, and which completion of its barriers belongs to it
      const int buf = kPipe ? (j & 1) : 0;
      const uint32_t par = kPipe ? ((j >> 1) & 1) : (j & 1);
      const uint32_t t_sj = t_s + buf * KV + lane_off;
      uint8_t* sp_row = sP + buf * kPBytes + row * 16;
      mbar_wait(&s_full[buf], par);
      tc_fence_after();
      if (warp == 0 && j < 4) ATL(16 + j);      // S_j complete
      if (j == 0) {
        const float mx = full ? row_max<false, KV>(t_sj, kvalid) : row_max<true, KV>(t_sj, kvalid);
        ref = mx * p.scale_log2;
      }
      if (j >= kBufs) {
        // this P buffer is free again once the P V of the block that used it last (j - kBufs) has completed (single
        // buffer: O is then quiescent as well)
        mbar_wait(&o_full[buf], par ^ 1);
        tc_fence_after();
      }
      float over, bsum = 0.f;
      const uint32_t t_pj = t_p + buf * kPCols + lane_off;
      if (kTS) over = full ? exp_store_tmem<false, KV>(t_sj, t_pj, p.scale_log2, -ref, kvalid, bsum)
                           : exp_store_tmem<true, KV>(t_sj, t_pj, p.scale_log2, -ref, kvalid, bsum);
      else if (packed) over = full ? exp_store<false, true, KV>(t_sj, sp_row, p.scale_log2, -ref, kvalid)
                              : exp_store<true, true, KV>(t_sj, sp_row, p.scale_log2, -ref, kvalid);
      else over = full ? exp_store<false, false, KV>(t_sj, sp_row, p.scale_log2, -ref, kvalid)
                       : exp_store<true, false, KV>(t_sj, sp_row, p.scale_log2, -ref, kvalid);
      // warp-uniform decision (the TMEM loads / stores below are warp-collective)
      if (__any_sync(0xffffffffu, over > kRescaleThreshold)) {
        // rare: a row's maximum moved by more than 2^8 -> every row of the warp takes its current maximum as the
        // new reference, O (with its denominator column) is rescaled, and the block's probabilities are redone
        if (kPipe && j >= 1) {
          // O must be quiescent: the P V of block j - 1 (other buffer) may still be running
          mbar_wait(&o_full[buf ^ 1], ((j - 1) >> 1) & 1);
          tc_fence_after();
        }
        const float mx = full ? row_max<false, KV>(t_sj, kvalid) : row_max<true, KV>(t_sj, kvalid);
        const float new_ref = fmaxf(ref, mx * p.scale_log2);
        const float fscale = ex2_approx(ref - new_ref);
        scale_o(t_o + lane_off, kTS ? D : kOCols, fscale);
        lsum *= fscale;
        ref = new_ref;
        if (kTS) (void)(full ? exp_store_tmem<false, KV>(t_sj, t_pj, p.scale_log2, -ref, kvalid, bsum)
                             : exp_store_tmem<true, KV>(t_sj, t_pj, p.scale_log2, -ref, kvalid, bsum));
        else if (packed) (void)(full ? exp_store<false, true, KV>(t_sj, sp_row, p.scale_log2, -ref, kvalid)
                                : exp_store<true, true, KV>(t_sj, sp_row, p.scale_log2, -ref, kvalid));
        else (void)(full ? exp_store<false, false, KV>(t_sj, sp_row, p.scale_log2, -ref, kvalid)
                         : exp_store<true, false, KV>(t_sj, sp_row, p.scale_log2, -ref, kvalid));
      }
      lsum += bsum;
      if (kTS) tmem_wait_st();
      else fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[buf]);
      if (warp == 0 && j < 4) ATL(24 + j);      // P_j written
      if (warp == 0 && j == p.nblk - 1) ATL(32);
    }
    {
      const int jl = p.nblk - 1;                // the last P V (commits complete in issue order: all earlier ones too)
      mbar_wait(&o_full[kPipe ? (jl & 1) : 0], kPipe ? ((jl >> 1) & 1) : (jl & 1));
    }
    tc_fence_after();
    if (warp == 0) ATL(33);                     // last P V complete
    // epilogue: O[:, :d] / O[:, d]
    const int qrow = q0 + row;
    float den = lsum;
    if (!kTS) {
      uint32_t rl[16];
      tmem_ld_x16(t_o + lane_off + D, rl);
      tmem_wait_ld();
      den = __uint_as_float(rl[0]);
    }
    const float inv = 1.0f / den;
    if (p.lse != nullptr && qrow < p.seq)      // p_ij = exp2(s_ij * scale_log2 - lse): what the backward kernel recomputes
      p.lse[(static_cast<size_t>(b) * p.heads + h) * p.seq + qrow] = ref + log2f(den);
    __nv_bfloat16* o = p.out + (static_cast<size_t>(b) * p.seq + qrow) * p.out_ld + h * D;
#pragma unroll
    for (int c = 0; c < D / 16; ++c) {
      uint32_t r[16];
      tmem_ld_x16(t_o + lane_off + c * 16, r);
      tmem_wait_ld();
      if (qrow < p.seq) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint4 pk;
          pk.x = pack_bf16x2(__uint_as_float(r[g * 8 + 0]) * inv, __uint_as_float(r[g * 8 + 1]) * inv);
          pk.y = pack_bf16x2(__uint_as_float(r[g * 8 + 2]) * inv, __uint_as_float(r[g * 8 + 3]) * inv);
          pk.z = pack_bf16x2(__uint_as_float(r[g * 8 + 4]) * inv, __uint_as_float(r[g * 8 + 5]) * inv);
          pk.w = pack_bf16x2(__uint_as_float(r[g * 8 + 6]) * inv, __uint_as_float(r[g * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(o + c * 16 + g * 8) = pk;
        }
      }
    }
    tc_fence_before();
  }

  if (warp == 0) ATL(34);                       // output rows stored
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
#if B200_ATTN_PROFILE
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    const long long t0 = s_tl[0];
    printf("attn<%d,%d> nblk %d stages %d: prologue %lld  pdl_wait %lld | kv0 %lld kv1 %lld kv2 %lld kv3 %lld | S0 %lld S1 %lld S2 %lld S3 %lld | "
           "P0 %lld P1 %lld P2 %lld P3 %lld | lastP %lld lastPV %lld stored %lld end %lld\n", D, KV, p.nblk, p.stages, s_tl[1] - t0, s_tl[2] - t0,
           s_tl[8] - t0, s_tl[9] - t0, s_tl[10] - t0, s_tl[11] - t0, s_tl[16] - t0, s_tl[17] - t0, s_tl[18] - t0, s_tl[19] - t0,
           s_tl[24] - t0, s_tl[25] - t0, s_tl[26] - t0, s_tl[27] - t0, s_tl[32] - t0, s_tl[33] - t0, s_tl[34] - t0,
           (long long)clock64() - t0);
  }
#endif
}

template <int D, int KV, bool PIPE, bool TS = false>
static int launch_attention(const CUtensorMap& tm_pieces, const CUtensorMap& tmkv, const void* qkv, AttnParams& p, cudaStream_t stream) {
  constexpr int kTileBytes = 128 * D * 2;
  constexpr bool kSwz = (D == 32 || D == 64);
  constexpr int kStageBytes = 2 * KV * D * 2 + (kSwz ? 0 : 2 * KV * 16);
  constexpr int kMinStages = PIPE ? 2 : 1;      // the pipelined form issues two QK^T before the first P V frees a stage
  const int fixed = kTileBytes + (TS ? 0 : (PIPE ? 2 : 1) * kQ * KV * 2) + 128 /*barriers*/ + (kSwz ? 512 : 0) /*ones*/ + 1024 /*align*/;
  // resident CTAs per SM are set by TMEM: 512 / tmem_cols (4, 2 or 1); give each its share of shared memory.  The K/V
  // ring wants >= 2 stages: with one, the next block's loads start only after the current block's P V has completed (the
  // CTA timeline showed exactly that at head_dim 48: profiles/r02_attn_timeline.md), so a fourth co-resident CTA is given
  // up when its share of shared memory would leave a single stage.
  int per_sm = 512 / p.tmem_cols;
  int budget = (220 * 1024) / per_sm;
  while (per_sm > 1 && fixed + (per_sm >= 4 ? 2 : kMinStages) * kStageBytes > budget) {
    --per_sm;
    budget = (220 * 1024) / per_sm;
  }
  p.stages = (budget - fixed) / kStageBytes;
  if (p.stages > 4) p.stages = 4;
  if (p.stages < kMinStages) p.stages = kMinStages;
  static const int dbg_stages = getenv("B200_ATTN_STAGES") ? atoi(getenv("B200_ATTN_STAGES")) : 0;       // A/B knob
  if (dbg_stages >= kMinStages && dbg_stages < p.stages) p.stages = dbg_stages;
  const int smem_bytes = fixed + p.stages * kStageBytes;
  if (smem_bytes > 227 * 1024) return fail(B200_ERR_UNSUPPORTED, "attention: %d bytes of shared memory needed", smem_bytes);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel<D, KV, PIPE, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail(B200_ERR_CUDA, "attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  // head_dim 32 / 64: Q and K tiles as whole rows (rank-3 maps over [3C, seq, batch], 64- / 128-byte swizzle)
  CUtensorMap tm = tm_pieces, tmks = tmkv;
  if (D == 32 || D == 64) {
    const uint64_t C3 = static_cast<uint64_t>(3) * p.heads * D;
    uint64_t dims[3] = {C3, (uint64_t)p.seq, (uint64_t)p.batch};
    uint64_t strides[2] = {C3, (uint64_t)p.seq * C3};
    const CUtensorMapSwizzle swz = D == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    uint32_t box_q[3] = {(uint32_t)D, 128, 1}, box_k[3] = {(uint32_t)D, (uint32_t)KV, 1};
    int rc = make_tmap_bf16(&tm, qkv, 3, dims, strides, box_q, swz);
    if (rc) return rc;
    rc = make_tmap_bf16(&tmks, qkv, 3, dims, strides, box_k, swz);
    if (rc) return rc;
  }
  dim3 grid((p.seq + kQ - 1) / kQ, p.batch * p.heads);
  B200_CHECK_PDL("attention", launch_pdl(attention_kernel<D, KV, PIPE, TS>, grid, dim3(kAttnThreads), (size_t)smem_bytes, stream, 0,
                                         tm, tmkv, tmks, p));
  return B200_OK;
}

}  // namespace b200

using namespace b200;

static int attention_impl(const void* qkv, void* out, float* lse, int batch, int seq, int heads, int head_dim, float scale,
                          int variant, void* stream_v);

extern "C" int b200_attention(const void* qkv, void* out, int batch, int seq, int heads, int head_dim, float scale,
                              int variant, void* stream_v) {
  return attention_impl(qkv, out, nullptr, batch, seq, heads, head_dim, scale, variant, stream_v);
}
// Training form: also writes the per-row log-sum-exp (log2 domain, scale folded in) the backward kernel needs.
extern "C" int b200_attention_lse(const void* qkv, void* out, float* lse, int batch, int seq, int heads, int head_dim,
                                  float scale, void* stream_v) {
  B200_CHECK_ARG(lse, "attention_lse: null lse");
  return attention_impl(qkv, out, lse, batch, seq, heads, head_dim, scale, 0, stream_v);
}

static int attention_impl(const void* qkv, void* out, float* lse, int batch, int seq, int heads, int head_dim, float scale,
                          int variant, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(qkv && out, "attention: null pointer");
  B200_CHECK_ARG(batch > 0 && seq > 0 && heads > 0, "attention: empty problem");
  const int C = heads * head_dim;
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.seq = seq; p.heads = heads; p.d = head_dim; p.batch = batch;
  // Keys per block.  Single-buffered form: 64 for head_dim <= 48, else 128 (B200_ATTN_KV=128 forces 128).  Pipelined
  // form: two S / P buffers of HALF as many keys -- the same TMEM columns, so the same CTAs per SM (B200_ATTN_PIPE=2: two
  // buffers of the single-buffered width; 0: never; 1: always).  Default (measured on B200, tools/attn_ab.py,
  // profiles/r02_attn_pipe.md): pipelined for head_dim 32 and >= 64 when there is more than one single-buffered block
  // (b32 s3000 d64: 1380 -> 956 us, s188 d160: 60.6 -> 41.7 us, s752 d96 equal; b16 s1000 d32: 58.0 -> 54.0 us together
  // with P through TMEM); single-buffered for head_dim 48 (three co-resident CTAs already overlap each other's hand-offs:
  // 12.2 vs 12.0 us) and for one-block problems.
  static const int kv_env = getenv("B200_ATTN_KV") ? atoi(getenv("B200_ATTN_KV")) : 0;     // A/B knob: 64 or 128
  static const int pipe_env = getenv("B200_ATTN_PIPE") ? atoi(getenv("B200_ATTN_PIPE")) : -1;
  const int kv1 = (head_dim <= 48 && kv_env != 128) ? 64 : 128;
  const int pipe = (variant & 8) ? 0 : (variant & 16) ? 1 : pipe_env >= 0 ? pipe_env : ((head_dim >= 64 || head_dim == 32) && seq > kv1) ? 1 : 0;
  const int kv = pipe == 1 ? kv1 / 2 : kv1;
  // TS form (head_dim 32 / 64): P through TMEM instead of shared memory (B200_ATTN_TS=0 / variant bit 5 turn it off)
  static const int ts_env = getenv("B200_ATTN_TS") ? atoi(getenv("B200_ATTN_TS")) : 1;
  const bool ts = (head_dim == 32 || head_dim == 64) && pipe != 2 && ((variant & 64) || (ts_env && !(variant & 32)));
  p.nblk = (seq + kv - 1) / kv;
  int cols = 32;
  while (cols < (ts ? (pipe ? 2 : 1) * (kv + kv / 2) + head_dim : (pipe ? 2 : 1) * kv + head_dim + 16)) cols *= 2;
  p.tmem_cols = cols;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.out_ld = C;
  p.variant = variant;
  p.lse = lse;

  CUtensorMap tm, tmkv;
  for (int which = 0; which < 2; ++which) {
    // dims (innermost first): 8 elems | seq rows | 16-byte chunks across [Q|K|V] | batch; box rows: 128 (Q) or kv (K, V)
    uint64_t dims[4] = {8, (uint64_t)seq, (uint64_t)(3 * C / 8), (uint64_t)batch};
    uint64_t strides[3] = {(uint64_t)3 * C, 8, (uint64_t)seq * 3 * C};
    uint32_t box[4] = {8, (uint32_t)(which ? kv : 128), (uint32_t)(head_dim / 8), 1};
    int rc = make_tmap_bf16(which ? &tmkv : &tm, qkv, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
  }
#define B200_ATTN_CASE(D_, KVS_, KVP_, KVP2_)                                                                         \
  case D_:                                                                                                           \
    if (!pipe) return kv == KVS_ ? launch_attention<D_, KVS_, false>(tm, tmkv, qkv, p, stream)                       \
                                 : launch_attention<D_, 128, false>(tm, tmkv, qkv, p, stream);                       \
    if (kv == KVP_) return launch_attention<D_, KVP_, true>(tm, tmkv, qkv, p, stream);                               \
    if (kv == KVP2_) return launch_attention<D_, KVP2_, true>(tm, tmkv, qkv, p, stream);                             \
    return fail(B200_ERR_UNSUPPORTED, "attention: no pipelined kernel for head_dim %d with %d-key blocks", D_, kv);
  if (ts) {
    if (head_dim == 32) return pipe ? launch_attention<32, 32, true, true>(tm, tmkv, qkv, p, stream)
                                    : kv == 64 ? launch_attention<32, 64, false, true>(tm, tmkv, qkv, p, stream)
                                               : launch_attention<32, 128, false, true>(tm, tmkv, qkv, p, stream);
    return pipe ? launch_attention<64, 64, true, true>(tm, tmkv, qkv, p, stream)
                : launch_attention<64, 128, false, true>(tm, tmkv, qkv, p, stream);
  }
  switch (head_dim) {
    B200_ATTN_CASE(32, 64, 32, 64)
    B200_ATTN_CASE(48, 64, 32, 64)
    B200_ATTN_CASE(64, 128, 64, 128)
    B200_ATTN_CASE(80, 128, 64, 128)
    B200_ATTN_CASE(96, 128, 64, 128)
    B200_ATTN_CASE(160, 128, 64, 128)
    default: return fail(B200_ERR_UNSUPPORTED, "attention: head_dim %d unsupported (32/48/64/80/96/160)", head_dim);
  }
#undef B200_ATTN_CASE
}
