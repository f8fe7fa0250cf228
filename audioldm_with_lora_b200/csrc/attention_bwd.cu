// Flash-attention backward for sm_100a (K14: the SDPA part of the fine-tuning step's backward pass).
//
// Given Q, K, V (the fused [B, S, 3C] projection output), dO, and from the forward pass the per-row log-sum-exp
// (log2 domain) and delta = rowsum(dO * O), it produces d[Q|K|V] in the same fused layout, so the following dgrad GEMM
// (d_ln = d_qkv . W_qkv) consumes it as is.  One kernel, two modes, both with the loop structure of the forward kernel:
//
//   mode KV  (CTA = 128 keys of one (batch, head); streams the query blocks)
//       S^T  = K Q^T            tcgen05 128x128xd   -> TMEM T1           P^T  = exp2(S^T c - lse[query])
//       dP^T = V dO^T           tcgen05 128x128xd   -> TMEM T2           dS^T = P^T * (dP^T - delta[query])
//       dV  += P^T  dO          tcgen05 128xdx128   (dO as MN-major B operand, accumulates in TMEM over the query blocks)
//       dK  += dS^T Q           tcgen05 128xdx128   (Q  as MN-major B operand)           epilogue: dK *= scale
//   mode Q   (CTA = 128 queries; streams the key blocks)
//       S  = Q K^T,  dP = dO V^T,  P = exp2(S c - lse[row]),  dS = P * (dP - delta[row])
//       dQ += dS K              tcgen05 128xdx128   (K as MN-major B operand)             epilogue: dQ *= scale
//
// i.e. T1 = X1 Y1^T, T2 = X2 Y2^T with (X1, X2) resident and (Y1, Y2) streamed through a TMA ring; the statistics are
// indexed by column (mode KV) or by row (mode Q), and mode Q skips the P.Y2 product.  Recomputing S in both modes costs
// 7 instead of 5 matrix products but needs no atomics: every output element has exactly one writer (deterministic).
// P^T / dS^T are written by 128 threads (one row each) straight into the no-swizzle K-major UMMA layout, like P in
// the forward kernel.  head_dim <= 128 (TMEM: 2 x 128 score columns + 2 x head_dim accumulator columns).
//
// Replaces the autograd backward of F.scaled_dot_product_attention reached from `accelerator.backward(loss)`,
// /root/reference/script/train/train_audioldm_lora.py:557.
#include "host_util.h"
#include "ptx.cuh"

namespace b200 {

static constexpr int kBT = 128;                // rows per CTA
// Two elementwise warps per TMEM lane quarter, each taking half of the block's columns (a warp reaches the 32 TMEM lanes
// warp % 4 owns): 8 instead of 4 elementwise warps per CTA.  With two (or one) CTAs per SM the elementwise pass was bound by
// the latency of each thread's tcgen05.ld -> exp -> pack -> store chain, not by a pipe.
static constexpr int kBwdEw = 8;               // elementwise warps
static constexpr int kBwdThreads = (kBwdEw + 2) * 32;   // warps 0-7 elementwise, warp 8 TMA, warp 9 MMA
// Columns per streamed block: template parameter BC (128, or 64 for head_dim <= 48: two 64-column score tiles + two
// accumulators then fit 256 TMEM columns and ~64 KB of shared memory, so two CTAs share an SM and overlap each other's
// MMA -> elementwise -> MMA hand-offs).

struct AttnBwdParams {
  int seq, heads, batch;
  int nblk;                 // ceil(seq / BC)
  int stages;
  int tmem_cols;
  float scale_log2;         // scale * log2(e)
  float scale;
  const float* lse;         // [batch, heads, seq] (log2 domain)
  const float* delta;       // [batch, heads, seq]
  __nv_bfloat16* dqkv;      // [batch, seq, 3C]
  int ld;                   // 3C
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// TS: P^T / dS^T go back to tensor memory (bf16 pairs, BC / 2 columns each, behind T1 / T2) and are the TMEM A operands of
// the accumulating products, instead of 2 x 16 KB of shared-memory stores and tensor-core reads per block (the forward
// kernel's TS form, attention.cu).  Needs 3 BC + 2 D <= 512 TMEM columns without giving up a co-resident CTA: head_dim 32
// (BC 64: 256 columns, two CTAs per SM as before) and 64 (BC 128: 512 columns, one CTA either way).
template <int D, int BC, bool kModeQ, bool TS = false>
__global__ void __launch_bounds__(kBwdThreads, ((TS ? 3 : 2) * BC + 2 * D <= 256) ? 2 : 1)      // two CTAs per SM where TMEM allows it
attention_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                     const __grid_constant__ CUtensorMap tmQKVc, const __grid_constant__ CUtensorMap tmDOc,
                     const AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  constexpr int kTileBytes = 128 * D * 2;       // resident X tiles (box of 128 rows: maps tmQKV / tmDO)
  constexpr int kYTile = BC * D * 2;            // streamed Y tiles (box of BC rows: maps tmQKVc / tmDOc)
  constexpr int kChunkY = BC * 16;              // bytes between 8-element chunks of a Y tile
  constexpr int kPBytes = kBT * BC * 2;
  uint8_t* sX1 = smem;
  uint8_t* sX2 = sX1 + kTileBytes;
  uint8_t* sP = sX2 + kTileBytes;               // P^T (mode KV only)
  uint8_t* sDS = sP + kPBytes;                  // dS^T / dS
  uint8_t* sY = TS ? sP : sDS + kPBytes;        // stages x {Y1, Y2} (TS: no P^T / dS^T tiles in shared memory)
  float* s_stat = reinterpret_cast<float*>(sY + p.stages * 2 * kYTile);         // [2][128]: lse, delta of the column block
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stat + 256);
  uint64_t* x_full = bars;
  uint64_t* y_full = bars + 1;                  // [2]
  uint64_t* y_empty = bars + 3;                 // [2]
  uint64_t* t_full = bars + 5;
  uint64_t* ps_full = bars + 6;
  uint64_t* acc_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * kBT;              // first row (key in mode KV, query in mode Q) of this CTA
  const int bh = blockIdx.y;
  const int b = bh / p.heads;
  const int h = bh % p.heads;
  const int C8 = p.heads * D / 8;
  const int chunk_q = h * (D / 8);
  const int chunk_k = C8 + chunk_q;
  const int chunk_v = 2 * C8 + chunk_q;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmQKVc);
    tma_prefetch_desc(&tmDOc);
    mbar_init(x_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&y_full[i], 1);
      mbar_init(&y_empty[i], 1);
    }
    mbar_init(t_full, 1);
    mbar_init(ps_full, kBwdEw * 32);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == kBwdEw) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t t_1 = tmem_base;                   // T1: columns [0, BC)
  const uint32_t t_2 = tmem_base + BC;              // T2: columns [BC, 2 BC)
  const uint32_t t_p = tmem_base + 2 * BC;          // TS: P^T as bf16 pairs, BC / 2 columns
  const uint32_t t_ds = t_p + BC / 2;               // TS: dS^T / dS
  const uint32_t t_a1 = tmem_base + (TS ? 3 : 2) * BC;      // Acc1 (dV): D columns behind the score (and P / dS) columns
  const uint32_t t_a2 = t_a1 + D;                   // Acc2 (dK or dQ): the next D columns

  if (warp == kBwdEw) {
    // ============================================================ TMA producer
    // (whole warp, uniform control flow, one lane elected at each use issues -- see conv_gemm.cu)
    {
      if (elect_one()) {
        mbar_expect_tx(x_full, 2 * kTileBytes);
        if (kModeQ) {
          tma_load_4d(sX1, &tmQKV, x_full, 0, r0, chunk_q, b);
          tma_load_4d(sX2, &tmDO, x_full, 0, r0, chunk_q, b);
        } else {
          tma_load_4d(sX1, &tmQKV, x_full, 0, r0, chunk_k, b);
          tma_load_4d(sX2, &tmQKV, x_full, 0, r0, chunk_v, b);
        }
      }
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < p.nblk; ++j) {
        mbar_wait(&y_empty[s], ph ^ 1);
        uint8_t* dst = sY + s * 2 * kYTile;
        if (elect_one()) {
          mbar_expect_tx(&y_full[s], 2 * kYTile);
          if (kModeQ) {
            tma_load_4d(dst, &tmQKVc, &y_full[s], 0, j * BC, chunk_k, b);
            tma_load_4d(dst + kYTile, &tmQKVc, &y_full[s], 0, j * BC, chunk_v, b);
          } else {
            tma_load_4d(dst, &tmQKVc, &y_full[s], 0, j * BC, chunk_q, b);
            tma_load_4d(dst + kYTile, &tmDOc, &y_full[s], 0, j * BC, chunk_q, b);
          }
        }
        if (++s == p.stages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == kBwdEw + 1) {
    // ============================================================ MMA issuer (whole warp, elected lane issues)
    {
      const uint32_t idesc_t = make_idesc_bf16(128, BC, 0, 0);        // T = X Y^T, both K-major
      const uint32_t idesc_a = make_idesc_bf16(128, D, 0, 1);         // Acc += P Y, Y MN-major
      const uint32_t x1_addr = smem_u32(sX1), x2_addr = smem_u32(sX2);
      const uint32_t p_addr = smem_u32(sP), ds_addr = smem_u32(sDS);
      mbar_wait(x_full, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < p.nblk; ++j) {
        const uint32_t y1_addr = smem_u32(sY + s * 2 * kYTile);
        const uint32_t y2_addr = y1_addr + kYTile;
        mbar_wait(&y_full[s], ph);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < D / 16; ++k) {
            const uint64_t a_desc = make_smem_desc(x1_addr + k * 4096, 2048, 128, SWZ_NONE);
            const uint64_t b_desc = make_smem_desc(y1_addr + k * 2 * kChunkY, kChunkY, 128, SWZ_NONE);
            umma_bf16_ss(t_1, a_desc, b_desc, idesc_t, k != 0);
          }
#pragma unroll
          for (int k = 0; k < D / 16; ++k) {
            const uint64_t a_desc = make_smem_desc(x2_addr + k * 4096, 2048, 128, SWZ_NONE);
            const uint64_t b_desc = make_smem_desc(y2_addr + k * 2 * kChunkY, kChunkY, 128, SWZ_NONE);
            umma_bf16_ss(t_2, a_desc, b_desc, idesc_t, k != 0);
          }
          umma_commit(t_full);
        }
        // the elementwise threads have turned T1 / T2 into P^T / dS^T in shared memory
        mbar_wait(ps_full, j & 1);
        tc_fence_after();
        if (elect_one()) {
          if (!kModeQ) {
#pragma unroll
            for (int k = 0; k < BC / 16; ++k) {
              const uint64_t a_desc = make_smem_desc(p_addr + k * 4096, 2048, 128, SWZ_NONE);
              const uint64_t b_desc = make_smem_desc(y2_addr + k * 256, 128, kChunkY, SWZ_NONE);
              if (TS) umma_bf16_ts(t_a1, t_p + k * 8, b_desc, idesc_a, (j | k) != 0);
              else umma_bf16_ss(t_a1, a_desc, b_desc, idesc_a, (j | k) != 0);
            }
          }
#pragma unroll
          for (int k = 0; k < BC / 16; ++k) {
            const uint64_t a_desc = make_smem_desc(ds_addr + k * 4096, 2048, 128, SWZ_NONE);
            const uint64_t b_desc = make_smem_desc(y1_addr + k * 256, 128, kChunkY, SWZ_NONE);
            if (TS) umma_bf16_ts(t_a2, t_ds + k * 8, b_desc, idesc_a, (j | k) != 0);
            else umma_bf16_ss(t_a2, a_desc, b_desc, idesc_a, (j | k) != 0);
          }
          umma_commit(&y_empty[s]);
          if (j == p.nblk - 1) umma_commit(acc_full);
        }
        if (++s == p.stages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else {
    // ============================================================ elementwise warps (row = thread)
    const int quarter = warp & 3, half = warp >> 2;      // TMEM lane quarter; which half of the block's columns
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    uint8_t* sp_row = sP + row * 16;
    uint8_t* sds_row = sDS + row * 16;
    const size_t stat_base = (static_cast<size_t>(b) * p.heads + h) * p.seq;
    float row_lse = 0.f, row_delta = 0.f;
    if (kModeQ && r0 + row < p.seq) {
      row_lse = p.lse[stat_base + r0 + row];
      row_delta = p.delta[stat_base + r0 + row];
    }
    for (int j = 0; j < p.nblk; ++j) {
      const int cvalid = min(BC, p.seq - j * BC);
      if (!kModeQ) {
        // statistics of this block's BC columns (= queries); the previous block's readers are past ps_full
        named_bar_sync(1, kBwdEw * 32);
        const int col = j * BC + row;
        if (half == 0 && row < BC) {
          s_stat[row] = col < p.seq ? p.lse[stat_base + col] : 0.f;
          s_stat[128 + row] = col < p.seq ? p.delta[stat_base + col] : 0.f;
        }
        named_bar_sync(1, kBwdEw * 32);
      }
      // T1 / T2 of this block are complete; every earlier MMA (incl. the previous block's reads of sP / sDS) too
      mbar_wait(t_full, j & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = half * (BC / 64); c < (half + 1) * (BC / 64); ++c) {
        uint32_t r1[32], r2[32];
        tmem_ld_x32(t_1 + lane_off + c * 32, r1);
        tmem_ld_x32(t_2 + lane_off + c * 32, r2);
        tmem_wait_ld();
        uint32_t pt[16], dt[16];                  // TS: the chunk's 32 P^T / dS^T values as bf16 pairs
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t pk[4], dk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float pv[2], dv[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int cc = c * 32 + g * 8 + 2 * i + e;
              const float l = kModeQ ? row_lse : s_stat[cc];
              const float dl = kModeQ ? row_delta : s_stat[128 + cc];
              float pe = ex2f(fmaf(__uint_as_float(r1[g * 8 + 2 * i + e]), p.scale_log2, -l));
              if (cc >= cvalid) pe = 0.f;
              pv[e] = pe;
              dv[e] = pe * (__uint_as_float(r2[g * 8 + 2 * i + e]) - dl);
            }
            pk[i] = pack_bf16x2(pv[0], pv[1]);
            dk[i] = pack_bf16x2(dv[0], dv[1]);
          }
          if (TS) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              pt[g * 4 + i] = pk[i];
              dt[g * 4 + i] = dk[i];
            }
          } else {
            if (!kModeQ) *reinterpret_cast<uint4*>(sp_row + (c * 4 + g) * 2048) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(sds_row + (c * 4 + g) * 2048) = make_uint4(dk[0], dk[1], dk[2], dk[3]);
          }
        }
        if (TS) {
          if (!kModeQ) tmem_st_x16(t_p + lane_off + c * 16, pt);
          tmem_st_x16(t_ds + lane_off + c * 16, dt);
        }
      }
      if (TS) tmem_wait_st();
      else fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(ps_full);
    }
    // ---- epilogue
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int grow = r0 + row;
    __nv_bfloat16* orow = p.dqkv + (static_cast<size_t>(b) * p.seq + grow) * p.ld + h * D;
    const int C = p.heads * D;
#pragma unroll 1
    for (int which = kModeQ ? 1 : 0; which < 2; ++which) {
      const uint32_t t_acc = (which == 0 ? t_a1 : t_a2) + lane_off;
      const float f = which == 0 ? 1.0f : p.scale;
      // mode KV: Acc1 -> dV (section 2), Acc2 -> dK (section 1); mode Q: Acc2 -> dQ (section 0)
      __nv_bfloat16* o = orow + (kModeQ ? 0 : (which == 0 ? 2 * C : C));
#pragma unroll
      for (int c = 0; c < D / 16; ++c) {
        if ((c & 1) != half) continue;              // the two warps of a lane quarter split the accumulator's column chunks
        uint32_t r[16];
        tmem_ld_x16(t_acc + c * 16, r);
        tmem_wait_ld();
        if (grow < p.seq) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint4 pk;
            pk.x = pack_bf16x2(__uint_as_float(r[g * 8 + 0]) * f, __uint_as_float(r[g * 8 + 1]) * f);
            pk.y = pack_bf16x2(__uint_as_float(r[g * 8 + 2]) * f, __uint_as_float(r[g * 8 + 3]) * f);
            pk.z = pack_bf16x2(__uint_as_float(r[g * 8 + 4]) * f, __uint_as_float(r[g * 8 + 5]) * f);
            pk.w = pack_bf16x2(__uint_as_float(r[g * 8 + 6]) * f, __uint_as_float(r[g * 8 + 7]) * f);
            *reinterpret_cast<uint4*>(o + c * 16 + g * 8) = pk;
          }
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == kBwdEw) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// delta[b, h, s] = sum_d dO[b, s, h*D + d] * O[b, s, h*D + d]
__global__ void __launch_bounds__(256)
attention_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout, int batch, int seq,
                       int heads, int d, float* __restrict__ delta) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t total = static_cast<size_t>(batch) * seq * heads;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int h = static_cast<int>(i % heads);
    const size_t bs = i / heads;                  // b * seq + s
    const size_t off = bs * heads * d + static_cast<size_t>(h) * d;
    float acc = 0.f;
    for (int c = 0; c < d; c += 8) {
      const uint4 a = *reinterpret_cast<const uint4*>(o + off + c);
      const uint4 g = *reinterpret_cast<const uint4*>(dout + off + c);
      acc += bf16_lo(a.x) * bf16_lo(g.x) + bf16_hi(a.x) * bf16_hi(g.x) + bf16_lo(a.y) * bf16_lo(g.y) + bf16_hi(a.y) * bf16_hi(g.y) +
             bf16_lo(a.z) * bf16_lo(g.z) + bf16_hi(a.z) * bf16_hi(g.z) + bf16_lo(a.w) * bf16_lo(g.w) + bf16_hi(a.w) * bf16_hi(g.w);
    }
    const size_t b = bs / seq, s = bs % seq;
    delta[(b * heads + h) * seq + s] = acc;
  }
}

template <int D, int BC, bool TS = false>
static int launch_attention_bwd(const CUtensorMap& tq, const CUtensorMap& td, const CUtensorMap& tqc, const CUtensorMap& tdc,
                                AttnBwdParams& p, cudaStream_t stream) {
  constexpr int kTileBytes = 128 * D * 2;
  constexpr int kYTile = BC * D * 2;
  const int fixed = 2 * kTileBytes + (TS ? 0 : 2 * kBT * BC * 2) + 1024 /*stats*/ + 128 /*barriers*/ + 128 /*align*/;
  int cols = 32;
  while (cols < (TS ? 3 : 2) * BC + 2 * D) cols *= 2;
  p.tmem_cols = cols;
  p.nblk = (p.seq + BC - 1) / BC;
  const int budget = (220 * 1024) / (512 / cols);
  p.stages = (fixed + 2 * 2 * kYTile <= budget) ? 2 : 1;
  const int smem_bytes = fixed + p.stages * 2 * kYTile;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_bwd_kernel<D, BC, false, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_bwd_kernel<D, BC, true, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail(B200_ERR_CUDA, "attention_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  dim3 grid((p.seq + kBT - 1) / kBT, p.batch * p.heads);
  B200_CHECK_PDL("attention_bwd(kv)", launch_pdl(attention_bwd_kernel<D, BC, false, TS>, grid, dim3(kBwdThreads), (size_t)smem_bytes,
                                                 stream, 0, tq, td, tqc, tdc, p));
  B200_CHECK_PDL("attention_bwd(q)", launch_pdl(attention_bwd_kernel<D, BC, true, TS>, grid, dim3(kBwdThreads), (size_t)smem_bytes,
                                                stream, 0, tq, td, tqc, tdc, p));
  return B200_OK;
}

}  // namespace b200

using namespace b200;

// qkv [B, S, 3C] bf16, o / dout [B, S, C] bf16, lse [B, heads, S] fp32 (from b200_attention_lse), delta [B, heads, S]
// fp32 scratch (written here), dqkv [B, S, 3C] bf16 out.
extern "C" int b200_attention_bwd(const void* qkv, const void* o, const void* dout, const float* lse, float* delta,
                                  void* dqkv, int batch, int seq, int heads, int head_dim, float scale, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(qkv && o && dout && lse && delta && dqkv, "attention_bwd: null pointer");
  B200_CHECK_ARG(batch > 0 && seq > 0 && heads > 0, "attention_bwd: empty problem");
  const int C = heads * head_dim;
  {
    const size_t total = static_cast<size_t>(batch) * seq * heads;
    size_t g = (total + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    B200_CHECK_PDL("attention_delta", launch_pdl(attention_delta_kernel, dim3(static_cast<unsigned>(g)), dim3(256), 0, stream, 0,
                                                 reinterpret_cast<const __nv_bfloat16*>(o), reinterpret_cast<const __nv_bfloat16*>(dout),
                                                 batch, seq, heads, head_dim, delta));
  }
  AttnBwdParams p;
  memset(&p, 0, sizeof(p));
  p.seq = seq; p.heads = heads; p.batch = batch;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = lse; p.delta = delta;
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  p.ld = 3 * C;
  static const int bc_env = getenv("B200_ATTN_BWD_BC") ? atoi(getenv("B200_ATTN_BWD_BC")) : 0;     // A/B knob: 64 or 128
  const int bc = (head_dim <= 48 && bc_env != 128) ? 64 : 128;
  CUtensorMap tq, td, tqc, tdc;
  for (int which = 0; which < 2; ++which) {
    const uint32_t rows = which ? (uint32_t)bc : 128u;
    {
      uint64_t dims[4] = {8, (uint64_t)seq, (uint64_t)(3 * C / 8), (uint64_t)batch};
      uint64_t strides[3] = {(uint64_t)3 * C, 8, (uint64_t)seq * 3 * C};
      uint32_t box[4] = {8, rows, (uint32_t)(head_dim / 8), 1};
      int rc = make_tmap_bf16(which ? &tqc : &tq, qkv, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
      if (rc) return rc;
    }
    {
      uint64_t dims[4] = {8, (uint64_t)seq, (uint64_t)(C / 8), (uint64_t)batch};
      uint64_t strides[3] = {(uint64_t)C, 8, (uint64_t)seq * C};
      uint32_t box[4] = {8, rows, (uint32_t)(head_dim / 8), 1};
      int rc = make_tmap_bf16(which ? &tdc : &td, dout, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
      if (rc) return rc;
    }
  }
  static const int ts_env = getenv("B200_ATTN_BWD_TS") ? atoi(getenv("B200_ATTN_BWD_TS")) : 1;     // A/B knob: 0 = P^T / dS^T in smem
  if (ts_env && head_dim == 32 && bc == 64) return launch_attention_bwd<32, 64, true>(tq, td, tqc, tdc, p, stream);
  if (ts_env && head_dim == 64) return launch_attention_bwd<64, 128, true>(tq, td, tqc, tdc, p, stream);
  switch (head_dim) {
    case 32: return bc == 64 ? launch_attention_bwd<32, 64>(tq, td, tqc, tdc, p, stream) : launch_attention_bwd<32, 128>(tq, td, tqc, tdc, p, stream);
    case 48: return bc == 64 ? launch_attention_bwd<48, 64>(tq, td, tqc, tdc, p, stream) : launch_attention_bwd<48, 128>(tq, td, tqc, tdc, p, stream);
    case 64: return launch_attention_bwd<64, 128>(tq, td, tqc, tdc, p, stream);
    case 80: return launch_attention_bwd<80, 128>(tq, td, tqc, tdc, p, stream);
    case 96: return launch_attention_bwd<96, 128>(tq, td, tqc, tdc, p, stream);
    default: return fail(B200_ERR_UNSUPPORTED, "attention_bwd: head_dim %d unsupported (32/48/64/80/96)", head_dim);
  }
}
