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
// * Epilogue (4 warps, one accumulator row per thread): + bias[n] + per-image row vector
//   (timestep/class embedding projection, K6) + residual, optional GEGLU (K3), bf16/fp32 store.
//
// Replaces (reference call sites -> diffusers/torch): F.conv2d / F.linear under
// UNet2DConditionModel.forward, /root/reference/script/train/train_audioldm_lora.py:539-546.
#include "host_util.h"
#include "ptx.cuh"

namespace b200 {

static constexpr int kBlockM = 128;
static constexpr int kBlockK = 64;              // 64 bf16 = one 128B swizzle row
static constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KB
static constexpr int kThreads = 192;            // warp0 TMA, warp1 MMA, warps2-5 epilogue

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
  int stride;                       // 1, or 2: keep even (h, w) only (Downsample2D: k3 s2 p1)
  int Hout, Wout;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB,
                 const ConvGemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024B alignment for the 128B swizzle atoms.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = kABytes + p.block_n * kBlockK * 2;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tfull_bar = empty_bar + p.stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int n_tile = t % p.num_n_tiles;
        const int m_tile = t / p.num_n_tiles;
        const int h0 = (m_tile % p.tiles_h) * p.BH;
        const int n0 = (m_tile / p.tiles_h) * p.BNI;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* a_dst = smem + s * stage_bytes;
          uint8_t* b_dst = a_dst + kABytes;
          mbar_expect_tx(&full_bar[s], stage_bytes);
          if (kb < p.seg_end0) {
            const int tap = kb / p.cb0;
            const int cb = kb - tap * p.cb0;
            int dh = 0, dw = 0;
            if (p.ntaps == 9) {
              dh = tap / 3 - 1;
              dw = tap % 3 - 1;
            }
            tma_load_4d(a_dst, &tmA0, &full_bar[s], cb * kBlockK, dw, h0 + dh, n0);
          } else if (kb < p.seg_end1) {
            tma_load_4d(a_dst, &tmA1, &full_bar[s], (kb - p.seg_end0) * kBlockK, 0, h0, n0);
          } else {
            tma_load_4d(a_dst, &tmA2, &full_bar[s], (kb - p.seg_end1) * kBlockK, 0, h0, n0);
          }
          tma_load_2d(b_dst, &tmB, &full_bar[s], kb * kBlockK, n_tile * p.block_n);
          if (++s == p.stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(kBlockM, p.block_n, 0, 0);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t use = it >> 1;
        mbar_wait(&tempty_bar[buf], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * p.block_n;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
          const uint32_t b_addr = a_addr + kABytes;
          // K-major, 128B swizzle: 8-row atom = 1024 B -> SBO = 1024; LBO unused (1).
          const uint64_t a_desc = make_smem_desc(a_addr, 16, 1024, SWZ_128B);
          const uint64_t b_desc = make_smem_desc(b_addr, 16, 1024, SWZ_128B);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the swizzle atom: +2 in (addr >> 4) units
            umma_bf16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[s]);
          if (++s == p.stages) {
            s = 0;
            ph ^= 1;
          }
        }
        umma_commit(&tfull_bar[buf]);
      }
    }
  } else {
    // ================================================================ epilogue warps
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;          // accumulator row == tile pixel
    const int wl = row % p.W;
    const int hl = (row / p.W) % p.BH;
    const int nl = row / (p.W * p.BH);
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = it >> 1;
      const int n_tile = t % p.num_n_tiles;
      const int m_tile = t / p.num_n_tiles;
      const int h = (m_tile % p.tiles_h) * p.BH + hl;
      const int n = (m_tile / p.tiles_h) * p.BNI + nl;
      bool row_ok = (h < p.H) && (n < p.NB);
      size_t pix = (static_cast<size_t>(n) * p.H + h) * p.W + wl;
      if (p.stride == 2) {
        row_ok = row_ok && ((h & 1) == 0) && ((wl & 1) == 0);
        pix = (static_cast<size_t>(n) * p.Hout + (h >> 1)) * p.Wout + (wl >> 1);
      }
      mbar_wait(&tfull_bar[buf], use & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * p.block_n;

      if (!p.geglu) {
        const int col_base = n_tile * p.block_n;
        for (int c = 0; c < p.block_n / 32; ++c) {
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
                if (p.bias) {
                  const float4 b0 = *reinterpret_cast<const float4*>(p.bias + cg);
                  const float4 b1 = *reinterpret_cast<const float4*>(p.bias + cg + 4);
                  v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                  v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
                }
                if (p.rowvec) {
                  const float* rv = p.rowvec + static_cast<size_t>(n) * p.rowvec_ld + cg;
                  const float4 b0 = *reinterpret_cast<const float4*>(rv);
                  const float4 b1 = *reinterpret_cast<const float4*>(rv + 4);
                  v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                  v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
                }
                if (p.residual) {
                  const uint4 rr = *reinterpret_cast<const uint4*>(p.residual + pix * p.res_ld + cg);
                  v[0] += bf16_lo(rr.x); v[1] += bf16_hi(rr.x); v[2] += bf16_lo(rr.y); v[3] += bf16_hi(rr.y);
                  v[4] += bf16_lo(rr.z); v[5] += bf16_hi(rr.z); v[6] += bf16_lo(rr.w); v[7] += bf16_hi(rr.w);
                }
                if (p.out_fp32) {
                  float* o = reinterpret_cast<float*>(p.out) + pix * p.out_ld + cg;
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
      } else {
        // GEGLU: tile columns [0, bn/2) are values, [bn/2, bn) the matching gates.
        const int half = p.block_n / 2;
        const int out_base = n_tile * half;
        for (int c = 0; c < half / 32; ++c) {
          uint32_t rv[32], rg[32];
          tmem_ld_x32(t_row + c * 32, rv);
          tmem_ld_x32(t_row + half + c * 32, rg);
          tmem_wait_ld();
          if (row_ok) {
            const int bcol = n_tile * p.block_n + c * 32;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int oc = out_base + c * 32 + g * 8;
              if (oc < p.n_valid) {
                float o8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  float val = __uint_as_float(rv[g * 8 + j]);
                  float gate = __uint_as_float(rg[g * 8 + j]);
                  if (p.bias) {
                    val += p.bias[bcol + g * 8 + j];
                    gate += p.bias[bcol + half + g * 8 + j];
                  }
                  o8[j] = val * gelu_erf(gate);
                }
                __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.out_ld + oc;
                uint4 pk;
                pk.x = pack_bf16x2(o8[0], o8[1]); pk.y = pack_bf16x2(o8[2], o8[3]);
                pk.z = pack_bf16x2(o8[4], o8[5]); pk.w = pack_bf16x2(o8[6], o8[7]);
                *reinterpret_cast<uint4*>(o) = pk;
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
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

// C-ABI: see include/b200ldm.h
extern "C" int b200_conv_gemm(const void* a0, int c0, const void* a1, int c1, const void* a2, int c2, int nb, int h,
                              int w, int ntaps, int stride, const void* wpacked, int n_pad, int n_valid,
                              const float* bias, const float* rowvec, int rowvec_ld, const void* residual, int res_ld,
                              void* out, int out_ld, int out_fp32, int geglu, int block_n, int max_ctas,
                              void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(stride == 1 || (stride == 2 && ntaps == 9 && !residual), "conv_gemm: stride %d unsupported", stride);
  B200_CHECK_ARG(a0 && wpacked && out, "conv_gemm: null pointer");
  B200_CHECK_ARG(ntaps == 1 || ntaps == 9, "conv_gemm: ntaps must be 1 or 9 (got %d)", ntaps);
  B200_CHECK_ARG(c0 > 0 && c0 % 64 == 0 && c1 % 64 == 0 && c2 % 64 == 0, "conv_gemm: channels must be multiples of 64 (%d,%d,%d)", c0, c1, c2);
  B200_CHECK_ARG((c1 == 0) == (a1 == nullptr) && (c2 == 0) == (a2 == nullptr), "conv_gemm: segment pointer/channel mismatch");
  B200_CHECK_ARG(block_n >= 32 && block_n <= 256 && block_n % 32 == 0, "conv_gemm: block_n %d unsupported", block_n);
  B200_CHECK_ARG(n_pad % block_n == 0, "conv_gemm: n_pad %d not a multiple of block_n %d", n_pad, block_n);
  B200_CHECK_ARG(n_valid % 8 == 0 && n_valid <= (geglu ? n_pad / 2 : n_pad), "conv_gemm: n_valid %d invalid", n_valid);
  B200_CHECK_ARG(!geglu || (block_n % 64 == 0 && !out_fp32 && !residual && !rowvec), "conv_gemm: geglu constraints");
  B200_CHECK_ARG(nb > 0 && h > 0 && w > 0, "conv_gemm: empty activation");

  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  if (pick_box(h, w, nb, &p.BH, &p.BNI) != 0) return fail(B200_ERR_UNSUPPORTED, "conv_gemm: width %d does not divide 128", w);
  p.cb0 = c0 / 64;
  p.ntaps = ntaps;
  p.seg_end0 = ntaps * p.cb0;
  p.seg_end1 = p.seg_end0 + c1 / 64;
  p.num_kb = p.seg_end1 + c2 / 64;
  p.H = h; p.W = w; p.NB = nb;
  p.stride = stride;
  p.Hout = stride == 2 ? (h - 1) / 2 + 1 : h;
  p.Wout = stride == 2 ? (w - 1) / 2 + 1 : w;
  p.tiles_h = (h + p.BH - 1) / p.BH;
  p.num_m_tiles = p.tiles_h * ((nb + p.BNI - 1) / p.BNI);
  p.num_n_tiles = n_pad / block_n;
  p.block_n = block_n;
  p.n_valid = n_valid;
  p.bias = bias; p.rowvec = rowvec; p.rowvec_ld = rowvec_ld;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual); p.res_ld = res_ld;
  p.out = out; p.out_ld = out_ld; p.out_fp32 = out_fp32; p.geglu = geglu;
  int tc = 32;
  while (tc < 2 * block_n) tc *= 2;
  p.tmem_cols = tc;
  const int stage_bytes = kABytes + block_n * kBlockK * 2;
  const int smem_budget = 200 * 1024;
  p.stages = smem_budget / stage_bytes;
  if (p.stages > 8) p.stages = 8;
  const int smem_bytes = p.stages * stage_bytes + 1024 /*align slack*/ + 256 /*barriers*/;

  CUtensorMap tA[3], tB;
  const void* srcs[3] = {a0, a1 ? a1 : a0, a2 ? a2 : a0};
  const int chans[3] = {c0, c1 ? c1 : c0, c2 ? c2 : c0};
  for (int i = 0; i < 3; ++i) {
    const uint64_t C = chans[i];
    uint64_t dims[4] = {C, (uint64_t)w, (uint64_t)h, (uint64_t)nb};
    uint64_t strides[3] = {C, C * w, C * w * h};
    uint32_t box[4] = {64, (uint32_t)w, (uint32_t)p.BH, (uint32_t)p.BNI};
    int rc = make_tmap_bf16(&tA[i], srcs[i], 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    const uint64_t K = static_cast<uint64_t>(p.num_kb) * 64;
    uint64_t dims[2] = {K, (uint64_t)n_pad};
    uint64_t strides[1] = {K};
    uint32_t box[2] = {64, (uint32_t)block_n};
    int rc = make_tmap_bf16(&tB, wpacked, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }

  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  }
  int grid = p.num_m_tiles * p.num_n_tiles;
  int cap = max_ctas > 0 ? max_ctas : num_sms;
  if (grid > cap) grid = cap;
  conv_gemm_kernel<<<grid, kThreads, smem_bytes, stream>>>(tA[0], tA[1], tA[2], tB, p);
  B200_CHECK_LAUNCH("conv_gemm");
  return B200_OK;
}
