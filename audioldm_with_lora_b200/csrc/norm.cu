// GroupNorm (+SiLU) and LayerNorm for NHWC bf16 activations -- HBM-bound, vectorised 16 B accesses.
//
// GroupNorm is two launches: gn_stats writes per-(image, slab, group) partial sum / sum-of-squares
// (deterministic: no atomics, no memset), gn_apply reduces the partials for its image in shared
// memory, then normalises, applies the affine and optional SiLU, and writes bf16.  Both read up to
// two source tensors so the up-block `torch.cat([h, skip], 1)` is consumed in place (K11).
// Statistics are fp32 whatever the storage type.
//
// Replaces F.group_norm + F.silu in ResnetBlock2D / Transformer2DModel / conv_norm_out and
// F.layer_norm in BasicTransformerBlock (diffusers, driven from
// /root/reference/script/train/train_audioldm_lora.py:539-546).
#include "host_util.h"
#include "ptx.cuh"

namespace b200 {

static constexpr int kGnThreads = 256;
static constexpr int kMaxC = 2560;          // cat(1280, 1280) at AudioLDM-L

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

// x0 [NB, HW, C0], x1 [NB, HW, C1] (nullable) ; partial [NB, nslab, groups, 2]
__global__ void __launch_bounds__(kGnThreads)
gn_stats_kernel(const __nv_bfloat16* __restrict__ x0, int C0, const __nv_bfloat16* __restrict__ x1, int C1, int HW,
                int groups, int nslab, float* __restrict__ partial) {
  const int C = C0 + C1;
  const int vec_per_pix = C / 8;
  const int n = blockIdx.y;
  const int slab = blockIdx.x;
  const int pix_per_slab = (HW + nslab - 1) / nslab;
  const int p_begin = slab * pix_per_slab;
  const int p_end = min(HW, p_begin + pix_per_slab);
  __shared__ float s_sum[kMaxC];
  __shared__ float s_sq[kMaxC];
  for (int c = threadIdx.x; c < C; c += kGnThreads) {
    s_sum[c] = 0.f;
    s_sq[c] = 0.f;
  }
  __syncthreads();
  // thread -> fixed channel vector, strided over pixels
  const int total_vec = (p_end - p_begin) * vec_per_pix;
  if (vec_per_pix <= kGnThreads) {
    const int v = threadIdx.x % vec_per_pix;
    const int pl = threadIdx.x / vec_per_pix;
    const int pstride = kGnThreads / vec_per_pix;       // whole pixel lanes; leftover threads idle
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int c = v * 8;
    const __nv_bfloat16* src = (c < C0) ? x0 + (static_cast<size_t>(n) * HW) * C0 + c
                                        : x1 + (static_cast<size_t>(n) * HW) * C1 + (c - C0);
    const int ld = (c < C0) ? C0 : C1;
    for (int p = p_begin + pl; pl < pstride && p < p_end; p += pstride) {
      const uint4 u = *reinterpret_cast<const uint4*>(src + static_cast<size_t>(p) * ld);
      float f[8];
      unpack8(u, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a[j] += f[j];
        b[j] += f[j] * f[j];
      }
    }
    if (pl < pstride) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&s_sum[c + j], a[j]);
        atomicAdd(&s_sq[c + j], b[j]);
      }
    }
  } else {
    for (int i = threadIdx.x; i < total_vec; i += kGnThreads) {
      const int p = p_begin + i / vec_per_pix;
      const int c = (i % vec_per_pix) * 8;
      const __nv_bfloat16* src = (c < C0) ? x0 + (static_cast<size_t>(n) * HW + p) * C0 + c
                                          : x1 + (static_cast<size_t>(n) * HW + p) * C1 + (c - C0);
      const uint4 u = *reinterpret_cast<const uint4*>(src);
      float f[8];
      unpack8(u, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&s_sum[c + j], f[j]);
        atomicAdd(&s_sq[c + j], f[j] * f[j]);
      }
    }
  }
  __syncthreads();
  const int cpg = C / groups;
  for (int g = threadIdx.x; g < groups; g += kGnThreads) {
    float s = 0.f, q = 0.f;
    for (int j = 0; j < cpg; ++j) {
      s += s_sum[g * cpg + j];
      q += s_sq[g * cpg + j];
    }
    float* dst = partial + ((static_cast<size_t>(n) * nslab + slab) * groups + g) * 2;
    dst[0] = s;
    dst[1] = q;
  }
}

// y [NB, HW, C] = act( (x - mean) * rstd * gamma + beta )
__global__ void __launch_bounds__(kGnThreads)
gn_apply_kernel(const __nv_bfloat16* __restrict__ x0, int C0, const __nv_bfloat16* __restrict__ x1, int C1, int HW,
                int groups, int nslab, const float* __restrict__ partial, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, int silu, __nv_bfloat16* __restrict__ y, int pix_per_block) {
  const int C = C0 + C1;
  const int n = blockIdx.y;
  __shared__ float s_scale[kMaxC];
  __shared__ float s_shift[kMaxC];
  __shared__ float s_mean[64], s_rstd[64];
  const int cpg = C / groups;
  if (threadIdx.x < groups) {
    float s = 0.f, q = 0.f;
    for (int k = 0; k < nslab; ++k) {
      const float* src = partial + ((static_cast<size_t>(n) * nslab + k) * groups + threadIdx.x) * 2;
      s += src[0];
      q += src[1];
    }
    const float cnt = static_cast<float>(HW) * cpg;
    const float mean = s / cnt;
    const float var = fmaxf(q / cnt - mean * mean, 0.f);
    s_mean[threadIdx.x] = mean;
    s_rstd[threadIdx.x] = rsqrtf(var + eps);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kGnThreads) {
    const int g = c / cpg;
    const float sc = gamma[c] * s_rstd[g];
    s_scale[c] = sc;
    s_shift[c] = beta[c] - s_mean[g] * sc;
  }
  __syncthreads();
  const int vec_per_pix = C / 8;
  const int p_begin = blockIdx.x * pix_per_block;
  const int p_end = min(HW, p_begin + pix_per_block);
  const int total_vec = (p_end - p_begin) * vec_per_pix;
  for (int i = threadIdx.x; i < total_vec; i += kGnThreads) {
    const int p = p_begin + i / vec_per_pix;
    const int c = (i % vec_per_pix) * 8;
    const __nv_bfloat16* src = (c < C0) ? x0 + (static_cast<size_t>(n) * HW + p) * C0 + c
                                        : x1 + (static_cast<size_t>(n) * HW + p) * C1 + (c - C0);
    const uint4 u = *reinterpret_cast<const uint4*>(src);
    float f[8];
    unpack8(u, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = f[j] * s_scale[c + j] + s_shift[c + j];
      if (silu) v = v / (1.0f + __expf(-v));
      f[j] = v;
    }
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(y + (static_cast<size_t>(n) * HW + p) * C + c) = o;
  }
}

// One warp per row; C % 8 == 0, C <= 1280.  y = (x - mean) * rstd * gamma + beta  (eps inside sqrt)
static constexpr int kLnMaxVec = 5;     // 5 * 32 lanes * 8 = 1280 channels
__global__ void __launch_bounds__(256)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, int M, int C, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  const int nvec = C / 8;
  const __nv_bfloat16* row = x + static_cast<size_t>(warp) * C;
  float f[kLnMaxVec][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      const uint4 u = *reinterpret_cast<const uint4*>(row + v * 8);
      unpack8(u, f[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[i][j];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = f[i][j] - mean;
        q += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / C + eps);
  __nv_bfloat16* out = y + static_cast<size_t>(warp) * C;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + v * 8);
      const float4 g1 = *reinterpret_cast<const float4*>(gamma + v * 8 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(beta + v * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(beta + v * 8 + 4);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = (f[i][j] - mean) * rstd * gg[j] + bb[j];
      uint4 o;
      o.x = pack_bf16x2(r[0], r[1]); o.y = pack_bf16x2(r[2], r[3]);
      o.z = pack_bf16x2(r[4], r[5]); o.w = pack_bf16x2(r[6], r[7]);
      *reinterpret_cast<uint4*>(out + v * 8) = o;
    }
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_gn_nslab(int hw) {
  int nslab = hw / 128;
  if (nslab < 1) nslab = 1;
  if (nslab > 32) nslab = 32;
  return nslab;
}

// partial must hold nb * b200_gn_nslab(hw) * groups * 2 floats.
extern "C" int b200_groupnorm_silu(const void* x0, int c0, const void* x1, int c1, int nb, int hw, int groups,
                                   const float* gamma, const float* beta, float eps, int silu, float* partial,
                                   void* y, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const int C = c0 + c1;
  B200_CHECK_ARG(x0 && y && partial && gamma && beta, "groupnorm: null pointer");
  B200_CHECK_ARG(c0 % 8 == 0 && c1 % 8 == 0 && C <= kMaxC, "groupnorm: channels (%d,%d) unsupported", c0, c1);
  B200_CHECK_ARG((c1 == 0) == (x1 == nullptr), "groupnorm: second source mismatch");
  B200_CHECK_ARG(groups > 0 && groups <= 64 && C % groups == 0, "groupnorm: %d channels not divisible by %d groups", C, groups);
  B200_CHECK_ARG(nb > 0 && hw > 0, "groupnorm: empty input");
  const int nslab = b200_gn_nslab(hw);
  gn_stats_kernel<<<dim3(nslab, nb), kGnThreads, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x0), c0,
                                                             reinterpret_cast<const __nv_bfloat16*>(x1), c1, hw, groups,
                                                             nslab, partial);
  B200_CHECK_LAUNCH("gn_stats");
  // ~16 KB of bf16 per block
  int pix_per_block = (8192 + C - 1) / C;
  if (pix_per_block < 1) pix_per_block = 1;
  const int nblk = (hw + pix_per_block - 1) / pix_per_block;
  gn_apply_kernel<<<dim3(nblk, nb), kGnThreads, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x0), c0,
                                                            reinterpret_cast<const __nv_bfloat16*>(x1), c1, hw, groups,
                                                            nslab, partial, gamma, beta, eps, silu,
                                                            reinterpret_cast<__nv_bfloat16*>(y), pix_per_block);
  B200_CHECK_LAUNCH("gn_apply");
  return B200_OK;
}

extern "C" int b200_layernorm(const void* x, int m, int c, const float* gamma, const float* beta, float eps, void* y,
                              void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(x && y && gamma && beta, "layernorm: null pointer");
  B200_CHECK_ARG(c % 8 == 0 && c <= kLnMaxVec * 256, "layernorm: C=%d unsupported", c);
  if (m == 0) return B200_OK;
  const int warps_per_block = 8;
  const int nblk = (m + warps_per_block - 1) / warps_per_block;
  layernorm_kernel<<<nblk, warps_per_block * 32, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), m, c, gamma,
                                                             beta, eps, reinterpret_cast<__nv_bfloat16*>(y));
  B200_CHECK_LAUNCH("layernorm");
  return B200_OK;
}
