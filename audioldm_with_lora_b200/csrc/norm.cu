// GroupNorm (+SiLU) and LayerNorm for NHWC bf16 activations -- HBM/L2-bound, vectorised 16 B accesses.
//
// GroupNorm is ONE launch: a thread-block cluster (up to 8 CTAs, distributed shared memory) owns one
// image.  Pass 1: every CTA streams its slab of pixels, each thread keeping fp32 sum / sum-of-squares
// of a fixed 8-channel vector in registers; fixed-order reductions (no atomics => bit-deterministic)
// give per-group partials, which the CTAs of the cluster exchange through DSMEM.  Pass 2 re-reads the
// slab (an L2 hit: the producing conv just wrote it), normalises, applies the affine and optional
// SiLU, and writes bf16.  Both passes read up to two source tensors so the up-block
// `torch.cat([h, skip], 1)` is consumed in place (K11).  Statistics are fp32 whatever the storage type.
//
// Replaces F.group_norm + F.silu in ResnetBlock2D / Transformer2DModel / conv_norm_out and
// F.layer_norm in BasicTransformerBlock (diffusers, driven from
// /root/reference/script/train/train_audioldm_lora.py:539-546).
#include <cooperative_groups.h>

#include "host_util.h"
#include "ptx.cuh"

namespace cg = cooperative_groups;

namespace b200 {

static constexpr int kGnThreads = 512;
static constexpr int kMaxC = 2560;          // cat(1280, 1280) at AudioLDM-L
static constexpr int kGnUnroll = 8;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

// SiLU with ONE MUFU op: x sigmoid(x) = h + h tanh(h), h = x / 2 (tanh.approx.f32: 2^-11 relative, i.e. at the level of the
// bf16 rounding of the result).  The x / (1 + exp(-x)) form costs an exponential AND a reciprocal plus the non-ftz range
// fix-ups of __expf / __fdividef -- measured XU-pipe-bound in the epilogues that evaluate it per element.
__device__ __forceinline__ float silu_fast(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

struct GnParams {
  const __nv_bfloat16* x0;
  const __nv_bfloat16* x1;
  int C0, C1, HW, groups;
  const float* gamma;
  const float* beta;
  float eps;
  int silu;
  __nv_bfloat16* y;
  float* stats;             // optional [nb, groups, 2] = (mean, rstd): kept for the backward pass (training)
};

// grid (cluster_size, NB, gsplit), cluster (cluster_size, 1, 1): blockIdx.x = rank of the CTA in its cluster
// (pixel slab), blockIdx.z = which 1/gsplit of the groups (= a contiguous channel range) the cluster owns.
__global__ void __launch_bounds__(kGnThreads)
groupnorm_silu_kernel(const GnParams p) {
  cg::cluster_group cluster = cg::this_cluster();
  pdl_launch_dependents();
  pdl_wait();
  const int nrank = gridDim.x;
  const int rank = blockIdx.x;
  const int n = blockIdx.y;
  const int C = p.C0 + p.C1;
  const int Cs = C / gridDim.z;                     // channels owned by this cluster
  const int c_base = blockIdx.z * Cs;
  const int gl = p.groups / gridDim.z;              // groups owned by this cluster
  const int vpp = Cs >> 3;                          // 16-byte vectors per pixel (this cluster's share)
  const int nlanes = kGnThreads / vpp;              // pixel lanes per CTA (>= 1: C <= 2560)
  const int tid = threadIdx.x;
  const bool active = tid < nlanes * vpp;
  const int v = tid % vpp;
  const int lane = tid / vpp;
  const int lc = v * 8;                             // channel inside the cluster's range
  const int c = c_base + lc;                        // channel of the (concatenated) tensor
  const int cpg = C / p.groups;

  __shared__ float s_sum[kGnThreads * 8];           // [lane][Cs]
  __shared__ float s_sq[kGnThreads * 8];
  __shared__ float s_piv[kMaxC];                    // per-channel pivot K_c (this cluster's channel range)
  __shared__ float s_gpart[128];                    // this CTA's per-group {sum, sumsq}; read by cluster peers
  __shared__ float s_mean[32], s_rstd[32];

  const int pps = (p.HW + nrank - 1) / nrank;
  const int p_begin = rank * pps;
  const int p_end = min(p.HW, p_begin + pps);
  const bool from0 = c < p.C0;
  const int ld = from0 ? p.C0 : p.C1;
  const __nv_bfloat16* src = from0 ? p.x0 + static_cast<size_t>(n) * p.HW * p.C0 + c
                                   : p.x1 + static_cast<size_t>(n) * p.HW * p.C1 + (c - p.C0);

  // ---- pass 1: per-thread channel sums over this CTA's slab.  Sums are taken of (x - K_c), K_c = the image's first
  //      pixel in that channel: E[x^2] - E[x]^2 in fp32 cancels catastrophically when |mean| >> std (real activations
  //      after a residual add); the shifted form does not (|K - mean| is of the order of std) and tracks torch's
  //      Welford-style F.group_norm to rounding.  Same pivot in every CTA / lane, so partial sums simply add; pixels
  //      past the slab's end are replaced by the pivot itself (contribute exactly zero: no select in the loop).
  float a[8], b[8], kv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = b[j] = kv[j] = 0.f;
  if (active) {
    const uint4 piv_u = *reinterpret_cast<const uint4*>(src);          // pixel 0 of image n, this thread's 8 channels
    unpack8(piv_u, kv);
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s_piv[lc + j] = kv[j];
    }
    for (int px = p_begin + lane; px < p_end; px += nlanes * kGnUnroll) {
      uint4 u[kGnUnroll];
#pragma unroll
      for (int k = 0; k < kGnUnroll; ++k) {
        const int pk = px + k * nlanes;
        u[k] = (pk < p_end) ? *reinterpret_cast<const uint4*>(src + static_cast<size_t>(pk) * ld) : piv_u;
      }
#pragma unroll
      for (int k = 0; k < kGnUnroll; ++k) {
        float f[8];
        unpack8(u[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = f[j] - kv[j];
          a[j] += d;
          b[j] = fmaf(d, d, b[j]);
        }
      }
    }
    float* ds = s_sum + lane * Cs + lc;
    float* dq = s_sq + lane * Cs + lc;
    *reinterpret_cast<float4*>(ds) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4*>(ds + 4) = make_float4(a[4], a[5], a[6], a[7]);
    *reinterpret_cast<float4*>(dq) = make_float4(b[0], b[1], b[2], b[3]);
    *reinterpret_cast<float4*>(dq + 4) = make_float4(b[4], b[5], b[6], b[7]);
  }
  __syncthreads();
  // lanes -> per-channel totals (fixed order), kept in lane 0's row
  for (int ch = tid; ch < Cs; ch += kGnThreads) {
    float s = 0.f, q = 0.f;
    for (int l = 0; l < nlanes; ++l) {
      s += s_sum[l * Cs + ch];
      q += s_sq[l * Cs + ch];
    }
    s_sum[ch] = s;
    s_sq[ch] = q;
  }
  __syncthreads();
  // channels -> groups, re-based on the group's first channel's pivot K_g:
  //   sum (x - K_g) = s_c + n d,  sum (x - K_g)^2 = q_c + 2 d s_c + n d^2,  d = K_c - K_g, n = pixels of this slab
  if (tid < gl) {
    const float npix = static_cast<float>(max(p_end - p_begin, 0));
    const float kg = s_piv[tid * cpg];
    float S = 0.f, Q = 0.f;
    for (int j = 0; j < cpg; ++j) {
      const int ch = tid * cpg + j;
      const float d = s_piv[ch] - kg, s = s_sum[ch];
      S += fmaf(npix, d, s);
      Q += s_sq[ch] + d * fmaf(npix, d, 2.f * s);
    }
    s_gpart[2 * tid] = S;
    s_gpart[2 * tid + 1] = Q;
  }
  cluster.sync();
  // ---- cluster exchange through DSMEM (fixed rank order)
  if (tid < 2 * gl) {
    float s = 0.f;
    for (int r = 0; r < nrank; ++r) s += cluster.map_shared_rank(s_gpart, r)[tid];
    s_gpart[64 + tid] = s;                          // local copy of the totals (peers read [0, 64) only)
  }
  __syncthreads();
  cluster.barrier_arrive();                         // peers may exit once everyone has read their partials
  if (tid < gl) {
    const float cnt = static_cast<float>(p.HW) * cpg;
    const float dm = s_gpart[64 + 2 * tid] / cnt;                    // mean of (x - K_g)
    const float var = fmaxf(s_gpart[64 + 2 * tid + 1] / cnt - dm * dm, 0.f);
    const float mean = s_piv[tid * cpg] + dm;
    s_mean[tid] = mean;
    s_rstd[tid] = rsqrtf(var + p.eps);
    if (p.stats != nullptr && rank == 0) {
      float* st = p.stats + (static_cast<size_t>(n) * p.groups + blockIdx.z * gl + tid) * 2;
      st[0] = mean;
      st[1] = s_rstd[tid];
    }
  }
  __syncthreads();

  // ---- pass 2: normalise + affine (+ SiLU)
  if (active) {
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (lc + j) / cpg;
      sc[j] = p.gamma[c + j] * s_rstd[g];
      sh[j] = p.beta[c + j] - s_mean[g] * sc[j];
    }
    __nv_bfloat16* dst = p.y + static_cast<size_t>(n) * p.HW * C + c;
    for (int px = p_begin + lane; px < p_end; px += nlanes * kGnUnroll) {
      uint4 u[kGnUnroll];
#pragma unroll
      for (int k = 0; k < kGnUnroll; ++k) {
        const int pk = px + k * nlanes;
        u[k] = (pk < p_end) ? *reinterpret_cast<const uint4*>(src + static_cast<size_t>(pk) * ld) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int k = 0; k < kGnUnroll; ++k) {
        const int pk = px + k * nlanes;
        if (pk < p_end) {
          float f[8];
          unpack8(u[k], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float val = fmaf(f[j], sc[j], sh[j]);
            if (p.silu) val = silu_fast(val);
            f[j] = val;
          }
          uint4 o;
          o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
          o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
          *reinterpret_cast<uint4*>(dst + static_cast<size_t>(pk) * C) = o;
        }
      }
    }
  }
  cluster.barrier_wait();
}

// One-pass GroupNorm (+SiLU): the statistics come from the PRODUCERS of x0 / x1 -- b200_conv_gemm_gnstat left, per image,
// per 32-pixel slab and per 4-channel unit, the (sum, sum of squares) of the values it stored -- so this kernel only merges
// those partials (in fp64: the fp32 partials cover <= 128 values each, the cancellation-prone steps are the merge and
// E[x^2] - E[x]^2) and streams the tensor once: no statistics pass, no cluster, no DSMEM exchange, any number of CTAs.
struct GnApplyParams {
  const __nv_bfloat16* x0;
  const __nv_bfloat16* x1;
  const float* st0;           // [nb, slabs0, C0 / 4, 2]
  const float* st1;           // [nb, slabs1, C1 / 4, 2]
  int C0, C1, HW, groups, slabs0, slabs1;
  const float* gamma;
  const float* beta;
  float eps;
  int silu;
  __nv_bfloat16* y;
};

// grid (ctas per image, NB)
__global__ void __launch_bounds__(kGnThreads)
groupnorm_apply_kernel(const GnApplyParams p) {
  pdl_launch_dependents();
  const int n = blockIdx.y;
  const int C = p.C0 + p.C1;
  const int cpg = C / p.groups;
  const int upg = cpg >> 2;                          // 4-channel units per group
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  __shared__ float s_mean[32], s_rstd[32];
  pdl_wait();
  // ---- group statistics: warp w merges groups w, w + 16 (lanes stride over (unit, slab); fixed-order xor tree)
  for (int g = warp; g < p.groups; g += kGnThreads / 32) {
    double S = 0.0, Q = 0.0;
    for (int uu = 0; uu < upg; ++uu) {
      const int c = (g * upg + uu) * 4;
      const bool from0 = c < p.C0;
      const int slabs = from0 ? p.slabs0 : p.slabs1;
      const int units = (from0 ? p.C0 : p.C1) >> 2;
      const float2* st = reinterpret_cast<const float2*>(from0 ? p.st0 : p.st1) +
                         static_cast<size_t>(n) * slabs * units + ((from0 ? c : c - p.C0) >> 2);
      for (int sl = lane; sl < slabs; sl += 32) {
        const float2 v = st[static_cast<size_t>(sl) * units];
        S += v.x;
        Q += v.y;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      S += __shfl_xor_sync(0xffffffffu, S, o);
      Q += __shfl_xor_sync(0xffffffffu, Q, o);
    }
    if (lane == 0) {
      const double cnt = static_cast<double>(p.HW) * cpg;
      const double mean = S / cnt;
      const double var = fmax(Q / cnt - mean * mean, 0.0);
      s_mean[g] = static_cast<float>(mean);
      s_rstd[g] = rsqrtf(static_cast<float>(var) + p.eps);
    }
  }
  __syncthreads();
  // ---- normalise + affine (+ SiLU), one streaming pass over this CTA's pixel slab
  const int vpp = C >> 3;
  const int nlanes = kGnThreads / vpp;
  if (tid >= nlanes * vpp) return;
  const int v = tid % vpp, pl = tid / vpp;
  const int c = v * 8;
  const bool from0 = c < p.C0;
  const int ld = from0 ? p.C0 : p.C1;
  const __nv_bfloat16* src = from0 ? p.x0 + static_cast<size_t>(n) * p.HW * p.C0 + c
                                   : p.x1 + static_cast<size_t>(n) * p.HW * p.C1 + (c - p.C0);
  const int pps = (p.HW + gridDim.x - 1) / gridDim.x;
  const int p_begin = blockIdx.x * pps;
  const int p_end = min(p.HW, p_begin + pps);
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (c + j) / cpg;
    sc[j] = p.gamma[c + j] * s_rstd[g];
    sh[j] = p.beta[c + j] - s_mean[g] * sc[j];
  }
  __nv_bfloat16* dst = p.y + static_cast<size_t>(n) * p.HW * C + c;
  for (int px = p_begin + pl; px < p_end; px += nlanes * kGnUnroll) {
    uint4 u[kGnUnroll];
#pragma unroll
    for (int k = 0; k < kGnUnroll; ++k) {
      const int pk = px + k * nlanes;
      u[k] = (pk < p_end) ? *reinterpret_cast<const uint4*>(src + static_cast<size_t>(pk) * ld) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < kGnUnroll; ++k) {
      const int pk = px + k * nlanes;
      if (pk < p_end) {
        float f[8];
        unpack8(u[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float val = fmaf(f[j], sc[j], sh[j]);
          if (p.silu) val = silu_fast(val);
          f[j] = val;
        }
        uint4 o;
        o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
        o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
        *reinterpret_cast<uint4*>(dst + static_cast<size_t>(pk) * C) = o;
      }
    }
  }
}

// Row softmax fp32 [rows, ld_s] -> bf16 [rows, ld_p] (columns >= cols written as zero up to cols_pad).  One 256-thread block
// per row; the row lives in registers between the three passes (max, sum of exp, normalise): cols <= 256 * 32.
// Used by the VAE decoder's single-head, head_dim-512 mid-block attention (AutoencoderKL.decode, loaded at
// /root/reference/script/train/train_audioldm_lora.py:370), whose scores S = Q K^T and O = P V run as b200_conv_gemm
// launches: head_dim 512 does not fit the fused attention kernel's TMEM budget, and the block runs once per clip.
static constexpr int kSmThreads = 256;
static constexpr int kSmMaxPer = 32;
__global__ void __launch_bounds__(kSmThreads)
softmax_rows_kernel(const float* __restrict__ s, int cols, int cols_pad, size_t ld_s, __nv_bfloat16* __restrict__ pout,
                    size_t ld_p, float scale_log2) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t row = blockIdx.x;
  const float* src = s + row * ld_s;
  __nv_bfloat16* dst = pout + row * ld_p;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  __shared__ float red[kSmThreads / 32];
  float v[kSmMaxPer];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < kSmMaxPer; ++i) {
    const int c = tid + i * kSmThreads;
    v[i] = c < cols ? src[c] * scale_log2 : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < kSmThreads / 32; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kSmMaxPer; ++i) {
    v[i] = exp2f(v[i] - mx);          // exp2(-inf) = 0 for the padding
    sum += v[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < kSmThreads / 32; ++w) sum += red[w];      // fixed order: deterministic
  const float inv = 1.0f / sum;
#pragma unroll
  for (int i = 0; i < kSmMaxPer; ++i) {
    const int c = tid + i * kSmThreads;
    if (c < cols_pad) dst[c] = __float2bfloat16(c < cols ? v[i] * inv : 0.f);
  }
}

// One warp per row, kLnRows rows per warp in flight (all loads issued before the first reduction: the kernel is bound by
// bytes in flight per SM, not by arithmetic); C % 8 == 0, C <= 1280.  y = (x - mean) * rstd * gamma + beta  (eps inside sqrt)
static constexpr int kLnMaxVec = 5;     // 5 * 32 lanes * 8 = 1280 channels
static constexpr int kLnRows = 2;      // rows per warp: 2 in flight for C <= 512, else the second row follows the first

template <int NV, int ROWS>
__device__ __forceinline__ void layernorm_rows(const __nv_bfloat16* __restrict__ x, int row0, int M, int C,
                                               const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                               __nv_bfloat16* __restrict__ y, int lane) {
  const int nvec = C / 8;
  uint4 u[ROWS][NV];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int row = row0 + r;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      u[r][i] = (row < M && v < nvec) ? *reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * C + v * 8)
                                      : make_uint4(0, 0, 0, 0);
    }
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int row = row0 + r;
    if (row >= M) break;                 // warp-uniform
    float f[NV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      unpack8(u[r][i], f[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[i][j];       // lanes past nvec hold zeros
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + i * 32 < nvec) {
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
    __nv_bfloat16* out = y + static_cast<size_t>(row) * C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
        const float4 g0 = *reinterpret_cast<const float4*>(gamma + v * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(gamma + v * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(beta + v * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(beta + v * 8 + 4);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o8[j] = (f[i][j] - mean) * rstd * gg[j] + bb[j];
        uint4 o;
        o.x = pack_bf16x2(o8[0], o8[1]); o.y = pack_bf16x2(o8[2], o8[3]);
        o.z = pack_bf16x2(o8[4], o8[5]); o.w = pack_bf16x2(o8[6], o8[7]);
        *reinterpret_cast<uint4*>(out + v * 8) = o;
      }
    }
  }
}

template <int ROWS>
__global__ void __launch_bounds__(256, 3)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, int M, int C, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();
  const int row0 = warp * ROWS;
  if (row0 >= M) return;
  const int nvec = C / 8;
  if (nvec <= 32) layernorm_rows<1, ROWS>(x, row0, M, C, gamma, beta, eps, y, lane);
  else if (nvec <= 64) layernorm_rows<2, ROWS>(x, row0, M, C, gamma, beta, eps, y, lane);
  else if (nvec <= 96) layernorm_rows<3, 1>(x, row0, M, C, gamma, beta, eps, y, lane);      // host: ROWS == 1 for C > 512
  else layernorm_rows<kLnMaxVec, 1>(x, row0, M, C, gamma, beta, eps, y, lane);
}

// Token embedding of a RoBERTa-style text encoder (the CLAP text tower in front of the pipeline):
//   y[row] = LayerNorm(word[ids[row]] + pos[pos_ids[row]] + type0) as bf16; fp32 tables, one warp per token row, C % 4 == 0.
// Three passes over the (L2-resident) table rows instead of registers: the row count is a few hundred per call.
__global__ void __launch_bounds__(256)
embed_layernorm_kernel(const int* __restrict__ ids, const int* __restrict__ pos_ids, int M, int C, int vocab, int npos,
                       const float* __restrict__ word, const float* __restrict__ pos, const float* __restrict__ type0,
                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                       __nv_bfloat16* __restrict__ y) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();
  if (row >= M) return;
  const int id = min(max(ids[row], 0), vocab - 1), pi = min(max(pos_ids[row], 0), npos - 1);
  const float4* w4 = reinterpret_cast<const float4*>(word + static_cast<size_t>(id) * C);
  const float4* p4 = reinterpret_cast<const float4*>(pos + static_cast<size_t>(pi) * C);
  const float4* t4 = reinterpret_cast<const float4*>(type0);
  const int nvec = C / 4;
  auto load = [&](int v) {
    const float4 a = w4[v], b = p4[v], c = t4[v];
    // the reference's order of additions: (word + type) + position
    return make_float4((a.x + c.x) + b.x, (a.y + c.y) + b.y, (a.z + c.z) + b.z, (a.w + c.w) + b.w);
  };
  float s = 0.f;
  for (int v = lane; v < nvec; v += 32) {
    const float4 e = load(v);
    s += (e.x + e.y) + (e.z + e.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float q = 0.f;
  for (int v = lane; v < nvec; v += 32) {
    const float4 e = load(v);
    const float d0 = e.x - mean, d1 = e.y - mean, d2 = e.z - mean, d3 = e.w - mean;
    q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / C + eps);
  __nv_bfloat16* out = y + static_cast<size_t>(row) * C;
  for (int v = lane; v < nvec; v += 32) {
    const float4 e = load(v);
    const float4 g = *reinterpret_cast<const float4*>(gamma + v * 4);
    const float4 b = *reinterpret_cast<const float4*>(beta + v * 4);
    uint2 o;
    o.x = pack_bf16x2((e.x - mean) * rstd * g.x + b.x, (e.y - mean) * rstd * g.y + b.y);
    o.y = pack_bf16x2((e.z - mean) * rstd * g.z + b.z, (e.w - mean) * rstd * g.w + b.w);
    *reinterpret_cast<uint2*>(out + v * 4) = o;
  }
}

}  // namespace b200

using namespace b200;

// SM share of the calling chain (0 = the whole GPU).  When several independent sub-batch chains run concurrently on
// parallel streams, a GroupNorm launch sized for all 148 SMs (8-CTA clusters x images x group splits) serialises the
// chains (measured: tools/concurrency_probe2.py); the host sets the budget around a chain's launches.
static thread_local int g_sm_budget = 0;
extern "C" int b200_set_sm_budget(int n) {
  g_sm_budget = n > 0 ? n : 0;
  return B200_OK;
}

static int groupnorm_impl(const void* x0, int c0, const void* x1, int c1, int nb, int hw, int groups, const float* gamma,
                          const float* beta, float eps, int silu, void* y, float* stats, void* stream_v);

extern "C" int b200_groupnorm_silu(const void* x0, int c0, const void* x1, int c1, int nb, int hw, int groups,
                                   const float* gamma, const float* beta, float eps, int silu, void* y,
                                   void* stream_v) {
  return groupnorm_impl(x0, c0, x1, c1, nb, hw, groups, gamma, beta, eps, silu, y, nullptr, stream_v);
}
// Training form: also writes (mean, rstd) per (image, group) for b200_groupnorm_silu_bwd.
extern "C" int b200_groupnorm_silu_stats(const void* x0, int c0, const void* x1, int c1, int nb, int hw, int groups,
                                         const float* gamma, const float* beta, float eps, int silu, void* y,
                                         float* stats, void* stream_v) {
  B200_CHECK_ARG(stats, "groupnorm_stats: null stats");
  return groupnorm_impl(x0, c0, x1, c1, nb, hw, groups, gamma, beta, eps, silu, y, stats, stream_v);
}

static int groupnorm_impl(const void* x0, int c0, const void* x1, int c1, int nb, int hw, int groups, const float* gamma,
                          const float* beta, float eps, int silu, void* y, float* stats, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const int C = c0 + c1;
  B200_CHECK_ARG(x0 && y && gamma && beta, "groupnorm: null pointer");
  B200_CHECK_ARG(c0 % 8 == 0 && c1 % 8 == 0 && C <= kMaxC, "groupnorm: channels (%d,%d) unsupported", c0, c1);
  B200_CHECK_ARG((c1 == 0) == (x1 == nullptr), "groupnorm: second source mismatch");
  B200_CHECK_ARG(groups > 0 && groups <= 32 && C % groups == 0, "groupnorm: %d channels not divisible by %d groups", C, groups);
  B200_CHECK_ARG(nb > 0 && hw > 0, "groupnorm: empty input");
  GnParams p;
  p.x0 = reinterpret_cast<const __nv_bfloat16*>(x0); p.x1 = reinterpret_cast<const __nv_bfloat16*>(x1);
  p.C0 = c0; p.C1 = c1; p.HW = hw; p.groups = groups; p.gamma = gamma; p.beta = beta; p.eps = eps; p.silu = silu;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.stats = stats;
  // Launch shape: cluster size cs (pixel slabs of one image) x gsplit (independent halves / quarters of the
  // groups), the combination with the most CTAs whose clusters are all co-resident (one wave).
  static int max_active[4] = {-1, -1, -1, -1};      // for cs = 1, 2, 4, 8
  static const int max_cs = getenv("B200_GN_CLUSTER") ? atoi(getenv("B200_GN_CLUSTER")) : 8;   // debugging knob
  static const int min_pix = getenv("B200_GN_MINPIX") ? atoi(getenv("B200_GN_MINPIX")) : 4;
  int best_cs = 1, best_gs = 1, best_ctas = 0;
  for (int ci = 3; ci >= 0; --ci) {
    const int cs = 1 << ci;
    if (cs > max_cs || (cs > 1 && hw < min_pix * cs)) continue;     // a slab of fewer pixels per CTA is all sync, no work
    if (max_active[ci] < 0) {
      cudaLaunchConfig_t qc;
      memset(&qc, 0, sizeof(qc));
      qc.gridDim = dim3(cs, 1, 1);
      qc.blockDim = dim3(kGnThreads, 1, 1);
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = cs; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      qc.attrs = qa;
      qc.numAttrs = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, groupnorm_silu_kernel, &qc) != cudaSuccess || nclusters <= 0) {
        cudaGetLastError();
        nclusters = 0;
      }
      max_active[ci] = nclusters;
    }
    for (int gs = 1; gs <= 4; gs *= 2) {
      if (groups % gs != 0 || (C / gs) % 8 != 0) continue;
      if (nb * gs > max_active[ci]) continue;
      const int ctas = nb * gs * cs;
      if (g_sm_budget > 0 && ctas > g_sm_budget && ctas > nb) continue;
      if (ctas > best_ctas) {                        // ties keep the larger cluster / smaller split seen first
        best_ctas = ctas; best_cs = cs; best_gs = gs;
      }
    }
  }
  if (best_ctas == 0) { best_cs = 1; best_gs = 1; }  // more images than co-resident CTAs: plain multi-wave launch
  B200_CHECK_PDL("groupnorm", launch_pdl(groupnorm_silu_kernel, dim3(best_cs, nb, best_gs), dim3(kGnThreads), 0, stream,
                                         best_cs, p));
  return B200_OK;
}

// One-pass GroupNorm (+SiLU) over cat(x0, x1) from producer-side statistics (see b200_conv_gemm_gnstat).
extern "C" int b200_groupnorm_apply(const void* x0, int c0, const float* st0, int slabs0, const void* x1, int c1,
                                    const float* st1, int slabs1, int nb, int hw, int groups, const float* gamma,
                                    const float* beta, float eps, int silu, void* y, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const int C = c0 + c1;
  B200_CHECK_ARG(x0 && y && gamma && beta && st0 && slabs0 > 0, "groupnorm_apply: null pointer");
  B200_CHECK_ARG((c1 == 0) == (x1 == nullptr) && (c1 == 0 || (st1 && slabs1 > 0)), "groupnorm_apply: second source mismatch");
  B200_CHECK_ARG(c0 % 8 == 0 && c1 % 8 == 0 && C <= kMaxC, "groupnorm_apply: channels (%d,%d) unsupported", c0, c1);
  B200_CHECK_ARG(groups > 0 && groups <= 32 && C % groups == 0 && (C / groups) % 4 == 0 && c0 % 4 == 0,
                 "groupnorm_apply: %d channels / %d groups must give groups of a multiple of 4 channels", C, groups);
  B200_CHECK_ARG(nb > 0 && hw > 0, "groupnorm_apply: empty input");
  GnApplyParams p;
  p.x0 = reinterpret_cast<const __nv_bfloat16*>(x0); p.x1 = reinterpret_cast<const __nv_bfloat16*>(x1);
  p.st0 = st0; p.st1 = st1; p.slabs0 = slabs0; p.slabs1 = slabs1;
  p.C0 = c0; p.C1 = c1; p.HW = hw; p.groups = groups; p.gamma = gamma; p.beta = beta; p.eps = eps; p.silu = silu;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  // CTAs per image: ~kGnUnroll 16-byte vectors per thread, at most ~4 CTAs per SM over the whole launch
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const long vectors = static_cast<long>(hw) * (C / 8);
  long per_image = (vectors + kGnThreads * kGnUnroll - 1) / (kGnThreads * kGnUnroll);
  const int budget = g_sm_budget > 0 ? g_sm_budget : num_sms;
  const long cap = (4L * budget + nb - 1) / nb;
  if (per_image > cap) per_image = cap;
  if (per_image < 1) per_image = 1;
  B200_CHECK_PDL("groupnorm_apply", launch_pdl(groupnorm_apply_kernel, dim3((unsigned)per_image, nb), dim3(kGnThreads), 0,
                                               stream, 0, p));
  return B200_OK;
}

// p[r, c] = softmax_c(scale * s[r, c]) for c < cols, 0 for cols <= c < cols_pad.  s fp32 (row stride ld_s), p bf16 (ld_p).
extern "C" int b200_softmax_rows(const float* s, int rows, int cols, int cols_pad, long ld_s, void* p, long ld_p, float scale,
                                 void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(s && p && rows > 0 && cols > 0 && cols_pad >= cols && ld_s >= cols && ld_p >= cols_pad, "softmax_rows: bad args");
  B200_CHECK_ARG(cols_pad <= kSmThreads * kSmMaxPer, "softmax_rows: at most %d columns", kSmThreads * kSmMaxPer);
  B200_CHECK_PDL("softmax_rows", launch_pdl(softmax_rows_kernel, dim3(rows), dim3(kSmThreads), 0, stream, 0, s, cols, cols_pad,
                                            static_cast<size_t>(ld_s), reinterpret_cast<__nv_bfloat16*>(p),
                                            static_cast<size_t>(ld_p), scale * 1.4426950408889634f));
  return B200_OK;
}

extern "C" int b200_layernorm(const void* x, int m, int c, const float* gamma, const float* beta, float eps, void* y,
                              void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(x && y && gamma && beta, "layernorm: null pointer");
  B200_CHECK_ARG(c % 8 == 0 && c <= kLnMaxVec * 256, "layernorm: C=%d unsupported", c);
  if (m == 0) return B200_OK;
  const int warps_per_block = 8;
  // Two rows per warp in flight once there are more rows than resident warps (measured on B200: 7.4 -> 6.7 us at
  // m = 16000, c = 256); below that one row per warp keeps more warps on the machine (3.1 vs 3.6 us at m = 1024).
  const int rows = (m >= 8192 && c <= 512) ? kLnRows : 1;
  const int rows_per_block = warps_per_block * rows;
  const int nblk = (m + rows_per_block - 1) / rows_per_block;
  if (rows == 1)
    B200_CHECK_PDL("layernorm", launch_pdl(layernorm_kernel<1>, dim3(nblk), dim3(warps_per_block * 32), 0, stream, 0,
                                           reinterpret_cast<const __nv_bfloat16*>(x), m, c, gamma, beta, eps,
                                           reinterpret_cast<__nv_bfloat16*>(y)));
  else
    B200_CHECK_PDL("layernorm", launch_pdl(layernorm_kernel<kLnRows>, dim3(nblk), dim3(warps_per_block * 32), 0, stream, 0,
                                           reinterpret_cast<const __nv_bfloat16*>(x), m, c, gamma, beta, eps,
                                           reinterpret_cast<__nv_bfloat16*>(y)));
  return B200_OK;
}

extern "C" int b200_embed_layernorm(const int* ids, const int* pos_ids, int m, int c, int vocab, int npos, const float* word,
                                    const float* pos, const float* type0, const float* gamma, const float* beta, float eps,
                                    void* y, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(ids && pos_ids && word && pos && type0 && gamma && beta && y, "embed_layernorm: null pointer");
  B200_CHECK_ARG(c > 0 && c % 4 == 0 && vocab > 0 && npos > 0, "embed_layernorm: C=%d vocab=%d npos=%d unsupported", c, vocab, npos);
  if (m == 0) return B200_OK;
  const int rows_per_block = 8;
  B200_CHECK_PDL("embed_layernorm", launch_pdl(embed_layernorm_kernel, dim3((m + rows_per_block - 1) / rows_per_block),
                                               dim3(rows_per_block * 32), 0, stream, 0, ids, pos_ids, m, c, vocab, npos, word, pos,
                                               type0, gamma, beta, eps, reinterpret_cast<__nv_bfloat16*>(y)));
  return B200_OK;
}
