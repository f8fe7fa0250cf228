// Backward-pass kernels of the LoRA fine-tuning step for sm_100a.
//
// The base UNet is frozen (train_audioldm_lora.py:373-376): the backward pass needs activation gradients
// (dgrad) everywhere downstream of the first adapted attention layer and weight gradients only for the rank-r
// LoRA matrices.  dgrad GEMMs / convolutions reuse b200_conv_gemm with transposed (and tap-flipped) packed
// weights; the kernels here are everything else:
//   * groupnorm_silu_bwd  -- d/dx of SiLU(GroupNorm(x)) over one or two NHWC sources, cluster per image (DSMEM)
//   * layernorm_bwd       -- d/dx of LayerNorm, warp per row, residual-stream gradient added in the same pass
//   * geglu_fwd / _bwd    -- GEGLU with the pre-activation kept for the backward pass
//   * lora_wgrad          -- dA = dT^T x, dB = s dY^T T: tall-skinny reductions over the token dimension written
//                            straight into the flat fp32 LoRA-gradient arena (the buffer NCCL all-reduces, C1)
//   * zero_insert / upsample_nearest_bwd / add_bf16 / mse_grad / lora_refresh
// All HBM/L2-bound: 16-byte vector accesses, fp32 statistics, fixed-order reductions except the final
// cross-CTA fp32 atomics of lora_wgrad and the loss sum.
//
// Replaces torch autograd under `accelerator.backward(loss)` --
// /root/reference/script/train/train_audioldm_lora.py:549,557 (loss, backward), :563 (optimizer step).
#include <cooperative_groups.h>

#include "host_util.h"
#include "ptx.cuh"

namespace cg = cooperative_groups;

namespace b200 {

static constexpr int kGnbThreads = 512;
static constexpr int kGnbUnroll = 4;

__device__ __forceinline__ void unpack8t(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8t(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
  return o;
}

// ---------------------------------------------------------------------------------- GroupNorm(+SiLU) backward
struct GnBwdParams {
  const __nv_bfloat16* x0;
  const __nv_bfloat16* x1;
  int C0, C1, HW, groups;
  const float* gamma;
  const float* beta;
  const float* stats;          // [nb, groups, 2] = (mean, rstd) saved by the forward kernel
  int silu;
  const __nv_bfloat16* dy;     // [nb*HW, C0+C1]
  const __nv_bfloat16* dres;   // optional gradient added to dx: [nb*HW, res_ld], channel c of the concatenation at column c
  int res_ld;
  __nv_bfloat16* dx0;          // [nb*HW, C0]
  __nv_bfloat16* dx1;          // [nb*HW, C1] or null (gradient of the second source not needed)
};

// With xh = (x - mean) rstd, yh = gamma xh + beta, g = dy * silu'(yh), dxh = g gamma:
//   dx = rstd * (dxh - mean_group(dxh) - xh * mean_group(dxh * xh))
// grid (cluster_size, NB, gsplit) exactly like the forward kernel.
__global__ void __launch_bounds__(kGnbThreads)
groupnorm_silu_bwd_kernel(const GnBwdParams p) {
  cg::cluster_group cluster = cg::this_cluster();
  pdl_launch_dependents();
  pdl_wait();
  const int nrank = gridDim.x;
  const int rank = blockIdx.x;
  const int n = blockIdx.y;
  const int C = p.C0 + p.C1;
  const int Cs = C / gridDim.z;
  const int c_base = blockIdx.z * Cs;
  const int gl = p.groups / gridDim.z;
  const int vpp = Cs >> 3;
  const int nlanes = kGnbThreads / vpp;
  const int tid = threadIdx.x;
  const bool active = tid < nlanes * vpp;
  const int v = tid % vpp;
  const int lane = tid / vpp;
  const int lc = v * 8;
  const int c = c_base + lc;
  const int cpg = C / p.groups;

  __shared__ float s_a[kGnbThreads * 8];
  __shared__ float s_b[kGnbThreads * 8];
  __shared__ float s_gpart[128];
  __shared__ float s_m1[32], s_m2[32];

  const int pps = (p.HW + nrank - 1) / nrank;
  const int p_begin = rank * pps;
  const int p_end = min(p.HW, p_begin + pps);
  const bool from0 = c < p.C0;
  const int ld = from0 ? p.C0 : p.C1;
  const __nv_bfloat16* src = from0 ? p.x0 + static_cast<size_t>(n) * p.HW * p.C0 + c
                                   : p.x1 + static_cast<size_t>(n) * p.HW * p.C1 + (c - p.C0);
  const __nv_bfloat16* dyp = p.dy + static_cast<size_t>(n) * p.HW * C + c;

  // per-channel constants: xh = x * rs + xo ; yh = xh * gm + bt
  float rs[8], xo[8], gm[8], bt[8];
  if (active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (c + j) / cpg;
      const float mean = p.stats[(static_cast<size_t>(n) * p.groups + g) * 2];
      const float rstd = p.stats[(static_cast<size_t>(n) * p.groups + g) * 2 + 1];
      rs[j] = rstd;
      xo[j] = -mean * rstd;
      gm[j] = p.gamma[c + j];
      bt[j] = p.beta[c + j];
    }
  }
  auto dxh_of = [&](const float (&xf)[8], const float (&df)[8], float (&xh)[8], float (&dxh)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      xh[j] = fmaf(xf[j], rs[j], xo[j]);
      float g = df[j];
      if (p.silu) {
        const float yh = fmaf(xh[j], gm[j], bt[j]);
        const float sg = __fdividef(1.0f, 1.0f + __expf(-yh));
        g *= sg * fmaf(yh, 1.0f - sg, 1.0f);
      }
      dxh[j] = g * gm[j];
    }
  };

  // ---- pass 1: per-channel sums of dxh and dxh * xh over this CTA's slab
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = b[j] = 0.f;
  if (active) {
    for (int px = p_begin + lane; px < p_end; px += nlanes * kGnbUnroll) {
      uint4 ux[kGnbUnroll], ud[kGnbUnroll];
#pragma unroll
      for (int k = 0; k < kGnbUnroll; ++k) {
        const int pk = px + k * nlanes;
        ux[k] = (pk < p_end) ? *reinterpret_cast<const uint4*>(src + static_cast<size_t>(pk) * ld) : make_uint4(0, 0, 0, 0);
        ud[k] = (pk < p_end) ? *reinterpret_cast<const uint4*>(dyp + static_cast<size_t>(pk) * C) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int k = 0; k < kGnbUnroll; ++k) {
        float xf[8], df[8], xh[8], dxh[8];
        unpack8t(ux[k], xf);
        unpack8t(ud[k], df);       // dy == 0 for padded pixels => no contribution
        dxh_of(xf, df, xh, dxh);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a[j] += dxh[j];
          b[j] = fmaf(dxh[j], xh[j], b[j]);
        }
      }
    }
    float* ds = s_a + lane * Cs + lc;
    float* dq = s_b + lane * Cs + lc;
    *reinterpret_cast<float4*>(ds) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4*>(ds + 4) = make_float4(a[4], a[5], a[6], a[7]);
    *reinterpret_cast<float4*>(dq) = make_float4(b[0], b[1], b[2], b[3]);
    *reinterpret_cast<float4*>(dq + 4) = make_float4(b[4], b[5], b[6], b[7]);
  }
  __syncthreads();
  for (int ch = tid; ch < Cs; ch += kGnbThreads) {
    float s = 0.f, q = 0.f;
    for (int l = 0; l < nlanes; ++l) {
      s += s_a[l * Cs + ch];
      q += s_b[l * Cs + ch];
    }
    s_a[ch] = s;
    s_b[ch] = q;
  }
  __syncthreads();
  if (tid < 2 * gl) {
    const int g = tid >> 1;
    const float* srcv = (tid & 1) ? s_b : s_a;
    float s = 0.f;
    for (int j = 0; j < cpg; ++j) s += srcv[g * cpg + j];
    s_gpart[tid] = s;
  }
  cluster.sync();
  if (tid < 2 * gl) {
    float s = 0.f;
    for (int r = 0; r < nrank; ++r) s += cluster.map_shared_rank(s_gpart, r)[tid];
    s_gpart[64 + tid] = s;
  }
  __syncthreads();
  cluster.barrier_arrive();
  if (tid < gl) {
    const float inv = 1.0f / (static_cast<float>(p.HW) * cpg);
    s_m1[tid] = s_gpart[64 + 2 * tid] * inv;
    s_m2[tid] = s_gpart[64 + 2 * tid + 1] * inv;
  }
  __syncthreads();

  // ---- pass 2: dx = rstd (dxh - m1 - xh m2) (+ dres)
  if (active && (from0 || p.dx1 != nullptr)) {
    float m1[8], m2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (lc + j) / cpg;
      m1[j] = s_m1[g];
      m2[j] = s_m2[g];
    }
    __nv_bfloat16* dst = from0 ? p.dx0 + static_cast<size_t>(n) * p.HW * p.C0 + c
                               : p.dx1 + static_cast<size_t>(n) * p.HW * p.C1 + (c - p.C0);
    const __nv_bfloat16* rp = p.dres ? p.dres + static_cast<size_t>(n) * p.HW * p.res_ld + c : nullptr;
    for (int px = p_begin + lane; px < p_end; px += nlanes * kGnbUnroll) {
      uint4 ux[kGnbUnroll], ud[kGnbUnroll], ur[kGnbUnroll];
#pragma unroll
      for (int k = 0; k < kGnbUnroll; ++k) {
        const int pk = px + k * nlanes;
        const bool ok = pk < p_end;
        ux[k] = ok ? *reinterpret_cast<const uint4*>(src + static_cast<size_t>(pk) * ld) : make_uint4(0, 0, 0, 0);
        ud[k] = ok ? *reinterpret_cast<const uint4*>(dyp + static_cast<size_t>(pk) * C) : make_uint4(0, 0, 0, 0);
        ur[k] = (ok && rp) ? *reinterpret_cast<const uint4*>(rp + static_cast<size_t>(pk) * p.res_ld) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int k = 0; k < kGnbUnroll; ++k) {
        const int pk = px + k * nlanes;
        if (pk < p_end) {
          float xf[8], df[8], xh[8], dxh[8], rf[8], o[8];
          unpack8t(ux[k], xf);
          unpack8t(ud[k], df);
          unpack8t(ur[k], rf);
          dxh_of(xf, df, xh, dxh);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(rs[j], dxh[j] - m1[j] - xh[j] * m2[j], rf[j]);
          *reinterpret_cast<uint4*>(dst + static_cast<size_t>(pk) * ld) = pack8t(o);
        }
      }
    }
  }
  cluster.barrier_wait();
}

// ---------------------------------------------------------------------------------- LayerNorm backward
// One warp per row.  dx = rstd (dxh - mean(dxh) - xh mean(dxh xh)) + dres, dxh = dy gamma.  Statistics recomputed
// from the saved input row (registers), exactly as the forward kernel computes them.
static constexpr int kLnbMaxVec = 5;
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, int M, int C,
                     const float* __restrict__ gamma, float eps, const __nv_bfloat16* __restrict__ dres,
                     __nv_bfloat16* __restrict__ dx) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();
  if (warp >= M) return;
  const int nvec = C / 8;
  const size_t roff = static_cast<size_t>(warp) * C;
  float f[kLnbMaxVec][8], d[kLnbMaxVec][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnbMaxVec; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      unpack8t(*reinterpret_cast<const uint4*>(x + roff + v * 8), f[i]);
      unpack8t(*reinterpret_cast<const uint4*>(dy + roff + v * 8), d[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[i][j];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kLnbMaxVec; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = f[i][j] - mean;
        q += t * t;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / C + eps);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < kLnbMaxVec; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + v * 8);
      const float4 g1 = *reinterpret_cast<const float4*>(gamma + v * 8 + 4);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f[i][j] = (f[i][j] - mean) * rstd;     // xh
        d[i][j] *= gg[j];                      // dxh
        s1 += d[i][j];
        s2 = fmaf(d[i][j], f[i][j], s2);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const float m1 = s1 / C, m2 = s2 / C;
#pragma unroll
  for (int i = 0; i < kLnbMaxVec; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (dres) unpack8t(*reinterpret_cast<const uint4*>(dres + roff + v * 8), r);
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(rstd, d[i][j] - m1 - f[i][j] * m2, r[j]);
      *reinterpret_cast<uint4*>(dx + roff + v * 8) = pack8t(o);
    }
  }
}

// ---------------------------------------------------------------------------------- GEGLU (training form)
// h [M, 2F] = [values | gates] (the un-fused ff.net.0.proj output, kept for the backward pass)
__device__ __forceinline__ float gelu_exact(float g) { return 0.5f * g * (1.0f + erff(g * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_exact_grad(float g) {
  return 0.5f * (1.0f + erff(g * 0.70710678118654752f)) + g * 0.3989422804014327f * __expf(-0.5f * g * g);
}
__global__ void __launch_bounds__(256)
geglu_fwd_kernel(const __nv_bfloat16* __restrict__ h, size_t M, int F, __nv_bfloat16* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int vpr = F >> 3;
  const size_t total = M * vpr;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t row = i / vpr;
    const int col = static_cast<int>(i % vpr) * 8;
    float v[8], g[8], o[8];
    unpack8t(*reinterpret_cast<const uint4*>(h + row * 2 * F + col), v);
    unpack8t(*reinterpret_cast<const uint4*>(h + row * 2 * F + F + col), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = v[j] * gelu_exact(g[j]);
    *reinterpret_cast<uint4*>(out + row * F + col) = pack8t(o);
  }
}
__global__ void __launch_bounds__(256)
geglu_bwd_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ dout, size_t M, int F,
                 __nv_bfloat16* __restrict__ dh) {
  pdl_launch_dependents();
  pdl_wait();
  const int vpr = F >> 3;
  const size_t total = M * vpr;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t row = i / vpr;
    const int col = static_cast<int>(i % vpr) * 8;
    float v[8], g[8], d[8], dv[8], dg[8];
    unpack8t(*reinterpret_cast<const uint4*>(h + row * 2 * F + col), v);
    unpack8t(*reinterpret_cast<const uint4*>(h + row * 2 * F + F + col), g);
    unpack8t(*reinterpret_cast<const uint4*>(dout + row * F + col), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dv[j] = d[j] * gelu_exact(g[j]);
      dg[j] = d[j] * v[j] * gelu_exact_grad(g[j]);
    }
    *reinterpret_cast<uint4*>(dh + row * 2 * F + col) = pack8t(dv);
    *reinterpret_cast<uint4*>(dh + row * 2 * F + F + col) = pack8t(dg);
  }
}

// ---------------------------------------------------------------------------------- LoRA weight gradients
// G[c, j] += scale * sum_m U[m, c] * V[m, j]   (c < C, j < r), written with fp32 atomics at out[c*ldc + j*ldj].
//   dB [Cout, r] = s * dY^T T      : U = dY, V = T,  ldc = r, ldj = 1
//   dA [r, Cin]  = dT^T x          : U = x,  V = dT, ldc = 1, ldj = Cin      (dT = s dY B already carries s)
// One launch processes up to 8 (U, V) pairs (blockIdx.z): the q/k/v/out adapters of one attention module.
struct WgradDesc {
  const __nv_bfloat16* u;
  const __nv_bfloat16* v;
  float* out;
  int ldu, ldv, C, r, ldc, ldj;
  float scale;
};
struct WgradParams {
  WgradDesc d[8];
  int n;
  int M;
  int rows_per_cta;
};
static constexpr int kWgThreads = 128;
static constexpr int kWgRows = 64;      // rows staged per smem tile
// CTA = 64 channels x r outputs over a slab of rows.  Thread = 8 channels x nj (<= 2) columns; for r < 16 the spare
// thread groups take interleaved rows of the tile (row groups), so all 128 threads have work at rank 2 as at rank 32.
// U is converted to fp32 once while it is staged; the inner loop is 2 LDS.128 + nj LDS.32 per 8 nj FMAs.
__global__ void __launch_bounds__(kWgThreads)
lora_wgrad_kernel(const WgradParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const WgradDesc& d = p.d[blockIdx.z];
  const int c0 = blockIdx.x * 64;
  if (c0 >= d.C) return;
  const int m_begin = blockIdx.y * p.rows_per_cta;
  const int m_end = min(p.M, m_begin + p.rows_per_cta);
  __shared__ __align__(16) float sU[kWgRows][64];
  __shared__ __align__(16) float sV[kWgRows][32];
  const int tid = threadIdx.x;
  const int c8 = (tid & 7) * 8;
  const int g = tid >> 3;                        // 0..15
  int jgroups = 1;
  while (jgroups < d.r && jgroups < 16) jgroups <<= 1;
  const int nj = (d.r + jgroups - 1) / jgroups;  // 1 or 2
  const int rowgroups = 16 / jgroups;
  const int jg = g % jgroups, rg = g / jgroups;
  const int j0 = jg * nj;
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
  for (int m0 = m_begin; m0 < m_end; m0 += kWgRows) {
    for (int i = tid; i < kWgRows * 8; i += kWgThreads) {
      const int r = i >> 3, cv = (i & 7) * 8;
      const int m = m0 + r;
      float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (m < m_end && c0 + cv < d.C) unpack8t(*reinterpret_cast<const uint4*>(d.u + static_cast<size_t>(m) * d.ldu + c0 + cv), f);
      *reinterpret_cast<float4*>(&sU[r][cv]) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(&sU[r][cv + 4]) = make_float4(f[4], f[5], f[6], f[7]);
    }
    for (int i = tid; i < kWgRows * 32; i += kWgThreads) {
      const int r = i >> 5, j = i & 31;
      const int m = m0 + r;
      sV[r][j] = (m < m_end && j < d.r) ? __bfloat162float(d.v[static_cast<size_t>(m) * d.ldv + j]) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = rg; r < kWgRows; r += rowgroups) {
      const float4 ua = *reinterpret_cast<const float4*>(&sU[r][c8]);
      const float4 ub = *reinterpret_cast<const float4*>(&sU[r][c8 + 4]);
      const float u[8] = {ua.x, ua.y, ua.z, ua.w, ub.x, ub.y, ub.z, ub.w};
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (j < nj) {
          const float vv = sV[r][j0 + j];       // columns >= r are zero
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(u[i], vv, acc[j][i]);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int jj = j0 + j;
    if (j >= nj || jj >= d.r) continue;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + c8 + i;
      if (c < d.C) atomicAdd(d.out + static_cast<size_t>(c) * d.ldc + static_cast<size_t>(jj) * d.ldj, acc[j][i] * d.scale);
    }
  }
}

// ---------------------------------------------------------------------------------- LoRA weight gradients on tcgen05
// The same reduction as a tensor-core GEMM over the token dimension:  G[128 channels x N] += U_tile^T . V_tile  with
// BOTH operands MN-major straight from their row-major [tokens, channels] tensors (no transposed copies): a token block
// is fetched by TMA as 8x16-byte core matrices (4-D box: 8 elements | 64 tokens | 16-byte channel chunks), i.e. the
// no-swizzle canonical layout whose M/N index is the contiguous one -- the layout V has in the attention kernels.
// CTA = 128 channels of one (U, V) pair over a slab of tokens; TMEM accumulator of N <= 32 columns; fp32 atomics into
// the flat gradient arena at the end.  Bound by streaming U once (~45 B/clk/SM): ~8 us per launch instead of ~90 us.
struct WgradTcMaps {
  CUtensorMap u[8];
  CUtensorMap v[8];
};
static constexpr int kWtTok = 64;          // tokens per pipeline stage
static constexpr int kWtStages = 4;
static constexpr int kWtThreads = 192;     // warps 0-3 epilogue (TMEM lane quarters), warp 4 TMA, warp 5 MMA
__global__ void __launch_bounds__(kWtThreads)
lora_wgrad_tc_kernel(const __grid_constant__ WgradTcMaps maps, const WgradParams p, int tiles_per_cta) {
  extern __shared__ __align__(1024) uint8_t wt_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(wt_smem_raw) + 127) & ~uintptr_t(127));
  const WgradDesc& d = p.d[blockIdx.z];
  const int N = (d.r + 15) & ~15;                      // MMA N: 16 or 32
  constexpr int kABytes = kWtTok * 128 * 2;            // [16 chunks][64 tokens][16 B]
  const int b_bytes = kWtTok * N * 2;                  // [N/8 chunks][64 tokens][16 B]
  const int stage_bytes = kABytes + 4096;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kWtStages * stage_bytes);
  uint64_t* empty_bar = full_bar + kWtStages;
  uint64_t* acc_bar = empty_bar + kWtStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * 128;
  const int ntiles_total = (p.M + kWtTok - 1) / kWtTok;
  const int t_begin = blockIdx.y * tiles_per_cta;
  const int t_end = min(ntiles_total, t_begin + tiles_per_cta);
  if (c0 >= d.C || t_begin >= t_end) return;           // (uniform per CTA: nothing allocated yet)
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&maps.u[blockIdx.z]);
    tma_prefetch_desc(&maps.v[blockIdx.z]);
    for (int s = 0; s < kWtStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_slot, 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();
  // producer and MMA issuer: the whole warp runs the loop (uniform control flow), one lane elected at each use issues
  if (warp == 4) {
    {
      int s = 0;
      uint32_t ph = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* dst = smem + s * stage_bytes;
        if (elect_one()) {
          mbar_expect_tx(&full_bar[s], kABytes + b_bytes);
          tma_load_4d(dst, &maps.u[blockIdx.z], &full_bar[s], 0, t * kWtTok, c0 / 8, 0);
          tma_load_4d(dst + kABytes, &maps.v[blockIdx.z], &full_bar[s], 0, t * kWtTok, 0, 0);
        }
        if (++s == kWtStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 5) {
    {
      const uint32_t idesc = make_idesc_bf16(128, N, 1, 1);      // A and B MN-major
      int s = 0;
      uint32_t ph = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
        const uint32_t b_addr = a_addr + kABytes;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kWtTok / 16; ++k) {
            // MN-major: next 8-token K group +128 B (LBO), next 8-channel chunk + 64 tokens x 16 B = 1024 B (SBO)
            const uint64_t a_desc = make_smem_desc(a_addr + k * 256, 128, kWtTok * 16, SWZ_NONE);
            const uint64_t b_desc = make_smem_desc(b_addr + k * 256, 128, kWtTok * 16, SWZ_NONE);
            umma_bf16_ss(tmem_base, a_desc, b_desc, idesc, (t != t_begin) || k != 0);
          }
          umma_commit(&empty_bar[s]);
        }
        if (++s == kWtStages) { s = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(acc_bar);
    }
  } else {
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const int c = c0 + warp * 32 + lane;                 // accumulator row == channel
    uint32_t r[32];
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    if (N == 16) {
      uint32_t r16[16];
      tmem_ld_x16(taddr, r16);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = r16[j];
    } else {
      tmem_ld_x32(taddr, r);
      tmem_wait_ld();
    }
    if (c < d.C) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < d.r) atomicAdd(d.out + static_cast<size_t>(c) * d.ldc + static_cast<size_t>(j) * d.ldj, __uint_as_float(r[j]) * d.scale);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

// ---------------------------------------------------------------------------------- small data-movement kernels
// z[n, 2i, 2j, :] = dy[n, i, j, :], zero elsewhere: the stride-2 Downsample2D dgrad becomes a stride-1 dgrad over z.
__global__ void __launch_bounds__(256)
zero_insert_kernel(const __nv_bfloat16* __restrict__ dy, int nb, int H, int W, int Ho, int Wo, int C,
                   __nv_bfloat16* __restrict__ z) {
  pdl_launch_dependents();
  pdl_wait();
  const int vpp = C >> 3;
  const size_t total = static_cast<size_t>(nb) * H * W * vpp;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vpp);
    size_t pix = i / vpp;
    const int w = static_cast<int>(pix % W); pix /= W;
    const int h = static_cast<int>(pix % H);
    const int n = static_cast<int>(pix / H);
    uint4 o = make_uint4(0, 0, 0, 0);
    if (((h | w) & 1) == 0)
      o = *reinterpret_cast<const uint4*>(dy + ((static_cast<size_t>(n) * Ho + (h >> 1)) * Wo + (w >> 1)) * C + v * 8);
    *reinterpret_cast<uint4*>(z + i * 8) = o;
  }
}

// dx[n, h, w, :] = sum over the output pixels (ho, wo) with floor(ho*H/Ho) == h, floor(wo*W/Wo) == w of dy[n, ho, wo, :]
__global__ void __launch_bounds__(256)
upsample_nearest_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int nb, int H, int W, int Ho, int Wo, int C,
                            __nv_bfloat16* __restrict__ dx) {
  pdl_launch_dependents();
  pdl_wait();
  const int vpp = C >> 3;
  const size_t total = static_cast<size_t>(nb) * H * W * vpp;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vpp);
    size_t pix = i / vpp;
    const int w = static_cast<int>(pix % W); pix /= W;
    const int h = static_cast<int>(pix % H);
    const int n = static_cast<int>(pix / H);
    // smallest ho with floor(ho*H/Ho) >= h is ceil(h*Ho/H)
    const int ho0 = (h * Ho + H - 1) / H, ho1 = ((h + 1) * Ho + H - 1) / H;
    const int wo0 = (w * Wo + W - 1) / W, wo1 = ((w + 1) * Wo + W - 1) / W;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int ho = ho0; ho < ho1 && ho < Ho; ++ho)
      for (int wo = wo0; wo < wo1 && wo < Wo; ++wo) {
        float f[8];
        unpack8t(*reinterpret_cast<const uint4*>(dy + ((static_cast<size_t>(n) * Ho + ho) * Wo + wo) * C + v * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    *reinterpret_cast<uint4*>(dx + i * 8) = pack8t(acc);
  }
}

__global__ void __launch_bounds__(256)
add_bf16_kernel(__nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ x, size_t nvec) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float a[8], b[8];
    unpack8t(*reinterpret_cast<const uint4*>(y + i * 8), a);
    unpack8t(*reinterpret_cast<const uint4*>(x + i * 8), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += b[j];
    *reinterpret_cast<uint4*>(y + i * 8) = pack8t(a);
  }
}

// loss_sum += sum (pred - noise)^2 ; d_eps[pix, 0:8] = 2 (pred - noise) * inv_count  (bf16, NHWC, c_pad columns per pixel,
// columns >= 8 are written as zero).  pred: fp32 NHWC [nb, hw, 8]; noise: fp32 NCHW [nb, 8, hw].
__global__ void __launch_bounds__(256)
mse_grad_kernel(const float* __restrict__ pred, const float* __restrict__ noise, int nb, int hw, int c_pad, float inv_count,
                float* __restrict__ loss_sum, __nv_bfloat16* __restrict__ deps) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t npix = static_cast<size_t>(nb) * hw;
  float local = 0.f;
  for (size_t pix = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; pix < npix;
       pix += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t n = pix / hw, s = pix % hw;
    const float4 p0 = *reinterpret_cast<const float4*>(pred + pix * 8);
    const float4 p1 = *reinterpret_cast<const float4*>(pred + pix * 8 + 4);
    const float pr[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    float g[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float d = pr[c] - noise[(n * 8 + c) * hw + s];
      local = fmaf(d, d, local);
      g[c] = 2.0f * d * inv_count;
    }
    __nv_bfloat16* o = deps + pix * c_pad;
    *reinterpret_cast<uint4*>(o) = pack8t(g);
    for (int c = 8; c < c_pad; c += 8) *reinterpret_cast<uint4*>(o + c) = make_uint4(0, 0, 0, 0);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  __shared__ float s_part[8];
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += s_part[i];
    atomicAdd(loss_sum, t);
  }
}

// After the optimizer step: re-materialise the bf16 packed LoRA operands (forward K segments, down-projection
// weights and their transposed backward counterparts) from the flat fp32 parameter arena.
struct RefreshDesc {
  __nv_bfloat16* dst;      // destination block origin
  long long src_off;       // element offset of the [src_rows, src_cols] row-major source inside the flat arena
  int dst_ld, src_rows, src_cols, transpose;
  float scale;
  int pad_;
};
__global__ void __launch_bounds__(256)
lora_refresh_kernel(const RefreshDesc* __restrict__ descs, const float* __restrict__ flat) {
  pdl_launch_dependents();
  pdl_wait();
  const RefreshDesc d = descs[blockIdx.x];
  const int total = d.src_rows * d.src_cols;
  const float* src = flat + d.src_off;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int r = i / d.src_cols, c = i % d.src_cols;
    const float val = src[i] * d.scale;
    const size_t o = d.transpose ? static_cast<size_t>(c) * d.dst_ld + r : static_cast<size_t>(r) * d.dst_ld + c;
    d.dst[o] = __float2bfloat16(val);
  }
}

static inline int grid1d(size_t n, int threads, int cap = 148 * 16) {
  size_t g = (n + threads - 1) / threads;
  if (g > static_cast<size_t>(cap)) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_groupnorm_silu_bwd(const void* x0, int c0, const void* x1, int c1, int nb, int hw, int groups,
                                       const float* gamma, const float* beta, const float* stats, int silu,
                                       const void* dy, const void* dres, int res_ld, void* dx0, void* dx1,
                                       void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const int C = c0 + c1;
  B200_CHECK_ARG(x0 && dy && dx0 && gamma && beta && stats, "groupnorm_bwd: null pointer");
  B200_CHECK_ARG(c0 % 8 == 0 && c1 % 8 == 0 && C <= 2560, "groupnorm_bwd: channels (%d,%d) unsupported", c0, c1);
  B200_CHECK_ARG((c1 == 0) == (x1 == nullptr), "groupnorm_bwd: second source mismatch");
  B200_CHECK_ARG(groups > 0 && groups <= 32 && C % groups == 0, "groupnorm_bwd: groups");
  B200_CHECK_ARG(nb > 0 && hw > 0, "groupnorm_bwd: empty input");
  B200_CHECK_ARG(!dres || res_ld % 8 == 0, "groupnorm_bwd: res_ld");
  GnBwdParams p;
  p.x0 = reinterpret_cast<const __nv_bfloat16*>(x0); p.x1 = reinterpret_cast<const __nv_bfloat16*>(x1);
  p.C0 = c0; p.C1 = c1; p.HW = hw; p.groups = groups; p.gamma = gamma; p.beta = beta; p.stats = stats; p.silu = silu;
  p.dy = reinterpret_cast<const __nv_bfloat16*>(dy); p.dres = reinterpret_cast<const __nv_bfloat16*>(dres); p.res_ld = res_ld;
  p.dx0 = reinterpret_cast<__nv_bfloat16*>(dx0); p.dx1 = reinterpret_cast<__nv_bfloat16*>(dx1);
  static int max_active[4] = {-1, -1, -1, -1};
  int best_cs = 1, best_gs = 1, best_ctas = 0;
  for (int ci = 3; ci >= 0; --ci) {
    const int cs = 1 << ci;
    if (cs > 1 && hw < 4 * cs) continue;
    if (max_active[ci] < 0) {
      cudaLaunchConfig_t qc;
      memset(&qc, 0, sizeof(qc));
      qc.gridDim = dim3(cs, 1, 1);
      qc.blockDim = dim3(kGnbThreads, 1, 1);
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = cs; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      qc.attrs = qa;
      qc.numAttrs = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, groupnorm_silu_bwd_kernel, &qc) != cudaSuccess || nclusters <= 0) {
        cudaGetLastError();
        nclusters = 0;
      }
      max_active[ci] = nclusters;
    }
    for (int gs = 1; gs <= 4; gs *= 2) {
      if (groups % gs != 0 || (C / gs) % 8 != 0) continue;
      // a cluster's channel range must not straddle the two sources unless it is vector-aligned (it is: c0 % 8 == 0)
      if (nb * gs > max_active[ci]) continue;
      const int ctas = nb * gs * cs;
      if (ctas > best_ctas) { best_ctas = ctas; best_cs = cs; best_gs = gs; }
    }
  }
  if (best_ctas == 0) { best_cs = 1; best_gs = 1; }
  B200_CHECK_PDL("groupnorm_bwd", launch_pdl(groupnorm_silu_bwd_kernel, dim3(best_cs, nb, best_gs), dim3(kGnbThreads), 0,
                                             stream, best_cs, p));
  return B200_OK;
}

extern "C" int b200_layernorm_bwd(const void* x, const void* dy, int m, int c, const float* gamma, float eps,
                                  const void* dres, void* dx, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(x && dy && dx && gamma, "layernorm_bwd: null pointer");
  B200_CHECK_ARG(c % 8 == 0 && c <= kLnbMaxVec * 256, "layernorm_bwd: C=%d unsupported", c);
  if (m == 0) return B200_OK;
  const int wpb = 8;
  B200_CHECK_PDL("layernorm_bwd", launch_pdl(layernorm_bwd_kernel, dim3((m + wpb - 1) / wpb), dim3(wpb * 32), 0, stream, 0,
                                             reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(dy),
                                             m, c, gamma, eps, reinterpret_cast<const __nv_bfloat16*>(dres),
                                             reinterpret_cast<__nv_bfloat16*>(dx)));
  return B200_OK;
}

extern "C" int b200_geglu_fwd(const void* h, long m, int f, void* out, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(h && out && m > 0 && f > 0 && f % 8 == 0, "geglu_fwd: bad args");
  B200_CHECK_PDL("geglu_fwd", launch_pdl(geglu_fwd_kernel, dim3(grid1d(static_cast<size_t>(m) * (f / 8), 256)), dim3(256), 0, stream, 0,
                                         reinterpret_cast<const __nv_bfloat16*>(h), static_cast<size_t>(m), f,
                                         reinterpret_cast<__nv_bfloat16*>(out)));
  return B200_OK;
}
extern "C" int b200_geglu_bwd(const void* h, const void* dout, long m, int f, void* dh, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(h && dout && dh && m > 0 && f > 0 && f % 8 == 0, "geglu_bwd: bad args");
  B200_CHECK_PDL("geglu_bwd", launch_pdl(geglu_bwd_kernel, dim3(grid1d(static_cast<size_t>(m) * (f / 8), 256)), dim3(256), 0, stream, 0,
                                         reinterpret_cast<const __nv_bfloat16*>(h), reinterpret_cast<const __nv_bfloat16*>(dout),
                                         static_cast<size_t>(m), f, reinterpret_cast<__nv_bfloat16*>(dh)));
  return B200_OK;
}

// descs: HOST array of n (<= 8) records {u, v, out, ldu, ldv, C, r, ldc, ldj, scale} laid out as struct B200WgradDesc.
extern "C" int b200_lora_wgrad(const void* descs_host, int n, int m, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(descs_host && n > 0 && n <= 8 && m > 0, "lora_wgrad: bad args");
  WgradParams p;
  memset(&p, 0, sizeof(p));
  const WgradDesc* d = reinterpret_cast<const WgradDesc*>(descs_host);
  int cmax = 0;
  for (int i = 0; i < n; ++i) {
    p.d[i] = d[i];
    B200_CHECK_ARG(d[i].u && d[i].v && d[i].out, "lora_wgrad: null pointer in descriptor %d", i);
    B200_CHECK_ARG(d[i].r > 0 && d[i].r <= 32, "lora_wgrad: rank %d unsupported (1..32)", d[i].r);
    B200_CHECK_ARG(d[i].C % 8 == 0 && d[i].ldu % 8 == 0, "lora_wgrad: C/ldu must be multiples of 8");
    B200_CHECK_ARG((reinterpret_cast<uintptr_t>(d[i].u) & 15) == 0, "lora_wgrad: U not 16-byte aligned");
    if (d[i].C > cmax) cmax = d[i].C;
  }
  p.n = n;
  p.M = m;
  // tensor-core path: needs 16-byte aligned V slices (rank a multiple of 8), channels a multiple of 8
  static const bool no_tc = getenv("B200_WGRAD_TC") && atoi(getenv("B200_WGRAD_TC")) == 0;
  bool tc_ok = !no_tc;
  for (int i = 0; i < n && tc_ok; ++i)
    tc_ok = (d[i].r % 8 == 0) && (reinterpret_cast<uintptr_t>(d[i].v) & 15) == 0 && d[i].ldv % 8 == 0 && d[i].C % 8 == 0;
  if (tc_ok) {
    WgradTcMaps maps;
    memset(&maps, 0, sizeof(maps));
    for (int i = 0; i < 8; ++i) {
      const WgradDesc& di = d[i < n ? i : 0];
      const int N = (di.r + 15) & ~15;
      {
        uint64_t dims[4] = {8, (uint64_t)m, (uint64_t)(di.C / 8), 1};
        uint64_t strides[3] = {(uint64_t)di.ldu, 8, (uint64_t)m * di.ldu};
        uint32_t box[4] = {8, (uint32_t)kWtTok, 16, 1};
        int rc = make_tmap_bf16(&maps.u[i], di.u, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
      }
      {
        uint64_t dims[4] = {8, (uint64_t)m, (uint64_t)(N / 8), 1};
        uint64_t strides[3] = {(uint64_t)di.ldv, 8, (uint64_t)m * di.ldv};
        uint32_t box[4] = {8, (uint32_t)kWtTok, (uint32_t)(N / 8), 1};
        int rc = make_tmap_bf16(&maps.v[i], di.v, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
      }
    }
    const int cblk = (cmax + 127) / 128;
    const int ntiles = (m + kWtTok - 1) / kWtTok;
    int splits = (148 * 2 + cblk * n - 1) / (cblk * n);
    if (splits > ntiles) splits = ntiles;
    if (splits < 1) splits = 1;
    const int tiles_per_cta = (ntiles + splits - 1) / splits;
    splits = (ntiles + tiles_per_cta - 1) / tiles_per_cta;
    const size_t smem = static_cast<size_t>(kWtStages) * (kWtTok * 128 * 2 + 4096) + 256 + 128;
    static bool configured = false;
    if (!configured) {
      cudaFuncSetAttribute(lora_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      configured = true;
    }
    B200_CHECK_PDL("lora_wgrad(tc)", launch_pdl(lora_wgrad_tc_kernel, dim3(cblk, splits, n), dim3(kWtThreads), smem, stream, 0,
                                                maps, p, tiles_per_cta));
    return B200_OK;
  }
  // ~8 resident CTAs per SM (128 threads, 24 KB smem each) over the token dimension
  const int cblk = (cmax + 63) / 64;
  int chunks = (148 * 8 + cblk * n - 1) / (cblk * n);
  int rows = (m + chunks - 1) / chunks;
  rows = (rows + kWgRows - 1) / kWgRows * kWgRows;
  if (rows < kWgRows) rows = kWgRows;
  p.rows_per_cta = rows;
  chunks = (m + rows - 1) / rows;
  B200_CHECK_PDL("lora_wgrad", launch_pdl(lora_wgrad_kernel, dim3(cblk, chunks, n), dim3(kWgThreads), 0, stream, 0, p));
  return B200_OK;
}

extern "C" int b200_zero_insert(const void* dy, int nb, int h, int w, int c, void* z, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(dy && z && nb > 0 && h > 0 && w > 0 && c % 8 == 0, "zero_insert: bad args");
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  const size_t total = static_cast<size_t>(nb) * h * w * (c / 8);
  B200_CHECK_PDL("zero_insert", launch_pdl(zero_insert_kernel, dim3(grid1d(total, 256)), dim3(256), 0, stream, 0,
                                           reinterpret_cast<const __nv_bfloat16*>(dy), nb, h, w, ho, wo, c,
                                           reinterpret_cast<__nv_bfloat16*>(z)));
  return B200_OK;
}

extern "C" int b200_upsample_nearest_bwd(const void* dy, int nb, int h, int w, int c, int ho, int wo, void* dx,
                                         void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(dy && dx && nb > 0 && h > 0 && w > 0 && ho >= h && wo >= w && c % 8 == 0, "upsample_nearest_bwd: bad args");
  const size_t total = static_cast<size_t>(nb) * h * w * (c / 8);
  B200_CHECK_PDL("upsample_nearest_bwd", launch_pdl(upsample_nearest_bwd_kernel, dim3(grid1d(total, 256)), dim3(256), 0, stream, 0,
                                                    reinterpret_cast<const __nv_bfloat16*>(dy), nb, h, w, ho, wo, c,
                                                    reinterpret_cast<__nv_bfloat16*>(dx)));
  return B200_OK;
}

extern "C" int b200_add_bf16(void* y, const void* x, long n, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(y && x && n > 0 && n % 8 == 0, "add_bf16: bad args");
  B200_CHECK_PDL("add_bf16", launch_pdl(add_bf16_kernel, dim3(grid1d(static_cast<size_t>(n) / 8, 256)), dim3(256), 0, stream, 0,
                                        reinterpret_cast<__nv_bfloat16*>(y), reinterpret_cast<const __nv_bfloat16*>(x),
                                        static_cast<size_t>(n) / 8));
  return B200_OK;
}

extern "C" int b200_mse_grad(const float* pred_nhwc, const float* noise_nchw, int nb, int hw, int c_pad, float inv_count,
                             float* loss_sum, void* deps, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(pred_nhwc && noise_nchw && loss_sum && deps && nb > 0 && hw > 0 && c_pad >= 8 && c_pad % 8 == 0,
                 "mse_grad: bad args");
  B200_CHECK_PDL("mse_grad", launch_pdl(mse_grad_kernel, dim3(grid1d(static_cast<size_t>(nb) * hw, 256)), dim3(256), 0, stream, 0,
                                        pred_nhwc, noise_nchw, nb, hw, c_pad, inv_count, loss_sum,
                                        reinterpret_cast<__nv_bfloat16*>(deps)));
  return B200_OK;
}

// descs: DEVICE array of n records (struct B200RefreshDesc), flat: DEVICE fp32 parameter arena.
extern "C" int b200_lora_refresh(const void* descs_dev, int n, const float* flat, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(descs_dev && flat && n > 0, "lora_refresh: bad args");
  B200_CHECK_PDL("lora_refresh", launch_pdl(lora_refresh_kernel, dim3(n), dim3(256), 0, stream, 0,
                                            reinterpret_cast<const RefreshDesc*>(descs_dev), flat));
  return B200_OK;
}
