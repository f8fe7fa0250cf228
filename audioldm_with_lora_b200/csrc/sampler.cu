// Sampler step, embeddings, layout and fine-tuning elementwise kernels (HBM / latency bound).
//
//  * b200_sampler_step: classifier-free-guidance combine + DDIM / PNDM(PLMS) latent update + the
//    CFG-duplicated bf16 UNet input for the next step, one launch (K13).  Replaces ~12 ATen launches
//    per step of AudioLDMPipeline.__call__ (/root/reference/app.py:14,
//    /root/reference/script/inference/generate_audio.py:47-52) and DDIMScheduler.step (scheduler class
//    pinned at /root/reference/script/train/train_audioldm_lora.py:367).
//  * b200_time_class_embed: Timesteps -> TimestepEmbedding -> cat with class_embedding(labels) (K12).
//  * pack / unpack / nearest-upsample layout kernels (K10 and the NCHW<->NHWC boundary).
//  * b200_add_noise, b200_mse_partial, b200_adamw_flat: fine-tuning step tail (K15),
//    train_audioldm_lora.py:504,549,563.
#include "host_util.h"
#include "ptx.cuh"

namespace b200 {

// ---------------------------------------------------------------------------------- sampler step
// table row (8 floats) for step s:
//   [0] a      [1] b      [2..5] w0..w3 (weights of e_now, hist[h1], hist[h2], hist[h3])
//   [6] flags  (bit0: x_base = x_saved instead of x ; bit1: save x into x_saved before the update ;
//               bits 4-6: push slot + 1 (0 = do not store e_now in the history ring))
//   [7] packed history slots  h1 | h2 << 4 | h3 << 8
// DDIM (eta = 0): a = sqrt(a_prev / a_t), b = sqrt(1 - a_prev) - sqrt(a_prev (1 - a_t) / a_t), w = (1,0,0,0).
__global__ void __launch_bounds__(256)
sampler_step_kernel(const float* __restrict__ eps, float* __restrict__ x, float* __restrict__ x_saved,
                    float* __restrict__ hist, const float* __restrict__ table, int* __restrict__ step_ptr,
                    float guidance, int do_cfg, int nb, int hw, int c, int c_pad,
                    __nv_bfloat16* __restrict__ xin_next, unsigned int* __restrict__ done_ctr) {
  pdl_launch_dependents();
  pdl_wait();
  const int step = *step_ptr;
  const float* row = table + static_cast<size_t>(step) * 8;
  const float a = row[0], b = row[1];
  const float w0 = row[2], w1 = row[3], w2 = row[4], w3 = row[5];
  const int flags = __float_as_int(row[6]);
  const int slots = __float_as_int(row[7]);
  const int push = ((flags >> 4) & 7) - 1;
  const size_t n = static_cast<size_t>(nb) * hw * c;          // elements of one latent batch
  const size_t npix = static_cast<size_t>(nb) * hw;
  // one thread per pixel (c == 8 channels = two float4)
  for (size_t pix = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; pix < npix;
       pix += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float out8[8];
    for (int v = 0; v < c; v += 4) {
      const size_t i = pix * c + v;
      float4 e;
      if (do_cfg) {
        const float4 eu = *reinterpret_cast<const float4*>(eps + i);
        const float4 et = *reinterpret_cast<const float4*>(eps + n + i);
        e = make_float4(eu.x + guidance * (et.x - eu.x), eu.y + guidance * (et.y - eu.y),
                        eu.z + guidance * (et.z - eu.z), eu.w + guidance * (et.w - eu.w));
      } else {
        e = *reinterpret_cast<const float4*>(eps + i);
      }
      float4 eh = make_float4(w0 * e.x, w0 * e.y, w0 * e.z, w0 * e.w);
      if (w1 != 0.f) {
        const float4 hh = *reinterpret_cast<const float4*>(hist + static_cast<size_t>(slots & 15) * n + i);
        eh.x += w1 * hh.x; eh.y += w1 * hh.y; eh.z += w1 * hh.z; eh.w += w1 * hh.w;
      }
      if (w2 != 0.f) {
        const float4 hh = *reinterpret_cast<const float4*>(hist + static_cast<size_t>((slots >> 4) & 15) * n + i);
        eh.x += w2 * hh.x; eh.y += w2 * hh.y; eh.z += w2 * hh.z; eh.w += w2 * hh.w;
      }
      if (w3 != 0.f) {
        const float4 hh = *reinterpret_cast<const float4*>(hist + static_cast<size_t>((slots >> 8) & 15) * n + i);
        eh.x += w3 * hh.x; eh.y += w3 * hh.y; eh.z += w3 * hh.z; eh.w += w3 * hh.w;
      }
      if (push >= 0) *reinterpret_cast<float4*>(hist + static_cast<size_t>(push) * n + i) = e;
      const float4 xc = *reinterpret_cast<const float4*>(x + i);
      if (flags & 2) *reinterpret_cast<float4*>(x_saved + i) = xc;
      const float4 xb = (flags & 1) ? *reinterpret_cast<const float4*>(x_saved + i) : xc;
      const float4 xn = make_float4(a * xb.x + b * eh.x, a * xb.y + b * eh.y, a * xb.z + b * eh.z, a * xb.w + b * eh.w);
      *reinterpret_cast<float4*>(x + i) = xn;
      out8[v] = xn.x; out8[v + 1] = xn.y; out8[v + 2] = xn.z; out8[v + 3] = xn.w;
    }
    if (xin_next) {
      for (int v = 0; v < c; v += 8) {
        uint4 pk;
        pk.x = pack_bf16x2(out8[v], out8[v + 1]); pk.y = pack_bf16x2(out8[v + 2], out8[v + 3]);
        pk.z = pack_bf16x2(out8[v + 4], out8[v + 5]); pk.w = pack_bf16x2(out8[v + 6], out8[v + 7]);
        *reinterpret_cast<uint4*>(xin_next + pix * c_pad + v) = pk;
        if (do_cfg) *reinterpret_cast<uint4*>(xin_next + (npix + pix) * c_pad + v) = pk;
      }
    }
  }
  // last block to finish advances the step counter (every block has read it by then)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(done_ctr, 1u);
    if (prev == gridDim.x - 1) {
      *done_ctr = 0;
      *step_ptr = step + 1;
    }
  }
}

// ---------------------------------------------------------------------------------- embeddings
// grid (nchunk, nb); block 256.  Weights arrive TRANSPOSED ([in, out] row-major) so that consecutive
// threads read consecutive floats of one k-row: every load of the k loops is coalesced and independent
// (no shuffles in the loop), i.e. the loops pipeline.  Every block recomputes the 1st MLP layer
// (tproj x ted MACs, 256 KB of L2-resident weights); each block then produces kEmbCols of the 2*ted
// outputs with a 4-way split of k across the thread block.
static constexpr int kEmbThreads = 256;
static constexpr int kEmbCols = 64;
__global__ void __launch_bounds__(kEmbThreads)
time_class_embed_kernel(const float* __restrict__ t_steps, const int* __restrict__ step_ptr, int per_sample,
                        const float* __restrict__ labels, int tproj, int ted, int class_in,
                        const float* __restrict__ w1t, const float* __restrict__ b1, const float* __restrict__ w2t,
                        const float* __restrict__ b2, const float* __restrict__ wct, const float* __restrict__ bc,
                        float* __restrict__ emb, __nv_bfloat16* __restrict__ silu_emb) {
  extern __shared__ float sm[];
  float* s_sin = sm;                  // [tproj]  ([cos | sin], flip_sin_to_cos=True)
  float* s_h = sm + tproj;            // [ted]
  float* s_lab = s_h + ted;           // [class_in]
  float* s_red = s_lab + class_in;    // [4][kEmbCols]
  pdl_launch_dependents();
  pdl_wait();
  const int bidx = blockIdx.y;
  const int total = 2 * ted;
  const int o_begin = blockIdx.x * kEmbCols;
  if (o_begin >= total) return;
  const bool time_half = o_begin < ted;         // ted % kEmbCols == 0: a block never straddles the halves
  const int col = threadIdx.x % kEmbCols;
  const int ks = threadIdx.x / kEmbCols;        // 0..3
  float acc = 0.f;
  if (time_half) {
    const float t = per_sample ? t_steps[bidx] : t_steps[step_ptr ? *step_ptr : 0];
    const int half = tproj / 2;
    for (int i = threadIdx.x; i < half; i += kEmbThreads) {
      const float freq = expf(-logf(10000.0f) * static_cast<float>(i) / static_cast<float>(half));
      const float arg = t * freq;
      s_sin[i] = cosf(arg);
      s_sin[half + i] = sinf(arg);
    }
    __syncthreads();
    for (int o = threadIdx.x; o < ted; o += kEmbThreads) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int k = 0;
#pragma unroll 8
      for (; k + 3 < tproj; k += 4) {
        a0 = fmaf(w1t[static_cast<size_t>(k) * ted + o], s_sin[k], a0);
        a1 = fmaf(w1t[static_cast<size_t>(k + 1) * ted + o], s_sin[k + 1], a1);
        a2 = fmaf(w1t[static_cast<size_t>(k + 2) * ted + o], s_sin[k + 2], a2);
        a3 = fmaf(w1t[static_cast<size_t>(k + 3) * ted + o], s_sin[k + 3], a3);
      }
      for (; k < tproj; ++k) a0 = fmaf(w1t[static_cast<size_t>(k) * ted + o], s_sin[k], a0);
      const float v = (a0 + a1) + (a2 + a3) + b1[o];
      s_h[o] = v / (1.0f + expf(-v));
    }
    __syncthreads();
    const int o = o_begin + col;
    const int kper = (ted + 3) / 4;
    const int k0 = ks * kper, k1 = min(ted, k0 + kper);
    float a0 = 0.f, a1 = 0.f;
    int k = k0;
#pragma unroll 16
    for (; k + 1 < k1; k += 2) {
      a0 = fmaf(w2t[static_cast<size_t>(k) * ted + o], s_h[k], a0);
      a1 = fmaf(w2t[static_cast<size_t>(k + 1) * ted + o], s_h[k + 1], a1);
    }
    if (k < k1) a0 = fmaf(w2t[static_cast<size_t>(k) * ted + o], s_h[k], a0);
    acc = a0 + a1;
  } else {
    for (int i = threadIdx.x; i < class_in; i += kEmbThreads) s_lab[i] = labels[static_cast<size_t>(bidx) * class_in + i];
    __syncthreads();
    const int o = o_begin - ted + col;
    const int kper = (class_in + 3) / 4;
    const int k0 = ks * kper, k1 = min(class_in, k0 + kper);
    float a0 = 0.f, a1 = 0.f;
    int k = k0;
#pragma unroll 16
    for (; k + 1 < k1; k += 2) {
      a0 = fmaf(wct[static_cast<size_t>(k) * ted + o], s_lab[k], a0);
      a1 = fmaf(wct[static_cast<size_t>(k + 1) * ted + o], s_lab[k + 1], a1);
    }
    if (k < k1) a0 = fmaf(wct[static_cast<size_t>(k) * ted + o], s_lab[k], a0);
    acc = a0 + a1;
  }
  s_red[ks * kEmbCols + col] = acc;
  __syncthreads();
  if (threadIdx.x < kEmbCols) {
    const int o = o_begin + col;
    const float bias = time_half ? b2[o] : bc[o - ted];
    const float v = ((s_red[col] + s_red[kEmbCols + col]) + (s_red[2 * kEmbCols + col] + s_red[3 * kEmbCols + col])) + bias;
    if (emb) emb[static_cast<size_t>(bidx) * total + o] = v;
    silu_emb[static_cast<size_t>(bidx) * total + o] = __float2bfloat16(v / (1.0f + expf(-v)));
  }
}

// ---------------------------------------------------------------------------------- layout kernels
__global__ void pack_nchw_to_nhwc_kernel(const float* __restrict__ x, int nb, int c, int hw, int c_pad,
                                         __nv_bfloat16* __restrict__ y) {
  const size_t total = static_cast<size_t>(nb) * hw;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t n = i / hw, p = i % hw;
    for (int ch = 0; ch < c; ++ch) y[i * c_pad + ch] = __float2bfloat16(x[(n * c + ch) * hw + p]);
  }
}
__global__ void unpack_nhwc_to_nchw_kernel(const float* __restrict__ x, int nb, int c, int hw, float* __restrict__ y) {
  const size_t total = static_cast<size_t>(nb) * hw;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t n = i / hw, p = i % hw;
    for (int ch = 0; ch < c; ++ch) y[(n * c + ch) * hw + p] = x[i * c + ch];
  }
}
// 16-byte vectors; src index = floor(dst * (in / out)) evaluated in fp32 like ATen's nearest kernel.
__global__ void upsample_nearest_kernel(const uint4* __restrict__ x, int nb, int h, int w, int cv, int ho, int wo,
                                        uint4* __restrict__ y) {
  pdl_launch_dependents();
  pdl_wait();
  const float sh = static_cast<float>(h) / static_cast<float>(ho);
  const float sw = static_cast<float>(w) / static_cast<float>(wo);
  const size_t total = static_cast<size_t>(nb) * ho * wo * cv;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % cv);
    size_t r = i / cv;
    const int xo = static_cast<int>(r % wo);
    r /= wo;
    const int yo = static_cast<int>(r % ho);
    const int n = static_cast<int>(r / ho);
    const int ys = min(static_cast<int>(floorf(yo * sh)), h - 1);
    const int xs = min(static_cast<int>(floorf(xo * sw)), w - 1);
    y[i] = x[((static_cast<size_t>(n) * h + ys) * w + xs) * cv + v];
  }
}

// HiFi-GAN: mean of the three residual blocks of an upsampling stage + the LeakyReLU in front of the next layer.  The
// block outputs are stored post-LeakyReLU(in_slope) (b200_conv1d's convention): x = min(a, a / in_slope).
__global__ void lrelu_mean3_kernel(const uint4* __restrict__ a0, const uint4* __restrict__ a1, const uint4* __restrict__ a2,
                                   size_t nvec, float in_neg_gain, float out_slope, uint4* __restrict__ y) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint4 u0 = a0[i], u1 = a1[i], u2 = a2[i];
    const uint32_t w0[4] = {u0.x, u0.y, u0.z, u0.w}, w1[4] = {u1.x, u1.y, u1.z, u1.w}, w2[4] = {u2.x, u2.y, u2.z, u2.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float r[2];
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        const float x0 = hlf ? bf16_hi(w0[j]) : bf16_lo(w0[j]);
        const float x1 = hlf ? bf16_hi(w1[j]) : bf16_lo(w1[j]);
        const float x2 = hlf ? bf16_hi(w2[j]) : bf16_lo(w2[j]);
        const float m = (fminf(x0, x0 * in_neg_gain) + fminf(x1, x1 * in_neg_gain) + fminf(x2, x2 * in_neg_gain)) * (1.0f / 3.0f);
        r[hlf] = fmaxf(m, m * out_slope);
      }
      o[j] = pack_bf16x2(r[0], r[1]);
    }
    y[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void f32_to_bf16_kernel(const float4* __restrict__ x, size_t nvec, uint2* __restrict__ y) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 v = x[i];
    y[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

// ---------------------------------------------------------------------------------- fine-tuning tail
__global__ void add_noise_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                 const float* __restrict__ sa, const float* __restrict__ sb, int nb, size_t per,
                                 float* __restrict__ out) {
  const size_t total = static_cast<size_t>(nb) * per;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t b = i / per;
    out[i] = sa[b] * x0[i] + sb[b] * noise[i];
  }
}
__global__ void __launch_bounds__(256)
mse_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n, float* __restrict__ out) {
  float s = 0.f;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float d = a[i] - b[i];
    s += d * d;
  }
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    s = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
    if (threadIdx.x == 0) atomicAdd(out, s);
  }
}
// torch.optim.AdamW semantics (decoupled weight decay, bias correction), grad pre-scaled by grad_scale
// (1/world after the NCCL sum all-reduce).
__global__ void adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                  float* __restrict__ v, size_t n, float lr, float b1, float b2, float eps, float wd,
                                  float bc1, float bc2, float gs) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float gr = g[i] * gs;
    float pi = p[i] * (1.0f - lr * wd);
    const float mi = b1 * m[i] + (1.0f - b1) * gr;
    const float vi = b2 * v[i] + (1.0f - b2) * gr * gr;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
  }
}

// Same update with the per-step scalars {lr, bias-correction 1, bias-correction 2, grad_scale} read from device memory:
// the launch is identical every step, so the whole training step can be one replayed CUDA graph.
__global__ void adamw_flat_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                      float* __restrict__ v, size_t n, const float* __restrict__ hyper, float b1, float b2,
                                      float eps, float wd) {
  const float lr = hyper[0], bc1 = hyper[1], bc2 = hyper[2], gs = hyper[3];
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float gr = g[i] * gs;
    float pi = p[i] * (1.0f - lr * wd);
    const float mi = b1 * m[i] + (1.0f - b1) * gr;
    const float vi = b2 * v[i] + (1.0f - b2) * gr * gr;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
  }
}

static inline int grid_for(size_t n, int block) {
  size_t g = (n + block - 1) / block;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

__device__ unsigned int g_sampler_done_ctr = 0;

}  // namespace b200

using namespace b200;

extern "C" int b200_version(void) { return 100; }
extern "C" const char* b200_last_error(void) { return last_error_buf(); }

extern "C" int b200_sampler_step(const float* eps, float* x, float* x_saved, float* hist, const float* table,
                                 int* step_ptr, float guidance, int do_cfg, int nb, int hw, int c, int c_pad,
                                 void* xin_next, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(eps && x && table && step_ptr, "sampler_step: null pointer");
  B200_CHECK_ARG(c == 8 && c_pad % 8 == 0 && c_pad >= c, "sampler_step: c=%d c_pad=%d unsupported", c, c_pad);
  B200_CHECK_ARG(nb > 0 && hw > 0, "sampler_step: empty latents");
  unsigned int* ctr = nullptr;
  cudaGetSymbolAddress(reinterpret_cast<void**>(&ctr), g_sampler_done_ctr);
  const size_t npix = static_cast<size_t>(nb) * hw;
  B200_CHECK_PDL("sampler_step", launch_pdl(sampler_step_kernel, dim3(grid_for(npix, 256)), dim3(256), 0, stream, 0, eps, x,
                                            x_saved, hist, table, step_ptr, guidance, do_cfg, nb, hw, c, c_pad,
                                            reinterpret_cast<__nv_bfloat16*>(xin_next), ctr));
  return B200_OK;
}

extern "C" int b200_time_class_embed(const float* t_steps, const int* step_ptr, int per_sample, const float* labels,
                                     int nb, int tproj, int ted, int class_in, const float* w1t, const float* b1,
                                     const float* w2t, const float* b2, const float* wct, const float* bc, float* emb,
                                     void* silu_emb, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(t_steps && labels && w1t && b1 && w2t && b2 && wct && bc && silu_emb, "time_class_embed: null pointer");
  B200_CHECK_ARG(nb > 0 && tproj % 2 == 0 && ted > 0 && ted % kEmbCols == 0 && class_in > 0, "time_class_embed: bad dims");
  const size_t smem = sizeof(float) * (tproj + ted + class_in + 4 * kEmbCols);
  B200_CHECK_PDL("time_class_embed",
                 launch_pdl(time_class_embed_kernel, dim3(2 * ted / kEmbCols, nb), dim3(kEmbThreads), smem, stream, 0,
                            t_steps, step_ptr, per_sample, labels, tproj, ted, class_in, w1t, b1, w2t, b2, wct, bc, emb,
                            reinterpret_cast<__nv_bfloat16*>(silu_emb)));
  return B200_OK;
}

extern "C" int b200_pack_nchw_to_nhwc(const float* x, int nb, int c, int hw, int c_pad, void* y, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(x && y && nb > 0 && c > 0 && hw > 0 && c_pad >= c, "pack_nchw_to_nhwc: bad args");
  pack_nchw_to_nhwc_kernel<<<grid_for(static_cast<size_t>(nb) * hw, 256), 256, 0, stream>>>(
      x, nb, c, hw, c_pad, reinterpret_cast<__nv_bfloat16*>(y));
  B200_CHECK_LAUNCH("pack_nchw_to_nhwc");
  return B200_OK;
}
extern "C" int b200_unpack_nhwc_to_nchw(const float* x, int nb, int c, int hw, float* y, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(x && y && nb > 0 && c > 0 && hw > 0, "unpack_nhwc_to_nchw: bad args");
  unpack_nhwc_to_nchw_kernel<<<grid_for(static_cast<size_t>(nb) * hw, 256), 256, 0, stream>>>(x, nb, c, hw, y);
  B200_CHECK_LAUNCH("unpack_nhwc_to_nchw");
  return B200_OK;
}
extern "C" int b200_upsample_nearest(const void* x, int nb, int h, int w, int c, int ho, int wo, void* y,
                                     void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(x && y && nb > 0 && h > 0 && w > 0 && ho > 0 && wo > 0 && c % 8 == 0, "upsample_nearest: bad args");
  const size_t total = static_cast<size_t>(nb) * ho * wo * (c / 8);
  B200_CHECK_PDL("upsample_nearest", launch_pdl(upsample_nearest_kernel, dim3(grid_for(total, 256)), dim3(256), 0, stream, 0,
                                                reinterpret_cast<const uint4*>(x), nb, h, w, c / 8, ho, wo,
                                                reinterpret_cast<uint4*>(y)));
  return B200_OK;
}

extern "C" int b200_lrelu_mean3(const void* a0, const void* a1, const void* a2, long n, float in_slope, float out_slope,
                                void* y, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(a0 && a1 && a2 && y && n > 0 && n % 8 == 0 && in_slope > 0.f && in_slope <= 1.f && out_slope >= 0.f &&
                 out_slope <= 1.f, "lrelu_mean3: bad args");
  const size_t nvec = static_cast<size_t>(n) / 8;
  B200_CHECK_PDL("lrelu_mean3", launch_pdl(lrelu_mean3_kernel, dim3(grid_for(nvec, 256)), dim3(256), 0, stream, 0,
                                           reinterpret_cast<const uint4*>(a0), reinterpret_cast<const uint4*>(a1),
                                           reinterpret_cast<const uint4*>(a2), nvec, 1.0f / in_slope, out_slope,
                                           reinterpret_cast<uint4*>(y)));
  return B200_OK;
}

extern "C" int b200_f32_to_bf16(const float* x, long n, void* y, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(x && y && n > 0 && n % 4 == 0, "f32_to_bf16: bad args");
  const size_t nvec = static_cast<size_t>(n) / 4;
  B200_CHECK_PDL("f32_to_bf16", launch_pdl(f32_to_bf16_kernel, dim3(grid_for(nvec, 256)), dim3(256), 0, stream, 0,
                                           reinterpret_cast<const float4*>(x), nvec, reinterpret_cast<uint2*>(y)));
  return B200_OK;
}

extern "C" int b200_add_noise(const float* x0, const float* noise, const float* sqrt_ac, const float* sqrt_1mac, int nb,
                              int c, int hw, float* out_nchw, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(x0 && noise && sqrt_ac && sqrt_1mac && out_nchw && nb > 0, "add_noise: bad args");
  const size_t per = static_cast<size_t>(c) * hw;
  add_noise_kernel<<<grid_for(per * nb, 256), 256, 0, stream>>>(x0, noise, sqrt_ac, sqrt_1mac, nb, per, out_nchw);
  B200_CHECK_LAUNCH("add_noise");
  return B200_OK;
}
extern "C" int b200_mse_partial(const float* pred, const float* target, long n, float* out_sum, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(pred && target && out_sum && n > 0, "mse_partial: bad args");
  mse_partial_kernel<<<grid_for(static_cast<size_t>(n), 256), 256, 0, stream>>>(pred, target, static_cast<size_t>(n),
                                                                               out_sum);
  B200_CHECK_LAUNCH("mse_partial");
  return B200_OK;
}
extern "C" int b200_adamw_flat(float* param, const float* grad, float* m, float* v, long n, float lr, float beta1,
                               float beta2, float eps, float weight_decay, int step, float grad_scale,
                               void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(param && grad && m && v && n > 0 && step >= 1, "adamw_flat: bad args");
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.0f - powf(beta2, static_cast<float>(step));
  adamw_flat_kernel<<<grid_for(static_cast<size_t>(n), 256), 256, 0, stream>>>(
      param, grad, m, v, static_cast<size_t>(n), lr, beta1, beta2, eps, weight_decay, bc1, bc2, grad_scale);
  B200_CHECK_LAUNCH("adamw_flat");
  return B200_OK;
}
extern "C" int b200_adamw_flat_dev(float* param, const float* grad, float* m, float* v, long n, const float* hyper_dev,
                                   float beta1, float beta2, float eps, float weight_decay, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(param && grad && m && v && hyper_dev && n > 0, "adamw_flat_dev: bad args");
  adamw_flat_dev_kernel<<<grid_for(static_cast<size_t>(n), 256), 256, 0, stream>>>(
      param, grad, m, v, static_cast<size_t>(n), hyper_dev, beta1, beta2, eps, weight_decay);
  B200_CHECK_LAUNCH("adamw_flat_dev");
  return B200_OK;
}
