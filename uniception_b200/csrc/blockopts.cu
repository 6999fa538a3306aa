// Block option kernels: qk_norm (LayerNorm over head_dim on q and k, utils/transformer_blocks.py:199-200,
// :222, :306-307, :347) fused with the 2-D RoPE that follows it, and LayerScale (:389-412).  HBM-bound: 16-byte accesses,
// 8-lane groups per 64-wide head, statistics in fp32; per-column reductions are combined in the block before the global
// atomics.
#include <algorithm>

#include "common.cuh"

namespace uc {
namespace {

__device__ __forceinline__ void unpack8(const uint4& t, float (&v)[8]) {
  v[0] = bf16_lo(t.x); v[1] = bf16_hi(t.x); v[2] = bf16_lo(t.y); v[3] = bf16_hi(t.y);
  v[4] = bf16_lo(t.z); v[5] = bf16_hi(t.z); v[6] = bf16_lo(t.w); v[7] = bf16_hi(t.w);
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 t;
  t.x = pack_bf16(v[0], v[1]); t.y = pack_bf16(v[2], v[3]); t.z = pack_bf16(v[4], v[5]); t.w = pack_bf16(v[6], v[7]);
  return t;
}
__device__ __forceinline__ float group8_sum(float s) {
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  return s;
}

// ------------------------------------------------------------------------------------------------
// head norm (+ RoPE) forward: y[row, h*64 + c] = rope(LN_64(x[row, h*64 + :]) * gamma + beta)
// an 8-lane group owns one (row, head); lane g holds columns 8g..8g+7.  RoPE pairs (i, i+16) inside each 32-column
// half live two lanes apart, so the partner values arrive by one xor-2 shuffle per element.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) headnorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx,
                                                           __nv_bfloat16* __restrict__ y, int64_t ldy,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const int32_t* __restrict__ pos, const float* __restrict__ table,
                                                           int rows, int H, float eps) {
  const int g = threadIdx.x & 7;
  float ga[8], be[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ga[j] = gamma[g * 8 + j];
    be[j] = beta[g * 8 + j];
  }
  const int64_t total = (int64_t)rows * H;
  const int64_t step = ((int64_t)gridDim.x * blockDim.x) >> 3;
  // warp-uniform trip count (the shuffles need every lane); groups past the end are predicated off
  for (int64_t base = ((int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) >> 3; base < total; base += step) {
    const int64_t gi = base + ((threadIdx.x & 31) >> 3);
    const bool on = gi < total;
    const int64_t row = on ? gi / H : 0;
    const int h = on ? (int)(gi % H) : 0;
    float v[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x + row * ldx + h * 64 + g * 8)), v);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
    const float mean = group8_sum(s) * (1.f / 64.f);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] -= mean;
      q += v[j] * v[j];
    }
    const float rstd = rsqrtf(group8_sum(q) * (1.f / 64.f) + eps);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = v[j] * rstd * ga[j] + be[j];
    if (pos != nullptr) {
      const int X = g >> 2, l = g & 3;
      const int p = pos[row * 2 + X];
      const float4* cs = reinterpret_cast<const float4*>(table + ((int64_t)p * 16 + (l & 1) * 8) * 2);
      const float sgn = (l < 2) ? -1.f : 1.f;  // u' = u c - v s ; v' = v c + u s
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const float4 t = __ldg(cs + (j >> 1));  // (cos_j, sin_j, cos_j+1, sin_j+1)
        const float o0 = __shfl_xor_sync(0xffffffffu, v[j], 2), o1 = __shfl_xor_sync(0xffffffffu, v[j + 1], 2);
        v[j] = v[j] * t.x + sgn * o0 * t.y;
        v[j + 1] = v[j + 1] * t.z + sgn * o1 * t.w;
      }
    }
    if (on) *reinterpret_cast<uint4*>(y + row * ldy + h * 64 + g * 8) = pack8(v);
  }
}

// ------------------------------------------------------------------------------------------------
// head norm backward, in place on g (the gradient w.r.t. the normalised, un-rotated values: the attention backward
// already applied the inverse RoPE): g <- rstd * (g*gamma - mean(g*gamma) - xhat * mean(g*gamma*xhat)); statistics are
// recomputed from the saved raw projection x.  dgamma/dbeta [64] are ACCUMULATED.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) headnorm_bwd_kernel(__nv_bfloat16* __restrict__ gr, int64_t ldg,
                                                           const __nv_bfloat16* __restrict__ x, int64_t ldx,
                                                           const float* __restrict__ gamma, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, int rows, int H, float eps) {
  __shared__ float red[2][64];
  if (threadIdx.x < 128) red[threadIdx.x >> 6][threadIdx.x & 63] = 0.f;
  __syncthreads();
  const int g = threadIdx.x & 7;
  float ga[8], adg[8], adb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ga[j] = gamma[g * 8 + j];
    adg[j] = 0.f;
    adb[j] = 0.f;
  }
  const int64_t total = (int64_t)rows * H;
  const int64_t step = ((int64_t)gridDim.x * blockDim.x) >> 3;
  for (int64_t base = ((int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) >> 3; base < total; base += step) {
    const int64_t gi = base + ((threadIdx.x & 31) >> 3);
    const bool on = gi < total;
    const int64_t row = on ? gi / H : 0;
    const int h = on ? (int)(gi % H) : 0;
    float v[8], d[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x + row * ldx + h * 64 + g * 8)), v);
    unpack8(*reinterpret_cast<const uint4*>(gr + row * ldg + h * 64 + g * 8), d);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
    const float mean = group8_sum(s) * (1.f / 64.f);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] -= mean;
      q += v[j] * v[j];
    }
    const float rstd = rsqrtf(group8_sum(q) * (1.f / 64.f) + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] *= rstd;  // xhat
      if (on) {
        adg[j] += d[j] * v[j];
        adb[j] += d[j];
      }
      d[j] *= ga[j];
      s1 += d[j];
      s2 += d[j] * v[j];
    }
    const float m1 = group8_sum(s1) * (1.f / 64.f), m2 = group8_sum(s2) * (1.f / 64.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = rstd * (d[j] - m1 - v[j] * m2);
    if (on) *reinterpret_cast<uint4*>(gr + row * ldg + h * 64 + g * 8) = pack8(d);
  }
  // the four groups of a warp, then the eight warps of the block, then one global atomic per column per block
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    adg[j] += __shfl_xor_sync(0xffffffffu, adg[j], 8);
    adg[j] += __shfl_xor_sync(0xffffffffu, adg[j], 16);
    adb[j] += __shfl_xor_sync(0xffffffffu, adb[j], 8);
    adb[j] += __shfl_xor_sync(0xffffffffu, adb[j], 16);
  }
  if ((threadIdx.x & 31) < 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&red[0][g * 8 + j], adg[j]);
      atomicAdd(&red[1][g * 8 + j], adb[j]);
    }
  }
  __syncthreads();
  if (threadIdx.x < 64) atomicAdd(dgamma + threadIdx.x, red[0][threadIdx.x]);
  else if (threadIdx.x < 128) atomicAdd(dbeta + threadIdx.x - 64, red[1][threadIdx.x - 64]);
}

// ------------------------------------------------------------------------------------------------
// LayerScale: out = res + gamma[c] * z   /   dz = gamma[c] * dy, dgamma[c] += sum_rows dy * z
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layerscale_fwd_kernel(const __nv_bfloat16* __restrict__ z, const __nv_bfloat16* __restrict__ res,
                                                             const float* __restrict__ gamma, __nv_bfloat16* __restrict__ out,
                                                             int64_t n8, int c8) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8) * 8;
    float a[8], r[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(z) + i), a);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    if (res != nullptr) {
      unpack8(__ldg(reinterpret_cast<const uint4*>(res) + i), r);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = r[j] + gg[j] * a[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = gg[j] * a[j];
    }
    reinterpret_cast<uint4*>(out)[i] = pack8(a);
  }
}

// block = 8 warps; a warp covers 256 columns (8 per lane); blockIdx.y strides over row slabs (as colsum_kernel)
__global__ void __launch_bounds__(256) layerscale_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ z,
                                                             const float* __restrict__ gamma, __nv_bfloat16* __restrict__ dz,
                                                             float* __restrict__ dgamma, int rows, int cols) {
  __shared__ float red[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col < cols) {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    for (int row = blockIdx.y * 8 + warp; row < rows; row += gridDim.y * 8) {
      float a[8], b[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(dy + (int64_t)row * cols + col)), a);
      unpack8(__ldg(reinterpret_cast<const uint4*>(z + (int64_t)row * cols + col)), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j] += a[j] * b[j];
        a[j] *= gg[j];
      }
      *reinterpret_cast<uint4*>(dz + (int64_t)row * cols + col) = pack8(a);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
  __syncthreads();
  const int c = threadIdx.x;
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w][c];
  if (blockIdx.x * 256 + c < cols) atomicAdd(dgamma + blockIdx.x * 256 + c, s);
}

inline int blocks_for(int64_t threads_needed, int per_sm) {
  int64_t b = (threads_needed + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * per_sm;
  if (b > cap) b = cap;
  return b < 1 ? 1 : (int)b;
}

}  // namespace
}  // namespace uc

using namespace uc;

extern "C" int uc_headnorm_fwd(const uc_headnorm_params* p, uc_stream_t stream_) {
  UC_REQUIRE(p && p->x && p->y && p->gamma && p->beta, UC_ERR_BAD_SHAPE, "uc_headnorm_fwd: null pointer");
  UC_REQUIRE(p->rows > 0 && p->heads > 0 && p->ldx % 8 == 0 && p->ldy % 8 == 0 && p->ldx >= p->heads * 64 && p->ldy >= p->heads * 64,
             UC_ERR_BAD_SHAPE, "uc_headnorm_fwd: bad shape (rows %d heads %d ldx %lld ldy %lld; head_dim is 64)", p->rows, p->heads,
             (long long)p->ldx, (long long)p->ldy);
  UC_REQUIRE(((uintptr_t)p->x % 16 == 0) && ((uintptr_t)p->y % 16 == 0), UC_ERR_BAD_SHAPE, "uc_headnorm_fwd: x / y must be 16-byte aligned");
  UC_REQUIRE((p->positions == nullptr) == (p->rope_table == nullptr), UC_ERR_BAD_SHAPE, "uc_headnorm_fwd: positions and rope_table go together");
  headnorm_fwd_kernel<<<blocks_for((int64_t)p->rows * p->heads * 8, 8), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const __nv_bfloat16*>(p->x), p->ldx, static_cast<__nv_bfloat16*>(p->y), p->ldy, p->gamma, p->beta, p->positions,
      p->rope_table, p->rows, p->heads, p->eps);
  return check_launch("uc_headnorm_fwd");
}

extern "C" int uc_headnorm_bwd(const uc_headnorm_params* p, uc_stream_t stream_) {
  UC_REQUIRE(p && p->x && p->y && p->gamma && p->dgamma && p->dbeta, UC_ERR_BAD_SHAPE, "uc_headnorm_bwd: null pointer");
  UC_REQUIRE(p->rows > 0 && p->heads > 0 && p->ldx % 8 == 0 && p->ldy % 8 == 0 && p->ldx >= p->heads * 64 && p->ldy >= p->heads * 64,
             UC_ERR_BAD_SHAPE, "uc_headnorm_bwd: bad shape (rows %d heads %d ldx %lld ldy %lld; head_dim is 64)", p->rows, p->heads,
             (long long)p->ldx, (long long)p->ldy);
  UC_REQUIRE(((uintptr_t)p->x % 16 == 0) && ((uintptr_t)p->y % 16 == 0), UC_ERR_BAD_SHAPE, "uc_headnorm_bwd: x / g must be 16-byte aligned");
  headnorm_bwd_kernel<<<blocks_for((int64_t)p->rows * p->heads * 8, 4), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<__nv_bfloat16*>(p->y), p->ldy, static_cast<const __nv_bfloat16*>(p->x), p->ldx, p->gamma, p->dgamma, p->dbeta, p->rows,
      p->heads, p->eps);
  return check_launch("uc_headnorm_bwd");
}

extern "C" int uc_layerscale_fwd(const void* z, const void* res, const float* gamma, void* out, int32_t rows, int32_t cols,
                                 uc_stream_t stream_) {
  UC_REQUIRE(z && gamma && out && rows > 0 && cols > 0 && cols % 8 == 0, UC_ERR_BAD_SHAPE, "uc_layerscale_fwd: bad arguments");
  const int64_t n8 = (int64_t)rows * cols / 8;
  layerscale_fwd_kernel<<<blocks_for(n8, 16), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const __nv_bfloat16*>(z), static_cast<const __nv_bfloat16*>(res), gamma, static_cast<__nv_bfloat16*>(out), n8, cols / 8);
  return check_launch("uc_layerscale_fwd");
}

extern "C" int uc_layerscale_bwd(const void* dy, const void* z, const float* gamma, void* dz, float* dgamma, int32_t rows, int32_t cols,
                                 uc_stream_t stream_) {
  UC_REQUIRE(dy && z && gamma && dz && dgamma && rows > 0 && cols > 0 && cols % 8 == 0, UC_ERR_BAD_SHAPE, "uc_layerscale_bwd: bad arguments");
  dim3 grid((cols + 255) / 256, 1);
  int slabs = (sm_count() * 4 + grid.x - 1) / grid.x;
  if (slabs > (rows + 7) / 8) slabs = (rows + 7) / 8;
  grid.y = slabs < 1 ? 1 : slabs;
  layerscale_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(z), gamma, static_cast<__nv_bfloat16*>(dz), dgamma, rows, cols);
  return check_launch("uc_layerscale_bwd");
}

// ------------------------------------------------------------------------------------------------
// Row softmax for the un-fused attention of head dims other than 64 (the DiffAttention family,
// utils/transformer_blocks.py:686-945: scores and their gradients are [queries, keys] matrices produced by uc_gemm).
// One warp per row, fp32 statistics; columns >= valid (padding up to a multiple of 64 keys) come out as exact zeros.
//   fwd:  P  = softmax(scale * S[:, :valid])                              S fp32 -> P bf16
//   bwd:  dS = scale * P o (dP - rowsum(P o dP))                          P bf16, dP fp32 -> dS bf16
// ------------------------------------------------------------------------------------------------
namespace uc {
namespace {
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__global__ void softmax_rows_fwd_kernel(const float* __restrict__ S, __nv_bfloat16* __restrict__ P, int rows, int valid, int ld, float scale) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps) {
    const float* s = S + (size_t)row * ld;
    __nv_bfloat16* p = P + (size_t)row * ld;
    float m = -INFINITY;
    for (int c = lane; c < valid; c += 32) m = fmaxf(m, s[c]);
    m = warp_max(m) * scale;
    float l = 0.f;
    for (int c = lane; c < valid; c += 32) l += __expf(s[c] * scale - m);
    l = warp_sum(l);
    const float inv = 1.f / l;
    for (int c = lane; c < ld; c += 32) p[c] = __float2bfloat16_rn(c < valid ? __expf(s[c] * scale - m) * inv : 0.f);
  }
}
__global__ void softmax_rows_bwd_kernel(const __nv_bfloat16* __restrict__ P, const float* __restrict__ dP, __nv_bfloat16* __restrict__ dS,
                                        int rows, int valid, int ld, float scale) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps) {
    const __nv_bfloat16* p = P + (size_t)row * ld;
    const float* dp = dP + (size_t)row * ld;
    __nv_bfloat16* ds = dS + (size_t)row * ld;
    float dot = 0.f;
    for (int c = lane; c < valid; c += 32) dot += __bfloat162float(p[c]) * dp[c];
    dot = warp_sum(dot);
    for (int c = lane; c < ld; c += 32) ds[c] = __float2bfloat16_rn(c < valid ? scale * __bfloat162float(p[c]) * (dp[c] - dot) : 0.f);
  }
}
}  // namespace
}  // namespace uc

extern "C" int uc_softmax_rows_fwd(const float* s, void* p, int32_t rows, int32_t valid, int32_t ld, float scale, uc_stream_t st) {
  using namespace uc;
  UC_REQUIRE(s && p && rows > 0 && valid > 0 && ld >= valid, UC_ERR_BAD_SHAPE, "uc_softmax_rows_fwd: bad arguments");
  const int blocks = (int)std::min<long long>(((long long)rows * 32 + 255) / 256, (long long)sm_count() * 8);
  softmax_rows_fwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(st)>>>(s, static_cast<__nv_bfloat16*>(p), rows, valid, ld, scale);
  return check_launch("uc_softmax_rows_fwd");
}

extern "C" int uc_softmax_rows_bwd(const void* p, const float* dp, void* ds, int32_t rows, int32_t valid, int32_t ld, float scale, uc_stream_t st) {
  using namespace uc;
  UC_REQUIRE(p && dp && ds && rows > 0 && valid > 0 && ld >= valid, UC_ERR_BAD_SHAPE, "uc_softmax_rows_bwd: bad arguments");
  const int blocks = (int)std::min<long long>(((long long)rows * 32 + 255) / 256, (long long)sm_count() * 8);
  softmax_rows_bwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(st)>>>(static_cast<const __nv_bfloat16*>(p), dp,
                                                                             static_cast<__nv_bfloat16*>(ds), rows, valid, ld, scale);
  return check_launch("uc_softmax_rows_bwd");
}
