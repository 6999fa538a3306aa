// Memory-bound kernels of the path: 2-D RoPE (standalone + table), LayerNorm fwd/bwd, patch gather,
// column sums, casts, NLC<->NCHW layout converters, pixel-shuffle + pointmap/confidence adaptor.
// All are HBM-bound: 128-bit loads/stores, warp-shuffle reductions, fp32 statistics, grids sized in
// multiples of the SM count.
#include "common.cuh"
#include <stdlib.h>

namespace uc {
namespace {

// ------------------------------------------------------------------------------------------------
// 2-D RoPE, standalone in-place (the reference's native op: curope/kernels.cu:17-82)
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

template <typename T>
__global__ void rope2d_kernel(T* __restrict__ tok, const int64_t* __restrict__ pos, int B, int N, int H, int D, int64_t sb,
                              int64_t sn, int64_t sh, float base, float fwd) {
  const int Q = D / 4;
  const int64_t total = (int64_t)B * N * H * 2 * Q;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = idx % Q;
    const int X = (idx / Q) % 2;
    const int h = (idx / (2 * Q)) % H;
    const int64_t bn = idx / (2 * Q * H);
    const int n = bn % N;
    const int b = bn / N;
    const float p = (float)pos[bn * 2 + X];
    const float ang = p * (fwd / powf(base, (float)i / (float)Q));
    float s, c;
    sincosf(ang, &s, &c);
    T* t = tok + b * sb + n * sn + h * sh + 2 * Q * X + i;
    const float u = to_f<T>(t[0]), v = to_f<T>(t[Q]);
    t[0] = from_f<T>(u * c - v * s);
    t[Q] = from_f<T>(v * c + u * s);
  }
}

__global__ void rope_table_kernel(float* __restrict__ table, int P, int Q, float base, float fwd) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P * Q) return;
  const int i = idx % Q, p = idx / Q;
  const float ang = (float)p * (fwd / powf(base, (float)i / (float)Q));
  float s, c;
  sincosf(ang, &s, &c);
  table[2 * idx] = c;
  table[2 * idx + 1] = s;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row held in registers (two-pass fp32 statistics, biased variance)
// ------------------------------------------------------------------------------------------------
constexpr int LN_MAX_CHUNKS = 8;  // C <= 32 lanes * 8 elems * 8 chunks = 2048

__device__ __forceinline__ void load8(const void* base, int dtype, int64_t elem_off, float (&v)[8]) {
  if (dtype == UC_DTYPE_BF16) {
    uint4 t = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(base) + elem_off));
    v[0] = bf16_lo(t.x); v[1] = bf16_hi(t.x); v[2] = bf16_lo(t.y); v[3] = bf16_hi(t.y);
    v[4] = bf16_lo(t.z); v[5] = bf16_hi(t.z); v[6] = bf16_lo(t.w); v[7] = bf16_hi(t.w);
  } else {
    const float4* q = reinterpret_cast<const float4*>(static_cast<const float*>(base) + elem_off);
    float4 a = __ldg(q), b = __ldg(q + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}
__device__ __forceinline__ void store8(void* base, int dtype, int64_t elem_off, const float (&v)[8]) {
  if (dtype == UC_DTYPE_BF16) {
    uint4 t;
    t.x = pack_bf16(v[0], v[1]); t.y = pack_bf16(v[2], v[3]); t.z = pack_bf16(v[4], v[5]); t.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(base) + elem_off) = t;
  } else {
    float4* q = reinterpret_cast<float4*>(static_cast<float*>(base) + elem_off);
    q[0] = make_float4(v[0], v[1], v[2], v[3]);
    q[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

template <int CH>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(uc_layernorm_fwd_params p) {
  // one warp per row, TWO rows in flight per warp (the kernel is latency-bound, not issue-bound)
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const float inv_c = 1.0f / (float)p.C;
  const int stride = gridDim.x * warps_per_block;
  for (int row0 = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row0 < p.rows; row0 += 2 * stride) {
    const int row1 = row0 + stride;
    const bool has1 = row1 < p.rows;
    float x[2][CH][8];
    float sum[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = r ? row1 : row0;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int col = c * 256 + lane * 8;
        if (col < p.C && (r == 0 || has1)) {
          load8(p.x, p.x_dtype, (int64_t)row * p.C + col, x[r][c]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) x[r][c][j] = 0.f;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) sum[r] += x[r][c][j];
    const float mean[2] = {warp_sum(sum[0]) * inv_c, warp_sum(sum[1]) * inv_c};
    float sq[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int col = c * 256 + lane * 8;
        if (col < p.C) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float d = x[r][c][j] - mean[r]; sq[r] += d * d; }
        }
      }
    const float rstd[2] = {rsqrtf(warp_sum(sq[0]) * inv_c + p.eps), rsqrtf(warp_sum(sq[1]) * inv_c + p.eps)};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = r ? row1 : row0;
      if (r == 1 && !has1) break;
      if (lane == 0) {
        if (p.mean) p.mean[row] = mean[r];
        if (p.rstd) p.rstd[row] = rstd[r];
      }
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int col = c * 256 + lane * 8;
        if (col < p.C) {
          float g[8], b[8], y[8];
          load8(p.gamma, UC_DTYPE_F32, col, g);
          load8(p.beta, UC_DTYPE_F32, col, b);
#pragma unroll
          for (int j = 0; j < 8; ++j) y[j] = (x[r][c][j] - mean[r]) * rstd[r] * g[j] + b[j];
          store8(p.y, p.y_dtype, (int64_t)row * p.C + col, y);
        }
      }
    }
  }
}

// Hot-path variant: bf16 in -> bf16 out, C == CH * 256.  The row stays in its packed 16-byte form (CH uint4 per
// lane), so the kernel needs < 64 registers and 32 warps are resident per SM; the NEXT row of the warp is requested
// before the current one is reduced (one exposed memory latency per warp, not per row).
template <int CH>
__global__ void __launch_bounds__(256, CH >= 4 ? 3 : 4) layernorm_fwd_bf16_kernel(uc_layernorm_fwd_params p) {
  const int lane = threadIdx.x & 31;
  const int stride = gridDim.x * (blockDim.x >> 5);
  const float inv_c = 1.0f / (float)(CH * 256);
  const uint4* xin = static_cast<const uint4*>(p.x);
  uint4* yout = static_cast<uint4*>(p.y);
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  pdl_launch_dependents();
  pdl_wait();
  if (row >= p.rows) return;
  uint4 cur[CH], nxt[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) cur[c] = __ldg(xin + (int64_t)row * (CH * 32) + c * 32 + lane);
  for (; row < p.rows; row += stride) {
    const int nrow = row + stride;
    if (nrow < p.rows) {
#pragma unroll
      for (int c = 0; c < CH; ++c) nxt[c] = __ldg(xin + (int64_t)nrow * (CH * 32) + c * 32 + lane);
    }
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      s0 += (bf16_lo(cur[c].x) + bf16_hi(cur[c].x)) + (bf16_lo(cur[c].y) + bf16_hi(cur[c].y));
      s1 += (bf16_lo(cur[c].z) + bf16_hi(cur[c].z)) + (bf16_lo(cur[c].w) + bf16_hi(cur[c].w));
    }
    const float mean = warp_sum(s0 + s1) * inv_c;
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const uint32_t w[4] = {cur[c].x, cur[c].y, cur[c].z, cur[c].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float d0 = bf16_lo(w[j]) - mean, d1 = bf16_hi(w[j]) - mean;
        q0 = fmaf(d0, d0, q0);
        q1 = fmaf(d1, d1, q1);
      }
    }
    const float rstd = rsqrtf(warp_sum(q0 + q1) * inv_c + p.eps);
    if (lane == 0) {
      if (p.mean) p.mean[row] = mean;
      if (p.rstd) p.rstd[row] = rstd;
    }
    const float nm = -mean * rstd;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int col = c * 256 + lane * 8;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + col)), g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + col) + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta + col)), b1 = __ldg(reinterpret_cast<const float4*>(p.beta + col) + 1);
      uint4 o;
      o.x = pack_bf16(fmaf(fmaf(bf16_lo(cur[c].x), rstd, nm), g0.x, b0.x), fmaf(fmaf(bf16_hi(cur[c].x), rstd, nm), g0.y, b0.y));
      o.y = pack_bf16(fmaf(fmaf(bf16_lo(cur[c].y), rstd, nm), g0.z, b0.z), fmaf(fmaf(bf16_hi(cur[c].y), rstd, nm), g0.w, b0.w));
      o.z = pack_bf16(fmaf(fmaf(bf16_lo(cur[c].z), rstd, nm), g1.x, b1.x), fmaf(fmaf(bf16_hi(cur[c].z), rstd, nm), g1.y, b1.y));
      o.w = pack_bf16(fmaf(fmaf(bf16_lo(cur[c].w), rstd, nm), g1.z, b1.z), fmaf(fmaf(bf16_hi(cur[c].w), rstd, nm), g1.w, b1.w));
      yout[(int64_t)row * (CH * 32) + c * 32 + lane] = o;
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) cur[c] = nxt[c];
  }
}

// dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat)) [+ dres];  dgamma += sum dy*xhat;  dbeta += sum dy
// One warp per row, latency-bound: x, dy and dres of a row are requested TOGETHER (one exposed memory latency per
// row) and kept in registers in their packed 16-byte form; only the 2*CH*8 dgamma/dbeta partial sums persist across
// rows.  Block partials are combined warp-by-warp in smem (no smem atomics), one global atomicAdd per column.
struct Raw8 { uint4 a; uint4 b; };  // 8 elements: bf16 -> a only, fp32 -> a,b
__device__ __forceinline__ Raw8 load_raw8(const void* base, int dtype, int64_t off) {
  Raw8 r;
  if (dtype == UC_DTYPE_BF16) {
    r.a = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(base) + off));
    r.b = make_uint4(0, 0, 0, 0);
  } else {
    const uint4* q = reinterpret_cast<const uint4*>(static_cast<const float*>(base) + off);
    r.a = __ldg(q);
    r.b = __ldg(q + 1);
  }
  return r;
}
__device__ __forceinline__ void unpack8(const Raw8& r, int dtype, float (&v)[8]) {
  if (dtype == UC_DTYPE_BF16) {
    v[0] = bf16_lo(r.a.x); v[1] = bf16_hi(r.a.x); v[2] = bf16_lo(r.a.y); v[3] = bf16_hi(r.a.y);
    v[4] = bf16_lo(r.a.z); v[5] = bf16_hi(r.a.z); v[6] = bf16_lo(r.a.w); v[7] = bf16_hi(r.a.w);
  } else {
    v[0] = __uint_as_float(r.a.x); v[1] = __uint_as_float(r.a.y); v[2] = __uint_as_float(r.a.z); v[3] = __uint_as_float(r.a.w);
    v[4] = __uint_as_float(r.b.x); v[5] = __uint_as_float(r.b.y); v[6] = __uint_as_float(r.b.z); v[7] = __uint_as_float(r.b.w);
  }
}

template <int CH, bool BF16_IN>
__global__ void __launch_bounds__(256, 2) layernorm_bwd_kernel(uc_layernorm_bwd_params p) {
  extern __shared__ float red[];  // [2][C]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  const float inv_c = 1.0f / (float)p.C;
  const int xdt = BF16_IN ? UC_DTYPE_BF16 : p.x_dtype, ydt = BF16_IN ? UC_DTYPE_BF16 : p.dy_dtype;
  for (int i = threadIdx.x; i < 2 * p.C; i += blockDim.x) red[i] = 0.f;
  float dg[CH][8], db[CH][8];
#pragma unroll
  for (int c = 0; c < CH; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) { dg[c][j] = 0.f; db[c][j] = 0.f; }

  for (int row = blockIdx.x * warps_per_block + warp; row < p.rows; row += gridDim.x * warps_per_block) {
    const int64_t base = (int64_t)row * p.C;
    Raw8 rx[CH], rdy[CH];
    uint4 rres[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int col = c * 256 + lane * 8;
      if (col < p.C) {
        rx[c] = load_raw8(p.x, xdt, base + col);
        rdy[c] = load_raw8(p.dy, ydt, base + col);
        if (p.dres) rres[c] = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.dres) + base + col));
      }
    }
    const float mean = p.mean[row], rstd = p.rstd[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int col = c * 256 + lane * 8;
      if (col < p.C) {
        float x[8], dy[8], g[8];
        unpack8(rx[c], xdt, x);
        unpack8(rdy[c], ydt, dy);
        load8(p.gamma, UC_DTYPE_F32, col, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (x[j] - mean) * rstd;
          const float gy = g[j] * dy[j];
          s1 += gy;
          s2 += gy * xh;
          dg[c][j] += dy[j] * xh;
          db[c][j] += dy[j];
        }
      }
    }
    s1 = warp_sum(s1) * inv_c;
    s2 = warp_sum(s2) * inv_c;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int col = c * 256 + lane * 8;
      if (col < p.C) {
        float x[8], dy[8], g[8], dx[8];
        unpack8(rx[c], xdt, x);
        unpack8(rdy[c], ydt, dy);
        load8(p.gamma, UC_DTYPE_F32, col, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) dx[j] = rstd * (g[j] * dy[j] - s1 - (x[j] - mean) * rstd * s2);
        if (p.dres) {
          Raw8 rr;
          rr.a = rres[c];
          float r[8];
          unpack8(rr, UC_DTYPE_BF16, r);
#pragma unroll
          for (int j = 0; j < 8; ++j) dx[j] += r[j];
        }
        store8(p.dx, UC_DTYPE_BF16, base + col, dx);
      }
    }
  }
  if (p.dgamma) {
    for (int w = 0; w < warps_per_block; ++w) {
      __syncthreads();
      if (warp == w) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const int col = c * 256 + lane * 8;
          if (col < p.C) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              red[col + j] += dg[c][j];
              red[p.C + col + j] += db[c][j];
            }
          }
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.C; i += blockDim.x) {
      atomicAdd(p.dgamma + i, red[i]);
      atomicAdd(p.dbeta + i, red[p.C + i]);
    }
  }
}

// Hot-path variant (bf16 x / dy / dres, C % 128 == 0, C <= 1024).  One 768-thread block per SM holds S = 768 / (C/4)
// independent TEAMS of C/4 threads; a team walks batches of G rows with thread <-> 4 columns, so only 3 x 4 column
// partials (dgamma, dbeta, colsum(dx)) persist in registers and G rows x 3 tensors of 8-byte loads are in flight per
// thread.  The 2G row sums of a batch are reduced with a value-halving shuffle butterfly per warp and two TEAM barriers
// (bar.sync with a team id).  At the end the teams' column partials are combined in shared memory and leave as ONE
// red.global.add.v4.f32 per 4 columns per block: same-address L2 atomics serialise (~30 ns each), so 148 of them per
// column instead of one per (team, column) is the difference between a 2 us and a 25 us tail.
template <int G>
__global__ void __launch_bounds__(768, 1) layernorm_bwd_rows_kernel(uc_layernorm_bwd_params p) {
  __shared__ float part[4][8][2 * G];
  __shared__ float tot[4][2 * G];
  __shared__ __align__(16) float colacc[3][1024];
  const int team_threads = p.C >> 2;
  const int team = threadIdx.x / team_threads, tid = threadIdx.x - team * team_threads;
  const int teams = blockDim.x / team_threads;
  const int lane = tid & 31, warp = tid >> 5, nwarps = team_threads >> 5;
  const int col = tid * 4;
  const float inv_c = 1.0f / (float)p.C;
  for (int i = threadIdx.x; i < 3 * 1024; i += blockDim.x) (&colacc[0][0])[i] = 0.f;
  pdl_launch_dependents();
  pdl_wait();
  const float4 gm = __ldg(reinterpret_cast<const float4*>(p.gamma + col));
  const uint2* xin = static_cast<const uint2*>(p.x) + tid;
  const uint2* dyin = static_cast<const uint2*>(p.dy) + tid;
  const uint2* rin = p.dres ? static_cast<const uint2*>(p.dres) + tid : nullptr;
  uint2* dxout = static_cast<uint2*>(p.dx) + tid;
  const int rowq = p.C >> 2;  // uint2 per row
  float dg[4] = {0, 0, 0, 0}, db[4] = {0, 0, 0, 0}, dc[4] = {0, 0, 0, 0};
  for (int row0 = (blockIdx.x * teams + team) * G; row0 < p.rows; row0 += gridDim.x * teams * G) {
    uint2 rx[G], rdy[G], rr[G];
    float mean[G], rstd[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int row = min(row0 + g, p.rows - 1);  // tail rows are clamped (re-read) and masked below
      rx[g] = __ldg(xin + (int64_t)row * rowq);
      rdy[g] = __ldg(dyin + (int64_t)row * rowq);
      rr[g] = rin ? __ldg(rin + (int64_t)row * rowq) : make_uint2(0u, 0u);
      mean[g] = __ldg(p.mean + row);
      rstd[g] = __ldg(p.rstd + row);
    }
    float s[2 * G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float live = (row0 + g < p.rows) ? 1.0f : 0.0f;
      const float x4[4] = {bf16_lo(rx[g].x), bf16_hi(rx[g].x), bf16_lo(rx[g].y), bf16_hi(rx[g].y)};
      const float d4[4] = {bf16_lo(rdy[g].x) * live, bf16_hi(rdy[g].x) * live, bf16_lo(rdy[g].y) * live, bf16_hi(rdy[g].y) * live};
      const float g4[4] = {gm.x, gm.y, gm.z, gm.w};
      const float nm = -mean[g] * rstd[g];
      float a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xh = fmaf(x4[j], rstd[g], nm);
        const float gy = g4[j] * d4[j];
        a1 += gy;
        a2 = fmaf(gy, xh, a2);
        dg[j] = fmaf(d4[j], xh, dg[j]);
        db[j] += d4[j];
      }
      s[g] = a1;
      s[G + g] = a2;
    }
    // warp reduction of 2G values with value halving: lane l ends with the warp total of value l / kRest
#pragma unroll
    for (int n = G, k = 16; n >= 1; n >>= 1, k >>= 1) {
      const bool up = (lane & k) != 0;
#pragma unroll
      for (int i = 0; i < n; ++i) {
        const float send = up ? s[i] : s[i + n];
        const float keep = up ? s[i + n] : s[i];
        s[i] = keep + __shfl_xor_sync(0xffffffffu, send, k);
      }
    }
    constexpr int kRest = 32 / (2 * G);  // lanes that still hold partial sums of the same value
#pragma unroll
    for (int k = kRest >> 1; k >= 1; k >>= 1) s[0] += __shfl_xor_sync(0xffffffffu, s[0], k);
    if ((lane & (kRest - 1)) == 0) part[team][warp][lane / kRest] = s[0];
    asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(team_threads) : "memory");
    if (tid < 2 * G) {
      float t = 0.f;
      for (int w = 0; w < nwarps; ++w) t += part[team][w][tid];
      tot[team][tid] = t * inv_c;
    }
    asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(team_threads) : "memory");
#pragma unroll
    for (int g = 0; g < G; ++g) {
      if (row0 + g >= p.rows) break;
      const float s1 = tot[team][g], s2 = tot[team][G + g];
      const float x4[4] = {bf16_lo(rx[g].x), bf16_hi(rx[g].x), bf16_lo(rx[g].y), bf16_hi(rx[g].y)};
      const float d4[4] = {bf16_lo(rdy[g].x), bf16_hi(rdy[g].x), bf16_lo(rdy[g].y), bf16_hi(rdy[g].y)};
      const float r4[4] = {bf16_lo(rr[g].x), bf16_hi(rr[g].x), bf16_lo(rr[g].y), bf16_hi(rr[g].y)};
      const float g4[4] = {gm.x, gm.y, gm.z, gm.w};
      const float nm = -mean[g] * rstd[g];
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xh = fmaf(x4[j], rstd[g], nm);
        const float t = fmaf(-xh, s2, fmaf(g4[j], d4[j], -s1));
        o[j] = round_bf16(fmaf(t, rstd[g], r4[j]));
        dc[j] += o[j];
      }
      uint2 w;
      w.x = pack_bf16(o[0], o[1]);
      w.y = pack_bf16(o[2], o[3]);
      dxout[(int64_t)(row0 + g) * rowq] = w;
    }
  }
  // combine the teams' column partials in shared memory, then one vector reduction per 4 columns per block
  __syncthreads();  // colacc zero-initialised
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    atomicAdd(&colacc[0][col + j], dg[j]);
    atomicAdd(&colacc[1][col + j], db[j]);
    atomicAdd(&colacc[2][col + j], dc[j]);
  }
  __syncthreads();
  if (team == 0) {
    if (p.dgamma) {
      const float4 a = *reinterpret_cast<const float4*>(&colacc[0][col]), b = *reinterpret_cast<const float4*>(&colacc[1][col]);
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.dgamma + col), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w) : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.dbeta + col), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
    }
    if (p.dx_colsum) {
      const float4 c = *reinterpret_cast<const float4*>(&colacc[2][col]);
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.dx_colsum + col), "f"(c.x), "f"(c.y), "f"(c.z), "f"(c.w) : "memory");
    }
  }
}

// ------------------------------------------------------------------------------------------------
// patch gather: img fp32 [B][C][H][W] -> cols bf16 [B*h*w][C*p*p]   (p % 8 == 0)
// ------------------------------------------------------------------------------------------------
__global__ void patchify_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ cols, int B, int C, int H, int W,
                                int p) {
  const int h = H / p, w = W / p, K = C * p * p, p8 = p / 8;
  const int64_t total = (int64_t)B * h * w * C * p * p8;  // groups of 8 consecutive j
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int j8 = idx % p8;
    const int i = (idx / p8) % p;
    const int c = (idx / ((int64_t)p8 * p)) % C;
    const int64_t tok = idx / ((int64_t)p8 * p * C);
    const int px = tok % w, py = (tok / w) % h, b = tok / ((int64_t)w * h);
    const float* src = img + (((int64_t)b * C + c) * H + (py * p + i)) * W + px * p + j8 * 8;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src)), d = __ldg(reinterpret_cast<const float4*>(src) + 1);
    uint4 t;
    t.x = pack_bf16(a.x, a.y); t.y = pack_bf16(a.z, a.w); t.z = pack_bf16(d.x, d.y); t.w = pack_bf16(d.z, d.w);
    *reinterpret_cast<uint4*>(cols + tok * K + (c * p + i) * p + j8 * 8) = t;
  }
}

// any patch size (e.g. 14 for the 518x518 / DINOv2-style grids): one thread per output element, row pitch Kp = K rounded
// up to 64 elements (GEMM K / N granularity), pad columns zero
__global__ void patchify_generic_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ cols, int B, int C, int H, int W,
                                        int p, int Kp) {
  const int h = H / p, w = W / p, K = C * p * p;
  const int64_t total = (int64_t)B * h * w * Kp;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = idx % Kp;
    const int64_t tok = idx / Kp;
    float v = 0.f;
    if (k < K) {
      const int j = k % p, i = (k / p) % p, c = k / (p * p);
      const int px = tok % w, py = (tok / w) % h;
      const int64_t b = tok / ((int64_t)w * h);
      v = __ldg(img + ((b * C + c) * H + (py * p + i)) * W + px * p + j);
    }
    cols[idx] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------------------------
// column sums (bias gradients): out[col] += sum_rows x[row][col]
// block = 8 warps; a warp covers 256 columns (8 per lane); blockIdx.y strides over row slabs
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_kernel(const void* __restrict__ x, int dtype, int64_t ld, int rows, int cols,
                                                     float* __restrict__ out) {
  __shared__ float red[8][256];
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col < cols) {
    const int stride = gridDim.y * 8;
    int row = blockIdx.y * 8 + warp;
    for (; row + 3 * stride < rows; row += 4 * stride) {  // 4 independent 16-byte loads in flight per lane
      float v0[8], v1[8], v2[8], v3[8];
      load8(x, dtype, (int64_t)row * ld + col, v0);
      load8(x, dtype, (int64_t)(row + stride) * ld + col, v1);
      load8(x, dtype, (int64_t)(row + 2 * stride) * ld + col, v2);
      load8(x, dtype, (int64_t)(row + 3 * stride) * ld + col, v3);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += (v0[j] + v1[j]) + (v2[j] + v3[j]);
    }
    for (; row < rows; row += stride) {
      float v[8];
      load8(x, dtype, (int64_t)row * ld + col, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
  __syncthreads();
  const int c = threadIdx.x;
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w][c];
  if (blockIdx.x * 256 + c < cols) atomicAdd(out + blockIdx.x * 256 + c, s);
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  const int64_t n8 = n / 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i), b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    uint4 t;
    t.x = pack_bf16(a.x, a.y); t.y = pack_bf16(a.z, a.w); t.z = pack_bf16(b.x, b.y); t.w = pack_bf16(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[i] = t;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) dst[n8 * 8 + threadIdx.x] = __float2bfloat16_rn(src[n8 * 8 + threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------
// [B][L][C] <-> fp32 [B][C][L] tiled transposes (bit-exact layout ops, plus dtype conversion)
// ------------------------------------------------------------------------------------------------
template <typename TS>
__global__ void nlc_to_nchw_kernel(const TS* __restrict__ src, float* __restrict__ dst, int L, int C) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, l0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int l = l0 + r, c = c0 + threadIdx.x;
    if (l < L && c < C) tile[r][threadIdx.x] = to_f<TS>(src[((int64_t)b * L + l) * C + c]);
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, l = l0 + threadIdx.x;
    if (l < L && c < C) dst[((int64_t)b * C + c) * L + l] = tile[threadIdx.x][r];
  }
}
template <typename TD>
__global__ void nchw_to_nlc_kernel(const float* __restrict__ src, TD* __restrict__ dst, int L, int C) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, l0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, l = l0 + threadIdx.x;
    if (l < L && c < C) tile[r][threadIdx.x] = src[((int64_t)b * C + c) * L + l];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int l = l0 + r, c = c0 + threadIdx.x;
    if (l < L && c < C) dst[((int64_t)b * L + l) * C + c] = from_f<TD>(tile[threadIdx.x][r]);
  }
}

// ------------------------------------------------------------------------------------------------
// linear-head post-processing: pixel_shuffle gather + pointmap(exp) / confidence(exp) adaptor, BHWC out
// ------------------------------------------------------------------------------------------------
__global__ void head_post_fwd_kernel(uc_head_post_fwd_params p) {
  const int Hh = p.h * p.patch, Ww = p.w * p.patch, pp = p.patch * p.patch;
  const int64_t total = (int64_t)p.B * Hh * Ww;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int X = idx % Ww, Y = (idx / Ww) % Hh, b = idx / ((int64_t)Ww * Hh);
    const int64_t tok = ((int64_t)b * p.h + Y / p.patch) * p.w + X / p.patch;
    const int64_t ld = p.ldy ? p.ldy : 4 * pp;
    const float* y = p.y + tok * ld + (Y % p.patch) * p.patch + (X % p.patch);
    const float x0 = y[0], x1 = y[pp], x2 = y[2 * pp], c = y[3 * pp];
    const float d = sqrtf(x0 * x0 + x1 * x1 + x2 * x2);
    const float s = expm1f(d) / fmaxf(d, 1e-8f);
    float* o = p.pts + idx * 3;
    o[0] = x0 * s; o[1] = x1 * s; o[2] = x2 * s;
    p.conf[idx] = p.conf_min + fminf(expf(c), p.conf_max - p.conf_min);
  }
}

__global__ void head_post_bwd_kernel(uc_head_post_bwd_params p) {
  const int Hh = p.h * p.patch, Ww = p.w * p.patch, pp = p.patch * p.patch;
  const int64_t total = (int64_t)p.B * Hh * Ww;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int X = idx % Ww, Y = (idx / Ww) % Hh, b = idx / ((int64_t)Ww * Hh);
    const int64_t tok = ((int64_t)b * p.h + Y / p.patch) * p.w + X / p.patch;
    const int64_t off = tok * (p.ldy ? p.ldy : 4 * pp) + (Y % p.patch) * p.patch + (X % p.patch);
    const float* y = p.y + off;
    const float x0 = y[0], x1 = y[pp], x2 = y[2 * pp], c = y[3 * pp];
    const float g0 = p.dpts[idx * 3], g1 = p.dpts[idx * 3 + 1], g2 = p.dpts[idx * 3 + 2];
    const float d = sqrtf(x0 * x0 + x1 * x1 + x2 * x2);
    const float ed = expf(d), em1 = expm1f(d);
    float s, sp_over_d;  // s(d) and s'(d)/d
    if (d >= 1e-8f) {
      s = em1 / d;
      sp_over_d = (ed * d - em1) / (d * d * d);
    } else {
      s = em1 * 1e8f;
      sp_over_d = d > 0.f ? ed * 1e8f / d : 0.f;
    }
    const float dot = g0 * x0 + g1 * x1 + g2 * x2;
    const float r0 = s * g0 + sp_over_d * x0 * dot;
    const float r1 = s * g1 + sp_over_d * x1 * dot;
    const float r2 = s * g2 + sp_over_d * x2 * dot;
    const float ec = expf(c);
    const float r3 = (ec <= p.conf_max - p.conf_min) ? p.dconf[idx] * ec : 0.f;
    if (p.dy_dtype == UC_DTYPE_BF16) {
      __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.dy) + off;
      o[0] = __float2bfloat16_rn(r0); o[pp] = __float2bfloat16_rn(r1);
      o[2 * pp] = __float2bfloat16_rn(r2); o[3 * pp] = __float2bfloat16_rn(r3);
    } else {
      float* o = static_cast<float*>(p.dy) + off;
      o[0] = r0; o[pp] = r1; o[2 * pp] = r2; o[3 * pp] = r3;
    }
  }
}

inline int grid_for(int64_t work_items, int threads, int per_sm = 8) {
  int64_t blocks = (work_items + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace
}  // namespace uc

using namespace uc;

extern "C" int uc_rope2d(const uc_rope2d_params* p, uc_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // same argument checks as the reference's TORCH_CHECKs (curope.cpp:54-59, kernels.cu:91-94)
  UC_REQUIRE(p && p->tokens && p->positions, UC_ERR_BAD_SHAPE, "uc_rope2d: null pointer");
  UC_REQUIRE(p->D % 4 == 0, UC_ERR_BAD_SHAPE, "uc_rope2d: token dim must be multiple of 4 (got %d)", p->D);
  UC_REQUIRE(p->B > 0 && p->N > 0 && p->H > 0, UC_ERR_BAD_SHAPE, "uc_rope2d: bad shape");
  const int64_t total = (int64_t)p->B * p->N * p->H * (p->D / 2);
  const int grid = grid_for(total, 256, 16);
  if (p->dtype == UC_DTYPE_F32)
    rope2d_kernel<float><<<grid, 256, 0, stream>>>(static_cast<float*>(p->tokens), p->positions, p->B, p->N, p->H, p->D,
                                                   p->stride_b, p->stride_n, p->stride_h, p->base, p->fwd);
  else if (p->dtype == UC_DTYPE_BF16)
    rope2d_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<__nv_bfloat16*>(p->tokens), p->positions, p->B, p->N,
                                                           p->H, p->D, p->stride_b, p->stride_n, p->stride_h, p->base, p->fwd);
  else if (p->dtype == UC_DTYPE_F16)
    rope2d_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<__half*>(p->tokens), p->positions, p->B, p->N, p->H, p->D,
                                                    p->stride_b, p->stride_n, p->stride_h, p->base, p->fwd);
  else
    UC_REQUIRE(false, UC_ERR_BAD_DTYPE, "uc_rope2d: unsupported dtype %d", p->dtype);
  return check_launch("uc_rope2d");
}

extern "C" int uc_rope2d_table(float* table, int32_t P, int32_t Q, float base, float fwd, uc_stream_t stream_) {
  UC_REQUIRE(table && P > 0 && Q > 0, UC_ERR_BAD_SHAPE, "uc_rope2d_table: bad arguments");
  rope_table_kernel<<<(P * Q + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream_)>>>(table, P, Q, base, fwd);
  return check_launch("uc_rope2d_table");
}

extern "C" int uc_layernorm_fwd(const uc_layernorm_fwd_params* p, uc_stream_t stream_) {
  UC_REQUIRE(p && p->x && p->y && p->gamma && p->beta, UC_ERR_BAD_SHAPE, "uc_layernorm_fwd: null pointer");
  UC_REQUIRE(p->C % 8 == 0 && p->C <= 256 * LN_MAX_CHUNKS && p->rows > 0, UC_ERR_BAD_SHAPE,
             "uc_layernorm_fwd: C=%d must be a multiple of 8 and <= %d", p->C, 256 * LN_MAX_CHUNKS);
  const int grid = grid_for((int64_t)p->rows * 32, 256, 8);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int ch = (p->C + 255) / 256;
  if (p->x_dtype == UC_DTYPE_BF16 && p->y_dtype == UC_DTYPE_BF16 && p->C % 256 == 0 && ch >= 2 && ch <= 4 &&
      ((uintptr_t)p->x % 16 == 0) && ((uintptr_t)p->y % 16 == 0)) {
    // persistent: 3-4 blocks of 8 warps per SM, each warp walks rows with a one-row prefetch
    int g = sm_count() * (ch >= 4 ? 3 : 4);
    if (g > (p->rows + 7) / 8) g = (p->rows + 7) / 8;
    cudaError_t le = ch == 2   ? launch_pdl(layernorm_fwd_bf16_kernel<2>, dim3(g), dim3(256), 0, stream, *p)
                     : ch == 3 ? launch_pdl(layernorm_fwd_bf16_kernel<3>, dim3(g), dim3(256), 0, stream, *p)
                               : launch_pdl(layernorm_fwd_bf16_kernel<4>, dim3(g), dim3(256), 0, stream, *p);
    UC_REQUIRE(le == cudaSuccess, UC_ERR_CUDA, "uc_layernorm_fwd: launch failed: %s", cudaGetErrorString(le));
    return check_launch("uc_layernorm_fwd");
  }
  if (ch <= 1) layernorm_fwd_kernel<1><<<grid, 256, 0, stream>>>(*p);
  else if (ch <= 2) layernorm_fwd_kernel<2><<<grid, 256, 0, stream>>>(*p);
  else if (ch <= 3) layernorm_fwd_kernel<3><<<grid, 256, 0, stream>>>(*p);
  else if (ch <= 4) layernorm_fwd_kernel<4><<<grid, 256, 0, stream>>>(*p);
  else layernorm_fwd_kernel<8><<<grid, 256, 0, stream>>>(*p);
  return check_launch("uc_layernorm_fwd");
}

extern "C" int uc_layernorm_bwd(const uc_layernorm_bwd_params* p, uc_stream_t stream_) {
  UC_REQUIRE(p && p->x && p->dy && p->dx && p->gamma && p->mean && p->rstd, UC_ERR_BAD_SHAPE, "uc_layernorm_bwd: null pointer");
  UC_REQUIRE(p->C % 8 == 0 && p->C <= 256 * LN_MAX_CHUNKS && p->rows > 0, UC_ERR_BAD_SHAPE,
             "uc_layernorm_bwd: C=%d must be a multiple of 8 and <= %d", p->C, 256 * LN_MAX_CHUNKS);
  UC_REQUIRE((p->dgamma == nullptr) == (p->dbeta == nullptr), UC_ERR_BAD_SHAPE, "uc_layernorm_bwd: dgamma/dbeta must both be set");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (p->x_dtype == UC_DTYPE_BF16 && p->dy_dtype == UC_DTYPE_BF16 && p->C % 128 == 0 && p->C <= 1024 &&
      ((uintptr_t)p->x % 8 == 0) && ((uintptr_t)p->dy % 8 == 0) && ((uintptr_t)p->dx % 8 == 0) && ((uintptr_t)p->dres % 8 == 0) &&
      ((uintptr_t)p->dgamma % 16 == 0) && ((uintptr_t)p->dbeta % 16 == 0) && ((uintptr_t)p->dx_colsum % 16 == 0)) {
    // one block per SM: 768 threads = 3 teams (C = 1024) / 4 teams (C = 768) / ... of C/4 threads
    const int team_threads = p->C / 4;
    int teams = 768 / team_threads;
    if (teams > 4) teams = 4;
    int grid = sm_count();
    const int need = (p->rows + 4 * teams - 1) / (4 * teams);
    if (grid > need) grid = need;
    cudaError_t le = launch_pdl(layernorm_bwd_rows_kernel<4>, dim3(grid), dim3(teams * team_threads), 0, stream, *p);
    UC_REQUIRE(le == cudaSuccess, UC_ERR_CUDA, "uc_layernorm_bwd: launch failed: %s", cudaGetErrorString(le));
    return check_launch("uc_layernorm_bwd");
  }
  UC_REQUIRE(p->dx_colsum == nullptr, UC_ERR_UNSUPPORTED, "uc_layernorm_bwd: dx_colsum needs bf16 inputs and C %% 128 == 0, C <= 1024");
  const int grid = grid_for((int64_t)p->rows * 32, 256, 2);
  const size_t sm = 2 * p->C * sizeof(float);
  const int ch = (p->C + 255) / 256;
  const bool bf = p->x_dtype == UC_DTYPE_BF16 && p->dy_dtype == UC_DTYPE_BF16;
#define UC_LN_BWD(CHV)                                                                  \
  do {                                                                                  \
    if (bf) layernorm_bwd_kernel<CHV, true><<<grid, 256, sm, stream>>>(*p);             \
    else layernorm_bwd_kernel<CHV, false><<<grid, 256, sm, stream>>>(*p);               \
  } while (0)
  if (ch <= 1) UC_LN_BWD(1);
  else if (ch <= 2) UC_LN_BWD(2);
  else if (ch <= 3) UC_LN_BWD(3);
  else if (ch <= 4) UC_LN_BWD(4);
  else UC_LN_BWD(8);
#undef UC_LN_BWD
  return check_launch("uc_layernorm_bwd");
}

extern "C" int uc_patchify(const float* img, void* cols, int32_t B, int32_t C, int32_t H, int32_t W, int32_t patch,
                           uc_stream_t stream_) {
  UC_REQUIRE(img && cols, UC_ERR_BAD_SHAPE, "uc_patchify: null pointer");
  // same divisibility assertions as PatchEmbedDust3R.forward (libs/croco/patch_embed.py:71-76)
  UC_REQUIRE(patch > 0 && H % patch == 0 && W % patch == 0, UC_ERR_BAD_SHAPE,
             "uc_patchify: image %dx%d is not a multiple of patch size %d", H, W, patch);
  if (patch % 8 == 0) {
    const int64_t total = (int64_t)B * C * H * W / 8;
    patchify_kernel<<<grid_for(total, 256, 16), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        img, static_cast<__nv_bfloat16*>(cols), B, C, H, W, patch);
  } else {
    const int Kp = (C * patch * patch + 63) / 64 * 64;
    const int64_t total = (int64_t)B * (H / patch) * (W / patch) * Kp;
    patchify_generic_kernel<<<grid_for(total, 256, 16), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        img, static_cast<__nv_bfloat16*>(cols), B, C, H, W, patch, Kp);
  }
  return check_launch("uc_patchify");
}

extern "C" int uc_colsum(const void* x, int32_t x_dtype, int64_t ld, int32_t rows, int32_t cols, float* out, uc_stream_t stream_) {
  UC_REQUIRE(x && out && rows > 0 && cols > 0 && cols % 8 == 0 && ld % 8 == 0, UC_ERR_BAD_SHAPE, "uc_colsum: bad arguments");
  dim3 grid((cols + 255) / 256, 1);
  int slabs = (sm_count() * 4 + grid.x - 1) / grid.x;
  if (slabs > (rows + 7) / 8) slabs = (rows + 7) / 8;
  grid.y = slabs < 1 ? 1 : slabs;
  cudaError_t le = launch_pdl(colsum_kernel, grid, dim3(256), 0, static_cast<cudaStream_t>(stream_), x, (int)x_dtype, (int64_t)ld,
                              (int)rows, (int)cols, out);
  UC_REQUIRE(le == cudaSuccess, UC_ERR_CUDA, "uc_colsum: launch failed: %s", cudaGetErrorString(le));
  return check_launch("uc_colsum");
}

extern "C" int uc_cast_bf16(const float* src, void* dst, int64_t n, uc_stream_t stream_) {
  UC_REQUIRE(src && dst && n > 0, UC_ERR_BAD_SHAPE, "uc_cast_bf16: bad arguments");
  cast_bf16_kernel<<<grid_for(n / 8 + 1, 256, 16), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      src, static_cast<__nv_bfloat16*>(dst), n);
  return check_launch("uc_cast_bf16");
}

extern "C" int uc_nlc_to_nchw(const void* src, int32_t src_dtype, float* dst, int32_t B, int32_t L, int32_t C, uc_stream_t stream_) {
  UC_REQUIRE(src && dst && B > 0 && L > 0 && C > 0, UC_ERR_BAD_SHAPE, "uc_nlc_to_nchw: bad arguments");
  dim3 grid((C + 31) / 32, (L + 31) / 32, B), block(32, 8);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (src_dtype == UC_DTYPE_BF16)
    nlc_to_nchw_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), dst, L, C);
  else if (src_dtype == UC_DTYPE_F32)
    nlc_to_nchw_kernel<float><<<grid, block, 0, stream>>>(static_cast<const float*>(src), dst, L, C);
  else
    UC_REQUIRE(false, UC_ERR_BAD_DTYPE, "uc_nlc_to_nchw: dtype %d", src_dtype);
  return check_launch("uc_nlc_to_nchw");
}

extern "C" int uc_nchw_to_nlc(const float* src, void* dst, int32_t dst_dtype, int32_t B, int32_t L, int32_t C, uc_stream_t stream_) {
  UC_REQUIRE(src && dst && B > 0 && L > 0 && C > 0, UC_ERR_BAD_SHAPE, "uc_nchw_to_nlc: bad arguments");
  dim3 grid((C + 31) / 32, (L + 31) / 32, B), block(32, 8);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (dst_dtype == UC_DTYPE_BF16)
    nchw_to_nlc_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>(src, static_cast<__nv_bfloat16*>(dst), L, C);
  else if (dst_dtype == UC_DTYPE_F32)
    nchw_to_nlc_kernel<float><<<grid, block, 0, stream>>>(src, static_cast<float*>(dst), L, C);
  else
    UC_REQUIRE(false, UC_ERR_BAD_DTYPE, "uc_nchw_to_nlc: dtype %d", dst_dtype);
  return check_launch("uc_nchw_to_nlc");
}

extern "C" int uc_head_post_fwd(const uc_head_post_fwd_params* p, uc_stream_t stream_) {
  UC_REQUIRE(p && p->y && p->pts && p->conf && p->B > 0 && p->h > 0 && p->w > 0 && p->patch > 0, UC_ERR_BAD_SHAPE,
             "uc_head_post_fwd: bad arguments");
  const int64_t total = (int64_t)p->B * p->h * p->w * p->patch * p->patch;
  head_post_fwd_kernel<<<grid_for(total, 256, 16), 256, 0, static_cast<cudaStream_t>(stream_)>>>(*p);
  return check_launch("uc_head_post_fwd");
}

extern "C" int uc_head_post_bwd(const uc_head_post_bwd_params* p, uc_stream_t stream_) {
  UC_REQUIRE(p && p->y && p->dpts && p->dconf && p->dy && p->B > 0 && p->h > 0 && p->w > 0 && p->patch > 0, UC_ERR_BAD_SHAPE,
             "uc_head_post_bwd: bad arguments");
  const int64_t total = (int64_t)p->B * p->h * p->w * p->patch * p->patch;
  head_post_bwd_kernel<<<grid_for(total, 256, 16), 256, 0, static_cast<cudaStream_t>(stream_)>>>(*p);
  return check_launch("uc_head_post_bwd");
}
