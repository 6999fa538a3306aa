// Memory-side kernels of the DPT head (prediction_heads/dpt.py:94-311, libs/croco/dpt_block.py:114-255).
// All feature maps are NHWC bf16 (== token-major [B*H*W, C]), so every convolution is `uc_gemm`:
//   1x1 conv            : GEMM on the tokens
//   3x3 conv (s=1,2)    : uc_im2col3x3 -> GEMM  (dgrad: GEMM -> uc_col2im3x3 gather; wgrad: GEMM on the saved columns)
//   ConvTranspose k=s   : GEMM to [(i,j,co)] columns -> uc_depth_to_space scatter (bwd: the inverse gather -> GEMM)
// plus bilinear resampling (align_corners=True) fwd/bwd and small elementwise ops.  128-bit accesses, 8 channels
// per thread, grid-stride loops sized in multiples of the SM count.
#include "common.cuh"

namespace uc {
namespace {

__device__ __forceinline__ void unpack_bf16x8(const uint4& t, float (&v)[8]) {
  v[0] = bf16_lo(t.x); v[1] = bf16_hi(t.x); v[2] = bf16_lo(t.y); v[3] = bf16_hi(t.y);
  v[4] = bf16_lo(t.z); v[5] = bf16_hi(t.z); v[6] = bf16_lo(t.w); v[7] = bf16_hi(t.w);
}
__device__ __forceinline__ uint4 pack_bf16x8(const float (&v)[8]) {
  uint4 t;
  t.x = pack_bf16(v[0], v[1]); t.y = pack_bf16(v[2], v[3]); t.z = pack_bf16(v[4], v[5]); t.w = pack_bf16(v[6], v[7]);
  return t;
}

inline int grid_for(int64_t work_items, int threads, int per_sm = 16) {
  int64_t blocks = (work_items + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// cols[(b,yo,xo)][tap*C + c] = x[b][yo*s + r - 1][xo*s + t - 1][c]  (zero outside), tap = r*3 + t
__global__ void im2col3x3_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ cols, int B, int H, int W, int C,
                                 int Ho, int Wo, int stride) {
  const int c8 = C / 8;
  const int64_t total = (int64_t)B * Ho * Wo * 9 * c8;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int cg = idx % c8;
    const int tap = (idx / c8) % 9;
    const int64_t pix = idx / ((int64_t)c8 * 9);
    const int xo = pix % Wo, yo = (pix / Wo) % Ho, b = pix / ((int64_t)Wo * Ho);
    const int y = yo * stride + tap / 3 - 1, xx = xo * stride + tap % 3 - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (y >= 0 && y < H && xx >= 0 && xx < W) v = __ldg(reinterpret_cast<const uint4*>(x + (((int64_t)b * H + y) * W + xx) * C) + cg);
    reinterpret_cast<uint4*>(cols + pix * 9 * C + (int64_t)tap * C)[cg] = v;
  }
}

// dx[b][y][x][c] = sum over taps of dcols[(b,yo,xo)][tap*C + c] with yo*s + r - 1 == y, xo*s + t - 1 == x  (gather)
__global__ void col2im3x3_kernel(const __nv_bfloat16* __restrict__ dcols, __nv_bfloat16* __restrict__ dx, int B, int H, int W, int C,
                                 int Ho, int Wo, int stride) {
  const int c8 = C / 8;
  const int64_t total = (int64_t)B * H * W * c8;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int cg = idx % c8;
    const int64_t pix = idx / c8;
    const int xx = pix % W, y = (pix / W) % H, b = pix / ((int64_t)W * H);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int ny = y + 1 - tap / 3, nx = xx + 1 - tap % 3;
      if (ny < 0 || nx < 0 || ny % stride != 0 || nx % stride != 0) continue;
      const int yo = ny / stride, xo = nx / stride;
      if (yo >= Ho || xo >= Wo) continue;
      const uint4 t = __ldg(reinterpret_cast<const uint4*>(dcols + (((int64_t)b * Ho + yo) * Wo + xo) * 9 * C + (int64_t)tap * C) + cg);
      float v[8];
      unpack_bf16x8(t, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
    reinterpret_cast<uint4*>(dx + pix * C)[cg] = pack_bf16x8(acc);
  }
}

// forward (to_space=1): out[b][y*s+i][x*s+j][c] = in[(b,y,x)][(i*s+j)*C + c];  to_space=0: the inverse gather
__global__ void depth_space_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int B, int h, int w, int C, int s,
                                   int to_space) {
  const int c8 = C / 8;
  const int64_t total = (int64_t)B * h * w * s * s * c8;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int cg = idx % c8;
    const int ij = (idx / c8) % (s * s);
    const int64_t pix = idx / ((int64_t)c8 * s * s);
    const int x = pix % w, y = (pix / w) % h, b = pix / ((int64_t)w * h);
    const int64_t deep = (pix * s * s + ij) * C;                                                         // [(b,y,x)][(i,j)][c]
    const int64_t wide = ((((int64_t)b * h * s + y * s + ij / s) * w * s) + x * s + ij % s) * (int64_t)C;  // NHWC at s x resolution
    if (to_space) reinterpret_cast<uint4*>(dst + wide)[cg] = __ldg(reinterpret_cast<const uint4*>(src + deep) + cg);
    else reinterpret_cast<uint4*>(dst + deep)[cg] = __ldg(reinterpret_cast<const uint4*>(src + wide) + cg);
  }
}

// bilinear, align_corners=True (F.interpolate, dpt_block.py:251-254, dpt.py:304): src coordinate = dst * (in-1)/(out-1)
template <bool F32IN>
__global__ void bilinear_fwd_kernel(const void* __restrict__ in_, __nv_bfloat16* __restrict__ out, int B, int Hi, int Wi, int Ho,
                                    int Wo, int C) {
  const int c8 = C / 8;
  const float sy = Ho > 1 ? float(Hi - 1) / float(Ho - 1) : 0.f, sx = Wo > 1 ? float(Wi - 1) / float(Wo - 1) : 0.f;
  const int64_t total = (int64_t)B * Ho * Wo * c8;
  auto load8 = [&](int64_t pixel, int cg, float (&v)[8]) {
    if (F32IN) {
      const float4* q = reinterpret_cast<const float4*>(static_cast<const float*>(in_) + pixel * C) + 2 * cg;
      const float4 lo = __ldg(q), hi = __ldg(q + 1);
      v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
    } else {
      unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(in_) + pixel * C) + cg), v);
    }
  };
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int cg = idx % c8;
    const int64_t pix = idx / c8;
    const int X = pix % Wo, Y = (pix / Wo) % Ho, b = pix / ((int64_t)Wo * Ho);
    const float fy = Y * sy, fx = X * sx;
    const int y0 = min((int)fy, Hi - 1), x0 = min((int)fx, Wi - 1);
    const int y1 = min(y0 + 1, Hi - 1), x1 = min(x0 + 1, Wi - 1);
    const float ly = fy - y0, lx = fx - x0;
    const int64_t base = (int64_t)b * Hi * Wi;
    float a[8], bq[8], c[8], d[8], r[8];
    load8(base + (int64_t)y0 * Wi + x0, cg, a);
    load8(base + (int64_t)y0 * Wi + x1, cg, bq);
    load8(base + (int64_t)y1 * Wi + x0, cg, c);
    load8(base + (int64_t)y1 * Wi + x1, cg, d);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      r[j] = (1.f - ly) * ((1.f - lx) * a[j] + lx * bq[j]) + ly * ((1.f - lx) * c[j] + lx * d[j]);
    reinterpret_cast<uint4*>(out + pix * C)[cg] = pack_bf16x8(r);
  }
}

// backward as a gather: every input pixel scans the output pixels whose 2x2 footprint can contain it
__global__ void bilinear_bwd_kernel(const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ din, int B, int Hi, int Wi, int Ho,
                                    int Wo, int C) {
  const int c8 = C / 8;
  const float sy = Ho > 1 ? float(Hi - 1) / float(Ho - 1) : 0.f, sx = Wo > 1 ? float(Wi - 1) / float(Wo - 1) : 0.f;
  const float iy = sy > 0.f ? 1.f / sy : 0.f, ix = sx > 0.f ? 1.f / sx : 0.f;
  const int64_t total = (int64_t)B * Hi * Wi * c8;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int cg = idx % c8;
    const int64_t pix = idx / c8;
    const int x = pix % Wi, y = (pix / Wi) % Hi, b = pix / ((int64_t)Wi * Hi);
    // output rows with floor(Y*sy) in {y-1, y}: Y in ((y-1)/sy, (y+1)/sy)
    const int Ya = sy > 0.f ? max(0, (int)floorf((y - 1) * iy)) : 0, Yb = sy > 0.f ? min(Ho - 1, (int)ceilf((y + 1) * iy)) : Ho - 1;
    const int Xa = sx > 0.f ? max(0, (int)floorf((x - 1) * ix)) : 0, Xb = sx > 0.f ? min(Wo - 1, (int)ceilf((x + 1) * ix)) : Wo - 1;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int Y = Ya; Y <= Yb; ++Y) {
      const float fy = Y * sy;
      const int y0 = min((int)fy, Hi - 1), y1 = min(y0 + 1, Hi - 1);
      const float ly = fy - y0;
      float wy = 0.f;
      if (y0 == y) wy += 1.f - ly;
      if (y1 == y) wy += ly;
      if (wy == 0.f) continue;
      for (int X = Xa; X <= Xb; ++X) {
        const float fx = X * sx;
        const int x0 = min((int)fx, Wi - 1), x1 = min(x0 + 1, Wi - 1);
        const float lx = fx - x0;
        float wx = 0.f;
        if (x0 == x) wx += 1.f - lx;
        if (x1 == x) wx += lx;
        if (wx == 0.f) continue;
        float v[8];
        unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(dout + (((int64_t)b * Ho + Y) * Wo + X) * C) + cg), v);
        const float wgt = wy * wx;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += wgt * v[j];
      }
    }
    reinterpret_cast<uint4*>(din + pix * C)[cg] = pack_bf16x8(acc);
  }
}

// op 0: out = a + b ; 1: out = relu(a) ; 2: out = a * (b > 0) ; 3: out = a + b + c
__global__ void ew_kernel(int op, const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, const __nv_bfloat16* __restrict__ c,
                          __nv_bfloat16* __restrict__ out, int64_t n8) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float x[8], y[8], z[8], r[8];
    unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(a) + i), x);
    if (op == 0 || op == 2 || op == 3) unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(b) + i), y);
    if (op == 3) unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(c) + i), z);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      r[j] = op == 0 ? x[j] + y[j] : op == 1 ? fmaxf(x[j], 0.f) : op == 2 ? (y[j] > 0.f ? x[j] : 0.f) : x[j] + y[j] + z[j];
    reinterpret_cast<uint4*>(out)[i] = pack_bf16x8(r);
  }
}

}  // namespace
}  // namespace uc

using namespace uc;

extern "C" int uc_im2col3x3(const void* x, void* cols, int32_t B, int32_t H, int32_t W, int32_t C, int32_t stride, uc_stream_t st) {
  UC_REQUIRE(x && cols && B > 0 && H > 0 && W > 0 && C % 8 == 0 && (stride == 1 || stride == 2), UC_ERR_BAD_SHAPE, "uc_im2col3x3: bad arguments");
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  im2col3x3_kernel<<<grid_for((int64_t)B * Ho * Wo * 9 * (C / 8), 256), 256, 0, static_cast<cudaStream_t>(st)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(cols), B, H, W, C, Ho, Wo, stride);
  return check_launch("uc_im2col3x3");
}

extern "C" int uc_col2im3x3(const void* dcols, void* dx, int32_t B, int32_t H, int32_t W, int32_t C, int32_t stride, uc_stream_t st) {
  UC_REQUIRE(dcols && dx && B > 0 && H > 0 && W > 0 && C % 8 == 0 && (stride == 1 || stride == 2), UC_ERR_BAD_SHAPE, "uc_col2im3x3: bad arguments");
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  col2im3x3_kernel<<<grid_for((int64_t)B * H * W * (C / 8), 256), 256, 0, static_cast<cudaStream_t>(st)>>>(
      static_cast<const __nv_bfloat16*>(dcols), static_cast<__nv_bfloat16*>(dx), B, H, W, C, Ho, Wo, stride);
  return check_launch("uc_col2im3x3");
}

extern "C" int uc_depth_space(const void* src, void* dst, int32_t B, int32_t h, int32_t w, int32_t C, int32_t s, int32_t to_space, uc_stream_t st) {
  UC_REQUIRE(src && dst && B > 0 && h > 0 && w > 0 && C % 8 == 0 && s > 0, UC_ERR_BAD_SHAPE, "uc_depth_space: bad arguments");
  depth_space_kernel<<<grid_for((int64_t)B * h * w * s * s * (C / 8), 256), 256, 0, static_cast<cudaStream_t>(st)>>>(
      static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), B, h, w, C, s, to_space);
  return check_launch("uc_depth_space");
}

extern "C" int uc_bilinear_fwd(const void* in, void* out, int32_t B, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo, int32_t C, uc_stream_t st) {
  UC_REQUIRE(in && out && B > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && C % 8 == 0, UC_ERR_BAD_SHAPE, "uc_bilinear_fwd: bad arguments");
  bilinear_fwd_kernel<false><<<grid_for((int64_t)B * Ho * Wo * (C / 8), 256), 256, 0, static_cast<cudaStream_t>(st)>>>(
      in, static_cast<__nv_bfloat16*>(out), B, Hi, Wi, Ho, Wo, C);
  return check_launch("uc_bilinear_fwd");
}

extern "C" int uc_bilinear_fwd_f32in(const void* in, void* out, int32_t B, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo, int32_t C, uc_stream_t st) {
  UC_REQUIRE(in && out && B > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && C % 8 == 0, UC_ERR_BAD_SHAPE, "uc_bilinear_fwd_f32in: bad arguments");
  bilinear_fwd_kernel<true><<<grid_for((int64_t)B * Ho * Wo * (C / 8), 256), 256, 0, static_cast<cudaStream_t>(st)>>>(
      in, static_cast<__nv_bfloat16*>(out), B, Hi, Wi, Ho, Wo, C);
  return check_launch("uc_bilinear_fwd_f32in");
}

extern "C" int uc_bilinear_bwd(const void* dout, void* din, int32_t B, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo, int32_t C, uc_stream_t st) {
  UC_REQUIRE(dout && din && B > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && C % 8 == 0, UC_ERR_BAD_SHAPE, "uc_bilinear_bwd: bad arguments");
  bilinear_bwd_kernel<<<grid_for((int64_t)B * Hi * Wi * (C / 8), 256), 256, 0, static_cast<cudaStream_t>(st)>>>(
      static_cast<const __nv_bfloat16*>(dout), static_cast<__nv_bfloat16*>(din), B, Hi, Wi, Ho, Wo, C);
  return check_launch("uc_bilinear_bwd");
}

extern "C" int uc_elementwise(int32_t op, const void* a, const void* b, const void* c, void* out, int64_t n, uc_stream_t st) {
  UC_REQUIRE(a && out && n > 0 && n % 8 == 0 && op >= 0 && op <= 3, UC_ERR_BAD_SHAPE, "uc_elementwise: bad arguments");
  UC_REQUIRE((op == 1) || b, UC_ERR_BAD_SHAPE, "uc_elementwise: op %d needs b", op);
  UC_REQUIRE(op != 3 || c, UC_ERR_BAD_SHAPE, "uc_elementwise: op 3 needs c");
  ew_kernel<<<grid_for(n / 8, 256), 256, 0, static_cast<cudaStream_t>(st)>>>(op, static_cast<const __nv_bfloat16*>(a),
                                                                            static_cast<const __nv_bfloat16*>(b),
                                                                            static_cast<const __nv_bfloat16*>(c),
                                                                            static_cast<__nv_bfloat16*>(out), n / 8);
  return check_launch("uc_elementwise");
}
