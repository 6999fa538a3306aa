/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference's native 2-D RoPE CPU loop.
 *
 * Follows uniception/models/libs/croco/curope/curope.cpp:11-47 (rope_2d_cpu) and the CUDA kernel
 * curope/kernels.cu:39-80: tokens [B,N,H,D] fp32 in place, positions [B,N,2] int64 (y,x), D = 4Q.
 *   for X in {0,1}, i < Q:  theta = pos[b,n,X] * fwd / base^(i/Q)
 *   (u,v) = (t[2QX+i], t[2QX+Q+i])  ->  (u cos - v sin, v cos + u sin)
 * `fwd` is +F0 for the forward pass and -F0 for the backward pass (curope2d.py:24-28).
 * Also: the index maps of the path (bit-exact checks): patch positions and pixel-shuffle gather.
 */
#include <math.h>
#include <stdint.h>

void uc_oracle_rope2d(float* tokens, const int64_t* pos, int B, int N, int H, int D, float base, float fwd) {
  const int Q = D / 4;
  for (int b = 0; b < B; ++b)
    for (int n = 0; n < N; ++n)
      for (int h = 0; h < H; ++h) {
        float* t = tokens + (((int64_t)b * N + n) * H + h) * D;
        for (int X = 0; X < 2; ++X) {
          const float p = (float)pos[((int64_t)b * N + n) * 2 + X];
          for (int i = 0; i < Q; ++i) {
            const float ang = p * (fwd / powf(base, (float)i / (float)Q));
            const float c = cosf(ang), s = sinf(ang);
            float* u = t + 2 * Q * X + i;
            float* v = u + Q;
            const float uu = *u, vv = *v;
            *u = uu * c - vv * s;
            *v = vv * c + uu * s;
          }
        }
      }
}

/* libs/croco/patch_embed.py:25-31: cartesian_prod(arange(h), arange(w)) -> (y,x), y outer */
void uc_oracle_positions(int64_t* pos, int B, int h, int w) {
  for (int b = 0; b < B; ++b)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        int64_t* p = pos + (((int64_t)b * h + y) * w + x) * 2;
        p[0] = y; p[1] = x;
      }
}

/* prediction_heads/linear.py:81-82: out[b,c,p*h+i,p*w+j] = in[b, c*p*p + i*p + j, h, w] */
void uc_oracle_pixel_shuffle(const float* in, float* out, int B, int C, int h, int w, int p) {
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int y = 0; y < h * p; ++y)
        for (int x = 0; x < w * p; ++x) {
          const int i = y % p, j = x % p, hh = y / p, ww = x / p;
          out[(((int64_t)b * C + c) * h * p + y) * w * p + x] =
              in[(((int64_t)b * C * p * p + (c * p * p + i * p + j)) * h + hh) * w + ww];
        }
}
