// Fused attention, second generation (uc_attn_fwd / uc_attn_bwd): softmax(q k^T * scale) v and its gradients, head_dim 64.
//
// Shape of the problem on this chip, per 128 x 128 score block of one (batch, head) ("unit"):
//   tensor pipe  QK^T 256 clk + PV 256 clk;   exp pipe (MUFU.EX2, 16 / clk / SM)  16384 exps = 1024 clk;
//   issue slots  4 / clk / SM: at the 6..11 instructions per score of the earlier kernels 770..1400 clk.
// d = 64 attention is bound by the exponentials, then by instruction issue, not by the MMAs.  Measured on the way here
// (profiles/r02b_attention_design_log.txt): the first-generation kernels and a straight 128-key-tile rewrite all sat at
// ~2000 clk / unit with the XU pipe 36 % busy -- the softmax warps were starved of issue slots and of each other:
// one softmax warp per scheduler and CTA, 6..11 instructions per score (mask selects, scalar FFMA / FADD, spills), and a
// third of every CTA's life spent in set-up and tear-down.  What this file does about it, identically in all three kernels:
//   * PERSISTENT: grid = 2 CTAs per SM; a CTA walks work items (batch*head, 128-row tile) i = cta, cta + grid, ...; barriers,
//     TMEM and descriptors are set up once and the TMA / MMA warps run ahead into the next item while the softmax warps
//     finish the current item's epilogue;
//   * 8 softmax warps per CTA (16 per SM, 4 per scheduler): two threads per row, one per half of the tile's columns;
//   * ~3 instructions per score: packed f32x2 arithmetic (FFMA2 / FADD2 / FMUL2 on register pairs), MUFU.EX2, F2FP packing,
//     3-input max; tail masking only on the ragged last tile;
//   * scores never leave TMEM / registers; P / dS (bf16) go back to TMEM and feed the next MMA as its A operand (TS mode).
//
// attn_fwd3_kernel    128-key tiles (N = 128 MMAs at the full 64-clk rate), QK^T(g+1) issued as soon as the softmax warps
//                     have read S(g), online softmax with lazy rescaling decided once per tile; the two threads of a row
//                     exchange their half-row maxima / sums through shared memory (64-thread named barrier).
// Backward = two kernels that share nothing but their inputs: no fp32 dQ accumulator, no atomics, no finish pass;
// dQ, dK and dV are bit-reproducible.
// attn_bwd_dq_kernel  query-outer, 64-key tiles:  S = Q K^T, dP = dO V^T (SS) -> dS = P o (dP - delta) -> dQ += dS K (TS);
//                     epilogue: x scale, inverse 2-D RoPE, bf16 store.
// attn_bwd_dkv_kernel key-outer, 64-query sub-tiles, transposed: S^T = K Q^T, dP^T = V dO^T (SS) -> P^T, dS^T in place ->
//                     dV += P^T dO, dK += dS^T Q (TS); epilogue: dK x scale + inverse RoPE, bf16 stores.
// attn_bwd_stats_kernel writes -lse * log2(e) and -delta = -rowsum(dO o O) in a 64-row padded layout
//                     [b*H + h][tile][2][64]: one 512-byte bulk copy per sub-tile for the dK,dV kernel.
// The recomputation costs 7 instead of 5 GEMM units and two exponentials per score; in exchange the 64 MB fp32 accumulator,
// its memset, the TMA reduce traffic (8 x 64 MB per call) and the finish kernel are gone.
#include "common.cuh"
#include <stdlib.h>

namespace uc {

int make_head_map(CUtensorMap* m, const void* base, int H, int N, int B, long long ld, int box_rows);
int attn_fwd_v1(const uc_attn_fwd_params* p, cudaStream_t stream);
int attn_bwd_v1(const uc_attn_bwd_params* p, cudaStream_t stream);

// Optional cycle trace of CTA 0 of the forward kernel (bring-up aid, tools/trace_attn2.py):
// trace[(role * 64 + tile) * 16 + event] = clock64().  One pointer load per thread at kernel start when disabled.
__device__ long long* g_attn2_trace = nullptr;
#define A3_TRACE(role, j, ev)                                                          \
  do {                                                                                 \
    if (trace_on && (j) < 64) trace_buf[((role) * 64 + (j)) * 16 + (ev)] = clock64();  \
  } while (0)

namespace {

constexpr int A3_THREADS = 320;  // warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 softmax (2..5: column half 0, 6..9: half 1)
constexpr float kLazyThreshold = 8.0f;  // log2 units: the running reference max may lag the true row max by 2^8
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// packed f32x2 arithmetic (sm_100: two fp32 lanes of a 64-bit register pair per instruction)
// (d0, d1) = (a0, a1) * b + c, scalars b and c broadcast
__device__ __forceinline__ void ffma2_ss(float& d0, float& d1, float a0, float a1, float b, float c) {
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\tmov.b64 rc, {%5, %5};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
// (d0, d1) = (a0, a1) * b + (c0, c1), scalar b broadcast
__device__ __forceinline__ void ffma2_sv(float& d0, float& d1, float a0, float a1, float b, float c0, float c1) {
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\tmov.b64 rc, {%5, %6};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fmul2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
// 1-D bulk copy global -> shared, completion credited to an mbarrier (16-byte aligned, size a multiple of 16)
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// Dynamic work distribution of the persistent kernels.  Two CTAs share an SM and the warp arbiter does not treat them alike
// (measured: with a static round-robin split one CTA of each pair needed 89 us for its 7 items, the other 127 us, and the SM
// idled half-empty for the difference), so items are handed out by an atomic counter: the TMA warp fetches the next item,
// publishes it in a 4-slot shared-memory ring (mbarrier = release / acquire) and the MMA / softmax warps follow.  The producer
// can never be more than two items ahead of the slowest consumer (operand buffers), so four slots need no "empty" barriers.
// counter[0] = next item, counter[1] = CTAs that drew their sentinel; the last of them resets both for the slot's next launch.
struct ItemQueue {
  uint32_t bar0;  // 4 "full" mbarriers, 8 bytes apart
  volatile int* slots;
  int* counter;
  __device__ __forceinline__ int produce(int n, int lane) const {  // warp 0, converged
    int it = 0;
    if (lane == 0) {
      it = atomicAdd(counter, 1);
      slots[n & 3] = it;
      mbar_arrive(bar0 + 8u * (n & 3));
    }
    return __shfl_sync(0xffffffffu, it, 0);
  }
  // producer, when it has drawn its sentinel (item >= n_items): this CTA takes no more items; once every CTA has said so
  // nobody touches the counters again and the last one resets them for the slot's next launch (off the kernel's exit path)
  __device__ __forceinline__ void retire(int lane) const {
    if (lane == 0) {
      const int done = atomicAdd(counter + 1, 1);
      if (done == (int)gridDim.x - 1) {
        counter[0] = 0;
        counter[1] = 0;
        __threadfence();
      }
    }
  }
  __device__ __forceinline__ int consume(int n) const {
    mbar_wait(bar0 + 8u * (n & 3), (n >> 2) & 1);
    return slots[n & 3];
  }
  __device__ __forceinline__ void init() const {
    for (int i = 0; i < 4; ++i) mbar_init(bar0 + 8u * i, 1);
  }
};

__device__ __forceinline__ void mask32(uint32_t (&s)[32], int valid) {  // columns >= valid -> -inf
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i >= valid) s[i] = 0xff800000u;
}
__device__ __forceinline__ float max32(const uint32_t (&s)[32]) {
  float m0 = fmaxf(__uint_as_float(s[0]), __uint_as_float(s[1])), m1 = fmaxf(__uint_as_float(s[2]), __uint_as_float(s[3]));
#pragma unroll
  for (int i = 4; i < 32; i += 4) {
    m0 = fmaxf(m0, fmaxf(__uint_as_float(s[i]), __uint_as_float(s[i + 1])));  // FMNMX3
    m1 = fmaxf(m1, fmaxf(__uint_as_float(s[i + 2]), __uint_as_float(s[i + 3])));
  }
  return fmaxf(m0, m1);
}

// ================================================================================================================
// forward
// ================================================================================================================
constexpr uint32_t F3_TILE = 128 * 64 * 2;  // 16 KB: a Q, K or V tile
constexpr uint32_t F3_OFF_K = 2 * F3_TILE, F3_OFF_V = F3_OFF_K + 2 * F3_TILE, F3_OFF_X = F3_OFF_V + 2 * F3_TILE,
                   F3_OFF_BAR = F3_OFF_X + 4096;
constexpr uint32_t F3_SMEM = F3_OFF_BAR + 256 + 1024;
constexpr uint32_t F3_TM_S = 0, F3_TM_P = 128, F3_TM_O = 192, F3_TM_COLS = 256;

struct FwdArgs {
  __nv_bfloat16* o;
  float* lse;
  int B, H, Nq, Nk;
  long long ldo;
  float scale_log2;  // scale * log2(e)
  float scale;
};

// (d0, d1) = (a0, a1) * (b0, b1) + c, scalar c broadcast
__device__ __forceinline__ void ffma2_vs(float& d0, float& d1, float a0, float a1, float b0, float b1, float c) {
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %6};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c));
}
// 2^x for a register pair on the FMA / ALU pipes (Cody-Waite split + degree-3 minimax polynomial, max rel. error 7.5e-5, far
// below the bf16 rounding of P): x = n + f, n = round(x) via the 1.5 * 2^23 magic constant, 2^f from the polynomial, 2^n by
// adding n to the exponent field ((t_bits << 23) keeps nothing but n).  Inputs are clamped to >= -125.
__device__ __forceinline__ void ex2_poly2(float x0, float x1, float& p0, float& p1) {
  x0 = fmaxf(x0, -125.0f);
  x1 = fmaxf(x1, -125.0f);
  float t0, t1, n0, n1, f0, f1;
  fadd2(t0, t1, x0, x1, 12582912.0f, 12582912.0f);
  fadd2(n0, n1, t0, t1, -12582912.0f, -12582912.0f);
  ffma2_ss(f0, f1, n0, n1, -1.0f, 0.0f);
  fadd2(f0, f1, f0, f1, x0, x1);
  float q0, q1;
  ffma2_ss(q0, q1, f0, f1, 0.055170949548482895f, 0.2426096349954605f);
  ffma2_vs(q0, q1, q0, q1, f0, f1, 0.6932609677314758f);
  ffma2_vs(q0, q1, q0, q1, f0, f1, 0.9999281764030457f);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}
// exps of 32 score columns in registers -> 16 packed bf16 P columns; (l0, l1) += the unrounded exps.  The first NPOLY of the
// 16 column pairs take the polynomial path (FMA pipe) instead of MUFU.EX2 (XU pipe).
template <int NPOLY>
__device__ __forceinline__ void exp32(const uint32_t (&s)[32], float sl2, float neg_m, uint32_t (&pk)[16], float& l0, float& l1) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float x0, x1, p0, p1;
    ffma2_ss(x0, x1, __uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]), sl2, neg_m);
    if (i < NPOLY) {
      ex2_poly2(x0, x1, p0, p1);
    } else {
      p0 = ex2(x0);
      p1 = ex2(x1);
    }
    fadd2(l0, l1, l0, l1, p0, p1);
    pk[i] = pack_bf16(p0, p1);
  }
}

template <int NPOLY>
__global__ void __launch_bounds__(A3_THREADS, 2)
attn_fwd3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const FwdArgs a, int* work) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  auto sQ = [&](int st) { return smem_base + st * F3_TILE; };
  auto sK = [&](int st) { return smem_base + F3_OFF_K + st * F3_TILE; };
  auto sV = [&](int st) { return smem_base + F3_OFF_V + st * F3_TILE; };
  float* xch = reinterpret_cast<float*>(smem_gen + F3_OFF_X);  // half-row maxima [parity 2][half 2][row 128]; +512: half-row sums
  const uint32_t bar = smem_base + F3_OFF_BAR;
  auto q_full = [&](int st) { return bar + 8u * st; };
  auto q_empty = [&](int st) { return bar + 8u * (2 + st); };
  auto k_full = [&](int st) { return bar + 8u * (4 + st); };
  auto k_empty = [&](int st) { return bar + 8u * (6 + st); };
  auto v_full = [&](int st) { return bar + 8u * (8 + st); };
  auto v_empty = [&](int st) { return bar + 8u * (10 + st); };
  const uint32_t s_full = bar + 8u * 12, s_free = bar + 8u * 13, p_ready = bar + 8u * 14, o_done = bar + 8u * 15,
                 tmem_slot = bar + 8u * 16;
  const ItemQueue queue{bar + 8u * 18, reinterpret_cast<volatile int*>(smem_gen + F3_OFF_BAR + 8 * 22), work};

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = (a.Nk + 127) / 128;   // key tiles per item
  const int nq = (a.Nq + 127) / 128;  // query tiles per (batch, head)
  const int n_items = nq * a.B * a.H;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int s = 0; s < 2; ++s) {
      mbar_init(q_full(s), 1); mbar_init(q_empty(s), 1);
      mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1);
      mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 8);
    mbar_init(p_ready, 8);
    mbar_init(o_done, 1);
    queue.init();
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, F3_TM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_launch_dependents();  // after the TMEM allocation (common.cuh: PDL rules)
  pdl_wait();
  long long* const trace_buf = g_attn2_trace;
  const bool trace_on = trace_buf != nullptr && blockIdx.x == 0 && lane == 0 && (warp == 1 || warp == 2 || warp == 6);
  const int trole = warp == 1 ? 0 : (warp == 2 ? 1 : 2);
  if (trace_buf != nullptr && threadIdx.x == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    trace_buf[3 * 64 * 16 + blockIdx.x * 4 + 0] = (long long)globaltimer_ns();
    trace_buf[3 * 64 * 16 + blockIdx.x * 4 + 2] = smid;
  }

  if (warp == 0) {
    // ---- TMA producer: runs up to two tiles / one item ahead ----
    int g = 0;  // tile counter of this CTA
    for (int n = 0;; ++n) {
      const int it = queue.produce(n, lane);
      if (it >= n_items) {
        queue.retire(lane);
        break;
      }
      const int bh = it / nq, qt = it % nq;
      const int b = bh / a.H, h = bh % a.H;
      mbar_wait(q_empty(n & 1), ((n >> 1) & 1) ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full(n & 1), F3_TILE);
        tma_load_3d(sQ(n & 1), &tmQ, q_full(n & 1), h * 64, qt * 128, b);
      }
      __syncwarp();
      for (int j = 0; j < T; ++j, ++g) {
        mbar_wait(k_empty(g & 1), ((g >> 1) & 1) ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(k_full(g & 1), F3_TILE);
          tma_load_3d(sK(g & 1), &tmK, k_full(g & 1), h * 64, j * 128, b);
        }
        __syncwarp();
        mbar_wait(v_empty(g & 1), ((g >> 1) & 1) ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(v_full(g & 1), F3_TILE);
          tma_load_3d(sV(g & 1), &tmV, v_full(g & 1), h * 64, j * 128, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: flat loop over this CTA's tiles; QK^T of the NEXT tile (possibly of the next item) goes first ----
    const uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);
    const uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);
    auto issue_qk = [&](int n, int j, int g) {  // S = Q(item n) K(tile g)^T
      if (j == 0) mbar_wait(q_full(n & 1), (n >> 1) & 1);
      mbar_wait(k_full(g & 1), (g >> 1) & 1);
      tc_fence_after();
      const uint64_t qd = umma_desc_kmajor(sQ(n & 1)), kd = umma_desc_kmajor(sK(g & 1));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem_base + F3_TM_S, qd + uint64_t(k * 2), kd + uint64_t(k * 2), idesc_qk, k > 0);
        umma_commit(s_full);
        umma_commit(k_empty(g & 1));
        if (j == T - 1) umma_commit(q_empty(n & 1));
      }
      __syncwarp();
    };
    int g = 0;
    if (queue.consume(0) < n_items) {
      issue_qk(0, 0, 0);
      for (int n = 0;; ++n) {
        bool more = true;
        for (int j = 0; j < T; ++j, ++g) {
          A3_TRACE(0, g, 0);
          // look-ahead: QK^T of the next tile -- of this item, or the first one of the next item
          if (j + 1 < T) {
            mbar_wait(s_free, g & 1);  // every softmax warp has finished reading S(g)
            A3_TRACE(0, g, 1);
            issue_qk(n, j + 1, g + 1);
          } else {
            more = queue.consume(n + 1) < n_items;
            if (more) {
              mbar_wait(s_free, g & 1);
              issue_qk(n + 1, 0, g + 1);
            }
          }
          A3_TRACE(0, g, 2);
          mbar_wait(p_ready, g & 1);
          A3_TRACE(0, g, 3);
          mbar_wait(v_full(g & 1), (g >> 1) & 1);
          tc_fence_after();
          const uint64_t vd = umma_desc_mnmajor(sV(g & 1), 8192);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_ts(tmem_base + F3_TM_O, tmem_base + F3_TM_P + k * 8, vd + uint64_t(k * 128), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
            umma_commit(v_empty(g & 1));
            umma_commit(o_done);
          }
          __syncwarp();
          A3_TRACE(0, g, 4);
        }
        if (!more) break;
      }
    }
  } else {
    // ---- softmax: two threads per query row (half = 64-column half of the key tile) ----
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;
    const int r = quad * 32 + lane;  // row inside the tile
    const uint32_t lane_addr = uint32_t(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + F3_TM_S + 64 * half, tP = tmem_base + lane_addr + F3_TM_P + 32 * half,
                   tO = tmem_base + lane_addr + F3_TM_O + 32 * half;
    float* xmax = xch;        // [2 parity][2 half][128] half-row maxima
    float* xsum = xch + 512;  // [2 half][128] half-row sums
    int g = 0;
    for (int n = 0;; ++n) {
      const int it = queue.consume(n);
      if (it >= n_items) break;
      const int bh = it / nq, qt = it % nq;
      const int b = bh / a.H, h = bh % a.H;
      float m_run = -INFINITY, l0 = 0.f, l1 = 0.f;  // l0 + l1: this thread's half of the row sum
      for (int j = 0; j < T; ++j, ++g) {
        A3_TRACE(trole, g, 0);
        mbar_wait(s_full, g & 1);
        tc_fence_after();
        A3_TRACE(trole, g, 1);
        const int valid = a.Nk - j * 128 - 64 * half;  // score columns of this half that are real keys (< 64: ragged tail)
        uint32_t s[32];
        float m_tile;
        {  // pass 1: row maximum of this half, then the partner thread's half through shared memory
          tmem_ld32(tS, s);
          tmem_ld_wait();
          if (valid < 32) mask32(s, valid);
          m_tile = max32(s);
          tmem_ld32(tS + 32, s);
          tmem_ld_wait();
          if (valid < 64) mask32(s, valid - 32);
          m_tile = fmaxf(m_tile, max32(s));
          xmax[((g & 1) * 2 + half) * 128 + r] = m_tile;
          A3_TRACE(trole, g, 2);
          bar_sync_named(1 + quad, 64);
          A3_TRACE(trole, g, 3);
          m_tile = fmaxf(m_tile, xmax[((g & 1) * 2 + (half ^ 1)) * 128 + r]);
        }
        // lazy rescaling (both threads of a row see the same numbers and decide alike)
        const bool grow = (m_tile - m_run) * a.scale_log2 > kLazyThreshold;  // true on an item's first tile (m_run = -inf)
        const float m_new = grow ? m_tile : m_run;
        const float alpha = grow ? ex2((m_run - m_new) * a.scale_log2) : 1.0f;
        const float neg_m = -m_new * a.scale_log2;
        if (grow) {
          l0 *= alpha;
          l1 *= alpha;
        }
        m_run = m_new;
        uint32_t pk[16];
        // pass 2: exponentials
        tmem_ld32(tS, s);
        tmem_ld_wait();
        if (valid < 32) mask32(s, valid);
        exp32<NPOLY>(s, a.scale_log2, neg_m, pk, l0, l1);
        A3_TRACE(trole, g, 4);
        if (g > 0) {  // PV(g-1) has retired: the P region may be overwritten and O is stable
          mbar_wait(o_done, (g - 1) & 1);
          tc_fence_after();
        }
        A3_TRACE(trole, g, 5);
        tmem_st16(tP, pk);
        tmem_ld32(tS + 32, s);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);  // this warp is done with S(g)
        A3_TRACE(trole, g, 6);
        if (valid < 64) mask32(s, valid - 32);
        exp32<NPOLY>(s, a.scale_log2, neg_m, pk, l0, l1);
        tmem_st16(tP + 16, pk);
        A3_TRACE(trole, g, 7);
        if (j > 0 && __any_sync(0xffffffffu, grow)) {  // rescale this thread's 32 output columns
          uint32_t o[32];
          tmem_ld32(tO, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st32(tO, o);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready);
        A3_TRACE(trole, g, 8);
      }
      // ---- item epilogue: O / l -> global (each thread its 32 columns), LSE ----
      const float l_half = l0 + l1;
      xsum[half * 128 + r] = l_half;
      A3_TRACE(trole, g - 1, 9);
      mbar_wait(o_done, (g - 1) & 1);
      tc_fence_after();
      A3_TRACE(trole, g - 1, 10);
      bar_sync_named(1 + quad, 64);
      A3_TRACE(trole, g - 1, 11);
      const float l_row = l_half + xsum[(half ^ 1) * 128 + r];
      const float inv_l = 1.0f / l_row;
      const int row = qt * 128 + r;
      const bool row_ok = row < a.Nq;
      {
        uint32_t o[32];
        tmem_ld32(tO, o);
        tmem_ld_wait();
        if (row_ok) {
          uint4* dst = reinterpret_cast<uint4*>(a.o + ((long long)b * a.Nq + row) * a.ldo + h * 64 + 32 * half);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 t;
            t.x = pack_bf16(__uint_as_float(o[8 * i + 0]) * inv_l, __uint_as_float(o[8 * i + 1]) * inv_l);
            t.y = pack_bf16(__uint_as_float(o[8 * i + 2]) * inv_l, __uint_as_float(o[8 * i + 3]) * inv_l);
            t.z = pack_bf16(__uint_as_float(o[8 * i + 4]) * inv_l, __uint_as_float(o[8 * i + 5]) * inv_l);
            t.w = pack_bf16(__uint_as_float(o[8 * i + 6]) * inv_l, __uint_as_float(o[8 * i + 7]) * inv_l);
            dst[i] = t;
          }
        }
      }
      A3_TRACE(trole, g - 1, 12);
      if (half == 0 && row_ok && a.lse) a.lse[((long long)b * a.H + h) * a.Nq + row] = m_run * a.scale + logf(l_row);
      A3_TRACE(trole, g - 1, 13);
      bar_sync_named(1 + quad, 64);  // the partner has read this item's row sum before the next item's is written
      A3_TRACE(trole, g - 1, 14);
    }
  }

  if (trace_buf != nullptr && threadIdx.x == 64) trace_buf[3 * 64 * 16 + blockIdx.x * 4 + 1] = (long long)globaltimer_ns();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, F3_TM_COLS);
  }
}

// ================================================================================================================
// forward, ping-pong: ONE CTA per SM works on TWO 128-row query tiles of the same (batch, head) at once
// ================================================================================================================
// attn_fwd3_kernel leaves the exp pipe ~40 % idle (profiles/r02b_attention_design_log.txt): its softmax warps read every score
// tile twice from TMEM (row maximum, then exponentials; four dependent TMEM round trips per tile), the next QK^T cannot start
// before the second read, and its two CTAs per SM fetch K and V separately.  Here:
//   * a softmax thread loads its 64 scores of a tile into registers ONCE and releases the S buffer at that point, so
//     QK^T(j+1) of its group runs underneath the whole softmax of tile j;
//   * one CTA per SM hosts two such groups (A and B: 8 warps each, two threads per query row) on two query tiles of the same
//     (batch, head); K / V tiles are fetched once and serve both; one MMA warp issues for both groups in a fixed order
//     QK_A(j+1) QK_B(j+1) PV_A(j) PV_B(j), each MMA gated by the barrier of the group it serves.
// TMEM (all 512 columns): S_A | S_B (128 fp32 each) | P_A | P_B (128 bf16 = 64 columns each) | O_A | O_B (64 each).
constexpr int F4_THREADS = 64 + 16 * 32;  // warp 0 TMA, warp 1 MMA, warps 2..9 group A, 10..17 group B
constexpr int F4_KST = 3;                 // K and V stages
constexpr uint32_t F4_OFF_K = 4 * F3_TILE, F4_OFF_V = F4_OFF_K + F4_KST * F3_TILE, F4_OFF_X = F4_OFF_V + F4_KST * F3_TILE,
                   F4_OFF_BAR = F4_OFF_X + 8192;
constexpr uint32_t F4_SMEM = F4_OFF_BAR + 512 + 1024;
constexpr uint32_t F4_TM_COLS = 512;

template <int NPOLY>
__global__ void __launch_bounds__(F4_THREADS, 1)
attn_fwd4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const FwdArgs a, int* work) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  auto sQ = [&](int st, int grp) { return smem_base + (2 * st + grp) * F3_TILE; };
  auto sK = [&](int st) { return smem_base + F4_OFF_K + st * F3_TILE; };
  auto sV = [&](int st) { return smem_base + F4_OFF_V + st * F3_TILE; };
  float* xch = reinterpret_cast<float*>(smem_gen + F4_OFF_X);  // per group 768 floats: maxima [parity 2][half 2][128], sums [half 2][128]
  const uint32_t bar = smem_base + F4_OFF_BAR;
  auto q_full = [&](int st) { return bar + 8u * st; };
  auto q_empty = [&](int st) { return bar + 8u * (2 + st); };
  auto k_full = [&](int st) { return bar + 8u * (4 + st); };
  auto k_empty = [&](int st) { return bar + 8u * (8 + st); };
  auto v_full = [&](int st) { return bar + 8u * (12 + st); };
  auto v_empty = [&](int st) { return bar + 8u * (16 + st); };
  auto s_full = [&](int grp) { return bar + 8u * (20 + grp); };
  auto p_ready = [&](int grp) { return bar + 8u * (22 + grp); };
  auto o_done = [&](int grp) { return bar + 8u * (24 + grp); };
  auto s_free = [&](int grp) { return bar + 8u * (26 + grp); };
  const uint32_t tmem_slot = bar + 8u * 28;
  const ItemQueue queue{bar + 8u * 30, reinterpret_cast<volatile int*>(smem_gen + F4_OFF_BAR + 8 * 34), work};
  auto tm_s = [](int grp) { return uint32_t(128 * grp); };
  auto tm_p = [](int grp) { return uint32_t(256 + 64 * grp); };
  auto tm_o = [](int grp) { return uint32_t(384 + 64 * grp); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = (a.Nk + 127) / 128;    // key tiles per item
  const int nq = (a.Nq + 127) / 128;   // query tiles per (batch, head)
  const int nq2 = (nq + 1) / 2;        // items (pairs of query tiles) per (batch, head)
  const int n_items = nq2 * a.B * a.H;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int s = 0; s < 2; ++s) { mbar_init(q_full(s), 1); mbar_init(q_empty(s), 1); }
    for (int s = 0; s < F4_KST; ++s) {
      mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1);
      mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1);
    }
    for (int grp = 0; grp < 2; ++grp) {
      mbar_init(s_full(grp), 1);
      mbar_init(p_ready(grp), 8);
      mbar_init(o_done(grp), 1);
      mbar_init(s_free(grp), 8);
    }
    queue.init();
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, F4_TM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_launch_dependents();  // after the TMEM allocation (common.cuh: PDL rules)
  pdl_wait();

  if (warp == 0) {
    // ---- TMA producer ----
    int g = 0;  // tile counter of this CTA
    for (int n = 0;; ++n) {
      const int it = queue.produce(n, lane);
      if (it >= n_items) {
        queue.retire(lane);
        break;
      }
      const int bh = it / nq2, qt2 = it % nq2;
      const int b = bh / a.H, h = bh % a.H;
      const int qa = 2 * qt2, qb = min(2 * qt2 + 1, nq - 1);  // odd tile count: group B repeats the last tile and stores nothing
      mbar_wait(q_empty(n & 1), ((n >> 1) & 1) ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full(n & 1), 2 * F3_TILE);
        tma_load_3d(sQ(n & 1, 0), &tmQ, q_full(n & 1), h * 64, qa * 128, b);
        tma_load_3d(sQ(n & 1, 1), &tmQ, q_full(n & 1), h * 64, qb * 128, b);
      }
      __syncwarp();
      for (int j = 0; j < T; ++j, ++g) {
        const int st = g % F4_KST;
        const uint32_t ph = ((g / F4_KST) & 1) ^ 1u;
        mbar_wait(k_empty(st), ph);
        if (elect_one()) {
          mbar_arrive_expect_tx(k_full(st), F3_TILE);
          tma_load_3d(sK(st), &tmK, k_full(st), h * 64, j * 128, b);
        }
        __syncwarp();
        mbar_wait(v_empty(st), ph);
        if (elect_one()) {
          mbar_arrive_expect_tx(v_full(st), F3_TILE);
          tma_load_3d(sV(st), &tmV, v_full(st), h * 64, j * 128, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: the issue order below is the ping-pong schedule ----
    const uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);
    const uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);
    auto issue_qk = [&](int grp, int n, int j, int g) {  // S_grp = Q_grp(item n) K(tile g)^T
      const int st = g % F4_KST;
      if (grp == 0) {  // group B's MMAs follow group A's in program order: same operands, already waited for
        if (j == 0) mbar_wait(q_full(n & 1), (n >> 1) & 1);
        mbar_wait(k_full(st), (g / F4_KST) & 1);
      }
      tc_fence_after();
      const uint64_t qd = umma_desc_kmajor(sQ(n & 1, grp)), kd = umma_desc_kmajor(sK(st));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem_base + tm_s(grp), qd + uint64_t(k * 2), kd + uint64_t(k * 2), idesc_qk, k > 0);
        umma_commit(s_full(grp));
        if (grp == 1) {
          umma_commit(k_empty(st));
          if (j == T - 1) umma_commit(q_empty(n & 1));
        }
      }
      __syncwarp();
    };
    auto issue_pv = [&](int grp, int j, int g) {  // O_grp += P_grp(g) V(g)
      const int st = g % F4_KST;
      mbar_wait(p_ready(grp), g & 1);
      if (grp == 0) mbar_wait(v_full(st), (g / F4_KST) & 1);
      tc_fence_after();
      const uint64_t vd = umma_desc_mnmajor(sV(st), 8192);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ts(tmem_base + tm_o(grp), tmem_base + tm_p(grp) + k * 8, vd + uint64_t(k * 128), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
        if (grp == 1) umma_commit(v_empty(st));
        umma_commit(o_done(grp));
      }
      __syncwarp();
    };
    int g = 0;
    if (queue.consume(0) < n_items) {
      issue_qk(0, 0, 0, 0);
      issue_qk(1, 0, 0, 0);
      for (int n = 0;; ++n) {
        bool more = true;
        for (int j = 0; j < T; ++j, ++g) {
          const bool last = j == T - 1;
          if (last) more = queue.consume(n + 1) < n_items;
          if (!last || more) {  // look-ahead: QK^T of the next tile as soon as the group has its scores in registers
#pragma unroll
            for (int grp = 0; grp < 2; ++grp) {
              mbar_wait(s_free(grp), g & 1);
              if (!last) issue_qk(grp, n, j + 1, g + 1);
              else issue_qk(grp, n + 1, 0, g + 1);
            }
          }
          issue_pv(0, j, g);
          issue_pv(1, j, g);
        }
        if (!more) break;
      }
    }
  } else {
    // ---- softmax: group = query tile, two threads per query row (half = 64-column half of the key tile) ----
    const int sw = warp - 2;
    const int grp = sw >> 3;
    const int half = (sw >> 2) & 1;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int r = quad * 32 + lane;  // row inside the tile
    const int bar_id = 1 + grp * 4 + quad;  // named barrier of the 64 threads that share 32 rows
    const uint32_t lane_addr = uint32_t(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_addr + tm_s(grp) + 64 * half, tP = tmem_base + lane_addr + tm_p(grp) + 32 * half,
                   tO = tmem_base + lane_addr + tm_o(grp) + 32 * half;
    float* xmax = xch + grp * 768;  // [2 parity][2 half][128] half-row maxima
    float* xsum = xmax + 512;       // [2 half][128] half-row sums
    int g = 0;
    for (int n = 0;; ++n) {
      const int it = queue.consume(n);
      if (it >= n_items) break;
      const int bh = it / nq2, qt = 2 * (it % nq2) + grp;
      const int b = bh / a.H, h = bh % a.H;
      float m_run = -INFINITY, l0 = 0.f, l1 = 0.f;  // l0 + l1: this thread's half of the row sum
      for (int j = 0; j < T; ++j, ++g) {
        mbar_wait(s_full(grp), g & 1);
        tc_fence_after();
        const int valid = a.Nk - j * 128 - 64 * half;  // score columns of this half that are real keys (< 64: ragged tail)
        uint32_t s0[32], s1[32];
        tmem_ld32(tS, s0);
        tmem_ld32(tS + 32, s1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free(grp));  // the scores are in registers: QK^T(g+1) may overwrite S
        if (valid < 64) {
          mask32(s0, valid);
          mask32(s1, valid - 32);
        }
        float m_tile = fmaxf(max32(s0), max32(s1));
        xmax[((g & 1) * 2 + half) * 128 + r] = m_tile;
        bar_sync_named(bar_id, 64);
        m_tile = fmaxf(m_tile, xmax[((g & 1) * 2 + (half ^ 1)) * 128 + r]);
        // lazy rescaling (both threads of a row see the same numbers and decide alike)
        const bool grow = (m_tile - m_run) * a.scale_log2 > kLazyThreshold;  // true on an item's first tile (m_run = -inf)
        const float m_new = grow ? m_tile : m_run;
        const float alpha = grow ? ex2((m_run - m_new) * a.scale_log2) : 1.0f;
        const float neg_m = -m_new * a.scale_log2;
        if (grow) {
          l0 *= alpha;
          l1 *= alpha;
        }
        m_run = m_new;
        uint32_t pk[16];
        exp32<NPOLY>(s0, a.scale_log2, neg_m, pk, l0, l1);
        if (g > 0) {  // PV(g-1) of this group has retired: its P region may be overwritten and O is stable
          mbar_wait(o_done(grp), (g - 1) & 1);
          tc_fence_after();
        }
        tmem_st16(tP, pk);
        exp32<NPOLY>(s1, a.scale_log2, neg_m, pk, l0, l1);
        tmem_st16(tP + 16, pk);
        if (j > 0 && __any_sync(0xffffffffu, grow)) {  // rescale this thread's 32 output columns
          uint32_t o[32];
          tmem_ld32(tO, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st32(tO, o);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready(grp));
      }
      // ---- item epilogue: O / l -> global (each thread its 32 columns), LSE ----
      const float l_half = l0 + l1;
      xsum[half * 128 + r] = l_half;
      mbar_wait(o_done(grp), (g - 1) & 1);
      tc_fence_after();
      bar_sync_named(bar_id, 64);
      const float l_row = l_half + xsum[(half ^ 1) * 128 + r];
      const float inv_l = 1.0f / l_row;
      const int row = qt * 128 + r;
      const bool row_ok = row < a.Nq;
      {
        uint32_t o[32];
        tmem_ld32(tO, o);
        tmem_ld_wait();
        if (row_ok) {
          uint4* dst = reinterpret_cast<uint4*>(a.o + ((long long)b * a.Nq + row) * a.ldo + h * 64 + 32 * half);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 t;
            t.x = pack_bf16(__uint_as_float(o[8 * i + 0]) * inv_l, __uint_as_float(o[8 * i + 1]) * inv_l);
            t.y = pack_bf16(__uint_as_float(o[8 * i + 2]) * inv_l, __uint_as_float(o[8 * i + 3]) * inv_l);
            t.z = pack_bf16(__uint_as_float(o[8 * i + 4]) * inv_l, __uint_as_float(o[8 * i + 5]) * inv_l);
            t.w = pack_bf16(__uint_as_float(o[8 * i + 6]) * inv_l, __uint_as_float(o[8 * i + 7]) * inv_l);
            dst[i] = t;
          }
        }
      }
      if (half == 0 && row_ok && a.lse) a.lse[((long long)b * a.H + h) * a.Nq + row] = m_run * a.scale + logf(l_row);
      bar_sync_named(bar_id, 64);  // the partner has read this item's row sum before the next item's is written
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, F4_TM_COLS);
  }
}

// ================================================================================================================
// backward: statistics
// ================================================================================================================
// stats[bh][tile][0][i] = -lse[q] * log2(e), stats[bh][tile][1][i] = -sum_d dO[q,d] * O[q,d], q = tile*64 + i; rows >= Nq hold
// (-inf, 0) so that padded queries contribute exp2(-inf) = 0.  8 lanes per (query, head), 8 elements per lane.
__global__ void attn_bwd_stats_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                                      const float* __restrict__ lse, float* __restrict__ stats, int B, int H, int N, int Npad,
                                      long long ldo, long long lddo) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = (long long)B * H * Npad * 8;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int sub = idx & 7;
    const long long t = idx >> 3;
    const int q = t % Npad;
    const long long bh = t / Npad;
    const int h = bh % H;
    const long long b = bh / H;
    float s = 0.f;
    if (q < N) {
      const long long tok = b * N + q;
      const uint4 x = __ldg(reinterpret_cast<const uint4*>(o + tok * ldo + h * 64 + sub * 8));
      const uint4 g = __ldg(reinterpret_cast<const uint4*>(d_o + tok * lddo + h * 64 + sub * 8));
      s = bf16_lo(x.x) * bf16_lo(g.x) + bf16_hi(x.x) * bf16_hi(g.x) + bf16_lo(x.y) * bf16_lo(g.y) + bf16_hi(x.y) * bf16_hi(g.y) +
          bf16_lo(x.z) * bf16_lo(g.z) + bf16_hi(x.z) * bf16_hi(g.z) + bf16_lo(x.w) * bf16_lo(g.w) + bf16_hi(x.w) * bf16_hi(g.w);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (sub == 0) {
      float* dst = stats + (bh * (Npad / 64) + q / 64) * 128 + (q & 63);
      dst[0] = q < N ? -lse[bh * N + q] * kLog2e : -INFINITY;
      dst[64] = -s;
    }
  }
}

struct BwdArgs {
  const float* stats;
  __nv_bfloat16* dq;
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
  int B, H, Nq, Nk, Nq_pad;
  long long lddq, lddk, lddv;
  float scale, scale_log2;
  const int* q_positions;
  const int* k_positions;
  const float* rope_table;
};

// 32 fp32 accumulator columns (one 32-wide half of a head) of this thread's row -> x scale -> optional inverse 2-D RoPE
// (pairs (i, i+16) inside the half; `pos` = the token's position along the axis this half encodes) -> 32 bf16 in global memory
__device__ __forceinline__ void store_grad_half(uint32_t taddr, __nv_bfloat16* dst_half, bool ok, float scale, const int* pos,
                                                const float* table) {
  uint32_t raw[32];
  tmem_ld32(taddr, raw);
  tmem_ld_wait();
  if (ok) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]) * scale;
    if (pos) {
      const float* tr = table + (long long)(*pos) * 32;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float cs = tr[2 * i], sn = -tr[2 * i + 1];  // inverse rotation
        const float u = v[i], w = v[i + 16];
        v[i] = u * cs - w * sn;
        v[i + 16] = w * cs + u * sn;
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(dst_half);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4 t;
      t.x = pack_bf16(v[8 * g], v[8 * g + 1]); t.y = pack_bf16(v[8 * g + 2], v[8 * g + 3]);
      t.z = pack_bf16(v[8 * g + 4], v[8 * g + 5]); t.w = pack_bf16(v[8 * g + 6], v[8 * g + 7]);
      dst[g] = t;
    }
  }
  __syncwarp();
}

// ================================================================================================================
// backward: dQ  (query-outer; 64-key tiles; thread <-> (query row, 32-key half of the tile))
// ================================================================================================================
constexpr uint32_t DQ_QT = 128 * 64 * 2, DQ_KT = 64 * 64 * 2;  // 16 KB query-side tiles, 8 KB key-side tiles
constexpr uint32_t DQ_OFF_DO = DQ_QT, DQ_OFF_K = 2 * DQ_QT, DQ_OFF_V = DQ_OFF_K + 3 * DQ_KT, DQ_OFF_BAR = DQ_OFF_V + 2 * DQ_KT;
constexpr uint32_t DQ_SMEM = DQ_OFF_BAR + 256 + 1024;
constexpr uint32_t DQ_TM_S = 0, DQ_TM_DP = 64, DQ_TM_DS = 128, DQ_TM_DQ = 192, DQ_TM_COLS = 256;

__global__ void __launch_bounds__(A3_THREADS, 2)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO, const BwdArgs a, int* work) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sQ = smem_base, sdO = smem_base + DQ_OFF_DO;
  auto sK = [&](int st) { return smem_base + DQ_OFF_K + st * DQ_KT; };
  auto sV = [&](int st) { return smem_base + DQ_OFF_V + st * DQ_KT; };
  const uint32_t bar = smem_base + DQ_OFF_BAR;
  const uint32_t qdo_full = bar, qdo_empty = bar + 8u;
  auto k_full = [&](int st) { return bar + 8u * (2 + st); };
  auto k_empty = [&](int st) { return bar + 8u * (5 + st); };
  auto v_full = [&](int st) { return bar + 8u * (8 + st); };
  auto v_empty = [&](int st) { return bar + 8u * (10 + st); };
  const uint32_t sdp_full = bar + 8u * 12, s_free = bar + 8u * 13, ds_ready = bar + 8u * 14, dq_done = bar + 8u * 15,
                 tmem_slot = bar + 8u * 16;
  const ItemQueue queue{bar + 8u * 18, reinterpret_cast<volatile int*>(smem_gen + DQ_OFF_BAR + 8 * 22), work};

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = (a.Nk + 63) / 64;     // key tiles per item
  const int nq = (a.Nq + 127) / 128;  // query tiles per (batch, head)
  const int n_items = nq * a.B * a.H;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    mbar_init(qdo_full, 1);
    mbar_init(qdo_empty, 1);
    for (int s = 0; s < 3; ++s) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1); }
    mbar_init(sdp_full, 1);
    mbar_init(s_free, 8);
    mbar_init(ds_ready, 8);
    mbar_init(dq_done, 1);
    queue.init();
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, DQ_TM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    int g = 0;
    for (int n = 0;; ++n) {
      const int it = queue.produce(n, lane);
      if (it >= n_items) {
        queue.retire(lane);
        break;
      }
      const int bh = it / nq, qt = it % nq;
      const int b = bh / a.H, h = bh % a.H;
      mbar_wait(qdo_empty, (n & 1) ^ 1u);  // S / dP MMAs of the previous item's last tile have read Q and dO
      if (elect_one()) {
        mbar_arrive_expect_tx(qdo_full, 2 * DQ_QT);
        tma_load_3d(sQ, &tmQ, qdo_full, h * 64, qt * 128, b);
        tma_load_3d(sdO, &tmdO, qdo_full, h * 64, qt * 128, b);
      }
      __syncwarp();
      for (int j = 0; j < T; ++j, ++g) {
        const int ks = g % 3, vs = g & 1;
        mbar_wait(k_empty(ks), ((g / 3) & 1) ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(k_full(ks), DQ_KT);
          tma_load_3d(sK(ks), &tmK, k_full(ks), h * 64, j * 64, b);
        }
        __syncwarp();
        mbar_wait(v_empty(vs), ((g >> 1) & 1) ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(v_full(vs), DQ_KT);
          tma_load_3d(sV(vs), &tmV, v_full(vs), h * 64, j * 64, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    const uint32_t id_ss = umma_idesc_bf16(128, 64, 0, 0);  // [128 queries] x [64 keys], both K-major (contraction over d)
    const uint32_t id_ts = umma_idesc_bf16(128, 64, 0, 1);  // A = dS in TMEM, B = K tile read MN-major (contraction over keys)
    const uint64_t qd = umma_desc_kmajor(sQ), dod = umma_desc_kmajor(sdO);
    auto issue_dq = [&](int g) {  // dQ (+)= dS(g) K(g)
      const int ks = g % 3;
      mbar_wait(ds_ready, g & 1);
      tc_fence_after();
      const uint64_t k_mn = umma_desc_mnmajor(sK(ks), 8192);
      const uint32_t acc = (g % T) > 0 ? 1u : 0u;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ts(tmem_base + DQ_TM_DQ, tmem_base + DQ_TM_DS + k * 8, k_mn + uint64_t(k * 128), id_ts, (acc || k > 0) ? 1u : 0u);
        umma_commit(k_empty(ks));
        umma_commit(dq_done);
      }
      __syncwarp();
    };
    int g = 0;
    for (int n = 0;; ++n) {
      if (queue.consume(n) >= n_items) break;
      for (int j = 0; j < T; ++j, ++g) {
        const int ks = g % 3, vs = g & 1;
        if (j == 0) mbar_wait(qdo_full, n & 1);
        mbar_wait(k_full(ks), (g / 3) & 1);
        mbar_wait(v_full(vs), (g >> 1) & 1);
        if (g > 0) mbar_wait(s_free, (g - 1) & 1);  // S(g-1) / dP(g-1) are in the softmax warps' registers
        tc_fence_after();
        const uint64_t kd = umma_desc_kmajor(sK(ks)), vd = umma_desc_kmajor(sV(vs));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_ss(tmem_base + DQ_TM_S, qd + uint64_t(k * 2), kd + uint64_t(k * 2), id_ss, k > 0);
            umma_ss(tmem_base + DQ_TM_DP, dod + uint64_t(k * 2), vd + uint64_t(k * 2), id_ss, k > 0);
          }
          umma_commit(sdp_full);
          umma_commit(v_empty(vs));
          if (j == T - 1) umma_commit(qdo_empty);
        }
        __syncwarp();
        if (g > 0) issue_dq(g - 1);
      }
    }
    if (g > 0) issue_dq(g - 1);
  } else {
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const uint32_t tbase = tmem_base + (uint32_t(quad * 32) << 16);
    int g = 0;
    for (int n = 0;; ++n) {
      const int it = queue.consume(n);
      if (it >= n_items) break;
      const int bh = it / nq, qt = it % nq;
      const int b = bh / a.H, h = bh % a.H;
      const int row = qt * 128 + r;
      float nlse = -INFINITY, ndelta = 0.f;  // rows beyond the padded statistics (Nq_pad is a multiple of 64, the tile has 128 rows)
      if (row < a.Nq_pad) {
        const float* st_row = a.stats + ((long long)bh * (a.Nq_pad / 64) + (row >> 6)) * 128 + (row & 63);
        nlse = st_row[0];
        ndelta = st_row[64];
      }
      for (int j = 0; j < T; ++j, ++g) {
        mbar_wait(sdp_full, g & 1);
        tc_fence_after();
        uint32_t s[32], dp[32], ds[16];
        tmem_ld32(tbase + DQ_TM_S + 32 * half, s);
        tmem_ld32(tbase + DQ_TM_DP + 32 * half, dp);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);  // this warp holds its part of S(g) / dP(g) in registers
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float x0, x1, t0, t1;
          ffma2_ss(x0, x1, __uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]), a.scale_log2, nlse);
          const float p0 = ex2(x0), p1 = ex2(x1);
          ffma2_ss(t0, t1, __uint_as_float(dp[2 * i]), __uint_as_float(dp[2 * i + 1]), 1.0f, ndelta);
          fmul2(t0, t1, t0, t1, p0, p1);
          ds[i] = pack_bf16(t0, t1);
        }
        if (g > 0) {  // dQ(g-1) has consumed the previous dS
          mbar_wait(dq_done, (g - 1) & 1);
          tc_fence_after();
        }
        tmem_st16(tbase + DQ_TM_DS + 16 * half, ds);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ds_ready);
      }
      mbar_wait(dq_done, (g - 1) & 1);
      tc_fence_after();
      const bool ok = row < a.Nq;
      const long long tok = (long long)b * a.Nq + row;
      store_grad_half(tbase + DQ_TM_DQ + 32 * half, a.dq + tok * a.lddq + h * 64 + 32 * half, ok, a.scale,
                      (ok && a.q_positions) ? a.q_positions + 2 * tok + half : nullptr, a.rope_table);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, DQ_TM_COLS);
  }
}

// ================================================================================================================
// backward: dK, dV  (key-outer; 64-query sub-tiles; transposed scores: thread <-> (key row, 32-query half of the sub-tile))
// ================================================================================================================
constexpr uint32_t KV_KT = 128 * 64 * 2, KV_QT = 64 * 64 * 2;
constexpr uint32_t KV_OFF_V = KV_KT, KV_OFF_QDO = 2 * KV_KT, KV_OFF_STATS = KV_OFF_QDO + 3 * 2 * KV_QT, KV_OFF_BAR = KV_OFF_STATS + 3 * 512;
constexpr uint32_t KV_SMEM = KV_OFF_BAR + 256 + 1024;
constexpr uint32_t KV_TM_S = 0, KV_TM_DP = 64, KV_TM_DV = 128, KV_TM_DK = 192, KV_TM_COLS = 256;

__global__ void __launch_bounds__(A3_THREADS, 2)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO, const BwdArgs a, int* work) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sK = smem_base, sV = smem_base + KV_OFF_V;
  auto sQ = [&](int st) { return smem_base + KV_OFF_QDO + st * 2 * KV_QT; };
  auto sdO = [&](int st) { return smem_base + KV_OFF_QDO + st * 2 * KV_QT + KV_QT; };
  auto sStats = [&](int st) { return smem_base + KV_OFF_STATS + st * 512; };
  const float* stats_gen = reinterpret_cast<const float*>(smem_gen + KV_OFF_STATS);
  const uint32_t bar = smem_base + KV_OFF_BAR;
  const uint32_t kv_full = bar, kv_empty = bar + 8u;
  auto qdo_full = [&](int st) { return bar + 8u * (2 + st); };
  auto qdo_empty = [&](int st) { return bar + 8u * (5 + st); };
  const uint32_t sdp_full = bar + 8u * 8, ds_ready = bar + 8u * 9, acc_done = bar + 8u * 10, tmem_slot = bar + 8u * 11;
  const ItemQueue queue{bar + 8u * 13, reinterpret_cast<volatile int*>(smem_gen + KV_OFF_BAR + 8 * 17), work};

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = a.Nq_pad / 64;        // query sub-tiles per item
  const int nk = (a.Nk + 127) / 128;  // key tiles per (batch, head)
  const int n_items = nk * a.B * a.H;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    mbar_init(kv_full, 1);
    mbar_init(kv_empty, 1);
    for (int s = 0; s < 3; ++s) { mbar_init(qdo_full(s), 1); mbar_init(qdo_empty(s), 1); }
    mbar_init(sdp_full, 1);
    mbar_init(ds_ready, 8);
    mbar_init(acc_done, 1);
    queue.init();
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, KV_TM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    int g = 0;
    for (int n = 0;; ++n) {
      const int it = queue.produce(n, lane);
      if (it >= n_items) {
        queue.retire(lane);
        break;
      }
      const int bh = it / nk, kt = it % nk;
      const int b = bh / a.H, h = bh % a.H;
      mbar_wait(kv_empty, (n & 1) ^ 1u);  // S^T / dP^T MMAs of the previous item's last sub-tile have read K and V
      if (elect_one()) {
        mbar_arrive_expect_tx(kv_full, 2 * KV_KT);
        tma_load_3d(sK, &tmK, kv_full, h * 64, kt * 128, b);
        tma_load_3d(sV, &tmV, kv_full, h * 64, kt * 128, b);
      }
      __syncwarp();
      for (int j = 0; j < T; ++j, ++g) {
        const int st = g % 3;
        mbar_wait(qdo_empty(st), ((g / 3) & 1) ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(qdo_full(st), 2 * KV_QT + 512);
          tma_load_3d(sQ(st), &tmQ, qdo_full(st), h * 64, j * 64, b);
          tma_load_3d(sdO(st), &tmdO, qdo_full(st), h * 64, j * 64, b);
          bulk_load_1d(sStats(st), a.stats + ((long long)bh * T + j) * 128, 512, qdo_full(st));
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    const uint32_t id_ss = umma_idesc_bf16(128, 64, 0, 0);  // [128 keys] x [64 queries], both K-major (contraction over d)
    const uint32_t id_ts = umma_idesc_bf16(128, 64, 0, 1);  // A = P^T / dS^T in TMEM, B = dO / Q sub-tile read MN-major
    const uint64_t kd = umma_desc_kmajor(sK), vd = umma_desc_kmajor(sV);
    int g = 0;
    for (int n = 0;; ++n) {
     if (queue.consume(n) >= n_items) break;
     for (int j = 0; j < T; ++j, ++g) {
      const int st = g % 3;
      if (j == 0) mbar_wait(kv_full, n & 1);
      mbar_wait(qdo_full(st), (g / 3) & 1);
      tc_fence_after();
      // S^T(g), dP^T(g) reuse the columns that P^T(g-1), dS^T(g-1) occupy: tcgen05.mma executes in issue order, so they
      // start behind the dV / dK MMAs of sub-tile g-1 that read those columns
      const uint64_t qd = umma_desc_kmajor(sQ(st)), dod = umma_desc_kmajor(sdO(st));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_ss(tmem_base + KV_TM_S, kd + uint64_t(k * 2), qd + uint64_t(k * 2), id_ss, k > 0);
          umma_ss(tmem_base + KV_TM_DP, vd + uint64_t(k * 2), dod + uint64_t(k * 2), id_ss, k > 0);
        }
        umma_commit(sdp_full);
        if (j == T - 1) umma_commit(kv_empty);
      }
      __syncwarp();
      mbar_wait(ds_ready, g & 1);
      tc_fence_after();
      const uint64_t q_mn = umma_desc_mnmajor(sQ(st), 8192), do_mn = umma_desc_mnmajor(sdO(st), 8192);
      const uint32_t acc = j > 0 ? 1u : 0u;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // packed P^T / dS^T: query half 0 at columns [0,16), half 1 at [32,48) of the S^T / dP^T regions
          const uint32_t acol = uint32_t((k >> 1) * 32 + (k & 1) * 8);
          umma_ts(tmem_base + KV_TM_DV, tmem_base + KV_TM_S + acol, do_mn + uint64_t(k * 128), id_ts, (acc || k > 0) ? 1u : 0u);
          umma_ts(tmem_base + KV_TM_DK, tmem_base + KV_TM_DP + acol, q_mn + uint64_t(k * 128), id_ts, (acc || k > 0) ? 1u : 0u);
        }
        umma_commit(qdo_empty(st));
        umma_commit(acc_done);
      }
      __syncwarp();
     }
    }
  } else {
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const uint32_t tbase = tmem_base + (uint32_t(quad * 32) << 16);
    int g = 0;
    for (int n = 0;; ++n) {
      const int it = queue.consume(n);
      if (it >= n_items) break;
      const int bh = it / nk, kt = it % nk;
      const int b = bh / a.H, h = bh % a.H;
      for (int j = 0; j < T; ++j, ++g) {
        const int st = g % 3;
        mbar_wait(qdo_full(st), (g / 3) & 1);  // the sub-tile's statistics are in shared memory
        const float4* nl4 = reinterpret_cast<const float4*>(stats_gen + st * 128) + 8 * half;
        const float4* nd4 = nl4 + 16;
        mbar_wait(sdp_full, g & 1);
        tc_fence_after();
        uint32_t s[32], dp[32], pt[16], dst[16];
        tmem_ld32(tbase + KV_TM_S + 32 * half, s);
        tmem_ld32(tbase + KV_TM_DP + 32 * half, dp);
        tmem_ld_wait();
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const float4 nl = nl4[q4], nd = nd4[q4];
          float x0, x1, x2, x3, t0, t1, t2, t3;
          ffma2_sv(x0, x1, __uint_as_float(s[4 * q4 + 0]), __uint_as_float(s[4 * q4 + 1]), a.scale_log2, nl.x, nl.y);
          ffma2_sv(x2, x3, __uint_as_float(s[4 * q4 + 2]), __uint_as_float(s[4 * q4 + 3]), a.scale_log2, nl.z, nl.w);
          const float p0 = ex2(x0), p1 = ex2(x1), p2 = ex2(x2), p3 = ex2(x3);
          fadd2(t0, t1, __uint_as_float(dp[4 * q4 + 0]), __uint_as_float(dp[4 * q4 + 1]), nd.x, nd.y);
          fadd2(t2, t3, __uint_as_float(dp[4 * q4 + 2]), __uint_as_float(dp[4 * q4 + 3]), nd.z, nd.w);
          fmul2(t0, t1, t0, t1, p0, p1);
          fmul2(t2, t3, t2, t3, p2, p3);
          pt[2 * q4] = pack_bf16(p0, p1);
          pt[2 * q4 + 1] = pack_bf16(p2, p3);
          dst[2 * q4] = pack_bf16(t0, t1);
          dst[2 * q4 + 1] = pack_bf16(t2, t3);
        }
        // packed P^T / dS^T overwrite the first 16 of this thread's own (fully consumed) 32 fp32 columns
        tmem_st16(tbase + KV_TM_S + 32 * half, pt);
        tmem_st16(tbase + KV_TM_DP + 32 * half, dst);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ds_ready);
      }
      mbar_wait(acc_done, (g - 1) & 1);
      tc_fence_after();
      const int kv = kt * 128 + r;
      const bool ok = kv < a.Nk;
      const long long tok = (long long)b * a.Nk + kv;
      store_grad_half(tbase + KV_TM_DV + 32 * half, a.dv + tok * a.lddv + h * 64 + 32 * half, ok, 1.0f, nullptr, nullptr);
      store_grad_half(tbase + KV_TM_DK + 32 * half, a.dk + tok * a.lddk + h * 64 + 32 * half, ok, a.scale,
                      (ok && a.k_positions) ? a.k_positions + 2 * tok + half : nullptr, a.rope_table);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, KV_TM_COLS);
  }
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

template <typename K>
int set_smem(K kernel, uint32_t bytes, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  UC_REQUIRE(e == cudaSuccess, UC_ERR_CUDA, "%s: cudaFuncSetAttribute failed: %s", what, cudaGetErrorString(e));
  return UC_OK;
}

dim3 persistent_grid(long long items) {
  const long long slots = 2ll * sm_count();
  return dim3((unsigned)(items < slots ? items : slots));
}

}  // namespace
}  // namespace uc

extern "C" __attribute__((visibility("default"))) int uc_debug_set_attn2_trace(long long* buf) {
  return cudaMemcpyToSymbol(uc::g_attn2_trace, &buf, sizeof(buf)) == cudaSuccess ? 0 : UC_ERR_CUDA;
}

extern "C" int uc_attn_fwd(const uc_attn_fwd_params* p, uc_stream_t stream_) {
  using namespace uc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  UC_REQUIRE(p && p->q && p->k && p->v && p->o, UC_ERR_BAD_SHAPE, "uc_attn_fwd: null pointer");
  UC_REQUIRE(p->B > 0 && p->H > 0 && p->Nq > 0 && p->Nk > 0, UC_ERR_BAD_SHAPE, "uc_attn_fwd: bad shape");
  UC_REQUIRE(p->ldq % 8 == 0 && p->ldk % 8 == 0 && p->ldv % 8 == 0 && p->ldo % 8 == 0, UC_ERR_BAD_SHAPE,
             "uc_attn_fwd: leading dimensions must be multiples of 8");
  UC_REQUIRE(((uintptr_t)p->q % 16 == 0) && ((uintptr_t)p->k % 16 == 0) && ((uintptr_t)p->v % 16 == 0) &&
                 ((uintptr_t)p->o % 16 == 0),
             UC_ERR_BAD_SHAPE, "uc_attn_fwd: pointers must be 16-byte aligned");
  // 1: first-generation kernel, 3: persistent kernel with two independent CTAs per SM (both kept as A/B baselines),
  // 4 (default): ping-pong kernel, one CTA per SM working on two query tiles
  static const int impl_env = env_int("UC_ATTN_FWD", 0);
  if (impl_env == 1) return attn_fwd_v1(p, stream);
  // default: the ping-pong kernel when the query tiles pair up exactly, else (odd tile count, e.g. 1369 tokens = 11 tiles: its
  // group B would repeat a tile for nothing) the two-CTAs-per-SM kernel.  Measured: profiles/r02e_attention_fwd4.log
  const int impl = impl_env ? impl_env : ((((p->Nq + 127) / 128) & 1) ? 3 : 4);
  CUtensorMap tmQ, tmK, tmV;
  int r;
  if ((r = make_head_map(&tmQ, p->q, p->H, p->Nq, p->B, p->ldq, 128))) return r;
  if ((r = make_head_map(&tmK, p->k, p->H, p->Nk, p->B, p->ldk, 128))) return r;
  if ((r = make_head_map(&tmV, p->v, p->H, p->Nk, p->B, p->ldv, 128))) return r;
  static bool configured = false;
  if (!configured) {
    if ((r = set_smem(attn_fwd3_kernel<0>, F3_SMEM, "uc_attn_fwd"))) return r;
    if ((r = set_smem(attn_fwd3_kernel<4>, F3_SMEM, "uc_attn_fwd"))) return r;
    if ((r = set_smem(attn_fwd3_kernel<8>, F3_SMEM, "uc_attn_fwd"))) return r;
    configured = true;
  }
  FwdArgs a;
  a.o = static_cast<__nv_bfloat16*>(p->o);
  a.lse = p->lse;
  a.B = p->B; a.H = p->H; a.Nq = p->Nq; a.Nk = p->Nk;
  a.ldo = p->ldo;
  a.scale = p->scale;
  a.scale_log2 = p->scale * kLog2e;
  const long long items = (long long)((p->Nq + 127) / 128) * p->B * p->H;
  static const int npoly = env_int("UC_ATTN_POLY", 4);  // column pairs (of 16 per 32-column chunk) exponentiated on the FMA pipe
  int* work = work_slot();
  UC_REQUIRE(work, UC_ERR_CUDA, "uc_attn_fwd: work counters unavailable");
  if (impl == 4) {
    static bool configured4 = false;
    if (!configured4) {
      if ((r = set_smem(attn_fwd4_kernel<0>, F4_SMEM, "uc_attn_fwd"))) return r;
      if ((r = set_smem(attn_fwd4_kernel<4>, F4_SMEM, "uc_attn_fwd"))) return r;
      if ((r = set_smem(attn_fwd4_kernel<8>, F4_SMEM, "uc_attn_fwd"))) return r;
      configured4 = true;
    }
    const long long items2 = (long long)(((p->Nq + 127) / 128 + 1) / 2) * p->B * p->H;
    const long long sms = sm_count();
    const dim3 grid4((unsigned)(items2 < sms ? items2 : sms));
    cudaError_t le4 = npoly >= 8   ? launch_pdl(attn_fwd4_kernel<8>, grid4, dim3(F4_THREADS), F4_SMEM, stream, tmQ, tmK, tmV, a, work)
                      : npoly >= 4 ? launch_pdl(attn_fwd4_kernel<4>, grid4, dim3(F4_THREADS), F4_SMEM, stream, tmQ, tmK, tmV, a, work)
                                   : launch_pdl(attn_fwd4_kernel<0>, grid4, dim3(F4_THREADS), F4_SMEM, stream, tmQ, tmK, tmV, a, work);
    UC_REQUIRE(le4 == cudaSuccess, UC_ERR_CUDA, "uc_attn_fwd: launch failed: %s", cudaGetErrorString(le4));
    return check_launch("uc_attn_fwd");
  }
  const dim3 grid = persistent_grid(items);
  cudaError_t le = npoly >= 8   ? launch_pdl(attn_fwd3_kernel<8>, grid, dim3(A3_THREADS), F3_SMEM, stream, tmQ, tmK, tmV, a, work)
                   : npoly >= 4 ? launch_pdl(attn_fwd3_kernel<4>, grid, dim3(A3_THREADS), F3_SMEM, stream, tmQ, tmK, tmV, a, work)
                                : launch_pdl(attn_fwd3_kernel<0>, grid, dim3(A3_THREADS), F3_SMEM, stream, tmQ, tmK, tmV, a, work);
  UC_REQUIRE(le == cudaSuccess, UC_ERR_CUDA, "uc_attn_fwd: launch failed: %s", cudaGetErrorString(le));
  return check_launch("uc_attn_fwd");
}

extern "C" int uc_attn_bwd(const uc_attn_bwd_params* p, uc_stream_t stream_) {
  using namespace uc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  UC_REQUIRE(p && p->q && p->k && p->v && p->o && p->d_o && p->lse && p->delta && p->dq && p->dk && p->dv, UC_ERR_BAD_SHAPE,
             "uc_attn_bwd: null pointer");
  UC_REQUIRE(p->B > 0 && p->H > 0 && p->Nq > 0 && p->Nk > 0, UC_ERR_BAD_SHAPE, "uc_attn_bwd: bad shape");
  UC_REQUIRE(p->ldq % 8 == 0 && p->ldk % 8 == 0 && p->ldv % 8 == 0 && p->ldo % 8 == 0 && p->lddq % 8 == 0 && p->lddk % 8 == 0 &&
                 p->lddv % 8 == 0,
             UC_ERR_BAD_SHAPE, "uc_attn_bwd: leading dimensions must be multiples of 8");
  UC_REQUIRE((p->q_positions == nullptr && p->k_positions == nullptr) || p->rope_table, UC_ERR_BAD_SHAPE,
             "uc_attn_bwd: positions given without rope_table");
  static const int impl = env_int("UC_ATTN_BWD", 1);  // 1: fused one-kernel backward (attention_bwd.cu; needs dq_acc), 3: the two
                                                      // bit-reproducible kernels of this file (no accumulator, no atomics)
  if (impl == 1) {
    UC_REQUIRE(p->dq_acc, UC_ERR_BAD_SHAPE, "uc_attn_bwd: null pointer (dq_acc, first-generation kernel)");
    return attn_bwd_v1(p, stream);
  }
  const int Nq_pad = (p->Nq + 63) / 64 * 64;
  int r;
  {
    const long long total = (long long)p->B * p->H * Nq_pad * 8;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    cudaError_t le = launch_pdl(attn_bwd_stats_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, static_cast<const __nv_bfloat16*>(p->o),
                                static_cast<const __nv_bfloat16*>(p->d_o), p->lse, p->delta, p->B, p->H, p->Nq, Nq_pad,
                                (long long)p->ldo, (long long)p->ldo);
    UC_REQUIRE(le == cudaSuccess, UC_ERR_CUDA, "uc_attn_bwd(stats): launch failed: %s", cudaGetErrorString(le));
    if ((r = check_launch("uc_attn_bwd(stats)"))) return r;
  }
  CUtensorMap tmQ, tmK, tmV, tmdO, tmQs, tmKs, tmVs, tmdOs;
  if ((r = make_head_map(&tmQ, p->q, p->H, p->Nq, p->B, p->ldq, 128))) return r;     // dQ kernel: 128-query tiles,
  if ((r = make_head_map(&tmdO, p->d_o, p->H, p->Nq, p->B, p->ldo, 128))) return r;
  if ((r = make_head_map(&tmKs, p->k, p->H, p->Nk, p->B, p->ldk, 64))) return r;     //            64-key tiles
  if ((r = make_head_map(&tmVs, p->v, p->H, p->Nk, p->B, p->ldv, 64))) return r;
  if ((r = make_head_map(&tmK, p->k, p->H, p->Nk, p->B, p->ldk, 128))) return r;     // dK,dV kernel: 128-key tiles,
  if ((r = make_head_map(&tmV, p->v, p->H, p->Nk, p->B, p->ldv, 128))) return r;
  if ((r = make_head_map(&tmQs, p->q, p->H, p->Nq, p->B, p->ldq, 64))) return r;     //               64-query sub-tiles
  if ((r = make_head_map(&tmdOs, p->d_o, p->H, p->Nq, p->B, p->ldo, 64))) return r;
  static bool configured = false;
  if (!configured) {
    if ((r = set_smem(attn_bwd_dq_kernel, DQ_SMEM, "uc_attn_bwd"))) return r;
    if ((r = set_smem(attn_bwd_dkv_kernel, KV_SMEM, "uc_attn_bwd"))) return r;
    configured = true;
  }
  BwdArgs a;
  a.stats = p->delta;
  a.dq = static_cast<__nv_bfloat16*>(p->dq);
  a.dk = static_cast<__nv_bfloat16*>(p->dk);
  a.dv = static_cast<__nv_bfloat16*>(p->dv);
  a.B = p->B; a.H = p->H; a.Nq = p->Nq; a.Nk = p->Nk; a.Nq_pad = Nq_pad;
  a.lddq = p->lddq; a.lddk = p->lddk; a.lddv = p->lddv;
  a.scale = p->scale;
  a.scale_log2 = p->scale * kLog2e;
  a.q_positions = p->q_positions;
  a.k_positions = p->k_positions;
  a.rope_table = p->rope_table;
  const long long kv_items = (long long)((p->Nk + 127) / 128) * p->B * p->H, q_items = (long long)((p->Nq + 127) / 128) * p->B * p->H;
  int* work_kv = work_slot();
  int* work_q = work_slot();
  UC_REQUIRE(work_kv && work_q, UC_ERR_CUDA, "uc_attn_bwd: work counters unavailable");
  cudaError_t le = launch_pdl(attn_bwd_dkv_kernel, persistent_grid(kv_items), dim3(A3_THREADS), KV_SMEM, stream, tmQs, tmK, tmV, tmdOs, a,
                              work_kv);
  UC_REQUIRE(le == cudaSuccess, UC_ERR_CUDA, "uc_attn_bwd(dkv): launch failed: %s", cudaGetErrorString(le));
  if ((r = check_launch("uc_attn_bwd(dkv)"))) return r;
  le = launch_pdl(attn_bwd_dq_kernel, persistent_grid(q_items), dim3(A3_THREADS), DQ_SMEM, stream, tmQ, tmKs, tmVs, tmdO, a, work_q);
  UC_REQUIRE(le == cudaSuccess, UC_ERR_CUDA, "uc_attn_bwd(dq): launch failed: %s", cudaGetErrorString(le));
  return check_launch("uc_attn_bwd(dq)");
}
