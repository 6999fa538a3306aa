// tcgen05 GEMM with fused epilogue (uc_gemm): the kernel behind every Linear / 1x1-conv / patch-embed
// of the DUSt3R path and their dgrad / wgrad.
//
//   warp 0      : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier full/empty)
//   warp 1      : MMA issuer     (one thread: tcgen05.mma kind::f16, 128 x BN x 16, fp32 accum in TMEM)
//   warps 2..9  : epilogue       (tcgen05.ld -> bias / RoPE / GELU / GELU' / residual -> global)
// Persistent CTAs (grid = #SMs), two TMEM accumulator buffers so the epilogue of tile i overlaps the
// main loop of tile i+1.  Both operands may be K-major or MN-major (UMMA descriptor major bits), so
// forward (x W^T), dgrad (dy W) and wgrad (dy^T x, split-K + fp32 red.add) share this one kernel without
// any transposition pass.
#include <cstdlib>

#include "common.cuh"

namespace uc {

// Optional launch-latency trace of the pair kernel (bring-up aid, tools/latency_probe.py --trace): per CTA 8 globaltimer stamps
// [entry, set-up done, dependency wait done, first operands landed, first accumulator complete, epilogue of last tile done, exit].
__device__ unsigned long long* g_gemm_trace = nullptr;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int NUM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + NUM_EPI_WARPS * 32;

struct GemmArgs {
  int m, n, k;
  int a_mn, b_mn;
  int num_m, num_n, num_kb, split_k, kb_per_split;
  int epilogue, c_f32, rope_cols, n_fastest;
  long long ldc;
  void* c;
  const float* bias;
  const __nv_bfloat16* residual;
  __nv_bfloat16* aux_out;
  const __nv_bfloat16* aux_in;
  const int* positions;
  const float* rope_table;
  float* colsum;  // optional [n]: += column sums of the bf16 C tile (pair kernel epilogue)
  int* work;      // pair kernel: (next item, clusters finished) counters of this launch (runtime.cu: work_slot)
  // implicit-GEMM 3x3 convolution (uc_conv3x3; pair kernel instantiated with CONV != 0), NHWC maps, stride 1, pad 1:
  //   CONV 1 (fwd / dgrad): the A tile of a CTA is a TW x TH pixel window (TW * TH = 128) of one image, fetched per k-block by a
  //     4-D TMA box at the tap's offset (out-of-image pixels zero-fill = the padding); C / aux tiles leave / arrive the same way.
  //   CONV 2 (wgrad): K runs over pixels; a k-block is a TW x TH window (TW * TH = 64) of dy (A, unshifted) and of x (B, shifted
  //     by the tap of each 64-column block of the [cout, 9 cin] weight gradient).
  //   CONV 3 (patch embedding, uc_patch_embed): the "map" is the patch grid; A comes straight from the fp32 NCHW image through a
  //     5-D TMA box (dx, dy, patch x, patch y, batch*3 + channel) -- a k-block is 32 fp32 = 32 / p rows of one patch of one
  //     channel -- and the MMAs run in TF32 on the fp32 master weights; C leaves like CONV 1.  No column buffer.
  int cv_H, cv_W, cv_tw_log2, cv_tiles_x, cv_tiles_y;
  int cv_kb_per_tap;    // CONV 1: channel blocks (of 64) per tap in A; CONV 3: k-blocks per image channel (p * p / 32)
  int cv_b_tap_stride;  // CONV 1: 0 = B is K-major with k = tap * C + c (fwd); > 0 = B is MN-major and tap t reads the column
                        // block (8 - t) * stride (dgrad: the spatially flipped filter)
  int cv_cin;           // CONV 2: channels of x (columns per tap); CONV 3: patch rows per k-block (32 / p)
};

template <int BN>
struct Cfg {
  static constexpr uint32_t A_BYTES = BM * BK * 2;
  static constexpr uint32_t B_BYTES = BN * BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr uint32_t SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr uint32_t TMEM_COLS = 2 * BN;
};

__device__ __forceinline__ void load_row32_bf16(const __nv_bfloat16* p, float (&v)[32]) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 t = __ldg(q + i);
    v[8 * i + 0] = bf16_lo(t.x); v[8 * i + 1] = bf16_hi(t.x);
    v[8 * i + 2] = bf16_lo(t.y); v[8 * i + 3] = bf16_hi(t.y);
    v[8 * i + 4] = bf16_lo(t.z); v[8 * i + 5] = bf16_hi(t.z);
    v[8 * i + 6] = bf16_lo(t.w); v[8 * i + 7] = bf16_hi(t.w);
  }
}
__device__ __forceinline__ void store_row32_bf16(__nv_bfloat16* p, const float (&v)[32]) {
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 t;
    t.x = pack_bf16(v[8 * i + 0], v[8 * i + 1]);
    t.y = pack_bf16(v[8 * i + 2], v[8 * i + 3]);
    t.z = pack_bf16(v[8 * i + 4], v[8 * i + 5]);
    t.w = pack_bf16(v[8 * i + 6], v[8 * i + 7]);
    q[i] = t;
  }
}

// Fused epilogue math on one 32-column chunk of one accumulator row (fp32 bits in r[]):
// bias -> 2-D RoPE -> GELU (bf16 pre-activation returned packed in prep[]) / GELU' -> residual.  No stores.
template <int MASK = -1>
__device__ __forceinline__ void epilogue_math(const GemmArgs& g, int row, int n, int pos_y, int pos_x, bool row_ok,
                                              const uint32_t (&r)[32], float (&v)[32], uint32_t (&prep)[16] /* packed bf16 pre-activation */,
                                              const float* hin = nullptr /* staged aux_in / residual values, or load */) {
  const int epi = g.epilogue & MASK;  // MASK: flags this instantiation can see (everything else is compiled out)
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  if ((epi & UC_EPI_BIAS) && n < g.n) {  // (n >= g.n: partial last column tile of a convolution, clipped by the TMA store)
    const float4* b4 = reinterpret_cast<const float4*>(g.bias + n);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 b = __ldg(b4 + j);
      v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
    }
  }
  if ((epi & UC_EPI_ROPE) && n < g.rope_cols) {
    const int p = ((n >> 5) & 1) ? pos_x : pos_y;
    const float4* t4 = reinterpret_cast<const float4*>(g.rope_table + (size_t)p * 32);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 cs = __ldg(t4 + j);  // (cos_{2j}, sin_{2j}, cos_{2j+1}, sin_{2j+1})
      float u0 = v[2 * j], w0 = v[2 * j + 16];
      v[2 * j] = u0 * cs.x - w0 * cs.y;
      v[2 * j + 16] = w0 * cs.x + u0 * cs.y;
      float u1 = v[2 * j + 1], w1 = v[2 * j + 17];
      v[2 * j + 1] = u1 * cs.z - w1 * cs.w;
      v[2 * j + 17] = w1 * cs.z + u1 * cs.w;
    }
  }
  const size_t off = (size_t)row * g.ldc + n;
  if (epi & UC_EPI_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (epi & UC_EPI_GELU) {
    // the pre-activation is rounded to bf16 by the pack that its store needs anyway (F2FP, not the XU-pipe F2F)
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const uint32_t pk = pack_bf16(v[2 * j], v[2 * j + 1]);
      prep[j] = pk;
      v[2 * j] = gelu_erf(bf16_lo(pk));
      v[2 * j + 1] = gelu_erf(bf16_hi(pk));
    }
  }
  if ((epi & UC_EPI_RELU_BWD) && row_ok) {
    float h[32];
    if (hin) {
#pragma unroll
      for (int j = 0; j < 32; ++j) h[j] = hin[j];
    } else {
      load_row32_bf16(g.aux_in + off, h);
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = h[j] > 0.f ? v[j] : 0.f;
  }
  if ((epi & UC_EPI_GELU_BWD) && row_ok) {
    float h[32];
    if (hin) {
#pragma unroll
      for (int j = 0; j < 32; ++j) h[j] = hin[j];
    } else {
      load_row32_bf16(g.aux_in + off, h);
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= gelu_erf_grad(h[j]);
  }
  if ((epi & UC_EPI_RESIDUAL) && row_ok) {
    float h[32];
    if (hin) {
#pragma unroll
      for (int j = 0; j < 32; ++j) h[j] = hin[j];
    } else if (epi & UC_EPI_RESIDUAL_F32) {
      const float4* q = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(g.residual) + off);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 t = __ldg(q + j);
        h[4 * j] = t.x; h[4 * j + 1] = t.y; h[4 * j + 2] = t.z; h[4 * j + 3] = t.w;
      }
    } else {
      load_row32_bf16(g.residual + off, h);
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += h[j];
  }
}

// epilogue math + direct row-per-thread global stores (single-CTA kernels)
__device__ __forceinline__ void epilogue_chunk(const GemmArgs& g, int row, int n, int pos_y, int pos_x, const uint32_t (&r)[32]) {
  float v[32];
  uint32_t prep[16];
  epilogue_math(g, row, n, pos_y, pos_x, true, r, v, prep);
  const size_t off = (size_t)row * g.ldc + n;
  if (g.epilogue & UC_EPI_GELU) {
    uint4* q = reinterpret_cast<uint4*>(g.aux_out + off);
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = make_uint4(prep[4 * i], prep[4 * i + 1], prep[4 * i + 2], prep[4 * i + 3]);
  }
  if (g.c_f32) {
    float* cp = reinterpret_cast<float*>(g.c) + off;
    if (g.epilogue & UC_EPI_ATOMIC) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp + 4 * j), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                     "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                     : "memory");
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        reinterpret_cast<float4*>(cp)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  } else {
    store_row32_bf16(reinterpret_cast<__nv_bfloat16*>(g.c) + off, v);
  }
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  using C = Cfg<BN>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * C::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_launch_dependents();  // after the TMEM allocation (common.cuh: PDL rules)
  pdl_wait();               // operands / epilogue inputs may come from the previous kernel in the stream

  const int total = g.num_m * g.num_n * g.split_k;

  if (warp == 0) {
    // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
      const int split = item % g.split_k;
      const int tile = item / g.split_k;
      const int m0 = g.n_fastest ? (tile / g.num_n) * BM : (tile % g.num_m) * BM;
      const int n0 = g.n_fastest ? (tile % g.num_n) * BN : (tile / g.num_m) * BN;
      const int kb0 = split * g.kb_per_split;
      const int kb1 = min(g.num_kb, kb0 + g.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
        const uint32_t sb = sa + C::A_BYTES;
        if (elect_one()) {
          mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
          if (!g.a_mn) {
            tma_load_2d(sa, &tmA, full_bar(stage), kb * BK, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * 8192, &tmA, full_bar(stage), m0 + j * 64, kb * BK);
          }
          if (!g.b_mn) {
            tma_load_2d(sb, &tmB, full_bar(stage), kb * BK, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &tmB, full_bar(stage), n0 + j * 64, kb * BK);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp loops, operands uniform, one elected lane issues) =====================
    const uint32_t idesc = umma_idesc_bf16(BM, BN, g.a_mn, g.b_mn);
    const uint32_t a_step = g.a_mn ? (2048u >> 4) : (32u >> 4);  // descriptor start-address step per UMMA_K
    const uint32_t b_step = g.b_mn ? (2048u >> 4) : (32u >> 4);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
      const int split = item % g.split_k;
      const int kb0 = split * g.kb_per_split;
      const int kb1 = min(g.num_kb, kb0 + g.kb_per_split);
      if (kb0 >= kb1) continue;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
        const uint32_t sb = sa + C::A_BYTES;
        const uint64_t adesc = g.a_mn ? umma_desc_mnmajor(sa, 8192) : umma_desc_kmajor(sa);
        const uint64_t bdesc = g.b_mn ? umma_desc_mnmajor(sb, 8192) : umma_desc_kmajor(sb);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_ss(d_tmem, adesc + uint64_t(k * a_step), bdesc + uint64_t(k * b_step), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(empty_bar(stage));  // smem slot reusable once these MMAs retire
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      if (elect_one()) umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else {
    // ===================== epilogue =====================
    const int e = warp - 2;
    const int lane_group = warp & 3;          // TMEM lanes this warp may touch: 32*(warp%4)..
    const int col_half = e >> 2;              // 0 / 1: which half of the BN columns
    constexpr int CHUNKS = BN / 2 / 32;       // 32-column chunks per warp
    const int epi = g.epilogue;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
      const int split = item % g.split_k;
      const int tile = item / g.split_k;
      const int m0 = g.n_fastest ? (tile / g.num_n) * BM : (tile % g.num_m) * BM;
      const int n0 = g.n_fastest ? (tile % g.num_n) * BN : (tile / g.num_m) * BN;
      const int kb0 = split * g.kb_per_split;
      const int kb1 = min(g.num_kb, kb0 + g.kb_per_split);
      if (kb0 >= kb1) continue;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int row = m0 + lane_group * 32 + lane;
      const bool row_ok = row < g.m;
      int pos_y = 0, pos_x = 0;
      if ((epi & UC_EPI_ROPE) && row_ok) {
        pos_y = g.positions[2 * row];
        pos_x = g.positions[2 * row + 1];
      }
#pragma unroll 1
      for (int ch = 0; ch < CHUNKS; ++ch) {
        const int col = col_half * (BN / 2) + ch * 32;
        const int n = n0 + col;
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_base + (uint32_t(lane_group * 32) << 16) + uint32_t(acc * BN + col), r);
        tmem_ld_wait();
        if (n >= g.n || !row_ok) continue;
        epilogue_chunk(g, row, n, pos_y, pos_x, r);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair GEMM: cluster (2,1,1), tcgen05.mma.cta_group::2, 256 x 256 output tile per pair.
// CTA rank r stages A rows [m0+128r, +128) and B columns [n0+128r, +128) (32 KB / stage instead of 48 KB, so the
// L2->SM operand stream per FLOP drops by a third: 128 FLOP/B instead of 85); the leader CTA's elected lane issues
// the 256x256x16 MMAs for both SMs; each CTA's TMEM holds its own 128 x 256 fp32 accumulator (double-buffered) and
// each CTA runs the same fused epilogue on its rows.  Barriers: TMA of both CTAs credit the leader's `full`;
// `empty` / `tmem_full` are released in both CTAs by a multicast tcgen05.commit; the peer's epilogue warps arrive
// remotely on the leader's `tmem_empty`.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t G2_A_BYTES = 128 * BK * 2;
constexpr uint32_t G2_STG_WARP = 2 * 4096;  // per epilogue warp: two 32-row x 128-byte staging tiles (see the epilogue)
constexpr uint32_t G2_STG_BYTES = NUM_EPI_WARPS * G2_STG_WARP;
template <int BN>
struct Cfg2 {  // BN = 256: 256 x 256 tiles (the only instantiation); BN = 128 compiles and is correct but slower per FLOP
  static constexpr int STAGES = BN == 256 ? 5 : 6;
  static constexpr uint32_t B_BYTES = (BN / 2) * BK * 2;
  static constexpr uint32_t STAGE_BYTES = G2_A_BYTES + B_BYTES;
  static constexpr uint32_t SMEM = STAGES * STAGE_BYTES + G2_STG_BYTES + 1024 + 512;
};

// MASK / F32: epilogue flags and output type this instantiation handles.  One generic kernel with run-time flags is
// 12 K SASS instructions (~200 KB): the epilogue warps then stall on instruction fetch (ncu: 1.0 "no_instruction" per issue).
template <int MASK, bool F32, int BN, int CONV = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAux, const GemmArgs g) {
  constexpr int STAGES = Cfg2<BN>::STAGES;
  constexpr uint32_t G2_STAGE_BYTES = Cfg2<BN>::STAGE_BYTES;
  constexpr int NCH = BN / 2 / 32;  // 32-column chunks per epilogue warp
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = smem_base + STAGES * G2_STAGE_BYTES;
  const uint32_t bar_base = stg_base + G2_STG_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  auto ld_bar = [&](int e, int which) { return bar_base + 8u * (2 * STAGES + 6 + 2 * e + which); };  // per epilogue warp: input tile A / B landed
  // Dynamic tile queue.  Items (tile, k-split) are handed out by an atomic counter instead of a static round robin: a
  // cluster that starts late -- its SMs were held by another stream's kernel or by an NCCL all-reduce CTA -- simply takes
  // fewer tiles, where the static split made the whole GEMM wait for a second wave.  The leader CTA's producer warp fetches
  // the next item and publishes it into an 8-slot ring in BOTH CTAs' shared memory (remote store + remote mbarrier arrive
  // with release semantics at cluster scope); every other role of both CTAs follows the ring.  The producer is never more than
  // three items ahead of the slowest role (smem stages, two TMEM accumulators), so eight slots need no "empty" barriers.
  auto q_bar = [&](int s) { return bar_base + 8u * (40 + s); };
  const uint32_t q_slot = bar_base + 8u * 48;
  auto next_item = [&](int n) -> int {  // consumer side (all lanes of the calling warp)
    // no cluster-scope fence here (ptxas turns one into MEMBAR.ALL.GPU + CCTL.IVALL per call): the slot lives in THIS CTA's
    // shared memory, the leader's remote store is ordered before its remote arrive (mbarrier.arrive.release.cluster), and the
    // volatile load below cannot move above the wait
    mbar_wait(q_bar(n & 7), (n >> 3) & 1);
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(q_slot + 4u * (n & 7)) : "memory");
    return (int)v;
  };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int num_clusters = gridDim.x >> 1;
  unsigned long long* const trc = g_gemm_trace ? g_gemm_trace + 16 * blockIdx.x : nullptr;
  if (trc && threadIdx.x == 0) trc[0] = globaltimer_ns();

  if (warp == 0) {
    // ~40 barriers: initialised by the 32 lanes in parallel (one thread doing them in sequence sits on the launch's critical path)
    if (lane == 0) tma_prefetch_desc(&tmA);
    if (lane == 1) tma_prefetch_desc(&tmB);
    if (lane == 2) tma_prefetch_desc(&tmC);
    if (lane == 3) tma_prefetch_desc(&tmAux);
    for (int i = lane; i < 48; i += 32) {
      // slots: [0, 2 S) full / empty (1: the leader producer's expect_tx / one multicast commit), 2 S + {0,1} tmem_full (1),
      // 2 S + {2,3} tmem_empty (epilogue warps of BOTH CTAs), 2 S + 4 / + 5 the TMEM address slot (not a barrier),
      // [2 S + 6, 2 S + 6 + 16) per-epilogue-warp input-tile barriers (1), [40, 48) work-queue ring (1)
      const bool is_slot = (i == 2 * STAGES + 4) || (i == 2 * STAGES + 5);
      const bool in_use = i < 2 * STAGES + 6 + 2 * NUM_EPI_WARPS || i >= 40;
      if (!is_slot && in_use) mbar_init(bar_base + 8u * i, (i == 2 * STAGES + 2 || i == 2 * STAGES + 3) ? 2 * NUM_EPI_WARPS : 1);
    }
    fence_barrier_init();
  }
  cluster_sync_all();  // barrier inits visible to the peer before any remote arrive / TMA credit
  if (warp == 1) {
    tmem_alloc2(tmem_slot, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (trc && threadIdx.x == 0) trc[1] = globaltimer_ns();
  pdl_launch_dependents();  // after the TMEM allocation (common.cuh: PDL rules)
  // the launch's work counter is private to it (not produced by the predecessor kernel): the leader's first fetch goes out
  // BEFORE the grid dependency wait, so its round trip hides under the predecessor's tail
  const int cluster_id = blockIdx.x >> 1;
  int first_item = cluster_id;  // g.work == nullptr: static round robin (item = cluster + k * clusters), no counters
  if (g.work && warp == 0 && rank == 0 && lane == 0) first_item = atomicAdd(g.work, 1);
  pdl_wait();
  if (trc && threadIdx.x == 0) trc[2] = globaltimer_ns();

  const int total = g.num_m * g.num_n * g.split_k;  // num_m counts 256-row blocks here

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    // leader: the fetch of item qn+1 is issued before the k-loop of item qn and published a few k-blocks into it, so neither
    // the atomic's round trip nor the remote store / arrive that tells the peer sits between two tiles' TMA streams
    auto publish = [&](int qn, int item) {
      if (lane == 0) {
        const uint32_t slot = q_slot + 4u * (qn & 7);
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(slot), "r"((uint32_t)item) : "memory");
        st_shared_cluster_u32(slot, 1, (uint32_t)item);
        mbar_arrive(q_bar(qn & 7));
        mbar_arrive_cluster(q_bar(qn & 7), 1);
      }
      __syncwarp();
    };
    int fetched = 0;
    if (rank == 0) {
      fetched = __shfl_sync(0xffffffffu, first_item, 0);
      publish(0, fetched);
    }
    for (int qn = 0;; ++qn) {
      int item;
      int fetched_next = 0;
      if (rank == 0) {
        item = fetched;
        if (item < total && lane == 0)  // consumed below, after the first k-blocks
          fetched_next = g.work ? atomicAdd(g.work, 1) : cluster_id + (qn + 1) * num_clusters;
      } else {
        item = next_item(qn);
      }
      if (item >= total) {
        // this cluster takes no more items.  Once every cluster has said so nobody touches the counters again: the last one
        // resets them for the slot's next launch -- here, in the producer warp, while the MMA / epilogue warps still work on
        // the last tiles, not on the kernel's exit path
        if (g.work && rank == 0 && lane == 0) {
          const int done = atomicAdd(g.work + 1, 1);
          if (done == num_clusters - 1) {
            g.work[0] = 0;
            g.work[1] = 0;
            __threadfence();
          }
        }
        break;
      }
      const int split = item % g.split_k;
      const int tile = item / g.split_k;
      const int m0 = (g.n_fastest ? (tile / g.num_n) : (tile % g.num_m)) * 256 + 128 * (int)rank;
      const int n0 = (g.n_fastest ? (tile % g.num_n) : (tile / g.num_m)) * BN + (BN / 2) * (int)rank;
      const int kb0 = split * g.kb_per_split;
      const int kb1 = min(g.num_kb, kb0 + g.kb_per_split);
      int cx0 = 0, cy0 = 0, cimg = 0;  // CONV 1 / 3: this CTA's pixel (patch) window
      if (CONV == 1 || CONV == 3) {
        const int tm = g.n_fastest ? (tile / g.num_n) : (tile % g.num_m);
        const int per_img = g.cv_tiles_x * g.cv_tiles_y;
        cimg = tm / per_img;
        const int r = tm - cimg * per_img;
        const int ty = r / g.cv_tiles_x;
        cx0 = (r - ty * g.cv_tiles_x) << g.cv_tw_log2;
        cy0 = (2 * ty + (int)rank) * (128 >> g.cv_tw_log2);
      }
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sa = smem_base + stage * G2_STAGE_BYTES;
        const uint32_t sb = sa + G2_A_BYTES;
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * G2_STAGE_BYTES);
          if (CONV == 1) {
            const int tap = kb / g.cv_kb_per_tap;
            const int c0 = (kb - tap * g.cv_kb_per_tap) * 64;
            const int dy = tap / 3, dx = tap - 3 * dy;
            tma2_load_4d(sa, &tmA, full_bar(stage), c0, cx0 + dx - 1, cy0 + dy - 1, cimg);
            if (!g.b_mn) {
              tma2_load_2d(sb, &tmB, full_bar(stage), kb * BK, n0);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 2 / 64; ++j)
                tma2_load_2d(sb + j * 8192, &tmB, full_bar(stage), n0 + j * 64 + (8 - tap) * g.cv_b_tap_stride, c0);
            }
          } else if (CONV == 3) {
            const int ch = kb / g.cv_kb_per_tap;
            const int dy0 = (kb - ch * g.cv_kb_per_tap) * g.cv_cin;
            // one box per patch row: [128 patches x p fp32]; 32 / p of them (64-byte rows: 64-byte swizzle) make the k-block
            for (int rr = 0; rr < g.cv_cin; ++rr)
              tma2_load_5d(sa + rr * (G2_A_BYTES / g.cv_cin), &tmA, full_bar(stage), 0, dy0 + rr, cx0, cy0, 3 * cimg + ch);
            tma2_load_2d(sb, &tmB, full_bar(stage), kb * 32, n0);
          } else if (CONV == 2) {
            const int per_img = g.cv_tiles_x * g.cv_tiles_y;
            const int img = kb / per_img;
            const int r = kb - img * per_img;
            const int ty = r / g.cv_tiles_x;
            const int x0 = (r - ty * g.cv_tiles_x) << g.cv_tw_log2;
            const int y0 = ty * (64 >> g.cv_tw_log2);
            tma2_load_4d(sa, &tmA, full_bar(stage), m0, x0, y0, img);
            tma2_load_4d(sa + 8192, &tmA, full_bar(stage), m0 + 64, x0, y0, img);
#pragma unroll
            for (int j = 0; j < BN / 2 / 64; ++j) {
              const int n = n0 + j * 64;
              const int tap = min(n / g.cv_cin, 8);  // columns >= 9 cin: partial last tile, clipped by the TMA reduce
              const int ci = n - (n / g.cv_cin) * g.cv_cin;
              const int dy = tap / 3, dx = tap - 3 * dy;
              tma2_load_4d(sb + j * 8192, &tmB, full_bar(stage), ci, x0 + dx - 1, y0 + dy - 1, img);
            }
          } else if (!g.a_mn) {
            tma2_load_2d(sa, &tmA, full_bar(stage), kb * BK, m0);
          } else {
            tma2_load_2d(sa, &tmA, full_bar(stage), m0, kb * BK);
            tma2_load_2d(sa + 8192, &tmA, full_bar(stage), m0 + 64, kb * BK);
          }
          if (CONV == 0) {
            if (!g.b_mn) {
              tma2_load_2d(sb, &tmB, full_bar(stage), kb * BK, n0);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 2 / 64; ++j) tma2_load_2d(sb + j * 8192, &tmB, full_bar(stage), n0 + j * 64, kb * BK);
            }
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        if (rank == 0 && kb == min(kb0 + 2, kb1 - 1)) {
          fetched = __shfl_sync(0xffffffffu, fetched_next, 0);
          publish(qn + 1, fetched);
        }
      }
      if (rank == 0 && kb0 >= kb1) {  // empty k-range (cannot happen for a scheduled split, kept for symmetry with the consumers)
        fetched = __shfl_sync(0xffffffffu, fetched_next, 0);
        publish(qn + 1, fetched);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      const uint32_t idesc = CONV == 3 ? umma_idesc_tf32(256, BN) : umma_idesc_bf16(256, BN, g.a_mn, g.b_mn);
      const uint32_t a_step = g.a_mn ? (2048u >> 4) : (32u >> 4);
      const uint32_t b_step = g.b_mn ? (2048u >> 4) : (32u >> 4);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int qn = 0;; ++qn) {
        const int item = next_item(qn);
        if (item >= total) break;
        const int split = item % g.split_k;
        const int kb0 = split * g.kb_per_split;
        const int kb1 = min(g.num_kb, kb0 + g.kb_per_split);
        if (kb0 >= kb1) continue;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (trc && qn == 0 && kb == kb0 && lane == 0) trc[3] = globaltimer_ns();
          const uint32_t sa = smem_base + stage * G2_STAGE_BYTES;
          const uint32_t sb = sa + G2_A_BYTES;
          const uint64_t adesc = g.a_mn ? umma_desc_mnmajor(sa, 8192) : umma_desc_kmajor(sa);
          const uint64_t bdesc = g.b_mn ? umma_desc_mnmajor(sb, 8192) : umma_desc_kmajor(sb);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              if (CONV == 3) {
                // A: one 128-byte-swizzled tile (patch 32) or two 64-byte-swizzled half tiles of 8 KB (patch 16), 8 tf32 per step
                const uint64_t ad = g.cv_cin == 2 ? umma_desc_kmajor_sw64(sa + (k >> 1) * 8192) + uint64_t((k & 1) * 2) : adesc + uint64_t(k * 2);
                umma2_ss_tf32(d_tmem, ad, bdesc + uint64_t(k * b_step), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              } else {
                umma2_ss(d_tmem, adesc + uint64_t(k * a_step), bdesc + uint64_t(k * b_step), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              }
            }
            umma2_commit_mc(empty_bar(stage));
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) umma2_commit_mc(tfull_bar(acc));
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    // TMEM -> registers -> fused math -> 128B-swizzled smem staging tile (32 rows x 128 B) -> TMA store / TMA
    // reduce-add.  A row-per-thread direct store would cost 32 partial-line transactions per instruction;
    // the staged path writes full 128-byte lines and runs asynchronously behind the next chunk's math.
    const int e = warp - 2;
    const int lane_group = warp & 3;
    const int col_half = e >> 2;
    const uint32_t stg = stg_base + e * G2_STG_WARP;
    // two tiles per warp.  plain: C of unit u -> buf[u];  GELU: C -> buf0, pre-activation -> buf1;  GELU'/ReLU'/residual:
    // the input tile of unit u lands in buf[u] (TMA load, prefetched one unit ahead) and is overwritten IN PLACE by the
    // unit's output (each lane has read its own 16-byte chunks before it writes them)
    const uint32_t buf0 = stg, buf1 = stg + 4096;
    constexpr bool f32 = F32;
    const int epi_flags = g.epilogue & MASK;
    const bool atomic = (epi_flags & UC_EPI_ATOMIC) != 0;
    const bool has_aux = (epi_flags & UC_EPI_GELU) != 0;
    const bool has_in = (epi_flags & (UC_EPI_GELU_BWD | UC_EPI_RELU_BWD | UC_EPI_RESIDUAL)) != 0;  // aux_in / residual tile, via TMA
    uint32_t ph_a = 0, ph_b = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    auto stage_and_store = [&](const CUtensorMap* tm, uint32_t buf, const uint32_t (&w)[32], int col, int row0, bool reduce) {
      if (lane == 0) tma_store_wait_read0();  // the previous store issued by this lane no longer reads the staging tiles
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf + lane * 128 + ((j ^ (lane & 7)) << 4)), "r"(w[4 * j]),
                     "r"(w[4 * j + 1]), "r"(w[4 * j + 2]), "r"(w[4 * j + 3])
                     : "memory");
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (reduce) tma_reduce_add_2d(tm, buf, col, row0);
        else tma_store_2d(tm, buf, col, row0);
        tma_store_commit();
      }
    };
    for (int qn = 0;; ++qn) {
      const int item = next_item(qn);
      if (item >= total) break;
      const int split = item % g.split_k;
      const int tile = item / g.split_k;
      const int m0 = (g.n_fastest ? (tile / g.num_n) : (tile % g.num_m)) * 256 + 128 * (int)rank;
      const int n0 = (g.n_fastest ? (tile % g.num_n) : (tile / g.num_m)) * BN;
      const int kb0 = split * g.kb_per_split;
      const int kb1 = min(g.num_kb, kb0 + g.kb_per_split);
      if (kb0 >= kb1) continue;
      const int row0 = m0 + lane_group * 32;
      int row = row0 + lane;
      bool row_ok = row < g.m;
      int cxs = 0, cys = 0, cimg = 0;  // CONV 1 / 3: pixel coordinates of this warp's 32-row group (its C / aux boxes)
      if (CONV == 1 || CONV == 3) {
        const int tm = g.n_fastest ? (tile / g.num_n) : (tile % g.num_m);
        const int per_img = g.cv_tiles_x * g.cv_tiles_y;
        cimg = tm / per_img;
        const int r = tm - cimg * per_img;
        const int ty = r / g.cv_tiles_x;
        const int x0 = (r - ty * g.cv_tiles_x) << g.cv_tw_log2;
        const int y0 = (2 * ty + (int)rank) * (128 >> g.cv_tw_log2);
        const int tw_mask = (1 << g.cv_tw_log2) - 1;
        const int rl0 = lane_group * 32, rl = rl0 + lane;
        cxs = x0 + (rl0 & tw_mask);
        cys = y0 + (rl0 >> g.cv_tw_log2);
        row_ok = (x0 + (rl & tw_mask)) < g.cv_W && (y0 + (rl >> g.cv_tw_log2)) < g.cv_H;
        row = 0;  // no row-addressed global access in this mode (aux tiles are staged, bias is per column)
      }
      auto load_aux = [&](uint32_t buf, uint32_t bar, int col) {
        if (CONV == 1 || CONV == 3) tma_load_4d(buf, &tmAux, bar, col, cxs, cys, cimg);
        else tma_load_2d(buf, &tmAux, bar, col, row0);
      };
      auto store_c = [&](uint32_t buf, int col) {
        if (CONV == 1 || CONV == 3) tma_store_4d(&tmC, buf, col, cxs, cys, cimg);
        else tma_store_2d(&tmC, buf, col, row0);
      };
      const int nw = n0 + col_half * (BN / 2);  // first column of this warp's half of the tile
      if (!f32 && has_in) {
        // input tile of unit 0 ([32 rows x 64 cols] bf16 of aux_in / residual) requested BEFORE waiting for the accumulator
        if (lane == 0) {
          if (NCH == 4) tma_store_wait_read1();  // the previous tile's unit-0 store has finished reading buf0
          else tma_store_wait_read0();
          mbar_arrive_expect_tx(ld_bar(e, 0), 4096);
          load_aux(buf0, ld_bar(e, 0), nw);
        }
        __syncwarp();
      }
      // Everything the epilogue math reads from global memory is fetched BEFORE the wait for the accumulator, while this warp
      // has nothing else to do: the row's RoPE positions, and one 16-byte piece per lane of the warp's 128 bias values and of
      // the row's two RoPE table lines, which pulls those lines into L1 -- epilogue_math's own loads then hit L1 (35 clk) instead
      // of L2 (300 clk) on the critical path of the first chunk.  (launch-latency trace: ~0.8 us per 32-column chunk before.)
      int pos_y = 0, pos_x = 0;
      if ((epi_flags & UC_EPI_ROPE) && row_ok) {
        pos_y = g.positions[2 * row];
        pos_x = g.positions[2 * row + 1];
      }
      if (epi_flags & UC_EPI_BIAS) {
        const int nb = nw + 4 * lane;
        if (nb < g.n) {
          float4 warm = __ldg(reinterpret_cast<const float4*>(g.bias + nb));
          asm volatile("" ::"f"(warm.x), "f"(warm.y), "f"(warm.z), "f"(warm.w));
        }
      }
      if ((epi_flags & UC_EPI_ROPE) && row_ok && nw < g.rope_cols) {
        float4 wy = __ldg(reinterpret_cast<const float4*>(g.rope_table + (size_t)pos_y * 32));
        float4 wx = __ldg(reinterpret_cast<const float4*>(g.rope_table + (size_t)pos_x * 32));
        asm volatile("" ::"f"(wy.x), "f"(wx.x));
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (trc && qn == 0 && warp == 2 && lane == 0) trc[4] = globaltimer_ns();
      const uint32_t tbase = tmem_base + (uint32_t(lane_group * 32) << 16) + uint32_t(acc * BN + col_half * (BN / 2));
      auto release_acc = [&]() {  // this warp has read all of its accumulator columns
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) mbar_arrive(tempty_bar(acc));
          else mbar_arrive_cluster(tempty_bar(acc), 0);
        }
      };
      if (f32) {
        // NCH units of 32 fp32 columns (128 B rows)
#pragma unroll 1
        for (int u = 0; u < NCH; ++u) {
          uint32_t r[32];
          __syncwarp();
          tmem_ld32(tbase + u * 32, r);
          tmem_ld_wait();
          const int n = nw + u * 32;
          float v[32];
          uint32_t prep[16];
          epilogue_math<MASK>(g, row, n, pos_y, pos_x, row_ok, r, v, prep);
          uint32_t w[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) w[j] = __float_as_uint(v[j]);
          stage_and_store(&tmC, buf0, w, n, row0, atomic);
        }
        release_acc();
      } else {
        // bf16 out: 4 chunks of 32 columns = 2 staged units of 64 columns (128 B rows), software-pipelined: the
        // TMEM load of chunk c+1 and the TMA load of the other unit's input tile are in flight behind chunk c's math
        uint32_t ra[32], rb[32];
        __syncwarp();
        tmem_ld32(tbase, ra);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int u = c >> 1, hf = c & 1;
          uint32_t(&rc)[32] = (c & 1) ? rb : ra;
          uint32_t(&rn)[32] = (c & 1) ? ra : rb;
          tmem_ld_wait();
          if (trc && qn == 0 && warp == 2 && lane == 0) trc[7 + 2 * c] = globaltimer_ns();  // chunk c in registers
          if (c + 1 < NCH) tmem_ld32(tbase + 32 * (c + 1), rn);
          else release_acc();
          const int n = nw + c * 32;
          float v[32], h[32];
          uint32_t prep[16];
          if (has_in) {
            if (hf == 0) {
              if (u == 0) {
                if (NCH == 4 && lane == 0) {
                  tma_store_wait_read0();  // the previous tile's unit-1 store has finished reading buf1
                  mbar_arrive_expect_tx(ld_bar(e, 1), 4096);
                  load_aux(buf1, ld_bar(e, 1), nw + 64);
                }
                mbar_wait(ld_bar(e, 0), ph_a);
                ph_a ^= 1u;
              } else {
                mbar_wait(ld_bar(e, 1), ph_b);
                ph_b ^= 1u;
              }
            }
            const uint32_t inb = u ? buf1 : buf0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t a0, a1, a2, a3;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                           : "r"(inb + lane * 128 + (((hf * 4 + j) ^ (lane & 7)) << 4)));
              h[8 * j + 0] = bf16_lo(a0); h[8 * j + 1] = bf16_hi(a0); h[8 * j + 2] = bf16_lo(a1); h[8 * j + 3] = bf16_hi(a1);
              h[8 * j + 4] = bf16_lo(a2); h[8 * j + 5] = bf16_hi(a2); h[8 * j + 6] = bf16_lo(a3); h[8 * j + 7] = bf16_hi(a3);
            }
            epilogue_math<MASK>(g, row, n, pos_y, pos_x, row_ok, rc, v, prep, h);
          } else {
            epilogue_math<MASK>(g, row, n, pos_y, pos_x, row_ok, rc, v, prep);
          }
          const uint32_t outb = (has_aux || u == 0) ? buf0 : buf1;
          if (hf == 0 && !has_in) {  // staging tile free once the bulk store that last read it has completed
            if (lane == 0) {
              if (has_aux || NCH != 4) tma_store_wait_read0();
              else tma_store_wait_read1();
            }
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t so = lane * 128 + (((hf * 4 + j) ^ (lane & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(outb + so), "r"(pack_bf16(v[8 * j], v[8 * j + 1])),
                         "r"(pack_bf16(v[8 * j + 2], v[8 * j + 3])), "r"(pack_bf16(v[8 * j + 4], v[8 * j + 5])),
                         "r"(pack_bf16(v[8 * j + 6], v[8 * j + 7]))
                         : "memory");
            if (has_aux)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf1 + so), "r"(prep[4 * j]), "r"(prep[4 * j + 1]),
                           "r"(prep[4 * j + 2]), "r"(prep[4 * j + 3])
                           : "memory");
          }
          if (trc && qn == 0 && warp == 2 && lane == 0) trc[8 + 2 * c] = globaltimer_ns();  // chunk c staged
          if (hf == 1) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (has_aux) tma_store_2d(&tmAux, buf1, n - 32, row0);
              store_c(outb, n - 32);
              tma_store_commit();
            }
            if (g.colsum) {
              // column sums of the staged [32 rows x 64 cols] bf16 tile: lane <-> columns 2*lane, 2*lane+1 (conflict-free:
              // the 128B swizzle spreads the 8 16-byte chunks of a row over all banks)
              float s0 = 0.f, s1 = 0.f;
              const int rows_valid = min(32, g.m - row0);
#pragma unroll 8
              for (int rr = 0; rr < rows_valid; ++rr) {
                uint32_t w2;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w2)
                             : "r"(outb + rr * 128 + ((uint32_t((lane >> 2) ^ (rr & 7))) << 4) + (lane & 3) * 4));
                s0 += bf16_lo(w2);
                s1 += bf16_hi(w2);
              }
              const int nc = n - 32 + 2 * lane;
              if (nc < g.n) {
                atomicAdd(g.colsum + nc, s0);
                atomicAdd(g.colsum + nc + 1, s1);
              }
            }
          }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (trc && warp == 2 && lane == 0) trc[15] = globaltimer_ns();  // tile loop left
    if (lane == 0) tma_store_wait0();  // all bulk stores of this lane have completed before the CTA may exit
    if (trc && warp == 2 && lane == 0) trc[5] = globaltimer_ns();
  }

  tc_fence_before();
  cluster_sync_all();  // nobody exits (or frees TMEM) while the peer can still touch this CTA's smem / barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
  if (trc && threadIdx.x == 0) trc[6] = globaltimer_ns();
}

template <int MASK, bool F32, int BN, int CONV = 0>
int launch2_inst(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmAux, const GemmArgs& g,
                 int grid, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_kernel<MASK, F32, BN, CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<BN>::SMEM);
    UC_REQUIRE(e == cudaSuccess, UC_ERR_CUDA, "uc_gemm: cudaFuncSetAttribute(gemm2) failed: %s", cudaGetErrorString(e));
    configured = true;
  }
  cudaError_t le = launch_pdl(gemm2_kernel<MASK, F32, BN, CONV>, dim3(grid), dim3(GEMM_THREADS), Cfg2<BN>::SMEM, stream, tmA, tmB, tmC, tmAux, g);
  UC_REQUIRE(le == cudaSuccess, UC_ERR_CUDA, "uc_gemm(cta_pair): launch failed: %s", cudaGetErrorString(le));
  return check_launch("uc_gemm(cta_pair)");
}

template <int BN>
int launch2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmAux, const GemmArgs& g,
            int grid, cudaStream_t stream) {
  constexpr int kGeneric = UC_EPI_BIAS | UC_EPI_ROPE | UC_EPI_RESIDUAL | UC_EPI_RELU | UC_EPI_RELU_BWD;
  const int e = g.epilogue;
  if (g.c_f32) {
    if (e == UC_EPI_ATOMIC) return launch2_inst<UC_EPI_ATOMIC, true, BN>(tmA, tmB, tmC, tmAux, g, grid, stream);
    if ((e & ~kGeneric) == 0) return launch2_inst<kGeneric, true, BN>(tmA, tmB, tmC, tmAux, g, grid, stream);
    return launch2_inst<-1, true, BN>(tmA, tmB, tmC, tmAux, g, grid, stream);
  }
  if ((e & ~(UC_EPI_BIAS | UC_EPI_GELU)) == 0 && (e & UC_EPI_GELU))
    return launch2_inst<UC_EPI_BIAS | UC_EPI_GELU, false, BN>(tmA, tmB, tmC, tmAux, g, grid, stream);
  if (e == UC_EPI_GELU_BWD) return launch2_inst<UC_EPI_GELU_BWD, false, BN>(tmA, tmB, tmC, tmAux, g, grid, stream);
  // every dgrad (no epilogue at all) and the plain Linear forwards: no run-time flag tests, no aux-tile path in the instance
  if ((e & ~UC_EPI_BIAS) == 0) return launch2_inst<UC_EPI_BIAS, false, BN>(tmA, tmB, tmC, tmAux, g, grid, stream);
  if ((e & ~kGeneric) == 0) return launch2_inst<kGeneric, false, BN>(tmA, tmB, tmC, tmAux, g, grid, stream);
  return launch2_inst<-1, false, BN>(tmA, tmB, tmC, tmAux, g, grid, stream);
}

template <int BN>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& g, int grid, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM);
    UC_REQUIRE(e == cudaSuccess, UC_ERR_CUDA, "uc_gemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    configured = true;
  }
  cudaError_t le = launch_pdl(gemm_kernel<BN>, dim3(grid), dim3(GEMM_THREADS), Cfg<BN>::SMEM, stream, tmA, tmB, g);
  UC_REQUIRE(le == cudaSuccess, UC_ERR_CUDA, "uc_gemm: launch failed: %s", cudaGetErrorString(le));
  return check_launch("uc_gemm");
}

}  // namespace

}  // namespace uc

// Tile distribution of the pair GEMM.  0 (default): static round robin -- fastest when the GEMM owns the GPU.  1: atomic tile
// queue -- a cluster whose SMs are held by somebody else's CTAs (an NCCL all-reduce overlapped with the backward pass, a kernel
// of another stream) takes fewer tiles instead of forcing a second wave.  Measured at N = 2 (profiles/r02c): no gain over the static split, so nothing switches it on by default.
static int g_gemm_dynamic = [] { const char* e = getenv("UC_GEMM_DYNAMIC"); return e ? atoi(e) : 0; }();
extern "C" __attribute__((visibility("default"))) int uc_debug_set_gemm_trace(unsigned long long* buf) {
  return cudaMemcpyToSymbol(uc::g_gemm_trace, &buf, sizeof(buf)) == cudaSuccess ? 0 : UC_ERR_CUDA;
}
extern "C" int uc_set_gemm_dynamic(int on) {
  const int prev = g_gemm_dynamic;
  g_gemm_dynamic = on ? 1 : 0;
  return prev;
}

extern "C" int uc_gemm(const uc_gemm_params* p, uc_stream_t stream_) {
  using namespace uc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  UC_REQUIRE(p && p->a && p->b && p->c, UC_ERR_BAD_SHAPE, "uc_gemm: null operand");
  UC_REQUIRE(p->m > 0 && p->n > 0 && p->k > 0, UC_ERR_BAD_SHAPE, "uc_gemm: bad sizes m=%d n=%d k=%d", p->m, p->n, p->k);
  UC_REQUIRE(p->n % 64 == 0, UC_ERR_BAD_SHAPE, "uc_gemm: n=%d must be a multiple of 64", p->n);
  UC_REQUIRE(p->lda % 8 == 0 && p->ldb % 8 == 0 && p->ldc % 8 == 0, UC_ERR_BAD_SHAPE,
             "uc_gemm: leading dimensions must be multiples of 8 elements (lda=%lld ldb=%lld ldc=%lld)", (long long)p->lda,
             (long long)p->ldb, (long long)p->ldc);
  UC_REQUIRE(((uintptr_t)p->a % 16 == 0) && ((uintptr_t)p->b % 16 == 0) && ((uintptr_t)p->c % 16 == 0), UC_ERR_BAD_SHAPE,
             "uc_gemm: operands must be 16-byte aligned");
  UC_REQUIRE(p->c_dtype == UC_DTYPE_BF16 || p->c_dtype == UC_DTYPE_F32, UC_ERR_BAD_DTYPE, "uc_gemm: c_dtype %d", p->c_dtype);
  const int epi = p->epilogue;
  UC_REQUIRE(!(epi & UC_EPI_BIAS) || p->bias, UC_ERR_BAD_SHAPE, "uc_gemm: UC_EPI_BIAS without bias");
  UC_REQUIRE(!(epi & UC_EPI_RESIDUAL) || p->residual, UC_ERR_BAD_SHAPE, "uc_gemm: UC_EPI_RESIDUAL without residual");
  UC_REQUIRE(!(epi & UC_EPI_GELU) || (p->aux_out && p->c_dtype == UC_DTYPE_BF16), UC_ERR_BAD_SHAPE,
             "uc_gemm: UC_EPI_GELU needs aux_out and bf16 C");
  UC_REQUIRE(!(epi & (UC_EPI_GELU_BWD | UC_EPI_RELU_BWD)) || p->aux_in, UC_ERR_BAD_SHAPE, "uc_gemm: GELU_BWD / RELU_BWD without aux_in");
  {
    const int users = ((epi & UC_EPI_GELU) ? 1 : 0) + ((epi & UC_EPI_GELU_BWD) ? 1 : 0) + ((epi & UC_EPI_RELU_BWD) ? 1 : 0) +
                      ((epi & UC_EPI_RESIDUAL) ? 1 : 0);
    UC_REQUIRE(users <= 1, UC_ERR_UNSUPPORTED, "uc_gemm: GELU / GELU_BWD / RELU_BWD / RESIDUAL epilogues are mutually exclusive");
  }
  UC_REQUIRE(!(epi & UC_EPI_ROPE) || (p->positions && p->rope_table && p->rope_cols % 64 == 0), UC_ERR_BAD_SHAPE,
             "uc_gemm: UC_EPI_ROPE needs positions, rope_table and rope_cols %% 64 == 0");
  UC_REQUIRE(!(epi & UC_EPI_ATOMIC) || p->c_dtype == UC_DTYPE_F32, UC_ERR_BAD_DTYPE, "uc_gemm: atomic epilogue needs fp32 C");
  UC_REQUIRE(!(epi & UC_EPI_RESIDUAL_F32) || ((epi & UC_EPI_RESIDUAL) && p->c_dtype == UC_DTYPE_F32), UC_ERR_BAD_DTYPE,
             "uc_gemm: UC_EPI_RESIDUAL_F32 needs UC_EPI_RESIDUAL and fp32 C");
  UC_REQUIRE(!p->c_colsum || (p->c_dtype == UC_DTYPE_BF16 && p->n % 8 == 0), UC_ERR_BAD_DTYPE, "uc_gemm: c_colsum needs bf16 C");

  const int num_kb = (p->k + BK - 1) / BK;
  const int sms = sm_count();
  // tile shape: maximise (wave efficiency) x (per-tile efficiency).  "pair" = 256x256 per CTA pair (cta_group::2):
  // a third less L2->SM operand traffic per FLOP; narrow single-CTA tiles are smem-bandwidth bound.
  const bool atomic = (epi & UC_EPI_ATOMIC) != 0;
  static const int pair_env = [] { const char* e = getenv("UC_GEMM_PAIR"); return e ? atoi(e) : -1; }();
  int bn = 0;
  bool pair = false;
  {
    double best = -1.0;
    // measured on B200 (profiles/r01h_gemm_pair_vs_single.log): with its staged, specialised epilogue and PDL the pair
    // kernel wins or ties on every shape of the path, including K = 768 (8192x3072x768: 33 us vs 48 us single-CTA;
    // 8192x768x768: 16.3 vs 17.1), so it is preferred unless its wave quantisation is much worse
    const double kf = num_kb <= 8 ? 0.0 : (num_kb >= 40 ? 1.0 : double(num_kb - 8) / 32.0);
    // (256 x 128 pair tiles were tried for the n = 768 shapes -- 3 waves of half tiles instead of 2 full ones -- and lost:
    // 102.6 vs 114.0 pairs/s when preferred; the kernel template still takes BN, only BN = 256 is instantiated.)
    const int cand[4] = {256, 128, 64, 256};
    const double rate[4] = {1.0, 0.9, 0.6, 1.3 + 0.1 * kf};
    for (int i = 0; i < 4; ++i) {
      const bool is_pair = (i == 3);
      if (is_pair && pair_env == 0) continue;
      if (p->n % cand[i] != 0) continue;
      const int bm = is_pair ? 256 : BM;
      const int slots = is_pair ? sms / 2 : sms;
      const long long tiles = (long long)((p->m + bm - 1) / bm) * (p->n / cand[i]);
      const long long waves = (tiles + slots - 1) / slots;
      // split-K (wgrad) fills the machine by itself, so only the per-tile efficiency matters there
      double eff = (atomic ? 1.0 : double(tiles) / double(waves * slots)) * rate[i];
      // the GELU / GELU' epilogues are issue-bound: only the pair kernel's staged, software-pipelined, specialised
      // epilogue keeps up with the MMA pipe (decoder fc1 at K=768: 86 us single-CTA vs 51 us)
      if (is_pair && (epi & (UC_EPI_GELU | UC_EPI_GELU_BWD))) eff *= 1.5;
      if (is_pair && pair_env == 1) eff = 10.0;
      if (eff > best) { best = eff; bn = cand[i]; pair = is_pair; }
    }
  }
  UC_REQUIRE(bn != 0, UC_ERR_BAD_SHAPE, "uc_gemm: n=%d must be a multiple of 64", p->n);
  const int bm_eff = pair ? 256 : BM;
  const int slots = pair ? sms / 2 : sms;
  const int num_m = (p->m + bm_eff - 1) / bm_eff;
  const int num_n = p->n / bn;
  int split_k = p->split_k;
  if (split_k <= 0) {
    split_k = 1;
    if (atomic) {
      // wgrad: pick the split that fills whole waves (768x768: 9 tiles x 8 splits = 72 of 74 cluster slots in ONE wave;
      // rounding up to 9 splits made it 81 items = two half-empty waves)
      const long long tiles = (long long)num_m * num_n;
      const int max_split = num_kb / 4 < 1 ? 1 : num_kb / 4;
      double best_u = -1.0;
      for (int sk = 1; sk <= max_split && sk <= 64; ++sk) {
        const long long items = tiles * sk;
        const long long waves = (items + slots - 1) / slots;
        const double u = double(items) / double(waves * slots) - 0.002 * sk;  // ties -> fewer splits (less reduction traffic)
        if (u > best_u) { best_u = u; split_k = sk; }
        if (items >= 4LL * slots) break;
      }
    }
  }
  UC_REQUIRE(split_k == 1 || ((epi & UC_EPI_ATOMIC) && epi == UC_EPI_ATOMIC), UC_ERR_BAD_SHAPE,
             "uc_gemm: split_k > 1 requires the pure UC_EPI_ATOMIC epilogue");
  if (split_k > num_kb) split_k = num_kb;
  int kb_per = (num_kb + split_k - 1) / split_k;
  split_k = (num_kb + kb_per - 1) / kb_per;

  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2], strides[1];
    uint32_t box[2];
    if (!p->a_layout) { dims[0] = p->k; dims[1] = p->m; box[0] = BK; box[1] = BM; }
    else { dims[0] = p->m; dims[1] = p->k; box[0] = 64; box[1] = BK; }
    strides[0] = (uint64_t)p->lda * 2;
    int r = make_tensor_map(&tmA, p->a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (r) return r;
    if (!p->b_layout) { dims[0] = p->k; dims[1] = p->n; box[0] = BK; box[1] = pair ? bn / 2 : bn; }
    else { dims[0] = p->n; dims[1] = p->k; box[0] = 64; box[1] = BK; }
    strides[0] = (uint64_t)p->ldb * 2;
    r = make_tensor_map(&tmB, p->b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (r) return r;
  }

  GemmArgs g;
  g.m = p->m; g.n = p->n; g.k = p->k;
  g.a_mn = p->a_layout ? 1 : 0;
  g.b_mn = p->b_layout ? 1 : 0;
  g.num_m = num_m; g.num_n = num_n; g.num_kb = num_kb; g.split_k = split_k; g.kb_per_split = kb_per;
  // rasterisation: tiles that run concurrently share the LARGER operand's blocks (read once from HBM);
  // the smaller operand is re-read from L2.  wgrad (atomic) keeps m-fastest: consecutive tiles share the B block.
  g.n_fastest = (!atomic && (long long)p->m >= (long long)p->n) ? 1 : 0;
  g.epilogue = epi; g.c_f32 = p->c_dtype == UC_DTYPE_F32; g.rope_cols = p->rope_cols;
  g.ldc = p->ldc; g.c = p->c; g.bias = p->bias;
  g.residual = static_cast<const __nv_bfloat16*>(p->residual);
  g.aux_out = static_cast<__nv_bfloat16*>(p->aux_out);
  g.aux_in = static_cast<const __nv_bfloat16*>(p->aux_in);
  g.positions = p->positions; g.rope_table = p->rope_table;
  g.colsum = (pair && p->c_dtype == UC_DTYPE_BF16) ? p->c_colsum : nullptr;
  g.work = nullptr;
  g.cv_H = g.cv_W = g.cv_tw_log2 = g.cv_tiles_x = g.cv_tiles_y = g.cv_kb_per_tap = g.cv_b_tap_stride = g.cv_cin = 0;

  const long long total = (long long)num_m * num_n * split_k;
  if (pair) {
    // output tiles leave through TMA: [32 rows x 128 B] boxes, 128B swizzle (bf16: 64 columns, fp32: 32 columns)
    CUtensorMap tmC, tmAux;
    const bool f32 = p->c_dtype == UC_DTYPE_F32;
    uint64_t dims[2] = {(uint64_t)p->n, (uint64_t)p->m};
    uint64_t strides[1] = {(uint64_t)p->ldc * (f32 ? 4 : 2)};
    uint32_t box[2] = {f32 ? 32u : 64u, 32u};
    int r = make_tensor_map(&tmC, p->c, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, strides, box,
                            CU_TENSOR_MAP_SWIZZLE_128B);
    if (r) return r;
    tmAux = tmC;
    const void* aux_ptr = (epi & UC_EPI_GELU) ? p->aux_out : (epi & (UC_EPI_GELU_BWD | UC_EPI_RELU_BWD)) ? p->aux_in
                          : (epi & UC_EPI_RESIDUAL) ? p->residual : nullptr;
    if (aux_ptr && !f32) {
      r = make_tensor_map(&tmAux, aux_ptr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
      if (r) return r;
    }
    const int clusters = (int)(total < slots ? total : slots);
    if (g_gemm_dynamic) {
      g.work = work_slot();
      UC_REQUIRE(g.work, UC_ERR_CUDA, "uc_gemm: work counters unavailable");
    }
    return launch2<256>(tmA, tmB, tmC, tmAux, g, 2 * clusters, stream);
  }
  const int grid = (int)(total < sms ? total : sms);
  int r = bn == 256 ? launch<256>(tmA, tmB, g, grid, stream) : bn == 128 ? launch<128>(tmA, tmB, g, grid, stream)
                                                                         : launch<64>(tmA, tmB, g, grid, stream);
  if (r == 0 && p->c_colsum)  // single-CTA kernels have no staged tile: one extra pass over C
    r = uc_colsum(p->c, UC_DTYPE_BF16, p->ldc, p->m, p->n, p->c_colsum, stream_);
  return r;
}

// ------------------------------------------------------------------------------------------------
// uc_conv3x3: 3x3 / stride 1 / pad 1 convolution on NHWC bf16 maps as an IMPLICIT GEMM on the CTA-pair kernel above -- no
// im2col buffer exists: the 9 taps are 9 shifted 4-D TMA boxes of the activation map accumulating into the same TMEM tile.
// ------------------------------------------------------------------------------------------------
namespace {
// window width (power of two, lo..hi) of area `area` pixels that wastes the fewest pixels on an H x W map;
// rows_per_tile = how many window heights one scheduling unit spans (2 for the CTA pair's 256-row tile)
int best_window_log2(int H, int W, int area, int rows_mult, int lo, int hi) {
  int best = lo;
  long long best_cost = -1;
  for (int l = lo; l <= hi; ++l) {
    const int tw = 1 << l, th = (area / tw) * rows_mult;
    if (area % tw) continue;
    const long long cost = (long long)((W + tw - 1) / tw) * tw * ((H + th - 1) / th) * th;
    if (best_cost < 0 || cost < best_cost || (cost == best_cost && tw <= 32)) { best_cost = cost; best = l; }
  }
  return best;
}
int act_map(CUtensorMap* tm, const void* base, int B, int H, int W, int C, int bw, int bh) {
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  uint32_t box[4] = {64u, (uint32_t)bw, (uint32_t)bh, 1u};
  return uc::make_tensor_map(tm, base, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}
}  // namespace

extern "C" int uc_conv3x3(const uc_conv3x3_params* p, uc_stream_t stream_) {
  using namespace uc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  UC_REQUIRE(p, UC_ERR_BAD_SHAPE, "uc_conv3x3: null params");
  UC_REQUIRE(p->mode >= 0 && p->mode <= 2, UC_ERR_BAD_SHAPE, "uc_conv3x3: mode %d", p->mode);
  UC_REQUIRE(p->B > 0 && p->H > 0 && p->W > 0, UC_ERR_BAD_SHAPE, "uc_conv3x3: bad map size B=%d H=%d W=%d", p->B, p->H, p->W);
  UC_REQUIRE(p->cin > 0 && p->cout > 0 && p->cin % 64 == 0 && p->cout % 64 == 0, UC_ERR_BAD_SHAPE,
             "uc_conv3x3: channel counts must be multiples of 64 (cin=%d cout=%d)", p->cin, p->cout);
  UC_REQUIRE(p->w || p->mode == 2, UC_ERR_BAD_SHAPE, "uc_conv3x3: null weight");
  const int sms = sm_count();
  const int slots = sms / 2;
  GemmArgs g{};
  g.cv_H = p->H; g.cv_W = p->W;
  g.bias = nullptr; g.residual = nullptr; g.aux_out = nullptr; g.aux_in = nullptr; g.positions = nullptr; g.rope_table = nullptr;
  g.colsum = nullptr; g.work = nullptr; g.rope_cols = 0;
  CUtensorMap tmA, tmB, tmC, tmAux;
  int r;
  if (p->mode == 2) {
    // dw[co, tap * cin + ci] += sum_pixels dy[pixel, co] * x[pixel + tap offset, ci]   (fp32, split over the pixels)
    UC_REQUIRE(p->x && p->dy && p->dw, UC_ERR_BAD_SHAPE, "uc_conv3x3(wgrad): x, dy and dw are required");
    const int l2 = best_window_log2(p->H, p->W, 64, 1, 1, 6);
    const int tw = 1 << l2, th = 64 / tw;
    g.cv_tw_log2 = l2; g.cv_tiles_x = (p->W + tw - 1) / tw; g.cv_tiles_y = (p->H + th - 1) / th; g.cv_cin = p->cin;
    g.m = p->cout; g.n = 9 * p->cin;
    g.num_kb = p->B * g.cv_tiles_x * g.cv_tiles_y;
    g.k = g.num_kb * 64;
    g.a_mn = 1; g.b_mn = 1;
    g.num_m = (g.m + 255) / 256; g.num_n = (g.n + 255) / 256;
    const long long tiles = (long long)g.num_m * g.num_n;
    int split_k = 1;
    {
      const int max_split = g.num_kb / 4 < 1 ? 1 : g.num_kb / 4;
      double best_u = -1.0;
      for (int sk = 1; sk <= max_split && sk <= 64; ++sk) {
        const long long items = tiles * sk;
        const long long waves = (items + slots - 1) / slots;
        const double u = double(items) / double(waves * slots) - 0.002 * sk;
        if (u > best_u) { best_u = u; split_k = sk; }
        if (items >= 4LL * slots) break;
      }
    }
    int kb_per = (g.num_kb + split_k - 1) / split_k;
    split_k = (g.num_kb + kb_per - 1) / kb_per;
    g.split_k = split_k; g.kb_per_split = kb_per;
    g.n_fastest = 0;
    g.epilogue = UC_EPI_ATOMIC; g.c_f32 = 1; g.ldc = 9LL * p->cin; g.c = p->dw;
    if ((r = act_map(&tmA, p->dy, p->B, p->H, p->W, p->cout, tw, th))) return r;
    if ((r = act_map(&tmB, p->x, p->B, p->H, p->W, p->cin, tw, th))) return r;
    uint64_t dims[2] = {(uint64_t)g.n, (uint64_t)g.m};
    uint64_t strides[1] = {(uint64_t)g.ldc * 4};
    uint32_t box[2] = {32u, 32u};
    if ((r = make_tensor_map(&tmC, p->dw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return r;
    tmAux = tmC;
    const long long total = tiles * split_k;
    const int clusters = (int)(total < slots ? total : slots);
    return launch2_inst<UC_EPI_ATOMIC, true, 256, 2>(tmA, tmB, tmC, tmAux, g, 2 * clusters, stream);
  }
  // fwd: y = conv(x, w) [+ bias] [relu | + residual]            A = x  (cin),  n = cout, B = w  K-major [cout, 9 cin]
  // dgrad: dx = conv(dy, flipped w^T) [* (relu_out > 0)]        A = dy (cout), n = cin,  B = w  MN-major view [cout, 9 cin]
  const bool fwd = p->mode == 0;
  const void* a_ptr = fwd ? p->x : p->dy;
  void* c_ptr = fwd ? p->y : p->dx;
  UC_REQUIRE(a_ptr && c_ptr, UC_ERR_BAD_SHAPE, "uc_conv3x3: null activation pointer");
  const int ca = fwd ? p->cin : p->cout;  // channels of the A map
  const int n = fwd ? p->cout : p->cin;
  const int l2 = best_window_log2(p->H, p->W, 128, 2, 3, 7);
  const int tw = 1 << l2, th = 128 / tw;
  g.cv_tw_log2 = l2; g.cv_tiles_x = (p->W + tw - 1) / tw; g.cv_tiles_y = (p->H + 2 * th - 1) / (2 * th);
  g.cv_kb_per_tap = ca / 64;
  g.cv_b_tap_stride = fwd ? 0 : p->cin;
  g.m = p->B * p->H * p->W; g.n = n; g.k = 9 * ca;
  g.a_mn = 0; g.b_mn = fwd ? 0 : 1;
  const int bn = (n % 256 == 0 || n > 384) ? 256 : 128;
  g.num_m = p->B * g.cv_tiles_x * g.cv_tiles_y;
  g.num_n = (n + bn - 1) / bn;
  g.num_kb = 9 * g.cv_kb_per_tap; g.split_k = 1; g.kb_per_split = g.num_kb;
  g.n_fastest = 1;
  int epi = 0;
  if (fwd) {
    UC_REQUIRE(!(p->relu && p->residual), UC_ERR_UNSUPPORTED, "uc_conv3x3: fused ReLU and residual are mutually exclusive");
    if (p->bias) { epi |= UC_EPI_BIAS; g.bias = p->bias; }
    if (p->relu) epi |= UC_EPI_RELU;
    if (p->residual) { epi |= UC_EPI_RESIDUAL; g.residual = static_cast<const __nv_bfloat16*>(p->residual); }
  } else if (p->relu_out) {
    epi |= UC_EPI_RELU_BWD; g.aux_in = static_cast<const __nv_bfloat16*>(p->relu_out);
  }
  g.epilogue = epi; g.c_f32 = 0; g.ldc = n; g.c = c_ptr;
  if ((r = act_map(&tmA, a_ptr, p->B, p->H, p->W, ca, tw, th))) return r;
  {
    uint64_t dims[2], strides[1] = {(uint64_t)9 * p->cin * 2};
    uint32_t box[2];
    if (fwd) { dims[0] = 9ull * p->cin; dims[1] = p->cout; box[0] = BK; box[1] = bn / 2; }
    else { dims[0] = 9ull * p->cin; dims[1] = p->cout; box[0] = 64; box[1] = BK; }
    if ((r = make_tensor_map(&tmB, p->w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return r;
  }
  // C / aux tiles: one epilogue warp's 32 rows = 32 pixels of a row (tw >= 32) or 32 / tw rows of tw pixels
  const int bw = tw >= 32 ? 32 : tw, bh = 32 / bw;
  if ((r = act_map(&tmC, c_ptr, p->B, p->H, p->W, n, bw, bh))) return r;
  tmAux = tmC;
  const void* aux_ptr = fwd ? p->residual : p->relu_out;
  if (aux_ptr && (r = act_map(&tmAux, aux_ptr, p->B, p->H, p->W, n, bw, bh))) return r;
  const long long total = (long long)g.num_m * g.num_n;
  const int clusters = (int)(total < slots ? total : slots);
  constexpr int kConvMask = UC_EPI_BIAS | UC_EPI_RESIDUAL | UC_EPI_RELU | UC_EPI_RELU_BWD;
  if (bn == 256) return launch2_inst<kConvMask, false, 256, 1>(tmA, tmB, tmC, tmAux, g, 2 * clusters, stream);
  return launch2_inst<kConvMask, false, 128, 1>(tmA, tmB, tmC, tmAux, g, 2 * clusters, stream);
}

// ------------------------------------------------------------------------------------------------
// uc_patch_embed: Conv2d(3, C, kernel = stride = p) + bias on the fp32 NCHW image, tokens out (bf16 [B * Hp * Wp, C]),
// without materialising the patch columns: 5-D TMA boxes of the image feed TF32 MMAs on the fp32 master weights.
// ------------------------------------------------------------------------------------------------
extern "C" int uc_patch_embed(const uc_patch_embed_params* p, uc_stream_t stream_) {
  using namespace uc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  UC_REQUIRE(p && p->img && p->w && p->out, UC_ERR_BAD_SHAPE, "uc_patch_embed: null pointer");
  UC_REQUIRE(p->patch == 16 || p->patch == 32, UC_ERR_UNSUPPORTED,
             "uc_patch_embed: patch %d (a patch row must be a 64- or 128-byte TMA box row: 16 or 32; use uc_patchify + uc_gemm)", p->patch);
  UC_REQUIRE(p->B > 0 && p->H > 0 && p->W > 0 && p->H % p->patch == 0 && p->W % p->patch == 0, UC_ERR_BAD_SHAPE,
             "uc_patch_embed: image %dx%d is not a multiple of the patch size %d", p->H, p->W, p->patch);
  UC_REQUIRE(p->n > 0 && p->n % 128 == 0, UC_ERR_BAD_SHAPE, "uc_patch_embed: n=%d must be a multiple of 128", p->n);
  UC_REQUIRE(((uintptr_t)p->img % 16 == 0) && ((uintptr_t)p->w % 16 == 0) && ((uintptr_t)p->out % 16 == 0), UC_ERR_BAD_SHAPE,
             "uc_patch_embed: pointers must be 16-byte aligned");
  const int ps = p->patch, Hp = p->H / ps, Wp = p->W / ps;
  const int sms = sm_count();
  const int slots = sms / 2;
  GemmArgs g{};
  g.cv_H = Hp; g.cv_W = Wp;
  const int l2 = best_window_log2(Hp, Wp, 128, 2, 3, 7);
  const int tw = 1 << l2, th = 128 / tw;
  g.cv_tw_log2 = l2; g.cv_tiles_x = (Wp + tw - 1) / tw; g.cv_tiles_y = (Hp + 2 * th - 1) / (2 * th);
  g.cv_kb_per_tap = ps * ps / 32; g.cv_cin = 32 / ps; g.cv_b_tap_stride = 0;
  g.m = p->B * Hp * Wp; g.n = p->n; g.k = 3 * ps * ps;
  g.a_mn = 0; g.b_mn = 0;
  const int bn = (p->n % 256 == 0) ? 256 : 128;
  g.num_m = p->B * g.cv_tiles_x * g.cv_tiles_y;
  g.num_n = p->n / bn;
  g.num_kb = 3 * g.cv_kb_per_tap; g.split_k = 1; g.kb_per_split = g.num_kb;
  g.n_fastest = 1;
  g.epilogue = p->bias ? UC_EPI_BIAS : 0; g.bias = p->bias;
  g.c_f32 = 0; g.ldc = p->n; g.c = p->out;
  CUtensorMap tmA, tmB, tmC, tmAux;
  int r;
  {
    uint64_t dims[5] = {(uint64_t)ps, (uint64_t)ps, (uint64_t)Wp, (uint64_t)Hp, (uint64_t)3 * p->B};
    uint64_t strides[4] = {(uint64_t)p->W * 4, (uint64_t)ps * 4, (uint64_t)ps * p->W * 4, (uint64_t)p->H * p->W * 4};
    uint32_t box[5] = {(uint32_t)ps, 1u, (uint32_t)tw, (uint32_t)th, 1u};  // [128 patches x one patch row]
    if ((r = make_tensor_map(&tmA, p->img, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, dims, strides, box,
                             ps == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B)))
      return r;
  }
  {
    uint64_t dims[2] = {(uint64_t)3 * ps * ps, (uint64_t)p->n};
    uint64_t strides[1] = {(uint64_t)3 * ps * ps * 4};
    uint32_t box[2] = {32u, (uint32_t)bn / 2};
    if ((r = make_tensor_map(&tmB, p->w, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return r;
  }
  const int bw = tw >= 32 ? 32 : tw, bh = 32 / bw;
  if ((r = act_map(&tmC, p->out, p->B, Hp, Wp, p->n, bw, bh))) return r;
  tmAux = tmC;
  const long long total = (long long)g.num_m * g.num_n;
  const int clusters = (int)(total < slots ? total : slots);
  if (bn == 256) return launch2_inst<UC_EPI_BIAS, false, 256, 3>(tmA, tmB, tmC, tmAux, g, 2 * clusters, stream);
  return launch2_inst<UC_EPI_BIAS, false, 128, 3>(tmA, tmB, tmC, tmAux, g, 2 * clusters, stream);
}
