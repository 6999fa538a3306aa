// Fused attention backward (uc_attn_bwd): dQ, dK, dV of softmax(q k^T * scale) v with recomputation,
// head_dim 64.  Three kernels:
//   1. attn_bwd_delta  : delta[q] = sum_d dO[q,d] * O[q,d]                       (memory-bound)
//   2. attn_bwd_pipe   : one CTA per (128-key tile, batch*head), loop over 64-query SUB-TILES, all five
//                        GEMMs on tcgen05, transposed formulation so that thread == key row:
//        S^T  = K  Q_j^T     (SS)      dP^T = V dO_j^T     (SS)       both double-buffered in TMEM
//        P^T  = exp2(S^T c - lse_j),   dS^T = P^T o (dP^T - delta_j)            (registers)
//        dV  += P^T  dO_j    (TS: P^T  from TMEM, dO_j MN-major from the same smem tile)
//        dK  += dS^T Q_j     (TS: dS^T from TMEM, Q_j  MN-major; or SS from the dS smem tile)
//        dQ_i = dS   K       (SS, once per 128 queries: dS written to smem as an MN-major A tile, K MN-major) -> TMEM ->
//                            smem -> cp.reduce.async.bulk.tensor (TMA add-reduction into the fp32 dq accumulator)
//   3. attn_bwd_finish : dq = bf16(scale * dq_acc) with optional inverse 2-D RoPE  (memory-bound)
// dK gets `scale` and the optional inverse RoPE in the main kernel's epilogue.
//   warp 0: TMA producer | warp 1: MMA issuer | warps 2..9: softmax/dS (two threads per key row, 32 query
//   columns each per sub-tile) | warps 10..13: dQ drain (TMEM -> swizzled smem -> TMA reduce-add)
// TMEM columns: S^T/P^T and dP^T/dS^T x 2 buffers [0,256) | dV [256,320) | dK [320,384) | dQ [384,448).
// Measured (profiles/r01f_*): a 128x64x16 MMA costs ~64 clk whatever its mode (A-operand fetch bound), so the five
// GEMMs of a 128x128 tile pair cost >= 40 x 64 = 2560 clk; the serial predecessor of this kernel took 5800, this one 3600.
#include "common.cuh"
#include <stdlib.h>

namespace uc {

// Optional cycle trace of CTA (0,0) of the pipelined backward kernel (bring-up aid, tools/trace_attn.py; build with
// -DUC_ATTN_TRACE): trace[role][sub-tile][event] = clock64().
__device__ long long* g_attn_trace = nullptr;
#ifdef UC_ATTN_TRACE
#define UC_TRACE(role, j, ev)                                                                       \
  do {                                                                                              \
    if (trace_on && (j) < 64) g_attn_trace[((role) * 64 + (j)) * 8 + (ev)] = clock64();             \
  } while (0)
#else
#define UC_TRACE(role, j, ev) do { } while (0)
#endif

int make_head_map(CUtensorMap* m, const void* base, int H, int N, int B, long long ld, int box_rows);
int attn_bwd_v1(const uc_attn_bwd_params* p, cudaStream_t stream);

namespace {

constexpr int BW_THREADS = 448;  // TMA, MMA, 8 softmax/dS warps (two threads per key row), 4 dQ-drain warps
constexpr uint32_t BW_TILE = 128 * 64 * 2;  // 16 KB
constexpr uint32_t BT_DV = 256, BT_DK = 320, BT_DQ = 384, BT_COLS = 512;  // accumulators; S^T / dP^T buffers: pt_sp / pt_dp

struct AttnBwdArgs {
  const float* lse;
  const float* delta;
  float* dq_acc;
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
  int B, H, Nq, Nk;
  long long lddk, lddv;
  float scale, scale_log2;
  const int* k_positions;
  const float* rope_table;
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// delta[b][h][q] = sum_d dO * O ; 8 lanes per (token, head), 8 elements per lane
__global__ void attn_bwd_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, float* __restrict__ delta,
                                      int B, int H, int N, long long ldo, long long lddo) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = (long long)B * N * H * 8;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int sub = idx & 7;
    const long long th = idx >> 3;
    const int h = th % H;
    const long long tok = th / H;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(o + tok * ldo + h * 64 + sub * 8));
    const uint4 g = __ldg(reinterpret_cast<const uint4*>(d_o + tok * lddo + h * 64 + sub * 8));
    float s = bf16_lo(a.x) * bf16_lo(g.x) + bf16_hi(a.x) * bf16_hi(g.x) + bf16_lo(a.y) * bf16_lo(g.y) + bf16_hi(a.y) * bf16_hi(g.y) +
              bf16_lo(a.z) * bf16_lo(g.z) + bf16_hi(a.z) * bf16_hi(g.z) + bf16_lo(a.w) * bf16_lo(g.w) + bf16_hi(a.w) * bf16_hi(g.w);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (sub == 0) {
      const int n = tok % N;
      const int b = tok / N;
      delta[((long long)b * H + h) * N + n] = s;
    }
  }
}

// dq = bf16(scale * dq_acc) with optional inverse RoPE; one thread per (token, head, 8-pair group)
__global__ void attn_bwd_finish_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq, long long rows, int H, long long lddq,
                                       float scale, const int* __restrict__ pos, const float* __restrict__ table) {
  // each thread: one 32-wide half-head: 16 (u,v) pairs
  pdl_launch_dependents();
  pdl_wait();
  const long long total = rows * H * 2;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int half = idx & 1;
    const int h = (idx >> 1) % H;
    const long long row = (idx >> 1) / H;
    const float4* src = reinterpret_cast<const float4*>(acc + row * (long long)H * 64 + h * 64 + half * 32);
    float v[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 t = __ldg(src + i);
      v[4 * i] = t.x * scale; v[4 * i + 1] = t.y * scale; v[4 * i + 2] = t.z * scale; v[4 * i + 3] = t.w * scale;
    }
    if (pos) {
      const float* tr = table + (long long)pos[2 * row + half] * 32;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float c = tr[2 * i], s = -tr[2 * i + 1];  // inverse rotation
        const float u = v[i], w = v[i + 16];
        v[i] = u * c - w * s;
        v[i + 16] = w * c + u * s;
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(dq + row * lddq + h * 64 + half * 32);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 t;
      t.x = pack_bf16(v[8 * i], v[8 * i + 1]); t.y = pack_bf16(v[8 * i + 2], v[8 * i + 3]);
      t.z = pack_bf16(v[8 * i + 4], v[8 * i + 5]); t.w = pack_bf16(v[8 * i + 6], v[8 * i + 7]);
      dst[i] = t;
    }
  }
}

// ---- pipelined main kernel ----------------------------------------------------------------------------------
// The query loop runs in SUB-TILES of 64 queries.  S^T / dP^T are double-buffered in TMEM (2 x (64 + 64) columns), so
// while the softmax warps turn sub-tile j into P^T / dS^T, the tensor pipe already runs  dV,dK(j-1) [, dQ]  and
// S^T,dP^T(j+1).  tcgen05.mma executes in issue order, which is what makes re-using buffer (j+1)&1 right behind the
// TS MMAs that read P^T(j-1) from it safe.  dS (smem, A operand of dQ) is double-buffered per 128-query tile because
// softmax(j+2) may start before dQ(tile) has been read.
// DK_SS: dK += dS^T Q as an SS MMA reading dS^T from the smem tile that dQ needs anyway (the tile is at the same time
// an MN-major A operand for dQ and a K-major A operand for dK), instead of a TS MMA reading dS^T from TMEM.
constexpr uint32_t P_OFF_K = 0, P_OFF_V = BW_TILE, P_OFF_QDO = 2 * BW_TILE, P_OFF_DS = 6 * BW_TILE /* 2 x 32 KB */,
                   P_OFF_DQ = 10 * BW_TILE, P_OFF_STATS = 12 * BW_TILE, P_OFF_BAR = 12 * BW_TILE + 2048;
constexpr uint32_t P_SMEM = P_OFF_BAR + 256 + 1024;
__host__ __device__ constexpr uint32_t pt_sp(int buf) { return uint32_t(buf) * 128u; }
__host__ __device__ constexpr uint32_t pt_dp(int buf) { return uint32_t(buf) * 128u + 64u; }

template <bool DK_SS>
__global__ void __launch_bounds__(BW_THREADS, 1)
attn_bwd_pipe_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                     const __grid_constant__ CUtensorMap tmDQ, const AttnBwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sK = smem_base + P_OFF_K, sV = smem_base + P_OFF_V;
  auto sQ = [&](int st) { return smem_base + P_OFF_QDO + st * 2 * BW_TILE; };
  auto sdO = [&](int st) { return smem_base + P_OFF_QDO + st * 2 * BW_TILE + BW_TILE; };
  auto sDS = [&](int db) { return smem_base + P_OFF_DS + db * 2 * BW_TILE; };
  float* stats = reinterpret_cast<float*>(smem_gen + P_OFF_STATS);  // [2 stages][2][128]
  const uint32_t bar = smem_base + P_OFF_BAR;
  const uint32_t kv_full = bar;
  auto qdo_full = [&](int st) { return bar + 8u * (1 + st); };
  auto qdo_empty = [&](int st) { return bar + 8u * (3 + st); };
  auto sdp_full = [&](int buf) { return bar + 8u * (5 + buf); };
  auto ds_ready = [&](int buf) { return bar + 8u * (7 + buf); };
  const uint32_t dq_full = bar + 8u * 9, dq_free = bar + 8u * 10, fin_full = bar + 8u * 11, tmem_slot = bar + 8u * 12;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * 128;
  const int bh = blockIdx.y;
  const int b = bh / a.H, h = bh % a.H;
  const int num_q_tiles = (a.Nq + 127) / 128;
  const int n_sub = 2 * num_q_tiles;
#ifdef UC_ATTN_TRACE
  const bool trace_on = g_attn_trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0;
#endif

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO); tma_prefetch_desc(&tmDQ);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(qdo_full(s), 1); mbar_init(qdo_empty(s), 1);
      mbar_init(sdp_full(s), 1); mbar_init(ds_ready(s), 8);
    }
    mbar_init(dq_full, 1);
    mbar_init(dq_free, 4);
    mbar_init(fin_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BT_COLS);
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
    if (elect_one()) {
      mbar_arrive_expect_tx(kv_full, 2 * BW_TILE);
      tma_load_3d(sK, &tmK, kv_full, h * 64, kv0, b);
      tma_load_3d(sV, &tmV, kv_full, h * 64, kv0, b);
    }
    __syncwarp();
    for (int i = 0; i < num_q_tiles; ++i) {
      const int st = i & 1;
      mbar_wait(qdo_empty(st), ((i >> 1) & 1) ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(qdo_full(st), 2 * BW_TILE);
        tma_load_3d(sQ(st), &tmQ, qdo_full(st), h * 64, i * 128, b);
        tma_load_3d(sdO(st), &tmdO, qdo_full(st), h * 64, i * 128, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t id_s = umma_idesc_bf16(128, 64, 0, 0);    // [128 keys] x [64 queries], both K-major
    const uint32_t id_ts = umma_idesc_bf16(128, 64, 0, 1);   // A in TMEM or K-major smem, B MN-major
    const uint32_t id_dq = umma_idesc_bf16(128, 64, 1, 1);   // A MN-major (dS in smem), B MN-major
    mbar_wait(kv_full, 0);
    const uint64_t kd = umma_desc_kmajor(sK), vd = umma_desc_kmajor(sV);
    const uint64_t k_mn = umma_desc_mnmajor(sK, 8192);
    auto issue_sdp = [&](int j) {  // S^T(j) = K Q_j^T, dP^T(j) = V dO_j^T into buffer j&1
      const int i = j >> 1, hq = j & 1, st = i & 1, buf = j & 1;
      if (hq == 0) {
        mbar_wait(qdo_full(st), (i >> 1) & 1);
        tc_fence_after();
      }
      const uint64_t qd = umma_desc_kmajor(sQ(st) + hq * 8192), dod = umma_desc_kmajor(sdO(st) + hq * 8192);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_ss(tmem_base + pt_sp(buf), kd + uint64_t(k * 2), qd + uint64_t(k * 2), id_s, k > 0);
          umma_ss(tmem_base + pt_dp(buf), vd + uint64_t(k * 2), dod + uint64_t(k * 2), id_s, k > 0);
        }
        umma_commit(sdp_full(buf));
      }
      __syncwarp();
    };
    issue_sdp(0);
    for (int j = 0; j < n_sub; ++j) {
      const int i = j >> 1, hq = j & 1, st = i & 1, buf = j & 1;
      UC_TRACE(0, j, 0);
      if (j + 1 < n_sub) issue_sdp(j + 1);
      UC_TRACE(0, j, 1);
      mbar_wait(ds_ready(buf), (j >> 1) & 1);
      UC_TRACE(0, j, 2);
      if (hq == 1 && i > 0) mbar_wait(dq_free, (i - 1) & 1);
      tc_fence_after();
      const uint64_t do_mn = umma_desc_mnmajor(sdO(st), 8192), q_mn = umma_desc_mnmajor(sQ(st), 8192);
      const uint64_t ds_k = umma_desc_kmajor(sDS(i & 1) + hq * 16384);
      const uint64_t ds_mn = umma_desc_mnmajor(sDS(i & 1), 16384);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t acol = uint32_t((k >> 1) * 32 + (k & 1) * 8);
          umma_ts(tmem_base + BT_DV, tmem_base + pt_sp(buf) + acol, do_mn + uint64_t((4 * hq + k) * 128), id_ts,
                  (j > 0 || k > 0) ? 1u : 0u);
          if (DK_SS)
            umma_ss(tmem_base + BT_DK, ds_k + uint64_t(k * 2), q_mn + uint64_t((4 * hq + k) * 128), id_ts, (j > 0 || k > 0) ? 1u : 0u);
          else
            umma_ts(tmem_base + BT_DK, tmem_base + pt_dp(buf) + acol, q_mn + uint64_t((4 * hq + k) * 128), id_ts,
                    (j > 0 || k > 0) ? 1u : 0u);
        }
        if (hq == 1) {
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_ss(tmem_base + BT_DQ, ds_mn + uint64_t(k * 128), k_mn + uint64_t(k * 128), id_dq, k > 0);
          umma_commit(dq_full);
          umma_commit(qdo_empty(st));
        }
      }
      __syncwarp();
      UC_TRACE(0, j, 3);
    }
    if (elect_one()) umma_commit(fin_full);
    __syncwarp();
  } else if (warp < 10) {
    // ===================== softmax / dS: two threads per key row, 32 queries each per sub-tile =====================
    const int lane_group = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t lane_addr = uint32_t(lane_group * 32) << 16;
    const int r = lane_group * 32 + lane;   // key row inside the tile
    const int ct = threadIdx.x - 64;        // 0..255
    auto load_stat = [&](int tile) -> float {  // threads 0..127: lse*log2e, 128..255: delta, of query (tile, ct&127)
      const int q = tile * 128 + (ct & 127);
      if (q >= a.Nq) return ct < 128 ? INFINITY : 0.f;
      const long long off = ((long long)b * a.H + h) * a.Nq + q;
      return ct < 128 ? a.lse[off] * 1.4426950408889634f : a.delta[off];
    };
    stats[ct] = load_stat(0);
    float next_stat = 0.f;
    for (int j = 0; j < n_sub; ++j) {
      const int i = j >> 1, hq = j & 1, st = i & 1, buf = j & 1;
      if (hq == 0) {
        named_bar_sync(1, 256);  // stats(i) visible; everyone is done with tile i-1
        next_stat = (i + 1 < num_q_tiles) ? load_stat(i + 1) : 0.f;  // latency hidden behind this tile's work
      }
      const float4* lse4 = reinterpret_cast<const float4*>(stats + st * 256 + hq * 64 + half * 32);
      const float4* dl4 = lse4 + 32;  // + 128 floats
      if (warp == 2) UC_TRACE(1, j, 0);
      mbar_wait(sdp_full(buf), (j >> 1) & 1);
      tc_fence_after();
      if (warp == 2) UC_TRACE(1, j, 1);
      uint32_t s[32], dp[32];
      tmem_ld32(tmem_base + lane_addr + pt_sp(buf) + 32 * half, s);
      tmem_ld32(tmem_base + lane_addr + pt_dp(buf) + 32 * half, dp);
      tmem_ld_wait();
      if (warp == 2) UC_TRACE(1, j, 2);
      uint32_t pk[16], dk[16];
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 l = lse4[g], d = dl4[g];
        const float p0 = fast_exp2(fmaf(__uint_as_float(s[4 * g + 0]), a.scale_log2, -l.x));
        const float p1 = fast_exp2(fmaf(__uint_as_float(s[4 * g + 1]), a.scale_log2, -l.y));
        const float p2 = fast_exp2(fmaf(__uint_as_float(s[4 * g + 2]), a.scale_log2, -l.z));
        const float p3 = fast_exp2(fmaf(__uint_as_float(s[4 * g + 3]), a.scale_log2, -l.w));
        pk[2 * g] = pack_bf16(p0, p1);
        pk[2 * g + 1] = pack_bf16(p2, p3);
        dk[2 * g] = pack_bf16(p0 * (__uint_as_float(dp[4 * g + 0]) - d.x), p1 * (__uint_as_float(dp[4 * g + 1]) - d.y));
        dk[2 * g + 1] = pack_bf16(p2 * (__uint_as_float(dp[4 * g + 2]) - d.z), p3 * (__uint_as_float(dp[4 * g + 3]) - d.w));
      }
      if (warp == 2) UC_TRACE(1, j, 3);
      // packed P^T / dS^T overwrite the first half of THIS thread's own (already consumed) 32 fp32 columns
      tmem_st16(tmem_base + lane_addr + pt_sp(buf) + 32 * half, pk);
      if (!DK_SS) tmem_st16(tmem_base + lane_addr + pt_dp(buf) + 32 * half, dk);
      // dS -> smem tile [q-block hq][key row r][64 q], 128B swizzle: MN-major A of dQ and K-major A of dK
      const uint32_t ds_row = sDS(i & 1) + hq * 16384 + r * 128;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint32_t addr = ds_row + ((uint32_t(half * 4 + g) ^ uint32_t(r & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(dk[4 * g]), "r"(dk[4 * g + 1]),
                     "r"(dk[4 * g + 2]), "r"(dk[4 * g + 3])
                     : "memory");
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_ready(buf));
      if (warp == 2) UC_TRACE(1, j, 4);
      if (hq == 1) stats[(st ^ 1) * 256 + ct] = next_stat;
    }
    // ---- epilogue: dV, dK (x scale, inverse RoPE) ----
    mbar_wait(fin_full, 0);
    tc_fence_after();
    const int kv = kv0 + r;
    const bool ok = kv < a.Nk;
    const long long tok = (long long)b * a.Nk + kv;
    {
      const int c = half;
      uint32_t v[32];
      tmem_ld32(tmem_base + lane_addr + BT_DV + c * 32, v);
      tmem_ld_wait();
      if (ok) {
        uint4* dst = reinterpret_cast<uint4*>(a.dv + tok * a.lddv + h * 64 + c * 32);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 t;
          t.x = pack_bf16(__uint_as_float(v[8 * g]), __uint_as_float(v[8 * g + 1]));
          t.y = pack_bf16(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3]));
          t.z = pack_bf16(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5]));
          t.w = pack_bf16(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7]));
          dst[g] = t;
        }
      }
      __syncwarp();
    }
    {
      const int c = half;
      uint32_t raw[32];
      tmem_ld32(tmem_base + lane_addr + BT_DK + c * 32, raw);
      tmem_ld_wait();
      if (ok) {
        float v[32];
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) v[jj] = __uint_as_float(raw[jj]) * a.scale;
        if (a.k_positions) {
          const float* tr = a.rope_table + (long long)a.k_positions[2 * tok + c] * 32;
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            const float cs = tr[2 * jj], sn = -tr[2 * jj + 1];
            const float u = v[jj], w = v[jj + 16];
            v[jj] = u * cs - w * sn;
            v[jj + 16] = w * cs + u * sn;
          }
        }
        uint4* dst = reinterpret_cast<uint4*>(a.dk + tok * a.lddk + h * 64 + c * 32);
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
  } else {
    // ===================== dQ drain: TMEM -> swizzled smem -> TMA reduce-add (once per 128-query tile) =====================
    const int lane_group = warp & 3;
    const uint32_t lane_addr = uint32_t(lane_group * 32) << 16;
    const int r = lane_group * 32 + lane;
    const uint32_t sDQ = smem_base + P_OFF_DQ;
    const bool issuer = (warp == 10 && lane == 0);
    for (int i = 0; i < num_q_tiles; ++i) {
      if (warp == 10) UC_TRACE(2, i, 0);
      mbar_wait(dq_full, i & 1);
      tc_fence_after();
      if (warp == 10) UC_TRACE(2, i, 1);
      uint32_t v0[32], v1[32];
      tmem_ld32(tmem_base + lane_addr + BT_DQ, v0);
      tmem_ld32(tmem_base + lane_addr + BT_DQ + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);          // TMEM dQ region may be overwritten by the next tile
      if (issuer) tma_store_wait_read0();            // previous reduction has finished reading the staging tile
      named_bar_sync(2, 128);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const uint32_t off = r * 128 + ((jj ^ (r & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sDQ + off), "r"(v0[4 * jj]), "r"(v0[4 * jj + 1]),
                     "r"(v0[4 * jj + 2]), "r"(v0[4 * jj + 3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sDQ + 16384 + off), "r"(v1[4 * jj]), "r"(v1[4 * jj + 1]),
                     "r"(v1[4 * jj + 2]), "r"(v1[4 * jj + 3]) : "memory");
      }
      fence_proxy_async_smem();
      named_bar_sync(2, 128);
      if (issuer) {
        tma_reduce_add_3d(&tmDQ, sDQ, h * 64, i * 128, b);
        tma_reduce_add_3d(&tmDQ, sDQ + 16384, h * 64 + 32, i * 128, b);
        tma_store_commit();
      }
      if (warp == 10) UC_TRACE(2, i, 2);
    }
    if (issuer) tma_store_wait0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BT_COLS);
  }
}

}  // namespace
}  // namespace uc

extern "C" __attribute__((visibility("default"))) int uc_debug_set_attn_trace(long long* buf) {
  return cudaMemcpyToSymbol(uc::g_attn_trace, &buf, sizeof(buf)) == cudaSuccess ? 0 : UC_ERR_CUDA;
}

// First-generation fused backward (one kernel for dQ, dK, dV + fp32 dQ accumulator), kept behind UC_ATTN_BWD=1 as the A/B
// baseline of attention2.cu (which owns the C entry point).
int uc::attn_bwd_v1(const uc_attn_bwd_params* p, cudaStream_t stream) {
  using namespace uc;
  UC_REQUIRE(p && p->q && p->k && p->v && p->o && p->d_o && p->lse && p->delta && p->dq_acc && p->dq && p->dk && p->dv,
             UC_ERR_BAD_SHAPE, "uc_attn_bwd: null pointer");
  UC_REQUIRE(p->B > 0 && p->H > 0 && p->Nq > 0 && p->Nk > 0, UC_ERR_BAD_SHAPE, "uc_attn_bwd: bad shape");
  UC_REQUIRE(p->ldq % 8 == 0 && p->ldk % 8 == 0 && p->ldv % 8 == 0 && p->ldo % 8 == 0 && p->lddq % 8 == 0 && p->lddk % 8 == 0 &&
                 p->lddv % 8 == 0,
             UC_ERR_BAD_SHAPE, "uc_attn_bwd: leading dimensions must be multiples of 8");
  UC_REQUIRE((p->q_positions == nullptr && p->k_positions == nullptr) || p->rope_table, UC_ERR_BAD_SHAPE,
             "uc_attn_bwd: positions given without rope_table");
  const long long rows_q = (long long)p->B * p->Nq;
  {
    const long long total = rows_q * p->H * 8;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    cudaError_t le = launch_pdl(attn_bwd_delta_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, static_cast<const __nv_bfloat16*>(p->o),
                                static_cast<const __nv_bfloat16*>(p->d_o), p->delta, p->B, p->H, p->Nq, (long long)p->ldo, (long long)p->ldo);
    UC_REQUIRE(le == cudaSuccess, UC_ERR_CUDA, "uc_attn_bwd(delta): launch failed: %s", cudaGetErrorString(le));
    int r = check_launch("uc_attn_bwd(delta)");
    if (r) return r;
  }
  cudaError_t e = cudaMemsetAsync(p->dq_acc, 0, sizeof(float) * rows_q * p->H * 64, stream);
  UC_REQUIRE(e == cudaSuccess, UC_ERR_CUDA, "uc_attn_bwd: memset failed: %s", cudaGetErrorString(e));

  CUtensorMap tmQ, tmK, tmV, tmdO, tmDQ;
  int r;
  if ((r = make_head_map(&tmQ, p->q, p->H, p->Nq, p->B, p->ldq, 128))) return r;
  if ((r = make_head_map(&tmK, p->k, p->H, p->Nk, p->B, p->ldk, 128))) return r;
  if ((r = make_head_map(&tmV, p->v, p->H, p->Nk, p->B, p->ldv, 128))) return r;
  if ((r = make_head_map(&tmdO, p->d_o, p->H, p->Nq, p->B, p->ldo, 128))) return r;
  {  // fp32 dq accumulator [B][Nq][H*64]: box (32 floats = 128 B, 128 rows), 128B swizzle; rows >= Nq are clipped by TMA
    uint64_t dims[3] = {(uint64_t)p->H * 64, (uint64_t)p->Nq, (uint64_t)p->B};
    uint64_t strides[2] = {(uint64_t)p->H * 64 * 4, (uint64_t)p->Nq * p->H * 64 * 4};
    uint32_t box[3] = {32, 128, 1};
    if ((r = make_tensor_map(&tmDQ, p->dq_acc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return r;
  }
  AttnBwdArgs a;
  a.lse = p->lse; a.delta = p->delta; a.dq_acc = p->dq_acc;
  a.dk = static_cast<__nv_bfloat16*>(p->dk);
  a.dv = static_cast<__nv_bfloat16*>(p->dv);
  a.B = p->B; a.H = p->H; a.Nq = p->Nq; a.Nk = p->Nk;
  a.lddk = p->lddk; a.lddv = p->lddv;
  a.scale = p->scale;
  a.scale_log2 = p->scale * 1.4426950408889634f;
  a.k_positions = p->k_positions;
  a.rope_table = p->rope_table;
  dim3 grid((p->Nk + 127) / 128, p->B * p->H);
  static const int dk_ss = [] { const char* ev = getenv("UC_ATTN_BWD_DK_SS"); return ev ? atoi(ev) : 0; }();  // bring-up switch
  static bool configured = false;
  if (!configured) {
    e = cudaFuncSetAttribute(attn_bwd_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM);
    UC_REQUIRE(e == cudaSuccess, UC_ERR_CUDA, "uc_attn_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    e = cudaFuncSetAttribute(attn_bwd_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM);
    UC_REQUIRE(e == cudaSuccess, UC_ERR_CUDA, "uc_attn_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    configured = true;
  }
  e = dk_ss ? launch_pdl(attn_bwd_pipe_kernel<true>, grid, dim3(BW_THREADS), P_SMEM, stream, tmQ, tmK, tmV, tmdO, tmDQ, a)
            : launch_pdl(attn_bwd_pipe_kernel<false>, grid, dim3(BW_THREADS), P_SMEM, stream, tmQ, tmK, tmV, tmdO, tmDQ, a);
  UC_REQUIRE(e == cudaSuccess, UC_ERR_CUDA, "uc_attn_bwd(main): launch failed: %s", cudaGetErrorString(e));
  if ((r = check_launch("uc_attn_bwd(main)"))) return r;
  {
    const long long total = rows_q * p->H * 2;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    cudaError_t le = launch_pdl(attn_bwd_finish_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, (const float*)p->dq_acc,
                                static_cast<__nv_bfloat16*>(p->dq), (long long)rows_q, p->H, (long long)p->lddq, p->scale,
                                (const int*)p->q_positions, (const float*)p->rope_table);
    UC_REQUIRE(le == cudaSuccess, UC_ERR_CUDA, "uc_attn_bwd(finish): launch failed: %s", cudaGetErrorString(le));
    r = check_launch("uc_attn_bwd(finish)");
  }
  return r;
}
