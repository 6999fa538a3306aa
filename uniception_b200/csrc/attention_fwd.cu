// Fused attention forward (uc_attn_fwd): softmax(q k^T * scale) v, head_dim 64, bf16 in / fp32 softmax.
// One CTA per (128-query tile, batch*head); two CTAs co-reside per SM (80 KB smem, 256 TMEM columns each).
//
//   warp 0     : TMA producer  (Q once; K/V tiles of 64 keys, 4-stage ring; 3-D maps so rows >= N zero-fill)
//   warp 1     : MMA issuer    (S = Q K^T  SS-mode 128x64x64 into a DOUBLE-BUFFERED S, issued one tile ahead of
//                               the softmax;  O += P V  TS-mode: P read from TMEM, V MN-major)
//   warps 2..5 : softmax       (thread == query row: no cross-thread reductions; online max/sum in fp32 with
//                               lazy rescaling -- O is only rescaled when the row max grows by > 2^8 --,
//                               P -> TMEM as packed bf16, epilogue O/l -> global, LSE)
// Because QK^T(j+1) is already in TMEM when softmax(j) ends, the exp pipe (MUFU, the true bound of d=64
// attention on this chip: 16 exp/clk/SM vs 256 MMA-FLOPs per score) never waits on the tensor pipe.
// TMEM columns: S0 [0,64) S1 [64,128) fp32 | P0 [128,160) P1 [160,192) packed bf16 | O [192,256) fp32.
#include "common.cuh"

namespace uc {
namespace {

constexpr int AT_BM = 128;      // queries per CTA
constexpr int AT_BN = 64;       // keys per tile
constexpr int AT_D = 64;        // head dim
constexpr int AT_STAGES = 4;
constexpr int AT_THREADS = 192;
constexpr uint32_t AT_QBYTES = 128 * 64 * 2;  // 16 KB
constexpr uint32_t AT_KBYTES = 64 * 64 * 2;   //  8 KB
constexpr uint32_t AT_OFF_BAR = AT_QBYTES + AT_STAGES * 2 * AT_KBYTES;
constexpr uint32_t AT_SMEM = AT_OFF_BAR + 256 + 1024;
constexpr uint32_t TM_S = 0, TM_P = 128, TM_O = 192, TM_COLS = 256;
constexpr float kLazyThreshold = 8.0f;  // log2 units

struct AttnFwdArgs {
  __nv_bfloat16* o;
  float* lse;
  int B, H, Nq, Nk;
  long long ldo;
  float scale_log2;  // scale * log2(e)
  float scale;
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(AT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnFwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  auto sK = [&](int st) { return smem_base + AT_QBYTES + st * 2 * AT_KBYTES; };
  auto sV = [&](int st) { return smem_base + AT_QBYTES + st * 2 * AT_KBYTES + AT_KBYTES; };
  const uint32_t bar = smem_base + AT_OFF_BAR;
  const uint32_t q_full = bar;
  auto kv_full = [&](int st) { return bar + 8u * (1 + st); };
  auto kv_empty = [&](int st) { return bar + 8u * (1 + AT_STAGES + st); };
  auto s_full = [&](int buf) { return bar + 8u * (1 + 2 * AT_STAGES + buf); };
  const uint32_t p_ready = bar + 8u * (3 + 2 * AT_STAGES);
  const uint32_t o_done = p_ready + 8, o_full = p_ready + 16, tmem_slot = p_ready + 24;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BM;
  const int bh = blockIdx.y;
  const int b = bh / a.H, h = bh % a.H;
  const int num_tiles = (a.Nk + AT_BN - 1) / AT_BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < AT_STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    mbar_init(s_full(0), 1);
    mbar_init(s_full(1), 1);
    mbar_init(p_ready, 4);
    mbar_init(o_done, 1);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TM_COLS);
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
    // whole warp loops (uniform control flow); one elected lane issues the TMA copies
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, AT_QBYTES);
      tma_load_3d(sQ, &tmQ, q_full, h * AT_D, q0, b);
    }
    __syncwarp();
    int st = 0;
    uint32_t ph = 0;
    for (int j = 0; j < num_tiles; ++j) {
      mbar_wait(kv_empty(st), ph ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(kv_full(st), 2 * AT_KBYTES);
        tma_load_3d(sK(st), &tmK, kv_full(st), h * AT_D, j * AT_BN, b);
        tma_load_3d(sV(st), &tmV, kv_full(st), h * AT_D, j * AT_BN, b);
      }
      __syncwarp();
      if (++st == AT_STAGES) { st = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // whole warp loops; operands stay in uniform registers; one elected lane issues tcgen05.mma / commit
    const uint32_t idesc_qk = umma_idesc_bf16(AT_BM, AT_BN, 0, 0);
    const uint32_t idesc_pv = umma_idesc_bf16(AT_BM, AT_D, 0, 1);
    const uint64_t qd = umma_desc_kmajor(sQ);
    auto issue_qk = [&](int j) {  // S[j&1] = Q K_j^T
      const int st = j % AT_STAGES;
      mbar_wait(kv_full(st), (j / AT_STAGES) & 1);
      tc_fence_after();
      const uint64_t kd = umma_desc_kmajor(sK(st));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k)
          umma_ss(tmem_base + TM_S + (j & 1) * 64, qd + uint64_t(k * 2), kd + uint64_t(k * 2), idesc_qk, k > 0);
        umma_commit(s_full(j & 1));
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_qk(0);
    for (int j = 0; j < num_tiles; ++j) {
      // look-ahead: S[(j+1)&1] was last read by softmax(j-1), which finished before p_ready(j-1) fired
      if (j + 1 < num_tiles) issue_qk(j + 1);
      mbar_wait(p_ready, j & 1);
      tc_fence_after();
      const int st = j % AT_STAGES;
      const uint64_t vd = umma_desc_mnmajor(sV(st), 8192);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < AT_BN / 16; ++k)
          umma_ts(tmem_base + TM_O, tmem_base + TM_P + (j & 1) * 32 + k * 8, vd + uint64_t(k * 128), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(kv_empty(st));
        umma_commit(o_done);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(o_full);
    __syncwarp();
  } else {
    // ===================== softmax: thread <-> query row =====================
    const int lane_group = warp & 3;
    const uint32_t lane_addr = uint32_t(lane_group * 32) << 16;
    const int row = q0 + lane_group * 32 + lane;
    float m_run = -INFINITY;  // reference max (raw score units); lags the true max by at most kLazyThreshold
    float l_run = 0.f;
    for (int j = 0; j < num_tiles; ++j) {
      mbar_wait(s_full(j & 1), (j >> 1) & 1);
      tc_fence_after();
      const int kv_valid = a.Nk - j * AT_BN;  // columns >= kv_valid are padding
      uint32_t s[64];
      tmem_ld32(tmem_base + lane_addr + TM_S + (j & 1) * 64, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      tmem_ld32(tmem_base + lane_addr + TM_S + (j & 1) * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      tmem_ld_wait();
      float m_tile = -INFINITY;
      if (kv_valid >= AT_BN) {
#pragma unroll
        for (int i = 0; i < 64; ++i) m_tile = fmaxf(m_tile, __uint_as_float(s[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          if (i >= kv_valid) s[i] = 0xff800000u;  // -inf
          m_tile = fmaxf(m_tile, __uint_as_float(s[i]));
        }
      }
      // lazy rescaling: keep the old reference unless the max grew by more than 2^kLazyThreshold
      const bool grow = (m_tile - m_run) * a.scale_log2 > kLazyThreshold;  // true on the first tile (m_run = -inf)
      const float m_new = grow ? m_tile : m_run;
      const float alpha = grow ? fast_exp2((m_run - m_new) * a.scale_log2) : 1.0f;
      const float m_scaled = m_new * a.scale_log2;
      float l_tile = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float p0 = fast_exp2(__uint_as_float(s[32 * c + 2 * i]) * a.scale_log2 - m_scaled);
          const float p1 = fast_exp2(__uint_as_float(s[32 * c + 2 * i + 1]) * a.scale_log2 - m_scaled);
          l_tile += p0 + p1;
          pk[i] = pack_bf16(p0, p1);
        }
        tmem_st16(tmem_base + lane_addr + TM_P + (j & 1) * 32 + 16 * c, pk);
      }
      l_run = l_run * alpha + l_tile;
      m_run = m_new;
      if (j > 0 && __any_sync(0xffffffffu, grow)) {
        // rescale the running O: PV(j-1) must have retired, PV(j) cannot start before p_ready(j)
        mbar_wait(o_done, (j - 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t o[32];
          tmem_ld32(tmem_base + lane_addr + TM_O + 32 * c, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st32(tmem_base + lane_addr + TM_O + 32 * c, o);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
    }
    // ---- epilogue: O / l -> global, LSE ----
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const bool row_ok = row < a.Nq;
    __nv_bfloat16* optr = a.o + ((long long)b * a.Nq + row) * a.ldo + h * AT_D;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_base + lane_addr + TM_O + c * 32, r);
      tmem_ld_wait();
      if (row_ok) {
        uint4* dst = reinterpret_cast<uint4*>(optr + c * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 t;
          t.x = pack_bf16(__uint_as_float(r[8 * i + 0]) * inv_l, __uint_as_float(r[8 * i + 1]) * inv_l);
          t.y = pack_bf16(__uint_as_float(r[8 * i + 2]) * inv_l, __uint_as_float(r[8 * i + 3]) * inv_l);
          t.z = pack_bf16(__uint_as_float(r[8 * i + 4]) * inv_l, __uint_as_float(r[8 * i + 5]) * inv_l);
          t.w = pack_bf16(__uint_as_float(r[8 * i + 6]) * inv_l, __uint_as_float(r[8 * i + 7]) * inv_l);
          dst[i] = t;
        }
      }
      __syncwarp();
    }
    if (row_ok && a.lse) a.lse[((long long)b * a.H + h) * a.Nq + row] = m_run * a.scale + logf(l_run);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}

}  // namespace

// 3-D map over a token-major bf16 matrix: dims (H*64 columns, N tokens, B), box (64, box_rows, 1).
int attn_fwd_v1(const uc_attn_fwd_params* p, cudaStream_t stream);

int make_head_map(CUtensorMap* m, const void* base, int H, int N, int B, long long ld, int box_rows) {
  uint64_t dims[3] = {(uint64_t)H * 64, (uint64_t)N, (uint64_t)B};
  uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)N * (uint64_t)ld * 2};
  uint32_t box[3] = {64, (uint32_t)box_rows, 1};
  return make_tensor_map(m, base, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace uc

// First-generation forward kernel, kept behind UC_ATTN_FWD=1 as the A/B baseline of attention2.cu (which owns the C entry point).
int uc::attn_fwd_v1(const uc_attn_fwd_params* p, cudaStream_t stream) {
  using namespace uc;
  UC_REQUIRE(p && p->q && p->k && p->v && p->o, UC_ERR_BAD_SHAPE, "uc_attn_fwd: null pointer");
  UC_REQUIRE(p->B > 0 && p->H > 0 && p->Nq > 0 && p->Nk > 0, UC_ERR_BAD_SHAPE, "uc_attn_fwd: bad shape");
  UC_REQUIRE(p->ldq % 8 == 0 && p->ldk % 8 == 0 && p->ldv % 8 == 0 && p->ldo % 8 == 0, UC_ERR_BAD_SHAPE,
             "uc_attn_fwd: leading dimensions must be multiples of 8");
  UC_REQUIRE(((uintptr_t)p->q % 16 == 0) && ((uintptr_t)p->k % 16 == 0) && ((uintptr_t)p->v % 16 == 0) &&
                 ((uintptr_t)p->o % 16 == 0),
             UC_ERR_BAD_SHAPE, "uc_attn_fwd: pointers must be 16-byte aligned");
  CUtensorMap tmQ, tmK, tmV;
  int r;
  if ((r = make_head_map(&tmQ, p->q, p->H, p->Nq, p->B, p->ldq, 128))) return r;
  if ((r = make_head_map(&tmK, p->k, p->H, p->Nk, p->B, p->ldk, AT_BN))) return r;
  if ((r = make_head_map(&tmV, p->v, p->H, p->Nk, p->B, p->ldv, AT_BN))) return r;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    UC_REQUIRE(e == cudaSuccess, UC_ERR_CUDA, "uc_attn_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    configured = true;
  }
  AttnFwdArgs a;
  a.o = static_cast<__nv_bfloat16*>(p->o);
  a.lse = p->lse;
  a.B = p->B; a.H = p->H; a.Nq = p->Nq; a.Nk = p->Nk;
  a.ldo = p->ldo;
  a.scale = p->scale;
  a.scale_log2 = p->scale * 1.4426950408889634f;
  dim3 grid((p->Nq + AT_BM - 1) / AT_BM, p->B * p->H);
  cudaError_t le = launch_pdl(attn_fwd_kernel, grid, dim3(AT_THREADS), AT_SMEM, stream, tmQ, tmK, tmV, a);
  UC_REQUIRE(le == cudaSuccess, UC_ERR_CUDA, "uc_attn_fwd: launch failed: %s", cudaGetErrorString(le));
  return check_launch("uc_attn_fwd");
}
