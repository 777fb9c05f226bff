// Residual GEMM + LayerNorm, version 2: the output columns of a 128-row tile are split over a CTA PAIR (sm_100a).
//
// gemm_ln.cuh gives one CTA all D = 384 columns of 128 rows: one TMEM buffer, so MMAs -> epilogue pass 1 -> statistics ->
// pass 2 run one after the other, and 512 tiles make 3.46 -> 4 rounds over 148 SMs.  Here a cluster of two CTAs shares the
// tile: CTA r accumulates columns [r D/2, (r+1) D/2) (UMMA 128 x D/2 x 16, cta_group::1 - each CTA is an independent GEMM
// over the same A rows), which leaves room for TWO accumulator stages in TMEM: the MMAs of the next tile run under the
// epilogue of this one.  Work items halve (1024 half tiles -> 6.92 -> 7 half-size rounds = 3.5).  The row statistics of the
// LayerNorm need both halves: every CTA reduces its D/2 columns to (mean, M2) per row, writes them into the PEER's shared
// memory (st.shared::cluster + a remote mbarrier arrive per lane, once per tile) and both merge the two halves with the
// same symmetric form of Chan's update - identical bits on both sides, no third pass.
//
//   warp 0      TMA producer: A tile 128x64 + this CTA's W rows (D/2 x 64) per 40 KB stage (same stage as gemm_ln.cuh)
//   warp 1      MMA issuer, accumulator stage as = tile parity: TMEM columns [as D/2, (as+1) D/2)
//   warps 2..9  epilogue (two per TMEM lane quarter, thread = row, alternate 32-column chunks): pass 1 as in gemm_ln.cuh
//               over D/2 columns (x chunk by TMA, v = (acc + bias) + x stored back by TMA and to TMEM, shifted sums),
//               warp pair -> CTA (mean, M2) -> exchange with the peer -> pass 2 (normalise, bf16, TMA store of xn)
// x is bit-identical to gemm_ln.cuh / the unfused pair (same k order, same adds); xn may differ from gemm_ln.cuh in the
// last bf16 bit where the statistics round differently (partition of the columns: 96 + 96 | 96 + 96 instead of alternating
// chunks) - the tests hold it to one bf16 ulp of LayerNorm(x) like the other implementations.
#pragma once
#include "gemm_ln.cuh"

namespace pq {

__device__ __forceinline__ void gln2_st_cluster_v2f(uint32_t addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
// wait with cluster-scope acquire: the data guarded by the barrier was written by the peer CTA's generic-proxy stores
__device__ __forceinline__ void gln2_mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  long long t0 = 0;
  uint32_t it = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (it == 0) t0 = clock64();
    if (((++it) & 0x3ffu) == 0 && (clock64() - t0) > PQ_SPIN_LIMIT_CYCLES) {
      printf("[parseq_b200] gemm_ln2 peer-statistics wait timeout: block %d thread %d\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

template <int D>
struct GemmLn2Cfg {
  static constexpr int kN = D / 2;                                    // columns of this CTA
  static constexpr int kABytes = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;     // 16 KB
  static constexpr int kBBytes = kN * GEMM_BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSlabBytes = GLN_EPI_WARPS * GLN_SLABS * 4096;
  // bias, gamma, beta (own columns); warp-pair exchange [4 q][2 w][32] float2; peer statistics [2 parities][128 rows] float2
  static constexpr int kParamBytes = 3 * kN * 4 + 4 * 2 * 32 * 8 + 2 * 128 * 8;
  static constexpr int kBarBytes = 512;
  static constexpr int kStagesRaw = (232448 - 1024 - kBarBytes - kSlabBytes - kParamBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 6 ? 6 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + kSlabBytes + kParamBytes + kBarBytes + 1024;
  static constexpr int kChunks = kN / 32;                             // pass-1 chunks of this CTA per row
  static constexpr int kMyChunks = kChunks / 2;                       // per epilogue warp
  static_assert(D == 384, "D = 192 would leave a warp pair with 2 + 1 chunks (unequal partial statistics): it stays on gemm_ln.cuh");
  static_assert(kN % 16 == 0 && kN <= 256, "UMMA N");
  static_assert(kBBytes % 1024 == 0 && kStageBytes % 1024 == 0, "1024-B aligned operand tiles");
  static_assert(kStages >= 3, "pipeline depth");
  static_assert(kChunks % 2 == 0 && kMyChunks <= GLN_SLABS, "every x chunk of a warp has its own slab: one load round per tile");
  static_assert(kN % 64 == 0, "pass 2 stores 64-column chunks");
};

template <int D>
__global__ void __launch_bounds__(GLN_THREADS, 1)
gemm_ln_split_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmN,
                     const GemmLnParams p) {
  using Cfg = GemmLn2Cfg<D>;
  constexpr int N = Cfg::kN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  uint8_t* smem = smem_raw + pad;
  uint8_t* slab_base = smem + Cfg::kStages * Cfg::kStageBytes;                 // 1024-B aligned
  float* s_bias = reinterpret_cast<float*>(slab_base + Cfg::kSlabBytes);
  float* s_gamma = s_bias + N;
  float* s_beta = s_gamma + N;
  float2* s_stat = reinterpret_cast<float2*>(s_beta + N);            // [4 quarters][2 warps][32 rows] (mean, M2) over N/2 columns
  float2* s_peer = s_stat + 4 * 2 * 32;                              // [2 tile parities][128 rows]: the PEER's (mean, M2) over its N columns
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(slab_base + Cfg::kSlabBytes + Cfg::kParamBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;      // [2] accumulator stage complete
  uint64_t* tempty_bar = tfull_bar + 2;                // [2] the epilogue has drained the stage
  uint64_t* x_bar = tempty_bar + 2;                    // [GLN_EPI_WARPS][GLN_SLABS]: x chunk landed
  uint64_t* peer_bar = x_bar + GLN_EPI_WARPS * GLN_SLABS;   // [2 parities][4 quarters]: the peer's statistics of 32 rows arrived
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(peer_bar + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (p.K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  const uint32_t rank = cluster_ctarank();
  const int first_tile = blockIdx.x >> 1;                // tiles of 128 rows, one per cluster pass
  const int tile_step = gridDim.x >> 1;
  const int col0 = static_cast<int>(rank) * N;           // this CTA's output columns [col0, col0 + N)

  grid_dep_launch();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB); prefetch_tmap(&tmX); prefetch_tmap(&tmN);
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], GLN_EPI_WARPS); }
    for (int i = 0; i < GLN_EPI_WARPS * GLN_SLABS; ++i) mbar_init(&x_bar[i], 1);
    for (int i = 0; i < 8; ++i) mbar_init(&peer_bar[i], 32);          // one remote arrive per row (lane) of the quarter
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  for (int j = threadIdx.x; j < N; j += GLN_THREADS) {
    s_bias[j] = (p.bias != nullptr) ? __ldg(p.bias + col0 + j) : 0.0f;
    s_gamma[j] = __ldg(p.gamma + col0 + j);
    s_beta[j] = __ldg(p.beta + col0 + j);
  }
  tc_fence_before();
  cluster_sync_all();                                    // the peer's barriers are initialised before anything is sent to it
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < p.num_m_tiles; tile += tile_step) {
        const int m0 = tile * GEMM_BLOCK_M;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * GEMM_BLOCK_K, m0);
          tma_load_2d(sa + Cfg::kABytes, &tmB, &full_bar[stage], kb * GEMM_BLOCK_K, col0);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BLOCK_M, N);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = first_tile; tile < p.num_m_tiles; tile += tile_step) {
        mbar_wait(&tempty_bar[as], aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * N);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t adesc = make_desc_k_sw128(sa);
          const uint64_t bdesc = make_desc_k_sw128(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k)
            umma_bf16(tmem_d, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc,
                      static_cast<uint32_t>((kb | k) != 0));
          umma_commit(&empty_bar[stage]);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tfull_bar[as]);
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue: thread = row; warp pair (w = 0, 1) of a quarter alternates chunks =====================
    const int quarter = warp & 3;
    const int ew = warp - 2;
    const int w = ew >> 2;
    uint8_t* my_slabs = slab_base + ew * (GLN_SLABS * 4096);
    uint64_t* my_xbar = x_bar + ew * GLN_SLABS;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    const int row_in_tile = quarter * 32 + lane;
    int as = 0;
    uint32_t aphase = 0;
    uint32_t n_local = 0;                              // tiles this cluster has done: slab-barrier phase, statistics parity
    // (Tried: rotate the slabs so that the one pass 2 does not use takes the next tile's first x chunk early - fc2 went from
    //  0.89x to 0.93x of the full-row kernel's time, i.e. worse; profiles/r2_gemm_ln_split_pair.txt.)
    for (int tile = first_tile; tile < p.num_m_tiles; tile += tile_step, ++n_local) {
      const int row0 = tile * GEMM_BLOCK_M + quarter * 32;
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(as * N);
      // every slab is free here (first tile, or bulk_wait_group_read<0> at the end of the previous tile)
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < Cfg::kMyChunks; ++i) {
          mbar_expect_tx(&my_xbar[i], 4096);
          tma_load_2d(my_slabs + i * 4096, &tmX, &my_xbar[i], col0 + (2 * i + w) * 32, row0);
        }
      }
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      float shift = 0.f, sum = 0.f, sq = 0.f;
#pragma unroll 1
      for (int i = 0; i < Cfg::kMyChunks; ++i) {
        const int c = 2 * i + w;                       // this warp's i-th 32-column chunk of the CTA's N columns
        mbar_wait(&my_xbar[i], n_local & 1u);
        uint32_t v[32];
        tmem_ld_32x32b_x32(trow + static_cast<uint32_t>(c * 32), v);
        tmem_ld_wait();
        uint8_t* slab = my_slabs + i * 4096;
        uint8_t* buf = slab + lane * 128;
        const float* bb = s_bias + c * 32;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          float4* px = reinterpret_cast<float4*>(buf + ((static_cast<uint32_t>(jj) ^ sw) << 4));
          const float4 xo = *px;
          float4 r;
          r.x = (__uint_as_float(v[jj * 4 + 0]) + bb[jj * 4 + 0]) + xo.x;
          r.y = (__uint_as_float(v[jj * 4 + 1]) + bb[jj * 4 + 1]) + xo.y;
          r.z = (__uint_as_float(v[jj * 4 + 2]) + bb[jj * 4 + 2]) + xo.z;
          r.w = (__uint_as_float(v[jj * 4 + 3]) + bb[jj * 4 + 3]) + xo.w;
          *px = r;
          if (i == 0 && jj == 0) shift = r.x;          // shifted single-pass variance (shift = first element seen)
          const float d0 = r.x - shift, d1 = r.y - shift, d2 = r.z - shift, d3 = r.w - shift;
          sum += (d0 + d1) + (d2 + d3);
          sq = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, sq))));
          v[jj * 4 + 0] = __float_as_uint(r.x); v[jj * 4 + 1] = __float_as_uint(r.y);
          v[jj * 4 + 2] = __float_as_uint(r.z); v[jj * 4 + 3] = __float_as_uint(r.w);
        }
        tmem_st_32x32b_x32(trow + static_cast<uint32_t>(c * 32), v);           // keep the updated row for pass 2
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmX, slab, col0 + c * 32, row0);
          bulk_commit_group();
        }
      }
      // ---- (mean, M2) of this warp's N/2 columns -> of the CTA's N columns (warp pair, through shared memory) ----
      constexpr float kQ = 0.25f * D;                  // columns per warp partial (N / 2)
      const float md = sum * (1.0f / kQ);
      const float my_mean = shift + md;
      const float my_m2 = fmaxf(sq - sum * md, 0.0f);
      s_stat[(quarter * 2 + w) * 32 + lane] = make_float2(my_mean, my_m2);
      tmem_st_wait();                                  // (also orders this warp's TMEM stores before the pair barrier)
      tc_fence_before();
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
      tc_fence_after();
      const float2 other = s_stat[(quarter * 2 + (w ^ 1)) * 32 + lane];
      const float dw = other.x - my_mean;              // symmetric forms: both warps of the pair get identical bits
      const float cta_mean = 0.5f * (my_mean + other.x);
      const float cta_m2 = (my_m2 + other.y) + dw * dw * (0.5f * kQ);
      // ---- exchange with the peer CTA: its (mean, M2) over the other N columns of the same rows ----
      const uint32_t par = n_local & 1u;
      if (w == 0) {                                    // one warp per quarter sends: a store + a releasing remote arrive per row
        const uint32_t peer = rank ^ 1u;
        gln2_st_cluster_v2f(mapa_cluster(smem_u32(&s_peer[par * 128 + row_in_tile]), peer), cta_mean, cta_m2);
        mbar_arrive_cluster(mapa_cluster(smem_u32(&peer_bar[par * 4 + quarter]), peer));
      }
      gln2_mbar_wait_cluster(&peer_bar[par * 4 + quarter], (n_local >> 1) & 1u);
      const float2 pr = s_peer[par * 128 + row_in_tile];
      constexpr float kHalfN = 0.5f * D;
      const float delta = pr.x - cta_mean;             // symmetric again: both CTAs compute identical (mean, var)
      const float mean = 0.5f * (cta_mean + pr.x);
      const float var = ((cta_m2 + pr.y) + delta * delta * (0.5f * kHalfN)) * (1.0f / D);
      const float rstd = 1.0f / sqrtf(var + p.eps);
      if (lane == 0) bulk_wait_group_read<0>();        // every slab of this warp is free again
      __syncwarp();
      // ---- pass 2: normalise, bf16, 64 columns (128 B) per row per TMA store; chunk c = w, w + 2, .. ----
      int k2 = 0;
#pragma unroll 1
      for (int c = w; c < N / 64; c += 2, ++k2) {
        uint8_t* slab = my_slabs + k2 * 4096;          // (N / 64 <= 3 chunks per CTA: at most 2 per warp)
        uint8_t* buf = slab + lane * 128;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(trow + static_cast<uint32_t>(c * 64 + h * 32), v);
          tmem_ld_wait();
          const float* gg = s_gamma + c * 64 + h * 32;
          const float* be = s_beta + c * 64 + h * 32;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            float f[8];
#pragma unroll
            for (int t = 0; t < 8; ++t)
              f[t] = (__uint_as_float(v[jj * 8 + t]) - mean) * rstd * gg[jj * 8 + t] + be[jj * 8 + t];
            uint4 q;
            q.x = pack_bf16(f[0], f[1]); q.y = pack_bf16(f[2], f[3]);
            q.z = pack_bf16(f[4], f[5]); q.w = pack_bf16(f[6], f[7]);
            *reinterpret_cast<uint4*>(buf + ((static_cast<uint32_t>(h * 4 + jj) ^ sw) << 4)) = q;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmN, slab, col0 + c * 64, row0);
          bulk_commit_group();
        }
      }
      tc_fence_before();
      if (lane == 0) bulk_wait_group_read<0>();        // slabs free for the next tile's x loads
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);     // this warp has read its last TMEM column of the stage
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
    if (lane == 0) bulk_wait_group<0>();
  }

  tc_fence_before();
  cluster_sync_all();                                  // no CTA exits while its peer may still send statistics to it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace pq
