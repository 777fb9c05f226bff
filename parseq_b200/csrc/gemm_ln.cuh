// Residual GEMM fused with the LayerNorm that follows it (sm_100a, tcgen05 + TMA).
//
//   x[M, D]  (fp32, in place)  <-  x + A[M, K] * W[D, K]^T + bias            (timm Block: x = x + attn(..) / mlp(..))
//   xn[M, D] (bf16)            <-  LayerNorm(x_new; gamma, beta, eps)         (norm2 / next block's norm1 / final norm)
//
// Why: as separate kernels the attention-projection and fc2 GEMMs are bound by the fp32 read-modify-write of the
// residual stream and the LayerNorm re-reads it from HBM; here each updated row is normalised while it is still on
// the SM (profiles/: proj 47 us + LN 31 us, fc2 92 us + LN 31 us at M = 65536 before the fusion).
//
// One CTA owns 128 full rows (N = D) so the row statistics never leave the thread that owns the row:
//   warp 0      TMA producer (A tile 128x64 + a HALF of W: D/2 x 64 per stage; the two column halves of a tile are
//               accumulated one after the other so that a stage stays 40 KB and shared memory is left for the epilogue)
//   warp 1      MMA issuer (UMMA 128 x D/2 x 16 into TMEM columns [h*D/2, (h+1)*D/2)), TMEM alloc / dealloc
//   warps 2..5  epilogue, one per TMEM lane quarter, thread = row:
//       pass 1 (32-column chunks): x tile chunk arrives by TMA in a 128B-swizzled slab (6 slabs per warp are kept in
//               flight), v = (acc + bias) + x is written back in place and stored with TMA, written back to TMEM, and
//               accumulated into the row's shifted sum / sum of squares; pass 1 of half 0 overlaps the MMAs of half 1;
//       pass 2 (64-column chunks): v is read back from TMEM, normalised, packed to bf16 and TMA-stored to xn.
// Rounding points are those of the unfused pair (TMA reduce-add epilogue + layernorm_kernel): fp32 x, bf16 xn.
#pragma once
#include "gemm.cuh"

namespace pq {

struct GemmLnParams {
  int M, K;
  const float* bias;     // [D] or nullptr
  const float* gamma;    // [D]
  const float* beta;     // [D]
  float eps;
  int num_m_tiles;
};

constexpr int GLN_EPI_WARPS = 4;
constexpr int GLN_THREADS = 64 + 32 * GLN_EPI_WARPS;
constexpr int GLN_SLABS = 6;          // x chunks in flight per epilogue warp

template <int D>
struct GemmLnCfg {
  static constexpr int kNH = D / 2;                                   // columns accumulated per pass over K
  static constexpr int kABytes = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;     // 16 KB
  static constexpr int kBBytes = kNH * GEMM_BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSlabBytes = GLN_EPI_WARPS * GLN_SLABS * 4096;
  static constexpr int kParamBytes = 3 * D * 4;                       // bias, gamma, beta
  static constexpr int kBarBytes = 512;
  static constexpr int kStagesRaw = (232448 - 1024 - kBarBytes - kSlabBytes - kParamBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 6 ? 6 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + kSlabBytes + kParamBytes + kBarBytes + 1024;
  static constexpr int kChunks = D / 32;                              // pass-1 chunks per row
  static_assert(D == 192 || D == 384, "full rows must fit 512 TMEM columns and the UMMA N range");
  static_assert(kNH % 16 == 0 && kNH <= 256, "UMMA N");
  static_assert(kBBytes % 1024 == 0 && kStageBytes % 1024 == 0, "1024-B aligned operand tiles");
  static_assert(kStages >= 3, "pipeline depth");
  static constexpr int kRounds = kChunks / GLN_SLABS;                 // phases every slab barrier completes per tile
  static_assert(kChunks % GLN_SLABS == 0, "whole rounds of slabs per tile");
};

template <int D>
__global__ void __launch_bounds__(GLN_THREADS, 1)
gemm_ln_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmN,
                     const GemmLnParams p) {
  using Cfg = GemmLnCfg<D>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  uint8_t* smem = smem_raw + pad;
  uint8_t* slab_base = smem + Cfg::kStages * Cfg::kStageBytes;                 // 1024-B aligned
  float* s_bias = reinterpret_cast<float*>(slab_base + Cfg::kSlabBytes);
  float* s_gamma = s_bias + D;
  float* s_beta = s_gamma + D;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(slab_base + Cfg::kSlabBytes + Cfg::kParamBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;      // [2]: accumulator half h complete
  uint64_t* tempty_bar = tfull_bar + 2;                // [1]: the epilogue is done with this tile's TMEM
  uint64_t* x_bar = tempty_bar + 1;                    // [GLN_EPI_WARPS][GLN_SLABS]: x chunk landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_bar + GLN_EPI_WARPS * GLN_SLABS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (p.K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;

  grid_dep_launch();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB); prefetch_tmap(&tmX); prefetch_tmap(&tmN);
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tfull_bar[0], 1);
    mbar_init(&tfull_bar[1], 1);
    mbar_init(&tempty_bar[0], GLN_EPI_WARPS);
    for (int i = 0; i < GLN_EPI_WARPS * GLN_SLABS; ++i) mbar_init(&x_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  // bias / gamma / beta are weights (never written by a preceding kernel): stage them before the dependency wait
  for (int j = threadIdx.x; j < D; j += GLN_THREADS) {
    s_bias[j] = (p.bias != nullptr) ? __ldg(p.bias + j) : 0.0f;
    s_gamma[j] = __ldg(p.gamma + j);
    s_beta[j] = __ldg(p.beta + j);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_m_tiles; tile += gridDim.x) {
        const int m0 = tile * GEMM_BLOCK_M;
        for (int h = 0; h < 2; ++h) {
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sa = smem + stage * Cfg::kStageBytes;
            mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * GEMM_BLOCK_K, m0);
            tma_load_2d(sa + Cfg::kABytes, &tmB, &full_bar[stage], kb * GEMM_BLOCK_K, h * Cfg::kNH);
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BLOCK_M, Cfg::kNH);
      int stage = 0;
      uint32_t phase = 0, tphase = 0;
      for (int tile = blockIdx.x; tile < p.num_m_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[0], tphase ^ 1u);        // previous tile's rows have left TMEM
        tc_fence_after();
        for (int h = 0; h < 2; ++h) {
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(h * Cfg::kNH);
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
          umma_commit(&tfull_bar[h]);
        }
        tphase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue: thread = row =====================
    const int quarter = warp & 3;
    const int ew = warp - 2;
    uint8_t* my_slabs = slab_base + ew * (GLN_SLABS * 4096);
    uint64_t* my_xbar = x_bar + ew * GLN_SLABS;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    uint32_t tphase = 0;
    uint32_t xround = 0;                               // slab-barrier phases completed before this tile
    for (int tile = blockIdx.x; tile < p.num_m_tiles; tile += gridDim.x, xround += Cfg::kRounds) {
      const int row0 = tile * GEMM_BLOCK_M + quarter * 32;
      // every slab is free here (first tile, or bulk_wait_group_read<0> at the end of the previous tile)
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < GLN_SLABS; ++c) {
          mbar_expect_tx(&my_xbar[c], 4096);
          tma_load_2d(my_slabs + c * 4096, &tmX, &my_xbar[c], c * 32, row0);
        }
      }
      float shift = 0.f, sum = 0.f, sq = 0.f;
#pragma unroll 1
      for (int c = 0; c < Cfg::kChunks; ++c) {
        if (c == 0 || c == Cfg::kChunks / 2) {         // accumulator half ready (pass 1 of half 0 overlaps half 1's MMAs)
          mbar_wait(&tfull_bar[c == 0 ? 0 : 1], tphase);
          tc_fence_after();
        }
        const int s = c % GLN_SLABS;
        mbar_wait(&my_xbar[s], (xround + static_cast<uint32_t>(c / GLN_SLABS)) & 1u);
        uint32_t v[32];
        tmem_ld_32x32b_x32(trow + static_cast<uint32_t>(c * 32), v);
        tmem_ld_wait();
        uint8_t* slab = my_slabs + s * 4096;
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
          if (c == 0 && jj == 0) shift = r.x;          // shifted single-pass variance (shift = first element of the row)
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
          tma_store_2d(&tmX, slab, c * 32, row0);
          bulk_commit_group();
          // the slab of chunk c-1 is reusable once its store (the second newest group) has read it
          if (c >= 1 && c - 1 + GLN_SLABS < Cfg::kChunks) {
            bulk_wait_group_read<1>();
            const int sp = (c - 1) % GLN_SLABS;
            mbar_expect_tx(&my_xbar[sp], 4096);
            tma_load_2d(my_slabs + sp * 4096, &tmX, &my_xbar[sp], (c - 1 + GLN_SLABS) * 32, row0);
          }
        }
      }
      const float mean_d = sum * (1.0f / D);
      const float mean = shift + mean_d;
      const float var = fmaxf(sq * (1.0f / D) - mean_d * mean_d, 0.0f);
      const float rstd = 1.0f / sqrtf(var + p.eps);
      tmem_st_wait();
      if (lane == 0) bulk_wait_group_read<0>();        // every slab is free again
      __syncwarp();
      // ---- pass 2: normalise, bf16, 64 columns (128 B) per row per TMA store ----
#pragma unroll 1
      for (int c = 0; c < D / 64; ++c) {
        uint8_t* slab = my_slabs + (c % GLN_SLABS) * 4096;
        uint8_t* buf = slab + lane * 128;
        if (c >= GLN_SLABS) {                           // (D = 384: 6 chunks, never taken)
          if (lane == 0) bulk_wait_group_read<GLN_SLABS - 1>();
          __syncwarp();
        }
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
          tma_store_2d(&tmN, slab, c * 64, row0);
          bulk_commit_group();
        }
      }
      tc_fence_before();
      if (lane == 0) bulk_wait_group_read<0>();        // slabs free for the next tile's x loads
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[0]);
      tphase ^= 1u;
    }
    if (lane == 0) bulk_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace pq
