// Persistent warp-specialised bf16 GEMM on tcgen05 (sm_100a):
//   out[M,N] = epilogue( A[M,K] * W[N,K]^T + bias )
// A and W are bf16, K-contiguous ("K-major"); both are fetched by TMA into 128B-swizzled shared-memory
// stages; one elected thread issues tcgen05.mma (UMMA 128 x BLOCK_N x 16) into a double-buffered fp32
// accumulator in TMEM; four epilogue warps drain TMEM with tcgen05.ld (one accumulator row per thread)
// while the next tile's MMAs run.  Every projection of the PARSeq path goes through this kernel:
// patch-embed (K=96), QKV / proj / fc1 / fc2 of the 12 ViT blocks (reference: timm Attention/Mlp via
// strhub/models/parseq/modules.py:145-165), the cross-attention K/V projection of the image memory,
// the decoder's q / out projections, MLP (modules.py:69-77) and the character head (model.py:63).
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace pq {

enum GemmEpilogue : int {
  EPI_F32 = 0,        // out_f32 = alpha*(acc+bias) (+ resid[row or row%resid_mod])
  EPI_BF16 = 1,       // out_bf16 = bf16(alpha*(acc+bias))
  EPI_GELU_BF16 = 2,  // out_bf16 = bf16(gelu(acc+bias))
};

struct GemmParams {
  int M, N, K;
  int mode;
  float alpha;
  const float* bias;   // [N] or nullptr
  const float* resid;  // fp32 residual (may alias out) or nullptr
  long long ldr;
  int resid_mod;       // >0: residual row = row % resid_mod (broadcast tables: pos_embed, pos_queries)
  void* out;
  long long ldo;       // elements
  int vec_ok;          // 16B-aligned rows: vector stores allowed
  int num_m_tiles, num_n_tiles;
};

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;   // 64 bf16 = 128 B = one swizzle row
constexpr int GEMM_THREADS = 192;  // warp0: TMA, warp1: MMA + TMEM alloc, warps 2-5: epilogue

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kStages = (BLOCK_N <= 128) ? 6 : 4;
  static constexpr int kABytes = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;  // 16 KB
  static constexpr int kBBytes = BLOCK_N * GEMM_BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(BLOCK_N % 32 == 0 && BLOCK_N >= 32 && BLOCK_N <= 256, "BLOCK_N");
};

template <int BLOCK_N>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  uint8_t* smem = smem_raw + pad;                         // 1024-B aligned (SWIZZLE_128B requirement)
  uint8_t* bar_base = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);         // [kStages] TMA -> MMA
  uint64_t* empty_bar = full_bar + Cfg::kStages;                       // [kStages] MMA -> TMA
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;                      // [2] MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;                                // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int num_kb = (p.K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.num_n_tiles) * GEMM_BLOCK_M;
        const int n0 = (tile % p.num_n_tiles) * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * GEMM_BLOCK_K, m0);
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * GEMM_BLOCK_K, n0);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[as], aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BLOCK_N);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t adesc = make_desc_k_sw128(sa);
          const uint64_t bdesc = make_desc_k_sw128(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the swizzle row: +2 in the (addr>>4) field
            umma_bf16(tmem_d, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc,
                      static_cast<uint32_t>((kb | k) != 0));
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tfull_bar[as]);       // accumulator complete
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int row_in_tile = quarter * 32 + lane;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / p.num_n_tiles) * GEMM_BLOCK_M;
      const int n0 = (tile % p.num_n_tiles) * BLOCK_N;
      const int row = m0 + row_in_tile;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(as * BLOCK_N);
      const long long rrow = (p.resid_mod > 0) ? (row % p.resid_mod) : row;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= p.N) break;            // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c * 32), v);
        tmem_ld_wait();
        if (row < p.M) {
          const bool full = p.vec_ok && (col0 + 32 <= p.N);
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (p.bias != nullptr) {
            if (col0 + 32 <= p.N) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
                f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) f[j] += __ldg(p.bias + col0 + j);
            }
          }
          if (p.mode == EPI_GELU_BF16) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
          } else if (p.alpha != 1.0f) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] *= p.alpha;
          }
          if (p.mode == EPI_F32) {
            float* o = reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo + col0;
            const float* r = (p.resid != nullptr) ? (p.resid + rrow * p.ldr + col0) : nullptr;
            if (full) {
              if (r != nullptr) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 x = *reinterpret_cast<const float4*>(r + j);
                  f[j] += x.x; f[j + 1] += x.y; f[j + 2] += x.z; f[j + 3] += x.w;
                }
              }
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) o[j] = f[j] + ((r != nullptr) ? r[j] : 0.0f);
            }
          } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(row) * p.ldo + col0;
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 q;
                q.x = pack_bf16(f[j], f[j + 1]);
                q.y = pack_bf16(f[j + 2], f[j + 3]);
                q.z = pack_bf16(f[j + 4], f[j + 5]);
                q.w = pack_bf16(f[j + 6], f[j + 7]);
                *reinterpret_cast<uint4*>(o + j) = q;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) o[j] = __float2bfloat16_rn(f[j]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);        // 128 arrivals free this accumulator stage
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace pq
