// ViT attention core on tcgen05 (T = 128 tokens, head dim 64): softmax(Q K^T / 8) V for one (image, head) per CTA.
//   TMA: Q, K, V tiles [128 x 64] bf16 (128B-swizzled) straight out of the packed qkv activation [B*128, 3D].
//   MMA1 (UMMA 128x128x16, 4 k-steps): S = Q K^T -> TMEM columns [0,128)      (both operands K-major)
//   softmax: one thread per query row reads S from TMEM (tcgen05.ld), fp32 max / exp2 / sum, writes P as bf16 into
//            shared memory in the K-major SW128 operand layout (overwriting the dead Q/K tiles)
//   MMA2 (UMMA 128x64x16, 8 k-steps): O = P V -> TMEM columns [0,64)           (A = P K-major, B = V MN-major)
//   epilogue: O / rowsum -> bf16 -> swizzled staging (over the dead V tile) -> TMA store into att [B*128, D].
// Barriers are single use (one tile per CTA); 48 KB smem and 128 TMEM columns per CTA -> 4 CTAs per SM overlap
// their load / MMA / softmax phases.  Rounding points identical to enc_attention_kernel (P rounded to bf16, row sum
// from the unrounded fp32 exponentials, output rounded to bf16).
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace pq {

constexpr int ATC_THREADS = 192;   // warp0: TMA + MMA issue, warp1: TMEM alloc, warps 2-5: softmax / epilogue

// MN-major (N contiguous) B operand, 128B swizzle: rows = K index (keys), each row = 64 contiguous N elements (128 B);
// 8-row groups are SBO = 1024 B apart (cute::UMMA canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in uint128 units).
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>(1024u >> 4) << 16;    // LBO: stride between 64-element N blocks (only one block here)
  d |= static_cast<uint64_t>(1024u >> 4) << 32;    // SBO: stride between 8-key groups
  d |= static_cast<uint64_t>(1u) << 46;
  d |= static_cast<uint64_t>(2u) << 61;            // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(ATC_THREADS, 4) enc_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV,
                                                                       const __grid_constant__ CUtensorMap tmO, int D,
                                                                       int heads) {
  extern __shared__ uint8_t atc_raw[];
  const uint32_t raw_addr = smem_u32(atc_raw);
  uint8_t* smem = atc_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;                 // 16 KB  (later: P k-block 0)
  uint8_t* sK = smem + 16384;         // 16 KB  (later: P k-block 1)
  uint8_t* sV = smem + 32768;         // 16 KB  (later: output staging)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 49152);   // [0] full, [1] s_full, [2] p_full, [3] o_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;

  grid_dep_launch();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQKV);
    prefetch_tmap(&tmO);
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 128);
    mbar_init(&bars[3], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  grid_dep_wait();

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&bars[0], 3 * 16384);
      tma_load_2d(sQ, &tmQKV, &bars[0], h * 64, b * 128);
      tma_load_2d(sK, &tmQKV, &bars[0], D + h * 64, b * 128);
      tma_load_2d(sV, &tmQKV, &bars[0], 2 * D + h * 64, b * 128);
      mbar_wait(&bars[0], 0);
      tc_fence_after();
      {   // S = Q K^T
        constexpr uint32_t idesc = make_idesc_bf16(128, 128);
        const uint64_t adesc = make_desc_k_sw128(smem_u32(sQ)), bdesc = make_desc_k_sw128(smem_u32(sK));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc, k != 0);
        umma_commit(&bars[1]);
      }
      mbar_wait(&bars[2], 0);          // P is in shared memory (generic-proxy writes fenced by the writers)
      tc_fence_after();
      {   // O = P V : A = P (two K-major [128 x 64] tiles over sQ, sK), B = V (MN-major)
        constexpr uint32_t idesc = make_idesc_bf16(128, 64) | (1u << 16);   // b_major = MN
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t adesc = make_desc_k_sw128(smem_u32(ks < 4 ? sQ : sK)) + static_cast<uint64_t>(2 * (ks & 3));
          const uint64_t bdesc = make_desc_mn_sw128(smem_u32(sV) + static_cast<uint32_t>(ks) * 2048u);
          umma_bf16(tmem, adesc, bdesc, idesc, ks != 0);
        }
        umma_commit(&bars[3]);
      }
    }
  } else if (warp >= 2) {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;                 // query row owned by this thread
    const uint32_t taddr = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    constexpr float kScaleLog2 = 0.125f * 1.4426950408889634f;
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    float sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c * 32), v);
      tmem_ld_wait();
      uint8_t* prow = (c < 2 ? sQ : sK) + row * 128;     // P k-block (c >> 1), this row
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float f[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          f[t] = ex2_approx((__uint_as_float(v[jj * 8 + t]) - mx) * kScaleLog2);
          sum += f[t];
        }
        uint4 q;
        q.x = pack_bf16(f[0], f[1]); q.y = pack_bf16(f[2], f[3]);
        q.z = pack_bf16(f[4], f[5]); q.w = pack_bf16(f[6], f[7]);
        const uint32_t chunk = static_cast<uint32_t>((c & 1) * 4 + jj);
        *reinterpret_cast<uint4*>(prow + ((chunk ^ sw) << 4)) = q;
      }
    }
    fence_proxy_async_smem();          // P (generic proxy) -> visible to the tensor core (async proxy)
    tc_fence_before();                 // TMEM reads of S complete before MMA2 overwrites the columns
    mbar_arrive(&bars[2]);
    mbar_wait(&bars[3], 0);
    tc_fence_after();
    const float inv = 1.0f / sum;
    uint8_t* slab = sV + quarter * 4096;                 // [32 rows][128 B], V is dead after MMA2
    uint8_t* orow = slab + lane * 128;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(hh * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        uint4 q;
        q.x = pack_bf16(__uint_as_float(v[jj * 8 + 0]) * inv, __uint_as_float(v[jj * 8 + 1]) * inv);
        q.y = pack_bf16(__uint_as_float(v[jj * 8 + 2]) * inv, __uint_as_float(v[jj * 8 + 3]) * inv);
        q.z = pack_bf16(__uint_as_float(v[jj * 8 + 4]) * inv, __uint_as_float(v[jj * 8 + 5]) * inv);
        q.w = pack_bf16(__uint_as_float(v[jj * 8 + 6]) * inv, __uint_as_float(v[jj * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(orow + ((static_cast<uint32_t>(hh * 4 + jj) ^ static_cast<uint32_t>(lane & 7)) << 4)) = q;
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(&tmO, slab, h * 64, b * 128 + quarter * 32);
      bulk_commit_group();
      bulk_wait_group<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<128>(tmem);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// General token count T <= 256 (ViTSTR: 129 tokens, parseq-patch16-224: 196, the 48x160 ViT-B-width config: 240): the
// same tcgen05 pipeline with NKB = 2 key blocks of 128 and one CTA per (image, head, 128-query tile).  qkv and the
// output are addressed through 3D tensor maps [image][token][channel], so rows past the image's T tokens read as zeros
// and are clipped on store; key columns >= T are masked in the softmax.
//   MMA1: S[:, 128 kb ..] = Q K_kb^T (UMMA 128x128x16 x 4 per key block) -> TMEM columns [0, 128 NKB)
//   P (bf16) -> NKB * 2 K-major [128 x 64] tiles over the dead Q / K tiles (+ one extra tile)
//   MMA2: O = P V (UMMA 128x64x16 x 8 NKB, V MN-major) -> TMEM columns [0, 64)
// 96 KB shared memory and 256 TMEM columns per CTA -> 2 CTAs per SM.
template <int NKB>
__global__ void __launch_bounds__(ATC_THREADS, 2) enc_attention_tc2_kernel(const __grid_constant__ CUtensorMap tmQKV,
                                                                        const __grid_constant__ CUtensorMap tmO, int D,
                                                                        int heads, int T) {
  extern __shared__ uint8_t atc_raw[];
  const uint32_t raw_addr = smem_u32(atc_raw);
  uint8_t* smem = atc_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;                                   // 16 KB            (later: P tile 0)
  uint8_t* sK = smem + 16384;                           // NKB x 16 KB      (later: P tiles 1 ..)
  uint8_t* sV = sK + NKB * 16384;                       // NKB x 16 KB      (later: output staging)
  uint8_t* sX = sV + NKB * 16384;                       // 16 KB: last P tile (NKB = 2: tiles 0..2 cover Q + K, tile 3 here)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sX + 16384);   // [0] full, [1] s_full, [2] p_full, [3] o_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  auto ptile = [&](int i) -> uint8_t* { return (i < 1 + NKB) ? (smem + i * 16384) : sX; };   // P k-tile i (64 keys)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int q0 = blockIdx.y * 128;                      // first query token of this CTA

  grid_dep_launch();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQKV);
    prefetch_tmap(&tmO);
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 128);
    mbar_init(&bars[3], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<128 * NKB>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  grid_dep_wait();

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&bars[0], (1 + 2 * NKB) * 16384);
      tma_load_3d(sQ, &tmQKV, &bars[0], h * 64, q0, b);
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb) {
        tma_load_3d(sK + kb * 16384, &tmQKV, &bars[0], D + h * 64, kb * 128, b);
        tma_load_3d(sV + kb * 16384, &tmQKV, &bars[0], 2 * D + h * 64, kb * 128, b);
      }
      mbar_wait(&bars[0], 0);
      tc_fence_after();
      {   // S = Q K^T, one 128-column block per key block
        constexpr uint32_t idesc = make_idesc_bf16(128, 128);
        const uint64_t adesc = make_desc_k_sw128(smem_u32(sQ));
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb) {
          const uint64_t bdesc = make_desc_k_sw128(smem_u32(sK + kb * 16384));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + static_cast<uint32_t>(kb * 128), adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k),
                      idesc, k != 0);
        }
        umma_commit(&bars[1]);
      }
      mbar_wait(&bars[2], 0);          // P is in shared memory (generic-proxy writes fenced by the writers)
      tc_fence_after();
      {   // O = P V : A = P (K-major [128 x 64] tiles), B = V (MN-major, 16 keys = 2048 B per k-step)
        constexpr uint32_t idesc = make_idesc_bf16(128, 64) | (1u << 16);   // b_major = MN
#pragma unroll
        for (int ks = 0; ks < 8 * NKB; ++ks) {
          const uint64_t adesc = make_desc_k_sw128(smem_u32(ptile(ks >> 2))) + static_cast<uint64_t>(2 * (ks & 3));
          const uint64_t bdesc = make_desc_mn_sw128(smem_u32(sV + (ks >> 3) * 16384) + static_cast<uint32_t>(ks & 7) * 2048u);
          umma_bf16(tmem, adesc, bdesc, idesc, ks != 0);
        }
        umma_commit(&bars[3]);
      }
    }
  } else if (warp >= 2) {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;                 // query row (inside the tile) owned by this thread
    const uint32_t taddr = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    constexpr float kScaleLog2 = 0.125f * 1.4426950408889634f;
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < 4 * NKB; ++c) {
      if (c * 32 >= T) break;                            // uniform
      uint32_t v[32];
      tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c * 32 + j < T) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    float sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < 4 * NKB; ++c) {
      uint32_t v[32];
      const bool live = c * 32 < T;                      // uniform; dead chunks write zeros (their P tiles are multiplied)
      if (live) {
        tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c * 32), v);
        tmem_ld_wait();
      }
      uint8_t* prow = ptile(c >> 1) + row * 128;         // P k-tile (c >> 1), this row
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float f[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const bool ok = live && (c * 32 + jj * 8 + t < T);
          f[t] = ok ? ex2_approx((__uint_as_float(v[jj * 8 + t]) - mx) * kScaleLog2) : 0.0f;
          sum += f[t];
        }
        uint4 q;
        q.x = pack_bf16(f[0], f[1]); q.y = pack_bf16(f[2], f[3]);
        q.z = pack_bf16(f[4], f[5]); q.w = pack_bf16(f[6], f[7]);
        const uint32_t chunk = static_cast<uint32_t>((c & 1) * 4 + jj);
        *reinterpret_cast<uint4*>(prow + ((chunk ^ sw) << 4)) = q;
      }
    }
    fence_proxy_async_smem();          // P (generic proxy) -> visible to the tensor core (async proxy)
    tc_fence_before();                 // TMEM reads of S complete before MMA2 overwrites the columns
    mbar_arrive(&bars[2]);
    mbar_wait(&bars[3], 0);
    tc_fence_after();
    const float inv = 1.0f / sum;
    uint8_t* slab = sV + quarter * 4096;                 // [32 rows][128 B], V is dead after MMA2
    uint8_t* orow = slab + lane * 128;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(hh * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        uint4 q;
        q.x = pack_bf16(__uint_as_float(v[jj * 8 + 0]) * inv, __uint_as_float(v[jj * 8 + 1]) * inv);
        q.y = pack_bf16(__uint_as_float(v[jj * 8 + 2]) * inv, __uint_as_float(v[jj * 8 + 3]) * inv);
        q.z = pack_bf16(__uint_as_float(v[jj * 8 + 4]) * inv, __uint_as_float(v[jj * 8 + 5]) * inv);
        q.w = pack_bf16(__uint_as_float(v[jj * 8 + 6]) * inv, __uint_as_float(v[jj * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(orow + ((static_cast<uint32_t>(hh * 4 + jj) ^ static_cast<uint32_t>(lane & 7)) << 4)) = q;
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && q0 + quarter * 32 < T) {            // rows past T inside the box are clipped by the tensor map
      tma_store_3d(&tmO, slab, h * 64, q0 + quarter * 32, b);
      bulk_commit_group();
      bulk_wait_group<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<128 * NKB>(tmem);
  }
}

template <int NKB>
constexpr int atc2_smem_bytes() { return (2 + 2 * NKB) * 16384 + 1024 + 64; }

constexpr int ATC_SMEM_BYTES = 49152 + 1024 + 64;

}  // namespace pq
