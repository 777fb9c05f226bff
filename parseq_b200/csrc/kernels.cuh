// Non-GEMM kernels of the PARSeq path: patch gather (im2col), LayerNorm, ViT attention core,
// decoder two-stream self-attention over the (position, token) K/V table, cross-attention over the
// cached image K/V, greedy argmax / refine-context construction, early-exit step count.
#pragma once
#include "ptx.cuh"

namespace pq {

// ---------------------------------------------------------------------------------------------
// Patch gather: images fp32 NCHW [B,3,H,W] -> A_pe bf16 [B*T, Kp] with token t = r*gw + c and
// k = ch*ph*pw + dy*pw + dx  (Conv2d(3,D,k=s=patch) as a GEMM; timm PatchEmbed via modules.py:145-161).
// One thread per (token, ch, dy): reads pw contiguous floats, writes pw contiguous bf16.
__global__ void im2col_patch_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int H,
                                    int W, int ph, int pw, int gh, int gw) {
  grid_dep_launch();
  grid_dep_wait();
  const long long total = static_cast<long long>(B) * gh * gw * 3 * ph;
  const int Kp = 3 * ph * pw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    // order: b, r, dy, ch, c  -> consecutive threads walk along an image row (coalesced reads)
    long long t = i;
    const int c = static_cast<int>(t % gw); t /= gw;
    const int ch = static_cast<int>(t % 3); t /= 3;
    const int dy = static_cast<int>(t % ph); t /= ph;
    const int r = static_cast<int>(t % gh); t /= gh;
    const int b = static_cast<int>(t);
    const float* src = img + ((static_cast<long long>(b) * 3 + ch) * H + (r * ph + dy)) * W + c * pw;
    __nv_bfloat16* dst = out + (static_cast<long long>(b) * gh * gw + r * gw + c) * Kp + ch * ph * pw + dy * pw;
    if (pw == 8 && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src));
      const float4 d = __ldg(reinterpret_cast<const float4*>(src) + 1);
      uint4 q;
      q.x = pack_bf16(a.x, a.y); q.y = pack_bf16(a.z, a.w);
      q.z = pack_bf16(d.x, d.y); q.w = pack_bf16(d.z, d.w);
      *reinterpret_cast<uint4*>(dst) = q;
    } else {
      for (int dx = 0; dx < pw; ++dx) dst[dx] = __float2bfloat16_rn(src[dx]);
    }
  }
}

// Same gather for raw uint8 HWC crops [B, H, W, 3] with the reference's input transform folded in:
// T.ToTensor() (u / 255) followed by T.Normalize(0.5, 0.5) ((x - 0.5) / 0.5)  (strhub/data/module.py:68-82), evaluated in
// fp32 with IEEE division exactly like torchvision, then rounded to bf16 like the float path.
__global__ void im2col_patch_u8_kernel(const uint8_t* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int H,
                                       int W, int ph, int pw, int gh, int gw) {
  grid_dep_launch();
  grid_dep_wait();
  const long long total = static_cast<long long>(B) * gh * gw * ph;     // one thread per (token, dy): pw*3 bytes
  const int Kp = 3 * ph * pw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long t = i;
    const int c = static_cast<int>(t % gw); t /= gw;
    const int dy = static_cast<int>(t % ph); t /= ph;
    const int r = static_cast<int>(t % gh); t /= gh;
    const int b = static_cast<int>(t);
    const uint8_t* src = img + ((static_cast<long long>(b) * H + (r * ph + dy)) * W + c * pw) * 3;
    __nv_bfloat16* dst = out + (static_cast<long long>(b) * gh * gw + r * gw + c) * Kp + dy * pw;
    for (int dx = 0; dx < pw; ++dx) {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const float x = __fdiv_rn(static_cast<float>(src[dx * 3 + ch]), 255.0f);
        dst[ch * ph * pw + dx] = __float2bfloat16_rn(__fdiv_rn(x - 0.5f, 0.5f));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over the last dim (biased variance, two-pass in registers), one warp per row.
// y_bf16 = bf16(LN(x)); optional fp32 copy (encoder output `memory`).
// Optional pre-add: x_row += add[(row % add_mod)] (broadcast table, e.g. pos_queries) and the sum is written back
// to xw (the residual stream) before normalising - lets the producing GEMM use its plain TMA-store epilogue.
template <int D>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, int M,
                                                        __nv_bfloat16* __restrict__ y, float* __restrict__ y32,
                                                        const float* __restrict__ add, int add_mod,
                                                        float* __restrict__ xw) {
  grid_dep_launch();
  grid_dep_wait();
  static_assert(D % 64 == 0, "D must be a multiple of 64");
  constexpr int NV = D / 64;  // float2 per lane
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const float2* xr = reinterpret_cast<const float2*>(x + static_cast<long long>(row) * D);
  float2 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = xr[i * 32 + lane];
  if (add != nullptr) {
    const float2* ar = reinterpret_cast<const float2*>(add + static_cast<long long>(row % add_mod) * D);
    float2* wr = reinterpret_cast<float2*>(xw + static_cast<long long>(row) * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float2 a = __ldg(ar + i * 32 + lane);
      v[i].x += a.x;
      v[i].y += a.y;
      wr[i * 32 + lane] = v[i];
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) s += v[i].x + v[i].y;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean;
    q += a * a + b * b;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = 1.0f / sqrtf(q * (1.0f / D) + eps);
  const float2* g2 = reinterpret_cast<const float2*>(gamma);
  const float2* b2 = reinterpret_cast<const float2*>(beta);
  uint32_t* yr = reinterpret_cast<uint32_t*>(y + static_cast<long long>(row) * D);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float2 g = __ldg(g2 + i * 32 + lane), b = __ldg(b2 + i * 32 + lane);
    const float o0 = (v[i].x - mean) * rstd * g.x + b.x;
    const float o1 = (v[i].y - mean) * rstd * g.y + b.y;
    yr[i * 32 + lane] = pack_bf16(o0, o1);
    if (y32 != nullptr) reinterpret_cast<float2*>(y32 + static_cast<long long>(row) * D)[i * 32 + lane] = make_float2(o0, o1);
  }
}

// ---------------------------------------------------------------------------------------------
// ViT attention core for T = 128 tokens, head dim 64 (timm Attention: softmax(QK^T / 8) V, no mask).
// One CTA per (image, head), 8 warps x 16 query rows.  Q/K/V tiles (128x64 bf16) are staged in
// XOR-swizzled shared memory with cp.async; S = QK^T and O = PV run on mma.sync m16n8k16 with the
// probabilities kept in registers (bf16 A fragments), fp32 row statistics, O/rowsum -> bf16.
constexpr int ATT_T = 128;
constexpr int ATT_DH = 64;
__device__ __forceinline__ uint32_t att_swz(int r, int c) {  // element offset of (row r, col c), c%8==0 chunks
  return static_cast<uint32_t>(r * ATT_DH + ((((c >> 3) ^ (r & 7)) << 3) | (c & 7)));
}
__global__ void __launch_bounds__(256, 2) enc_attention_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                            __nv_bfloat16* __restrict__ out, int D, int heads) {
  grid_dep_launch();
  grid_dep_wait();
  __shared__ __align__(128) __nv_bfloat16 sQ[ATT_T * ATT_DH];
  __shared__ __align__(128) __nv_bfloat16 sK[ATT_T * ATT_DH];
  __shared__ __align__(128) __nv_bfloat16 sV[ATT_T * ATT_DH];
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long ld = 3ll * D;
  const __nv_bfloat16* base = qkv + static_cast<long long>(b) * ATT_T * ld + h * ATT_DH;
  // 3 matrices x 128 rows x 8 chunks of 16 B
  for (int i = tid; i < 3 * ATT_T * 8; i += 256) {
    const int m = i / (ATT_T * 8);
    const int r = (i / 8) % ATT_T;
    const int ck = i % 8;
    const __nv_bfloat16* src = base + static_cast<long long>(r) * ld + m * D + ck * 8;
    __nv_bfloat16* dstm = (m == 0) ? sQ : (m == 1) ? sK : sV;
    cp_async_16(smem_u32(dstm + att_swz(r, ck * 8)), src);
  }
  cp_async_wait_all();
  __syncthreads();

  const int r0 = warp * 16;
  // ---- Q fragments for the 4 k-steps over d ----
  uint32_t qf[4][4];
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
    const int row = r0 + (lane & 7) + ((lane >> 3) & 1) * 8;
    const int col = kt * 16 + (lane >> 4) * 8;
    ldmatrix_x4(smem_u32(sQ + att_swz(row, col)), qf[kt][0], qf[kt][1], qf[kt][2], qf[kt][3]);
  }
  // ---- S = Q K^T : 16 n-tiles of 8 keys ----
  float sacc[16][4];
#pragma unroll
  for (int nt = 0; nt < 16; ++nt) { sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f; }
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
#pragma unroll
    for (int np = 0; np < 8; ++np) {
      const int key = np * 16 + (lane & 7) + (lane >> 4) * 8;
      const int col = kt * 16 + ((lane >> 3) & 1) * 8;
      uint32_t b0, b1, b2, b3;
      ldmatrix_x4(smem_u32(sK + att_swz(key, col)), b0, b1, b2, b3);
      mma_bf16_16816(sacc[2 * np], qf[kt][0], qf[kt][1], qf[kt][2], qf[kt][3], b0, b1);
      mma_bf16_16816(sacc[2 * np + 1], qf[kt][0], qf[kt][1], qf[kt][2], qf[kt][3], b2, b3);
    }
  }
  // ---- softmax over 128 keys; rows g (c0,c1) and g+8 (c2,c3) ----
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 16; ++nt) {
    mx0 = fmaxf(mx0, fmaxf(sacc[nt][0], sacc[nt][1]));
    mx1 = fmaxf(mx1, fmaxf(sacc[nt][2], sacc[nt][3]));
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  constexpr float kScaleLog2 = 0.125f * 1.4426950408889634f;  // d^-0.5 * log2(e)
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 16; ++nt) {
    sacc[nt][0] = exp2f((sacc[nt][0] - mx0) * kScaleLog2);
    sacc[nt][1] = exp2f((sacc[nt][1] - mx0) * kScaleLog2);
    sacc[nt][2] = exp2f((sacc[nt][2] - mx1) * kScaleLog2);
    sacc[nt][3] = exp2f((sacc[nt][3] - mx1) * kScaleLog2);
    sum0 += sacc[nt][0] + sacc[nt][1];
    sum1 += sacc[nt][2] + sacc[nt][3];
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  // ---- O = P V : 8 k-steps over keys, 8 n-tiles over d ----
  float oacc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { oacc[nt][0] = oacc[nt][1] = oacc[nt][2] = oacc[nt][3] = 0.f; }
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    const uint32_t a0 = pack_bf16(sacc[2 * kk][0], sacc[2 * kk][1]);
    const uint32_t a1 = pack_bf16(sacc[2 * kk][2], sacc[2 * kk][3]);
    const uint32_t a2 = pack_bf16(sacc[2 * kk + 1][0], sacc[2 * kk + 1][1]);
    const uint32_t a3 = pack_bf16(sacc[2 * kk + 1][2], sacc[2 * kk + 1][3]);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      const int key = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      const int col = np * 16 + (lane >> 4) * 8;
      uint32_t b0, b1, b2, b3;
      ldmatrix_x4_trans(smem_u32(sV + att_swz(key, col)), b0, b1, b2, b3);
      mma_bf16_16816(oacc[2 * np], a0, a1, a2, a3, b0, b1);
      mma_bf16_16816(oacc[2 * np + 1], a0, a1, a2, a3, b2, b3);
    }
  }
  // ---- normalise, stage through this warp's (now dead) Q rows, coalesced 16-B stores ----
  const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
  const int g = lane >> 2, t = lane & 3;
  __syncwarp();
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int col = nt * 8 + 2 * t;
    *reinterpret_cast<uint32_t*>(sQ + att_swz(r0 + g, col)) = pack_bf16(oacc[nt][0] * inv0, oacc[nt][1] * inv0);
    *reinterpret_cast<uint32_t*>(sQ + att_swz(r0 + g + 8, col)) = pack_bf16(oacc[nt][2] * inv1, oacc[nt][3] * inv1);
  }
  __syncwarp();
  __nv_bfloat16* obase = out + static_cast<long long>(b) * ATT_T * D + h * ATT_DH;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = i * 32 + lane;       // 16 rows x 8 chunks
    const int r = r0 + (idx >> 3), ck = idx & 7;
    const uint4 val = *reinterpret_cast<const uint4*>(sQ + att_swz(r, ck * 8));
    *reinterpret_cast<uint4*>(obase + static_cast<long long>(r) * D + ck * 8) = val;
  }
}

// ---------------------------------------------------------------------------------------------
// ViT attention core for ANY token count T (e.g. 196 = 224x224/16x16, 240 = 48x160/4x8), head dim 64: correctness
// path for the geometries that do not fill one 128-row tile per image (the T = 128 fast path is attn_tc.cuh).
// grid = (B*heads, ceil(T/128)): one CTA per 128-query tile; keys are visited in blocks of 128 in two passes
// (pass 1: row maxima, pass 2: P = exp(S - max), O += P V) so the rounding points equal the single-tile kernel's.
// Rows / keys >= T are masked; same mma.sync fragment scheme as enc_attention_kernel.
__global__ void __launch_bounds__(256, 2) enc_attention_any_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                   __nv_bfloat16* __restrict__ out, int T, int D,
                                                                   int heads) {
  grid_dep_launch();
  grid_dep_wait();
  __shared__ __align__(128) __nv_bfloat16 sQ[ATT_T * ATT_DH];
  __shared__ __align__(128) __nv_bfloat16 sK[ATT_T * ATT_DH];
  __shared__ __align__(128) __nv_bfloat16 sV[ATT_T * ATT_DH];
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int q0 = blockIdx.y * ATT_T;                         // first query row of this tile
  const int nkb = (T + ATT_T - 1) / ATT_T;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long ld = 3ll * D;
  const __nv_bfloat16* base = qkv + static_cast<long long>(b) * T * ld + h * ATT_DH;
  auto load_tile = [&](__nv_bfloat16* dst, int mat, int row0, int nrows) {   // rows clamped to T-1 (masked later)
    for (int i = tid; i < nrows * 8; i += 256) {
      const int r = i >> 3, ck = i & 7;
      int row = row0 + r;
      if (row >= T) row = T - 1;
      cp_async_16(smem_u32(dst + att_swz(r, ck * 8)), base + static_cast<long long>(row) * ld + mat * D + ck * 8);
    }
  };
  load_tile(sQ, 0, q0, ATT_T);
  cp_async_wait_all();
  __syncthreads();
  const int r0 = warp * 16;
  const bool active = (q0 + r0) < T;     // warp-uniform: a warp whose 16 query rows are all padding only helps loading
  uint32_t qf[4][4];
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
    const int row = r0 + (lane & 7) + ((lane >> 3) & 1) * 8;
    const int col = kt * 16 + (lane >> 4) * 8;
    ldmatrix_x4(smem_u32(sQ + att_swz(row, col)), qf[kt][0], qf[kt][1], qf[kt][2], qf[kt][3]);
  }
  const int t4 = lane & 3;
  constexpr float kScaleLog2 = 0.125f * 1.4426950408889634f;
  float mx0 = -INFINITY, mx1 = -INFINITY, sum0 = 0.f, sum1 = 0.f;
  float oacc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { oacc[nt][0] = oacc[nt][1] = oacc[nt][2] = oacc[nt][3] = 0.f; }
  for (int pass = 0; pass < 2; ++pass) {
    for (int kb = 0; kb < nkb; ++kb) {
      __syncthreads();                                       // previous tile fully consumed
      // only the 16-key groups that hold at least one real key are loaded and multiplied (T = 129: 1 of 8 in block 1)
      const int nvalid = (T - kb * ATT_T < ATT_T) ? (T - kb * ATT_T) : ATT_T;
      const int np_max = (nvalid + 15) >> 4;
      load_tile(sK, 1, kb * ATT_T, np_max * 16);
      if (pass == 1) load_tile(sV, 2, kb * ATT_T, np_max * 16);
      cp_async_wait_all();
      __syncthreads();
      if (!active) continue;
      float sacc[16][4];
#pragma unroll
      for (int nt = 0; nt < 16; ++nt) { sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f; }
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
#pragma unroll
        for (int np = 0; np < 8; ++np) {
          if (np >= np_max) break;
          const int key = np * 16 + (lane & 7) + (lane >> 4) * 8;
          const int col = kt * 16 + ((lane >> 3) & 1) * 8;
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4(smem_u32(sK + att_swz(key, col)), b0, b1, b2, b3);
          mma_bf16_16816(sacc[2 * np], qf[kt][0], qf[kt][1], qf[kt][2], qf[kt][3], b0, b1);
          mma_bf16_16816(sacc[2 * np + 1], qf[kt][0], qf[kt][1], qf[kt][2], qf[kt][3], b2, b3);
        }
      }
      // mask keys >= T (columns 8*nt + 2*t4, +1 of this key block)
#pragma unroll
      for (int nt = 0; nt < 16; ++nt) {
        const int k0 = kb * ATT_T + nt * 8 + 2 * t4;
        if (k0 >= T) { sacc[nt][0] = -INFINITY; sacc[nt][2] = -INFINITY; }
        if (k0 + 1 >= T) { sacc[nt][1] = -INFINITY; sacc[nt][3] = -INFINITY; }
      }
      if (pass == 0) {
#pragma unroll
        for (int nt = 0; nt < 16; ++nt) {
          mx0 = fmaxf(mx0, fmaxf(sacc[nt][0], sacc[nt][1]));
          mx1 = fmaxf(mx1, fmaxf(sacc[nt][2], sacc[nt][3]));
        }
      } else {
#pragma unroll
        for (int nt = 0; nt < 16; ++nt) {
          sacc[nt][0] = exp2f((sacc[nt][0] - mx0) * kScaleLog2);
          sacc[nt][1] = exp2f((sacc[nt][1] - mx0) * kScaleLog2);
          sacc[nt][2] = exp2f((sacc[nt][2] - mx1) * kScaleLog2);
          sacc[nt][3] = exp2f((sacc[nt][3] - mx1) * kScaleLog2);
          sum0 += sacc[nt][0] + sacc[nt][1];
          sum1 += sacc[nt][2] + sacc[nt][3];
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          if (kk >= np_max) break;           // P is exactly 0 on the padded keys
          const uint32_t a0 = pack_bf16(sacc[2 * kk][0], sacc[2 * kk][1]);
          const uint32_t a1 = pack_bf16(sacc[2 * kk][2], sacc[2 * kk][3]);
          const uint32_t a2 = pack_bf16(sacc[2 * kk + 1][0], sacc[2 * kk + 1][1]);
          const uint32_t a3 = pack_bf16(sacc[2 * kk + 1][2], sacc[2 * kk + 1][3]);
#pragma unroll
          for (int np = 0; np < 4; ++np) {
            const int key = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
            const int col = np * 16 + (lane >> 4) * 8;
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4_trans(smem_u32(sV + att_swz(key, col)), b0, b1, b2, b3);
            mma_bf16_16816(oacc[2 * np], a0, a1, a2, a3, b0, b1);
            mma_bf16_16816(oacc[2 * np + 1], a0, a1, a2, a3, b2, b3);
          }
        }
      }
    }
    if (pass == 0) {
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    }
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
  const int g = lane >> 2;
  __nv_bfloat16* obase = out + static_cast<long long>(b) * T * D + h * ATT_DH;
  const int row_lo = q0 + r0 + g, row_hi = row_lo + 8;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int col = nt * 8 + 2 * t4;
    if (row_lo < T) *reinterpret_cast<uint32_t*>(obase + static_cast<long long>(row_lo) * D + col) = pack_bf16(oacc[nt][0] * inv0, oacc[nt][1] * inv0);
    if (row_hi < T) *reinterpret_cast<uint32_t*>(obase + static_cast<long long>(row_hi) * D + col) = pack_bf16(oacc[nt][2] * inv1, oacc[nt][3] * inv1);
  }
}

// ---------------------------------------------------------------------------------------------
// Decoder context rows for the (position, token) K/V table:
//   ctx[pos*V + tok] = sqrt(D)*E[tok] + (pos >= 1 ? pos_queries[pos-1] : 0)     (model.py:94-99, modules.py:175-176)
__global__ void build_ctx_rows_kernel(const float* __restrict__ emb, const float* __restrict__ posq, float* __restrict__ ctx,
                                      int L, int V, int D, float sqrtD) {
  const long long total = static_cast<long long>(L) * V * D;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % D);
    const int tok = static_cast<int>((i / D) % V);
    const int pos = static_cast<int>(i / (static_cast<long long>(D) * V));
    float v = sqrtD * emb[static_cast<long long>(tok) * D + c];
    if (pos >= 1) v = posq[static_cast<long long>(pos - 1) * D + c] + v;
    ctx[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Greedy argmax (first maximum wins, torch.argmax semantics) of logits rows -> token ids.
// One warp per row.  Row r = (b, s): reads logits[b, src_pos0 + s, :C], writes ids[b*ids_ld + dst_pos0 + s].
// If `forced` != nullptr the written id is forced[b*forced_ld + dst_pos0 + s] (teacher forcing).
__global__ void argmax_rows_kernel(const float* __restrict__ logits, int L, int C, int B, int nrows_per_b, int src_pos0,
                                   int* __restrict__ ids, int ids_ld, int dst_pos0, const int* __restrict__ forced,
                                   int forced_ld) {
  grid_dep_launch();
  grid_dep_wait();
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= B * nrows_per_b) return;
  const int lane = threadIdx.x & 31;
  const int b = w / nrows_per_b, s = w % nrows_per_b;
  const float* row = logits + (static_cast<long long>(b) * L + src_pos0 + s) * C;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = lane; j < C; j += 32) {
    const float v = row[j];
    if (v > best) { best = v; bi = j; }    // strictly greater: keeps the lowest index within the lane
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) {
    int v = bi;
    if (forced != nullptr) v = forced[static_cast<long long>(b) * forced_ld + dst_pos0 + s];
    ids[static_cast<long long>(b) * ids_ld + dst_pos0 + s] = v;
  }
}

__global__ void fill_ids_kernel(int* __restrict__ ids, int B, int ld, int bos, int pad) {
  grid_dep_launch();
  grid_dep_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * ld) ids[i] = ((i % ld) == 0) ? bos : pad;
}

// S = number of AR steps the reference returns under its batch-wide early exit (model.py:144):
// smallest j>=1 such that every row has an EOS among ids[b,1..j]  == max_b first_eos_pos(b); L if any row has none.
__global__ void ar_steps_kernel(const int* __restrict__ ids, int ids_ld, int B, int L, int eos_id, int* __restrict__ steps) {
  grid_dep_launch();
  grid_dep_wait();
  int worst = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int first = L;
    for (int j = 1; j < L; ++j)
      if (ids[static_cast<long long>(b) * ids_ld + j] == eos_id) { first = j; break; }
    worst = max(worst, first);
  }
  atomicMax(steps, worst);
}
__global__ void set_int_kernel(int* p, int v) {
  grid_dep_launch();
  grid_dep_wait(); if (threadIdx.x == 0 && blockIdx.x == 0) *p = v; }

// ---------------------------------------------------------------------------------------------
// Decoder self-attention, one CTA per image, one warp per head (head dim 32), ALL nq queries of the pass:
// the context K rows (lane = key, <= 32 keys) and V columns (lane = channel) of the head are gathered once from the
// (position, token) table into registers, then every query costs ~130 warp instructions.
// (DecoderLayer.forward_stream step 1, modules.py:69-72)
//   q      : Qs[qpos] fp32 (W_q LN_q(pos_queries[qpos]) + b, pre-scaled by 1/sqrt(32); input independent)
//   K/V    : kvtab[(k*V + ids[b,k]) * 2D + {0, D} + c] bf16   (the content stream is a function of
//            (position, token) only at decoder depth 1)
//   mask   : mode 0 (AR step / NAR): keys 0..nkeys-1 all visible (model.py:130-136: the sliced causal row is
//            all-False);  mode 1 (cloze refine): key k masked iff k == q+1 or an EOS occurs in ids[b, 0..k]
//            (model.py:157,163)
// out bf16 [B*nq, D] (A operand of the out-projection GEMM).
//            mode 2 (PARSeq.decode with caller-supplied masks, model.py:86-103): Qs holds one query row per (image,
//            query) [B*nq, D]; key k of query qi is masked iff qmask[qi*nkeys + k] or pmask[b*nkeys + k] (either may be
//            null); a row with every key masked yields NaN, as torch's softmax over -inf does
__global__ void dec_self_attn2_kernel(const float* __restrict__ Qs, const __nv_bfloat16* __restrict__ kvtab,
                                      const int* __restrict__ ids, int ids_ld, int V, int D, int nq, int q0, int nkeys,
                                      int mode, int eos_id, __nv_bfloat16* __restrict__ out, int qsplit,
                                      const unsigned char* __restrict__ qmask = nullptr,
                                      const unsigned char* __restrict__ pmask = nullptr) {
  // grid = B * qsplit: CTA (b, part) handles queries [part*nq/qsplit, (part+1)*nq/qsplit) of image b
  __shared__ int s_ids[32];
  __shared__ int s_first_eos;
  grid_dep_launch();
  grid_dep_wait();
  const int b = blockIdx.x / qsplit, part = blockIdx.x % qsplit;
  const int q_begin = (part * nq) / qsplit, q_end = ((part + 1) * nq) / qsplit;
  const int lane = threadIdx.x & 31;
  const int heads = D >> 5, nwarps = blockDim.x >> 5;   // blockDim.x = min(D, 384): heads are looped when D > 384
  if (threadIdx.x < 32) {
    const int id = (threadIdx.x < nkeys) ? ids[static_cast<long long>(b) * ids_ld + threadIdx.x] : -1;
    s_ids[threadIdx.x] = id;
    const unsigned m = __ballot_sync(0xffffffffu, id == eos_id);
    if (threadIdx.x == 0) s_first_eos = (m != 0u) ? (__ffs(m) - 1) : (1 << 30);
  }
  __syncthreads();
  const int first_eos = s_first_eos;
  for (int h = threadIdx.x >> 5; h < heads; h += nwarps) {
  float kreg[32], vreg[32];
  if (lane < nkeys) {
    const uint4* kr = reinterpret_cast<const uint4*>(kvtab + (static_cast<long long>(lane) * V + s_ids[lane]) * 2 * D + h * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 u = __ldg(kr + j);
      const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(p2[e]);
        kreg[j * 8 + e * 2] = f.x;
        kreg[j * 8 + e * 2 + 1] = f.y;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) kreg[j] = 0.f;
  }
#pragma unroll
  for (int k = 0; k < 32; ++k)
    vreg[k] = (k < nkeys) ? __bfloat162float(kvtab[(static_cast<long long>(k) * V + s_ids[k]) * 2 * D + D + h * 32 + lane]) : 0.f;
  for (int qi = q_begin; qi < q_end; ++qi) {
    const int qpos = q0 + qi;
    const long long qrow = (mode == 2) ? (static_cast<long long>(b) * nq + qi) : qpos;
    const float qv = __ldg(Qs + qrow * D + h * 32 + lane);   // lane j holds q_j
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) s = fmaf(__shfl_sync(0xffffffffu, qv, j), kreg[j], s);
    bool masked = (lane >= nkeys) || ((mode == 1) && (lane == qpos + 1 || lane >= first_eos));
    if (mode == 2 && lane < nkeys) {
      if (qmask != nullptr && qmask[qi * nkeys + lane] != 0) masked = true;
      if (pmask != nullptr && pmask[static_cast<long long>(b) * nkeys + lane] != 0) masked = true;
    }
    if (masked) s = -INFINITY;
    float mx = s;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float e = masked ? 0.f : expf(s - mx);
    float sum = e;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float pme = e / sum;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) acc = fmaf(__shfl_sync(0xffffffffu, pme, k), vreg[k], acc);
    out[(static_cast<long long>(b) * nq + qi) * D + h * 32 + lane] = __float2bfloat16_rn(acc);
  }
  }
}

// ---------------------------------------------------------------------------------------------
// Decoder cross-attention: one CTA per (image, head), 4 warps.  The head's K (padded pitch, lane = key reads are
// conflict-free) and V tiles are staged once in shared memory, so the per-image K/V cache is read once per decode
// pass; the nq queries are distributed over the warps and each query is handled entirely inside one warp (scores
// for 4 keys per lane, shuffle softmax, lane = channel for P.V): no block-level synchronisation after the load.
// q fp32 [B*nq, D] pre-scaled by 1/sqrt(32); kv bf16 [B, T, 2D]; out bf16 [B*nq, D].  T <= 128, head dim 32.
template <int NR>   // keys per lane: T <= 32 * NR (NR = 4: T <= 128, NR = 8: T <= 256)
__global__ void __launch_bounds__(128) dec_cross_attn3_kernel(const float* __restrict__ q,
                                                              const __nv_bfloat16* __restrict__ kv, long long kv_rows,
                                                              int b_first, int T, int D, int heads, int nq,
                                                              __nv_bfloat16* __restrict__ out) {
  // kv: column-blocked cross K/V cache [2D/64][kv_rows][64] (ptx.cuh: blocked_off), row = image * T + key; the
  // queries / outputs of this launch belong to images b_first, b_first + 1, ...
  constexpr int TK = 32 * NR;
  __shared__ uint32_t sK[TK * 17];                        // bf16x2 words, pitch 17 (odd)
  __shared__ __align__(16) __nv_bfloat16 sV[TK * 32];
  grid_dep_launch();
  grid_dep_wait();
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long row_b = static_cast<long long>(b_first + b) * T;
  for (int t = tid; t < TK; t += 128) {
    uint4* vd = reinterpret_cast<uint4*>(sV + t * 32);
    if (t < T) {
      const uint4* kr = reinterpret_cast<const uint4*>(kv + blocked_off(kv_rows, row_b + t, h * 32));
      const uint4* vr = reinterpret_cast<const uint4*>(kv + blocked_off(kv_rows, row_b + t, D + h * 32));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 u = __ldg(kr + j);
        sK[t * 17 + j * 4 + 0] = u.x; sK[t * 17 + j * 4 + 1] = u.y;
        sK[t * 17 + j * 4 + 2] = u.z; sK[t * 17 + j * 4 + 3] = u.w;
        vd[j] = __ldg(vr + j);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) sK[t * 17 + j] = 0u;
#pragma unroll
      for (int j = 0; j < 4; ++j) vd[j] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  __syncthreads();
  for (int qi = warp; qi < nq; qi += 4) {
    const long long row = static_cast<long long>(b) * nq + qi;
    const float qv = q[row * D + h * 32 + lane];          // lane j holds q_j
    float sc[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) sc[r] = 0.f;
#pragma unroll
    for (int w = 0; w < 16; ++w) {
      const float qa = __shfl_sync(0xffffffffu, qv, 2 * w), qb = __shfl_sync(0xffffffffu, qv, 2 * w + 1);
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const uint32_t kw = sK[(r * 32 + lane) * 17 + w];
        sc[r] = fmaf(qb, __uint_as_float(kw & 0xffff0000u), fmaf(qa, __uint_as_float(kw << 16), sc[r]));
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      if (r * 32 + lane >= T) sc[r] = -INFINITY;
      mx = fmaxf(mx, sc[r]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      sc[r] = expf(sc[r] - mx);
      sum += sc[r];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    // P.V with 16-byte smem reads: lane = (key group kg = lane>>2, 8-channel chunk cc = lane&3); 4*NR iterations cover
    // the keys; the 8 key groups are then summed with xor-shuffles and lanes 0..3 hold the 32 output channels.
    const int kg = lane >> 2, cc = lane & 3;
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll
    for (int it = 0; it < 4 * NR; ++it) {              // key = it*8 + kg -> register sc[it>>2], source lane (it&3)*8 + kg
      const int key = it * 8 + kg;
      const uint4 vvv = *reinterpret_cast<const uint4*>(sV + key * 32 + cc * 8);
      const float pk = __shfl_sync(0xffffffffu, sc[it >> 2], (it & 3) * 8 + kg);
      const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&vvv);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(p2[e]);
        o[e * 2] = fmaf(pk, f.x, o[e * 2]);
        o[e * 2 + 1] = fmaf(pk, f.y, o[e * 2 + 1]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j] += __shfl_xor_sync(0xffffffffu, o[j], 4);
      o[j] += __shfl_xor_sync(0xffffffffu, o[j], 8);
      o[j] += __shfl_xor_sync(0xffffffffu, o[j], 16);
    }
    if (kg == 0) {
      const float inv = 1.0f / sum;
      uint4 q4;
      q4.x = pack_bf16(o[0] * inv, o[1] * inv); q4.y = pack_bf16(o[2] * inv, o[3] * inv);
      q4.z = pack_bf16(o[4] * inv, o[5] * inv); q4.w = pack_bf16(o[6] * inv, o[7] * inv);
      *reinterpret_cast<uint4*>(out + row * D + h * 32 + cc * 8) = q4;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// ViTSTR (vitstr/model.py:14-28 over timm VisionTransformer._pos_embed): token 0 of every image is
// cls_token + pos_embed[0]; tokens 1..Tp are the patch embeddings (pos_embed[1..Tp] already added by the patch GEMM).
__global__ void cls_assemble_kernel(const float4* __restrict__ patches, const float4* __restrict__ cls,
                                    const float4* __restrict__ pos0, float4* __restrict__ x, int B, int Tp, int D4) {
  grid_dep_launch();
  grid_dep_wait();                 // launched with the PDL attribute: the patch GEMM's output is complete from here on
  const long long total = static_cast<long long>(B) * (Tp + 1) * D4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % D4);
    const long long row = i / D4;
    const int t = static_cast<int>(row % (Tp + 1));
    const long long b = row / (Tp + 1);
    float4 v;
    if (t == 0) {
      const float4 a = cls[c], p = pos0[c];
      v = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    } else {
      v = patches[(b * Tp + (t - 1)) * D4 + c];
    }
    x[i] = v;
  }
}
// out[b, j, :] = x[b, first + j, :], j < n (the `x[:, :seqlen]` / `logits[:, 1:]` slices of vitstr/model.py:21,
// vitstr/system.py:70 applied BEFORE norm + head: both are row-wise, so only the kept rows are computed).
__global__ void gather_token_rows_kernel(const float4* __restrict__ x, float4* __restrict__ out, int B, int T, int first,
                                         int n, int D4) {
  grid_dep_launch();
  grid_dep_wait();
  const long long total = static_cast<long long>(B) * n * D4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % D4);
    const long long row = i / D4;
    const int j = static_cast<int>(row % n);
    const long long b = row / n;
    out[i] = x[(b * T + first + j) * D4 + c];
  }
}

// ---------------------------------------------------------------------------------------------
// Fused tail of a decoder pass: out = LayerNorm(y; decoder.norm) -> logits = head(out) -> greedy argmax
// (modules.py:124 `Decoder.norm`, model.py:138 `self.head`, model.py:142 argmax).  4 rows per CTA; the head weight
// (C x D bf16, 73 KB for 95 x 384) is staged in shared memory with an odd word pitch (lane = class reads are
// conflict-free); each warp computes a (32 classes x 4 rows) partial over a quarter of K; the normalised rows are
// rounded to bf16 exactly like a tensor-core A operand.
// logits fp32: row r -> logits[r * logits_ld .. + C); ids (optional): row r = (b, qi) -> ids[b*ids_ld + dst_off + qi].
constexpr int HEAD_ROWS = 4;
template <int D>
__global__ void __launch_bounds__(384) dec_ln_head_argmax_kernel(
    const float* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
    const __nv_bfloat16* __restrict__ Wh, const float* __restrict__ bh, int M, int C, float* __restrict__ logits,
    long long logits_ld, int* __restrict__ ids, int ids_ld, int nq, int dst_off, const int* __restrict__ forced,
    int forced_ld) {
  extern __shared__ __align__(16) unsigned char head_smem[];
  constexpr int WP = D / 2 + 1;                         // weight row pitch in 32-bit words (odd -> conflict-free)
  uint32_t* sW = reinterpret_cast<uint32_t*>(head_smem);            // [C][WP] bf16x2
  float* sy = reinterpret_cast<float*>(head_smem + ((static_cast<size_t>(C) * WP * 4 + 15) / 16) * 16);   // [4][D]
  float* spart = sy + HEAD_ROWS * D;                                // [4 k-quarters][4 rows][128]
  float* sl = spart + 4 * HEAD_ROWS * 128;                          // [4][128] logits
  grid_dep_launch();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nwarps = blockDim.x >> 5;
  // (weights do not depend on the previous kernel: stage them before the dependency wait)
  const uint32_t* Wg = reinterpret_cast<const uint32_t*>(Wh);
  for (int i = tid; i < C * (D / 2); i += blockDim.x) {
    const int c = i / (D / 2), k = i % (D / 2);
    sW[c * WP + k] = __ldg(Wg + i);
  }
  grid_dep_wait();
  const int row0 = blockIdx.x * HEAD_ROWS;
  // ---- LayerNorm of up to 4 rows (warp per row), rounded to bf16 ----
  constexpr int NV = D / 64;
  if (warp < HEAD_ROWS) {
    const int r = warp;
    const int row = row0 + r;
    float2 v[NV];
    if (row < M) {
      const float2* xr = reinterpret_cast<const float2*>(y + static_cast<long long>(row) * D);
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = xr[i * 32 + lane];
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = make_float2(0.f, 0.f);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += v[i].x + v[i].y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / D);
    float qv = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = v[i].x - mean, b2 = v[i].y - mean;
      qv += a * a + b2 * b2;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) qv += __shfl_xor_sync(0xffffffffu, qv, o);
    const float rstd = 1.0f / sqrtf(qv * (1.0f / D) + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float2 g = __ldg(reinterpret_cast<const float2*>(gamma) + i * 32 + lane);
      const float2 bb = __ldg(reinterpret_cast<const float2*>(beta) + i * 32 + lane);
      const float o0 = (v[i].x - mean) * rstd * g.x + bb.x;
      const float o1 = (v[i].y - mean) * rstd * g.y + bb.y;
      reinterpret_cast<float2*>(sy + r * D)[i * 32 + lane] =
          make_float2(__bfloat162float(__float2bfloat16_rn(o0)), __bfloat162float(__float2bfloat16_rn(o1)));
    }
  }
  __syncthreads();
  // ---- head partials: unit = (class pass of 32, quarter of K); lane = class; 4 rows at once ----
  const int passes = (C + 31) / 32;
  constexpr int KQ = D / 8;                               // words per K quarter
  for (int u = warp; u < passes * 4; u += nwarps) {
    const int c = (u >> 2) * 32 + lane;
    const int kq = u & 3;
    const int cc = (c < C) ? c : (C - 1);
    const uint32_t* wr = sW + cc * WP + kq * KQ;
    const float* y0 = sy + kq * KQ * 2;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
    for (int k = 0; k < KQ; ++k) {
      const uint32_t w2 = wr[k];
      const float wx = __uint_as_float(w2 << 16), wy = __uint_as_float(w2 & 0xffff0000u);
      const float2 p0 = *reinterpret_cast<const float2*>(y0 + 2 * k);
      const float2 p1 = *reinterpret_cast<const float2*>(y0 + D + 2 * k);
      const float2 p2 = *reinterpret_cast<const float2*>(y0 + 2 * D + 2 * k);
      const float2 p3 = *reinterpret_cast<const float2*>(y0 + 3 * D + 2 * k);
      a0 = fmaf(wy, p0.y, fmaf(wx, p0.x, a0));
      a1 = fmaf(wy, p1.y, fmaf(wx, p1.x, a1));
      a2 = fmaf(wy, p2.y, fmaf(wx, p2.x, a2));
      a3 = fmaf(wy, p3.y, fmaf(wx, p3.x, a3));
    }
    float* sp = spart + (kq * HEAD_ROWS) * 128 + (c & 127);
    sp[0] = a0; sp[128] = a1; sp[256] = a2; sp[384] = a3;
  }
  __syncthreads();
  for (int i = tid; i < HEAD_ROWS * C; i += blockDim.x) {
    const int r = i / C, c = i % C;
    const float l = ((spart[(0 * HEAD_ROWS + r) * 128 + c] + spart[(1 * HEAD_ROWS + r) * 128 + c]) +
                     (spart[(2 * HEAD_ROWS + r) * 128 + c] + spart[(3 * HEAD_ROWS + r) * 128 + c])) + __ldg(bh + c);
    sl[r * 128 + c] = l;
    const int row = row0 + r;
    if (row < M) logits[static_cast<long long>(row) * logits_ld + c] = l;
  }
  if (ids == nullptr) return;
  __syncthreads();
  // ---- greedy argmax (first maximum wins), warp per row ----
  if (warp < HEAD_ROWS) {
    const int r = warp;
    const int row = row0 + r;
    if (row < M) {
      float best = -INFINITY;
      int bi = 0x7fffffff;
      for (int j = lane; j < C; j += 32) {
        const float v = sl[r * 128 + j];
        if (v > best) { best = v; bi = j; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (lane == 0) {
        const int b = row / nq, qi = row % nq;
        int v = bi;
        if (forced != nullptr) v = forced[static_cast<long long>(b) * forced_ld + dst_off + qi];
        ids[static_cast<long long>(b) * ids_ld + dst_off + qi] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fused post-processing of the reference's test path (strhub/models/base.py:132-142 + Tokenizer._filter,
// strhub/data/utils.py:120-129): per image  ids[i] = argmax_c logits[i, c]  (first maximum),  length = index of the
// first EOS (L if none),  confidence = prod_{i <= min(length, L-1)} max_c softmax(logits[i])_c  (the EOS probability is
// included).  One warp per image; only (length, confidence, ids) cross PCIe instead of the [B, L, C] probabilities.
__global__ void postprocess_kernel(const float* __restrict__ logits, int B, int L, int C, int eos_id, int* __restrict__ ids,
                                   int* __restrict__ lengths, float* __restrict__ confidence) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  float conf = 1.0f;
  int len = L;
  bool done = false;
  for (int i = 0; i < L; ++i) {
    const float* row = logits + (static_cast<long long>(b) * L + i) * C;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int j = lane; j < C; j += 32) {
      const float v = row[j];
      if (v > best) { best = v; bi = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    float se = 0.f;
    for (int j = lane; j < C; j += 32) se += expf(row[j] - best);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    if (lane == 0) ids[static_cast<long long>(b) * L + i] = bi;
    if (!done) {
      conf *= 1.0f / se;                 // max softmax probability of position i
      if (bi == eos_id) { len = i; done = true; }
    }
  }
  if (lane == 0) {
    lengths[b] = len;
    confidence[b] = conf;
  }
}

// ---------------------------------------------------------------------------------------------
// Boundary helpers of the module API (PARSeq.decode / head / text_embed called on their own).
__global__ void f32_to_bf16_kernel(const float4* __restrict__ x, uint2* __restrict__ y, long long n4) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = x[i];
    y[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}
// dst[b, :] = src[:] for b < B (n floats, n % 4 == 0)
__global__ void bcast_rows_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int n4, int B) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < static_cast<long long>(B) * n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] = src[i % n4];
}
// ids [B, J] (row pitch J) -> context buffer [B, 32]
__global__ void copy_ids_kernel(const int* __restrict__ src, int J, int* __restrict__ dst, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * 32) dst[i] = ((i & 31) < J) ? src[(i >> 5) * J + (i & 31)] : 0;
}
// TokenEmbedding.forward (modules.py:175-176): out[i, :] = sqrt(D) * E[ids[i], :]
__global__ void text_embed_kernel(const int* __restrict__ ids, const float* __restrict__ E, float* __restrict__ out, int n,
                                  int D, int V, float scale) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < static_cast<long long>(n) * D;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    int tok = ids[i / D];
    tok = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);
    out[i] = scale * E[static_cast<long long>(tok) * D + (i % D)];
  }
}

}  // namespace pq
