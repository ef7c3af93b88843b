// vtk_attention.cu -- masked variable-length flash-style attention forward on tcgen05 (sm_100a).
//
// Replaces modules/attention.py:109-127 (flash_attn_func / F.scaled_dot_product_attention) of the
// reference: out = softmax(q k^T / sqrt(d)) v per (image, head), q/k already QK-normed + RoPE'd by the
// QKV GEMM epilogue.  The reference's sdpa backend masks keys with patch_mask (ae.py:173-187) by
// materialising a [B,1,N,N] bool mask; here the mask is a per-image key length (padded keys are never
// loaded -- whole kv tiles past kv_len[b] are skipped) plus an optional per-key byte mask for
// non-prefix masks.  The flash backend's semantics (no mask) = kv_len null.
//
// One CTA per (128-query tile, head, image), 6 warps:
//   warp 0 : TMA producer (Q once, K/V tiles of 128 keys, 2-stage ring each)
//   warp 1 : MMA issuer   S = Q K^T  (128 x 128 x d)  and  O += P V  (128 x d x 128), fp32 in TMEM
//   warps 2-5 : softmax, one query row per thread (tcgen05.ld 32x32b): online max/sum in fp32, P -> bf16
//               -> 128B-swizzled smem (A operand of the PV MMA), lazy rescale of O in TMEM.
// S is double-buffered in TMEM so S_{j+1} is computed while softmax_j runs.
// V is consumed as an MN-major B operand straight from its natural [key, d] layout (no transpose).
#include <math.h>
#include <stdio.h>

#include "vtk_common.cuh"
#include "vtk_kernels.h"

namespace vtk {

static constexpr int ATT_BQ = 128;   // queries per CTA
static constexpr int ATT_BKV = 128;  // keys per tile
static constexpr int BLK = 16384;    // one [128 x 64] bf16 swizzled block

template <int DH> struct AttnShape {
  static constexpr int NB = DH / 64;               // 64-column blocks per head
  static constexpr int NP = (DH == 64) ? 2 : 1;    // P buffers
  static constexpr int Q_BYTES = NB * BLK;
  static constexpr int KV_BYTES = NB * BLK;        // one K (or V) tile
  static constexpr int P_BYTES = 2 * BLK;          // 128 x 128 bf16
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_BYTES;
  static constexpr int OFF_V = OFF_K + 2 * KV_BYTES;
  static constexpr int OFF_P = OFF_V + 2 * KV_BYTES;
  static constexpr int OFF_BAR = OFF_P + NP * P_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static constexpr uint32_t TMEM_COLS = 512;       // S0 [0,128) S1 [128,256) O [256, 256+DH)
};

struct AttnParams {
  bf16* out; long long ld_out;
  const int* kv_len; const uint8_t* key_mask; const int* prefix_flag;
  int N, heads, zero_invalid;
  float scale_log2;   // (1/sqrt(d)) * log2(e)
  int q_col0, k_col0, v_col0;  // column offsets of head 0 inside the respective tensor maps
};

template <int DH>
__global__ void __launch_bounds__(192, 1)
attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
            const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using S = AttnShape<DH>;
  const int q0 = blockIdx.x * ATT_BQ;
  const int head = blockIdx.y;
  const int img = blockIdx.z;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int N = p.N;
  int kvlen = p.kv_len ? p.kv_len[img] : N;
  kvlen = kvlen < N ? kvlen : N;
  const long long row0 = (long long)img * N;

  const int T = (kvlen + ATT_BKV - 1) / ATT_BKV;
  if (T == 0 || (p.zero_invalid && q0 >= kvlen)) {
    // nothing to attend to (padded query tile / empty image): define the output as 0
    if (warp >= 2) {
      const int r = ((warp & 3) << 5) + lane;
      if (q0 + r < N) {
        bf16* op = p.out + (row0 + q0 + r) * p.ld_out + head * DH;
        for (int c = 0; c < DH; c += 8) st_global_v4(op + c, 0u, 0u, 0u, 0u);
      }
    }
    return;
  }

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + S::OFF_Q;
  uint8_t* sK = smem + S::OFF_K;
  uint8_t* sV = smem + S::OFF_V;
  uint8_t* sP = smem + S::OFF_P;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]
  uint64_t* s_empty = bars + 11;  // [2]
  uint64_t* p_full = bars + 13;   // [2]
  uint64_t* p_empty = bars + 15;  // [2]
  uint64_t* o_done = bars + 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 128);
      mbar_init(&p_full[i], 128);
      mbar_init(&p_empty[i], 1);
    }
    mbar_init(o_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, S::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;          // + 128 * buffer
  const uint32_t tmem_O = tmem_base + 256;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      mbar_expect_tx(q_full, S::Q_BYTES);
      for (int nb = 0; nb < S::NB; ++nb)
        tma_load_2d(sQ + nb * BLK, &tmQ, q_full, p.q_col0 + head * DH + nb * 64, (int)(row0 + q0));
      for (int j = 0; j < T; ++j) {
        const int st = j & 1;
        const uint32_t use_ph = (uint32_t)(j >> 1) & 1u;
        const int krow = (int)(row0 + (long long)j * ATT_BKV);
        mbar_wait(&k_empty[st], use_ph ^ 1);
        mbar_expect_tx(&k_full[st], S::KV_BYTES);
        for (int nb = 0; nb < S::NB; ++nb)
          tma_load_2d(sK + st * S::KV_BYTES + nb * BLK, &tmK, &k_full[st], p.k_col0 + head * DH + nb * 64, krow);
        mbar_wait(&v_empty[st], use_ph ^ 1);
        mbar_expect_tx(&v_full[st], S::KV_BYTES);
        for (int nb = 0; nb < S::NB; ++nb)
          tma_load_2d(sV + st * S::KV_BYTES + nb * BLK, &tmV, &v_full[st], p.v_col0 + head * DH + nb * 64, krow);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      const uint32_t idesc_s = make_idesc_bf16(ATT_BQ, ATT_BKV, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(ATT_BQ, DH, 0, 1);   // B (= V) is MN-major
      auto issue_pv = [&](int i) {
        const int pb = i % S::NP;
        const int vs = i & 1;
        mbar_wait(&p_full[pb], (uint32_t)(i / S::NP) & 1u);
        mbar_wait(&v_full[vs], (uint32_t)(i >> 1) & 1u);
        tc_fence_after();
        const uint32_t pa = smem_u32(sP + pb * S::P_BYTES);
        const uint32_t va = smem_u32(sV + vs * S::KV_BYTES);
#pragma unroll
        for (int kk = 0; kk < ATT_BKV / 16; ++kk) {
          const uint64_t adesc = make_desc_kmajor_sw128(pa + (kk >> 2) * BLK + (kk & 3) * 32);
          // V tile: NB blocks of [128 keys x 64 d]; MN-major: LBO = block stride, SBO = 8-key group stride
          const uint64_t bdesc = make_smem_desc(va + kk * 2048, BLK, 1024, 2);
          umma_bf16_ss(tmem_O, adesc, bdesc, idesc_o, (i | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&v_empty[vs]);
        umma_commit(&p_empty[pb]);
        umma_commit(o_done);
      };
      mbar_wait(q_full, 0);
      for (int j = 0; j < T; ++j) {
        const int sb = j & 1;
        const uint32_t use_ph = (uint32_t)(j >> 1) & 1u;
        mbar_wait(&k_full[sb], use_ph);
        mbar_wait(&s_empty[sb], use_ph ^ 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(sQ);
        const uint32_t ka = smem_u32(sK + sb * S::KV_BYTES);
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk) {
          const uint32_t off = (kk >> 2) * BLK + (kk & 3) * 32;
          umma_bf16_ss(tmem_S + sb * 128, make_desc_kmajor_sw128(qa + off), make_desc_kmajor_sw128(ka + off), idesc_s,
                       kk != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[sb]);
        umma_commit(&k_empty[sb]);
        if (j > 0) issue_pv(j - 1);
      }
      issue_pv(T - 1);
    }
  } else {
    // ===== softmax (warps 2..5): thread <-> query row =====
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const bool general_mask = p.key_mask != nullptr && !(p.prefix_flag != nullptr && p.prefix_flag[img] != 0);
    const uint8_t* kmask = general_mask ? p.key_mask + row0 : nullptr;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < T; ++j) {
      const int sb = j & 1;
      const int kv0 = j * ATT_BKV;
      const bool need_mask = (kv0 + ATT_BKV > kvlen) || (kmask != nullptr);
      mbar_wait(&s_full[sb], (uint32_t)(j >> 1) & 1u);
      __syncwarp();
      tc_fence_after();
      const uint32_t ts = tmem_S + lane_base + sb * 128;
      // pass 1: row max
      float mx = -INFINITY;
      for (int c = 0; c < ATT_BKV; c += 32) {
        uint32_t v[32];
        tmem_ld32(ts + c, v);
        tmem_wait_ld();
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int kc = kv0 + c + i;
            const bool ok = kc < kvlen && (kmask == nullptr || kmask[kc] != 0);
            mx = fmaxf(mx, ok ? __uint_as_float(v[i]) : -INFINITY);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
      }
      const float m_new = fmaxf(m_run, mx * p.scale_log2);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = exp2f(m_run - m_use);   // m_run = -inf -> 0
      // P buffer free?
      const int pb = j % S::NP;
      mbar_wait(&p_empty[pb], ((uint32_t)(j / S::NP) & 1u) ^ 1u);
      __syncwarp();
      uint8_t* prow = sP + pb * S::P_BYTES + r * 128;
      float rowsum = 0.f;
      for (int c = 0; c < ATT_BKV; c += 32) {
        uint32_t v[32];
        tmem_ld32(ts + c, v);
        tmem_wait_ld();
        uint32_t o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float e0 = exp2f(__uint_as_float(v[2 * i]) * p.scale_log2 - m_use);
          float e1 = exp2f(__uint_as_float(v[2 * i + 1]) * p.scale_log2 - m_use);
          if (need_mask) {
            const int kc = kv0 + c + 2 * i;
            if (!(kc < kvlen && (kmask == nullptr || kmask[kc] != 0))) e0 = 0.f;
            if (!(kc + 1 < kvlen && (kmask == nullptr || kmask[kc + 1] != 0))) e1 = 0.f;
          }
          rowsum += e0 + e1;
          o[i] = pack_bf16x2(e0, e1);
        }
        // 32 columns = 4 chunks of 16 B; swizzle-128B: chunk' = chunk ^ (row & 7)
        uint8_t* pblk = prow + (c >> 6) * BLK;
        const int ch0 = (c & 63) >> 3;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int ch = (ch0 + g) ^ (r & 7);
          *reinterpret_cast<uint4*>(pblk + (ch << 4)) = make_uint4(o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(&s_empty[sb]);   // S buffer consumed
      l_run = l_run * alpha + rowsum;
      m_run = m_new;
      // rescale O (needs PV_{j-1} complete)
      if (j > 0) {
        mbar_wait(o_done, (uint32_t)(j - 1) & 1u);
        __syncwarp();
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.f)) {
          for (int c = 0; c < DH; c += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_O + lane_base + c, v);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st32(tmem_O + lane_base + c, v);
          }
          tmem_wait_st();
        }
      }
      fence_proxy_async_smem();   // P (generic-proxy smem writes) -> visible to the MMA (async proxy)
      tc_fence_before();
      mbar_arrive(&p_full[pb]);
    }
    // epilogue: O / l
    mbar_wait(o_done, (uint32_t)(T - 1) & 1u);
    __syncwarp();
    tc_fence_after();
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    const int qi = q0 + r;
    bool zero_row = false;
    if (p.zero_invalid) zero_row = (qi >= kvlen) || (kmask != nullptr && qi < N && kmask[qi] == 0);
    const float sc = zero_row ? 0.f : inv;
    bf16* op = p.out + (row0 + qi) * p.ld_out + head * DH;
    for (int c = 0; c < DH; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_O + lane_base + c, v);
      tmem_wait_ld();
      if (qi < N) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            o[i] = pack_bf16x2(__uint_as_float(v[8 * g + 2 * i]) * sc, __uint_as_float(v[8 * g + 2 * i + 1]) * sc);
          st_global_v4(op + c + 8 * g, o[0], o[1], o[2], o[3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, S::TMEM_COLS);
  }
}

template <int DH>
static int launch_attention_t(const AttnArgs& a, cudaStream_t stream) {
  using S = AttnShape<DH>;
  const long long Mrows = (long long)a.B * a.N;
  const long long cols = (long long)a.heads * a.d;
  CUtensorMap tmQ, tmK, tmV;
  if (encode_tmap_bf16_sw128(&tmQ, a.q, cols, Mrows, a.ld_qkv, ATT_BQ)) return -1;
  if (encode_tmap_bf16_sw128(&tmK, a.k, cols, Mrows, a.ld_qkv, ATT_BKV)) return -1;
  if (encode_tmap_bf16_sw128(&tmV, a.v, cols, Mrows, a.ld_qkv, ATT_BKV)) return -1;
  AttnParams p;
  p.out = a.out; p.ld_out = a.ld_out;
  p.kv_len = a.kv_len; p.key_mask = a.key_mask; p.prefix_flag = a.prefix_flag;
  p.N = a.N; p.heads = a.heads; p.zero_invalid = a.zero_invalid_rows;
  p.scale_log2 = (float)((1.0 / sqrt((double)a.d)) * 1.4426950408889634);
  p.q_col0 = p.k_col0 = p.v_col0 = 0;
  auto kern = attn_kernel<DH>;
  static bool attr_set = false;
  if (!attr_set) {
    if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM_BYTES),
                   "cudaFuncSetAttribute(attn)"))
      return -1;
    attr_set = true;
  }
  dim3 grid((a.N + ATT_BQ - 1) / ATT_BQ, a.heads, a.B);
  kern<<<grid, 192, S::SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
  return check_cuda(cudaGetLastError(), "attention launch");
}

int launch_attention(const AttnArgs& a, cudaStream_t stream) {
  if (a.B <= 0 || a.N <= 0 || a.heads <= 0) { set_error("attention: empty problem"); return -2; }
  if ((a.ld_qkv % 8) || (a.ld_out % 8)) { set_error("attention: row strides must be multiples of 8"); return -2; }
  if (a.d == 64) return launch_attention_t<64>(a, stream);
  if (a.d == 128) return launch_attention_t<128>(a, stream);
  set_error("attention: head_dim %d unsupported (64 or 128)", a.d);
  return -3;
}

}  // namespace vtk
