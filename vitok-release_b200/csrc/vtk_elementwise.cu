// vtk_elementwise.cu -- HBM-bound helper kernels of the AE path (coalesced, 16-byte vectorised).
//
//   rmsnorm      Block.norm1            vitok/models/modules/norm.py:17-25   (bf16 in/out, fp32 math)
//   rope_table   compute_2d_freqs_cis   vitok/models/modules/rotary_embedding.py:46-75 (+ the bf16 cast at :118-119)
//   kv_len       _get_attn_mask         vitok/models/ae.py:173-187 reduced to a per-image key count
//   casts        callers' .to(bf16)     scripts/train_vae.py:305-306
#include "vtk_common.cuh"
#include "vtk_kernels.h"

namespace vtk {

// one warp per row; D % 8 == 0.  Two passes over the row (second pass hits L1).
__global__ void __launch_bounds__(256) rmsnorm_kernel(const bf16* __restrict__ x, long long ldx,
                                                      const bf16* __restrict__ w, bf16* __restrict__ y, long long ldy,
                                                      int M, int D, float eps) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nvec = D >> 3;
  for (long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < M;
       row += (long long)gridDim.x * warps_per_block) {
    const bf16* xr = x + row * ldx;
    float ss = 0.f;
    for (int v = lane; v < nvec; v += 32) {
      uint4 q = ld_global_v4(xr + 8 * v);
      const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float a = bf16_lo(u[i]), b = bf16_hi(u[i]);
        ss += a * a + b * b;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss / (float)D + eps);
    bf16* yr = y + row * ldy;
    for (int v = lane; v < nvec; v += 32) {
      uint4 q = ld_global_v4(xr + 8 * v);
      uint4 g = ld_global_nc_v4(w + 8 * v);
      const uint32_t u[4] = {q.x, q.y, q.z, q.w};
      const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
      uint32_t o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        o[i] = pack_bf16x2(bf16_lo(u[i]) * rstd * bf16_lo(gw[i]), bf16_hi(u[i]) * rstd * bf16_hi(gw[i]));
      st_global_v4(yr + 8 * v, o[0], o[1], o[2], o[3]);
    }
  }
}

// Single-pass variant: the row (D <= NV * 256 elements) stays in registers between the sum of squares and the
// scaling, so x is read from HBM/L2 exactly once.  One warp per row, 16-byte vectors, lane-strided (coalesced).
template <int NV>
__global__ void __launch_bounds__(256) rmsnorm_reg_kernel(const bf16* __restrict__ x, long long ldx,
                                                          const bf16* __restrict__ w, bf16* __restrict__ y,
                                                          long long ldy, int M, int D, float eps) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nvec = D >> 3;
  const float inv_d = 1.f / (float)D;
  for (long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < M;
       row += (long long)gridDim.x * warps_per_block) {
    const bf16* xr = x + row * ldx;
    uint4 q[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      q[i] = v < nvec ? ld_global_v4(xr + 8 * v) : make_uint4(0, 0, 0, 0);
    }
    uint64_t ss2 = 0ull;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint32_t u[4] = {q[i].x, q[i].y, q[i].z, q[i].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t t = bf2_to_f2(u[k]);
        ss2 = f2_fma(t, t, ss2);
      }
    }
    float s0, s1;
    f2_unpack(ss2, s0, s1);
    float ss = s0 + s1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss * inv_d + eps);
    const uint64_t rstd2 = f2_pack(rstd, rstd);
    bf16* yr = y + row * ldy;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < nvec) {
        const uint4 g = ld_global_nc_v4(w + 8 * v);
        const uint32_t u[4] = {q[i].x, q[i].y, q[i].z, q[i].w};
        const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float a, b;
          f2_unpack(f2_mul(f2_mul(bf2_to_f2(u[k]), rstd2), bf2_to_f2(gw[k])), a, b);
          o[k] = bf2_cvt(a, b);
        }
        st_global_v4(yr + 8 * v, o[0], o[1], o[2], o[3]);
      }
    }
  }
}

int launch_rmsnorm(const bf16* x, long long ldx, const bf16* w, bf16* y, long long ldy, int M, int D, float eps,
                   cudaStream_t stream) {
  if (D % 8 || ldx % 8 || ldy % 8) { set_error("rmsnorm: D and strides must be multiples of 8"); return -2; }
  if (M <= 0) return 0;
  const int wpb = 8;
  long long blocks = ((long long)M + wpb - 1) / wpb;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (D <= 1024) rmsnorm_reg_kernel<4><<<(int)blocks, wpb * 32, 0, stream>>>(x, ldx, w, y, ldy, M, D, eps);
  else if (D <= 3072) rmsnorm_reg_kernel<12><<<(int)blocks, wpb * 32, 0, stream>>>(x, ldx, w, y, ldy, M, D, eps);
  else if (D <= 4096) rmsnorm_reg_kernel<16><<<(int)blocks, wpb * 32, 0, stream>>>(x, ldx, w, y, ldy, M, D, eps);
  else rmsnorm_kernel<<<(int)blocks, wpb * 32, 0, stream>>>(x, ldx, w, y, ldy, M, D, eps);
  return check_cuda(cudaGetLastError(), "rmsnorm launch");
}

// Pair-expanded, chunk-major RoPE table consumed by the QKV GEMM epilogue (vtk_gemm.cu, rope_row_ptr):
//   per token row 2d bf16 = C2[d] = (c_0, c_0, c_1, c_1, ...) | S2[d] = (-s_0, +s_0, -s_1, +s_1, ...),
//   c_j = bf16(cos(angle_j)), s_j = bf16(sin(angle_j)), angle_j = row*f_j for j < d/4, col*f_{j-d/4} for j >= d/4
//   (fp32 angles like the reference, rotary_embedding.py:46-75; the bf16 cast is the one at :118-119);
//   bf16 element offset(row m, element e) = ((m >> 5) * (d / 4) + (e >> 3)) * 256 + (m & 31) * 8 + (e & 7).
// Rows are grouped by 32 so that one warp of the epilogue (32 consecutive rows) reads each 16-byte chunk
// as 512 contiguous bytes.  The table holds ceil(M / 32) * 32 rows.
__global__ void rope_table_kernel(const int64_t* __restrict__ row_idx, const int64_t* __restrict__ col_idx,
                                  const float* __restrict__ inv_freq, bf16* __restrict__ table, int M, int d) {
  const int half = d >> 1, quarter = d >> 2;
  const long long groups = ((long long)M + 31) >> 5;
  const long long total = groups * half * 32;   // thread <-> (group, pair j, row-in-group): coalesced-ish writes
  uint32_t* tw = reinterpret_cast<uint32_t*>(table);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ml = (int)(i & 31);
    const long long t = i >> 5;
    const int j = (int)(t % half);
    const long long grp = t / half;
    const long long m = grp * 32 + ml;
    if (m >= M) continue;
    const float pos = j < quarter ? (float)row_idx[m] : (float)col_idx[m];
    const float f = inv_freq[j < quarter ? j : j - quarter];
    const float ang = pos * f;
    const uint32_t c = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(cosf(ang)));
    const uint32_t sn = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(sinf(ang)));
    // 32-bit word w of the row (w < d/2: C2 pair w; w >= d/2: S2 pair w - d/2) lives in chunk w/4, slot w%4
    const long long base = grp * (long long)(d >> 2);
    const int wc = j, ws = half + j;
    tw[((base + (wc >> 2)) * 32 + ml) * 4 + (wc & 3)] = c | (c << 16);
    tw[((base + (ws >> 2)) * 32 + ml) * 4 + (ws & 3)] = (sn ^ 0x8000u) | (sn << 16);
  }
}

int launch_rope_table(const int64_t* row_idx, const int64_t* col_idx, const float* inv_freq, bf16* table, int M, int d,
                      cudaStream_t stream) {
  if (d % 8) { set_error("rope: head dimension must be a multiple of 8 (2D RoPE needs d %% 4 == 0; table chunks need 8)"); return -2; }
  if (M <= 0) return 0;
  const long long total = (((long long)M + 31) >> 5) * (d / 2) * 32;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  rope_table_kernel<<<(int)blocks, 256, 0, stream>>>(row_idx, col_idx, inv_freq, table, M, d);
  return check_cuda(cudaGetLastError(), "rope_table launch");
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long n) {
  const long long nvec = n >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const float4 a = reinterpret_cast<const float4*>(in)[2 * v];
    const float4 b = reinterpret_cast<const float4*>(in)[2 * v + 1];
    st_global_v4(out + 8 * v, pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
  }
  for (long long i = (nvec << 3) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void cast_bf16_f32_kernel(const bf16* __restrict__ in, float* __restrict__ out, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = __bfloat162float(in[i]);
}
static int cast_grid(long long n) {
  long long blocks = (n / 8 + 255) / 256 + 1;
  const long long cap = (long long)num_sms() * 16;
  return (int)(blocks > cap ? cap : blocks);
}
int launch_cast_f32_bf16(const float* in, bf16* out, long long n, cudaStream_t stream) {
  if (n <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) { set_error("cast: pointers must be 16-byte aligned"); return -2; }
  cast_f32_bf16_kernel<<<cast_grid(n), 256, 0, stream>>>(in, out, n);
  return check_cuda(cudaGetLastError(), "cast launch");
}
int launch_cast_bf16_f32(const bf16* in, float* out, long long n, cudaStream_t stream) {
  if (n <= 0) return 0;
  cast_bf16_f32_kernel<<<cast_grid(n), 256, 0, stream>>>(in, out, n);
  return check_cuda(cudaGetLastError(), "cast launch");
}

// one warp per image: kv_len = 1 + index of the last valid token (0 if none); is_prefix = mask is all-ones
// on [0, kv_len).
__global__ void kv_len_kernel(const uint8_t* __restrict__ mask, int* __restrict__ kv_len, int* __restrict__ is_prefix,
                              int B, int N) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  int last = -1, count = 0;
  for (int i = lane; i < N; i += 32) {
    if (mask[(long long)b * N + i]) { last = i; ++count; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
    count += __shfl_xor_sync(0xffffffffu, count, o);
  }
  if (lane == 0) {
    kv_len[b] = last + 1;
    if (is_prefix) is_prefix[b] = (count == last + 1) ? 1 : 0;
  }
}
int launch_kv_len(const uint8_t* mask, int* kv_len, int* is_prefix, int B, int N, cudaStream_t stream) {
  if (B <= 0) return 0;
  kv_len_kernel<<<(B + 3) / 4, 128, 0, stream>>>(mask, kv_len, is_prefix, B, N);
  return check_cuda(cudaGetLastError(), "kv_len launch");
}

}  // namespace vtk
