// vtk_pp.cu -- NaFlex patchify / unpatchify / index packing as coalesced, vectorised gather kernels.
//
//   patchify     vitok/pp/ops.py:217-285 + patch_collate_fn (vitok/data.py:77-94): writes straight into the
//                batched [B, T, 3p^2] buffers (no per-image dicts, no torch.stack), builds patch_mask /
//                row_idx / col_idx / time_idx / orig_height / orig_width / grid_rows / grid_cols.
//                Optional fused front end: uint8 HWC -> to_tensor | normalize(minus_one_to_one)
//                (ops.py:140-155) and bf16 output.
//   unpatchify   vitok/pp/ops.py:295-335: scatter + fold restated as a gather through a cell->token map,
//                with _convert_format (vitok/pp/io.py:91-121) optionally fused.
// All integer/byte work is bit-exact against the reference.
#include "vtk_common.cuh"
#include "vtk_kernels.h"

namespace vtk {

// ToTensor (u / 255, fp32 division) then Normalize(0.5, 0.5) ((t - 0.5) / 0.5 == (t - 0.5) * 2 exactly), ops.py:140-161
__device__ __forceinline__ float norm_u8(uint32_t u) {
  return __fmul_rn(__fsub_rn(__fdiv_rn((float)u, 255.0f), 0.5f), 2.0f);
}

template <typename OutT>
__device__ __forceinline__ void store4(OutT* dst, float4 v) {
  if (sizeof(OutT) == 4) *reinterpret_cast<float4*>(dst) = v;
  else *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

// One warp per token: the token's 3p^2 output elements are written as contiguous 16-byte (fp32) / 8-byte (bf16) chunks,
// lane after lane (fully coalesced), and gathered from the image as 16-pixel row segments.  All per-token index work
// (image table, grid, row / col) is done once per token; PT = compile-time patch size (16 / 32; 0 = any multiple of 4)
// turns the per-chunk divisions into shifts -- the first version of this kernel spent its time in 64-bit div / mod
// (issue-bound at 1.9 TB/s).  uint8 HWC input: a lane reads the 12 contiguous bytes of 4 pixels (3 x 32-bit loads) and
// writes one chunk per channel; the 256-entry normalisation table lives in shared memory (bit-exact by construction).
template <typename OutT, int PT>
__global__ void __launch_bounds__(256) patchify_kernel(const PatchifyArgs a) {
  pdl_wait();      // PDL (vtk_common.cuh): this kernel may have been launched before its predecessor finished
  pdl_trigger();
  __shared__ float lut[256];
  if (a.in_dtype == 1) {
    lut[threadIdx.x] = norm_u8(threadIdx.x);
    __syncthreads();
  }
  const int p = PT ? PT : a.patch, T = a.max_tokens;
  const int pp = p * p, P = 3 * pp, p4 = p >> 2, chunks = pp >> 2;   // chunks = 4-pixel groups per channel
  const int lane = threadIdx.x & 31;
  const unsigned ntok = (unsigned)a.B * (unsigned)T;
  const unsigned nwarp = gridDim.x * (blockDim.x >> 5);
  for (unsigned bt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); bt < ntok; bt += nwarp) {
    const int b = (int)(bt / (unsigned)T), t = (int)(bt - (unsigned)b * (unsigned)T);
    const long long off = a.img_table[3 * b + 0];
    const int H = (int)a.img_table[3 * b + 1], W = (int)a.img_table[3 * b + 2];
    const int gr = (H + p - 1) / p, gc = (W + p - 1) / p;
    const int n = gr * gc;
    const bool valid = t < n && n <= T;
    const int r = valid ? t / gc : 0, c = valid ? t - (t / gc) * gc : 0;
    OutT* dst = reinterpret_cast<OutT*>(a.patches) + (long long)bt * P;
    if (!valid) {
      for (int e4 = lane; e4 < 3 * chunks; e4 += 32) store4(dst + (e4 << 2), make_float4(0.f, 0.f, 0.f, 0.f));
    } else if (a.in_dtype == 0) {
      const float* img = reinterpret_cast<const float*>(a.images) + off;
      const bool al = ((reinterpret_cast<uintptr_t>(img) | ((uintptr_t)W << 2)) & 15) == 0;   // rows start 16-byte aligned
#pragma unroll 2
      for (int e4 = lane; e4 < 3 * chunks; e4 += 32) {
        const int ch = e4 / chunks, rem = e4 - ch * chunks;
        const int dy = rem / p4, dx = (rem - dy * p4) << 2;
        const int y = r * p + dy, x = c * p + dx;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);      // zero padding to the patch boundary (ops.py:235-238)
        if (y < H) {
          const float* src = img + ((long long)ch * H + y) * W + x;
          if (al && x + 3 < W) {
            v = __ldg(reinterpret_cast<const float4*>(src));
          } else {
            if (x < W) v.x = __ldg(src);
            if (x + 1 < W) v.y = __ldg(src + 1);
            if (x + 2 < W) v.z = __ldg(src + 2);
            if (x + 3 < W) v.w = __ldg(src + 3);
          }
        }
        store4(dst + (e4 << 2), v);
      }
    } else {
      const uint8_t* img = reinterpret_cast<const uint8_t*>(a.images) + off;
      const bool al = (reinterpret_cast<uintptr_t>(img) & 3) == 0;
      for (int q = lane; q < chunks; q += 32) {
        const int dy = q / p4, dx = (q - dy * p4) << 2;
        const int y = r * p + dy, x = c * p + dx;
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0, v2 = v0;
        if (y < H && x < W) {
          const uint8_t* src = img + ((long long)y * W + x) * 3;    // 12 contiguous bytes: r g b r g b r g b r g b
          uint32_t w0 = 0, w1 = 0, w2 = 0;
          if (al && x + 3 < W && (((long long)y * W + x) & 3) == 0) {
            const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
            w0 = __ldg(s32); w1 = __ldg(s32 + 1); w2 = __ldg(s32 + 2);
          } else {
            const int nb = min(4, W - x) * 3;
            for (int k = 0; k < nb; ++k) {
              const uint32_t u = src[k];
              if (k < 4) w0 |= u << (8 * k); else if (k < 8) w1 |= u << (8 * (k - 4)); else w2 |= u << (8 * (k - 8));
            }
          }
          const int np = min(4, W - x);   // pixels inside the image; the rest stay 0.0 (padding is applied AFTER normalize)
          v0.x = lut[w0 & 255]; v1.x = lut[(w0 >> 8) & 255]; v2.x = lut[(w0 >> 16) & 255];
          if (np > 1) { v0.y = lut[w0 >> 24]; v1.y = lut[w1 & 255]; v2.y = lut[(w1 >> 8) & 255]; }
          if (np > 2) { v0.z = lut[(w1 >> 16) & 255]; v1.z = lut[w1 >> 24]; v2.z = lut[w2 & 255]; }
          if (np > 3) { v0.w = lut[(w2 >> 8) & 255]; v1.w = lut[(w2 >> 16) & 255]; v2.w = lut[w2 >> 24]; }
        }
        store4(dst + (q << 2), v0);
        store4(dst + pp + (q << 2), v1);
        store4(dst + 2 * pp + (q << 2), v2);
      }
    }
    if (lane == 0) {  // index packing, one lane per token
      a.patch_mask[bt] = valid ? 1 : 0;
      a.row_idx[bt] = r;
      a.col_idx[bt] = c;
      a.time_idx[bt] = 0;
      if (t == 0) {
        a.meta[0 * a.B + b] = H;
        a.meta[1 * a.B + b] = W;
        a.meta[2 * a.B + b] = gr;
        a.meta[3 * a.B + b] = gc;
        if (n > T && a.status) atomicExch(a.status, 1);
      }
    }
  }
}

template <typename OutT>
static void patchify_dispatch(const PatchifyArgs& a, int blocks, cudaStream_t stream) {
  if (a.patch == 16) (void)launch_k(patchify_kernel<OutT, 16>, dim3(blocks), dim3(256), 0, stream, a);
  else if (a.patch == 32) (void)launch_k(patchify_kernel<OutT, 32>, dim3(blocks), dim3(256), 0, stream, a);
  else (void)launch_k(patchify_kernel<OutT, 0>, dim3(blocks), dim3(256), 0, stream, a);
}

int launch_patchify(const PatchifyArgs& a, cudaStream_t stream) {
  if (a.patch % 4 || a.patch <= 0) { set_error("patchify: patch size must be a positive multiple of 4 (got %d)", a.patch); return -2; }
  if (a.B <= 0 || a.max_tokens <= 0) return 0;
  const long long ntok = (long long)a.B * a.max_tokens;
  if (ntok >= (1ll << 31)) { set_error("patchify: B * max_tokens too large"); return -2; }
  long long blocks = (ntok + 7) / 8;                   // one warp per token, 8 warps per CTA
  const long long cap = (long long)num_sms() * 8;     // a multiple of the SM count (8 resident CTAs per SM), warp-stride beyond
  if (blocks > cap) blocks = cap;
  if (a.out_dtype == 0) patchify_dispatch<float>(a, (int)blocks, stream);
  else patchify_dispatch<bf16>(a, (int)blocks, stream);
  return check_cuda(cudaGetLastError(), "patchify launch");
}

// ---------------------------------------------------------------------------------------------
// unpatchify
// ---------------------------------------------------------------------------------------------
__global__ void cellmap_kernel(const UnpatchifyArgs a) {
  pdl_wait();      // PDL (vtk_common.cuh): this kernel may have been launched before its predecessor finished
  pdl_trigger();
  const long long total = (long long)a.B * a.N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (!a.patch_mask[i]) continue;
    const int b = (int)(i / a.N), t = (int)(i % a.N);
    const long long r = a.row_idx[i], c = a.col_idx[i];
    const long long cell = r * a.gx + c;   // flat_idx = row * max_x + col (ops.py:328)
    if (r < 0 || c < 0 || cell >= (long long)a.gy * a.gx) {
      if (a.status) atomicExch(a.status, 1);   // the reference's scatter raises an index error here
      continue;
    }
    atomicMax(&a.cell_map[(long long)b * a.gy * a.gx + cell], t);
  }
}

__device__ __forceinline__ float convert_px(float x, int fmt, bool bf16_math) {
  // vitok/pp/io.py:91-121 from "minus_one_to_one"; each eager op rounds to the tensor dtype
  if (fmt == 1) {  // 0_255: ((clamp(x,-1,1) + 1) / 2 * 255).round()
    float t = fminf(fmaxf(x, -1.f), 1.f);
    if (bf16_math) { t = bf16r(t + 1.f); t = bf16r(t / 2.f); t = bf16r(t * 255.f); }
    else { t = __fadd_rn(t, 1.f); t = __fdiv_rn(t, 2.f); t = __fmul_rn(t, 255.f); }
    return rintf(t);
  }
  if (fmt == 2) {  // zero_to_one: ((x + 1) / 2).clamp(0, 1)
    float t;
    if (bf16_math) { t = bf16r(x + 1.f); t = bf16r(t / 2.f); }
    else { t = __fdiv_rn(__fadd_rn(x, 1.f), 2.f); }
    return fminf(fmaxf(t, 0.f), 1.f);
  }
  return x;
}

// One warp per canvas cell (b, r, c): the token that owns the cell is looked up once, its 3p^2 elements are read as
// contiguous 8 / 16-byte chunks (coalesced) and written as p-pixel row segments of the three channel planes.  PT = compile-
// time patch size (16 / 32; 0 = any multiple of 4) so the per-chunk index math is shifts, not divisions.
template <typename T, int PT>
__global__ void __launch_bounds__(256) unpatchify_kernel(const UnpatchifyArgs a) {
  pdl_wait();      // PDL (vtk_common.cuh): this kernel may have been launched before its predecessor finished
  pdl_trigger();
  const int p = PT ? PT : a.patch, pp = p * p, P = 3 * pp, p4 = p >> 2, chunks = pp >> 2;
  const int Hc = a.gy * p, Wc = a.gx * p;
  const int cells_per_img = a.gy * a.gx;
  const unsigned ncell = (unsigned)a.B * (unsigned)cells_per_img;
  const bool bfm = sizeof(T) == 2;
  const int lane = threadIdx.x & 31;
  const unsigned nwarp = gridDim.x * (blockDim.x >> 5);
  for (unsigned bc = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); bc < ncell; bc += nwarp) {
    const int b = (int)(bc / (unsigned)cells_per_img), cell = (int)(bc - (unsigned)b * (unsigned)cells_per_img);
    const int r = cell / a.gx, c = cell - r * a.gx;
    // token 0 is re-scattered into cell 0 after the main scatter (ops.py:332-333)
    const int tok = (cell == 0) ? (a.patch_mask[(long long)b * a.N] ? 0 : -1) : a.cell_map[bc];
    const T* src = reinterpret_cast<const T*>(a.patches) + ((long long)b * a.N + (tok >= 0 ? tok : 0)) * P;
#pragma unroll 2
    for (int e4 = lane; e4 < 3 * chunks; e4 += 32) {
      const int ch = e4 / chunks, rem = e4 - ch * chunks;
      const int dy = rem / p4, dx = (rem - dy * p4) << 2;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (tok >= 0) {
        if (sizeof(T) == 4) {
          const float4 f = __ldg(reinterpret_cast<const float4*>(src + (e4 << 2)));
          v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
        } else {
          const uint2 u = __ldg(reinterpret_cast<const uint2*>(src + (e4 << 2)));
          v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
        }
      }
      const long long o = (((long long)b * 3 + ch) * Hc + r * p + dy) * Wc + c * p + dx;
      if (a.out_format == 1) {
        uchar4 u;
        u.x = (unsigned char)convert_px(v[0], 1, bfm); u.y = (unsigned char)convert_px(v[1], 1, bfm);
        u.z = (unsigned char)convert_px(v[2], 1, bfm); u.w = (unsigned char)convert_px(v[3], 1, bfm);
        *reinterpret_cast<uchar4*>(reinterpret_cast<uint8_t*>(a.out) + o) = u;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = convert_px(v[k], a.out_format, bfm);
        if (sizeof(T) == 4) {
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) + o) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
          *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(a.out) + o) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
        }
      }
    }
  }
}

template <typename T>
static void unpatchify_dispatch(const UnpatchifyArgs& a, int blocks, cudaStream_t stream) {
  if (a.patch == 16) (void)launch_k(unpatchify_kernel<T, 16>, dim3(blocks), dim3(256), 0, stream, a);
  else if (a.patch == 32) (void)launch_k(unpatchify_kernel<T, 32>, dim3(blocks), dim3(256), 0, stream, a);
  else (void)launch_k(unpatchify_kernel<T, 0>, dim3(blocks), dim3(256), 0, stream, a);
}

int launch_unpatchify(const UnpatchifyArgs& a, cudaStream_t stream) {
  if (a.patch % 4 || a.patch <= 0) { set_error("unpatchify: patch size must be a positive multiple of 4 (got %d)", a.patch); return -2; }
  if (a.gy <= 0 || a.gx <= 0) { set_error("unpatchify: empty canvas %dx%d", a.gy, a.gx); return -2; }
  if (a.B <= 0) return 0;
  const long long cells = (long long)a.B * a.gy * a.gx;
  if (check_cuda(cudaMemsetAsync(a.cell_map, 0xFF, cells * sizeof(int), stream), "cell_map memset")) return -1;
  const long long cap = (long long)num_sms() * 16;
  {
    long long blocks = ((long long)a.B * a.N + 255) / 256;
    if (blocks > cap) blocks = cap;
    (void)launch_k(cellmap_kernel, dim3((int)blocks), dim3(256), 0, stream, a);
  }
  if (cells >= (1ll << 31)) { set_error("unpatchify: canvas too large"); return -2; }
  long long blocks = (cells + 7) / 8;                    // one warp per cell, 8 warps per CTA
  const long long cap2 = (long long)num_sms() * 8;
  if (blocks > cap2) blocks = cap2;
  if (a.dtype == 0) unpatchify_dispatch<float>(a, (int)blocks, stream);
  else unpatchify_dispatch<bf16>(a, (int)blocks, stream);
  return check_cuda(cudaGetLastError(), "unpatchify launch");
}

// out2[0] = max(row)+1, out2[1] = max(col)+1 over valid tokens (ops.py:319-321); out2 must be zeroed.
__global__ void grid_extent_kernel(const uint8_t* __restrict__ mask, const int64_t* __restrict__ row,
                                   const int64_t* __restrict__ col, long long total, int* out2) {
  pdl_wait();      // PDL (vtk_common.cuh): this kernel may have been launched before its predecessor finished
  pdl_trigger();
  int my = 0, mx = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (mask[i]) {
      my = max(my, (int)row[i] + 1);
      mx = max(mx, (int)col[i] + 1);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    my = max(my, __shfl_xor_sync(0xffffffffu, my, o));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&out2[0], my);
    atomicMax(&out2[1], mx);
  }
}
int launch_grid_extent(const uint8_t* mask, const int64_t* row, const int64_t* col, int B, int N, int* out2,
                       cudaStream_t stream) {
  if (check_cuda(cudaMemsetAsync(out2, 0, 2 * sizeof(int), stream), "grid_extent memset")) return -1;
  const long long total = (long long)B * N;
  if (total <= 0) return 0;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 4;
  if (blocks > cap) blocks = cap;
  (void)launch_k(grid_extent_kernel, dim3((int)blocks), dim3(256), 0, stream, mask, row, col, total, out2);
  return check_cuda(cudaGetLastError(), "grid_extent launch");
}

}  // namespace vtk
