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

__device__ __forceinline__ float load_pixel(const void* base, long long off, int in_dtype, int H, int W, int ch, int y,
                                            int x) {
  if (y >= H || x >= W) return 0.f;  // zero padding to the patch boundary (ops.py:235-238)
  if (in_dtype == 0) return reinterpret_cast<const float*>(base)[off + ((long long)ch * H + y) * W + x];
  const uint8_t u = reinterpret_cast<const uint8_t*>(base)[off + ((long long)y * W + x) * 3 + ch];
  // ToTensor: u/255 (fp32 division); Normalize(0.5, 0.5): (t - 0.5) / 0.5
  const float t = __fdiv_rn((float)u, 255.0f);
  return __fdiv_rn(__fsub_rn(t, 0.5f), 0.5f);
}

template <typename OutT>
__global__ void __launch_bounds__(256) patchify_kernel(const PatchifyArgs a) {
  const int p = a.patch, T = a.max_tokens;
  const int pp = p * p, P = 3 * pp, P4 = P >> 2, p4 = p >> 2;
  const long long total = (long long)a.B * T * P4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e4 = (int)(i % P4);
    const long long bt = i / P4;
    const int t = (int)(bt % T);
    const int b = (int)(bt / T);
    const long long off = a.img_table[3 * b + 0];
    const int H = (int)a.img_table[3 * b + 1], W = (int)a.img_table[3 * b + 2];
    const int gr = (H + p - 1) / p, gc = (W + p - 1) / p;
    const int n = gr * gc;
    const bool valid = t < n && n <= T;
    const int r = valid ? t / gc : 0, c = valid ? t % gc : 0;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
      const int ch = e4 / (p * p4);
      const int rem = e4 - ch * (p * p4);
      const int dy = rem / p4, dx = (rem - dy * p4) << 2;
      const int y = r * p + dy, x = c * p + dx;
      bool fast = false;
      if (a.in_dtype == 0 && y < H && x + 3 < W) {
        const long long idx = off + ((long long)ch * H + y) * W + x;
        const float* src = reinterpret_cast<const float*>(a.images) + idx;
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
          v = *reinterpret_cast<const float4*>(src);
          fast = true;
        }
      }
      if (!fast) {
        v.x = load_pixel(a.images, off, a.in_dtype, H, W, ch, y, x);
        v.y = load_pixel(a.images, off, a.in_dtype, H, W, ch, y, x + 1);
        v.z = load_pixel(a.images, off, a.in_dtype, H, W, ch, y, x + 2);
        v.w = load_pixel(a.images, off, a.in_dtype, H, W, ch, y, x + 3);
      }
    }
    OutT* dst = reinterpret_cast<OutT*>(a.patches) + bt * P + ((long long)e4 << 2);
    if (sizeof(OutT) == 4) {
      *reinterpret_cast<float4*>(dst) = v;
    } else {
      *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
    if (e4 == 0) {  // index packing, one thread per token
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

int launch_patchify(const PatchifyArgs& a, cudaStream_t stream) {
  if (a.patch % 4 || a.patch <= 0) { set_error("patchify: patch size must be a positive multiple of 4 (got %d)", a.patch); return -2; }
  if (a.B <= 0 || a.max_tokens <= 0) return 0;
  const long long total = (long long)a.B * a.max_tokens * (3 * a.patch * a.patch / 4);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;   // multiple of the SM count, grid-stride beyond
  if (blocks > cap) blocks = cap;
  if (a.out_dtype == 0) patchify_kernel<float><<<(int)blocks, 256, 0, stream>>>(a);
  else patchify_kernel<bf16><<<(int)blocks, 256, 0, stream>>>(a);
  return check_cuda(cudaGetLastError(), "patchify launch");
}

// ---------------------------------------------------------------------------------------------
// unpatchify
// ---------------------------------------------------------------------------------------------
__global__ void cellmap_kernel(const UnpatchifyArgs a) {
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

template <typename T>
__global__ void __launch_bounds__(256) unpatchify_kernel(const UnpatchifyArgs a) {
  const int p = a.patch, pp = p * p, P = 3 * pp;
  const int Hc = a.gy * p, Wc = a.gx * p, W4 = Wc >> 2;
  const long long total = (long long)a.B * 3 * Hc * W4;
  const bool bfm = sizeof(T) == 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W4) << 2;
    long long q = i / W4;
    const int y = (int)(q % Hc); q /= Hc;
    const int ch = (int)(q % 3);
    const int b = (int)(q / 3);
    const int r = y / p, dy = y - r * p, c = x / p, dx = x - c * p;
    const int cell = r * a.gx + c;
    // token 0 is re-scattered into cell 0 after the main scatter (ops.py:332-333)
    int tok = (cell == 0) ? (a.patch_mask[(long long)b * a.N] ? 0 : -1) : a.cell_map[(long long)b * a.gy * a.gx + cell];
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (tok >= 0) {
      const T* src = reinterpret_cast<const T*>(a.patches) + ((long long)b * a.N + tok) * P + ch * pp + dy * p + dx;
      if (sizeof(T) == 4) {
        const float4 f = *reinterpret_cast<const float4*>(src);
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
      } else {
        const uint2 u = *reinterpret_cast<const uint2*>(src);
        v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
      }
    }
    const long long o = (((long long)b * 3 + ch) * Hc + y) * Wc + x;
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
    cellmap_kernel<<<(int)blocks, 256, 0, stream>>>(a);
  }
  const long long total = (long long)a.B * 3 * a.gy * a.patch * (a.gx * a.patch / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > cap) blocks = cap;
  if (a.dtype == 0) unpatchify_kernel<float><<<(int)blocks, 256, 0, stream>>>(a);
  else unpatchify_kernel<bf16><<<(int)blocks, 256, 0, stream>>>(a);
  return check_cuda(cudaGetLastError(), "unpatchify launch");
}

// out2[0] = max(row)+1, out2[1] = max(col)+1 over valid tokens (ops.py:319-321); out2 must be zeroed.
__global__ void grid_extent_kernel(const uint8_t* __restrict__ mask, const int64_t* __restrict__ row,
                                   const int64_t* __restrict__ col, long long total, int* out2) {
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
  grid_extent_kernel<<<(int)blocks, 256, 0, stream>>>(mask, row, col, total, out2);
  return check_cuda(cudaGetLastError(), "grid_extent launch");
}

}  // namespace vtk
