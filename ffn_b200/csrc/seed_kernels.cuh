// seed_kernels.cuh — PolicyPeaks on the device: once the flood fill runs at
// thousands of FoV steps/s the host seed policy (Sobel -> adaptive threshold -> Euclidean distance
// transform -> local maxima; ffn/inference/seed.py:142-199) dominates the end-to-end time.
// Every stage is an HBM-bound grid-stride or line-parallel kernel over the canvas:
//   sobel_mag      27-point stencil, reflect boundary           8 B/voxel  (read f32/u8, write f32)
//   gauss_pass x3  1-D correlation, radius 33, reflect          8 B/voxel per pass
//   edges_kernel   edges > threshold, masks                     9 B/voxel
//   edt_x / edt_line x2  exact squared EDT (two sweeps, then lower envelope of parabolas per line)
//   peaks_kernel   7x7x7 maximum of (distance, tie-break noise) keys, plateau-free arg-max test
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace ffn {
namespace seedk {

constexpr float kBig = 1e18f;

__device__ __forceinline__ int reflect(int i, int n) {   // scipy 'reflect': d c b a | a b c d | d c b a
  if (n == 1) return 0;
  const int period = 2 * n;
  i = i % period;
  if (i < 0) i += period;
  return i < n ? i : period - 1 - i;
}

__device__ __forceinline__ float load_image(const void* img, int is_u8, float mean, float stddev, size_t i) {
  if (is_u8) return __fdiv_rn(__fsub_rn((float)reinterpret_cast<const uint8_t*>(img)[i], mean), stddev);
  return reinterpret_cast<const float*>(img)[i];
}

// ndimage.generic_gradient_magnitude(image, ndimage.sobel): per axis a derivative [-1, 0, 1] along the
// axis and [1, 2, 1] smoothing along the other two (ascending axis order), each 1-D pass accumulated in
// float64 and stored as float32 exactly like ndimage.correlate1d; then float32 squares, sum and sqrt.
__device__ __forceinline__ float smooth3(float a, float b, float c) {
  return (float)((double)b * 2.0 + ((double)a + (double)c));
}

__global__ void sobel_mag(const void* img, int is_u8, float mean, float stddev, float* out, int sz, int sy, int sx) {
  const size_t n = (size_t)sz * sy * sx;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % sx), y = (int)((i / sx) % sy), z = (int)(i / ((size_t)sx * sy));
    float v[3][3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int zz = reflect(z + a - 1, sz);
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const int yy = reflect(y + b - 1, sy);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int xx = reflect(x + c - 1, sx);
          v[a][b][c] = load_image(img, is_u8, mean, stddev, ((size_t)zz * sy + yy) * sx + xx);
        }
      }
    }
    float t[3], g[3];
    // axis 0 (z): derivative along z, smooth along y, then along x
#pragma unroll
    for (int c = 0; c < 3; ++c) t[c] = smooth3(v[2][0][c] - v[0][0][c], v[2][1][c] - v[0][1][c], v[2][2][c] - v[0][2][c]);
    g[0] = smooth3(t[0], t[1], t[2]);
    // axis 1 (y): derivative along y, smooth along z, then along x
#pragma unroll
    for (int c = 0; c < 3; ++c) t[c] = smooth3(v[0][2][c] - v[0][0][c], v[1][2][c] - v[1][0][c], v[2][2][c] - v[2][0][c]);
    g[1] = smooth3(t[0], t[1], t[2]);
    // axis 2 (x): derivative along x, smooth along z, then along y
#pragma unroll
    for (int b = 0; b < 3; ++b) t[b] = smooth3(v[0][b][2] - v[0][b][0], v[1][b][2] - v[1][b][0], v[2][b][2] - v[2][b][0]);
    g[2] = smooth3(t[0], t[1], t[2]);
    float acc = __fmul_rn(g[0], g[0]);
    acc = __fadd_rn(acc, __fmul_rn(g[1], g[1]));
    acc = __fadd_rn(acc, __fmul_rn(g[2], g[2]));
    out[i] = sqrtf(acc);
  }
}

// One separable pass of ndimage.gaussian_filter (mode='reflect'): float64 accumulation over the symmetric
// pairs from the outermost tap inwards (ndimage.correlate1d's symmetric branch), float32 store.
// weights[0..radius]: weights[0] is the centre tap; normalised on the host.
__global__ void gauss_pass(const float* in, float* out, const double* weights, int radius, int axis, int sz, int sy,
                           int sx) {
  const size_t n = (size_t)sz * sy * sx;
  const int len = axis == 0 ? sz : (axis == 1 ? sy : sx);
  const size_t st = axis == 0 ? (size_t)sy * sx : (axis == 1 ? (size_t)sx : 1);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % sx), y = (int)((i / sx) % sy), z = (int)(i / ((size_t)sx * sy));
    const int pos = axis == 0 ? z : (axis == 1 ? y : x);
    const size_t base = i - (size_t)pos * st;
    double acc = (double)in[i] * weights[0];
    if (pos >= radius && pos + radius < len) {
      for (int k = radius; k >= 1; --k)
        acc += ((double)in[i - (size_t)k * st] + (double)in[i + (size_t)k * st]) * weights[k];
    } else {
      for (int k = radius; k >= 1; --k)
        acc += ((double)in[base + (size_t)reflect(pos - k, len) * st] + (double)in[base + (size_t)reflect(pos + k, len) * st]) *
               weights[k];
    }
    out[i] = (float)acc;
  }
}

// filt_edges = edges > thresh (| masks); the distance transform input is 0 at edges, "infinite" elsewhere.
__global__ void edges_kernel(const float* edges, const float* thresh, const uint8_t* mask, const uint8_t* seed_mask,
                             float* d, size_t n, unsigned long long* n_free) {
  unsigned long long local = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const bool edge = edges[i] > thresh[i] || (mask && mask[i]) || (seed_mask && seed_mask[i]);
    d[i] = edge ? 0.f : kBig;
    local += edge ? 0 : 1;
  }
  if (local) atomicAdd(n_free, local);
}

// First EDT pass along x: squared distance (in physical units) to the nearest zero of the same x-line.
__global__ void edt_x(float* d, int sz, int sy, int sx, float wx) {
  const size_t lines = (size_t)sz * sy;
  for (size_t l = (size_t)blockIdx.x * blockDim.x + threadIdx.x; l < lines; l += (size_t)gridDim.x * blockDim.x) {
    float* row = d + l * sx;
    float dist = kBig;
    for (int x = 0; x < sx; ++x) {
      dist = row[x] == 0.f ? 0.f : (dist >= kBig ? kBig : dist + wx);
      row[x] = dist;
    }
    dist = kBig;
    for (int x = sx - 1; x >= 0; --x) {
      dist = row[x] == 0.f ? 0.f : (dist >= kBig ? kBig : dist + wx);
      const float m = fminf(row[x], dist);
      row[x] = m >= kBig ? kBig : m * m;
    }
  }
}

// Later EDT passes (axis = 1: y, axis = 0: z): D(q) = min_p f(p) + (w (q - p))^2 via the lower envelope of
// parabolas (Felzenszwalb & Huttenlocher).  One thread per line; the envelope arrays live in global
// scratch laid out [k][line] so that neighbouring threads touch neighbouring addresses.
__global__ void edt_line(float* d, int axis, int sz, int sy, int sx, float w, int* vbuf, double* zbuf) {
  const int len = axis == 0 ? sz : sy;
  const size_t st = axis == 0 ? (size_t)sy * sx : (size_t)sx;
  const size_t lines = axis == 0 ? (size_t)sy * sx : (size_t)sz * sx;
  const double w2 = (double)w * (double)w;
  for (size_t l = (size_t)blockIdx.x * blockDim.x + threadIdx.x; l < lines; l += (size_t)gridDim.x * blockDim.x) {
    size_t base;
    if (axis == 0) {
      base = l;                                            // (y, x) fixed
    } else {
      const size_t z = l / sx, x = l % sx;
      base = z * (size_t)sy * sx + x;                      // (z, x) fixed
    }
    int k = -1;
    for (int q = 0; q < len; ++q) {
      const float fq = d[base + (size_t)q * st];
      if (fq >= kBig) continue;
      double s = -CUDART_INF;
      while (k >= 0) {
        const int vk = vbuf[(size_t)k * lines + l];
        const double fv = (double)d[base + (size_t)vk * st];   // still the INPUT value: outputs are written later
        s = (((double)fq + w2 * q * q) - (fv + w2 * vk * vk)) / (2.0 * w2 * (q - vk));
        if (s <= zbuf[(size_t)k * lines + l]) {
          --k;
          s = -CUDART_INF;
        } else {
          break;
        }
      }
      ++k;
      vbuf[(size_t)k * lines + l] = q;
      zbuf[(size_t)k * lines + l] = s;
    }
    if (k < 0) continue;   // no finite value on this line: stays "infinite"
    // The outputs overwrite the inputs, so cache the envelope's f(v) first.
    const int nk = k + 1;
    for (int j = 0; j < nk; ++j) {
      const int vj = vbuf[(size_t)j * lines + l];
      zbuf[(size_t)(len + j) * lines + l] = (double)d[base + (size_t)vj * st];
    }
    int j = 0;
    for (int q = 0; q < len; ++q) {
      while (j + 1 < nk && zbuf[(size_t)(j + 1) * lines + l] < (double)q) ++j;
      const int vj = vbuf[(size_t)j * lines + l];
      const double dq = (double)(q - vj);
      d[base + (size_t)q * st] = (float)(w2 * dq * dq + zbuf[(size_t)(len + j) * lines + l]);
    }
  }
}

// dt = sqrt(squared distance); excluded voxels and "infinite" distances become -1 (seed.py:187-188).
__global__ void finish_dt(float* d, const int* seg, const uint8_t* mask, const uint8_t* seed_mask, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const bool excl = (seg && seg[i] > 0) || (mask && mask[i]) || (seed_mask && seed_mask[i]);
    const float v = d[i];
    d[i] = (excl || v >= kBig) ? -1.f : sqrtf(v);
  }
}

// peak_local_max(dt + noise * 1e-4, min_distance = 3, threshold_abs = 0): a voxel is a seed iff its key
// (dt, noise) is the maximum of its 7x7x7 neighbourhood (edge-clamped) and dt + 1e-4 * noise > 0.
// Comparing (dt, noise) lexicographically equals comparing the float64 sums because distinct dt
// values differ by far more than 1e-4.
__global__ void peaks_kernel(const float* dt, const double* noise, int sz, int sy, int sx, int radius, int* coords,
                             unsigned long long cap, unsigned long long* count) {
  const size_t n = (size_t)sz * sy * sx;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = dt[i];
    const double nv = noise ? noise[i] : 0.0;
    if (!((double)v + nv * 1e-4 > 0.0)) continue;
    const int x = (int)(i % sx), y = (int)((i / sx) % sy), z = (int)(i / ((size_t)sx * sy));
    bool is_max = true;
    for (int dz = -radius; dz <= radius && is_max; ++dz) {
      const int zz = min(max(z + dz, 0), sz - 1);
      for (int dy = -radius; dy <= radius && is_max; ++dy) {
        const int yy = min(max(y + dy, 0), sy - 1);
        const size_t rb = ((size_t)zz * sy + yy) * sx;
        for (int dx = -radius; dx <= radius; ++dx) {
          const int xx = min(max(x + dx, 0), sx - 1);
          const size_t j = rb + xx;
          const float u = dt[j];
          if (u > v || (u == v && noise && noise[j] > nv)) {
            is_max = false;
            break;
          }
        }
      }
    }
    if (!is_max) continue;
    const unsigned long long slot = atomicAdd(count, 1ull);
    if (slot < cap) {
      coords[3 * slot + 0] = z;
      coords[3 * slot + 1] = y;
      coords[3 * slot + 2] = x;
    }
  }
}

}  // namespace seedk
}  // namespace ffn
