// The step BEFORE the hot path: the key-frame detector of frame_generator
// (ofgen_pixel_inpaint.py:127-176, 300-312; same code in ofgen_keyframe_inpaint.py).  Per frame the reference runs
//
//   lum   = cv2.split(cv2.cvtColor(frame, cv2.COLOR_BGR2HSV))[2]          (V of 8-bit HSV = max(B,G,R))
//   low, high = int(max(0, (1 - 1/3) * median(lum))), int(min(255, (1 + 1/3) * median(lum)))
//   edges = cv2.dilate(cv2.Canny(lum, low, high), ones(k, k))
//   delta = sum(|edges - key_edges|) / (H*W)                               (mean_pixel_distance)
//
// on the CPU.  All of it is integer work on H*W bytes; here it stays on the device, bit-exact to OpenCV:
//   lum_hist_kernel      : V channel + 256-bin histogram (shared-memory atomics)
//   canny_thresholds_kernel : numpy's median from the histogram (mean of the two middle values for even counts)
//                          and the two thresholds in the reference's double arithmetic
//   canny_nms_kernel     : 3x3 Sobel (replicated border), L1 magnitude, non-maximum suppression with OpenCV's
//                          fixed-point tangent test (TG22 = 13573, shift 15); state 2 = strong, 0 = candidate, 1 = no edge
//   canny_hysteresis_kernel : candidates 8-connected to a strong pixel become strong; a CTA iterates its 32x32 tile (Jacobi sweeps)
//                          (+1 halo) to a local fixed point in shared memory; the host relaunches until no tile changes
//   edges_dilate_kernel  : 255 where strong, then the k x k rectangular dilation
//   abs_diff_sum_kernel  : integer sum of |a - b| (mean_pixel_distance numerator)
#include "sdof_common.cuh"

namespace sdof {

__global__ void __launch_bounds__(256) lum_hist_kernel(const unsigned char* __restrict__ bgr, int64_t npix, unsigned char* __restrict__ lum,
                                                       unsigned* __restrict__ hist) {
  __shared__ unsigned sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    const unsigned char* q = bgr + p * 3;
    const unsigned v = max(max((unsigned)q[0], (unsigned)q[1]), (unsigned)q[2]);
    lum[p] = (unsigned char)v;
    atomicAdd(&sh[v], 1u);
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

// thr[0] = low, thr[1] = high (ints).  np.median: sorted[(n-1)/2] and sorted[n/2] averaged.
__global__ void canny_thresholds_kernel(const unsigned* __restrict__ hist, int64_t npix, int* __restrict__ thr) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int64_t ia = (npix - 1) / 2, ib = npix / 2;
  int va = -1, vb = -1;
  int64_t cum = 0;
  for (int v = 0; v < 256; ++v) {
    cum += hist[v];
    if (va < 0 && cum > ia) va = v;
    if (vb < 0 && cum > ib) vb = v;
  }
  const double median = ((double)va + (double)vb) / 2.0;
  const double sigma = 1.0 / 3.0;
  const double lo = (1.0 - sigma) * median, hi = (1.0 + sigma) * median;
  thr[0] = (int)(lo > 0.0 ? lo : 0.0);
  thr[1] = (int)(hi < 255.0 ? hi : 255.0);
}

constexpr int kCnT = 32;  // tile edge

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// state map: 2 strong, 0 candidate, 1 suppressed (OpenCV's encoding).  lum [H,W], one image.
__global__ void __launch_bounds__(256) canny_nms_kernel(const unsigned char* __restrict__ lum, int H, int W, const int* __restrict__ thr,
                                                        int low_fixed, int high_fixed, unsigned char* __restrict__ state) {
  __shared__ unsigned char sl[kCnT + 4][kCnT + 4];  // luma, halo 2 (replicated border)
  __shared__ short sdx[kCnT + 2][kCnT + 2], sdy[kCnT + 2][kCnT + 2];
  __shared__ int smag[kCnT + 2][kCnT + 2];           // magnitude, halo 1 (0 outside the image)
  const int tx0 = blockIdx.x * kCnT, ty0 = blockIdx.y * kCnT;
  int low = thr ? thr[0] : low_fixed, high = thr ? thr[1] : high_fixed;
  if (low > high) { const int t = low; low = high; high = t; }
  for (int i = threadIdx.x; i < (kCnT + 4) * (kCnT + 4); i += blockDim.x) {
    const int ly = i / (kCnT + 4), lx = i - ly * (kCnT + 4);
    sl[ly][lx] = lum[(int64_t)clampi(ty0 + ly - 2, 0, H - 1) * W + clampi(tx0 + lx - 2, 0, W - 1)];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (kCnT + 2) * (kCnT + 2); i += blockDim.x) {
    const int ly = i / (kCnT + 2), lx = i - ly * (kCnT + 2);
    const int gy = ty0 + ly - 1, gx = tx0 + lx - 1;
    int dx = 0, dy = 0, m = 0;
    if ((unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W) {
      // sl index of pixel (gy, gx) is (ly + 1, lx + 1); its neighbours were clamp-loaded, but the clamp of the HALO
      // load is relative to the image, which is exactly BORDER_REPLICATE for in-image centres
      const int a = sl[ly][lx], b = sl[ly][lx + 1], c = sl[ly][lx + 2];
      const int d = sl[ly + 1][lx], f = sl[ly + 1][lx + 2];
      const int g = sl[ly + 2][lx], h = sl[ly + 2][lx + 1], k = sl[ly + 2][lx + 2];
      dx = (c + 2 * f + k) - (a + 2 * d + g);
      dy = (g + 2 * h + k) - (a + 2 * b + c);
      m = abs(dx) + abs(dy);
    }
    sdx[ly][lx] = (short)dx;
    sdy[ly][lx] = (short)dy;
    smag[ly][lx] = m;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kCnT * kCnT; i += blockDim.x) {
    const int ly = i / kCnT + 1, lx = i % kCnT + 1;
    const int gy = ty0 + ly - 1, gx = tx0 + lx - 1;
    if (gy >= H || gx >= W) continue;
    const int m = smag[ly][lx];
    unsigned char st = 1;
    if (m > low) {
      const int xs = sdx[ly][lx], ys = sdy[ly][lx];
      const int x = abs(xs), y = abs(ys) << 15;
      const int tg22x = x * 13573;
      bool keep;
      if (y < tg22x) {
        keep = m > smag[ly][lx - 1] && m >= smag[ly][lx + 1];
      } else {
        const int tg67x = tg22x + (x << 16);
        if (y > tg67x) {
          keep = m > smag[ly - 1][lx] && m >= smag[ly + 1][lx];
        } else {
          const int s = (xs ^ ys) < 0 ? -1 : 1;
          keep = m > smag[ly - 1][lx - s] && m > smag[ly + 1][lx + s];
        }
      }
      if (keep) st = m > high ? 2 : 0;
    }
    state[(int64_t)gy * W + gx] = st;
  }
}

__global__ void __launch_bounds__(256) canny_hysteresis_kernel(unsigned char* __restrict__ state, int H, int W, int* __restrict__ changed) {
  __shared__ unsigned char s[kCnT + 2][kCnT + 2];
  const int tx0 = blockIdx.x * kCnT, ty0 = blockIdx.y * kCnT;
  for (int i = threadIdx.x; i < (kCnT + 2) * (kCnT + 2); i += blockDim.x) {
    const int ly = i / (kCnT + 2), lx = i - ly * (kCnT + 2);
    const int gy = ty0 + ly - 1, gx = tx0 + lx - 1;
    s[ly][lx] = ((unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W) ? state[(int64_t)gy * W + gx] : 1;
  }
  __syncthreads();
  bool any = false;
  for (;;) {
    // Jacobi sweep: all reads of this sweep happen before any write (no shared-memory race), 4 pixels per thread
    bool upd[kCnT * kCnT / 256];
#pragma unroll
    for (int k = 0; k < kCnT * kCnT / 256; ++k) {
      const int i = threadIdx.x + k * 256;
      const int ly = i / kCnT + 1, lx = i % kCnT + 1;
      upd[k] = s[ly][lx] == 0 &&
               (s[ly - 1][lx - 1] == 2 || s[ly - 1][lx] == 2 || s[ly - 1][lx + 1] == 2 || s[ly][lx - 1] == 2 ||
                s[ly][lx + 1] == 2 || s[ly + 1][lx - 1] == 2 || s[ly + 1][lx] == 2 || s[ly + 1][lx + 1] == 2);
    }
    __syncthreads();
    int ch = 0;
#pragma unroll
    for (int k = 0; k < kCnT * kCnT / 256; ++k)
      if (upd[k]) {
        const int i = threadIdx.x + k * 256;
        s[i / kCnT + 1][i % kCnT + 1] = 2;
        ch = 1;
      }
    if (!__syncthreads_or(ch)) break;
    any = true;
  }
  if (!any) return;
  for (int i = threadIdx.x; i < kCnT * kCnT; i += blockDim.x) {
    const int ly = i / kCnT + 1, lx = i % kCnT + 1;
    const int gy = ty0 + ly - 1, gx = tx0 + lx - 1;
    if (gy < H && gx < W && s[ly][lx] == 2) state[(int64_t)gy * W + gx] = 2;
  }
  if (threadIdx.x == 0) *changed = 1;
}

// edges = 255 where state == 2; out = k x k rectangular dilation (anchor centre; outside the image never wins).
// k == 1 gives the Canny edge map itself.
__global__ void __launch_bounds__(256) edges_dilate_kernel(const unsigned char* __restrict__ state, int H, int W, int k,
                                                           unsigned char* __restrict__ out) {
  const int r = k >> 1;
  const int64_t total = (int64_t)H * W;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
    const int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
    unsigned char v = 0;
    for (int dy = -r; dy <= r && !v; ++dy) {
      const int yy = y + dy;
      if ((unsigned)yy >= (unsigned)H) continue;
      for (int dx = -r; dx <= r; ++dx) {
        const int xx = x + dx;
        if ((unsigned)xx < (unsigned)W && state[(int64_t)yy * W + xx] == 2) {
          v = 255;
          break;
        }
      }
    }
    out[p] = v;
  }
}

__global__ void __launch_bounds__(256) abs_diff_sum_kernel(const unsigned char* __restrict__ a, const unsigned char* __restrict__ b,
                                                           int64_t n, unsigned long long* __restrict__ sum) {
  unsigned long long acc = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc += (unsigned)abs((int)a[i] - (int)b[i]);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(sum, acc);
}

}  // namespace sdof

extern "C" {

int64_t sdof_detect_edges_workspace_bytes(int H, int W) {
  if (H < 1 || W < 1) return -1;
  return 2 * (((int64_t)H * W + 255) & ~255LL) + 256 * 4 + 64;  // lum, state, histogram, thresholds + flag
}

int sdof_detect_edges(const uint8_t* frame_bgr, int H, int W, int dilate_k, int low, int high, uint8_t* edges, void* workspace,
                      int64_t workspace_bytes, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(frame_bgr && edges && workspace, "sdof_detect_edges: NULL pointer");
  SDOF_REQUIRE(H >= 1 && W >= 1 && H <= 32767 && W <= 32767, "sdof_detect_edges: bad size %dx%d", H, W);
  SDOF_REQUIRE(dilate_k >= 1 && (dilate_k & 1) && dilate_k <= 63, "sdof_detect_edges: dilate_k must be odd in [1,63], got %d", dilate_k);
  SDOF_REQUIRE(workspace_bytes >= sdof_detect_edges_workspace_bytes(H, W), "sdof_detect_edges: workspace too small");
  SDOF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "sdof_detect_edges: workspace must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  const int64_t npix = (int64_t)H * W;
  const int64_t plane = (npix + 255) & ~255LL;
  unsigned char* lum = static_cast<unsigned char*>(workspace);
  unsigned char* state = lum + plane;
  unsigned* hist = reinterpret_cast<unsigned*>(state + plane);
  int* thr = reinterpret_cast<int*>(hist + 256);
  int* changed = thr + 2;
  SDOF_CUDA(cudaMemsetAsync(hist, 0, 256 * 4 + 64, st));
  lum_hist_kernel<<<grid_for(npix, 256, 4), 256, 0, st>>>(frame_bgr, npix, lum, hist);
  SDOF_LAUNCH_CHECK("lum_hist_kernel");
  const bool auto_thr = low < 0 || high < 0;  // the reference's median rule
  if (auto_thr) {
    canny_thresholds_kernel<<<1, 32, 0, st>>>(hist, npix, thr);
    SDOF_LAUNCH_CHECK("canny_thresholds_kernel");
  }
  dim3 grid(ceil_div(W, kCnT), ceil_div(H, kCnT));
  canny_nms_kernel<<<grid, 256, 0, st>>>(lum, H, W, auto_thr ? thr : nullptr, low, high, state);
  SDOF_LAUNCH_CHECK("canny_nms_kernel");
  // hysteresis to the global fixed point: tiles converge locally, the host relaunches while any tile changed
  for (int it = 0; it < 4096; ++it) {
    SDOF_CUDA(cudaMemsetAsync(changed, 0, 4, st));
    canny_hysteresis_kernel<<<grid, 256, 0, st>>>(state, H, W, changed);
    SDOF_LAUNCH_CHECK("canny_hysteresis_kernel");
    int h_changed = 0;
    SDOF_CUDA(cudaMemcpyAsync(&h_changed, changed, 4, cudaMemcpyDeviceToHost, st));
    SDOF_CUDA(cudaStreamSynchronize(st));
    if (!h_changed) break;
  }
  edges_dilate_kernel<<<grid_for(npix, 256, 8), 256, 0, st>>>(state, H, W, dilate_k, edges);
  SDOF_LAUNCH_CHECK("edges_dilate_kernel");
  return SDOF_OK;
}

int sdof_abs_diff_sum_u8(const uint8_t* a, const uint8_t* b, int64_t n, unsigned long long* sum, sdof_stream_t stream) {
  using namespace sdof;
  SDOF_REQUIRE(a && b && sum, "sdof_abs_diff_sum_u8: NULL pointer");
  SDOF_REQUIRE(n >= 0, "sdof_abs_diff_sum_u8: bad size");
  SDOF_CUDA(cudaMemsetAsync(sum, 0, 8, as_stream(stream)));
  if (n == 0) return SDOF_OK;
  abs_diff_sum_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(a, b, n, sum);
  SDOF_LAUNCH_CHECK("abs_diff_sum_kernel");
  return SDOF_OK;
}

}  // extern "C"
