// Library plumbing: error text, launch counter, device query, host-side tables.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "sdof_common.cuh"

namespace sdof {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SDOF_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

size_t l2_persist_bytes() {
  static long long cached[64];
  static bool done[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (!done[dev]) {
    long long v = 0;
    const char* e = getenv("SDOF_L2_PERSIST");
    if (e && e[0] == '1') {   // opt-in (experiment): see profiles/README.md
      int max_persist = 0, max_window = 0;
      if (cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev) == cudaSuccess &&
          cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev) == cudaSuccess && max_persist > 0 &&
          cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist) == cudaSuccess)
        v = max_persist < max_window ? max_persist : max_window;
    }
    (void)cudaGetLastError();
    cached[dev] = v;
    done[dev] = true;
  }
  return (size_t)cached[dev];
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ---------------------------------------------------------------------------
// OpenCV's INTER_CUBIC remap weight tables (imgproc initInterTab2D, fixed point):
// 32x32 sub-pixel positions, 4x4 taps, a = -0.75 cubic evaluated in fp32
// (c3 = 1 - c0 - c1 - c2), outer product in fp32, scaled by 2^15 and rounded
// half-to-even; if the 16 integers do not sum to 2^15 the difference goes to
// the largest (sum too small) or smallest (sum too large) of the four taps
// ky,kx in {2,3}.  Pinned against cv2.remap by tests/test_capi_cpu.py.
static int16_t g_tab_i16[1024 * 16];
static float g_tab_f32[1024 * 16];
static std::once_flag g_tab_once;

static void cubic_1d(float x, volatile float* c) {
  const float A = -0.75f;
  volatile float xp1 = x + 1.0f;
  volatile float t;
  t = A * xp1;
  t = t - 5.0f * A;
  t = t * xp1;
  t = t + 8.0f * A;
  t = t * xp1;
  c[0] = t - 4.0f * A;
  t = (A + 2.0f) * x;
  t = t - (A + 3.0f);
  t = t * x;
  t = t * x;
  c[1] = t + 1.0f;
  volatile float omx = 1.0f - x;
  t = (A + 2.0f) * omx;
  t = t - (A + 3.0f);
  t = t * omx;
  t = t * omx;
  c[2] = t + 1.0f;
  t = 1.0f - c[0];
  t = t - c[1];
  c[3] = t - c[2];
}

static void build_tables() {
  float t1[32][4];
  for (int i = 0; i < 32; ++i) {
    volatile float c[4];
    cubic_1d((float)i * (1.0f / 32.0f), c);
    for (int k = 0; k < 4; ++k) t1[i][k] = c[k];
  }
  for (int fy = 0; fy < 32; ++fy)
    for (int fx = 0; fx < 32; ++fx) {
      int q[4][4];
      int isum = 0;
      float* tf = g_tab_f32 + (fy * 32 + fx) * 16;
      for (int k1 = 0; k1 < 4; ++k1)
        for (int k2 = 0; k2 < 4; ++k2) {
          volatile float v = t1[fy][k1] * t1[fx][k2];
          tf[k1 * 4 + k2] = v;
          volatile float s = v * 32768.0f;
          long r = lrintf(s);  // round half to even (default rounding mode) = cvRound
          if (r > 32767) r = 32767;
          if (r < -32768) r = -32768;
          q[k1][k2] = (int)r;
          isum += (int)r;
        }
      if (isum != 32768) {
        int diff = isum - 32768;
        int Mk1 = 2, Mk2 = 2, mk1 = 2, mk2 = 2;
        for (int k1 = 2; k1 < 4; ++k1)
          for (int k2 = 2; k2 < 4; ++k2) {
            if (q[k1][k2] < q[mk1][mk2]) {
              mk1 = k1;
              mk2 = k2;
            } else if (q[k1][k2] > q[Mk1][Mk2]) {
              Mk1 = k1;
              Mk2 = k2;
            }
          }
        if (diff < 0)
          q[Mk1][Mk2] -= diff;
        else
          q[mk1][mk2] -= diff;
      }
      int16_t* ti = g_tab_i16 + (fy * 32 + fx) * 16;
      for (int k1 = 0; k1 < 4; ++k1)
        for (int k2 = 0; k2 < 4; ++k2) ti[k1 * 4 + k2] = (int16_t)q[k1][k2];
    }
}

const int16_t* cubic_table_i16_host() {
  std::call_once(g_tab_once, build_tables);
  return g_tab_i16;
}
const float* cubic_table_f32_host() {
  std::call_once(g_tab_once, build_tables);
  return g_tab_f32;
}

// cv2.getStructuringElement(MORPH_ELLIPSE,(k,k)): row i spans c-dx..c+dx with
// dx = round(c*sqrt((r^2-dy^2)/r^2)), r = c = k/2.
void ellipse_half_widths_host(int ksize, int32_t* out) {
  int r = ksize / 2, c = ksize / 2;
  double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
  for (int i = 0; i < ksize; ++i) {
    int dy = i - r;
    int dx = (int)lrint((double)c * sqrt(((double)r * r - (double)dy * dy) * inv_r2));
    out[i] = dx;
  }
}

}  // namespace sdof

extern "C" {

int sdof_abi_version(void) { return SDOF_ABI_VERSION; }
const char* sdof_last_error(void) { return sdof::g_err; }
int64_t sdof_launch_count(void) { return sdof::g_launches.load(std::memory_order_relaxed); }

int sdof_cubic_table_i16(int16_t* out) {
  SDOF_REQUIRE(out != nullptr, "sdof_cubic_table_i16: out is NULL");
  memcpy(out, sdof::cubic_table_i16_host(), sizeof(int16_t) * 1024 * 16);
  return SDOF_OK;
}

int sdof_ellipse_half_widths(int ksize, int32_t* out) {
  SDOF_REQUIRE(out != nullptr, "sdof_ellipse_half_widths: out is NULL");
  SDOF_REQUIRE(ksize >= 1 && (ksize & 1) && ksize <= 31, "sdof_ellipse_half_widths: ksize must be odd in [1,31], got %d",
               ksize);
  sdof::ellipse_half_widths_host(ksize, out);
  return SDOF_OK;
}

int sdof_corr_pyramid_layout(int64_t rows, int h2, int w2, int levels, sdof_pyramid_layout* out) {
  SDOF_REQUIRE(out != nullptr, "sdof_corr_pyramid_layout: out is NULL");
  SDOF_REQUIRE(levels >= 1 && levels <= SDOF_MAX_LEVELS, "sdof_corr_pyramid_layout: levels must be in [1,%d], got %d",
               SDOF_MAX_LEVELS, levels);
  SDOF_REQUIRE(rows >= 0 && h2 >= 1 && w2 >= 1, "sdof_corr_pyramid_layout: bad sizes rows=%lld h2=%d w2=%d",
               (long long)rows, h2, w2);
  memset(out, 0, sizeof(*out));
  out->levels = levels;
  int64_t off = 0;
  for (int l = 0; l < levels; ++l) {
    int h = h2 >> l, w = w2 >> l;
    int wp = (w + 3) & ~3;
    out->h[l] = h;
    out->w[l] = w;
    out->wp[l] = wp;
    out->pitch[l] = (int64_t)h * wp;  // (padding the pitch to skew L2 slices was measured: no gain, slightly slower)
    out->offset[l] = off;
    off += rows * out->pitch[l];
    off = (off + 31) & ~(int64_t)31;  // keep every level 128-byte aligned
  }
  out->total_floats = off;
  return SDOF_OK;
}

}  // extern "C"
