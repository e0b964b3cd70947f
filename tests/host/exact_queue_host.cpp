// Host build of the lane algorithm of marex_b200/csrc/exact_queue.cuh (one-lane environment) for
// tests/test_exact_queue_host.py.  Test infrastructure: compiled by the test with g++, never loaded by the package.
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

// event log: [gridpoint][day of year][kind]: 0 raise, 1 lowering, 2 bracketing rebuild, 3 bracketing pass, 4 selection pass,
// 5 queue samples read by the selection passes
static short* g_events = nullptr;
static long long g_lane = 0;
static int g_day = 0;
static inline void xq_event(int kind, int count = 1) {
  if (g_events) g_events[(g_lane * 366 + g_day) * 6 + kind] += (short)count;
}
#define XQ_EVENT(kind) xq_event(kind)
#define XQ_EVENT_N(kind, count) xq_event(kind, count)
#include "exact_queue.cuh"

struct HostEnv {
  const float* col;
  int64_t pitch;
  const int32_t* doy_ptr;
  const int32_t* doy_rows;
  int w;
  std::vector<float> q;
  std::vector<uint8_t> c;
  HostEnv(const float* col_, int64_t pitch_, const int32_t* p_, const int32_t* r_, int w_, int Q, float qf_)
      : col(col_), pitch(pitch_), doy_ptr(p_), doy_rows(r_), w(w_), q(Q, -1234.f), c(3 * w_, 0), qf(qf_) {}
  static float inf() { return std::numeric_limits<float>::infinity(); }
  static float nan() { return std::numeric_limits<float>::quiet_NaN(); }
  static bool finite(float v) { return std::fabs(v) < inf(); }
  static float fmin(float a, float b) { return std::fmin(a, b); }
  static float fmax(float a, float b) { return std::fmax(a, b); }
  static float level(float lob, float top, int k) { return k == 8 ? top : (k == 0 ? lob : lob + (top - lob) * ((float)k * 0.125f)); }
  float qf;
  void rank(int n, int& r0, int& r1, float& g) const {  // as f32_rank (thresholds.cu)
    const float vi = (float)(n - 1) * qf;
    if (vi >= (float)(n - 1)) { r0 = r1 = n - 1; g = 0.f; return; }
    if (vi < 0.f) { r0 = r1 = 0; g = 0.f; return; }
    const float lo = std::floor(vi);
    r0 = (int)lo;
    r1 = r0 + 1;
    g = vi - lo;
  }
  float finish(float a, float b, float g) const {  // as f32_lerp (thresholds.cu)
    const float diff = b - a;
    volatile float t = diff * g;
    float r = a + t;
    if (g >= 0.5f) { volatile float u = diff * (1.f - g); r = b - u; }
    return r;
  }
  float load(int j) const { return col[(int64_t)doy_rows[j] * pitch]; }
  int doy_begin(int dd) const { return doy_ptr[dd]; }
  float& que(int pos) { return q[pos]; }
  uint8_t& cnt(int s) { return c[s]; }
  uint8_t& eqc(int s) { return c[w + s]; }
  uint8_t& nvc(int s) { return c[2 * w + s]; }
  int wmax(int v) { return v; }
  bool any(bool p) { return p; }
  bool all(bool p) { return p; }
};

template <int Q>
static int run_all(const float* anom, int64_t N, int64_t pitch, const int32_t* doy_ptr, const int32_t* doy_rows, int w,
                   float qf, float* thr, int32_t* failed) {
  int nfail = 0;
  for (int64_t c = 0; c < N; ++c) {
    HostEnv env(anom + c, pitch, doy_ptr, doy_rows, w, Q, qf);
    marex::ExactQueue<Q, HostEnv> lane(env, w);
    g_lane = c;
    g_day = 0;
    const bool ok = lane.run([&](int d, float v) {
      thr[(int64_t)d * N + c] = v;
      g_day = d + 1 < 366 ? d + 1 : 365;
    });
    failed[c] = ok ? 0 : 1;
    nfail += ok ? 0 : 1;
  }
  return nfail;
}

// events: optional [N][366][6] int16, zeroed by the caller
extern "C" int xq_host(const float* anom, int64_t N, int64_t pitch, const int32_t* doy_ptr, const int32_t* doy_rows,
                       int w, float percentile, int Q, float* thr, int32_t* failed, short* events) {
  const float qf = percentile / 100.0f;
  int r = -1;
  g_events = events;
  if (Q == 64) r = run_all<64>(anom, N, pitch, doy_ptr, doy_rows, w, qf, thr, failed);
  if (Q == 128) r = run_all<128>(anom, N, pitch, doy_ptr, doy_rows, w, qf, thr, failed);
  if (Q == 16) r = run_all<16>(anom, N, pitch, doy_ptr, doy_rows, w, qf, thr, failed);
  g_events = nullptr;
  return r;
}

extern "C" int xq_kk_max_host(int rows, float percentile) {
  HostEnv env(nullptr, 0, nullptr, nullptr, 1, 1, percentile / 100.0f);
  int r0, r1;
  float g;
  env.rank(rows < 1 ? 1 : rows, r0, r1, g);
  return (rows < 1 ? 1 : rows) - r0;
}
