// Exact day-of-year percentiles (np.nanpercentile over the +-w/2 day-of-year window of every gridpoint,
// detect.py:1921-1956) without keeping the window or a histogram of it: the q-quantile of n samples is decided by
// the kk = n - floor(q (n - 1)) LARGEST of them, so a gridpoint only keeps the samples above a PIVOT a little below
// its threshold.
//
//   * Samples leave the window in the order they entered (whole days of year), so the kept samples of a gridpoint
//     live in a first-in first-out queue of Q floats; per window day three byte counters say how many of that day's
//     samples are in the queue (> pivot), equal the pivot (ties are counted, not stored) and are valid (not NaN).
//     A day leaving the window pops its count from the head: no search, no histogram, 2 compares per entering sample.
//   * The two order statistics the percentile interpolates between are found by walking from the previous day's
//     answer: one pass over the queue counts the samples above / equal to the guess and finds the next distinct
//     values above and below it; the guess moves by one distinct value per pass (the window changes by 2/w of its
//     samples per step, so the answer rarely moves more than a rank or two).  The 32 lanes of a warp run these passes
//     in lockstep (a lane that is done feeds NaN, which compares false everywhere).
//   * The pivot follows the threshold through the year, lane by lane:
//       - queue filling up (threshold rising): the pivot is RAISED in place -- the queue is compacted in shared
//         memory, no global reads;
//       - fewer than kk samples left (threshold falling): the pivot is LOWERED by a fraction of the gridpoint's scale
//         and the lane's window re-read once; lanes that do not need it neither load nor change state;
//       - anything else (first window, a guess that missed, overflow): the pivot is found by BRACKETING -- 8 levels
//         per pass between the window's minimum and maximum, counting samples >= each level, until between
//         kk + SLACK and kk + ROOM samples lie at or above the pivot, or the bracket collapses onto a block of equal
//         values, which the tie counter absorbs.
//   * Exactness never depends on these heuristics: every state change is verified by counts, and a warp that cannot
//     find a pivot reports failure; its 32 gridpoints are recomputed by the histogram kernel
//     (hobday_exact_win_kernel, list mode).
//
// The lane algorithm is written against an environment `Env` (row loads, queue / counter storage, warp votes, the
// rank rule `rank(n, r0, r1, g)` -- ascending 0-based ranks of the two order statistics and the interpolation weight
// -- and `finish(a, b, g)`, which interpolates) so that the very same code runs per lane in the CUDA kernel and, with
// a one-lane environment, on the host for tests/test_exact_queue_host.py.
// (The same queue on uint16 bin codes for the approximate path without pooling was built and checked on the host: it
// needs a second selection pass per day for the bin's count, ~3,700 warp instructions per warp and day against the
// 1,230 of hobday_hist_kernel's byte histogram -- not pursued.)
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define XQ_HD __host__ __device__ __forceinline__
#else
#define XQ_HD inline
#endif

#ifndef XQ_SLACK_V
#define XQ_SLACK_V 14
#endif
#ifndef XQ_ROOM_V
#define XQ_ROOM_V 30
#endif
#ifndef XQ_RAISE_KEEP_V
#define XQ_RAISE_KEEP_V 20
#endif
#ifndef XQ_LOWER_PCT_V
#define XQ_LOWER_PCT_V 20
#endif
#ifndef XQ_DEPTH_V
#define XQ_DEPTH_V 4
#endif
#ifndef XQ_EVENT
#define XQ_EVENT(kind)
#define XQ_EVENT_N(kind, count)
#endif

namespace marex {

constexpr int XQ_SLACK = XQ_SLACK_V;            // bracketing keeps at least kk + SLACK samples at or above the pivot ...
constexpr int XQ_ROOM = XQ_ROOM_V;              // ... and at most kk + ROOM
constexpr int XQ_HEAD = 4;                      // kk + ROOM <= Q - HEAD (host-side dispatch rule)
#ifndef XQ_RAISE_GAP_V
#define XQ_RAISE_GAP_V 8
#endif
constexpr int XQ_RAISE_GAP = XQ_RAISE_GAP_V;                // raise the pivot when fewer than this many queue places are free
constexpr int XQ_RAISE_KEEP = XQ_RAISE_KEEP_V;  // a raise aims at kk + RAISE_KEEP kept samples
constexpr float XQ_LOWER = XQ_LOWER_PCT_V * 0.01f;  // a lowering moves the pivot down by this fraction of the scale
constexpr int XQ_DEPTH = XQ_DEPTH_V;  // neighbours of the guess a selection pass keeps on either side
constexpr int XQ_NDOY = 366;
constexpr int XQ_PASSES = 24;  // bracketing passes before a lane gives up

// WC: window_days_hobday as a compile-time constant (0 = run-time): slot arithmetic and the per-slot loops fold.
template <int Q, class Env, int WC = 0>
struct ExactQueue {
  Env& e;
  const int w, half;
  float pivot, x;  // x: the previous answer (lower order statistic), the guess of the next selection
  float scale;     // window maximum - pivot when the pivot was last bracketed: the step of a lowering
  int h, m_gt, m_eq, n;

  XQ_HD ExactQueue(Env& env, int w_)
      : e(env), w(WC ? WC : w_), half((WC ? WC : w_) / 2), pivot(-Env::inf()), x(-Env::inf()), scale(0.f), h(0), m_gt(0), m_eq(0), n(0) {}

  static XQ_HD int wrap(int d) { return ((d % XQ_NDOY) + XQ_NDOY) % XQ_NDOY; }

  // fn(v) for the rows of day of year dd; lanes with mine == false see NaN and load nothing.  Full batches of 8
  // independent loads, then the remainder one by one.
  template <class F>
  XQ_HD void scan_day(int dd, bool mine, F&& fn) {
    const int b0 = e.doy_begin(dd), b1 = e.doy_begin(dd + 1);
    int j = b0;
    for (; j + 8 <= b1; j += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = mine ? e.load(j + u) : Env::nan();
#pragma unroll
      for (int u = 0; u < 8; ++u) fn(v[u]);
    }
    for (; j < b1; ++j) fn(mine ? e.load(j) : Env::nan());
  }
  template <class F>
  XQ_HD void scan_window(int d, bool mine, F&& fn) {
    for (int k = -half; k <= half; ++k) scan_day(wrap(d + k), mine, fn);
  }

  // The day of year dd enters the window as window slot `slot`.  Branch-free: every sample is written to the tail
  // of the queue and the tail only advances when the sample is kept.  A full queue (m_gt >= Q) loses samples: the
  // caller sees it and re-reads the window.  Lanes with mine == false keep their state (their tail place is free).
  XQ_HD void enter_day(int slot, int dd, bool mine) {
    const int g0 = m_gt;
    int eq = 0, nv = 0;
    scan_day(dd, mine, [&](float v) {
      nv += (v == v) ? 1 : 0;
      eq += (v == pivot) ? 1 : 0;
      e.que((h + m_gt) & (Q - 1)) = v;
      m_gt += (v > pivot) ? 1 : 0;
    });
    if (mine) {
      e.cnt(slot) = (uint8_t)(m_gt - g0);
      e.eqc(slot) = (uint8_t)eq;
      e.nvc(slot) = (uint8_t)nv;
    }
    m_eq += eq;
    n += nv;
  }
  XQ_HD void leave_day(int slot) {
    const int lc = e.cnt(slot);
    h = (h + lc) & (Q - 1);
    m_gt -= lc;
    m_eq -= e.eqc(slot);
    n -= e.nvc(slot);
  }
  // Queue and counters of the lanes with mine == true refilled from the window centred on d, with the current pivot.
  XQ_HD void refill(int d, bool mine) {
    if (mine) { h = 0; m_gt = 0; m_eq = 0; n = 0; }
    for (int k = -half; k <= half; ++k) enter_day((d + k + half) % w, wrap(d + k), mine);
  }
  XQ_HD bool holds(int kk) const { return m_gt < Q && m_gt + m_eq >= kk; }

  // Pivot raised in place to np (> pivot) for the lanes with mine == true: the queue is compacted slot by slot,
  // oldest day first (slot d % w once day d + half has entered), so the first-in first-out order survives.
  XQ_HD void raise(int d, bool mine, float np) {
    int ri = 0, wi = 0, neq = 0;
    for (int t = 0; t < w; ++t) {
      const int slot = (d + t) % w;
      const int c = mine ? (int)e.cnt(slot) : 0;
      const int cmax = e.wmax(c);
      int kept = 0, eqs = 0;
      for (int j = 0; j < cmax; ++j) {
        const bool in = j < c;
        const float v = in ? e.que((h + ri) & (Q - 1)) : Env::nan();
        if (in) e.que((h + wi) & (Q - 1)) = v;
        ri += in ? 1 : 0;
        const int k = (v > np) ? 1 : 0;
        wi += k;
        kept += k;
        eqs += (v == np) ? 1 : 0;
      }
      if (mine) {
        e.cnt(slot) = (uint8_t)kept;
        e.eqc(slot) = (uint8_t)eqs;
      }
      neq += eqs;
    }
    if (mine) { m_gt = wi; m_eq = neq; pivot = np; }
  }

  // New pivot by bracketing for the lanes with mine == true (window centred on d), queue and counters refilled.
  // Returns false for a lane whose kk largest samples do not fit.
  XQ_HD bool rebuild(int d, bool mine) {
    XQ_EVENT(2);
    int nn = 0;
    float mn = Env::inf(), mx = -Env::inf();
    scan_window(d, mine, [&](float v) {
      nn += (v == v) ? 1 : 0;
      if (Env::finite(v)) { mn = Env::fmin(mn, v); mx = Env::fmax(mx, v); }
    });
    int kk = 0;
    bool ok = true;
    bool done = true;
    float lo = -Env::inf(), hi = Env::inf();
    bool hi_real = false;
    int clo = nn, need = 0, mhi = 0;
    if (mine && nn > 0) {
      int r0, r1;
      float g;
      e.rank(nn, r0, r1, g);
      kk = nn - r0;
      need = kk + XQ_SLACK;
      mhi = kk + XQ_ROOM;
      if (nn > mhi) {
        if (mn <= mx) done = false;  // bracket between the finite extremes
        else ok = false;             // only infinities, and too many of them
      }
    }
    for (int it = 0; it < XQ_PASSES && e.any(!done); ++it) {
      XQ_EVENT(3);
      // levels lob + (top - lob) * k / 8: k = 1..8 above a pivot candidate that already qualifies, k = 0..7 (the
      // minimum itself included: a block of equal values at the bottom) while none does
      const bool open = lo == -Env::inf();
      const float lob = open ? mn : lo, top = hi_real ? hi : mx;
      float f[8];
      int cge[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f[j] = Env::level(lob, top, open ? j : j + 1);
        cge[j] = 0;
      }
      int cgt = 0;  // samples strictly above the qualifying candidate
      scan_window(d, !done, [&](float v) {
        cgt += (v > lob) ? 1 : 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) cge[j] += (v >= f[j]) ? 1 : 0;
      });
      if (!done && !open && cgt < need) {
        done = true;  // a block of equal values at lo holds the rank: the tie counter absorbs it
      } else if (!done) {
        int jq = -1;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (cge[j] >= need) jq = j;
        float nlo = lo, nhi = hi;
        bool nhr = hi_real;
        int nclo = clo;
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // (static indexing: f and cge stay in registers)
          if (j == jq) { nlo = f[j]; nclo = cge[j]; }
          if (j == jq + 1) { nhi = f[j]; nhr = true; }
        }
        const bool progress = (nlo != lo) || (nhi != hi) || (nhr != hi_real);
        lo = nlo; hi = nhi; hi_real = nhr; clo = nclo;
        if (lo != -Env::inf() && clo <= mhi) done = true;
        else if (!progress) done = true;  // equal values at lo (checked below) or nothing to find
      }
    }
    if (mine) {  // (a lane that ran out of passes keeps its best candidate: the counts below decide)
      pivot = lo;
      scale = (mn <= mx && lo != -Env::inf()) ? mx - lo : 0.f;
    }
    refill(d, mine);
    if (mine && nn > 0 && !holds(kk)) ok = false;
    return ok;
  }

  // a = kk-th largest, b = kk1-th largest (kk1 = kk or kk - 1) of the window, from the queue and the tie counter,
  // for the lanes with act == true; needs holds(kk) there.  Warp-synchronous.  One pass over the queue counts the
  // samples above (G) and equal to (E) the guess g and keeps the DEPTH smallest samples above it (u0 <= u1 <= ...)
  // and the DEPTH largest below it (d0 >= d1 >= ...), multiplicities included; in descending order the window reads
  // ... u1 u0 [g x E] d0 d1 ... [pivot x m_eq], so the sample at position k (1 = largest) is u[G - k] for k <= G,
  // g up to G + E, d[k - G - E - 1] after that.  A position more than DEPTH samples away moves the guess there and
  // takes another pass.
  XQ_HD void select(bool act, int kk, int kk1, float& a, float& b) {
    float g = Env::fmax(x, pivot);  // a guess below the pivot (or NaN) starts at the pivot
    for (int it = 0; it < Q + 8 && e.any(act); ++it) {
      XQ_EVENT(4);
      int G = 0, E = 0;
      float u[XQ_DEPTH], dn[XQ_DEPTH];
#pragma unroll
      for (int k = 0; k < XQ_DEPTH; ++k) { u[k] = Env::inf(); dn[k] = -Env::inf(); }
      const int mlim = act ? m_gt : 0;
      const int mmax = e.wmax(mlim);
      XQ_EVENT_N(5, mmax);
      for (int i = 0; i < mmax; ++i) {
        const float v = (i < mlim) ? e.que((h + i) & (Q - 1)) : Env::nan();
        const bool gt = v > g, lt = v < g;
        G += gt ? 1 : 0;
        E += (v == g) ? 1 : 0;
        float c = gt ? v : Env::inf();
#pragma unroll
        for (int k = 0; k < XQ_DEPTH - 1; ++k) { const float t = Env::fmin(u[k], c); c = Env::fmax(u[k], c); u[k] = t; }
        u[XQ_DEPTH - 1] = Env::fmin(u[XQ_DEPTH - 1], c);
        c = lt ? v : -Env::inf();
#pragma unroll
        for (int k = 0; k < XQ_DEPTH - 1; ++k) { const float t = Env::fmax(dn[k], c); c = Env::fmin(dn[k], c); dn[k] = t; }
        dn[XQ_DEPTH - 1] = Env::fmax(dn[XQ_DEPTH - 1], c);
      }
      if (act) {
        const int L = m_gt - G - E;        // queue samples below the guess
        if (g == pivot) E += m_eq;         // (then L == 0: every queue sample is above the pivot)
        // sample at position k of the descending order; false when it is more than DEPTH samples from the guess
        auto at = [&](int k, float& out) -> bool {
          const int iu = G - k, id = k - G - E - 1;
          if (iu >= 0) {
            if (iu >= XQ_DEPTH) return false;
#pragma unroll
            for (int t = 0; t < XQ_DEPTH; ++t) if (t == iu) out = u[t];
          } else if (id < 0) {
            out = g;
          } else if (id >= L) {
            out = pivot;  // among the ties at the pivot (holds(kk): id < L + m_eq)
          } else {
            if (id >= XQ_DEPTH) return false;
#pragma unroll
            for (int t = 0; t < XQ_DEPTH; ++t) if (t == id) out = dn[t];
          }
          return true;
        };
        if (at(kk, a) && at(kk1, b)) act = false;
        else g = (kk <= G) ? u[XQ_DEPTH - 1] : dn[XQ_DEPTH - 1];
      }
    }
  }

  // All 366 days of year; out(d, value).  Returns false (for every lane of the warp) when some lane failed.
  template <class F>
  XQ_HD bool run(F&& out) {
    if (!e.all(rebuild(0, true))) return false;
    for (int d = 0; d < XQ_NDOY; ++d) {
      if (d > 0) {
        const int slot = (d - 1) % w;  // holds day d - 1 - half, which leaves; day d + half takes its place
        leave_day(slot);
        enter_day(slot, wrap(d + half), true);
      }
      int r0 = 0, r1 = 0, kk = 0, kk1 = 0;
      float gw = 0.f;
      const bool has = n > 0;
      if (has) {
        e.rank(n, r0, r1, gw);
        kk = n - r0;
        kk1 = n - r1;
      }
      // threshold rising: raise the pivot in place, towards kk + RAISE_KEEP kept samples if they were evenly spread
      // between the pivot and the previous answer
      const bool full = has && holds(kk) && m_gt > Q - XQ_RAISE_GAP && m_gt > kk + XQ_RAISE_KEEP && x > pivot;
      if (e.any(full)) {
        XQ_EVENT(0);
        const float np = x - (x - pivot) * ((float)XQ_RAISE_KEEP / (float)(m_gt - kk > 0 ? m_gt - kk : 1));
        raise(d, full && np > pivot && np < x, np);
      }
      // threshold falling: lower the pivot by a fraction of the scale and re-read the window once
      const bool low = has && m_gt < Q && m_gt + m_eq < kk && scale > 0.f;
      if (e.any(low)) {
        XQ_EVENT(1);
        if (low) pivot = pivot - XQ_LOWER * scale;
        refill(d, low);
      }
      const bool bad = has && !holds(kk);
      if (e.any(bad)) {
        if (!e.all(rebuild(d, bad))) return false;
      }
      float res = Env::nan(), a = res, b = res;
      select(has, kk, kk1, a, b);
      if (has) {
        res = e.finish(a, b, gw);
        x = a;
      }
      out(d, res);
    }
    return true;
  }
};

}  // namespace marex
