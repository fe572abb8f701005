// TEST INFRASTRUCTURE ONLY — see scan_oracle.h for the scope and the pinning status of each piece.
// Every routine below cites the reference file:line whose arithmetic it restates (paths relative to /root/reference).
#include "scan_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <dlfcn.h>
#include <pthread.h>
#include <sched.h>
#include <complex>
#include <cstring>
#include <deque>
#include <limits>
#include <map>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

// sources/radio/blocks/psd.cpp:18-20, all fp32: 10 * log10(|z|^2 / fs); `lin` (optional) receives |z|^2 / fs
static inline float psdDb(std::complex<float> z, float fs, float* lin) {
  const float mag = std::abs(z);
  const float pw = std::pow(mag, 2.0f) / fs;
  if (lin) *lin = pw;
  return 10.0f * std::log10(pw);
}

namespace {

constexpr float kNoData = -100.0f;  // setNoData, sources/utils/radio_utils.cpp:72-76

// ---------------------------------------------------------------------------------------------------------------
// helpers that the reference keeps in sources/utils
// ---------------------------------------------------------------------------------------------------------------

// sources/utils/utils.cpp:31-53 — centred boxcar, ONE running sum swept left to right, shrinking at the edges.
void boxcar(const float* in, float* out, int size, int group) {
  const int half = group / 2;
  float running = 0.0;
  int terms = 0;
  for (int pos = -half; pos < size + half - 1; ++pos) {
    const int leaving = pos - half - 1;
    const int entering = pos + half;
    if (0 <= leaving && leaving < size) {
      running -= in[leaving];
      terms--;
    }
    if (0 <= entering && entering < size) {
      running += in[entering];
      terms++;
    }
    if (0 <= pos && pos < size) {
      out[pos] = running / terms;
    }
  }
}

// sources/utils/collection_utils.h:9-14 — first maximum inside [index-g/2, index+g/2] clipped to the row.
int maxIndex(const float* data, int size, int index, int group) {
  const int lo = std::max(0, index - group / 2);
  const int hi = std::min(size, index + group / 2 + 1);
  int best = lo;
  for (int i = lo + 1; i < hi; ++i) {
    if (data[best] < data[i]) best = i;
  }
  return best;
}

// sources/utils/collection_utils.h:17-27 — lower_bound on the ordered keys; margin is halved, odd margins round up.
template <typename Map>
bool withinMargin(const Map& keys, int index, int margin, int* found) {
  const int sub = (margin % 2 == 0) ? margin / 2 : margin / 2 + 1;
  auto it = keys.lower_bound(index - sub);
  if (it != keys.end() && it->first <= index + sub) {
    if (found) *found = it->first;
    return true;
  }
  return false;
}

// sources/utils/collection_utils.h:30-50 — mode; among equally frequent values the upper median of the tied set.
// The reference indexes buffer[0] of an empty vector when data is empty (UB); the caller defines that case.
int modeValue(const std::vector<int>& data) {
  std::map<int, int> hist;
  for (int v : data) hist[v]++;
  int top = 0;
  for (const auto& kv : hist) top = std::max(top, kv.second);
  std::vector<int> tied;
  for (const auto& kv : hist) {
    if (kv.second == top) tied.push_back(kv.first);  // std::map iterates ascending == the reference's sort by value
  }
  return tied[tied.size() / 2];
}

// sources/utils/radio_utils.cpp:98-104
int fftFor(int32_t sampleRate, int32_t maxStep) {
  uint32_t n = 1;
  while (maxStep < static_cast<double>(sampleRate) / n) n <<= 1;
  return static_cast<int>(n);
}

// sources/utils/radio_utils.cpp:86-96 — nearest multiple of step, ties go up, negative-safe.
int32_t tuned(int32_t f, int32_t step) {
  const int32_t rest = f < 0 ? f % step + step : f % step;
  const int32_t down = f - rest;
  return (rest < step - rest) ? down : down + step;
}

// ---------------------------------------------------------------------------------------------------------------
// window + FFT: out of tree in the reference (GNU Radio 3.10 gr::fft::window::hamming / fft_v<gr_complex,true>,
// call site sources/radio/sdr_device.cpp:164). Restated from the documented behaviour.
// ---------------------------------------------------------------------------------------------------------------
void hamming(int n, float* w) {
  if (n == 1) {
    w[0] = 1.0f;
    return;
  }
  const double m = static_cast<double>(n - 1);
  for (int i = 0; i < n; ++i) w[i] = static_cast<float>(0.54 - 0.46 * std::cos((2.0 * M_PI * i) / m));
}

template <typename T>
struct FftPlan {
  int n = 0;
  std::vector<std::complex<T>> tw;  // exp(-2*pi*i*k/n), k < n/2
  std::vector<int> rev;
  explicit FftPlan(int size) : n(size), tw(size / 2), rev(size) {
    for (int k = 0; k < n / 2; ++k) {
      const double a = -2.0 * M_PI * k / n;
      tw[k] = std::complex<T>(static_cast<T>(std::cos(a)), static_cast<T>(std::sin(a)));
    }
    int bits = 0;
    while ((1 << bits) < n) ++bits;
    for (int i = 0; i < n; ++i) {
      int r = 0;
      for (int b = 0; b < bits; ++b) r |= ((i >> b) & 1) << (bits - 1 - b);
      rev[i] = r;
    }
  }
  // unnormalised forward DFT, in place, decimation in time
  void run(std::complex<T>* x) const {
    for (int i = 0; i < n; ++i) {
      if (i < rev[i]) std::swap(x[i], x[rev[i]]);
    }
    for (int len = 2; len <= n; len <<= 1) {
      const int half = len / 2;
      const int stride = n / len;
      for (int base = 0; base < n; base += len) {
        for (int k = 0; k < half; ++k) {
          const std::complex<T> w = tw[k * stride];
          const std::complex<T> a = x[base + k];
          const std::complex<T> b = x[base + k + half];
          const std::complex<T> t(b.real() * w.real() - b.imag() * w.imag(), b.real() * w.imag() + b.imag() * w.real());
          x[base + k] = a + t;
          x[base + k + half] = a - t;
        }
      }
    }
  }
};

// fp32 Stockham radix-4 (+ one radix-2) used only as the TIMED cpu baseline: a faithful stand-in for the
// reference's single-threaded FFTW3f plan (which is unavailable here), not a parity source.
struct FftF32 {
  int n;
  std::vector<std::complex<float>> tw;  // exp(-2 pi i k / n), k < n
  mutable std::vector<std::complex<float>> tmp;
  explicit FftF32(int size) : n(size), tw(size), tmp(size) {
    for (int k = 0; k < n; ++k) {
      const double a = -2.0 * M_PI * k / n;
      tw[k] = std::complex<float>(static_cast<float>(std::cos(a)), static_cast<float>(std::sin(a)));
    }
  }
  void run(std::complex<float>* x) const {
    std::complex<float>* src = x;
    std::complex<float>* dst = tmp.data();
    int l = 1;  // product of the radices already done
    int remaining = n;
    while (remaining > 1) {
      if (remaining % 4 == 0) {
        const int m = n / 4;  // butterflies
        for (int i = 0; i < m; ++i) {
          const int k = i & (l - 1);
          const int j = ((i - k) << 2) + k;
          const int tstep = n / (4 * l);
          std::complex<float> a = src[i], b = src[i + m], c = src[i + 2 * m], d = src[i + 3 * m];
          if (k) {
            b *= tw[k * tstep];
            c *= tw[2 * k * tstep];
            d *= tw[3 * k * tstep];
          }
          const std::complex<float> s0 = a + c, s1 = a - c, s2 = b + d, s3 = b - d;
          const std::complex<float> s3r(s3.imag(), -s3.real());  // -i * s3
          dst[j] = s0 + s2;
          dst[j + l] = s1 + s3r;
          dst[j + 2 * l] = s0 - s2;
          dst[j + 3 * l] = s1 - s3r;
        }
        l *= 4;
        remaining /= 4;
      } else {
        const int m = n / 2;
        for (int i = 0; i < m; ++i) {
          const int k = i & (l - 1);
          const int j = ((i - k) << 1) + k;
          const int tstep = n / (2 * l);
          std::complex<float> a = src[i], b = src[i + m];
          if (k) b *= tw[k * tstep];
          dst[j] = a + b;
          dst[j + l] = a - b;
        }
        l *= 2;
        remaining /= 2;
      }
      std::swap(src, dst);
    }
    if (src != x) std::memcpy(x, src, sizeof(std::complex<float>) * n);
  }
};

// FFTW3f, when the box has it (dlopen: no build-time dependency): the reference's real FFT engine (gr::fft::fft_v plans FFTW3f with
// FFTW_MEASURE, one thread). Used by the TIMED baseline only — never for parity (its rounding differs from the fp64 oracle's like any
// fp32 FFT). Absent in the build image; probed at run time so that a box that has libfftw3f.so.3 reports the "FFTW path" row.
struct Fftw {
  using plan_t = void*;
  plan_t (*plan_dft_1d)(int, float (*)[2], float (*)[2], int, unsigned) = nullptr;
  void (*execute_dft)(plan_t, float (*)[2], float (*)[2]) = nullptr;
  void (*destroy_plan)(plan_t) = nullptr;
  void* (*malloc_)(size_t) = nullptr;
  void (*free_)(void*) = nullptr;
  bool ok = false;
  std::mutex planner;  // FFTW's planner is not thread safe
  Fftw() {
    void* h = dlopen("libfftw3f.so.3", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libfftw3f.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return;
    plan_dft_1d = reinterpret_cast<decltype(plan_dft_1d)>(dlsym(h, "fftwf_plan_dft_1d"));
    execute_dft = reinterpret_cast<decltype(execute_dft)>(dlsym(h, "fftwf_execute_dft"));
    destroy_plan = reinterpret_cast<decltype(destroy_plan)>(dlsym(h, "fftwf_destroy_plan"));
    malloc_ = reinterpret_cast<decltype(malloc_)>(dlsym(h, "fftwf_malloc"));
    free_ = reinterpret_cast<decltype(free_)>(dlsym(h, "fftwf_free"));
    ok = plan_dft_1d && execute_dft && destroy_plan && malloc_ && free_;
  }
  static Fftw& get() {
    static Fftw f;
    return f;
  }
};
struct FftwPlan {
  Fftw::plan_t plan = nullptr;
  float (*buf)[2] = nullptr;
  int n = 0;
  explicit FftwPlan(int size) : n(size) {
    Fftw& f = Fftw::get();
    if (!f.ok) return;
    std::lock_guard<std::mutex> lk(f.planner);
    buf = static_cast<float (*)[2]>(f.malloc_(sizeof(float) * 2 * n));
    if (buf) plan = f.plan_dft_1d(n, buf, buf, -1 /* FFTW_FORWARD */, 0u /* FFTW_MEASURE */);
  }
  ~FftwPlan() {
    Fftw& f = Fftw::get();
    if (!f.ok) return;
    std::lock_guard<std::mutex> lk(f.planner);
    if (plan) f.destroy_plan(plan);
    if (buf) f.free_(buf);
  }
  bool usable() const { return plan != nullptr; }
  void run(std::complex<float>* x) const {  // in place through the aligned planning buffer
    std::memcpy(buf, static_cast<const void*>(x), sizeof(float) * 2 * n);
    Fftw::get().execute_dft(plan, buf, buf);
    std::memcpy(static_cast<void*>(x), buf, sizeof(float) * 2 * n);
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Averager — sources/radio/averager.cpp:7-60 (state must be bit-exact)
// ---------------------------------------------------------------------------------------------------------------
struct RingAverager {
  int size, group, frames = 0;
  std::vector<float> sum, mean;
  std::deque<std::vector<float>> rows;  // oldest at the front
  RingAverager(int s, int g) : size(s), group(g), sum(s, 0.0f), mean(s, 0.0f) {
    for (int i = 0; i < group; ++i) rows.emplace_back(size, 0.0f);
    refresh();
  }
  void push(const float* data) {  // averager.cpp:14-25
    frames = std::min(frames + 1, group);
    std::vector<float> row = std::move(rows.front());
    for (int i = 0; i < size; ++i) sum[i] -= row[i];  // subtract(), :46-50
    rows.pop_front();
    std::memcpy(row.data(), data, sizeof(float) * size);
    for (int i = 0; i < size; ++i) sum[i] += row[i];  // add(), :40-44
    rows.push_back(std::move(row));
    refresh();
  }
  void reset() {  // averager.cpp:27-34
    std::fill(sum.begin(), sum.end(), 0.0f);
    for (auto& r : rows) std::fill(r.begin(), r.end(), 0.0f);
    frames = 0;
    refresh();
  }
  void refresh() {  // updateAverage(), averager.cpp:52-60 — float / int => true float division
    if (group <= frames) {
      for (int i = 0; i < size; ++i) mean[i] = sum[i] / group;
    } else {
      for (int i = 0; i < size; ++i) mean[i] = kNoData;
    }
  }
};

// sources/radio/blocks/noise_learner.cpp:11-28 on the injected frame clock. learningMs > 0: the reference's own rule - m_startLearningTime
// is the stamp of the first frame this centre sees (Noise() runs inside the first work() call, :9,42), and the frame whose stamp reaches
// start + NOISE_LEARNING_TIME is the last learning frame (:23). learningMs == 0: a frame count instead of the clock.
struct NoiseState {
  std::vector<float> threshold;
  int samples = 0;
  bool ready = false;
  bool started = false;
  int64_t start = 0;
  // returns true when this frame completes (or already completed) learning
  bool add(const float* data, int size, int learnFrames, int64_t now, int64_t learningMs) {
    if (ready) return true;
    if (!started) {
      started = true;
      start = now;
    }
    if (static_cast<int>(threshold.size()) < size) threshold.resize(size, -std::numeric_limits<float>::max());
    for (int i = 0; i < size; ++i) threshold[i] = std::max(threshold[i], data[i]);
    samples++;
    if (learningMs > 0 ? start + learningMs <= now : learnFrames <= samples) {
      ready = true;
      return true;
    }
    return false;
  }
};

// sources/radio/signal.cpp:6-40 (times in ms, injected)
struct TrackedSignal {
  int64_t first, last;
  float power = 0.0f;
};

// sources/radio/blocks/spectrogram.cpp:9 — the reference leaves m_counter uninitialised; defined as 0 here.
struct SpectrogramBin {
  std::vector<float> sum;
  int counter = 0;
  int64_t lastSend;
  SpectrogramBin(int size, int64_t now) : sum(size, 0.0f), lastSend(now) {}
};

struct SentRow {
  int64_t time;
  int32_t center;
  std::vector<int8_t> row;
};

}  // namespace

struct orc_averager {
  RingAverager impl;
  orc_averager(int s, int g) : impl(s, g) {}
};

struct orc_chain {
  orc_config cfg;
  std::vector<float> window;
  FftPlan<double> plan64;
  FftF32 plan32;
  FftwPlan planw;  // the timed baseline prefers FFTW3f when the box has it
  double stageSeconds[3] = {0.0, 0.0, 0.0};  // timed baseline only: unpack+window+FFT, PSD, noise+averager+detect(+spectrogram)
  bool timeStages = false;
  std::map<int32_t, NoiseState> noise;  // keyed by centre frequency, noise_learner.cpp:41-42
  RingAverager averager;
  std::map<int, TrackedSignal> signals;  // transmission.h:49
  std::map<int32_t, SpectrogramBin> spectro;
  std::vector<SentRow> sent;
  int32_t center, rangeLo, rangeHi;
  // scratch
  std::vector<std::complex<double>> work64;
  std::vector<std::complex<float>> work32;
  std::vector<float> psd, sub, box;

  explicit orc_chain(const orc_config& c)
      : cfg(c),
        window(c.fft_size),
        plan64(c.fft_size),
        plan32(c.fft_size),
        planw(c.fft_size),
        averager(c.fft_size, c.grouping_y),
        center(c.center_hz),
        rangeLo(c.range_lo_hz),
        rangeHi(c.range_hi_hz),
        work64(c.fft_size),
        work32(c.fft_size),
        psd(c.fft_size),
        sub(c.fft_size),
        box(c.fft_size) {
    if (c.window_kind == 1 && c.window_taps) {
      std::memcpy(window.data(), c.window_taps, sizeof(float) * c.fft_size);
    } else {
      hamming(c.fft_size, window.data());
    }
    cfg.window_taps = nullptr;
  }

  // sources/radio/sdr_device.cpp:150,153-154 — index <-> frequency lambdas
  double step() const { return static_cast<double>(cfg.sample_rate_hz) / cfg.fft_size; }
  int32_t indexToShift(int i) const { return static_cast<int32_t>(step() * (i + 0.5)) - cfg.sample_rate_hz / 2; }
  int32_t indexToFrequency(int i) const { return center + static_cast<int32_t>(step() * (i + 0.5)) - cfg.sample_rate_hz / 2; }
  bool inRange(int i) const {  // sdr_device.cpp:155-158
    const int32_t f = indexToFrequency(i);
    return rangeLo <= f && f <= rangeHi;
  }
  bool ignored(int i) const {  // transmission.cpp:156-164
    const int32_t f = indexToFrequency(i);
    for (int r = 0; r < cfg.n_ignored; ++r) {
      if (cfg.ignored_lo_hz[r] <= f && f <= cfg.ignored_hi_hz[r]) return true;
    }
    return false;
  }

  // unpack -> window -> FFT -> fftshift -> PSD for one frame
  void framePsd(const void* iqFrame, float* out, float* lin) {
    const int n = cfg.fft_size;
    const bool f32 = (cfg.flags & 1) != 0;
    const auto ts0 = timeStages ? std::chrono::steady_clock::now() : std::chrono::steady_clock::time_point();
    for (int i = 0; i < n; ++i) {
      float re, im;
      if (cfg.iq_format == 0) {  // CS8; the unpack lives in the SoapySDR driver in the reference (sdr_source.cpp:52)
        const int8_t* p = static_cast<const int8_t*>(iqFrame);
        re = static_cast<float>(p[2 * i]) * cfg.iq_scale;
        im = static_cast<float>(p[2 * i + 1]) * cfg.iq_scale;
      } else {
        const float* p = static_cast<const float*>(iqFrame);
        re = p[2 * i];
        im = p[2 * i + 1];
      }
      // gr::fft::fft_v window multiply (VOLK volk_32fc_32f_multiply_32fc): fp32 products
      const float wr = re * window[i];
      const float wi = im * window[i];
      if (f32) {
        work32[i] = std::complex<float>(wr, wi);
      } else {
        work64[i] = std::complex<double>(wr, wi);
      }
    }
    if (f32) {
      if (planw.usable()) planw.run(work32.data()); else plan32.run(work32.data());
    } else {
      plan64.run(work64.data());
    }
    const auto ts1 = timeStages ? std::chrono::steady_clock::now() : std::chrono::steady_clock::time_point();
    const float fs = static_cast<float>(cfg.sample_rate_hz);  // int32 promoted to float in psd.cpp:19
    for (int j = 0; j < n; ++j) {
      const int k = (j + n / 2) % n;  // fft_v shift=true: out[j] = X[(j + N/2) mod N]
      std::complex<float> z;
      if (f32) {
        z = work32[k];
      } else {
        z = std::complex<float>(static_cast<float>(work64[k].real()), static_cast<float>(work64[k].imag()));
      }
      out[j] = psdDb(z, fs, lin ? &lin[j] : nullptr);
    }
    if (timeStages) {
      const auto ts2 = std::chrono::steady_clock::now();
      stageSeconds[0] += std::chrono::duration<double>(ts1 - ts0).count();
      stageSeconds[1] += std::chrono::duration<double>(ts2 - ts1).count();
    }
  }

  // sources/radio/blocks/transmission.cpp:132-154
  int bestIndex(int index) const {
    std::vector<int> votes;
    const size_t total = averager.rows.size();
    for (size_t r = total / 2; r < total; ++r) {
      const auto& row = averager.rows.at(r);
      const int best = maxIndex(row.data(), static_cast<int>(row.size()), index, cfg.group_size_bins);
      if (cfg.start_level <= row[best]) votes.push_back(best);
    }
    if (votes.empty()) return index;  // reference: UB (collection_utils.h:46-49); defined as "keep the candidate"
    return modeValue(votes);
  }

  // sources/radio/blocks/transmission.cpp:57-68, with `now` injected
  int64_t lastNow = 0;  // clock of the most recent frame (for orc_chain_get_transmissions)
  void detect(const float* power, int64_t now, int frameNo, const orc_outputs* out) {
    const int n = cfg.fft_size;
    lastNow = now;
    averager.push(power);
    boxcar(averager.mean.data(), box.data(), n, cfg.grouping_x);
    if (out && out->avg_db) std::memcpy(out->avg_db + static_cast<size_t>(frameNo) * n, averager.mean.data(), sizeof(float) * n);
    if (out && out->box_db) std::memcpy(out->box_db + static_cast<size_t>(frameNo) * n, box.data(), sizeof(float) * n);

    // addSignals, transmission.cpp:88-111. std::sort there is unstable; ties are defined as "lower index first".
    std::vector<int> cand;
    for (int i = 0; i < n; ++i) {
      if (cfg.start_level <= box[i] && inRange(i) && !ignored(i)) cand.push_back(i);
    }
    std::stable_sort(cand.begin(), cand.end(), [&](int a, int b) { return box[a] > box[b]; });
    for (int idx : cand) {
      if (!withinMargin(signals, idx, cfg.group_size_bins, nullptr)) {
        const int key = bestIndex(idx);
        signals.insert({key, TrackedSignal{now, now, 0.0f}});
      }
    }
    // updateSignals, transmission.cpp:113-130 + Signal::newData, signal.cpp:16-24
    for (auto& kv : signals) {
      const int bestAvg = maxIndex(box.data(), n, kv.first, cfg.group_size_bins);
      const float p = box[bestAvg];
      kv.second.power = p;
      if (cfg.stop_level <= p) kv.second.last = now;
    }
    // clearSignals, transmission.cpp:70-86 (isTimeout / isMaximalTime, signal.cpp:28-30)
    for (auto it = signals.begin(); it != signals.end();) {
      const bool timeout = it->second.last + cfg.timeout_ms <= now;
      const bool tooLong = it->second.first + cfg.max_time_ms <= now;
      if (timeout || tooLong) {
        it = signals.erase(it);
      } else {
        ++it;
      }
    }
    // getSortedTransmissions, transmission.cpp:166-176 (power descending; ties: lower key first)
    if (out && out->tx_count) {
      std::vector<int> keys;
      for (const auto& kv : signals) keys.push_back(kv.first);
      std::stable_sort(keys.begin(), keys.end(), [&](int a, int b) { return signals.at(a).power > signals.at(b).power; });
      const int count = std::min<int>(static_cast<int>(keys.size()), ORC_MAX_TX);
      out->tx_count[frameNo] = static_cast<int32_t>(keys.size());
      for (int s = 0; s < count; ++s) {
        const auto& sig = signals.at(keys[s]);
        const size_t o = static_cast<size_t>(frameNo) * ORC_MAX_TX + s;
        const bool flush = (sig.last == now) && (sig.first + cfg.min_time_ms <= now);  // signal.cpp:26,32
        if (out->tx_freq) out->tx_freq[o] = tuned(indexToShift(keys[s]), cfg.tuning_step_hz);
        if (out->tx_flush) out->tx_flush[o] = flush ? 1 : 0;
        if (out->tx_key) out->tx_key[o] = keys[s];
        if (out->tx_power) out->tx_power[o] = sig.power;
      }
    }
  }

  // sources/radio/blocks/spectrogram.cpp:29-75
  void spectrogram(const float* raw, int64_t now) {
    const int outN = cfg.spectrogram_out_size;
    if (outN <= 0) return;
    const int d = cfg.fft_size / outN;
    auto it = spectro.find(center);
    if (it == spectro.end()) it = spectro.emplace(center, SpectrogramBin(outN, now)).first;
    SpectrogramBin& c = it->second;
    if (d == 1) {
      for (int i = 0; i < outN; ++i) c.sum[i] += raw[i];
    } else {
      for (int i = 0; i < outN; ++i) {
        float s = 0.0;
        for (int j = 0; j < d; ++j) s += raw[i * d + j];
        c.sum[i] += s / d;
      }
    }
    c.counter++;
    if (c.lastSend + cfg.spectrogram_interval_ms < now) {
      SentRow r{now, center, std::vector<int8_t>(outN)};
      for (int j = 0; j < outN; ++j) r.row[j] = static_cast<int8_t>(c.sum[j] / c.counter);  // float -> int8 truncation
      sent.push_back(std::move(r));
      std::fill(c.sum.begin(), c.sum.end(), 0.0f);
      c.counter = 0;
      c.lastSend = now;
    }
  }

  // psdRows != nullptr: the frames' PSD rows are given (everything after PSD::work); otherwise they are computed from iq
  int push(const void* iq, const float* psdRows, size_t nFrames, int64_t t0, double period, const orc_outputs* out) {
    const int n = cfg.fft_size;
    const size_t bytesPerSample = cfg.iq_format == 0 ? 2 : 8;
    for (size_t k = 0; k < nFrames; ++k) {
      const int64_t now = t0 + static_cast<int64_t>(std::floor(static_cast<double>(k) * period + 0.5));
      if (psdRows) {
        std::memcpy(psd.data(), psdRows + k * n, sizeof(float) * n);
      } else {
        // stream_to_vector + Decimator: first N samples of each r*N group (decimator.h:16-22)
        const char* frame = static_cast<const char*>(iq) + k * static_cast<size_t>(cfg.frame_stride_samples) * bytesPerSample;
        framePsd(frame, psd.data(), nullptr);
      }
      const auto td0 = timeStages ? std::chrono::steady_clock::now() : std::chrono::steady_clock::time_point();
      if (out && out->psd_db) std::memcpy(out->psd_db + k * n, psd.data(), sizeof(float) * n);
      spectrogram(psd.data(), now);  // wired to the raw PSD, sdr_device.cpp:170-171

      // NoiseLearner::work, noise_learner.cpp:36-67
      NoiseState& ns = noise[center];
      int peak = -1;
      if (!ns.ready) {
        ns.add(psd.data(), n, cfg.learn_frames, now, cfg.noise_learning_ms);
        for (int j = 0; j < n; ++j) sub[j] = kNoData;  // also on the frame that completes learning (:45-51)
      } else {
        peak = 0;
        for (int j = 0; j < n; ++j) {
          sub[j] = psd[j] - ns.threshold[j];
          if (psd[peak] < psd[j]) peak = j;
        }
      }
      if (out && out->noise_sub_db) std::memcpy(out->noise_sub_db + k * n, sub.data(), sizeof(float) * n);
      if (out && out->peak_index) out->peak_index[k] = peak;
      detect(sub.data(), now, static_cast<int>(k), out);
      if (timeStages) stageSeconds[2] += std::chrono::duration<double>(std::chrono::steady_clock::now() - td0).count();
    }
    return 0;
  }
};

extern "C" {

orc_chain* orc_chain_create(const orc_config* cfg) {
  if (!cfg || cfg->fft_size < 2 || (cfg->fft_size & (cfg->fft_size - 1)) || cfg->learn_frames < 1) return nullptr;
  return new orc_chain(*cfg);
}
void orc_chain_destroy(orc_chain* c) { delete c; }
int orc_chain_push(orc_chain* c, const void* iq, size_t n, int64_t t0, double period, const orc_outputs* out) { return c->push(iq, nullptr, n, t0, period, out); }
int orc_chain_push_psd(orc_chain* c, const float* psd_rows, size_t n, int64_t t0, double period, const orc_outputs* out) {
  return psd_rows ? c->push(nullptr, psd_rows, n, t0, period, out) : -1;
}
void orc_chain_reset(orc_chain* c) {
  c->signals.clear();
  c->averager.reset();
}
void orc_chain_set_center(orc_chain* c, int32_t center, int32_t lo, int32_t hi) {
  c->center = center;
  c->rangeLo = lo;
  c->rangeHi = hi;
}
void orc_chain_get_averager(orc_chain* c, float* sum, float* avg, float* ring, int32_t* frames) {
  const int n = c->cfg.fft_size;
  if (sum) std::memcpy(sum, c->averager.sum.data(), sizeof(float) * n);
  if (avg) std::memcpy(avg, c->averager.mean.data(), sizeof(float) * n);
  if (ring) {
    size_t o = 0;
    for (const auto& r : c->averager.rows) {
      std::memcpy(ring + o, r.data(), sizeof(float) * n);
      o += n;
    }
  }
  if (frames) *frames = c->averager.frames;
}
int orc_chain_get_noise(orc_chain* c, float* thr, int32_t* samples) {
  auto it = c->noise.find(c->center);
  if (it == c->noise.end()) {
    if (samples) *samples = 0;
    return 0;
  }
  if (thr && !it->second.threshold.empty()) std::memcpy(thr, it->second.threshold.data(), sizeof(float) * c->cfg.fft_size);
  if (samples) *samples = it->second.samples;
  return it->second.ready ? 1 : 0;
}
int orc_chain_get_spectrogram(orc_chain* c, int64_t* times, int32_t* centers, int8_t* rows, int cap) {
  const int outN = c->cfg.spectrogram_out_size;
  const int count = static_cast<int>(c->sent.size());
  for (int i = 0; i < count && i < cap; ++i) {
    if (times) times[i] = c->sent[i].time;
    if (centers) centers[i] = c->sent[i].center;
    if (rows) std::memcpy(rows + static_cast<size_t>(i) * outN, c->sent[i].row.data(), outN);
  }
  return count;
}
void orc_chain_clear_spectrogram(orc_chain* c) { c->sent.clear(); }

// the live std::map<Index, Signal> (transmission.h:49) in key order
int orc_chain_get_signals(orc_chain* c, int32_t* keys, int64_t* first, int64_t* last, float* power, int cap) {
  int i = 0;
  for (const auto& kv : c->signals) {
    if (i < cap) {
      if (keys) keys[i] = kv.first;
      if (first) first[i] = kv.second.first;
      if (last) last[i] = kv.second.last;
      if (power) power[i] = kv.second.power;
    }
    ++i;
  }
  return i;
}
// getSortedTransmissions (transmission.cpp:166-176) after the most recent frame, without the ORC_MAX_TX bound of orc_outputs
int orc_chain_get_transmissions(orc_chain* c, int32_t* freq, int32_t* flush, int32_t* key, float* power, int cap) {
  std::vector<int> keys;
  for (const auto& kv : c->signals) keys.push_back(kv.first);
  std::stable_sort(keys.begin(), keys.end(), [&](int a, int b) { return c->signals.at(a).power > c->signals.at(b).power; });
  const int64_t now = c->lastNow;
  for (int s = 0; s < static_cast<int>(keys.size()) && s < cap; ++s) {
    const auto& sig = c->signals.at(keys[s]);
    if (freq) freq[s] = tuned(c->indexToShift(keys[s]), c->cfg.tuning_step_hz);
    if (flush) flush[s] = ((sig.last == now) && (sig.first + c->cfg.min_time_ms <= now)) ? 1 : 0;
    if (key) key[s] = keys[s];
    if (power) power[s] = sig.power;
  }
  return static_cast<int>(keys.size());
}

void orc_hamming(int n, float* w) { hamming(n, w); }
void orc_fft_f64(int n, const float* in, float* out) {
  FftPlan<double> plan(n);
  std::vector<std::complex<double>> x(n);
  for (int i = 0; i < n; ++i) x[i] = std::complex<double>(in[2 * i], in[2 * i + 1]);
  plan.run(x.data());
  for (int i = 0; i < n; ++i) {
    out[2 * i] = static_cast<float>(x[i].real());
    out[2 * i + 1] = static_cast<float>(x[i].imag());
  }
}
void orc_fft_f32(int n, const float* in, float* out) {
  FftF32 plan(n);
  std::vector<std::complex<float>> x(n);
  for (int i = 0; i < n; ++i) x[i] = std::complex<float>(in[2 * i], in[2 * i + 1]);
  plan.run(x.data());
  std::memcpy(out, x.data(), sizeof(float) * 2 * n);
}
void orc_psd_frame(const orc_config* cfg, const float* window, const void* iq, float* psd, float* lin) {
  orc_config c = *cfg;
  if (window) {
    c.window_kind = 1;
    c.window_taps = window;
  }
  if (c.learn_frames < 1) c.learn_frames = 1;
  if (c.grouping_y < 1) c.grouping_y = 1;
  orc_chain chain(c);
  chain.framePsd(iq, psd, lin);
}
void orc_average(const float* in, float* out, int size, int group) { boxcar(in, out, size, group); }
int orc_get_max_index(const float* data, int size, int index, int group) { return maxIndex(data, size, index, group); }
int orc_contains_with_margin(const int* keys, int nKeys, int index, int margin, int* found) {
  std::map<int, bool> m;
  for (int i = 0; i < nKeys; ++i) m[keys[i]] = false;
  return withinMargin(m, index, margin, found) ? 1 : 0;
}
int orc_most_frequent_value(const int* data, int n) {
  if (n <= 0) return -1;
  return modeValue(std::vector<int>(data, data + n));
}
int orc_get_fft(int32_t sampleRate, int32_t maxStep) { return fftFor(sampleRate, maxStep); }
int32_t orc_get_tuned_frequency(int32_t f, int32_t step) { return tuned(f, step); }

// DataController::pushSpectrogram (network/data_controller.cpp:44-57): fields appended in order, native (little-endian) layout
size_t orc_spectrogram_message(int64_t timeMs, int32_t frequency, int32_t sampleRate, const int8_t* data, int size, uint8_t* out) {
  std::vector<uint8_t> bytes;
  auto append = [&bytes](const void* p, size_t n) { bytes.insert(bytes.end(), static_cast<const uint8_t*>(p), static_cast<const uint8_t*>(p) + n); };
  const uint64_t stamp = static_cast<uint64_t>(timeMs);
  const int32_t start = frequency - sampleRate / 2, stop = frequency + sampleRate / 2, step = sampleRate / size;
  const uint32_t count = static_cast<uint32_t>(size);
  append(&stamp, 8);
  append(&start, 4);
  append(&stop, 4);
  append(&step, 4);
  append(&count, 4);
  append(data, static_cast<size_t>(size));
  std::memcpy(out, bytes.data(), bytes.size());
  return bytes.size();
}
// DataController::pushTransmission (network/data_controller.cpp:27-42): header, then the int8 IQ bytes with the sign bit flipped
size_t orc_transmission_message(int64_t timeMs, int32_t frequency, int32_t sampleRate, const int8_t* iq, int size, uint8_t* out) {
  std::vector<uint8_t> bytes;
  auto append = [&bytes](const void* p, size_t n) { bytes.insert(bytes.end(), static_cast<const uint8_t*>(p), static_cast<const uint8_t*>(p) + n); };
  const uint64_t stamp = static_cast<uint64_t>(timeMs);
  const int32_t start = frequency - sampleRate / 2, stop = frequency + sampleRate / 2;
  const uint32_t rate = static_cast<uint32_t>(sampleRate);
  append(&stamp, 8);
  append(&start, 4);
  append(&stop, 4);
  append(&rate, 4);
  const size_t header = bytes.size();
  append(iq, 2 * static_cast<size_t>(size));
  for (size_t i = header; i < bytes.size(); ++i) bytes[i] ^= 0x80;
  std::memcpy(out, bytes.data(), bytes.size());
  return bytes.size();
}

// PSD::work (psd.cpp:11-22) on `items` vectors of n complex values (no shift: fft_v does that)
void orc_psd_from_spectrum(int n, int32_t sampleRate, const float* x, float* out, int items) {
  const float fs = static_cast<float>(sampleRate);
  for (int i = 0; i < n * items; ++i) out[i] = psdDb(std::complex<float>(x[2 * i], x[2 * i + 1]), fs, nullptr);
}

orc_averager* orc_averager_create(int size, int group) { return new orc_averager(size, group); }
void orc_averager_destroy(orc_averager* a) { delete a; }
void orc_averager_push(orc_averager* a, const float* d) { a->impl.push(d); }
void orc_averager_reset(orc_averager* a) { a->impl.reset(); }
void orc_averager_average(orc_averager* a, float* out) { std::memcpy(out, a->impl.mean.data(), sizeof(float) * a->impl.size); }
void orc_averager_data(orc_averager* a, float* out) {
  size_t o = 0;
  for (const auto& r : a->impl.rows) {
    std::memcpy(out + o, r.data(), sizeof(float) * a->impl.size);
    o += a->impl.size;
  }
}
void orc_averager_sum(orc_averager* a, float* out) { std::memcpy(out, a->impl.sum.data(), sizeof(float) * a->impl.size); }
int orc_averager_frames(orc_averager* a) { return a->impl.frames; }

// stage split of the most recent orc_bench_run, summed over its threads: [0] unpack+window+FFT, [1] PSD (hypot/pow/log10), [2] noise +
// Averager + boxcar + detection + spectrogram
static double g_stage_seconds[3] = {0.0, 0.0, 0.0};
void orc_bench_stage_seconds(double* out) {
  for (int i = 0; i < 3; ++i) out[i] = g_stage_seconds[i];
}
const char* orc_fft_backend() { return Fftw::get().ok ? "FFTW3f (dlopen libfftw3f.so.3, FFTW_MEASURE, 1 thread per chain)" : "in-repo fp32 radix-4 Stockham (FFTW3f not on this box)"; }

double orc_bench_run(const orc_config* cfg, const void* iq, size_t nFrames, double period, int threads) {
  if (threads < 1) threads = 1;
  orc_config c = *cfg;
  c.flags |= 1;  // fp32 FFT: the timed baseline
  const size_t bytesPerSample = c.iq_format == 0 ? 2 : 8;
  std::vector<orc_chain*> chains;
  for (int t = 0; t < threads; ++t) {
    chains.push_back(new orc_chain(c));
    chains.back()->timeStages = true;
  }
  // one chain per thread, each thread pinned to its own core of the process's affinity mask (reproducible across boxes)
  std::vector<int> cpus;
  {
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0) {
      for (int i = 0; i < CPU_SETSIZE; ++i) {
        if (CPU_ISSET(i, &set)) cpus.push_back(i);
      }
    }
  }
  const size_t per = nFrames / threads;
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) {
    const size_t begin = per * t;
    const size_t count = (t == threads - 1) ? nFrames - begin : per;
    pool.emplace_back([&, t, begin, count]() {
      if (!cpus.empty()) {
        cpu_set_t one;
        CPU_ZERO(&one);
        CPU_SET(cpus[t % cpus.size()], &one);
        pthread_setaffinity_np(pthread_self(), sizeof(one), &one);
      }
      const char* base = static_cast<const char*>(iq) + begin * static_cast<size_t>(c.frame_stride_samples) * bytesPerSample;
      orc_outputs out{};
      std::vector<int32_t> cnt(count), freq(count * ORC_MAX_TX), fl(count * ORC_MAX_TX);
      out.tx_count = cnt.data();
      out.tx_freq = freq.data();
      out.tx_flush = fl.data();
      chains[t]->push(base, nullptr, count, 0, period, &out);
    });
  }
  for (auto& th : pool) th.join();
  const auto t1 = std::chrono::steady_clock::now();
  for (int i = 0; i < 3; ++i) g_stage_seconds[i] = 0.0;
  for (auto* ch : chains) {
    for (int i = 0; i < 3; ++i) g_stage_seconds[i] += ch->stageSeconds[i];
    delete ch;
  }
  return std::chrono::duration<double>(t1 - t0).count();
}
}
