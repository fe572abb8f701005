// TEST INFRASTRUCTURE ONLY — C-ABI veneer over the REFERENCE's own objects.
//
// This file is compiled together with the reference's unmodified sources, taken where they lie under
// /root/reference (never copied into this repo), by oracle/Makefile -> oracle/_ref/libref.so:
//   sources/radio/averager.cpp, sources/utils/utils.cpp, sources/utils/radio_utils.cpp, sources/logger.cpp
//   + header-only sources/utils/collection_utils.h
// It lets pytest pin the CPU restatement in oracle/scan_oracle.cpp against the real reference code for every
// piece of the hot path that is buildable here. The reference's block objects (PSD, NoiseLearner,
// Transmission, Spectrogram, DataController) are driven by oracle/ref_blocks_shim.cpp, which lives in the same library.
#include <logger.h>
#include <radio/averager.h>
#include <utils/collection_utils.h>
#include <utils/radio_utils.h>
#include <utils/utils.h>

#include <cstring>
#include <map>
#include <vector>

namespace {
void ensureLogger() {
  static bool done = false;
  if (!done) {
    // tests/test_main.cpp:5 — the static logger is null until configured
    Logger::configure(spdlog::level::off, spdlog::level::off, "", 0, 0, true);
    done = true;
  }
}
}  // namespace

extern "C" {

// ---- Averager (sources/radio/averager.h:8-28) ----
void* ref_averager_create(int size, int groupSize) {
  ensureLogger();
  return new Averager(size, groupSize);
}
void ref_averager_destroy(void* h) { delete static_cast<Averager*>(h); }
void ref_averager_push(void* h, const float* data) { static_cast<Averager*>(h)->push(data); }
void ref_averager_reset(void* h) { static_cast<Averager*>(h)->reset(); }
void ref_averager_average(void* h, float* out) {
  const auto& v = static_cast<Averager*>(h)->average();
  std::memcpy(out, v.data(), sizeof(float) * v.size());
}
// rows oldest -> newest, out[groupSize][size]
void ref_averager_data(void* h, float* out) {
  const auto& d = static_cast<Averager*>(h)->data();
  size_t off = 0;
  for (const auto& row : d) {
    std::memcpy(out + off, row.data(), sizeof(float) * row.size());
    off += row.size();
  }
}

// ---- utils (sources/utils/utils.cpp:31-62) ----
void ref_average(const float* in, float* out, int size, int groupSize) { average(in, out, size, groupSize); }
int ref_round_up(int v, int f) { return roundUp(v, f); }
int ref_round_down(int v, int f) { return roundDown(v, f); }

// ---- collection utils (sources/utils/collection_utils.h:9-50) ----
int ref_get_max_index(const float* data, int size, int index, int groupSize) { return getMaxIndex(data, size, index, groupSize); }
// returns 1 and writes *found when a key lies within the margin, else 0
int ref_contains_with_margin(const int* keys, int nKeys, int index, int margin, int* found) {
  std::map<int, bool> m;
  for (int i = 0; i < nKeys; ++i) m[keys[i]] = false;
  const auto r = containsWithMargin(m, index, margin);
  if (r) {
    *found = *r;
    return 1;
  }
  return 0;
}
int ref_most_frequent_value(const int* data, int n) {
  std::vector<int> v(data, data + n);
  return mostFrequentValue(v);
}

// ---- radio utils (sources/utils/radio_utils.cpp:72-195) ----
void ref_set_no_data(float* data, int size) { setNoData(data, size); }
int ref_get_tuned_frequency(int f, int step) { return getTunedFrequency(f, step); }
int ref_get_fft(int sampleRate, int maxStep) { return getFft(sampleRate, maxStep); }
int ref_get_range_split_sample_rate(int sampleRate) { return getRangeSplitSampleRate(sampleRate); }
// out pairs (first, second); returns count
int ref_split_range(int lo, int hi, int sampleRate, int* out, int cap) {
  const auto r = splitRange({lo, hi}, sampleRate);
  int n = 0;
  for (const auto& p : r) {
    if (n < cap) {
      out[2 * n] = p.first;
      out[2 * n + 1] = p.second;
    }
    ++n;
  }
  return n;
}
int ref_get_resamplers_factors(int sampleRate, int bandwidth, int threshold, int* out, int cap) {
  ensureLogger();
  const auto r = getResamplersFactors(sampleRate, bandwidth, threshold);
  int n = 0;
  for (const auto& p : r) {
    if (n < cap) {
      out[2 * n] = p.first;
      out[2 * n + 1] = p.second;
    }
    ++n;
  }
  return n;
}
}
