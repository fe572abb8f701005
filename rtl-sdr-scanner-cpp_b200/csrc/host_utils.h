// Host-side helpers with the reference's semantics (product code: used by the tracker and exported through the C-ABI).
// Each function names the reference routine it stands in for; the arithmetic (integer truncation, tie rules) is the
// reference's, the code is not.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <vector>

namespace b2s {
namespace host {

// getFft — sources/utils/radio_utils.cpp:98-104: smallest power of two N with fs / N <= maxStep
inline int fft_size_for(int32_t sample_rate, int32_t max_step) {
  uint32_t n = 1;
  while (static_cast<double>(max_step) < static_cast<double>(sample_rate) / n) n <<= 1;
  return static_cast<int>(n);
}

// getTunedFrequency — sources/utils/radio_utils.cpp:86-96: round to the nearest multiple of step, halves go up
inline int32_t tuned_frequency(int32_t f, int32_t step) {
  int32_t r = f % step;
  if (f < 0) r += step;
  const int32_t below = f - r;
  return (r < step - r) ? below : below + step;
}

// getMaxIndex — sources/utils/collection_utils.h:9-14: first maximum of data[index-g/2 .. index+g/2] (clipped)
inline int max_index(const float* data, int size, int index, int group) {
  const int from = std::max(0, index - group / 2);
  const int to = std::min(size, index + group / 2 + 1);
  return static_cast<int>(std::max_element(data + from, data + to) - data);
}

// containsWithMargin — sources/utils/collection_utils.h:17-27
inline int margin_for(int group) { return (group % 2 == 0) ? group / 2 : group / 2 + 1; }

template <typename V>
inline bool key_within_margin(const std::map<int, V>& keys, int index, int group, int* found = nullptr) {
  const int m = margin_for(group);
  auto it = keys.lower_bound(index - m);
  if (it == keys.end() || it->first > index + m) return false;
  if (found) *found = it->first;
  return true;
}

// mostFrequentValue — sources/utils/collection_utils.h:30-50: the mode; among equally frequent values, the element at
// position size/2 of the ascending tied set. Empty input is undefined in the reference; callers handle it.
inline int most_frequent(std::vector<int> values) {
  std::sort(values.begin(), values.end());
  std::vector<int> tied;
  int best = 0;
  for (size_t i = 0; i < values.size();) {
    size_t j = i;
    while (j < values.size() && values[j] == values[i]) ++j;
    const int run = static_cast<int>(j - i);
    if (run > best) {
      best = run;
      tied.clear();
    }
    if (run == best) tied.push_back(values[i]);
    i = j;
  }
  return tied[tied.size() / 2];
}

// decimatorFactor — sources/radio/sdr_device.cpp:150-152
inline int decimator_factor(int32_t sample_rate, int32_t fft_size, int fps = 50) {
  const double step = static_cast<double>(sample_rate) / fft_size;
  return std::max(1, static_cast<int>(step / fps));
}

// frame clock: now_k = t0 + floor(k * period + 0.5)
inline int64_t frame_time(int64_t t0_ms, double period_ms, size_t k) { return t0_ms + static_cast<int64_t>(std::floor(static_cast<double>(k) * period_ms + 0.5)); }

}  // namespace host
}  // namespace b2s
