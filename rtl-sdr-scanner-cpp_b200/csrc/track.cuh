// K4 — the Transmission block's signal map on the device (one CTA per band and push).
//   addSignals / getBestIndex        reference sources/radio/blocks/transmission.cpp:88-111,132-154
//   updateSignals / Signal::newData  reference sources/radio/blocks/transmission.cpp:113-130, sources/radio/signal.cpp:16-24
//   clearSignals / isTimeout ...     reference sources/radio/blocks/transmission.cpp:70-86, sources/radio/signal.cpp:26-32
//   getSortedTransmissions           reference sources/radio/blocks/transmission.cpp:166-176
// The reference runs this per frame on the host; tracker.h does the same from K2's detection entries (and stays the path for
// callers that ask for every frame's list). Here the map lives in device memory across pushes and the host only reads the
// mailbox of the last frame, so a push costs the host one small device->host copy and no bookkeeping.
//
// The map only changes at EVENTS: a frame with a start-level bin that no live key covers (containsWithMargin), a time-out,
// or the 10-minute limit. Between events every frame is independent: a key's m_lastDataTime is the time of the latest frame
// whose window [key - g/2, key + g/2] holds a bin at or above the stop level. So the kernel walks the push in blocks of 1024
// frames, one thread per frame, finds the first event frame of the block in parallel, commits the frames before it in
// parallel, replays the event frame exactly like the reference (candidates by power, getBestIndex on the Averager ring rows,
// update, clear) and continues behind it. k_runs first folds each frame's detection entries into runs of consecutive bins
// (one per emitter and level), so the per-frame work does not grow with the width of a signal.
//
// m_power (only read when the list is emitted) is taken for the last frame of the push from K2's boxcar row of that frame.
// Tie rules left open by the reference's unstable std::sort are the oracle's: candidates (power desc, bin asc), transmissions
// (power desc, key asc).
#pragma once
#include "../../include/b2s.h"
#include "detect.cuh"

namespace b2s {

constexpr int kMaxSignals = 1024;  // live signals per band (keys are at least g/2 bins apart); beyond it the push fails loudly
constexpr int kRunCap = 24;        // runs kept per frame and level; a frame with more is replayed from its raw entries
constexpr int kTrackThreads = 1024, kTrackFrames = 1024, kTrackWords = kTrackFrames / 32;
constexpr int kKeyChunk = 128;     // keys whose per-frame hit bits are held in shared memory at a time
constexpr int kMaxCand = 4096;     // candidates of one event frame (= the largest detect_capacity)

struct TrackParams {  // Transmission's construction-time parameters (transmission.h:17-25) + the index lambdas of sdr_device.cpp:153-158
  int n, sample_rate, center, range_lo, range_hi;
  int n_ignored, ignored_lo[B2S_MAX_IGNORED], ignored_hi[B2S_MAX_IGNORED];
  int group_size, group_y;
  float start_level, stop_level;
  int tuning_step;
  long long min_time, timeout, max_time;
};

struct TrackState {  // std::map<Index, Signal> (transmission.h:49), keys ascending
  int n, error;      // error: 1 = more than kMaxSignals live signals, 2 = more than kMaxCand candidates in one frame
  int key[kMaxSignals];
  long long first[kMaxSignals], last[kMaxSignals];
  float power[kMaxSignals];
};

struct TrackResult {  // what the host reads back per push
  int n_tx, n_entries, max_count, error;
  long long last_now;
  b2s_transmission tx[kMaxSignals];  // getSortedTransmissions after the last frame
};

struct FrameRuns {  // runs of consecutive bins of one frame: [0] at or above the stop level, [1] start-level candidates
  int count[2];     // may exceed kRunCap (then the frame is "complex")
  int lo[2][kRunCap], hi[2][kRunCap];
};

struct TrackArgs {
  TrackParams p;
  int n_frames;
  long long t0_ms;
  double period_ms;
  long long frame_offset;
  const DetectEntry* entries;  // ordered by (frame, bin)
  const int* offsets;          // [T + 1]
  const int* max_count;        // largest per-frame entry count (overflow report)
  const FrameRuns* runs;       // [T]
  const float* box_last;       // [N] boxcar row of the last frame
  // getBestIndex inputs: noise-subtracted rows = the Averager ring
  const float* psd;            // [T][N]
  const float* threshold;      // [N]
  int noise_samples, learn_frames;
  const float* ring_before;    // [Y][N] ring before the push, oldest -> newest
  TrackState* state;
  TrackResult* result;
};

__device__ __forceinline__ long long track_frame_time(long long t0, double period, long long k) {
  return t0 + static_cast<long long>(floor(__dadd_rn(__dmul_rn(static_cast<double>(k), period), 0.5)));  // host::frame_time
}
__device__ __forceinline__ int track_index_to_shift(const TrackParams& p, int i) {  // sdr_device.cpp:150,154
  const double step = __ddiv_rn(static_cast<double>(p.sample_rate), static_cast<double>(p.n));
  return static_cast<int>(__dmul_rn(step, __dadd_rn(static_cast<double>(i), 0.5))) - p.sample_rate / 2;
}
__device__ __forceinline__ bool track_candidate_bin(const TrackParams& p, int i) {  // isIndexInRange && !isIndexIgnored, transmission.cpp:91,156-164
  const int f = p.center + track_index_to_shift(p, i);
  if (f < p.range_lo || f > p.range_hi) return false;
  for (int r = 0; r < p.n_ignored; ++r) {
    if (p.ignored_lo[r] <= f && f <= p.ignored_hi[r]) return false;
  }
  return true;
}
__device__ __forceinline__ int track_tuned(int f, int step) {  // getTunedFrequency, radio_utils.cpp:86-96
  int r = f % step;
  if (f < 0) r += step;
  const int below = f - r;
  return (r < step - r) ? below : below + step;
}
__device__ __forceinline__ int track_margin(int g) { return (g % 2 == 0) ? g / 2 : g / 2 + 1; }  // collection_utils.h:17-27

// first index with keys[i] >= v
__device__ __forceinline__ int track_lower_bound(const int* keys, int n, int v) {
  int a = 0, b = n;
  while (a < b) {
    const int m = (a + b) >> 1;
    if (keys[m] < v) a = m + 1; else b = m;
  }
  return a;
}
__device__ __forceinline__ bool track_within_margin(const int* keys, int n, int index, int margin) {
  const int i = track_lower_bound(keys, n, index - margin);
  return i < n && keys[i] <= index + margin;
}

// ------------------------------------------------------------------------------------------------------------
// k_runs: one warp per frame folds the frame's entries (ascending bins) into runs of consecutive bins per level
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_runs(const DetectEntry* entries, const int* offsets, int n_frames, TrackParams p, FrameRuns* runs) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= n_frames) return;
  const int e0 = offsets[t], e1 = offsets[t + 1];
  FrameRuns& fr = runs[t];
  int count[2] = {0, 0};
  for (int base = e0; base < e1; base += 32) {
    const int e = base + lane;
    const bool valid = e < e1;
    const DetectEntry cur = valid ? entries[e] : DetectEntry{-10, 0.0f};
    const DetectEntry prev = (valid && e > e0) ? entries[e - 1] : DetectEntry{-10, 0.0f};
    const DetectEntry next = (valid && e + 1 < e1) ? entries[e + 1] : DetectEntry{-10, 0.0f};
#pragma unroll
    for (int L = 0; L < 2; ++L) {
      auto in = [&](const DetectEntry& d) { return d.bin >= 0 && (L == 0 ? p.stop_level <= d.value : (p.start_level <= d.value && track_candidate_bin(p, d.bin))); };
      const bool me = valid && in(cur);
      const bool starts = me && !(prev.bin == cur.bin - 1 && in(prev));
      const bool ends = me && !(next.bin == cur.bin + 1 && in(next));
      const unsigned sm = __ballot_sync(0xffffffffu, starts), em = __ballot_sync(0xffffffffu, ends);
      const unsigned below = (1u << lane) - 1u;
      if (starts) {
        const int r = count[L] + __popc(sm & below);
        if (r < kRunCap) fr.lo[L][r] = cur.bin;
      }
      if (ends) {  // the run that ends here started at or before this lane: its index is (#starts up to and including me) - 1
        const int r = count[L] + __popc(sm & (below | (1u << lane))) - 1;
        if (r < kRunCap) fr.hi[L][r] = cur.bin;
      }
      count[L] += __popc(sm);
      (void)em;
    }
  }
  if (lane == 0) {
    fr.count[0] = count[0];
    fr.count[1] = count[1];
  }
}

// ------------------------------------------------------------------------------------------------------------
// k_track
// ------------------------------------------------------------------------------------------------------------
struct TrackShared {
  int n;                            // live signals
  int key[kMaxSignals];
  long long first[kMaxSignals], last[kMaxSignals];
  unsigned int hit[kKeyChunk][kTrackWords];  // bit f of word w: frame (block start + 32 w + f) has a stop-level bin in the key's window
  int event_frame;                  // first event frame found in this evaluation (or INT_MAX)
  int error;
  // event frame scratch
  int n_cand;
  int cand_bin[kMaxCand];
  float cand_val[kMaxCand];
  int cand_order[kMaxCand];
  float tx_power[kMaxSignals];
  int tx_order[kMaxSignals];
};

// getBestIndex (transmission.cpp:132-154) for candidate bin `index` at in-push frame `frame`, by one warp
__device__ int track_best_index(const TrackArgs& a, int index, int frame, int lane) {
  const TrackParams& p = a.p;
  const int total = p.group_y, rows = total - total / 2;  // rows [total/2, total) of the ring = the newest `rows` frames
  const int lo = max(0, index - p.group_size / 2), hi = min(p.n - 1, index + p.group_size / 2);
  constexpr int kMaxVotes = 128;  // rows = Y - Y/2 <= 128 (grouping_y <= 256 is validated on the host)
  int votes[kMaxVotes];            // identical in every lane of the warp
  int n_votes = 0;
  for (int r = 0; r < rows; ++r) {
    const int f = frame - rows + 1 + r;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int b = lo + lane; b <= hi; b += 32) {
      float v;
      if (f >= 0) {
        v = (a.noise_samples + f < a.learn_frames) ? kNoData : __fsub_rn(a.psd[static_cast<size_t>(f) * p.n + b], a.threshold[b]);
      } else if (total + f >= 0) {
        v = a.ring_before[static_cast<size_t>(total + f) * p.n + b];
      } else {
        v = 0.0f;
      }
      argmax_combine(bv, bi, v, b);  // first maximum: ties keep the lower bin (getMaxIndex, collection_utils.h:9-14)
    }
    warp_argmax(bv, bi);
    if (p.start_level <= bv && n_votes < kMaxVotes) votes[n_votes++] = bi;
  }
  if (n_votes == 0) return index;  // the reference indexes an empty vector here (collection_utils.h:46-49); defined as "keep the candidate"
  // mostFrequentValue (collection_utils.h:30-50): the mode; among equally frequent values the one at position size/2 of the
  // ascending tied set
  for (int i = 1; i < n_votes; ++i) {  // insertion sort, ascending
    const int v = votes[i];
    int j = i - 1;
    while (j >= 0 && votes[j] > v) {
      votes[j + 1] = votes[j];
      --j;
    }
    votes[j + 1] = v;
  }
  int best = 0, n_tied = 0;
  int tied[kMaxVotes];
  for (int i = 0; i < n_votes;) {
    int j = i;
    while (j < n_votes && votes[j] == votes[i]) ++j;
    const int run = j - i;
    if (run > best) {
      best = run;
      n_tied = 0;
    }
    if (run == best) tied[n_tied++] = votes[i];
    i = j;
  }
  return tied[n_tied / 2];
}

__global__ void __launch_bounds__(kTrackThreads) k_track(const TrackArgs a) {
  extern __shared__ __align__(16) unsigned char track_smem[];
  TrackShared& s = *reinterpret_cast<TrackShared*>(track_smem);
  const TrackParams& p = a.p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = a.n_frames;
  const int gh = p.group_size / 2, margin = track_margin(p.group_size);

  // ---- load the map ----
  if (tid == 0) {
    s.n = min(a.state->n, kMaxSignals);
    s.error = a.state->error;
  }
  __syncthreads();
  for (int i = tid; i < s.n; i += kTrackThreads) {
    s.key[i] = a.state->key[i];
    s.first[i] = a.state->first[i];
    s.last[i] = a.state->last[i];
  }
  __syncthreads();

  int ts = 0;  // first frame not yet applied (uniform)
  while (ts < T) {
    const int bs = ts & ~(kTrackFrames - 1);              // block of frames [bs, bs + 1024) holding ts
    const int be = min(T, bs + kTrackFrames);
    const int t = bs + tid;                               // my frame
    const bool mine = t >= ts && t < be;
    const int K = s.n;
    if (tid == 0) s.event_frame = 0x7fffffff;
    __syncthreads();
    const FrameRuns* fr = mine ? a.runs + t : nullptr;
    const int n_stop = mine ? fr->count[0] : 0, n_start = mine ? fr->count[1] : 0;
    const long long now = track_frame_time(a.t0_ms, a.period_ms, a.frame_offset + t);
    bool event = false;
    if (mine) {
      if (n_stop > kRunCap || n_start > kRunCap) event = true;  // complex frame: replayed from its raw entries
      // addSignals can fire when a candidate bin lies outside every key's margin interval (containsWithMargin)
      for (int r = 0; r < min(n_start, kRunCap) && !event; ++r) {
        int pos = fr->lo[1][r];
        const int end = fr->hi[1][r];
        int i = track_lower_bound(s.key, K, pos - margin);
        while (pos <= end) {
          if (i >= K || s.key[i] - margin > pos) {
            event = true;
            break;
          }
          pos = s.key[i] + margin + 1;
          ++i;
        }
      }
    }
    // ---- pass 1: hit bits per key chunk, then the first time-out of every key ----
    for (int c0 = 0; c0 < K || c0 == 0; c0 += kKeyChunk) {
      const int kc = min(kKeyChunk, K - c0);
      for (int q = 0; q < kc; ++q) {
        const int key = s.key[c0 + q];
        const int lo = max(0, key - gh), hi = min(p.n - 1, key + gh);
        bool h = false;
        for (int r = 0; r < min(n_stop, kRunCap); ++r) h = h || (fr->lo[0][r] <= hi && fr->hi[0][r] >= lo);
        const unsigned m = __ballot_sync(0xffffffffu, mine && h);
        if (lane == 0) s.hit[q][warp] = m;
      }
      __syncthreads();
      if (mine && !event) {
        for (int q = 0; q < kc; ++q) {
          // time of the key's latest hit at or before my frame (Signal::newData sets m_lastDataTime = now on such frames)
          long long last = s.last[c0 + q];
          unsigned m = s.hit[q][warp] & (0xffffffffu >> (31 - lane));
          int w = warp;
          while (m == 0u && w > 0) m = s.hit[q][--w];
          if (m != 0u) last = track_frame_time(a.t0_ms, a.period_ms, a.frame_offset + bs + 32 * w + (31 - __clz(m)));
          if (last + p.timeout <= now || s.first[c0 + q] + p.max_time <= now) {  // isTimeout / isMaximalTime, signal.cpp:28-30
            event = true;
            break;
          }
        }
      }
      if (c0 + kKeyChunk < K) __syncthreads();  // the hit words are rewritten by the next chunk
      if (K == 0) break;
    }
    if (event) atomicMin(&s.event_frame, t);
    __syncthreads();
    const int te = min(s.event_frame, be);  // frames [ts, te) are steady
    // ---- pass 2: commit the steady frames: every key's m_lastDataTime ----
    for (int c0 = 0; c0 < K; c0 += kKeyChunk) {
      const int kc = min(kKeyChunk, K - c0);
      if (K > kKeyChunk) {  // several chunks: the words of this chunk have to be rebuilt
        __syncthreads();
        for (int q = 0; q < kc; ++q) {
          const int key = s.key[c0 + q];
          const int lo = max(0, key - gh), hi = min(p.n - 1, key + gh);
          bool h = false;
          for (int r = 0; r < min(n_stop, kRunCap); ++r) h = h || (fr->lo[0][r] <= hi && fr->hi[0][r] >= lo);
          const unsigned m = __ballot_sync(0xffffffffu, mine && h);
          if (lane == 0) s.hit[q][warp] = m;
        }
        __syncthreads();
      }
      if (tid < kc && te > ts) {
        const int last_f = te - 1 - bs;  // newest steady frame, relative to the block
        int w = last_f >> 5;
        unsigned m = s.hit[tid][w] & (0xffffffffu >> (31 - (last_f & 31)));
        while (m == 0u && w > 0) m = s.hit[tid][--w];
        if (m != 0u) s.last[c0 + tid] = track_frame_time(a.t0_ms, a.period_ms, a.frame_offset + bs + 32 * w + (31 - __clz(m)));
      }
    }
    __syncthreads();
    if (te >= be) {
      ts = be;
      continue;
    }
    // ---- the event frame te, exactly as Transmission::process orders it (transmission.cpp:57-68) ----
    const long long ev_now = track_frame_time(a.t0_ms, a.period_ms, a.frame_offset + te);
    const int e0 = a.offsets[te], e1 = a.offsets[te + 1];
    if (tid == 0) s.n_cand = 0;
    __syncthreads();
    // addSignals: candidates = start-level bins in range and not ignored, strongest first (transmission.cpp:88-96)
    for (int e = e0 + tid; e < e1; e += kTrackThreads) {
      const DetectEntry d = a.entries[e];
      if (p.start_level <= d.value && track_candidate_bin(p, d.bin)) {
        const int i = atomicAdd(&s.n_cand, 1);
        if (i < kMaxCand) {
          s.cand_bin[i] = d.bin;
          s.cand_val[i] = d.value;
        }
      }
    }
    __syncthreads();
    if (s.n_cand > kMaxCand && tid == 0) s.error |= 2;
    const int nc = min(s.n_cand, kMaxCand);
    for (int i = tid; i < nc; i += kTrackThreads) {  // rank sort: (value desc, bin asc); bins are distinct
      const float v = s.cand_val[i];
      const int b = s.cand_bin[i];
      int rank = 0;
      for (int k = 0; k < nc; ++k) rank += (s.cand_val[k] > v || (s.cand_val[k] == v && s.cand_bin[k] < b)) ? 1 : 0;
      s.cand_order[rank] = i;
    }
    __syncthreads();
    if (warp == 0) {
      for (int c = 0; c < nc; ++c) {
        const int idx = s.cand_bin[s.cand_order[c]];
        if (track_within_margin(s.key, s.n, idx, margin)) continue;  // containsWithMargin, transmission.cpp:99
        const int key = track_best_index(a, idx, te, lane);
        if (lane == 0) {
          const int pos = track_lower_bound(s.key, s.n, key);
          if (!(pos < s.n && s.key[pos] == key)) {  // std::map::insert keeps an existing element
            if (s.n >= kMaxSignals) {
              s.error |= 1;
            } else {
              for (int i = s.n; i > pos; --i) {
                s.key[i] = s.key[i - 1];
                s.first[i] = s.first[i - 1];
                s.last[i] = s.last[i - 1];
              }
              s.key[pos] = key;
              s.first[pos] = ev_now;  // Signal(now): m_firstDataTime = m_lastDataTime = now (signal.cpp:6-14)
              s.last[pos] = ev_now;
              s.n += 1;
            }
          }
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // updateSignals: a stop-level bin inside the key's window refreshes m_lastDataTime (transmission.cpp:113-130, signal.cpp:16-24)
    for (int i = tid; i < s.n; i += kTrackThreads) {
      const int lo = max(0, s.key[i] - gh), hi = min(p.n - 1, s.key[i] + gh);
      int x = e0, y = e1;
      while (x < y) {  // first entry of the frame with bin >= lo
        const int m = (x + y) >> 1;
        if (a.entries[m].bin < lo) x = m + 1; else y = m;
      }
      bool h = false;
      for (int e = x; e < e1 && a.entries[e].bin <= hi; ++e) h = h || (p.stop_level <= a.entries[e].value);
      if (h) s.last[i] = ev_now;
    }
    __syncthreads();
    // clearSignals (transmission.cpp:70-86)
    if (tid == 0) {
      int w = 0;
      for (int i = 0; i < s.n; ++i) {
        if (s.last[i] + p.timeout <= ev_now || s.first[i] + p.max_time <= ev_now) continue;
        s.key[w] = s.key[i];
        s.first[w] = s.first[i];
        s.last[w] = s.last[i];
        ++w;
      }
      s.n = w;
    }
    __syncthreads();
    ts = te + 1;
  }

  // ---- after the last frame: m_power, getSortedTransmissions, state back to global memory ----
  const int K = s.n;
  const long long last_now = T > 0 ? track_frame_time(a.t0_ms, a.period_ms, a.frame_offset + T - 1) : a.result->last_now;
  for (int i = tid; i < K; i += kTrackThreads) {
    float pw = a.state->power[i];  // (overwritten below when the push had frames)
    if (T > 0) {
      const int lo = max(0, s.key[i] - gh), hi = min(p.n - 1, s.key[i] + gh);
      pw = a.box_last[lo];
      for (int b = lo + 1; b <= hi; ++b) pw = fmaxf(pw, a.box_last[b]);  // getMaxIndex(avgPower, ...): the window maximum
    }
    s.tx_power[i] = pw;
  }
  __syncthreads();
  for (int i = tid; i < K; i += kTrackThreads) {  // power descending, equal powers by ascending key
    const float v = s.tx_power[i];
    int rank = 0;
    for (int k = 0; k < K; ++k) rank += (s.tx_power[k] > v || (s.tx_power[k] == v && k < i)) ? 1 : 0;
    s.tx_order[rank] = i;
  }
  __syncthreads();
  for (int r = tid; r < K; r += kTrackThreads) {
    const int i = s.tx_order[r];
    b2s_transmission tx;
    tx.shift_hz = track_tuned(track_index_to_shift(p, s.key[i]), p.tuning_step);
    tx.flush = (s.last[i] == last_now && s.first[i] + p.min_time <= last_now) ? 1 : 0;  // Signal::needFlush, signal.cpp:26,32
    tx.key = s.key[i];
    tx.power = s.tx_power[i];
    a.result->tx[r] = tx;
    a.state->key[i] = s.key[i];
    a.state->first[i] = s.first[i];
    a.state->last[i] = s.last[i];
    a.state->power[i] = s.tx_power[i];
  }
  if (tid == 0) {
    a.state->n = K;
    a.state->error = s.error;
    a.result->n_tx = K;
    a.result->n_entries = T > 0 ? a.offsets[T] : 0;
    a.result->max_count = a.max_count ? *a.max_count : 0;
    a.result->error = s.error;
    a.result->last_now = last_now;
  }
}

}  // namespace b2s
