// K2/K3 — the time-ordered half of the hot path, one thread per FFT bin marching through the frames of a push:
//   NoiseLearner (learn max / subtract)        reference sources/radio/blocks/noise_learner.cpp:11-28,36-67
//   Averager (ring + running sum, bit-exact)   reference sources/radio/averager.cpp:14-25,40-60
//   average(avg, GROUPING_X) frequency boxcar  reference sources/utils/utils.cpp:31-53
//   threshold predicate -> detection entries   reference sources/radio/blocks/transmission.cpp:88-96,113-130
//   Spectrogram decimate + accumulate + send   reference sources/radio/blocks/spectrogram.cpp:45-75
// The signal-map bookkeeping (transmission.cpp:70-176) consumes the compact detection entries on the host (tracker.h).
//
// Arithmetic contract: every float operation that feeds Averager state is the reference's operation in the reference's
// order (separate sub / add / IEEE division, no FMA contraction) so m_sum, the ring and m_average are bit-identical.
// The frequency boxcar is evaluated per aligned segment of 8 bins: the first bin's window is summed left to right, the
// next 7 slide it exactly like the reference does (sum -= leaving; sum += entering), see boxcar_segment(). The
// reference carries ONE running sum across the whole row (a 16384-long serial float chain); restarting it every
// 8 bins changes rounding only (<= 1e-3 dB vs the reference's own drift, asserted in tests) and keeps the work
// parallel. b2s_average(..., exact=1) provides the serial form for operator-level bit parity.
#pragma once
#include <cstdio>
#include <cstring>

#include "b2s_device.cuh"

namespace b2s {

constexpr int kDetectBinsPerCta = 128;  // bins owned by one CTA (also the largest spectrogram decimation supported)
constexpr int kDetectTileFrames = 32;   // frames per shared-memory tile
constexpr int kDetectThreads = 192 + 512;  // 6 march warps + 16 box warps (one per 8-bin segment)
constexpr int kDetectBuffers = 3;       // PSD tiles resident: current, previous (ring look-back), next (in flight)
constexpr int kMaxSpecEmits = 16;       // spectrogram rows that one push (chunk) may complete
constexpr int kMaxWatch = 16;           // live signal keys whose window maxima K2 reports directly
constexpr int kCheckpointEvery = 64;    // frames between Averager-sum checkpoints (replay points for K3)

struct DetectEntry {  // one bin whose boxcar power reached min(start, stop) in one frame
  int bin;
  float value;  // boxcar-averaged power (dB above learned noise)
};

struct DetectArgs {
  // geometry
  int n;         // N
  int n_frames;  // T
  int group_y;   // Averager depth Y
  int group_x;   // boxcar width X
  // inputs
  const float* psd;  // [T][N] raw PSD rows from K1
  // noise state (per centre frequency)
  float* threshold;   // [N], updated in place while learning
  int noise_samples;  // samples learned before this push
  int learn_frames;   // frames 0.. with noise_samples + t < learn_frames are learning frames
  // averager state
  float* avg_sum;         // [N] m_sum, updated in place
  const float* ring_in;   // [Y][N] ring before the push, oldest -> newest
  float* ring_out;        // [Y][N] ring after the push (a different buffer)
  int avg_frames;         // m_frames before the push
  float* avg_last;        // [N] m_average after the last frame
  float* checkpoints;     // [ceil(T/64)][N] m_sum before frame 64*c
  // detection: per-frame slot lists
  float detect_level;     // min(start, stop)
  DetectEntry* slots;     // [T][slot_capacity]
  int* slot_count;        // [T] (zeroed before launch); may exceed slot_capacity -> overflow, reported by the host
  int slot_capacity;
  // spectrogram
  int spec_out;           // M (0 = off)
  float* spec_sum;        // [M]
  // watch list: the signal-map keys known when the push was enqueued. For each, K2 reports the maximum of the boxcar row
  // over [key - g/2, key + g/2] per frame (getMaxIndex(avgPower, N, key, groupSize) of Transmission::updateSignals,
  // transmission.cpp:114-117) and, per frame, whether any bin at or above the start level lies outside every key's
  // containsWithMargin interval (collection_utils.h:17-27) — i.e. whether addSignals could create a signal at all.
  int n_watch;
  int watch_key[kMaxWatch];
  int group_size;            // m_groupSize in bins
  float start_level;
  unsigned int* watch_max;   // [T][kMaxWatch] order-preserving encoding of the float maximum (0 = nothing written)
  int* cand_flag;            // [T] set to 1 when an uncovered candidate exists
  // rows to emit during this push, planned by the host from the frame clock (Spectrogram::send, spectrogram.cpp:62-75).
  // Carried in the kernel arguments so that no small host->device copy sits on the critical path behind the bulk IQ copy.
  int n_emit;                      // <= kMaxSpecEmits
  int emit_frame[kMaxSpecEmits];   // frame after which row i is emitted (ascending)
  int emit_div[kMaxSpecEmits];     // Container::m_counter at that moment
  signed char* spec_rows;          // [n_emit][M]
  // optional dense rows [T][N]
  float* dense_q;
  float* dense_avg;
  float* dense_box;
};

// noise-subtracted power of in-push frame t for bin j (NoiseLearner output), recomputable anywhere from the PSD rows
__device__ __forceinline__ float noise_sub(float p, float thr, bool learning) { return learning ? kNoData : __fsub_rn(p, thr); }

// One Averager::push for one bin: subtract the leaving value, add the new one (two separate roundings, in this order),
// then m_average = m_sum / groupSize (IEEE division) once groupSize frames were seen — averager.cpp:14-25,40-60.
__device__ __forceinline__ float averager_step(float& sum, float leaving, float entering, int frames_after, int group) {
  sum = __fsub_rn(sum, leaving);
  sum = __fadd_rn(sum, entering);
  return (frames_after >= group) ? __fdiv_rn(sum, static_cast<float>(group)) : kNoData;
}

// x / D for a small integer constant D, bit-identical to IEEE division: Markstein's sequence q = x*r, e = fma(-q, D, x),
// q' = fma(e, r, q) with r = RN(1/D). Verified exhaustively against x / D for every float with |x| in [2^-60, 2^60] and
// for +0 (D = 1..21); anything outside that range takes the IEEE path. Three FMA-pipe instructions
// instead of the ~10 instruction + subroutine-call sequence the compiler emits for a correctly rounded division.
template <int D>
__device__ __forceinline__ float div_const(float x) {
  constexpr float d = static_cast<float>(D);
  constexpr float r = 1.0f / d;
  const uint32_t bits = __float_as_uint(x);
  const uint32_t ex = (bits >> 23) & 0xffu;
  if ((ex - 67u) <= 120u || bits == 0u) {
    const float q = __fmul_rn(x, r);
    const float e = __fmaf_rn(-q, d, x);
    return __fmaf_rn(e, r, q);
  }
  return __fdiv_rn(x, d);
}

// Same, without the range guard: for callers whose operands are sums of finite dB values (|x| < 2^60 by construction).
template <int D>
__device__ __forceinline__ float div_const_fast(float x) {
  constexpr float d = static_cast<float>(D);
  constexpr float r = 1.0f / d;
  const float q = __fmul_rn(x, r);
  const float e = __fmaf_rn(-q, d, x);
  return __fmaf_rn(e, r, q);
}

// order-preserving map float -> unsigned (so atomicMax on the image is max on the floats); never yields 0 for a real number
__host__ __device__ __forceinline__ unsigned int float_to_ordered(float f) {
#ifdef __CUDA_ARCH__
  const unsigned int b = __float_as_uint(f);
#else
  unsigned int b;
  memcpy(&b, &f, 4);
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
inline float ordered_to_float(unsigned int u) {
  const unsigned int b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  float f;
  memcpy(&f, &b, 4);
  return f;
}

constexpr int kBoxSegment = 8;  // bins per boxcar segment (segments are aligned to multiples of 8 bins)

// Boxcar running sums of one aligned segment of 8 bins over the ZERO-EXTENDED row: w[i] holds the averaged value of bin
// (b0 - H + i), i in [0, 8 + 2H), with 0.0f wherever that bin lies outside [0, N) (adding or subtracting 0.0f is exact,
// so clipped windows come out as the left-to-right sum of their valid bins). sums[0] = w[0] + w[1] + ... + w[2H];
// sums[k] continues the running sum exactly like utils.cpp:41-48: drop the leaving element, then add the entering one.
template <int H>
__device__ __forceinline__ void boxcar_segment(const float (&w)[kBoxSegment + 2 * H], float (&sums)[kBoxSegment]) {
  float s = w[0];
#pragma unroll
  for (int i = 1; i <= 2 * H; ++i) s = __fadd_rn(s, w[i]);
  sums[0] = s;
#pragma unroll
  for (int k = 1; k < kBoxSegment; ++k) {
    s = __fsub_rn(s, w[k - 1]);
    s = __fadd_rn(s, w[k + 2 * H]);
    sums[k] = s;
  }
}
// number of valid bins in the window of bin j (the reference's `count`, utils.cpp:34-49)
__device__ __forceinline__ int boxcar_count(int j, int n, int half) { return min(n - 1, j + half) - max(0, j - half) + 1; }
// a segment is interior when none of its 8 windows is clipped by the row ends (then every count is 2*half + 1)
__device__ __forceinline__ bool segment_interior(int b0, int n, int half) { return b0 - half >= 0 && b0 + kBoxSegment - 1 + half < n; }

// Boxcar value of ONE bin under the same definition, for any half; `at(bin)` returns the averaged value of a valid bin.
template <typename At>
__device__ __forceinline__ float boxcar_value(At at, int j, int n, int half) {
  if (half == 0 && j == n - 1) return 0.0f;  // reference quirk: groupSize 1 never writes the last element (utils.cpp:38)
  const int b0 = j & ~(kBoxSegment - 1);
  auto z = [&](int bin) { return (bin >= 0 && bin < n) ? at(bin) : 0.0f; };
  float s = z(b0 - half);
  for (int i = 1; i <= 2 * half; ++i) s = __fadd_rn(s, z(b0 - half + i));
  for (int k = 1; k <= j - b0; ++k) {
    s = __fsub_rn(s, z(b0 - half + k - 1));
    s = __fadd_rn(s, z(b0 + half + k));
  }
  return __fdiv_rn(s, static_cast<float>(boxcar_count(j, n, half)));
}

// 16-byte asynchronous global->shared copy (LDGSTS); used to stream PSD tiles ahead of the march
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One CTA owns kDetectBinsPerCta bins plus a halo of X/2 bins (rounded up to 4) on each side, computed redundantly.
// PSD tiles of kDetectTileFrames rows are streamed into shared memory with cp.async two tiles ahead; one thread per
// column marches the tile through noise -> Averager (the only serial chain: two dependent FADDs per frame); then all
// threads evaluate boxcar + threshold for the tile's (frame, bin) grid.
// named barriers (id 0 is __syncthreads)
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

constexpr int kMarchThreads = 192;                              // warps 0-5: one thread per column (<= 192 columns)
constexpr int kBoxThreads = kDetectThreads - kMarchThreads;     // warps 6-15: boxcar + threshold + emission
constexpr int kBarMarch = 1, kBarBox = 2, kBarFull = 3 /*,4*/, kBarEmpty = 5 /*,6*/;

// Y_T / HALF_T: Averager depth and X/2 as compile-time constants (21 / 10 = the reference's GROUPING_Y / GROUPING_X),
// or 0 / -1 for the generic runtime-parameter instantiation.
//
// Warp-specialised, software-pipelined over tiles of 32 frames:
//   MARCH warps (one thread per column): stream the PSD tile in with cp.async, NoiseLearner subtraction, the serial
//     Averager chain (m_sum -= leaving; m_sum += entering; m_average = m_sum / Y), spectrogram accumulation; they write the
//     tile of averaged values into one of two shared buffers and move on to the next tile.
//   BOX warps: boxcar over 8-bin segments + threshold + watched-window maxima for the tile the march warps finished one
//     step earlier; detection entries are staged in shared memory and flushed with one global atomic per (CTA, frame).
// The two groups meet only through FULL/EMPTY named barriers on the double-buffered average tile, so the serial chain of
// tile i+1 overlaps the throughput work of tile i. Steady-state tiles (no learning frame, ring look-back inside the push,
// full tile, no dense debug rows) take a branch-free register-resident march; all others a generic per-column march with
// the same float operations in the same order (bit-identical, tested).
template <int Y_T, int HALF_T>
__global__ void __launch_bounds__(kDetectThreads) k_detect(const DetectArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int half = HALF_T >= 0 ? HALF_T : a.group_x / 2;
  const int hp = (half + 3) & ~3;                   // halo padded to a 16-byte multiple
  const int width = kDetectBinsPerCta + 2 * hp;     // columns held by this CTA (<= kMarchThreads)
  const int tile_elems = kDetectTileFrames * width;
  float* psd_tiles = sm;                                            // [kDetectBuffers][TF][width] raw PSD (cp.async target)
  float* q_tiles = psd_tiles + kDetectBuffers * tile_elems;         // [2][TF][width] noise-subtracted rows (current, previous)
  // averaged values handed to the box warps, TRANSPOSED: [2][width][kAvgPitch] (column-major, pitch 33). The march thread of
  // column c writes avg[c*33 + f] (lane stride 33: conflict-free); a box warp reads one column for 32 frames at once
  // (lane = frame: consecutive words, conflict-free).
  constexpr int kAvgPitch = kDetectTileFrames + 1;
  const int avg_elems = width * kAvgPitch;
  float* avg_tiles = q_tiles + 2 * tile_elems;
  int* stage_count = reinterpret_cast<int*>(avg_tiles + 2 * avg_elems);  // [TF] detection entries staged per frame of the tile
  int* stage_base = stage_count + kDetectTileFrames;                // [TF] where this CTA's block starts in the frame's slot list
  DetectEntry* stage = reinterpret_cast<DetectEntry*>(stage_base + kDetectTileFrames);  // [TF][kDetectBinsPerCta]
  __shared__ int rel_n, rel_key[kMaxWatch], rel_slot[kMaxWatch];    // watched keys that touch this CTA's bins

  const int n = a.n, T = a.n_frames, Y = Y_T > 0 ? Y_T : a.group_y;
  const int j0 = blockIdx.x * kDetectBinsPerCta;
  const int col0 = j0 - hp;  // bin of column 0
  const int tid = threadIdx.x;
  const int n_tiles = (T + kDetectTileFrames - 1) / kDetectTileFrames;
  const bool dense = a.dense_q || a.dense_avg || a.dense_box;

  // ---- one-time setup (all threads) ----
  if (tid == 0) {
    int cnt = 0;
    const int reach = a.group_size / 2 + 1;
    for (int w = 0; w < a.n_watch; ++w) {
      if (a.watch_key[w] + reach >= j0 && a.watch_key[w] - reach < j0 + kDetectBinsPerCta) {
        rel_key[cnt] = a.watch_key[w];
        rel_slot[cnt] = w;
        ++cnt;
      }
    }
    rel_n = cnt;
  }
  if (tid < kDetectTileFrames) stage_count[tid] = 0;
  if (tid < width && (col0 + tid < 0 || col0 + tid >= n)) {  // columns outside the row: the boxcar sees the zero-extended row
    for (int f = 0; f < kAvgPitch; ++f) {
      avg_tiles[tid * kAvgPitch + f] = 0.0f;
      avg_tiles[avg_elems + tid * kAvgPitch + f] = 0.0f;
    }
  }
  __syncthreads();

  if (tid < kMarchThreads) {
    // =========================================== MARCH warps ===========================================
    const int j = col0 + tid;  // my column's bin
    const bool active = tid < width && j >= 0 && j < n;
    const bool owner = active && tid >= hp && tid < hp + kDetectBinsPerCta;
    const int chunks_per_row = width / 4;
    auto issue_tile = [&](int tile) {
      if (tile < n_tiles) {
        float* dst = psd_tiles + (tile % kDetectBuffers) * tile_elems;
        const int t_base = tile * kDetectTileFrames;
        for (int c = tid; c < kDetectTileFrames * chunks_per_row; c += kMarchThreads) {
          const int f = c / chunks_per_row, x = (c - f * chunks_per_row) * 4;
          const int t = t_base + f, col = col0 + x;
          if (t < T && col >= 0 && col + 3 < n) cp_async16(dst + f * width + x, a.psd + static_cast<size_t>(t) * n + col);
        }
      }
      cp_async_commit();
    };
    float thr = active ? a.threshold[j] : 0.0f;
    float sum = active ? a.avg_sum[j] : 0.0f;
    float last_avg = kNoData;
    const int d = a.spec_out > 0 ? n / a.spec_out : 0;
    const bool spec_owner = owner && d > 0 && (j % d) == 0;
    float spec = spec_owner ? a.spec_sum[j / d] : 0.0f;
    const bool ring_in_smem = Y <= kDetectTileFrames;
    int next_emit = 0;  // index of the first planned spectrogram row not yet emitted (rows are in frame order)

    issue_tile(0);
    issue_tile(1);
    for (int tile = 0; tile < n_tiles; ++tile) {
      const int t0 = tile * kDetectTileFrames;
      const int tf = min(kDetectTileFrames, T - t0);
      cp_async_wait<1>();                     // my part of tile `tile` has landed (tile+1 may still be in flight)
      bar_sync(kBarMarch, kMarchThreads);     // ... and everybody else's part; also: tile-1 is fully consumed
      if (tile >= 2) bar_sync(kBarEmpty + (tile & 1), kDetectThreads);  // the box warps are done with this average buffer
      const float* __restrict__ cur = psd_tiles + (tile % kDetectBuffers) * tile_elems;
      const float* __restrict__ prev = psd_tiles + ((tile + kDetectBuffers - 1) % kDetectBuffers) * tile_elems;
      float* __restrict__ q_cur = q_tiles + (tile & 1) * tile_elems;
      const float* __restrict__ q_prev = q_tiles + ((tile & 1) ^ 1) * tile_elems;
      float* __restrict__ avg_col = avg_tiles + (tile & 1) * avg_elems + tid * kAvgPitch;  // my column of the transposed tile
      // planned spectrogram row inside this tile (at most the first one is handled by the fast path)
      while (next_emit < a.n_emit && a.emit_frame[next_emit] < t0) ++next_emit;
      const int emit_f = (d > 0 && next_emit < a.n_emit && a.emit_frame[next_emit] < t0 + tf) ? a.emit_frame[next_emit] - t0 : -1;
      auto slot_of = [&](int f) {  // planned row emitted after frame t0 + f, or -1
        int slot = -1;
        for (int i = next_emit; i < a.n_emit && a.emit_frame[i] <= t0 + f; ++i) slot = (a.emit_frame[i] == t0 + f) ? i : slot;
        return slot;
      };
      // steady state: whole tile, ring look-back inside the push, no learning frame in the tile
      // (the look-back reads q_prev, which every earlier tile wrote — learning frames as -100 — whichever path it took)
      const bool steady = Y_T > 0 && HALF_T > 0 && tf == kDetectTileFrames && t0 >= kDetectTileFrames && Y <= kDetectTileFrames &&
                          (a.noise_samples + t0 >= a.learn_frames) && !dense;
      if (steady) {
        if (active) {
          if (owner && (t0 % kCheckpointEvery) == 0) a.checkpoints[static_cast<size_t>(t0 / kCheckpointEvery) * n + j] = sum;  // m_sum before frame t0
          constexpr int YC = Y_T > 0 ? Y_T : 1;
          constexpr int TF = kDetectTileFrames;
          const bool full = a.avg_frames + t0 + 1 >= YC;  // m_frames has reached groupSize
          float q[TF], lead[YC];
#pragma unroll
          for (int f = 0; f < TF; ++f) q[f] = cur[f * width + tid];
#pragma unroll
          for (int f = 0; f < YC; ++f) lead[f] = q_prev[(f - YC + TF) * width + tid];  // rows leaving the ring during the first Y frames
          // Spectrogram::process on the RAW rows (spectrogram.cpp:46-49)
          if (d == 1 && owner) {
            if (emit_f < 0) {
#pragma unroll
              for (int f = 0; f < TF; ++f) spec = __fadd_rn(spec, q[f]);
            } else {
              for (int f = 0; f < TF; ++f) {
                spec = __fadd_rn(spec, cur[f * width + tid]);
                const int slot = slot_of(f);
                if (slot >= 0) {  // Spectrogram::send, spectrogram.cpp:66-72: float -> int8 truncation, then clear
                  a.spec_rows[static_cast<size_t>(slot) * a.spec_out + j] = static_cast<signed char>(static_cast<int>(__fdiv_rn(spec, static_cast<float>(a.emit_div[slot]))));
                  spec = 0.0f;
                }
              }
            }
          }
#pragma unroll
          for (int f = 0; f < TF; ++f) q[f] = __fsub_rn(q[f], thr);  // NoiseLearner::work, noise_learner.cpp:54
#pragma unroll
          for (int f = 0; f < TF; ++f) {
            const float old = (f >= YC) ? q[f - YC] : lead[f];
            sum = __fsub_rn(sum, old);   // Averager::subtract, averager.cpp:46-50
            sum = __fadd_rn(sum, q[f]);  // Averager::add, averager.cpp:40-44
            const float avg = full ? div_const_fast<YC>(sum) : kNoData;
            avg_col[f] = avg;
            q_cur[f * width + tid] = q[f];  // the next tile looks back into this one
            if (f == TF - 1) last_avg = avg;
          }
        }
      } else if (active) {
        // ---- generic per-column march (learning frames, first tile of a push, partial tiles, dense debug rows) ----
        for (int f = 0; f < tf; ++f) {
          const int t = t0 + f;
          const float p = cur[f * width + tid];
          const bool learning = a.noise_samples + t < a.learn_frames;
          if (learning) thr = fmaxf(thr, p);  // Noise::add, noise_learner.cpp:19-21
          const float q = noise_sub(p, thr, learning);
          q_cur[f * width + tid] = q;  // a later steady tile looks back into this one
          // value leaving the ring (frame t - Y): thr is final for every frame that was not a learning frame
          float old;
          if (t >= Y) {
            float po;
            if (ring_in_smem) {
              po = (f >= Y) ? cur[(f - Y) * width + tid] : prev[(f - Y + kDetectTileFrames) * width + tid];
            } else {
              po = a.psd[static_cast<size_t>(t - Y) * n + j];
            }
            old = noise_sub(po, thr, a.noise_samples + (t - Y) < a.learn_frames);
          } else {
            old = a.ring_in[static_cast<size_t>(t) * n + j];  // the t-th oldest row of the pre-push ring
          }
          if (owner && (t % kCheckpointEvery) == 0) a.checkpoints[static_cast<size_t>(t / kCheckpointEvery) * n + j] = sum;  // m_sum before frame t
          const float avg = averager_step(sum, old, q, min(a.avg_frames + t + 1, Y), Y);
          avg_col[f] = avg;
          last_avg = avg;
          if (owner) {
            if (a.dense_q) a.dense_q[static_cast<size_t>(t) * n + j] = q;
            if (a.dense_avg) a.dense_avg[static_cast<size_t>(t) * n + j] = avg;
            if (d == 1) {
              spec = __fadd_rn(spec, p);  // Spectrogram::process, spectrogram.cpp:46-49
              const int slot = emit_f >= 0 ? slot_of(f) : -1;
              if (slot >= 0) {  // Spectrogram::send, spectrogram.cpp:66-72
                a.spec_rows[static_cast<size_t>(slot) * a.spec_out + j] = static_cast<signed char>(static_cast<int>(__fdiv_rn(spec, static_cast<float>(a.emit_div[slot]))));
                spec = 0.0f;
              }
            }
          }
        }
      }
      if (spec_owner && d > 1) {  // decimating spectrogram: mean of d adjacent raw bins, then accumulate (spectrogram.cpp:50-58)
        for (int f = 0; f < tf; ++f) {
          float s = 0.0f;
          for (int i = 0; i < d; ++i) s = __fadd_rn(s, cur[f * width + tid + i]);
          spec = __fadd_rn(spec, __fdiv_rn(s, static_cast<float>(d)));
          const int slot = emit_f >= 0 ? slot_of(f) : -1;
          if (slot >= 0) {
            a.spec_rows[static_cast<size_t>(slot) * a.spec_out + j / d] = static_cast<signed char>(static_cast<int>(__fdiv_rn(spec, static_cast<float>(a.emit_div[slot]))));
            spec = 0.0f;
          }
        }
      }
      __threadfence_block();
      bar_arrive(kBarFull + (tile & 1), kDetectThreads);  // hand the averaged tile to the box warps
      bar_sync(kBarMarch, kMarchThreads);                 // every march thread is done with tile-1's PSD buffer
      issue_tile(tile + 2);                               // ... which tile+2 reuses
    }
    cp_async_wait<0>();
    if (owner) {
      a.threshold[j] = thr;
      a.avg_sum[j] = sum;
      a.avg_last[j] = last_avg;
      // ring after the push, oldest -> newest: row i is in-push frame T - Y + i, or a surviving row of ring_in
      for (int i = 0; i < Y; ++i) {
        const int t = T - Y + i;
        float q;
        if (t >= 0) {
          q = noise_sub(a.psd[static_cast<size_t>(t) * n + j], thr, a.noise_samples + t < a.learn_frames);
        } else {
          q = a.ring_in[static_cast<size_t>(T + i) * n + j];
        }
        a.ring_out[static_cast<size_t>(i) * n + j] = q;
      }
    }
    if (spec_owner) a.spec_sum[j / d] = spec;
  } else {
    // ============================================ BOX warps ============================================
    // warp w owns the 8-bin segment w of the CTA's 128 bins; lane = frame of the tile
    const int btid = tid - kMarchThreads;
    const int lane = btid & 31, seg = btid >> 5;
    constexpr int SEG = kBoxSegment;
    static_assert(kBoxThreads / 32 == kDetectBinsPerCta / kBoxSegment && kDetectTileFrames == 32, "one box warp per segment, one lane per frame");
    const int b0 = seg * SEG, bin0 = j0 + b0;
    for (int tile = 0; tile < n_tiles; ++tile) {
      const int t0 = tile * kDetectTileFrames;
      const int tf = min(kDetectTileFrames, T - t0);
      bar_sync(kBarFull + (tile & 1), kDetectThreads);  // the march warps have written this average tile
      const float* avg_tile = avg_tiles + (tile & 1) * avg_elems;
      const int f = lane, t = t0 + f;
      if (f < tf && bin0 < n) {
        float box[SEG];
        if (HALF_T > 0) {
          constexpr int H = HALF_T > 0 ? HALF_T : 1;
          float w[SEG + 2 * H];
#pragma unroll
          for (int i = 0; i < SEG + 2 * H; ++i) w[i] = avg_tile[(hp + b0 - H + i) * kAvgPitch + f];  // columns outside [0, N) hold 0.0f
          boxcar_segment<H>(w, box);
          if (segment_interior(bin0, n, half)) {
#pragma unroll
            for (int k = 0; k < SEG; ++k) box[k] = div_const_fast<2 * H + 1>(box[k]);
          } else {
#pragma unroll
            for (int k = 0; k < SEG; ++k) box[k] = __fdiv_rn(box[k], static_cast<float>(boxcar_count(bin0 + k, n, half)));
          }
        } else {
#pragma unroll
          for (int k = 0; k < SEG; ++k) {
            const int bin = bin0 + k;
            box[k] = (bin >= n) ? -INFINITY : boxcar_value([&](int bb) { return avg_tile[(hp + (bb - j0)) * kAvgPitch + f]; }, bin, n, half);
          }
        }
        float top = box[0];
#pragma unroll
        for (int k = 1; k < SEG; ++k) top = fmaxf(top, box[k]);
        // watched keys: window maxima over ALL bins (also below the detection level), and the uncovered-candidate flag
        if (rel_n > 0 || top >= a.start_level) {
          const int gh = a.group_size / 2, margin = (a.group_size % 2 == 0) ? gh : gh + 1;
          unsigned int covered = 0;  // bit k: bin0 + k lies inside some key's containsWithMargin interval
          for (int r = 0; r < rel_n; ++r) {
            const int key = rel_key[r], w = rel_slot[r];
            if (bin0 + SEG - 1 >= key - gh && bin0 <= key + gh) {
              float m = -INFINITY;
#pragma unroll
              for (int k = 0; k < SEG; ++k) {
                const int bin = bin0 + k;
                if (bin < n && bin >= key - gh && bin <= key + gh) m = fmaxf(m, box[k]);
              }
              if (m > -INFINITY) atomicMax(a.watch_max + static_cast<size_t>(t) * kMaxWatch + w, float_to_ordered(m));
            }
            if (bin0 + SEG - 1 >= key - margin && bin0 <= key + margin) {
#pragma unroll
              for (int k = 0; k < SEG; ++k) covered |= (bin0 + k >= key - margin && bin0 + k <= key + margin) ? (1u << k) : 0u;
            }
          }
          if (top >= a.start_level) {
            bool uncovered = false;
#pragma unroll
            for (int k = 0; k < SEG; ++k) uncovered |= (bin0 + k < n) && box[k] >= a.start_level && !((covered >> k) & 1u);
            if (uncovered) a.cand_flag[t] = 1;
          }
        }
        if (a.dense_box || top >= a.detect_level) {
#pragma unroll
          for (int k = 0; k < SEG; ++k) {
            const int bin = bin0 + k;
            if (bin < n) {
              if (a.dense_box) a.dense_box[static_cast<size_t>(t) * n + bin] = box[k];
              if (box[k] >= a.detect_level) {
                const int pos = atomicAdd(stage_count + f, 1);  // shared-memory counter: at most 128 entries per frame per CTA
                stage[f * kDetectBinsPerCta + pos] = DetectEntry{bin, box[k]};
              }
            }
          }
        }
      }
      bar_arrive(kBarEmpty + (tile & 1), kDetectThreads);  // this average buffer may be overwritten (tile + 2)
      bar_sync(kBarBox, kBoxThreads);                      // all entries of the tile are staged
      // flush: ONE global atomic per (CTA, frame) reserves a block of the frame's slot list
      if (btid < tf) {
        const int cnt = stage_count[btid];
        stage_base[btid] = cnt > 0 ? atomicAdd(a.slot_count + t0 + btid, cnt) : 0;
      }
      bar_sync(kBarBox, kBoxThreads);
      for (int f = btid >> 5; f < tf; f += kBoxThreads >> 5) {  // one warp per frame
        const int cnt = stage_count[f], base = stage_base[f];
        for (int i = btid & 31; i < cnt; i += 32) {
          if (base + i < a.slot_capacity) a.slots[static_cast<size_t>(t0 + f) * a.slot_capacity + base + i] = stage[f * kDetectBinsPerCta + i];
        }
        __syncwarp();
        if ((btid & 31) == 0) stage_count[f] = 0;
      }
      bar_sync(kBarBox, kBoxThreads);  // staging area re-armed before the next tile's entries arrive
    }
  }
}

// Exclusive prefix of min(slot_count[t], capacity) over the T frames (one CTA), then per-frame ordering of the slot
// lists by bin into one dense array (one warp per frame, rank sort: bins are distinct inside a frame).
__global__ void __launch_bounds__(1024) k_entries_prefix(const int* slot_count, int capacity, int n_frames, int* offsets /*[T+1]*/, int* max_count) {
  __shared__ int part[1024];
  const int tid = threadIdx.x;
  const int per = (n_frames + 1023) / 1024;
  const int begin = tid * per, end = min(n_frames, begin + per);
  int local = 0, biggest = 0;
  for (int t = begin; t < end; ++t) {
    local += min(slot_count[t], capacity);
    biggest = max(biggest, slot_count[t]);
  }
  part[tid] = local;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int v = tid >= o ? part[tid - o] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  int run = part[tid] - local;
  for (int t = begin; t < end; ++t) {
    offsets[t] = run;
    run += min(slot_count[t], capacity);
  }
  if (tid == 1023) offsets[n_frames] = part[1023];
  atomicMax(max_count, biggest);
}

__global__ void __launch_bounds__(256) k_entries_sort(const DetectEntry* slots, const int* slot_count, int capacity, int n_frames, const int* offsets, DetectEntry* out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_frames) return;
  const int count = min(slot_count[warp], capacity);
  const DetectEntry* src = slots + static_cast<size_t>(warp) * capacity;
  DetectEntry* dst = out + offsets[warp];
  for (int i = lane; i < count; i += 32) {
    const DetectEntry e = src[i];
    int rank = 0;
    for (int k = 0; k < count; ++k) rank += (src[k].bin < e.bin) ? 1 : 0;
    dst[rank] = e;
  }
}

// ------------------------------------------------------------------------------------------------------------
// K3 — window query: max / first-argmax of the boxcar row over [bin_lo, bin_hi] for a range of frames of the last push.
// This is getMaxIndex(avgPower, N, key, groupSize) of Transmission::updateSignals (transmission.cpp:114-117) for the
// (rare) frames where no bin of the window reached the detection level, so no detection entry carries the value.
// Each work item replays the Averager for the window's bins from the nearest m_sum checkpoint (same device functions
// as k_detect => bit-identical values).
// ------------------------------------------------------------------------------------------------------------
struct WindowWork {
  int bin_lo, bin_hi;      // inclusive window, already clipped to [0, N)
  int frame_lo, frame_hi;  // [frame_lo, frame_hi) inside one checkpoint interval
  int out_offset;          // results for frame f go to out[out_offset + (f - frame_lo)]
};

struct WindowArgs {
  int n, group_y, group_x;
  const float* psd;
  const float* threshold;  // final threshold of the push
  int noise_samples, learn_frames;
  const float* ring_in;
  int avg_frames;
  const float* checkpoints;
  const WindowWork* work;
  float* out_value;
  int* out_index;
};

__global__ void __launch_bounds__(256) k_window_query(const WindowArgs a) {
  extern __shared__ float sm[];
  const WindowWork w = a.work[blockIdx.x];
  const int n = a.n, Y = a.group_y, half = a.group_x / 2;
  // columns needed: the windows of every 8-bin segment that intersects [bin_lo, bin_hi]
  const int lo = max(0, (w.bin_lo & ~(kBoxSegment - 1)) - half), hi = min(n - 1, (w.bin_hi | (kBoxSegment - 1)) + half);
  const int width = hi - lo + 1;
  float* sum_s = sm;          // [width]
  float* avg_s = sm + width;  // [width]
  __shared__ float red_v[8];
  __shared__ int red_i[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = w.frame_lo / kCheckpointEvery;
  for (int i = tid; i < width; i += blockDim.x) sum_s[i] = a.checkpoints[static_cast<size_t>(c) * n + lo + i];
  __syncthreads();
  for (int t = c * kCheckpointEvery; t < w.frame_hi; ++t) {
    for (int i = tid; i < width; i += blockDim.x) {
      const int j = lo + i;
      const float thr = a.threshold[j];
      const float q = noise_sub(a.psd[static_cast<size_t>(t) * n + j], thr, a.noise_samples + t < a.learn_frames);
      float old;
      if (t >= Y) {
        old = noise_sub(a.psd[static_cast<size_t>(t - Y) * n + j], thr, a.noise_samples + (t - Y) < a.learn_frames);
      } else {
        old = a.ring_in[static_cast<size_t>(t) * n + j];
      }
      float s = sum_s[i];
      avg_s[i] = averager_step(s, old, q, min(a.avg_frames + t + 1, Y), Y);
      sum_s[i] = s;
    }
    __syncthreads();
    if (t >= w.frame_lo) {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int j = w.bin_lo + tid; j <= w.bin_hi; j += blockDim.x) {
        const float box = boxcar_value([&](int bb) { return avg_s[bb - lo]; }, j, n, half);
        argmax_combine(bv, bi, box, j);
      }
      warp_argmax(bv, bi);
      if (lane == 0) {
        red_v[warp] = bv;
        red_i[warp] = bi;
      }
      __syncthreads();
      if (warp == 0) {
        const int nw = blockDim.x >> 5;
        bv = lane < nw ? red_v[lane] : -INFINITY;
        bi = lane < nw ? red_i[lane] : 0x7fffffff;
        warp_argmax(bv, bi);
        if (lane == 0) {
          a.out_value[w.out_offset + (t - w.frame_lo)] = bv;
          a.out_index[w.out_offset + (t - w.frame_lo)] = bi;
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------
// stand-alone operators (operator-level parity with tests/test_averager.cpp and tests/test_utils.cpp)
// ------------------------------------------------------------------------------------------------------------
// Averager::push for `count` rows: same per-bin step as k_detect. ring_in/ring_out are [group][size], oldest first.
__global__ void k_averager_push(const float* rows, int count, int size, int group, float* sum, const float* ring_in, float* ring_out, int frames_before,
                                float* avg_out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= size) return;
  float s = sum[j];
  float avg = kNoData;
  for (int t = 0; t < count; ++t) {
    const float old = (t >= group) ? rows[static_cast<size_t>(t - group) * size + j] : ring_in[static_cast<size_t>(t) * size + j];
    avg = averager_step(s, old, rows[static_cast<size_t>(t) * size + j], min(frames_before + t + 1, group), group);
  }
  sum[j] = s;
  avg_out[j] = avg;
  for (int i = 0; i < group; ++i) {
    const int t = count - group + i;
    ring_out[static_cast<size_t>(i) * size + j] = (t >= 0) ? rows[static_cast<size_t>(t) * size + j] : ring_in[static_cast<size_t>(count + i) * size + j];
  }
}

// average(in, out, size, groupSize), engine form: zero-extended 8-bin segments (same definition as k_detect)
__global__ void k_boxcar(const float* in, float* out, int size, int group, int rows) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (j >= size || r >= rows) return;
  const int half = group / 2;
  const float* row = in + static_cast<size_t>(r) * size;
  out[static_cast<size_t>(r) * size + j] = boxcar_value([&](int bb) { return row[bb]; }, j, size, half);
}

// average(in, out, size, groupSize), reference form (utils.cpp:31-53): one serial running sum per row, bit-exact.
__global__ void k_boxcar_serial(const float* in, float* out, int size, int group, int rows) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* x = in + static_cast<size_t>(r) * size;
  float* y = out + static_cast<size_t>(r) * size;
  const int half = group / 2;
  float running = 0.0f;
  int terms = 0;
  if (half == 0 && size > 0) y[size - 1] = 0.0f;
  for (int pos = -half; pos < size + half - 1; ++pos) {
    const int leaving = pos - half - 1, entering = pos + half;
    if (0 <= leaving && leaving < size) {
      running = __fsub_rn(running, x[leaving]);
      terms--;
    }
    if (0 <= entering && entering < size) {
      running = __fadd_rn(running, x[entering]);
      terms++;
    }
    if (0 <= pos && pos < size) y[pos] = __fdiv_rn(running, static_cast<float>(terms));
  }
}

}  // namespace b2s
