/* TEST INFRASTRUCTURE ONLY — CPU oracle for the spectrum-scan hot path.
 *
 * A plain restatement of the reference's per-band chain
 *   decimator -> fft(window, shift) -> psd -> noiseLearner -> transmission (+ spectrogram)
 * (reference: sources/radio/sdr_device.cpp:161-171), used as the checker in tests/, in
 * __graft_entry__.smoke() and as bench.py's cpu_baseline / --impl reference arm. Nothing in the product path
 * (rtl-sdr-scanner-cpp_b200/) may include, link or call this.
 *
 * Pinning status:
 *   - Averager, average(), getMaxIndex, containsWithMargin, mostFrequentValue, getFft, getTunedFrequency,
 *     setNoData: PINNED against the reference's own gtest vectors (tests/golden/) and against the reference's
 *     own objects compiled from /root/reference into oracle/_ref/libref.so (oracle/Makefile).
 *   - PSD::work, NoiseLearner::work, Transmission::work (+ Signal), Spectrogram::work, DataController payloads: PINNED
 *     against the reference's own block objects — sources/radio/blocks/{psd,noise_learner,transmission,spectrogram}.cpp,
 *     sources/radio/signal.cpp, sources/network/data_controller.cpp compiled UNMODIFIED into oracle/_ref/libref.so behind
 *     name-only stand-ins for the absent GNU Radio / Paho headers and an injected clock (oracle/ref_blocks_shim.cpp);
 *     tests/test_oracle_vs_reference_blocks.py drives both with the same PSD rows, frame by frame: rows bit for bit,
 *     transmission lists and published payloads identical.
 *   - window, FFT, fftshift (gr::fft::fft_v = GNU Radio + FFTW, not under /root/reference): PARITY UNPINNED for these
 *     pieces; they follow the documented GNU Radio 3.10 behaviour, cross-checked against numpy.fft, closed-form known
 *     answers and scripts/converter.py:17-21 semantics.
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_IGNORED 16
#define ORC_MAX_TX 64

/* Same field order as b2s_band_config (include/b2s.h) so one ctypes.Structure serves both. */
typedef struct orc_config {
  int32_t fft_size;             /* N, power of two */
  int32_t sample_rate_hz;       /* Frequency = int32_t, help_structures.h:13 */
  int32_t frame_stride_samples; /* r*N complex samples between frame starts (decimator.h:16) */
  int32_t iq_format;            /* 0 = CS8 interleaved int8 I,Q ; 1 = CF32 interleaved float I,Q */
  float iq_scale;               /* CS8: x = (float)i8 * iq_scale */
  int32_t window_kind;          /* 0 = Hamming (gr::fft::window::hamming), 1 = user taps */
  const float* window_taps;     /* N floats when window_kind == 1 */
  int32_t grouping_x;           /* GROUPING_X = 21, config.h:28 */
  int32_t grouping_y;           /* GROUPING_Y = 21, config.h:29 */
  int32_t group_size_bins;      /* indexStep, sdr_device.cpp:151 */
  float start_level;            /* Device::m_startLevel */
  float stop_level;             /* Device::m_stopLevel */
  int32_t learn_frames;         /* noise learning length in frames (>=1), replaces NOISE_LEARNING_TIME */
  int32_t center_hz;
  int32_t range_lo_hz;
  int32_t range_hi_hz;
  int32_t n_ignored;
  int32_t ignored_lo_hz[ORC_MAX_IGNORED];
  int32_t ignored_hi_hz[ORC_MAX_IGNORED];
  int32_t tuning_step_hz;       /* recordingTuningStep */
  int64_t min_time_ms;          /* recordingMinTime */
  int64_t timeout_ms;           /* recordingTimeout */
  int64_t max_time_ms;          /* TRANSMISSION_MAX_TIME = 10 min, config.h:21 */
  int32_t spectrogram_out_size; /* min(16384, getFft(fs, 1000)), spectrogram.cpp:14 ; 0 = disabled */
  int64_t spectrogram_interval_ms; /* SPECTROGRAM_SEND_INTERVAL = 1000 */
  int32_t flags;                /* oracle: bit0 = use the fp32 FFT (timed CPU baseline) instead of fp64 */
  int32_t engine_max_frames_per_push; /* (engine-only sizing fields of b2s_band_config: same layout, not read here) */
  int32_t engine_detect_capacity;
  int64_t noise_learning_ms;    /* > 0: Noise::add's wall-clock rule on the injected frame clock (noise_learner.cpp:11,23): ready after the first
                                   frame stamped >= (stamp of the centre's first frame) + noise_learning_ms; 0: learn_frames frames */
} orc_config;

/* Optional dense per-frame outputs; any pointer may be NULL. */
typedef struct orc_outputs {
  float* psd_db;        /* [n][N] raw PSD, psd.cpp:18 */
  float* noise_sub_db;  /* [n][N] NoiseLearner output */
  float* avg_db;        /* [n][N] Averager::average() after the push */
  float* box_db;        /* [n][N] average(avg, X) */
  int32_t* peak_index;  /* [n] argmax of raw PSD (noise_learner.cpp:53-59); -1 on learning frames */
  int32_t* tx_count;    /* [n] */
  int32_t* tx_freq;     /* [n][ORC_MAX_TX] tuned shift, transmission.cpp:172 */
  int32_t* tx_flush;    /* [n][ORC_MAX_TX] */
  int32_t* tx_key;      /* [n][ORC_MAX_TX] signal map key (bin index) */
  float* tx_power;      /* [n][ORC_MAX_TX] Signal::getPower() */
} orc_outputs;

typedef struct orc_chain orc_chain;

orc_chain* orc_chain_create(const orc_config* cfg);
void orc_chain_destroy(orc_chain* c);
/* frame k is stamped now_k = t0_ms + floor(k*frame_period_ms + 0.5) */
int orc_chain_push(orc_chain* c, const void* iq, size_t n_frames, int64_t t0_ms, double frame_period_ms, const orc_outputs* out);
/* same from the PSD rows on (psd_rows[n_frames][N] = what PSD::work emitted): noiseLearner -> transmission, psd -> spectrogram */
int orc_chain_push_psd(orc_chain* c, const float* psd_rows, size_t n_frames, int64_t t0_ms, double frame_period_ms, const orc_outputs* out);
void orc_chain_reset(orc_chain* c);                                    /* Transmission::resetBuffers, transmission.cpp:42-55 */
void orc_chain_set_center(orc_chain* c, int32_t center, int32_t lo, int32_t hi); /* sdr_device.cpp:77 */
void orc_chain_get_averager(orc_chain* c, float* sum, float* avg, float* ring, int32_t* frames);
int orc_chain_get_noise(orc_chain* c, float* thr, int32_t* samples);  /* returns ready flag for the current centre */
/* spectrogram rows sent so far (spectrogram.cpp:62-75): returns count; copies up to cap rows */
int orc_chain_get_spectrogram(orc_chain* c, int64_t* times, int32_t* centers, int8_t* rows, int cap);
void orc_chain_clear_spectrogram(orc_chain* c);
/* live signals in key order (the std::map of transmission.h:49); returns the count, copies up to cap */
int orc_chain_get_signals(orc_chain* c, int32_t* keys, int64_t* first_ms, int64_t* last_ms, float* power, int cap);
/* the complete getSortedTransmissions list after the most recent frame (not bounded by ORC_MAX_TX); returns the count */
int orc_chain_get_transmissions(orc_chain* c, int32_t* freq, int32_t* flush, int32_t* key, float* power, int cap);

/* stand-alone operators */
void orc_hamming(int n, float* w);
void orc_fft_f64(int n, const float* in_interleaved, float* out_interleaved); /* unnormalised forward DFT, fp64 inside */
void orc_fft_f32(int n, const float* in_interleaved, float* out_interleaved); /* same in fp32 (timed baseline) */
void orc_psd_frame(const orc_config* cfg, const float* window, const void* iq_frame, float* psd_db, float* power_lin);
void orc_psd_from_spectrum(int n, int32_t sample_rate, const float* x_interleaved, float* out, int items); /* PSD::work, psd.cpp:11-22 */
void orc_average(const float* in, float* out, int size, int group_size);      /* utils.cpp:31-53 */
int orc_get_max_index(const float* data, int size, int index, int group_size); /* collection_utils.h:9-14 */
int orc_contains_with_margin(const int* keys, int n_keys, int index, int margin, int* found); /* collection_utils.h:17-27 */
int orc_most_frequent_value(const int* data, int n);                          /* collection_utils.h:30-50 */
int orc_get_fft(int32_t sample_rate, int32_t max_step);                       /* radio_utils.cpp:98-104 */
int32_t orc_get_tuned_frequency(int32_t f, int32_t step);                     /* radio_utils.cpp:86-96 */

/* MQTT payload layouts, network/data_controller.cpp:27-57; return the number of bytes written into out (sized by the caller) */
size_t orc_spectrogram_message(int64_t time_ms, int32_t frequency, int32_t sample_rate, const int8_t* data, int size, uint8_t* out);
size_t orc_transmission_message(int64_t time_ms, int32_t frequency, int32_t sample_rate, const int8_t* iq, int size, uint8_t* out);

typedef struct orc_averager orc_averager;
orc_averager* orc_averager_create(int size, int group_size);
void orc_averager_destroy(orc_averager* a);
void orc_averager_push(orc_averager* a, const float* data);
void orc_averager_reset(orc_averager* a);
void orc_averager_average(orc_averager* a, float* out);
void orc_averager_data(orc_averager* a, float* out); /* [group][size] oldest -> newest */
void orc_averager_sum(orc_averager* a, float* out);
int orc_averager_frames(orc_averager* a);

/* multi-threaded throughput run of the fp32 path for bench.py (cpu_baseline / --impl reference):
 * splits n_frames into `threads` contiguous segments, each with its own chain. Returns seconds. */
double orc_bench_run(const orc_config* cfg, const void* iq, size_t n_frames, double frame_period_ms, int threads);
void orc_bench_stage_seconds(double* out3); /* of the last run, summed over threads: unpack+window+FFT, PSD, noise+averager+detect */
const char* orc_fft_backend(void);           /* FFTW3f when the box has libfftw3f.so.3 (probed with dlopen), else the in-repo fp32 FFT */

#ifdef __cplusplus
}
#endif
