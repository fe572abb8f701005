/* b2s — B200 spectrum-scan engine: the C-ABI drop-in boundary.
 *
 * This library replaces, for ONE band (= one SDR device chain), the reference's GNU Radio block chain
 *     decimator -> fft_v(hamming, shift) -> PSD -> NoiseLearner -> Transmission   (+ PSD -> Spectrogram)
 * that is assembled in exactly one place, sources/radio/sdr_device.cpp:161-171 (reference paths are relative to
 * /root/reference). Everything upstream (SoapySDR source, stream_to_vector, Blocker) and downstream
 * (Scanner::worker, SdrDevice::updateRecordings, Recorder, DataController/MQTT) stays the reference's own code;
 * INTEGRATION.md shows the ~40-line gr::sync_block adaptor a maintainer would add.
 *
 * Conventions
 *   - every entry point returns 0 on success or a negative B2S_E_* code; nothing throws or aborts across the ABI.
 *     b2s_last_error() returns a thread-local message for the last failure on the calling thread.
 *     (reference: C++ exceptions at construction, main.cpp:60; work() has no error channel.)
 *   - plain pointers and sizes only. Device pointers are accepted where flagged.
 *   - a band handle is externally synchronised, except b2s_band_reset / b2s_band_set_center, which may be called
 *     from another thread while a push is running (same guarantee as the reference's per-block mutexes,
 *     transmission.cpp:34,43; noise_learner.cpp:40,70). Different bands are independent (own stream + state).
 *   - the hot path runs ONLY on the GPU (sm_100a). There is no CPU fallback: if no CUDA device is usable,
 *     b2s_engine_create fails with B2S_E_CUDA.
 *   - time is injected: frame k of a push is stamped now_k = t0_ms + floor(k * frame_period_ms + 0.5)
 *     (replaces getTime() in noise_learner.cpp:18, transmission.cpp:63, spectrogram.cpp:63).
 */
#ifndef B2S_H
#define B2S_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2S_VERSION 100

#define B2S_OK 0
#define B2S_E_INVALID (-1)   /* bad argument / unsupported configuration */
#define B2S_E_CUDA (-2)      /* CUDA runtime error (sticky errors: destroy and re-create the engine) */
#define B2S_E_NOMEM (-3)
#define B2S_E_OVERFLOW (-4)  /* a frame had more detection entries than detect_capacity: the push completed on truncated lists (state
                                stays consistent), the capacity grows before the next push; or more live signals than the engine tracks */
#define B2S_E_STATE (-5)

#define B2S_MAX_IGNORED 16
#define B2S_MAX_TX 64        /* transmissions held INSIDE b2s_result; the signal map itself is not limited to this:
                                n_transmissions_total reports the real count and b2s_band_get_transmissions returns the whole list.
                                The device-resident map holds up to 256 live signals per band (the reference's std::map is
                                unbounded); beyond that a push fails with B2S_E_OVERFLOW instead of dropping signals */

#define B2S_IQ_CS8 0         /* interleaved int8 I,Q (help_structures.h:17 SimpleComplex) */
#define B2S_IQ_CF32 1        /* interleaved float I,Q (what SdrSource delivers, sdr_source.cpp:52) */

#define B2S_WINDOW_HAMMING 0 /* gr::fft::window::hamming(N), sdr_device.cpp:164 */
#define B2S_WINDOW_USER 1

/* b2s_band_config.flags */
#define B2S_FLAG_IQ_ON_DEVICE 0x100 /* `iq` passed to b2s_band_push is a device pointer */
#define B2S_FLAG_ASYNC 0x200        /* b2s_band_push returns once the GPU work is enqueued; the signal bookkeeping of push k
                                       runs on a worker thread while push k+1 is in the kernels (the reference decouples the
                                       same two halves through its 1-slot mailbox, notification.h:14-26). Results are
                                       collected with b2s_band_sync; b2s_band_push must be given out = NULL. */

/* Construction-time parameters. The reference takes them from Config / Device / the setupChains lambdas
 * (sdr_device.cpp:148-167, transmission.h:17-25, config.h:24-38). */
typedef struct b2s_band_config {
  int32_t fft_size;             /* N = getFft(fs, SIGNAL_DETECTION_MAX_STEP); power of two, 256..262144 (sizes above 16384 run as 16384-point
                                   residue classes of the bin index, see csrc/spectral3.cuh) */
  int32_t sample_rate_hz;       /* Device::m_sampleRate (Frequency = int32_t) */
  int32_t frame_stride_samples; /* fftSize * decimatorFactor complex samples between frame starts (sdr_device.cpp:161-163) */
  int32_t iq_format;            /* B2S_IQ_* */
  float iq_scale;               /* CS8: x = (float)i8 * iq_scale; default 1/127 (inverse of recorder.cpp:36) */
  int32_t window_kind;          /* B2S_WINDOW_* */
  const float* window_taps;     /* host pointer, N floats, only read during b2s_band_create when B2S_WINDOW_USER */
  int32_t grouping_x;           /* GROUPING_X (21): frequency boxcar width, transmission.cpp:61 */
  int32_t grouping_y;           /* GROUPING_Y (21): Averager depth, transmission.cpp:24 */
  int32_t group_size_bins;      /* indexStep = ceil(recordingBandwidth / (fs/N)), sdr_device.cpp:151 */
  float start_level;            /* Device::m_startLevel (8 dB) */
  float stop_level;             /* Device::m_stopLevel (5 dB) */
  int32_t learn_frames;         /* noise-learning length in frames (>= 1); see b2s_learn_frames_from_ms */
  int32_t center_hz;            /* SdrDevice::getFrequency() */
  int32_t range_lo_hz;          /* m_frequencyRange.first */
  int32_t range_hi_hz;          /* m_frequencyRange.second */
  int32_t n_ignored;            /* Config::ignoredRanges() */
  int32_t ignored_lo_hz[B2S_MAX_IGNORED];
  int32_t ignored_hi_hz[B2S_MAX_IGNORED];
  int32_t tuning_step_hz;       /* Config::recordingTuningStep() */
  int64_t min_time_ms;          /* Config::recordingMinTime() */
  int64_t timeout_ms;           /* Config::recordingTimeout() */
  int64_t max_time_ms;          /* TRANSMISSION_MAX_TIME (600000) */
  int32_t spectrogram_out_size; /* min(SPECTROGRAM_MAX_FFT, getFft(fs, 1000)), spectrogram.cpp:14; 0 disables */
  int64_t spectrogram_interval_ms; /* SPECTROGRAM_SEND_INTERVAL (1000) */
  int32_t flags;                /* B2S_FLAG_* */
  /* ---- engine-only sizing (ignored by the oracle) ---- */
  int32_t max_frames_per_push;  /* capacity of the per-push device buffers; 0 -> 4096 */
  int32_t detect_capacity;      /* detection entries kept per FRAME (bins >= min(start,stop)); 0 -> clamp(N/8, 256, 4096); grows on overflow */
  /* ---- noise learning on the frame clock (read by the engine AND the oracle) ---- */
  int64_t noise_learning_ms;    /* > 0: NoiseLearner's own rule (noise_learner.cpp:11,23): a centre frequency is learned from its first frame
                                   (stamped s) up to and including the first frame stamped >= s + noise_learning_ms, however many frames
                                   that is - under a hop schedule the time spent on other centres counts, as in the reference;
                                   learn_frames is ignored. 0: learn_frames frames per centre (equal for a band that never hops).
                                   b2s_default_config sets NOISE_LEARNING_TIME = 2000 (config.h:24). */
} b2s_band_config;

/* Fill cfg with the reference's defaults for a device with this sample rate, exactly as setupChains sizes the chain
 * (sdr_device.cpp:148-152: N, indexStep, decimatorFactor) and config.h:24-38 / config.example.json:9-13. */
void b2s_default_config(b2s_band_config* cfg, int32_t sample_rate_hz, int32_t center_hz, int32_t recording_bandwidth_hz);

/* One FrequencyFlush (help_structures.h:15) as Transmission::getSortedTransmissions emits it (transmission.cpp:166-176). */
typedef struct b2s_transmission {
  int32_t shift_hz; /* getTunedFrequency(indexToShift(key), tuningStep) */
  int32_t flush;    /* Signal::needFlush(now) */
  int32_t key;      /* map key (bin index) — extra, for diagnostics */
  float power;      /* Signal::getPower() — extra */
} b2s_transmission;

/* Result buffers for one push; every pointer is optional (NULL = not wanted) and caller-owned HOST memory. */
typedef struct b2s_result {
  /* the mailbox content after the last frame (what Notification::notify last received, transmission.cpp:67) */
  int32_t n_transmissions;
  b2s_transmission transmissions[B2S_MAX_TX];
  /* per-frame lists for parity tests / callers that want every notification */
  int32_t* frame_tx_count;          /* [n_frames] */
  b2s_transmission* frame_tx;       /* [n_frames][B2S_MAX_TX] */
  int32_t* peak_index;              /* [n_frames] argmax of the raw PSD row (noise_learner.cpp:53-59) */
  float* peak_value;                /* [n_frames] raw PSD at peak_index */
  /* dense rows, [n_frames][N] each — debug / parity only (they cost PCIe time) */
  float* psd_db;                    /* PSD::work output (psd.cpp:18) */
  float* noise_sub_db;              /* NoiseLearner::work output */
  float* avg_db;                    /* Averager::average() after each push */
  float* box_db;                    /* average(avg, GROUPING_X) */
  /* statistics */
  int32_t n_detect_entries;         /* bins >= min(start,stop) level found in this push */
  int32_t n_spectrogram_rows;       /* rows completed during this push (fetch with b2s_band_get_spectrogram) */
  int32_t n_transmissions_total;    /* live transmissions after the last frame; when > B2S_MAX_TX, transmissions[] holds the
                                       B2S_MAX_TX strongest and b2s_band_get_transmissions the complete list */
} b2s_result;

typedef struct b2s_engine b2s_engine;
typedef struct b2s_band b2s_band;

const char* b2s_last_error(void);
int b2s_version(void);

/* ---- engine / band lifetime (reference: SdrDevice ctor/dtor, sdr_device.cpp:17-52) ---- */
int b2s_engine_create(int cuda_device, b2s_engine** out);
int b2s_engine_destroy(b2s_engine* e);
int b2s_engine_device_name(b2s_engine* e, char* buf, size_t cap);
int b2s_band_create(b2s_engine* e, const b2s_band_config* cfg, b2s_band** out);
int b2s_band_destroy(b2s_band* b);
/* run this band's kernels on a caller-owned CUDA stream (cudaStream_t); NULL restores the band's own stream */
int b2s_band_set_stream(b2s_band* b, void* cuda_stream);

/* ---- data path: replaces the work() calls of Decimator..Transmission (+Spectrogram) for n_frames input items ----
 * iq: n_frames frames, frame k starting at sample k*frame_stride_samples; host memory (pageable or pinned) or,
 * with B2S_FLAG_IQ_ON_DEVICE, device memory. Read-only; may be reused as soon as the call returns. */
int b2s_band_push(b2s_band* b, const void* iq, size_t n_frames, int64_t t0_ms, double frame_period_ms, b2s_result* out);

/* Async mode: wait for every outstanding push; `out` (optional) receives the mailbox after the last frame pushed so far
 * and the statistics accumulated since the previous sync. A no-op returning the last mailbox in synchronous mode. */
int b2s_band_sync(b2s_band* b, b2s_result* out);

/* ---- profiling (bench.py): per-kernel device time measured with CUDA events on the band's stream ---- */
typedef struct b2s_profile {
  double spectral_ms;        /* K1 k_spectrum: unpack+window+FFT+PSD */
  double detect_ms;          /* K2 k_detect: noise/averager/boxcar/threshold/spectrogram */
  double window_ms;          /* K3 k_window_query (only when the tracker needs sub-threshold window maxima) */
  double tracker_host_ms;    /* host time spent on the results of a push (wall clock): reading K4's result, or tracker.h when every frame's list is wanted */
  int64_t spectral_launches, detect_launches, window_launches;
  int64_t pushes, frames;
  int64_t h2d_bytes, d2h_bytes; /* bytes moved by b2s_band_push itself */
  /* load balance of K2 (one CTA per 128 bins): per-CTA run time in ms, median and slowest, summed over launches */
  double detect_cta_median_ms, detect_cta_max_ms;
  double track_ms;           /* K4 k_runs + k_track: the signal map on the device (runs beside the next push's K1) */
  int64_t track_launches;
  int64_t track_evals, track_events, track_best_index;  /* K4 work counters: block evaluations, event frames replayed, getBestIndex calls */
} b2s_profile;
int b2s_band_set_profiling(b2s_band* b, int enable); /* 0 off, 1 kernel times and byte counts, 2 also K2 per-CTA run times */
int b2s_band_get_profile(b2s_band* b, b2s_profile* out, int reset);

/* ---- side channels ---- */
int b2s_band_reset(b2s_band* b); /* Transmission::resetBuffers (transmission.cpp:42-55): drop signals, Averager::reset; noise kept */
int b2s_band_set_center(b2s_band* b, int32_t center_hz, int32_t range_lo_hz, int32_t range_hi_hz); /* retune, sdr_device.cpp:66-77 */

/* ---- state introspection (reference: Averager::average()/data(), averager.h:15-16) ---- */
int b2s_band_get_averager(b2s_band* b, float* sum /*[N]*/, float* avg /*[N]*/, float* ring /*[Y][N] oldest->newest*/, int32_t* frames);
int b2s_band_get_noise(b2s_band* b, float* threshold /*[N]*/, int32_t* samples, int32_t* ready);
/* completed spectrogram rows (Spectrogram::send, spectrogram.cpp:62-75), oldest first: up to `cap` rows are copied, *count is
 * the number available; with consume != 0 the rows copied out (and only those) are dropped from the band's list */
int b2s_band_get_spectrogram(b2s_band* b, int64_t* times_ms, int32_t* centers_hz, int8_t* rows /*[cap][out_size]*/, int cap, int consume, int* count);
/* the complete mailbox after the last finished push (Transmission::getSortedTransmissions, transmission.cpp:166-176), strongest
 * first; up to `cap` entries are copied, *count is the number of live transmissions */
int b2s_band_get_transmissions(b2s_band* b, b2s_transmission* out, int cap, int* count);
/* live signals (the std::map<Index, Signal> of transmission.h:49) */
int b2s_band_get_signals(b2s_band* b, int32_t* keys, int64_t* first_ms, int64_t* last_ms, float* power, int cap, int* count);

/* ---- stand-alone operators (operator-level parity with the reference's unit tests) ---- */
/* device-backed Averager with the reference's surface (averager.h:8-28) */
typedef struct b2s_averager b2s_averager;
int b2s_averager_create(b2s_engine* e, int size, int group_size, b2s_averager** out);
int b2s_averager_destroy(b2s_averager* a);
int b2s_averager_push(b2s_averager* a, const float* data);                /* Averager::push, one row */
int b2s_averager_push_many(b2s_averager* a, const float* rows, int count); /* count rows in one launch */
int b2s_averager_reset(b2s_averager* a);
int b2s_averager_average(b2s_averager* a, float* out);                    /* Averager::average() */
int b2s_averager_data(b2s_averager* a, float* out);                       /* Averager::data(), [group][size] oldest->newest */
int b2s_averager_sum(b2s_averager* a, float* out, int32_t* frames);       /* m_sum, m_frames */
/* average(in,out,size,groupSize) (utils.cpp:31-53) for `rows` rows on the device.
 * exact == 0: the engine's fused form (independent window sums, <= 1e-5 dB from the reference's running sum);
 * exact != 0: the reference's serial running sum, bit-exact (one thread per row). */
int b2s_average(b2s_engine* e, const float* in, float* out, int size, int group_size, int rows, int exact);
/* IQ -> raw PSD rows (unpack, window, FFT, shift, dB) only; power_lin optional (|X|^2/fs) */
int b2s_psd(b2s_engine* e, const b2s_band_config* cfg, const void* iq, size_t n_frames, float* psd_db, float* power_lin);

/* ---- recorder chain: what one reference Recorder computes (sources/radio/recorder.cpp:22-40,58-73) ----
 * rotator_cc(-shift) -> rational_resampler(f1, f2) per pair of getResamplersFactors(fs, bandwidth, RESAMPLER_THRESHOLD = 125)
 * -> complex_to_interleaved_char(x 127): a continuous IQ stream at fs in, int8 IQ at `bandwidth` samples/s out. The resamplers use
 * GNU Radio's default taps (Kaiser low-pass, beta 7, fractional bandwidth 0.4); history is zero at b2s_recorder_start.
 * Chunking into messages (recorder.cpp:35) and the wire format stay host work: b2s_pack_transmission_message. */
typedef struct b2s_recorder b2s_recorder;
int b2s_get_resamplers_factors(int32_t sample_rate_hz, int32_t bandwidth_hz, int threshold, int32_t* interp, int32_t* decim, int cap); /* radio_utils.cpp:129-152; returns the count */
int b2s_recorder_create(b2s_engine* e, int32_t sample_rate_hz, int32_t bandwidth_hz, int iq_format, float iq_scale, int flags /* B2S_FLAG_IQ_ON_DEVICE */,
                        size_t max_samples_per_push /* 0 -> 4 Mi */, b2s_recorder** out);
int b2s_recorder_destroy(b2s_recorder* r);
int b2s_recorder_start(b2s_recorder* r, int32_t shift_hz); /* Recorder::startRecording: rotator phase_inc = 2 pi (-shift) / fs, empty buffers */
int b2s_recorder_stop(b2s_recorder* r);                    /* Recorder::stopRecording */
/* n_samples consecutive IQ samples of the stream (host memory, or device memory with B2S_FLAG_IQ_ON_DEVICE) -> *n_out int8 I/Q pairs in out_iq (host) */
int b2s_recorder_push(b2s_recorder* r, const void* iq, size_t n_samples, int8_t* out_iq, size_t cap_samples, size_t* n_out);
int b2s_recorder_stages(b2s_recorder* r, int32_t* interp, int32_t* decim, int32_t* n_taps, int cap); /* returns the number of stages */
int b2s_recorder_taps(b2s_recorder* r, int stage, float* taps, int cap);                            /* returns the number of taps */

/* ---- scan policy: Scanner's hop rule and SdrDevice's recorder assignment as a host state machine ----
 * Scanner::worker (scanner.cpp:36-64): stay on a range while now <= start + RANGE_SCANNING_TIME or the last notification was not
 * empty. SdrDevice::updateRecordings (sdr_device.cpp:82-144): stop recorders whose shift left the list, flush / start the others.
 * Ranges are split like Scanner's constructor does (splitRanges(ranges, getRangeSplitSampleRate(fs)), radio_utils.cpp:162-199).
 * The caller owns the loop: each mailbox list obtained from b2s_band_push / b2s_band_sync is one notification. */
#define B2S_REC_START 1      /* Recorder::startRecording(frequency, shift) on recorder `recorder` */
#define B2S_REC_STOP 2       /* Recorder::stopRecording; duration_ms = Recorder::getDuration() */
#define B2S_REC_FLUSH 3      /* Recorder::flush */
#define B2S_REC_NONE_FREE 4  /* no recorder available for this shift (logged once, sdr_device.cpp:129-132) */
typedef struct b2s_recorder_action {
  int32_t kind;      /* B2S_REC_* */
  int32_t recorder;  /* index into the pool, -1 for B2S_REC_NONE_FREE */
  int32_t shift_hz;
  int64_t duration_ms;
} b2s_recorder_action;
typedef struct b2s_scan_policy b2s_scan_policy;
int b2s_scan_policy_create(const int32_t* range_lo_hz, const int32_t* range_hi_hz, int n_ranges, int32_t sample_rate_hz, int n_recorders, int64_t scanning_time_ms /* 0 -> 500 */,
                           b2s_scan_policy** out);
int b2s_scan_policy_destroy(b2s_scan_policy* p);
int b2s_scan_policy_ranges(b2s_scan_policy* p, int32_t* lo_hz, int32_t* hi_hz, int cap);              /* the split ranges; returns their number */
int b2s_scan_policy_begin(b2s_scan_policy* p, int64_t now_ms, int32_t* lo_hz, int32_t* hi_hz);          /* first setFrequencyRange */
/* one notification: recorder actions in the reference's order; *hop != 0 when the scanner retunes, then next_lo/next_hi hold the range */
int b2s_scan_policy_notify(b2s_scan_policy* p, int64_t now_ms, const b2s_transmission* list, int n, b2s_recorder_action* actions, int cap, int* n_actions, int* hop,
                           int32_t* next_lo_hz, int32_t* next_hi_hz);
int32_t b2s_get_range_split_sample_rate(int32_t sample_rate_hz);                                         /* radio_utils.cpp:162-172 */

/* Self-test: the 3-instruction exact division by a small constant that the Averager (m_sum / GROUPING_Y, averager.cpp:52-60) and
 * boxcar fast paths use, compared with IEEE division for EVERY float with |x| in [2^-60, 2^61) and +-0. *mismatches must be 0. */
int b2s_selftest_div_const(b2s_engine* e, int divisor, uint64_t* mismatches);

/* ---- host helpers with the reference's semantics (used by the tracker; exported for the adaptor and for tests) ---- */
int b2s_get_fft(int32_t sample_rate_hz, int32_t max_step_hz);                                  /* radio_utils.cpp:98-104 */
int32_t b2s_get_tuned_frequency(int32_t frequency_hz, int32_t step_hz);                        /* radio_utils.cpp:86-96 */
int b2s_get_max_index(const float* data, int size, int index, int group_size);                 /* collection_utils.h:9-14 */
int b2s_contains_with_margin(const int* keys, int n_keys, int index, int margin, int* found);  /* collection_utils.h:17-27 */
int b2s_most_frequent_value(const int* data, int n);                                           /* collection_utils.h:30-50 */
int b2s_learn_frames_from_ms(int64_t learning_ms, double frame_period_ms);                     /* NOISE_LEARNING_TIME -> frames */
int b2s_decimator_factor(int32_t sample_rate_hz, int32_t fft_size);                            /* sdr_device.cpp:150-152 */


/* ---- Transmission bookkeeping on HOST rows (operator-level parity with transmission.cpp:57-176; no GPU involved) ----
 * The same host tracker that follows K2's detection entries inside b2s_band_push, fed from dense rows instead:
 * box_rows[n_frames][N] = average(Averager::average(), GROUPING_X) and q_rows[n_frames][N] = the NoiseLearner output rows
 * (the Averager ring that getBestIndex votes on). Frames are stamped like b2s_band_push stamps them. With use_watch != 0 the
 * per-frame window maxima / candidate flags that K2 reports for the live keys are emulated as well (the path the band
 * takes in steady state); the lists must not depend on it. tx_count[n_frames], tx[n_frames][B2S_MAX_TX] (either may be
 * NULL). State (signal map, last Y rows of q) carries over between calls. */
typedef struct b2s_host_transmission b2s_host_transmission;
int b2s_host_transmission_create(const b2s_band_config* cfg, b2s_host_transmission** out);
int b2s_host_transmission_destroy(b2s_host_transmission* h);
int b2s_host_transmission_reset(b2s_host_transmission* h); /* Transmission::resetBuffers: drop the signals and the ring */
double b2s_host_transmission_last_run_ms(b2s_host_transmission* h); /* wall time of the bookkeeping of the last push (measurement) */
int b2s_host_transmission_push(b2s_host_transmission* h, const float* box_rows, const float* q_rows, int n_frames, int64_t t0_ms,
                               double frame_period_ms, int use_watch, int32_t* tx_count, b2s_transmission* tx);

/* ---- wire formats of the reference's MQTT payloads (network/data_controller.cpp:27-57), little-endian, packed ----
 * so that rows / recordings produced here can be published to an unchanged sdr-hub. Both return 0 and the payload length in
 * *written, or B2S_E_INVALID when `cap` is too small (then *written holds the required size).
 * spectrogram ("sdr/<dev>/spectrogram"):        u64 time_ms, i32 start_hz, i32 stop_hz, i32 step_hz, u32 size, int8[size]
 * transmission ("sdr/<dev>/transmission/uint8"): u64 time_ms, i32 start_hz, i32 stop_hz, u32 sample_rate, uint8 IQ pairs (int8 ^ 0x80) */
int b2s_pack_spectrogram_message(int64_t time_ms, int32_t center_hz, int32_t sample_rate_hz, const int8_t* row, int size, uint8_t* out, size_t cap,
                                 size_t* written); /* DataController::pushSpectrogram, data_controller.cpp:44-57 */
int b2s_pack_transmission_message(int64_t time_ms, int32_t frequency_hz, int32_t sample_rate_hz, const int8_t* iq, int n_samples, uint8_t* out,
                                  size_t cap, size_t* written); /* DataController::pushTransmission, data_controller.cpp:27-42 */

#ifdef __cplusplus
}
#endif
#endif /* B2S_H */
