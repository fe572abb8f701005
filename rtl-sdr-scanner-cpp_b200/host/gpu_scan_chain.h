// GpuScanChain — the host C++ block that puts libb2s.so where the reference's detection chain sits.
//
// The reference assembles  decimator -> fft_v(hamming, shift) -> PSD -> NoiseLearner -> Transmission  and  PSD -> Spectrogram
// in SdrDevice::setupChains (reference sources/radio/sdr_device.cpp:161-171). This block has the io_signature of the Blocker's
// output (one item = fftSize * decimatorFactor complex floats, sdr_device.cpp:161-162) and no output stream, exactly like
// Transmission (transmission.cpp:18); its work() forwards the items to b2s_band_push and hands the result to the same two
// consumers the reference's blocks feed: TransmissionNotification::notify (transmission.cpp:67) and
// DataController::pushSpectrogram (spectrogram.cpp:70). Everything else of the reference stays as it is.
//
// It is written against the reference's own headers (config.h, notification.h, radio/help_structures.h, network/data_controller.h,
// utils/utils.h, <gnuradio/sync_block.h>) and include/b2s.h; a maintainer adds this one file to sources/radio/blocks/ and links
// libb2s.so. Nothing here computes: the hot path runs on the GPU behind the C ABI, and construction fails with
// std::runtime_error when no B200 is usable (there is no CPU fallback).
#pragma once

#include <b2s.h>
#include <config.h>
#include <gnuradio/sync_block.h>
#include <logger.h>
#include <network/data_controller.h>
#include <radio/help_structures.h>
#include <utils/utils.h>

#include <algorithm>
#include <cstdlib>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

class GpuScanChain : virtual public gr::sync_block {
 public:
  // Same inputs as the blocks it replaces get in setupChains: the Config / Device pair, the notification mailbox of the Scanner, the
  // DataController the Spectrogram publishes through, and the two getters SdrDevice binds (getFrequency, m_frequencyRange).
  GpuScanChain(
      const Config& config,
      const Device& device,
      TransmissionNotification& notification,
      DataController& dataController,
      std::function<Frequency()> getFrequency,
      std::function<FrequencyRange()> getFrequencyRange,
      int cudaDevice = 0)
      : gr::sync_block("GpuScanChain", gr::io_signature::make(1, 1, static_cast<int>(itemBytes(device.m_sampleRate))), gr::io_signature::make(0, 0, 0)),
        m_notification(notification),
        m_dataController(dataController),
        m_getFrequency(getFrequency),
        m_getFrequencyRange(getFrequencyRange),
        m_sampleRate(device.m_sampleRate) {
    if (b2s_engine_create(cudaDevice, &m_engine) != B2S_OK) throw std::runtime_error(std::string("GpuScanChain: ") + b2s_last_error());  // like sdr_device_reader.cpp:44
    b2s_band_config cfg;
    b2s_default_config(&cfg, device.m_sampleRate, getFrequency(), config.recordingBandwidth());  // N, indexStep, decimatorFactor as setupChains computes them
    cfg.iq_format = B2S_IQ_CF32;                                                                    // SdrSource delivers CF32 (sdr_source.cpp:52)
    cfg.start_level = device.m_startLevel;
    cfg.stop_level = device.m_stopLevel;
    cfg.tuning_step_hz = config.recordingTuningStep();
    cfg.min_time_ms = config.recordingMinTime().count();
    cfg.timeout_ms = config.recordingTimeout().count();
    const auto range = getFrequencyRange();
    cfg.range_lo_hz = range.first;
    cfg.range_hi_hz = range.second;
    const auto ignored = config.ignoredRanges();
    if (ignored.size() > B2S_MAX_IGNORED) {
      b2s_engine_destroy(m_engine);
      throw std::runtime_error("GpuScanChain: more ignored ranges than B2S_MAX_IGNORED");
    }
    cfg.n_ignored = static_cast<int32_t>(ignored.size());
    for (size_t i = 0; i < ignored.size(); ++i) {
      cfg.ignored_lo_hz[i] = ignored[i].first;
      cfg.ignored_hi_hz[i] = ignored[i].second;
    }
    m_periodMs = 1000.0 * cfg.frame_stride_samples / device.m_sampleRate;
    cfg.learn_frames = b2s_learn_frames_from_ms(NOISE_LEARNING_TIME.count(), m_periodMs);  // config.h:24, noise_learner.cpp:23
    cfg.max_frames_per_push = 256;  // GNU Radio hands work() a few items at a time
    m_fftSize = cfg.fft_size;
    m_decimatorFactor = cfg.frame_stride_samples / cfg.fft_size;
    m_spectrogramSize = cfg.spectrogram_out_size;
    m_row.resize(std::max(m_spectrogramSize, 1));
    if (b2s_band_create(m_engine, &cfg, &m_band) != B2S_OK) {
      const std::string why = b2s_last_error();
      b2s_engine_destroy(m_engine);
      throw std::runtime_error("GpuScanChain: " + why);
    }
  }
  ~GpuScanChain() override {
    b2s_band_destroy(m_band);
    b2s_engine_destroy(m_engine);
  }
  GpuScanChain(const GpuScanChain&) = delete;
  GpuScanChain& operator=(const GpuScanChain&) = delete;

  // bytes of one input item for a device: what stream_to_vector / Blocker produce (sdr_device.cpp:161-162)
  static size_t itemBytes(Frequency sampleRate) {
    const int n = b2s_get_fft(sampleRate, SIGNAL_DETECTION_MAX_STEP);
    return sizeof(gr_complex) * static_cast<size_t>(n) * b2s_decimator_factor(sampleRate, n);
  }
  int fftSize() const { return m_fftSize; }
  int decimatorFactor() const { return m_decimatorFactor; }

  int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star&) override {
    const auto range = m_getFrequencyRange();
    const Frequency center = m_getFrequency();
    bool ok = b2s_band_set_center(m_band, center, range.first, range.second) == B2S_OK;  // the lambdas' captures, sdr_device.cpp:146,153-158
    b2s_result result{};
    ok = ok && b2s_band_push(m_band, input_items[0], static_cast<size_t>(noutput_items), getTime().count(), m_periodMs, &result) == B2S_OK;
    int total = 0;
    std::vector<b2s_transmission> list(std::max(result.n_transmissions_total, 1));
    ok = ok && b2s_band_get_transmissions(m_band, list.data(), static_cast<int>(list.size()), &total) == B2S_OK;
    if (!ok) {  // work() has no error channel: the reference logs and exits on a fatal runtime error (sdr_source.cpp:38-40)
      Logger::error("gpu", "{}", b2s_last_error());
      std::exit(1);
    }
    std::vector<FrequencyFlush> transmissions;
    for (int i = 0; i < std::min<int>(total, static_cast<int>(list.size())); ++i) transmissions.emplace_back(list[i].shift_hz, list[i].flush != 0);
    m_notification.notify(transmissions);  // transmission.cpp:67
    // Spectrogram::send (spectrogram.cpp:62-75): rows completed during this call, oldest first
    int available = 0;
    int64_t time = 0;
    int32_t rowCenter = 0;
    while (m_spectrogramSize > 0 && b2s_band_get_spectrogram(m_band, &time, &rowCenter, m_row.data(), 1, 1, &available) == B2S_OK && available > 0) {
      m_dataController.pushSpectrogram(std::chrono::milliseconds(time), rowCenter, m_sampleRate, m_row.data(), m_spectrogramSize);
    }
    return noutput_items;
  }

  void resetBuffers() { b2s_band_reset(m_band); }  // Transmission::resetBuffers, called from SdrDevice::setFrequencyRange (sdr_device.cpp:74)

 private:
  TransmissionNotification& m_notification;
  DataController& m_dataController;
  const std::function<Frequency()> m_getFrequency;
  const std::function<FrequencyRange()> m_getFrequencyRange;
  const Frequency m_sampleRate;
  b2s_engine* m_engine = nullptr;
  b2s_band* m_band = nullptr;
  double m_periodMs = 0.0;
  int m_fftSize = 0, m_decimatorFactor = 1, m_spectrogramSize = 0;
  std::vector<int8_t> m_row;
};
