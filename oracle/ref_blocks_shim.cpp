// TEST INFRASTRUCTURE ONLY — drives the REFERENCE's own block objects so that the oracle can be pinned against them.
//
// Compiled by oracle/Makefile (`make ref`) together with the reference's UNMODIFIED sources, taken where they lie under
// /root/reference (never copied into this repo):
//   sources/radio/blocks/psd.cpp, noise_learner.cpp, transmission.cpp, spectrogram.cpp, sources/radio/signal.cpp,
//   sources/performance_logger.cpp, sources/network/data_controller.cpp (+ averager.cpp, utils/*.cpp, logger.cpp as before)
// What is NOT the reference's code, and why:
//   * oracle/shim/gnuradio/sync_block.h, oracle/shim/mqtt/client.h: name-only stand-ins for absent libraries;
//   * Config / Mqtt member functions below: sources/config.cpp and sources/network/mqtt.cpp pull in SoapySDR and Paho. The
//     stand-in Config returns the values the test passes; the stand-in Mqtt::publish records the payload;
//   * getTime(): sources/utils/utils.cpp is compiled with -DgetTime=ref_wallclock_getTime so that the blocks' calls to
//     getTime() resolve to the INJECTED frame clock below (the wall clock is an input of the path, not part of it);
//   * the three index lambdas of SdrDevice::setupChains (sdr_device.cpp:153-158; that file needs SoapySDR) are restated.
#include <config.h>
#include <logger.h>
#include <network/data_controller.h>
#include <network/mqtt.h>
#include <radio/blocks/noise_learner.h>
#include <radio/blocks/psd.h>
#include <radio/blocks/spectrogram.h>
#include <radio/blocks/transmission.h>

#include <cstring>
#include <memory>

// ---- injected clock ----
static std::chrono::milliseconds g_now{0};
std::chrono::milliseconds getTime() { return g_now; }

// ---- stand-in Config (sources/config.h:45-85 declares it; the values come from the test) ----
Config::Config(const nlohmann::json& json)
    : m_json(json),
      m_devices(),
      m_isColorLogEnabled(false),
      m_consoleLogLevel(spdlog::level::off),
      m_fileLogLevel(spdlog::level::off),
      m_ignoredRanges([&json]() {
        std::vector<FrequencyRange> r;
        for (const auto& p : json.at("ignored")) r.emplace_back(p.at(0).get<Frequency>(), p.at(1).get<Frequency>());
        return r;
      }()),
      m_recordingBandwidth(json.at("bandwidth").get<Frequency>()),
      m_recordingMinTime(json.at("min_time_ms").get<int64_t>()),
      m_recordingTimeout(json.at("timeout_ms").get<int64_t>()),
      m_recordingTuningStep(json.at("tuning_step").get<Frequency>()),
      m_workers(0),
      m_mqttUrl(),
      m_mqttUsername(),
      m_mqttPassword() {}
Config Config::loadFromFile(const std::string& text) { return Config(nlohmann::json::parse(text)); }  // here: the JSON text itself
std::vector<FrequencyRange> Config::ignoredRanges() const { return m_ignoredRanges; }
Frequency Config::recordingBandwidth() const { return m_recordingBandwidth; }
std::chrono::milliseconds Config::recordingMinTime() const { return m_recordingMinTime; }
std::chrono::milliseconds Config::recordingTimeout() const { return m_recordingTimeout; }
Frequency Config::recordingTuningStep() const { return m_recordingTuningStep; }

// ---- stand-in Mqtt (sources/network/mqtt.h:15-41): records what DataController publishes ----
namespace {
struct Published {
  std::string topic;
  std::vector<uint8_t> payload;
};
std::vector<Published> g_published;
void ensureLogger() {
  static bool done = false;
  if (!done) {
    Logger::configure(spdlog::level::off, spdlog::level::off, "", 0, 0, true);  // tests/test_main.cpp:5
    done = true;
  }
}
}  // namespace
Mqtt::Mqtt(const Config& config) : m_config(config), m_client(), m_isRunning(false) {}
Mqtt::~Mqtt() {}
void Mqtt::publish(const std::string& topic, const std::string& data, int) { g_published.push_back({topic, std::vector<uint8_t>(data.begin(), data.end())}); }
void Mqtt::publish(const std::string& topic, const std::vector<uint8_t>& data, int) { g_published.push_back({topic, data}); }
void Mqtt::publish(const std::string& topic, const std::vector<uint8_t>&& data, int) { g_published.push_back({topic, data}); }
void Mqtt::setMessageCallback(const std::string&, std::function<void(const std::string&)>) {}

// ---- one detection chain built from the reference's objects, wired like SdrDevice::setupChains (sdr_device.cpp:147-171) ----
namespace {
struct RefChain {
  Config config;
  Device device;
  int n;
  Frequency sampleRate, center, rangeLo, rangeHi;
  double step;
  TransmissionNotification notification;
  Mqtt mqtt;
  DataController dataController;
  std::unique_ptr<NoiseLearner> noiseLearner;
  std::unique_ptr<Transmission> transmission;
  std::unique_ptr<Spectrogram> spectrogram;

  RefChain(const std::string& configJson, int n_, Frequency fs, Frequency c, Frequency lo, Frequency hi, int groupSize, float start, float stop, bool withSpectrogram)
      : config(Config::loadFromFile(configJson)), n(n_), sampleRate(fs), center(c), rangeLo(lo), rangeHi(hi), step(static_cast<double>(fs) / n_), mqtt(config), dataController(mqtt, "dev") {
    device.m_sampleRate = fs;
    device.m_startLevel = start;
    device.m_stopLevel = stop;
    // sdr_device.cpp:153-158
    const auto indexToFrequency = [this](const int index) { return center + static_cast<Frequency>(step * (index + 0.5)) - sampleRate / 2; };
    const auto indexToShift = [this](const int index) { return static_cast<Frequency>(step * (index + 0.5)) - sampleRate / 2; };
    const auto isIndexInRange = [this, indexToFrequency](const int index) {
      const auto f = indexToFrequency(index);
      return rangeLo <= f && f <= rangeHi;
    };
    const auto getFrequency = [this]() { return center; };
    noiseLearner = std::make_unique<NoiseLearner>(n, getFrequency, indexToFrequency);
    transmission = std::make_unique<Transmission>(config, device, n, groupSize, notification, indexToFrequency, indexToShift, isIndexInRange);
    if (withSpectrogram) spectrogram = std::make_unique<Spectrogram>(n, fs, dataController, getFrequency);
  }
};
}  // namespace

extern "C" {

void ref_set_time(int64_t ms) { g_now = std::chrono::milliseconds(ms); }

// PSD::work (psd.cpp:11-22) on `items` vectors of n complex values
void ref_psd_work(int n, int sampleRate, const float* spectrumInterleaved, float* out, int items) {
  ensureLogger();
  PSD psd(n, sampleRate);
  gr_vector_const_void_star in{spectrumInterleaved};
  gr_vector_void_star o{out};
  psd.work(items, in, o);
}

void* ref_chain_create(const char* configJson, int n, int sampleRate, int center, int rangeLo, int rangeHi, int groupSize, float start, float stop, int withSpectrogram, int64_t nowMs) {
  ensureLogger();
  g_now = std::chrono::milliseconds(nowMs);
  return new RefChain(configJson, n, sampleRate, center, rangeLo, rangeHi, groupSize, start, stop, withSpectrogram != 0);
}
void ref_chain_destroy(void* h) { delete static_cast<RefChain*>(h); }
void ref_chain_set_center(void* h, int center, int lo, int hi) {
  auto* c = static_cast<RefChain*>(h);
  c->center = center;
  c->rangeLo = lo;
  c->rangeHi = hi;
}
void ref_chain_reset(void* h) { static_cast<RefChain*>(h)->transmission->resetBuffers(); }  // sdr_device.cpp:74

// One PSD row through psd -> noiseLearner -> transmission and psd -> spectrogram at time nowMs (one work() call each,
// as the flat-out GNU Radio scheduler would do with one item available). Returns the number of transmissions.
int ref_chain_push_row(void* h, const float* psdRow, int64_t nowMs, float* noiseSubOut, int32_t* txFreq, int32_t* txFlush, int cap) {
  auto* c = static_cast<RefChain*>(h);
  g_now = std::chrono::milliseconds(nowMs);
  std::vector<float> q(c->n);
  {
    gr_vector_const_void_star in{psdRow};
    gr_vector_void_star out{q.data()};
    c->noiseLearner->work(1, in, out);
  }
  if (noiseSubOut) std::memcpy(noiseSubOut, q.data(), sizeof(float) * c->n);
  {
    gr_vector_const_void_star in{q.data()};
    gr_vector_void_star out;
    c->transmission->work(1, in, out);
  }
  if (c->spectrogram) {
    gr_vector_const_void_star in{psdRow};
    gr_vector_void_star out;
    c->spectrogram->work(1, in, out);
  }
  const auto list = c->notification.wait();  // notify() ran inside work(): returns at once
  int count = 0;
  for (const auto& ff : list) {
    if (count < cap) {
      txFreq[count] = ff.first;
      txFlush[count] = ff.second ? 1 : 0;
    }
    ++count;
  }
  return count;
}

// payloads recorded by the stand-in Mqtt::publish since the last clear
int ref_published_count() { return static_cast<int>(g_published.size()); }
int ref_published_get(int i, char* topic, int topicCap, uint8_t* payload, int payloadCap) {
  const auto& p = g_published.at(i);
  std::snprintf(topic, topicCap, "%s", p.topic.c_str());
  const int n = static_cast<int>(p.payload.size());
  if (n <= payloadCap) std::memcpy(payload, p.payload.data(), n);
  return n;
}
void ref_published_clear() { g_published.clear(); }

// DataController::pushSpectrogram / pushTransmission directly (data_controller.cpp:27-57)
void ref_push_spectrogram(int64_t timeMs, int frequency, int sampleRate, const int8_t* data, int size) {
  ensureLogger();
  static Config cfg = Config::loadFromFile("{\"ignored\":[],\"bandwidth\":0,\"min_time_ms\":0,\"timeout_ms\":0,\"tuning_step\":1}");
  static Mqtt mqtt(cfg);
  DataController dc(mqtt, "dev");
  dc.pushSpectrogram(std::chrono::milliseconds(timeMs), frequency, sampleRate, data, size);
}
void ref_push_transmission(int64_t timeMs, int frequency, int sampleRate, const int8_t* iq, int size) {
  ensureLogger();
  static Config cfg = Config::loadFromFile("{\"ignored\":[],\"bandwidth\":0,\"min_time_ms\":0,\"timeout_ms\":0,\"tuning_step\":1}");
  static Mqtt mqtt(cfg);
  DataController dc(mqtt, "dev");
  dc.pushTransmission(std::chrono::milliseconds(timeMs), frequency, sampleRate, reinterpret_cast<const SimpleComplex*>(iq), size);
}
}
