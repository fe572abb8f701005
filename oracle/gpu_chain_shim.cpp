// TEST INFRASTRUCTURE ONLY — drives the product's host wrapper (rtl-sdr-scanner-cpp_b200/host/gpu_scan_chain.h) the way GNU Radio
// and SdrDevice would: constructed from the reference's Config / Device / TransmissionNotification / DataController types, fed
// through work() with CF32 items, read back through the notification mailbox and the MQTT stand-in of ref_blocks_shim.cpp.
// Compiled into oracle/_ref/libref.so next to the reference's own objects (same injected clock, same recorded publishes), so a
// test can run the reference's NoiseLearner/Transmission/Spectrogram and this block side by side.
#include <network/mqtt.h>

#include "../rtl-sdr-scanner-cpp_b200/host/gpu_scan_chain.h"

namespace {
struct GpuChainHarness {
  Config config;
  Device device;
  Frequency center, rangeLo, rangeHi;
  TransmissionNotification notification;
  Mqtt mqtt;
  DataController dataController;
  std::unique_ptr<GpuScanChain> chain;
  GpuChainHarness(const std::string& json, Frequency fs, Frequency c, Frequency lo, Frequency hi, float start, float stop)
      : config(Config::loadFromFile(json)), center(c), rangeLo(lo), rangeHi(hi), mqtt(config), dataController(mqtt, "dev") {
    device.m_sampleRate = fs;
    device.m_startLevel = start;
    device.m_stopLevel = stop;
    chain = std::make_unique<GpuScanChain>(
        config, device, notification, dataController, [this]() { return center; }, [this]() { return FrequencyRange(rangeLo, rangeHi); });
  }
};
thread_local std::string g_chain_error;
}  // namespace

extern "C" {
void ref_set_time(int64_t ms);  // ref_blocks_shim.cpp: the injected getTime()

const char* gpuchain_last_error() { return g_chain_error.c_str(); }
void* gpuchain_create(const char* configJson, int sampleRate, int center, int rangeLo, int rangeHi, float start, float stop, int64_t nowMs) {
  ref_set_time(nowMs);
  try {
    return new GpuChainHarness(configJson, sampleRate, center, rangeLo, rangeHi, start, stop);
  } catch (const std::exception& e) {  // constructors throw like the reference's (main.cpp:60 catches per device)
    g_chain_error = e.what();
    return nullptr;
  }
}
void gpuchain_destroy(void* h) { delete static_cast<GpuChainHarness*>(h); }
int gpuchain_fft_size(void* h) { return static_cast<GpuChainHarness*>(h)->chain->fftSize(); }
int gpuchain_decimator(void* h) { return static_cast<GpuChainHarness*>(h)->chain->decimatorFactor(); }
long gpuchain_item_bytes(int sampleRate) { return static_cast<long>(GpuScanChain::itemBytes(sampleRate)); }
void gpuchain_set_center(void* h, int center, int lo, int hi) {
  auto* c = static_cast<GpuChainHarness*>(h);
  c->center = center;
  c->rangeLo = lo;
  c->rangeHi = hi;
}
void gpuchain_reset(void* h) { static_cast<GpuChainHarness*>(h)->chain->resetBuffers(); }
// one work() call with n_items input items at time nowMs; returns the number of transmissions the mailbox received
int gpuchain_work(void* h, const float* items, int n_items, int64_t nowMs, int32_t* txFreq, int32_t* txFlush, int cap) {
  auto* c = static_cast<GpuChainHarness*>(h);
  ref_set_time(nowMs);
  gr_vector_const_void_star in{items};
  gr_vector_void_star out;
  if (c->chain->work(n_items, in, out) != n_items) return -1;
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
}
