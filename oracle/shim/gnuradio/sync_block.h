// TEST INFRASTRUCTURE ONLY — stand-in for <gnuradio/sync_block.h> (GNU Radio is not installable in this image).
// It provides only the names the reference's block sources mention — the base class, the I/O signature factory and
// the work() buffer types — so that sources/radio/blocks/{psd,noise_learner,transmission,spectrogram}.cpp compile
// UNMODIFIED from /root/reference and their work() can be driven directly by oracle/ref_blocks_shim.cpp. No
// scheduling, no buffers, no DSP lives here: every float that comes out is computed by the reference's own code.
#pragma once
#include <chrono>
#include <complex>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

typedef std::complex<float> gr_complex;
typedef std::vector<const void*> gr_vector_const_void_star;
typedef std::vector<void*> gr_vector_void_star;

namespace gr {

class io_signature {
 public:
  typedef std::shared_ptr<io_signature> sptr;
  static sptr make(int /*min_streams*/, int /*max_streams*/, int /*sizeof_stream_item*/) { return std::make_shared<io_signature>(); }
};

class sync_block {
 public:
  sync_block() {}
  sync_block(const std::string& /*name*/, io_signature::sptr /*input*/, io_signature::sptr /*output*/) {}
  virtual ~sync_block() {}
  virtual int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) = 0;
};

}  // namespace gr
