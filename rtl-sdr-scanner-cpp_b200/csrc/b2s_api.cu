// b2s C-ABI implementation (include/b2s.h): engine / band objects, kernel launches, host tracker glue.
// The compute path is CUDA only; there is deliberately no CPU fallback anywhere in this file.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "../../include/b2s.h"
#include "detect.cuh"
#include "host_utils.h"
#include "recorder.cuh"
#include "scan_policy.h"
#include "spectral3.cuh"
#include "track.cuh"
#include "tracker.h"

namespace {

thread_local std::string g_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
  return code;
}

#define CU(call)                                                                                        \
  do {                                                                                                  \
    cudaError_t err__ = (call);                                                                         \
    if (err__ != cudaSuccess) return fail(B2S_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__); \
  } while (0)

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  int alloc(size_t count) {
    if (count <= n) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    if (cudaMalloc(&p, count * sizeof(T)) != cudaSuccess) return fail(B2S_E_NOMEM, "cudaMalloc of %zu bytes failed", count * sizeof(T));
    n = count;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

template <typename T>
struct PinBuf {
  T* p = nullptr;
  size_t n = 0;
  int alloc(size_t count) {
    if (count <= n) return 0;
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = 0;
    if (cudaMallocHost(&p, count * sizeof(T)) != cudaSuccess) return fail(B2S_E_NOMEM, "cudaMallocHost of %zu bytes failed", count * sizeof(T));
    n = count;
    return 0;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = 0;
  }
};

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

void make_window(const b2s_band_config& cfg, std::vector<float>& w) {
  const int n = cfg.fft_size;
  w.resize(n);
  if (cfg.window_kind == B2S_WINDOW_USER && cfg.window_taps) {
    std::memcpy(w.data(), cfg.window_taps, sizeof(float) * n);
  } else {
    // gr::fft::window::hamming(N) (reference call site sources/radio/sdr_device.cpp:164): symmetric, double -> f32
    const double m = static_cast<double>(n - 1);
    for (int i = 0; i < n; ++i) w[i] = (n == 1) ? 1.0f : static_cast<float>(0.54 - 0.46 * std::cos((2.0 * M_PI * i) / m));
  }
}

}  // namespace

using namespace b2s;

// ------------------------------------------------------------------------------------------------------------
// engine
// ------------------------------------------------------------------------------------------------------------
struct b2s_engine {
  int device = 0;
  cudaDeviceProp prop{};
  int sm_count = 0;
  // kernels whose function attributes have been set on THIS device -> resident CTAs per SM. Function attributes are per
  // device, so the cache lives in the engine (one engine per GPU; several engines may share a process).
  std::mutex attr_mutex;
  std::map<const void*, int> kernel_ctas;
};

namespace {

// Opt the kernel into `smem` bytes of dynamic shared memory on the engine's device (once per engine) and report how many
// CTAs of `threads` threads fit on an SM.
template <typename K>
int prepare_kernel(b2s_engine* e, K kernel, int threads, size_t smem, int* ctas_per_sm) {
  std::lock_guard<std::mutex> lk(e->attr_mutex);
  const void* key = reinterpret_cast<const void*>(kernel);
  auto it = e->kernel_ctas.find(key);
  if (it == e->kernel_ctas.end()) {
    int ctas = 0;
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kernel, threads, smem));
    if (ctas < 1) return fail(B2S_E_CUDA, "a kernel needing %zu bytes of shared memory and %d threads does not fit on an SM of this device", smem, threads);
    it = e->kernel_ctas.emplace(key, ctas).first;
  }
  if (ctas_per_sm) *ctas_per_sm = it->second;
  return 0;
}

// K1 launcher -------------------------------------------------------------------------------------------------
template <int N, int MODE, bool LIN>
int launch_spectrum_v(b2s_engine* e, const SpectralArgs& a, cudaStream_t stream) {
  using PL = FftPlanT<N>;
  constexpr int T = N / PL::E;
  const size_t smem = sizeof(float2) * (exchange_elems<N>() + TwiddleLayout<N>::SMEM) + (MODE == kModeCs8Tma ? 2 * N : 0);
  int ctas_per_sm = 1;
  int rc = prepare_kernel(e, k_spectrum<N, MODE, LIN>, T, smem, &ctas_per_sm);
  if (rc) return rc;
  const int grid = std::min(a.n_frames, e->sm_count * ctas_per_sm);
  k_spectrum<N, MODE, LIN><<<grid, T, smem, stream>>>(a);
  CU(cudaGetLastError());
  return 0;
}
// k_spectrum3: N = RA * 1024 directly (RA = 4, 8, 16), or N = S * 16384 through the split mode (S = a.split > 1)
template <int RA, int MODE, bool LIN, int S>
int launch_spectrum3_v(b2s_engine* e, const SpectralArgs& a, cudaStream_t stream) {
  constexpr int T = RA * 32;
  constexpr bool SPLIT = S > 1;
  const size_t smem = sizeof(float2) * (RA * kBlockPitch + 31 * 32) + (MODE == kModeCs8Tma ? (SPLIT ? 2 * kSplitStageBytes : 2 * RA * 1024) : 0);
  int ctas_per_sm = 1;
  int rc = prepare_kernel(e, k_spectrum3<RA, MODE, LIN, S>, T, smem, &ctas_per_sm);
  if (rc) return rc;
  if (!a.work_counter) return fail(B2S_E_INVALID, "k_spectrum3 needs a work counter");
  const int items = a.n_frames * S;
  if (a.split != S) return fail(B2S_E_INVALID, "split tables were built for another fft_size");
  const int grid = std::min(items, std::max(1, e->sm_count - a.reserve_sms) * ctas_per_sm);
  if (SPLIT) {
    if (!a.peak_packed || !a.split_tw || !a.split_ws) return fail(B2S_E_INVALID, "split-mode tables are missing");
    CU(cudaMemsetAsync(a.peak_packed, 0, sizeof(unsigned long long) * a.n_frames, stream));
  }
  k_spectrum3<RA, MODE, LIN, S><<<grid, T, smem, stream>>>(a);
  CU(cudaGetLastError());
  if (SPLIT) {
    k_peak_unpack<<<(a.n_frames + 255) / 256, 256, 0, stream>>>(a.peak_packed, a.n_frames, a.peak_index, a.peak_value);
    CU(cudaGetLastError());
  }
  return 0;
}

// Which K1 serves which FFT size: k_spectrum (Stockham, two barriers per pass) below 4096, k_spectrum3 (warp-local passes) for
// 4096 / 8192 / 16384, and k_spectrum3's split mode (S residue classes x 16384 points) for 32768 ... 262144.
constexpr int kMaxFft = 16 * kSplitM;
constexpr bool k1_is_v3(int n) { return n >= 4096; }

template <int N, int MODE>
int launch_spectrum_t(b2s_engine* e, const SpectralArgs& a, cudaStream_t stream) {
  // the |X|^2/fs debug rows come from a debug twin of each instantiation (parity tests); the product one stays lean
  if constexpr (N > kSplitM) {
    if (a.power_lin) return launch_spectrum3_v<16, MODE, true, N / kSplitM>(e, a, stream);
    return launch_spectrum3_v<16, MODE, false, N / kSplitM>(e, a, stream);
  } else if constexpr (k1_is_v3(N)) {
    if (a.power_lin) return launch_spectrum3_v<N / 1024, MODE, true, 1>(e, a, stream);
    return launch_spectrum3_v<N / 1024, MODE, false, 1>(e, a, stream);
  } else {
    if (a.power_lin) return launch_spectrum_v<N, MODE, true>(e, a, stream);
    return launch_spectrum_v<N, MODE, false>(e, a, stream);
  }
}

template <int MODE>
int launch_spectrum_n(b2s_engine* e, int n, const SpectralArgs& a, cudaStream_t stream) {
  if (!a.peak_index || !a.peak_value) return fail(B2S_E_INVALID, "peak buffers are required");
  switch (n) {
    case 256: return launch_spectrum_t<256, MODE>(e, a, stream);
    case 512: return launch_spectrum_t<512, MODE>(e, a, stream);
    case 1024: return launch_spectrum_t<1024, MODE>(e, a, stream);
    case 2048: return launch_spectrum_t<2048, MODE>(e, a, stream);
    case 4096: return launch_spectrum_t<4096, MODE>(e, a, stream);
    case 8192: return launch_spectrum_t<8192, MODE>(e, a, stream);
    case 16384: return launch_spectrum_t<16384, MODE>(e, a, stream);
    case 32768: return launch_spectrum_t<32768, MODE>(e, a, stream);
    case 65536: return launch_spectrum_t<65536, MODE>(e, a, stream);
    case 131072: return launch_spectrum_t<131072, MODE>(e, a, stream);
    case 262144: return launch_spectrum_t<262144, MODE>(e, a, stream);
    default: return fail(B2S_E_INVALID, "fft_size %d is not supported (256..%d)", n, kMaxFft);
  }
}

int launch_spectrum(b2s_engine* e, int n, int iq_format, const SpectralArgs& a, cudaStream_t stream) {
  if (iq_format == B2S_IQ_CF32) return launch_spectrum_n<kModeCf32>(e, n, a, stream);
  const bool aligned = (reinterpret_cast<uintptr_t>(a.iq) % 16 == 0) && (a.frame_stride_bytes % 16 == 0);
  if (aligned) return launch_spectrum_n<kModeCs8Tma>(e, n, a, stream);
  return launch_spectrum_n<kModeCs8Direct>(e, n, a, stream);
}

template <int N>
void plan_radices_t(int* r) {
  r[0] = FftPlanT<N>::R0;
  r[1] = FftPlanT<N>::R1;
  r[2] = FftPlanT<N>::R2;
  r[3] = FftPlanT<N>::R3;
}
void plan_radices(int n, int* r) {
  switch (n) {
    case 256: plan_radices_t<256>(r); break;
    case 512: plan_radices_t<512>(r); break;
    case 1024: plan_radices_t<1024>(r); break;
    case 2048: plan_radices_t<2048>(r); break;
    case 4096: plan_radices_t<4096>(r); break;
    case 8192: plan_radices_t<8192>(r); break;
    default: plan_radices_t<2048>(r); break;  // larger sizes run k_spectrum3, which has its own tables
  }
}

struct SpectralTables {
  DevBuf<float> wscale;
  DevBuf<float2> twiddle, split_tw, split_ws;
  DevBuf<int> work_counter;  // k_spectrum3's {next item, finished CTAs}; the kernel leaves both at zero
  int split = 1;
  int build(const b2s_band_config& cfg) {
    const int n = cfg.fft_size;
    std::vector<float> w;
    make_window(cfg, w);
    if (cfg.iq_format == B2S_IQ_CS8) {
      // unpack scale folded into the window: x*scale*w -> x*(scale*w); differs from the two-step product by < 1 ulp
      for (int i = 0; i < n; ++i) w[i] = w[i] * cfg.iq_scale;
    }
    split = n > kSplitM ? n / kSplitM : 1;
    const int m = n / split;  // length of the transform the passes run (the whole FFT, or one residue class of it)
    std::vector<float2> tw;
    auto root = [](double num, double den) {
      const double ang = -2.0 * M_PI * num / den;
      return make_float2(static_cast<float>(std::cos(ang)), static_cast<float>(std::sin(ang)));
    };
    if (k1_is_v3(m)) {
      // k_spectrum3 (TwiddleLayout3): pass A  W_m^(b*k0) as [k0-1][b], b < 1024;  pass B  W_1024^(n2*k1) as [k1-1][n2]
      const int ra = m / 1024;
      for (int k0 = 1; k0 < ra; ++k0)
        for (int b = 0; b < 1024; ++b) tw.push_back(root(static_cast<double>(b) * k0, m));
      for (int k1 = 1; k1 < 32; ++k1)
        for (int n2 = 0; n2 < 32; ++n2) tw.push_back(root(static_cast<double>(n2) * k1, 1024.0));
    } else {
      // per-pass compact tables [m-1][k] = exp(-2 pi i k m / (P R)), interleaved (re, im), in the order of TwiddleLayout<N>
      int radix[4] = {0, 0, 0, 0};
      plan_radices(n, radix);
      int P = radix[0];
      for (int pass = 1; pass < 4 && radix[pass] > 1; ++pass) {
        const int R = radix[pass];
        for (int mm = 1; mm < R; ++mm)
          for (int k = 0; k < P; ++k) tw.push_back(root(static_cast<double>(k) * mm, static_cast<double>(P) * R));
        P *= R;
      }
    }
    int rc = wscale.alloc(n);
    if (rc) return rc;
    rc = twiddle.alloc(tw.size() + 1);
    if (rc) return rc;
    CU(cudaMemcpy(wscale.p, w.data(), sizeof(float) * n, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(twiddle.p, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice));
    if ((rc = work_counter.alloc(2))) return rc;
    CU(cudaMemset(work_counter.p, 0, sizeof(int) * 2));
    if (split > 1) {
      // split mode: class twiddles W_N^(n' c) as [c][n'] (n' c < N^2 fits a double exactly), and the S-th roots of unity
      std::vector<float2> tc(static_cast<size_t>(split) * m), ws(split);
      for (int c = 0; c < split; ++c)
        for (int i = 0; i < m; ++i) tc[static_cast<size_t>(c) * m + i] = root(std::fmod(static_cast<double>(i) * c, static_cast<double>(n)), n);
      for (int j = 0; j < split; ++j) ws[j] = root(j, split);
      if ((rc = split_tw.alloc(tc.size()))) return rc;
      if ((rc = split_ws.alloc(ws.size()))) return rc;
      CU(cudaMemcpy(split_tw.p, tc.data(), sizeof(float2) * tc.size(), cudaMemcpyHostToDevice));
      CU(cudaMemcpy(split_ws.p, ws.data(), sizeof(float2) * ws.size(), cudaMemcpyHostToDevice));
    }
    return 0;
  }
  // the table pointers of a K1 launch
  void fill(SpectralArgs& sa) const {
    sa.wscale = wscale.p;
    sa.twiddle = twiddle.p;
    sa.work_counter = work_counter.p;
    sa.split = split;
    sa.split_tw = split_tw.p;
    sa.split_ws = split_ws.p;
  }
  void release() {
    wscale.release();
    twiddle.release();
    split_tw.release();
    split_ws.release();
    work_counter.release();
  }
};

int validate_config(const b2s_band_config& c) {
  if (!is_pow2(c.fft_size) || c.fft_size < 256 || c.fft_size > kMaxFft) return fail(B2S_E_INVALID, "fft_size must be a power of two in 256..%d (got %d)", kMaxFft, c.fft_size);
  if (c.sample_rate_hz <= 0) return fail(B2S_E_INVALID, "sample_rate_hz must be positive");
  if (c.frame_stride_samples < c.fft_size) return fail(B2S_E_INVALID, "frame_stride_samples (%d) < fft_size", c.frame_stride_samples);
  if (c.iq_format != B2S_IQ_CS8 && c.iq_format != B2S_IQ_CF32) return fail(B2S_E_INVALID, "unknown iq_format %d", c.iq_format);
  if (c.window_kind == B2S_WINDOW_USER && !c.window_taps) return fail(B2S_E_INVALID, "window_taps is NULL");
  if (c.grouping_x < 1 || c.grouping_x > 65) return fail(B2S_E_INVALID, "grouping_x must be in 1..65");
  if (c.grouping_y < 1 || c.grouping_y > 256) return fail(B2S_E_INVALID, "grouping_y must be in 1..256");
  if (c.group_size_bins < 0 || c.group_size_bins > 4096) return fail(B2S_E_INVALID, "group_size_bins must be in 0..4096");
  if (c.learn_frames < 1) return fail(B2S_E_INVALID, "learn_frames must be >= 1");
  if (c.noise_learning_ms < 0) return fail(B2S_E_INVALID, "noise_learning_ms must be >= 0 (0 = count learn_frames frames)");
  if (c.n_ignored < 0 || c.n_ignored > B2S_MAX_IGNORED) return fail(B2S_E_INVALID, "n_ignored out of range");
  if (c.tuning_step_hz <= 0) return fail(B2S_E_INVALID, "tuning_step_hz must be positive");
  if (c.spectrogram_out_size < 0 || (c.spectrogram_out_size > 0 && (!is_pow2(c.spectrogram_out_size) || c.spectrogram_out_size > c.fft_size ||
                                                                   c.fft_size / c.spectrogram_out_size > kDetectBinsPerCta)))
    return fail(B2S_E_INVALID, "spectrogram_out_size must be 0 or a power of two with N/out <= %d", kDetectBinsPerCta);
  return 0;
}

}  // namespace

// 2-D tensor map over a row-major fp32 matrix [rows][cols] with box [box_rows][box_cols] (cuTensorMapEncodeTiled through the
// runtime's driver entry point: no link-time dependency on libcuda)
static int make_tile_map(CUtensorMap* out, const float* base, size_t cols, size_t rows, int box_cols, int box_rows) {
  using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return fail(B2S_E_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * sizeof(float)};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t elem[2] = {1, 1};
  CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  if (const char* e = getenv("B2S_K2_L2PROMO")) {  // A/B measurements
    const int v = atoi(e);
    promo = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : v == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : promo;
  }
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B2S_E_CUDA, "cuTensorMapEncodeTiled failed (%d) for box %dx%d", static_cast<int>(r), box_cols, box_rows);
  return 0;
}

// Least float x with fl(x / divisor) >= level (IEEE single division is monotonic in x, so the bins whose quotient reaches a level
// are exactly those whose undivided sum reaches this x): lets K2 compare boxcar SUMS and divide only what leaves the kernel.
static float least_sum_reaching(float level, int divisor) {
  if (!std::isfinite(level)) return level;
  const float d = static_cast<float>(divisor);
  float x = level * d;
  for (int i = 0; i < 64 && !(x / d >= level); ++i) x = std::nextafterf(x, INFINITY);
  for (int i = 0; i < 64 && std::nextafterf(x, -INFINITY) / d >= level; ++i) x = std::nextafterf(x, -INFINITY);
  return x;
}

// ------------------------------------------------------------------------------------------------------------
// band
// ------------------------------------------------------------------------------------------------------------
#include "band.cuh"

// ------------------------------------------------------------------------------------------------------------
// recorder chain (recorder.cuh)
// ------------------------------------------------------------------------------------------------------------
// new carry = the last `hc` elements of (old carry ++ fresh[0 .. n_new)); one CTA, read everything before writing anything
template <typename T>
__global__ void k_shift_carry(T* carry, const T* fresh, int hc, long long n_new) {
  extern __shared__ unsigned char carry_smem[];
  T* tmp = reinterpret_cast<T*>(carry_smem);
  for (int i = threadIdx.x; i < hc; i += blockDim.x) {
    const long long j = i + n_new;  // position in (old carry ++ fresh)
    tmp[i] = j < hc ? carry[j] : fresh[j - hc];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < hc; i += blockDim.x) carry[i] = tmp[i];
}

struct b2s_recorder {
  b2s_engine* engine = nullptr;
  int32_t sample_rate = 0, bandwidth = 0;
  int iq_format = B2S_IQ_CS8;
  float iq_scale = 1.0f;
  bool on_device = false, recording = false;
  size_t max_in = 0;
  unsigned long long phase_inc = 0;
  cudaStream_t stream = nullptr;
  struct Stage {
    int interp = 1, decim = 1, n_taps = 0, hc = 0;
    std::vector<float> h_taps;
    DevBuf<float> taps, taps_pq;  // taps_pq: [decim][kPolyQ] polyphase layout for k_decimate_poly (decimating stages), else empty
    DevBuf<float2> buf;      // stages >= 1: [hc carried samples | the previous stage's outputs of this push]
    long long consumed = 0;  // input samples seen since startRecording
    long long produced = 0;  // output samples produced since startRecording
    size_t max_in = 0;
  };
  std::vector<Stage> stages;
  DevBuf<unsigned char> carry_raw, staging;  // stage 0: carried raw samples; host input staging
  DevBuf<signed char> d_out;
  ~b2s_recorder() {
    for (auto& st : stages) {
      st.taps.release();
      st.taps_pq.release();
      st.buf.release();
    }
    carry_raw.release();
    staging.release();
    d_out.release();
    if (stream) cudaStreamDestroy(stream);
  }
  size_t raw_bytes() const { return iq_format == B2S_IQ_CS8 ? 2 : 8; }
};

// ------------------------------------------------------------------------------------------------------------
// stand-alone device Averager
// ------------------------------------------------------------------------------------------------------------
struct b2s_averager {
  b2s_engine* engine = nullptr;
  int size = 0, group = 0, frames = 0, cur = 0;
  DevBuf<float> sum, ring[2], avg, rows;
  ~b2s_averager() {
    sum.release();
    ring[0].release();
    ring[1].release();
    avg.release();
    rows.release();
  }
  int reset() {
    CU(cudaMemset(sum.p, 0, sizeof(float) * size));
    CU(cudaMemset(ring[0].p, 0, sizeof(float) * size * group));
    CU(cudaMemset(ring[1].p, 0, sizeof(float) * size * group));
    std::vector<float> nd(size, kNoData);
    CU(cudaMemcpy(avg.p, nd.data(), sizeof(float) * size, cudaMemcpyHostToDevice));
    frames = 0;
    cur = 0;
    return 0;
  }
};

// ---- self-test of the exact constant division used on the Averager / boxcar fast paths (detect.cuh: div_const) ----
template <int D>
static int run_div_check(unsigned long long* d_bad) {
  k_check_div_const<D><<<148 * 8, 256>>>(d_bad);
  CU(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------------------
extern "C" {

const char* b2s_last_error(void) { return g_error.c_str(); }
int b2s_version(void) { return B2S_VERSION; }

int b2s_engine_create(int cuda_device, b2s_engine** out) {
  if (!out) return fail(B2S_E_INVALID, "out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t err = cudaGetDeviceCount(&count);
  if (err != cudaSuccess || count == 0) return fail(B2S_E_CUDA, "no usable CUDA device (%s); this engine has no CPU fallback", cudaGetErrorString(err));
  if (cuda_device < 0 || cuda_device >= count) return fail(B2S_E_INVALID, "cuda_device %d out of range (0..%d)", cuda_device, count - 1);
  b2s_engine* e = new b2s_engine();
  e->device = cuda_device;
  CU(cudaSetDevice(cuda_device));
  CU(cudaGetDeviceProperties(&e->prop, cuda_device));
  if (e->prop.major < 10) {
    const int major = e->prop.major, minor = e->prop.minor;
    delete e;
    return fail(B2S_E_CUDA, "device is sm_%d%d; this build targets sm_100a (B200) only", major, minor);
  }
  e->sm_count = e->prop.multiProcessorCount;
  *out = e;
  return 0;
}
int b2s_engine_destroy(b2s_engine* e) {
  delete e;
  return 0;
}
int b2s_engine_device_name(b2s_engine* e, char* buf, size_t cap) {
  if (!e || !buf || cap == 0) return fail(B2S_E_INVALID, "bad argument");
  snprintf(buf, cap, "%s (%d SMs)", e->prop.name, e->sm_count);
  return 0;
}

void b2s_default_config(b2s_band_config* cfg, int32_t sample_rate_hz, int32_t center_hz, int32_t recording_bandwidth_hz) {
  std::memset(cfg, 0, sizeof(*cfg));
  const int n = host::fft_size_for(sample_rate_hz, 250);  // SIGNAL_DETECTION_MAX_STEP, config.h:33
  const double step = static_cast<double>(sample_rate_hz) / n;
  cfg->fft_size = n;
  cfg->sample_rate_hz = sample_rate_hz;
  cfg->frame_stride_samples = n * host::decimator_factor(sample_rate_hz, n);
  cfg->iq_format = B2S_IQ_CS8;
  cfg->iq_scale = 1.0f / 127.0f;
  cfg->window_kind = B2S_WINDOW_HAMMING;
  cfg->grouping_x = 21;
  cfg->grouping_y = 21;
  cfg->group_size_bins = static_cast<int32_t>(std::ceil(recording_bandwidth_hz / step));  // sdr_device.cpp:151
  cfg->start_level = 8.0f;
  cfg->stop_level = 5.0f;
  const double period = static_cast<double>(cfg->frame_stride_samples) * 1000.0 / sample_rate_hz;
  cfg->learn_frames = b2s_learn_frames_from_ms(2000, period);
  cfg->noise_learning_ms = 2000;  // NOISE_LEARNING_TIME (config.h:24): the reference's wall-clock rule on the frame clock
  cfg->center_hz = center_hz;
  cfg->range_lo_hz = center_hz - sample_rate_hz / 2;
  cfg->range_hi_hz = center_hz + sample_rate_hz / 2;
  cfg->tuning_step_hz = 2500;
  cfg->min_time_ms = 2000;
  cfg->timeout_ms = 2000;
  cfg->max_time_ms = 600000;
  cfg->spectrogram_out_size = std::min(n, std::min(16384, host::fft_size_for(sample_rate_hz, 1000)));
  cfg->spectrogram_interval_ms = 1000;
}

int b2s_band_create(b2s_engine* e, const b2s_band_config* cfg, b2s_band** out) {
  if (!e || !cfg || !out) return fail(B2S_E_INVALID, "NULL argument");
  *out = nullptr;
  int rc = validate_config(*cfg);
  if (rc) return rc;
  b2s_band* b = new b2s_band();
  rc = b->init(e, *cfg);
  if (rc) {
    delete b;
    return rc;
  }
  *out = b;
  return 0;
}
int b2s_band_destroy(b2s_band* b) {
  if (b) {
    cudaSetDevice(b->engine->device);
    delete b;
  }
  return 0;
}
int b2s_band_set_stream(b2s_band* b, void* cuda_stream) {
  if (!b) return fail(B2S_E_INVALID, "NULL band");
  std::lock_guard<std::mutex> lock(b->mutex);
  int rc = b->drain();
  if (rc) return rc;
  b->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : b->own_stream;
  return 0;
}

int b2s_band_push(b2s_band* b, const void* iq, size_t n_frames, int64_t t0_ms, double frame_period_ms, b2s_result* out) {
  if (!b || (!iq && n_frames)) return fail(B2S_E_INVALID, "NULL argument");
  std::lock_guard<std::mutex> lock(b->mutex);
  CU(cudaSetDevice(b->engine->device));
  if (b->async_mode && out) return fail(B2S_E_INVALID, "with B2S_FLAG_ASYNC results are collected by b2s_band_sync; pass out = NULL to b2s_band_push");
  if (out) {
    out->n_transmissions = 0;
    out->n_transmissions_total = 0;
    out->n_detect_entries = 0;
    out->n_spectrogram_rows = 0;
  }
  b->prof.pushes += 1;
  {
    int rc = b->grow_capacity();  // a previous push overflowed its per-frame entry lists
    if (rc) return rc;
  }
  const size_t bytes_per_sample = b->cfg.iq_format == B2S_IQ_CS8 ? 2 : 8;
  const size_t stride_bytes = static_cast<size_t>(b->cfg.frame_stride_samples) * bytes_per_sample;
  const bool on_device = (b->cfg.flags & B2S_FLAG_IQ_ON_DEVICE) != 0;
  if (on_device) {
    for (size_t done = 0; done < n_frames;) {
      const size_t chunk = std::min(n_frames - done, static_cast<size_t>(b->max_frames));
      int rc = b->push_chunk(static_cast<const char*>(iq) + done * stride_bytes, chunk, t0_ms, frame_period_ms, done, out);
      if (rc) return rc;
      done += chunk;
    }
    return 0;
  }
  // Host input: the push is cut into pipeline chunks; the host->device copy of chunk i+1 runs on a second stream
  // while chunk i is in the kernels / tracker (double-buffered staging), so PCIe time hides behind compute (or vice versa).
  // In async mode one chunk per push is enough: the copy of push k+1 already overlaps the work of push k.
  const size_t pipe = b->async_mode ? std::min<size_t>(b->max_frames, std::max<size_t>(n_frames, 1))
                                    : std::max<size_t>(1, std::min<size_t>(b->max_frames, n_frames >= 512 ? (n_frames + 3) / 4 : n_frames));
  if (!b->copy_stream) {
    CU(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&b->copy_done[0], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&b->copy_done[1], cudaEventDisableTiming));
  }
  const size_t buf_bytes = pipe * stride_bytes;
  for (int i = 0; i < 2; ++i) {
    int rc = b->d_iq[i].alloc(buf_bytes);
    if (rc) return rc;
  }
  auto chunk_len = [&](size_t done) { return std::min(n_frames - done, pipe); };
  auto start_copy = [&](size_t done, int slot) -> int {
    const size_t chunk = chunk_len(done);
    const size_t bytes = (chunk - 1) * stride_bytes + static_cast<size_t>(b->cfg.fft_size) * bytes_per_sample;  // the last frame needs N samples only
    CU(cudaMemcpyAsync(b->d_iq[slot].p, static_cast<const char*>(iq) + done * stride_bytes, bytes, cudaMemcpyHostToDevice, b->copy_stream));
    CU(cudaEventRecord(b->copy_done[slot], b->copy_stream));
    b->prof.h2d_bytes += bytes;
    return 0;
  };
  if (b->async_mode) {
    // staging buffer iq_slot was last read by the K1 of the chunk two chunks ago; iq_prev_use[slot] marks the end of that
    // chunk's kernels on b->stream, and the copy stream waits for it before overwriting the buffer
    for (size_t done = 0; done < n_frames;) {
      const size_t chunk = chunk_len(done);
      const int slot = b->iq_slot;
      CU(cudaStreamWaitEvent(b->copy_stream, b->iq_prev_use[slot], 0));
      int rc = start_copy(done, slot);
      if (rc) return rc;
      CU(cudaStreamWaitEvent(b->stream, b->copy_done[slot], 0));
      rc = b->push_chunk(b->d_iq[slot].p, chunk, t0_ms, frame_period_ms, done, nullptr);
      if (rc) return rc;
      CU(cudaEventRecord(b->iq_prev_use[slot], b->stream));  // K1 (and the rest) of this chunk: the buffer may be overwritten after it
      CU(cudaStreamSynchronize(b->copy_stream));             // the caller may reuse `iq` as soon as the call returns
      b->iq_slot ^= 1;
      done += chunk;
    }
    return 0;
  }
  int slot = 0;
  if (n_frames > 0) {
    int rc = start_copy(0, 0);
    if (rc) return rc;
  }
  for (size_t done = 0; done < n_frames;) {
    const size_t chunk = chunk_len(done);
    CU(cudaStreamWaitEvent(b->stream, b->copy_done[slot], 0));
    if (done + chunk < n_frames) {  // the other staging buffer was released when the previous chunk finished (push_chunk is synchronous)
      int rc = start_copy(done + chunk, slot ^ 1);
      if (rc) return rc;
    }
    int rc = b->push_chunk(b->d_iq[slot].p, chunk, t0_ms, frame_period_ms, done, out);
    if (rc) return rc;
    done += chunk;
    slot ^= 1;
  }
  return 0;
}

int b2s_band_sync(b2s_band* b, b2s_result* out) {
  if (!b) return fail(B2S_E_INVALID, "NULL band");
  std::lock_guard<std::mutex> lock(b->mutex);
  CU(cudaSetDevice(b->engine->device));
  int rc = b->drain();
  if (rc) return rc;
  if (out) {
    out->n_transmissions_total = static_cast<int32_t>(b->mailbox.size());
    out->n_transmissions = std::min<int32_t>(out->n_transmissions_total, B2S_MAX_TX);
    std::memcpy(out->transmissions, b->mailbox.data(), sizeof(b2s_transmission) * out->n_transmissions);
    out->n_detect_entries = b->stat_entries;
    out->n_spectrogram_rows = b->stat_rows;
  }
  b->stat_entries = 0;
  b->stat_rows = 0;
  return 0;
}

int b2s_band_set_profiling(b2s_band* b, int enable) {
  if (!b) return fail(B2S_E_INVALID, "NULL band");
  std::lock_guard<std::mutex> lock(b->mutex);
  CU(cudaSetDevice(b->engine->device));
  int rc = b->drain();
  if (rc) return rc;
  b->profiling = enable != 0;
  b->profile_ctas = enable >= 2;
  return 0;
}
int b2s_band_get_profile(b2s_band* b, b2s_profile* out, int reset) {
  if (!b || !out) return fail(B2S_E_INVALID, "NULL argument");
  std::lock_guard<std::mutex> lock(b->mutex);
  int rc = b->drain();
  if (rc) return rc;
  *out = b->prof;
  if (reset) b->prof = b2s_profile{};
  return 0;
}

int b2s_band_reset(b2s_band* b) {
  if (!b) return fail(B2S_E_INVALID, "NULL band");
  std::lock_guard<std::mutex> lock(b->mutex);
  CU(cudaSetDevice(b->engine->device));
  return b->reset_buffers();
}

int b2s_band_set_center(b2s_band* b, int32_t center_hz, int32_t lo, int32_t hi) {
  if (!b) return fail(B2S_E_INVALID, "NULL band");
  std::lock_guard<std::mutex> lock(b->mutex);
  int rc = b->drain();
  if (rc) return rc;
  b->center = center_hz;
  b->tracker.p.center = center_hz;
  b->tracker.p.range_lo = lo;
  b->tracker.p.range_hi = hi;
  return 0;
}

int b2s_band_get_averager(b2s_band* b, float* sum, float* avg, float* ring, int32_t* frames) {
  if (!b) return fail(B2S_E_INVALID, "NULL band");
  std::lock_guard<std::mutex> lock(b->mutex);
  CU(cudaSetDevice(b->engine->device));
  int rc = b->drain();
  if (rc) return rc;
  CU(cudaStreamSynchronize(b->stream));
  const size_t n = b->cfg.fft_size, Y = b->cfg.grouping_y;
  if (sum) CU(cudaMemcpy(sum, b->d_sum[b->sum_cur].p, sizeof(float) * n, cudaMemcpyDeviceToHost));
  if (avg) CU(cudaMemcpy(avg, b->d_avg_last.p, sizeof(float) * n, cudaMemcpyDeviceToHost));
  if (ring) CU(cudaMemcpy(ring, b->d_ring[b->ring_cur].p, sizeof(float) * n * Y, cudaMemcpyDeviceToHost));
  if (frames) *frames = b->avg_frames;
  return 0;
}

int b2s_band_get_noise(b2s_band* b, float* threshold, int32_t* samples, int32_t* ready) {
  if (!b) return fail(B2S_E_INVALID, "NULL band");
  std::lock_guard<std::mutex> lock(b->mutex);
  CU(cudaSetDevice(b->engine->device));
  int rc = b->drain();
  if (rc) return rc;
  CU(cudaStreamSynchronize(b->stream));
  auto it = b->noise.find(b->center);
  if (it == b->noise.end()) {
    if (samples) *samples = 0;
    if (ready) *ready = 0;
    if (threshold) {
      for (int i = 0; i < b->cfg.fft_size; ++i) threshold[i] = -std::numeric_limits<float>::max();
    }
    return 0;
  }
  if (threshold) CU(cudaMemcpy(threshold, it->second.now(), sizeof(float) * b->cfg.fft_size, cudaMemcpyDeviceToHost));
  if (samples) *samples = it->second.samples;
  if (ready) *ready = it->second.ready ? 1 : 0;
  return 0;
}

int b2s_band_get_spectrogram(b2s_band* b, int64_t* times, int32_t* centers, int8_t* rows, int cap, int consume, int* count) {
  if (!b || !count) return fail(B2S_E_INVALID, "NULL argument");
  std::lock_guard<std::mutex> lock(b->mutex);
  int rc = b->drain();
  if (rc) return rc;
  const int M = b->cfg.spectrogram_out_size;
  const int total = static_cast<int>(b->sent.size());
  for (int i = 0; i < total && i < cap; ++i) {
    if (times) times[i] = b->sent[i].time;
    if (centers) centers[i] = b->sent[i].center;
    if (rows) std::memcpy(rows + static_cast<size_t>(i) * M, b->sent[i].row.data(), M);
  }
  *count = total;
  if (consume) b->sent.erase(b->sent.begin(), b->sent.begin() + std::min(total, std::max(cap, 0)));  // only the rows handed out
  return 0;
}

int b2s_band_get_transmissions(b2s_band* b, b2s_transmission* out, int cap, int* count) {
  if (!b || !count) return fail(B2S_E_INVALID, "NULL argument");
  std::lock_guard<std::mutex> lock(b->mutex);
  int rc = b->drain();
  if (rc) return rc;
  const int total = static_cast<int>(b->mailbox.size());
  if (out) std::memcpy(out, b->mailbox.data(), sizeof(b2s_transmission) * std::max(0, std::min(cap, total)));
  *count = total;
  return 0;
}

int b2s_band_get_signals(b2s_band* b, int32_t* keys, int64_t* first, int64_t* last, float* power, int cap, int* count) {
  if (!b || !count) return fail(B2S_E_INVALID, "NULL argument");
  std::lock_guard<std::mutex> lock(b->mutex);
  CU(cudaSetDevice(b->engine->device));
  int rc = b->drain();
  if (rc) return rc;
  std::vector<TrackState> h(1);  // the map is device resident (K4)
  if ((rc = b->download_state(h[0]))) return rc;
  for (int i = 0; i < h[0].n && i < cap; ++i) {
    if (keys) keys[i] = h[0].key[i];
    if (first) first[i] = h[0].first[i];
    if (last) last[i] = h[0].last[i];
    if (power) power[i] = h[0].power[i];
  }
  *count = h[0].n;
  return 0;
}

// ---- stand-alone operators ----
int b2s_averager_create(b2s_engine* e, int size, int group_size, b2s_averager** out) {
  if (!e || !out || size < 1 || group_size < 1) return fail(B2S_E_INVALID, "bad argument");
  CU(cudaSetDevice(e->device));
  b2s_averager* a = new b2s_averager();
  a->engine = e;
  a->size = size;
  a->group = group_size;
  int rc = a->sum.alloc(size);
  if (!rc) rc = a->ring[0].alloc(static_cast<size_t>(size) * group_size);
  if (!rc) rc = a->ring[1].alloc(static_cast<size_t>(size) * group_size);
  if (!rc) rc = a->avg.alloc(size);
  if (!rc) rc = a->reset();
  if (rc) {
    delete a;
    return rc;
  }
  *out = a;
  return 0;
}
int b2s_averager_destroy(b2s_averager* a) {
  delete a;
  return 0;
}
int b2s_averager_push_many(b2s_averager* a, const float* rows, int count) {
  if (!a || !rows || count < 0) return fail(B2S_E_INVALID, "bad argument");
  if (count == 0) return 0;
  CU(cudaSetDevice(a->engine->device));
  int rc = a->rows.alloc(static_cast<size_t>(count) * a->size);
  if (rc) return rc;
  CU(cudaMemcpy(a->rows.p, rows, sizeof(float) * count * a->size, cudaMemcpyHostToDevice));
  k_averager_push<<<(a->size + 127) / 128, 128>>>(a->rows.p, count, a->size, a->group, a->sum.p, a->ring[a->cur].p, a->ring[a->cur ^ 1].p, a->frames, a->avg.p);
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  a->cur ^= 1;
  a->frames = std::min(a->frames + count, a->group);
  return 0;
}
int b2s_averager_push(b2s_averager* a, const float* data) { return b2s_averager_push_many(a, data, 1); }
int b2s_averager_reset(b2s_averager* a) {
  if (!a) return fail(B2S_E_INVALID, "NULL averager");
  CU(cudaSetDevice(a->engine->device));
  return a->reset();
}
int b2s_averager_average(b2s_averager* a, float* out) {
  if (!a || !out) return fail(B2S_E_INVALID, "bad argument");
  CU(cudaMemcpy(out, a->avg.p, sizeof(float) * a->size, cudaMemcpyDeviceToHost));
  return 0;
}
int b2s_averager_data(b2s_averager* a, float* out) {
  if (!a || !out) return fail(B2S_E_INVALID, "bad argument");
  CU(cudaMemcpy(out, a->ring[a->cur].p, sizeof(float) * a->size * a->group, cudaMemcpyDeviceToHost));
  return 0;
}
int b2s_averager_sum(b2s_averager* a, float* out, int32_t* frames) {
  if (!a) return fail(B2S_E_INVALID, "bad argument");
  if (out) CU(cudaMemcpy(out, a->sum.p, sizeof(float) * a->size, cudaMemcpyDeviceToHost));
  if (frames) *frames = a->frames;
  return 0;
}

int b2s_average(b2s_engine* e, const float* in, float* out, int size, int group_size, int rows, int exact) {
  if (!e || !in || !out || size < 1 || group_size < 1 || rows < 1) return fail(B2S_E_INVALID, "bad argument");
  CU(cudaSetDevice(e->device));
  DevBuf<float> din, dout;
  int rc = din.alloc(static_cast<size_t>(size) * rows);
  if (!rc) rc = dout.alloc(static_cast<size_t>(size) * rows);
  if (rc) {
    din.release();
    dout.release();
    return rc;
  }
  cudaError_t err = cudaMemcpy(din.p, in, sizeof(float) * size * rows, cudaMemcpyHostToDevice);
  if (err == cudaSuccess) {
    if (exact) {
      k_boxcar_serial<<<(rows + 31) / 32, 32>>>(din.p, dout.p, size, group_size, rows);
    } else {
      dim3 grid((size + 127) / 128, rows);
      k_boxcar<<<grid, 128>>>(din.p, dout.p, size, group_size, rows);
    }
    err = cudaGetLastError();
  }
  if (err == cudaSuccess) err = cudaMemcpy(out, dout.p, sizeof(float) * size * rows, cudaMemcpyDeviceToHost);
  din.release();
  dout.release();
  if (err != cudaSuccess) return fail(B2S_E_CUDA, "b2s_average: %s", cudaGetErrorString(err));
  return 0;
}

int b2s_psd(b2s_engine* e, const b2s_band_config* cfg, const void* iq, size_t n_frames, float* psd_db, float* power_lin) {
  if (!e || !cfg || !iq || !psd_db || n_frames == 0) return fail(B2S_E_INVALID, "bad argument");
  b2s_band_config c = *cfg;
  if (c.learn_frames < 1) c.learn_frames = 1;
  if (c.grouping_x < 1) c.grouping_x = 1;
  if (c.grouping_y < 1) c.grouping_y = 1;
  if (c.tuning_step_hz < 1) c.tuning_step_hz = 1;
  int rc = validate_config(c);
  if (rc) return rc;
  CU(cudaSetDevice(e->device));
  SpectralTables tables;
  DevBuf<unsigned char> diq;
  DevBuf<float> dpsd, dlin, dpv;
  DevBuf<int> dpi;
  DevBuf<unsigned long long> dpacked;
  const size_t n = c.fft_size;
  const size_t bps = c.iq_format == B2S_IQ_CS8 ? 2 : 8;
  const size_t stride = static_cast<size_t>(c.frame_stride_samples) * bps;
  const size_t bytes = (n_frames - 1) * stride + n * bps;
  rc = tables.build(c);
  if (!rc) rc = diq.alloc(bytes);
  if (!rc) rc = dpsd.alloc(n_frames * n);
  if (!rc && power_lin) rc = dlin.alloc(n_frames * n);
  if (!rc) rc = dpv.alloc(n_frames);
  if (!rc) rc = dpi.alloc(n_frames);
  if (!rc && tables.split > 1) rc = dpacked.alloc(n_frames);
  cudaError_t err = cudaSuccess;
  if (!rc) err = cudaMemcpy(diq.p, iq, bytes, cudaMemcpyHostToDevice);
  if (!rc && err == cudaSuccess) {
    SpectralArgs sa{};
    sa.iq = diq.p;
    sa.frame_stride_bytes = static_cast<long long>(stride);
    sa.n_frames = static_cast<int>(n_frames);
    tables.fill(sa);
    sa.inv_fs = 1.0f / static_cast<float>(c.sample_rate_hz);
    sa.psd_db = dpsd.p;
    sa.power_lin = power_lin ? dlin.p : nullptr;
    sa.peak_index = dpi.p;
    sa.peak_value = dpv.p;
    sa.peak_packed = dpacked.p;
    rc = launch_spectrum(e, c.fft_size, c.iq_format, sa, nullptr);
    if (!rc) err = cudaDeviceSynchronize();
    if (!rc && err == cudaSuccess) err = cudaMemcpy(psd_db, dpsd.p, sizeof(float) * n_frames * n, cudaMemcpyDeviceToHost);
    if (!rc && err == cudaSuccess && power_lin) err = cudaMemcpy(power_lin, dlin.p, sizeof(float) * n_frames * n, cudaMemcpyDeviceToHost);
  }
  tables.release();
  diq.release();
  dpsd.release();
  dlin.release();
  dpv.release();
  dpi.release();
  dpacked.release();
  if (rc) return rc;
  if (err != cudaSuccess) return fail(B2S_E_CUDA, "b2s_psd: %s", cudaGetErrorString(err));
  return 0;
}

// ---- scan policy: Scanner's hop rule + SdrDevice::updateRecordings (scan_policy.h) ----
struct b2s_scan_policy {
  host::ScanPolicy policy;
  b2s_scan_policy(const int32_t* lo, const int32_t* hi, int n, int32_t fs, int rec, int64_t t) : policy(lo, hi, n, fs, rec, t) {}
};
int32_t b2s_get_range_split_sample_rate(int32_t sample_rate_hz) { return host::range_split_sample_rate(sample_rate_hz); }
int b2s_scan_policy_create(const int32_t* lo, const int32_t* hi, int n_ranges, int32_t sample_rate_hz, int n_recorders, int64_t scanning_time_ms, b2s_scan_policy** out) {
  if (!out || n_ranges < 0 || (n_ranges && (!lo || !hi)) || sample_rate_hz <= 0 || n_recorders < 0) return fail(B2S_E_INVALID, "b2s_scan_policy_create: bad argument");
  *out = new b2s_scan_policy(lo, hi, n_ranges, sample_rate_hz, n_recorders, scanning_time_ms > 0 ? scanning_time_ms : 500);
  return 0;
}
int b2s_scan_policy_destroy(b2s_scan_policy* p) {
  delete p;
  return 0;
}
int b2s_scan_policy_ranges(b2s_scan_policy* p, int32_t* lo, int32_t* hi, int cap) {
  if (!p) return fail(B2S_E_INVALID, "NULL policy");
  const auto& r = p->policy.ranges;
  for (size_t i = 0; i < r.size() && static_cast<int>(i) < cap; ++i) {
    if (lo) lo[i] = r[i].first;
    if (hi) hi[i] = r[i].second;
  }
  return static_cast<int>(r.size());
}
int b2s_scan_policy_begin(b2s_scan_policy* p, int64_t now_ms, int32_t* lo, int32_t* hi) {
  if (!p || p->policy.ranges.empty()) return fail(B2S_E_INVALID, "b2s_scan_policy_begin: no ranges to scan");
  p->policy.current = 0;
  p->policy.start = now_ms;
  if (lo) *lo = p->policy.ranges[0].first;
  if (hi) *hi = p->policy.ranges[0].second;
  return 0;
}
int b2s_scan_policy_notify(b2s_scan_policy* p, int64_t now_ms, const b2s_transmission* list, int n, b2s_recorder_action* actions, int cap, int* n_actions, int* hop,
                           int32_t* next_lo, int32_t* next_hi) {
  if (!p || n < 0 || (n && !list) || !n_actions || !hop) return fail(B2S_E_INVALID, "b2s_scan_policy_notify: bad argument");
  std::vector<b2s_recorder_action> acts;
  p->policy.update_recordings(now_ms, list, n, acts);
  for (size_t i = 0; i < acts.size() && static_cast<int>(i) < cap; ++i) actions[i] = acts[i];
  *n_actions = static_cast<int>(acts.size());
  *hop = p->policy.dwell_over(now_ms, n == 0) ? 1 : 0;
  if (*hop) {
    p->policy.hop(now_ms);
    if (next_lo) *next_lo = p->policy.ranges[p->policy.current].first;
    if (next_hi) *next_hi = p->policy.ranges[p->policy.current].second;
  }
  return 0;
}

// ---- recorder chain: rotate -> rational resamplers -> int8 (sources/radio/recorder.cpp:22-40,58-73) ----
int b2s_get_resamplers_factors(int32_t sample_rate_hz, int32_t bandwidth_hz, int threshold, int32_t* interp, int32_t* decim, int cap) {
  if (sample_rate_hz <= 0 || bandwidth_hz <= 0 || threshold < 1) return fail(B2S_E_INVALID, "b2s_get_resamplers_factors: bad argument");
  const auto f = host::resamplers_factors(sample_rate_hz, bandwidth_hz, threshold);
  for (size_t i = 0; i < f.size() && static_cast<int>(i) < cap; ++i) {
    if (interp) interp[i] = f[i].first;
    if (decim) decim[i] = f[i].second;
  }
  return static_cast<int>(f.size());
}
int b2s_recorder_create(b2s_engine* e, int32_t sample_rate_hz, int32_t bandwidth_hz, int iq_format, float iq_scale, int flags, size_t max_samples_per_push, b2s_recorder** out) {
  if (!e || !out || sample_rate_hz <= 0 || bandwidth_hz <= 0 || bandwidth_hz > sample_rate_hz) return fail(B2S_E_INVALID, "b2s_recorder_create: bad argument");
  if (iq_format != B2S_IQ_CS8 && iq_format != B2S_IQ_CF32) return fail(B2S_E_INVALID, "unknown iq_format %d", iq_format);
  *out = nullptr;
  CU(cudaSetDevice(e->device));
  auto* r = new b2s_recorder();
  r->engine = e;
  r->sample_rate = sample_rate_hz;
  r->bandwidth = bandwidth_hz;
  r->iq_format = iq_format;
  r->iq_scale = iq_scale;
  r->on_device = (flags & B2S_FLAG_IQ_ON_DEVICE) != 0;
  r->max_in = max_samples_per_push ? max_samples_per_push : (size_t(1) << 22);
  int rc = 0;
  cudaError_t err = cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking);
  if (err != cudaSuccess) rc = fail(B2S_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(err));
  size_t n_in = r->max_in;
  for (const auto& f : host::resamplers_factors(sample_rate_hz, bandwidth_hz, 125)) {  // RESAMPLER_THRESHOLD, config.h
    if (rc) break;
    r->stages.emplace_back();
    auto& st = r->stages.back();
    st.interp = f.first;
    st.decim = f.second;
    st.h_taps = host::design_resampler_taps(st.interp, st.decim);
    st.n_taps = static_cast<int>(st.h_taps.size());
    st.hc = (st.n_taps - 1 + st.decim) / st.interp + 2;
    st.max_in = n_in;
    if (st.hc > 4096) rc = fail(B2S_E_INVALID, "resampler stage %d/%d needs %d samples of history", st.interp, st.decim, st.hc);
    if (!rc) rc = st.taps.alloc(st.n_taps);
    if (!rc && cudaMemcpy(st.taps.p, st.h_taps.data(), sizeof(float) * st.n_taps, cudaMemcpyHostToDevice) != cudaSuccess) rc = fail(B2S_E_CUDA, "taps upload failed");
    if (!rc && st.interp == 1 && st.decim >= 2 && (st.n_taps + st.decim - 1) / st.decim <= kPolyQ && !getenv("B2S_RECORDER_GENERIC")) {
      std::vector<float> pq(static_cast<size_t>(st.decim) * kPolyQ, 0.0f);  // h[q D + p] at [p][q], zero padded
      for (int k = 0; k < st.n_taps; ++k) pq[static_cast<size_t>(k % st.decim) * kPolyQ + k / st.decim] = st.h_taps[k];
      rc = st.taps_pq.alloc(pq.size());
      if (!rc && cudaMemcpy(st.taps_pq.p, pq.data(), sizeof(float) * pq.size(), cudaMemcpyHostToDevice) != cudaSuccess) rc = fail(B2S_E_CUDA, "taps upload failed");
    }
    if (!rc && r->stages.size() > 1) rc = st.buf.alloc(st.hc + n_in + 2);
    n_in = (n_in * st.interp) / st.decim + 2;
  }
  if (!rc) rc = r->carry_raw.alloc(static_cast<size_t>(r->stages[0].hc) * r->raw_bytes());
  if (!rc) rc = r->d_out.alloc(2 * n_in);
  if (!rc && !r->on_device) rc = r->staging.alloc(r->max_in * r->raw_bytes());
  if (rc) {
    delete r;
    return rc;
  }
  *out = r;
  return 0;
}
int b2s_recorder_destroy(b2s_recorder* r) {
  if (r) {
    cudaSetDevice(r->engine->device);
    delete r;
  }
  return 0;
}
int b2s_recorder_stages(b2s_recorder* r, int32_t* interp, int32_t* decim, int32_t* n_taps, int cap) {
  if (!r) return fail(B2S_E_INVALID, "NULL recorder");
  for (size_t i = 0; i < r->stages.size() && static_cast<int>(i) < cap; ++i) {
    if (interp) interp[i] = r->stages[i].interp;
    if (decim) decim[i] = r->stages[i].decim;
    if (n_taps) n_taps[i] = r->stages[i].n_taps;
  }
  return static_cast<int>(r->stages.size());
}
int b2s_recorder_taps(b2s_recorder* r, int stage, float* taps, int cap) {
  if (!r || stage < 0 || stage >= static_cast<int>(r->stages.size()) || !taps) return fail(B2S_E_INVALID, "b2s_recorder_taps: bad argument");
  const auto& h = r->stages[stage].h_taps;
  std::memcpy(taps, h.data(), sizeof(float) * std::min<size_t>(h.size(), std::max(cap, 0)));
  return static_cast<int>(h.size());
}
// Recorder::startRecording (recorder.cpp:58-73): set the rotator to -shift, start from empty buffers
int b2s_recorder_start(b2s_recorder* r, int32_t shift_hz) {
  if (!r) return fail(B2S_E_INVALID, "NULL recorder");
  // phase increment per sample in turns: -shift / fs, as a 64-bit binary fraction (exact to 2^-65 turns)
  const long double turns = -static_cast<long double>(shift_hz) / static_cast<long double>(r->sample_rate);
  long double frac = turns - floorl(turns);  // [0, 1)
  r->phase_inc = static_cast<unsigned long long>(frac * 18446744073709551616.0L + 0.5L);
  if (shift_hz != 0 && r->phase_inc == 0) r->phase_inc = 1;
  for (auto& st : r->stages) st.consumed = st.produced = 0;
  r->recording = true;
  return 0;
}
int b2s_recorder_stop(b2s_recorder* r) {
  if (!r) return fail(B2S_E_INVALID, "NULL recorder");
  r->recording = false;
  return 0;
}
int b2s_recorder_push(b2s_recorder* r, const void* iq, size_t n_samples, int8_t* out_iq, size_t cap_samples, size_t* n_out) {
  if (!r || (!iq && n_samples) || !n_out) return fail(B2S_E_INVALID, "b2s_recorder_push: NULL argument");
  *n_out = 0;
  if (!r->recording) return fail(B2S_E_STATE, "b2s_recorder_push: the recorder is not recording (b2s_recorder_start)");
  if (n_samples > r->max_in) return fail(B2S_E_INVALID, "b2s_recorder_push: %zu samples exceed max_samples_per_push %zu", n_samples, r->max_in);
  if (n_samples == 0) return 0;
  CU(cudaSetDevice(r->engine->device));
  const void* in = iq;
  if (!r->on_device) {
    CU(cudaMemcpyAsync(r->staging.p, iq, n_samples * r->raw_bytes(), cudaMemcpyHostToDevice, r->stream));
    in = r->staging.p;
  }
  long long n_in = static_cast<long long>(n_samples);
  const void* cur_in = in;
  const void* cur_carry = r->carry_raw.p;
  int kind = r->iq_format == B2S_IQ_CS8 ? 0 : 1;
  long long produced_last = 0;
  for (size_t si = 0; si < r->stages.size(); ++si) {
    auto& st = r->stages[si];
    const bool last = si + 1 == r->stages.size();
    const long long g0 = st.consumed, g1 = g0 + n_in;
    const long long m_end = (g1 * st.interp - 1) / st.decim + 1;  // outputs whose newest input has arrived
    const long long n_new = m_end - st.produced;
    ResampleArgs a{};
    a.in = cur_in;
    a.carry = cur_carry;
    a.kind = kind;
    a.iq_scale = r->iq_scale;
    a.g0 = g0;
    a.n_in = static_cast<int>(n_in);
    a.hc = st.hc;
    a.phase_inc = si == 0 ? r->phase_inc : 0ull;
    a.taps = st.taps.p;
    a.n_taps = st.n_taps;
    a.interp = st.interp;
    a.decim = st.decim;
    a.m0 = st.produced;
    a.n_out = static_cast<int>(n_new);
    a.per_cta = std::max(1, std::min(kResampleThreads, static_cast<int>((static_cast<long long>(kResampleTile - 2) - st.n_taps / st.interp) * st.interp / st.decim)));
    float2* next_buf = last ? nullptr : r->stages[si + 1].buf.p;
    a.out_f = last ? nullptr : next_buf + r->stages[si + 1].hc;
    a.out_i8 = last ? r->d_out.p : nullptr;
    if (n_new > 0 && st.taps_pq.p) {  // decimating stage: polyphase kernel
      k_decimate_poly<<<static_cast<int>((n_new + kPolyOut - 1) / kPolyOut), kPolyThreads, 0, r->stream>>>(a, st.taps_pq.p);
      CU(cudaGetLastError());
    } else if (n_new > 0) {
      const int grid = static_cast<int>((n_new + a.per_cta - 1) / a.per_cta);
      k_resample<<<grid, kResampleThreads, sizeof(float2) * kResampleTile, r->stream>>>(a);
      CU(cudaGetLastError());
    }
    // carry the newest inputs of this stage over to the next push
    if (si == 0) {
      if (kind == 0) k_shift_carry<short><<<1, 1024, st.hc * 2, r->stream>>>(static_cast<short*>(static_cast<void*>(r->carry_raw.p)), static_cast<const short*>(cur_in), st.hc, n_in);
      else k_shift_carry<double><<<1, 1024, st.hc * 8, r->stream>>>(static_cast<double*>(static_cast<void*>(r->carry_raw.p)), static_cast<const double*>(cur_in), st.hc, n_in);
    } else {
      k_shift_carry<double><<<1, 1024, st.hc * 8, r->stream>>>(reinterpret_cast<double*>(st.buf.p), reinterpret_cast<const double*>(st.buf.p + st.hc), st.hc, n_in);
    }
    CU(cudaGetLastError());
    st.consumed = g1;
    st.produced = m_end;
    if (!last) {
      cur_in = a.out_f;
      cur_carry = next_buf;
      kind = 2;
      n_in = n_new;
    } else {
      produced_last = n_new;
    }
  }
  if (static_cast<size_t>(produced_last) > cap_samples) {
    CU(cudaStreamSynchronize(r->stream));
    return fail(B2S_E_INVALID, "b2s_recorder_push: %lld output samples, the buffer holds %zu", produced_last, cap_samples);
  }
  if (produced_last > 0 && out_iq) CU(cudaMemcpyAsync(out_iq, r->d_out.p, 2 * static_cast<size_t>(produced_last), cudaMemcpyDeviceToHost, r->stream));
  CU(cudaStreamSynchronize(r->stream));
  *n_out = static_cast<size_t>(produced_last);
  return 0;
}

// self-test of the exact constant division (k_check_div_const above)
int b2s_selftest_div_const(b2s_engine* e, int divisor, uint64_t* mismatches) {
  if (!e || !mismatches) return fail(B2S_E_INVALID, "NULL argument");
  CU(cudaSetDevice(e->device));
  DevBuf<unsigned long long> bad;
  int rc = bad.alloc(1);
  if (rc) return rc;
  cudaMemset(bad.p, 0, sizeof(unsigned long long));
  switch (divisor) {
    case 2: rc = run_div_check<2>(bad.p); break;
    case 3: rc = run_div_check<3>(bad.p); break;
    case 5: rc = run_div_check<5>(bad.p); break;
    case 7: rc = run_div_check<7>(bad.p); break;
    case 9: rc = run_div_check<9>(bad.p); break;
    case 11: rc = run_div_check<11>(bad.p); break;
    case 13: rc = run_div_check<13>(bad.p); break;
    case 15: rc = run_div_check<15>(bad.p); break;
    case 17: rc = run_div_check<17>(bad.p); break;
    case 19: rc = run_div_check<19>(bad.p); break;
    case 21: rc = run_div_check<21>(bad.p); break;
    default: rc = fail(B2S_E_INVALID, "no div_const instantiation for divisor %d", divisor);
  }
  unsigned long long h = 0;
  cudaError_t err = cudaSuccess;
  if (!rc) err = cudaMemcpy(&h, bad.p, sizeof(h), cudaMemcpyDeviceToHost);
  bad.release();
  if (rc) return rc;
  if (err != cudaSuccess) return fail(B2S_E_CUDA, "b2s_selftest_div_const: %s", cudaGetErrorString(err));
  *mismatches = h;
  return 0;
}

// ---- host helpers ----
int b2s_get_fft(int32_t sample_rate_hz, int32_t max_step_hz) { return host::fft_size_for(sample_rate_hz, max_step_hz); }
int32_t b2s_get_tuned_frequency(int32_t f, int32_t step) { return host::tuned_frequency(f, step); }
int b2s_get_max_index(const float* data, int size, int index, int group_size) { return host::max_index(data, size, index, group_size); }
int b2s_contains_with_margin(const int* keys, int n_keys, int index, int margin, int* found) {
  std::map<int, char> m;
  for (int i = 0; i < n_keys; ++i) m[keys[i]] = 0;
  return host::key_within_margin(m, index, margin, found) ? 1 : 0;
}
int b2s_most_frequent_value(const int* data, int n) {
  if (!data || n <= 0) return -1;
  return host::most_frequent(std::vector<int>(data, data + n));
}
int b2s_learn_frames_from_ms(int64_t learning_ms, double frame_period_ms) {
  // Noise::add (noise_learner.cpp:23): learning completes on the first frame k with t_k >= t_0 + learning_ms
  size_t k = 0;
  while (host::frame_time(0, frame_period_ms, k) < learning_ms) ++k;
  return static_cast<int>(k + 1);
}
int b2s_decimator_factor(int32_t sample_rate_hz, int32_t fft_size) { return host::decimator_factor(sample_rate_hz, fft_size); }

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------
// Transmission bookkeeping on host rows: the band's tracker behind a DeviceQueries that reads dense host arrays
// ------------------------------------------------------------------------------------------------------------
struct b2s_host_transmission : DeviceQueries {
  b2s_band_config cfg{};
  Tracker tracker;
  std::vector<float> history;  // the last Y rows of q before the current call, oldest -> newest (zeros before any data)
  // current call
  const float* box = nullptr;
  const float* q = nullptr;
  int frames = 0;
  double last_run_ms = 0.0;  // wall time of the last Tracker::run (the bookkeeping alone, without building its inputs)

  int fetch_ring_window(int frame_first, int rows, int bin_lo, int width, float* out) override {
    const int n = cfg.fft_size, Y = cfg.grouping_y;
    for (int r = 0; r < rows; ++r) {
      const int f = frame_first + r;
      float* dst = out + static_cast<size_t>(r) * width;
      const float* src = nullptr;
      if (f >= 0) {
        src = q + static_cast<size_t>(f) * n;
      } else if (Y + f >= 0) {
        src = history.data() + static_cast<size_t>(Y + f) * n;  // f = -1 is the newest row of the previous call
      }
      for (int i = 0; i < width; ++i) dst[i] = src ? src[bin_lo + i] : 0.0f;
    }
    return 0;
  }
  int query_windows(const std::vector<Window>& w, std::vector<std::vector<float>>& values, std::vector<std::vector<int>>& indices) override {
    const int n = cfg.fft_size;
    values.assign(w.size(), {});
    indices.assign(w.size(), {});
    for (size_t i = 0; i < w.size(); ++i) {
      for (int f = w[i].frame_lo; f < w[i].frame_hi; ++f) {
        const float* row = box + static_cast<size_t>(f) * n;
        int best = w[i].bin_lo;
        for (int b = w[i].bin_lo + 1; b <= w[i].bin_hi; ++b) {
          if (row[best] < row[b]) best = b;  // first maximum, collection_utils.h:9-14
        }
        values[i].push_back(row[best]);
        indices[i].push_back(best);
      }
    }
    return 0;
  }
};

extern "C" {

int b2s_host_transmission_create(const b2s_band_config* cfg, b2s_host_transmission** out) {
  if (!cfg || !out) return fail(B2S_E_INVALID, "NULL argument");
  int rc = validate_config(*cfg);
  if (rc) return rc;
  auto* h = new b2s_host_transmission();
  h->cfg = *cfg;
  h->cfg.window_taps = nullptr;
  TrackerParams& p = h->tracker.p;
  p.n = cfg->fft_size;
  p.sample_rate = cfg->sample_rate_hz;
  p.center = cfg->center_hz;
  p.range_lo = cfg->range_lo_hz;
  p.range_hi = cfg->range_hi_hz;
  p.n_ignored = cfg->n_ignored;
  for (int i = 0; i < cfg->n_ignored; ++i) {
    p.ignored_lo[i] = cfg->ignored_lo_hz[i];
    p.ignored_hi[i] = cfg->ignored_hi_hz[i];
  }
  p.group_size = cfg->group_size_bins;
  p.group_y = cfg->grouping_y;
  p.start_level = cfg->start_level;
  p.stop_level = cfg->stop_level;
  p.tuning_step = cfg->tuning_step_hz;
  p.min_time = cfg->min_time_ms;
  p.timeout = cfg->timeout_ms;
  p.max_time = cfg->max_time_ms;
  h->history.assign(static_cast<size_t>(cfg->grouping_y) * cfg->fft_size, 0.0f);  // Averager::reset fills the ring with zeros
  *out = h;
  return 0;
}
int b2s_host_transmission_destroy(b2s_host_transmission* h) {
  delete h;
  return 0;
}
double b2s_host_transmission_last_run_ms(b2s_host_transmission* h) { return h ? h->last_run_ms : 0.0; }
int b2s_host_transmission_reset(b2s_host_transmission* h) {
  if (!h) return fail(B2S_E_INVALID, "NULL handle");
  h->tracker.reset();
  std::fill(h->history.begin(), h->history.end(), 0.0f);
  return 0;
}
int b2s_host_transmission_push(b2s_host_transmission* h, const float* box_rows, const float* q_rows, int n_frames, int64_t t0_ms,
                               double frame_period_ms, int use_watch, int32_t* tx_count, b2s_transmission* tx) {
  if (!h || !box_rows || !q_rows || n_frames < 0) return fail(B2S_E_INVALID, "b2s_host_transmission_push: bad argument");
  const int n = h->cfg.fft_size, Y = h->cfg.grouping_y, T = n_frames;
  const TrackerParams& p = h->tracker.p;
  h->box = box_rows;
  h->q = q_rows;
  h->frames = T;
  // what K2 hands to the tracker: per frame the bins at or above min(start, stop), ascending
  const float level = std::min(p.start_level, p.stop_level);
  std::vector<DetectEntry> entries;
  std::vector<int> begin(T + 1, 0);
  for (int t = 0; t < T; ++t) {
    const float* row = box_rows + static_cast<size_t>(t) * n;
    begin[t] = static_cast<int>(entries.size());
    for (int b = 0; b < n; ++b) {
      if (row[b] >= level) entries.push_back(DetectEntry{b, row[b]});
    }
  }
  begin[T] = static_cast<int>(entries.size());
  // ... and, for the keys alive when the call starts, the window maxima and the "uncovered candidate" flags (k_detect's box warps)
  Tracker::Watch watch;
  std::vector<int> keys, flags;
  std::vector<unsigned int> maxima;
  if (use_watch) {
    for (const auto& kv : h->tracker.signals) {
      if (static_cast<int>(keys.size()) < kMaxWatch) keys.push_back(kv.first);
    }
    const int gh = p.group_size / 2, margin = (p.group_size % 2 == 0) ? gh : gh + 1;
    maxima.assign(static_cast<size_t>(T) * kMaxWatch, 0u);
    flags.assign(T, 0);
    for (int t = 0; t < T; ++t) {
      const float* row = box_rows + static_cast<size_t>(t) * n;
      for (size_t i = 0; i < keys.size(); ++i) {
        float m = -INFINITY;
        for (int b = std::max(0, keys[i] - gh); b <= std::min(n - 1, keys[i] + gh); ++b) m = std::max(m, row[b]);
        maxima[static_cast<size_t>(t) * kMaxWatch + i] = float_to_ordered(m);
      }
      for (int b = 0; b < n && !flags[t]; ++b) {
        if (row[b] < p.start_level) continue;
        bool covered = false;
        for (int key : keys) covered = covered || (b >= key - margin && b <= key + margin);
        if (!covered) flags[t] = 1;
      }
    }
    watch = Tracker::Watch{static_cast<int>(keys.size()), keys.data(), maxima.data(), flags.data()};
  }
  std::vector<Tracker::FrameState> states;
  const DetectEntry* ep = entries.empty() ? nullptr : entries.data();
  const auto run_t0 = std::chrono::steady_clock::now();
  int rc = h->tracker.run(ep, begin.data(), static_cast<size_t>(T), t0_ms, frame_period_ms, 0, *h, tx_count != nullptr || tx != nullptr, watch, states);
  h->last_run_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - run_t0).count();
  if (rc) return rc;
  if (tx_count) std::fill(tx_count, tx_count + T, 0);
  for (const auto& fs : states) {
    const int total = h->tracker.sorted_transmissions(fs, tx ? tx + static_cast<size_t>(fs.frame) * B2S_MAX_TX : nullptr, tx ? B2S_MAX_TX : 0);
    if (tx_count) tx_count[fs.frame] = total;
  }
  // keep the newest Y rows of q for the next call's getBestIndex look-back
  std::vector<float> next(static_cast<size_t>(Y) * n);
  for (int i = 0; i < Y; ++i) {
    const int f = T - Y + i;
    const float* src = f >= 0 ? q_rows + static_cast<size_t>(f) * n : h->history.data() + static_cast<size_t>(T + i) * n;
    std::memcpy(next.data() + static_cast<size_t>(i) * n, src, sizeof(float) * n);
  }
  h->history.swap(next);
  h->box = h->q = nullptr;
  return 0;
}

// append the object representation of a field (the reference writes through reinterpret_cast on a little-endian host)
static void put_bytes(uint8_t* out, size_t& at, const void* v, size_t n) {
  std::memcpy(out + at, v, n);
  at += n;
}
#define B2S_PUT(type, expr)          \
  do {                               \
    const type field_ = (expr);      \
    put_bytes(out, at, &field_, sizeof(type)); \
  } while (0)

int b2s_pack_spectrogram_message(int64_t time_ms, int32_t center_hz, int32_t sample_rate_hz, const int8_t* row, int size, uint8_t* out, size_t cap,
                                 size_t* written) {
  if (!row || !out || !written || size <= 0) return fail(B2S_E_INVALID, "b2s_pack_spectrogram_message: NULL argument or size <= 0");
  const size_t need = sizeof(uint64_t) + 3 * sizeof(int32_t) + sizeof(uint32_t) + static_cast<size_t>(size);
  *written = need;
  if (cap < need) return fail(B2S_E_INVALID, "spectrogram message needs %zu bytes, buffer holds %zu", need, cap);
  size_t at = 0;
  B2S_PUT(uint64_t, time_ms);
  B2S_PUT(int32_t, center_hz - sample_rate_hz / 2);  // start
  B2S_PUT(int32_t, center_hz + sample_rate_hz / 2);  // stop
  B2S_PUT(int32_t, sample_rate_hz / size);           // step
  B2S_PUT(uint32_t, size);
  std::memcpy(out + at, row, static_cast<size_t>(size));
  return 0;
}

int b2s_pack_transmission_message(int64_t time_ms, int32_t frequency_hz, int32_t sample_rate_hz, const int8_t* iq, int n_samples, uint8_t* out,
                                  size_t cap, size_t* written) {
  if (!out || !written || n_samples < 0 || (n_samples > 0 && !iq)) return fail(B2S_E_INVALID, "b2s_pack_transmission_message: bad argument");
  const size_t header = sizeof(uint64_t) + 2 * sizeof(int32_t) + sizeof(uint32_t), body = 2 * static_cast<size_t>(n_samples);
  *written = header + body;
  if (cap < header + body) return fail(B2S_E_INVALID, "transmission message needs %zu bytes, buffer holds %zu", header + body, cap);
  size_t at = 0;
  B2S_PUT(uint64_t, time_ms);
  B2S_PUT(int32_t, frequency_hz - sample_rate_hz / 2);
  B2S_PUT(int32_t, frequency_hz + sample_rate_hz / 2);
  B2S_PUT(uint32_t, sample_rate_hz);
  for (size_t i = 0; i < body; ++i) out[at + i] = static_cast<uint8_t>(iq[i]) ^ 0x80u;  // offset-binary bytes, data_controller.cpp:38-40
  return 0;
}
#undef B2S_PUT

}  // extern "C"
