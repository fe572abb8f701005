// K1 (packed-f32x2 variant, N >= 4096): same transform, same pass plan and same epilogue as k_spectrum (spectral.cuh), but
// every thread works on PAIRS of adjacent butterflies (b, b+1) and all arithmetic is issued as Blackwell packed-fp32
// instructions (FADD2 / FMUL2 / FFMA2, PTX add/mul/fma.rn.f32x2): lane .x of every register pair belongs to butterfly b,
// lane .y to butterfly b+1. That halves the floating-point instruction count of the FFT, which is what bounds K1
// (profiles/r01_k1_v1.1_ncu_summary.txt: 70 % of the executed instructions are FADD/FMUL/FFMA at 49 % issue utilisation).
//
// To make adjacent butterflies land in adjacent words the exchange buffer is PLANAR (all real parts, then all imaginary
// parts) and padded by 2 floats per 32 (keeps 8-byte alignment of the pairs; makes the 32-float-strided stores of the first
// pass 2-way instead of 32-way conflicted and keeps every other access pattern at its minimum wavefront count).
// Twiddle tables are planar too ([m-1][k], real plane then imaginary plane), so the pair (k, k+1) is one 8-byte load.
#pragma once
#include "spectral.cuh"

namespace b2s {

__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, neg2(b)); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 bc(float c) { return make_float2(c, c); }

// (r, i) *= (wr + i*wi), per lane
__device__ __forceinline__ void cmul2(float2& r, float2& i, float2 wr, float2 wi) {
  const float2 nr = fma2(r, wr, neg2(mul2(i, wi)));
  const float2 ni = fma2(r, wi, mul2(i, wr));
  r = nr;
  i = ni;
}
// multiply by the compile-time 16th root of unity W16^e = exp(-2 pi i e / 16)
template <int E16>
__device__ __forceinline__ void cmulw16(float2& r, float2& i) {
  constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, H = 0.70710678118654752f;
  constexpr int e = E16 & 15;
  if (e == 0) return;
  if (e == 4) {  // -i
    const float2 t = r;
    r = i;
    i = neg2(t);
    return;
  }
  if (e == 8) {
    r = neg2(r);
    i = neg2(i);
    return;
  }
  if (e == 12) {  // +i
    const float2 t = r;
    r = neg2(i);
    i = t;
    return;
  }
  constexpr float wr = (e == 1) ? C1 : (e == 2) ? H : (e == 3) ? S1 : (e == 5) ? -S1 : (e == 6) ? -H : (e == 7) ? -C1 : (e == 9) ? -C1 : (e == 10) ? -H
                       : (e == 11) ? -S1 : (e == 13) ? S1 : (e == 14) ? H : C1;
  constexpr float wi = (e == 1) ? -S1 : (e == 2) ? -H : (e == 3) ? -C1 : (e == 5) ? -C1 : (e == 6) ? -H : (e == 7) ? -S1 : (e == 9) ? S1 : (e == 10) ? H
                       : (e == 11) ? C1 : (e == 13) ? C1 : (e == 14) ? H : S1;
  cmul2(r, i, bc(wr), bc(wi));
}

template <int R>
struct Dft2x;
template <>
struct Dft2x<1> {
  __device__ __forceinline__ static void run(float2*, float2*) {}
};
template <>
struct Dft2x<2> {
  __device__ __forceinline__ static void run(float2* r, float2* i) {
    const float2 ar = r[0], ai = i[0];
    r[0] = add2(ar, r[1]);
    i[0] = add2(ai, i[1]);
    r[1] = sub2(ar, r[1]);
    i[1] = sub2(ai, i[1]);
  }
};
template <>
struct Dft2x<4> {
  __device__ __forceinline__ static void run(float2* r, float2* i) {
    const float2 t0r = add2(r[0], r[2]), t0i = add2(i[0], i[2]);
    const float2 t1r = sub2(r[0], r[2]), t1i = sub2(i[0], i[2]);
    const float2 t2r = add2(r[1], r[3]), t2i = add2(i[1], i[3]);
    const float2 t3r = sub2(r[1], r[3]), t3i = sub2(i[1], i[3]);
    r[0] = add2(t0r, t2r);
    i[0] = add2(t0i, t2i);
    r[2] = sub2(t0r, t2r);
    i[2] = sub2(t0i, t2i);
    r[1] = add2(t1r, t3i);  // t1 + (-i) t3
    i[1] = sub2(t1i, t3r);
    r[3] = sub2(t1r, t3i);  // t1 - (-i) t3
    i[3] = add2(t1i, t3r);
  }
};
// Cooley-Tukey R = 4 * (R/4): n = N2*n1 + n2, k = k1 + 4*k2
template <int R>
struct Dft2x {
  template <int N2, int n2>
  __device__ __forceinline__ static void column(const float2* r, const float2* i, float2 (&yr)[R / 4][4], float2 (&yi)[R / 4][4]) {
    float2 ar[4], ai[4];
#pragma unroll
    for (int n1 = 0; n1 < 4; ++n1) {
      ar[n1] = r[N2 * n1 + n2];
      ai[n1] = i[N2 * n1 + n2];
    }
    Dft2x<4>::run(ar, ai);
    cmulw16<(16 / R) * n2 * 1>(ar[1], ai[1]);
    cmulw16<(16 / R) * n2 * 2>(ar[2], ai[2]);
    cmulw16<(16 / R) * n2 * 3>(ar[3], ai[3]);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
      yr[n2][k1] = ar[k1];
      yi[n2][k1] = ai[k1];
    }
    if constexpr (n2 + 1 < N2) column<N2, n2 + 1>(r, i, yr, yi);
  }
  __device__ __forceinline__ static void run(float2* r, float2* i) {
    constexpr int N2 = R / 4;
    float2 yr[N2][4], yi[N2][4];
    column<N2, 0>(r, i, yr, yi);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
      float2 br[N2], bi[N2];
#pragma unroll
      for (int n2 = 0; n2 < N2; ++n2) {
        br[n2] = yr[n2][k1];
        bi[n2] = yi[n2][k1];
      }
      Dft2x<N2>::run(br, bi);
#pragma unroll
      for (int k2 = 0; k2 < N2; ++k2) {
        r[k1 + 4 * k2] = br[k2];
        i[k1 + 4 * k2] = bi[k2];
      }
    }
  }
};

// planar exchange buffer padding: 2 floats per 32
__host__ __device__ constexpr int padp(int i) { return i + 2 * (i >> 5); }
template <int N>
__host__ __device__ constexpr int plane_elems() { return N + N / 16; }

// planar twiddle layout: for each pass with P > 1: real plane [(R-1)*P] then imaginary plane [(R-1)*P]
template <int N>
struct TwiddleLayout2 {
  using PL = FftPlanT<N>;
  static constexpr int P1 = PL::R0, P2 = PL::R0 * PL::R1, P3 = PL::R0 * PL::R1 * PL::R2;
  static constexpr int E1 = 2 * twiddle_entries<PL::R1, P1>();
  static constexpr int E2 = PL::R2 > 1 ? 2 * twiddle_entries<PL::R2, P2>() : 0;
  static constexpr int E3 = PL::R3 > 1 ? 2 * twiddle_entries<PL::R3, P3>() : 0;
  static constexpr int O1 = 0, O2 = E1, O3 = E1 + E2, TOTAL = E1 + E2 + E3;  // float offsets into the global table
  static constexpr bool S1 = twiddle_in_smem<PL::R1, P1>();
  static constexpr bool S2 = PL::R2 > 1 && twiddle_in_smem<PL::R2, P2>();
  static constexpr bool S3 = PL::R3 > 1 && twiddle_in_smem<PL::R3, P3>();
  static constexpr int SO1 = 0, SO2 = S1 ? E1 : 0, SO3 = SO2 + (S2 ? E2 : 0);
  static constexpr int SMEM = SO3 + (S3 ? E3 : 0);  // floats
};

template <int N, int R, int P, int E, int T>
__device__ __forceinline__ void pass2_twiddle_butterfly(float2 (&re)[E / 2], float2 (&im)[E / 2], const float* __restrict__ tw, int tid) {
  constexpr int PP = E / (2 * R);
#pragma unroll
  for (int u = 0; u < PP; ++u) {
    if (P > 1) {
      const int k = (2 * (tid + u * T)) & (P - 1);
      const float* twr = tw + k;
      const float* twi = tw + (R - 1) * P + k;
#pragma unroll
      for (int m = 1; m < R; ++m) {
        const float2 wr = *reinterpret_cast<const float2*>(twr + (m - 1) * P);
        const float2 wi = *reinterpret_cast<const float2*>(twi + (m - 1) * P);
        cmul2(re[u * R + m], im[u * R + m], wr, wi);
      }
    }
    Dft2x<R>::run(&re[u * R], &im[u * R]);
  }
}

template <int N, int R, int E, int T>
__device__ __forceinline__ void pass2_load(const float* XR, const float* XI, float2 (&re)[E / 2], float2 (&im)[E / 2], int tid) {
  constexpr int NB = N / R, PP = E / (2 * R);
  static_assert(NB % 32 == 0 && (2 * T) % 32 == 0, "padding arithmetic assumes multiples of 32");
  const int base = padp(2 * tid);
#pragma unroll
  for (int u = 0; u < PP; ++u) {
#pragma unroll
    for (int m = 0; m < R; ++m) {
      constexpr int dummy = 0;
      (void)dummy;
      const int off = padp(u * 2 * T + m * NB);  // compile-time: both terms are multiples of 32
      re[u * R + m] = *reinterpret_cast<const float2*>(XR + base + off);
      im[u * R + m] = *reinterpret_cast<const float2*>(XI + base + off);
    }
  }
}

template <int N, int R, int P, int E, int T>
__device__ __forceinline__ void pass2_store(float* XR, float* XI, const float2 (&re)[E / 2], const float2 (&im)[E / 2], int tid) {
  constexpr int PP = E / (2 * R);
#pragma unroll
  for (int u = 0; u < PP; ++u) {
    const int b = 2 * (tid + u * T);
    if (P == 1) {
      // butterfly b writes R*b + m, butterfly b+1 writes R*(b+1) + m: two 4-byte stores per plane
#pragma unroll
      for (int m = 0; m < R; ++m) {
        const int i0 = padp(R * b + m), i1 = padp(R * (b + 1) + m);
        XR[i0] = re[u * R + m].x;
        XR[i1] = re[u * R + m].y;
        XI[i0] = im[u * R + m].x;
        XI[i1] = im[u * R + m].y;
      }
    } else {
      const int k = b & (P - 1);
      const int j = ((b - k) * R) + k;  // even; butterfly b+1 writes j + 1
      if (P % 32 == 0) {
        const int pj = padp(j);
#pragma unroll
        for (int m = 0; m < R; ++m) {
          *reinterpret_cast<float2*>(XR + pj + padp(m * P)) = re[u * R + m];
          *reinterpret_cast<float2*>(XI + pj + padp(m * P)) = im[u * R + m];
        }
      } else {
#pragma unroll
        for (int m = 0; m < R; ++m) {
          const int idx = padp(j + m * P);
          *reinterpret_cast<float2*>(XR + idx) = re[u * R + m];
          *reinterpret_cast<float2*>(XI + idx) = im[u * R + m];
        }
      }
    }
  }
}

template <int N, int MODE, bool DEBUG_LIN>
__global__ void __launch_bounds__(N / FftPlanT<N>::E) k_spectrum2(const SpectralArgs a) {
  using PL = FftPlanT<N>;
  using TL = TwiddleLayout2<N>;
  constexpr int E = PL::E, T = N / E;
  static_assert(E == 32, "the packed variant holds 16 butterfly-pair elements per thread");
  constexpr int R0 = PL::R0, R1 = PL::R1, R2 = PL::R2, R3 = PL::R3;
  constexpr int NP = (R3 > 1) ? 4 : (R2 > 1 ? 3 : 2);
  constexpr int P1 = R0, P2 = R0 * R1, P3 = R0 * R1 * R2;
  constexpr int RL = (NP == 4) ? R3 : (NP == 3 ? R2 : R1);
  static_assert(R0 * R1 * R2 * R3 == N, "plan");

  extern __shared__ __align__(128) unsigned char smem[];
  float* XR = reinterpret_cast<float*>(smem);
  float* XI = XR + plane_elems<N>();
  float* tws = XI + plane_elems<N>();                                     // shared copy of the small twiddle tables
  unsigned char* raw = reinterpret_cast<unsigned char*>(tws + TL::SMEM);  // 2N bytes (TMA mode only), 16-byte aligned
  __shared__ __align__(8) uint64_t full_bar;
  __shared__ float red_v[32];
  __shared__ int red_i[32];

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const char* base = static_cast<const char*>(a.iq);
  const float* twg = reinterpret_cast<const float*>(a.twiddle);

  if (MODE == kModeCs8Tma) {
    if (tid == 0) {
      mbar_init(&full_bar, 1);
      fence_barrier_init();
    }
  }
  if (TL::S1) {
    for (int i = tid; i < TL::E1; i += T) tws[TL::SO1 + i] = twg[TL::O1 + i];
  }
  if (TL::S2) {
    for (int i = tid; i < TL::E2; i += T) tws[TL::SO2 + i] = twg[TL::O2 + i];
  }
  if (TL::S3) {
    for (int i = tid; i < TL::E3; i += T) tws[TL::SO3 + i] = twg[TL::O3 + i];
  }
  __syncthreads();
  const float* tw1 = TL::S1 ? tws + TL::SO1 : twg + TL::O1;
  const float* tw2 = TL::S2 ? tws + TL::SO2 : twg + TL::O2;
  const float* tw3 = TL::S3 ? tws + TL::SO3 : twg + TL::O3;
  if (MODE == kModeCs8Tma) {
    if (tid == 0 && static_cast<int>(blockIdx.x) < a.n_frames) {
      mbar_arrive_expect_tx(&full_bar, 2 * N);
      bulk_g2s(raw, base + static_cast<long long>(blockIdx.x) * a.frame_stride_bytes, 2 * N, &full_bar);
    }
  }

  uint32_t parity = 0;
  for (int frame = blockIdx.x; frame < a.n_frames; frame += gridDim.x) {
    float2 re[E / 2], im[E / 2];
    // ---------------- pass 0: unpack + window, radix R0, no twiddles (P = 1); pair p = samples (2p, 2p+1) + m*NB ----------------
    {
      constexpr int NB = N / R0, PP = E / (2 * R0);
      if (MODE == kModeCs8Tma) mbar_wait(&full_bar, parity);
      parity ^= 1;
      const float2* w2 = reinterpret_cast<const float2*>(a.wscale);
#pragma unroll
      for (int u = 0; u < PP; ++u) {
        const int p = tid + u * T;
#pragma unroll
        for (int m = 0; m < R0; ++m) {
          const int q = p + m * (NB / 2);  // pair index: samples 2q, 2q+1
          const float2 w = __ldg(&w2[q]);
          float2 xr, xi;
          if (MODE == kModeCs8Tma) {
            const char4 s = reinterpret_cast<const char4*>(raw)[q];
            xr = make_float2(static_cast<float>(s.x), static_cast<float>(s.z));
            xi = make_float2(static_cast<float>(s.y), static_cast<float>(s.w));
          } else if (MODE == kModeCs8Direct) {
            const signed char* fp = reinterpret_cast<const signed char*>(base + static_cast<long long>(frame) * a.frame_stride_bytes) + 4 * q;
            xr = make_float2(static_cast<float>(fp[0]), static_cast<float>(fp[2]));
            xi = make_float2(static_cast<float>(fp[1]), static_cast<float>(fp[3]));
          } else {
            const float* fp = reinterpret_cast<const float*>(base + static_cast<long long>(frame) * a.frame_stride_bytes) + 4 * q;
            xr = make_float2(fp[0], fp[2]);
            xi = make_float2(fp[1], fp[3]);
          }
          re[u * R0 + m] = mul2(xr, w);
          im[u * R0 + m] = mul2(xi, w);
        }
      }
      pass2_twiddle_butterfly<N, R0, 1, E, T>(re, im, nullptr, tid);
      pass2_store<N, R0, 1, E, T>(XR, XI, re, im, tid);
    }
    __syncthreads();
    if (MODE == kModeCs8Tma && tid == 0) {  // staging buffer consumed: fetch this CTA's next frame behind the remaining passes
      const int next = frame + gridDim.x;
      if (next < a.n_frames) {
        mbar_arrive_expect_tx(&full_bar, 2 * N);
        bulk_g2s(raw, base + static_cast<long long>(next) * a.frame_stride_bytes, 2 * N, &full_bar);
      }
    }
    if (NP >= 3) {
      pass2_load<N, R1, E, T>(XR, XI, re, im, tid);
      __syncthreads();
      pass2_twiddle_butterfly<N, R1, P1, E, T>(re, im, tw1, tid);
      pass2_store<N, R1, P1, E, T>(XR, XI, re, im, tid);
      __syncthreads();
    }
    if (NP >= 4) {
      pass2_load<N, R2, E, T>(XR, XI, re, im, tid);
      __syncthreads();
      pass2_twiddle_butterfly<N, R2, P2, E, T>(re, im, tw2, tid);
      pass2_store<N, R2, P2, E, T>(XR, XI, re, im, tid);
      __syncthreads();
    }
    // ---------------- last pass + epilogue ----------------
    pass2_load<N, RL, E, T>(XR, XI, re, im, tid);
    constexpr int PL_ = (NP == 4) ? P3 : (NP == 3 ? P2 : P1);
    pass2_twiddle_butterfly<N, RL, PL_, E, T>(re, im, NP == 4 ? tw3 : (NP == 3 ? tw2 : tw1), tid);

    // pair (b, b+1) holds bins k = b + m*NB and k+1: |X|^2/fs -> dB (psd.cpp:18), stored as one 8-byte word at (k + N/2) mod N
    float* row = a.psd_db + static_cast<size_t>(frame) * N;
    float best_v = -INFINITY;
    {
      constexpr int NB = N / RL, PP = E / (2 * RL);
      constexpr float kDbPerLog2 = 3.0102999566398120f;
      const float2 inv = bc(a.inv_fs);
#pragma unroll
      for (int u = 0; u < PP; ++u) {
        const int b = 2 * (tid + u * T);
#pragma unroll
        for (int m = 0; m < RL; ++m) {
          const int j = (b + m * NB + N / 2) & (N - 1);
          const float2 pw = mul2(fma2(re[u * RL + m], re[u * RL + m], mul2(im[u * RL + m], im[u * RL + m])), inv);
          const float2 db = make_float2(kDbPerLog2 * fast_log2(pw.x), kDbPerLog2 * fast_log2(pw.y));
          *reinterpret_cast<float2*>(row + j) = db;
          if (DEBUG_LIN) *reinterpret_cast<float2*>(a.power_lin + static_cast<size_t>(frame) * N + j) = pw;
          re[u * RL + m] = db;  // kept for the argmax resolution below
          best_v = fmaxf(best_v, fmaxf(db.x, db.y));
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best_v = fmaxf(best_v, __shfl_xor_sync(0xffffffffu, best_v, o));
    if (lane == 0) red_v[warp] = best_v;
    __syncthreads();  // also: all reads of the exchange buffer are done before the next frame's first pass overwrites it
    {
      constexpr int NW = (T + 31) / 32;
      float row_max = red_v[0];
#pragma unroll
      for (int w = 1; w < NW; ++w) row_max = fmaxf(row_max, red_v[w]);
      int best_i = 0x7fffffff;
      constexpr int NB = N / RL, PP = E / (2 * RL);
#pragma unroll
      for (int u = 0; u < PP; ++u) {
#pragma unroll
        for (int m = 0; m < RL; ++m) {
          const int j = (2 * (tid + u * T) + m * NB + N / 2) & (N - 1);
          if (re[u * RL + m].x == row_max) best_i = min(best_i, j);
          if (re[u * RL + m].y == row_max) best_i = min(best_i, j + 1);
        }
      }
      if (tid == 0) red_i[0] = 0x7fffffff;
      __syncthreads();
      if (best_i != 0x7fffffff) atomicMin(&red_i[0], best_i);
      __syncthreads();
      if (tid == 0) {
        a.peak_index[frame] = red_i[0];
        a.peak_value[frame] = row_max;
      }
    }
  }
}

}  // namespace b2s
