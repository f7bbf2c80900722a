// hode_gen4gi.cu — on-device cohort generation: the 8-state "4GI" glucose / insulin / GLP-1 / glucagon / GIP
// simulator that produces the reference's training data (reference data/generate4GI.py: parameters :15-62,
// baselines :64-70, equations :72-160, per-interval integration and meal distribution :162-211).
//
// the step BEFORE the hot path — the reference integrates one subject at a time with
// scipy.integrate.odeint (LSODA) per 5-minute interval; here one thread integrates one subject with an
// adaptive Dormand-Prince 5(4) in float64 (the data are generated once; float64 keeps the result within
// LSODA's own tolerance of the reference's), the meal rate being piecewise constant per sampling interval
// exactly as the reference distributes it.  Output: the 5 concentration series at the sampling times.
#include <math.h>

#include "hode_common.cuh"
#include "hode_kernels.h"

namespace hode {

namespace {

// model constants (reference data/generate4GI.py:15-62); [0] = T2DM, [1] = HV where they differ
struct G4 {
  double CLglc, CLglci, pow_below;
  static constexpr double Qglc = 26.5, VCglc = 9.33, VPglc = 8.56;
  static constexpr double CLins = 73.2, VCins = 6.09;
  static constexpr double VCglp = 16.0;
  static constexpr double CLglg = 453.2, VCglg = 64.6;
  static constexpr double CLgip = 86.8, VCgip = 9.21, Qgip = 49.4, VPgip = 22.8;
  static constexpr double GLCINS_S = 2.46, HILL_1 = 1.79, EMAX_4 = 6.73;
  static constexpr double FDGLP = 0.0102, FDGIP = 0.0343, FDGLG = 0.00329;
  static constexpr double POW_ABOVE = 0.925;
};

struct Subject {
  double bglc, bins, bglp, bglg, bgip;                   // baselines
  double KINglc, KINins, KINglp, KINglg, KINgip, S0glg;  // baseline production rates, 1 + GLGGLC_S0
};

struct Derived {   // exp() constants evaluated once on the host
  double Ke0ins, VM_GLP, KM_GLP, EMAX_1, EC50_1, EC50_4;
};

__device__ __forceinline__ void rhs4gi(const G4& g, const Derived& d, const Subject& s, const double* y, double meal,
                                       double* f) {
  const double Gc = y[0], Ins = y[1], GLP = y[2], Glg = y[3], GIP = y[4], Gp = y[5], InsE = y[6], GIPp = y[7];
  const double Cglc = Gc / G4::VCglc, Cins = Ins / G4::VCins, Cglp = GLP / G4::VCglp, Cglg = Glg / G4::VCglg;
  const double hill = pow(Cglp / d.EC50_1, G4::HILL_1);
  const double GLPINS_S = d.EMAX_1 * hill / (1.0 + hill);
  const double GLGGLC_S = G4::EMAX_4 * (Cglg / d.EC50_4) / (1.0 + Cglg / d.EC50_4);
  const double glgEFFglc = (1.0 + GLGGLC_S) / s.S0glg;
  const double pw = Cglc >= s.bglc ? G4::POW_ABOVE : g.pow_below;
  const double glcEFFglg = Cglc > 0.0 ? pow(s.bglc / Cglc, pw) : 1.0;
  const double me = meal * 10.0;   // food effects act on 10x the meal rate (:117)
  const double fglp = me > 0.0 ? G4::FDGLP * me : 0.0, fgip = me > 0.0 ? G4::FDGIP * me : 0.0,
               fglg = me > 0.0 ? G4::FDGLG * me : 0.0;
  const double K27 = G4::Qglc / G4::VCglc, K72 = G4::Qglc / G4::VPglc;
  const double K612 = G4::Qgip / G4::VCgip, K126 = G4::Qgip / G4::VPgip;
  f[0] = meal + s.KINglc * glgEFFglc - K27 * Gc + K72 * Gp - (g.CLglc / G4::VCglc) * Gc -
         (g.CLglci * InsE / G4::VCglc) * Gc;
  f[1] = s.KINins * (1.0 + GLPINS_S * pow(Cglc, G4::GLCINS_S)) - (G4::CLins / G4::VCins) * Ins;
  f[2] = s.KINglp * (1.0 + fglp) - d.VM_GLP * Cglp / (d.KM_GLP + Cglp);
  f[3] = s.KINglg * (1.0 + fglg) * glcEFFglg - (G4::CLglg / G4::VCglg) * Glg;
  f[4] = s.KINgip * (1.0 + fgip) - (G4::CLgip / G4::VCgip) * GIP - K612 * GIP + K126 * GIPp;
  f[5] = K27 * Gc - K72 * Gp;
  f[6] = d.Ke0ins * (Cins - InsE);
  f[7] = K612 * GIP - K126 * GIPp;
}

constexpr int NY = 8;

// double-precision DP5(4) tableau (the float one lives in hode_common.cuh)
constexpr double A21 = 1.0 / 5, A31 = 3.0 / 40, A32 = 9.0 / 40, A41 = 44.0 / 45, A42 = -56.0 / 15, A43 = 32.0 / 9,
                 A51 = 19372.0 / 6561, A52 = -25360.0 / 2187, A53 = 64448.0 / 6561, A54 = -212.0 / 729,
                 A61 = 9017.0 / 3168, A62 = -355.0 / 33, A63 = 46732.0 / 5247, A64 = 49.0 / 176, A65 = -5103.0 / 18656,
                 B1 = 35.0 / 384, B3 = 500.0 / 1113, B4 = 125.0 / 192, B5 = -2187.0 / 6784, B6 = 11.0 / 84,
                 E1 = -71.0 / 57600, E3 = 71.0 / 16695, E4 = -71.0 / 1920, E5 = 17253.0 / 339200, E6 = -22.0 / 525,
                 E7 = 1.0 / 40;

__global__ void __launch_bounds__(128) gen4gi_kernel(G4 g, Derived d, int n, int T, double dt, double rtol, double atol,
                                                     const float* __restrict__ baselines,
                                                     const float* __restrict__ meal_rate, float* __restrict__ out,
                                                     int32_t* __restrict__ status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Subject s;
  s.bglc = baselines[i * 5 + 0]; s.bins = baselines[i * 5 + 1]; s.bglp = baselines[i * 5 + 2];
  s.bglg = baselines[i * 5 + 3]; s.bgip = baselines[i * 5 + 4];
  {   // baseline production rates (:106-110) and the baseline effects they are built from (:95-101)
    const double h0 = pow(s.bglp / d.EC50_1, G4::HILL_1);
    const double GLPINS_S0 = d.EMAX_1 * h0 / (1.0 + h0);
    s.S0glg = 1.0 + G4::EMAX_4 * (s.bglg / d.EC50_4) / (1.0 + s.bglg / d.EC50_4);
    s.KINglc = s.bglc * (g.CLglc + g.CLglci * s.bins);
    s.KINins = s.bins * G4::CLins / (1.0 + GLPINS_S0 * pow(s.bglc, G4::GLCINS_S));
    s.KINglp = d.VM_GLP * s.bglp * G4::VCglp / (d.KM_GLP + s.bglp);
    s.KINglg = s.bglg * G4::CLglg;
    s.KINgip = s.bgip * G4::CLgip;
  }
  double y[NY] = {s.bglc * G4::VCglc, s.bins * G4::VCins, s.bglp * G4::VCglp, s.bglg * G4::VCglg,
                  s.bgip * G4::VCgip, s.bglc * G4::VPglc, s.bins, s.bgip * G4::VPgip};   // :176-185
  float* o = out + (size_t)i * T * 5;
  auto emit = [&](int k) {
    o[k * 5 + 0] = (float)(y[0] / G4::VCglc); o[k * 5 + 1] = (float)(y[1] / G4::VCins);
    o[k * 5 + 2] = (float)(y[2] / G4::VCglp); o[k * 5 + 3] = (float)(y[3] / G4::VCglg);
    o[k * 5 + 4] = (float)(y[4] / G4::VCgip);
  };
  emit(0);
  int st = 0;
  double h = dt * 0.25;
  for (int k = 0; k + 1 < T; ++k) {
    const double meal = meal_rate ? (double)meal_rate[(size_t)i * (T - 1) + k] : 0.0;
    double t = 0.0;   // local time inside the interval (the system is autonomous between meal changes)
    double k1[NY], k2[NY], k3[NY], k4[NY], k5[NY], k6[NY], k7[NY], ys[NY], yn[NY];
    rhs4gi(g, d, s, y, meal, k1);
    int guard = 0;
    while (t < dt && st == 0) {
      if (++guard > 100000) { st = HODE_ST_MAX_STEPS; break; }
      const double hh = fmin(h, dt - t);
#pragma unroll
      for (int c = 0; c < NY; ++c) ys[c] = y[c] + hh * A21 * k1[c];
      rhs4gi(g, d, s, ys, meal, k2);
#pragma unroll
      for (int c = 0; c < NY; ++c) ys[c] = y[c] + hh * (A31 * k1[c] + A32 * k2[c]);
      rhs4gi(g, d, s, ys, meal, k3);
#pragma unroll
      for (int c = 0; c < NY; ++c) ys[c] = y[c] + hh * (A41 * k1[c] + A42 * k2[c] + A43 * k3[c]);
      rhs4gi(g, d, s, ys, meal, k4);
#pragma unroll
      for (int c = 0; c < NY; ++c) ys[c] = y[c] + hh * (A51 * k1[c] + A52 * k2[c] + A53 * k3[c] + A54 * k4[c]);
      rhs4gi(g, d, s, ys, meal, k5);
#pragma unroll
      for (int c = 0; c < NY; ++c)
        ys[c] = y[c] + hh * (A61 * k1[c] + A62 * k2[c] + A63 * k3[c] + A64 * k4[c] + A65 * k5[c]);
      rhs4gi(g, d, s, ys, meal, k6);
#pragma unroll
      for (int c = 0; c < NY; ++c) yn[c] = y[c] + hh * (B1 * k1[c] + B3 * k3[c] + B4 * k4[c] + B5 * k5[c] + B6 * k6[c]);
      rhs4gi(g, d, s, yn, meal, k7);
      double e2 = 0.0;
      bool finite = true;
#pragma unroll
      for (int c = 0; c < NY; ++c) {
        const double sc = atol + rtol * fmax(fabs(y[c]), fabs(yn[c]));
        const double e = hh * (E1 * k1[c] + E3 * k3[c] + E4 * k4[c] + E5 * k5[c] + E6 * k6[c] + E7 * k7[c]) / sc;
        e2 += e * e;
        finite = finite && isfinite(yn[c]);
      }
      const double err = finite ? sqrt(e2 / NY) : 1e30;
      if (err < 1.0) {
        t += hh;
#pragma unroll
        for (int c = 0; c < NY; ++c) { y[c] = yn[c]; k1[c] = k7[c]; }
        if (hh == h) h = hh * fmin(5.0, fmax(0.2, 0.9 * pow(fmax(err, 1e-12), -0.2)));
      } else {
        h = hh * fmax(0.2, 0.9 * pow(err, -0.2));
        if (!(h > 1e-14 * dt)) st = HODE_ST_STEP_TOO_SMALL;
      }
    }
    if (st != 0) {   // like the rollout: rows after a failure are zero
      for (int kk = k + 1; kk < T; ++kk)
        for (int c = 0; c < 5; ++c) o[kk * 5 + c] = 0.f;
      break;
    }
    emit(k + 1);
  }
  if (status) status[i] = st;
}

}  // namespace

cudaError_t launch_gen4gi(int n, int T, double dt_hours, int patient_type, double rtol, double atol,
                          const float* baselines, const float* meal_rate, float* out, int32_t* status,
                          cudaStream_t stream) {
  G4 g;
  if (patient_type == 0) { g.CLglc = 1.72; g.CLglci = 0.0256; g.pow_below = 0.0; }      // T2DM (:19-21, :103-104)
  else { g.CLglc = 5.36; g.CLglci = 0.072; g.pow_below = 0.327; }                       // HV   (:22-24, :105-106)
  Derived d;
  d.Ke0ins = exp(-0.159); d.VM_GLP = exp(7.97); d.KM_GLP = exp(4.91);
  d.EMAX_1 = exp(2.37); d.EC50_1 = exp(3.29); d.EC50_4 = exp(4.59);
  if (n <= 0) return cudaSuccess;
  count_launch();
  gen4gi_kernel<<<(n + 127) / 128, 128, 0, stream>>>(g, d, n, T, dt_hours, rtol, atol, baselines, meal_rate, out, status);
  return cudaGetLastError();
}

}  // namespace hode
