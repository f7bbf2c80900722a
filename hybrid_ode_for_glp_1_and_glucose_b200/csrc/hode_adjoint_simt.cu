// hode_adjoint_simt.cu — gradients of the rollout path on FP32 CUDA cores.
//
//   rollout_bwd_kernel : discrete adjoint of hode_rollout_fwd over the recorded accepted steps
//                        (BASELINE.json north_star item 3).  The reference has no through-solver
//                        gradient (models/hybrid_ode_nn.py:248 returns a graph-free tensor); the
//                        semantics are those of autograd through the unrolled RK steps with the
//                        step sizes frozen, which is what tests/ check against
//                        (oracle/torch_restate.py).
//   rhs_vjp_kernel     : vector-Jacobian product of one batched f_physio + g_NN evaluation —
//                        the backward of HybridODENN.ode_residual, i.e. the only differentiable
//                        model call of the reference's loss (models/hybrid_ode_nn.py:327).
//
// One trajectory per thread, 128 trajectories per CTA, persistent CTAs (grid-stride over
// trajectory blocks).  Per RK stage, in reverse order, the cotangent of the stage derivative is
// pulled back through the mechanistic RHS (closed form, registers) and through the MLP:
//   * activations of every stage are stashed by the forward recomputation in a per-thread
//     global scratch column (coalesced, L1/L2 resident: 7 stages x L x H floats per thread);
//   * delta_{l-1} = relu'(a_{l-1}) * W_l^T delta_l  per thread from the shared-memory weights;
//   * dW_l += delta_l^T a_{l-1} is a [n_out x 128] x [128 x n_in] product over the CTA's
//     trajectories, staged through shared memory and accumulated into a per-CTA copy of the
//     gradient in shared memory — fixed summation order, no float atomics;
//   * per-CTA partial gradients go to the workspace and a second kernel adds them in CTA order,
//     so results are bit-reproducible run to run.
#include <math.h>

#include "hode_common.cuh"
#include "hode_kernels.h"

namespace hode {

namespace {

constexpr int ABLK = 128;      // trajectories (= threads) per CTA
constexpr int MAX_STAGES = 7;  // DP5(4): 6 stages + the FSAL evaluation used by dense output

struct AdjCtx {
  const float* img;  // shared: forward weight image (Wt[k][ldo] + bias per layer)
  float* dW;         // shared: gradient accumulator in the packed W layout of include/hode.h
  float* Abuf;       // shared: [ABLK][ld] layer inputs a_{l-1}, one row per trajectory
  float* D0;         // shared: [ABLK][ld] delta rows (ping)
  float* D1;         // shared: [ABLK][ld] delta rows (pong)
  float* stash;      // global: this thread's activation column; element e at stash[e * NT]
  size_t NT;         // total threads of the launch
  int ld, H, L, P, img_floats;
};

__device__ __forceinline__ float& stash_at(const AdjCtx& c, int st, int l, int k) {
  return c.stash[((size_t)(st * c.L + l) * c.H + k) * c.NT];
}

// ---- MLP forward, stashing every hidden activation (reference models/nn_residual.py:136-146) ----
__device__ void mlp_fwd_stash(const AdjCtx& c, int st, const float* x9, float* out6) {
  const int H = c.H, ldh = mlp_ldo(H);
  const float* w = c.img;
  {  // layer 0: 9 -> H, inputs in registers
    const float* bias = w + HODE_NN_IN * ldh;
    for (int j0 = 0; j0 < ldh; j0 += 8) {
      float acc[8];
      const float4 b0 = *reinterpret_cast<const float4*>(bias + j0);
      const float4 b1 = *reinterpret_cast<const float4*>(bias + j0 + 4);
      acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
      acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
#pragma unroll
      for (int k = 0; k < HODE_NN_IN; ++k) {
        const float4 w0 = *reinterpret_cast<const float4*>(w + k * ldh + j0);
        const float4 w1 = *reinterpret_cast<const float4*>(w + k * ldh + j0 + 4);
        acc[0] = fmaf(x9[k], w0.x, acc[0]); acc[1] = fmaf(x9[k], w0.y, acc[1]);
        acc[2] = fmaf(x9[k], w0.z, acc[2]); acc[3] = fmaf(x9[k], w0.w, acc[3]);
        acc[4] = fmaf(x9[k], w1.x, acc[4]); acc[5] = fmaf(x9[k], w1.y, acc[5]);
        acc[6] = fmaf(x9[k], w1.z, acc[6]); acc[7] = fmaf(x9[k], w1.w, acc[7]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (j0 + i < H) stash_at(c, st, 0, j0 + i) = fmaxf(acc[i], 0.f);
    }
    w = bias + ldh;
  }
  for (int l = 1; l < c.L; ++l) {  // hidden layers: H -> H, inputs from the stash
    const float* bias = w + H * ldh;
    for (int j0 = 0; j0 < ldh; j0 += 8) {
      float acc[8];
      const float4 b0 = *reinterpret_cast<const float4*>(bias + j0);
      const float4 b1 = *reinterpret_cast<const float4*>(bias + j0 + 4);
      acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
      acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
#pragma unroll 4
      for (int k = 0; k < H; ++k) {
        const float a = stash_at(c, st, l - 1, k);
        const float4 w0 = *reinterpret_cast<const float4*>(w + k * ldh + j0);
        const float4 w1 = *reinterpret_cast<const float4*>(w + k * ldh + j0 + 4);
        acc[0] = fmaf(a, w0.x, acc[0]); acc[1] = fmaf(a, w0.y, acc[1]);
        acc[2] = fmaf(a, w0.z, acc[2]); acc[3] = fmaf(a, w0.w, acc[3]);
        acc[4] = fmaf(a, w1.x, acc[4]); acc[5] = fmaf(a, w1.y, acc[5]);
        acc[6] = fmaf(a, w1.z, acc[6]); acc[7] = fmaf(a, w1.w, acc[7]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (j0 + i < H) stash_at(c, st, l, j0 + i) = fmaxf(acc[i], 0.f);
    }
    w = bias + ldh;
  }
  {  // output layer: H -> 6 (image padded to 8 columns)
    const float* bias = w + H * 8;
    float acc[8];
    const float4 b0 = *reinterpret_cast<const float4*>(bias);
    const float4 b1 = *reinterpret_cast<const float4*>(bias + 4);
    acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
    acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
#pragma unroll 4
    for (int k = 0; k < H; ++k) {
      const float a = stash_at(c, st, c.L - 1, k);
      const float4 w0 = *reinterpret_cast<const float4*>(w + k * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(w + k * 8 + 4);
      acc[0] = fmaf(a, w0.x, acc[0]); acc[1] = fmaf(a, w0.y, acc[1]);
      acc[2] = fmaf(a, w0.z, acc[2]); acc[3] = fmaf(a, w0.w, acc[3]);
      acc[4] = fmaf(a, w1.x, acc[4]); acc[5] = fmaf(a, w1.y, acc[5]);
    }
#pragma unroll
    for (int i = 0; i < NS; ++i) out6[i] = acc[i];
  }
}

// ---- dW += D^T A over the CTA's 128 trajectories (block-collective, after a barrier) -----------
// D rows hold delta_l (n_out values), A rows hold a_{l-1} (n_in values); dWw is [n_out][n_in]
// row-major (the reference's weight layout), dWb the bias gradient.  Every output element is
// owned by one thread and summed over trajectories in index order: deterministic.
__device__ void gemm_acc(float* dWw, float* dWb, const float* __restrict__ D,
                         const float* __restrict__ A, int n_out, int n_in, int ld, int tid) {
  if ((n_out & 7) == 0 && (n_in & 3) == 0) {
    const int nkb = n_in >> 2, nblk = (n_out >> 3) * nkb;
    for (int blk = tid; blk < nblk; blk += ABLK) {
      const int jb = blk / nkb, kb = blk - jb * nkb;
      float acc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;
      const float* dp_ = D + jb * 8;
      const float* ap = A + kb * 4;
#pragma unroll 2
      for (int r = 0; r < ABLK; ++r) {
        const float4 d0 = *reinterpret_cast<const float4*>(dp_ + r * ld);
        const float4 d1 = *reinterpret_cast<const float4*>(dp_ + r * ld + 4);
        const float4 a = *reinterpret_cast<const float4*>(ap + r * ld);
        const float dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        const float aa[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[i][q] = fmaf(dd[i], aa[q], acc[i][q]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) dWw[(jb * 8 + i) * n_in + kb * 4 + q] += acc[i][q];
    }
  } else {
    for (int e = tid; e < n_out * n_in; e += ABLK) {
      const int j = e / n_in, k = e - j * n_in;
      float acc = 0.f;
      for (int r = 0; r < ABLK; ++r) acc = fmaf(D[r * ld + j], A[r * ld + k], acc);
      dWw[e] += acc;
    }
  }
  for (int j = tid; j < n_out; j += ABLK) {
    float acc = 0.f;
    for (int r = 0; r < ABLK; ++r) acc += D[r * ld + j];
    dWb[j] += acc;
  }
}

// ---- MLP backward for one stage (block-collective: every thread of the CTA must call it) ---------
// g6: cotangent of the 6 network outputs of this thread's trajectory (zeros for idle threads).
// gx[1..7]: cotangent of the input features 1..7 (the states and the duplicated GLP1 column).
__device__ void mlp_bwd(const AdjCtx& c, int st, const float* x9, const float* g6, float* gx) {
  const int tid = threadIdx.x, H = c.H, ld = c.ld;
  float* Dbase = c.D0;
  float* Dnext = c.D1;
  float* Ar = c.Abuf + tid * ld;
  {
    float* Dc = Dbase + tid * ld;
#pragma unroll
    for (int i = 0; i < 8; ++i) Dc[i] = (i < NS) ? g6[i] : 0.f;
  }
  int pk_off = c.P, img_off = c.img_floats;
  for (int l = c.L; l >= 0; --l) {
    const int n_out = (l == c.L) ? NS : H;
    const int n_in = (l == 0) ? HODE_NN_IN : H;
    const int ldo = mlp_ldo(n_out);
    pk_off -= n_out * n_in + n_out;
    img_off -= n_in * ldo + ldo;
    // own row of layer inputs
    if (l == 0) {
#pragma unroll
      for (int k = 0; k < 12; ++k) Ar[k] = (k < HODE_NN_IN) ? x9[k] : 0.f;
    } else {
      for (int k = 0; k < H; ++k) Ar[k] = stash_at(c, st, l - 1, k);
      for (int k = H; k < ((H + 3) & ~3); ++k) Ar[k] = 0.f;
    }
    __syncthreads();
    gemm_acc(c.dW + pk_off, c.dW + pk_off + n_out * n_in, Dbase, c.Abuf, n_out, n_in, ld, tid);
    // pull delta back through W_l (own rows only)
    const float* Dc = Dbase + tid * ld;
    const float* wt = c.img + img_off;
    if (l > 0) {
      float* Dn = Dnext + tid * ld;
      for (int k0 = 0; k0 < H; k0 += 4) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j4 = 0; j4 < ldo; j4 += 4) {
          const float4 d = *reinterpret_cast<const float4*>(Dc + j4);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (k0 + q < H) {
              const float4 w = *reinterpret_cast<const float4*>(wt + (k0 + q) * ldo + j4);
              acc[q] = fmaf(d.x, w.x, fmaf(d.y, w.y, fmaf(d.z, w.z, fmaf(d.w, w.w, acc[q]))));
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (k0 + q < H) Dn[k0 + q] = (Ar[k0 + q] > 0.f) ? acc[q] : 0.f;
      }
      for (int k = H; k < mlp_ldo(H); ++k) Dn[k] = 0.f;
    } else {
#pragma unroll
      for (int k = 1; k <= 7; ++k) {
        float acc = 0.f;
        for (int j4 = 0; j4 < ldo; j4 += 4) {
          const float4 d = *reinterpret_cast<const float4*>(Dc + j4);
          const float4 w = *reinterpret_cast<const float4*>(wt + k * ldo + j4);
          acc = fmaf(d.x, w.x, fmaf(d.y, w.y, fmaf(d.z, w.z, fmaf(d.w, w.w, acc))));
        }
        gx[k] = acc;
      }
    }
    __syncthreads();
    float* tmp = Dbase; Dbase = Dnext; Dnext = tmp;
  }
}

// ---- closed-form VJP of f_physio (reference models/ode_core.py:122-161) -----------------------------
// c: cotangent of the 6 derivatives; gy += J_y^T c; gth += J_theta^T c.
__device__ __forceinline__ void rhs_mech_vjp(const Theta& p, const float* y, float GD, bool gd_present,
                                             const float* c, float* gy, float* gth) {
  const float G = y[0], I = y[1], Glu = y[2], GLP1 = y[3], FFA = y[5];
  const float Pi = 1.0f + p.rho * GLP1;
  const float dG = G - p.G_b, dI = I - p.I_b, dGlu = Glu - p.Glu_b;
  const float den_e = p.EC_50 + GLP1, inv_e = 1.0f / den_e;
  const float frac_e = GLP1 * inv_e;
  const float ge = p.E_max * frac_e;
  const float den_m = p.K_m + G, inv_m = 1.0f / den_m;
  float r = 0.f, dr_du = 0.f, dr_dv = 0.f, u = 0.f;
  const float v = p.igd_pow;
  if (gd_present) {
    u = powf(GD, p.g);
    const float inv = 1.0f / (v + u);
    r = u * inv;
    dr_du = v * inv * inv;
    dr_dv = -u * inv * inv;
  }
  const float k_GE = p.k_GE0 * (1.0f - r);
  const float lin5 = -p.p_7 - p.p_8 * I + p.p_9 * G;
  // state cotangents
  gy[0] += c[1] * Pi * p.a_GI + c[3] * p.V_max * p.K_m * inv_m * inv_m + c[5] * FFA * p.p_9 - c[0] * k_GE;
  gy[1] += -c[1] * p.k_I - c[5] * FFA * p.p_8 - 0.01f * c[0];
  gy[2] += -c[2] * ge + 0.005f * c[0];
  gy[3] += c[1] * p.rho * p.a_GI * dG - c[2] * dGlu * p.E_max * p.EC_50 * inv_e * inv_e - c[3] * p.k_L;
  gy[5] += c[5] * lin5;
  // parameter cotangents (theta order of include/hode.h)
  gth[0] += c[1] * Pi * dG;                                   // a_GI
  gth[1] += -c[1] * dI;                                       // k_I
  gth[2] += c[1] * GLP1 * p.a_GI * dG;                        // rho
  gth[3] += -c[1] * Pi * p.a_GI;                              // G_b
  gth[4] += c[1] * p.k_I + 0.01f * c[0];                      // I_b
  gth[5] += -c[2] * dGlu * frac_e;                            // E_max
  gth[6] += c[2] * dGlu * p.E_max * GLP1 * inv_e * inv_e;     // EC_50
  gth[7] += c[2] * ge - 0.005f * c[0];                        // Glu_b
  gth[8] += c[3] * G * inv_m;                                 // V_max
  gth[9] += -c[3] * p.V_max * G * inv_m * inv_m;              // K_m
  gth[10] += -c[3] * GLP1;                                    // k_L
  gth[11] += -c[0] * G * (1.0f - r);                          // k_GE0
  if (gd_present) {
    const float g_r = c[0] * p.k_GE0 * G;                     // d d0 / d r
    const float dv_dIGD = p.g * powf(p.IGD_50, p.g - 1.0f);
    gth[12] += g_r * dr_dv * dv_dIGD;                         // IGD_50
    float dg = dr_dv * v * logf(p.IGD_50);
    if (GD > 0.f) dg += dr_du * u * logf(GD);
    gth[13] += g_r * dg;                                      // g
  }
  gth[14] += -c[5] * FFA;                                     // p_7
  gth[15] += -c[5] * FFA * I;                                 // p_8
  gth[16] += c[5] * FFA * G;                                  // p_9
}

// ---- Butcher tableaux ------------------------------------------------------------------------------------
template <int SOLVER> struct Tab;
template <> struct Tab<HODE_SOLVER_RK4> {
  static constexpr int N = 4;
  __device__ static constexpr float a(int i, int j) {
    return (i == 1 && j == 0) ? 0.5f : (i == 2 && j == 1) ? 0.5f : (i == 3 && j == 2) ? 1.0f : 0.f;
  }
  __device__ static constexpr float b(int i) { return (i == 0 || i == 3) ? (1.0f / 6.0f) : (1.0f / 3.0f); }
  __device__ static constexpr float c(int i) { return i == 0 ? 0.f : (i == 3 ? 1.0f : 0.5f); }
  __device__ static constexpr float p(int, int) { return 0.f; }  // no dense output
};
template <> struct Tab<HODE_SOLVER_DOPRI5> {
  static constexpr int N = 7;  // stage 7 = f(t_new, y_new): only dense output depends on it
  __device__ static constexpr float b(int i) {
    return i == 0 ? dp::b1 : i == 2 ? dp::b3 : i == 3 ? dp::b4 : i == 4 ? dp::b5 : i == 5 ? dp::b6 : 0.f;
  }
  __device__ static constexpr float a(int i, int j) {
    switch (i * 8 + j) {
      case 1 * 8 + 0: return dp::a21;
      case 2 * 8 + 0: return dp::a31; case 2 * 8 + 1: return dp::a32;
      case 3 * 8 + 0: return dp::a41; case 3 * 8 + 1: return dp::a42; case 3 * 8 + 2: return dp::a43;
      case 4 * 8 + 0: return dp::a51; case 4 * 8 + 1: return dp::a52; case 4 * 8 + 2: return dp::a53;
      case 4 * 8 + 3: return dp::a54;
      case 5 * 8 + 0: return dp::a61; case 5 * 8 + 1: return dp::a62; case 5 * 8 + 2: return dp::a63;
      case 5 * 8 + 3: return dp::a64; case 5 * 8 + 4: return dp::a65;
      default: return (i == 6) ? b(j) : 0.f;
    }
  }
  __device__ static constexpr float c(int i) {
    return i == 0 ? 0.f : i == 1 ? dp::c2 : i == 2 ? dp::c3 : i == 3 ? dp::c4 : i == 4 ? dp::c5 : 1.0f;
  }
  // dense-output matrix P (rk.py:554-565), row i (0-based stage), column j
  __device__ static constexpr float p(int i, int j) {
    switch (i * 4 + j) {
      case 0: return dp::p11; case 1: return dp::p12; case 2: return dp::p13; case 3: return dp::p14;
      case 2 * 4 + 1: return dp::p32; case 2 * 4 + 2: return dp::p33; case 2 * 4 + 3: return dp::p34;
      case 3 * 4 + 1: return dp::p42; case 3 * 4 + 2: return dp::p43; case 3 * 4 + 3: return dp::p44;
      case 4 * 4 + 1: return dp::p52; case 4 * 4 + 2: return dp::p53; case 4 * 4 + 3: return dp::p54;
      case 5 * 4 + 1: return dp::p62; case 5 * 4 + 2: return dp::p63; case 5 * 4 + 3: return dp::p64;
      case 6 * 4 + 1: return dp::p72; case 6 * 4 + 2: return dp::p73; case 6 * 4 + 3: return dp::p74;
      default: return 0.f;
    }
  }
};

}  // namespace

struct AdjArgs {
  RolloutArgs R;
  const float* grad_traj;  // [S,B,T,6]
  float* grad_y0;          // [S,B,6] or nullptr
  float* partials;         // [gridDim.y * gridDim.x][P + 17]
  float* scratch;          // activation stash
  // rhs_vjp mode
  const float* tt;         // [B]
  const float* state;      // [B,6]
  const float* grad_out;   // [B,6]
  float* grad_state;       // [B,6] or nullptr
  int has_nn;
};

namespace {

// shared-memory carve-up common to both kernels
__device__ void adj_setup(AdjCtx& c, float* sm, const AdjArgs& G, int s) {
  const RolloutArgs& A = G.R;
  c.H = A.H; c.L = A.L; c.P = A.P;
  c.img_floats = 0; c.ld = 0;
  c.img = nullptr; c.dW = nullptr; c.Abuf = c.D0 = c.D1 = nullptr;
  if (G.has_nn) {
    c.img_floats = mlp_image_floats(A.H, A.L);
    const int imgp = (c.img_floats + 3) & ~3;
    stage_mlp_image(sm, A.W + (size_t)s * A.P, A.H, A.L);
    c.img = sm; sm += imgp;
    c.dW = sm; sm += (A.P + 3) & ~3;
    for (int i = threadIdx.x; i < A.P; i += blockDim.x) c.dW[i] = 0.f;
    const int hp = ((A.H > 16 ? A.H : 16) + 7) & ~7;
    c.ld = hp + 4;
    c.Abuf = sm; sm += ABLK * c.ld;
    c.D0 = sm; sm += ABLK * c.ld;
    c.D1 = sm; sm += ABLK * c.ld;
  }
  const size_t gt = ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * ABLK + threadIdx.x;
  c.NT = (size_t)gridDim.x * gridDim.y * ABLK;
  c.stash = G.scratch ? G.scratch + gt : nullptr;
}

// deterministic CTA reduction of the per-thread theta gradients + dump of the CTA's partials
__device__ void adj_finish(const AdjCtx& c, const AdjArgs& G, float* red /*[17][ABLK]*/, const float* gth) {
  const int tid = threadIdx.x;
  if (G.has_nn) red = c.Abuf;   // the staging rows are free once the last stage has been pulled back
  __syncthreads();
#pragma unroll
  for (int i = 0; i < HODE_N_THETA; ++i) red[i * ABLK + tid] = gth[i];
  __syncthreads();
  float* out = G.partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (size_t)(G.R.P + HODE_N_THETA);
  if (G.has_nn)
    for (int i = tid; i < G.R.P; i += ABLK) out[i] = c.dW[i];
  if (tid < HODE_N_THETA) {
    float acc = 0.f;
    for (int r = 0; r < ABLK; ++r) acc += red[tid * ABLK + r];
    out[G.R.P + tid] = acc;
  }
}

// One stage evaluation for the forward recomputation: d = f_physio + g_NN at (te, ys).
template <bool HAS_NN>
__device__ __forceinline__ void stage_eval(const AdjCtx& c, const Theta& th, TrajInputs& in, int st,
                                           double te, const float* ys, float* d, float& tv_out,
                                           float& gd_out) {
  const float t32 = (float)te;
  int idx = 0;
  if (any_series(in)) idx = grid_index_from(in, t32, in.cur);
  const float meal = input_channel(in, HODE_CH_MEAL, t32, idx);
  const float tvns = input_channel(in, HODE_CH_TVNS, t32, idx);
  const float gd = input_channel(in, HODE_CH_GD, t32, idx);
  tv_out = tvns; gd_out = gd;
  rhs_mech(th, ys, meal, gd, in.mode[HODE_CH_GD] != HODE_IN_ABSENT, d);
  if (HAS_NN) {
    float x[HODE_NN_IN], r[NS];
    x[0] = t32;
#pragma unroll
    for (int i = 0; i < NS; ++i) x[1 + i] = ys[i];
    x[7] = ys[3];
    x[8] = tvns;
    mlp_fwd_stash(c, st, x, r);
#pragma unroll
    for (int i = 0; i < NS; ++i) d[i] = __fadd_rn(d[i], r[i]);
  }
}

// Pull the cotangent gk of one stage derivative back to the stage state (block-collective).
template <bool HAS_NN>
__device__ __forceinline__ void stage_vjp(const AdjCtx& c, const Theta& th, bool gd_present, int st,
                                          double te, const float* ys, float tvns, float gd,
                                          const float* gk, float* gys, float* gth, bool mech) {
#pragma unroll
  for (int i = 0; i < NS; ++i) gys[i] = 0.f;
  if (mech) rhs_mech_vjp(th, ys, gd, gd_present, gk, gys, gth);
  if (HAS_NN) {
    float x[HODE_NN_IN], gx[8];
    x[0] = (float)te;
#pragma unroll
    for (int i = 0; i < NS; ++i) x[1 + i] = ys[i];
    x[7] = ys[3];
    x[8] = tvns;
    mlp_bwd(c, st, x, gk, gx);
#pragma unroll
    for (int i = 0; i < NS; ++i) gys[i] += gx[1 + i];
    gys[3] += gx[7];
  }
}

__device__ __forceinline__ void bind_inputs(TrajInputs& in, const RolloutArgs& A, const float* t_shared, long b) {
  in.T = A.T;
  in.cur = 0;
  in.t_obs = A.t_per_traj ? A.t_obs + b * A.T : (t_shared ? t_shared : A.t_obs);
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    in.mode[ch] = A.in_mode[ch];
    in.u[ch] = in.mode[ch] == HODE_IN_SERIES ? A.u[ch] + b * A.T
             : in.mode[ch] == HODE_IN_CONST ? A.u[ch] + b : nullptr;
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// Discrete adjoint of the rollout.  grid = (ctas per parameter set, S), block = 128.
// ---------------------------------------------------------------------------------------------------
template <int SOLVER, bool HAS_NN>
__global__ void __launch_bounds__(ABLK, 1) rollout_bwd_kernel(const AdjArgs G) {
  using TB = Tab<SOLVER>;
  constexpr int N = TB::N;
  extern __shared__ __align__(16) float smem[];
  __shared__ int s_nmax;
  const RolloutArgs& A = G.R;
  const int tid = threadIdx.x, s = blockIdx.y, T = A.T;

  float* red = smem;                         // [17][ABLK] (only carved out when there is no MLP)
  float* sm = smem + (G.has_nn ? 0 : HODE_N_THETA * ABLK);
  const float* t_shared = nullptr;
  if (!A.t_per_traj && T <= HODE_SIMT_MAX_SHARED_T) {
    for (int i = tid; i < T; i += ABLK) sm[i] = A.t_obs[i];
    t_shared = sm;
    sm += (T + 3) & ~3;
  }
  AdjCtx c;
  adj_setup(c, sm, G, s);
  __syncthreads();

  const Theta th = load_theta(A.theta + (size_t)s * HODE_N_THETA);
  const bool gd_present = A.in_mode[HODE_CH_GD] != HODE_IN_ABSENT;
  const int nsub = A.n_substeps > 0 ? A.n_substeps : 1;
  float gth[HODE_N_THETA];
#pragma unroll
  for (int i = 0; i < HODE_N_THETA; ++i) gth[i] = 0.f;

  for (long blk = blockIdx.x; blk * ABLK < A.B; blk += gridDim.x) {
    const long b = blk * ABLK + tid;
    const bool valid = b < A.B;
    const long bs = valid ? b : 0;
    const long unit = (long)s * A.B + bs;
    int n = valid ? A.save_n[unit] : 0;
    const bool ok = valid && n >= 0;   // failed trajectories (negative count) carry no gradient
    if (n < 0) n = 0;
    TrajInputs in;
    bind_inputs(in, A, t_shared, bs);
    const float* gtraj = G.grad_traj + (size_t)unit * T * NS;
    const double t_bound = (double)in.t_obs[T - 1];

    if (tid == 0) s_nmax = 0;
    __syncthreads();
    if (n > 0) atomicMax(&s_nmax, n);
    __syncthreads();
    const int nmax = s_nmax;

    float lam[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) lam[i] = 0.f;
    int ei = T - 1;  // DP5: next observation (walking backwards) whose gradient is still unassigned

    for (int it = 0; it < nmax; ++it) {
      const int sidx = n - 1 - it;
      const bool act = ok && sidx >= 0;
      // ---- step data ------------------------------------------------------------------------------
      double t = (double)in.t_obs[0], t_new = t, h = 0.0;
      float y[NS];
#pragma unroll
      for (int i = 0; i < NS; ++i) y[i] = 0.f;
      if (act) {
        const float* rec = step_rec(A, unit, sidx);
        float h_rec;
        step_rec_load(rec, t, h_rec, y, nullptr);
        if (SOLVER == HODE_SOLVER_RK4) {
          h = (double)h_rec;
          t_new = t + h;
        } else {
          t_new = (sidx + 1 < n) ? step_rec_t(rec + A.rec_floats) : t_bound;
          h = t_new - t;
        }
      }
      const float hf = (float)h;
      // input cursor: grid points known to be < any stage time of this step (one point of slack)
      {
        int lo = 0, hi = T;
        const float t32 = (float)t;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (in.t_obs[mid] < t32) lo = mid + 1; else hi = mid; }
        in.cur = lo > 0 ? lo - 1 : 0;
      }
      // ---- forward recomputation of the stages (activations -> stash) ---------------------------
      float k[N][NS], tv[N], gdv[N];
#pragma unroll
      for (int i = 0; i < N; ++i) {
        float ys[NS];
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) {
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < i; ++j)
            if (TB::a(i, j) != 0.f) acc = fmaf(TB::a(i, j), k[j][cc], acc);
          ys[cc] = fmaf(hf, acc, y[cc]);
        }
        const double te = (i == 0) ? t : (TB::c(i) == 1.0f ? t_new : t + (double)TB::c(i) * h);
        stage_eval<HAS_NN>(c, th, in, i, te, ys, k[i], tv[i], gdv[i]);
      }
      // ---- cotangents entering this step ----------------------------------------------------------
      float gy[NS], gk[N][NS];
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) gk[i][cc] = 0.f;
      if (SOLVER == HODE_SOLVER_RK4) {
        if (act && (sidx + 1) % nsub == 0) {
          const float* g = gtraj + (size_t)((sidx + 1) / nsub) * NS;
#pragma unroll
          for (int cc = 0; cc < NS; ++cc) lam[cc] += g[cc];
        }
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) gy[cc] = act ? lam[cc] : 0.f;
      } else {
        float gnew[NS];
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) { gnew[cc] = act ? lam[cc] : 0.f; gy[cc] = 0.f; }
        if (act) {
          while (ei >= 0 && (double)in.t_obs[ei] > t) {
            const double te = (double)in.t_obs[ei];
            const float* g = gtraj + (size_t)ei * NS;
            if (te >= t_new) {
              if (te == t_new) {
#pragma unroll
                for (int cc = 0; cc < NS; ++cc) gnew[cc] += g[cc];
              }
            } else {
              const float x = (float)((te - t) / h);
#pragma unroll
              for (int cc = 0; cc < NS; ++cc) gy[cc] += g[cc];
#pragma unroll
              for (int i = 0; i < N; ++i) {
                if (i == 1) continue;
                const float wgt = hf * x * fmaf(x, fmaf(x, fmaf(x, TB::p(i, 3), TB::p(i, 2)), TB::p(i, 1)), TB::p(i, 0));
#pragma unroll
                for (int cc = 0; cc < NS; ++cc) gk[i][cc] = fmaf(wgt, g[cc], gk[i][cc]);
              }
            }
            --ei;
          }
        }
        // y_new carries lam plus the gradient of observations that sit exactly on t_new; it is the
        // input of stage 7 (a_7j = b_j), so it is folded into that stage's state cotangent below
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) {
          gy[cc] += gnew[cc];
#pragma unroll
          for (int j = 0; j < N - 1; ++j)
            if (TB::b(j) != 0.f) gk[j][cc] = fmaf(hf * TB::b(j), gnew[cc], gk[j][cc]);
        }
      }
      if (SOLVER == HODE_SOLVER_RK4) {
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
          for (int cc = 0; cc < NS; ++cc) gk[j][cc] = hf * TB::b(j) * gy[cc];
      }
      // ---- reverse sweep over the stages -----------------------------------------------------------
#pragma unroll
      for (int i = N - 1; i >= 0; --i) {
        float ys[NS], gys[NS];
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) {
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < i; ++j)
            if (TB::a(i, j) != 0.f) acc = fmaf(TB::a(i, j), k[j][cc], acc);
          ys[cc] = fmaf(hf, acc, y[cc]);
        }
        const double te = (i == 0) ? t : (TB::c(i) == 1.0f ? t_new : t + (double)TB::c(i) * h);
        stage_vjp<HAS_NN>(c, th, gd_present, i, te, ys, tv[i], gdv[i], gk[i], gys, gth, true);
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) {
          gy[cc] += gys[cc];
#pragma unroll
          for (int j = 0; j < i; ++j)
            if (TB::a(i, j) != 0.f) gk[j][cc] = fmaf(hf * TB::a(i, j), gys[cc], gk[j][cc]);
        }
      }
      if (act) {
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) lam[cc] = gy[cc];
      }
    }
    // observations at or before the first step start are copies of y0
    if (ok) {
      if (SOLVER == HODE_SOLVER_RK4) {
        const float* g = gtraj;
#pragma unroll
        for (int cc = 0; cc < NS; ++cc) lam[cc] += g[cc];
      } else {
        const double t0 = (double)in.t_obs[0];
        for (; ei >= 0; --ei) {
          if ((double)in.t_obs[ei] > t0) continue;   // unreachable for a completed trajectory
          const float* g = gtraj + (size_t)ei * NS;
#pragma unroll
          for (int cc = 0; cc < NS; ++cc) lam[cc] += g[cc];
        }
      }
    }
    if (G.grad_y0 && valid) {
      float* o = G.grad_y0 + (size_t)unit * NS;
#pragma unroll
      for (int cc = 0; cc < NS; ++cc) o[cc] = ok ? lam[cc] : 0.f;
    }
  }
  adj_finish(c, G, red, gth);
}

// ---------------------------------------------------------------------------------------------------
// VJP of one batched RHS evaluation.  grid = (ctas, 1), block = 128.
// ---------------------------------------------------------------------------------------------------
template <bool HAS_NN>
__global__ void __launch_bounds__(ABLK, 1) rhs_vjp_kernel(const AdjArgs G) {
  extern __shared__ __align__(16) float smem[];
  const RolloutArgs& A = G.R;
  const int tid = threadIdx.x;
  float* red = smem;
  float* sm = smem + (G.has_nn ? 0 : HODE_N_THETA * ABLK);
  AdjCtx c;
  adj_setup(c, sm, G, 0);
  __syncthreads();
  const Theta th = load_theta(A.theta);
  const bool gd_present = A.in_mode[HODE_CH_GD] != HODE_IN_ABSENT;
  const bool mech = A.rhs_part != HODE_RHS_NN_ONLY;
  float gth[HODE_N_THETA];
#pragma unroll
  for (int i = 0; i < HODE_N_THETA; ++i) gth[i] = 0.f;
  for (long blk = blockIdx.x; blk * ABLK < A.B; blk += gridDim.x) {
    const long b = blk * ABLK + tid;
    const bool valid = b < A.B;
    const long bs = valid ? b : 0;
    float y[NS], gk[NS], gys[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      y[i] = G.state[bs * NS + i];
      gk[i] = valid ? G.grad_out[bs * NS + i] : 0.f;
    }
    const float t32 = G.tt[bs];
    const float tvns = A.in_mode[HODE_CH_TVNS] != HODE_IN_ABSENT ? A.u[HODE_CH_TVNS][bs] : 0.f;
    const float gd = gd_present ? A.u[HODE_CH_GD][bs] : 0.f;
    if (HAS_NN) {
      float x[HODE_NN_IN], r[NS];
      x[0] = t32;
#pragma unroll
      for (int i = 0; i < NS; ++i) x[1 + i] = y[i];
      x[7] = y[3];
      x[8] = tvns;
      mlp_fwd_stash(c, 0, x, r);
    }
    stage_vjp<HAS_NN>(c, th, gd_present, 0, (double)t32, y, tvns, gd, gk, gys, gth, mech);
    if (G.grad_state && valid) {
#pragma unroll
      for (int i = 0; i < NS; ++i) G.grad_state[b * NS + i] = gys[i];
    }
  }
  adj_finish(c, G, red, gth);
}

// out[s][i] = sum over the CTAs of parameter set s, in CTA order (deterministic)
__global__ void reduce_partials_kernel(const float* __restrict__ partials, int ctas_per_set, int P,
                                       float* __restrict__ grad_W, float* __restrict__ grad_theta) {
  const int s = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = P + HODE_N_THETA;
  if (i >= n) return;
  const float* p = partials + (size_t)s * ctas_per_set * n + i;
  float acc = 0.f;
  for (int cta = 0; cta < ctas_per_set; ++cta) acc += p[(size_t)cta * n];
  if (i < P) { if (grad_W) grad_W[(size_t)s * P + i] = acc; }
  else if (grad_theta) grad_theta[(size_t)s * HODE_N_THETA + (i - P)] = acc;
}

cudaError_t launch_reduce_partials(const float* partials, int ctas_per_set, int S, int P, float* grad_W,
                                   float* grad_theta, cudaStream_t stream) {
  const int n = P + HODE_N_THETA;
  count_launch();
  reduce_partials_kernel<<<dim3((n + 255) / 256, S), 256, 0, stream>>>(partials, ctas_per_set, P, grad_W, grad_theta);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static size_t adj_smem_bytes(int H, int L, int P, bool has_nn, int T_shared) {
  size_t floats = (size_t)(has_nn ? 0 : HODE_N_THETA * ABLK) + ((T_shared + 3) & ~3);
  if (has_nn) {
    floats += (mlp_image_floats(H, L) + 3) & ~3;
    floats += (P + 3) & ~3;
    const int hp = ((H > 16 ? H : 16) + 7) & ~7;
    floats += (size_t)3 * ABLK * (hp + 4);
  }
  return floats * sizeof(float);
}

AdjPlan adj_plan(int B, int S, int H, int L, int P, bool has_nn, int T, int t_per_traj, int n_stages) {
  AdjPlan p{};
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms < 1) sms = 148;
  const long blocks = ((long)B + ABLK - 1) / ABLK;
  long gx = sms / (S > 0 ? S : 1);
  if (gx < 1) gx = 1;
  if (gx > blocks) gx = blocks;
  if (gx < 1) gx = 1;
  // balance: every CTA walks the same number of trajectory blocks (fewer, equally loaded CTAs
  // finish at the same time as more, unequally loaded ones, with less scratch)
  {
    const long rounds = blocks > 0 ? (blocks + gx - 1) / gx : 1;
    gx = blocks > 0 ? (blocks + rounds - 1) / rounds : 1;
  }
  p.grid_x = (int)gx;
  p.grid_y = S;
  const int Tsh = (!t_per_traj && T <= HODE_SIMT_MAX_SHARED_T) ? T : 0;
  p.smem = adj_smem_bytes(H, L, P, has_nn, Tsh);
  p.partial_floats = (size_t)gx * S * (size_t)(P + HODE_N_THETA);
  p.scratch_floats = has_nn ? (size_t)gx * S * ABLK * (size_t)n_stages * L * H : 0;
  return p;
}

cudaError_t launch_rollout_bwd(const RolloutArgs& A, bool has_nn, const float* grad_traj, float* grad_y0,
                               float* grad_theta, float* grad_W, void* workspace, cudaStream_t stream) {
  const int nst = A.solver == HODE_SOLVER_RK4 ? 4 : MAX_STAGES;
  const AdjPlan p = adj_plan(A.B, A.S, A.H, A.L, A.P, has_nn, A.T, A.t_per_traj, nst);
  if (p.smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  AdjArgs G{};
  G.R = A;
  G.grad_traj = grad_traj;
  G.grad_y0 = grad_y0;
  G.partials = reinterpret_cast<float*>(workspace);
  G.scratch = G.partials + ((p.partial_floats + 63) & ~(size_t)63);
  G.has_nn = has_nn ? 1 : 0;
  dim3 grid(p.grid_x, p.grid_y);
  auto launch = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return e;
    count_launch();
    kern<<<grid, ABLK, p.smem, stream>>>(G);
    return cudaGetLastError();
  };
  cudaError_t e;
  if (A.solver == HODE_SOLVER_RK4)
    e = has_nn ? launch(rollout_bwd_kernel<HODE_SOLVER_RK4, true>) : launch(rollout_bwd_kernel<HODE_SOLVER_RK4, false>);
  else
    e = has_nn ? launch(rollout_bwd_kernel<HODE_SOLVER_DOPRI5, true>) : launch(rollout_bwd_kernel<HODE_SOLVER_DOPRI5, false>);
  if (e != cudaSuccess) return e;
  const int n = A.P + HODE_N_THETA;
  count_launch();
  reduce_partials_kernel<<<dim3((n + 255) / 256, A.S), 256, 0, stream>>>(G.partials, p.grid_x, A.P, grad_W, grad_theta);
  return cudaGetLastError();
}

cudaError_t launch_rhs_vjp(const RolloutArgs& A, bool has_nn, const float* t, const float* state,
                           const float* grad_out, float* grad_state, float* grad_theta, float* grad_W,
                           void* workspace, cudaStream_t stream) {
  const AdjPlan p = adj_plan(A.B, 1, A.H, A.L, A.P, has_nn, 0, 1, 1);
  if (p.smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  AdjArgs G{};
  G.R = A;
  G.R.S = 1;
  G.partials = reinterpret_cast<float*>(workspace);
  G.scratch = G.partials + ((p.partial_floats + 63) & ~(size_t)63);
  G.tt = t; G.state = state; G.grad_out = grad_out; G.grad_state = grad_state;
  G.has_nn = has_nn ? 1 : 0;
  dim3 grid(p.grid_x, 1);
  auto launch = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return e;
    count_launch();
    kern<<<grid, ABLK, p.smem, stream>>>(G);
    return cudaGetLastError();
  };
  cudaError_t e = has_nn ? launch(rhs_vjp_kernel<true>) : launch(rhs_vjp_kernel<false>);
  if (e != cudaSuccess) return e;
  const int n = A.P + HODE_N_THETA;
  count_launch();
  reduce_partials_kernel<<<dim3((n + 255) / 256, 1), 256, 0, stream>>>(G.partials, p.grid_x, A.P, grad_W, grad_theta);
  return cudaGetLastError();
}

size_t adj_workspace_bytes(const AdjPlan& p) {
  return (((p.partial_floats + 63) & ~(size_t)63) + p.scratch_floats) * sizeof(float);
}

}  // namespace hode
