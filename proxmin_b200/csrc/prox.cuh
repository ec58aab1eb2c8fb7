// Proximal-operator chains and the fused "forward step + prox + norms" kernel family.
//
// Replaces proxmin/operators.py:20-160 (elementwise maps), AlternatingProjections
// (:187-211), the update line of pgm (algorithms.py:107-108), the convergence norms
// (algorithms.py:130-133, utils.py:257-260) and the proximal sub-iteration of adaprox
// (algorithms.py:386-393).  All of these are HBM-bound elementwise passes, so one
// kernel family does "load -> transform -> prox chain -> store -> norms" in a single
// pass whenever the chain allows it.
#pragma once
#include "common.cuh"

struct ProxChain {  // device-side copy of pmx_prox (passed by value as a kernel argument)
  int n;
  int op[PMX_MAX_OPS];
  int rel[PMX_MAX_OPS];
  int axis[PMX_MAX_OPS];
  float thr[PMX_MAX_OPS];
};

static inline ProxChain make_chain(const pmx_prox* p) {
  ProxChain c;
  memset(&c, 0, sizeof(c));
  if (!p) return c;
  c.n = p->n_ops;
  for (int i = 0; i < p->n_ops && i < PMX_MAX_OPS; ++i) {
    c.op[i] = p->ops[i].op;
    c.rel[i] = p->ops[i].relative;
    c.axis[i] = p->ops[i].axis;
    c.thr[i] = p->ops[i].thresh;
  }
  return c;
}

// position of the first UNITY op at or after `from` (or c.n)
__host__ __device__ inline int chain_next_unity(const ProxChain& c, int from) {
  int i = from;
  while (i < c.n && c.op[i] != PMX_OP_UNITY) ++i;
  return i;
}
static inline int chain_unity_axis(const ProxChain& c) {  // -1 none, 0/1 the single axis used, 2 mixed
  int ax = -1;
  for (int i = 0; i < c.n; ++i)
    if (c.op[i] == PMX_OP_UNITY) {
      if (ax == -1) ax = c.axis[i];
      else if (ax != c.axis[i]) ax = 2;
    }
  return ax;
}

// Principal branch of the Lambert W function for real z >= 0 (scipy.special.lambertw of operators.py:183, whose
// argument here is always a positive real): Halley iterations in fp64 from log1p(z) / the asymptotic series.
// NOT inlined, like pmx_maxent_elem below: the fp64 exp / log expansions are ~3000 instructions and prox_elem is
// inlined into every unrolled chain of every fused kernel (the first version grew k_pgm_tail from 8 K to 42 K
// instructions); the rare operator pays a function call instead.
static __device__ __noinline__ double pmx_lambertw0(double z) {
  if (!(z > 0.0)) return z;                 // 0 -> 0, NaN -> NaN
  if (isinf(z)) return z;
  double w;
  if (z < 2.718281828459045) {
    w = log1p(z);
    w = w - (w * exp(w) - z) / (exp(w) * (w + 1.0));   // one Newton step: log1p overshoots for z ~ 1
  } else {
    const double l1 = log(z), l2 = log(l1);
    w = l1 - l2 + l2 / l1;
  }
  for (int it = 0; it < 8; ++it) {
    const double ew = exp(w), f = w * ew - z;
    const double wp1 = w + 1.0;
    const double dw = f / (ew * wp1 - (w + 2.0) * f / (2.0 * wp1));
    w -= dw;
    if (fabs(dw) <= 1e-16 * fabs(w)) break;
  }
  return w;
}

// prox_max_entropy of one element with x > 0 (operators.py:182-183); f64 != 0: the caller's array is float64
static __device__ __noinline__ float pmx_maxent_elem(float x, float t, int f64) {
  if (!f64) {
    // NumPy evaluates  gamma_ * real(lambertw(exp(X / gamma_ - 1) / gamma_))  with the fp32 array X: the argument of
    // lambertw is an fp32 value (np.exp of an fp32 array: overflows to inf for X / gamma_ > 89), W and the product
    // with gamma_ are fp64, the assignment rounds to fp32
    const float a = __fsub_rn(__fdiv_rn(x, t), 1.0f);
    const float e = (float)exp((double)a);
    const float z = __fdiv_rn(e, t);
    return (float)((double)t * pmx_lambertw0((double)z));
  }
  const double a = (double)x / (double)t - 1.0;
  if (a < 700.0) return (float)((double)t * pmx_lambertw0(exp(a) / (double)t));
  // exp(a) would overflow even in fp64 (the reference returns inf only beyond a = 709): solve w + log(w) = a - log(t)
  const double Lr = a - log((double)t);
  double w = Lr - log(Lr);
  for (int it = 0; it < 6; ++it) w -= (w + log(w) - Lr) / (1.0 + 1.0 / w);
  return a > 709.78 ? __int_as_float(0x7f800000) : (float)((double)t * w);
}

// One elementwise primitive.  The comparisons are written exactly like the NumPy masks of
// the reference so that NaN, +-inf and -0.0 behave identically (support sets are bit-exact).
__device__ __forceinline__ float prox_elem(float x, int op, float t) {
  switch (op) {
    case PMX_OP_ZERO: return 0.0f;                                   // operators.py:29
    case PMX_OP_PLUS: return (x < 0.0f) ? 0.0f : x;                   // operators.py:36-37
    case PMX_OP_MIN:  return (x - t < 0.0f) ? t : x;                  // operators.py:67-68
    case PMX_OP_MAX:  return (x - t > 0.0f) ? t : x;                  // operators.py:82-83
    case PMX_OP_HARD: return (fabsf(x) < t) ? 0.0f : x;               // operators.py:124-125
    case PMX_OP_SOFT: {                                               // operators.py:150
      float a = fabsf(x) - t;
      a = (a < 0.0f) ? 0.0f : a;                                      // prox_plus of |X|-t
      float s = (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : ((x == 0.0f) ? 0.0f : x));  // np.sign
      return s * a;
    }
    case PMX_OP_MAXENT:                                               // operators.py:182-183: only X[X > 0] is touched
    case PMX_OP_MAXENT64:
      return (x > 0.0f) ? pmx_maxent_elem(x, t, op == PMX_OP_MAXENT64) : x;
    default: return x;
  }
}

// apply the elementwise ops [a, b) of the chain
__device__ __forceinline__ float chain_segment(const ProxChain& c, int a, int b, float x, float step) {
  for (int i = a; i < b; ++i) {
    float t = c.rel[i] ? __fmul_rn(c.thr[i], step) : c.thr[i];  // operators.py:4-14 in fp32 (NumPy weak-scalar rule)
    x = prox_elem(x, c.op[i], t);
  }
  return x;
}

// How the per-element step / threshold scale is obtained.
struct StepSpec {
  const float* ptr;  // device pointer (modes 1..3)
  int mode;          // 0 host scalar `value`; 1 *ptr; 2 ptr[col]; 3 ptr[row]
  float scale;       // multiplies the device value (backtracking T_j, slack)
  float value;
};
__device__ __forceinline__ float step_at(const StepSpec& s, int row, int col) {
  switch (s.mode) {
    case 1: return s.ptr[0] * s.scale;
    case 2: return s.ptr[col] * s.scale;
    case 3: return s.ptr[row] * s.scale;
    default: return s.value;
  }
}

enum { IN_PLAIN = 0, IN_PGM = 1, IN_ADASUB = 2 };

struct UpdIO {
  const float* Xin;    // point the transform starts from (X, extrapolated X, or z)
  const float* G;      // IN_PGM: gradient;  IN_ADASUB: Psi
  const float* X0;     // IN_ADASUB: X after the moment step (algorithms.py:378)
  const float* Xprev;  // norms compare the result against this (may alias Xout)
  float* Xout;
  float* Xold_out;     // optional: receives a copy of Xprev (algorithms.py:102)
  double* norms;       // optional: += { |out-prev|^2, |out|^2, |prev|^2 }
  const int* done;     // optional early-exit flag(s): skip when *done != 0
  const int* done2;
  const float* psimax; // IN_ADASUB: max(Psi) (algorithms.py:384)
  unsigned short* hi;  // optional: bf16 (hi, lo) split of the result, element (r, c) at [r * ld_split + c]
  unsigned short* lo;  //   (operands of the tcgen05 gradient GEMMs; saves the separate split pass)
  int ld_split;
  float* gram_part;    // optional (cols kernel, rows <= 64): per-block partial Gram  sum_c X[i,c] X[j,c]  of the result,
                       //   [gridDim.x][rows*rows] -- the Lipschitz constant of the next iteration without another pass
  int rows, cols;
  StepSpec step;       // IN_PGM: step; IN_ADASUB: Alpha; IN_PLAIN: step handed to the prox
};

int launch_update(pmx_ctx* ctx, int in_kind, const ProxChain& chain, const UpdIO& io);
