/* CPU ORACLE (plain C) -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-loop restatements of the third-party primitives the reference's hot path reaches through torch
 * (conv2d / conv backward-data / max_pool2d / bmm / Adam), used to cross-check the torch primitives that
 * oracle/st_oracle.py builds on (tests/test_oracle_c.py).  Only tests/, __graft_entry__ and bench.py's CPU legs
 * may load this; the product never does.  Follows:
 *   conv3x3 + bias (+ReLU)        torchvision vgg.py:73-87 (zero pad), ST:39,52-59 (replicate pad on conv0)
 *   conv3x3 backward-data         autograd of the above (weights frozen, ST:49)
 *   max_pool2d 2x2 fwd/bwd        ST:21 (ATen: first maximum in window scan order wins; floor mode)
 *   gram / mean                   ST:163-168
 *   sqrtm_ns / Lyapunov backward  sqrtm.py:9-25, 36-47
 *   adam + clamp + ema            torch/optim/adam.py:413-546, ST:483-486, ST:250-253
 * Layout: NCHW float32, batch 1, accumulation in double.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* out[co][y][x] = b[co] + sum_{ci,ky,kx} in[ci][y+ky-1][x+kx-1] * w[co][ci][ky][kx]; pad: 0 = zeros, 1 = replicate */
EXPORT void stc_conv3x3(const float* in, const float* w, const float* b, float* out, int Cin, int Cout, int H, int W,
                        int replicate, int relu) {
#pragma omp parallel for collapse(2)
  for (int co = 0; co < Cout; ++co)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        double acc = b ? b[co] : 0.0;
        for (int ci = 0; ci < Cin; ++ci)
          for (int ky = 0; ky < 3; ++ky)
            for (int kx = 0; kx < 3; ++kx) {
              int yy = y + ky - 1, xx = x + kx - 1;
              if (replicate) { yy = clampi(yy, 0, H - 1); xx = clampi(xx, 0, W - 1); }
              else if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
              acc += (double)in[((size_t)ci * H + yy) * W + xx] * w[(((size_t)co * Cin + ci) * 3 + ky) * 3 + kx];
            }
        float v = (float)acc;
        out[((size_t)co * H + y) * W + x] = (relu && v < 0.f) ? 0.f : v;
      }
}

/* gin[ci][yi][xi] = sum over (co,ky,kx,yo,xo) with src(yo+ky-1, xo+kx-1) == (yi,xi) of gout[co][yo][xo]*w[co][ci][ky][kx] */
EXPORT void stc_conv3x3_dgrad(const float* gout, const float* w, float* gin, int Cin, int Cout, int H, int W,
                              int replicate) {
  memset(gin, 0, sizeof(float) * (size_t)Cin * H * W);
#pragma omp parallel for
  for (int ci = 0; ci < Cin; ++ci) {
    double* acc = (double*)calloc((size_t)H * W, sizeof(double));
    for (int co = 0; co < Cout; ++co)
      for (int yo = 0; yo < H; ++yo)
        for (int xo = 0; xo < W; ++xo) {
          const double g = gout[((size_t)co * H + yo) * W + xo];
          if (g == 0.0) continue;
          for (int ky = 0; ky < 3; ++ky)
            for (int kx = 0; kx < 3; ++kx) {
              int yy = yo + ky - 1, xx = xo + kx - 1;
              if (replicate) { yy = clampi(yy, 0, H - 1); xx = clampi(xx, 0, W - 1); }
              else if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
              acc[(size_t)yy * W + xx] += g * w[(((size_t)co * Cin + ci) * 3 + ky) * 3 + kx];
            }
        }
    for (size_t i = 0; i < (size_t)H * W; ++i) gin[(size_t)ci * H * W + i] = (float)acc[i];
    free(acc);
  }
}

EXPORT void stc_maxpool2(const float* in, float* out, int C, int H, int W) {
  const int Ho = H / 2, Wo = W / 2;
  for (int c = 0; c < C; ++c)
    for (int y = 0; y < Ho; ++y)
      for (int x = 0; x < Wo; ++x) {
        float m = -INFINITY;
        for (int i = 0; i < 2; ++i)
          for (int j = 0; j < 2; ++j) {
            float v = in[((size_t)c * H + 2 * y + i) * W + 2 * x + j];
            if (v > m) m = v;
          }
        out[((size_t)c * Ho + y) * Wo + x] = m;
      }
}

EXPORT void stc_maxpool2_bwd(const float* gout, const float* in, float* gin, int C, int H, int W) {
  const int Ho = H / 2, Wo = W / 2;
  memset(gin, 0, sizeof(float) * (size_t)C * H * W);
  for (int c = 0; c < C; ++c)
    for (int y = 0; y < Ho; ++y)
      for (int x = 0; x < Wo; ++x) {
        int bi = 0, bj = 0;
        float m = -INFINITY;
        for (int i = 0; i < 2; ++i)
          for (int j = 0; j < 2; ++j) {
            float v = in[((size_t)c * H + 2 * y + i) * W + 2 * x + j];
            if (v > m) { m = v; bi = i; bj = j; } /* strict > : first maximum wins */
          }
        gin[((size_t)c * H + 2 * y + bi) * W + 2 * x + bj] = gout[((size_t)c * Ho + y) * Wo + x];
      }
}

/* mean[c] and srm[c][d] = sum_p f[c][p] f[d][p] / N */
EXPORT void stc_style_stats(const float* f, float* mean, float* srm, int C, int N) {
#pragma omp parallel for
  for (int c = 0; c < C; ++c) {
    double s = 0;
    for (int p = 0; p < N; ++p) s += f[(size_t)c * N + p];
    mean[c] = (float)(s / N);
    for (int d = 0; d < C; ++d) {
      double a = 0;
      for (int p = 0; p < N; ++p) a += (double)f[(size_t)c * N + p] * f[(size_t)d * N + p];
      srm[(size_t)c * C + d] = (float)(a / N);
    }
  }
}

static void matmul(const float* a, const float* b, float* c, int n) { /* c = a b, fp32 storage, double accumulate */
#pragma omp parallel for
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0;
      for (int k = 0; k < n; ++k) s += (double)a[(size_t)i * n + k] * b[(size_t)k * n + j];
      c[(size_t)i * n + j] = (float)s;
    }
}

EXPORT void stc_sqrtm_ns(const float* a, float* out, int n, int iters) {
  size_t nn = (size_t)n * n;
  float *y = malloc(nn * 4), *z = malloc(nn * 4), *t = malloc(nn * 4), *tmp = malloc(nn * 4);
  double ss = 0;
  for (size_t i = 0; i < nn; ++i) ss += (double)a[i] * a[i];
  const float norm = (float)sqrt(ss);
  for (size_t i = 0; i < nn; ++i) { y[i] = a[i] / norm; z[i] = 0.f; }
  for (int i = 0; i < n; ++i) z[(size_t)i * n + i] = 1.f;
  for (int it = 0; it < iters; ++it) {
    matmul(z, y, tmp, n);
    for (size_t i = 0; i < nn; ++i) t[i] = -tmp[i] / 2;
    for (int i = 0; i < n; ++i) t[(size_t)i * n + i] += 1.5f;
    matmul(y, t, tmp, n); memcpy(y, tmp, nn * 4);
    matmul(t, z, tmp, n); memcpy(z, tmp, nn * 4);
  }
  const float s = sqrtf(norm);
  for (size_t i = 0; i < nn; ++i) out[i] = y[i] * s;
  free(y); free(z); free(t); free(tmp);
}

static void transpose(const float* a, float* t, int n) {
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) t[(size_t)j * n + i] = a[(size_t)i * n + j];
}

EXPORT void stc_sqrtm_lyap_bwd(const float* z, const float* gout, float* gin, int n, int iters) {
  size_t nn = (size_t)n * n;
  float *a = malloc(nn * 4), *at = malloc(nn * 4), *q = malloc(nn * 4), *e = malloc(nn * 4), *t1 = malloc(nn * 4),
        *t2 = malloc(nn * 4), *t3 = malloc(nn * 4);
  double ss = 0;
  for (size_t i = 0; i < nn; ++i) ss += (double)z[i] * z[i];
  const float norm = (float)sqrt(ss);
  for (size_t i = 0; i < nn; ++i) { a[i] = z[i] / norm; q[i] = gout[i] / norm; }
  for (int it = 0; it < iters; ++it) {
    matmul(a, a, e, n);
    for (size_t i = 0; i < nn; ++i) e[i] = -e[i];
    for (int i = 0; i < n; ++i) e[(size_t)i * n + i] += 3.f;
    transpose(a, at, n);
    matmul(q, e, t1, n);          /* q E */
    matmul(at, q, t2, n);         /* a^T q */
    matmul(q, a, t3, n);          /* q a */
    for (size_t i = 0; i < nn; ++i) t2[i] -= t3[i];
    matmul(at, t2, t3, n);        /* a^T (a^T q - q a) */
    for (size_t i = 0; i < nn; ++i) q[i] = (t1[i] - t3[i]) / 2;
    if (it < iters - 1) { matmul(a, e, t1, n); for (size_t i = 0; i < nn; ++i) a[i] = t1[i] / 2; }
  }
  for (size_t i = 0; i < nn; ++i) gin[i] = q[i] / 2;
  free(a); free(at); free(q); free(e); free(t1); free(t2); free(t3);
}

EXPORT void stc_adam_clamp_ema(float* x, const float* g, float* m, float* v, float* ema, long n, int step, float lr,
                               float b1, float b2, float eps, float decay) {
  const double bc1 = 1.0 - pow(b1, step), bc2 = 1.0 - pow(b2, step);
  const float step_size = (float)(lr / bc1), bc2s = (float)sqrt(bc2);
  for (long i = 0; i < n; ++i) {
    m[i] = m[i] + (g[i] - m[i]) * (1.f - b1);
    v[i] = v[i] * b2 + (1.f - b2) * g[i] * g[i];
    const float denom = sqrtf(v[i]) / bc2s + eps;
    float p = x[i] - step_size * (m[i] / denom);
    p = p < 0.f ? 0.f : (p > 1.f ? 1.f : p);
    x[i] = p;
    ema[i] = ema[i] * decay + (1.f - decay) * p;
  }
}
