// Wasserstein-2 style loss on the five tap covariances, forward and backward, in true fp32:
//   StyleLossW2.__init__/forward      /root/reference/style_transfer/style_transfer.py:149-181  (ST)
//   sqrtm_ns (12 Newton-Schulz its)   /root/reference/style_transfer/sqrtm.py:9-25              (SQ)
//   _MatrixSquareRootNSLyap.backward  SQ:36-47 (iterative Lyapunov solve, 12 its)
// plus the closed-form backward of ST:163-181 down to  G = d loss / d srm  and  d loss / d mean.
//
// The loss is a cancellation (tr(St + S - 2 sqrt(.))): TF32/bf16 operands are not accurate
// enough, so these C x C chains are evaluated with fp32 operands and an error-compensated 3xTF32 split on the tensor
// cores (fp32-level accuracy).  All five layers advance in lock-step: one "round" = one grouped launch whose CTAs are
// 64x64 tiles of every layer's GEMM.
#include <vector>

#include "kernels.h"
#include "ptx.cuh"

namespace stb {

namespace {

constexpr int TS = 64;     // output tile
constexpr int KC = 8;      // k chunk per smem stage
constexpr int KG = 8;      // in-CTA split-K groups (64 threads each) -> 512 threads
constexpr int NRED = 64;   // max reduction partials per layer (tiles of a 512x512 problem / helper CTAs)

// D = alpha*op(A)*op(B) + alpha2*op(A2)*op(B2) + beta*Cadd + gamma*I on 64x64 tiles, fp32 in / fp32 out.
//
// The products run on the tensor cores with an error-compensated 3xTF32 split done in registers:
//     x = hi + lo,  hi = x with the low 13 mantissa bits cleared (exactly representable in TF32),  lo = x - hi,
//     a*b ~= lo_a*hi_b + hi_a*lo_b + hi_a*hi_b          (the dropped lo*lo term is 2^-22 relative),
// accumulated in fp32 (mma.sync.m16n8k8.tf32).  Plain TF32 is NOT accurate enough for this chain;
// the split restores fp32-level accuracy (tests/test_gpu_kernels.py::test_w2_*).
//
// One CTA = 8 k-groups x 2 warps: a group owns 1/8 of K with private smem double buffers; warp w of a group owns
// rows [32w, 32w+32) of the tile as 2 x 8 m16n8 fragments.  The groups' partial tiles are tree-reduced through
// smem, group 0 applies the epilogue.  Symmetric results are computed for tiles ti <= tj only and mirrored.
// Optionally writes the tile's {sum of squares, trace} partial for the Newton-Schulz norms (deterministic order).
constexpr int LDT = TS + 8;  // smem row stride (floats): fragment loads hit 32 distinct banks

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

__global__ void __launch_bounds__(KG * 64, 1)
sgemm_grouped_kernel(const GemmProb* __restrict__ probs, const uint32_t* __restrict__ tiles) {
  extern __shared__ __align__(16) float smem_f[];
  const uint32_t t = tiles[blockIdx.x];
  const GemmProb pr = probs[t >> 16];
  const int ti = (t >> 8) & 0xFF, tj = t & 0xFF;
  const int n = pr.n;
  const int grp = threadIdx.x >> 6, lt = threadIdx.x & 63;
  const int w2 = lt >> 5, lane = lt & 31, g = lane >> 2, tq = lane & 3;
  float* As = smem_f + grp * (4 * KC * LDT);  // [2][KC][LDT]
  float* Bs = As + 2 * KC * LDT;              // [2][KC][LDT]
  // acc[mi][ni][r]: rows 32*w2 + 16*mi + g (+8 for r >= 2), cols 8*ni + 2*tq + (r & 1)
  float acc[2][8][4];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 8; ++ni)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[mi][ni][r] = 0.f;

  const int kper = n / KG;           // k range of this group (n is a multiple of 64)
  const int kbeg = grp * kper;
  const int n_prod = (pr.A2 != nullptr) ? 2 : 1;
  for (int pi = 0; pi < n_prod; ++pi) {
    const float* __restrict__ A = pi ? pr.A2 : pr.A;
    const float* __restrict__ B = pi ? pr.B2 : pr.B;
    const int tA = pi ? pr.transA2 : pr.transA, tB = pi ? pr.transB2 : pr.transB;
    const float alpha = pi ? pr.alpha2 : pr.alpha;
    float4 ra[2], rb[2];
    auto gload = [&](int k0) {
      if (!tA) {  // A[i][k]: thread = row i, 8 consecutive k
        const float4* p = reinterpret_cast<const float4*>(A + (size_t)(ti * TS + lt) * n + k0);
        ra[0] = p[0]; ra[1] = p[1];
      } else {    // A^T(i,k) = A[k][i]: thread = (k = lt>>3, 8 consecutive i)
        const float4* p = reinterpret_cast<const float4*>(A + (size_t)(k0 + (lt >> 3)) * n + ti * TS + (lt & 7) * 8);
        ra[0] = p[0]; ra[1] = p[1];
      }
      if (!tB) {  // B[k][j]
        const float4* p = reinterpret_cast<const float4*>(B + (size_t)(k0 + (lt >> 3)) * n + tj * TS + (lt & 7) * 8);
        rb[0] = p[0]; rb[1] = p[1];
      } else {    // B^T(k,j) = B[j][k]
        const float4* p = reinterpret_cast<const float4*>(B + (size_t)(tj * TS + lt) * n + k0);
        rb[0] = p[0]; rb[1] = p[1];
      }
    };
    auto sstore = [&](int buf) {
      float* a = As + buf * KC * LDT;
      float* b = Bs + buf * KC * LDT;
      const float av[8] = {ra[0].x * alpha, ra[0].y * alpha, ra[0].z * alpha, ra[0].w * alpha,
                           ra[1].x * alpha, ra[1].y * alpha, ra[1].z * alpha, ra[1].w * alpha};
      const float bv[8] = {rb[0].x, rb[0].y, rb[0].z, rb[0].w, rb[1].x, rb[1].y, rb[1].z, rb[1].w};
      if (!tA) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k * LDT + lt] = av[k];
      } else {
        float4* d = reinterpret_cast<float4*>(a + (lt >> 3) * LDT + (lt & 7) * 8);
        d[0] = make_float4(av[0], av[1], av[2], av[3]);
        d[1] = make_float4(av[4], av[5], av[6], av[7]);
      }
      if (!tB) {
        float4* d = reinterpret_cast<float4*>(b + (lt >> 3) * LDT + (lt & 7) * 8);
        d[0] = make_float4(bv[0], bv[1], bv[2], bv[3]);
        d[1] = make_float4(bv[4], bv[5], bv[6], bv[7]);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) b[k * LDT + lt] = bv[k];
      }
    };
    gload(kbeg);
    sstore(0);
    named_bar_sync(1 + grp, 64);
    int buf = 0;
    for (int k0 = 0; k0 < kper; k0 += KC) {
      const bool more = (k0 + KC) < kper;
      if (more) gload(kbeg + k0 + KC);
      const float* a = As + buf * KC * LDT;
      const float* b = Bs + buf * KC * LDT;
      // one m16n8k8 k-step per chunk (KC == 8)
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int rb_ = w2 * 32 + mi * 16 + g;
        split_tf32(a[tq * LDT + rb_], ah[mi][0], al[mi][0]);
        split_tf32(a[tq * LDT + rb_ + 8], ah[mi][1], al[mi][1]);
        split_tf32(a[(tq + 4) * LDT + rb_], ah[mi][2], al[mi][2]);
        split_tf32(a[(tq + 4) * LDT + rb_ + 8], ah[mi][3], al[mi][3]);
      }
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) {
        uint32_t bh[2], bl[2];
        split_tf32(b[tq * LDT + ni * 8 + g], bh[0], bl[0]);
        split_tf32(b[(tq + 4) * LDT + ni * 8 + g], bh[1], bl[1]);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
          mma_tf32(acc[mi][ni], al[mi], bh);
          mma_tf32(acc[mi][ni], ah[mi], bl);
          mma_tf32(acc[mi][ni], ah[mi], bh);
        }
      }
      if (more) sstore(buf ^ 1);
      named_bar_sync(1 + grp, 64);
      buf ^= 1;
    }
  }
  // ---- tree reduction of the 8 partial tiles through smem (reuses the staging buffers)
  __syncthreads();
  float* red = smem_f;  // [4][64 threads][64] floats
  float* accf = &acc[0][0][0];
  for (int half = KG / 2; half >= 1; half >>= 1) {
    if (grp >= half && grp < 2 * half) {
      float* dst = red + ((grp - half) * 64 + lt) * 64;
#pragma unroll
      for (int e = 0; e < 64; e += 4)
        *reinterpret_cast<float4*>(dst + (e ^ ((lt & 7) * 4))) = make_float4(accf[e], accf[e + 1], accf[e + 2], accf[e + 3]);
    }
    __syncthreads();
    if (grp < half) {
      const float* src = red + (grp * 64 + lt) * 64;
#pragma unroll
      for (int e = 0; e < 64; e += 4) {
        const float4 v = *reinterpret_cast<const float4*>(src + (e ^ ((lt & 7) * 4)));
        accf[e] += v.x; accf[e + 1] += v.y; accf[e + 2] += v.z; accf[e + 3] += v.w;
      }
    }
    __syncthreads();
  }
  if (grp == 0) {
    float ssq = 0.f, tr = 0.f;
    const bool mirror = pr.sym && ti != tj;
    float* tp = smem_f + 4096;  // [64][65] transpose buffer inside the idle staging area
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 8; ++ni)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int lr = w2 * 32 + mi * 16 + g + h * 8, lc = ni * 8 + 2 * tq;  // local row / col of (c[2h], c[2h+1])
          const int gi = ti * TS + lr, gj = tj * TS + lc;
          float2 o = make_float2(acc[mi][ni][2 * h], acc[mi][ni][2 * h + 1]);
          if (pr.Cadd != nullptr) {
            const float2 c = *reinterpret_cast<const float2*>(pr.Cadd + (size_t)gi * n + gj);
            o.x = fmaf(pr.beta, c.x, o.x);
            o.y = fmaf(pr.beta, c.y, o.y);
          }
          if (gi == gj) { o.x += pr.gamma; tr += o.x; }
          if (gi == gj + 1) { o.y += pr.gamma; tr += o.y; }
          ssq = fmaf(o.x, o.x, ssq);
          ssq = fmaf(o.y, o.y, ssq);
          *reinterpret_cast<float2*>(pr.D + (size_t)gi * n + gj) = o;
          if (mirror) { tp[lc * 65 + lr] = o.x; tp[(lc + 1) * 65 + lr] = o.y; }
        }
    if (mirror) {
      // symmetric result (product of commuting symmetric matrices): the mirror tile D[tj][ti] = tile^T is written
      // through the padded smem transpose (thread lt writes row lt, coalesced float4)
      ssq *= 2.f;
      named_bar_sync(1, 64);
      float* drow = pr.D + (size_t)(tj * TS + lt) * n + ti * TS;
#pragma unroll
      for (int c4 = 0; c4 < 64; c4 += 4)
        *reinterpret_cast<float4*>(drow + c4) =
            make_float4(tp[lt * 65 + c4], tp[lt * 65 + c4 + 1], tp[lt * 65 + c4 + 2], tp[lt * 65 + c4 + 3]);
    }
    if (pr.red_out != nullptr) {  // warp-uniform: group 0 = warps 0, 1
      ssq = warp_sum(ssq);
      tr = warp_sum(tr);
      float* sh = smem_f + 12288;  // scratch inside the idle staging area (disjoint from the transpose buffer)
      if ((lt & 31) == 0) { sh[(lt >> 5) * 2] = ssq; sh[(lt >> 5) * 2 + 1] = tr; }
      named_bar_sync(1, 64);
      if (lt == 0) {
        const int nt = n / TS;
        pr.red_out[(ti * nt + tj) * 2] = sh[0] + sh[2];
        pr.red_out[(ti * nt + tj) * 2 + 1] = sh[1] + sh[3];
        if (mirror) {  // consumers sum all (n/64)^2 slots: the mirror tile's share is already doubled above
          pr.red_out[(tj * nt + ti) * 2] = 0.f;
          pr.red_out[(tj * nt + ti) * 2 + 1] = 0.f;
        }
      }
    }
  }
}
constexpr int SGEMM_SMEM = KG * 4 * KC * LDT * 4;  // 72 KiB (>= the 64 KiB the tree reduction needs)

// ---- helpers: grid (5 layers, NB CTAs), 256 threads; reductions via fixed-order partials (deterministic)
constexpr int NB = 32;
__device__ __forceinline__ float block_sum_256(float v, float* s_red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += s_red[i];
  return t;
}
__device__ __forceinline__ void sum_partials(const float* red, int count, float& a, float& b) {
  a = 0.f; b = 0.f;
  for (int i = 0; i < count; ++i) { a += red[2 * i]; b += red[2 * i + 1]; }
}

// covariance from (reduced) raw sums:  mu = sums/N; cov = S_raw/N - mu mu^T + eps I     (ST:171-173, 177)
// target mode: cov_t from (mean_t, srm_t), plus sum-of-squares partials of cov_t for the NS normalisation.
__global__ void __launch_bounds__(256) w2_cov_kernel(const W2Layer* __restrict__ layers, int from_target) {
  __shared__ float s_red[8];
  const W2Layer L = layers[blockIdx.x];
  const int n = L.n;
  const float inv_n = from_target ? 1.f : 1.f / L.npix;
  const float* S = from_target ? L.srm_t : L.S_raw;
  const float* sm = from_target ? L.mean_t : L.sums;
  float* cov = from_target ? L.cov_t : L.cov;
  float ssq = 0.f;
  for (int e = blockIdx.y * 256 + threadIdx.x; e < n * n; e += NB * 256) {
    const int i = e / n, j = e - i * n;
    float v = S[e] * inv_n - (sm[i] * inv_n) * (sm[j] * inv_n);
    if (i == j) v += L.eps;
    cov[e] = v;
    ssq = fmaf(v, v, ssq);
  }
  ssq = block_sum_256(ssq, s_red);
  if (threadIdx.x == 0) { L.red[blockIdx.y * 2] = ssq; L.red[blockIdx.y * 2 + 1] = 0.f; }
  if (blockIdx.y == 0) {
    float tr = 0.f, md = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) {
      const float m = sm[i] * inv_n;
      tr += S[(size_t)i * n + i] * inv_n - m * m + L.eps;
      if (!from_target) {
        L.mu[i] = m;
        const float d = m - L.mean_t[i];
        md += d * d;
      }
    }
    tr = block_sum_256(tr, s_red);
    md = block_sum_256(md, s_red);
    if (threadIdx.x == 0) {
      if (from_target) L.scal[W2S_TR_COV_T] = tr;
      else { L.scal[W2S_TR_COV] = tr; L.scal[W2S_MEAN_DIFF] = md / n; }
    }
  }
}

// Y = M / ||M||_F, Z = I        (SQ:15-19).  ||M||^2 arrives as partials (cov kernel: NB; GEMM tiles: (n/64)^2)
__global__ void __launch_bounds__(256) w2_ns_init_kernel(const W2Layer* __restrict__ layers, int from_target) {
  const W2Layer L = layers[blockIdx.x];
  const int n = L.n;
  const float* M = from_target ? L.cov_t : L.M;
  float ss, dummy;
  sum_partials(L.red, from_target ? NB : (n / TS) * (n / TS), ss, dummy);
  const float norm = sqrtf(ss);
  for (int e = blockIdx.y * 256 + threadIdx.x; e < n * n; e += NB * 256) {
    const int i = e / n, j = e - i * n;
    L.Y[0][e] = M[e] / norm;
    L.Z[0][e] = (i == j) ? 1.f : 0.f;
  }
  if (blockIdx.y == 0 && threadIdx.x == 0) L.scal[W2S_NORM_A] = norm;
}

// target: P = Y sqrt(norm)                                              (ST:159, SQ:25)
__global__ void __launch_bounds__(256) w2_target_finish_kernel(const W2Layer* __restrict__ layers) {
  const W2Layer L = layers[blockIdx.x];
  const int n = L.n;
  const float s = sqrtf(L.scal[W2S_NORM_A]);
  for (int e = blockIdx.y * 256 + threadIdx.x; e < n * n; e += NB * 256) L.P[e] = L.Y[0][e] * s;
}

// forward finish: R = Y sqrt(normA); loss; seeds of the Lyapunov backward    (SQ:25, ST:178-181, SQ:37-41).
// ||Y||^2 and tr(Y) arrive as per-tile partials written by the last NS round.
__global__ void __launch_bounds__(256) w2_fwd_finish_kernel(const W2Layer* __restrict__ layers, float* loss_terms) {
  const W2Layer L = layers[blockIdx.x];
  const int n = L.n;
  const float* Y = L.Y[0];
  float ss, tr;
  sum_partials(L.red, (n / TS) * (n / TS), ss, tr);
  const float sq = sqrtf(L.scal[W2S_NORM_A]);
  const float norm_y = sqrtf(ss);
  const float norm_r = sq * norm_y;                     // ||R||_F
  const float tr_r = tr * sq;
  const float seed = -2.f * L.weight / (n * norm_r);    // grad_output / ||z|| with grad_output = -2 w / C * I
  for (int e = blockIdx.y * 256 + threadIdx.x; e < n * n; e += NB * 256) {
    const int i = e / n, j = e - i * n;
    L.A[0][e] = Y[e] / norm_y;                          // a = z / ||z||
    L.Q[0][e] = (i == j) ? seed : 0.f;
  }
  if (blockIdx.y == 0 && threadIdx.x == 0) {
    const float cov_diff = (L.scal[W2S_TR_COV_T] + L.scal[W2S_TR_COV] - 2.f * tr_r) / n;
    const float l = (L.scal[W2S_MEAN_DIFF] + cov_diff) * L.weight;
    L.scal[W2S_LOSS] = l;
    loss_terms[blockIdx.x] = l;
  }
}

// backward finish 1: Gs = Gc + Gc^T; bf16 Gs/N ([C][C]) for the tap-gradient GEMM
__global__ void __launch_bounds__(256) w2_bwd_finish_kernel(const W2Layer* __restrict__ layers) {
  const W2Layer L = layers[blockIdx.x];
  const int n = L.n;
  const float inv_n = 1.f / L.npix;
  for (int e = blockIdx.y * 256 + threadIdx.x; e < n * n; e += NB * 256) {
    const int i = e / n, j = e - i * n;
    const float gs = L.Gc[e] + L.Gc[(size_t)j * n + i];
    L.Gs[e] = gs;
    L.gs_bf16[e] = __float2bfloat16(gs * inv_n);
  }
}
// backward finish 2: gmu = 2w(mu - mu_t)/C - Gs mu  (one warp per row), emitted divided by N
__global__ void __launch_bounds__(256) w2_gmu_kernel(const W2Layer* __restrict__ layers) {
  const W2Layer L = layers[blockIdx.x];
  const int n = L.n;
  const int lane = threadIdx.x & 31;
  const float inv_n = 1.f / L.npix;
  for (int i = blockIdx.y * 8 + (threadIdx.x >> 5); i < n; i += NB * 8) {
    float s = 0.f;
    for (int j = lane; j < n; j += 32) s = fmaf(L.Gs[(size_t)i * n + j], L.mu[j], s);
    s = warp_sum(s);
    if (lane == 0) L.gmu_bias[i] = (2.f * L.weight * (L.mu[i] - L.mean_t[i]) / n - s) * inv_n;
  }
}

}  // namespace

// ================================================================================================ host engine
size_t W2Engine::layer_floats(int n) {
  // cov, M, X, Y[2], Z[2], T, A[2], Q[2], E, X1, X23, U, Gc, Gs, P, cov_t, srm_t  (21 matrices) + vectors
  return (size_t)21 * n * n + 8 * (size_t)n + 64 + 2 * NRED + 256;
}

size_t W2Engine::workspace_bytes() {
  size_t fl = 0;
  const int ns[5] = {64, 128, 256, 512, 512};
  for (int l = 0; l < 5; ++l) fl += layer_floats(ns[l]) + (size_t)ns[l] * ns[l] / 2 + 64;  // + bf16 Gs
  // device copies of layer table, problems and tiles
  return fl * 4 + (size_t(1) << 20);
}

static void add_prob(std::vector<GemmProb>& probs, std::vector<uint32_t>& tiles, const GemmProb& p) {
  const int idx = (int)probs.size();
  probs.push_back(p);
  const int nt = p.n / TS;
  for (int i = 0; i < nt; ++i)
    for (int j = (p.sym ? i : 0); j < nt; ++j) tiles.push_back((uint32_t)idx << 16 | (uint32_t)i << 8 | (uint32_t)j);
}

static GemmProb mk(int n, float* D, const float* A, int tA, const float* B, int tB, float alpha, float gamma = 0.f,
                   const float* Cadd = nullptr, float beta = 0.f, const float* A2 = nullptr, int tA2 = 0,
                   const float* B2 = nullptr, int tB2 = 0, float alpha2 = 0.f, float* red_out = nullptr, int sym = 1) {
  GemmProb p{};
  p.red_out = red_out;
  p.sym = sym;
  p.A = A; p.B = B; p.A2 = A2; p.B2 = B2; p.Cadd = Cadd; p.D = D; p.n = n;
  p.transA = tA; p.transB = tB; p.transA2 = tA2; p.transB2 = tB2;
  p.alpha = alpha; p.alpha2 = alpha2; p.beta = beta; p.gamma = gamma;
  return p;
}

int W2Engine::init(void* ws, size_t bytes, const int n_per_layer[5]) {
  STB_CHECK(bytes >= workspace_bytes(), STB_ERR_WORKSPACE, "W2 workspace too small");
  uint8_t* base = static_cast<uint8_t*>(ws);
  size_t off = 0;
  auto take = [&](size_t nbytes) { void* p = base + off; off += (nbytes + 255) & ~size_t(255); return p; };
  for (int l = 0; l < 5; ++l) {
    W2Layer& L = host_layers[l];
    const int n = n_per_layer[l];
    const size_t nn = (size_t)n * n * 4;
    L.n = n;
    L.eps = 1e-4f;
    L.cov = (float*)take(nn); L.M = (float*)take(nn); L.X = (float*)take(nn);
    L.Y[0] = (float*)take(nn); L.Y[1] = (float*)take(nn); L.Z[0] = (float*)take(nn); L.Z[1] = (float*)take(nn);
    L.T = (float*)take(nn); L.A[0] = (float*)take(nn); L.A[1] = (float*)take(nn);
    L.Q[0] = (float*)take(nn); L.Q[1] = (float*)take(nn); L.E = (float*)take(nn); L.X1 = (float*)take(nn);
    L.X23 = (float*)take(nn); L.U = (float*)take(nn); L.Gc = (float*)take(nn); L.Gs = (float*)take(nn);
    L.P = (float*)take(nn); L.cov_t = (float*)take(nn); L.srm_t = (float*)take(nn);
    L.S_raw = nullptr; L.sums = nullptr;  // bound per plan (stats buffer)
    L.mu = (float*)take(n * 4); L.mean_t = (float*)take(n * 4); L.gmu_bias = (float*)take(n * 4);
    L.scal = (float*)take(64 * 4);
    L.red = (float*)take(2 * NRED * 4);
    L.gs_bf16 = (bf16*)take((size_t)n * n * 2);
    L.weight = 0.f; L.npix = 1.f;
  }
  d_layers = (W2Layer*)take(sizeof(W2Layer) * 5);

  // ---- build the round lists once (pointers are stable)
  std::vector<GemmProb> probs;
  std::vector<uint32_t> tiles;
  rounds.clear();
  auto begin_round = [&]() { rounds.push_back({(int)tiles.size(), 0}); };
  auto end_round = [&]() { rounds.back().n_tiles = (int)tiles.size() - rounds.back().first_tile; };
  auto ns_rounds = [&]() {  // 12 x { T = 1.5 I - 0.5 Z Y ; Y' = Y T, Z' = T Z }, result ends in Y[0], Z[0]
    for (int it = 0; it < 12; ++it) {
      const int s = it & 1, d = s ^ 1;
      begin_round();
      for (int l = 0; l < 5; ++l) {
        W2Layer& L = host_layers[l];
        add_prob(probs, tiles, mk(L.n, L.T, L.Z[s], 0, L.Y[s], 0, -0.5f, 1.5f));
      }
      end_round();
      begin_round();
      for (int l = 0; l < 5; ++l) {
        W2Layer& L = host_layers[l];
        add_prob(probs, tiles, mk(L.n, L.Y[d], L.Y[s], 0, L.T, 0, 1.f, 0.f, nullptr, 0.f, nullptr, 0, nullptr, 0, 0.f,
                                  it == 11 ? L.red : nullptr));
        if (it < 11) add_prob(probs, tiles, mk(L.n, L.Z[d], L.T, 0, L.Z[s], 0, 1.f));  // Z is dead after the last Y
      }
      end_round();
    }
  };
  // (a) target chain: NS on cov_t
  r_target_begin = (int)rounds.size();
  ns_rounds();
  r_target_end = (int)rounds.size();
  // (b) iterate forward: X = P cov; M = X P; NS
  r_fwd_begin = (int)rounds.size();
  begin_round();
  for (int l = 0; l < 5; ++l) { W2Layer& L = host_layers[l]; add_prob(probs, tiles, mk(L.n, L.X, L.P, 0, L.cov, 0, 1.f, 0.f, nullptr, 0.f, nullptr, 0, nullptr, 0, 0.f, nullptr, 0)); }
  end_round();
  begin_round();
  for (int l = 0; l < 5; ++l) {
    W2Layer& L = host_layers[l];
    add_prob(probs, tiles, mk(L.n, L.M, L.X, 0, L.P, 0, 1.f, 0.f, nullptr, 0.f, nullptr, 0, nullptr, 0, 0.f, L.red));
  }
  end_round();
  r_fwd_ns_begin = (int)rounds.size();
  ns_rounds();
  r_fwd_end = (int)rounds.size();
  // (c) backward, SQ:42-46:  E = 3I - a a;  q' = (q E - a^T (a^T q - q a)) / 2;  a' = a E / 2.
  // On this path grad_output is always a multiple of I (d/dR of -2 w tr(R)/C), `a` is symmetric and every q is a
  // polynomial in `a`, so the commutator a^T q - q a is identically zero in exact arithmetic (the reference
  // evaluates its rounding noise, ~1e-7 relative); the schedule below drops it:  q' = q E / 2.
  r_bwd_begin = (int)rounds.size();
  for (int it = 0; it < 12; ++it) {
    const int s = it & 1, d = s ^ 1;
    begin_round();
    for (int l = 0; l < 5; ++l) { W2Layer& L = host_layers[l]; add_prob(probs, tiles, mk(L.n, L.E, L.A[s], 0, L.A[s], 0, -1.f, 3.f)); }
    end_round();
    begin_round();
    for (int l = 0; l < 5; ++l) {
      W2Layer& L = host_layers[l];
      add_prob(probs, tiles, mk(L.n, L.Q[d], L.Q[s], 0, L.E, 0, 0.5f));
      if (it < 11) add_prob(probs, tiles, mk(L.n, L.A[d], L.A[s], 0, L.E, 0, 0.5f));
    }
    end_round();
  }
  // after 12 its q is in Q[0].  U = P^T q ; Gc = 0.5 U P^T + (w/C) I  (gamma patched per layer in set_weights)
  begin_round();
  for (int l = 0; l < 5; ++l) { W2Layer& L = host_layers[l]; add_prob(probs, tiles, mk(L.n, L.U, L.P, 1, L.Q[0], 0, 1.f, 0.f, nullptr, 0.f, nullptr, 0, nullptr, 0, 0.f, nullptr, 0)); }
  end_round();
  begin_round();
  gc_prob_first = (int)probs.size();
  for (int l = 0; l < 5; ++l) { W2Layer& L = host_layers[l]; add_prob(probs, tiles, mk(L.n, L.Gc, L.U, 0, L.P, 1, 0.5f, 0.f)); }
  end_round();
  r_bwd_end = (int)rounds.size();

  host_probs = probs;
  d_probs = (GemmProb*)take(sizeof(GemmProb) * probs.size());
  d_tiles = (uint32_t*)take(sizeof(uint32_t) * tiles.size());
  STB_CHECK(off <= bytes, STB_ERR_WORKSPACE, "W2 workspace overflow (%zu > %zu)", off, bytes);
  STB_CUDA_CHECK(cudaMemcpy(d_tiles, tiles.data(), sizeof(uint32_t) * tiles.size(), cudaMemcpyHostToDevice));
  STB_CUDA_CHECK(cudaMemcpy(d_probs, probs.data(), sizeof(GemmProb) * probs.size(), cudaMemcpyHostToDevice));
  STB_CUDA_CHECK(cudaMemcpy(d_layers, host_layers, sizeof(W2Layer) * 5, cudaMemcpyHostToDevice));
  return STB_OK;
}

int W2Engine::upload_layers(cudaStream_t s) {
  // layer weights enter the Gc round through gamma = w / C
  for (int l = 0; l < 5; ++l) host_probs[gc_prob_first + l].gamma = host_layers[l].weight / host_layers[l].n;
  STB_CUDA_CHECK(cudaMemcpyAsync(d_probs + gc_prob_first, host_probs.data() + gc_prob_first, sizeof(GemmProb) * 5,
                                 cudaMemcpyHostToDevice, s));
  STB_CUDA_CHECK(cudaMemcpyAsync(d_layers, host_layers, sizeof(W2Layer) * 5, cudaMemcpyHostToDevice, s));
  return STB_OK;
}

int W2Engine::run_rounds(int r0, int r1, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    STB_CUDA_CHECK(cudaFuncSetAttribute(sgemm_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SGEMM_SMEM));
    attr_set = true;
  }
  for (int r = r0; r < r1; ++r) {
    sgemm_grouped_kernel<<<rounds[r].n_tiles, KG * 64, SGEMM_SMEM, s>>>(d_probs, d_tiles + rounds[r].first_tile);
  }
  STB_CUDA_CHECK(cudaGetLastError());
  return STB_OK;
}

int W2Engine::build_targets(cudaStream_t s) {
  // srm_t / mean_t already hold the blended target moments (ST:443-450)
  const dim3 grid(5, NB);
  w2_cov_kernel<<<grid, 256, 0, s>>>(d_layers, 1);
  w2_ns_init_kernel<<<grid, 256, 0, s>>>(d_layers, 1);
  STB_TRY(run_rounds(r_target_begin, r_target_end, s));
  w2_target_finish_kernel<<<grid, 256, 0, s>>>(d_layers);
  STB_CUDA_CHECK(cudaGetLastError());
  return STB_OK;
}

int W2Engine::forward_backward(float* loss_terms, cudaStream_t s) {
  const dim3 grid(5, NB);
  w2_cov_kernel<<<grid, 256, 0, s>>>(d_layers, 0);
  STB_TRY(run_rounds(r_fwd_begin, r_fwd_ns_begin, s));
  w2_ns_init_kernel<<<grid, 256, 0, s>>>(d_layers, 0);
  STB_TRY(run_rounds(r_fwd_ns_begin, r_fwd_end, s));
  w2_fwd_finish_kernel<<<grid, 256, 0, s>>>(d_layers, loss_terms);
  STB_TRY(run_rounds(r_bwd_begin, r_bwd_end, s));
  w2_bwd_finish_kernel<<<grid, 256, 0, s>>>(d_layers);
  w2_gmu_kernel<<<grid, 256, 0, s>>>(d_layers);
  STB_CUDA_CHECK(cudaGetLastError());
  return STB_OK;
}

}  // namespace stb
