// Wasserstein-2 style loss on the five tap covariances, forward and backward, in true fp32:
//   StyleLossW2.__init__/forward      /root/reference/style_transfer/style_transfer.py:149-181  (ST)
//   sqrtm_ns (12 Newton-Schulz its)   /root/reference/style_transfer/sqrtm.py:9-25              (SQ)
//   _MatrixSquareRootNSLyap.backward  SQ:36-47 (iterative Lyapunov solve, 12 its)
// plus the closed-form backward of ST:163-181 down to  G = d loss / d srm  and  d loss / d mean.
//
// The loss is a cancellation (tr(St + S - 2 sqrt(.))): TF32/bf16 operands are not accurate
// enough, so these C x C chains run on the FP32 FMA pipe.  All five layers advance in lock-step: one "round" = one
// grouped launch whose CTAs are 64x64 tiles of every layer's GEMM (1+4+16+64+64 = 149 tiles ~ one wave of 148 SMs).
#include <vector>

#include "kernels.h"
#include "ptx.cuh"

namespace stb {

namespace {

constexpr int TS = 64;   // tile size
constexpr int KS = 16;   // k step

__global__ void __launch_bounds__(256)
sgemm_grouped_kernel(const GemmProb* __restrict__ probs, const uint32_t* __restrict__ tiles) {
  __shared__ __align__(16) float As[2][KS][TS + 4];
  __shared__ __align__(16) float Bs[2][KS][TS + 4];
  const uint32_t t = tiles[blockIdx.x];
  const GemmProb pr = probs[t >> 16];
  const int ti = (t >> 8) & 0xFF, tj = t & 0xFF;
  const int n = pr.n;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int n_prod = (pr.A2 != nullptr) ? 2 : 1;
  for (int pi = 0; pi < n_prod; ++pi) {
    const float* __restrict__ A = pi ? pr.A2 : pr.A;
    const float* __restrict__ B = pi ? pr.B2 : pr.B;
    const int tA = pi ? pr.transA2 : pr.transA, tB = pi ? pr.transB2 : pr.transB;
    const float alpha = pi ? pr.alpha2 : pr.alpha;
    float4 ra, rb;
    auto gload = [&](int k0) {
      if (!tA) {  // A[i][k]: thread -> row i = tid & 63, k quad = tid >> 6
        ra = *reinterpret_cast<const float4*>(A + (size_t)(ti * TS + (tid & 63)) * n + k0 + (tid >> 6) * 4);
      } else {    // A^T: element (i,k) = A[k][i]: thread -> k = tid >> 4, i quad = tid & 15
        ra = *reinterpret_cast<const float4*>(A + (size_t)(k0 + (tid >> 4)) * n + ti * TS + (tid & 15) * 4);
      }
      if (!tB) {  // B[k][j]
        rb = *reinterpret_cast<const float4*>(B + (size_t)(k0 + (tid >> 4)) * n + tj * TS + (tid & 15) * 4);
      } else {    // B^T: element (k,j) = B[j][k]
        rb = *reinterpret_cast<const float4*>(B + (size_t)(tj * TS + (tid & 63)) * n + k0 + (tid >> 6) * 4);
      }
    };
    auto sstore = [&](int buf) {
      if (!tA) {
        const int i = tid & 63, kq = (tid >> 6) * 4;
        As[buf][kq + 0][i] = ra.x * alpha; As[buf][kq + 1][i] = ra.y * alpha;
        As[buf][kq + 2][i] = ra.z * alpha; As[buf][kq + 3][i] = ra.w * alpha;
      } else {
        *reinterpret_cast<float4*>(&As[buf][tid >> 4][(tid & 15) * 4]) =
            make_float4(ra.x * alpha, ra.y * alpha, ra.z * alpha, ra.w * alpha);
      }
      if (!tB) {
        *reinterpret_cast<float4*>(&Bs[buf][tid >> 4][(tid & 15) * 4]) = rb;
      } else {
        const int j = tid & 63, kq = (tid >> 6) * 4;
        Bs[buf][kq + 0][j] = rb.x; Bs[buf][kq + 1][j] = rb.y; Bs[buf][kq + 2][j] = rb.z; Bs[buf][kq + 3][j] = rb.w;
      }
    };
    gload(0);
    sstore(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < n; k0 += KS) {
      const bool more = (k0 + KS) < n;
      if (more) gload(k0 + KS);
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      if (more) sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = ti * TS + ty * 4 + i;
    const int gj = tj * TS + tx * 4;
    float4 o = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    if (pr.Cadd != nullptr) {
      const float4 c = *reinterpret_cast<const float4*>(pr.Cadd + (size_t)gi * n + gj);
      o.x = fmaf(pr.beta, c.x, o.x); o.y = fmaf(pr.beta, c.y, o.y);
      o.z = fmaf(pr.beta, c.z, o.z); o.w = fmaf(pr.beta, c.w, o.w);
    }
    if (gi >= gj && gi < gj + 4) (&o.x)[gi - gj] += pr.gamma;
    *reinterpret_cast<float4*>(pr.D + (size_t)gi * n + gj) = o;
  }
}

// ---- block reduction helper (1024 threads)
__device__ float block_sum_1024(float v, float* s_red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < 32; ++i) t += s_red[i];
  return t;
}

// covariance from (reduced) raw sums:  mu = sums/N; cov = S_raw/N - mu mu^T + eps I     (ST:171-173, 177)
__global__ void __launch_bounds__(1024) w2_cov_kernel(const W2Layer* __restrict__ layers, int from_target) {
  __shared__ float s_red[32];
  const W2Layer L = layers[blockIdx.x];
  const int n = L.n;
  const float inv_n = from_target ? 1.f : 1.f / L.npix;
  const float* S = from_target ? L.srm_t : L.S_raw;
  const float* sm = from_target ? L.mean_t : L.sums;
  float* mu = from_target ? L.mean_t : L.mu;
  float* cov = from_target ? L.cov_t : L.cov;
  if (!from_target)
    for (int i = threadIdx.x; i < n; i += blockDim.x) mu[i] = sm[i] * inv_n;
  __syncthreads();
  float tr = 0.f;
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int i = e / n, j = e - i * n;
    float v = S[e] * inv_n - mu[i] * mu[j];
    if (i == j) { v += L.eps; tr += v; }
    cov[e] = v;
  }
  tr = block_sum_1024(tr, s_red);
  float md = 0.f;
  if (!from_target)
    for (int i = threadIdx.x; i < n; i += blockDim.x) { const float d = mu[i] - L.mean_t[i]; md += d * d; }
  md = block_sum_1024(md, s_red);
  if (threadIdx.x == 0) {
    if (from_target) L.scal[W2S_TR_COV_T] = tr;
    else { L.scal[W2S_TR_COV] = tr; L.scal[W2S_MEAN_DIFF] = md / n; }
  }
}

// Y = M / ||M||_F, Z = I        (SQ:15-19)
__global__ void __launch_bounds__(1024) w2_ns_init_kernel(const W2Layer* __restrict__ layers, int from_target) {
  __shared__ float s_red[32];
  const W2Layer L = layers[blockIdx.x];
  const int n = L.n;
  const float* M = from_target ? L.cov_t : L.M;
  float ss = 0.f;
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) ss = fmaf(M[e], M[e], ss);
  ss = block_sum_1024(ss, s_red);
  const float norm = sqrtf(ss);
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int i = e / n, j = e - i * n;
    L.Y[0][e] = M[e] / norm;
    L.Z[0][e] = (i == j) ? 1.f : 0.f;
  }
  if (threadIdx.x == 0) L.scal[W2S_NORM_A] = norm;
}

// target: P = Y sqrt(norm)                                              (ST:159, SQ:25)
__global__ void __launch_bounds__(1024) w2_target_finish_kernel(const W2Layer* __restrict__ layers) {
  const W2Layer L = layers[blockIdx.x];
  const int n = L.n;
  const float s = sqrtf(L.scal[W2S_NORM_A]);
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) L.P[e] = L.Y[0][e] * s;
}

// forward finish: R = Y sqrt(normA); loss; seeds of the Lyapunov backward    (SQ:25, ST:178-181, SQ:37-41)
__global__ void __launch_bounds__(1024) w2_fwd_finish_kernel(const W2Layer* __restrict__ layers, float* loss_terms) {
  __shared__ float s_red[32];
  const W2Layer L = layers[blockIdx.x];
  const int n = L.n;
  const float* Y = L.Y[0];
  float ss = 0.f, tr = 0.f;
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const float v = Y[e];
    ss = fmaf(v, v, ss);
    if (e / n == e % n) tr += v;
  }
  ss = block_sum_1024(ss, s_red);
  tr = block_sum_1024(tr, s_red);
  const float sq = sqrtf(L.scal[W2S_NORM_A]);
  const float norm_y = sqrtf(ss);
  const float norm_r = sq * norm_y;         // ||R||_F
  const float tr_r = tr * sq;
  const float seed = -2.f * L.weight / (n * norm_r);  // grad_output / norm_z with grad_output = -2 w / C * I
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    L.A[0][e] = Y[e] / norm_y;                        // a = z / ||z||
    L.Q[0][e] = (e / n == e % n) ? seed : 0.f;
  }
  if (threadIdx.x == 0) {
    const float cov_diff = (L.scal[W2S_TR_COV_T] + L.scal[W2S_TR_COV] - 2.f * tr_r) / n;
    const float l = (L.scal[W2S_MEAN_DIFF] + cov_diff) * L.weight;
    L.scal[W2S_LOSS] = l;
    loss_terms[blockIdx.x] = l;
  }
}

// backward finish: Gs = Gc + Gc^T; gmu = 2w(mu - mu_t)/C - Gs mu; emit bf16 Gs/N ([C][C]) and fp32 gmu/N
__global__ void __launch_bounds__(1024) w2_bwd_finish_kernel(const W2Layer* __restrict__ layers) {
  const W2Layer L = layers[blockIdx.x];
  const int n = L.n;
  const float inv_n = 1.f / L.npix;
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int i = e / n, j = e - i * n;
    const float gs = L.Gc[e] + L.Gc[j * n + i];
    L.Gs[e] = gs;
    L.gs_bf16[e] = __float2bfloat16(gs * inv_n);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < n; ++j) s = fmaf(L.Gs[i * n + j], L.mu[j], s);
    const float gm = 2.f * L.weight * (L.mu[i] - L.mean_t[i]) / n - s;
    L.gmu_bias[i] = gm * inv_n;
  }
}

}  // namespace

// ================================================================================================ host engine
size_t W2Engine::layer_floats(int n) {
  // cov, M, X, Y[2], Z[2], T, A[2], Q[2], E, X1, X23, U, Gc, Gs, P, cov_t, srm_t  (21 matrices) + vectors
  return (size_t)21 * n * n + 8 * (size_t)n + 64;
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
    for (int j = 0; j < nt; ++j) tiles.push_back((uint32_t)idx << 16 | (uint32_t)i << 8 | (uint32_t)j);
}

static GemmProb mk(int n, float* D, const float* A, int tA, const float* B, int tB, float alpha, float gamma = 0.f,
                   const float* Cadd = nullptr, float beta = 0.f, const float* A2 = nullptr, int tA2 = 0,
                   const float* B2 = nullptr, int tB2 = 0, float alpha2 = 0.f) {
  GemmProb p{};
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
        add_prob(probs, tiles, mk(L.n, L.Y[d], L.Y[s], 0, L.T, 0, 1.f));
        add_prob(probs, tiles, mk(L.n, L.Z[d], L.T, 0, L.Z[s], 0, 1.f));
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
  for (int l = 0; l < 5; ++l) { W2Layer& L = host_layers[l]; add_prob(probs, tiles, mk(L.n, L.X, L.P, 0, L.cov, 0, 1.f)); }
  end_round();
  begin_round();
  for (int l = 0; l < 5; ++l) { W2Layer& L = host_layers[l]; add_prob(probs, tiles, mk(L.n, L.M, L.X, 0, L.P, 0, 1.f)); }
  end_round();
  r_fwd_ns_begin = (int)rounds.size();
  ns_rounds();
  r_fwd_end = (int)rounds.size();
  // (c) backward: 12 x { E = 3I - a a ; X1 = q E, X23 = a^T q - q a, a' = a E / 2 ; q' = X1/2 - a^T X23 / 2 }
  r_bwd_begin = (int)rounds.size();
  for (int it = 0; it < 12; ++it) {
    const int s = it & 1, d = s ^ 1;
    begin_round();
    for (int l = 0; l < 5; ++l) { W2Layer& L = host_layers[l]; add_prob(probs, tiles, mk(L.n, L.E, L.A[s], 0, L.A[s], 0, -1.f, 3.f)); }
    end_round();
    begin_round();
    for (int l = 0; l < 5; ++l) {
      W2Layer& L = host_layers[l];
      add_prob(probs, tiles, mk(L.n, L.X1, L.Q[s], 0, L.E, 0, 1.f));
      add_prob(probs, tiles, mk(L.n, L.X23, L.A[s], 1, L.Q[s], 0, 1.f, 0.f, nullptr, 0.f, L.Q[s], 0, L.A[s], 0, -1.f));
      if (it < 11) add_prob(probs, tiles, mk(L.n, L.A[d], L.A[s], 0, L.E, 0, 0.5f));
    }
    end_round();
    begin_round();
    for (int l = 0; l < 5; ++l) {
      W2Layer& L = host_layers[l];
      add_prob(probs, tiles, mk(L.n, L.Q[d], L.A[s], 1, L.X23, 0, -0.5f, 0.f, L.X1, 0.5f));
    }
    end_round();
  }
  // after 12 its q is in Q[0].  U = P^T q ; Gc = 0.5 U P^T + (w/C) I  (gamma patched per layer in set_weights)
  begin_round();
  for (int l = 0; l < 5; ++l) { W2Layer& L = host_layers[l]; add_prob(probs, tiles, mk(L.n, L.U, L.P, 1, L.Q[0], 0, 1.f)); }
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
  for (int r = r0; r < r1; ++r) {
    sgemm_grouped_kernel<<<rounds[r].n_tiles, 256, 0, s>>>(d_probs, d_tiles + rounds[r].first_tile);
  }
  STB_CUDA_CHECK(cudaGetLastError());
  return STB_OK;
}

int W2Engine::build_targets(cudaStream_t s) {
  // srm_t / mean_t already hold the blended target moments (ST:443-450)
  w2_cov_kernel<<<5, 1024, 0, s>>>(d_layers, 1);
  w2_ns_init_kernel<<<5, 1024, 0, s>>>(d_layers, 1);
  STB_TRY(run_rounds(r_target_begin, r_target_end, s));
  w2_target_finish_kernel<<<5, 1024, 0, s>>>(d_layers);
  STB_CUDA_CHECK(cudaGetLastError());
  return STB_OK;
}

int W2Engine::forward_backward(float* loss_terms, cudaStream_t s) {
  w2_cov_kernel<<<5, 1024, 0, s>>>(d_layers, 0);
  STB_TRY(run_rounds(r_fwd_begin, r_fwd_ns_begin, s));
  w2_ns_init_kernel<<<5, 1024, 0, s>>>(d_layers, 0);
  STB_TRY(run_rounds(r_fwd_ns_begin, r_fwd_end, s));
  w2_fwd_finish_kernel<<<5, 1024, 0, s>>>(d_layers, loss_terms);
  STB_TRY(run_rounds(r_bwd_begin, r_bwd_end, s));
  w2_bwd_finish_kernel<<<5, 1024, 0, s>>>(d_layers);
  STB_CUDA_CHECK(cudaGetLastError());
  return STB_OK;
}

}  // namespace stb
