// Wasserstein-2 style loss on the five tap covariances, forward and backward, at fp32 accuracy on tcgen05:
//   StyleLossW2.__init__/forward      /root/reference/style_transfer/style_transfer.py:149-181  (ST)
//   sqrtm_ns (12 Newton-Schulz its)   /root/reference/style_transfer/sqrtm.py:9-25              (SQ)
//   _MatrixSquareRootNSLyap.backward  SQ:36-47 (iterative Lyapunov solve, 12 its)
// plus the closed-form backward of ST:163-181 down to  G = d loss / d srm  and  d loss / d mean.
//
// The loss is a cancellation (tr(St + S - 2 sqrt(.))): plain TF32/bf16 operands are not
// accurate enough.  Every C x C matrix of the chain therefore lives as TWO fp32 planes, x = hi + lo, where hi has its
// low 13 mantissa bits cleared (exactly representable in TF32) and lo = x - hi; a product is evaluated as
//     a*b ~= lo_a*hi_b + hi_a*lo_b + hi_a*hi_b        (3 x tcgen05.mma kind::tf32, fp32 accumulate in TMEM),
// the dropped lo*lo term being 2^-22 relative.  The split of a result is done by the producing GEMM's epilogue, so a
// GEMM only ever streams ready-made planes through TMA.
//
// All five layers advance in lock-step: one "round" = one grouped launch whose CTAs are 128 x 64 output tiles of
// every layer's GEMM (75 or 150 CTAs, largest matrices first).  Both tcgen05 operands are K-major SW128 tiles:
// A[m][k] from the row-major planes of A, B[n][k] from the planes of B^T -- every matrix that is later used as a
// right factor is therefore written TWICE by its producer: the planes of D (TMA store of the staged tile) and the
// planes of D^T (scalar stores, coalesced across the warp because TMEM lane = row).
//
// The products are formed in exactly the reference's order, Y <- Y T, Z <- T Z, and NOTHING is symmetrised, although
// every matrix here is symmetric in exact arithmetic.  Two shortcuts were tried and measured to break parity on real
// (ill-conditioned, eps = 1e-4) covariances, where the coupled Newton-Schulz iteration amplifies a perturbation of
// the small-eigenvalue directions by 1.5^12:
//   * reading a computed product as its own transpose turns the error recursion E' = E/2 into
//     E' = E - A^(1/2) E A^(-1/2) / 2: NaN at 2048^2;
//   * computing only the upper triangle and mirroring it bit-for-bit (38 % fewer tiles) is stable but biases the
//     style terms by -1e-3 ... -3e-3 at 512^2 and above (the CPU emulation of that schedule in fp32 shows the same),
//     i.e. above the 1e-3 loss bar.  The full products stay within 3e-4 of the fp64 chain.
// Tensor maps are encoded once per workspace binding.
#include <cuda_fp16.h>

#include <cstdlib>
#include <map>
#include <set>
#include <utility>
#include <vector>

#include "kernels.h"
#include "ptx.cuh"

namespace stb {

namespace {

constexpr int TM = 128, TN = 64, TKF = 32;      // tile rows / cols, k floats per stage (128 B swizzle row)
constexpr int T_STAGES = 4;
// TMEM: T_MAX_ACC hi*hi accumulators + two for the cross terms (lo*hi and hi*lo apart), 64 columns each.  The k-steps
// are dealt to the hi*hi accumulators ROUND-ROBIN: each still sums n_steps / T_MAX_ACC products (short truncating
// chains, section 5 of DESIGN.md), but consecutive MMAs never target the same accumulator -- a tcgen05.mma that
// accumulates into the tile its predecessor is still writing waits for it (measured: ~125 cycles per MMA in dependent
// order vs the 32 its math takes).
constexpr int T_MAX_ACC = 6;
constexpr int T_TMEM_COLS = 512;
constexpr int T_CROSS_COL = T_MAX_ACC * TN;        // lo*hi at T_CROSS_COL, hi*lo at T_CROSS_COL + TN
constexpr int A_PLANE_BYTES = TM * 128;          // 16 KiB
constexpr int B_PLANE_BYTES = TN * 128;          // 8 KiB
constexpr int T_STAGE_BYTES = 2 * A_PLANE_BYTES + 2 * B_PLANE_BYTES;  // A_hi, A_lo, B_hi, B_lo = 48 KiB
constexpr int T_OFF_BAR = T_STAGES * T_STAGE_BYTES;
constexpr int T_OFF_TMEMPTR = T_OFF_BAR + (2 * T_STAGES + 1) * 8;
constexpr int T_OFF_RED = T_OFF_TMEMPTR + 16;
constexpr int T_SMEM_BYTES = T_OFF_RED + 64 + 1024;
constexpr int T_EPI_THREADS = 256;                 // two epilogue groups of four warps: one per 32-column half of the tile
constexpr int T_THREADS = 64 + T_EPI_THREADS;
constexpr int NRED = 128;  // max reduction partials per layer
constexpr int NB = 128;    // helper-kernel CTAs per layer: they are latency-bound element-wise passes (32 CTAs: 48 us)
static_assert(NB <= NRED, "the covariance kernel writes one partial per CTA");

// x = hi + lo with both parts rounded to nearest TF32 (the tensor core would otherwise truncate them: a biased
// error that the loss' cancellation amplifies); x - hi is exact in fp32.
__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float v, float& h, float& l) {
  h = rna_tf32(v);
  l = rna_tf32(v - h);
}
__device__ __forceinline__ void store_split(float* m, size_t nn, size_t e, float v) {
  float h, l;
  split_tf32(v, h, l);
  m[e] = h;
  m[nn + e] = l;
}
__device__ __forceinline__ float load2(const float* m, size_t nn, size_t e) { return m[e] + m[nn + e]; }

// fp16 plane pairs (the Newton-Schulz / Lyapunov iterates): x = hi + lo' * 2^-11 with hi = fp16(x) and
// lo' = fp16((x - hi) * 2^11) -- 22 significant bits like the TF32 pair, in HALF the bytes, and kind::f16 MMAs.  The
// residual is stored scaled so that it lives in the same exponent range as x (no fp16 underflow); the products
// lo'_a hi_b + hi_a lo'_b go to their own accumulator, which the epilogue scales by 2^-11.  Every matrix kept this way
// is normalised (|entries| <~ 1e2), far inside the fp16 range; the un-normalised ones (cov, P, X, M, U, Gc and the
// final q) stay TF32 pairs.
constexpr float H_LO_SCALE = 2048.f, H_LO_INV = 1.f / 2048.f;
__device__ __forceinline__ void split_half(float v, __half& h, __half& l) {
  h = __float2half_rn(v);
  l = __float2half_rn((v - __half2float(h)) * H_LO_SCALE);
}
__device__ __forceinline__ void store_split_h(float* m, size_t nn, size_t e, float v) {
  __half* mh = reinterpret_cast<__half*>(m);
  __half h, l;
  split_half(v, h, l);
  mh[e] = h;
  mh[nn + e] = l;
}
__device__ __forceinline__ float load2_h(const float* m, size_t nn, size_t e) {
  const __half* mh = reinterpret_cast<const __half*>(m);
  return __half2float(mh[e]) + __half2float(mh[nn + e]) * H_LO_INV;
}

// D = alpha * A * B + gamma * I on one 128 x 64 tile.  A matrix is 4 planes of n*n floats: hi, lo, hi^T, lo^T.
// smem per stage: A planes [128 rows][32 k] and B^T planes [64 rows (n)][32 k], all K-major SW128.
constexpr int STG_OFF = 0;                          // 4 store tiles (2 column halves x hi/lo) of 16 KiB, SW128, in
static_assert(4 * A_PLANE_BYTES <= T_STAGES * T_STAGE_BYTES, "epilogue smem");  // the drained pipeline buffers

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// The whole chain of rounds [r0, r1) in ONE persistent, cooperatively launched kernel: one CTA per SM, tile t of a
// round goes to CTA t % gridDim.x, a grid-wide barrier separates two rounds.  Replaces 24-26 dependent launches per
// phase: barrier init, TMEM allocation and the launch / drain latency of a grid are paid once, a round boundary costs
// one atomic + one acquire spin (~1.5 us instead of ~4 us of launch gap).  The arithmetic of a tile is unchanged.
__global__ void __launch_bounds__(T_THREADS, 1)
w2_chain_kernel(const W2Round* __restrict__ rounds, int r0, int r1, unsigned* __restrict__ grid_counter,
                unsigned long long* __restrict__ trace) {
  // trace (STB_W2_TRACE=1, diagnostics): CTA 0 stamps %globaltimer at 8 points of every round --
  // 0 round start, 1 first stage landed, 2 accumulators complete, 3 tile staged + TMA stores issued, 4 stores complete,
  // 5 arrived at the grid barrier, 6 barrier passed
  auto stamp = [&](int round, int k) {
    if (trace != nullptr && blockIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      trace[round * 8 + k] = t;
    }
  };
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + T_OFF_BAR);
  uint64_t* empty = full + T_STAGES;
  uint64_t* t_full = empty + T_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + T_OFF_TMEMPTR);
  float* s_red = reinterpret_cast<float*>(smem + T_OFF_RED);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < T_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(t_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<T_TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // pipeline ring state, persistent across tiles and rounds (the producer and the MMA warp advance identically)
  int s = 0;
  uint32_t ph = 0, tile_parity = 0;
  unsigned barriers_done = 0;

  // The tile -> CTA map is static, so everything a CTA needs for its first tile of the NEXT round (problem record,
  // tensor-map descriptors) is fetched BEFORE the grid barrier: after it the producer can issue its TMA loads at once
  // instead of first walking rounds[] -> tiles[] -> probs[] -> descriptor through dependent L2 reads.
  uint32_t pre_t = 0;
  int pre_n_tiles = 0;
  TcProb pre_pr = {};
  auto prefetch_round = [&](int round) {
    const W2Round& nx = rounds[round];
    pre_n_tiles = nx.n_tiles;
    if ((int)blockIdx.x < pre_n_tiles) {
      pre_t = nx.tiles[blockIdx.x];
      pre_pr = nx.probs[pre_t >> 16];
      if (threadIdx.x == 0) {
        tma_prefetch_desc(pre_pr.amap); tma_prefetch_desc(pre_pr.amap + 1);
        tma_prefetch_desc(pre_pr.bmap); tma_prefetch_desc(pre_pr.bmap + 1);
        tma_prefetch_desc(pre_pr.dmap); tma_prefetch_desc(pre_pr.dmap + 1);
      }
    }
  };
  prefetch_round(r0);

  for (int round = r0; round < r1; ++round) {
    const W2Round& rp = rounds[round];
    const int n_tiles = pre_n_tiles;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const bool first = tile == (int)blockIdx.x;
      const uint32_t t = first ? pre_t : rp.tiles[tile];
      const TcProb pr = first ? pre_pr : rp.probs[t >> 16];
      const int ti = (t >> 8) & 0xFF, tj = t & 0xFF;
      const int n = pr.n;
      const int k_elems = pr.in_half ? 2 * TKF : TKF;          // elements per 128-byte stage row
      const int n_k = n / k_elems;
      const int k_steps = n / (pr.in_half ? 16 : 8);           // MMA K steps of the whole tile

      if (warp == 0) {
        // ---- TMA producer: four planes per stage
        const bool leader = elect_one();   // every lane walks the ring state, the elected one issues
        if (leader && first) stamp(round, 0);
        for (int k = 0; k < n_k; ++k) {
          if (leader) {
            mbar_wait(&empty[s], ph ^ 1);
            mbar_expect_tx(&full[s], T_STAGE_BYTES);
            uint8_t* st = smem + s * T_STAGE_BYTES;
            tma_load_2d(st, pr.amap, &full[s], k * k_elems, ti * TM);
            tma_load_2d(st + A_PLANE_BYTES, pr.amap + 1, &full[s], k * k_elems, ti * TM);
            tma_load_2d(st + 2 * A_PLANE_BYTES, pr.bmap, &full[s], k * k_elems, tj * TN);  // rows of B^T
            tma_load_2d(st + 2 * A_PLANE_BYTES + B_PLANE_BYTES, pr.bmap + 1, &full[s], k * k_elems, tj * TN);
          }
          if (++s == T_STAGES) { s = 0; ph ^= 1; }
        }
        __syncwarp();
      } else if (warp == 1) {
        // ---- MMA issuer: 4 k-steps (K = 8) x 3 split products per stage.  The tensor core aligns every product to
        // the accumulator and truncates, so a chain of m products loses ~2^-25 * m of the sum, systematically; on the
        // ill-conditioned covariances of real activations the Newton-Schulz iteration turns that into a -1e-3 bias of
        // the style terms (measured).  hi*hi therefore runs in short chains: the n/8 k-steps are dealt out in equal
        // runs to seven TMEM accumulators (2 steps each at C = 64 ... 10 at C = 512), the small cross terms go to an
        // eighth, and the epilogue adds all of them in round-to-nearest fp32.
        constexpr uint32_t idesc = umma_idesc_tf32(TM, TN);
        constexpr uint32_t idesc_h = umma_idesc_f16(TM, TN);
        constexpr uint32_t dhi = umma_desc_hi_sw128(1024);
        const bool leader = elect_one();
        const bool half_in = pr.in_half != 0;
        for (int k = 0; k < n_k; ++k) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          if (leader && first && k == 0) stamp(round, 1);
          if (leader) {
            const uint32_t base = smem_u32(smem + s * T_STAGE_BYTES);
            const uint32_t a_h = umma_desc_lo(base), a_l = umma_desc_lo(base + A_PLANE_BYTES);
            const uint32_t b_h = umma_desc_lo(base + 2 * A_PLANE_BYTES);
            const uint32_t b_l = umma_desc_lo(base + 2 * A_PLANE_BYTES + B_PLANE_BYTES);
            if (!half_in) {
#pragma unroll
              for (int kk = 0; kk < TKF / 8; ++kk) {  // 8 floats = 32 bytes per K step -> +2 in the descriptor
                const int step = k * (TKF / 8) + kk;
                const uint32_t t_main = tmem_base + (step % T_MAX_ACC) * TN, t_cross = tmem_base + T_CROSS_COL;
                umma_tf32_split(t_cross, a_l + 2 * kk, dhi, b_h + 2 * kk, dhi, idesc, step > 0);
                umma_tf32_split(t_main, a_h + 2 * kk, dhi, b_h + 2 * kk, dhi, idesc, step >= T_MAX_ACC);
                umma_tf32_split(t_cross + TN, a_h + 2 * kk, dhi, b_l + 2 * kk, dhi, idesc, step > 0);
              }
            } else {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {        // 16 halfs = 32 bytes per K step -> +2 in the descriptor
                const int step = k * 4 + kk;
                const uint32_t t_main = tmem_base + (step % T_MAX_ACC) * TN, t_cross = tmem_base + T_CROSS_COL;
                umma_bf16_split(t_cross, a_l + 2 * kk, dhi, b_h + 2 * kk, dhi, idesc_h, step > 0);
                umma_bf16_split(t_main, a_h + 2 * kk, dhi, b_h + 2 * kk, dhi, idesc_h, step >= T_MAX_ACC);
                umma_bf16_split(t_cross + TN, a_h + 2 * kk, dhi, b_l + 2 * kk, dhi, idesc_h, step > 0);
              }
            }
            umma_commit(&empty[s]);
          }
          __syncwarp();
          if (++s == T_STAGES) { s = 0; ph ^= 1; }
        }
        if (leader) umma_commit(t_full);
        __syncwarp();
      } else {
        // ---- epilogue (2 groups x 128 threads, TMEM lane = tile row; group g drains columns [32 g, 32 g + 32))
        const int wq = warp & 3;
        const int grp = (warp - 2) >> 2;
        const bool issuer = warp == 2 && lane == 0;   // the one thread that issues the TMA stores / writes the partials
        const int r = wq * 32 + lane;
        const int gi = ti * TM + r;
        const size_t nn = (size_t)n * n;
        uint8_t* stg = smem + STG_OFF;
        mbar_wait(t_full, tile_parity);
        tc_fence_after();
        if (issuer && first) stamp(round, 2);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
        const int n_chunks = k_steps < T_MAX_ACC ? k_steps : T_MAX_ACC;   // hi*hi accumulators in use
        const bool valid = gi < n;
        const float cross_scale = pr.in_half ? H_LO_INV : 1.f;   // fp16 pairs keep the residual plane times 2^11
        const bool half_out = pr.out_half != 0;
        __half* Dh = reinterpret_cast<__half*>(pr.D);
        float ssq = 0.f, tr = 0.f;
        {
          const int h = grp;
          uint32_t v[32], v2[32];
          float acc[32];
          // chunk sums in the same order as ever (c = 0, 1, 2, ...); the TMEM loads go out in pairs under one wait
          tmem_ld_32x32(taddr + h * 32, v);
          if (n_chunks > 1) tmem_ld_32x32(taddr + TN + h * 32, v2);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) acc[e] = __uint_as_float(v[e]);
          if (n_chunks > 1) {
#pragma unroll
            for (int e = 0; e < 32; ++e) acc[e] += __uint_as_float(v2[e]);
          }
          for (int c = 2; c < n_chunks; c += 2) {
            tmem_ld_32x32(taddr + c * TN + h * 32, v);
            if (c + 1 < n_chunks) tmem_ld_32x32(taddr + (c + 1) * TN + h * 32, v2);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) acc[e] += __uint_as_float(v[e]);
            if (c + 1 < n_chunks) {
#pragma unroll
              for (int e = 0; e < 32; ++e) acc[e] += __uint_as_float(v2[e]);
            }
          }
          tmem_ld_32x32(taddr + T_CROSS_COL + h * 32, v);
          tmem_ld_32x32(taddr + T_CROSS_COL + TN + h * 32, v2);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(v2[e]));
          const int gj0 = tj * TN + h * 32;
#pragma unroll
          for (int e = 0; e < 32; ++e) {   // finished values of this row's 32 columns
            float o = (acc[e] + __uint_as_float(v[e]) * cross_scale) * pr.alpha;
            if (gi == gj0 + e) { o += pr.gamma; tr += o; }
            if (valid) ssq = fmaf(o, o, ssq);
            acc[e] = o;
          }
          if (!half_out) {
            uint8_t* row_hi = stg + (h * 2) * A_PLANE_BYTES + r * 128;
            uint8_t* row_lo = row_hi + A_PLANE_BYTES;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float vh[4], vl[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) split_tf32(acc[4 * q + e], vh[e], vl[e]);
              const int chunk = (q ^ (r & 7)) * 16;  // 128-byte swizzle, as the TMA store expects
              *reinterpret_cast<float4*>(row_hi + chunk) = make_float4(vh[0], vh[1], vh[2], vh[3]);
              *reinterpret_cast<float4*>(row_lo + chunk) = make_float4(vl[0], vl[1], vl[2], vl[3]);
              if (pr.write_t && valid) {  // D^T planes: for a fixed column the 32 lanes (consecutive rows) write 128 B
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const size_t at = 2 * nn + (size_t)(gj0 + 4 * q + e) * n + gi;
                  pr.D[at] = vh[e];
                  pr.D[nn + at] = vl[e];
                }
              }
            }
          } else {
            // fp16 pair: one 128 x 64 tile per plane (row = 64 halfs = 128 bytes), hi plane at stg, lo plane after it;
            // this half of the row fills 16-byte chunks 4h .. 4h+3
            uint8_t* row_hi = stg + r * 128;
            uint8_t* row_lo = row_hi + A_PLANE_BYTES;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              __half hh[8], ll[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) split_half(acc[8 * q + e], hh[e], ll[e]);
              const int chunk = ((h * 4 + q) ^ (r & 7)) * 16;
              *reinterpret_cast<uint4*>(row_hi + chunk) = *reinterpret_cast<const uint4*>(hh);
              *reinterpret_cast<uint4*>(row_lo + chunk) = *reinterpret_cast<const uint4*>(ll);
              if (pr.write_t && valid) {  // D^T planes: for a fixed column the 32 lanes write 64 contiguous bytes
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const size_t at = 2 * nn + (size_t)(gj0 + 8 * q + e) * n + gi;
                  Dh[at] = hh[e];
                  Dh[nn + at] = ll[e];
                }
              }
            }
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        // the D^T planes were written through the generic proxy; the next round reads them through TMA (async proxy)
        asm volatile("fence.proxy.async;" ::: "memory");
        named_bar_sync(1, T_EPI_THREADS);
        if (issuer) {
          const CUtensorMap* dm = pr.dmap;
          if (!half_out) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              tma_store_2d(dm, stg + (h * 2) * A_PLANE_BYTES, tj * TN + h * 32, ti * TM);
              tma_store_2d(dm + 1, stg + (h * 2 + 1) * A_PLANE_BYTES, tj * TN + h * 32, ti * TM);
            }
          } else {
            tma_store_2d(dm, stg, tj * TN, ti * TM);
            tma_store_2d(dm + 1, stg + A_PLANE_BYTES, tj * TN, ti * TM);
          }
          tma_store_commit();
          if (first) stamp(round, 3);
        }
        if (pr.red_out != nullptr) {
          ssq = warp_sum(ssq);
          tr = warp_sum(tr);
          if (lane == 0) { s_red[(grp * 4 + wq) * 2] = ssq; s_red[(grp * 4 + wq) * 2 + 1] = tr; }
          named_bar_sync(1, T_EPI_THREADS);
          if (issuer) {
            const int ntj = n / TN;
            pr.red_out[(ti * ntj + tj) * 2] = ((s_red[0] + s_red[2]) + (s_red[4] + s_red[6])) +
                                              ((s_red[8] + s_red[10]) + (s_red[12] + s_red[14]));
            pr.red_out[(ti * ntj + tj) * 2 + 1] = ((s_red[1] + s_red[3]) + (s_red[5] + s_red[7])) +
                                                  ((s_red[9] + s_red[11]) + (s_red[13] + s_red[15]));
          }
        }
        if (issuer) {
          tma_store_wait_all0();   // stores complete (not only read): the staging smem is free, the data is out
          if (first) stamp(round, 4);
        }
      }
      // ---- end of tile: TMEM drained, staging smem free, every role done
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      tile_parity ^= 1;
    }
    // ---- grid-wide barrier between two rounds (the last round of the launch ends with the kernel)
    if (round + 1 < r1) {
      prefetch_round(round + 1);
      __syncthreads();
      if (threadIdx.x == 0) {
        ++barriers_done;
        const unsigned target = barriers_done * gridDim.x;
        stamp(round, 5);
        asm volatile("fence.proxy.async;" ::: "memory");
        __threadfence();
        atomicAdd(grid_counter, 1u);
        if (ld_acquire_gpu_u32(grid_counter) < target) {
          const long long t0 = clock64();
          while (ld_acquire_gpu_u32(grid_counter) < target)
            if (clock64() - t0 > (8ll << 30)) __trap();   // ~4 s at 2 GHz: a lost CTA must not hang the stream
        }
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
        stamp(round, 6);
      }
      __syncthreads();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<T_TMEM_COLS>(tmem_base);
}

// ---- helpers: grid (5 layers, NB CTAs), 256 threads; reductions via fixed-order partials (deterministic)
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
// Sum of the (<= NRED) reduction partials {a_i, b_i}, by the whole 256-thread block in a fixed tree (every CTA of a
// layer gets the same bits).  Was a serial loop in every thread: 32-128 dependent-issue global loads at the head of
// each element-wise helper, ~20 us of the 40-50 us those kernels took.
__device__ __forceinline__ void sum_partials(const float* red, int count, float& a, float& b, float* s_red16) {
  float va = 0.f, vb = 0.f;
  if ((int)threadIdx.x < count) { va = red[2 * threadIdx.x]; vb = red[2 * threadIdx.x + 1]; }
  va = warp_sum(va);
  vb = warp_sum(vb);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { s_red16[threadIdx.x >> 5] = va; s_red16[8 + (threadIdx.x >> 5)] = vb; }
  __syncthreads();
  a = 0.f; b = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { a += s_red16[i]; b += s_red16[8 + i]; }
}
__device__ __forceinline__ int gemm_tiles(int n) { return ((n + TM - 1) / TM) * (n / TN); }

// covariance from (reduced) raw sums:  mu = sums/N; cov = S_raw/N - mu mu^T + eps I     (ST:171-173, 177)
// target mode: cov_t from (mean_t, srm_t), plus sum-of-squares partials of cov_t for the NS normalisation.
__global__ void __launch_bounds__(256) w2_cov_kernel(const W2Layer* __restrict__ layers, int from_target) {
  __shared__ float s_red[8];
  const W2Layer L = layers[blockIdx.x];
  const int n = L.n;
  const size_t nn = (size_t)n * n;
  const float inv_n = from_target ? 1.f : 1.f / L.npix;
  const float* S = from_target ? L.srm_t : L.S_raw;
  const float* sm = from_target ? L.mean_t : L.sums;
  float* cov = from_target ? L.cov_t : L.cov;
  float ssq = 0.f;
  for (int e = blockIdx.y * 256 + threadIdx.x; e < n * n; e += NB * 256) {
    const int i = e / n, j = e - i * n;
    const int lo = min(i, j), hi = max(i, j);  // read the upper triangle of S: cov must be symmetric bit-for-bit
    float v = S[(size_t)lo * n + hi] * inv_n - (sm[lo] * inv_n) * (sm[hi] * inv_n);
    if (i == j) v += L.eps;
    store_split(cov, nn, e, v);
    store_split(cov + 2 * nn, nn, e, v);  // cov^T = cov bit-for-bit (upper triangle of S read for both)
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

// Y = M / ||M||_F, Z = I        (SQ:15-19).  ||M||^2 arrives as partials (cov kernel: NB; GEMM tiles otherwise)
__global__ void __launch_bounds__(256) w2_ns_init_kernel(const W2Layer* __restrict__ layers, int from_target) {
  __shared__ float s_red16[16];
  const W2Layer L = layers[blockIdx.x];
  const int n = L.n;
  const size_t nn = (size_t)n * n;
  const float* M = from_target ? L.cov_t : L.M;
  float ss, dummy;
  sum_partials(L.red, from_target ? NB : gemm_tiles(n), ss, dummy, s_red16);
  const float norm = sqrtf(ss);
  for (int e = blockIdx.y * 256 + threadIdx.x; e < n * n; e += NB * 256) {
    const int i = e / n, j = e - i * n;
    // Y, Z are fp16 plane pairs: planes hi, lo', hi^T, lo'^T of n*n HALFS each
    store_split_h(L.Y[0], nn, e, load2(M, nn, e) / norm);
    store_split_h(L.Y[0], nn, 2 * nn + e, load2(M + 2 * nn, nn, e) / norm);  // Y^T from the planes of M^T (cov_t: symmetric)
    const __half z0 = __float2half_rn((i == j) ? 1.f : 0.f), zero = __float2half_rn(0.f);
    __half* Zh = reinterpret_cast<__half*>(L.Z[0]);
    Zh[e] = z0; Zh[nn + e] = zero; Zh[2 * nn + e] = z0; Zh[3 * nn + e] = zero;
  }
  if (blockIdx.y == 0 && threadIdx.x == 0) L.scal[W2S_NORM_A] = norm;
}

// target: P = Y sqrt(norm)                                              (ST:159, SQ:25)
__global__ void __launch_bounds__(256) w2_target_finish_kernel(const W2Layer* __restrict__ layers) {
  const W2Layer L = layers[blockIdx.x];
  const int n = L.n;
  const size_t nn = (size_t)n * n;
  const float s = sqrtf(L.scal[W2S_NORM_A]);
  for (int e = blockIdx.y * 256 + threadIdx.x; e < n * n; e += NB * 256) {
    store_split(L.P, nn, e, load2_h(L.Y[0], nn, e) * s);
    store_split(L.P + 2 * nn, nn, e, load2_h(L.Y[0], nn, 2 * nn + e) * s);  // P^T from the planes of Y^T
  }
}

// forward finish: R = Y sqrt(normA); loss; seeds of the Lyapunov backward    (SQ:25, ST:178-181, SQ:37-41).
// ||Y||^2 and tr(Y) arrive as per-tile partials written by the last NS round.
__global__ void __launch_bounds__(256) w2_fwd_finish_kernel(const W2Layer* __restrict__ layers, float* loss_terms) {
  __shared__ float s_red16[16];
  const W2Layer L = layers[blockIdx.x];
  const int n = L.n;
  const size_t nn = (size_t)n * n;
  float ss, tr;
  sum_partials(L.red, gemm_tiles(n), ss, tr, s_red16);
  const float sq = sqrtf(L.scal[W2S_NORM_A]);
  const float norm_y = sqrtf(ss);
  const float norm_r = sq * norm_y;                     // ||R||_F
  const float tr_r = tr * sq;
  const float seed = -2.f * L.weight / (n * norm_r);    // grad_output / ||z|| with grad_output = -2 w / C * I
  // q only ever enters linearly (q' = q E / 2, U = P^T q, Gc = U P^T / 2): the iteration runs on q * 2^k with
  // |seed| * 2^k in [0.5, 1), which keeps the fp16 pair of a 1e-5-sized seed out of the subnormals; 2^-k goes into the
  // alpha of the Gc GEMM (exact powers of two: no rounding changes)
  const float qscale = (fabsf(seed) > 0.f && isfinite(seed)) ? exp2f(-floorf(log2f(fabsf(seed))) - 1.f) : 1.f;
  for (int e = blockIdx.y * 256 + threadIdx.x; e < n * n; e += NB * 256) {
    const int i = e / n, j = e - i * n;
    store_split_h(L.A[0], nn, e, load2_h(L.Y[0], nn, e) / norm_y);   // a = z / ||z||
    store_split_h(L.A[0], nn, 2 * nn + e, load2_h(L.Y[0], nn, 2 * nn + e) / norm_y);
    const float q0 = (i == j) ? seed * qscale : 0.f;
    store_split_h(L.Q[0], nn, e, q0);
    store_split_h(L.Q[0], nn, 2 * nn + e, q0);
  }
  if (blockIdx.y == 0 && threadIdx.x == 0) {
    L.scal[W2S_QSCALE] = qscale;
    *L.gc_alpha = 0.5f / qscale;
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
  const size_t nn = (size_t)n * n;
  const float inv_n = 1.f / L.npix;
  for (int e = blockIdx.y * 256 + threadIdx.x; e < n * n; e += NB * 256) {
    const int i = e / n, j = e - i * n;
    const float gs = load2(L.Gc, nn, e) + load2(L.Gc, nn, (size_t)j * n + i);
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

__global__ void sum_planes_kernel(float* __restrict__ dst, const float* __restrict__ pair, size_t nn) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nn; e += (size_t)gridDim.x * blockDim.x)
    dst[e] = pair[e] + pair[nn + e];
}

}  // namespace

int preload_w2_kernels() {
  cudaFuncAttributes fa;
#define STB_PRELOAD(k) STB_CUDA_CHECK(cudaFuncGetAttributes(&fa, reinterpret_cast<const void*>(k)))
  STB_PRELOAD(w2_chain_kernel); STB_PRELOAD(w2_cov_kernel); STB_PRELOAD(w2_ns_init_kernel);
  STB_PRELOAD(w2_target_finish_kernel); STB_PRELOAD(w2_fwd_finish_kernel); STB_PRELOAD(w2_bwd_finish_kernel);
  STB_PRELOAD(w2_gmu_kernel); STB_PRELOAD(sum_planes_kernel);
#undef STB_PRELOAD
  return STB_OK;
}

int W2Engine::read_matrix(float* dst, const float* pair, int n, cudaStream_t s) {
  sum_planes_kernel<<<64, 256, 0, s>>>(dst, pair, (size_t)n * n);
  STB_CUDA_CHECK(cudaGetLastError());
  return STB_OK;
}

// ================================================================================================ host engine
size_t W2Engine::layer_floats(int n) {
  // 4 planes (hi, lo, hi^T, lo^T): cov, M, X, Y[2], Z[2], T, A[2], Q[2], E, U, Gc, P, cov_t, Qf (18); single: Gs, srm_t, X1
  // (the fp16 pairs Y, Z, T, A, Q, E use half of their slot)
  return (size_t)(18 * 4 + 3) * n * n + 8 * (size_t)n + 64 + 2 * NRED + 1024;
}

size_t W2Engine::workspace_bytes() {
  size_t fl = 0;
  const int ns[5] = {64, 128, 256, 512, 512};
  for (int l = 0; l < 5; ++l) fl += layer_floats(ns[l]) + (size_t)ns[l] * ns[l] / 2 + 64;  // + bf16 Gs
  // + device copies of the layer table, tensor maps, problems and tiles
  return fl * 4 + (size_t(1) << 20);
}

namespace {
struct Builder {
  std::vector<W2Round>* rounds = nullptr;
  std::vector<CUtensorMap> maps;
  std::map<const float*, int> map_index;
  std::set<const float*> half_mats;   // matrices kept as fp16 plane pairs (4 planes of n*n halfs at the slot's start)
  CUtensorMap* d_maps = nullptr;
  int rc = 0;
  bool is_half(const float* m) const { return half_mats.count(m) != 0; }
  int maps_for(const float* m, int n) {
    auto it = map_index.find(m);
    if (it != map_index.end()) return it->second;
    const int idx = (int)maps.size();
    maps.resize(idx + 4);
    const size_t nn = (size_t)n * n;
    int r;
    if (is_half(m)) {
      const __half* h = reinterpret_cast<const __half*>(m);
      r = make_tmap_f16_2d(&maps[idx + 0], h, n, n, 2 * TKF, TM);               // 64 halfs = 128 bytes per box row
      if (!r) r = make_tmap_f16_2d(&maps[idx + 1], h + nn, n, n, 2 * TKF, TM);
      if (!r) r = make_tmap_f16_2d(&maps[idx + 2], h + 2 * nn, n, n, 2 * TKF, TN);
      if (!r) r = make_tmap_f16_2d(&maps[idx + 3], h + 3 * nn, n, n, 2 * TKF, TN);
    } else {
      r = make_tmap_f32_2d(&maps[idx + 0], m, n, n, TKF, TM);               // 128-row boxes: A loads, D stores
      if (!r) r = make_tmap_f32_2d(&maps[idx + 1], m + nn, n, n, TKF, TM);
      if (!r) r = make_tmap_f32_2d(&maps[idx + 2], m + 2 * nn, n, n, TKF, TN);  // 64-row boxes on the planes of the
      if (!r) r = make_tmap_f32_2d(&maps[idx + 3], m + 3 * nn, n, n, TKF, TN);  // transpose: B loads
    }
    if (r) rc = r;
    map_index[m] = idx;
    return idx;
  }
  void begin_round() { rounds->emplace_back(); rounds->back().n_tiles = 0; rounds->back().n_probs = 0; }
  // D = alpha op(A) B' + gamma I.  a_t: op(A) = A^T (its planes sit 2*n*n floats after A's).  b_t: B' = B^T, whose
  // K-major rows are B's own planes 0,1 -- registered under the key B - 2*n*n so that "+2" lands on them.
  // write_t: also emit the planes of D^T (D is later used as a right factor).
  void add(int n, float* D, const float* A, const float* B, float alpha, float gamma = 0.f, float* red_out = nullptr,
           int write_t = 1, int a_t = 0, int b_t = 0) {
    W2Round& R = rounds->back();
    if (R.n_probs >= W2_MAX_PROBS) { rc = STB_ERR_STATE; return; }
    const size_t nn = (size_t)n * n;
    TcProb& p = R.probs[R.n_probs];
    p.amap = d_maps + maps_for(a_t ? A + 2 * nn : A, n);
    p.bmap = d_maps + maps_for(b_t ? B - 2 * nn : B, n) + 2;
    p.dmap = d_maps + maps_for(D, n);
    p.D = D; p.red_out = red_out; p.n = n; p.alpha = alpha; p.gamma = gamma; p.write_t = write_t;
    p.in_half = is_half(A) ? 1 : 0;
    p.out_half = is_half(D) ? 1 : 0;
    if (is_half(A) != is_half(B) || ((a_t || b_t) && is_half(A))) { rc = STB_ERR_STATE; return; }
    for (int i = 0; i < (n + TM - 1) / TM; ++i)
      for (int j = 0; j < n / TN; ++j) {
        if (R.n_tiles >= W2_MAX_TILES) { rc = STB_ERR_STATE; return; }
        R.tiles[R.n_tiles++] = (uint32_t)R.n_probs << 16 | (uint32_t)i << 8 | (uint32_t)j;
      }
    ++R.n_probs;
  }
};
constexpr int MAX_MAPS = 5 * 32 * 4;
}  // namespace

int W2Engine::init(void* ws, size_t bytes, const int n_per_layer[5]) {
  STB_CHECK(bytes >= workspace_bytes(), STB_ERR_WORKSPACE, "W2 workspace too small");
  uint8_t* base = static_cast<uint8_t*>(ws);
  size_t off = 0;
  auto take = [&](size_t nbytes) { void* p = base + off; off += (nbytes + 255) & ~size_t(255); return p; };
  for (int l = 0; l < 5; ++l) {
    W2Layer& L = host_layers[l];
    const int n = n_per_layer[l];
    const size_t nn = (size_t)n * n * 4, pp = 4 * nn;  // single plane / (hi, lo, hi^T, lo^T)
    L.n = n;
    L.eps = 1e-4f;
    L.cov = (float*)take(pp); L.M = (float*)take(pp); L.X = (float*)take(pp);
    L.Y[0] = (float*)take(pp); L.Y[1] = (float*)take(pp); L.Z[0] = (float*)take(pp); L.Z[1] = (float*)take(pp);
    L.T = (float*)take(pp); L.A[0] = (float*)take(pp); L.A[1] = (float*)take(pp);
    L.Q[0] = (float*)take(pp); L.Q[1] = (float*)take(pp); L.E = (float*)take(pp);
    L.U = (float*)take(pp); L.Gc = (float*)take(pp); L.P = (float*)take(pp); L.cov_t = (float*)take(pp);
    L.Qf = (float*)take(pp);
    L.Gs = (float*)take(nn); L.srm_t = (float*)take(nn); L.X1 = (float*)take(nn); L.X23 = nullptr;
    L.S_raw = nullptr; L.sums = nullptr;  // bound per plan (stats buffer)
    L.mu = (float*)take(n * 4); L.mean_t = (float*)take(n * 4); L.gmu_bias = (float*)take(n * 4);
    L.scal = (float*)take(64 * 4);
    L.red = (float*)take(2 * NRED * 4);
    L.gs_bf16 = (bf16*)take((size_t)n * n * 2);
    L.weight = 0.f; L.npix = 1.f;
  }
  d_layers = (W2Layer*)take(sizeof(W2Layer) * 5);
  d_maps = (CUtensorMap*)take(sizeof(CUtensorMap) * MAX_MAPS);
  d_grid_counter = (unsigned*)take(256);
  d_trace = (unsigned long long*)take(W2_TRACE_WORDS * 8);

  // ---- build the round lists once (pointers are stable)
  Builder b;
  b.d_maps = d_maps;
  b.rounds = &rounds;
  rounds.clear();
  for (int l = 0; l < 5; ++l) {  // the normalised iterates of both iterations live as fp16 plane pairs
    W2Layer& L = host_layers[l];
    for (float* m : {L.Y[0], L.Y[1], L.Z[0], L.Z[1], L.T, L.A[0], L.A[1], L.Q[0], L.Q[1], L.E}) b.half_mats.insert(m);
  }
  auto begin_round = [&]() { b.begin_round(); };
  auto end_round = [&]() {};
  auto ns_rounds = [&]() {  // 12 x { T = 1.5 I - 0.5 Z Y ; Y' = Y T, Z' = T Z }, result ends in Y[0]
    for (int it = 0; it < 12; ++it) {
      const int s = it & 1, d = s ^ 1;
      begin_round();
      for (int l = 4; l >= 0; --l) { W2Layer& L = host_layers[l]; b.add(L.n, L.T, L.Z[s], L.Y[s], -0.5f, 1.5f); }
      end_round();
      begin_round();
      for (int l = 4; l >= 0; --l) {  // largest matrices first: the two CTAs beyond 148 are short ones
        W2Layer& L = host_layers[l];
        b.add(L.n, L.Y[d], L.Y[s], L.T, 1.f, 0.f, it == 11 ? L.red : nullptr);
        if (it < 11) b.add(L.n, L.Z[d], L.T, L.Z[s], 1.f);  // Z is dead after the last Y
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
  for (int l = 4; l >= 0; --l) { W2Layer& L = host_layers[l]; b.add(L.n, L.X, L.P, L.cov, 1.f, 0.f, nullptr, 0); }
  end_round();
  begin_round();
  for (int l = 4; l >= 0; --l) { W2Layer& L = host_layers[l]; b.add(L.n, L.M, L.X, L.P, 1.f, 0.f, L.red, 1); }  // M^T planes: ns_init reads them coalesced
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
    for (int l = 4; l >= 0; --l) { W2Layer& L = host_layers[l]; b.add(L.n, L.E, L.A[s], L.A[s], -1.f, 3.f); }
    end_round();
    begin_round();
    for (int l = 4; l >= 0; --l) {
      W2Layer& L = host_layers[l];
      b.add(L.n, it == 11 ? L.Qf : L.Q[d], L.Q[s], L.E, 0.5f);   // the final q leaves the iteration as a TF32 pair
      if (it < 11) b.add(L.n, L.A[d], L.A[s], L.E, 0.5f);
    }
    end_round();
  }
  // after 12 its q (times the layer's power-of-two q scale) is in Qf.  U = P^T q ; Gc = alpha U P^T + (w/C) I with
  // alpha = 0.5 / qscale patched on the device by w2_fwd_finish_kernel, gamma per layer in upload_layers
  begin_round();
  for (int l = 4; l >= 0; --l) { W2Layer& L = host_layers[l]; b.add(L.n, L.U, L.P, L.Qf, 1.f, 0.f, nullptr, 0, 1, 0); }
  end_round();
  begin_round();
  gc_round = (int)rounds.size() - 1;
  for (int l = 4; l >= 0; --l) { W2Layer& L = host_layers[l]; b.add(L.n, L.Gc, L.U, L.P, 0.5f, 0.f, nullptr, 0, 0, 1); }
  end_round();
  r_bwd_end = (int)rounds.size();
  STB_CHECK(b.rc == 0, STB_ERR_CUDA, "W2 round construction failed (%d): %s", b.rc, last_error_string().c_str());
  STB_CHECK((int)b.maps.size() <= MAX_MAPS, STB_ERR_WORKSPACE, "W2 tensor map table overflow (%zu)", b.maps.size());

  d_rounds = (W2Round*)take(sizeof(W2Round) * rounds.size());
  STB_CHECK(off <= bytes, STB_ERR_WORKSPACE, "W2 workspace overflow (%zu > %zu)", off, bytes);
  STB_CUDA_CHECK(cudaMemcpy(d_rounds, rounds.data(), sizeof(W2Round) * rounds.size(), cudaMemcpyHostToDevice));
  for (int l = 0; l < 5; ++l) host_layers[l].gc_alpha = &d_rounds[gc_round].probs[4 - l].alpha;
  STB_CUDA_CHECK(cudaMemcpy(d_maps, b.maps.data(), sizeof(CUtensorMap) * b.maps.size(), cudaMemcpyHostToDevice));
  STB_CUDA_CHECK(cudaMemcpy(d_layers, host_layers, sizeof(W2Layer) * 5, cudaMemcpyHostToDevice));
  return STB_OK;
}

int W2Engine::upload_layers(cudaStream_t s) {
  // layer weights enter the Gc round through gamma = w / C (rounds travel as kernel parameters)
  for (int l = 0; l < 5; ++l) rounds[gc_round].probs[4 - l].gamma = host_layers[l].weight / host_layers[l].n;
  STB_CUDA_CHECK(cudaMemcpyAsync(d_rounds + gc_round, &rounds[gc_round], sizeof(W2Round), cudaMemcpyHostToDevice, s));
  STB_CUDA_CHECK(cudaMemcpyAsync(d_layers, host_layers, sizeof(W2Layer) * 5, cudaMemcpyHostToDevice, s));
  return STB_OK;
}

int W2Engine::run_rounds(int r0, int r1, cudaStream_t s) {
  if (r1 <= r0) return STB_OK;
  STB_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(w2_chain_kernel), T_SMEM_BYTES));
  // STB_W2_CHAIN=0: one launch per round (no grid barrier, no cooperative launch) -- diagnostic / fallback
  static const bool chain = [] { const char* e = getenv("STB_W2_CHAIN"); return !(e && e[0] == '0'); }();
  const int sms = num_sms();
  if (!chain) {
    for (int r = r0; r < r1; ++r) {
      const int grid = rounds[r].n_tiles < sms ? rounds[r].n_tiles : sms;
      w2_chain_kernel<<<grid, T_THREADS, T_SMEM_BYTES, s>>>(d_rounds, r, r + 1, d_grid_counter, nullptr);
    }
    STB_CUDA_CHECK(cudaGetLastError());
    return STB_OK;
  }
  int max_tiles = 1;
  for (int r = r0; r < r1; ++r) max_tiles = rounds[r].n_tiles > max_tiles ? rounds[r].n_tiles : max_tiles;
  STB_CUDA_CHECK(cudaMemsetAsync(d_grid_counter, 0, sizeof(unsigned), s));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(max_tiles < sms ? max_tiles : sms);   // one CTA per SM: all co-resident (cooperative launch)
  cfg.blockDim = dim3(T_THREADS);
  cfg.dynamicSmemBytes = T_SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  const W2Round* dr = d_rounds;
  unsigned* ctr = d_grid_counter;
  static const bool tracing = [] { const char* e = getenv("STB_W2_TRACE"); return e && e[0] == '1'; }();
  unsigned long long* tr = tracing ? d_trace : nullptr;
  STB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, w2_chain_kernel, dr, r0, r1, ctr, tr));
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
