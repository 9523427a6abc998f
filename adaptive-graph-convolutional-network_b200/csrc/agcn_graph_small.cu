// Per-graph kernels of the SGC-LL layer for graphs with n <= AGCN_SMALL_MAX nodes: every n x n matrix
// of one graph lives in shared memory; one CTA per (graph, feature chunk) for the Chebyshev
// recurrences, one CTA per graph for the Laplacian construction and its gradient.
//
//   graph_build_laplacian : graphconv.py:163-216 / graphconv_reslap.py:136-195
//   graph_chebyshev_fwd   : graphconv.py:221-236
//   graph_recurrence_bwd  : reverse-mode of graphconv.py:221-236 (dX through T_k, dL_all)
//   graph_laplacian_bwd   : reverse-mode of graphconv.py:212-216 / graphconv_reslap.py:185-195 and,
//                           with metric_grad == FULL, of the metric block (paper semantics)
#include <algorithm>
#include <mutex>
#include <set>

#include "agcn_internal.cuh"

namespace agcn {

bool literal_shortcut(int variant, int lap_mode) {
  return variant == AGCN_VARIANT_SGC_LL && lap_mode == AGCN_LAP_REFERENCE_LITERAL;
}

__host__ __device__ __forceinline__ int round4(int x) { return (x + 3) & ~3; }
// pitch of an n x n shared-memory matrix: multiple of 4 floats, == 4 (mod 8)
__host__ __device__ __forceinline__ int lap_pitch(int n4) { return (n4 % 8 == 0) ? n4 + 4 : n4; }

struct PlanPtrs {
  const int32_t* n_nodes;
  const int32_t* node_off;
  const int32_t* order;
  const int64_t* lap_off;
};

static PlanPtrs plan_ptrs(const agcn_plan* p) { return PlanPtrs{p->d_n, p->d_node_off, p->d_order, p->d_lap_off}; }

// ------------------------------------------------------------------------------------------------
// shared-memory micro kernels
// ------------------------------------------------------------------------------------------------
// acc[q][t] = sum_j M(r0+q, j) * In[j][c0+t]   (TRANS: M(i,j) = sM[j][i]).  sM is n4 x pl, zero padded;
// In is n4 x pf, zero padded.
template <bool TRANS>
__device__ __forceinline__ void mm_tile(const float* __restrict__ sM, int pl, int n4, const float* __restrict__ sIn,
                                        int pf, int r0, int c0, float acc[4][4]) {
  for (int j = 0; j < n4; j += 4) {
    float4 b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) b[u] = *reinterpret_cast<const float4*>(&sIn[(j + u) * pf + c0]);
    if (!TRANS) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 a = *reinterpret_cast<const float4*>(&sM[(r0 + q) * pl + j]);
        acc[q][0] += a.x * b[0].x + a.y * b[1].x + a.z * b[2].x + a.w * b[3].x;
        acc[q][1] += a.x * b[0].y + a.y * b[1].y + a.z * b[2].y + a.w * b[3].y;
        acc[q][2] += a.x * b[0].z + a.y * b[1].z + a.z * b[2].z + a.w * b[3].z;
        acc[q][3] += a.x * b[0].w + a.y * b[1].w + a.z * b[2].w + a.w * b[3].w;
      }
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 a = *reinterpret_cast<const float4*>(&sM[(j + u) * pl + r0]);
        acc[0][0] += a.x * b[u].x; acc[0][1] += a.x * b[u].y; acc[0][2] += a.x * b[u].z; acc[0][3] += a.x * b[u].w;
        acc[1][0] += a.y * b[u].x; acc[1][1] += a.y * b[u].y; acc[1][2] += a.y * b[u].z; acc[1][3] += a.y * b[u].w;
        acc[2][0] += a.z * b[u].x; acc[2][1] += a.z * b[u].y; acc[2][2] += a.z * b[u].z; acc[2][3] += a.z * b[u].w;
        acc[3][0] += a.w * b[u].x; acc[3][1] += a.w * b[u].y; acc[3][2] += a.w * b[u].z; acc[3][3] += a.w * b[u].w;
      }
    }
  }
}

constexpr int NT_COLS = (AGCN_SMALL_MAX + 31) / 32;  // columns per lane in the "NT" products

// Row block r0..r0+3 (one warp) against all rows j of sBm: lane owns j = lane + 32*u.
//   DIST = false: acc[q][u] = sum_c sAm[r0+q][c] * sBm[j][c]
//   DIST = true : acc[q][u] = sum_c (sAm[r0+q][c] - sBm[j][c])^2
// pitch pc must be a multiple of 4 with pc/4 odd (conflict-free float4 reads).
template <bool DIST>
__device__ __forceinline__ void nt_rowblock(const float* __restrict__ sAm, const float* __restrict__ sBm, int pc, int w,
                                            int n4, int r0, int lane, float acc[4][NT_COLS]) {
  for (int c = 0; c < w; c += 4) {
    float4 a[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) a[q] = *reinterpret_cast<const float4*>(&sAm[(r0 + q) * pc + c]);
#pragma unroll
    for (int u = 0; u < NT_COLS; ++u) {
      const int j = lane + 32 * u;
      if (j < n4) {
        const float4 b = *reinterpret_cast<const float4*>(&sBm[j * pc + c]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (DIST) {
            const float dx = a[q].x - b.x, dy = a[q].y - b.y, dz = a[q].z - b.z, dw = a[q].w - b.w;
            acc[q][u] += dx * dx + dy * dy + dz * dz + dw * dw;
          } else {
            acc[q][u] += a[q].x * b.x + a[q].y * b.y + a[q].z * b.z + a[q].w * b.w;
          }
        }
      }
    }
  }
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  // red: >= 33 floats of shared memory.  All threads must call.
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    float s = (lane < nw) ? red[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[32] = s;
  }
  __syncthreads();
  return red[32];
}

__device__ __forceinline__ float leaky(float x, float alpha) { return fmaxf(x, 0.f) - alpha * fmaxf(-x, 0.f); }
__device__ __forceinline__ float leaky_grad(float x, float alpha) { return x > 0.f ? 1.f : (x < 0.f ? alpha : 0.f); }

// ---- asynchronous global -> shared copies (LDGSTS): every element of a tile is in flight at once, so a
// CTA pays one memory latency per tile instead of one per loop iteration.
__device__ __forceinline__ void cp_async4(float* sdst, const float* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(sdst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(float* sdst, const float* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(sdst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// n x n packed matrix -> n4 x pl zero-padded shared tile (async; finish with cp_async_wait_all + barrier)
__device__ __forceinline__ void load_square(float* sM, int pl, int n, int n4, const float* __restrict__ src) {
  for (int idx = threadIdx.x; idx < n4 * pl; idx += blockDim.x) {
    const int i = idx / pl, j = idx - i * pl;
    if (src != nullptr && i < n && j < n)
      cp_async4(&sM[idx], src + i * n + j);
    else
      sM[idx] = 0.f;
  }
}

__device__ __forceinline__ void add_identity(float* sM, int pl, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) sM[i * pl + i] += 1.f;
}

// rows [row0, row0+n) x cols [f0, f0+fc) of a [R, ld] matrix -> n4 x pf zero padded tile (async)
__device__ __forceinline__ void load_chunk(float* sT, int pf, int n, int n4, const float* __restrict__ src, int ld,
                                           int64_t row0, int f0, int fc) {
  const bool vec = ((ld & 3) == 0) && ((f0 & 3) == 0) && ((fc & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  if (vec) {
    const int pf4 = pf >> 2, fc4 = fc >> 2;
    for (int idx = threadIdx.x; idx < n4 * pf4; idx += blockDim.x) {
      const int i = idx / pf4, c4 = idx - i * pf4;
      float* d = &sT[i * pf + c4 * 4];
      if (i < n && c4 < fc4)
        cp_async16(d, src + (row0 + i) * ld + f0 + c4 * 4);
      else
        *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
    for (int idx = threadIdx.x; idx < n4 * pf; idx += blockDim.x) {
      const int i = idx / pf, c = idx - i * pf;
      if (i < n && c < fc)
        cp_async4(&sT[idx], src + (row0 + i) * ld + f0 + c);
      else
        sT[idx] = 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Chebyshev recurrence, forward
// ------------------------------------------------------------------------------------------------
struct ChebArgs {
  PlanPtrs pp;
  int order_start, chunks, FC, F, K;
  const float* X;
  const float* L;  // packed: Lint (add_identity) or L_all
  int add_identity;
  float* T;
  int64_t slice;
};

__global__ void cheb_fwd_kernel(ChebArgs p) {
  extern __shared__ __align__(16) float smem[];
  const int gi = blockIdx.x / p.chunks, ch = blockIdx.x - gi * p.chunks;
  const int g = p.pp.order[p.order_start + gi];
  const int n = p.pp.n_nodes[g], n4 = round4(n), pl = lap_pitch(n4);
  const int f0 = ch * p.FC;
  if (f0 >= p.F) return;
  const int fc = min(p.FC, p.F - f0), pf = p.FC + 4;
  float* sL = smem;
  float* sA = sL + n4 * pl;
  float* sB = sA + n4 * pf;
  const int64_t row0 = p.pp.node_off[g];
  load_square(sL, pl, n, n4, p.L + p.pp.lap_off[g]);
  load_chunk(sA, pf, n, n4, p.X, p.F, row0, f0, fc);
  for (int idx = threadIdx.x; idx < n4 * pf; idx += blockDim.x) sB[idx] = 0.f;
  cp_async_wait_all();
  __syncthreads();
  if (p.add_identity) {
    add_identity(sL, pl, n);
    __syncthreads();
  }
  const int tc = (fc + 3) / 4, tiles = (n4 / 4) * tc;
  float* src = sA;
  float* dst = sB;
  for (int k = 1; k < p.K; ++k) {
    float* __restrict__ Tk = p.T + (int64_t)(k - 1) * p.slice;
    for (int t = threadIdx.x; t < tiles; t += blockDim.x) {
      const int r0 = (t / tc) * 4, c0 = (t % tc) * 4;
      float acc[4][4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[q][u] = 0.f;
      mm_tile<false>(sL, pl, n4, src, pf, r0, c0, acc);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = r0 + q;
        if (i >= n) continue;
        float* d = &dst[i * pf + c0];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float o = (k == 1) ? acc[q][u] : 2.f * acc[q][u] - d[u];  // graphconv.py:231,234
          d[u] = (c0 + u < fc) ? o : 0.f;
          if (c0 + u < fc) Tk[(row0 + i) * p.F + f0 + c0 + u] = o;
        }
      }
    }
    __syncthreads();
    float* tmp = src; src = dst; dst = tmp;
  }
}

// ------------------------------------------------------------------------------------------------
// Chebyshev recurrence, backward
//   U_{K-1} = G_{K-1};  U_j = G_j + c_{j+1} L^T U_{j+1} - [j+2 <= K-1] U_{j+2},  c_1 = 1, c_k = 2 (k >= 2)
//   dX = U_0;  dL = sum_{k>=1} c_k U_k T_{k-1}^T (+ dLall_in)
// ------------------------------------------------------------------------------------------------
struct RecurArgs {
  PlanPtrs pp;
  int order_start, chunks, FC, F, K;
  const float* L;
  int add_identity;
  const float* G;
  int64_t gslice;
  float* dX;
  int need_dL;  // the `chunks` CTAs of a graph take the feature chunks ch, ch + chunks, ... and write their partial
                // dL (+ dLall_in for ch == 0) to dL + ch * dl_stride; graph_laplacian_bwd adds the partials up
  const float* X;
  const float* T;
  int64_t tslice;
  const float* dLall_in;
  float* dL;
  int64_t dl_stride;
};

__global__ void recur_bwd_kernel(RecurArgs p) {
  extern __shared__ __align__(16) float smem[];
  const int gi = blockIdx.x / p.chunks, ch = blockIdx.x - gi * p.chunks;
  const int g = p.pp.order[p.order_start + gi];
  const int n = p.pp.n_nodes[g], n4 = round4(n), pl = lap_pitch(n4);
  const int pf = p.FC + 4;
  float* sL = smem;
  float* sA = sL + n4 * pl;
  float* sB = sA + n4 * pf;
  float* sT = sB + n4 * pf;   // need_dL only
  float* sdL = sT + n4 * pf;  // need_dL only
  const int64_t row0 = p.pp.node_off[g];
  const int64_t loff = p.pp.lap_off[g];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  load_square(sL, pl, n, n4, p.L + loff);
  if (p.need_dL) load_square(sdL, pl, n, n4, (p.dLall_in && ch == 0) ? p.dLall_in + loff : nullptr);
  cp_async_wait_all();
  __syncthreads();
  if (p.add_identity) add_identity(sL, pl, n);
  const int f_begin = ch * p.FC;
  const int f_end = p.need_dL ? p.F : min(p.F, f_begin + p.FC);
  const int f_step = p.need_dL ? p.chunks * p.FC : p.FC;
  for (int f0 = f_begin; f0 < f_end; f0 += f_step) {
    const int fc = min(p.FC, p.F - f0);
    const int tc = (fc + 3) / 4, tiles = (n4 / 4) * tc;
    __syncthreads();
    load_chunk(sB, pf, n, n4, p.G + (int64_t)(p.K - 1) * p.gslice, p.F, row0, f0, fc);  // U_{K-1}
    for (int idx = threadIdx.x; idx < n4 * pf; idx += blockDim.x) sA[idx] = 0.f;
    float* sU1 = sB;  // U_{j+1}
    float* sU2 = sA;  // U_{j+2} -> receives U_j
    for (int j = p.K - 2; j >= 0; --j) {
      const float cmul = (j + 1 >= 2) ? 2.f : 1.f;
      if (p.need_dL) {
        // dL += c_{j+1} U_{j+1} T_j^T
        const float* Tj = (j == 0) ? p.X : p.T + (int64_t)(j - 1) * p.tslice;
        load_chunk(sT, pf, n, n4, Tj, p.F, row0, f0, fc);
        cp_async_wait_all();
        __syncthreads();
        for (int rb = wid; rb < n4 / 4; rb += nw) {
          float acc[4][NT_COLS];
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int u = 0; u < NT_COLS; ++u) acc[q][u] = 0.f;
          nt_rowblock<false>(sU1, sT, pf, round4(fc), n4, rb * 4, lane, acc);
#pragma unroll
          for (int u = 0; u < NT_COLS; ++u) {
            const int jj = lane + 32 * u;
            if (jj < n) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (rb * 4 + q < n) sdL[(rb * 4 + q) * pl + jj] += cmul * acc[q][u];
            }
          }
        }
      } else {
        cp_async_wait_all();
        __syncthreads();
      }
      const float* __restrict__ Gj = p.G + (int64_t)j * p.gslice;
      const bool has_u2 = (j + 2 <= p.K - 1);
      for (int t = threadIdx.x; t < tiles; t += blockDim.x) {
        const int r0 = (t / tc) * 4, c0 = (t % tc) * 4;
        float acc[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[q][u] = 0.f;
        mm_tile<true>(sL, pl, n4, sU1, pf, r0, c0, acc);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = r0 + q;
          if (i >= n) continue;
          float* d = &sU2[i * pf + c0];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (c0 + u < fc) {
              float o = Gj[(row0 + i) * p.F + f0 + c0 + u] + cmul * acc[q][u];
              if (has_u2) o -= d[u];
              d[u] = o;
              if (j == 0) p.dX[(row0 + i) * p.F + f0 + c0 + u] = o;
            } else {
              d[u] = 0.f;
            }
          }
        }
      }
      __syncthreads();
      float* tmp = sU1; sU1 = sU2; sU2 = tmp;
    }
  }
  if (p.need_dL) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
      const int i = idx / n, j = idx - i * n;
      p.dL[(int64_t)ch * p.dl_stride + loff + idx] = sdL[i * pl + j];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Laplacian construction (one CTA per graph)
// ------------------------------------------------------------------------------------------------
struct BuildArgs {
  PlanPtrs pp;
  int order_start, F, FCD;
  int variant, lap_mode, need_W;
  const float* XW;
  const float* Lint;
  const float* Lprev;
  const float* alpha;
  const float* beta;
  float* Lall;
  float* Lall2;  // second destination (user output) or null
  float* resL;
  float* resW;
  float* dist;
  float* dis;
  float* stats;
};

__global__ void build_lap_kernel(BuildArgs p) {
  extern __shared__ __align__(16) float smem[];
  const int g = p.pp.order[p.order_start + blockIdx.x];
  const int n = p.pp.n_nodes[g], n4 = round4(n), pl = lap_pitch(n4);
  const int pc = p.FCD + 4;
  float* sM = smem;
  float* sS = sM + n4 * pl;
  float* sdis = sS + n4 * pc;
  float* red = sdis + n4;
  const int64_t row0 = p.pp.node_off[g];
  const int64_t loff = p.pp.lap_off[g];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float alpha = p.alpha[0];
  const bool paper = (p.lap_mode == AGCN_LAP_PAPER);

  if (p.need_W) {
    // squared distances by direct differences, fp32 (graphconv.py:174-176: np.linalg.norm(u - v))
    for (int idx = threadIdx.x; idx < n4 * pl; idx += blockDim.x) sM[idx] = 0.f;
    for (int f0 = 0; f0 < p.F; f0 += p.FCD) {
      const int fc = min(p.FCD, p.F - f0);
      __syncthreads();
      load_chunk(sS, pc, n, n4, p.XW, p.F, row0, f0, fc);
      cp_async_wait_all();
      __syncthreads();
      for (int rb = wid; rb < n4 / 4; rb += nw) {
        float acc[4][NT_COLS];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int u = 0; u < NT_COLS; ++u) acc[q][u] = 0.f;
        nt_rowblock<true>(sS, sS, pc, round4(fc), n4, rb * 4, lane, acc);
#pragma unroll
        for (int u = 0; u < NT_COLS; ++u) {
          const int j = lane + 32 * u;
          if (j < n4) {
#pragma unroll
            for (int q = 0; q < 4; ++q) sM[(rb * 4 + q) * pl + j] += acc[q][u];
          }
        }
      }
    }
    __syncthreads();
    // W_ij = exp(-dist), W_ii = 0 (graphconv.py:171-178)
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
      const int i = idx / n, j = idx - i * n;
      const float d = sqrtf(sM[i * pl + j]);
      const float w = (i == j) ? 0.f : expf(-d);
      sM[i * pl + j] = w;
      if (p.resW) p.resW[loff + idx] = w;
      if (p.dist) p.dist[loff + idx] = (i == j) ? 0.f : d;
    }
    __syncthreads();
  }
  float normR2;
  if (paper) {
    // d = W.sum(axis=0) (+eps), d^-1/2 (graphconv.py:195-197); d == 0 -> 0
    for (int j = threadIdx.x; j < n4; j += blockDim.x) {
      float d = 0.f;
      if (j < n)
        for (int i = 0; i < n; ++i) d += sM[i * pl + j];
      const float v = (d > AGCN_DEGREE_FLOOR) ? 1.0f / sqrtf(d) : 0.f;
      sdis[j] = v;
      if (j < n && p.dis) p.dis[row0 + j] = v;
    }
    __syncthreads();
    float part = 0.f;
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
      const int i = idx / n, j = idx - i * n;
      const float r = ((i == j) ? 1.f : 0.f) - (sdis[i] * sM[i * pl + j]) * sdis[j];  // I - D W D
      sM[i * pl + j] = r;
      part += r * r;
    }
    normR2 = block_sum(part, red);
  } else {
    __syncthreads();
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
      const int i = idx / n, j = idx - i * n;
      sM[i * pl + j] = (i == j) ? 1.f : 0.f;  // D * W * D elementwise == 0 (graphconv.py:198-200)
    }
    normR2 = (float)n;
    __syncthreads();
  }
  // tf.clip_by_average_norm (graphconv.py:212) / tf.clip_by_norm (graphconv_reslap.py:185)
  const float inv1 = (normR2 > 0.f) ? rsqrtf(normR2) : INFINITY;
  const float s1 = (p.variant == AGCN_VARIANT_SGC_LL) ? fminf(inv1 * (float)(n * n), 1.f) : fminf(inv1, 1.f);
  const bool reslap = (p.variant == AGCN_VARIANT_SGC_LL_RESLAP);
  const float beta = (reslap && p.Lprev) ? p.beta[0] : 0.f;
  float part = 0.f;
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int i = idx / n, j = idx - i * n;
    const float rl = leaky(sM[i * pl + j] * s1, alpha);  // graphconv.py:213
    if (p.resL) p.resL[loff + idx] = rl;
    float z = rl + p.Lint[loff + idx];                   // graphconv.py:216
    if (reslap && p.Lprev) z += p.Lprev[loff + idx] * beta;  // graphconv_reslap.py:190
    sM[i * pl + j] = z;
    part += z * z;
  }
  float s2 = 1.f, normZ2 = 0.f;
  if (reslap) {
    normZ2 = block_sum(part, red);
    const float inv2 = (normZ2 > 0.f) ? rsqrtf(normZ2) : INFINITY;
    s2 = fminf(inv2, 1.f);  // graphconv_reslap.py:194
  }
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int i = idx / n, j = idx - i * n;
    float v = sM[i * pl + j];
    if (reslap) v = leaky(v * s2, alpha);  // graphconv_reslap.py:195
    if (p.Lall) p.Lall[loff + idx] = v;
    if (p.Lall2) p.Lall2[loff + idx] = v;
  }
  if (threadIdx.x == 0 && p.stats) {
    p.stats[4 * g + 0] = s1;
    p.stats[4 * g + 1] = s2;
    p.stats[4 * g + 2] = normR2;
    p.stats[4 * g + 3] = normZ2;
  }
}

// ------------------------------------------------------------------------------------------------
// Laplacian chain, backward (one CTA per graph)
// ------------------------------------------------------------------------------------------------
struct LapBwdArgs {
  PlanPtrs pp;
  int order_start, F, FCD;
  int variant, lap_mode, metric_full;
  const float* XW;
  const float* Lint;
  const float* Lprev;
  const float* alpha;
  const float* beta;
  const float* dist;
  const float* dis;
  const float* stats;
  const float* dL;       // dl_parts partial sums, dl_stride elements apart (graph_recurrence_bwd)
  int dl_parts;
  int64_t dl_stride;
  float* dLprev;
  float* dXW;
  float* dalpha_part;
  float* dbeta_part;
};

__global__ void lap_bwd_kernel(LapBwdArgs p) {
  extern __shared__ __align__(16) float smem[];
  const int g = p.pp.order[p.order_start + blockIdx.x];
  const int n = p.pp.n_nodes[g], n4 = round4(n), pl = lap_pitch(n4);
  const int pc = p.FCD + 4;
  const bool paper = (p.lap_mode == AGCN_LAP_PAPER);
  const bool reslap = (p.variant == AGCN_VARIANT_SGC_LL_RESLAP);
  const bool full = paper && p.metric_full;
  float* sG = smem;                               // gradient matrix
  float* sW = sG + n4 * pl;                       // similarity matrix (paper)
  float* sS = sW + (paper ? n4 * pl : 0);         // feature chunk of XW (full)
  float* sdis = sS + (full ? n4 * pc : 0);
  float* svec = sdis + n4;                        // dd / rowsum(C)
  float* red = svec + n4;
  const int64_t row0 = p.pp.node_off[g];
  const int64_t loff = p.pp.lap_off[g];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float alpha = p.alpha[0];
  const float beta = (reslap && p.Lprev) ? p.beta[0] : 0.f;
  const float s1 = p.stats[4 * g + 0], s2 = p.stats[4 * g + 1];
  const float normR2 = p.stats[4 * g + 2], normZ2 = p.stats[4 * g + 3];
  const float cavg = (p.variant == AGCN_VARIANT_SGC_LL) ? (float)(n * n) : 1.f;
  const bool clipped1 = s1 < 1.f, clipped2 = s2 < 1.f;

  for (int j = threadIdx.x; j < n4; j += blockDim.x) sdis[j] = (paper && j < n) ? p.dis[row0 + j] : 0.f;
  for (int idx = threadIdx.x; idx < n4 * pl; idx += blockDim.x) {
    sG[idx] = 0.f;
    if (paper) sW[idx] = 0.f;
  }
  __syncthreads();
  if (paper) {
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
      const int i = idx / n, j = idx - i * n;
      sW[i * pl + j] = (i == j) ? 0.f : expf(-p.dist[loff + idx]);
    }
    __syncthreads();
  }
  auto Rval = [&](int i, int j) -> float {
    const float eye = (i == j) ? 1.f : 0.f;
    return paper ? eye - (sdis[i] * sW[i * pl + j]) * sdis[j] : eye;
  };
  float acc_alpha = 0.f, acc_beta = 0.f;
  auto dL_at = [&](int64_t e) -> float {   // fixed summation order: deterministic
    float v = p.dL[e];
    for (int c = 1; c < p.dl_parts; ++c) v += p.dL[(int64_t)c * p.dl_stride + e];
    return v;
  };
  if (reslap) {
    // L_all = leaky(s2 * Z): gradient w.r.t. v = s2 Z, and <gv, Z>
    float ip2 = 0.f;
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
      const int i = idx / n, j = idx - i * n;
      float z = leaky(Rval(i, j) * s1, alpha) + p.Lint[loff + idx];
      if (p.Lprev) z += p.Lprev[loff + idx] * beta;
      const float v = z * s2;
      const float gd = dL_at(loff + idx);
      acc_alpha -= gd * fmaxf(-v, 0.f);
      const float gv = gd * leaky_grad(v, alpha);
      ip2 += gv * z;
      sG[i * pl + j] = gv;
    }
    ip2 = block_sum(ip2, red);
    const float k2 = clipped2 ? s2 * s2 * s2 * ip2 : 0.f;
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
      const int i = idx / n, j = idx - i * n;
      float gz = sG[i * pl + j];
      if (clipped2) {
        float z = leaky(Rval(i, j) * s1, alpha) + p.Lint[loff + idx];
        if (p.Lprev) z += p.Lprev[loff + idx] * beta;
        gz = s2 * gz - z * k2;
      }
      if (p.Lprev) {
        acc_beta += gz * p.Lprev[loff + idx];
        if (p.dLprev) p.dLprev[loff + idx] = beta * gz;
      }
      sG[i * pl + j] = gz;  // = d res_L'
    }
  } else {
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
      const int i = idx / n, j = idx - i * n;
      sG[i * pl + j] = dL_at(loff + idx);
    }
  }
  // res_L' = leaky(s1 R)
  float ip1 = 0.f;
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int i = idx / n, j = idx - i * n;
    const float r = Rval(i, j);
    const float u = r * s1;
    const float gd = sG[i * pl + j];
    acc_alpha -= gd * fmaxf(-u, 0.f);
    const float gu = gd * leaky_grad(u, alpha);
    ip1 += gu * r;
    sG[i * pl + j] = gu;
  }
  acc_alpha = block_sum(acc_alpha, red);
  acc_beta = block_sum(acc_beta, red);
  if (threadIdx.x == 0) {
    p.dalpha_part[g] = acc_alpha;
    if (p.dbeta_part) p.dbeta_part[g] = acc_beta;
  }
  if (!full) return;

  // ---- metric block (paper semantics, differentiable)
  ip1 = block_sum(ip1, red);
  // s1 = c * (sum R^2)^-1/2 when clipped: dR = s1 gu - R (s1^3 / c^2) <gu, R>
  const float k1 = clipped1 ? (s1 * s1 * s1 / (cavg * cavg)) * ip1 : 0.f;
  (void)normR2; (void)normZ2;
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int i = idx / n, j = idx - i * n;
    float gr = sG[i * pl + j];
    if (clipped1) gr = s1 * gr - Rval(i, j) * k1;
    sG[i * pl + j] = gr;  // dR
  }
  __syncthreads();
  // R_ij = delta_ij - dis_i W_ij dis_j ; d_j = sum_i W_ij ; dis = d^-1/2
  // ddis_m = -sum_j dR_mj W_mj dis_j - sum_j dR_jm W_jm dis_j ; dd_m = -1/2 dis_m^3 ddis_m
  for (int m = wid; m < n; m += nw) {
    float a = 0.f;
    for (int j = lane; j < n; j += 32) a += sG[m * pl + j] * sW[m * pl + j] * sdis[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) svec[m] = a;
  }
  __syncthreads();
  for (int m = threadIdx.x; m < n; m += blockDim.x) {
    float b = 0.f;
    for (int j = 0; j < n; ++j) b += sG[j * pl + m] * sW[j * pl + m] * sdis[j];
    const float ddis = -(svec[m] + b);
    const float dm = sdis[m];
    svec[m] = -0.5f * dm * dm * dm * ddis;  // dd_m (0 when d_m == 0 because dis_m == 0)
  }
  __syncthreads();
  // C_ij = C_ji = (ddist_ij + ddist_ji) / dist_ij
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int i = idx / n, j = idx - i * n;
    if (i < j) {
      const float dst = p.dist[loff + idx];
      const float w = sW[i * pl + j];
      const float dw_ij = -sG[i * pl + j] * sdis[i] * sdis[j] + svec[j];
      const float dw_ji = -sG[j * pl + i] * sdis[i] * sdis[j] + svec[i];
      const float dd = -w * (dw_ij + dw_ji);
      const float c = (dst > 0.f) ? dd / dst : 0.f;  // sub-gradient 0 at exact duplicates
      sG[i * pl + j] = c;
      sG[j * pl + i] = c;
    } else if (i == j) {
      sG[i * pl + i] = 0.f;
    }
  }
  __syncthreads();
  for (int m = wid; m < n; m += nw) {
    float a = 0.f;
    for (int j = lane; j < n; j += 32) a += sG[m * pl + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) svec[m] = a;
  }
  // dXW_i = rowsum(C)_i xw_i - (C XW)_i
  for (int f0 = 0; f0 < p.F; f0 += p.FCD) {
    const int fc = min(p.FCD, p.F - f0);
    __syncthreads();
    load_chunk(sS, pc, n, n4, p.XW, p.F, row0, f0, fc);
    cp_async_wait_all();
    __syncthreads();
    const int tc = (fc + 3) / 4, tiles = (n4 / 4) * tc;
    for (int t = threadIdx.x; t < tiles; t += blockDim.x) {
      const int r0 = (t / tc) * 4, c0 = (t % tc) * 4;
      float acc[4][4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[q][u] = 0.f;
      mm_tile<false>(sG, pl, n4, sS, pc, r0, c0, acc);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = r0 + q;
        if (i >= n) continue;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (c0 + u < fc) p.dXW[(row0 + i) * p.F + f0 + c0 + u] = svec[i] * sS[i * pc + c0 + u] - acc[q][u];
      }
    }
  }
}

__global__ void reduce_scalar_kernel(const float* __restrict__ parts, int B, float* __restrict__ out) {
  __shared__ float red[33];
  float s = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) s += parts[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = s;
}

int reduce_scalar_parts(const float* parts, int B, float* out, cudaStream_t st) {
  reduce_scalar_kernel<<<1, 256, 0, st>>>(parts, B, out);
  AGCN_LAUNCH_CHECK();
  return AGCN_OK;
}

// ------------------------------------------------------------------------------------------------
// host-side launch configuration
// ------------------------------------------------------------------------------------------------
static int g_smem_optin = -1;

static int smem_limit() {
  if (g_smem_optin < 0) {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    g_smem_optin = v;
  }
  return g_smem_optin;
}

// opt every kernel in to the full dynamic shared memory of the device, once
template <typename Kern>
static int set_smem(Kern k, size_t bytes) {
  static std::mutex mu;
  static std::set<const void*> done;
  if (bytes > (size_t)smem_limit()) {
    set_error("shared memory request exceeds the device limit");
    return AGCN_ERR_INVALID;
  }
  std::lock_guard<std::mutex> lock(mu);
  if (!done.count((const void*)k)) {
    AGCN_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit()));
    done.insert((const void*)k);
  }
  return AGCN_OK;
}

struct ChunkCfg {
  int FC, threads, chunks;
  size_t smem;
};

// feature-chunk width / CTA size per bucket for the recurrence kernels
static ChunkCfg chunk_cfg(int max_n, int F, int nbuf, int nsq, bool single_chunk) {
  const int n4 = round4(max_n), pl = lap_pitch(n4);
  const int F4 = round4(F);
  int FC;
  if (max_n <= 32)
    FC = std::min(F4, 128);
  else if (max_n <= 64)
    FC = std::min(F4, 64);
  else
    FC = std::min(F4, 16);  // big graphs: many narrow feature chunks = many CTAs per graph
  if (single_chunk) FC = std::min(F4, max_n <= 64 ? 32 : 16);
  FC = (FC + 7) & ~7;  // (FC + 4) / 4 odd: conflict-free float4 rows
  ChunkCfg c;
  c.FC = FC;
  c.chunks = single_chunk ? 1 : (F + FC - 1) / FC;
  // one 4x4 output tile per thread where possible
  const int tiles = (n4 / 4) * (std::min(FC, F4) / 4);
  c.threads = std::min(512, std::max(64, (tiles + 31) / 32 * 32));
  c.smem = ((size_t)nsq * n4 * pl + (size_t)nbuf * n4 * (FC + 4)) * sizeof(float);
  return c;
}

int graph_chebyshev_fwd(const GraphArgs& a, cudaStream_t st, int above_n) {
  if (a.K <= 1) return AGCN_OK;
  const agcn_plan* plan = a.plan;
  const bool shortcut = (a.Lall == nullptr);
  const int nb = (int)plan->buckets.size();
  int rc = fork_streams(plan, st, nb - 1);
  if (rc) return rc;
  for (int b = 0; b < nb; ++b) {
    const Bucket& bk = plan->buckets[b];
    if (bk.max_n > plan->cheb_small_max) continue;  // row-tiled below
    if (bk.limit <= above_n) continue;              // owned by the fused tile kernel
    ChunkCfg c = chunk_cfg(bk.max_n, a.F, 2, 1, false);
    rc = set_smem(cheb_fwd_kernel, c.smem);
    if (rc) return rc;
    ChebArgs k{plan_ptrs(plan), bk.start, c.chunks, c.FC, a.F, a.K, a.X, shortcut ? a.Lint : a.Lall, shortcut ? 1 : 0,
               a.T, (int64_t)plan->R * a.F};
    cudaStream_t s = (b == 0) ? st : plan->aux[(b - 1) % 3];
    {
      ProfScope prof("cheb_fwd_kernel", s);
      cheb_fwd_kernel<<<bk.count * c.chunks, c.threads, c.smem, s>>>(k);
    }
    AGCN_LAUNCH_CHECK();
  }
  if ((rc = large_chebyshev_fwd(a, st))) return rc;  // graphs that do not fit in shared memory
  return join_streams(plan, st, nb - 1);
}

int graph_recurrence_bwd(const GraphArgs& a, bool need_dL, cudaStream_t st, int above_n) {
  const agcn_plan* plan = a.plan;
  const bool shortcut = (a.Lall == nullptr);
  const int nb = (int)plan->buckets.size();
  int rc = fork_streams(plan, st, nb - 1);
  if (rc) return rc;
  for (int b = 0; b < nb; ++b) {
    const Bucket& bk = plan->buckets[b];
    if (!need_dL && bk.max_n > plan->cheb_small_max) continue;  // row-tiled below
    if (bk.limit <= above_n) continue;                          // owned by the fused tile kernel
    ChunkCfg c = chunk_cfg(bk.max_n, a.F, need_dL ? 3 : 2, need_dL ? 2 : 1, need_dL);
    if (need_dL) c.chunks = std::max(1, std::min(a.dl_parts, (a.F + c.FC - 1) / c.FC));
    rc = set_smem(recur_bwd_kernel, c.smem);
    if (rc) return rc;
    RecurArgs k{plan_ptrs(plan), bk.start, c.chunks, c.FC, a.F, a.K, shortcut ? a.Lint : a.Lall, shortcut ? 1 : 0,
                a.G, (int64_t)plan->R * a.F, a.dX, need_dL ? 1 : 0, a.X, a.T, (int64_t)plan->R * a.F, a.dLall_in, a.dL,
                a.dl_stride};
    cudaStream_t s = (b == 0) ? st : plan->aux[(b - 1) % 3];
    {
      ProfScope prof("recur_bwd_kernel", s);
      recur_bwd_kernel<<<bk.count * c.chunks, c.threads, c.smem, s>>>(k);
    }
    AGCN_LAUNCH_CHECK();
  }
  // row-tiled reverse recurrence: every graph above AGCN_CHEB_SMALL_MAX, or, when dL is needed, only those
  // above AGCN_SMALL_MAX (the mid-size ones accumulated dL in shared memory above)
  if ((rc = large_recurrence_bwd(a, const_cast<float*>(a.G), need_dL, st))) return rc;
  if (need_dL && (rc = big_dL(a, a.G, st))) return rc;
  return join_streams(plan, st, nb - 1);
}

static int dist_chunk(int max_n) { return max_n <= 64 ? 56 : 24; }  // (FCD + 4) / 4 odd

int graph_build_laplacian(const GraphArgs& a, bool need_W, cudaStream_t st) {
  const agcn_plan* plan = a.plan;
  const int nb = (int)plan->buckets.size();
  int rc = fork_streams(plan, st, nb - 1);
  if (rc) return rc;
  for (int b = 0; b < nb; ++b) {
    const Bucket& bk = plan->buckets[b];
    const int n4 = round4(bk.max_n), pl = lap_pitch(n4);
    const int FCD = dist_chunk(bk.max_n);
    const size_t smem = ((size_t)n4 * pl + (size_t)n4 * (FCD + 4) + n4 + 64) * sizeof(float);
    rc = set_smem(build_lap_kernel, smem);
    if (rc) return rc;
    BuildArgs k{plan_ptrs(plan), bk.start, a.F, FCD, a.variant, a.lap_mode, need_W ? 1 : 0, a.XW, a.Lint, a.Lprev,
                a.alpha, a.beta, a.Lall, a.Lall_out, a.resL, a.resW, a.dist, a.dis, a.stats};
    cudaStream_t s = (b == 0) ? st : plan->aux[(b - 1) % 3];
    const int threads = (bk.max_n <= 32) ? 128 : (bk.max_n <= 64 ? 256 : 512);
    {
      ProfScope prof("build_lap_kernel", s);
      build_lap_kernel<<<bk.count, threads, smem, s>>>(k);
    }
    AGCN_LAUNCH_CHECK();
  }
  if ((rc = big_build_laplacian(a, need_W, a.big_work, st))) return rc;
  return join_streams(plan, st, nb - 1);
}

int graph_laplacian_bwd(const GraphArgs& a, cudaStream_t st) {
  const agcn_plan* plan = a.plan;
  const int nb = (int)plan->buckets.size();
  const bool paper = (a.lap_mode == AGCN_LAP_PAPER);
  const bool full = paper && a.metric_full;
  int rc = fork_streams(plan, st, nb - 1);
  if (rc) return rc;
  for (int b = 0; b < nb; ++b) {
    const Bucket& bk = plan->buckets[b];
    const int n4 = round4(bk.max_n), pl = lap_pitch(n4);
    const int FCD = dist_chunk(bk.max_n);
    const size_t smem =
        ((size_t)(paper ? 2 : 1) * n4 * pl + (full ? (size_t)n4 * (FCD + 4) : 0) + 2 * n4 + 64) * sizeof(float);
    rc = set_smem(lap_bwd_kernel, smem);
    if (rc) return rc;
    LapBwdArgs k{plan_ptrs(plan), bk.start, a.F, FCD, a.variant, a.lap_mode, a.metric_full, a.XW, a.Lint, a.Lprev,
                 a.alpha, a.beta, a.dist, a.dis, a.stats, a.dL, a.dl_parts, a.dl_stride, a.dLprev, a.dXW, a.dalpha_part,
                 a.dbeta_part};
    cudaStream_t s = (b == 0) ? st : plan->aux[(b - 1) % 3];
    const int threads = (bk.max_n <= 32) ? 128 : (bk.max_n <= 64 ? 256 : 512);
    {
      ProfScope prof("lap_bwd_kernel", s);
      lap_bwd_kernel<<<bk.count, threads, smem, s>>>(k);
    }
    AGCN_LAUNCH_CHECK();
  }
  if ((rc = big_laplacian_bwd(a, a.big_work, st))) return rc;
  return join_streams(plan, st, nb - 1);
}

}  // namespace agcn
