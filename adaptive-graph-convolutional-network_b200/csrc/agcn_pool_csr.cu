// Callers / data formats either side of the SGC-LL path:
//
//   agcn_pack_lap_csr   the intrinsic Laplacian arrives as the scipy CSR matrix Graph.compute_laplacian built
//                       (graph_structure.py:100-107) instead of the dense zero-padded [Nmax, Nmax] array of
//                       pad_Lap2sparse (graph_topology.py:92-98): ~4 stored entries per row cross PCIe, the packed
//                       dense matrices the kernels consume are expanded here, on the device.
//   agcn_graph_pool     GraphPoolMol (graphpool.py:55-110): every node takes the feature-wise maximum over the nodes
//                       its Laplacian row marks (non-zero entries: itself and its neighbours); a row with no
//                       non-zero keeps its own features.  The reference runs this as a Python loop inside
//                       tf.py_func; the argmax is recorded for callers that want a gradient (the reference has none).
#include <algorithm>

#include "agcn_internal.cuh"

namespace agcn {

// graph of a packed row: largest g with node_off[g] <= row
__device__ __forceinline__ int graph_of_row(const int32_t* __restrict__ node_off, int B, int row) {
  int lo = 0, hi = B - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (node_off[mid] <= row) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// one warp per packed row: clear the n columns of the row, then scatter its stored entries
__global__ void __launch_bounds__(256) csr_expand_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                                         const float* __restrict__ values, float* __restrict__ packed,
                                                         const int32_t* __restrict__ n_nodes,
                                                         const int32_t* __restrict__ node_off,
                                                         const int64_t* __restrict__ lap_off, int B, int R) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= R) return;
  const int g = graph_of_row(node_off, B, row);
  const int n = n_nodes[g];
  float* __restrict__ dst = packed + lap_off[g] + (int64_t)(row - node_off[g]) * n;
  for (int j = lane; j < n; j += 32) dst[j] = 0.f;
  __syncwarp();
  const int e0 = indptr[row], e1 = indptr[row + 1];
  for (int e = e0 + lane; e < e1; e += 32) {
    const int j = indices[e];
    if (j >= 0 && j < n) dst[j] = values[e];   // canonical CSR: one entry per (row, column)
  }
}

// one warp per packed row i; lanes stride over the features
__global__ void __launch_bounds__(256) graph_pool_kernel(const float* __restrict__ X, const float* __restrict__ L,
                                                         float* __restrict__ Y, int32_t* __restrict__ arg,
                                                         const int32_t* __restrict__ n_nodes,
                                                         const int32_t* __restrict__ node_off,
                                                         const int64_t* __restrict__ lap_off, int B, int R, int F) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= R) return;
  const int g = graph_of_row(node_off, B, row);
  const int n = n_nodes[g], r0 = node_off[g], i = row - r0;
  const float* __restrict__ lrow = L + lap_off[g] + (int64_t)i * n;
  for (int f0 = 0; f0 < F; f0 += 128) {   // four features per lane and sweep of the row
    float best[4];
    int who[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { best[u] = 0.f; who[u] = -1; }
    for (int j0 = 0; j0 < n; j0 += 32) {
      const float mine = (j0 + lane < n) ? lrow[j0 + lane] : 0.f;
      unsigned mask = __ballot_sync(0xffffffffu, mine != 0.f);   // graphpool.py:93 np.nonzero(l)
      while (mask) {
        const int j = j0 + __ffs(mask) - 1;
        mask &= mask - 1;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int f = f0 + 32 * u + lane;
          if (f < F) {
            const float v = X[(int64_t)(r0 + j) * F + f];
            if (who[u] < 0 || v > best[u]) { best[u] = v; who[u] = j; }   // first maximum wins, like np.amax's value
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int f = f0 + 32 * u + lane;
      if (f >= F) continue;
      if (who[u] < 0) { best[u] = X[(int64_t)row * F + f]; who[u] = i; }   // graphpool.py:97-98: no marked node
      Y[(int64_t)row * F + f] = best[u];
      if (arg) arg[(int64_t)row * F + f] = r0 + who[u];   // packed row that supplied the maximum
    }
  }
}

// dX[arg[r, f], f] += dY[r, f]
__global__ void graph_pool_bwd_kernel(const float* __restrict__ dY, const int32_t* __restrict__ arg,
                                      float* __restrict__ dX, int64_t total, int F) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(dX + (int64_t)arg[e] * F + (e % F), dY[e]);
}

}  // namespace agcn

using namespace agcn;

extern "C" {

int agcn_pack_lap_csr(const agcn_plan* plan, const int32_t* d_indptr, const int32_t* d_indices, const float* d_values,
                      float* d_packed, void* stream) {
  AGCN_REQUIRE(plan && d_indptr && d_indices && d_values && d_packed, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = plan_use(plan, st)) return rc;
  const int wpb = 8, R = (int)plan->R;
  csr_expand_kernel<<<(R + wpb - 1) / wpb, wpb * 32, 0, st>>>(d_indptr, d_indices, d_values, d_packed, plan->d_n,
                                                             plan->d_node_off, plan->d_lap_off, plan->B, R);
  AGCN_LAUNCH_CHECK();
  return AGCN_OK;
}

int agcn_graph_pool(const agcn_plan* plan, const float* d_X, const float* d_L, float* d_Y, int32_t* d_argmax, int32_t F,
                    void* stream) {
  AGCN_REQUIRE(plan && d_X && d_L && d_Y && F >= 1, "null pointer or F < 1");
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = plan_use(plan, st)) return rc;
  const int wpb = 8, R = (int)plan->R;
  graph_pool_kernel<<<(R + wpb - 1) / wpb, wpb * 32, 0, st>>>(d_X, d_L, d_Y, d_argmax, plan->d_n, plan->d_node_off,
                                                             plan->d_lap_off, plan->B, R, F);
  AGCN_LAUNCH_CHECK();
  return AGCN_OK;
}

int agcn_graph_pool_backward(const agcn_plan* plan, const float* d_dY, const int32_t* d_argmax, float* d_dX, int32_t F,
                             void* stream) {
  AGCN_REQUIRE(plan && d_dY && d_argmax && d_dX && F >= 1, "null pointer or F < 1");
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = plan_use(plan, st)) return rc;
  const int64_t total = plan->R * (int64_t)F;
  if (int rc = zero_async(d_dX, (size_t)total, st)) return rc;
  graph_pool_bwd_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, 1184), 256, 0, st>>>(d_dY, d_argmax, d_dX,
                                                                                              total, F);
  AGCN_LAUNCH_CHECK();
  return AGCN_OK;
}

}  // extern "C"
