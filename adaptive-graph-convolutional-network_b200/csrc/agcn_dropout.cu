// Dropout on the packed node matrices (models/layers/dropout.py:27-41 -> tf.nn.dropout(x, keep_prob = 1 - p, seed)):
// y = x * keep / (1 - p), keep ~ Bernoulli(1 - p) per element.  The mask is a counter-based Philox4x32-10 stream keyed
// by (seed, element index), so the backward pass regenerates it from the seed instead of storing it: the same call
// on dY with the same seed is the gradient.
#include "agcn_internal.cuh"

namespace agcn {

__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
  const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}

__device__ __forceinline__ uint4 philox4x32_10(uint64_t counter, uint64_t seed) {
  uint32_t c0 = (uint32_t)counter, c1 = (uint32_t)(counter >> 32), c2 = 0u, c3 = 0u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// one thread per group of 4 consecutive elements (one Philox block)
__global__ void __launch_bounds__(256) dropout_kernel(const float* __restrict__ x, float* __restrict__ y, long long n,
                                                      float p, float inv_keep, unsigned long long seed) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i0 = q * 4;
  if (i0 >= n) return;
  const uint4 r = philox4x32_10((uint64_t)q, seed);
  const uint32_t bits[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (i0 + e < n) {
      const float u = (float)(bits[e] >> 8) * (1.0f / 16777216.0f);   // uniform in [0, 1)
      y[i0 + e] = (u >= p) ? x[i0 + e] * inv_keep : 0.f;
    }
  }
}

}  // namespace agcn

extern "C" int agcn_dropout(const float* d_X, float* d_Y, int64_t n, float p, uint64_t seed, void* stream) {
  AGCN_REQUIRE(d_X && d_Y && n >= 0, "dropout: bad arguments");
  AGCN_REQUIRE(p >= 0.f && p < 1.f, "dropout: p must be in [0, 1)");
  if (n == 0) return AGCN_OK;
  const long long groups = (n + 3) / 4;
  agcn::dropout_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_X, d_Y, (long long)n, p,
                                                                                           1.f / (1.f - p), seed);
  AGCN_LAUNCH_CHECK();
  return AGCN_OK;
}
