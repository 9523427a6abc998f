// Fused SGC-LL tile kernels: the Chebyshev recurrence and the feature transform of one layer in ONE
// kernel per direction (north_star item 3), for the molecule-sized graphs that dominate the Tox21 /
// ToxCast shapes.
//
// A "tile" is 128 rows of the packed node matrix that belong to whole graphs (agcn_plan.cu packs graphs
// with n <= AGCN_FUSE_MAX_N into tiles, first-fit decreasing).  One CTA owns one tile:
//
//   forward  (graphconv.py:221-247, :118-123)
//     warps 2-5 (one thread per tile row): T_0 chunk (32 feature columns) -> shared memory, then
//       T_1 = L T_0, T_k = 2 L T_{k-1} - T_{k-2} on the CUDA cores in exact fp32 with the per-graph L
//       matrices of the tile resident in shared memory; every T_k chunk is written once to HBM (saved for
//       backward) and, split into hi/lo TF32 halves, into a 128x32 K-major swizzled operand stage;
//     warp 1: tcgen05.mma (3xTF32) Y += T_k[:, chunk] W_k[chunk, :] into a TMEM accumulator;
//     warp 0: TMA producer of the pre-split W tiles;
//     epilogue: TMEM -> registers -> bias + activation -> Y.
//   backward (reverse mode of the same lines, dX chain only)
//     mainloop: G_z = dYpre W_z^T for z = 0..K-1 into K TMEM accumulators (A = dYpre chunks split by the
//       workers, B = W_z by TMA);
//     epilogue: U_{K-1} = G_{K-1}, U_j = G_j + c_{j+1} L^T U_{j+1} - U_{j+2} per 32-column chunk, reading
//       G_j straight from TMEM; dX = U_0.  The K [R,F] G matrices never touch HBM.
//
// Graphs with n > AGCN_FUSE_MAX_N own "pre" tiles (128-row ranges of one graph): their recurrences run in
// the per-graph / row-tiled kernels (agcn_graph_small.cu, agcn_graph_large.cu) and the tile kernel only
// does the tensor-core part for those rows (forward: T_k read from HBM; backward: G_z written to HBM).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <utility>
#include <vector>

#include "agcn_internal.cuh"

namespace agcn {
namespace ft {

constexpr int TM = 128;            // rows per tile (UMMA M)
constexpr int CH = 32;             // feature columns per k-block = 128 bytes = one swizzle row
constexpr int UMMA_K = 8;          // tf32
constexpr int A_BYTES = TM * CH * 4;  // 16 KB: one operand half (hi or lo) of a k-block
constexpr int LCAP = AGCN_FUSE_LCAP;  // floats of per-graph L matrices per tile

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { mbar_wait_shared(smem_u32(bar), parity); }
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row atoms of 1024 bytes
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, int cols) {
  const uint32_t s = smem_u32(slot);
  switch (cols) {
    case 32: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;\n" ::"r"(s) : "memory"); break;
    case 64: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;\n" ::"r"(s) : "memory"); break;
    case 128: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;\n" ::"r"(s) : "memory"); break;
    case 256: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n" ::"r"(s) : "memory"); break;
    default: asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(s) : "memory"); break;
  }
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, int cols) {
  switch (cols) {
    case 32: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;\n" ::"r"(base) : "memory"); break;
    case 64: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;\n" ::"r"(base) : "memory"); break;
    case 128: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;\n" ::"r"(base) : "memory"); break;
    case 256: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" ::"r"(base) : "memory"); break;
    default: asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(base) : "memory"); break;
  }
}
// 32 consecutive fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float v[32]) {
  uint32_t u[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];\n"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
        "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
        "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(u[i]);
}

// ---- explicit shared-space accesses (32-bit shared addresses): the operand stages and chunk buffers are carved
// from one dynamic allocation with integer arithmetic, so plain pointers would compile to generic LD/ST
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ int ldsi32(uint32_t a) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];\n" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};\n" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t sdst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sdst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t sdst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sdst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// ---- fp32 chunk buffers: [128 rows][32 floats] with a row pitch of 36 floats (144 bytes): rows r and r' read by
// one warp (different graphs) collide only if r == r' (mod 8), and a row's 16-byte groups sit at immediate
// offsets from one base register
constexpr int CPITCH = 144;
constexpr int CBUF_BYTES = TM * CPITCH;  // 18 KB
__device__ __forceinline__ uint32_t chunk_addr(uint32_t buf, int row, int group) {
  return buf + (uint32_t)(row * CPITCH + (group << 4));
}

// Worker warp (quarter q, column half h): its 32 rows x 16 columns of chunk c, global rows s_grow[row] -> buf.
__device__ __forceinline__ void load_rows_async(uint32_t buf, const float* __restrict__ M, int ld, int ncols, int c,
                                                uint32_t s_grow, int q, int h, int lane, bool vec) {
  if (vec) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = q * 32 + 8 * i + (lane >> 2), grp = 4 * h + (lane & 3), col = c * CH + grp * 4;
      const int grow = ldsi32(s_grow + 4 * row);
      const uint32_t d = chunk_addr(buf, row, grp);
      if (grow >= 0 && col < ncols)
        cp_async16(d, M + (int64_t)grow * ld + col);
      else
        sts128(d, make_float4(0.f, 0.f, 0.f, 0.f));
    }
  } else {
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
      const int row = q * 32 + 2 * i + (lane >> 4), cc = 16 * h + (lane & 15), col = c * CH + cc;
      const int grow = ldsi32(s_grow + 4 * row);
      const uint32_t d = chunk_addr(buf, row, cc >> 2) + 4 * (cc & 3);
      if (grow >= 0 && col < ncols)
        cp_async4(d, M + (int64_t)grow * ld + col);
      else
        sts32(d, 0.f);
    }
  }
}

// The same 32 x 16 block, buf -> global rows (coalesced).  Callers bracket with __syncwarp.
__device__ __forceinline__ void store_rows(uint32_t buf, float* __restrict__ M, int ld, int ncols, int c, uint32_t s_grow,
                                           int q, int h, int lane, bool vec) {
  if (vec) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = q * 32 + 8 * i + (lane >> 2), grp = 4 * h + (lane & 3), col = c * CH + grp * 4;
      const int grow = ldsi32(s_grow + 4 * row);
      if (grow >= 0 && col < ncols) *reinterpret_cast<float4*>(M + (int64_t)grow * ld + col) = lds128(chunk_addr(buf, row, grp));
    }
  } else {
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
      const int row = q * 32 + 2 * i + (lane >> 4), cc = 16 * h + (lane & 15), col = c * CH + cc;
      const int grow = ldsi32(s_grow + 4 * row);
      if (grow >= 0 && col < ncols) M[(int64_t)grow * ld + col] = lds32(chunk_addr(buf, row, cc >> 2) + 4 * (cc & 3));
    }
  }
}

// my half row (16 columns: groups 4h .. 4h+3)
__device__ __forceinline__ void read_half(uint32_t buf, int row, int h, float v[16]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float4 x = lds128(chunk_addr(buf, row, 4 * h + g));
    v[4 * g] = x.x; v[4 * g + 1] = x.y; v[4 * g + 2] = x.z; v[4 * g + 3] = x.w;
  }
}
__device__ __forceinline__ void write_half(uint32_t buf, int row, int h, const float v[16]) {
#pragma unroll
  for (int g = 0; g < 4; ++g)
    sts128(chunk_addr(buf, row, 4 * h + g), make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]));
}

// acc[:] += sum_j L[laddr + j * lstride_bytes] * src[r0 + j][16h .. 16h+15]
// Four rows per trip with every shared-memory load issued before the first FMA that needs it (2 warps per
// scheduler cannot hide the LDS latency otherwise).
__device__ __forceinline__ void fma_row(float a, const float4 b[4], float acc[16]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    acc[4 * g] = fmaf(a, b[g].x, acc[4 * g]);
    acc[4 * g + 1] = fmaf(a, b[g].y, acc[4 * g + 1]);
    acc[4 * g + 2] = fmaf(a, b[g].z, acc[4 * g + 2]);
    acc[4 * g + 3] = fmaf(a, b[g].w, acc[4 * g + 3]);
  }
}
__device__ __forceinline__ void lap_times_rows(uint32_t laddr, int lstride_bytes, uint32_t src, int r0, int n, int h,
                                               float acc[16]) {
  uint32_t ta = src + (uint32_t)(r0 * CPITCH + h * 64);
  uint32_t la = laddr;
  int j = 0;
  for (; j + 4 <= n; j += 4) {
    float a[4];
    float4 b[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      a[u] = lds32(la + (uint32_t)(u * lstride_bytes));
#pragma unroll
      for (int g = 0; g < 4; ++g) b[u][g] = lds128(ta + (uint32_t)(u * CPITCH + 16 * g));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) fma_row(a[u], b[u], acc);
    la += 4 * lstride_bytes;
    ta += 4 * CPITCH;
  }
  for (; j < n; ++j) {
    const float a = lds32(la);
    float4 b[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) b[g] = lds128(ta + 16 * g);
    fma_row(a, b, acc);
    la += lstride_bytes;
    ta += CPITCH;
  }
}

// ---- sparse rows.  The Laplacian of a molecule has ~3 non-zeros per row (reference_literal mode multiplies by
// I + L_int, the normalised Laplacian of a graph of degree <= 4): a dense n x n product spends 80 .. 95 % of its shared
// memory loads and FMAs on exact zeros.  Every worker keeps the non-zero pattern of its row (forward) / column
// (backward) as a 64-bit mask (n <= AGCN_FUSE_MAX_N = 64) and walks the set bits; the values stay where they are.
// Skipping a term whose coefficient is exactly 0 does not change an fp32 sum of finite values.
__device__ __forceinline__ unsigned long long nonzero_mask(uint32_t laddr, int lstride_bytes, int n) {
  unsigned long long m = 0;
#pragma unroll 4
  for (int j = 0; j < n; ++j)
    if (lds32(laddr + (uint32_t)(j * lstride_bytes)) != 0.f) m |= 1ull << j;
  return m;
}
// Warp-uniform choice between the dense loop (n x 21 instructions) and the masked loop (28 per non-zero); the warp runs
// as long as its slowest lane either way.
__device__ __forceinline__ bool prefer_masked(unsigned long long mask, int n) {
  const unsigned dense_cost = __reduce_max_sync(0xffffffffu, (unsigned)(n * 3));
  const unsigned masked_cost = __reduce_max_sync(0xffffffffu, (unsigned)(__popcll(mask) * 4));
  return masked_cost < dense_cost;
}
// acc[:] += sum over the set bits j of L[laddr + j * lstride_bytes] * src[r0 + j][16h .. 16h+15], two terms per trip
__device__ __forceinline__ void lap_times_rows_masked(uint32_t laddr, int lstride_bytes, uint32_t src, int r0,
                                                      unsigned long long mask, int h, float acc[16]) {
  const uint32_t tbase = src + (uint32_t)(r0 * CPITCH + h * 64);
  while (mask) {
    const int j0 = __ffsll((long long)mask) - 1;
    mask &= mask - 1;
    const bool two = mask != 0;
    const int j1 = two ? __ffsll((long long)mask) - 1 : j0;
    mask &= mask - 1;   // no-op when mask is already 0
    const float a0 = lds32(laddr + (uint32_t)(j0 * lstride_bytes));
    const float a1 = two ? lds32(laddr + (uint32_t)(j1 * lstride_bytes)) : 0.f;
    float4 b0[4], b1[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      b0[g] = lds128(tbase + (uint32_t)(j0 * CPITCH + 16 * g));
      b1[g] = lds128(tbase + (uint32_t)(j1 * CPITCH + 16 * g));
    }
    fma_row(a0, b0, acc);
    fma_row(a1, b1, acc);
  }
}

// My half of one row of a 128x32 K-major SWIZZLE_128B operand tile, split into hi / lo TF32 halves.
__device__ __forceinline__ void write_operand_half(uint32_t a_hi, uint32_t a_lo, int row, int h, const float v[16]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float4 hi, lo;
    hi.x = tf32_rn(v[4 * g]);
    hi.y = tf32_rn(v[4 * g + 1]);
    hi.z = tf32_rn(v[4 * g + 2]);
    hi.w = tf32_rn(v[4 * g + 3]);
    lo.x = tf32_rn(v[4 * g] - hi.x);
    lo.y = tf32_rn(v[4 * g + 1] - hi.y);
    lo.z = tf32_rn(v[4 * g + 2] - hi.z);
    lo.w = tf32_rn(v[4 * g + 3] - hi.w);
    const uint32_t off = (uint32_t)(row * 128 + (((4 * h + g) ^ (row & 7)) << 4));
    sts128(a_hi + off, hi);
    sts128(a_lo + off, lo);
  }
}

// 16 consecutive fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
  uint32_t u[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(u[i]);
}

// ------------------------------------------------------------------------------------------------
// shared memory carve-up (identical for both directions)
// ------------------------------------------------------------------------------------------------
constexpr int WORKERS = 256;  // warps 2..9: two threads per tile row (column halves)
constexpr int THREADS = 64 + WORKERS;

__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }

struct SmemPlan {
  int stages, stage_bytes, b_bytes;
  int off_bufs;   // 3 fp32 chunk buffers (forward) -- the backward recurrence reuses the operand stages
  int off_L;      // LCAP floats
  int off_glist;  // 128 entries of 2 x int4
  int off_grow;   // 128 int
  int off_rowinfo;  // 128 int4 {r0, n, lbase, i}
  int off_bars;
  int total;
};

__host__ __device__ inline SmemPlan smem_plan(int N, bool forward) {
  SmemPlan s;
  s.b_bytes = N * 128;
  s.stage_bytes = 2 * A_BYTES + 2 * s.b_bytes;
  const int fixed = (forward ? 3 * CBUF_BYTES : 0) + LCAP * 4 + 128 * 32 + 128 * 4 + 128 * 16 + 256 + 1024;
  int stages = (227 * 1024 - fixed) / s.stage_bytes;
  s.stages = stages > 4 ? 4 : stages;
  int off = s.stages * s.stage_bytes;
  s.off_bufs = off;
  off += forward ? 3 * CBUF_BYTES : 0;
  s.off_L = off;
  off += LCAP * 4;
  s.off_glist = off;
  off += 128 * 32;
  s.off_grow = off;
  off += 128 * 4;
  s.off_rowinfo = off;
  off += 128 * 16;
  s.off_bars = off;
  off += 256;
  s.total = off + 1024;  // alignment slack
  return s;
}

struct TileRow {
  int grow;   // global packed row or -1
  int n;      // nodes of my graph (0: pre tile or padding row)
  int r0;     // tile row of my graph's first node
  int lbase;  // float offset of my graph's matrix in the tile's L region
  int i;      // my index inside the graph
  int pitch;
  bool pre;
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}
// timeline slots per tile (debug builds of a launch: agcn_fused_debug_set)
constexpr int DBG_SLOTS = 128;
#define FT_STAMP(slot)                                                        \
  do {                                                                        \
    if (p.t.dbg) p.t.dbg[(long long)tile * DBG_SLOTS + (slot)] = gtime();     \
  } while (0)

struct TileArgs {
  unsigned long long* dbg;     // optional timeline buffer [tiles][DBG_SLOTS] (nanoseconds), NULL in production
  const int4* tile_graphs;     // 2 x int4 per entry: {g, r0 | row_start, n | nrows, lbase | -1}, {node_off, lap_off lo, hi, 0}
  const int32_t* tile_gstart;  // [tiles + 1]
  const float* L;              // packed Laplacians (Lint or L_all)
  int add_identity;
  int tile0;                   // first tile of this launch
  int F, Fo, K;
  int N;                       // MMA N (padded output columns of the mainloop)
  int nchunks;                 // k-blocks per slice
};

// Prologue shared by both kernels (worker threads): graph list and row table of the tile; the copies of the
// per-graph L matrices are left in flight (cp.async): callers wait + worker_sync before the first use.
__device__ __forceinline__ TileRow tile_prologue(const TileArgs& p, int tile, int r, int h, int wt, uint32_t s_glist,
                                                 uint32_t s_grow, uint32_t s_rowinfo, uint32_t sL, int* ng_out) {
  const int gs = p.tile_gstart[tile], ng = p.tile_gstart[tile + 1] - gs;
  *ng_out = ng;
  for (int e = wt; e < 2 * ng; e += WORKERS) {
    const int4 v = __ldg(p.tile_graphs + 2 * gs + e);
    asm volatile("st.shared.v4.s32 [%0], {%1,%2,%3,%4};\n" ::"r"(s_glist + 16 * e), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  }
  worker_sync();
  TileRow t;
  t.grow = -1; t.n = 0; t.r0 = 0; t.lbase = 0; t.i = 0; t.pitch = 1; t.pre = false;
  for (int e = 0; e < ng; ++e) {
    int gx, gy, gz, gw;
    asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];\n" : "=r"(gx), "=r"(gy), "=r"(gz), "=r"(gw) : "r"(s_glist + 32 * e) : "memory");
    const int noff = ldsi32(s_glist + 32 * e + 16);
    if (gw < 0) {  // pre tile: rows [gy, gy + gz) of graph gx
      t.pre = true;
      if (r < gz) t.grow = noff + gy + r;
    } else if (r >= gy && r < gy + gz) {
      t.grow = noff + (r - gy);
      t.n = gz; t.r0 = gy; t.lbase = gw; t.i = r - gy; t.pitch = gz | 1;
    }
  }
  if (h == 0) {
    asm volatile("st.shared.s32 [%0], %1;\n" ::"r"(s_grow + 4 * r), "r"(t.grow) : "memory");
    asm volatile("st.shared.v4.s32 [%0], {%1,%2,%3,%4};\n" ::"r"(s_rowinfo + 16 * r), "r"(t.r0), "r"(t.n), "r"(t.lbase), "r"(t.i) : "memory");
  }
  // per-graph matrices, row pitch n | 1 (odd: the rows read by neighbouring lanes sit in different banks)
  for (int e = 0; e < ng; ++e) {
    int gx, gy, gz, gw, lo, hi;
    asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];\n" : "=r"(gx), "=r"(gy), "=r"(gz), "=r"(gw) : "r"(s_glist + 32 * e) : "memory");
    if (gw < 0) continue;
    lo = ldsi32(s_glist + 32 * e + 20);
    hi = ldsi32(s_glist + 32 * e + 24);
    const long long loff = (long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo);
    const int n = gz, pitch = n | 1;
    const float* __restrict__ src = p.L + loff;
    const uint32_t dst = sL + 4 * gw;
    int i = wt / n, j = wt - i * n;  // element wt of the n x n matrix, then steps of 256
    const int di = WORKERS / n, dj = WORKERS - di * n;
    for (int idx = wt; idx < n * n; idx += WORKERS) {
      cp_async4(dst + 4 * (i * pitch + j), src + idx);
      i += di; j += dj;
      if (j >= n) { j -= n; ++i; }
    }
  }
  cp_async_commit();
  return t;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
struct FwdArgs {
  TileArgs t;
  const float* X;     // [R,F]
  float* T;           // [K-1][R][F] saved Chebyshev terms
  long long tslice;
  const float* bias;
  int act;
  float* Y;           // [R,Fo]
};

__global__ void __launch_bounds__(THREADS, 1)
fused_fwd_kernel(const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo, FwdArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(base);
  const SmemPlan sp = smem_plan(p.t.N, true);
  const uint32_t bufs = sbase + sp.off_bufs, sL = sbase + sp.off_L, s_glist = sbase + sp.off_glist,
                 s_grow = sbase + sp.off_grow, s_rowinfo = sbase + sp.off_rowinfo;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + sp.off_bars);
  uint64_t* full_bar = bars;        // W tiles landed (TMA)
  uint64_t* split_bar = bars + 4;   // operand rows written by the 256 workers
  uint64_t* empty_bar = bars + 8;   // MMAs that read the stage retired
  uint64_t* tmem_full_bar = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = p.t.tile0 + blockIdx.x;
  const int F = p.t.F, Fo = p.t.Fo, K = p.t.K, N = p.t.N, nc = p.t.nchunks;
  const int num_kb = nc * K;
  int tmem_cols = 32;
  while (tmem_cols < N) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < sp.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&split_bar[s], WORKERS / 32);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer: W_s[chunk c] as a [N, 32] K-major tile, hi and lo halves =================
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int stage = kb % sp.stages, phase = (kb / sp.stages) & 1;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = base + stage * sp.stage_bytes + 2 * A_BYTES;
        const int c = kb / K, s = kb - c * K;
        mbar_expect_tx(&full_bar[stage], 2 * sp.b_bytes);
        tma_load_2d(st, &tmBhi, &full_bar[stage], c * CH, s * N);
        tma_load_2d(st + sp.b_bytes, &tmBlo, &full_bar[stage], c * CH, s * N);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int stage = kb % sp.stages, phase = (kb / sp.stages) & 1;
        mbar_wait(&full_bar[stage], phase);
        if (kb < 16) FT_STAMP(72 + 3 * kb);
        mbar_wait(&split_bar[stage], phase);
        if (kb < 16) FT_STAMP(73 + 3 * kb);
        tc_fence_after();
        const uint32_t sa = sbase + stage * sp.stage_bytes;
        const uint32_t sa_lo = sa + A_BYTES, sb_hi = sa + 2 * A_BYTES, sb_lo = sb_hi + sp.b_bytes;
#pragma unroll
        for (int k = 0; k < CH / UMMA_K; ++k) {
          const uint32_t koff = k * UMMA_K * 4;
          const uint64_t a_hi = make_desc(sa + koff), a_lo = make_desc(sa_lo + koff);
          const uint64_t b_hi = make_desc(sb_hi + koff), b_lo = make_desc(sb_lo + koff);
          umma_tf32(tmem_base, a_lo, b_hi, idesc, (kb | k) != 0);
          umma_tf32(tmem_base, a_hi, b_lo, idesc, 1);
          umma_tf32(tmem_base, a_hi, b_hi, idesc, 1);
        }
        umma_commit(&empty_bar[stage]);
        if (kb < 16) FT_STAMP(74 + 3 * kb);
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ================= workers: recurrence + operand production, then epilogue =================
    const int q = warp & 3;             // TMEM lane quarter of this warp
    const int h = (warp - 2) >> 2;      // column half
    const int r = q * 32 + lane;        // my tile row
    const int wt = (warp - 2) * 32 + lane;
    int ng;
    if (wt == 0) FT_STAMP(0);
    const TileRow me = tile_prologue(p.t, tile, r, h, wt, s_glist, s_grow, s_rowinfo, sL, &ng);
    worker_sync();  // row table visible
    if (wt == 0) FT_STAMP(1);
    const bool vecX = ((F & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.X) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(p.T) & 15) == 0) && ((p.tslice & 3) == 0);
    const uint32_t xbuf[2] = {bufs, bufs + CBUF_BYTES};
    const uint32_t tbuf = bufs + 2 * CBUF_BYTES;
    const uint32_t lrow = sL + 4 * (me.lbase + me.i * me.pitch);

    auto emit = [&](int kb, const float v[16]) {
      const int stage = kb % sp.stages, phase = (kb / sp.stages) & 1;
      if (lane == 0) mbar_wait(&empty_bar[stage], phase ^ 1);
      __syncwarp();
      const uint32_t st = sbase + stage * sp.stage_bytes;
      write_operand_half(st, st + A_BYTES, r, h, v);
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy writes -> tensor-core reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&split_bar[stage]);  // one arrival per worker warp
    };

    if (me.pre) {
      // 128-row range of a big graph: the per-graph / row-tiled kernels produced T_1..T_{K-1}; this tile only feeds
      // the tensor core.  My half row of k-block kb + 1 is in flight while k-block kb is split and handed over.
      cp_async_wait_all();  // (nothing of this tile, but keeps the group accounting of the prologue simple)
      auto load_kb = [&](int kb, float v[16]) {
        const int c = kb / K, sl = kb - c * K;
        const float* src = (sl == 0) ? p.X : p.T + (long long)(sl - 1) * p.tslice;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int col = c * CH + 16 * h + u;
          v[u] = (me.grow >= 0 && col < F) ? __ldg(src + (long long)me.grow * F + col) : 0.f;
        }
      };
      float nxt[16];
      load_kb(0, nxt);
      for (int kb = 0; kb < num_kb; ++kb) {
        float v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = nxt[u];
        if (kb + 1 < num_kb) load_kb(kb + 1, nxt);
        emit(kb, v);
      }
    } else {
      load_rows_async(xbuf[0], p.X, F, F, 0, s_grow, q, h, lane, vecX);
      cp_async_commit();
      unsigned long long row_mask = 0;
      bool masked = false;
      for (int c = 0; c < nc; ++c) {
        const int cur = c & 1;
        cp_async_wait_all();
        worker_sync();  // T_0 chunk c (and, first time, the L matrices) complete; chunk c-1 is finished everywhere
        if (c == 0) {
          row_mask = nonzero_mask(lrow, 4, me.n);
          masked = prefer_masked(row_mask, me.n);
        }
        if (wt == 0 && c < 8) FT_STAMP(2 + 8 * c);
        if (c + 1 < nc) {
          load_rows_async(xbuf[cur ^ 1], p.X, F, F, c + 1, s_grow, q, h, lane, vecX);
          cp_async_commit();
        }
        float tm2[16], tm1[16];  // my half row of T_{s-2}, T_{s-1}
        read_half(xbuf[cur], r, h, tm1);
        emit(c * K, tm1);
        if (wt == 0 && c < 8) FT_STAMP(3 + 8 * c);
        uint32_t src = xbuf[cur], dst = tbuf;
        for (int s = 1; s < K; ++s) {
          float t[16];
          {
  #pragma unroll
            for (int u = 0; u < 16; ++u) t[u] = p.t.add_identity ? tm1[u] : 0.f;  // L_all = I + L_int (literal mode)
            if (masked)
              lap_times_rows_masked(lrow, 4, src, me.r0, row_mask, h, t);
            else
              lap_times_rows(lrow, 4, src, me.r0, me.n, h, t);  // graphconv.py:231
            if (s >= 2) {
  #pragma unroll
              for (int u = 0; u < 16; ++u) t[u] = 2.f * t[u] - tm2[u];  // graphconv.py:234
            }
            if (wt == 0 && c < 8 && s < 3) FT_STAMP(2 + 8 * c + 2 * s);
            emit(c * K + s, t);       // the tensor core gets its operand first ...
            write_half(dst, r, h, t);  // ... then the next step's input and the copy saved for backward
            __syncwarp();
            store_rows(dst, p.T + (long long)(s - 1) * p.tslice, F, F, c, s_grow, q, h, lane, vecX);
          }
          if (wt == 0 && c < 8 && s < 3) FT_STAMP(3 + 8 * c + 2 * s);
  #pragma unroll
          for (int u = 0; u < 16; ++u) { tm2[u] = tm1[u]; tm1[u] = t[u]; }
          if (s + 1 < K) {
            worker_sync();  // T_s rows of every graph of the tile are in `dst`
            const uint32_t tmp = src; src = dst; dst = tmp;
          }
        }
      }
    }
    // ---- epilogue: Y = act(acc + bias)   graphconv.py:245-247, :118-123
    if (wt == 0) FT_STAMP(66);
    if (lane == 0) mbar_wait(tmem_full_bar, 0);
    __syncwarp();
    tc_fence_after();
    if (wt == 0) FT_STAMP(67);
    // Every worker's operand rows were consumed before tmem_full_bar fired (operand write -> split_bar -> MMA ->
    // commit), so the stages are free; the named barrier states that ordering between the workers explicitly
    // (compute-sanitizer racecheck does not follow the tcgen05.commit edge and reports the reuse otherwise).
    worker_sync();
    const uint32_t stg = sbase + (uint32_t)((warp - 2) * (32 * 36 * 4));  // operand stages are free now
    const bool vecY = ((Fo & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.Y) & 15) == 0);
    for (int c0 = 32 * h; c0 < N && c0 < Fo; c0 += 64) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        sts128(stg + 4 * (lane * 36 + 4 * u), make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]));
      __syncwarp();
      const int cc = c0 + 4 * (lane & 7);
      float bv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (p.bias && cc + e < Fo) bv[e] = __ldg(p.bias + cc + e);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + (lane >> 3);
        const int grow = ldsi32(s_grow + 4 * (q * 32 + rr));
        if (grow < 0 || cc >= Fo) continue;
        const float4 o4 = lds128(stg + 4 * (rr * 36 + 4 * (lane & 7)));
        float o[4] = {o4.x + bv[0], o4.y + bv[1], o4.z + bv[2], o4.w + bv[3]};
        if (p.act == AGCN_ACT_RELU) {
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = fmaxf(o[e], 0.f);
        }
        float* dstp = p.Y + (long long)grow * Fo + cc;
        if (vecY && cc + 3 < Fo) {
          *reinterpret_cast<float4*>(dstp) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (cc + e < Fo) dstp[e] = o[e];
        }
      }
      __syncwarp();
    }
  }
  if (threadIdx.x == 64) FT_STAMP(68);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// backward (dX chain)
// ------------------------------------------------------------------------------------------------
struct BwdArgs {
  TileArgs t;          // t.N = padded F (MMA N), t.nchunks = ceil(Fo / 32)
  const float* dYp;    // [R,Fo]  dY
  const float* Y;      // [R,Fo]  activated output: dYpre = dY * [Y > 0]; NULL: linear activation
  float* G;            // [K][R][F]: written for pre tiles only (their recurrence runs in the per-graph kernels)
  long long gslice;
  float* dX;           // [R,F]
  int acc_stride;      // TMEM columns between the K accumulators
};

__global__ void __launch_bounds__(THREADS, 1)
fused_bwd_kernel(const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo, BwdArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(base);
  const SmemPlan sp = smem_plan(p.t.N, false);
  const uint32_t sL = sbase + sp.off_L, s_glist = sbase + sp.off_glist, s_grow = sbase + sp.off_grow,
                 s_rowinfo = sbase + sp.off_rowinfo;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + sp.off_bars);
  uint64_t* full_bar = bars;
  uint64_t* split_bar = bars + 4;
  uint64_t* empty_bar = bars + 8;
  uint64_t* tmem_full_bar = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = p.t.tile0 + blockIdx.x;
  const int F = p.t.F, Fo = p.t.Fo, K = p.t.K, N = p.t.N, nc = p.t.nchunks;
  const int num_kb = nc * K;
  int tmem_cols = 32;
  while (tmem_cols < K * p.acc_stride) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < sp.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&split_bar[s], WORKERS / 32);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer: W_z[:, chunk c] as a [N = F, 32] K-major tile =================
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int stage = kb % sp.stages, phase = (kb / sp.stages) & 1;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = base + stage * sp.stage_bytes + 2 * A_BYTES;
        const int z = kb / nc, c = kb - z * nc;
        mbar_expect_tx(&full_bar[stage], 2 * sp.b_bytes);
        tma_load_2d(st, &tmBhi, &full_bar[stage], c * CH, z * N);
        tma_load_2d(st + sp.b_bytes, &tmBlo, &full_bar[stage], c * CH, z * N);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: accumulator z at TMEM column z * acc_stride =================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int stage = kb % sp.stages, phase = (kb / sp.stages) & 1;
        const int z = kb / nc, c = kb - z * nc;
        mbar_wait(&full_bar[stage], phase);
        mbar_wait(&split_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = sbase + stage * sp.stage_bytes;
        const uint32_t sa_lo = sa + A_BYTES, sb_hi = sa + 2 * A_BYTES, sb_lo = sb_hi + sp.b_bytes;
        const uint32_t d = tmem_base + (uint32_t)(z * p.acc_stride);
#pragma unroll
        for (int k = 0; k < CH / UMMA_K; ++k) {
          const uint32_t koff = k * UMMA_K * 4;
          const uint64_t a_hi = make_desc(sa + koff), a_lo = make_desc(sa_lo + koff);
          const uint64_t b_hi = make_desc(sb_hi + koff), b_lo = make_desc(sb_lo + koff);
          umma_tf32(d, a_lo, b_hi, idesc, (c | k) != 0);
          umma_tf32(d, a_hi, b_lo, idesc, 1);
          umma_tf32(d, a_hi, b_hi, idesc, 1);
        }
        umma_commit(&empty_bar[stage]);
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int wt = (warp - 2) * 32 + lane;
    int ng;
    const TileRow me = tile_prologue(p.t, tile, r, h, wt, s_glist, s_grow, s_rowinfo, sL, &ng);
    const bool vecD = ((Fo & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.dYp) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(p.Y) & 15) == 0);
    // mainloop: my half row of dYpre chunk c (L1-resident across the K passes) -> hi/lo operand rows.  The rows of the
    // next TWO k-blocks are in flight while this one is split and stored: dY and Y travel as RAW values and the
    // relu' mask is applied only when the row is consumed -- masking at load time makes the compare wait for the
    // load and turns the prefetch into a 1 - 2 us stall per k-block (ncu source view: every top stall of the
    // round-1 kernel sat on those FSETPs).
    auto load_raw = [&](int kb, float d[16], float m[16]) {
      if (kb >= num_kb) return;
      const int c = kb % nc;
      if (vecD) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col = c * CH + 16 * h + 4 * g;
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = make_float4(1.f, 1.f, 1.f, 1.f);
          if (me.grow >= 0 && col < Fo) {
            x = __ldg(reinterpret_cast<const float4*>(p.dYp + (long long)me.grow * Fo + col));
            if (p.Y) y = __ldg(reinterpret_cast<const float4*>(p.Y + (long long)me.grow * Fo + col));
          }
          d[4 * g] = x.x; d[4 * g + 1] = x.y; d[4 * g + 2] = x.z; d[4 * g + 3] = x.w;
          m[4 * g] = y.x; m[4 * g + 1] = y.y; m[4 * g + 2] = y.z; m[4 * g + 3] = y.w;
        }
      } else {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int col = c * CH + 16 * h + u;
          float x = 0.f, y = 1.f;
          if (me.grow >= 0 && col < Fo) {
            x = __ldg(p.dYp + (long long)me.grow * Fo + col);
            if (p.Y) y = __ldg(p.Y + (long long)me.grow * Fo + col);
          }
          d[u] = x;
          m[u] = y;
        }
      }
    };
    auto hand_over = [&](int kb, float d[16], float m[16]) {
      if (kb >= num_kb) return;
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = (m[u] > 0.f) ? d[u] : 0.f;   // relu'(0) = 0 (TF's ReluGrad)
      load_raw(kb + 2, d, m);
      const int stage = kb % sp.stages, phase = (kb / sp.stages) & 1;
      if (lane == 0) mbar_wait(&empty_bar[stage], phase ^ 1);
      __syncwarp();
      const uint32_t st = sbase + stage * sp.stage_bytes;
      write_operand_half(st, st + A_BYTES, r, h, v);
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&split_bar[stage]);
    };
    {
      float d0[16], m0[16], d1[16], m1[16];
      load_raw(0, d0, m0);
      load_raw(1, d1, m1);
      for (int kb = 0; kb < num_kb; kb += 2) {
        hand_over(kb, d0, m0);
        hand_over(kb + 1, d1, m1);
      }
    }
    // ---- epilogue: reverse recurrence on the accumulators
    cp_async_wait_all();  // the L matrices of the tile (issued in the prologue)
    if (lane == 0) mbar_wait(tmem_full_bar, 0);
    __syncwarp();
    tc_fence_after();
    worker_sync();
    const uint32_t ub[2] = {sbase, sbase + CBUF_BYTES};  // operand stages are free now
    const bool vecX = ((F & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.dX) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(p.G) & 15) == 0) && ((p.gslice & 3) == 0);
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(16 * h);
    const uint32_t lcol = sL + 4 * (me.lbase + me.i);  // column i of my graph's matrix: (L^T U)_i = sum_j L[j][i] U_j
    const unsigned long long col_mask = me.pre ? 0ull : nonzero_mask(lcol, 4 * me.pitch, me.n);
    const bool masked = prefer_masked(col_mask, me.n);
    const int nfc = (F + CH - 1) / CH;
    for (int fc = 0; fc < nfc; ++fc) {
      if (me.pre) {
        // big graph: hand G_z to the per-graph / row-tiled reverse recurrence
        for (int z = 0; z < K; ++z) {
          float g[16];
          tmem_ld16(lane_base + (uint32_t)(z * p.acc_stride + fc * CH), g);
          write_half(ub[0], r, h, g);
          __syncwarp();
          store_rows(ub[0], p.G + (long long)z * p.gslice, F, F, fc, s_grow, q, h, lane, vecX);
          __syncwarp();
        }
        continue;
      }
      float u1[16], u2[16];
      tmem_ld16(lane_base + (uint32_t)((K - 1) * p.acc_stride + fc * CH), u1);  // U_{K-1} = G_{K-1}
#pragma unroll
      for (int u = 0; u < 16; ++u) u2[u] = 0.f;
      int cur = 0;
      for (int j = K - 2; j >= 0; --j) {
        write_half(ub[cur], r, h, u1);
        worker_sync();  // U_{j+1} rows of every graph of the tile are visible
        float acc[16], g[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) acc[u] = p.t.add_identity ? u1[u] : 0.f;  // (I + L)^T U = U + L^T U
        if (masked)
          lap_times_rows_masked(lcol, 4 * me.pitch, ub[cur], me.r0, col_mask, h, acc);
        else
          lap_times_rows(lcol, 4 * me.pitch, ub[cur], me.r0, me.n, h, acc);
        tmem_ld16(lane_base + (uint32_t)(j * p.acc_stride + fc * CH), g);
        const float cmul = (j + 1 >= 2) ? 2.f : 1.f;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const float o = g[u] + cmul * acc[u] - u2[u];
          u2[u] = u1[u];
          u1[u] = o;
        }
        cur ^= 1;
      }
      // dX = U_0
      worker_sync();  // every read of the buffers is done before they are reused for the store / next chunk
      write_half(ub[cur], r, h, u1);
      __syncwarp();
      store_rows(ub[cur], p.dX, F, F, fc, s_grow, q, h, lane, vecX);
      worker_sync();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// W prep: out_{hi,lo}[(z * N + n) * Kp + k] = split(W[n * sn + k * sk + z * sz])  (zero outside n < Nv, k < Kv)
// ------------------------------------------------------------------------------------------------
__global__ void prep_w_kernel(const float* __restrict__ W, long long sn, long long sk, long long sz, int Nv, int Kv,
                              int N, int Kp, int Z, float* __restrict__ hi, float* __restrict__ lo) {
  const long long total = (long long)Z * N * Kp;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e % Kp);
    const long long rr = e / Kp;
    const int n = (int)(rr % N), z = (int)(rr / N);
    float x = 0.f;
    if (n < Nv && k < Kv) x = W[n * sn + k * sk + z * sz];
    const float h = tf32_rn(x);
    hi[e] = h;
    lo[e] = tf32_rn(x - h);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// [rows, cols] fp32 row-major, box = 32 columns x box_rows, 128-byte swizzle
static int make_map(CUtensorMap* map, const float* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return AGCN_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)CH, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return AGCN_ERR_CUDA;
  }
  return AGCN_OK;
}

static int pad16(int x) { return (x + 15) & ~15; }
static int pad32(int x) { return (x + 31) & ~31; }

static unsigned long long* g_dbg = nullptr;

static TileArgs tile_args(const agcn_plan* plan, const float* L, int add_identity, int F, int Fo, int K) {
  TileArgs t;
  t.dbg = g_dbg;
  t.tile_graphs = reinterpret_cast<const int4*>(plan->d_ft_entries);
  t.tile_gstart = plan->d_ft_gstart;
  t.L = L;
  t.add_identity = add_identity;
  t.tile0 = 0;
  t.F = F; t.Fo = Fo; t.K = K;
  t.N = 0; t.nchunks = 0;
  return t;
}

}  // namespace ft

// ------------------------------------------------------------------------------------------------
// host API
// ------------------------------------------------------------------------------------------------
void rows_debug_set(void* d_buf);
void cheb_debug_set(void* d_buf);
void fused_debug_set(void* d_buf) {
  ft::g_dbg = reinterpret_cast<unsigned long long*>(d_buf);
  rows_debug_set(d_buf);   // the all-rows contraction (agcn_pre_tile.cu) stamps the same buffer
  cheb_debug_set(d_buf);   // ... and the recurrence tiles (agcn_cheb_tile.cu), slots 100..
}

bool fused_enabled() {
  static const bool off = ab_env("AGCN_DISABLE_FUSED") != nullptr || ab_env("AGCN_DISABLE_TCGEN05") != nullptr;
  return !off;
}

bool fused_fwd_supported(const agcn_plan* plan, int F, int Fo, int K) {
  return fused_enabled() && plan->ft_tiles > 0 && K >= 2 && Fo >= 1 && Fo <= 128 && F >= 1;
}

bool fused_bwd_supported(const agcn_plan* plan, int F, int Fo, int K) {
  if (!fused_enabled() || plan->ft_tiles <= 0 || K < 2 || F > 128) return false;
  const int N = ft::pad16(F), stride = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
  return K * stride <= 512;
}

size_t fused_w_floats(int Nv, int Kv, int Z) { return 2 * (size_t)Z * ft::pad16(Nv) * ft::pad32(Kv); }

// forward operand: B_s[n, k] = weight[(k*K + s)*Fo + n]  (n < Fo output columns, k < F)
int fused_fwd_prep(const float* weight, int F, int Fo, int K, float* scratch, cudaStream_t st) {
  const int N = ft::pad16(Fo), Kp = ft::pad32(F);
  const long long total = (long long)K * N * Kp;
  const int blocks = (int)std::min<long long>((total + 255) / 256, 1184);
  ft::prep_w_kernel<<<blocks, 256, 0, st>>>(weight, 1, (long long)K * Fo, Fo, Fo, F, N, Kp, K, scratch, scratch + total);
  AGCN_LAUNCH_CHECK();
  return AGCN_OK;
}

// backward operand: B_z[f, o] = weight[(f*K + z)*Fo + o]  (f < F rows of G_z, o < Fo contraction)
int fused_bwd_prep(const float* weight, int F, int Fo, int K, float* scratch, cudaStream_t st) {
  const int N = ft::pad16(F), Kp = ft::pad32(Fo);
  const long long total = (long long)K * N * Kp;
  const int blocks = (int)std::min<long long>((total + 255) / 256, 1184);
  ft::prep_w_kernel<<<blocks, 256, 0, st>>>(weight, (long long)K * Fo, 1, Fo, F, Fo, N, Kp, K, scratch, scratch + total);
  AGCN_LAUNCH_CHECK();
  return AGCN_OK;
}

template <typename Kern>
static int opt_in_smem(Kern k, int bytes) {
  static std::mutex mu;
  static int done_bytes = 0;
  std::lock_guard<std::mutex> lock(mu);
  if (bytes > done_bytes) {
    AGCN_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done_bytes = bytes;
  }
  return AGCN_OK;
}

int fused_forward(const agcn_plan* plan, int tile0, int ntiles, const float* X, const float* L, int add_identity,
                  const float* wsplit, const float* bias, int act, int F, int Fo, int K, float* T, float* Y,
                  cudaStream_t st) {
  using namespace ft;
  if (ntiles <= 0) return AGCN_OK;
  const int N = pad16(Fo), Kp = pad32(F);
  const long long half = (long long)K * N * Kp;
  CUtensorMap mhi, mlo;
  int rc;
  if ((rc = make_map(&mhi, wsplit, (uint64_t)K * N, (uint64_t)Kp, (uint32_t)N))) return rc;
  if ((rc = make_map(&mlo, wsplit + half, (uint64_t)K * N, (uint64_t)Kp, (uint32_t)N))) return rc;
  FwdArgs a;
  a.t = tile_args(plan, L, add_identity, F, Fo, K);
  a.t.N = N;
  a.t.nchunks = Kp / CH;
  a.t.tile0 = tile0;
  a.X = X; a.T = T; a.tslice = (long long)plan->R * F;
  a.bias = bias; a.act = act; a.Y = Y;
  const SmemPlan sp = smem_plan(N, true);
  if ((rc = opt_in_smem(fused_fwd_kernel, 227 * 1024))) return rc;
  {
    ProfScope prof(tile0 == 0 ? "ft::fused_fwd_kernel" : "ft::fused_fwd_kernel(pre tiles)", st);
    fused_fwd_kernel<<<ntiles, THREADS, sp.total, st>>>(mhi, mlo, a);
  }
  AGCN_LAUNCH_CHECK();
  return AGCN_OK;
}

int fused_backward(const agcn_plan* plan, int tile0, int ntiles, const float* dYp, const float* Y, const float* L,
                   int add_identity, const float* wsplit, int F, int Fo, int K, float* G, float* dX, cudaStream_t st) {
  using namespace ft;
  if (ntiles <= 0) return AGCN_OK;
  const int N = pad16(F), Kp = pad32(Fo);
  const long long half = (long long)K * N * Kp;
  CUtensorMap mhi, mlo;
  int rc;
  if ((rc = make_map(&mhi, wsplit, (uint64_t)K * N, (uint64_t)Kp, (uint32_t)N))) return rc;
  if ((rc = make_map(&mlo, wsplit + half, (uint64_t)K * N, (uint64_t)Kp, (uint32_t)N))) return rc;
  BwdArgs a;
  a.t = tile_args(plan, L, add_identity, F, Fo, K);
  a.t.N = N;
  a.t.nchunks = Kp / CH;
  a.t.tile0 = tile0;
  a.dYp = dYp; a.Y = Y; a.G = G; a.gslice = (long long)plan->R * F; a.dX = dX;
  a.acc_stride = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
  const SmemPlan sp = smem_plan(N, false);
  if ((rc = opt_in_smem(fused_bwd_kernel, 227 * 1024))) return rc;
  {
    ProfScope prof(tile0 == 0 ? "ft::fused_bwd_kernel" : "ft::fused_bwd_kernel(pre tiles)", st);
    fused_bwd_kernel<<<ntiles, THREADS, sp.total, st>>>(mhi, mlo, a);
  }
  AGCN_LAUNCH_CHECK();
  return AGCN_OK;
}

}  // namespace agcn
