// Fused SGC-LL tile kernel: the Chebyshev recurrence AND the feature transform of one layer on the tensor cores, in
// ONE kernel per direction (north_star item 3), for every graph of up to 128 nodes.
//
// A "tile" is up to 128 rows of the packed node matrix that belong to whole graphs (agcn_plan.cu packs the graphs
// first-fit decreasing).  One CTA owns one tile and runs, for V_0 = the tile's input rows,
//
//     V_1 = Lt V_0,   V_s = 2 Lt V_{s-1} - V_{s-2},      Out = sum_s V_s B_s                         (*)
//
//   forward  (graphconv.py:221-247, :118-123): V = T (Chebyshev terms of X), Lt = L_all (block diagonal over the
//            tile's graphs), B_s = W_s, Out = Y before bias + activation; T_1 .. T_{K-1} are saved for dweight.
//   backward (reverse mode of the same lines, dX chain): sum_s T_s(L) X W_s is linear in X, so
//            dX = sum_s T_s(L^T) dYpre W_s^T: the SAME recurrence on dYpre = dY * relu'(Y) with Lt = L_all^T and
//            B_s = W_s^T.  Nothing is saved.
//
// Where the operands live:
//   Lt     the tile's block-diagonal 128 x 128 matrix, split into hi / lo TF32 halves, in TENSOR MEMORY (2 x 128
//          columns), written once per tile with tcgen05.st straight from the packed Laplacians: it is the A operand of
//          every recurrence product (tcgen05.mma with A in TMEM), so the 128 KB it would take in shared memory stay
//          free for the pipeline and the recurrence costs no CUDA-core FMAs and no shared-memory bandwidth;
//   V_s    produced 32 feature columns ("chunk") at a time by the 8 worker warps (thread = tile row x column half):
//          read from global (s = 0) or from the recurrence accumulator in TMEM (s >= 1, tcgen05.ld), combined with
//          V_{s-2}, and written ONCE in the two shapes the tensor core wants -- rows x chunk (K-major A operand of the
//          transform product V_s[:, chunk] B_s[chunk, :]) and chunk x rows (K-major B operand of the next recurrence
//          product Lt V_s[:, chunk]) -- each split hi / lo (3xTF32: a_lo b_hi + a_hi b_lo + a_hi b_hi);
//   B_s    pre-split parameter tiles streamed by TMA;
//   Out    one TMEM accumulator for the whole tile (all chunks, all s); the recurrence has one 32-column accumulator
//          per chunk in flight.
// Column chunks are independent chains (chunk c of V_s depends on chunk c of V_{s-1} only): two chains are in flight
// on two shared-memory slots, so the hand-off latency of one (tcgen05.ld -> split -> st.shared -> fence -> mma ->
// commit) hides behind the other.
//
// Graphs with more than 128 nodes own "pre" tiles (128-row ranges of one graph): their recurrences run in the
// per-graph / row-tiled kernels (agcn_graph_small.cu, agcn_big_tc.cu) and the tile kernel does the transform part for
// those rows (forward: V_s read from the saved T; backward: G_z = dYpre W_z^T written for the reverse recurrence).
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
constexpr int CH = 32;             // feature columns per chunk = 128 bytes = one swizzle row
constexpr int UMMA_K = 8;          // tf32
constexpr int A_BYTES = TM * CH * 4;   // 16 KB: one half (hi or lo) of a rows x chunk operand
constexpr int R_BYTES = CH * TM * 4;   // 16 KB: one half of a chunk x rows operand (4 k-blocks of [32][32])
constexpr int WORKERS = 256;       // warps 2..9: two threads per tile row (column halves)
constexpr int THREADS = 64 + WORKERS;
constexpr int SLOTS = 2;
constexpr uint32_t SPIN_LIMIT = 1u << 26;
// tensor memory columns
constexpr int TM_LT_HI = 0, TM_LT_LO = 128, TM_REC = 256 /* + 64 * slot: [hh + lh | hl] */, TM_OUT = 384;

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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done && ++spins > SPIN_LIMIT) __trap();  // a protocol bug becomes an error, not a hang
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(dst),
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
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]   (A: lane = row, one column per K element)
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}
// 16 consecutive fp32 columns of this thread's TMEM lane
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
// 32 consecutive columns of this thread's TMEM lane <- registers
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float v[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31,%32};\n" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
      "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
      "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};\n" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ int ldsi32(uint32_t a) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];\n" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }

// My half (16 columns: 16-byte groups 4h .. 4h+3) of row `row` of a rows x chunk operand ([128][32] K-major,
// SWIZZLE_128B), split into hi / lo TF32 halves.
__device__ __forceinline__ void write_rows_operand(uint32_t a_hi, uint32_t a_lo, int row, int h, const float v[16]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float4 hi, lo;
    hi.x = tf32_rn(v[4 * g]);     hi.y = tf32_rn(v[4 * g + 1]);
    hi.z = tf32_rn(v[4 * g + 2]); hi.w = tf32_rn(v[4 * g + 3]);
    lo.x = tf32_rn(v[4 * g] - hi.x);     lo.y = tf32_rn(v[4 * g + 1] - hi.y);
    lo.z = tf32_rn(v[4 * g + 2] - hi.z); lo.w = tf32_rn(v[4 * g + 3] - hi.w);
    const uint32_t off = (uint32_t)(row * 128 + (((4 * h + g) ^ (row & 7)) << 4));
    sts128(a_hi + off, hi);
    sts128(a_lo + off, lo);
  }
}
// The same 16 values (times `scale`, a power of two) as my column `row` of the chunk x rows operand: 4 k-blocks of
// [64][32 tile rows] (K-major, SWIZZLE_128B) whose rows 0..31 are the hi halves of the chunk's 32 features and rows
// 32..63 the lo halves, so that ONE N = 64 product with Lt_hi yields [Lt_hi v_hi | Lt_hi v_lo]; feature f of tile row
// r sits in k-block r / 32, row f (+ 32), element r % 32.  The 32 lanes of a warp (consecutive tile rows) fill one
// 128-byte row: no bank conflicts.
__device__ __forceinline__ void write_cols_operand(uint32_t b_hi, uint32_t b_lo, int row, int h, const float v[16], float scale) {
  const uint32_t kb = (uint32_t)(row >> 5) * 8192u, e = (uint32_t)(row & 31);
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const int f = 16 * h + u;
    const float x = v[u] * scale;
    const float hi = tf32_rn(x);
    const uint32_t off = kb + (uint32_t)(f * 128) + ((((e >> 2) ^ (uint32_t)(f & 7))) << 4) + ((e & 3) << 2);
    sts32(b_hi + off, hi);
    sts32(b_lo + off, tf32_rn(x - hi));
  }
}

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}
// timeline slots per tile (debug launches: agcn_fused_debug_set)
constexpr int DBG_SLOTS = 128;
#define FT_STAMP(slot)                                                      \
  do {                                                                      \
    if (p.dbg) p.dbg[(long long)tile * DBG_SLOTS + (slot)] = gtime();       \
  } while (0)

struct SmemPlan {
  int w_bytes;     // one half of a parameter tile: N x 128 bytes
  int slot_bytes;  // rows operand (hi, lo) + cols operand (hi, lo) + parameter tile (hi, lo)
  int off_glist;   // 128 entries of 2 x int4
  int off_bars;
  int total;
};
__host__ __device__ inline SmemPlan smem_plan(int N) {
  SmemPlan s;
  s.w_bytes = N * 128;
  s.slot_bytes = 2 * A_BYTES + 2 * R_BYTES + 2 * s.w_bytes;
  s.off_glist = SLOTS * s.slot_bytes;
  s.off_bars = s.off_glist + 128 * 32;
  s.total = s.off_bars + 256 + 1024;  // + alignment slack
  return s;
}

struct TileArgs {
  unsigned long long* dbg;     // optional timeline buffer [tiles][DBG_SLOTS] (nanoseconds), NULL in production
  const int4* tile_graphs;     // 2 x int4 per entry: {g, r0 | row_start, n | nrows, >= 0 | -1}, {node_off, lap_off lo, hi, 0}
  const int32_t* tile_gstart;  // [tiles + 1]
  const float* L;              // packed Laplacians (Lint or L_all)
  int add_identity;            // Lt = I + L (literal SGC_LL: L_all = I + L_int)
  int transL;                  // Lt = (.)^T  (backward)
  int tile0;                   // first tile of this launch
  int Fr;                      // recurrence width (columns of V): F forward, Fo backward
  int Fout;                    // columns of Out: Fo forward, F backward
  int K;
  int N;                       // MMA N of the transform product (Fout padded to 16)
  int nchunks;                 // ceil(Fr / 32)
  const float* In;             // [R, Fr]   forward: X, backward: dY
  const float* Mask;           // backward: Y (dYpre = dY * [Y > 0]) or NULL
  float* Save;                 // forward: T [K-1][R][Fr] (written; read for pre tiles); backward: scratch for K >= 4
  long long save_slice;
  const float* bias;           // forward
  int act;
  float* Out;                  // [R, Fout]   forward: Y, backward: dX (not written for pre tiles)
  float* G;                    // backward, pre tiles: [K][R][Fout]
  long long g_slice;
  int forward;
};

struct RowInfo {
  int grow;   // global packed row or -1
  int n;      // nodes of my graph (0: pre tile or padding row)
  int r0;     // tile row of my graph's first node
  int i;      // my index inside the graph
  long long lap;  // element offset of my graph's matrix
};

__global__ void __launch_bounds__(THREADS, 1)
fused_tile_kernel(const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo, TileArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(base);
  const SmemPlan sp = smem_plan(p.N);
  const uint32_t s_glist = sbase + sp.off_glist;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + sp.off_bars);
  uint64_t* full_bar = bars;        // [SLOTS] parameter tile landed (TMA)
  uint64_t* ops_bar = bars + 2;     // [SLOTS] operands written by the 8 worker warps
  uint64_t* done_bar = bars + 4;    // [SLOTS] the MMAs that read the slot (and wrote its recurrence accumulator) retired
  uint64_t* out_bar = bars + 6;     // Out accumulator complete (all items, or one z of a backward pre tile)
  uint64_t* outfree_bar = bars + 7; // backward pre tiles: Out accumulator drained by the workers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  int* s_pre = reinterpret_cast<int*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = p.tile0 + blockIdx.x;
  const int K = p.K, N = p.N, nc = p.nchunks, Fr = p.Fr, Fout = p.Fout;
  const int gs = p.tile_gstart[tile], ng = p.tile_gstart[tile + 1] - gs;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SLOTS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&ops_bar[s], WORKERS / 32);
      mbar_init(&done_bar[s], 1);
    }
    mbar_init(out_bar, 1);
    mbar_init(outfree_bar, WORKERS / 32);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    *s_pre = (p.tile_graphs[2 * gs].w < 0) ? 1 : 0;
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool pre = *s_pre != 0;
  // backward pre tiles produce one G_z at a time (z outer, chunks inner); everything else runs the chains
  const bool zloop = pre && !p.forward;

  if (warp == 0) {
    // ================= TMA producer: the parameter tile of every item =================
    if (lane == 0) {
      int use[SLOTS] = {0, 0};
      auto item = [&](int cc, int s, int slot) {
        const int u = use[slot]++;
        if (u > 0) mbar_wait(&done_bar[slot], (uint32_t)((u - 1) & 1));
        const uint32_t dst = sbase + slot * sp.slot_bytes + 2 * A_BYTES + 2 * R_BYTES;
        mbar_expect_tx(&full_bar[slot], 2 * sp.w_bytes);
        tma_load_2d(dst, &tmBhi, &full_bar[slot], cc * CH, s * N);
        tma_load_2d(dst + sp.w_bytes, &tmBlo, &full_bar[slot], cc * CH, s * N);
      };
      if (zloop) {
        for (int z = 0; z < K; ++z)
          for (int cc = 0; cc < nc; ++cc) item(cc, z, cc & 1);
      } else {
        for (int g0 = 0; g0 < nc; g0 += SLOTS)
          for (int s = 0; s < K; ++s)
            for (int ci = 0; ci < SLOTS && g0 + ci < nc; ++ci) item(g0 + ci, s, ci);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      // D = f32, A = B = tf32, both K-major
      const uint32_t idesc_x = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      const uint32_t idesc_r = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(CH >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      const uint32_t idesc_r2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * CH) >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      int use[SLOTS] = {0, 0};
      int tcount = 0;
      bool first = true;
      auto item = [&](int cc, int s, int slot, bool recur) {
        const int u = use[slot]++;
        const int t = tcount++;
        mbar_wait(&full_bar[slot], (uint32_t)(u & 1));
        if (t < 14) FT_STAMP(72 + 3 * t);
        mbar_wait(&ops_bar[slot], (uint32_t)(u & 1));
        if (t < 14) FT_STAMP(73 + 3 * t);
        tc_fence_after();
        const uint32_t sa = sbase + slot * sp.slot_bytes, sa_lo = sa + A_BYTES;
        const uint32_t sr = sa + 2 * A_BYTES;
        const uint32_t sw = sr + 2 * R_BYTES, sw_lo = sw + sp.w_bytes;
        // transform: Out += V_s[:, chunk] B_s[chunk, :]
#pragma unroll
        for (int k = 0; k < CH / UMMA_K; ++k) {
          const uint32_t koff = k * UMMA_K * 4;
          const uint64_t a_hi = make_desc(sa + koff), a_lo = make_desc(sa_lo + koff);
          const uint64_t b_hi = make_desc(sw + koff), b_lo = make_desc(sw_lo + koff);
          umma_ss(tmem_base + TM_OUT, a_lo, b_hi, idesc_x, !(first && k == 0));
          umma_ss(tmem_base + TM_OUT, a_hi, b_lo, idesc_x, 1);
          umma_ss(tmem_base + TM_OUT, a_hi, b_hi, idesc_x, 1);
        }
        first = false;
        // recurrence: Rec[slot] = Lt (c V_s[:, chunk]), the tile's block-diagonal Lt from tensor memory.  A product with
        // the A operand in tensor memory costs its A fetch (128 x 8 elements) whatever N is, so the two products that
        // share Lt_hi are ONE N = 64 instruction over the operand's [v_hi | v_lo] rows: columns 0..31 of the
        // accumulator collect Lt_hi v_hi + Lt_lo v_hi, columns 32..63 Lt_hi v_lo; the workers add the two halves.
        if (recur) {
          const uint32_t d = tmem_base + TM_REC + 64 * slot;
#pragma unroll 4
          for (int k = 0; k < TM / UMMA_K; ++k) {
            const uint64_t b = make_desc(sr + (uint32_t)(k >> 2) * 8192u + (uint32_t)(k & 3) * (UMMA_K * 4));
            const uint32_t a_hi = tmem_base + TM_LT_HI + k * UMMA_K, a_lo = tmem_base + TM_LT_LO + k * UMMA_K;
            umma_ts(d, a_hi, b, idesc_r2, k != 0);
            umma_ts(d, a_lo, b, idesc_r, 1);
          }
        }
        umma_commit(&done_bar[slot]);
        if (t < 14) FT_STAMP(74 + 3 * t);
      };
      if (zloop) {
        for (int z = 0; z < K; ++z) {
          if (z > 0) {
            mbar_wait(outfree_bar, (uint32_t)((z - 1) & 1));
            tc_fence_after();
          }
          first = true;
          for (int cc = 0; cc < nc; ++cc) item(cc, z, cc & 1, false);
          umma_commit(out_bar);
        }
      } else {
        for (int g0 = 0; g0 < nc; g0 += SLOTS)
          for (int s = 0; s < K; ++s)
            for (int ci = 0; ci < SLOTS && g0 + ci < nc; ++ci) item(g0 + ci, s, ci, !pre && s + 1 < K);
        umma_commit(out_bar);
      }
    }
  } else {
    // ================= workers =================
    const int q = warp & 3;             // TMEM lane quarter of this warp
    const int h = (warp - 2) >> 2;      // column half
    const int r = q * 32 + lane;        // my tile row
    const int wt = (warp - 2) * 32 + lane;
    if (wt == 0) FT_STAMP(0);
    // ---- graph list of the tile -> my row
    for (int e = wt; e < 2 * ng; e += WORKERS) {
      const int4 v = __ldg(p.tile_graphs + 2 * gs + e);
      asm volatile("st.shared.v4.s32 [%0], {%1,%2,%3,%4};\n" ::"r"(s_glist + 16 * e), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
    worker_sync();
    RowInfo me;
    me.grow = -1; me.n = 0; me.r0 = 0; me.i = 0; me.lap = 0;
    for (int e = 0; e < ng; ++e) {
      int gx, gy, gz, gw;
      asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];\n" : "=r"(gx), "=r"(gy), "=r"(gz), "=r"(gw) : "r"(s_glist + 32 * e) : "memory");
      const int noff = ldsi32(s_glist + 32 * e + 16);
      if (gw < 0) {  // pre tile: rows [gy, gy + gz) of graph gx
        if (r < gz) me.grow = noff + gy + r;
      } else if (r >= gy && r < gy + gz) {
        me.grow = noff + (r - gy);
        me.n = gz; me.r0 = gy; me.i = r - gy;
        const int lo = ldsi32(s_glist + 32 * e + 20), hi = ldsi32(s_glist + 32 * e + 24);
        me.lap = (long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo);
      }
    }
    // ---- Lt -> tensor memory.  The tile's matrices are first staged in shared memory (the rows / cols operand area of
    // slot 0, 64 KB, idle now; the parameter tiles that TMA is already bringing sit behind it) with coalesced
    // asynchronous copies, row pitch n | 1 (odd: rows and columns are both conflict-free to read; a 128-node graph
    // fills the area exactly with pitch 128 and rotates row i by i elements instead);
    // then every thread builds its row of the block-diagonal matrix -- row i of L, or column i for Lt = L^T -- and
    // stores its half (h = 0: hi, 1: lo) with tcgen05.st.  Reading L row by row straight from global memory costs
    // 16 .. 46 us per tile in a cold first wave (profiles/r02_a_tile_v2_timeline_ts_n32.txt).
    if (!pre && K > 1) {
      int my_base = 0;
      {
        int acc = 0;
        for (int e = 0; e < ng; ++e) {
          int gx, gy, gz, gw;
          asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];\n" : "=r"(gx), "=r"(gy), "=r"(gz), "=r"(gw) : "r"(s_glist + 32 * e) : "memory");
          const int lo = ldsi32(s_glist + 32 * e + 20), hi = ldsi32(s_glist + 32 * e + 24);
          const long long loff = (long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo);
          const int n = gz, pitch = (n == TM) ? TM : (n | 1);
          if (r >= gy && r < gy + gz) my_base = acc;
          const float* __restrict__ src = p.L + loff;
          const uint32_t dst = sbase + 4u * (uint32_t)acc;
          int i = wt / n, j = wt - i * n;  // element wt of the n x n matrix, then steps of 256
          const int di = WORKERS / n, dj = WORKERS - di * n;
          for (int idx = wt; idx < n * n; idx += WORKERS) {
            const int jj = (n == TM) ? ((i + j) & (TM - 1)) : j;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst + 4u * (uint32_t)(i * pitch + jj)), "l"(src + idx) : "memory");
            i += di; j += dj;
            if (j >= n) { j -= n; ++i; }
          }
          acc += n * pitch;
          (void)gx; (void)gw;
        }
      }
      asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
      worker_sync();
      const uint32_t lt = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h ? TM_LT_LO : TM_LT_HI);
      const bool rot = me.n == TM;
      const int pitch = rot ? TM : (me.n | 1);
      const uint32_t mine0 = sbase + 4u * (uint32_t)my_base;
#pragma unroll 1
      for (int jb = 0; jb < 4; ++jb) {
        float v[32];
        const int j0 = 32 * jb - me.r0;   // graph-local column of the group's first element
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const int j = j0 + jj;
          float x = 0.f;
          if (me.n > 0 && j >= 0 && j < me.n) {
            // element (row, col) = (i, j) of L, or (j, i) for the transpose
            const int er = p.transL ? j : me.i, ec = p.transL ? me.i : j;
            const int ecs = rot ? ((er + ec) & (TM - 1)) : ec;
            asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(x) : "r"(mine0 + 4u * (uint32_t)(er * pitch + ecs)) : "memory");
            if (p.add_identity && j == me.i) x += 1.f;
          }
          const float hi = tf32_rn(x);
          v[jj] = h ? tf32_rn(x - hi) : hi;
        }
        tmem_st32(lt + 32 * jb, v);
      }
      tmem_st_wait();
      worker_sync();   // the staging area is about to become operand slot 0
    }
    if (wt == 0) FT_STAMP(1);

    const bool vecIn = ((Fr & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.In) & 15) == 0) &&
                       (!p.Mask || (reinterpret_cast<uintptr_t>(p.Mask) & 15) == 0);
    const bool vecSave = ((Fr & 3) == 0) && p.Save && ((reinterpret_cast<uintptr_t>(p.Save) & 15) == 0) && ((p.save_slice & 3) == 0);
    // my 16 columns of chunk cc of the input (backward: masked by relu'(Y), relu'(0) = 0 like TF's ReluGrad)
    auto load_in = [&](int cc, float v[16]) {
      const int c0 = cc * CH + 16 * h;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = c0 + 4 * g;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (me.grow >= 0 && col < Fr) {
          const long long o = (long long)me.grow * Fr + col;
          if (vecIn) {
            x = __ldg(reinterpret_cast<const float4*>(p.In + o));
            if (p.Mask) {
              const float4 y = __ldg(reinterpret_cast<const float4*>(p.Mask + o));
              x.x = y.x > 0.f ? x.x : 0.f; x.y = y.y > 0.f ? x.y : 0.f;
              x.z = y.z > 0.f ? x.z : 0.f; x.w = y.w > 0.f ? x.w : 0.f;
            }
          } else {
            float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (col + e < Fr) {
                t[e] = __ldg(p.In + o + e);
                if (p.Mask && !(__ldg(p.Mask + o + e) > 0.f)) t[e] = 0.f;
              }
            x = make_float4(t[0], t[1], t[2], t[3]);
          }
        }
        v[4 * g] = x.x; v[4 * g + 1] = x.y; v[4 * g + 2] = x.z; v[4 * g + 3] = x.w;
      }
    };
    // my 16 columns of chunk cc of saved term s (V_s, s >= 1): plain loads (this very thread wrote them)
    auto load_saved = [&](int cc, int s, float v[16]) {
      const int c0 = cc * CH + 16 * h;
      const float* src = p.Save + (long long)(s - 1) * p.save_slice;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = c0 + 4 * g;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (me.grow >= 0 && col < Fr) {
          const long long o = (long long)me.grow * Fr + col;
          if (vecSave) {
            x = *reinterpret_cast<const float4*>(src + o);
          } else {
            float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (col + e < Fr) t[e] = src[o + e];
            x = make_float4(t[0], t[1], t[2], t[3]);
          }
        }
        v[4 * g] = x.x; v[4 * g + 1] = x.y; v[4 * g + 2] = x.z; v[4 * g + 3] = x.w;
      }
    };
    auto store_saved = [&](int cc, int s, const float v[16]) {
      const int c0 = cc * CH + 16 * h;
      float* dst = p.Save + (long long)(s - 1) * p.save_slice;
      if (me.grow < 0) return;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = c0 + 4 * g;
        if (col >= Fr) continue;
        const long long o = (long long)me.grow * Fr + col;
        if (vecSave) {
          *reinterpret_cast<float4*>(dst + o) = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (col + e < Fr) dst[o + e] = v[4 * g + e];
        }
      }
    };
    // Out accumulator -> global rows (bias + activation in the forward direction)
    auto drain_out = [&](float* dstm) {
      const bool vecO = ((Fout & 3) == 0) && ((reinterpret_cast<uintptr_t>(dstm) & 15) == 0);
      for (int c0 = 16 * h; c0 < N && c0 < Fout; c0 += 32) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(TM_OUT + c0), v);
        if (me.grow < 0) continue;
        if (p.forward) {
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            float o = v[u] + ((p.bias && c0 + u < Fout) ? __ldg(p.bias + c0 + u) : 0.f);
            if (p.act == AGCN_ACT_RELU) o = fmaxf(o, 0.f);
            v[u] = o;
          }
        }
        float* dst = dstm + (long long)me.grow * Fout + c0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (c0 + 4 * g >= Fout) continue;
          if (vecO && c0 + 4 * g + 3 < Fout) {
            *reinterpret_cast<float4*>(dst + 4 * g) = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (c0 + 4 * g + e < Fout) dst[4 * g + e] = v[4 * g + e];
          }
        }
      }
    };

    int use[SLOTS] = {0, 0};
    int tcount = 0;
    // one hand-off: V_s[:, chunk cc] -> rows operand (transform) and, when another step follows, cols operand
    auto item = [&](int cc, int s, int slot, const float* pre0) {
      const int u = use[slot]++;
      const int t = tcount++;
      if (wt == 0 && t < 14) FT_STAMP(8 + 4 * t);
      if (u > 0) {
        if (lane == 0) mbar_wait(&done_bar[slot], (uint32_t)((u - 1) & 1));
        __syncwarp();
        tc_fence_after();
      }
      if (wt == 0 && t < 14) FT_STAMP(9 + 4 * t);
      float v[16];
      if (s == 0) {
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = pre0[e];
      } else if (pre) {
        load_saved(cc, s, v);
      } else {
        {
          float v2[16];
          const uint32_t rec = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(TM_REC + 64 * slot + 16 * h);
          tmem_ld16(rec, v);
          tmem_ld16(rec + 32, v2);
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] += v2[e];
        }
        if (s >= 2) {
          float w[16];
          if (s == 2) load_in(cc, w); else load_saved(cc, s - 2, w);
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] -= w[e];     // the operand was pre-scaled by 2: acc = 2 Lt V_{s-1}
        }
        if (p.forward || s + 2 < K) store_saved(cc, s, v);   // forward: saved for dweight; backward: V_{s+2} needs it
      }
      if (wt == 0 && t < 14) FT_STAMP(10 + 4 * t);
      const uint32_t st = sbase + slot * sp.slot_bytes;
      write_rows_operand(st, st + A_BYTES, r, h, v);
      if (!pre && s + 1 < K) write_cols_operand(st + 2 * A_BYTES, st + 2 * A_BYTES + 4096, r, h, v, s == 0 ? 1.f : 2.f);
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy writes -> tensor-core reads
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ops_bar[slot]);
      if (wt == 0 && t < 14) FT_STAMP(11 + 4 * t);
    };

    if (zloop) {
      // backward, 128-row range of a big graph: G_z = dYpre W_z^T for the per-graph / row-tiled reverse recurrence
      for (int z = 0; z < K; ++z) {
        for (int cc = 0; cc < nc; ++cc) {
          float v0[16];
          load_in(cc, v0);
          item(cc, 0, cc & 1, v0);
        }
        if (lane == 0) mbar_wait(out_bar, (uint32_t)(z & 1));
        __syncwarp();
        tc_fence_after();
        drain_out(p.G + (long long)z * p.g_slice);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(outfree_bar);
      }
    } else {
      float nxt[SLOTS][16];
      for (int ci = 0; ci < SLOTS && ci < nc; ++ci) load_in(ci, nxt[ci]);
      for (int g0 = 0; g0 < nc; g0 += SLOTS) {
        float cur[SLOTS][16];
#pragma unroll
        for (int ci = 0; ci < SLOTS; ++ci)
#pragma unroll
          for (int e = 0; e < 16; ++e) cur[ci][e] = nxt[ci][e];
        for (int s = 0; s < K; ++s) {
          if (s == K - 1)   // the next group's inputs travel while this group's last step is handed over
            for (int ci = 0; ci < SLOTS && g0 + SLOTS + ci < nc; ++ci) load_in(g0 + SLOTS + ci, nxt[ci]);
          for (int ci = 0; ci < SLOTS && g0 + ci < nc; ++ci) item(g0 + ci, s, ci, cur[ci]);
        }
      }
      if (wt == 0) FT_STAMP(66);
      if (lane == 0) mbar_wait(out_bar, 0);
      __syncwarp();
      tc_fence_after();
      if (wt == 0) FT_STAMP(67);
      drain_out(p.Out);
    }
  }
  if (threadIdx.x == 64) FT_STAMP(68);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// parameter prep: out_{hi,lo}[(z * N + n) * Kp + k] = split(W[n * sn + k * sk + z * sz])  (zero outside n < Nv, k < Kv)
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

}  // namespace ft

// ------------------------------------------------------------------------------------------------
// host API
// ------------------------------------------------------------------------------------------------
void fused_debug_set(void* d_buf) { ft::g_dbg = reinterpret_cast<unsigned long long*>(d_buf); }

bool fused_fwd_supported(const agcn_plan* plan, int F, int Fo, int K) {
  return plan->ft_tiles > 0 && K >= 2 && Fo >= 1 && Fo <= 128 && F >= 1;
}

bool fused_bwd_supported(const agcn_plan* plan, int F, int Fo, int K) {
  return plan->ft_tiles > 0 && K >= 2 && F >= 1 && F <= 128 && Fo >= 1;
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

// backward operand: B_z[f, o] = weight[(f*K + z)*Fo + o]  (f < F rows of dX, o < Fo contraction)
int fused_bwd_prep(const float* weight, int F, int Fo, int K, float* scratch, cudaStream_t st) {
  const int N = ft::pad16(F), Kp = ft::pad32(Fo);
  const long long total = (long long)K * N * Kp;
  const int blocks = (int)std::min<long long>((total + 255) / 256, 1184);
  ft::prep_w_kernel<<<blocks, 256, 0, st>>>(weight, (long long)K * Fo, 1, Fo, F, Fo, N, Kp, K, scratch, scratch + total);
  AGCN_LAUNCH_CHECK();
  return AGCN_OK;
}

static int opt_in_smem(int bytes) {
  static std::mutex mu;
  static int done_bytes = 0;
  std::lock_guard<std::mutex> lock(mu);
  if (bytes > done_bytes) {
    AGCN_CUDA(cudaFuncSetAttribute(ft::fused_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done_bytes = bytes;
  }
  return AGCN_OK;
}

static int launch_tiles(const agcn_plan* plan, int tile0, int ntiles, ft::TileArgs& a, const float* wsplit, int Kp,
                        const char* name, cudaStream_t st) {
  using namespace ft;
  const long long half = (long long)a.K * a.N * Kp;
  CUtensorMap mhi, mlo;
  int rc;
  if ((rc = make_map(&mhi, wsplit, (uint64_t)a.K * a.N, (uint64_t)Kp, (uint32_t)a.N))) return rc;
  if ((rc = make_map(&mlo, wsplit + half, (uint64_t)a.K * a.N, (uint64_t)Kp, (uint32_t)a.N))) return rc;
  a.dbg = g_dbg;
  a.tile_graphs = reinterpret_cast<const int4*>(plan->d_ft_entries);
  a.tile_gstart = plan->d_ft_gstart;
  a.tile0 = tile0;
  a.nchunks = Kp / CH;
  const SmemPlan sp = smem_plan(a.N);
  if ((rc = opt_in_smem(227 * 1024))) return rc;
  {
    ProfScope prof(name, st);
    fused_tile_kernel<<<ntiles, THREADS, sp.total, st>>>(mhi, mlo, a);
  }
  AGCN_LAUNCH_CHECK();
  return AGCN_OK;
}

// T_1..T_{K-1} (saved) and Y = act(sum_k T_k W_k + b) for the tiles [tile0, tile0 + ntiles); L = Lint (add_identity)
// or L_all
int fused_forward(const agcn_plan* plan, int tile0, int ntiles, const float* X, const float* L, int add_identity,
                  const float* wsplit, const float* bias, int act, int F, int Fo, int K, float* T, float* Y,
                  cudaStream_t st) {
  if (ntiles <= 0) return AGCN_OK;
  ft::TileArgs a{};
  a.L = L; a.add_identity = add_identity; a.transL = 0;
  a.Fr = F; a.Fout = Fo; a.K = K; a.N = ft::pad16(Fo);
  a.In = X; a.Mask = nullptr;
  a.Save = T; a.save_slice = (long long)plan->R * F;
  a.bias = bias; a.act = act; a.Out = Y;
  a.G = nullptr; a.g_slice = 0; a.forward = 1;
  return launch_tiles(plan, tile0, ntiles, a, wsplit, ft::pad32(F),
                      tile0 == 0 ? "ft::fused_fwd_kernel" : "ft::fused_fwd_kernel(pre tiles)", st);
}

// dX = sum_z T_z(L^T) dYpre W_z^T, dYpre = dY * [Y > 0] (Y == NULL: dYpre = dY), for whole-graph tiles; for 128-row
// ranges of graphs above AGCN_FUSE_MAX_N, G_z = dYpre W_z^T goes to G instead.  `scratch` ([K][R][Fo], only touched
// when K >= 4) keeps V_s for the three-term recurrence.
int fused_backward(const agcn_plan* plan, int tile0, int ntiles, const float* dY, const float* Y, const float* L,
                   int add_identity, const float* wsplit, int F, int Fo, int K, float* G, float* dX, float* scratch,
                   cudaStream_t st) {
  if (ntiles <= 0) return AGCN_OK;
  ft::TileArgs a{};
  a.L = L; a.add_identity = add_identity; a.transL = 1;
  a.Fr = Fo; a.Fout = F; a.K = K; a.N = ft::pad16(F);
  a.In = dY; a.Mask = Y;
  a.Save = scratch; a.save_slice = (long long)plan->R * Fo;
  a.bias = nullptr; a.act = AGCN_ACT_LINEAR; a.Out = dX;
  a.G = G; a.g_slice = (long long)plan->R * F; a.forward = 0;
  return launch_tiles(plan, tile0, ntiles, a, wsplit, ft::pad32(Fo),
                      tile0 == 0 ? "ft::fused_bwd_kernel" : "ft::fused_bwd_kernel(pre tiles)", st);
}

}  // namespace agcn
