// libmarsgb: one tile of a stable multi-way split, shared by the shuffle's partition kernel (mgb_split.cu) and the
// bucket split of the high-cardinality aggregation (mgb_bucket.cu).
//
// A tile is SP_TILE consecutive rows of a set of 8-byte columns.  Every row has a digit in [0, R), R <= 256.  The
// rows of a digit leave the tile as ONE run per column:
//   A  digit of every row; rank of the row among the rows of its digit (warp match + per-warp counters: stable,
//      no atomics)
//   B  runs: digit d gets count + 1 slots of the on-chip staging row; its run starts on the slot whose parity equals
//      the parity of its first destination row, so that (even, odd) slot pairs are 16-byte aligned pairs in memory
//   C  slot of every row
//   then column by column: the column tile is brought into shared memory by the TMA unit (cp.async.bulk + mbarrier,
//   double buffered: column c + 1 is in flight while column c is handled), permuted into slot order on chip, and
//   written out run by run -- consecutive lanes write consecutive addresses, two rows per 16-byte store.
#pragma once
#include "mgb_common.cuh"

#define SP_THREADS 256
#define SP_WARPS 8
#define SP_ROUNDS 8
#define SP_TILE (SP_THREADS * SP_ROUNDS)
#define SP_MAX_R 256
#define SP_MAX_COLS 80
#define SP_SLOTS (SP_TILE + SP_MAX_R + 2)
#define SP_PAD 0xffffu

// ---- TMA (bulk asynchronous copy) + mbarrier -------------------------------------------------------------------
__device__ __forceinline__ uint32_t sp_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sp_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sp_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sp_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sp_smem(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   sp_smem(dst)),
               "l"(src), "r"(bytes), "r"(sp_smem(bar))
               : "memory");
}
__device__ __forceinline__ void sp_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SP_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SP_DONE;\n"
      "bra SP_WAIT;\n"
      "SP_DONE:\n"
      "}\n" ::"r"(sp_smem(bar)),
      "r"(parity)
      : "memory");
}

// destinations are global memory (this GPU's or an IPC-mapped peer's): typed stores, not generic ones
__device__ __forceinline__ void sp_st128(uint64_t addr, const ulonglong2& v) {
  asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(addr), "l"(v.x), "l"(v.y) : "memory");
}
__device__ __forceinline__ void sp_st64(uint64_t addr, uint64_t v) {
  asm volatile("st.global.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}

// THREADS threads split a tile of 8 * THREADS rows; NBUF = 2: the next column's tile is in flight while this one is
// handled (256 threads, 2048 rows, 3 CTAs per SM); NBUF = 1: one staging tile, refilled while the previous column is
// written out (1024 threads, 8192 rows, 1 CTA per SM: four times longer runs for stores that cross NVLink)
template <int THREADS, int NBUF>
struct SplitSmemT {
  static constexpr int TILE = THREADS * SP_ROUNDS;
  static constexpr int WARPS = THREADS / 32;
  static constexpr int SLOTS = TILE + SP_MAX_R + 2;
  alignas(128) uint64_t a[NBUF][TILE];     // column tiles as they lie in memory (TMA destinations)
  alignas(16) uint64_t b[SLOTS + 2];       // the same rows in partition order, runs padded to destination parity
  alignas(8) uint64_t colbase[SP_MAX_R];   // address of slot 0's row if slot 0 belonged to partition p (this column)
  int64_t gof[SP_MAX_R];                   // first output row of this tile in partition p
  uint64_t bar[2];
  int32_t rsm[SP_MAX_R];                   // first slot of partition p's run
  uint16_t wc[WARPS][SP_MAX_R];            // rows of partition p in the warps before this one
  alignas(4) uint16_t spart[SLOTS + 2];    // partition of the row in slot s (SP_PAD: no row)
  int32_t cnt32[SP_MAX_R];                 // unordered ranking: rows of digit d seen so far
  int32_t wsum[WARPS];
  int32_t nslots;
};
using SplitSmem = SplitSmemT<SP_THREADS, 2>;

// F supplies the tile:
//   uint32_t        digit(int j)        digit of row j (j < nt)
//   int64_t         dest_row(int d)     destination row of the tile's first row with digit d (called by thread d < R)
//   const uint64_t* src(int c)          row 0 of the tile in source column c
//   uint64_t        dst(int c, int d)   address of destination row 0 of column c for digit d
// Called by all SP_THREADS threads of the CTA.  STABLE: rows of a digit keep their input order (what the shuffle's
// blocks promise).  Otherwise the rank of a row is the return value of one shared-memory atomic -- MATCH.ANY costs a
// round per DISTINCT digit in the warp, which at 128-256 ways is most of a tile's time -- and only the order of the
// rows inside a run is unspecified (the bucket split: the fold of a bucket does not depend on it).
template <bool STABLE, int THREADS, int NBUF, class F>
__device__ __forceinline__ void sp_tile_split_t(SplitSmemT<THREADS, NBUF>& sm, const int nt, const int R, const int ncols,
                                                F& f) {
  constexpr int WARPS = THREADS / 32, TILE = THREADS * SP_ROUNDS, SLOTS = TILE + SP_MAX_R + 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bulk_bytes = (uint32_t)(nt & ~1) * 8u;
  if (tid == 0) {
    sp_mbar_init(&sm.bar[0], 1);
    sp_mbar_init(&sm.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < WARPS * SP_MAX_R; i += THREADS) (&sm.wc[0][0])[i] = 0;
  for (int i = tid; i < SLOTS + 2; i += THREADS) sm.spart[i] = (uint16_t)SP_PAD;
  if (tid < SP_MAX_R) sm.cnt32[tid] = 0;
  __syncthreads();
  // a column tile is fetched by the TMA unit when its address is 16-byte aligned
  auto tma_ok = [&](int c) { return ((uintptr_t)f.src(c) & 15) == 0 && bulk_bytes != 0; };
  if (tid == 0 && ncols > 0 && tma_ok(0)) sp_bulk_load(sm.a[0], f.src(0), bulk_bytes, &sm.bar[0]);

  const int64_t g_first = tid < R ? f.dest_row(tid) : 0;   // issued now: its latency overlaps the key loads of phase A
  // ---- A ---------------------------------------------------------------------------------------------------------
  const unsigned lt = mgb_lanemask_lt();
  const int jbase = warp * (TILE / WARPS) + lane;
  uint32_t dr[SP_ROUNDS];   // digit | rank << 16, later the slot of the row
#pragma unroll
  for (int r = 0; r < SP_ROUNDS; ++r) {   // all key loads of the thread in flight together
    const int j = jbase + r * 32;
    dr[r] = j < nt ? f.digit(j) : (uint32_t)(SP_MAX_R + lane);
  }
  if (STABLE) {
#pragma unroll
    for (int r = 0; r < SP_ROUNDS; ++r) {
      const bool valid = jbase + r * 32 < nt;
      const uint32_t d = dr[r];
      const unsigned peers = __match_any_sync(0xffffffffu, d);
      uint32_t before = 0;
      if (valid) before = sm.wc[warp][d];
      __syncwarp();
      if (valid && (peers & lt) == 0) sm.wc[warp][d] = (uint16_t)(before + __popc(peers));
      __syncwarp();
      dr[r] = d | ((before + __popc(peers & lt)) << 16);
    }
  } else {
#pragma unroll
    for (int r = 0; r < SP_ROUNDS; ++r)
      if (jbase + r * 32 < nt) dr[r] |= (uint32_t)atomicAdd(&sm.cnt32[dr[r]], 1) << 16;
  }
  __syncthreads();
  // ---- B ---------------------------------------------------------------------------------------------------------
  {
    int cnt = 0;
    int64_t g = 0;
    if (tid < R) {
      if (STABLE) {
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
          const int t = sm.wc[w][tid];
          sm.wc[w][tid] = (uint16_t)cnt;
          cnt += t;
        }
      } else {
        cnt = sm.cnt32[tid];
      }
      g = g_first;
    }
    int v = tid < R ? cnt + 1 : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) sm.wsum[warp] = inc;
    __syncthreads();
    int add = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w)
      if (w < warp) add += sm.wsum[w];
    const int excl = add + inc - v;
    if (tid < R) {
      const int rs = excl + (int)((excl ^ (int)g) & 1);
      sm.rsm[tid] = rs;
      sm.gof[tid] = g;
    }
    if (tid == THREADS - 1) sm.nslots = (add + inc + 2) & ~1;
  }
  __syncthreads();
  // ---- C ---------------------------------------------------------------------------------------------------------
#pragma unroll
  for (int r = 0; r < SP_ROUNDS; ++r) {
    const int j = jbase + r * 32;
    if (j < nt) {
      const uint32_t d = dr[r] & 0xffffu;
      const uint32_t slot = (uint32_t)sm.rsm[d] + (STABLE ? (uint32_t)sm.wc[warp][d] : 0u) + (dr[r] >> 16);
      sm.spart[slot] = (uint16_t)d;
      dr[r] = slot;
    }
  }
  // ---- columns ---------------------------------------------------------------------------------------------------
  uint32_t uses0 = 0, uses1 = 0;
  for (int c = 0; c < ncols; ++c) {
    const int buf = NBUF == 2 ? (c & 1) : 0;
    if (tid < R) sm.colbase[tid] = f.dst(c, tid) + 8ull * (uint64_t)(sm.gof[tid] - (int64_t)sm.rsm[tid]);
    if (NBUF == 2 && tid == 0 && c + 1 < ncols && tma_ok(c + 1))
      sp_bulk_load(sm.a[buf ^ (NBUF - 1)], f.src(c + 1), bulk_bytes, &sm.bar[buf ^ (NBUF - 1)]);
    const uint64_t* src = f.src(c);
    if (tma_ok(c)) {
      const uint32_t use = buf ? uses1++ : uses0++;
      sp_mbar_wait(&sm.bar[buf], use & 1);
#pragma unroll
      for (int r = 0; r < SP_ROUNDS; ++r) {
        const int j = jbase + r * 32;
        if (j < nt) sm.b[dr[r]] = (j == nt - 1 && (nt & 1)) ? mgb_ld_stream(src + j) : sm.a[buf][j];
      }
    } else {
#pragma unroll
      for (int r = 0; r < SP_ROUNDS; ++r) {
        const int j = jbase + r * 32;
        if (j < nt) sm.b[dr[r]] = mgb_ld_stream(src + j);
      }
    }
    __syncthreads();
    // one staging tile: it is free now -- the next column's tile arrives while this one is written out
    if (NBUF == 1 && tid == 0 && c + 1 < ncols && tma_ok(c + 1)) sp_bulk_load(sm.a[0], f.src(c + 1), bulk_bytes, &sm.bar[0]);
    const int npairs = sm.nslots >> 1;
    for (int q = tid; q < npairs; q += THREADS) {
      const uint32_t pp = *reinterpret_cast<const uint32_t*>(&sm.spart[2 * q]);
      const uint32_t p0 = pp & 0xffffu, p1 = pp >> 16;
      if (pp == 0xffffffffu) continue;
      const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(&sm.b[2 * q]);
      const uint64_t a0 = p0 != SP_PAD ? sm.colbase[p0] + 16ull * (uint64_t)q : 0;
      if (p0 == p1 && (a0 & 15) == 0) {
        sp_st128(a0, v);
      } else {
        if (p0 != SP_PAD) sp_st64(a0, v.x);
        if (p1 != SP_PAD) sp_st64(sm.colbase[p1] + 16ull * (uint64_t)q + 8, v.y);
      }
    }
    __syncthreads();
  }
}

// the default geometry: 256 threads, 2048-row tiles, double-buffered staging
template <bool STABLE, class F>
__device__ __forceinline__ void sp_tile_split(SplitSmem& sm, const int nt, const int R, const int ncols, F& f) {
  sp_tile_split_t<STABLE, SP_THREADS, 2>(sm, nt, R, ncols, f);
}
