// libmarsgb: map-side pre-aggregation for bands with at most TN_G (4) groups per CTA -- the TPC-H Q1 shape -- a4,
// DataFrameGroupByAgg._execute_map (mars/dataframe/groupby/aggregation.py:1043-1128), sum / mean / count over
// float64 columns.
//
// k_preagg_dense keeps its accumulators in shared memory, privatised per warp and per lane (no atomics), and streams
// the rows through registers: the register file then bounds the bytes in flight (57 KB per SM, about what HBM latency
// asks for at 5 TB/s, not more), and shared memory has no room left for staging.  With four groups the accumulators
// fit in REGISTERS instead -- per thread 4 x (NV sums + NV NaN counters + rows), updated under a branch on the group
// id -- and all of shared memory becomes a three-stage ring of column tiles filled by the TMA unit (cp.async.bulk +
// mbarrier): two tiles (2 x 56 KB for Q1) are in flight per SM while the third is consumed.
//
//   thread 0        before it consumes tile i: wait for the stage of tile i + 2 to be empty (every warp releases a stage
//                   as soon as the tile is in its registers), arm that stage's "full" barrier with the tile's bytes,
//                   issue one bulk copy per column (8 KB each)
//   16 warps        wait "full", two rows per thread from shared memory (conflict-free 8-byte reads), arrive on
//                   "empty"; group id by comparing with the CTA's <= 4 known key tuples (a new tuple is appended under
//                   a lock by lane 0 of the warp that met it), accumulate
//   epilogue        warp shuffles -> shared memory -> one staged row per (CTA, group); atomic ticket; the last CTA
//                   merges the staged rows of all CTAs (<= 16 distinct tuples in total) and writes the result block
//
// A CTA that meets a fifth key tuple raises the fail flag: the call then reports -1 through the SAME device-side
// count, and the dense kernel that the host enqueued right behind this one runs (it exits at once otherwise).  A
// failure is remembered per spec on the device, so that the following launches of that shape skip this kernel.
#include "mgb_dense.cuh"
#include "mgb_split.cuh"   // sp_mbar_init / sp_bulk_load / sp_mbar_wait

#define TN_G 4
#define TN_CONSUMERS 512
#define TN_THREADS TN_CONSUMERS      // thread 0 also issues the bulk copies (two tiles ahead)
#define TN_STAGES 3
#define TN_TILE 1024
#define TN_MAX_TOTAL 16

struct TinyParams {
  DenseProgram prog;        // PROG instantiations: row filter + derived columns (TPC-H Q1 pre-path), as in k_preagg_dense
  uint64_t* stg;            // [grid][TN_G][NK + 1 + 2 NV]: keys | rows | sums | NaN counts
  int* cta_ng;              // [grid]
  unsigned int* ticket;
  int* fail;                // 1: not a shape for this kernel (the dense kernel behind it takes over)
  int* sticky;              // per spec, survives the call: set on failure, checked at kernel start
  OutCols out;
  uint8_t out_col[MGB_MAX_ACCS];
  uint8_t out_src[MGB_MAX_ACCS];
  long long* out_n;
};

__device__ __forceinline__ void tn_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sp_smem(bar)) : "memory");
}

template <int NK, int NV>
struct TinySmem {
  alignas(128) uint64_t tile[TN_STAGES][NK + NV][TN_TILE];
  uint64_t full[TN_STAGES], empty[TN_STAGES];
  uint64_t gkeys[MGB_MAX_KEYS][TN_G];
  int ng, lock, last;
  uint32_t nan[TN_G][NV];      // NaN values per (group, column): rare, counted with shared-memory atomics
  alignas(8) uint64_t red[TN_CONSUMERS / 32][TN_G][1 + NV];
};

// which chunk a virtual tile belongs to (tiles ascend per CTA: the cursor only moves forward)
__device__ __forceinline__ void tn_locate(const DenseBatch& bt, long long v, int& q, long long& row0) {
  while (q < bt.n_chunks - 1 && bt.tile_end[q] <= v) ++q;
  row0 = (v - (q ? bt.tile_end[q - 1] : 0)) * TN_TILE;
}

// PROG: the tile holds the n_phys PHYSICAL columns of the row program; the NV logical columns a row contributes are
// computed from them in registers (products of affine factors, each operation rounded on its own like the chain of
// pandas operators it replaces), rows the filter rejects take no part in anything
template <int NK, int NV, bool PROG>
__global__ void __launch_bounds__(TN_THREADS, 1) k_preagg_tiny(const TinyParams p, const DenseBatch bt) {
  extern __shared__ __align__(128) unsigned char tn_raw[];
  TinySmem<NK, NV>& sm = *reinterpret_cast<TinySmem<NK, NV>*>(tn_raw);
  constexpr int W = NK + 1 + 2 * NV;   // words of a staged row
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grid = gridDim.x, b = blockIdx.x;
  if (*(volatile const int*)p.sticky) {   // this shape failed before: leave it to the dense kernel
    if (b == 0 && tid == 0) {
      *p.fail = 1;
      *p.out_n = -1;
    }
    return;
  }
  if (tid == 0) {
    for (int s = 0; s < TN_STAGES; ++s) {
      sp_mbar_init(&sm.full[s], 1);
      sp_mbar_init(&sm.empty[s], TN_CONSUMERS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    sm.ng = 0;
    sm.lock = 0;
    sm.last = 0;
  }
  if (tid < TN_G * NV) (&sm.nan[0][0])[tid] = 0;
  __syncthreads();
  const long long total_tiles = bt.tile_end[bt.n_chunks - 1];
  const int n_load = PROG ? p.prog.n_phys : NV;      // value columns a tile carries

  double sum[TN_G][NV];
  uint32_t rows[TN_G];
#pragma unroll
  for (int g = 0; g < TN_G; ++g) {
    rows[g] = 0;
#pragma unroll
    for (int c = 0; c < NV; ++c) sum[g][c] = 0.0;
  }
  bool failed = false;
  // the CTA's known key tuples, cached in registers; refreshed when the shared list has grown
  uint64_t gk0[TN_G], gk1[TN_G];
  int ng_reg = 0;
#pragma unroll
  for (int g = 0; g < TN_G; ++g) gk0[g] = gk1[g] = 0;
  auto refresh = [&]() {
    const int ng = *(volatile int*)&sm.ng;
    if (ng != ng_reg) {
#pragma unroll
      for (int g = 0; g < TN_G; ++g) {
        if (g < ng) {
          gk0[g] = ((volatile uint64_t*)sm.gkeys[0])[g];
          if (NK == 2) gk1[g] = ((volatile uint64_t*)sm.gkeys[1])[g];
        }
      }
      ng_reg = ng;
    }
  };
  // group id of a row among the cached tuples, -1: unknown
  auto take_row = [&](const uint64_t* k) {
    int gid = -1;
#pragma unroll
    for (int g = 0; g < TN_G; ++g) {
      bool eq = g < ng_reg && gk0[g] == k[0];
      if (NK == 2) eq = eq && gk1[g] == k[1];
      if (eq) gid = g;
    }
    return gid;
  };
  // NaN values are taken out of the row once (they add 0 and are counted), then the group's registers are updated
  auto update = [&](int gid, const uint64_t* x) {
    double v[NV];
    uint32_t nanm = 0;
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      v[c] = __longlong_as_double((long long)x[c]);
      if (v[c] != v[c]) {
        nanm |= 1u << c;
        v[c] = 0.0;
      }
    }
    if (nanm && gid >= 0) {
#pragma unroll
      for (int c = 0; c < NV; ++c)
        if ((nanm >> c) & 1u) atomicAdd(&sm.nan[gid][c], 1u);
    }
#pragma unroll
    for (int g = 0; g < TN_G; ++g) {
      if (gid == g) {
        ++rows[g];
#pragma unroll
        for (int c = 0; c < NV; ++c) sum[g][c] += v[c];
      }
    }
  };
  // row program: physical values in x[] -> logical values in x[]; returns whether the row passes the filter
  auto apply_program = [&](uint64_t* x) -> bool {
    double ph[NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) ph[c] = __longlong_as_double((long long)x[c]);
    bool keep = true;
    if (p.prog.filter_col >= 0) {
      uint64_t raw = 0;
#pragma unroll
      for (int c = 0; c < NV; ++c)
        if (p.prog.filter_col == c) raw = x[c];
      int cmp;   // NaN compares as "unordered": every comparison but != fails, like pandas
      bool unordered = false;
      if (p.prog.filter_f64) {
        const double v = __longlong_as_double((long long)raw);
        unordered = v != v;
        cmp = v < p.prog.filter_d ? -1 : (v > p.prog.filter_d ? 1 : 0);
      } else {
        const long long v = (long long)raw;
        cmp = v < p.prog.filter_i ? -1 : (v > p.prog.filter_i ? 1 : 0);
      }
      const int op = p.prog.filter_op;
      keep = op == 0 ? cmp <= 0 : op == 1 ? cmp < 0 : op == 2 ? cmp >= 0 : op == 3 ? cmp > 0 : op == 4 ? cmp == 0 : cmp != 0;
      if (unordered) keep = op == 5;
    }
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      const int nf = p.prog.nf[c];
      if (nf == 0) continue;            // logical column c == physical column c
      double v = 1.0;
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        if (f < nf) {
          double o = 0.0;
#pragma unroll
          for (int j = 0; j < NV; ++j)
            if (p.prog.fcol[c][f] == j) o = ph[j];
          const double term = __dadd_rn(p.prog.falpha[c][f], __dmul_rn(p.prog.fbeta[c][f], o));
          v = f == 0 ? term : __dmul_rn(v, term);
        }
      }
      x[c] = (uint64_t)__double_as_longlong(v);
    }
    return keep;
  };
  // warp-collective: lanes with live && gid < 0 hold key tuples the CTA does not know yet
  auto slow = [&](bool live, int gid, const uint64_t* k) {
    unsigned need = __ballot_sync(0xffffffffu, live && gid < 0);
    while (need) {
      const int src = __ffs(need) - 1;
      const uint64_t s0 = __shfl_sync(0xffffffffu, k[0], src);
      const uint64_t s1 = __shfl_sync(0xffffffffu, NK == 2 ? k[1] : 0ULL, src);
      int gi = -1;
      if (lane == 0) {
        while (atomicCAS(&sm.lock, 0, 1) != 0) {
        }
        __threadfence_block();
        const int ng = *(volatile int*)&sm.ng;
        for (int g = 0; g < ng; ++g) {
          bool eq = ((volatile uint64_t*)sm.gkeys[0])[g] == s0;
          if (NK == 2) eq = eq && ((volatile uint64_t*)sm.gkeys[1])[g] == s1;
          if (eq) gi = g;
        }
        if (gi < 0) {
          if (ng < TN_G) {
            ((volatile uint64_t*)sm.gkeys[0])[ng] = s0;
            if (NK == 2) ((volatile uint64_t*)sm.gkeys[1])[ng] = s1;
            __threadfence_block();
            *(volatile int*)&sm.ng = ng + 1;
            gi = ng;
          } else {
            gi = -2;
          }
        }
        __threadfence_block();
        atomicExch(&sm.lock, 0);
      }
      gi = __shfl_sync(0xffffffffu, gi, 0);
      bool mine = live && gid < 0 && k[0] == s0;
      if (NK == 2) mine = mine && k[1] == s1;
      if (mine) gid = gi == -2 ? -3 : gi;      // -3: a fifth tuple -- the row is dropped, the call fails
      if (gi == -2) failed = true;
      need = __ballot_sync(0xffffffffu, live && gid == -1);
    }
    refresh();
    return gid;
  };

  {
    // issue the bulk copies of the i-th tile of this CTA (virtual tile v)
    int pq = 0;
    auto issue = [&](long long i, long long v) {
      const int s = (int)(i % TN_STAGES);
      if (i >= TN_STAGES) sp_mbar_wait(&sm.empty[s], (uint32_t)((i / TN_STAGES - 1) & 1));
      long long row0;
      tn_locate(bt, v, pq, row0);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sp_smem(&sm.full[s])),
                   "r"((uint32_t)((NK + n_load) * TN_TILE * 8))
                   : "memory");
#pragma unroll
      for (int c = 0; c < NK + NV; ++c) {
        if (c >= NK + n_load) break;
        const uint64_t* src = bt.ptr[pq][c < NK ? c : MGB_MAX_KEYS + (c - NK)] + row0;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         sp_smem(&sm.tile[s][c][0])),
                     "l"(src), "r"((uint32_t)(TN_TILE * 8)), "r"(sp_smem(&sm.full[s]))
                     : "memory");
      }
    };
    if (tid == 0) {
      for (int a = 0; a < TN_STAGES - 1; ++a)
        if (b + (long long)a * grid < total_tiles) issue(a, b + (long long)a * grid);
    }
    long long i = 0;
    for (long long v = b; v < total_tiles; v += grid, ++i) {
      if (tid == 0) {
        const long long va = v + (long long)(TN_STAGES - 1) * grid;
        if (va < total_tiles) issue(i + TN_STAGES - 1, va);
      }
      const int s = (int)(i % TN_STAGES);
      sp_mbar_wait(&sm.full[s], (uint32_t)((i / TN_STAGES) & 1));
      refresh();
      uint64_t k[2][2], x[2][NV];
      int gid[2];
      {   // rows 2 tid and 2 tid + 1 of the tile: one 16-byte read per column
#pragma unroll
        for (int q = 0; q < NK; ++q) {
          const ulonglong2 t2 = reinterpret_cast<const ulonglong2*>(&sm.tile[s][q][0])[tid];
          k[0][q] = t2.x;
          k[1][q] = t2.y;
        }
        if (NK == 1) k[0][1] = k[1][1] = 0;
#pragma unroll
        for (int c = 0; c < NV; ++c) {
          ulonglong2 t2 = make_ulonglong2(0, 0);
          if (c < n_load) t2 = reinterpret_cast<const ulonglong2*>(&sm.tile[s][NK + c][0])[tid];
          x[0][c] = t2.x;
          x[1][c] = t2.y;
        }
      }
      __syncwarp();
      if (lane == 0) tn_mbar_arrive(&sm.empty[s]);   // the tile is in registers: the stage may be refilled
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const bool live = PROG ? apply_program(x[r]) : true;
        gid[r] = live ? take_row(k[r]) : -4;
        if (__any_sync(0xffffffffu, live && gid[r] < 0)) gid[r] = slow(live, gid[r], k[r]);
        if (live) update(gid[r], x[r]);
      }
    }
    // remainders of the chunks (< 1024 rows each): chunk q by CTA q % grid, plain loads
    for (int q = b; q < bt.n_chunks; q += grid) {
      const long long n = bt.n[q], r0 = (n / TN_TILE) * TN_TILE;
      for (long long j0 = r0; j0 < n; j0 += TN_CONSUMERS) {   // uniform trip count per warp: slow() is collective
        const long long j = j0 + tid;
        bool live = j < n;
        uint64_t k[2] = {0, 0}, x[NV];
#pragma unroll
        for (int c = 0; c < NV; ++c) x[c] = 0;
        if (live) {
#pragma unroll
          for (int qq = 0; qq < NK; ++qq) k[qq] = mgb_ld_stream(bt.ptr[q][qq] + j);
#pragma unroll
          for (int c = 0; c < NV; ++c)
            if (c < n_load) x[c] = mgb_ld_stream(bt.ptr[q][MGB_MAX_KEYS + c] + j);
          if (PROG) live = apply_program(x);
        }
        refresh();
        int gid = live ? take_row(k) : -4;
        if (__any_sync(0xffffffffu, live && gid < 0)) gid = slow(live, gid, k);
        if (live) update(gid, x);
      }
    }
  }
  // ---- epilogue: lanes -> warp -> CTA -> staged rows ------------------------------------------------------------
  {
#pragma unroll
    for (int g = 0; g < TN_G; ++g) {
      uint32_t r = rows[g];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
      if (lane == 0) sm.red[warp][g][0] = r;
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        double s_ = sum[g][c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s_ += __shfl_xor_sync(0xffffffffu, s_, o);
        if (lane == 0) sm.red[warp][g][1 + c] = (uint64_t)__double_as_longlong(s_);
      }
    }
    if (__any_sync(0xffffffffu, failed) && lane == 0) atomicExch(p.fail, 1);
  }
  __syncthreads();
  const int ng = sm.ng;
  for (int idx = tid; idx < TN_G * (1 + 2 * NV); idx += TN_THREADS) {
    const int g = idx / (1 + 2 * NV), w = idx - g * (1 + 2 * NV);
    uint64_t* dst = p.stg + ((long long)b * TN_G + g) * W;
    if (w >= 1 && w <= NV) {
      double a = 0.0;
      for (int q = 0; q < TN_CONSUMERS / 32; ++q) a += __longlong_as_double((long long)sm.red[q][g][w]);
      dst[NK + w] = (uint64_t)__double_as_longlong(a);
    } else if (w == 0) {
      uint64_t a = 0;
      for (int q = 0; q < TN_CONSUMERS / 32; ++q) a += sm.red[q][g][0];
      dst[NK] = a;
      dst[0] = sm.gkeys[0][g];
      if (NK == 2) dst[1] = sm.gkeys[1][g];
    } else {
      dst[NK + w] = sm.nan[g][w - 1 - NV];
    }
  }
  if (tid == 0) p.cta_ng[b] = ng;
  __threadfence();
  __syncthreads();
  if (tid == 0) sm.last = atomicAdd(p.ticket, 1u) == (unsigned)(grid - 1);
  __syncthreads();
  if (!sm.last) return;
  // ---- last CTA: merge the staged rows of all CTAs ----------------------------------------------------------------
  __threadfence();
  // shared memory is free now: reuse the tile ring
  int* map = reinterpret_cast<int*>(&sm.tile[0][0][0]);             // [grid * TN_G] group of every staged row, -1 unused
  int* pick = map + 1024 * TN_G;                                     // [1] lowest unassigned staged row
  uint64_t* okeys = reinterpret_cast<uint64_t*>(pick + 8);           // [2][TN_MAX_TOTAL]
  uint64_t* res = okeys + 2 * TN_MAX_TOTAL;                          // [TN_MAX_TOTAL][1 + 2 NV]
  const int n_stg = grid * TN_G;
  for (int r = tid; r < n_stg; r += TN_THREADS) map[r] = (r % TN_G) < __ldcg(p.cta_ng + r / TN_G) ? -1 : -2;
  __syncthreads();
  int G = 0;
  bool too_many = false;
  for (;;) {   // one distinct key tuple per round, in order of its first staged row (deterministic)
    if (tid == 0) *pick = 0x7fffffff;
    __syncthreads();
    for (int r = tid; r < n_stg; r += TN_THREADS)
      if (map[r] == -1) atomicMin(pick, r);
    __syncthreads();
    const int first = *pick;
    if (first == 0x7fffffff) break;
    if (G == TN_MAX_TOTAL) {
      too_many = true;
      break;
    }
    const uint64_t f0 = ld_cg(p.stg + (long long)first * W), f1 = NK == 2 ? ld_cg(p.stg + (long long)first * W + 1) : 0;
    for (int r = tid; r < n_stg; r += TN_THREADS) {
      if (map[r] != -1) continue;
      bool eq = ld_cg(p.stg + (long long)r * W) == f0;
      if (NK == 2) eq = eq && ld_cg(p.stg + (long long)r * W + 1) == f1;
      if (eq) map[r] = G;
    }
    if (tid == 0) {
      okeys[G] = f0;
      okeys[TN_MAX_TOTAL + G] = f1;
    }
    ++G;
    __syncthreads();
  }
  const bool bad = too_many || *(volatile int*)p.fail != 0;
  if (bad) {
    if (tid == 0) {
      *p.fail = 1;
      *p.sticky = 1;
      *p.out_n = -1;
    }
    return;
  }
  // one warp per (output group, word): staged rows in fixed order
  for (int item = warp; item < G * (1 + 2 * NV); item += TN_THREADS / 32) {
    const int g = item / (1 + 2 * NV), w = item - g * (1 + 2 * NV);
    const bool f = w >= 1 && w <= NV;
    double fa = 0.0;
    uint64_t ia = 0;
    for (int r = lane; r < n_stg; r += 32) {
      if (map[r] != g) continue;
      const uint64_t v = ld_cg(p.stg + (long long)r * W + NK + w);
      if (f)
        fa += __longlong_as_double((long long)v);
      else
        ia += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      fa += __shfl_xor_sync(0xffffffffu, fa, o);
      ia += __shfl_xor_sync(0xffffffffu, ia, o);
    }
    if (lane == 0) res[g * (1 + 2 * NV) + w] = f ? (uint64_t)__double_as_longlong(fa) : ia;
  }
  __syncthreads();
  for (int idx = tid; idx < G * (NK + p.out.na); idx += TN_THREADS) {
    const int g = idx / (NK + p.out.na), a = idx - g * (NK + p.out.na);
    if (a < NK) {
      p.out.key[a][g] = okeys[a * TN_MAX_TOTAL + g];
      continue;
    }
    const int acc = a - NK, col = p.out_col[acc];
    const uint64_t* rr = res + g * (1 + 2 * NV);
    uint64_t v;
    if (p.out_src[acc] == MGB_COUNT)
      v = col == 0xFF ? rr[0] : rr[0] - rr[1 + NV + col];
    else
      v = rr[1 + col];
    p.out.acc[acc][g] = v;
  }
  if (tid == 0) *p.out_n = G;
}

template <int NK, int NV, bool PROG>
static int tiny_launch(const TinyParams& p, const DenseBatch& bt, int grid, cudaStream_t st) {
  static bool attr_set[8] = {false};
  int dev = 0;
  MGB_CUDA(cudaGetDevice(&dev));
  const size_t smem = sizeof(TinySmem<NK, NV>);
  if (!attr_set[dev & 7]) {
    MGB_CUDA(cudaFuncSetAttribute(k_preagg_tiny<NK, NV, PROG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[dev & 7] = true;
  }
  k_preagg_tiny<NK, NV, PROG><<<grid, TN_THREADS, smem, st>>>(p, bt);
  MGB_CHECK_LAUNCH();
  return MGB_OK;
}

// eligible: sum / count shapes over <= 6 float64 (logical) columns, every loaded column 16-byte aligned (bulk copies);
// n_load: the columns a tile carries (== ncol, or the physical columns of a row program)
bool mgb_tiny_eligible(const DenseBatch& bt, int nk, int ncol, int n_load) {
  if (ncol < 1 || ncol > 6 || n_load < 1 || n_load > ncol) return false;
  for (int q = 0; q < bt.n_chunks; ++q) {
    for (int j = 0; j < nk; ++j)
      if ((uintptr_t)bt.ptr[q][j] & 15) return false;
    for (int c = 0; c < n_load; ++c)
      if ((uintptr_t)bt.ptr[q][MGB_MAX_KEYS + c] & 15) return false;
  }
  return true;
}

int mgb_tiny_launch(const DenseParams& dp, const DenseBatch& bt, int nk, int ncol, int grid, int* ticket_fail2,
                    int* sticky, cudaStream_t st) {
  TinyParams p;
  memset(&p, 0, sizeof(p));
  p.stg = dp.stg;
  p.cta_ng = dp.cta_ng;
  p.ticket = (unsigned int*)ticket_fail2;
  p.fail = ticket_fail2 + 1;
  p.sticky = sticky;
  p.out = dp.out;
  memcpy(p.out_col, dp.out_col, sizeof(p.out_col));
  memcpy(p.out_src, dp.out_src, sizeof(p.out_src));
  p.out_n = dp.out_n;
  p.prog = dp.prog;
#define TN_PICK(NK_, PROG_)                                          \
  switch (ncol) {                                                    \
    case 1: return tiny_launch<NK_, 1, PROG_>(p, bt, grid, st);      \
    case 2: return tiny_launch<NK_, 2, PROG_>(p, bt, grid, st);      \
    case 3: return tiny_launch<NK_, 3, PROG_>(p, bt, grid, st);      \
    case 4: return tiny_launch<NK_, 4, PROG_>(p, bt, grid, st);      \
    case 5: return tiny_launch<NK_, 5, PROG_>(p, bt, grid, st);      \
    default: return tiny_launch<NK_, 6, PROG_>(p, bt, grid, st);     \
  }
  if (dp.prog.n_phys > 0) {
    if (nk == 1) {
      TN_PICK(1, true)
    }
    TN_PICK(2, true)
  }
  if (nk == 1) {
    TN_PICK(1, false)
  }
  TN_PICK(2, false)
#undef TN_PICK
}
