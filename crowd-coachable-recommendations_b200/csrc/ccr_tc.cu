// Tensor-core score + select kernel for sm_100a.
//
//   scores[128 queries x 256 items] per tile = Q_tile[128 x D] . Items_tile[256 x D]^T
//
// * operands staged by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) into a 4-stage
//   shared-memory ring of (16 KB Q block + 32 KB item block), K advanced 64 elements a stage;
// * one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (bf16 x bf16 -> fp32,
//   M=128, N=256, K=16) into one of two 256-column TMEM accumulators;
// * four epilogue warps read the other accumulator with tcgen05.ld (thread == query row),
//   compare against the row's running threshold and append survivors to the row's candidate
//   buffer; a nearly full buffer is pruned in place to its exact top-k (warp radix select),
//   which raises the threshold.  The B x N score matrix never leaves the SM.
//
// Persistent: grid = #SMs, each CTA walks units (query tile, item split) round-robin with the
// query tile varying fastest so CTAs running concurrently share item tiles through L2.
#include <cuda.h>
#include <stdlib.h>

#include "ccr_params.cuh"

namespace ccr {

constexpr int kStages = 4;
constexpr int kBytesA = kQTile * kKBlock * 2;   // 16384
constexpr int kBytesB = kITile * kKBlock * 2;   // 32768
constexpr int kStageBytes = kBytesA + kBytesB;  // 49152
constexpr int kEpiWarps = 8;                    // 2 per TMEM lane quadrant, one per column half
constexpr int kTcThreads = 64 + 32 * kEpiWarps; // warp0 TMA, warp1 MMA, warps 2..9 epilogue
constexpr int kStageKeys = 256;                 // candidate buffers up to this size are pruned in smem
constexpr int kTmemCols = 512;
constexpr unsigned long long kWaitLimitNs = 10ull * 1000ull * 1000ull * 1000ull;  // 10 s

struct __align__(8) TcShared {
  u64 full[kStages];
  u64 empty[kStages];
  u64 tmem_full[2];
  u64 tmem_empty[2];
  u32 tmem_base;
  u32 pad;
  u32 hist[kEpiWarps][256];
  u64 stage[kEpiWarps][kStageKeys];
};
constexpr size_t kTcSmemBytes = (size_t)kStages * kStageBytes + sizeof(TcShared) + 1024;

// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(u64* bar, u32 parity) {
  u32 ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a pipeline bug must end in a trap + error record, never in a hung GPU
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity, DeviceStatus* st, int where) {
  u32 spins = 0;
  unsigned long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > kWaitLimitNs) {
        if (st) { st->code = 1; st->where = where; st->block = blockIdx.x; st->extra = (int)parity; }
        __threadfence_system();
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, u64* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(u64* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(u32 tmem_d, u64 adesc, u64 bdesc, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(u32 taddr, u32 (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile (rows x 64 bf16, row pitch 128 B, 8-row atoms 1024 B
// apart).  Field layout per the sm_100 shared-memory matrix descriptor: start address [0,14)
// (>>4), LBO [16,30) (ignored for swizzled K-major; 1 like CUTLASS), SBO [32,46) = 1024>>4,
// version [46,48) = 1, layout [61,64) = 2 (SWIZZLE_128B).
__device__ __forceinline__ u64 make_sw128_desc(u32 saddr) {
  u64 d = 0;
  d |= (u64)((saddr & 0x3FFFFu) >> 4);
  d |= (u64)1 << 16;
  d |= (u64)(1024 >> 4) << 32;
  d |= (u64)1 << 46;
  d |= (u64)2 << 61;
  return d;
}
// instruction descriptor, kind::f16: c_format f32 (bit 4), a/b format bf16 (bits 7, 10),
// both K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
constexpr u32 kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((u32)(kITile >> 3) << 17) | ((u32)(kQTile >> 4) << 24);

// ---------------------------------------------------------------------------------------
// per-thread selection state and the per-chunk filter (32 columns of one query row)
// ---------------------------------------------------------------------------------------
struct SelState {
  u64* buf;        // this (row, split, half)'s candidate buffer
  int cnt;
  float tau_f;     // score of the current k-th best (or -inf / +inf for padding rows)
  u64 tau_key;
  long long mbeg, mend;  // the row's slice of the mask CSR
};

__device__ __forceinline__ float max8(const u32 (&v)[32], int g) {
  const float a = fmaxf(fmaxf(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1])), __uint_as_float(v[g * 8 + 2]));
  const float b = fmaxf(fmaxf(__uint_as_float(v[g * 8 + 3]), __uint_as_float(v[g * 8 + 4])), __uint_as_float(v[g * 8 + 5]));
  const float c = fmaxf(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
  return fmaxf(fmaxf(a, b), c);
}

__device__ __forceinline__ void select_chunk(const u32 (&v)[32], long long col0, SelState& st, const SelectParams& p,
                                             int k, int C, u32 hist_s, u32 stage_s) {
  const int lane = threadIdx.x & 31;
  if (p.debug & 1) return;
  float m8[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) m8[g] = max8(v, g);
  const float m = fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3]));
  if (!__any_sync(0xffffffffu, m >= st.tau_f)) return;
  // slow path: some row of this warp has a candidate in these 32 columns
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (m8[g] >= st.tau_f) {
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float s = __uint_as_float(v[g * 8 + jj]);
        if (s >= st.tau_f) {
          const long long col = col0 + g * 8 + jj;
          if (col < p.n_items)
            st.cnt = cand_insert(s, (u32)col, st.cnt, st.tau_key, st.buf, p.mask_cols, st.mbeg, st.mend);
        }
      }
    }
  }
  unsigned need = __ballot_sync(0xffffffffu, st.cnt > C - 32);
  while (need) {
    const int src = __ffs(need) - 1;
    need &= need - 1;
    u64* b = reinterpret_cast<u64*>(__shfl_sync(0xffffffffu, (u64)(uintptr_t)st.buf, src));
    const int n = __shfl_sync(0xffffffffu, st.cnt, src);
    const u64 pivot = warp_prune(b, n, k, hist_s, stage_s);
    if (lane == src) { st.cnt = k; st.tau_key = pivot; st.tau_f = key_score(pivot); }
  }
}

// ---------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads, 1)
select_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_items,
                 SelectParams p) {
  extern __shared__ unsigned char smem_dyn[];
  // 1024-byte alignment for the swizzled tiles
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
  TcShared* sh = reinterpret_cast<TcShared*>(smem + (size_t)kStages * kStageBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (p.D + kKBlock - 1) / kKBlock;
  const long long tiles_total = (p.n_items + kITile - 1) / kITile;
  const int n_units = p.n_q_tiles * p.S;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&sh->full[s], 1); mbar_init(&sh->empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&sh->tmem_full[s], 1); mbar_init(&sh->tmem_empty[s], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_items) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)),
                 "r"((u32)kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const u32 tmem_base = sh->tmem_base;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0; u32 phase = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int qt = unit % p.n_q_tiles, u = unit / p.n_q_tiles;
        const long long t0 = (long long)u * tiles_total / p.S, t1 = (long long)(u + 1) * tiles_total / p.S;
        for (long long t = t0; t < t1; ++t) {
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&sh->empty[stage], phase ^ 1u, p.status, 100 + stage);
            unsigned char* sa = smem + (size_t)stage * kStageBytes;
            mbar_expect_tx(&sh->full[stage], kStageBytes);
            tma_load_2d(sa, &tmap_q, kb * kKBlock, qt * kQTile, &sh->full[stage]);
            tma_load_2d(sa + kBytesA, &tmap_items, kb * kKBlock, (int)(t * kITile), &sh->full[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      int stage = 0; u32 phase = 0;
      int acc = 0; u32 acc_phase = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int u = unit / p.n_q_tiles;
        const long long t0 = (long long)u * tiles_total / p.S, t1 = (long long)(u + 1) * tiles_total / p.S;
        for (long long t = t0; t < t1; ++t) {
          mbar_wait(&sh->tmem_empty[acc], acc_phase ^ 1u, p.status, 200 + acc);
          tc_fence_after();
          const u32 d_tmem = tmem_base + (u32)(acc * kITile);
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&sh->full[stage], phase, p.status, 300 + stage);
            tc_fence_after();
            const u32 sa = smem_u32(smem + (size_t)stage * kStageBytes);
            const u64 adesc = make_sw128_desc(sa);
            const u64 bdesc = make_sw128_desc(sa + kBytesA);
#pragma unroll
            for (int kk = 0; kk < kKBlock / 16; ++kk) {
              // +32 bytes (16 bf16) along K inside the 128-byte swizzle row: +2 in 16-byte units
              tc_mma_bf16(d_tmem, adesc + (u64)(kk * 2), bdesc + (u64)(kk * 2), kIdesc, (kb | kk) ? 1u : 0u);
            }
            tc_commit(&sh->empty[stage]);  // frees the smem slot once these MMAs retire
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          tc_commit(&sh->tmem_full[acc]);  // accumulator complete
          if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
      }
    }
  } else {
    // ===== epilogue / selection: 8 warps; thread == (query row, 128-column half of the tile) =====
    const int ew = warp - 2;
    const int quad = warp & 3;   // TMEM lane quadrant this warp may access
    const int half = ew >> 2;    // 0: tile columns [0,128), 1: [128,256)
    const int row_in_tile = quad * 32 + lane;
    const u32 hist_s = smem_addr(sh->hist[ew]);
    const int k = p.k, C = p.C;
    const u32 stage_s = (C <= kStageKeys) ? smem_addr(sh->stage[ew]) : 0u;
    int acc = 0; u32 acc_phase = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      const int qt = unit % p.n_q_tiles, u = unit / p.n_q_tiles;
      const long long t0 = (long long)u * tiles_total / p.S, t1 = (long long)(u + 1) * tiles_total / p.S;
      const int row = qt * kQTile + row_in_tile;
      const bool valid_row = row < p.B;
      SelState st;
      st.buf = p.cand + (((long long)row * p.S + u) * 2 + half) * C;
      st.cnt = 0;
      st.tau_f = valid_row ? -INFINITY : INFINITY;
      st.tau_key = 0ull;
      st.mbeg = 0; st.mend = 0;
      if (p.mask_indptr && valid_row) { st.mbeg = p.mask_indptr[row]; st.mend = p.mask_indptr[row + 1]; }

      for (long long t = t0; t < t1; ++t) {
        mbar_wait(&sh->tmem_full[acc], acc_phase, p.status, 400 + acc);
        tc_fence_after();
        const u32 taddr0 = tmem_base + ((u32)(quad * 32) << 16) + (u32)(acc * kITile + half * 128);
        const long long col0 = t * kITile + half * 128;
        u32 va[32], vb[32];
        tmem_ld_32x32b_x32(taddr0, va);
        tmem_ld_wait();
        tmem_ld_32x32b_x32(taddr0 + 32, vb);
        select_chunk(va, col0, st, p, k, C, hist_s, stage_s);
        tmem_ld_wait();
        tmem_ld_32x32b_x32(taddr0 + 64, va);
        select_chunk(vb, col0 + 32, st, p, k, C, hist_s, stage_s);
        tmem_ld_wait();
        tmem_ld_32x32b_x32(taddr0 + 96, vb);
        select_chunk(va, col0 + 64, st, p, k, C, hist_s, stage_s);
        tmem_ld_wait();
        // every column of this half is in registers: hand the accumulator back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh->tmem_empty[acc]);
        select_chunk(vb, col0 + 96, st, p, k, C, hist_s, stage_s);
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
      p.counts[((long long)row * p.S + u) * 2 + half] = st.cnt;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((u32)kTmemCols)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D bf16 row-major [rows, D] with row pitch ld elements; box = 64 x box_rows, 128B swizzle
static int make_tmap(CUtensorMap* m, const void* base, long long rows, int D, long long ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)kKBlock, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}

int launch_select_tc(const SelectParams& p, cudaStream_t st, int num_sms) {
  CUtensorMap tq, ti;
  int r = make_tmap(&tq, p.q, p.B, p.D, p.ldq, kQTile);
  if (r) return r;
  r = make_tmap(&ti, p.items, p.n_items, p.D, p.ldi, kITile);
  if (r) return r;
  cudaError_t e = cudaFuncSetAttribute(select_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kTcSmemBytes);
  if (e != cudaSuccess) return (int)e;
  int n_units = p.n_q_tiles * p.S;
  int grid = n_units < num_sms ? n_units : num_sms;
  if (const char* g = getenv("CCR_DEBUG_GRID")) { int v = atoi(g); if (v > 0 && v < grid) grid = v; }
  if (grid < 1) grid = 1;
  select_tc_kernel<<<grid, kTcThreads, kTcSmemBytes, st>>>(tq, ti, p);
  return (int)cudaGetLastError();
}

}  // namespace ccr
