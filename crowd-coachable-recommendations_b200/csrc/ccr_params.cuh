// Kernel parameter blocks and launcher prototypes shared by the .cu translation units.
#pragma once
#include "ccr_common.cuh"

namespace ccr {

// One "unit" = (query tile, item split).  Each unit owns, per query row, a private candidate
// buffer of `C` 64-bit keys in the workspace:  cand[(row * S + split) * C + i],
// counts[row * S + split].  rows are padded to a multiple of the row-tile of the kernel.
struct SelectParams {
  const __nv_bfloat16* q;
  long long ldq;
  int B;
  const __nv_bfloat16* items;
  long long ldi;
  long long n_items;
  int D;
  int k;
  int C;          // candidate capacity per (row, split)
  int S;          // item splits
  int n_q_tiles;  // query tiles (tensor-core kernel) or query groups (SIMT kernel)
  const long long* mask_indptr;  // null when no mask
  const int* mask_cols;
  u64* cand;
  int* counts;
  DeviceStatus* status;
  int debug;  // CCR_DEBUG bits (env, diagnostics only): 1 = skip selection, keep pipeline
};

struct FinalizeParams {
  int B;
  int k;
  int C;
  int S;
  const u64* cand;
  const int* counts;
  // mask overrides (null when no mask): per mask entry e, value and ~col
  const long long* mask_indptr;
  const u64* ovr_hi;
  const u32* ovr_lo;
  long long id_offset;
  float* out_scores;
  double* out_scores64;
  long long* out_ids;
};

struct OverrideParams {
  const __nv_bfloat16* q;
  long long ldq;
  int B;
  const __nv_bfloat16* items;
  long long ldi;
  long long n_items;
  int D;
  const long long* mask_indptr;
  const int* mask_cols;
  const double* mask_vals;
  long long nnz;
  int mode;
  u64* ovr_hi;
  u32* ovr_lo;
};

// launchers (return cudaError_t as int)
int launch_select_simt(const SelectParams& p, cudaStream_t st);
int launch_select_tc(const SelectParams& p, cudaStream_t st, int num_sms);
int launch_overrides(const OverrideParams& p, cudaStream_t st);
int launch_finalize(const FinalizeParams& p, cudaStream_t st);
int launch_merge_topk(const double* s, const long long* ids, int G, long long B, int k_in, int k_out,
                      float* os, double* os64, long long* oi, cudaStream_t st);
int launch_ingest_f32(const float* src, long long n, int D, long long ld_src, __nv_bfloat16* dst,
                      long long ld_dst, int normalize, cudaStream_t st);
int launch_normalize_bf16(const __nv_bfloat16* src, long long n, int D, long long ld_src,
                          __nv_bfloat16* dst, long long ld_dst, cudaStream_t st);
int launch_dense_f32(const __nv_bfloat16* q, long long B, long long ldq, const __nv_bfloat16* items,
                     long long n, long long ldi, int D, float* out, long long ld_out, cudaStream_t st);

}  // namespace ccr
