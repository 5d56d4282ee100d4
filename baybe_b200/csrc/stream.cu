// stream.cu -- the end-to-end pass over a candidate set that lives in (pinned) HOST memory.
//
// bb_score_fused_host copies row blocks on the caller's copy stream while the previous block is expanded
// (level-coded rows: bb_decode_codes) and scored (the fused kernel) on the caller's compute stream; two staging
// slots, four events created and destroyed inside the call.  One C call per pass: the Python loop it replaces spent
// ~70 us of host time per block, more than the GPU needs for a block of 125,000 rows (60 us).
// Replaces: SubspaceDiscrete.transform + to_tensor + optimize_acqf_discrete over a host-resident comp-rep
// (baybe/recommenders/pure/bayesian/botorch/discrete.py:120-126).
#include "fused_common.cuh"

namespace bb {
int launch_fused(const bb_model* m, const void* d_x, int32_t layout, int64_t N, int64_t ldx, const bb_acq_spec* acq,
                 const float* d_z, int32_t S, const uint8_t* d_keep, float* d_mu, float* d_var, float* d_score,
                 int64_t* d_best_key, int64_t index_offset, cudaStream_t stream, const WideCross* wc);
}
extern "C" int bb_decode_codes(const uint8_t* d_codes, int32_t bits, int64_t N, int32_t d, int64_t ld_bytes,
                               const float* d_table, int32_t table_ld, float* d_out, int64_t ldo, void* stream);
extern "C" int bb_best_init(int64_t* d_best_key, void* stream);

using namespace bb;

extern "C" int bb_score_fused_host(const bb_model* m, const bb_acq_spec* a, const void* h_x, int32_t host_format,
                                   int64_t N, int64_t ld, const float* d_table, int32_t table_ld,
                                   void* const* d_stage, float* const* d_rows, int64_t block_rows,
                                   const uint8_t* d_keep, const float* d_z, int32_t S, float* d_score,
                                   int64_t* d_best_key, int64_t index_offset, void* stream_, void* copy_stream_) {
  cudaStream_t stream = (cudaStream_t)stream_, copy = (cudaStream_t)copy_stream_;
  BB_CHECK_ARG(m && m->abi_version == BB_ABI_VERSION && a, "bb_score_fused_host: model / acquisition spec missing");
  BB_CHECK_ARG(N >= 0 && block_rows >= 128 && block_rows % 128 == 0, "bb_score_fused_host: block_rows must be a positive multiple of 128");
  BB_CHECK_ARG(d_stage && d_stage[0] && d_stage[1] && d_best_key, "bb_score_fused_host: staging buffers / key missing");
  BB_CHECK_ARG(stream != copy, "bb_score_fused_host: the copy stream must differ from the compute stream");
  const int d = m->d;
  size_t row_bytes;
  int32_t dev_layout = BB_ROW_MAJOR_F32;
  int bits = 0;
  switch (host_format) {
    case BB_HOST_ROWS_F32: row_bytes = (size_t)ld * 4; dev_layout = BB_ROW_MAJOR_F32; break;
    case BB_HOST_ROWS_F64: row_bytes = (size_t)ld * 8; dev_layout = BB_ROW_MAJOR_F64; break;
    case BB_HOST_CODES4: row_bytes = (size_t)ld; bits = 4; break;
    case BB_HOST_CODES8: row_bytes = (size_t)ld; bits = 8; break;
    default: set_error("bb_score_fused_host: unknown host format %d", host_format); return BB_ERR_INVALID;
  }
  if (bits) {
    BB_CHECK_ARG(d_table && d_rows && d_rows[0] && d_rows[1], "bb_score_fused_host: value table / row buffers missing");
    BB_CHECK_ARG(ld >= (bits == 8 ? d : (d + 1) / 2), "bb_score_fused_host: code rows shorter than d columns");
  } else {
    BB_CHECK_ARG(ld >= d, "bb_score_fused_host: leading dimension smaller than d");
  }
  int rc = bb_best_init(d_best_key, stream);
  if (rc != BB_OK || N == 0) return rc;
  BB_CHECK_ARG(h_x != nullptr, "bb_score_fused_host: host matrix is null");
  cudaEvent_t ev[5];
  for (int i = 0; i < 5; ++i) BB_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
  cudaEvent_t* ready = ev;       // [2] block landed in its staging slot
  cudaEvent_t* freed = ev + 2;   // [2] staging slot consumed
  BB_CUDA(cudaEventRecord(ev[4], stream));  // staging buffers may still be in use by earlier work on `stream`
  BB_CUDA(cudaStreamWaitEvent(copy, ev[4], 0));
  rc = BB_OK;
  int b = 0;
  for (int64_t lo = 0; lo < N && rc == BB_OK; lo += block_rows, ++b) {
    const int64_t rows = (N - lo) < block_rows ? (N - lo) : block_rows;
    const int slot = b & 1;
    if (b >= 2) BB_CUDA(cudaStreamWaitEvent(copy, freed[slot], 0));
    BB_CUDA(cudaMemcpyAsync(d_stage[slot], reinterpret_cast<const uint8_t*>(h_x) + (size_t)lo * row_bytes,
                            (size_t)rows * row_bytes, cudaMemcpyHostToDevice, copy));
    BB_CUDA(cudaEventRecord(ready[slot], copy));
    BB_CUDA(cudaStreamWaitEvent(stream, ready[slot], 0));
    const void* x = d_stage[slot];
    int64_t ldx = ld;
    if (bits) {
      rc = bb_decode_codes(reinterpret_cast<const uint8_t*>(d_stage[slot]), bits, rows, d, ld, d_table, table_ld,
                           d_rows[slot], d, stream);
      if (rc != BB_OK) break;
      x = d_rows[slot];
      ldx = d;
    }
    rc = launch_fused(m, x, dev_layout, rows, ldx, a, d_z, S, d_keep ? d_keep + lo : nullptr, nullptr, nullptr,
                      d_score ? d_score + lo : nullptr, d_best_key, index_offset + lo, stream, nullptr);
    if (rc != BB_OK) break;
    BB_CUDA(cudaEventRecord(freed[slot], stream));
  }
  for (int i = 0; i < 5; ++i) cudaEventDestroy(ev[i]);
  return rc;
}
