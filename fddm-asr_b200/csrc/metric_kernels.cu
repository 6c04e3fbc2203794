// metric_kernels.cu -- batched Levenshtein distance for CER / WER.
//
// Replaces the reference's per-pair numpy/Python O(n*m) loops (models/evaluate.py:94-134: calculate_cer on
// characters, calculate_wer on whitespace-split words; called once per utterance at evaluate.py:185,331,448).
// Symbols arrive as int32 (code points / word ids), all pairs of an evaluation set in ONE launch: a thread owns
// a pair and rolls a single DP row (evaluate.py:104-115: deletion, insertion, substitution, unit costs); the
// rows of all pairs are interleaved in the workspace ([column][pair]) so that the threads of a warp touch
// consecutive words.  Integer arithmetic: results are exact.
#include <algorithm>

#include "common.cuh"

namespace fddm {
namespace {

__global__ void __launch_bounds__(128) edit_distance_kernel(const int32_t* __restrict__ ref,
                                                            const int64_t* __restrict__ ref_off,
                                                            const int32_t* __restrict__ hyp,
                                                            const int64_t* __restrict__ hyp_off, int64_t n_pairs,
                                                            int32_t* __restrict__ row, int32_t* __restrict__ dist) {
  const int64_t pr = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pr >= n_pairs) return;
  const int32_t* r = ref + ref_off[pr];
  const int32_t* h = hyp + hyp_off[pr];
  const int n = static_cast<int>(ref_off[pr + 1] - ref_off[pr]);
  const int m = static_cast<int>(hyp_off[pr + 1] - hyp_off[pr]);
  int32_t* dp = row + pr;                                  // dp[j] lives at row[j * n_pairs + pr]
  for (int j = 0; j <= m; ++j) dp[static_cast<int64_t>(j) * n_pairs] = j;           // evaluate.py:106-107
  for (int i = 1; i <= n; ++i) {
    const int32_t ri = r[i - 1];
    int32_t diag = dp[0];                                  // dp[i-1][0]
    int32_t left = i;                                      // dp[i][0] = i          evaluate.py:104-105
    dp[0] = i;
    for (int j = 1; j <= m; ++j) {
      const int32_t up = dp[static_cast<int64_t>(j) * n_pairs];                      // dp[i-1][j]
      const int32_t cost = (ri == h[j - 1]) ? 0 : 1;
      const int32_t v = min(min(up + 1, left + 1), diag + cost);                    // evaluate.py:111-115
      dp[static_cast<int64_t>(j) * n_pairs] = v;
      diag = up;
      left = v;
    }
  }
  dist[pr] = dp[static_cast<int64_t>(m) * n_pairs];
}

}  // namespace
}  // namespace fddm

extern "C" {

size_t fddm_edit_distance_workspace_bytes(int64_t n_pairs, int64_t max_hyp_len) {
  if (n_pairs <= 0 || max_hyp_len < 0) return 0;
  return static_cast<size_t>(n_pairs) * static_cast<size_t>(max_hyp_len + 1) * sizeof(int32_t);
}

int fddm_edit_distance(const int32_t* ref, const int64_t* ref_off, const int32_t* hyp, const int64_t* hyp_off,
                       int64_t n_pairs, int64_t max_hyp_len, void* workspace, int32_t* dist_out, fddm_stream_t stream_) {
  FDDM_API_RANGE();
  using namespace fddm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FDDM_CHECK_ARG(ref_off && hyp_off && workspace && dist_out, "edit_distance: null pointer argument");
  FDDM_CHECK_ARG(n_pairs > 0 && n_pairs < (1ll << 31) && max_hyp_len >= 0 && max_hyp_len < (1ll << 24),
                 "edit_distance: bad size");
  KernelScope ks("edit_distance_kernel", stream);
  edit_distance_kernel<<<static_cast<unsigned>((n_pairs + 127) / 128), 128, 0, stream>>>(
      ref, ref_off, hyp, hyp_off, n_pairs, static_cast<int32_t*>(workspace), dist_out);
  FDDM_LAUNCH_OK();
  return FDDM_OK;
}

}  // extern "C"
