// Host-side launch prototypes of every kernel in the library (internal; the public ABI is include/tmae.h).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tmae.h"
#include "gemm.cuh"

namespace tmae {

// Per-call pointers of one forward, resident in device memory: kernels captured in a CUDA graph read the caller's
// buffers through it, so one instantiated graph serves every call of that batch size.
struct IoBlock {
    const float* imgs;
    const float* scores;
    tmae_outputs out;
};

// gemm_tc.cu
cudaError_t gemm_tc_configure();
int gemm_pick_stages(int block_n, int total_ctas, bool share_sm, int* smem_bytes, int* kgroup);
int gemm_epi_kind(const GemmParams& p);
cudaError_t gemm_launch(const GemmParams* d_params, int groups, int max_M, int max_N, int block_n, int act, int epi,
                        bool simt, bool share_sm, cudaStream_t stream,
                        const GemmParams* d_next = nullptr, int next_groups = 0, int conv_reuse_stage_bytes = 0, int pair = 0);   // pair: 1 = CTA pairs, 2 = persistent CTA pairs
bool gemm_use_pair(int groups, int epi, int act, int max_M, int block_n, bool pair_ok, bool conv);
bool gemm_use_pair_persistent(int groups, int epi, int act, int max_M, int max_N, int block_n, bool share_sm, bool conv);
int gemm_reuse_stages(int stage_bytes, int total_ctas, bool share_sm, int* smem_bytes);

// mask.cu
cudaError_t launch_mask_select(const float* scores, int N, int L, int K, int softmax_isa, int64_t* ids_shuffle,
                               int64_t* ids_restore, int64_t* ids_keep, cudaStream_t st, const IoBlock* io = nullptr);

// attention.cu
cudaError_t attention_configure(int T);
cudaError_t launch_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, int N, int T, int H, int C, float scale,
                             cudaStream_t st);
// tcgen05 / TMEM attention (T <= 192): Q / K / V by TMA out of the qkv matrix, S and O in tensor memory
bool attention_tc_supported(int T);   // kernel limits (Tp <= 384: shared memory and the 512 TMEM columns)
bool attention_tc_eligible(int T);    // policy: the tcgen05 kernel whenever it is supported (TMAE_NO_TC_ATTN=1 -> mma.sync kernel)
int attention_tc_tp(int T);
// mode: 1 = full form, 2 = several streams share the GPU (lite form for short rows), 3 = duo (two heads per tile, T = 65; opt-in)
bool attention_tc_describe(int T, int H, int N, int mode, int* out12);        // host-only: the plan for (T, H, N) in `mode`
void attention_tc_boxes(int T, int H, int mode, int* q_rows, int* kv_rows);   // TMA box rows of the Q map and of the K / V map
cudaError_t launch_attention_tc(const CUtensorMap* map_q, const CUtensorMap* map_kv, const __nv_bfloat16* qkv, __nv_bfloat16* out, int N,
                                int T, int H, int C, float scale, cudaStream_t st, long long* dbg = nullptr, int mode = 1);
// precise mode: fp32 softmax attention on CUDA cores over split-bf16 (hi + lo plane) q, k, v; writes both planes
cudaError_t launch_attention_f32(const __nv_bfloat16* qkv, long long qkv_lo, __nv_bfloat16* out, long long out_lo, int N,
                                 int T, int H, int C, float scale, cudaStream_t st);

// elementwise.cu
cudaError_t launch_gather_patches(const float* imgs, const int64_t* ids_keep, __nv_bfloat16* patches, float* x,
                                  const float* cls_token, const float* pos_embed, int N, int S, int grid_w, int K,
                                  int T, int C, int in_chans, int patch, long long lo_off, cudaStream_t st,
                                  const IoBlock* io = nullptr);
cudaError_t launch_layernorm(const float* x, const float* gamma, const float* beta, __nv_bfloat16* out, float* out_f32,
                             int rows, int C, int T, int drop_cls, float eps, long long lo_off, cudaStream_t st,
                             const IoBlock* io = nullptr);
cudaError_t launch_bottleneck(const float* z, const float* eb_tab, long long rows, int Cz, float* lik, int32_t* sym,
                              float* zhat, __nv_bfloat16* zhat_bf, long long lo_off, int s4, double* rate_acc, int rows_per_image,
                              int16_t* sym16,
                              cudaStream_t st, const IoBlock* io = nullptr);
cudaError_t launch_gaussian_slice(const float* y, const float* mu, const float* sigma, long long rows, int ld, int col0,
                                  int cs, float* lik, int32_t* sym, float* yhat, __nv_bfloat16* yhat_bf, long long lo_off,
                                  int ld_bf, int s, double* rate_acc, const float* scale_table, int n_table, int16_t* sym16,
                                  int32_t* idx, cudaStream_t st, const IoBlock* io = nullptr);
cudaError_t launch_pack_nchw_i32(const int32_t* src, int32_t* dst, int N, int hw, int C, cudaStream_t st);
cudaError_t launch_gaussian_flat(const float* y, const float* mu, const float* sigma, long long n, float* lik,
                                 int32_t* sym, float* yhat, cudaStream_t st);
cudaError_t launch_rate_finalize(const double* rate_acc, int N, double pixels_per_image, float* bpp,
                                 double* rate_sums, cudaStream_t st, const IoBlock* io = nullptr);
// copies the workspace-resident results the caller asked for (io->out.{y,z,mu,sigma,y_hat,ids_keep})
cudaError_t launch_copy_outputs(const IoBlock* io, const float* y, const float* z, const float* mu, const float* sigma,
                                const float* yhat, const int64_t* ids_keep, long long n_y, long long n_z, long long n_ids,
                                cudaStream_t st);
cudaError_t launch_prepack_weight(const float* w, __nv_bfloat16* out, int Cout, int Cin_total, int taps, int nseg,
                                  const int* segc, int shuffle, int planes, cudaStream_t st, const float* gamma = nullptr);
cudaError_t launch_fold_ln(const float* w, const float* beta, const float* bias, const __nv_bfloat16* packed, int Kp, int Cout, int Cin,
                           float* bias_out, float* wsum, cudaStream_t st);
cudaError_t launch_build_blockdiag2(const float* wa, const float* wb, const float* ba, const float* bb, float* w_out, float* b_out,
                                    int half, int Cin, int taps, cudaStream_t st);
cudaError_t launch_permute_bias_shuffle(const float* b, float* out, int Cout, cudaStream_t st);
cudaError_t launch_eb_table(const float* const* ptrs, float* tab, int Cz, cudaStream_t st);
cudaError_t launch_f32_to_bf16(const float* a, __nv_bfloat16* o, long long n, long long lo_off, cudaStream_t st);
cudaError_t launch_f32_to_bf16_cols(const float* a, __nv_bfloat16* o, long long rows, int cols, int ld, long long lo_off, cudaStream_t st);

// ---- patch-score generation (scores.cu; generate_scores_file.py:19-31, utils/map.py, utils/distribution.py) ----
constexpr int kScoreMaxLevels = 16;
struct ScoreGeom {
    int H, W, S;                           // grey image, side of the resized maps (224)
    int levels;                            // quadtree levels whose nodes may split (min(h, w) > 5); deeper nodes are leaves
    int flag_total;                        // split-decision flags per image = sum of 4^d over those levels
    int large_levels, large_nodes, large_cta_total, small_nodes;
    int h[kScoreMaxLevels + 1], w[kScoreMaxLevels + 1];       // node size per level: int(h / 2) of the level above
    int flag_off[kScoreMaxLevels + 1];
    int chunks[kScoreMaxLevels], large_node_off[kScoreMaxLevels], large_ctas[kScoreMaxLevels];
    double scale_x_crop, scale_y_crop, scale_x_full, scale_y_full;   // cv2.resize source/destination ratios
};
bool score_geometry(int H, int W, int S, ScoreGeom* out);
size_t score_workspace_bytes(const ScoreGeom& g, int n);
cudaError_t launch_generate_scores(const uint8_t* gray, int n, const ScoreGeom& g, float* scores, uint8_t* s_map, uint8_t* t_map,
                                   uint8_t* seg_out, void* workspace, cudaStream_t st);

}  // namespace tmae
