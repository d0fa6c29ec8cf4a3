// Internal declarations shared by the translation units of libpst_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "pst_abi.h"

#define PST_D 128
#define PST_FFN 512
#define PST_TRANS 256
#define PST_HEADS 4
#define PST_HEAD_DIM 32
#define PST_C8 PST_MAX_LEVELS
#define PST_FEAT_PAD 32   // 27 edge features padded to 32 GEMM rows
#define PST_PREP_STRIDE 16 // doubles per residue in the prep record
#define PST_PROF_MAX_SPANS 512
#define PST_PROF_KINDS 8   // 0 featurise+knn, 1 message MLP, 2 edge-update MLP, 3 node update, 4 input embeddings, 5 resampler + head, 6 FSQ quantiser

// ---- prepared weight blob (fp32), see pst/weights.py for the packer ----------
struct PstLayerW {
  const float *msg_w1, *msg_b1, *msg_w2, *msg_b2, *msg_w3, *msg_b3;
  const float *ln0_s, *ln0_o;
  const float *ffn_w1, *ffn_b1, *ffn_w2, *ffn_b2;
  const float *ln1_s, *ln1_o;
  const float *edge_w1, *edge_b1, *edge_w2, *edge_b2, *edge_w3, *edge_b3;
  const float *ln2_s, *ln2_o;
};
struct PstBlockW {
  const float *qn_s, *qn_o, *dn_s, *dn_o;
  const float *wq, *wk, *wv, *wg, *bg, *wo, *bo;
  const float *rt_ln_s, *rt_ln_o, *rt_w1, *rt_b1, *rt_w2, *rt_b2;
  const float *ot_ln_s, *ot_ln_o, *ot_w1, *ot_b1, *ot_w2, *ot_b2;
};
#define PST_MAX_LAYERS 4
#define PST_MAX_BLOCKS 4
struct PstWeights {
  const float* node_table;    // [seq_max_size,128]   PE_node.W + b
  const float* edge_pe_table; // [2*seq_max_size-1,128] PE_edge.W[0:128] + b
  const float* edge_feat_w;   // [32,128]  W[128:155], rows 27..31 zero
  PstLayerW layer[PST_MAX_LAYERS];
  const float* token_table;   // [max_out_len,128]
  PstBlockW block[PST_MAX_BLOCKS];
  const float* down_w;        // [128,8]
  const float* down_b;        // [8]
};

// operand-precision copies of the edge-level MLP weights for the tensor-core path:
// per MLP three 128x128 matrices, 16-bit, stored [n][k] (K-major "B" operand),
// W1 restricted to its edge rows (256..383).
struct PstTcWeights {
  const uint16_t* w;  // [layers][2 (msg,edge)][3][128*128]
};

struct PstLinearRegistry;
struct PstNodeChain;

struct pst_model {
  pst_config cfg;
  int device;
  int num_sms;
  float* blob_dev;
  float* blob_host;          // host copy of the blob: small parameter vectors are passed to kernels by value
  size_t blob_floats;
  PstWeights w;
  uint16_t* tc_dev;
  PstTcWeights tc;
  uint16_t* embed_img_dev;   // hi/lo fp16 images of W_edge[128:155] (tensor-core input embedding)
  uint16_t* table16_dev;     // fp16 copy of edge_pe_table
  uint16_t* layer0_tables;   // [2][seq_max][128] fp16: layer-1 message-MLP addend tables by position (tensor-core modes)
  PstLinearRegistry* linear_tc;  // split-fp16 operand images of the node-level weights (tensor-core modes)
  PstNodeChain* node_chain;      // weight streaming schedules of the fused node-level kernels (node_chain_tc.cu)
  // FSQ constants (model/quantize.py:175-181), fp32
  float half_l[PST_C8], fsq_offset[PST_C8], fsq_shift[PST_C8];
  int32_t basis[PST_C8], half_width[PST_C8];
  bool use_fused_resampler;  // df > 1: the two fused chain kernels instead of the per-op path (PST_FUSED_RESAMPLER=0 switches them off)
  bool use_fused_fsq;        // pst_tokenize: the quantiser is the epilogue of the fused resampler kernels (PST_FUSED_FSQ=0: separate launch)
  bool use_msg_t;            // message MLPs through the transposed kernel (edge_msg_t_kernel; PST_MSG_T=0 switches it off)
  mutable int launch_count;
  // CUDA-graph cache of the fused hot call (api.cu): a pst_tokenize call whose arguments (pointers and sizes) repeat
  // is captured once and replayed, which removes the launch gaps between its dependent kernels
  mutable struct PstGraphCache* graphs;
  // optional per-kernel-group timing (pst_profile_*): CUDA events recorded on the call's stream
  mutable bool prof_on;
  mutable int prof_n;
  mutable cudaEvent_t prof_ev[2 * PST_PROF_MAX_SPANS];
  mutable int prof_kind[PST_PROF_MAX_SPANS];
};

// RAII span: records an event pair around a group of launches when profiling is enabled
struct PstSpan {
  const pst_model* m; cudaStream_t st; int idx;
  PstSpan(const pst_model* m_, cudaStream_t st_, int kind) : m(m_), st(st_), idx(-1) {
    if (m->prof_on && m->prof_n < PST_PROF_MAX_SPANS) {
      idx = m->prof_n++;
      m->prof_kind[idx] = kind;
      cudaEventRecord(m->prof_ev[2 * idx], st);
    }
  }
  ~PstSpan() { if (idx >= 0) cudaEventRecord(m->prof_ev[2 * idx + 1], st); }
};

size_t pst_fill_weight_pointers(const pst_config& cfg, const float* base, PstWeights* w);

// ---- workspace carving -------------------------------------------------------
struct PstWorkspace {
  int32_t* status;       // [4]
  int32_t* row_base;     // [R]
  int32_t* redo;         // [1 + R] count, then the rows the packed-key k-NN kernel hands to the exact kernel
  double* prep;          // [R,16]
  double* cen4;          // centroids as x[R], y[R], z[R] (room for 4R doubles): coalesced copy for the k-NN scan
  int32_t* senders;      // [E]
  float* edge_feat;      // [E,27]
  float* e;              // [E,128] fp32 (fp32 mode) or 16-bit (tensor-core modes: same pointer, half the bytes)
  float* t1;             // [E,128]   (fp32 mode only)
  float* t2;             // [E,128]   (fp32 mode only)
  float* partial;        // [tiles,4,128] (tensor-core modes only)
  int32_t* senders_abs;  // [E] absolute sender rows (tensor-core modes only)
  float* h;              // [R,128]
  float* agg;            // [R,128]
  float* ps;             // [R,128]
  float* pr;             // [R,128]
  float* tmp;            // [R,128]
  float* u;              // [max(R,T),512]
  float* orig;           // [R,128]
  float* dn;             // [R,128]
  float* kx;             // [R,128]
  float* vx;             // [R,128]
  float* res;            // [T,128]
  float* qn;             // [T,128]
  float* q;              // [T,128]
  float* g;              // [T,128]
  float* wa;             // [T,128]
  float* z;              // [T,8]
  size_t bytes;
};
PstWorkspace pst_carve_workspace(const pst_model* m, void* base, int R, int T_upper);

// ---- kernel launchers (each returns the number of kernels enqueued) --------------
int pst_launch_featurize(const pst_model* m, cudaStream_t st, const float* atoms,
                         const uint8_t* mask, int apr, const int32_t* offsets, int B, int R,
                         int32_t* senders, float* edge_feat, double* prep, double* cen4, int32_t* status,
                         int32_t* redo, int compact = 0);
bool pst_featurize_compact_ok(const pst_model* m);

int pst_launch_encode_fp32(const pst_model* m, cudaStream_t st, const float* edge_feat,
                           const int32_t* senders, const int32_t* offsets,
                           const int32_t* token_offsets, int B, int R, int T, float* z_out,
                           PstWorkspace& ws, int compact_features = 0, int32_t* fused_tokens = nullptr,
                           bool* fused_tokens_done = nullptr);  // fused_tokens: let the resampler's head emit the token ids too

int pst_launch_quantize(const pst_model* m, cudaStream_t st, const float* z, int n,
                        int32_t* tokens, float* bounded, int32_t* status = nullptr);
int pst_launch_fsq_pack(const pst_model* m, cudaStream_t st, const float* bounded, int n,
                        int32_t* tokens);
int pst_launch_indexes_to_codes(const pst_model* m, cudaStream_t st, const int32_t* tokens, int n,
                                float* codes);

// tensor-core edge MLP (edge_mlp_tc.cu).  mode 0: message MLP -> agg[R,128] = mean_K;
// mode 1: edge update -> e = LN(e + MLP).  `senders` are ABSOLUTE rows (pst_launch_abs_senders).
// Returns kernels launched, <0 on error.
int pst_launch_edge_mlp_tc(const pst_model* m, cudaStream_t st, int layer, int mode, uint16_t* e,
                           const uint16_t* ps, const uint16_t* pr, const int32_t* senders,
                           const int32_t* row_base, float* partial, int R);
size_t pst_tc_partial_floats(int R, int K);
// transposed message MLP (round 2): the sender term is a product with the TMA-gathered rows of h16 = fp16(h)
bool pst_edge_msg_t_ok(const pst_model* m);
int pst_launch_to_half(cudaStream_t st, const float* src, uint16_t* dst, size_t n);
int pst_launch_edge_msg_t(const pst_model* m, cudaStream_t st, int layer, const uint16_t* e, const uint16_t* h16, const uint16_t* pr,
                          const int32_t* senders_abs, float* partial, int R);
int pst_launch_abs_senders(const pst_model* m, cudaStream_t st, const int32_t* senders, const int32_t* row_base, int R,
                           int32_t* senders_abs);
int pst_launch_edge_embed_tc(const pst_model* m, cudaStream_t st, const float* feat, const int32_t* senders,
                             const int32_t* row_base, int R, uint16_t* e, int compact = 0);

// node-level linears on tensor cores (linear_tc.cu)
int pst_prepare_linear_tc(pst_model* m);
void pst_destroy_linear_tc(pst_model* m);
int pst_launch_linear_tc(const pst_model* m, cudaStream_t st, const float* A, const float* W, float* C, int M, int N,
                         int K, const float* bias, const float* residual, float scale, int act, int out_half);

// fused node-level chains (node_chain_tc.cu); prepare after pst_prepare_linear_tc
const uint8_t* pst_linear_tc_image(const pst_model* m, const float* W, int K, int N);
int pst_prepare_node_chain(pst_model* m);
void pst_destroy_node_chain(pst_model* m);
int pst_prepare_layer0_tables(pst_model* m);  // encoder_fp32.cu
int pst_launch_node_update(const pst_model* m, cudaStream_t st, int layer, const float* partial, int partial_tile_shift, float* h, int R,
                           uint16_t* out_edge_s, uint16_t* out_edge_r, uint16_t* out_msg_s, uint16_t* out_msg_r,
                           uint16_t* h16 = nullptr);
// tokens / status (both or neither): the FSQ epilogue writes the token ids as well (fused tokenize call)
int pst_launch_resampler_df1(const pst_model* m, cudaStream_t st, const float* h, const int32_t* row_base, int R, float* z,
                             int32_t* tokens = nullptr, int32_t* status = nullptr);
// downsampling_ratio > 1: two chain kernels (residue track, token track); kv = six [R,128] fp32 buffers, info = int2 [T]
int pst_launch_resampler_dfn(const pst_model* m, cudaStream_t st, const float* h, const int32_t* offsets, const int32_t* token_offsets,
                             int B, int R, int T, float* const* kv, void* info, float* z, int32_t* tokens = nullptr,
                             int32_t* status = nullptr);

#define PST_CUDA_OK(expr)                                  \
  do {                                                     \
    cudaError_t _e = (expr);                               \
    if (_e != cudaSuccess) return PST_ERR_CUDA;            \
  } while (0)
