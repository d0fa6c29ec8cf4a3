// C-ABI entry points (include/pst_abi.h): model handle, workspace carving, call sequencing.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "pst_internal.h"

#include <mutex>
#include <vector>

// The cache is keyed by the BUFFERS of a call (a staging slot of a chunk pipeline, a resident batch), not by its sizes:
// when the same buffers come back with another (B, R, T) -- the next ragged chunk of a stream -- the launch sequence is
// captured again and the instantiated graph is UPDATED in place (cudaGraphExecUpdate: same kernels in the same order,
// new parameters and grids), so a stream of distinct chunk shapes also runs as one graph launch per call.
struct PstGraphKey {
  const void *atoms, *mask, *offsets, *token_offsets, *tokens, *workspace;
  size_t ws_bytes;
  int apr;
  bool operator==(const PstGraphKey& o) const {
    return atoms == o.atoms && mask == o.mask && offsets == o.offsets && token_offsets == o.token_offsets &&
           tokens == o.tokens && workspace == o.workspace && ws_bytes == o.ws_bytes && apr == o.apr;
  }
};
struct PstGraphEntry {
  PstGraphKey key;
  cudaGraphExec_t exec = nullptr;
  int B = -1, R = -1, T = -1;  // sizes the instantiated graph currently holds
  int launches = 0;
  bool failed = false;
  unsigned long long last_use = 0;
};
struct PstGraphCache {
  static constexpr size_t kMaxEntries = 64;  // buffer sets (a resident pass over many chunks has one per chunk)
  std::mutex mu;
  std::vector<PstGraphEntry> entries;
  unsigned long long clock = 0;
  bool enabled = true;
  int n_replay = 0, n_update = 0, n_instantiate = 0, n_eager = 0;  // pst_graph_cache_stats
  PstGraphEntry* find(const PstGraphKey& k) {
    for (auto& e : entries)
      if (e.key == k) return &e;
    return nullptr;
  }
  void insert(const PstGraphKey& k) {
    if (entries.size() >= kMaxEntries) {  // evict the least recently used entry
      size_t lru = 0;
      for (size_t i = 1; i < entries.size(); ++i)
        if (entries[i].last_use < entries[lru].last_use) lru = i;
      if (entries[lru].exec) cudaGraphExecDestroy(entries[lru].exec);
      entries.erase(entries.begin() + lru);
    }
    PstGraphEntry e;
    e.key = k;
    e.last_use = ++clock;
    entries.push_back(e);
  }
  void clear() {
    for (auto& e : entries)
      if (e.exec) cudaGraphExecDestroy(e.exec);
    entries.clear();
  }
};

int pst_prepare_tc_weights(pst_model* m);  // edge_mlp_tc.cu

namespace {

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

bool config_ok(const pst_config* c) {
  if (!c || c->abi_version != PST_ABI_VERSION) return false;
  if (c->seq_max_size < 1 || c->seq_max_size > 8192) return false;
  if (c->max_len < 1 || c->max_len > c->seq_max_size) return false;
  if (c->num_neighbor < 1 || c->num_neighbor > 64) return false;
  if (c->downsampling_ratio < 1 || c->downsampling_ratio > 8) return false;
  if (c->max_out_len < 1 || c->max_out_len < c->max_len / c->downsampling_ratio) return false;
  if (c->num_levels < 1 || c->num_levels > PST_MAX_LEVELS) return false;
  if (c->gnn_layers < 1 || c->gnn_layers > PST_MAX_LAYERS) return false;
  if (c->num_blocks < 1 || c->num_blocks > PST_MAX_BLOCKS) return false;
  if (c->precision < PST_PREC_FP32 || c->precision > PST_PREC_BF16) return false;
  long long codes = 1;
  for (int i = 0; i < c->num_levels; ++i) {
    if (c->levels[i] < 2 || c->levels[i] > 64) return false;
    codes *= c->levels[i];
    if (codes > 0x7fffffffLL) return false;
  }
  return true;
}

}  // namespace

// Walks the blob in the order documented in pst/weights.py; returns the float count.
size_t pst_fill_weight_pointers(const pst_config& cfg, const float* base, PstWeights* w) {
  size_t off = 0;
  auto take = [&](size_t n) {
    const float* p = base ? base + off : nullptr;
    off += n;
    return p;
  };
  const size_t D = PST_D;
  PstWeights tmp;
  PstWeights& W = w ? *w : tmp;
  W.node_table = take((size_t)cfg.seq_max_size * D);
  W.edge_pe_table = take((size_t)(2 * cfg.seq_max_size - 1) * D);
  W.edge_feat_w = take(PST_FEAT_PAD * D);
  for (int l = 0; l < cfg.gnn_layers; ++l) {
    PstLayerW& L = W.layer[l];
    L.msg_w1 = take(3 * D * D); L.msg_b1 = take(D);
    L.msg_w2 = take(D * D);     L.msg_b2 = take(D);
    L.msg_w3 = take(D * D);     L.msg_b3 = take(D);
    L.ln0_s = take(D);          L.ln0_o = take(D);
    L.ffn_w1 = take(D * PST_FFN); L.ffn_b1 = take(PST_FFN);
    L.ffn_w2 = take(PST_FFN * D); L.ffn_b2 = take(D);
    L.ln1_s = take(D);          L.ln1_o = take(D);
    L.edge_w1 = take(3 * D * D); L.edge_b1 = take(D);
    L.edge_w2 = take(D * D);    L.edge_b2 = take(D);
    L.edge_w3 = take(D * D);    L.edge_b3 = take(D);
    L.ln2_s = take(D);          L.ln2_o = take(D);
  }
  W.token_table = take((size_t)cfg.max_out_len * D);
  for (int b = 0; b < cfg.num_blocks; ++b) {
    PstBlockW& B = W.block[b];
    B.qn_s = take(D); B.qn_o = take(D); B.dn_s = take(D); B.dn_o = take(D);
    B.wq = take(D * D); B.wk = take(D * D); B.wv = take(D * D); B.wg = take(D * D);
    B.bg = take(D);     B.wo = take(D * D); B.bo = take(D);
    B.rt_ln_s = take(D); B.rt_ln_o = take(D);
    B.rt_w1 = take(D * PST_TRANS); B.rt_b1 = take(PST_TRANS);
    B.rt_w2 = take(PST_TRANS * D); B.rt_b2 = take(D);
    B.ot_ln_s = take(D); B.ot_ln_o = take(D);
    B.ot_w1 = take(D * PST_TRANS); B.ot_b1 = take(PST_TRANS);
    B.ot_w2 = take(PST_TRANS * D); B.ot_b2 = take(D);
  }
  W.down_w = take(D * PST_C8);
  W.down_b = take(PST_C8);
  return off;
}

PstWorkspace pst_carve_workspace(const pst_model* m, void* base, int R, int T) {
  PstWorkspace ws{};
  const size_t K = m->cfg.num_neighbor;
  const size_t E = (size_t)R * K;
  const size_t D = PST_D;
  const bool fp32 = m->cfg.precision == PST_PREC_FP32;
  size_t off = 0;
  char* b = static_cast<char*>(base);
  auto take = [&](size_t bytes) {
    void* p = b ? b + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  const size_t RT = (size_t)(R > T ? R : T);
  ws.status = (int32_t*)take(4 * sizeof(int32_t));
  ws.row_base = (int32_t*)take((size_t)R * sizeof(int32_t));
  ws.redo = (int32_t*)take((size_t)(R + 1) * sizeof(int32_t));  // [0] = count, [1..] = rows
  ws.prep = (double*)take((size_t)R * PST_PREP_STRIDE * sizeof(double));
  ws.cen4 = (double*)take((size_t)R * 4 * sizeof(double));
  ws.senders = (int32_t*)take(E * sizeof(int32_t));
  ws.edge_feat = (float*)take(E * PST_EDGE_FEATURES * sizeof(float));
  ws.e = (float*)take(E * D * (fp32 ? sizeof(float) : sizeof(uint16_t)));
  ws.t1 = (float*)take(fp32 ? E * D * sizeof(float) : 0);
  ws.t2 = (float*)take(fp32 ? E * D * sizeof(float) : 0);
  ws.partial = (float*)take(fp32 ? 0 : pst_tc_partial_floats(R, (int)K) * sizeof(float));
  ws.senders_abs = (int32_t*)take(fp32 ? 0 : E * sizeof(int32_t));
  ws.h = (float*)take((size_t)R * D * sizeof(float));
  ws.agg = (float*)take((size_t)R * D * sizeof(float));
  ws.ps = (float*)take((size_t)R * D * sizeof(float));
  ws.pr = (float*)take((size_t)R * D * sizeof(float));
  ws.tmp = (float*)take((size_t)R * D * sizeof(float));
  ws.u = (float*)take(RT * PST_FFN * sizeof(float));
  ws.orig = (float*)take((size_t)R * D * sizeof(float));
  ws.dn = (float*)take((size_t)R * D * sizeof(float));
  ws.kx = (float*)take((size_t)R * D * sizeof(float));
  ws.vx = (float*)take((size_t)R * D * sizeof(float));
  ws.res = (float*)take((size_t)T * D * sizeof(float));
  ws.qn = (float*)take((size_t)T * D * sizeof(float));
  ws.q = (float*)take((size_t)T * D * sizeof(float));
  ws.g = (float*)take((size_t)T * D * sizeof(float));
  ws.wa = (float*)take((size_t)T * D * sizeof(float));
  ws.z = (float*)take((size_t)T * PST_C8 * sizeof(float));
  ws.bytes = off;
  return ws;
}

extern "C" {

int pst_abi_version(void) { return PST_ABI_VERSION; }

const char* pst_status_string(int status) {
  switch (status) {
    case PST_OK: return "ok";
    case PST_ERR_BAD_ARGUMENT: return "bad argument";
    case PST_ERR_UNSUPPORTED_CONFIG: return "unsupported config";
    case PST_ERR_LENGTH_OUT_OF_RANGE: return "structure length outside [num_neighbor, max_len]";
    case PST_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
    case PST_ERR_CUDA: return "CUDA error";
    case PST_ERR_NO_DEVICE: return "no CUDA device";
    case PST_ERR_BAD_WEIGHTS: return "weight blob has the wrong size";
    case PST_ERR_PDB_MODEL_COUNT: return "Only single model PDBs are supported";
    case PST_ERR_PDB_INSERTION_CODE: return "PDB contains an insertion code; these are not supported";
    case PST_ERR_PDB_MALFORMED: return "malformed ATOM / HETATM record";
    case PST_ERR_FILE_NOT_FOUND: return "file could not be opened";
    case PST_ERR_NON_FINITE: return "a latent is Inf / NaN (activations exceeded the 16-bit operand range; use PST_PREC_FP32)";
    default: return "unknown status";
  }
}

size_t pst_weight_blob_floats(const pst_config* cfg) {
  if (!config_ok(cfg)) return 0;
  return pst_fill_weight_pointers(*cfg, nullptr, nullptr);
}

int pst_model_create(const pst_config* cfg, const float* blob_host, size_t blob_floats, int device,
                     pst_model** out) {
  if (!out) return PST_ERR_BAD_ARGUMENT;
  *out = nullptr;
  if (!config_ok(cfg)) return PST_ERR_UNSUPPORTED_CONFIG;
  if (!blob_host) return PST_ERR_BAD_ARGUMENT;
  if (blob_floats != pst_fill_weight_pointers(*cfg, nullptr, nullptr)) return PST_ERR_BAD_WEIGHTS;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return PST_ERR_NO_DEVICE;
  if (device < 0 || device >= ndev) return PST_ERR_BAD_ARGUMENT;
  PST_CUDA_OK(cudaSetDevice(device));
  pst_model* m = new (std::nothrow) pst_model();
  if (!m) return PST_ERR_BAD_ARGUMENT;
  m->cfg = *cfg;
  m->device = device;
  m->launch_count = 0;
  m->tc_dev = nullptr;
  m->blob_host = nullptr;
  m->blob_dev = nullptr;
  m->graphs = new (std::nothrow) PstGraphCache();
  if (m->graphs) { const char* g = getenv("PST_CUDA_GRAPH"); m->graphs->enabled = !(g && g[0] == '0'); }
  { const char* t = getenv("PST_MSG_T"); m->use_msg_t = !(t && t[0] == '0'); }
  { const char* t = getenv("PST_FUSED_RESAMPLER"); m->use_fused_resampler = !(t && t[0] == '0'); }
  { const char* t = getenv("PST_FUSED_FSQ"); m->use_fused_fsq = !(t && t[0] == '0'); }
  m->linear_tc = nullptr;
  m->node_chain = nullptr;
  m->embed_img_dev = nullptr;
  m->table16_dev = nullptr;
  m->layer0_tables = nullptr;
  m->prof_on = false;
  m->prof_n = 0;
  for (int i = 0; i < 2 * PST_PROF_MAX_SPANS; ++i) m->prof_ev[i] = nullptr;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete m; return PST_ERR_CUDA; }
  m->num_sms = prop.multiProcessorCount;
  if (cudaMalloc(&m->blob_dev, blob_floats * sizeof(float)) != cudaSuccess) { delete m; return PST_ERR_CUDA; }
  m->blob_floats = blob_floats;
  if (cudaMemcpy(m->blob_dev, blob_host, blob_floats * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(m->blob_dev); delete m; return PST_ERR_CUDA;
  }
  m->blob_host = static_cast<float*>(malloc(blob_floats * sizeof(float)));
  if (!m->blob_host) { cudaFree(m->blob_dev); delete m; return PST_ERR_BAD_ARGUMENT; }
  memcpy(m->blob_host, blob_host, blob_floats * sizeof(float));
  pst_fill_weight_pointers(m->cfg, m->blob_dev, &m->w);
  // FSQ constants in fp32, operation order of model/quantize.py:177-180
  int basis = 1;
  for (int c = 0; c < PST_C8; ++c) {
    m->half_l[c] = 0.f; m->fsq_offset[c] = 0.f; m->fsq_shift[c] = 0.f; m->basis[c] = 0; m->half_width[c] = 0;
    if (c < cfg->num_levels) {
      int L = cfg->levels[c];
      float half_l = ((float)(L - 1) * (float)(1.0 - 1e-3)) / 2.0f;
      float offset = (L % 2 == 0) ? 0.5f : 0.0f;
      m->half_l[c] = half_l;
      m->fsq_offset[c] = offset;
      m->fsq_shift[c] = tanf(offset / half_l);
      m->basis[c] = basis;
      m->half_width[c] = L / 2;
      basis *= L;
    }
  }
  if (cfg->precision != PST_PREC_FP32) {
    int rc = pst_prepare_tc_weights(m);
    if (rc == PST_OK) rc = pst_prepare_linear_tc(m);
    if (rc == PST_OK) rc = pst_prepare_node_chain(m);
    if (rc == PST_OK) rc = pst_prepare_layer0_tables(m);
    if (rc != PST_OK) { pst_model_destroy(m); return rc; }
  }
  if (cudaDeviceSynchronize() != cudaSuccess) { pst_model_destroy(m); return PST_ERR_CUDA; }
  *out = m;
  return PST_OK;
}

void pst_model_destroy(pst_model* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  if (m->graphs) { m->graphs->clear(); delete m->graphs; }
  if (m->blob_dev) cudaFree(m->blob_dev);
  free(m->blob_host);
  if (m->tc_dev) cudaFree(m->tc_dev);
  pst_destroy_node_chain(m);
  pst_destroy_linear_tc(m);
  if (m->embed_img_dev) cudaFree(m->embed_img_dev);
  if (m->table16_dev) cudaFree(m->table16_dev);
  if (m->layer0_tables) cudaFree(m->layer0_tables);
  for (int i = 0; i < 2 * PST_PROF_MAX_SPANS; ++i)
    if (m->prof_ev[i]) cudaEventDestroy(m->prof_ev[i]);
  delete m;
}

size_t pst_workspace_bytes(const pst_model* m, int total_residues, int num_structures) {
  if (!m || total_residues < 0 || num_structures < 0) return 0;
  if ((long long)total_residues * m->cfg.num_neighbor > (long long)PST_MAX_EDGES_PER_CALL) return 0;
  (void)num_structures;
  return pst_carve_workspace(m, nullptr, total_residues, total_residues).bytes;
}

int pst_last_launch_count(const pst_model* m) { return m ? m->launch_count : 0; }

int pst_profile_enable(const pst_model* m, int enable) {
  if (!m) return PST_ERR_BAD_ARGUMENT;
  PST_CUDA_OK(cudaSetDevice(m->device));
  if (enable && !m->prof_ev[0])
    for (int i = 0; i < 2 * PST_PROF_MAX_SPANS; ++i) PST_CUDA_OK(cudaEventCreate(&m->prof_ev[i]));
  m->prof_on = enable != 0;
  m->prof_n = 0;
  return PST_OK;
}

int pst_profile_collect(const pst_model* m, float* ms_out, int* count_out) {
  if (!m || !ms_out || !count_out) return PST_ERR_BAD_ARGUMENT;
  for (int i = 0; i < m->prof_n; ++i) {
    float ms = 0.f;
    PST_CUDA_OK(cudaEventSynchronize(m->prof_ev[2 * i + 1]));
    PST_CUDA_OK(cudaEventElapsedTime(&ms, m->prof_ev[2 * i], m->prof_ev[2 * i + 1]));
    int k = m->prof_kind[i];
    if (k >= 0 && k < PST_PROF_KINDS) { ms_out[k] += ms; count_out[k] += 1; }
  }
  m->prof_n = 0;
  return PST_OK;
}

int pst_read_status(const pst_model* m, void* stream, void* workspace) {
  if (!m || !workspace) return PST_ERR_BAD_ARGUMENT;
  int32_t v = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PST_CUDA_OK(cudaMemcpyAsync(&v, workspace, sizeof(v), cudaMemcpyDeviceToHost, st));
  PST_CUDA_OK(cudaStreamSynchronize(st));
  return v;
}

static int check_batch(const pst_model* m, const void* offsets, int B, int R, void* ws, size_t ws_bytes, int T) {
  if (!m || !offsets || B < 0 || R < 0 || T < 0 || T > R) return PST_ERR_BAD_ARGUMENT;
  if ((long long)R * m->cfg.num_neighbor > (long long)PST_MAX_EDGES_PER_CALL) return PST_ERR_BAD_ARGUMENT;
  if (!ws || (reinterpret_cast<uintptr_t>(ws) & 255u) != 0) return PST_ERR_BAD_ARGUMENT;  // TMA maps / 32-byte vector loads
  if (ws_bytes < pst_carve_workspace(m, nullptr, R, R).bytes) return PST_ERR_WORKSPACE_TOO_SMALL;
  return PST_OK;
}

int pst_featurize_knn(const pst_model* m, void* stream, const float* atoms, const uint8_t* atom_mask,
                      int atoms_per_residue, const int32_t* offsets, int num_structures, int total_residues,
                      int32_t* senders_out, float* edge_features_out, void* workspace, size_t workspace_bytes) {
  int rc = check_batch(m, offsets, num_structures, total_residues, workspace, workspace_bytes, 0);
  if (rc != PST_OK) return rc;
  if (!atoms || !senders_out || atoms_per_residue < 4) return PST_ERR_BAD_ARGUMENT;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PST_CUDA_OK(cudaSetDevice(m->device));
  PstWorkspace ws = pst_carve_workspace(m, workspace, total_residues, total_residues);
  PST_CUDA_OK(cudaMemsetAsync(ws.status, 0, 4 * sizeof(int32_t), st));
  m->launch_count = pst_launch_featurize(m, st, atoms, atom_mask, atoms_per_residue, offsets, num_structures,
                                         total_residues, senders_out, edge_features_out, ws.prep, ws.cen4, ws.status, ws.redo);
  return cudaGetLastError() == cudaSuccess ? PST_OK : PST_ERR_CUDA;
}

int pst_encode_graph(const pst_model* m, void* stream, const float* edge_features, const int32_t* senders,
                     const int32_t* offsets, const int32_t* token_offsets, int num_structures, int total_residues,
                     int total_tokens, float* latents_out, void* workspace, size_t workspace_bytes) {
  int rc = check_batch(m, offsets, num_structures, total_residues, workspace, workspace_bytes, total_tokens);
  if (rc != PST_OK) return rc;
  if (!edge_features || !senders || !token_offsets || !latents_out) return PST_ERR_BAD_ARGUMENT;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PST_CUDA_OK(cudaSetDevice(m->device));
  PstWorkspace ws = pst_carve_workspace(m, workspace, total_residues, total_residues);
  PST_CUDA_OK(cudaMemsetAsync(ws.status, 0, 4 * sizeof(int32_t), st));
  int n = pst_launch_encode_fp32(m, st, edge_features, senders, offsets, token_offsets, num_structures,
                                 total_residues, total_tokens, latents_out, ws);
  if (n < 0) return n;
  m->launch_count = n;
  return cudaGetLastError() == cudaSuccess ? PST_OK : PST_ERR_CUDA;
}

int pst_quantize(const pst_model* m, void* stream, const float* latents, int n_tokens, int32_t* tokens_out,
                 float* bounded_out) {
  if (!m || !latents || !tokens_out || n_tokens < 0) return PST_ERR_BAD_ARGUMENT;
  PST_CUDA_OK(cudaSetDevice(m->device));
  m->launch_count = pst_launch_quantize(m, static_cast<cudaStream_t>(stream), latents, n_tokens, tokens_out, bounded_out);
  return cudaGetLastError() == cudaSuccess ? PST_OK : PST_ERR_CUDA;
}

int pst_fsq_pack(const pst_model* m, void* stream, const float* bounded, int n_tokens, int32_t* tokens_out) {
  if (!m || !bounded || !tokens_out || n_tokens < 0) return PST_ERR_BAD_ARGUMENT;
  PST_CUDA_OK(cudaSetDevice(m->device));
  m->launch_count = pst_launch_fsq_pack(m, static_cast<cudaStream_t>(stream), bounded, n_tokens, tokens_out);
  return cudaGetLastError() == cudaSuccess ? PST_OK : PST_ERR_CUDA;
}

int pst_indexes_to_codes(const pst_model* m, void* stream, const int32_t* tokens, int n_tokens, float* codes_out) {
  if (!m || !tokens || !codes_out || n_tokens < 0) return PST_ERR_BAD_ARGUMENT;
  PST_CUDA_OK(cudaSetDevice(m->device));
  m->launch_count = pst_launch_indexes_to_codes(m, static_cast<cudaStream_t>(stream), tokens, n_tokens, codes_out);
  return cudaGetLastError() == cudaSuccess ? PST_OK : PST_ERR_CUDA;
}

// Enqueues the fused call on `st` (eagerly, or into the graph being captured on it).
static int tokenize_enqueue(const pst_model* m, cudaStream_t st, const float* atoms, const uint8_t* atom_mask,
                            int atoms_per_residue, const int32_t* offsets, const int32_t* token_offsets,
                            int num_structures, int total_residues, int total_tokens, int32_t* tokens_out,
                            void* workspace) {
  PstWorkspace ws = pst_carve_workspace(m, workspace, total_residues, total_residues);
  PST_CUDA_OK(cudaMemsetAsync(ws.status, 0, 4 * sizeof(int32_t), st));
  // Tensor-core modes: the features travel to the embedding kernel in the compact layout (16 floats per edge; the
  // RBFs are evaluated there in fp32).  The fp32 mode and the two-call path (pst_featurize_knn + pst_encode_graph)
  // use the 27 fp32 features of the reference.
  const int compact = (m->cfg.precision != PST_PREC_FP32 && pst_featurize_compact_ok(m)) ? 1 : 0;
  int count;
  {
    PstSpan span(m, st, 0);
    count = pst_launch_featurize(m, st, atoms, atom_mask, atoms_per_residue, offsets, num_structures,
                                 total_residues, ws.senders, ws.edge_feat, ws.prep, ws.cen4, ws.status, ws.redo, compact);
  }
  if (count < 0) return count;
  // The quantiser (bound, round, pack) is the epilogue of the fused resampler kernels where they run (tensor-core modes):
  // the head that forms a token's latent also emits its int32 id.  Other paths quantise in a launch of their own.
  bool tokens_done = false;
  int n = pst_launch_encode_fp32(m, st, ws.edge_feat, ws.senders, offsets, token_offsets, num_structures,
                                 total_residues, total_tokens, ws.z, ws, compact, m->use_fused_fsq ? tokens_out : nullptr, &tokens_done);
  if (n < 0) return n;
  count += n;
  if (!tokens_done) {
    PstSpan span(m, st, 6);
    count += pst_launch_quantize(m, st, ws.z, total_tokens, tokens_out, nullptr, ws.status);
  }
  m->launch_count = count;
  return cudaGetLastError() == cudaSuccess ? PST_OK : PST_ERR_CUDA;
}

int pst_tokenize(const pst_model* m, void* stream, const float* atoms, const uint8_t* atom_mask,
                 int atoms_per_residue, const int32_t* offsets, const int32_t* token_offsets, int num_structures,
                 int total_residues, int total_tokens, int32_t* tokens_out, void* workspace,
                 size_t workspace_bytes) {
  int rc = check_batch(m, offsets, num_structures, total_residues, workspace, workspace_bytes, total_tokens);
  if (rc != PST_OK) return rc;
  if (!atoms || !token_offsets || !tokens_out || atoms_per_residue < 4) return PST_ERR_BAD_ARGUMENT;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PST_CUDA_OK(cudaSetDevice(m->device));
  // ---- CUDA-graph path: the launch sequence depends only on the arguments below (the host never reads device
  // data), so a call that repeats them replays the graph captured on their second occurrence.  Not used on the
  // legacy / per-thread default streams (capture is not allowed there), while the caller is itself capturing
  // (the launches then go into the caller's graph), or while the profiling spans are on.
  PstGraphCache* gc = m->graphs;
  const bool named_stream = st != nullptr && st != cudaStreamLegacy && st != cudaStreamPerThread;
  if (gc && gc->enabled && named_stream && !m->prof_on) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusNone) {
      PstGraphKey key{atoms, atom_mask, offsets, token_offsets, tokens_out, workspace, workspace_bytes, atoms_per_residue};
      std::lock_guard<std::mutex> lock(gc->mu);
      PstGraphEntry* e = gc->find(key);
      if (!e) {
        gc->insert(key);  // first sighting: run eagerly below, capture if the buffers come back
      } else if (e->exec && e->B == num_structures && e->R == total_residues && e->T == total_tokens) {
        e->last_use = ++gc->clock;
        m->launch_count = e->launches;
        ++gc->n_replay;
        return cudaGraphLaunch(e->exec, st) == cudaSuccess ? PST_OK : PST_ERR_CUDA;
      } else if (!e->failed) {
        e->last_use = ++gc->clock;
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
          const int crc = tokenize_enqueue(m, st, atoms, atom_mask, atoms_per_residue, offsets, token_offsets,
                                           num_structures, total_residues, total_tokens, tokens_out, workspace);
          const cudaError_t end = cudaStreamEndCapture(st, &graph);
          if (crc == PST_OK && end == cudaSuccess && graph) {
            bool ready = false;
            if (e->exec) {  // same buffers, new sizes: retarget the instantiated graph
              cudaGraphExecUpdateResultInfo info;
              ready = cudaGraphExecUpdate(e->exec, graph, &info) == cudaSuccess;
              if (ready) ++gc->n_update;
              if (!ready) {
                cudaGetLastError();
                cudaGraphExecDestroy(e->exec);
                e->exec = nullptr;
              }
            }
            if (!ready) {
              ready = cudaGraphInstantiate(&e->exec, graph, 0) == cudaSuccess;
              if (ready) ++gc->n_instantiate;
            }
            if (ready) {
              e->B = num_structures; e->R = total_residues; e->T = total_tokens;
              e->launches = m->launch_count;
              cudaGraphDestroy(graph);
              return cudaGraphLaunch(e->exec, st) == cudaSuccess ? PST_OK : PST_ERR_CUDA;
            }
          }
          if (graph) cudaGraphDestroy(graph);
        }
        cudaGetLastError();  // a failed capture is not an error of the call: enqueue eagerly from now on
        if (e->exec) cudaGraphExecDestroy(e->exec);
        e->exec = nullptr;
        e->failed = true;
      }
    }
  }
  if (gc) ++gc->n_eager;
  return tokenize_enqueue(m, st, atoms, atom_mask, atoms_per_residue, offsets, token_offsets, num_structures,
                          total_residues, total_tokens, tokens_out, workspace);
}

int pst_graph_cache_stats(const pst_model* m, int* counts4) {
  if (!m || !m->graphs || !counts4) return PST_ERR_BAD_ARGUMENT;
  std::lock_guard<std::mutex> lock(m->graphs->mu);
  counts4[0] = m->graphs->n_replay; counts4[1] = m->graphs->n_update;
  counts4[2] = m->graphs->n_instantiate; counts4[3] = m->graphs->n_eager;
  return PST_OK;
}

int pst_graph_cache_enable(const pst_model* m, int enable) {
  if (!m || !m->graphs) return PST_ERR_BAD_ARGUMENT;
  std::lock_guard<std::mutex> lock(m->graphs->mu);
  m->graphs->enabled = enable != 0;
  if (!enable) m->graphs->clear();
  return PST_OK;
}

}  // extern "C"
