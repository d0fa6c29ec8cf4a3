/*
 * pst_abi.h -- C ABI of the B200-native structure-tokenization hot path.
 *
 * The reference (xwang112358/protein-structure-tokenizer) is pure Python/JAX and has
 * no FFI for this path; these entry points are what a JAX FFI custom call (or any
 * ctypes / cffi host) binds in place of the reference functions cited on each one.
 * File:line citations are relative to the reference repository root.
 *
 * Conventions
 *   - Every array pointer is a DEVICE pointer owned by the caller unless its name
 *     ends in `_host`.  The library never allocates device memory inside a hot call
 *     (the caller passes a workspace sized by pst_workspace_bytes), never synchronises,
 *     and only enqueues work on the given cudaStream_t (passed as void*): kernel by
 *     kernel, or as one CUDA-graph launch when pst_tokenize sees a repeated argument
 *     set (pst_graph_cache_enable below).
 *   - Structures are stored ragged and concatenated: `offsets[B+1]` (int32, device)
 *     gives the first residue row of each structure; R = offsets[B] rows in total.
 *     Only VALID residues are stored (N, CA, C and O all present; the host drops the
 *     others exactly like data/preprocessing.py:99-117).  Each structure must have
 *     K <= L_b <= max_len residues (the reference raises NotImplementedError outside
 *     [K, 512]: scripts/inference_runner.py:52-62); the host checks this, the kernels
 *     additionally raise the device status word (pst_read_status).
 *   - Neighbour indices (`senders`) are LOCAL to the structure (0 .. L_b-1), exactly
 *     the values the reference stores in ProteinGraph.senders (types.py:48-75);
 *     `receivers` is implicit: edge e of row i is (i*K + e), receivers = repeat(arange).
 *   - Tokens are ragged as well: structure b owns floor(L_b / df) tokens starting at
 *     token_offsets[b].
 *   - Return value: 0 on success, a negative pst_status otherwise.  No exceptions,
 *     no global mutable state: apart from its mutex-guarded graph cache a pst_model
 *     is immutable after creation and may be used from one host thread per device.
 */
#ifndef PST_ABI_H_
#define PST_ABI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PST_ABI_VERSION 1
#define PST_CHANNELS 128       /* encoding / hidden / out_emb size of every released config */
#define PST_EDGE_FEATURES 27   /* 15 RBF + p,q,k,t (utils/protein_utils.py:403-434) */
#define PST_MAX_LEVELS 8

typedef enum pst_status {
  PST_OK = 0,
  PST_ERR_BAD_ARGUMENT = -1,
  PST_ERR_UNSUPPORTED_CONFIG = -2,
  PST_ERR_LENGTH_OUT_OF_RANGE = -3, /* a structure has L < K or L > max_len */
  PST_ERR_WORKSPACE_TOO_SMALL = -4,
  PST_ERR_CUDA = -5,
  PST_ERR_NO_DEVICE = -6,
  PST_ERR_BAD_WEIGHTS = -7,
  PST_ERR_PDB_MODEL_COUNT = -8,     /* the file does not hold exactly one model */
  PST_ERR_PDB_INSERTION_CODE = -9,  /* a residue carries an insertion code */
  PST_ERR_PDB_MALFORMED = -10,      /* an ATOM / HETATM record is too short or has non-numeric fields */
  PST_ERR_FILE_NOT_FOUND = -11,     /* pst_parse_pdb_files: the path could not be opened */
  PST_ERR_NON_FINITE = -12          /* device status: a latent reached the quantiser as Inf / NaN (16-bit operand range exceeded) */
} pst_status;

/* GEMM operand precision of the edge-level MLPs (accumulation is always fp32;
 * node-level linears, LayerNorms, softmax and the quantiser always run in fp32). */
typedef enum pst_precision {
  PST_PREC_FP32 = 0, /* CUDA-core fp32 everywhere: on-device reference mode          */
  PST_PREC_FP16 = 1, /* tcgen05 kind::f16, fp16 operands (default; >= 99.5 % tokens) */
  PST_PREC_BF16 = 2  /* tcgen05 kind::f16, bf16 operands (the north-star dtype)      */
} pst_precision;

/* Hyper-parameters the tokenize path reads:
 * config/structure_tokenizer/data/ablation_df_*.yaml:15-23 (seq_max_size, graph_max_neighbor,
 * downsampling_ratio), model/gnn/ablation_*_df_*.yaml (max_out_len, codebook.levels),
 * model/shared.yaml (gnn_number_layers, sc_num_block, num_head). */
typedef struct pst_config {
  int32_t abi_version;        /* PST_ABI_VERSION */
  int32_t seq_max_size;       /* base n of node/edge positional encodings (512)          */
  int32_t max_out_len;        /* base n of the resampled-token positional encoding       */
  int32_t num_neighbor;       /* K = graph_max_neighbor (50)                             */
  int32_t downsampling_ratio; /* df in {1,2,4}                                           */
  int32_t num_levels;         /* C = len(codebook.levels) (5 or 6)                       */
  int32_t levels[PST_MAX_LEVELS];
  int32_t gnn_layers;         /* 3 */
  int32_t num_blocks;         /* sc_num_block = 3 */
  int32_t precision;          /* pst_precision */
  int32_t max_len;            /* longest structure accepted, <= seq_max_size             */
} pst_config;

typedef struct pst_model pst_model; /* opaque */

/* Number of fp32 values in the prepared weight blob for `cfg` (layout: see
 * protein-structure-tokenizer_b200/pst/weights.py, which builds it from Haiku-named arrays). */
size_t pst_weight_blob_floats(const pst_config* cfg);

/* Replaces InferenceRunner.load_params + jax.device_put_replicated for one device
 * (scripts/inference_runner.py:236-248): uploads the prepared fp32 blob (host pointer)
 * to `device` and builds the operand-precision copies of the edge-MLP weights. */
int pst_model_create(const pst_config* cfg, const float* blob_host, size_t blob_floats,
                     int device, pst_model** out);
void pst_model_destroy(pst_model* model);

/* Largest batch one hot call accepts: total_residues * num_neighbor edges.  Kernels index edge rows and
 * their feature floats with 32-bit integers; beyond this the call returns PST_ERR_BAD_ARGUMENT (split the
 * batch: pst/tokenizer.py sends chunks of <= 131 072 residues, i.e. 6.5 M edges at K = 50). */
#define PST_MAX_EDGES_PER_CALL (1 << 26)

/* Bytes of caller-provided scratch needed by any hot call below for a batch of
 * `total_residues` rows in `num_structures` structures (0 if the batch exceeds PST_MAX_EDGES_PER_CALL).
 * The workspace pointer must be 256-byte aligned (cudaMalloc, torch and XLA allocations are); a misaligned
 * one is refused with PST_ERR_BAD_ARGUMENT. */
size_t pst_workspace_bytes(const pst_model* model, int total_residues, int num_structures);

/* Replaces the host featuriser: frames (model/quat_affine.py:406-522), centroid and
 * k-NN (utils/protein_utils.py:373-399), RBF + orientation features (:257-281,:403-434)
 * as called from data/preprocessing.py:85-189.
 *   atoms        f32 [R, atoms_per_residue, 3]; slots 0,1,2 = N, CA, C.  atoms_per_residue
 *                is 4 (N,CA,C,O: synthetic / backbone-only input) or 37 (atom37 layout).
 *   atom_mask    u8  [R, atoms_per_residue] (gt_exists & atom_exists, preprocessing.py:72),
 *                or NULL = every slot present.
 *   senders_out  i32 [R*K]  local neighbour indices, ascending (distance, index)   BIT-EXACT
 *   edge_features_out f32 [R*K, 27]  (fp64 arithmetic, rounded once to fp32), may be NULL */
int pst_featurize_knn(const pst_model* model, void* stream, const float* atoms,
                      const uint8_t* atom_mask, int atoms_per_residue, const int32_t* offsets,
                      int num_structures, int total_residues, int32_t* senders_out,
                      float* edge_features_out, void* workspace, size_t workspace_bytes);

/* Replaces the pmapped callable Vq3D.encode (model/model.py:357-420) on a ProteinGraph
 * (boundary B1): input embeddings (model/structure_encoder.py:77-105), 3 MPNN layers
 * (model/gnn_layers.py:325-438), CrossAttentionScaler (model/modules.py:438-636), spherical
 * norm + down_proj (model/model.py:169-174,148-164).
 *   latents_out  f32 [T_total, PST_MAX_LEVELS]  pre-quantisation z, first C columns valid. */
int pst_encode_graph(const pst_model* model, void* stream, const float* edge_features,
                     const int32_t* senders, const int32_t* offsets, const int32_t* token_offsets,
                     int num_structures, int total_residues, int total_tokens,
                     float* latents_out, void* workspace, size_t workspace_bytes);

/* Replaces FiniteScalarCodebook.__call__ (model/quantize.py:141-209): bound, round half to
 * even, mixed-radix pack.  z is f32 [n_tokens, PST_MAX_LEVELS]; tokens_out i32 [n_tokens].
 * bounded_out (f32 [n_tokens, PST_MAX_LEVELS], the reference's `continuous_embedding`) may be NULL. */
int pst_quantize(const pst_model* model, void* stream, const float* latents, int n_tokens,
                 int32_t* tokens_out, float* bounded_out);

/* The integer half of the quantiser alone (model/quantize.py:188,209,113-120): round + pack
 * already-bounded values.  BIT-EXACT. */
int pst_fsq_pack(const pst_model* model, void* stream, const float* bounded, int n_tokens,
                 int32_t* tokens_out);

/* Inverse of the packing (model/quantize.py:122-139, renorm=False): tokens -> integer codes
 * f32 [n_tokens, PST_MAX_LEVELS]. */
int pst_indexes_to_codes(const pst_model* model, void* stream, const int32_t* tokens,
                         int n_tokens, float* codes_out);

/* Boundary B2 -- replaces make_graph_from_pdb's featurisation plus the pmapped callable
 * Vq3D.encode_and_quantize (scripts/inference_runner.py:288-305; model/model.py:453-479):
 * atoms in, token ids out.  tokens_out i32 [total_tokens]. */
int pst_tokenize(const pst_model* model, void* stream, const float* atoms,
                 const uint8_t* atom_mask, int atoms_per_residue, const int32_t* offsets,
                 const int32_t* token_offsets, int num_structures, int total_residues,
                 int total_tokens, int32_t* tokens_out, void* workspace, size_t workspace_bytes);

/* CUDA-graph cache of pst_tokenize (on by default; environment PST_CUDA_GRAPH=0 or enable = 0 turns it off and drops
 * the cached graphs).  The launch sequence of pst_tokenize depends only on its arguments (the host never reads device
 * data), so when a call repeats the pointers and sizes of an earlier one on a named stream, the library captures the
 * sequence into a CUDA graph once (on the second occurrence) and replays it afterwards: the dependent kernels of a
 * call (16 in the tensor-core modes at df = 1) otherwise pay a launch gap each.  The cache is keyed by the BUFFERS of a call, up to 64 sets per model (LRU); when the same buffers come back with other sizes (the next ragged chunk of a stream) the instantiated graph is updated in place (cudaGraphExecUpdate);
 * the cache is the only mutable state of a pst_model and is guarded by a mutex.  Calls on the legacy / per-thread
 * default stream, calls made while the caller is itself capturing `stream`, and calls with profiling enabled are
 * enqueued kernel by kernel as before. */
int pst_graph_cache_enable(const pst_model* model, int enable);

/* How the pst_tokenize calls of this model were enqueued so far: counts4 = {graph replays with unchanged sizes, graph
 * launches after an in-place update to new sizes, launches after a fresh instantiation, kernel-by-kernel calls}. */
int pst_graph_cache_stats(const pst_model* model, int* counts4);

/* HOST function (no CUDA): PDB text -> atom37 arrays.  Replaces protein_structure_from_pdb_string
 * (structure_tokenizer/data/protein_structure_sample.py:166-248) together with the BioPython PDBParser semantics it
 * relies on (single model, chains and residues in order of first appearance, highest-occupancy altloc, atoms
 * outside atom37 dropped, unknown residues -> UNK, insertion codes rejected).  Outputs are caller-owned host
 * arrays with room for max_residues residues: positions f32 [max,37,3], gt_exists / atom_exists u8 [max,37],
 * aatype i32 [max] (20 = UNK).  Call with the array pointers NULL to get *n_residues_out only.
 * Returns PST_OK, PST_ERR_PDB_* or PST_ERR_WORKSPACE_TOO_SMALL (n_residues_out is then the needed count). */
int pst_parse_pdb(const char* text, size_t text_bytes, int max_residues, float* atom37_positions,
                  uint8_t* gt_exists, uint8_t* atom_exists, int32_t* aatype, int32_t* n_residues_out);

/* The same with the reference's `chain_id` argument (protein_structure_sample.py:166-168,201-203; the second caller,
 * data_pipeline.py:69-128, passes it): only the chain with this one-character id is emitted; 0 = every chain
 * (= pst_parse_pdb).  An insertion code in a chain that is skipped is not an error, as in the reference. */
int pst_parse_pdb_chain(const char* text, size_t text_bytes, char chain_id, int max_residues, float* atom37_positions,
                        uint8_t* gt_exists, uint8_t* atom_exists, int32_t* aatype, int32_t* n_residues_out);

/* mmCIF (the next ingest format; the reference itself reads PDB text and sample .npy files only).
 * The `_atom_site` loop goes through the same selection as the PDB records, with BioPython's MMCIFParser conventions
 * where the formats differ (chain = auth_asym_id, residue number = auth_seq_id, atom name = label_atom_id, altloc /
 * insertion code '.' / '?' = none, one pdbx_PDB_model_num).  chain_id: NULL or "" = every chain; ids may be longer than
 * one character.  pst_parse_pdb, pst_parse_pdb_chain, pst_parse_pdb_batch and pst_parse_pdb_files accept mmCIF text as
 * well: a text whose first token is a `data_` block header is parsed as mmCIF. */
int pst_parse_mmcif(const char* text, size_t text_bytes, const char* chain_id, int max_residues, float* atom37_positions,
                    uint8_t* gt_exists, uint8_t* atom_exists, int32_t* aatype, int32_t* n_residues_out);

/* HOST function: n_files PDB texts parsed side by side on up to n_threads host threads (<= 0: all hardware threads),
 * the feeder in front of pst_tokenize (the reference parses one file at a time in Python,
 * scripts/inference_runner.py:40-74 called from :288-296).  The output arrays are shared: file i's residues are rows
 * residue_offsets_out[i] .. residue_offsets_out[i+1] (n_files + 1 entries); status_out[i] is PST_OK or that file's
 * PST_ERR_PDB_* (it then contributes no rows).  Call with the array pointers NULL to get the offsets only.
 * Returns PST_OK, PST_ERR_BAD_ARGUMENT or PST_ERR_WORKSPACE_TOO_SMALL (residue_offsets_out[n_files] rows are needed). */
int pst_parse_pdb_batch(const char* const* texts, const size_t* text_bytes, int n_files, int n_threads,
                        int max_residues_total, float* atom37_positions, uint8_t* gt_exists, uint8_t* atom_exists,
                        int32_t* aatype, int32_t* residue_offsets_out, int32_t* status_out);

/* The same over files: the worker threads read paths[i] themselves (a PST_ERR_WORKSPACE_TOO_SMALL retry reads
 * and parses the files again).  status_out[i] may also be PST_ERR_FILE_NOT_FOUND. */
int pst_parse_pdb_files(const char* const* paths, int n_files, int n_threads, int max_residues_total,
                        float* atom37_positions, uint8_t* gt_exists, uint8_t* atom_exists, int32_t* aatype,
                        int32_t* residue_offsets_out, int32_t* status_out);

/* Device status word raised by kernels (0 = fine, PST_ERR_LENGTH_OUT_OF_RANGE ...).
 * Synchronises `stream`; not part of the hot path. */
int pst_read_status(const pst_model* model, void* stream, void* workspace);

/* How many kernels of this library the last hot call enqueued (for bench.py's gpu_launches). */
int pst_last_launch_count(const pst_model* model);

/* Optional per-kernel-group timing for bench.py's roofline: when enabled, the hot calls record
 * CUDA event pairs on their stream around (kind 0) featurise + k-NN, (1) message-MLP launches,
 * (2) edge-update-MLP launches, (3) node updates, (4) input embeddings, (5) the fused df = 1 resampler + head, (6) the FSQ quantiser.
 * pst_profile_collect synchronises those events, adds the elapsed milliseconds and launch-group counts per
 * kind into ms_out[8] / count_out[8], and resets. */
int pst_profile_enable(const pst_model* model, int enable);
int pst_profile_collect(const pst_model* model, float* ms_out_host, int* count_out_host);

const char* pst_status_string(int status);
int pst_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PST_ABI_H_ */
