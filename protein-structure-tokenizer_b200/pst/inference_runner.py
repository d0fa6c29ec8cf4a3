"""Drop-in mirror of the reference's `InferenceRunner` (scripts/inference_runner.py:168-324), tokenize side.

Same method names, argument meaning and error behaviour; the JAX/Haiku machinery underneath is
replaced by the CUDA hot path:

  prepare_devices(backend)                 -> (devices, n)     CUDA devices of this process (:169-177)
  prepare_tokenize_fn(cfg, devices)        -> callable         builds the per-device tokenizer once params
                                                               are known (:179-191)
  load_params(model_dir, local_devices)    -> params           `params.npz` with Haiku-style names (:236-248);
                                                               falls back to reference-rule random init when the
                                                               checkpoint directory does not exist and
                                                               allow_random_init=True (HF weights are not
                                                               reachable offline)
  tokenize(random_key, quantize, model_params, pdbs, token_save_path, num_device, data_config,
           batch_size_per_device, logger)                      file loop, `<stem>_tokens.npy` uint32 (1, n_tok)
                                                               (:250-324)
`make_graph_from_pdb` (:40-74) becomes `load_structure`: it parses the file and applies the same length
guards (NotImplementedError) but returns atoms, because featurisation now runs on the GPU.
"""
from __future__ import annotations

import logging
import os
import time
from concurrent.futures import ThreadPoolExecutor
from itertools import cycle, islice
from typing import Any, Callable, List, Optional, Sequence

import numpy as np

from .config import TokenizerConfig
from .pdb import structure_from_pdb_file_native as structure_from_pdb_file  # C++ parser behind the C ABI (pst_parse_pdb)
from .pdb import structure_from_sample_file
from .pdb import structures_from_pdb_files_native as structures_from_pdb_files  # pst_parse_pdb_batch
from .weights import init_params, load_params_npz


def load_structure(pdb_file_path: str, num_neighbor: int, padding_num_residue: int):
    """Counterpart of make_graph_from_pdb's host part: parse + length guards.  Note the reference checks
    nb_residues *before* dropping incomplete residues (scripts/inference_runner.py:52-62); the device path
    additionally needs n_valid >= num_neighbor, which the tokenizer re-checks."""
    return _guard_lengths(structure_from_pdb_file(pdb_file_path), num_neighbor, padding_num_residue)


def load_structures(pdb_file_paths: Sequence[str], num_neighbor: int, padding_num_residue: int, n_threads: int = 0):
    """`load_structure` for a batch of files: ONE call into the library parses them side by side on host threads
    (`pst_parse_pdb_batch`: no GIL, no Python work per file while it runs).  The first bad file in list order raises,
    as a loop over `load_structure` would.  Paths ending in .npy are read as the reference's preprocessed
    `ProteinStructureSample` files (data/protein_structure_sample.py:46-54) instead of being parsed."""
    paths = list(pdb_file_paths)
    is_npy = [p.endswith(".npy") for p in paths]  # preprocessed ProteinStructureSample files need no parsing
    parsed = iter(structures_from_pdb_files([p for p, s in zip(paths, is_npy) if not s], n_threads))
    out = []
    for path, npy in zip(paths, is_npy):
        sample = structure_from_sample_file(path) if npy else next(parsed)
        if isinstance(sample, Exception):
            raise sample
        out.append(_guard_lengths(sample, num_neighbor, padding_num_residue))
    return out


def _guard_lengths(sample, num_neighbor: int, padding_num_residue: int):
    if sample.nb_residues > padding_num_residue:
        raise NotImplementedError(
            f"We currently don't support protein with more than {padding_num_residue} residues"
            f"given: {sample.nb_residues}")
    if sample.nb_residues < num_neighbor:
        raise NotImplementedError(
            f"We currently don't support protein with less than {num_neighbor} residues"
            f"given: {sample.nb_residues}")
    return sample.device_arrays()


def load_and_build_batch(files_paths: Sequence[str], max_seq_len: int, pad_token_id: int) -> np.ndarray:
    """Token files back in: `<stem>_tokens.npy` (uint32 (1, n)) -> int32 [len(files), max_seq_len], truncated to
    max_seq_len and right-padded with pad_token_id (scripts/inference_runner.py:114-133; the pad id comes from
    config/structure_tokenizer/data/ablation_df_*.yaml `pad_token_id`)."""
    out = np.full((len(files_paths), max_seq_len), pad_token_id, np.int32)
    for i, path in enumerate(files_paths):
        seq = np.load(path).astype(np.int32).reshape(1, -1)[:, :max_seq_len]
        out[i, : seq.shape[-1]] = seq[0]
    return out


def _leaf(graph, name):
    return graph[name] if isinstance(graph, dict) else getattr(graph, name)


def is_padded_graph(batch) -> bool:
    """True for the reference's own batch: a `ProteinGraph` (types.py:48-75), a `BatchDataVQ3D` wrapping one
    (types.py:78-87, what data_pipeline.py saves) or a dict / object with the same leaves."""
    g = getattr(batch, "graph", batch)
    if isinstance(g, dict):
        return all(k in g for k in ("edge_features", "senders", "n_node"))
    return all(hasattr(g, k) for k in ("edge_features", "senders", "n_node"))


def padded_graph_to_ragged(batch, num_neighbor: int):
    """`ProteinGraph` leaves with any leading batch dims ([Dev, B, ...] after `batch_collate`,
    scripts/inference_runner.py:77-83; [...] for one structure) -> the ragged B1 inputs.

    Padding layout (data/preprocessing.py:191-283): the n_node valid residues come first, their K edges each occupy
    the first n_node * K rows of `edge_features` / `senders` (receivers = repeat(arange(n), K)); the rest is padding
    (zero features, self loops), which never influences a valid token.  Returns (lead_shape, n_valid[Btot],
    edge_features f32 [sum n*K, 27], senders i32 [sum n*K], offsets i32 [Btot+1], padded_residues N)."""
    g = getattr(batch, "graph", batch)
    K = int(num_neighbor)
    ef = np.asarray(_leaf(g, "edge_features"))
    se = np.asarray(_leaf(g, "senders"))
    nn = np.asarray(_leaf(g, "n_node"))
    if ef.ndim < 2 or ef.shape[-1] != 27:
        raise ValueError(f"edge_features must end in [N*K, 27], got {ef.shape}")
    lead = ef.shape[:-2]
    EK = ef.shape[-2]
    if EK % K:
        raise ValueError(f"{EK} padded edges is not a multiple of num_neighbor = {K}")
    ef = ef.reshape(-1, EK, 27)
    se = se.reshape(-1, EK)
    n_valid = nn.reshape(-1).astype(np.int64)
    if n_valid.shape[0] != ef.shape[0]:
        raise ValueError(f"n_node has {n_valid.shape[0]} entries for {ef.shape[0]} structures")
    if (n_valid * K > EK).any():
        raise ValueError("n_node exceeds the padded size")
    offsets = np.zeros(len(n_valid) + 1, np.int32)
    offsets[1:] = np.cumsum(n_valid)
    feats = np.concatenate([ef[b, : n_valid[b] * K] for b in range(len(n_valid))]).astype(np.float32)
    send = np.concatenate([se[b, : n_valid[b] * K] for b in range(len(n_valid))]).astype(np.int32)
    return lead, n_valid, np.ascontiguousarray(feats), np.ascontiguousarray(send), offsets, EK // K


class _TokenizeFn:
    """The callable `prepare_tokenize_fn` returns.  quantize(params, rng, batch) -> {"tokens": ...}.

    `batch` is either a list of (atoms, mask) pairs, one per structure (boundary B2: featurisation on the GPU; tokens
    come back as a list of uint32 arrays), or the reference's own padded `ProteinGraph` / `BatchDataVQ3D` with leading
    dims [Dev, B] (boundary B1, scripts/inference_runner.py:299-306; tokens come back as uint32 [Dev, B, T] with the
    masked-token code in the padded tail, exactly the array the reference's loop slices with `tokens_mask`)."""

    def __init__(self, cfg: TokenizerConfig, devices: Sequence[int]):
        self.cfg = cfg
        self.devices = list(devices)
        self._tok = {}
        self._params_id = None

    def _get(self, params, device: int):
        from .tokenizer import StructureTokenizer

        if self._params_id != id(params):
            for t in self._tok.values():
                t.close()
            self._tok = {}
            self._params_id = id(params)
        if device not in self._tok:
            self._tok[device] = StructureTokenizer(self.cfg, params, device=device)
        return self._tok[device]

    def _run_sharded(self, params, lengths, run_one):
        """Splits structure indices over this process's devices by LPT on the length cost model (ragged batches
        balance; the reference's [Dev, B] reshape is a contiguous split) and runs `run_one(tokenizer, indices)` on
        one host thread per device, like pmap's per-device dispatch (the GIL is released inside CUDA calls)."""
        from .distributed import lpt_partition

        shards = [(dev, idx) for dev, idx in zip(self.devices, lpt_partition(lengths, len(self.devices))) if idx]
        for dev, _ in shards:  # tokenizers are created on the calling thread (one model per device)
            self._get(params, dev)
        if len(shards) > 1:
            with ThreadPoolExecutor(max_workers=len(shards)) as pool:
                outs = list(pool.map(lambda s: run_one(self._tok[s[0]], s[1]), shards))
        else:
            outs = [run_one(self._tok[dev], idx) for dev, idx in shards]
        return [(idx, out) for (_, idx), out in zip(shards, outs)]

    def _call_padded(self, params, batch):
        cfg = self.cfg
        K, df = cfg.num_neighbor, cfg.downsampling_ratio
        lead, n_valid, feats, send, offsets, n_pad = padded_graph_to_ragged(batch, K)
        if n_pad != cfg.seq_max_size:
            raise ValueError(f"graph padded to {n_pad} residues, config says seq_max_size = {cfg.seq_max_size}")
        if (n_valid < K).any():
            raise NotImplementedError(f"We currently don't support protein with less than {K} residues")
        T_pad = n_pad // df

        def run_one(tok, idx):
            sub_off = np.zeros(len(idx) + 1, np.int32)
            sub_off[1:] = np.cumsum(n_valid[idx])
            f = np.concatenate([feats[offsets[i] * K : offsets[i + 1] * K] for i in idx])
            s = np.concatenate([send[offsets[i] * K : offsets[i + 1] * K] for i in idx])
            return tok.encode_graph(f, s, sub_off)

        tokens = np.empty((len(n_valid), T_pad), np.uint32)
        for idx, outs in self._run_sharded(params, n_valid.tolist(), run_one):
            pad_code = next(iter(self._tok.values())).masked_token_code()
            for i, t in zip(idx, outs):
                tokens[i, : t.size] = t
                tokens[i, t.size :] = pad_code
        return {"tokens": tokens.reshape(*lead, T_pad)}

    def __call__(self, params, rng, batch):
        del rng  # inference is deterministic (is_training=False: dropout off)
        if is_padded_graph(batch):
            return self._call_padded(params, batch)
        tokens: List[Optional[np.ndarray]] = [None] * len(batch)

        def run_one(tok, idx):
            return tok.tokenize([batch[i][0] for i in idx], [batch[i][1] for i in idx])

        for idx, outs in self._run_sharded(params, [int(a.shape[0]) for a, _ in batch], run_one):
            for i, t in zip(idx, outs):
                tokens[i] = t
        return {"tokens": tokens}


class InferenceRunner:
    @staticmethod
    def prepare_devices(backend: str = "gpu"):
        import torch

        if backend not in ("gpu", "cuda"):
            raise RuntimeError(f"backend '{backend}' is not available: this implementation runs on CUDA devices only "
                               "(there is no CPU fallback; the CPU path is the reference itself)")
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device visible")
        n = torch.cuda.device_count()
        devices = list(range(n))
        print("---Devices---\n" + f"\tlocal device count: {n}")
        return devices, n

    @staticmethod
    def prepare_tokenize_fn(cfg: Any, devices: Sequence[int], precision: str = "fp16") -> Callable:
        tcfg = cfg if isinstance(cfg, TokenizerConfig) else TokenizerConfig.from_reference_cfg(cfg, precision=precision)
        return _TokenizeFn(tcfg, devices)

    @staticmethod
    def load_params(model_dir: str, local_devices: Sequence[int] = (), cfg: Optional[TokenizerConfig] = None,
                    allow_random_init: bool = False, seed: int = 0):
        """Looks for `<model_dir>/params_named.npz` (what INTEGRATION.md's flatten recipe writes), then `params.npz`.
        A released checkpoint's `params.npz` holds positional leaves (`arr_0`, ...) whose names live in the pickled
        jax PyTreeDef of `state_variables.npy` (scripts/inference_runner.py:136-150,236-248): that pair cannot be
        read without jax, so it is recognised and rejected with the conversion step instead of failing later."""
        named = os.path.join(model_dir, "params_named.npz")
        path = named if os.path.exists(named) else os.path.join(model_dir, "params.npz")
        if os.path.exists(path):
            with np.load(path) as f:
                keys = list(f.files)
            positional = bool(keys) and all(k.startswith("arr_") and k[4:].isdigit() for k in keys)
            if positional:
                raise ValueError(
                    f"{path} holds positional leaves ({keys[0]} ... {keys[-1]}), i.e. a released checkpoint whose "
                    "parameter names are stored in the jax PyTreeDef of state_variables.npy.  Flatten it once where jax "
                    "is installed (INTEGRATION.md, section 'Checkpoints') and put the resulting params_named.npz next to it.")
            return load_params_npz(path)
        if allow_random_init and cfg is not None:
            return init_params(cfg, seed, "spread")
        raise FileNotFoundError(
            f"neither {named} nor {os.path.join(model_dir, 'params.npz')} found.  Expected an .npz of Haiku-named arrays "
            "(see pst/weights.py); the released checkpoints store a pickled jax treedef next to params.npz and must be "
            "flattened to names first (INTEGRATION.md).")

    @staticmethod
    def tokenize(random_key, quantize: Callable, model_params, pdbs: List[str], token_save_path: str, num_device: int,
                 data_config: Any, batch_size_per_device: int = 8, logger: Optional[logging.Logger] = None):
        """Same loop as scripts/inference_runner.py:250-324.  Under `torchrun` (an initialised torch.distributed group
        of world size W) the file list is sharded over the ranks by LPT on file size, each rank runs the loop on its
        shard with its own GPU(s), and the tokens are gathered over NCCL to rank 0, which writes every file: the one
        collective of the job (pst/distributed.py)."""
        from .distributed import dist_env, gather_tokens, lpt_partition

        rank, world, _ = dist_env()
        if logger is not None:
            logger.info(f"Starting tokenization of {pdbs}")
        if rank == 0:
            os.makedirs(token_save_path, exist_ok=False)  # same behaviour: an existing directory aborts the run
        effective_batch_size = batch_size_per_device * num_device
        all_pdbs = list(pdbs)
        if world > 1:
            sizes = [max(1, os.path.getsize(f) // 324) if os.path.exists(f) else 1 for f in all_pdbs]  # ~ residues (all-atom text)
            mine = lpt_partition(sizes, world)[rank]
        else:
            mine = list(range(len(all_pdbs)))
        local = [all_pdbs[i] for i in mine]
        num_iteration = len(local) // effective_batch_size + int((len(local) % effective_batch_size) > 0)
        total = num_iteration * effective_batch_size
        n_real = len(local)
        local = list(islice(cycle(local), total)) if local else []  # repeat the list to a multiple of the batch size
        # Host pipeline around the device call: the next batch is read and parsed (C++ parser, the GIL is released
        # inside the ctypes call) and the previous batch's files are written while the GPU works on the current one.
        # Errors keep the reference's order: a batch's parse error is raised when that batch's turn comes.
        # The files of one batch are parsed side by side inside ONE library call (pst_parse_pdb_batch, host threads, no
        # GIL): a single parser thread delivers ~0.8 M residues/s, a B200 tokenizes 16 M/s.  The first bad file of a
        # batch, in list order, is still the one that raises.
        n_parse = max(1, min(16, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1),
                             effective_batch_size))

        def load(it: int):
            files = local[it * effective_batch_size : (it + 1) * effective_batch_size]
            return files, load_structures(files, data_config.graph_max_neighbor, data_config.seq_max_size, n_parse)

        def save(files, tokens):
            for f, tok in zip(files, tokens):
                name = os.path.basename(f).split(".pdb")[0]  # scripts/inference_runner.py:316
                for ext in (".npy", ".cif", ".mmcif"):  # sample files and mmCIF files: the stem names the token file too
                    if name.endswith(ext):
                        name = name[: -len(ext)]
                np.save(os.path.join(token_save_path, name + "_tokens"), np.asarray(tok, np.uint32).reshape(1, -1))

        kept: List[np.ndarray] = []  # world > 1: this rank's tokens, in shard order, for the gather
        with ThreadPoolExecutor(max_workers=2) as pool:
            nxt = pool.submit(load, 0) if num_iteration else None
            pending_save = None
            for it in range(num_iteration):
                t0 = time.perf_counter()
                files, batch = nxt.result()
                nxt = pool.submit(load, it + 1) if it + 1 < num_iteration else None
                out = quantize(model_params, random_key, batch)
                if world > 1:
                    kept.extend(out["tokens"])
                else:
                    if pending_save is not None:
                        pending_save.result()
                    pending_save = pool.submit(save, files, out["tokens"])
                if logger is not None:
                    logger.info(f"Took {time.perf_counter() - t0}s to tokenize")
            if pending_save is not None:
                pending_save.result()
        if world > 1:
            import torch

            dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() and \
                torch.distributed.get_backend() == "nccl" else None
            everything = gather_tokens(mine, kept[:n_real], len(all_pdbs), rank, world, device=dev)
            if rank == 0:
                save(all_pdbs, everything)
            torch.distributed.barrier()
