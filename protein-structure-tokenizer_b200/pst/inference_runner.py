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


class _TokenizeFn:
    """The callable `prepare_tokenize_fn` returns.  quantize(params, rng, batch) -> {"tokens": [...]};
    `batch` is a list of (atoms, mask) pairs (one per structure)."""

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

    def __call__(self, params, rng, batch):
        del rng  # inference is deterministic (is_training=False: dropout off)
        n_dev = len(self.devices)
        tokens: List[Optional[np.ndarray]] = [None] * len(batch)
        # structures are independent: contiguous shards, one per device (the reference's [Dev, B] reshape)
        per = (len(batch) + n_dev - 1) // n_dev
        shards = [(d, dev, batch[d * per : (d + 1) * per]) for d, dev in enumerate(self.devices)]
        shards = [s for s in shards if s[2]]
        for _, dev, _ in shards:  # tokenizers are created on the calling thread (one model per device)
            self._get(params, dev)

        def run(shard):
            d, dev, part = shard
            return d, self._tok[dev].tokenize([a for a, _ in part], [m for _, m in part])

        if len(shards) > 1:
            # one host thread per device, like pmap's per-device dispatch: the devices work concurrently (the GIL is
            # released inside the CUDA calls and copies)
            with ThreadPoolExecutor(max_workers=len(shards)) as pool:
                results = list(pool.map(run, shards))
        else:
            results = [run(s) for s in shards]
        for d, out in results:
            tokens[d * per : d * per + len(out)] = out
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
        path = os.path.join(model_dir, "params.npz")
        if os.path.exists(path):
            return load_params_npz(path)
        if allow_random_init and cfg is not None:
            return init_params(cfg, seed, "spread")
        raise FileNotFoundError(
            f"{path} not found.  Expected an .npz of Haiku-named arrays (see pst/weights.py); the released "
            "checkpoints store a pickled jax treedef next to params.npz and must be flattened to names first "
            "(INTEGRATION.md).")

    @staticmethod
    def tokenize(random_key, quantize: Callable, model_params, pdbs: List[str], token_save_path: str, num_device: int,
                 data_config: Any, batch_size_per_device: int = 8, logger: Optional[logging.Logger] = None):
        if logger is not None:
            logger.info(f"Starting tokenization of {pdbs}")
        os.makedirs(token_save_path, exist_ok=False)  # same behaviour: an existing directory aborts the run
        effective_batch_size = batch_size_per_device * num_device
        num_iteration = len(pdbs) // effective_batch_size + int((len(pdbs) % effective_batch_size) > 0)
        total = num_iteration * effective_batch_size
        pdbs = list(islice(cycle(pdbs), total))  # repeat the list to a multiple of the batch size
        # Host pipeline around the device call: the next batch is read and parsed (C++ parser, the GIL is released
        # inside the ctypes call) and the previous batch's files are written while the GPU works on the current one.
        # Errors keep the reference's order: a batch's parse error is raised when that batch's turn comes.
        # The files of one batch are parsed side by side inside ONE library call (pst_parse_pdb_batch, host threads, no
        # GIL): a single parser thread delivers ~0.8 M residues/s, a B200 tokenizes 16 M/s.  The first bad file of a
        # batch, in list order, is still the one that raises.
        n_parse = max(1, min(16, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1),
                             effective_batch_size))

        def load(it: int):
            files = pdbs[it * effective_batch_size : (it + 1) * effective_batch_size]
            return files, load_structures(files, data_config.graph_max_neighbor, data_config.seq_max_size, n_parse)

        def save(files, tokens):
            for f, tok in zip(files, tokens):
                name = os.path.basename(f).split(".pdb")[0]  # scripts/inference_runner.py:316
                if name.endswith(".npy"):
                    name = name[: -len(".npy")]
                np.save(os.path.join(token_save_path, name + "_tokens"), np.asarray(tok, np.uint32).reshape(1, -1))

        with ThreadPoolExecutor(max_workers=2) as pool:
            nxt = pool.submit(load, 0) if num_iteration else None
            pending_save = None
            for it in range(num_iteration):
                t0 = time.perf_counter()
                files, batch = nxt.result()
                nxt = pool.submit(load, it + 1) if it + 1 < num_iteration else None
                out = quantize(model_params, random_key, batch)
                if pending_save is not None:
                    pending_save.result()
                pending_save = pool.submit(save, files, out["tokens"])
                if logger is not None:
                    logger.info(f"Took {time.perf_counter() - t0}s to tokenize")
            if pending_save is not None:
                pending_save.result()
