"""Host-side driver of the CUDA hot path: owns the `pst_model` handle and a grow-only
workspace, and exposes the two drop-in boundaries of SURVEY section 8b on top of the C ABI:

  B1  `encode_graph`  ProteinGraph arrays (edge_features, senders)  -> latents / tokens
      (the callable the reference builds in InferenceRunner.prepare_tokenize_fn,
       scripts/inference_runner.py:179-191)
  B2  `tokenize`      atoms (backbone or atom37) -> token ids
      (make_graph_from_pdb's featurisation + that callable, scripts/inference_runner.py:288-305)

torch is used for device memory and streams only.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .config import PRECISIONS, TokenizerConfig
from .weights import pack_weights


def _ptr(t) -> Optional[int]:
    return None if t is None else t.data_ptr()


class StructureTokenizer:
    def __init__(self, cfg: TokenizerConfig, params, device: int = 0, max_rows_per_call: int = 131072):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("StructureTokenizer needs a CUDA device (sm_100a); there is no CPU fallback")
        self.torch = torch
        self.cfg = cfg
        self.device = torch.device("cuda", device)
        self.lib = _lib.load()
        self.max_rows_per_call = int(max_rows_per_call)
        c = _lib.PstConfig()
        c.abi_version = _lib.PST_ABI_VERSION
        c.seq_max_size, c.max_out_len = cfg.seq_max_size, cfg.max_out_len
        c.num_neighbor, c.downsampling_ratio = cfg.num_neighbor, cfg.downsampling_ratio
        c.num_levels = len(cfg.levels)
        for i, l in enumerate(cfg.levels):
            c.levels[i] = int(l)
        c.gnn_layers, c.num_blocks = cfg.gnn_layers, cfg.num_blocks
        c.precision = PRECISIONS[cfg.precision]
        c.max_len = cfg.max_len
        self._c = c
        blob = np.ascontiguousarray(pack_weights(params, cfg), np.float32)
        want = self.lib.pst_weight_blob_floats(C.byref(c))
        if want == 0:
            raise _lib.PstError(-2, "pst_weight_blob_floats")
        if want != blob.size:
            raise ValueError(f"weight blob has {blob.size} floats, library expects {want}")
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pst_model_create(C.byref(c), blob.ctypes.data, blob.size, device, C.byref(handle)),
                       "pst_model_create")
        self._h = handle
        self._ws = None
        self.launches = 0
        # pst_tokenize replays a CUDA graph for repeated argument sets, which needs a named stream: when the caller is
        # on the default stream the fused call runs on this side stream, ordered with the caller's stream on both sides
        self._side = torch.cuda.Stream(device=self.device)
        self._copy = torch.cuda.Stream(device=self.device)  # H2D / D2H copies of the chunk pipeline (tokenize)
        self._slots = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.pst_model_destroy(self._h)
            self._h = C.c_void_p()
        self._ws = None
        self._slots = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ plumbing
    def _stream(self) -> int:
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def _workspace(self, R: int, B: int):
        need = self.lib.pst_workspace_bytes(self._h, int(R), int(B))
        if need == 0:  # include/pst_abi.h: PST_MAX_EDGES_PER_CALL
            raise ValueError(f"a batch of {R} residues exceeds the per-call limit of the C ABI; use tokenize(), which sends chunks")
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = self.torch.empty(int(need), dtype=self.torch.uint8, device=self.device)
        return self._ws

    def token_offsets(self, offsets: np.ndarray) -> np.ndarray:
        lens = np.diff(np.asarray(offsets, np.int64))
        out = np.zeros(len(lens) + 1, np.int32)
        out[1:] = np.cumsum(lens // self.cfg.downsampling_ratio)
        return out

    def check_lengths(self, offsets: np.ndarray) -> None:
        """Same rejection rule as scripts/inference_runner.py:52-62 (NotImplementedError)."""
        lens = np.diff(np.asarray(offsets, np.int64))
        if (lens > self.cfg.max_len).any():
            raise NotImplementedError(
                f"We currently don't support protein with more than {self.cfg.max_len} residues given: {int(lens.max())}")
        if (lens < self.cfg.num_neighbor).any():
            raise NotImplementedError(
                f"We currently don't support protein with less than {self.cfg.num_neighbor} residues given: {int(lens.min())}")

    # ------------------------------------------------------------------ device-level API
    def featurize_device(self, atoms, atom_mask, offsets_dev, B: int, R: int, want_features: bool = True):
        """atoms f32 [R, A, 3] (A = 4 or 37), atom_mask u8 [R, A] or None, offsets i32 [B+1], all on device."""
        t = self.torch
        K = self.cfg.num_neighbor
        senders = t.empty((R * K,), dtype=t.int32, device=self.device)
        feats = t.empty((R * K, 27), dtype=t.float32, device=self.device) if want_features else None
        ws = self._workspace(R, B)
        _lib.check(self.lib.pst_featurize_knn(self._h, self._stream(), atoms.data_ptr(), _ptr(atom_mask), int(atoms.shape[1]),
                                              offsets_dev.data_ptr(), B, R, senders.data_ptr(), _ptr(feats),
                                              ws.data_ptr(), ws.numel()), "pst_featurize_knn")
        self.launches = self.lib.pst_last_launch_count(self._h)
        return senders, feats

    def encode_graph_device(self, edge_features, senders, offsets_dev, token_offsets_dev, B: int, R: int, T: int):
        t = self.torch
        z = t.empty((T, 8), dtype=t.float32, device=self.device)
        ws = self._workspace(R, B)
        _lib.check(self.lib.pst_encode_graph(self._h, self._stream(), edge_features.data_ptr(), senders.data_ptr(),
                                             offsets_dev.data_ptr(), token_offsets_dev.data_ptr(), B, R, T,
                                             z.data_ptr(), ws.data_ptr(), ws.numel()), "pst_encode_graph")
        self.launches = self.lib.pst_last_launch_count(self._h)
        return z

    def quantize_device(self, z, want_bounded: bool = False):
        t = self.torch
        n = int(z.shape[0])
        tokens = t.empty((n,), dtype=t.int32, device=self.device)
        bounded = t.empty((n, 8), dtype=t.float32, device=self.device) if want_bounded else None
        _lib.check(self.lib.pst_quantize(self._h, self._stream(), z.data_ptr(), n, tokens.data_ptr(), _ptr(bounded)), "pst_quantize")
        return (tokens, bounded) if want_bounded else tokens

    def fsq_pack_device(self, bounded):
        t = self.torch
        n = int(bounded.shape[0])
        tokens = t.empty((n,), dtype=t.int32, device=self.device)
        _lib.check(self.lib.pst_fsq_pack(self._h, self._stream(), bounded.data_ptr(), n, tokens.data_ptr()), "pst_fsq_pack")
        return tokens

    def indexes_to_codes_device(self, tokens):
        t = self.torch
        n = int(tokens.shape[0])
        codes = t.empty((n, 8), dtype=t.float32, device=self.device)
        _lib.check(self.lib.pst_indexes_to_codes(self._h, self._stream(), tokens.data_ptr(), n, codes.data_ptr()), "pst_indexes_to_codes")
        return codes

    def tokenize_device(self, atoms, atom_mask, offsets_dev, token_offsets_dev, B: int, R: int, T: int, out=None):
        """Fused B2 call: atoms -> int32 tokens [T], everything resident on the device."""
        t = self.torch
        tokens = out if out is not None else t.empty((T,), dtype=t.int32, device=self.device)
        ws = self._workspace(R, B)
        cur = t.cuda.current_stream(self.device)
        on_default = cur.cuda_stream == 0
        run = self._side if on_default else cur
        if on_default:
            run.wait_stream(cur)
        rc = self.lib.pst_tokenize(self._h, run.cuda_stream, atoms.data_ptr(), _ptr(atom_mask), int(atoms.shape[1]),
                                   offsets_dev.data_ptr(), token_offsets_dev.data_ptr(), B, R, T,
                                   tokens.data_ptr(), ws.data_ptr(), ws.numel())
        if on_default:
            cur.wait_stream(run)  # later work on the caller's stream (and reuse of these buffers) is ordered after the call
        _lib.check(rc, "pst_tokenize")
        self.launches = self.lib.pst_last_launch_count(self._h)
        return tokens

    def masked_token_code(self) -> int:
        """Token id the reference emits at padded positions: the bounded latent is multiplied by the mask before
        rounding (model/quantize.py:188-209), so every digit is 0 + floor(L/2) (2 730 for 4 096
        codes, 32 036 for 64 000)."""
        code, basis = 0, 1
        for L in self.cfg.levels:
            code += (int(L) // 2) * basis
            basis *= int(L)
        return code

    def encode_graph(self, edge_features: np.ndarray, senders: np.ndarray, offsets: np.ndarray) -> List[np.ndarray]:
        """Boundary B1 on host arrays: ragged `edge_features` f32 [sum n*K, 27], `senders` i32 [sum n*K] (indices local
        to their structure, as in `ProteinGraph.senders`), `offsets` i32 [B+1] -> one uint32 token array per structure.
        Sent in chunks of at most `max_rows_per_call` residues (pst_encode_graph + pst_quantize per chunk)."""
        t = self.torch
        K = self.cfg.num_neighbor
        offsets = np.asarray(offsets, np.int64)
        self.check_lengths(offsets)
        lengths = np.diff(offsets).tolist()
        out: List[np.ndarray] = []
        for a, b in self._chunks(lengths):
            offs = (offsets[a : b + 1] - offsets[a]).astype(np.int32)
            toff = self.token_offsets(offs)
            R, T, B = int(offs[-1]), int(toff[-1]), b - a
            e0, e1 = int(offsets[a]) * K, int(offsets[b]) * K
            f = t.from_numpy(np.ascontiguousarray(edge_features[e0:e1], np.float32)).to(self.device)
            s = t.from_numpy(np.ascontiguousarray(senders[e0:e1], np.int32)).to(self.device)
            z = self.encode_graph_device(f, s, t.from_numpy(offs).to(self.device), t.from_numpy(toff).to(self.device), B, R, T)
            tokens = self.quantize_device(z).cpu().numpy()
            st = self.read_status()
            if st == 0 and not bool(t.isfinite(z).all()):
                st = _lib.PST_ERR_NON_FINITE  # the standalone quantiser has no status word: same rule as the fused call
            if st != 0:
                raise _lib.PstError(st, "pst_encode_graph (device status)")
            out.extend(tokens[toff[i] : toff[i + 1]].astype(np.uint32) for i in range(B))
        return out

    def graph_cache_enable(self, on: bool = True) -> None:
        """CUDA-graph replay of repeated pst_tokenize calls (on by default; see include/pst_abi.h)."""
        _lib.check(self.lib.pst_graph_cache_enable(self._h, int(on)), "pst_graph_cache_enable")

    def graph_cache_stats(self) -> dict:
        """{'replay', 'update', 'instantiate', 'eager'}: how the fused calls of this model were enqueued so far."""
        c = (C.c_int * 4)()
        _lib.check(self.lib.pst_graph_cache_stats(self._h, c), "pst_graph_cache_stats")
        return {"replay": c[0], "update": c[1], "instantiate": c[2], "eager": c[3]}

    def profile_enable(self, on: bool = True) -> None:
        _lib.check(self.lib.pst_profile_enable(self._h, int(on)), "pst_profile_enable")

    def profile_collect(self):
        """-> (ms[8], groups[8]) accumulated since the last collect; kinds: 0 featurise+k-NN, 1 message MLP,
        2 edge-update MLP, 3 node update, 4 input embeddings, 5 fused df=1 resampler + head, 6 FSQ quantiser."""
        ms = (C.c_float * 8)()
        cnt = (C.c_int * 8)()
        _lib.check(self.lib.pst_profile_collect(self._h, ms, cnt), "pst_profile_collect")
        return list(ms), list(cnt)

    def read_status(self) -> int:
        return self.lib.pst_read_status(self._h, self._stream(), self._ws.data_ptr()) if self._ws is not None else 0

    # ------------------------------------------------------------------ host-level API
    def _chunks(self, lengths: Sequence[int]) -> List[Tuple[int, int]]:
        out, start, rows = [], 0, 0
        for i, L in enumerate(lengths):
            if rows + L > self.max_rows_per_call and i > start:
                out.append((start, i))
                start, rows = i, 0
            rows += L
        if start < len(lengths):
            out.append((start, len(lengths)))
        return out

    def tokenize(self, structures: Sequence[np.ndarray], masks: Optional[Sequence[Optional[np.ndarray]]] = None,
                 flat: bool = False):
        """structures: one fp32 array [L, A, 3] per protein (A = 4: N,CA,C,O; A = 37: atom37),
        only valid residues; masks: optional u8/bool [L, A].  Returns one uint32 token
        array [floor(L/df)] per protein (the dtype the reference saves,
        scripts/inference_runner.py:315-321); with `flat=True` the same tokens as ONE int32 array (structures back
        to back) plus the per-structure counts, which is what the multi-GPU gather sends (pst/distributed.py).

        The structures are processed in chunks of at most `max_rows_per_call` residues through a two-slot pipeline:
        while the GPU works on chunk i, the host packs chunk i + 1 into pinned staging buffers and its H2D copy runs on
        a second stream; tokens and the device status word come back the same way."""
        t = self.torch
        lengths = [int(s.shape[0]) for s in structures]
        offs_all = np.zeros(len(lengths) + 1, np.int64)
        offs_all[1:] = np.cumsum(lengths)
        self.check_lengths(offs_all)
        df = self.cfg.downsampling_ratio
        counts = np.asarray(lengths, np.int64) // df
        tok_all = np.zeros(len(lengths) + 1, np.int64)
        tok_all[1:] = np.cumsum(counts)
        out_flat = np.empty(int(tok_all[-1]), np.int32)
        if not lengths:
            return (out_flat, counts) if flat else []
        chunks = self._chunks(lengths)
        A = int(structures[0].shape[1])
        use_mask = masks is not None and any(m is not None for m in masks)
        cap_rows = max(int(offs_all[b] - offs_all[a]) for a, b in chunks)
        cap_b = max(b - a for a, b in chunks)
        slots = self._pipeline_slots(min(2, len(chunks)), cap_rows, cap_b, A, use_mask)
        self._workspace(cap_rows, cap_b)  # sized for the largest chunk up front: no reallocation inside the pipeline
        main = t.cuda.current_stream(self.device)

        def finish(sl):
            a, b, T = sl["job"]
            sl["done"].synchronize()
            st = int(sl["status_pin"][0])
            if st != 0:
                raise _lib.PstError(st, "pst_tokenize (device status)")
            out_flat[tok_all[a] : tok_all[a] + T] = sl["tokens_pin"].numpy()[:T]
            sl["job"] = None

        for ci, (a, b) in enumerate(chunks):
            sl = slots[ci % len(slots)]
            if sl["job"] is not None:
                finish(sl)
            offsets = (offs_all[a : b + 1] - offs_all[a]).astype(np.int32)
            tok_off = (tok_all[a : b + 1] - tok_all[a]).astype(np.int32)
            R, T, B = int(offsets[-1]), int(tok_off[-1]), b - a
            atoms_np = sl["atoms_pin"].numpy()
            if B == 1:
                atoms_np[:R] = structures[a]
            else:
                np.concatenate(structures[a:b], axis=0, out=atoms_np[:R])
            if use_mask:
                mask_np = sl["mask_pin"].numpy()
                for i in range(a, b):
                    m = masks[i]
                    mask_np[offsets[i - a] : offsets[i - a + 1]] = 1 if m is None else np.asarray(m, np.uint8)
            sl["offs_pin"].numpy()[: B + 1] = offsets
            sl["toff_pin"].numpy()[: B + 1] = tok_off
            with t.cuda.stream(self._copy):
                self._copy.wait_event(sl["computed"])  # the slot's previous call has consumed its device inputs
                sl["atoms"][:R].copy_(sl["atoms_pin"][:R], non_blocking=True)
                if use_mask:
                    sl["mask"][:R].copy_(sl["mask_pin"][:R], non_blocking=True)
                sl["offs"][: B + 1].copy_(sl["offs_pin"][: B + 1], non_blocking=True)
                sl["toff"][: B + 1].copy_(sl["toff_pin"][: B + 1], non_blocking=True)
                sl["in_ready"].record(self._copy)
            main.wait_event(sl["in_ready"])
            self.tokenize_device(sl["atoms"][:R], sl["mask"][:R] if use_mask else None, sl["offs"], sl["toff"], B, R, T,
                                 out=sl["tokens"][:T])
            sl["status_pin"].copy_(self._ws[:4].view(t.int32)[:1], non_blocking=True)  # status word: start of the workspace
            sl["computed"].record(main)
            with t.cuda.stream(self._copy):
                self._copy.wait_event(sl["computed"])
                sl["tokens_pin"][:T].copy_(sl["tokens"][:T], non_blocking=True)
                sl["done"].record(self._copy)
            sl["job"] = (a, b, T)
        order = sorted((sl for sl in slots if sl["job"] is not None), key=lambda sl: sl["job"][0])
        for sl in order:
            finish(sl)
        if flat:
            return out_flat, counts
        u = out_flat.view(np.uint32)
        return [u[a:b] for a, b in zip(tok_all[:-1].tolist(), tok_all[1:].tolist())]

    def _pipeline_slots(self, n: int, rows: int, nb: int, A: int, use_mask: bool):
        """Pinned host + device staging buffers of the chunk pipeline (grow-only, kept between calls)."""
        t = self.torch
        key = (A, use_mask)
        st = getattr(self, "_slots", None)
        if st is None or st["key"] != key or st["rows"] < rows or st["nb"] < nb or len(st["slots"]) < n:
            rows = max(rows, st["rows"] if st and st["key"] == key else 0)
            nb = max(nb, st["nb"] if st and st["key"] == key else 0)
            slots = []
            for _ in range(max(n, len(st["slots"]) if st and st["key"] == key else 0)):
                slots.append({
                    "atoms_pin": t.empty((rows, A, 3), dtype=t.float32).pin_memory(),
                    "mask_pin": t.empty((rows, A), dtype=t.uint8).pin_memory() if use_mask else None,
                    "offs_pin": t.empty((nb + 1,), dtype=t.int32).pin_memory(),
                    "toff_pin": t.empty((nb + 1,), dtype=t.int32).pin_memory(),
                    "tokens_pin": t.empty((rows,), dtype=t.int32).pin_memory(),
                    "status_pin": t.zeros((1,), dtype=t.int32).pin_memory(),
                    "atoms": t.empty((rows, A, 3), dtype=t.float32, device=self.device),
                    "mask": t.empty((rows, A), dtype=t.uint8, device=self.device) if use_mask else None,
                    "offs": t.empty((nb + 1,), dtype=t.int32, device=self.device),
                    "toff": t.empty((nb + 1,), dtype=t.int32, device=self.device),
                    "tokens": t.empty((rows,), dtype=t.int32, device=self.device),
                    "in_ready": t.cuda.Event(), "computed": t.cuda.Event(), "done": t.cuda.Event(), "job": None,
                })
            st = {"key": key, "rows": rows, "nb": nb, "slots": slots}
            self._slots = st
        for sl in st["slots"]:
            sl["job"] = None
        return st["slots"][:n]
