"""Mirror of the reference's `DataPipeline` (data_pipeline.py:36-332), the second caller of the featurisation path:
raw structure (PDB file / PDB string / ProteinStructureSample .npy) -> validated sample -> padded `BatchDataVQ3D`
ready for the model, saved as .npy / .npz.

Same method names, arguments, config keys and defaults.  What differs underneath: `preprocess` runs the frames,
centroids, k-NN and 27 edge features on the GPU (`pst_featurize_knn`, csrc/featurize.cu) instead of the NumPy code of
data/preprocessing.py:42-189 + utils/protein_utils.py:325-438, and then pads exactly like data/preprocessing.py:191-283
(zero features and self loops for padded residues, `nodes_mask`, `tokens_mask`, `n_node`, `n_edge`).  The result feeds
`InferenceRunner.prepare_tokenize_fn`'s callable directly (boundary B1) and has the leaves `save_output` writes.

Not reproduced (training-side features of `preprocess_sample`, rejected loudly): random cropping of chains longer than
`crop_index` (:113-135), coordinate noise (`noise_level` > 0), `residue_loc_is_alphac=False`, and
`make_protein_features` (decoder-loss inputs: `features` is an empty dict, nothing on the tokenize path reads it).
"""
from __future__ import annotations

import os
from typing import Any, Dict, NamedTuple, Optional, Union

import numpy as np

from .pdb import StructureSample, structure_from_pdb_bytes_native, structure_from_pdb_file_native, structure_from_sample_file


class ProteinGraph(NamedTuple):
    """The leaves of the reference's `ProteinGraph` (types.py:48-75), host arrays."""

    n_node: np.ndarray
    n_edge: np.ndarray
    nodes_mask: np.ndarray
    nodes_original_coordinates: np.ndarray
    node_features: np.ndarray
    edge_features: np.ndarray
    tokens_mask: np.ndarray
    senders: np.ndarray
    receivers: np.ndarray


class BatchDataVQ3D(NamedTuple):
    """types.py:78-87"""

    graph: ProteinGraph
    features: dict


def filter_out_sample(sample: StructureSample, min_number_valid_residues: int, max_number_residues: int) -> bool:
    """data/preprocessing.py:29-39"""
    known = int(sample.valid_backbone().sum())
    return bool(known < min_number_valid_residues or sample.nb_residues > max_number_residues)


class DataPipeline:
    DEFAULTS = {
        "num_neighbor": 30, "downsampling_ratio": 4, "residue_loc_is_alphac": True, "padding_num_residue": 512,
        "crop_index": 512, "noise_level": 0.0, "min_number_valid_residues": 10, "max_number_residues": 1000,
        "chain_id": None, "save_intermediate": False, "output_format": "npy",
    }

    def __init__(self, config: Optional[Dict[str, Any]] = None, device: int = 0):
        self.config = dict(self.DEFAULTS)
        if config:
            self.config.update(config)
        self.device = device
        self._tok = None

    # ------------------------------------------------------------------ loading (data_pipeline.py:69-154)
    def load_from_pdb_file(self, pdb_file_path: str, chain_id: Optional[str] = None) -> StructureSample:
        if not os.path.exists(pdb_file_path):
            raise FileNotFoundError(f"PDB file not found: {pdb_file_path}")
        return structure_from_pdb_file_native(pdb_file_path, chain_id or self.config["chain_id"])

    def load_from_pdb_string(self, pdb_string: str, chain_id: Optional[str] = None) -> StructureSample:
        return structure_from_pdb_bytes_native(pdb_string.encode(), chain_id or self.config["chain_id"])

    def load_from_npy_file(self, npy_file_path: str) -> StructureSample:
        return structure_from_sample_file(npy_file_path)

    def validate_sample(self, sample: StructureSample) -> bool:
        return not filter_out_sample(sample, self.config["min_number_valid_residues"], self.config["max_number_residues"])

    # ------------------------------------------------------------------ preprocessing (data_pipeline.py:186-214)
    def _featurizer(self):
        if self._tok is None:
            from .config import TokenizerConfig
            from .tokenizer import StructureTokenizer
            from .weights import init_params

            c = self.config
            # the featuriser only reads num_neighbor / max_len; the weights are placeholders and never used
            cfg = TokenizerConfig(seq_max_size=max(int(c["padding_num_residue"]), int(c["num_neighbor"])),
                                  max_out_len=max(int(c["padding_num_residue"]), int(c["num_neighbor"])),
                                  num_neighbor=int(c["num_neighbor"]), downsampling_ratio=1, precision="fp32")
            self._tok = StructureTokenizer(cfg, init_params(cfg, 0, "spread"), device=self.device)
        return self._tok

    def preprocess(self, sample: StructureSample) -> BatchDataVQ3D:
        c = self.config
        K, N, df = int(c["num_neighbor"]), int(c["padding_num_residue"]), int(c["downsampling_ratio"])
        if not c["residue_loc_is_alphac"]:
            raise NotImplementedError("residue_loc_is_alphac=False is not supported (every released data config uses CA)")
        if float(c["noise_level"]) != 0.0:
            raise NotImplementedError("noise_level > 0 is a training-time augmentation; not part of the tokenize path")
        atoms, mask = sample.device_arrays()  # valid residues only (data/preprocessing.py:99-117)
        n = int(atoms.shape[0])
        if n > int(c["crop_index"]):
            raise NotImplementedError(f"{n} valid residues > crop_index {c['crop_index']}: random cropping is a training-time step")
        if n < K:
            raise NotImplementedError(f"We currently don't support protein with less than {K} residues given: {n}")
        if n > N:
            raise NotImplementedError(f"We currently don't support protein with more than {N} residues given: {n}")
        import torch

        tok = self._featurizer()
        dev = tok.device
        offs = torch.tensor([0, n], dtype=torch.int32, device=dev)
        senders, feats = tok.featurize_device(torch.from_numpy(atoms).to(dev), torch.from_numpy(mask).to(dev), offs, 1, n)
        st = tok.read_status()
        if st != 0:
            from . import _lib

            raise _lib.PstError(st, "pst_featurize_knn (device status)")
        senders, feats = senders.cpu().numpy().astype(np.int64), feats.cpu().numpy()
        # padding exactly as data/preprocessing.py:191-283
        edge_features = np.zeros((N * K, 27), np.float32)
        edge_features[: n * K] = feats
        send = np.repeat(np.arange(N, dtype=np.int64), K)
        send[: n * K] = senders
        recv = np.repeat(np.arange(N, dtype=np.int64), K)
        ca = np.zeros((N, 3), np.float32)
        ca[:n] = atoms[:, 1]
        graph = ProteinGraph(
            n_node=np.array([n]), n_edge=np.array([n * K]), nodes_mask=(np.arange(N) < n)[:, None],
            nodes_original_coordinates=ca, node_features=ca, edge_features=edge_features,
            tokens_mask=(np.arange(N // df) < n // df)[:, None], senders=send, receivers=recv)
        return BatchDataVQ3D(graph=graph, features={})

    # ------------------------------------------------------------------ output (data_pipeline.py:216-272)
    def save_output(self, data: BatchDataVQ3D, output_path: str, save_raw_sample: bool = False,
                    raw_sample: Optional[StructureSample] = None) -> None:
        from .pdb import structure_to_sample_file

        os.makedirs(os.path.dirname(output_path) if os.path.dirname(output_path) else ".", exist_ok=True)
        g = data.graph
        if self.config["output_format"] == "npz":
            np.savez_compressed(output_path, graph_n_node=g.n_node, graph_n_edge=g.n_edge, graph_nodes_mask=g.nodes_mask,
                                graph_nodes_original_coordinates=g.nodes_original_coordinates,
                                graph_node_features=g.node_features, graph_edge_features=g.edge_features,
                                graph_tokens_mask=g.tokens_mask, graph_senders=g.senders, graph_receivers=g.receivers)
        else:
            np.save(output_path, {"graph": g._asdict(), "features": dict(data.features)})
        if save_raw_sample and raw_sample is not None:
            structure_to_sample_file(raw_sample, output_path.replace(".npy", "_raw.npy").replace(".npz", "_raw.npy"))

    @staticmethod
    def load_output(path: str) -> BatchDataVQ3D:
        """Reads back what `save_output` wrote (either format) as a `BatchDataVQ3D`."""
        if path.endswith(".npz"):
            with np.load(path) as f:
                g = {k[len("graph_"):]: f[k] for k in f.files if k.startswith("graph_")}
        else:
            from .pdb import _load_sample_dict  # the restricted unpickler: plain dicts of NumPy arrays only

            g = _load_sample_dict(path)["graph"]
        return BatchDataVQ3D(graph=ProteinGraph(**g), features={})

    def process_single(self, input_source: Union[str, StructureSample], output_path: str, input_type: str = "auto") -> BatchDataVQ3D:
        if isinstance(input_source, StructureSample):
            sample = input_source
        elif input_type == "auto":
            if os.path.exists(str(input_source)):
                sample = self.load_from_npy_file(str(input_source)) if str(input_source).endswith(".npy") \
                    else self.load_from_pdb_file(str(input_source))
            else:
                sample = self.load_from_pdb_string(str(input_source))
        elif input_type == "pdb_file":
            sample = self.load_from_pdb_file(str(input_source))
        elif input_type == "pdb_string":
            sample = self.load_from_pdb_string(str(input_source))
        elif input_type == "npy_file":
            sample = self.load_from_npy_file(str(input_source))
        else:
            raise ValueError(f"Unknown input_type: {input_type}")
        if not self.validate_sample(sample):
            raise ValueError("Sample failed validation - does not meet filtering criteria")
        data = self.preprocess(sample)
        self.save_output(data, output_path, save_raw_sample=self.config.get("save_intermediate", False), raw_sample=sample)
        return data

    def get_sample_info(self, sample: StructureSample) -> Dict[str, Any]:
        missing = ~sample.valid_backbone()
        return {"chain_id": None, "total_residues": sample.nb_residues, "valid_residues": int((~missing).sum()),
                "missing_residues": int(missing.sum()), "resolution": 0.0, "pdb_cluster_size": 1,
                "sequence": "".join("ACDEFGHIKLMNPQRSTVWYX"[int(a)] for a in sample.aatype)}
