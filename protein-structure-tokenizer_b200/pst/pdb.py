"""PDB text -> atom37 arrays for the tokenizer (host side of boundary B2).

Mirrors `protein_structure_from_pdb_string` of the reference
(structure_tokenizer/data/protein_structure_sample.py:166-248) and the BioPython PDBParser
behaviour it depends on: ATOM/HETATM records of a single model, every chain concatenated in
file order, atoms kept iff their name is one of the 37 atom37 types
(data/residue_constants.py:539-577), unknown residue names mapped to UNK (N, CA, C, CB expected),
ValueError on insertion codes and on multi-model files, coordinates stored as float32.
Also the residue filter of data/preprocessing.py:72,99-117 (`valid_backbone`).
"""
from __future__ import annotations

import pickle
import threading
from typing import Dict, List, NamedTuple, Optional, Tuple

import numpy as np

ATOM_TYPES: Tuple[str, ...] = (
    "N", "CA", "C", "CB", "O", "CG", "CG1", "CG2", "OG", "OG1", "SG", "CD", "CD1", "CD2", "ND1", "ND2", "OD1", "OD2",
    "SD", "CE", "CE1", "CE2", "CE3", "NE", "NE1", "NE2", "OE1", "OE2", "CH2", "NH1", "NH2", "OH", "CZ", "CZ2", "CZ3",
    "NZ", "OXT",
)
ATOM_ORDER: Dict[str, int] = {n: i for i, n in enumerate(ATOM_TYPES)}
N_INDEX, CA_INDEX, C_INDEX, CB_INDEX, O_INDEX = 0, 1, 2, 3, 4

RESIDUE_ATOMS: Dict[str, Tuple[str, ...]] = {
    "ALA": ("C", "CA", "CB", "N", "O"),
    "ARG": ("C", "CA", "CB", "CG", "CD", "CZ", "N", "NE", "O", "NH1", "NH2"),
    "ASP": ("C", "CA", "CB", "CG", "N", "O", "OD1", "OD2"),
    "ASN": ("C", "CA", "CB", "CG", "N", "ND2", "O", "OD1"),
    "CYS": ("C", "CA", "CB", "N", "O", "SG"),
    "GLU": ("C", "CA", "CB", "CG", "CD", "N", "O", "OE1", "OE2"),
    "GLN": ("C", "CA", "CB", "CG", "CD", "N", "NE2", "O", "OE1"),
    "GLY": ("C", "CA", "N", "O"),
    "HIS": ("C", "CA", "CB", "CG", "CD2", "CE1", "N", "ND1", "NE2", "O"),
    "ILE": ("C", "CA", "CB", "CG1", "CG2", "CD1", "N", "O"),
    "LEU": ("C", "CA", "CB", "CG", "CD1", "CD2", "N", "O"),
    "LYS": ("C", "CA", "CB", "CG", "CD", "CE", "N", "NZ", "O"),
    "MET": ("C", "CA", "CB", "CG", "CE", "N", "O", "SD"),
    "PHE": ("C", "CA", "CB", "CG", "CD1", "CD2", "CE1", "CE2", "CZ", "N", "O"),
    "PRO": ("C", "CA", "CB", "CG", "CD", "N", "O"),
    "SER": ("C", "CA", "CB", "N", "O", "OG"),
    "THR": ("C", "CA", "CB", "CG2", "N", "O", "OG1"),
    "TRP": ("C", "CA", "CB", "CG", "CD1", "CD2", "CE2", "CE3", "CZ2", "CZ3", "CH2", "N", "NE1", "O"),
    "TYR": ("C", "CA", "CB", "CG", "CD1", "CD2", "CE1", "CE2", "CZ", "N", "O", "OH"),
    "VAL": ("C", "CA", "CB", "CG1", "CG2", "N", "O"),
}
RESTYPE_ORDER = {r: i for i, r in enumerate(
    ("ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIS", "ILE", "LEU", "LYS", "MET", "PHE", "PRO", "SER",
     "THR", "TRP", "TYR", "VAL"))}
_EXISTS = {r: np.isin(np.array(ATOM_TYPES), atoms) for r, atoms in RESIDUE_ATOMS.items()}
_EXISTS["UNK"] = np.array([True] * 4 + [False] * 33)


class StructureSample(NamedTuple):
    nb_residues: int
    aatype: np.ndarray               # int32 [n] (20 = unknown)
    atom37_positions: np.ndarray     # float32 [n, 37, 3]
    atom37_gt_exists: np.ndarray     # bool [n, 37]
    atom37_atom_exists: np.ndarray   # bool [n, 37]

    def valid_backbone(self) -> np.ndarray:
        g = self.atom37_gt_exists
        return g[:, CA_INDEX] & g[:, N_INDEX] & g[:, C_INDEX] & g[:, O_INDEX]

    def device_arrays(self) -> Tuple[np.ndarray, np.ndarray]:
        """(atoms f32 [n_valid, 37, 3], mask u8 [n_valid, 37]) -- what pst_featurize_knn / pst_tokenize take."""
        keep = self.valid_backbone()
        mask = (self.atom37_gt_exists & self.atom37_atom_exists)[keep]
        return np.ascontiguousarray(self.atom37_positions[keep], np.float32), np.ascontiguousarray(mask, np.uint8)


def structure_from_pdb_string(pdb_str: str, chain_id: Optional[str] = None) -> StructureSample:
    """`chain_id`: only that chain is parsed (protein_structure_sample.py:166-168,201-203); None = all chains."""
    residues: Dict[Tuple[str, str, int, str], dict] = {}
    chain_order: List[str] = []
    models = 0
    in_model = False
    loose_atoms = False
    for line in pdb_str.splitlines():
        tag = line[:6]
        if tag.startswith("MODEL"):
            models += 1
            in_model = True
            continue
        if tag.startswith("ENDMDL"):
            in_model = False
            continue
        if tag != "ATOM  " and tag != "HETATM":
            continue
        if not in_model:
            loose_atoms = True
        resname = line[17:20].strip()
        chain = line[21]
        het = " " if tag == "ATOM  " else ("W" if resname in ("HOH", "WAT") else "H_" + resname)
        key = (chain, het, int(line[22:26]), line[26])
        res = residues.get(key)
        if res is None:
            if chain not in chain_order:
                chain_order.append(chain)
            res = residues[key] = {"resname": resname, "atoms": {}}
        name = line[12:16].strip()
        altloc = line[16]
        try:
            occ = float(line[54:60])
        except ValueError:
            occ = 1.0
        prev = res["atoms"].get(name)
        if prev is None or (altloc != " " and prev[2] != " " and occ > prev[1]):
            xyz = (np.float32(line[30:38]), np.float32(line[38:46]), np.float32(line[46:54]))
            res["atoms"][name] = (xyz, occ, altloc)
    n_models = models if models > 0 else (1 if loose_atoms else 0)
    if n_models != 1:
        raise ValueError(f"Only single model PDBs are supported. Found {n_models} models.")

    pos_l, gt_l, ex_l, aa_l = [], [], [], []
    for chain in chain_order:  # BioPython groups residues by chain, chains in order of first appearance
        for (c, het, resseq, icode), res in residues.items():
            if c != chain:
                continue
            if chain_id is not None and c != chain_id:
                continue
            if icode != " ":
                raise ValueError(f"PDB contains an insertion code at chain {chain} and residue index {resseq}. These are not supported.")
            rn = res["resname"] if res["resname"] in RESIDUE_ATOMS else "UNK"
            pos = np.zeros((37, 3), np.float32)
            gt = np.zeros(37, bool)
            for name, (xyz, _, _) in res["atoms"].items():
                slot = ATOM_ORDER.get(name)
                if slot is not None:
                    pos[slot] = xyz
                    gt[slot] = True
            if not gt.any():
                continue
            pos_l.append(pos)
            gt_l.append(gt)
            ex_l.append(_EXISTS[rn])
            aa_l.append(RESTYPE_ORDER.get(rn, 20))
    n = len(pos_l)
    return StructureSample(
        nb_residues=n,
        aatype=np.asarray(aa_l, np.int32),
        atom37_positions=np.asarray(pos_l, np.float32).reshape(n, 37, 3),
        atom37_gt_exists=np.asarray(gt_l, bool).reshape(n, 37),
        atom37_atom_exists=np.asarray(ex_l, bool).reshape(n, 37),
    )


def structure_from_pdb_file(path: str, chain_id: Optional[str] = None) -> StructureSample:
    with open(path, "r") as fh:
        return structure_from_pdb_string(fh.read(), chain_id)


_scratch = threading.local()


def looks_like_mmcif(data: bytes) -> bool:
    """True when the first token of the text (after blank lines and # comments) is a `data_` block header: the test the
    C++ parser applies (csrc/pdb_parse.cc `looks_like_mmcif`)."""
    for line in data[:4096].splitlines():
        t = line.strip()
        if not t or t.startswith(b"#"):
            continue
        return t.startswith(b"data_")
    return False


def structure_from_pdb_bytes_native(data: bytes, chain_id: Optional[str] = None) -> StructureSample:
    """The same result through the C++ parser of the C ABI (`pst_parse_pdb`, csrc/pdb_parse.cc): ~100x faster than
    the pure-Python loop above, which is kept as an independent restatement for the tests.  Raises ValueError where
    the reference does (protein_structure_sample.py:187-190, :205-209).  `chain_id`: only that chain (`pst_parse_pdb_chain`)."""
    import ctypes as C

    from . import _lib

    lib = _lib.load()
    n = C.c_int32(0)
    is_cif = looks_like_mmcif(data)  # mmCIF text (pst_parse_mmcif): chain ids may be longer than one character there
    if chain_id is not None and len(chain_id) != 1 and not is_cif:
        raise ValueError(f"chain_id must be one character, got {chain_id!r}")
    chain = C.c_char(chain_id.encode()) if (chain_id is not None and not is_cif) else C.c_char(b"\0")
    cif_chain = chain_id.encode() if (is_cif and chain_id) else None
    # Per-thread scratch arrays, grown on demand and reused from file to file (the runner parses files side by side:
    # fresh multi-hundred-KB arrays per file are mmap'ed and page-faulted every time, which serialises the threads in the
    # kernel), no scan of the text in Python (the GIL is only released inside the C call): capacity guess = one
    # residue per four 81-byte records, the exact count on the rare retry.
    cap = len(data) // 324 + 16
    while True:
        sc = getattr(_scratch, "arrays", None)
        if sc is None or sc[0].shape[0] < cap:
            grow = max(cap, 1024)
            sc = _scratch.arrays = (np.empty((grow, 37, 3), np.float32), np.empty((grow, 37), np.uint8),
                                    np.empty((grow, 37), np.uint8), np.empty((grow,), np.int32))
        pos, gt, ex, aa = sc
        if is_cif:
            rc = lib.pst_parse_mmcif(data, len(data), cif_chain, pos.shape[0], pos.ctypes.data, gt.ctypes.data, ex.ctypes.data,
                                     aa.ctypes.data, C.byref(n))
        else:
            rc = lib.pst_parse_pdb_chain(data, len(data), chain, pos.shape[0], pos.ctypes.data, gt.ctypes.data, ex.ctypes.data,
                                         aa.ctypes.data, C.byref(n))
        if rc == _lib.PST_ERR_WORKSPACE_TOO_SMALL:
            cap = int(n.value)
            continue
        break
    if rc in (_lib.PST_ERR_PDB_MODEL_COUNT, _lib.PST_ERR_PDB_INSERTION_CODE, _lib.PST_ERR_PDB_MALFORMED):
        raise ValueError(lib.pst_status_string(rc).decode())
    _lib.check(rc, "pst_parse_pdb")
    k = int(n.value)
    return StructureSample(nb_residues=k, aatype=aa[:k].copy(), atom37_positions=pos[:k].copy(),
                           atom37_gt_exists=gt[:k].astype(np.bool_), atom37_atom_exists=ex[:k].astype(np.bool_))


def structure_from_pdb_file_native(path: str, chain_id: Optional[str] = None) -> StructureSample:
    with open(path, "rb") as fh:
        return structure_from_pdb_bytes_native(fh.read(), chain_id)


def _parse_batch_native(n_files: int, guess_rows: int, call, names):
    """Shared driver of the two batch entry points: capacity retry, per-file status -> sample or exception."""
    from . import _lib

    lib = _lib.load()
    offs = np.zeros(n_files + 1, np.int32)
    status = np.zeros(n_files, np.int32)
    cap = guess_rows
    while True:
        pos = np.empty((cap, 37, 3), np.float32)
        gt = np.empty((cap, 37), np.uint8)
        ex = np.empty((cap, 37), np.uint8)
        aa = np.empty((cap,), np.int32)
        rc = call(lib, cap, pos, gt, ex, aa, offs, status)
        if rc == _lib.PST_ERR_WORKSPACE_TOO_SMALL:
            cap = int(offs[n_files])
            continue
        break
    _lib.check(rc, "pst_parse_pdb_batch")
    gtb, exb = gt.view(np.bool_), ex.view(np.bool_)
    out = []
    for i in range(n_files):
        st = int(status[i])
        if st in (_lib.PST_ERR_PDB_MODEL_COUNT, _lib.PST_ERR_PDB_INSERTION_CODE, _lib.PST_ERR_PDB_MALFORMED):
            out.append(ValueError(lib.pst_status_string(st).decode()))
            continue
        if st == _lib.PST_ERR_FILE_NOT_FOUND:
            out.append(FileNotFoundError(f"{names[i]} does not exist or cannot be read"))
            continue
        _lib.check(st, "pst_parse_pdb_batch")
        a, b = int(offs[i]), int(offs[i + 1])
        out.append(StructureSample(nb_residues=b - a, aatype=aa[a:b], atom37_positions=pos[a:b],
                                   atom37_gt_exists=gtb[a:b], atom37_atom_exists=exb[a:b]))
    return out


def structures_from_pdb_bytes_batch_native(datas, n_threads: int = 0):
    """Many PDB texts through ONE C call (`pst_parse_pdb_batch`: host threads inside the library, no GIL, no Python
    work per file while they run).  Returns a list with a StructureSample per text, or the ValueError the reference
    would raise for that file (not raised here: the caller decides which file's error comes first)."""
    import ctypes as C

    nf = len(datas)
    if nf == 0:
        return []
    texts = (C.c_char_p * nf)(*datas)
    sizes = (C.c_size_t * nf)(*[len(d) for d in datas])

    def call(lib, cap, pos, gt, ex, aa, offs, status):
        return lib.pst_parse_pdb_batch(texts, sizes, nf, int(n_threads), cap, pos.ctypes.data, gt.ctypes.data, ex.ctypes.data,
                                       aa.ctypes.data, offs.ctypes.data, status.ctypes.data)

    return _parse_batch_native(nf, sum(len(d) for d in datas) // 324 + 16 * nf, call, [f"text {i}" for i in range(nf)])


def structures_from_pdb_files_native(paths, n_threads: int = 0):
    """The same over files (`pst_parse_pdb_files`): the library's worker threads read the files too, so nothing per file
    happens in Python until the arrays are sliced.  A missing file yields a FileNotFoundError in its slot."""
    import ctypes as C
    import os

    paths = [os.fspath(p) for p in paths]
    nf = len(paths)
    if nf == 0:
        return []
    arr = (C.c_char_p * nf)(*[os.fsencode(p) for p in paths])
    guess = 0
    for p in paths:
        try:
            guess += os.path.getsize(p) // 324 + 16
        except OSError:
            guess += 16

    def call(lib, cap, pos, gt, ex, aa, offs, status):
        return lib.pst_parse_pdb_files(arr, nf, int(n_threads), cap, pos.ctypes.data, gt.ctypes.data, ex.ctypes.data,
                                       aa.ctypes.data, offs.ctypes.data, status.ctypes.data)

    return _parse_batch_native(nf, guess, call, paths)


class _SampleUnpickler(pickle.Unpickler):
    """Unpickler for `ProteinStructureSample.to_file` payloads: a dict of str / int / float / NumPy arrays.  Only the
    NumPy reconstruction helpers those need may be resolved; any other global (what a hostile pickle would use to run
    code) is refused, so a stray .npy in --pdb_dir cannot execute anything (np.load(allow_pickle=True) would)."""

    _ALLOWED = {
        ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
        ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
        ("numpy.core.numeric", "_frombuffer"), ("numpy._core.numeric", "_frombuffer"),
        ("numpy", "ndarray"), ("numpy", "dtype"),
    }

    def find_class(self, module, name):
        if (module, name) in self._ALLOWED:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"refusing to resolve {module}.{name} while reading a structure sample file")


def _load_sample_dict(path: str) -> dict:
    """Reads the 0-d object .npy that `np.save(path, dict)` writes without np.load's unrestricted unpickling."""
    with open(path, "rb") as fh:
        version = np.lib.format.read_magic(fh)
        if version == (1, 0):
            shape, _, dtype = np.lib.format.read_array_header_1_0(fh)
        elif version in ((2, 0), (3, 0)):
            shape, _, dtype = np.lib.format.read_array_header_2_0(fh)
        else:
            raise ValueError(f"{path}: unsupported .npy version {version}")
        if not dtype.hasobject or shape != ():
            raise ValueError(f"{path}: not a ProteinStructureSample file (expected a 0-d object array holding a dict)")
        obj = _SampleUnpickler(fh).load()
    if isinstance(obj, np.ndarray) and obj.shape == ():
        obj = obj[()]
    if not isinstance(obj, dict):
        raise ValueError(f"{path}: not a ProteinStructureSample file (payload is {type(obj).__name__}, expected dict)")
    return obj



def structure_from_sample_file(path: str) -> StructureSample:
    """The reference's preprocessed sample format: `ProteinStructureSample.to_file` / `.from_file`
    (structure_tokenizer/data/protein_structure_sample.py:46-62), an .npy holding the pickled dict of the NamedTuple
    (chain_id, nb_residues, aatype one-hot bool [n, 21], atom37_positions f32 [n, 37, 3], atom37_gt_exists /
    atom37_atom_exists bool [n, 37], resolution, pdb_cluster_size).  444 bytes per residue instead of ~700 bytes of PDB
    text and no parsing: the ingest format for large runs (data/test_data/cif_raw_data.npy is the bundled example)."""
    import os

    if not os.path.isfile(path):
        raise FileNotFoundError(f"{path} does not exist")
    d = _load_sample_dict(path)
    n = int(d["nb_residues"])
    aatype = np.asarray(d["aatype"])
    if aatype.ndim == 2:  # one-hot over the 20 residue types + unknown
        aatype = aatype.argmax(axis=-1)
    pos = np.ascontiguousarray(d["atom37_positions"], np.float32)
    gt = np.ascontiguousarray(d["atom37_gt_exists"], np.bool_)
    ex = np.ascontiguousarray(d["atom37_atom_exists"], np.bool_)
    if pos.shape != (n, 37, 3) or gt.shape != (n, 37) or ex.shape != (n, 37) or aatype.shape != (n,):
        raise ValueError(f"{path}: arrays do not match nb_residues = {n}")
    return StructureSample(nb_residues=n, aatype=aatype.astype(np.int32), atom37_positions=pos, atom37_gt_exists=gt,
                           atom37_atom_exists=ex)


def structure_to_sample_file(sample: StructureSample, path: str, chain_id: str = "", resolution: float = 0.0,
                             pdb_cluster_size: int = 1) -> None:
    """Writes what `ProteinStructureSample.from_file` of the reference reads back (protein_structure_sample.py:46-62)."""
    onehot = np.zeros((sample.nb_residues, 21), np.bool_)
    onehot[np.arange(sample.nb_residues), np.asarray(sample.aatype, np.int64)] = True
    np.save(path, {"chain_id": chain_id, "nb_residues": int(sample.nb_residues), "aatype": onehot,
                   "atom37_positions": np.asarray(sample.atom37_positions, np.float32),
                   "atom37_gt_exists": np.asarray(sample.atom37_gt_exists, np.bool_),
                   "atom37_atom_exists": np.asarray(sample.atom37_atom_exists, np.bool_),
                   "resolution": float(resolution), "pdb_cluster_size": int(pdb_cluster_size)})
