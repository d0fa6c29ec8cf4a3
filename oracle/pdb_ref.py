"""Oracle (TEST INFRASTRUCTURE): PDB text -> atom37 arrays, restating
structure_tokenizer/data/protein_structure_sample.py:166-248 on top of the
BioPython ``PDBParser`` semantics it relies on (biopython==1.80, not vendored
in the reference): fixed columns, ATOM/HETATM records of the single model,
chains in order of first appearance, residues keyed by (hetero flag, resseq,
icode), first-seen atom wins on a duplicate name unless a later alternate
location has a strictly higher occupancy, coordinates parsed to float32.

Tables follow structure_tokenizer/data/residue_constants.py:539-577 (atom37
order) and :733-737 (which atom37 slots exist per residue type; UNK = N, CA,
C, CB).  Parity: unpinned against BioPython itself (absent from the image);
the bundled CASP14 files have no HETATM / altloc / insertion codes, so only the
plain path is exercised by the fixtures.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

ATOM_TYPES: List[str] = (
    "N CA C CB O CG CG1 CG2 OG OG1 SG CD CD1 CD2 ND1 ND2 OD1 OD2 SD CE CE1 CE2 CE3 "
    "NE NE1 NE2 OE1 OE2 CH2 NH1 NH2 OH CZ CZ2 CZ3 NZ OXT"
).split()
ATOM_ORDER: Dict[str, int] = {a: i for i, a in enumerate(ATOM_TYPES)}

# side-chain heavy atoms per residue type (backbone N, CA, C, O are implied)
_SIDE: Dict[str, str] = {
    "ALA": "CB", "ARG": "CB CG CD NE CZ NH1 NH2", "ASN": "CB CG OD1 ND2",
    "ASP": "CB CG OD1 OD2", "CYS": "CB SG", "GLN": "CB CG CD OE1 NE2",
    "GLU": "CB CG CD OE1 OE2", "GLY": "", "HIS": "CB CG ND1 CD2 CE1 NE2",
    "ILE": "CB CG1 CG2 CD1", "LEU": "CB CG CD1 CD2", "LYS": "CB CG CD CE NZ",
    "MET": "CB CG SD CE", "PHE": "CB CG CD1 CD2 CE1 CE2 CZ", "PRO": "CB CG CD",
    "SER": "CB OG", "THR": "CB OG1 CG2",
    "TRP": "CB CG CD1 CD2 NE1 CE2 CE3 CZ2 CZ3 CH2",
    "TYR": "CB CG CD1 CD2 CE1 CE2 CZ OH", "VAL": "CB CG1 CG2",
}
RESTYPES_3 = ["ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIS", "ILE",
              "LEU", "LYS", "MET", "PHE", "PRO", "SER", "THR", "TRP", "TYR", "VAL"]
RESTYPE_ORDER = {r: i for i, r in enumerate(RESTYPES_3)}


def atom37_exists(resname: str) -> np.ndarray:
    out = np.zeros(37, bool)
    if resname in _SIDE:
        for a in ["N", "CA", "C", "O"] + _SIDE[resname].split():
            out[ATOM_ORDER[a]] = True
    else:  # UNK: N, CA, C, CB (residue_constants.py:737)
        out[:4] = True
    return out


def parse_pdb(text: str) -> Dict[str, np.ndarray]:
    chains: Dict[str, Dict[tuple, dict]] = {}
    n_models = 0
    saw_atom_outside_model = False
    in_model = False
    for line in text.splitlines():
        rec = line[:6]
        if rec.startswith("MODEL"):
            n_models += 1
            in_model = True
            continue
        if rec.startswith("ENDMDL"):
            in_model = False
            continue
        if rec not in ("ATOM  ", "HETATM"):
            continue
        if not in_model:
            saw_atom_outside_model = True
        name = line[12:16].strip()
        altloc = line[16]
        resname = line[17:20].strip()
        chain = line[21]
        resseq = int(line[22:26])
        icode = line[26]
        hetero = " "
        if rec == "HETATM":
            hetero = "W" if resname in ("HOH", "WAT") else "H_" + resname
        xyz = np.array([float(line[30:38]), float(line[38:46]), float(line[46:54])], np.float32)
        try:
            occ = float(line[54:60])
        except ValueError:
            occ = 1.0
        res = chains.setdefault(chain, {}).setdefault(
            (hetero, resseq, icode), {"resname": resname, "atoms": {}, "chain": chain}
        )
        prev = res["atoms"].get(name)
        if prev is None:
            res["atoms"][name] = (xyz, occ, altloc)
        elif altloc != " " and prev[2] != " " and occ > prev[1]:
            res["atoms"][name] = (xyz, occ, altloc)
    total_models = n_models + (1 if (saw_atom_outside_model and n_models == 0) else 0)
    if n_models > 1 or total_models != 1:
        raise ValueError(f"Only single model PDBs are supported. Found {max(n_models, total_models)} models.")

    pos_l, gt_l, ex_l, aa_l = [], [], [], []
    for chain_id, residues in chains.items():
        for (hetero, resseq, icode), res in residues.items():
            if icode != " ":
                raise ValueError(
                    f"PDB contains an insertion code at chain {chain_id} and residue index {resseq}. "
                    "These are not supported."
                )
            rn = res["resname"] if res["resname"] in _SIDE else "UNK"
            pos = np.zeros((37, 3), np.float64)
            gt = np.zeros(37, bool)
            for name, (xyz, _, _) in res["atoms"].items():
                slot = ATOM_ORDER.get(name)
                if slot is None:
                    continue
                pos[slot] = xyz
                gt[slot] = True
            if not gt.any():
                continue
            pos_l.append(pos)
            gt_l.append(gt)
            ex_l.append(atom37_exists(rn))
            aa_l.append(RESTYPE_ORDER.get(rn, 20))
    n = len(pos_l)
    return {
        "nb_residues": n,
        "atom37_positions": np.asarray(pos_l, np.float64).reshape(n, 37, 3),
        "atom37_gt_exists": np.asarray(gt_l, bool).reshape(n, 37),
        "atom37_atom_exists": np.asarray(ex_l, bool).reshape(n, 37),
        "aatype": np.asarray(aa_l, np.int32),
    }


def _cif_tokens(text: str):
    """STAR tokens of an mmCIF text: (value, quoted).  Whitespace-separated values, '...' / "..." values closed by a
    quote followed by whitespace, ;...; text fields starting in column 1, # comments."""
    i, n = 0, len(text)
    while i < n:
        c = text[i]
        if c in " \t\r\n":
            i += 1
            continue
        if c == "#":
            while i < n and text[i] != "\n":
                i += 1
            continue
        if c == ";" and (i == 0 or text[i - 1] == "\n"):
            j = text.find("\n;", i)
            if j < 0:
                yield text[i + 1:], True
                return
            yield text[i + 1:j], True
            i = j + 2
            continue
        if c in "'\"":
            j = i + 1
            while j < n and text[j] != "\n" and not (text[j] == c and (j + 1 >= n or text[j + 1] in " \t\r\n")):
                j += 1
            yield text[i + 1:j], True
            i = j + 1
            continue
        j = i
        while j < n and text[j] not in " \t\r\n":
            j += 1
        yield text[i:j], False
        i = j


def parse_mmcif(text: str, chain_id: str = None) -> Dict[str, np.ndarray]:
    """mmCIF `_atom_site` loop -> the same arrays as parse_pdb, with BioPython's MMCIFParser conventions (chain =
    auth_asym_id, residue number = auth_seq_id, atom name = label_atom_id, '.' / '?' = no altloc / insertion code,
    hetero flag 'W' / 'H' from group_PDB, one pdbx_PDB_model_num).  The reference has no mmCIF reader: this is the
    independent restatement the C++ parser (csrc/pdb_parse.cc) is checked against."""
    toks = list(_cif_tokens(text))
    i = 0
    tags = None
    while i < len(toks):
        v, q = toks[i]
        if not q and v == "loop_":
            j = i + 1
            cur = []
            while j < len(toks) and not toks[j][1] and toks[j][0].startswith("_"):
                cur.append(toks[j][0])
                j += 1
            i = j
            if cur and cur[0].startswith("_atom_site."):
                tags = [t[len("_atom_site."):] for t in cur]
                break
            continue
        i += 1
    if tags is None:
        raise ValueError("Only single model PDBs are supported. Found 0 models.")

    def col(*names):
        for nm in names:
            if nm in tags:
                return tags.index(nm)
        return None

    c_group, c_atom, c_alt = col("group_PDB"), col("label_atom_id", "auth_atom_id"), col("label_alt_id")
    c_comp, c_chain = col("label_comp_id", "auth_comp_id"), col("auth_asym_id", "label_asym_id")
    c_seq, c_seq_label, c_ins = col("auth_seq_id", "label_seq_id"), col("label_seq_id"), col("pdbx_PDB_ins_code")
    c_x, c_y, c_z, c_occ, c_model = col("Cartn_x"), col("Cartn_y"), col("Cartn_z"), col("occupancy"), col("pdbx_PDB_model_num")
    nc = len(tags)
    chains: Dict[str, Dict[tuple, dict]] = {}
    models = set()
    null = lambda t: (not t[1]) and t[0] in (".", "?")
    while i < len(toks):
        v, q = toks[i]
        if not q and (v.startswith("_") or v == "loop_" or v.startswith("data_") or v.startswith("save_")):
            break
        row = toks[i:i + nc]
        if len(row) < nc:
            raise ValueError("malformed atom_site row")
        i += nc
        if c_model is not None:
            models.add(int(row[c_model][0]))
            if len(models) > 1:
                raise ValueError(f"Only single model PDBs are supported. Found {len(models)} models.")
        het = c_group is not None and row[c_group][0] == "HETATM"
        resname = row[c_comp][0]
        hetero = " " if not het else ("W" if resname in ("HOH", "WAT") else "H")
        seq = row[c_seq]
        if null(seq) and c_seq_label is not None:
            seq = row[c_seq_label]
        resseq = int(seq[0])
        icode = " " if (c_ins is None or null(row[c_ins]) or not row[c_ins][0]) else row[c_ins][0][0]
        altloc = " " if (c_alt is None or null(row[c_alt]) or not row[c_alt][0]) else row[c_alt][0][0]
        name = row[c_atom][0]
        xyz = np.array([float(row[c_x][0]), float(row[c_y][0]), float(row[c_z][0])], np.float32)
        try:
            occ = float(row[c_occ][0]) if c_occ is not None else 1.0
        except ValueError:
            occ = 1.0
        chain = row[c_chain][0]
        res = chains.setdefault(chain, {}).setdefault((hetero, resseq, icode), {"resname": resname, "atoms": {}})
        prev = res["atoms"].get(name)
        if prev is None or (altloc != " " and prev[2] != " " and occ > prev[1]):
            res["atoms"][name] = (xyz, occ, altloc)
    if not chains:
        raise ValueError("Only single model PDBs are supported. Found 0 models.")
    pos_l, gt_l, ex_l, aa_l = [], [], [], []
    for cid, residues in chains.items():
        if chain_id is not None and cid != chain_id:
            continue
        for (hetero, resseq, icode), res in residues.items():
            if icode != " ":
                raise ValueError(f"PDB contains an insertion code at chain {cid} and residue index {resseq}. These are not supported.")
            rn = res["resname"] if res["resname"] in _SIDE else "UNK"
            pos = np.zeros((37, 3), np.float64)
            gt = np.zeros(37, bool)
            for name, (xyz, _, _) in res["atoms"].items():
                slot = ATOM_ORDER.get(name)
                if slot is None:
                    continue
                pos[slot] = xyz
                gt[slot] = True
            if not gt.any():
                continue
            pos_l.append(pos)
            gt_l.append(gt)
            ex_l.append(atom37_exists(rn))
            aa_l.append(RESTYPE_ORDER.get(rn, 20))
    n = len(pos_l)
    return {
        "nb_residues": n,
        "atom37_positions": np.asarray(pos_l, np.float64).reshape(n, 37, 3),
        "atom37_gt_exists": np.asarray(gt_l, bool).reshape(n, 37),
        "atom37_atom_exists": np.asarray(ex_l, bool).reshape(n, 37),
        "aatype": np.asarray(aa_l, np.int32),
    }
