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
