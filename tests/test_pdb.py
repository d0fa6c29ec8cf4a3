"""PDB parser of the product against the oracle's independent restatement (CPU)."""
import numpy as np
import pytest

from oracle import pdb_ref
from pst import pdb as ppdb


def _atom(serial, name, resname, chain, resseq, xyz, rec="ATOM", altloc=" ", icode=" ", occ=1.0, element="C"):
    name_f = name if len(name) == 4 else " " + name.ljust(3)
    return (f"{rec:<6}{serial:>5} {name_f}{altloc}{resname:>3} {chain}{resseq:>4}{icode}   "
            f"{xyz[0]:8.3f}{xyz[1]:8.3f}{xyz[2]:8.3f}{occ:6.2f}{20.0:6.2f}          {element:>2}")


def _pdb_from_backbone(bb, resname="GLY", chain="A"):
    lines, s = [], 1
    for i, res in enumerate(bb):
        for name, xyz in zip(("N", "CA", "C", "O"), res):
            lines.append(_atom(s, name, resname, chain, i + 1, xyz, element=name[0]))
            s += 1
    return "\n".join(lines + ["END"])


def _same(a, b):
    assert a.nb_residues == b["nb_residues"]
    assert np.array_equal(a.atom37_positions.astype(np.float64), b["atom37_positions"])
    assert np.array_equal(a.atom37_gt_exists, b["atom37_gt_exists"])
    assert np.array_equal(a.atom37_atom_exists, b["atom37_atom_exists"])
    assert np.array_equal(a.aatype, b["aatype"])


def test_round_trip_of_a_synthetic_chain():
    from pst import synthetic as syn

    bb = syn.make_backbones(9, [60])[0]
    text = _pdb_from_backbone(bb)
    s = ppdb.structure_from_pdb_string(text)
    _same(s, pdb_ref.parse_pdb(text))
    atoms, mask = s.device_arrays()
    assert atoms.shape == (60, 37, 3) and mask.sum() == 240
    assert np.array_equal(atoms[:, [0, 1, 2, 4]], bb)  # 3-decimal fp32 coordinates survive the text round trip
    assert (s.aatype == 7).all()


def test_tricky_records():
    L = []
    L.append(_atom(1, "N", "ALA", "A", 1, (0, 0, 0)))
    L.append(_atom(2, "CA", "ALA", "A", 1, (1.4, 0, 0)))
    L.append(_atom(3, "C", "ALA", "A", 1, (2.0, 1.4, 0)))
    L.append(_atom(4, "O", "ALA", "A", 1, (1.5, 2.4, 0)))
    L.append(_atom(5, "CB", "ALA", "A", 1, (1.9, -0.8, 1.2)))
    L.append(_atom(6, "HA", "ALA", "A", 1, (1.6, -0.5, -0.9), element="H"))      # hydrogen: dropped
    L.append(_atom(7, "N", "MSE", "A", 2, (3.3, 1.5, 0)))                          # unknown residue -> UNK
    L.append(_atom(8, "CA", "MSE", "A", 2, (4.0, 2.8, 0)))
    L.append(_atom(9, "C", "MSE", "A", 2, (5.5, 2.6, 0)))
    L.append(_atom(10, "O", "MSE", "A", 2, (6.0, 1.5, 0)))
    L.append(_atom(11, "SE", "MSE", "A", 2, (4.0, 4.8, 1.0)))                      # not an atom37 name
    L.append(_atom(12, "N", "SER", "A", 3, (6.2, 3.7, 0)))
    L.append(_atom(13, "CA", "SER", "A", 3, (7.6, 3.7, 0), altloc="A", occ=0.4))
    L.append(_atom(14, "CA", "SER", "A", 3, (7.7, 3.8, 0.1), altloc="B", occ=0.6))  # higher occupancy wins
    L.append(_atom(15, "C", "SER", "A", 3, (8.2, 5.1, 0)))                          # O missing -> invalid residue
    L.append(_atom(16, "O", "HOH", "A", 101, (20, 20, 20), rec="HETATM", element="O"))  # water: kept by the parser, dropped later
    L.append(_atom(17, "N", "GLY", "B", 1, (30, 0, 0)))
    L.append(_atom(18, "CA", "GLY", "B", 1, (31.4, 0, 0)))
    L.append(_atom(19, "C", "GLY", "B", 1, (32.0, 1.4, 0)))
    L.append(_atom(20, "O", "GLY", "B", 1, (31.5, 2.4, 0)))
    text = "\n".join(L)
    s = ppdb.structure_from_pdb_string(text)
    _same(s, pdb_ref.parse_pdb(text))
    assert s.nb_residues == 5
    assert list(s.aatype) == [0, 20, 15, 20, 7]
    assert list(s.valid_backbone()) == [True, True, False, False, True]
    assert np.allclose(s.atom37_positions[2, 1], [7.7, 3.8, 0.1])
    atoms, mask = s.device_arrays()
    assert atoms.shape[0] == 3
    assert mask[0].sum() == 5 and mask[1].sum() == 3  # UNK: O is excluded from the centroid mask


def test_errors_match_the_reference():
    ok = [_atom(1, "N", "GLY", "A", 1, (0, 0, 0)), _atom(2, "CA", "GLY", "A", 1, (1, 0, 0))]
    ins = ok + [_atom(3, "N", "GLY", "A", 2, (3, 0, 0), icode="A")]
    for parse in (ppdb.structure_from_pdb_string, pdb_ref.parse_pdb):
        with pytest.raises(ValueError, match="insertion code"):
            parse("\n".join(ins))
        multi = "MODEL        1\n" + "\n".join(ok) + "\nENDMDL\nMODEL        2\n" + "\n".join(ok) + "\nENDMDL"
        with pytest.raises(ValueError, match="single model"):
            parse(multi)
        single = "MODEL        1\n" + "\n".join(ok) + "\nENDMDL"
        assert (parse(single).nb_residues if parse is ppdb.structure_from_pdb_string else parse(single)["nb_residues"]) == 1


def test_parsed_casp14_fixture_is_what_the_parser_produces(casp14):
    # the committed atom37 fixtures were produced by oracle/pdb_ref.py from the bundled files; sanity-check their shape
    assert len(casp14) == 31
    assert sum(e["pos"].shape[0] for e in casp14.values()) == 5618
    assert sum(e["n_valid"] for e in casp14.values()) == 5616


# ---- the C++ parser behind the C ABI (pst_parse_pdb, csrc/pdb_parse.cc) against both Python restatements ----------
def _same_samples(a, b):
    assert a.nb_residues == b.nb_residues
    assert np.array_equal(a.atom37_positions, b.atom37_positions)
    assert np.array_equal(a.atom37_gt_exists, b.atom37_gt_exists)
    assert np.array_equal(a.atom37_atom_exists, b.atom37_atom_exists)
    assert np.array_equal(a.aatype, b.aatype)


def _tricky_text():
    L = [_atom(1, "N", "ALA", "A", 1, (0, 0, 0)), _atom(2, "CA", "ALA", "A", 1, (1.4, 0, 0)), _atom(3, "C", "ALA", "A", 1, (2.0, 1.4, 0)),
         _atom(4, "O", "ALA", "A", 1, (1.5, 2.4, 0)), _atom(5, "CB", "ALA", "A", 1, (1.9, -0.8, 1.2)),
         _atom(6, "HA", "ALA", "A", 1, (1.6, -0.5, -0.9), element="H"),
         _atom(7, "N", "MSE", "A", 2, (3.3, 1.5, 0)), _atom(8, "CA", "MSE", "A", 2, (4.0, 2.8, 0)), _atom(9, "C", "MSE", "A", 2, (5.5, 2.6, 0)),
         _atom(10, "O", "MSE", "A", 2, (6.0, 1.5, 0)), _atom(11, "SE", "MSE", "A", 2, (4.0, 4.8, 1.0)),
         _atom(12, "N", "SER", "A", 3, (6.2, 3.7, 0)), _atom(13, "CA", "SER", "A", 3, (7.6, 3.7, 0), altloc="A", occ=0.4),
         _atom(14, "CA", "SER", "A", 3, (7.7, 3.8, 0.1), altloc="B", occ=0.6), _atom(15, "C", "SER", "A", 3, (8.2, 5.1, 0)),
         _atom(16, "O", "HOH", "A", 101, (20, 20, 20), rec="HETATM", element="O"),
         _atom(17, "N", "GLY", "B", 1, (30, 0, 0)), _atom(18, "CA", "GLY", "B", 1, (31.4, 0, 0)), _atom(19, "C", "GLY", "B", 1, (32.0, 1.4, 0)),
         _atom(20, "O", "GLY", "B", 1, (31.5, 2.4, 0)),
         _atom(21, "OXT", "ALA", "A", 1, (0.5, 3.0, 0.5), element="O"),       # chain A again after chain B: grouped with chain A
         _atom(22, "N", "TRP", "A", 4, (9.0, 6.0, 0)), _atom(23, "CH2", "TRP", "A", 4, (12.0, 9.0, 1.0))]
    return "\n".join(L) + "\nTER\nEND\n"


def test_native_parser_matches_python_parser(built_lib):
    text = _tricky_text()
    a = ppdb.structure_from_pdb_bytes_native(text.encode())
    _same_samples(a, ppdb.structure_from_pdb_string(text))
    _same(a, pdb_ref.parse_pdb(text))
    assert a.nb_residues == 6 and list(a.aatype) == [0, 20, 15, 20, 17, 7]


def test_native_parser_errors_match_the_reference(built_lib):
    ok = [_atom(1, "N", "GLY", "A", 1, (0, 0, 0)), _atom(2, "CA", "GLY", "A", 1, (1, 0, 0))]
    with pytest.raises(ValueError, match="insertion code"):
        ppdb.structure_from_pdb_bytes_native("\n".join(ok + [_atom(3, "N", "GLY", "A", 2, (3, 0, 0), icode="A")]).encode())
    multi = "MODEL        1\n" + "\n".join(ok) + "\nENDMDL\nMODEL        2\n" + "\n".join(ok) + "\nENDMDL"
    with pytest.raises(ValueError, match="single model"):
        ppdb.structure_from_pdb_bytes_native(multi.encode())
    with pytest.raises(ValueError, match="single model"):
        ppdb.structure_from_pdb_bytes_native(b"REMARK nothing here\nEND\n")
    single = "MODEL        1\n" + "\n".join(ok) + "\nENDMDL"
    assert ppdb.structure_from_pdb_bytes_native(single.encode()).nb_residues == 1


def test_native_parser_round_trips_the_casp14_fixture(built_lib, casp14):
    """atom37 arrays of the 31 bundled structures -> PDB text -> C++ parser: identical arrays (5 618 residues)"""
    total = 0
    for name, e in casp14.items():
        lines, serial = [], 1
        n = e["pos"].shape[0]
        for i in range(n):
            # a residue name whose canonical atom set is exactly the committed atom_exists row
            rn = next((r for r, ex in ppdb._EXISTS.items() if np.array_equal(ex, e["exists"][i])), None)
            assert rn is not None
            for a in range(37):
                if e["gt"][i, a]:
                    lines.append(_atom(serial % 100000, ppdb.ATOM_TYPES[a], rn if rn != "UNK" else "XYZ", "A", i + 1, e["pos"][i, a]))
                    serial += 1
        s = ppdb.structure_from_pdb_bytes_native(("\n".join(lines) + "\nEND\n").encode())
        assert s.nb_residues == n
        assert np.array_equal(s.atom37_positions, e["pos"]) and np.array_equal(s.atom37_gt_exists, e["gt"])
        assert np.array_equal(s.atom37_atom_exists, e["exists"])
        total += n
    assert total == 5618


def test_native_parser_on_the_bundled_pdb_files(built_lib):
    import glob
    import os

    files = sorted(glob.glob("/root/reference/casp14_pdbs/*.pdb"))
    if not files:
        pytest.skip("reference checkout not present (development container only)")
    for f in files:
        with open(f, "rb") as fh:
            data = fh.read()
        _same_samples(ppdb.structure_from_pdb_bytes_native(data), ppdb.structure_from_pdb_string(data.decode()))


def test_batch_parser_equals_the_single_file_parser(built_lib, tmp_path):
    """pst_parse_pdb_batch (host threads inside the library): every file's rows equal the single-file call, a bad file
    reports its own status without disturbing its neighbours, and the capacity retry / counts-only modes work."""
    import ctypes as C

    from pst import _lib
    from pst import synthetic as syn

    bbs = syn.make_backbones(23, [60, 131, 77, 52, 300, 64, 90])
    texts = [_pdb_from_backbone(bb, resname=rn) for bb, rn in zip(bbs, ("GLY", "ALA", "TRP", "XYZ", "SER", "LYS", "VAL"))]
    texts.insert(2, _tricky_text())
    ok = [_atom(1, "N", "GLY", "A", 1, (0, 0, 0)), _atom(2, "CA", "GLY", "A", 1, (1, 0, 0))]
    texts.insert(4, "\n".join(ok + [_atom(3, "N", "GLY", "A", 2, (3, 0, 0), icode="A")]))  # insertion code
    texts.insert(6, "REMARK nothing here\nEND\n")                                            # no model
    datas = [t.encode() for t in texts]
    for n_threads in (1, 4, 0):
        out = ppdb.structures_from_pdb_bytes_batch_native(datas, n_threads)
        assert len(out) == len(datas)
        for i, (d, o) in enumerate(zip(datas, out)):
            if i in (4, 6):
                assert isinstance(o, ValueError) and ("insertion code" in str(o) if i == 4 else "single model" in str(o))
            else:
                _same_samples(o, ppdb.structure_from_pdb_bytes_native(d))
    # the C entry point itself: counts only, then a capacity that is too small
    lib = _lib.load()
    nf = len(datas)
    arr = (C.c_char_p * nf)(*datas)
    sizes = (C.c_size_t * nf)(*[len(d) for d in datas])
    offs = np.zeros(nf + 1, np.int32)
    status = np.zeros(nf, np.int32)
    assert lib.pst_parse_pdb_batch(arr, sizes, nf, 3, 0, None, None, None, None, offs.ctypes.data, status.ctypes.data) == 0
    good = [o for o in ppdb.structures_from_pdb_bytes_batch_native(datas, 2) if not isinstance(o, Exception)]
    assert offs[-1] == sum(o.nb_residues for o in good) and list(status[[4, 6]]) == [-9, -8]
    total = int(offs[-1])
    pos = np.empty((total - 1, 37, 3), np.float32)
    gt = np.empty((total - 1, 37), np.uint8)
    ex = np.empty((total - 1, 37), np.uint8)
    aa = np.empty((total - 1,), np.int32)
    rc = lib.pst_parse_pdb_batch(arr, sizes, nf, 3, total - 1, pos.ctypes.data, gt.ctypes.data, ex.ctypes.data, aa.ctypes.data,
                                 offs.ctypes.data, status.ctypes.data)
    assert rc == -4 and offs[-1] == total  # PST_ERR_WORKSPACE_TOO_SMALL, rows needed
    assert ppdb.structures_from_pdb_bytes_batch_native([], 4) == []
    # files on disk through the runner's helper
    paths = []
    for i in (0, 1, 3):
        f = tmp_path / f"f{i}.pdb"
        f.write_bytes(datas[i])
        paths.append(str(f))
    for o, i in zip(ppdb.structures_from_pdb_files_native(paths, 2), (0, 1, 3)):
        _same_samples(o, ppdb.structure_from_pdb_bytes_native(datas[i]))
    # a missing file takes its own slot and leaves the neighbours alone (pst_parse_pdb_files reads inside the library)
    out = ppdb.structures_from_pdb_files_native([paths[0], str(tmp_path / "missing.pdb"), paths[2]], 3)
    assert isinstance(out[1], FileNotFoundError)
    _same_samples(out[0], ppdb.structure_from_pdb_bytes_native(datas[0]))
    _same_samples(out[2], ppdb.structure_from_pdb_bytes_native(datas[3]))


def test_chain_filter_matches_the_reference_argument(built_lib):
    """`chain_id` of protein_structure_from_pdb_string (protein_structure_sample.py:166-168,201-203): only that chain
    is emitted, an insertion code in a skipped chain is not an error, the model count is still checked on the file."""
    from pst import pdb as ppdb
    from pst import synthetic as syn

    bbs = syn.make_backbones(3, [64, 70])
    a = _pdb_from_backbone(bbs[0], chain="A")
    b = _pdb_from_backbone(bbs[1], chain="B")
    text = "".join(l + "\n" for l in (a + "\n" + b).splitlines() if l.startswith("ATOM"))
    for parse in (ppdb.structure_from_pdb_string, lambda t, c=None: ppdb.structure_from_pdb_bytes_native(t.encode(), c)):
        both = parse(text)
        only_a, only_b = parse(text, "A"), parse(text, "B")
        assert both.nb_residues == 134 and only_a.nb_residues == 64 and only_b.nb_residues == 70
        assert np.array_equal(only_b.atom37_positions, both.atom37_positions[64:])
        assert parse(text, "Z").nb_residues == 0
        lines = text.splitlines()
        bad = "\n".join(l[:26] + "X" + l[27:] if l[21] == "B" and int(l[22:26]) == 5 else l for l in lines) + "\n"
        assert parse(bad, "A").nb_residues == 64          # the insertion code sits in the skipped chain
        with pytest.raises(ValueError):
            parse(bad, "B")
        with pytest.raises(ValueError):
            parse(bad)


# ---- mmCIF (csrc/pdb_parse.cc parse_mmcif_text) --------------------------------------------------------------
_CIF_COLS = ("group_PDB id type_symbol label_atom_id label_alt_id label_comp_id label_asym_id label_entity_id label_seq_id "
             "pdbx_PDB_ins_code Cartn_x Cartn_y Cartn_z occupancy B_iso_or_equiv auth_seq_id auth_comp_id auth_asym_id "
             "auth_atom_id pdbx_PDB_model_num").split()


def _cif_quote(v):
    return f'"{v}"' if ("'" in v or " " in v) else v


def _pdb_to_mmcif(text, name="T", chain_map=None, model=1):
    """The ATOM / HETATM records of a PDB text as an mmCIF `_atom_site` loop (the columns wwPDB files carry)."""
    rows = []
    for line in text.splitlines():
        if line[:6] not in ("ATOM  ", "HETATM"):
            continue
        chain = line[21]
        chain = chain_map.get(chain, chain) if chain_map else chain
        chain = "A" if chain == " " else chain  # CASP models leave the chain column blank; mmCIF needs a value
        alt = line[16] if line[16] != " " else "."
        ins = line[26] if line[26] != " " else "?"
        atom = _cif_quote(line[12:16].strip())
        rn = line[17:20].strip()
        rows.append(" ".join([line[:6].strip(), line[6:11].strip(), (line[76:78].strip() or atom[0]), atom, alt, rn, chain, "1",
                              line[22:26].strip() if line[:4] == "ATOM" else ".", ins, line[30:38].strip(), line[38:46].strip(),
                              line[46:54].strip(), line[54:60].strip() or "1.00", line[60:66].strip() or "0.00",
                              line[22:26].strip(), rn, chain, atom, str(model)]))
    head = [f"data_{name}", "#", f"_entry.id {name}", "#", "_struct.title", ";a title that mentions loop_ and _atom_site.x", "on two lines",
            ";", "#", "loop_", "_entity.id", "_entity.type", "1 polymer", "2 water", "#", "loop_"] + [f"_atom_site.{c}" for c in _CIF_COLS]
    return "\n".join(head + rows + ["#", "loop_", "_pdbx_poly_seq_scheme.asym_id", "_pdbx_poly_seq_scheme.seq_id", "A 1", "#"]) + "\n"


def _same_dict(s, d):
    assert s.nb_residues == d["nb_residues"]
    assert np.array_equal(s.atom37_positions.astype(np.float64), d["atom37_positions"])
    assert np.array_equal(s.atom37_gt_exists, d["atom37_gt_exists"])
    assert np.array_equal(s.atom37_atom_exists, d["atom37_atom_exists"])
    assert np.array_equal(s.aatype, d["aatype"])


def test_mmcif_gives_the_arrays_of_the_same_structure_as_pdb(built_lib):
    """The tricky PDB text (altlocs, HETATM water, UNK, hydrogens, two chains) converted to mmCIF: the C++ parser returns
    the arrays it returns for the PDB text, and both equal the independent Python restatements."""
    from pst import synthetic as syn

    for text in (_tricky_text(), _pdb_from_backbone(syn.make_backbones(3, [75])[0], resname="TRP")):
        cif = _pdb_to_mmcif(text)
        a = ppdb.structure_from_pdb_bytes_native(text.encode())
        b = ppdb.structure_from_pdb_bytes_native(cif.encode())
        _same_samples(a, b)
        _same_dict(b, pdb_ref.parse_mmcif(cif))
        _same_dict(b, pdb_ref.parse_pdb(text))


def test_mmcif_chain_ids_models_insertion_codes_and_wrapped_rows(built_lib):
    text = _tricky_text()
    cif = _pdb_to_mmcif(text, chain_map={"A": "AA", "B": "B2"})
    full = ppdb.structure_from_pdb_bytes_native(cif.encode())
    only = ppdb.structure_from_pdb_bytes_native(cif.encode(), chain_id="B2")  # multi-character chain id
    _same_dict(only, pdb_ref.parse_mmcif(cif, "B2"))
    _same_samples(only, ppdb.structure_from_pdb_bytes_native(text.encode(), chain_id="B"))
    assert 0 < only.nb_residues < full.nb_residues
    assert ppdb.structure_from_pdb_bytes_native(cif.encode(), chain_id="ZZ").nb_residues == 0
    # a row may continue on the next line, values may be quoted
    wrapped = cif.replace(" 1.00 ", " 1.00\n  ").replace(" CA ", " 'CA' ")
    _same_samples(ppdb.structure_from_pdb_bytes_native(wrapped.encode()), full)
    # two models -> the reference's single-model error; an insertion code -> its insertion-code error
    two = cif.rstrip("#\n")
    rows = [l for l in cif.splitlines() if l.startswith(("ATOM", "HETATM"))]
    second = "\n".join(r.rsplit(" ", 1)[0] + " 2" for r in rows)
    head, tail = cif.split(rows[-1])
    with pytest.raises(ValueError, match="single model"):
        ppdb.structure_from_pdb_bytes_native((head + rows[-1] + "\n" + second + tail).encode())
    with pytest.raises(ValueError, match="insertion code"):
        ppdb.structure_from_pdb_bytes_native(cif.replace(" ? ", " A ", 1).encode())
    with pytest.raises(ValueError, match="single model"):
        ppdb.structure_from_pdb_bytes_native(b"data_EMPTY\n_entry.id EMPTY\n")
    assert two  # keep the linter quiet about the helper variable


def test_mmcif_files_go_through_the_batch_and_file_entry_points(built_lib, tmp_path):
    from pst import synthetic as syn

    bbs = syn.make_backbones(31, [60, 90, 64])
    texts = [_pdb_from_backbone(bb, resname=rn) for bb, rn in zip(bbs, ("GLY", "ALA", "LYS"))]
    paths = []
    for i, t in enumerate(texts):
        f = tmp_path / (f"s{i}.cif" if i != 1 else f"s{i}.pdb")
        f.write_text(_pdb_to_mmcif(t) if i != 1 else t)
        paths.append(str(f))
    out = ppdb.structures_from_pdb_files_native(paths, 2)
    for o, t in zip(out, texts):
        _same_samples(o, ppdb.structure_from_pdb_bytes_native(t.encode()))


def test_mmcif_of_the_bundled_casp14_files(built_lib):
    import glob

    files = sorted(glob.glob("/root/reference/casp14_pdbs/*.pdb"))
    if not files:
        pytest.skip("reference checkout not present (development container only)")
    for f in files[:8]:
        with open(f) as fh:
            text = fh.read()
        _same_samples(ppdb.structure_from_pdb_bytes_native(_pdb_to_mmcif(text).encode()), ppdb.structure_from_pdb_bytes_native(text.encode()))


def test_mmcif_detection_by_content_agrees_between_the_wrapper_and_the_library(built_lib):
    """Leading blank lines and # comments before the `data_` header; a PDB text that merely mentions data_ later."""
    import ctypes as C

    from pst import _lib

    text = _tricky_text()
    cif = "\n# produced by a test\n\n" + _pdb_to_mmcif(text)
    assert ppdb.looks_like_mmcif(cif.encode()) and not ppdb.looks_like_mmcif(text.encode())
    pdb_with_remark = "REMARK data_block mentioned in a remark\n" + text
    assert not ppdb.looks_like_mmcif(pdb_with_remark.encode())
    ref = ppdb.structure_from_pdb_bytes_native(text.encode())
    _same_samples(ppdb.structure_from_pdb_bytes_native(cif.encode()), ref)
    _same_samples(ppdb.structure_from_pdb_bytes_native(pdb_with_remark.encode()), ref)
    # the plain PDB entry point of the C ABI takes the mmCIF text as well (detection inside the library)
    lib = _lib.load()
    n = C.c_int32(0)
    data = cif.encode()
    assert lib.pst_parse_pdb(data, len(data), 0, None, None, None, None, C.byref(n)) == 0
    assert n.value == ref.nb_residues
