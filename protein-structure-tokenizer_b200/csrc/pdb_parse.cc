// Host ingest: PDB text -> atom37 arrays.  Plain C++ (no CUDA): at > 10 M residues/s on
// the device, BioPython's parser (tens of ms per file) is what bounds real-file throughput.
//
// Behaviour restated from the reference (structure_tokenizer/data/protein_structure_sample.py:166-248) and the
// BioPython 1.80 PDBParser semantics it relies on:
//   * ATOM / HETATM records of a single model (:187-190: != 1 model -> error); fixed columns: atom name [12:16],
//     altloc [16], resname [17:20], chain [21], resseq [22:26], icode [26], x / y / z [30:38] [38:46] [46:54],
//     occupancy [54:60];
//   * residues are keyed by (chain, hetero flag, resseq, icode) - BioPython's residue id, where the hetero flag is
//     ' ' for ATOM, 'W' for HOH/WAT and 'H_<resname>' for other HETATM - chains in order of first appearance, residues
//     in order of first appearance inside their chain (:201-204);
//   * an insertion code != ' ' is an error (:205-209);
//   * atoms kept iff their name is one of the 37 atom37 types (data/residue_constants.py:539-577, :223-224);
//     disordered atoms: the altloc with the highest occupancy wins (BioPython DisorderedAtom.disordered_add +
//     selected child), a blank-altloc duplicate keeps the first record;
//   * unknown residue names -> UNK (aatype 20; expected atoms N, CA, C, CB) (:211-215, residue_constants.py:733-737);
//   * residues without any kept atom are skipped (:228-230); coordinates are float32 (:216).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <atomic>
#include <thread>
#include <utility>
#include <vector>

#include "pst_abi.h"

namespace {

const char* const kAtomTypes[37] = {"N",   "CA",  "C",   "CB",  "O",   "CG",  "CG1", "CG2", "OG",  "OG1", "SG",  "CD",  "CD1",
                                    "CD2", "ND1", "ND2", "OD1", "OD2", "SD",  "CE",  "CE1", "CE2", "CE3", "NE",  "NE1", "NE2",
                                    "OE1", "OE2", "CH2", "NH1", "NH2", "OH",  "CZ",  "CZ2", "CZ3", "NZ",  "OXT"};
struct ResType {
  const char* name;
  const char* atoms;  // space separated
};
// restype order of residue_constants.py:596-617; atom lists :341-362
const ResType kResTypes[20] = {
    {"ALA", "C CA CB N O"},
    {"ARG", "C CA CB CG CD CZ N NE O NH1 NH2"},
    {"ASN", "C CA CB CG N ND2 O OD1"},
    {"ASP", "C CA CB CG N O OD1 OD2"},
    {"CYS", "C CA CB N O SG"},
    {"GLN", "C CA CB CG CD N NE2 O OE1"},
    {"GLU", "C CA CB CG CD N O OE1 OE2"},
    {"GLY", "C CA N O"},
    {"HIS", "C CA CB CG CD2 CE1 N ND1 NE2 O"},
    {"ILE", "C CA CB CG1 CG2 CD1 N O"},
    {"LEU", "C CA CB CG CD1 CD2 N O"},
    {"LYS", "C CA CB CG CD CE N NZ O"},
    {"MET", "C CA CB CG CE N O SD"},
    {"PHE", "C CA CB CG CD1 CD2 CE1 CE2 CZ N O"},
    {"PRO", "C CA CB CG CD N O"},
    {"SER", "C CA CB N O OG"},
    {"THR", "C CA CB CG2 N O OG1"},
    {"TRP", "C CA CB CG CD1 CD2 CE2 CE3 CZ2 CZ3 CH2 N NE1 O"},
    {"TYR", "C CA CB CG CD1 CD2 CE1 CE2 CZ N O OH"},
    {"VAL", "C CA CB CG1 CG2 N O"},
};

// A stripped field of at most three characters packed into 32 bits (character i in byte i): atom and residue names
// are compared as integers in the per-line loop (no std::string, no strlen).  Returns 0 for an empty field and
// 0xFFFFFFFF for one that is longer than three characters after stripping.
inline uint32_t pack_field(const char* s, size_t n) {
  size_t a = 0, b = n;
  while (a < b && s[a] == ' ') ++a;
  while (b > a && s[b - 1] == ' ') --b;
  const size_t m = b - a;
  if (m > 3) return 0xFFFFFFFFu;
  uint32_t code = 0;
  for (size_t i = 0; i < m; ++i) code |= static_cast<uint32_t>(static_cast<unsigned char>(s[a + i])) << (8 * i);
  return code;
}
inline uint32_t pack_cstr(const char* s) { return pack_field(s, strlen(s)); }

struct Tables {
  uint32_t atom_code[37];
  uint32_t res_code[20];
  uint8_t exists[21][37];  // atom37_atom_exists per residue type (20 = UNK: N, CA, C, CB)
  Tables() {
    for (int i = 0; i < 37; ++i) atom_code[i] = pack_cstr(kAtomTypes[i]);
    memset(exists, 0, sizeof(exists));
    for (int t = 0; t < 20; ++t) {
      res_code[t] = pack_cstr(kResTypes[t].name);
      const char* q = kResTypes[t].atoms;
      while (*q) {
        const char* sp = strchr(q, ' ');
        const size_t m = sp ? static_cast<size_t>(sp - q) : strlen(q);
        const uint32_t code = pack_field(q, m);
        for (int i = 0; i < 37; ++i)
          if (atom_code[i] == code) exists[t][i] = 1;
        q += m + (sp ? 1 : 0);
      }
    }
    exists[20][0] = exists[20][1] = exists[20][2] = exists[20][3] = 1;
  }
};
const Tables& tables() {
  static const Tables t;
  return t;
}

// the same for a fixed-column field, without allocating: atom37 names start with N, C, O or S, which rejects the
// hydrogens (about half of the records of an all-atom file) on one character
// atom name code -> atom37 slot through a 128-entry open-addressing table (the 37-way linear search was a tenth of
// the per-record time)
struct AtomHash {
  uint32_t key[128];
  int8_t slot[128];
  static uint32_t h(uint32_t c) { return (c * 0x9E3779B1u) >> 25; }
  AtomHash() {
    for (int i = 0; i < 128; ++i) { key[i] = 0; slot[i] = -1; }
    const uint32_t* t = tables().atom_code;
    for (int i = 0; i < 37; ++i) {
      uint32_t j = h(t[i]);
      while (slot[j] >= 0) j = (j + 1) & 127;
      key[j] = t[i];
      slot[j] = static_cast<int8_t>(i);
    }
  }
};
int atom_slot_field(const char* s, size_t n) {
  static const AtomHash H;
  const uint32_t code = pack_field(s, n);
  const char c0 = static_cast<char>(code & 0xFF);
  if (c0 != 'C' && c0 != 'N' && c0 != 'O' && c0 != 'S') return -1;
  for (uint32_t j = AtomHash::h(code);; j = (j + 1) & 127) {
    if (H.slot[j] < 0) return -1;
    if (H.key[j] == code) return H.slot[j];
  }
}

std::string strip(const char* s, size_t n) {
  size_t a = 0, b = n;
  while (a < b && (s[a] == ' ' || s[a] == '\t')) ++a;
  while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r')) --b;
  return std::string(s + a, b - a);
}

// float(field): Python's float() of a fixed-column field, rounded to float32 like np.float32(str)
// Fast path for plain fixed-point fields ("  -12.345"): mantissa / 10^k with both exactly representable in a double
// is one correctly rounded division, i.e. the same double strtod (and Python's float()) returns.
bool parse_float(const char* s, size_t n, float* out) {
  {
    size_t a = 0, b = n;
    while (a < b && s[a] == ' ') ++a;
    while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\r')) --b;
    if (a < b) {
      bool neg = false;
      size_t i = a;
      if (s[i] == '-' || s[i] == '+') { neg = s[i] == '-'; ++i; }
      unsigned long long mant = 0;
      int digits = 0, frac = -1;
      bool ok = i < b;
      for (; i < b; ++i) {
        const char c = s[i];
        if (c >= '0' && c <= '9') { mant = mant * 10 + (unsigned)(c - '0'); ++digits; if (frac >= 0) ++frac; }
        else if (c == '.' && frac < 0) frac = 0;
        else { ok = false; break; }
      }
      if (ok && digits > 0 && digits <= 15) {
        static const double p10[16] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};
        double v = (double)mant;
        if (frac > 0) v /= p10[frac];
        *out = static_cast<float>(neg ? -v : v);
        return true;
      }
    }
  }
  std::string t = strip(s, n);
  if (t.empty()) return false;
  char* end = nullptr;
  const double v = std::strtod(t.c_str(), &end);
  if (end == t.c_str() || *end != '\0') return false;
  *out = static_cast<float>(v);
  return true;
}

// residue id -> position in the residue list: open addressing in one allocation (an unordered_map costs a heap node per
// residue, which is also where parser threads running side by side meet in the allocator)
struct FlatIndex {
  std::vector<unsigned long long> keys;
  std::vector<int> vals;
  size_t mask = 0, used = 0;
  explicit FlatIndex(size_t expected) {
    size_t cap = 64;
    while (cap < expected * 2) cap <<= 1;
    keys.assign(cap, ~0ull);  // ~0 is never a valid id (the hetero-kind field is at most 2)
    vals.assign(cap, -1);
    mask = cap - 1;
  }
  static size_t hash(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33;
    return static_cast<size_t>(k);
  }
  int find(unsigned long long k) const {
    for (size_t i = hash(k) & mask;; i = (i + 1) & mask) {
      if (keys[i] == k) return vals[i];
      if (keys[i] == ~0ull) return -1;
    }
  }
  void insert(unsigned long long k, int v) {
    if ((used + 1) * 2 > keys.size()) {
      FlatIndex bigger(keys.size());
      for (size_t i = 0; i < keys.size(); ++i)
        if (keys[i] != ~0ull) bigger.insert(keys[i], vals[i]);
      *this = std::move(bigger);
    }
    size_t i = hash(k) & mask;
    while (keys[i] != ~0ull) i = (i + 1) & mask;
    keys[i] = k;
    vals[i] = v;
    ++used;
  }
};

struct Atom {
  float xyz[3];
  float occ;
  char altloc;
  bool set;
};
struct Residue {
  uint32_t resname;  // pack_field of columns [17:20]
  char chain, icode;
  int resseq;
  int chain_rank;
  Atom atoms[37];
};

// Phase 1: the records of one file -> `residues` (order of first appearance) and `order`, the positions of the residues
// that are emitted, in output order (chains by first appearance, residues without a kept atom skipped).
// only_chain != 0: only that chain is emitted (protein_structure_sample.py:201-203: the chain filter comes BEFORE the
// insertion-code check, so an insertion code in a skipped chain is not an error; the model count is checked on the file).
int parse_text(const char* text, size_t len, std::vector<Residue>& residues, std::vector<int>& order, char only_chain = 0) {
  residues.clear();
  order.clear();
  residues.reserve(len / 640 + 16);  // ~8 records of 81 bytes per residue in an all-atom file: no regrowth copies of the 750-byte entries
  FlatIndex index(len / 640 + 16);   // residue id -> position in `residues`
  unsigned long long last_key = ~0ull;
  int last_idx = -1;
  std::string chain_order;
  int models = 0;
  bool in_model = false, loose_atoms = false;
  size_t pos = 0;
  while (pos < len) {
    const char* line = text + pos;
    const char* nl = static_cast<const char*>(memchr(line, '\n', len - pos));
    const size_t n = nl ? static_cast<size_t>(nl - line) : len - pos;
    pos += n + 1;
    if (n >= 5 && memcmp(line, "MODEL", 5) == 0) { ++models; in_model = true; continue; }
    if (n >= 6 && memcmp(line, "ENDMDL", 6) == 0) { in_model = false; continue; }
    const bool is_atom = n >= 6 && memcmp(line, "ATOM  ", 6) == 0;
    const bool is_het = n >= 6 && memcmp(line, "HETATM", 6) == 0;
    if (!is_atom && !is_het) continue;
    if (n < 54) return PST_ERR_PDB_MALFORMED;
    if (!in_model) loose_atoms = true;
    const uint32_t resname = pack_field(line + 17, 3);
    const char chain = line[21];
    long resseq = 0;
    {  // int(line[22:26]): optional sign and decimal digits between blanks
      size_t a = 22, b = 26;
      while (a < b && line[a] == ' ') ++a;
      while (b > a && line[b - 1] == ' ') --b;
      bool neg = false;
      if (a < b && (line[a] == '-' || line[a] == '+')) { neg = line[a] == '-'; ++a; }
      if (a == b) return PST_ERR_PDB_MALFORMED;
      for (; a < b; ++a) {
        if (line[a] < '0' || line[a] > '9') return PST_ERR_PDB_MALFORMED;
        resseq = resseq * 10 + (line[a] - '0');
      }
      if (neg) resseq = -resseq;
    }
    const char icode = line[26];
    // residue id (chain, hetero flag, resseq, icode) packed into 64 bits: chain 8 | icode 8 | resseq + 2^15 16 |
    // hetero kind 2 (0 ATOM, 1 water, 2 other HETATM) | the 3 resname characters 24 (other HETATM only)
    unsigned long long key = ((unsigned long long)(unsigned char)chain << 56) | ((unsigned long long)(unsigned char)icode << 48) |
                             ((unsigned long long)(unsigned)(resseq + 32768) << 32);
    if (!is_atom) {
      static const uint32_t kHoh = pack_cstr("HOH"), kWat = pack_cstr("WAT");
      if (resname == kHoh || resname == kWat) key |= 1ull << 30;
      else key |= (2ull << 30) | (resname & 0xFFFFFFu);
    }
    Residue* res;
    int found = -1;
    if (key == last_key) found = last_idx;
    else {
      found = index.find(key);
    }
    if (found < 0) {
      size_t rank = chain_order.find(chain);
      if (rank == std::string::npos) { rank = chain_order.size(); chain_order.push_back(chain); }
      found = static_cast<int>(residues.size());
      index.insert(key, found);
      residues.emplace_back();
      res = &residues.back();
      res->resname = resname;
      res->chain = chain;
      res->icode = icode;
      res->resseq = static_cast<int>(resseq);
      res->chain_rank = static_cast<int>(rank);
      memset(res->atoms, 0, sizeof(res->atoms));
    } else {
      res = &residues[found];
    }
    last_key = key;
    last_idx = found;
    const int slot = atom_slot_field(line + 12, 4);
    if (slot < 0) continue;  // hydrogens and names outside atom37 are dropped
    const char altloc = line[16];
    Atom& a = res->atoms[slot];
    float occ = 1.0f;  // only ever compared between two records that both carry an altloc
    if (altloc != ' ' && (n < 60 || !parse_float(line + 54, 6, &occ))) occ = 1.0f;
    if (!a.set || (altloc != ' ' && a.altloc != ' ' && occ > a.occ)) {
      float x, y, z;
      if (!parse_float(line + 30, 8, &x) || !parse_float(line + 38, 8, &y) || !parse_float(line + 46, 8, &z)) return PST_ERR_PDB_MALFORMED;
      a.xyz[0] = x; a.xyz[1] = y; a.xyz[2] = z;
      a.occ = occ;
      a.altloc = altloc;
      a.set = true;
    }
  }
  const int n_models = models > 0 ? models : (loose_atoms ? 1 : 0);
  if (n_models != 1) return PST_ERR_PDB_MODEL_COUNT;
  const int n_chains = static_cast<int>(chain_order.size());
  for (int c = 0; c < n_chains; ++c) {
    for (size_t r = 0; r < residues.size(); ++r) {
      const Residue& res = residues[r];
      if (res.chain_rank != c) continue;
      if (only_chain && res.chain != only_chain) continue;
      if (res.icode != ' ') return PST_ERR_PDB_INSERTION_CODE;
      bool any = false;
      for (int s = 0; s < 37; ++s) any = any || res.atoms[s].set;
      if (any) order.push_back(static_cast<int>(r));
    }
  }
  return PST_OK;
}

// ---- mmCIF ---------------------------------------------------------------------------------------------------
// The reference itself has no mmCIF reader (its two entry points take PDB text and ProteinStructureSample .npy files);
// SURVEY section 8f lists one as the next ingest format.  The `_atom_site` loop goes through the same residue / atom
// selection as the PDB records above, with the BioPython 1.80 MMCIFParser conventions where the formats differ:
// chain = auth_asym_id (label_asym_id if absent; may be longer than one character), residue number = auth_seq_id
// (label_seq_id if absent), atom name = label_atom_id, altloc '.' / '?' = none, insertion code '?' / '.' = none, hetero
// flag from group_PDB ('W' for HOH / WAT, 'H' for other HETATM: the residue name is not part of the id here), model
// = pdbx_PDB_model_num (a file with more than one distinct value is rejected like a multi-model PDB file).
struct CifTok {
  const char* p;
  size_t n;
  bool quoted;
};
// Next token at or after `pos` (STAR syntax: whitespace-separated values, '...' / "..." quoted values that end at a
// quote followed by whitespace, ;...; text fields that start in column 1, # comments).  Returns false at the end.
struct CifWs {
  bool ws[256];
  CifWs() {
    for (int i = 0; i < 256; ++i) ws[i] = false;
    ws[(unsigned char)' '] = ws[(unsigned char)'\t'] = ws[(unsigned char)'\r'] = ws[(unsigned char)'\n'] = true;
  }
};
const bool* cif_ws() {
  static const CifWs t;
  return t.ws;
}
bool cif_next(const char* t, size_t len, size_t& pos, CifTok& tok) {
  const bool* ws = cif_ws();
  for (;;) {
    while (pos < len && ws[(unsigned char)t[pos]]) ++pos;
    if (pos >= len) return false;
    if (t[pos] == '#') {  // comment to the end of the line
      const char* nl = static_cast<const char*>(memchr(t + pos, '\n', len - pos));
      pos = nl ? static_cast<size_t>(nl - t) : len;
      continue;
    }
    break;
  }
  const char c0 = t[pos];
  if (c0 != ';' && c0 != '\'' && c0 != '"') {  // the common case: a bare value
    const size_t a = pos;
    while (pos < len && !ws[(unsigned char)t[pos]]) ++pos;
    tok = {t + a, pos - a, false};
    return true;
  }
  const bool line_start = pos == 0 || t[pos - 1] == '\n';
  if (c0 == ';' && line_start) {  // text field: up to the next line that starts with ';'
    const size_t a = pos + 1;
    size_t q = a;
    for (;;) {
      const char* nl = static_cast<const char*>(memchr(t + q, '\n', len - q));
      if (!nl) { q = len; break; }
      q = static_cast<size_t>(nl - t) + 1;
      if (q < len && t[q] == ';') break;
    }
    tok = {t + a, (q > a ? q - 1 : a) - a, true};
    pos = q < len ? q + 1 : len;
    return true;
  }
  if (c0 == '\'' || c0 == '"') {
    const size_t a = pos + 1;
    size_t q = a;
    while (q < len && t[q] != '\n' && !(t[q] == c0 && (q + 1 >= len || ws[(unsigned char)t[q + 1]]))) ++q;
    tok = {t + a, q - a, true};
    pos = q < len && t[q] == c0 ? q + 1 : q;
    return true;
  }
  const size_t a = pos;  // a ';' that is not at the start of a line is an ordinary character
  while (pos < len && !ws[(unsigned char)t[pos]]) ++pos;
  tok = {t + a, pos - a, false};
  return true;
}
inline bool tok_is(const CifTok& k, const char* s) { return k.n == strlen(s) && memcmp(k.p, s, k.n) == 0; }
inline bool tok_starts(const CifTok& k, const char* s) { const size_t m = strlen(s); return k.n >= m && memcmp(k.p, s, m) == 0; }
inline bool tok_null(const CifTok& k) { return !k.quoted && k.n == 1 && (k.p[0] == '.' || k.p[0] == '?'); }

bool looks_like_mmcif(const char* t, size_t len) {
  size_t pos = 0;
  CifTok k;
  return cif_next(t, len, pos, k) && !k.quoted && tok_starts(k, "data_");
}

int parse_mmcif_text(const char* text, size_t len, std::vector<Residue>& residues, std::vector<int>& order, const char* only_chain) {
  residues.clear();
  order.clear();
  size_t pos = 0;
  CifTok k;
  // find the `_atom_site.` loop
  std::vector<std::string> tags;
  bool have = cif_next(text, len, pos, k);
  bool found_loop = false;
  while (have) {
    if (!k.quoted && tok_is(k, "loop_")) {
      tags.clear();
      have = cif_next(text, len, pos, k);
      while (have && !k.quoted && k.n > 0 && k.p[0] == '_') {
        tags.emplace_back(k.p, k.n);
        have = cif_next(text, len, pos, k);
      }
      if (!tags.empty() && tags[0].compare(0, 11, "_atom_site.") == 0) { found_loop = true; break; }
      continue;  // k is the first value of some other loop (or the next keyword)
    }
    have = cif_next(text, len, pos, k);
  }
  if (!found_loop) return PST_ERR_PDB_MODEL_COUNT;  // no atoms: zero models, as for a PDB file without ATOM records
  auto col = [&](const char* name) {
    for (size_t i = 0; i < tags.size(); ++i)
      if (tags[i].compare(11, std::string::npos, name) == 0) return static_cast<int>(i);
    return -1;
  };
  auto col2 = [&](const char* a, const char* b) { const int c = col(a); return c >= 0 ? c : col(b); };
  const int c_group = col("group_PDB"), c_atom = col2("label_atom_id", "auth_atom_id"), c_alt = col("label_alt_id"),
            c_comp = col2("label_comp_id", "auth_comp_id"), c_chain = col2("auth_asym_id", "label_asym_id"),
            c_seq = col2("auth_seq_id", "label_seq_id"), c_seq_label = col("label_seq_id"), c_ins = col("pdbx_PDB_ins_code"),
            c_x = col("Cartn_x"), c_y = col("Cartn_y"), c_z = col("Cartn_z"), c_occ = col("occupancy"),
            c_model = col("pdbx_PDB_model_num");
  if (c_atom < 0 || c_comp < 0 || c_chain < 0 || c_seq < 0 || c_x < 0 || c_y < 0 || c_z < 0) return PST_ERR_PDB_MALFORMED;
  const size_t nc = tags.size();
  residues.reserve(len / 700 + 16);
  FlatIndex index(len / 700 + 16);
  std::vector<std::string> chains;  // in order of first appearance
  std::vector<long> models;
  std::vector<CifTok> row(nc);
  unsigned long long last_key = ~0ull;
  int last_idx = -1, last_chain = -1;
  static const uint32_t kHoh = pack_cstr("HOH"), kWat = pack_cstr("WAT");
  while (have) {
    if (!k.quoted && (k.p[0] == '_' || tok_is(k, "loop_") || tok_starts(k, "data_") || tok_starts(k, "save_"))) break;
    size_t got = 0;
    while (have && got < nc) {
      row[got++] = k;
      have = cif_next(text, len, pos, k);
    }
    if (got < nc) return PST_ERR_PDB_MALFORMED;  // truncated row
    if (c_model >= 0) {
      long mnum = 0;
      const CifTok& m = row[c_model];
      for (size_t i = 0; i < m.n; ++i) {
        if (m.p[i] < '0' || m.p[i] > '9') return PST_ERR_PDB_MALFORMED;
        mnum = mnum * 10 + (m.p[i] - '0');
      }
      bool seen = false;
      for (long v : models) seen = seen || v == mnum;
      if (!seen) models.push_back(mnum);
      if (models.size() > 1) return PST_ERR_PDB_MODEL_COUNT;
    }
    const bool is_atom = c_group < 0 || !tok_is(row[c_group], "HETATM");
    const uint32_t resname = pack_field(row[c_comp].p, row[c_comp].n);
    int chain_idx = -1;
    {
      const CifTok& c = row[c_chain];
      if (last_chain >= 0 && chains[last_chain].size() == c.n && memcmp(chains[last_chain].data(), c.p, c.n) == 0) chain_idx = last_chain;
      else {
        for (size_t i = 0; i < chains.size(); ++i)
          if (chains[i].size() == c.n && memcmp(chains[i].data(), c.p, c.n) == 0) { chain_idx = static_cast<int>(i); break; }
        if (chain_idx < 0) {
          if (chains.size() >= 65535) return PST_ERR_PDB_MALFORMED;
          chain_idx = static_cast<int>(chains.size());
          chains.emplace_back(c.p, c.n);
        }
      }
      last_chain = chain_idx;
    }
    long long resseq = 0;
    {
      const CifTok* q = &row[c_seq];
      if (tok_null(*q) && c_seq_label >= 0) q = &row[c_seq_label];
      size_t a = 0;
      bool neg = false;
      if (a < q->n && (q->p[a] == '-' || q->p[a] == '+')) { neg = q->p[a] == '-'; ++a; }
      if (a == q->n || q->n - a > 12) return PST_ERR_PDB_MALFORMED;
      for (; a < q->n; ++a) {
        if (q->p[a] < '0' || q->p[a] > '9') return PST_ERR_PDB_MALFORMED;
        resseq = resseq * 10 + (q->p[a] - '0');
      }
      if (neg) resseq = -resseq;
    }
    const char icode = (c_ins < 0 || tok_null(row[c_ins]) || row[c_ins].n == 0) ? ' ' : row[c_ins].p[0];
    // residue id packed into 64 bits: chain index 16 | icode 8 | hetero kind 2 | residue number (offset) 38
    const unsigned long long het = is_atom ? 0ull : ((resname == kHoh || resname == kWat) ? 1ull : 2ull);
    const unsigned long long key = ((unsigned long long)chain_idx << 48) | ((unsigned long long)(unsigned char)icode << 40) | (het << 38) |
                                   ((unsigned long long)(resseq + (1LL << 36)) & ((1ull << 38) - 1));
    int found = key == last_key ? last_idx : index.find(key);
    Residue* res;
    if (found < 0) {
      found = static_cast<int>(residues.size());
      index.insert(key, found);
      residues.emplace_back();
      res = &residues.back();
      res->resname = resname;
      res->chain = 0;
      res->icode = icode;
      res->resseq = static_cast<int>(resseq);
      res->chain_rank = chain_idx;
      memset(res->atoms, 0, sizeof(res->atoms));
    } else {
      res = &residues[found];
    }
    last_key = key;
    last_idx = found;
    const int slot = atom_slot_field(row[c_atom].p, row[c_atom].n);
    if (slot < 0) continue;
    const char altloc = (c_alt < 0 || tok_null(row[c_alt]) || row[c_alt].n == 0) ? ' ' : row[c_alt].p[0];
    Atom& a = res->atoms[slot];
    float occ = 1.0f;
    if (altloc != ' ' && (c_occ < 0 || !parse_float(row[c_occ].p, row[c_occ].n, &occ))) occ = 1.0f;
    if (!a.set || (altloc != ' ' && a.altloc != ' ' && occ > a.occ)) {
      float x, y, z;
      if (!parse_float(row[c_x].p, row[c_x].n, &x) || !parse_float(row[c_y].p, row[c_y].n, &y) || !parse_float(row[c_z].p, row[c_z].n, &z))
        return PST_ERR_PDB_MALFORMED;
      a.xyz[0] = x; a.xyz[1] = y; a.xyz[2] = z;
      a.occ = occ;
      a.altloc = altloc;
      a.set = true;
    }
  }
  if (residues.empty()) return PST_ERR_PDB_MODEL_COUNT;
  int want = -1;  // chain filter (the reference's chain_id argument)
  if (only_chain && only_chain[0]) {
    for (size_t i = 0; i < chains.size(); ++i)
      if (chains[i] == only_chain) want = static_cast<int>(i);
    if (want < 0) return PST_OK;  // no such chain: nothing is emitted
  }
  // chains in order of first appearance, residues in order of first appearance inside their chain (stable bucket sort)
  std::vector<int> start(chains.size() + 1, 0);
  for (const Residue& r : residues) ++start[r.chain_rank + 1];
  for (size_t c = 0; c < chains.size(); ++c) start[c + 1] += start[c];
  std::vector<int> sorted(residues.size());
  {
    std::vector<int> fill(start.begin(), start.end() - 1);
    for (size_t r = 0; r < residues.size(); ++r) sorted[fill[residues[r].chain_rank]++] = static_cast<int>(r);
  }
  for (int r : sorted) {
    const Residue& res = residues[r];
    if (want >= 0 && res.chain_rank != want) continue;
    if (res.icode != ' ') return PST_ERR_PDB_INSERTION_CODE;
    bool any = false;
    for (int s = 0; s < 37; ++s) any = any || res.atoms[s].set;
    if (any) order.push_back(r);
  }
  return PST_OK;
}

// PDB or mmCIF by content: an mmCIF file starts (after comments) with a `data_` block header
int parse_any(const char* text, size_t len, std::vector<Residue>& residues, std::vector<int>& order, char only_chain = 0) {
  if (looks_like_mmcif(text, len)) {
    const char chain[2] = {only_chain, 0};
    return parse_mmcif_text(text, len, residues, order, only_chain ? chain : nullptr);
  }
  return parse_text(text, len, residues, order, only_chain);
}

// Phase 2: one residue -> row `row` of the caller's arrays
void emit_residue(const Residue& res, size_t row, float* atom37_positions, uint8_t* gt_exists, uint8_t* atom_exists, int32_t* aatype) {
  const Tables& T = tables();
  int rt = 20;
  for (int t = 0; t < 20; ++t)
    if (res.resname == T.res_code[t]) { rt = t; break; }
  float* p = atom37_positions + row * 37 * 3;
  uint8_t* g = gt_exists + row * 37;
  for (int s = 0; s < 37; ++s) {
    const Atom& a = res.atoms[s];
    p[s * 3] = a.set ? a.xyz[0] : 0.f;
    p[s * 3 + 1] = a.set ? a.xyz[1] : 0.f;
    p[s * 3 + 2] = a.set ? a.xyz[2] : 0.f;
    g[s] = a.set ? 1 : 0;
  }
  memcpy(atom_exists + row * 37, T.exists[rt], 37);
  aatype[row] = rt;
}

}  // namespace

extern "C" int pst_parse_pdb(const char* text, size_t len, int max_residues, float* atom37_positions, uint8_t* gt_exists,
                             uint8_t* atom_exists, int32_t* aatype, int32_t* n_residues_out) {
  return pst_parse_pdb_chain(text, len, 0, max_residues, atom37_positions, gt_exists, atom_exists, aatype, n_residues_out);
}

extern "C" int pst_parse_pdb_chain(const char* text, size_t len, char chain_id, int max_residues, float* atom37_positions,
                                   uint8_t* gt_exists, uint8_t* atom_exists, int32_t* aatype, int32_t* n_residues_out) {
  if (!text || !n_residues_out) return PST_ERR_BAD_ARGUMENT;
  *n_residues_out = 0;
  std::vector<Residue> residues;
  std::vector<int> order;
  const int rc = parse_any(text, len, residues, order, chain_id);
  if (rc != PST_OK) return rc;
  const int n_out = static_cast<int>(order.size());
  *n_residues_out = n_out;
  if (!(atom37_positions && gt_exists && atom_exists && aatype)) return PST_OK;
  const int n_emit = n_out < max_residues ? n_out : (max_residues > 0 ? max_residues : 0);
  for (int i = 0; i < n_emit; ++i) emit_residue(residues[order[i]], static_cast<size_t>(i), atom37_positions, gt_exists, atom_exists, aatype);
  return n_out > max_residues ? PST_ERR_WORKSPACE_TOO_SMALL : PST_OK;
}

extern "C" int pst_parse_mmcif(const char* text, size_t len, const char* chain_id, int max_residues, float* atom37_positions,
                              uint8_t* gt_exists, uint8_t* atom_exists, int32_t* aatype, int32_t* n_residues_out) {
  if (!text || !n_residues_out) return PST_ERR_BAD_ARGUMENT;
  *n_residues_out = 0;
  std::vector<Residue> residues;
  std::vector<int> order;
  const int rc = parse_mmcif_text(text, len, residues, order, chain_id);
  if (rc != PST_OK) return rc;
  const int n_out = static_cast<int>(order.size());
  *n_residues_out = n_out;
  if (!(atom37_positions && gt_exists && atom_exists && aatype)) return PST_OK;
  const int n_emit = n_out < max_residues ? n_out : (max_residues > 0 ? max_residues : 0);
  for (int i = 0; i < n_emit; ++i) emit_residue(residues[order[i]], static_cast<size_t>(i), atom37_positions, gt_exists, atom_exists, aatype);
  return n_out > max_residues ? PST_ERR_WORKSPACE_TOO_SMALL : PST_OK;
}

// Many files side by side on host threads (the feeder).  Each worker parses whole files into
// its own residue lists (phase 1), the residue counts are prefix-summed, then the workers write the rows (phase 2).
// `paths` != nullptr: the workers also read the files (one fread per file), texts / text_bytes are ignored.
namespace {

int parse_batch(const char* const* texts, const size_t* text_bytes, const char* const* paths, int n_files, int n_threads,
                int max_residues_total, float* atom37_positions, uint8_t* gt_exists, uint8_t* atom_exists, int32_t* aatype,
                int32_t* residue_offsets_out, int32_t* status_out) {
  if (n_files < 0 || !residue_offsets_out || !status_out) return PST_ERR_BAD_ARGUMENT;
  if (n_files > 0 && !paths && (!texts || !text_bytes)) return PST_ERR_BAD_ARGUMENT;
  residue_offsets_out[0] = 0;
  if (n_files == 0) return PST_OK;
  if (n_threads <= 0) n_threads = static_cast<int>(std::thread::hardware_concurrency());
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n_files) n_threads = n_files;
  std::vector<std::vector<Residue>> residues(n_files);
  std::vector<std::vector<int>> order(n_files);
  std::atomic<int> next{0};
  auto run = [&](auto&& body) {
    next.store(0);
    auto worker = [&]() {
      std::string file;  // reused from file to file by this worker
      for (int i = next.fetch_add(1); i < n_files; i = next.fetch_add(1)) body(i, file);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(worker);
    worker();
    for (std::thread& t : pool) t.join();
  };
  run([&](int i, std::string& file) {
    const char* text = nullptr;
    size_t len = 0;
    int rc = PST_OK;
    if (paths) {
      FILE* fh = paths[i] ? fopen(paths[i], "rb") : nullptr;
      if (!fh) rc = PST_ERR_FILE_NOT_FOUND;
      else {
        file.clear();
        char buf[1 << 16];
        size_t got;
        while ((got = fread(buf, 1, sizeof buf, fh)) > 0) file.append(buf, got);
        fclose(fh);
        text = file.data();
        len = file.size();
      }
    } else if (!texts[i]) {
      rc = PST_ERR_BAD_ARGUMENT;
    } else {
      text = texts[i];
      len = text_bytes[i];
    }
    status_out[i] = rc == PST_OK ? parse_any(text, len, residues[i], order[i]) : rc;
    if (status_out[i] != PST_OK) order[i].clear();
  });
  long long total = 0;
  for (int i = 0; i < n_files; ++i) {
    total += static_cast<long long>(order[i].size());
    if (total > 0x7fffffffLL) return PST_ERR_BAD_ARGUMENT;
    residue_offsets_out[i + 1] = static_cast<int32_t>(total);
  }
  if (!(atom37_positions && gt_exists && atom_exists && aatype)) return PST_OK;  // counts only
  if (total > max_residues_total) return PST_ERR_WORKSPACE_TOO_SMALL;
  run([&](int i, std::string&) {
    const size_t base = static_cast<size_t>(residue_offsets_out[i]);
    for (size_t k = 0; k < order[i].size(); ++k)
      emit_residue(residues[i][order[i][k]], base + k, atom37_positions, gt_exists, atom_exists, aatype);
  });
  return PST_OK;
}

}  // namespace

extern "C" int pst_parse_pdb_batch(const char* const* texts, const size_t* text_bytes, int n_files, int n_threads,
                                   int max_residues_total, float* atom37_positions, uint8_t* gt_exists, uint8_t* atom_exists,
                                   int32_t* aatype, int32_t* residue_offsets_out, int32_t* status_out) {
  return parse_batch(texts, text_bytes, nullptr, n_files, n_threads, max_residues_total, atom37_positions, gt_exists, atom_exists,
                     aatype, residue_offsets_out, status_out);
}

extern "C" int pst_parse_pdb_files(const char* const* paths, int n_files, int n_threads, int max_residues_total,
                                   float* atom37_positions, uint8_t* gt_exists, uint8_t* atom_exists, int32_t* aatype,
                                   int32_t* residue_offsets_out, int32_t* status_out) {
  if (n_files > 0 && !paths) return PST_ERR_BAD_ARGUMENT;
  return parse_batch(nullptr, nullptr, paths, n_files, n_threads, max_residues_total, atom37_positions, gt_exists, atom_exists,
                     aatype, residue_offsets_out, status_out);
}
