// ingest.cpp -- native FASTA / FASTQ ingest, pure host code.
//
// The reference parses with one regular expression per record type (/root/reference/src/records.py:141-199,
// 212-302) and books every input character in a Python set; 2 MB of FASTQ take it more than a second.  This file
// parses the CANONICAL form of both formats -- what sequencers and assemblers write -- straight into the packed
// arrays the device path consumes, and reports "not canonical" for everything else, so that the caller falls back to
// the regex restatement in records.py and every unusual input keeps the reference's exact acceptance rules, error
// types, messages and precedence.  For canonical text the regex provably yields the same records:
//   FASTQ  '@' + identifier line | [ACGT]+ line | '+' line (then dots only: the reference's class is "[.]") | quality
//          line (ASCII 33..126) of the same length; the next line starts with '@' or the text ends (one optional line
//          end).  No group can cross a line end (none of the character classes contains '\n'), the lazy groups must
//          stop at the line ends because of the literal "\r?\n" / "\r?\n+" that follow, and the quality group stops
//          where the look-ahead (?=\r?\n@|(\r?\n)?\Z) first holds: the end of its line.  Duplicate identifiers are
//          left to the fallback (DuplicateRecordError).
//   FASTA  '>' + description line, then lines over ACGTN and blanks up to the next line starting with '>' or the end;
//          at least one base.  The genome is the body with all whitespace removed (records.py:192 re.sub(\s)).
// Identifier / description: the reference strips them (records.py:192); so does this parser.  Only ASCII text is
// handled here (the caller checks), control characters other than \t \r \n make the text non-canonical.
//
// Parallel: the text is cut into one piece per worker at record starts found locally (FASTQ: a line starting with
// '@' whose second-next line starts with '+' -- in canonical text only a record's first line has that shape, because
// the line two below a quality line is a sequence line; FASTA: a line starting with '>').  Every piece is parsed
// strictly and must end exactly where the next one begins; by induction from position 0 the pieces then are the
// records of the sequential parse, whatever the cut heuristic did -- a wrong cut can only make a piece fail, i.e.
// send the text to the regex.
#include "../../include/pa_b200.h"
#include "hostpack.h"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace {

struct Piece {
  std::vector<uint8_t> seq, qual;            // concatenated sequences / qualities of the piece
  std::vector<uint64_t> seq_len;             // per record
  std::vector<uint64_t> name_beg, name_len;  // stripped identifier / description inside the input text
  std::vector<uint64_t> plus_beg, plus_len;  // FASTQ: the text after '+'
};

struct Parsed {
  std::vector<Piece> pieces;
  uint64_t n_records = 0, n_bases = 0;
  bool fastq = false;
};

inline bool name_char(uint8_t c) { return (c >= 0x21 && c <= 0x7E) || c == ' ' || c == '\t'; }
inline bool blank(uint8_t c) { return c == ' ' || c == '\t'; }

// end of the line starting at `at` (a '\r' directly before the '\n' belongs to the line end); *next = start of the
// following line (n when the text ends without '\n')
inline void line_at(const uint8_t* t, uint64_t n, uint64_t at, uint64_t* end, uint64_t* next, bool* has_newline) {
  const void* nl = at < n ? memchr(t + at, '\n', n - at) : nullptr;
  if (nl) {
    uint64_t e = (uint64_t)((const uint8_t*)nl - t);
    *next = e + 1;
    *has_newline = true;
    if (e > at && t[e - 1] == '\r') --e;
    *end = e;
  } else {
    *end = n; *next = n; *has_newline = false;
  }
}

bool parse_name(const uint8_t* t, uint64_t beg, uint64_t end, uint64_t* sb, uint64_t* sl) {
  if (end <= beg) return false;
  for (uint64_t i = beg; i < end; ++i) if (!name_char(t[i])) return false;
  while (beg < end && blank(t[beg])) ++beg;
  while (end > beg && blank(t[end - 1])) --end;
  *sb = beg; *sl = end - beg;
  return true;
}

// all bytes of [p, p + n) inside the class given by a 256-entry table (8 at a time: one branch per word)
inline bool all_in(const uint8_t* p, uint64_t n, const uint8_t* table) {
  uint64_t i = 0;
  for (; i + 8 <= n; i += 8) {
    const uint8_t ok = table[p[i]] & table[p[i + 1]] & table[p[i + 2]] & table[p[i + 3]] & table[p[i + 4]] & table[p[i + 5]] &
                       table[p[i + 6]] & table[p[i + 7]];
    if (!ok) return false;
  }
  for (; i < n; ++i) if (!table[p[i]]) return false;
  return true;
}

struct Tables {
  uint8_t acgt[256], qualc[256], base[256];
  Tables() {
    for (int c = 0; c < 256; ++c) {
      acgt[c] = c == 'A' || c == 'C' || c == 'G' || c == 'T';
      qualc[c] = c >= 33 && c <= 126;
      base[c] = acgt[c] || c == 'N';
    }
  }
};
const Tables& tables() { static const Tables t; return t; }

// strict parse of the records in [beg, end); must land exactly on end
bool parse_fastq_range(const uint8_t* t, uint64_t n, uint64_t beg, uint64_t end, Piece& out) {
  const Tables& tb = tables();
  uint64_t pos = beg;
  while (pos < end) {
    if (t[pos] != '@') return false;
    uint64_t e1, n1, e2, n2, e3, n3, e4, n4;
    bool nl;
    line_at(t, n, pos, &e1, &n1, &nl);
    if (!nl) return false;
    uint64_t ib, il;
    if (!parse_name(t, pos + 1, e1, &ib, &il)) return false;
    line_at(t, n, n1, &e2, &n2, &nl);
    if (!nl || e2 <= n1 || !all_in(t + n1, e2 - n1, tb.acgt)) return false;
    if (n2 >= n || t[n2] != '+') return false;
    line_at(t, n, n2, &e3, &n3, &nl);
    if (!nl) return false;
    for (uint64_t i = n2 + 1; i < e3; ++i) if (t[i] != '.') return false;
    line_at(t, n, n3, &e4, &n4, &nl);
    if (e4 <= n3 || e4 - n3 != e2 - n1 || !all_in(t + n3, e4 - n3, tb.qualc)) return false;
    if (nl && n4 < n && t[n4] != '@') return false;     // something other than a record follows
    out.seq.insert(out.seq.end(), t + n1, t + e2);
    out.qual.insert(out.qual.end(), t + n3, t + e4);
    out.seq_len.push_back(e2 - n1);
    out.name_beg.push_back(ib); out.name_len.push_back(il);
    out.plus_beg.push_back(n2 + 1); out.plus_len.push_back(e3 - (n2 + 1));
    pos = n4;
  }
  return pos == end;
}

bool parse_fasta_range(const uint8_t* t, uint64_t n, uint64_t beg, uint64_t end, Piece& out) {
  const Tables& tb = tables();
  uint64_t pos = beg;
  while (pos < end) {
    if (t[pos] != '>') return false;
    uint64_t e1, n1;
    bool nl;
    line_at(t, n, pos, &e1, &n1, &nl);
    if (!nl) return false;
    uint64_t ib, il;
    if (!parse_name(t, pos + 1, e1, &ib, &il)) return false;
    const uint64_t before = out.seq.size();
    uint64_t at = n1;
    while (at < n && t[at] != '>') {        // body lines
      uint64_t e, nx;
      line_at(t, n, at, &e, &nx, &nl);
      if (all_in(t + at, e - at, tb.base)) {
        out.seq.insert(out.seq.end(), t + at, t + e);
      } else {
        for (uint64_t i = at; i < e; ++i) {
          const uint8_t c = t[i];
          if (tb.base[c]) out.seq.push_back(c);
          else if (c != ' ' && c != '\t' && c != '\r') return false;
        }
      }
      at = nx;
    }
    if (out.seq.size() == before) return false;   // no base at all: leave it to the regex
    out.seq_len.push_back(out.seq.size() - before);
    out.name_beg.push_back(ib); out.name_len.push_back(il);
    pos = at;
  }
  return pos == end;
}

// first plausible record start at or after `from` (n when there is none)
uint64_t find_start(const uint8_t* t, uint64_t n, uint64_t from, bool fastq) {
  if (from == 0) return 0;
  if (from >= n) return n;
  const void* nl = memchr(t + from - 1, '\n', n - (from - 1));   // begin at a line start
  uint64_t at = nl ? (uint64_t)((const uint8_t*)nl - t) + 1 : n;
  while (at < n) {
    uint64_t e, nx, e2, nx2;
    bool has;
    line_at(t, n, at, &e, &nx, &has);
    if (fastq) {
      if (t[at] == '@' && has) {
        line_at(t, n, nx, &e2, &nx2, &has);
        if (has && nx2 < n && t[nx2] == '+') return at;
      }
    } else if (t[at] == '>') {
      return at;
    }
    if (nx >= n) break;
    at = nx;
  }
  return n;
}

struct IdTable {   // open addressing over (hash, record) pairs; equality is checked on the bytes
  std::vector<uint64_t> slot_hash, slot_beg, slot_len;
  uint64_t mask = 0;
  explicit IdTable(uint64_t n) {
    uint64_t cap = 16;
    while (cap < 2 * n + 2) cap <<= 1;
    slot_hash.assign(cap, 0); slot_beg.assign(cap, 0); slot_len.assign(cap, UINT64_MAX);
    mask = cap - 1;
  }
  static uint64_t hash(const uint8_t* p, uint64_t n) {
    uint64_t h = 0x9E3779B97F4A7C15ULL ^ (n * 0xff51afd7ed558ccdULL);
    uint64_t i = 0;
    for (; i + 8 <= n; i += 8) { uint64_t w; memcpy(&w, p + i, 8); h = (h ^ w) * 0x9FB21C651E98DF25ULL; h ^= h >> 32; }
    uint64_t w = 0;
    if (i < n) { memcpy(&w, p + i, n - i); h = (h ^ w) * 0x9FB21C651E98DF25ULL; h ^= h >> 32; }
    return h * 0xD6E8FEB86659FD93ULL;
  }
  bool insert(const uint8_t* t, uint64_t beg, uint64_t len) {   // false: already present
    const uint64_t h = hash(t + beg, len);
    for (uint64_t i = (h >> 20) & mask;; i = (i + 1) & mask) {
      if (slot_len[i] == UINT64_MAX) { slot_hash[i] = h; slot_beg[i] = beg; slot_len[i] = len; return true; }
      if (slot_hash[i] == h && slot_len[i] == len && memcmp(t + slot_beg[i], t + beg, len) == 0) return false;
    }
  }
};

bool parse_all(const uint8_t* t, uint64_t n, bool fastq, Parsed& out) {
  if (n == 0) return false;
  out.fastq = fastq;
  uint64_t piece_bytes = 1u << 20;   // PA_INGEST_PIECE_BYTES: the tests cut small texts into many pieces
  if (const char* e = getenv("PA_INGEST_PIECE_BYTES")) { const uint64_t v = strtoull(e, nullptr, 10); if (v) piece_bytes = v; }
  const int n_pieces = (int)std::min<uint64_t>((uint64_t)pa::host_pack_threads(), std::max<uint64_t>(1, n / piece_bytes));
  std::vector<uint64_t> cut(n_pieces + 1, n);
  cut[0] = 0;
  pa::host_parallel_for(n_pieces - 1, [&](int i) { cut[i + 1] = find_start(t, n, n * (uint64_t)(i + 1) / n_pieces, fastq); });
  cut[n_pieces] = n;
  for (int i = 1; i <= n_pieces; ++i) cut[i] = std::max(cut[i], cut[i - 1]);
  out.pieces.assign(n_pieces, Piece());
  std::atomic<bool> ok{true};
  pa::host_parallel_for(n_pieces, [&](int i) {
    if (cut[i] >= cut[i + 1]) return;
    const bool good = fastq ? parse_fastq_range(t, n, cut[i], cut[i + 1], out.pieces[i])
                            : parse_fasta_range(t, n, cut[i], cut[i + 1], out.pieces[i]);
    if (!good) ok.store(false, std::memory_order_relaxed);
  });
  if (!ok.load()) return false;
  for (const Piece& p : out.pieces) { out.n_records += p.name_beg.size(); out.n_bases += p.seq.size(); }
  if (out.n_records == 0) return false;
  if (fastq) {   // identifiers are a unique index (records.py:195-198): a duplicate goes to the fallback, which raises
    IdTable ids(out.n_records);
    for (const Piece& p : out.pieces)
      for (size_t r = 0; r < p.name_beg.size(); ++r)
        if (!ids.insert(t, p.name_beg[r], p.name_len[r])) return false;
  }
  return true;
}

}  // namespace

struct pa_parsed { Parsed p; };

extern "C" {

int32_t pa_parse_records(const uint8_t* text, uint64_t n, int32_t fastq, pa_parsed** out, int32_t* canonical,
                         uint64_t* n_records, uint64_t* n_bases) {
  if (!out || !canonical || !n_records || !n_bases || (n && !text)) return PA_ERR_INVALID_ARG;
  *out = nullptr; *canonical = 0; *n_records = 0; *n_bases = 0;
  pa_parsed* h = new (std::nothrow) pa_parsed();
  if (!h) return PA_ERR_NOMEM;
  bool ok = false;
  try {
    ok = parse_all(text, n, fastq != 0, h->p);
  } catch (const std::bad_alloc&) { delete h; return PA_ERR_NOMEM; }
  if (!ok) { delete h; return PA_OK; }
  *out = h; *canonical = 1;
  *n_records = h->p.n_records;
  *n_bases = h->p.n_bases;
  return PA_OK;
}

int32_t pa_parsed_copy(pa_parsed* h, uint8_t* seq, uint8_t* qual, uint64_t* seq_off, uint64_t* name_beg, uint64_t* name_len,
                       uint64_t* plus_beg, uint64_t* plus_len) {
  if (!h) return PA_ERR_INVALID_ARG;
  const Parsed& P = h->p;
  const int np = (int)P.pieces.size();
  std::vector<uint64_t> rec0(np + 1, 0), base0(np + 1, 0);
  for (int i = 0; i < np; ++i) { rec0[i + 1] = rec0[i] + P.pieces[i].name_beg.size(); base0[i + 1] = base0[i] + P.pieces[i].seq.size(); }
  pa::host_parallel_for(np, [&](int i) {
    const Piece& p = P.pieces[i];
    const uint64_t r0 = rec0[i], b0 = base0[i], nr = p.name_beg.size();
    if (seq && !p.seq.empty()) memcpy(seq + b0, p.seq.data(), p.seq.size());
    if (qual && !p.qual.empty()) memcpy(qual + b0, p.qual.data(), p.qual.size());
    if (seq_off) { uint64_t o = b0; for (uint64_t r = 0; r < nr; ++r) { seq_off[r0 + r] = o; o += p.seq_len[r]; } }
    if (name_beg && nr) memcpy(name_beg + r0, p.name_beg.data(), nr * 8);
    if (name_len && nr) memcpy(name_len + r0, p.name_len.data(), nr * 8);
    if (plus_beg && !p.plus_beg.empty()) memcpy(plus_beg + r0, p.plus_beg.data(), nr * 8);
    if (plus_len && !p.plus_len.empty()) memcpy(plus_len + r0, p.plus_len.data(), nr * 8);
  });
  if (seq_off) seq_off[P.n_records] = P.n_bases;
  return PA_OK;
}

int32_t pa_parsed_free(pa_parsed* h) { delete h; return PA_OK; }

}  // extern "C"
