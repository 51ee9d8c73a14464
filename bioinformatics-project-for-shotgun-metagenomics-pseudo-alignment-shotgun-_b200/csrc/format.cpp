// format.cpp -- `dumpref` without Python dictionaries, pure host code.
//
// KmerReference.get_summary (/root/reference/src/kmer.py:300-329) builds {"Kmers": {kmer: {description: [positions]}}}
// as nested Python objects and main.py:127 prints it with json.dumps(indent=4).  This file writes that "Kmers" object
// as text straight from the exported CSR, byte for byte what json.dumps produces:
//   * k-mers in dict insertion order (`order` = ascending first occurrence),
//   * inside a k-mer one entry per DESCRIPTION in order of first appearance (ascending genome index); genomes that
//     share a description share the entry, and -- dict assignment semantics -- the positions of the LAST such genome
//     win (kmer.py:311-313),
//   * positions ascending, one per line; item separator ",\n", key separator ": ", `indent` spaces per level.
// The descriptions arrive already JSON-escaped (json.dumps of each string, quotes included) from the Python side.
#include "../../include/pa_b200.h"
#include "hostpack.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

namespace {

inline void append_uint(std::string& s, uint64_t v) {
  char buf[24];
  int n = 0;
  do { buf[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  while (n) s.push_back(buf[--n]);
}

// inverse of the bijective key mix of common.cuh (kept in sync by tests/test_cpu_host.py against pa_decode_kmers)
inline uint64_t unmix(uint64_t x, int k) {
  const uint64_t mask = (2 * k >= 64) ? ~0ULL : ((1ULL << (2 * k)) - 1);
  const uint32_t shift = (uint32_t)k;
  auto unxs = [&](uint64_t v) { uint64_t r = v; for (uint32_t s = shift; s < 64; s += shift) r = v ^ (r >> shift); return r & mask; };
  x = unxs(x);
  x = (x * 0xCFEE444D8B59A89BULL) & mask;
  x = unxs(x);
  x = (x * 0xF1DE83E19937733DULL) & mask;
  x = unxs(x);
  return x;
}

}  // namespace

extern "C" {

int32_t pa_format_kmers_json(int32_t k, uint64_t n_keys, const uint64_t* keys, const uint32_t* order, const uint64_t* run_off,
                             const uint32_t* run_genome, const uint64_t* pos_off, const uint32_t* pos,
                             const uint32_t* desc_class, const uint8_t* desc_json, const uint64_t* desc_json_off,
                             int32_t indent, int32_t level, uint8_t** out_text, uint64_t* out_len) {
  if (!out_text || !out_len) return PA_ERR_INVALID_ARG;
  *out_text = nullptr; *out_len = 0;
  if (n_keys && (!keys || !order || !run_off || !run_genome || !pos_off || !pos || !desc_class || !desc_json || !desc_json_off))
    return PA_ERR_INVALID_ARG;
  if (k < 0 || k > 31 || indent < 0 || level < 0) return PA_ERR_INVALID_ARG;
  static const char dec[4] = {'A', 'C', 'T', 'G'};
  const std::string ind0((size_t)indent * level, ' '), ind1((size_t)indent * (level + 1), ' '),
      ind2((size_t)indent * (level + 2), ' '), ind3((size_t)indent * (level + 3), ' ');
  const int n_tasks = (int)std::min<uint64_t>((uint64_t)pa::host_pack_threads() * 4, std::max<uint64_t>(1, n_keys / 4096));
  std::vector<std::string> parts(n_tasks);
  try {
    pa::host_parallel_for(n_tasks, [&](int t) {
      const uint64_t a = n_keys * (uint64_t)t / n_tasks, b = n_keys * (uint64_t)(t + 1) / n_tasks;
      std::string& s = parts[t];
      s.reserve((size_t)(b - a) * 96);
      std::vector<uint32_t> cls;      // description classes of this k-mer, in order of first appearance
      std::vector<uint64_t> run_of;   // the run whose positions the class shows (the last genome of the class)
      for (uint64_t i = a; i < b; ++i) {
        const uint64_t u = order[i];
        s += ind1; s.push_back('"');
        const uint64_t raw = k >= 1 ? unmix(keys[u], k) : 0;
        for (int j = 0; j < k; ++j) s.push_back(dec[(((raw >> (k + j)) & 1) << 1) | ((raw >> j) & 1)]);
        s += "\": {\n";
        cls.clear(); run_of.clear();
        for (uint64_t r = run_off[u]; r < run_off[u + 1]; ++r) {
          const uint32_t c = desc_class[run_genome[r]];
          size_t at = 0;
          while (at < cls.size() && cls[at] != c) ++at;
          if (at == cls.size()) { cls.push_back(c); run_of.push_back(r); } else run_of[at] = r;
        }
        for (size_t e = 0; e < cls.size(); ++e) {
          s += ind2;
          s.append(reinterpret_cast<const char*>(desc_json + desc_json_off[cls[e]]), (size_t)(desc_json_off[cls[e] + 1] - desc_json_off[cls[e]]));
          s += ": [\n";
          const uint64_t r = run_of[e];
          for (uint64_t q = pos_off[r]; q < pos_off[r + 1]; ++q) {
            s += ind3; append_uint(s, pos[q]);
            s += (q + 1 < pos_off[r + 1]) ? ",\n" : "\n";
          }
          s += ind2; s += (e + 1 < cls.size()) ? "],\n" : "]\n";
        }
        s += ind1; s += (i + 1 < n_keys) ? "},\n" : "}\n";
      }
    });
  } catch (const std::bad_alloc&) { return PA_ERR_NOMEM; }
  uint64_t total = n_keys ? 2 + ind0.size() + 1 : 2;
  for (const std::string& p : parts) total += p.size();
  uint8_t* out = static_cast<uint8_t*>(malloc(total ? total : 1));
  if (!out) return PA_ERR_NOMEM;
  uint64_t at = 0;
  if (n_keys == 0) {
    out[at++] = '{'; out[at++] = '}';
  } else {
    out[at++] = '{'; out[at++] = '\n';
    for (const std::string& p : parts) { memcpy(out + at, p.data(), p.size()); at += p.size(); }
    memcpy(out + at, ind0.data(), ind0.size()); at += ind0.size();
    out[at++] = '}';
  }
  *out_text = out; *out_len = at;
  return PA_OK;
}

int32_t pa_free_text(uint8_t* p) { free(p); return PA_OK; }

}  // extern "C"
