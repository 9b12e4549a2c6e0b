// pm_cli_common.hpp -- what the two command line tools (primer_match_main.cc,
// pcr_match_main.cc) share: option helpers, pattern input, the per-pattern
// constraint rules and the rendering of an alignment record into the strings
// the reference's alignformat() functions print.
#ifndef PM_CLI_COMMON_HPP
#define PM_CLI_COMMON_HPP

#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "pm_pattern_match.hpp"

namespace pmcli {

inline int tilde_int(const char *a) { return a[0] == '~' ? -std::atoi(a + 1) : std::atoi(a); }

// exact_start_bases / exact_end_bases of a pattern and of its reverse complement from the
// -s / -e / -5 / -3 options (a negative value is the `~n` form: all but n characters),
// primer_match.cc:990-1060 = pcr_match.cc:835-909; for a reverse complement 5' and 3' swap ends
inline void pattern_constraints(int len, int stlen, int edlen, int fplen, int tplen, std::pair<int, int> &fwd,
                                std::pair<int, int> &rev) {
  int first = stlen > 0 ? stlen : 0;
  if (fplen > first) first = fplen;
  if (edlen < 0 && len + edlen > first) first = len + edlen;
  if (tplen < 0 && len + tplen > first) first = len + tplen;
  int second = edlen > 0 ? edlen : 0;
  if (tplen > second) second = tplen;
  if (stlen < 0 && len + stlen > second) second = len + stlen;
  if (fplen < 0 && len + fplen > second) second = len + fplen;
  fwd = std::make_pair(first, second);
  int f2 = stlen > 0 ? stlen : 0;
  if (tplen > f2) f2 = tplen;
  if (edlen < 0 && len + edlen > f2) f2 = len + edlen;
  if (fplen < 0 && len + fplen > f2) f2 = len + fplen;
  int s2 = edlen > 0 ? edlen : 0;
  if (fplen > s2) s2 = fplen;
  if (stlen < 0 && len + stlen > s2) s2 = len + stlen;
  if (tplen < 0 && len + tplen > s2) s2 = len + tplen;
  rev = std::make_pair(f2, s2);
}

inline std::string reverse_comp(const std::string &p) {
  std::vector<char> buf(p.size() + 1);
  pm_reverse_comp(p.c_str(), buf.data());
  return std::string(buf.data());
}

// pattern_alignment::alignment_string / alignment_text / alignment_pattern
// (pattern_alignment.h:124-197)
inline char code_char(uint8_t c) {
  switch (c) {
    case 1: return '|';
    case 2: return '+';
    case 3: return '*';
    case 4: return '^';
    case 5: return 'v';
    case 6: return '!';
    default: return ' ';
  }
}

struct Aligned {
  std::string A, T, Q, t;
};

inline Aligned render(const pm_hit &h, const std::string &pat, const std::string &matching) {
  Aligned a;
  a.t = matching;
  size_t tp = 0, pp = 0;
  for (uint32_t i = 0; i < h.ncodes; i++) {
    const uint8_t c = h.codes[i];
    a.A += code_char(c);
    if (c != 5) a.T += (tp < matching.size() ? matching[tp++] : ' '); else a.T += '-';
    if (c != 4) a.Q += (pp < pat.size() ? pat[pp++] : ' '); else a.Q += '-';
  }
  return a;
}

// backslash escapes of the format strings (primer_match.cc:729-745, pcr_match.cc:727-743)
inline void escape(std::ostream &os, const std::string &f, size_t &pos) {
  pos++;
  if (pos < f.size()) {
    switch (f[pos]) {
      case 'n': os << '\n'; break;
      case 't': os << '\t'; break;
      case '\\': os << '\\'; break;
      default: os << f[pos];
    }
  } else {
    os << '\\';
  }
}

inline void read_words(std::istream &is, std::vector<std::string> &out) {
  std::string w;
  while (is >> w) out.push_back(w);
}

// fasta_entry input (sts_io / fasta_io operator>>): defline after '>', sequence = joined lines
inline void read_fasta_patterns(std::istream &is, std::vector<std::string> &pats, std::vector<std::string> &defs) {
  std::string line, seq, def;
  bool have = false;
  while (std::getline(is, line)) {
    if (!line.empty() && line[line.size() - 1] == '\r') line.erase(line.size() - 1);
    if (!line.empty() && line[0] == '>') {
      if (have && !seq.empty()) { pats.push_back(seq); defs.push_back(def); }
      def = line.substr(1);
      seq.clear();
      have = true;
    } else if (have) {
      for (char ch : line)
        if (ch != ' ' && ch != '\t') seq += ch;
    }
  }
  if (have && !seq.empty()) { pats.push_back(seq); defs.push_back(def); }
}

// sts_entry (sts_io.h:11-37, sts_io.cc:12-47): one line of a UniSTS file.  One object is
// read into over and over, as the reference does: fields a short line does not have keep
// the previous line's values.
struct StsEntry {
  std::string id, forward_primer, reverse_primer, acc, chrom, altacc, species;
  unsigned long sizelb = 0, sizeub = 0;
  bool read(std::istream &is) {
    std::string line;
    if (!std::getline(is, line)) return false;
    std::istringstream iss(line);
    std::string size;
    iss >> id >> forward_primer >> reverse_primer >> size >> acc >> chrom >> altacc;
    const std::string::size_type p = size.find("-");
    if (p != std::string::npos) {
      sizelb = (unsigned long)std::atoi(size.substr(0, p).c_str());
      sizeub = (unsigned long)std::atoi(size.substr(p + 1).c_str());
    } else {
      sizelb = sizeub = (unsigned long)std::atoi(size.c_str());
    }
    std::string rest;
    std::getline(iss, rest);
    species = rest;
    return true;
  }
};

// the codes both tools print from an sts_entry (primer_match.cc:500-521 / 791-812, pcr_match.cc:498-549);
// false when `code` is not one of them
inline bool sts_code(std::ostream &os, char code, const StsEntry &sts) {
  switch (code) {
    case 'I': os << sts.id; return true;
    case 'a': os << sts.acc; return true;
    case 'O': os << sts.species; return true;
    case '&': os << sts.altacc; return true;
    case 'X': os << sts.chrom; return true;
    default: return false;
  }
}

}  // namespace pmcli
#endif
