// pm_pattern_match.hpp -- C++ host mirror of the reference's PatternMatch
// plugin surface (pattern_match.h:83-157, select.h:8-19) over the C ABI of
// include/pm_b200.h.  Same method names, argument meaning and error behaviour
// (fatal errors print the reference's message and exit(1)), so code written
// against the reference classes reads the same here.
//
// Header-only; link with -lpm_b200.
#ifndef PM_PATTERN_MATCH_HPP
#define PM_PATTERN_MATCH_HPP

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <utility>
#include <vector>

#include "pm_b200.h"

namespace pmb200 {

inline void fatal(const std::string &msg) {
  std::cerr << msg << std::endl;
  std::exit(1);
}

// The CharacterProducer of this library: a sequence database as pick_fasta_file opens it
// (select.t:23-190) -- the FASTA file itself (StreamedFastaFile, fasta_io.t:448-751) or the
// .seq / .sqn / .sqz files of compress_seq (IndexedFastaFile) -- with the header index FastaFile<H>
// keeps (fasta_io.t:140-215).  The file is read by the library's own reader (pm_set_text_database:
// a ring of page-locked chunk buffers feeding the copy engine) when an index adopts the database
// (b200_index::init); the stream is built on the device and only the sequence table comes back.
class FastaStream {
 public:
  FastaStream() : ctx_(0), stream_len_(0), dbind_(0), upper_case_(false), eos_('\n') {}
  // the database and how to open it (-D: 0 automatic, 1 FASTA, 2 .seq, 3 .sqn, 4 .sqz)
  bool load(const std::string &database, bool upper_case, unsigned char eos, int dbind = 0) {
    std::ifstream f(database.c_str(), std::ios::binary);
    std::ifstream g((database + ".idb").c_str(), std::ios::binary);
    if (!f && !g) return false;
    database_ = database;
    upper_case_ = upper_case;
    eos_ = eos;
    dbind_ = dbind;
    return true;
  }
  // CharacterProducer::size() as pick_pattern_index sees it (select.cc:120-123): 256 for the raw
  // FASTA file, the size of the character table for a .sqn / .sqz database
  int alphabet_size() const {
    int kind = dbind_;
    if (kind == 0) kind = exists(database_ + ".sqn") ? 3 : (exists(database_ + ".sqz") ? 4 : (exists(database_ + ".seq") ? 2 : 1));
    if (kind != 3 && kind != 4) return 256;
    std::ifstream t((database_ + (kind == 3 ? ".tbl" : ".tbz")).c_str(), std::ios::binary);
    if (!t) return 256;
    t.seekg(0, std::ios::end);
    const long n = (long)t.tellg();
    return (n >= 1 && n <= 256) ? (int)n : 256;
  }
  // hand the database to the device and keep its sequence table
  bool normalise_on_device(pm_ctx *ctx) {
    if (pm_set_text_database(ctx, database_.c_str(), dbind_, upper_case_) != PM_OK) return false;
    const int64_t *k = 0, *o = 0, *l = 0;
    uint64_t ns = 0, len = 0;
    if (pm_get_sequences(ctx, &k, &o, &l, &ns, &len) != PM_OK) return false;
    key_.assign(k, k + ns);
    stream_len_ = len;
    ctx_ = ctx;
    return true;
  }
  // the sequence table (stream position of each entry's first character), for pm_pcr_pairs
  const int64_t *seq_keys() const { return key_.data(); }
  uint64_t n_seqs() const { return key_.size(); }
  uint64_t size() const { return stream_len_; }
  // set_header_item (fasta_io.t:152-168): the entry whose key is the last <= pos-1
  long locate(int64_t pos) const {
    long lo = 0, hi = (long)key_.size();
    while (lo < hi) {
      const long mid = (lo + hi) / 2;
      if (key_[mid] <= pos - 1) lo = mid + 1; else hi = mid;
    }
    return lo - 1;
  }
  int64_t get_seq_pos(int64_t pos) const {  // fasta_io.t:194-202
    const long i = locate(pos);
    return i < 0 ? 0 : pos - key_[i];
  }
  // FastaFile::is_subseq (fasta_io.t:203-213): [start, end) lies inside one entry
  bool is_subseq(int64_t start, int64_t end) const {
    const long a = locate(start + 1);
    return a >= 0 && a == locate(end);
  }
  unsigned long index(int64_t pos) const { const long i = locate(pos); return i < 0 ? 0 : (unsigned long)(i + 1); }
  std::string header(int64_t pos) const {
    const long i = locate(pos);
    if (i < 0 || !ctx_) return "";
    const char *t = 0;
    uint64_t n = 0;
    if (pm_get_header(ctx_, (uint64_t)i, &t, &n) != PM_OK) fatal(std::string("Fatal error: ") + pm_last_error());
    return std::string(t, (size_t)n);
  }
  std::string short_header(int64_t pos) const {  // first word, fasta_io.t:118-126
    const std::string h = header(pos);
    const size_t p = h.find_first_of(" \t");
    return p == std::string::npos ? h : h.substr(0, p);
  }
  // the texts of a batch of hits in one round trip: text(a, b) below then serves them from memory
  void prefetch_texts(const std::vector<const pm_hit *> &batch) {
    pre_begin_.clear(); pre_off_.clear(); pre_.clear();
    if (!ctx_ || batch.empty()) return;
    std::vector<int64_t> len;
    int64_t total = 0;
    for (const pm_hit *h : batch) {
      int64_t a = h->start < 0 ? 0 : h->start, b = h->end > (int64_t)stream_len_ ? (int64_t)stream_len_ : h->end;
      if (b < a) b = a;
      pre_begin_.push_back(a);
      len.push_back(b - a);
      pre_off_.push_back(total);
      total += b - a;
    }
    pre_off_.push_back(total);
    pre_.resize((size_t)total);
    if (pm_gather_text(ctx_, pre_begin_.data(), len.data(), batch.size(), (uint8_t *)(total ? &pre_[0] : 0)) != PM_OK)
      fatal(std::string("Fatal error: ") + pm_last_error());
  }
  // characters [a, b) of the stream (the text primer_match prints next to an alignment)
  std::string text(int64_t a, int64_t b, long batch_index = -1) const {
    if (a < 0) a = 0;
    if (b > (int64_t)stream_len_) b = (int64_t)stream_len_;
    if (a >= b || !ctx_) return std::string();
    if (batch_index >= 0 && (size_t)batch_index + 1 < pre_off_.size() && pre_begin_[batch_index] == a &&
        pre_off_[batch_index + 1] - pre_off_[batch_index] == b - a)
      return pre_.substr((size_t)pre_off_[batch_index], (size_t)(b - a));
    std::string out((size_t)(b - a), '\0');
    if (pm_get_text(ctx_, (uint64_t)a, (uint64_t)(b - a), (uint8_t *)&out[0]) != PM_OK)
      fatal(std::string("Fatal error: ") + pm_last_error());
    return out;
  }

 private:
  static bool exists(const std::string &p) { std::ifstream f(p.c_str(), std::ios::binary); return (bool)f; }
  std::string database_;
  std::vector<int64_t> key_, pre_begin_, pre_off_;
  std::string pre_;
  pm_ctx *ctx_;
  uint64_t stream_len_;
  int dbind_;
  bool upper_case_;
  unsigned char eos_;
};

// pattern_hit_vector element (pattern_match.h:81): key = end position of the hit
struct pattern_hit {
  int64_t key;
  unsigned long id;
  unsigned char value;
};
typedef std::vector<pattern_hit> pattern_hit_vector;

class PatternMatch {
 public:
  PatternMatch() : next_pattern_id_(0) {}
  virtual ~PatternMatch() {}
  virtual unsigned long add_pattern(std::string const &pat, unsigned long id = 0, int exact_start_bases = 0,
                                    int exact_end_bases = 0) = 0;
  virtual void init(FastaStream &cp) = 0;
  virtual bool find_patterns(FastaStream &cp, pattern_hit_vector &pas, unsigned long minka = 1) = 0;
  virtual void reset() = 0;

 protected:
  unsigned long next_pattern_id_;
};

// One index on one GPU.  find_patterns() scans the whole stream on its first
// call and then hands the hits over `minka` at a time, like the reference's
// indexes do (e.g. filter_bitvec.cc:229-231).
class b200_index : public PatternMatch {
 public:
  b200_index(int index, unsigned k, char eos, bool wc, bool tn, bool indels, int device = 0)
      : ctx_(0), hits_(0), nhits_(0), next_(0), k_((int)k), scanned_(false), tolerate_bogus_(false) {
    pm_options o;
    std::memset(&o, 0, sizeof o);
    o.device = device; o.index = index; o.k = (int)k; o.indels = indels; o.wildcards = wc; o.text_n = tn;
    o.eos = (unsigned char)eos;
    if (pm_create(&ctx_, &o) != PM_OK) fatal(std::string("Fatal error: ") + pm_last_error());
  }
  ~b200_index() { pm_destroy(ctx_); }
  unsigned long add_pattern(std::string const &pat, unsigned long id = 0, int esb = 0, int eeb = 0) {
    if (id == 0) id = ++next_pattern_id_;
    if (pm_add_pattern(ctx_, pat.c_str(), (uint32_t)id, esb, eeb) != PM_OK) {
      // select.cc:88-91 prints this one when k >= pattern length
      fatal(std::string("Fatal error: ") + pm_last_error());
    }
    return id;
  }
  void init(FastaStream &cp) {
    if (pm_set_alphabet_size(ctx_, cp.alphabet_size()) != PM_OK) fatal(std::string("Fatal error: ") + pm_last_error());
    if (pm_init(ctx_) != PM_OK) fatal(std::string("Fatal error: ") + pm_last_error());
    if (!cp.normalise_on_device(ctx_)) fatal(std::string("Fatal error: ") + pm_last_error());
  }
  bool find_patterns(FastaStream &, pattern_hit_vector &pas, unsigned long minka = 1) {
    if (!scanned_) {
      const int rc = pm_find_patterns(ctx_, 0, INT64_MAX, &hits_, &nhits_);
      if (rc == PM_E_BOGUS && !tolerate_bogus_)
        fatal("Bogus hit returned to primer_match main()");  // primer_match.cc:1249
      if (rc != PM_OK && rc != PM_E_BOGUS) fatal(std::string("Fatal error: ") + pm_last_error());
      scanned_ = true;
      // the order filter_bitvec reports its hits in (hits that begin at one position: pm_hit_order);
      // primer_match prints them as they come, pcr_match sorts them by position alone
      const pm_candidate *cands = 0;
      uint64_t ncands = 0;
      last_candidates(&cands, &ncands);
      order_.assign(nhits_, 0);
      if (nhits_ && pm_hit_order(hits_, nhits_, cands, ncands, k_, order_.data()) != PM_OK)
        fatal(std::string("Fatal error: ") + pm_last_error());
    }
    while (next_ < nhits_ && pas.size() < minka) {
      const pm_hit &h = hits_[order_[next_++]];
      pattern_hit ph = {h.end, h.pattern_id, (unsigned char)h.value};
      pas.push_back(ph);
      batch_.push_back(&h);
    }
    return next_ < nhits_;
  }
  void reset() { pm_reset(ctx_); scanned_ = false; next_ = 0; batch_.clear(); }
  // A hit whose final alignment fails (editdist > k) ends primer_match; pcr_match only
  // skips the pairs it is part of (pcr_match.cc:1114): hand such hits out instead, the
  // caller tests pm_hit::editdist.
  void tolerate_bogus(bool yes) { tolerate_bogus_ = yes; }
  // the raw candidates behind the hits of the last find_patterns (none for exact_halves / exact_bases)
  void last_candidates(const pm_candidate **cands, uint64_t *n) {
    *cands = 0;
    *n = 0;
    if (pm_last_candidates(ctx_, cands, n) != PM_OK) { *cands = 0; *n = 0; }
  }
  // PM_ALIGN_PCR_MATCH: the alignment pcr_match makes of a hit (see pm_set_alignment_rules)
  void alignment_rules(int rules) {
    if (pm_set_alignment_rules(ctx_, rules) != PM_OK) fatal(std::string("Fatal error: ") + pm_last_error());
  }
  // the full records (alignment included) of the hits handed out by the last
  // find_patterns() call, same order as `pas`
  std::vector<const pm_hit *> &batch() { return batch_; }
  pm_ctx *ctx() { return ctx_; }

 private:
  pm_ctx *ctx_;
  const pm_hit *hits_;
  uint64_t nhits_, next_;
  int k_;
  bool scanned_, tolerate_bogus_;
  std::vector<uint64_t> order_;
  std::vector<const pm_hit *> batch_;
};

// pick_pattern_index (select.cc:18-298) for the selections this library has.
// pmselect 0 = automatic.  Selections whose *hit semantics* differ from what is
// implemented (hash tables 6/15, suffix
// tree) are refused rather than silently replaced.
inline b200_index *pick_pattern_index(int pmselect, int nmismatch, bool wildcard, bool textn, bool indels,
                                      char eos, int device = 0) {
  int index;
  switch (pmselect) {
    case 0: index = PM_INDEX_AUTO; break;
    case 1: case 2: case 3: index = PM_INDEX_KEYWORD_TREE; break;
    case 4: index = PM_INDEX_SHIFT_AND; break;
    case 5: index = PM_INDEX_FILTER_BITVEC; break;
    case 7: case 8: case 9: case 10: index = PM_INDEX_EXACT_BASES; break;
    case 11: case 12: case 13: case 14: index = PM_INDEX_EXACT_HALVES; break;
    default:
      fatal("Fatal error: pattern index selection not available in the B200 build (use -N 0, 3, 4, 5 or 7-14)");
      return 0;
  }
  if (index == PM_INDEX_KEYWORD_TREE && wildcard) index = PM_INDEX_SHIFT_AND;
  return new b200_index(index, (unsigned)nmismatch, eos, wildcard, textn, indels, device);
}

}  // namespace pmb200
#endif
