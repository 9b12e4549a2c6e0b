// b200_match.h -- the PatternMatch a maintainer adds to the reference tree so that its OWN
// primer_match.cc / pcr_match.cc (unchanged) run their scan on libpm_b200.so.  Compiled in
// tests/adapter/ against the reference's headers (pattern_match.h:83-157) and linked with the
// reference's own objects; tests/test_adapter.py compares the text the reference's main()
// prints through it with the text the unmodified reference printed.
//
// The database goes to the device as it lies on disk: the CharacterProducer names its file
// (char_io.h:54-56), <db>.sqz / <db>.sqn / <db>.seq / the FASTA file itself, which is read into
// page-locked memory and handed to pm_set_text_*_host -- no getch() per base.  A prefix of the
// device's stream is then compared with what the CharacterProducer itself delivers (this also
// finds out whether the file is read with -u); a producer the library cannot mirror (character
// mapping, translation) is streamed through getch() instead.
#ifndef B200_MATCH_H
#define B200_MATCH_H

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "pattern_match.h"
#include "pm_b200.h"

class b200_match : public PatternMatch {
  pm_ctx *ctx_;
  std::vector<pattern_list::const_iterator> plit_;
  const pm_hit *hits_;
  uint64_t nhits_, next_;
  std::vector<uint64_t> order_;   // the hits in the order filter_bitvec would report them (pm_hit_order)
  int k_;
  bool loaded_, scanned_;
  unsigned char eos_;
  void *host_;       // page-locked copy of the database file

  void die(const char *what) {
    std::string msg = std::string("b200_match: ") + what + ": " + pm_last_error();
    timestamp(msg.c_str());
    exit(1);
  }
  static bool read_file(const std::string &path, void **pinned, uint64_t *size) {
    int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return false; }
    void *p = 0;
    if (pm_host_alloc((uint64_t)st.st_size, 1, &p) != PM_OK) { close(fd); return false; }
    uint64_t done = 0;
    while (done < (uint64_t)st.st_size) {
      ssize_t r = read(fd, (char *)p + done, (size_t)std::min<uint64_t>((uint64_t)st.st_size - done, 64u << 20));
      if (r <= 0) { close(fd); pm_host_free(p); return false; }
      done += (uint64_t)r;
    }
    close(fd);
    *pinned = p;
    *size = done;
    return true;
  }
  static std::string with_extn(const std::string &fn, const char *ext) {
    const size_t dot = fn.rfind('.');
    return (dot == std::string::npos ? fn : fn.substr(0, dot)) + "." + ext;
  }
  // The first characters the producer delivers, read ONCE when the index is initialised.  The
  // producer is never rewound: a StreamedFastaFile cannot go back to position 0 after its first
  // lookup (fasta_io.t:700-713), and the callers restore `ff->pos()` as they found it after
  // find_patterns (primer_match.cc:1121, 1262), which therefore must not be 0 either.
  std::string probe_;
  void read_probe(CharacterProducer &cp) {
    probe_.clear();
    while (probe_.size() < 16384 && !cp.eof()) probe_ += cp.getch();
  }
  // does the device hold what this producer delivers?
  bool prefix_agrees(bool *case_only) {
    const int64_t *key, *ho, *hl;
    uint64_t nseq = 0, n = 0;
    if (pm_get_sequences(ctx_, &key, &ho, &hl, &nseq, &n) != PM_OK) return false;
    // (a producer may report its end before it has delivered the closing end-of-sequence character)
    if (n < probe_.size() || (probe_.size() < 16384 && n > probe_.size() + 1)) return false;
    const uint64_t take = probe_.size();
    std::vector<uint8_t> dev(take + 1);
    if (take && pm_get_text(ctx_, 0, take, dev.data()) != PM_OK) return false;
    bool same = true, same_upper = true;
    for (uint64_t i = 0; i < take; i++) {
      const unsigned char c = (unsigned char)probe_[i];
      if (c != dev[i]) same = false;
      if (toupper(c) != toupper(dev[i])) same_upper = false;
    }
    if (case_only) *case_only = !same && same_upper;
    return same;
  }
  void load(CharacterProducer &cp) {
    bool ok = false;
    read_probe(cp);
    if (cp.has_filename()) {
      const std::string fn = cp.filename();
      const std::string ext = fn.size() > 4 ? fn.substr(fn.size() - 4) : "";
      uint64_t size = 0;
      if (read_file(fn, &host_, &size)) {
        if (ext == ".sqz" || ext == ".sqn") {
          void *tab = 0;
          uint64_t tsize = 0;
          if (read_file(with_extn(fn, ext == ".sqz" ? "tbz" : "tbl"), &tab, &tsize) && tsize >= 1 && tsize <= 256) {
            ok = (ext == ".sqz" ? pm_set_text_sqz_host(ctx_, (const uint8_t *)host_, size, (const uint8_t *)tab,
                                                        (uint32_t)tsize, 0, 0)
                                : pm_set_text_sqn_host(ctx_, (const uint8_t *)host_, size, (const uint8_t *)tab,
                                                        (uint32_t)tsize, 0)) == PM_OK;
            pm_host_free(tab);
          }
        } else if (ext == ".seq") {
          ok = pm_set_text_host(ctx_, (const uint8_t *)host_, size, 0) == PM_OK;
        } else {
          ok = pm_set_text_fasta_host(ctx_, (const uint8_t *)host_, size, 0) == PM_OK;
          bool case_only = false;
          if (ok && !prefix_agrees(&case_only) && case_only)     // primer_match -u
            ok = pm_set_text_fasta_host(ctx_, (const uint8_t *)host_, size, 1) == PM_OK;
        }
        if (ok) ok = prefix_agrees(0);
      }
    }
    if (getenv("B200_MATCH_TRACE"))
      fprintf(stderr, "b200_match: database %s\n", ok ? "handed over as it lies on disk" : "streamed through getch()");
    if (!ok) {
      // a producer that cannot be mirrored from its file: its characters, one by one
      std::string s = probe_;
      while (!cp.eof()) s += cp.getch();
      if (pm_set_text_host(ctx_, (const uint8_t *)s.data(), s.size(), 0)) die("pm_set_text_host");
    }
    loaded_ = true;
  }

 public:
  b200_match(int index, unsigned k, char eos, bool wc, bool tn, bool indels, int alphabet_size, bool pcr_rules)
      : hits_(0), nhits_(0), next_(0), k_((int)k), loaded_(false), scanned_(false), eos_((unsigned char)eos), host_(0) {
    pm_options o = {0, index, (int)k, indels ? 1 : 0, wc ? 1 : 0, tn ? 1 : 0, (unsigned char)eos};
    if (const char *d = getenv("PM_B200_DEVICE")) o.device = atoi(d);
    if (pm_create(&ctx_, &o)) die("pm_create");
    if (pm_set_alphabet_size(ctx_, alphabet_size)) die("pm_set_alphabet_size");
    if (pcr_rules && pm_set_alignment_rules(ctx_, PM_ALIGN_PCR_MATCH)) die("pm_set_alignment_rules");
  }
  ~b200_match() {
    pm_destroy(ctx_);
    if (host_) pm_host_free(host_);
  }
  long unsigned int add_pattern(std::string const &pat, unsigned long id = 0, int esb = 0, int eeb = 0) {
    pattern_list::const_iterator it = add_pattern_(pat, id, esb, eeb);
    if (plit_.size() <= id) plit_.resize(id + 1);
    plit_[id] = it;
    if (pm_add_pattern(ctx_, pat.c_str(), (uint32_t)id, esb, eeb)) die("pm_add_pattern");
    return id;
  }
  void init(CharacterProducer &cp) {
    if (pm_init(ctx_)) die("pm_init");
    load(cp);
  }
  // the hits of the whole database, handed over in report_interval batches: key = the position
  // the reference's own final alignment (primer_match.cc:1132-1153, pcr_match.cc:1096-1113)
  // restarts from
  bool find_patterns(CharacterProducer &cp, pattern_hit_vector &phs, long unsigned minka = 1) {
    if (!loaded_) load(cp);
    if (!scanned_) {
      const int rc = pm_find_patterns(ctx_, 0, INT64_MAX, &hits_, &nhits_);
      if (rc != PM_OK && rc != PM_E_BOGUS) die("pm_find_patterns");   // a bogus hit is the caller's to judge
      scanned_ = true;
      next_ = 0;
      // hits that end at one position: the callers sort by position alone (std::sort), so the
      // order they are handed over in is part of the answer (pcr_match pairs a hit only with
      // partners that come later)
      const pm_candidate *cands = 0;
      uint64_t ncands = 0;
      if (pm_last_candidates(ctx_, &cands, &ncands) != PM_OK) { cands = 0; ncands = 0; }
      order_.assign(nhits_, 0);
      if (nhits_ && pm_hit_order(hits_, nhits_, cands, ncands, k_, order_.data())) die("pm_hit_order");
    }
    int64_t frontier = 0;
    while (next_ < nhits_ && phs.size() < minka) {
      const pm_hit &h = hits_[order_[next_++]];
      phs.push_back((FILE_POSITION_TYPE)h.end, std::make_pair(plit_[h.pattern_id], (unsigned char)h.value));
      if (h.end > frontier) frontier = h.end;
    }
    // the reference's indexes leave the producer at their scan frontier, behind every hit they
    // report: a StreamedFastaFile only knows the sequences it has read past (get_seq_pos /
    // get_header_data, fasta_io.t:163-215), so it is moved there before the caller looks them up
    if (frontier > (int64_t)cp.pos()) cp.pos((FILE_POSITION_TYPE)frontier);
    return next_ < nhits_;
  }
  void reset() {
    pm_reset(ctx_);
    loaded_ = scanned_ = false;
    next_ = nhits_ = 0;
  }
};

#endif
