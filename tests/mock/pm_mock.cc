// TEST INFRASTRUCTURE ONLY -- a stand-in for libpm_b200.so that lets the host
// logic of the command line tools (option handling, primer input, pairing of
// hits into amplicons, output formats) run on a machine without a GPU.  It
// implements the entry points of include/pm_b200.h the tools call and serves
// canned data a test wrote into $PM_MOCK_DIR:
//   stream.bin  the normalised database stream (bytes)
//   seqs.bin    int64 n, then key[n], header offset[n], header length[n]
//   hits.bin    pm_hit records, in the order the index reports them
// Never built into or linked with the product; tests/test_cli_host_logic.py
// compiles it into tests/mock/_build/.
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "pm_b200.h"

struct pm_ctx {
  std::vector<uint8_t> stream;
  std::vector<int64_t> key, hoff, hlen;
  std::vector<pm_hit> hits;
  std::vector<uint8_t> raw;   // the database file (for the header lines)
  bool bogus = false;
  int k = 0;
};

static std::vector<uint8_t> slurp(const char *name) {
  const char *dir = getenv("PM_MOCK_DIR");
  std::string path = std::string(dir ? dir : ".") + "/" + name;
  std::vector<uint8_t> out;
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) { fprintf(stderr, "pm_mock: cannot read %s\n", path.c_str()); exit(2); }
  uint8_t buf[65536];
  size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) out.insert(out.end(), buf, buf + n);
  fclose(f);
  return out;
}

extern "C" {

const char *pm_last_error(void) { return "pm_mock"; }

int pm_create(pm_ctx **ctx, const pm_options *opt) {
  *ctx = new pm_ctx;
  (*ctx)->k = opt->k;
  return PM_OK;
}
void pm_destroy(pm_ctx *ctx) { delete ctx; }
int pm_add_pattern(pm_ctx *, const char *, uint32_t, int, int) { return PM_OK; }
int pm_init(pm_ctx *) { return PM_OK; }
int pm_reset(pm_ctx *) { return PM_OK; }
int pm_set_alignment_rules(pm_ctx *, int) { return PM_OK; }

int pm_set_alphabet_size(pm_ctx *, int) { return PM_OK; }
int pm_host_alloc(uint64_t bytes, int, void **out) { *out = malloc(bytes ? bytes : 1); return *out ? PM_OK : PM_E_NOMEM; }
int pm_host_free(void *p) { free(p); return PM_OK; }

static int load_canned(pm_ctx *c);
// whatever form the database arrives in, the canned stream / sequence table / hits are served
int pm_set_text_fasta_host(pm_ctx *c, const uint8_t *, uint64_t, int) { return load_canned(c); }
int pm_set_text_host(pm_ctx *c, const uint8_t *, uint64_t, int64_t) { return load_canned(c); }
int pm_set_text_sqz_host(pm_ctx *c, const uint8_t *, uint64_t, const uint8_t *, uint32_t, uint64_t, int64_t) { return load_canned(c); }
int pm_set_text_sqn_host(pm_ctx *c, const uint8_t *, uint64_t, const uint8_t *, uint32_t, int64_t) { return load_canned(c); }

int pm_set_text_database(pm_ctx *c, const char *database, int, int) {
  c->raw.clear();
  if (FILE *f = fopen(database, "rb")) {
    uint8_t buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) c->raw.insert(c->raw.end(), buf, buf + n);
    fclose(f);
  }
  return load_canned(c);
}
int pm_get_header(pm_ctx *c, uint64_t i, const char **text, uint64_t *len) {
  if (i >= c->key.size() || (uint64_t)(c->hoff[i] + c->hlen[i]) > c->raw.size()) return PM_E_ARG;
  *text = (const char *)c->raw.data() + c->hoff[i];
  *len = (uint64_t)(c->hlen[i] > 0 ? c->hlen[i] : 0);
  return PM_OK;
}

static int load_canned(pm_ctx *c) {
  c->stream = slurp("stream.bin");
  std::vector<uint8_t> s = slurp("seqs.bin");
  int64_t n = 0;
  memcpy(&n, s.data(), 8);
  c->key.resize(n); c->hoff.resize(n); c->hlen.resize(n);
  memcpy(c->key.data(), s.data() + 8, 8 * n);
  memcpy(c->hoff.data(), s.data() + 8 + 8 * n, 8 * n);
  memcpy(c->hlen.data(), s.data() + 8 + 16 * n, 8 * n);
  std::vector<uint8_t> h = slurp("hits.bin");
  c->hits.resize(h.size() / sizeof(pm_hit));
  memcpy(c->hits.data(), h.data(), c->hits.size() * sizeof(pm_hit));
  for (const pm_hit &x : c->hits) c->bogus |= (x.editdist < 0 || x.editdist > c->k);
  return PM_OK;
}

int pm_get_sequences(pm_ctx *c, const int64_t **key, const int64_t **hoff, const int64_t **hlen, uint64_t *ns,
                     uint64_t *len) {
  *key = c->key.data(); *hoff = c->hoff.data(); *hlen = c->hlen.data();
  *ns = c->key.size(); *len = c->stream.size();
  return PM_OK;
}

int pm_find_patterns(pm_ctx *c, int64_t, int64_t, const pm_hit **hits, uint64_t *nhits) {
  *hits = c->hits.data();
  *nhits = c->hits.size();
  return c->bogus ? PM_E_BOGUS : PM_OK;
}

int pm_get_text(pm_ctx *c, uint64_t begin, uint64_t len, uint8_t *out) {
  if (begin + len > c->stream.size()) return PM_E_ARG;
  memcpy(out, c->stream.data() + begin, len);
  return PM_OK;
}

int pm_gather_text(pm_ctx *c, const int64_t *begin, const int64_t *len, uint64_t n, uint8_t *out) {
  for (uint64_t i = 0; i < n; i++) {
    if (begin[i] < 0 || (uint64_t)(begin[i] + len[i]) > c->stream.size()) return PM_E_ARG;
    memcpy(out, c->stream.data() + begin[i], (size_t)len[i]);
    out += len[i];
  }
  return PM_OK;
}

int pm_normalize_fasta(const uint8_t *, uint64_t, int, unsigned char, uint8_t *, uint64_t *, int64_t *, int64_t *,
                       int64_t *, uint64_t, uint64_t *) {
  return PM_E_UNSUPPORTED;
}

// pure host functions of the product: forwarded to the real library ($PM_REAL_LIB), which
// loads without a GPU
static void *real_symbol(const char *name) {
  static void *h = 0;
  const char *lib = getenv("PM_REAL_LIB");
  if (!h && lib) h = dlopen(lib, RTLD_NOW | RTLD_LOCAL);
  void *f = h ? dlsym(h, name) : 0;
  if (!f) { fprintf(stderr, "pm_mock: cannot load %s from %s\n", name, lib ? lib : "(PM_REAL_LIB unset)"); exit(2); }
  return f;
}

int pm_reverse_comp(const char *in, char *out) {
  typedef int (*fn_t)(const char *, char *);
  static fn_t fn = (fn_t)real_symbol("pm_reverse_comp");
  return fn(in, out);
}

int pm_pcr_pairs(const pm_hit *hits, uint64_t nhits, uint32_t n_primers, const uint32_t *patlen, const int64_t *seq_key,
                 uint64_t n_seqs, const pm_pcr_options *opt, pm_pcr_pair **pairs, uint64_t *npairs) {
  typedef int (*fn_t)(const pm_hit *, uint64_t, uint32_t, const uint32_t *, const int64_t *, uint64_t,
                      const pm_pcr_options *, pm_pcr_pair **, uint64_t *);
  static fn_t fn = (fn_t)real_symbol("pm_pcr_pairs");
  return fn(hits, nhits, n_primers, patlen, seq_key, n_seqs, opt, pairs, npairs);
}

// the hits of hits.bin already come in the order the index reports them
int pm_last_candidates(pm_ctx *, const pm_candidate **cands, uint64_t *n) {
  *cands = 0;
  *n = 0;
  return PM_E_UNSUPPORTED;
}

int pm_hit_order(const pm_hit *hits, uint64_t nhits, const pm_candidate *cands, uint64_t ncands, int k, uint64_t *order) {
  typedef int (*fn_t)(const pm_hit *, uint64_t, const pm_candidate *, uint64_t, int, uint64_t *);
  static fn_t fn = (fn_t)real_symbol("pm_hit_order");
  return fn(hits, nhits, cands, ncands, k, order);
}

void pm_free(void *p) {
  typedef void (*fn_t)(void *);
  static fn_t fn = (fn_t)real_symbol("pm_free");
  fn(p);
}

}  // extern "C"
