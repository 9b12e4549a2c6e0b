// CPU check: the banded alignment (editdist_align_band) against the full-matrix one
// (editdist_align) of csrc/pm_align.cuh on random inputs.  Test infrastructure:
//   g++ -O2 -std=c++17 -o /tmp/align_band_check tools/align_band_check.cc && /tmp/align_band_check
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <random>
#include <vector>
#include "../sequence-alignment-tools_b200/csrc/pm_align.cuh"

using namespace pm;

int main(int argc, char **argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 200000;
  std::mt19937_64 g(12345);
  // iupac-ish compat: a few wildcard rows
  std::vector<uint32_t> compat(128 * 4, 0);
  auto setc = [&](int r, const char *cols) { for (const char *p = cols; *p; ++p) compat[r * 4 + (*p >> 5)] |= 1u << (*p & 31); };
  setc('N', "ACGTN"); setc('R', "AGRN"); setc('Y', "CTYN"); setc('A', "ARN"); setc('C', "CYN"); setc('G', "GRN"); setc('T', "TYN");
  long checked = 0, okc = 0;
  for (int it = 0; it < iters; it++) {
    const int k = (int)(g() % 4);
    const int indels = (int)(g() % 2);
    const int wc = (g() % 4) == 0;
    const int tn = (int)(g() % 2);
    const int m = k + 1 + (int)(g() % (64 - k));
    const int n = 80 + (int)(g() % 200);
    const char *alpha = (g() % 3 == 0) ? "AC" : "ACGT";
    const int na = (int)strlen(alpha);
    std::vector<uint8_t> text(n), pat(PAT_STRIDE * 2, 0);
    for (auto &c : text) c = (uint8_t)alpha[g() % na];
    text[0] = '\n'; text[n - 1] = '\n';
    if (g() % 3 == 0) text[1 + g() % (n - 2)] = '\n';
    if (g() % 3 == 0) text[1 + g() % (n - 2)] = 'N';
    if (wc && g() % 2) text[1 + g() % (n - 2)] = 'R';
    // pattern: a mutated copy of some text window (so that alignments succeed often)
    const int at = 1 + (int)(g() % (n - 2));
    for (int j = 0; j < m; j++) pat[PAT_STRIDE + j] = (at + j < n - 1 && text[at + j] != '\n') ? text[at + j] : (uint8_t)alpha[g() % na];
    for (int e = 0; e < (int)(g() % (k + 2)); e++) pat[PAT_STRIDE + g() % m] = (uint8_t)alpha[g() % na];
    if (wc && g() % 2) pat[PAT_STRIDE + g() % m] = "NRY"[g() % 3];
    uint8_t plen[2] = {0, (uint8_t)m};
    int16_t lc[2] = {0, 0}, rc[2] = {0, 0};
    if (g() % 4 == 0) lc[1] = (int16_t)(g() % (m - k));
    if (g() % 4 == 0) rc[1] = (int16_t)(g() % (m - k));
    AlignEnv env;
    env.text = text.data(); env.n = n; env.pat_chars = pat.data(); env.plen = plen; env.lconst = lc; env.rconst = rc;
    env.compat = compat.data(); env.k = k; env.indels = indels; env.wc = wc; env.tn = tn; env.eos = '\n';
    // chain [e1, e2]: ends near the planted copy
    int64_t e1 = at + m + (int64_t)(g() % 7) - 3;
    if (e1 < 0) e1 = 0;
    if (e1 > n) e1 = n;
    int64_t e2 = e1 + ((g() % 3) ? (int64_t)(g() % (2 * k + 2)) : (int64_t)(g() % 40));
    if (e2 > n) e2 = n;
    const int64_t buflen = e2 - align_textstart(e1, m, k);
    std::vector<uint8_t> ws(align_ws_bytes(m, buflen < 0 ? 0 : buflen) + 64);
    const int64_t wb = align_band_width(e1, e2, k, indels);
    std::vector<uint8_t> best((size_t)(m + 1) * wb, 0xEE);
    std::vector<int> ra(wb + 2, 0x7777777), rb(wb + 2, 0x7777777);
    uint8_t c1[80], c2[80];
    memset(c1, 0, sizeof c1); memset(c2, 0, sizeof c2);
    AlignResult r1, r2;
    editdist_align(env, 1, e1, e2, ws.data(), c1, 72, r1);
    editdist_align_band(env, 1, e1, e2, best.data(), ra.data(), rb.data(), c2, 72, r2);
    checked++;
    if (r1.ok == 1) okc++;
    if (r1.ok != r2.ok || r1.ncodes != r2.ncodes || r1.start != r2.start || r1.end != r2.end || r1.val != r2.val ||
        r1.editdist != r2.editdist || memcmp(c1, c2, 72) != 0) {
      printf("MISMATCH it=%d k=%d indels=%d m=%d e1=%ld e2=%ld ok %d/%d end %ld/%ld val %d/%d ed %d/%d nc %d/%d\n", it, k, indels, m,
             (long)e1, (long)e2, r1.ok, r2.ok, (long)r1.end, (long)r2.end, r1.val, r2.val, r1.editdist, r2.editdist, r1.ncodes, r2.ncodes);
      return 1;
    }
  }
  printf("align_band_check: %ld alignments identical (%ld with ok == 1)\n", checked, okc);
  return 0;
}
