// CPU check of csrc/pm_lv.cuh (seed verification by diagonal extension) against the oracle's
// restatement of the reference recurrence (oracle/pm_oracle.c: pmo_wm_scan, indels on).
// For random texts with end-of-sequence and N characters and random / planted patterns:
//   ref   = candidates (pos, pattern) of pmo_wm_scan
//   lv    = union over every PAIR seed (two pieces of the pattern's k+2 that occur exactly at a
//           distance some edit script allows -- what the pair index proposes) of the ends
//           lv_verify reports for the pair's first piece as anchor
// and the two sets must be equal.  Test infrastructure:
//   g++ -O2 -std=c++17 -o /tmp/lv_check tools/lv_check.cc oracle/pm_oracle.c && /tmp/lv_check
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <random>
#include <set>
#include <string>
#include <vector>
#include "../sequence-alignment-tools_b200/csrc/pm_lv.cuh"

extern "C" {
typedef struct { int64_t pos; int32_t pid; int32_t val; } pmo_hit;
long pmo_wm_scan(const uint8_t *T, long L, int npat, const char *const *pats, int k, int indels, int wc, int tn,
                 uint8_t eos, pmo_hit *out, long cap);
int pmo_scan_match(unsigned pc, unsigned tc, int wc, int tn);
}

using namespace pm;

static int code_of(uint8_t ch) { return (ch >> 1) & 3; }   // A 0, C 1, T 2, G 3

template <int K>
static uint32_t run_lv(const LvCtx &t, int a, int L) { return lv_verify<K>(t, a, L); }

int main(int argc, char **argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 3000;
  std::mt19937_64 g(argc > 2 ? atoll(argv[2]) : 777);
  long n_ref = 0, n_seeds = 0;
  for (int it = 0; it < iters; it++) {
    const int K = 1 + (int)(g() % 3);
    const int P = K + 2;
    int m = 4 * P + (int)(g() % (33 - 4 * P));      // 4P .. 32
    int L = m / P; if (L > 8) L = 8;
    if (L > (32 - K) / P) L = (32 - K) / P;
    const int n = 300 + (int)(g() % 400);
    const bool small_alpha = (g() % 3) == 0;
    const char *alpha = small_alpha ? "AC" : "ACGT";
    const int na = (int)strlen(alpha);
    std::vector<uint8_t> T(n);
    for (auto &c : T) c = (uint8_t)alpha[g() % na];
    T[0] = '\n'; T[n - 1] = '\n';
    for (int r = 0; r < 4; r++) if (g() % 2) T[1 + g() % (n - 2)] = '\n';
    for (int r = 0; r < 3; r++) if (g() % 2) T[1 + g() % (n - 2)] = 'N';
    // patterns: random, and mutated copies of text windows (next to sequence ends too)
    const int npat = 3;
    std::vector<std::string> pats;
    const int wc = (int)(g() % 3 == 0);     // IUPAC pattern characters, scanned with -w
    for (int p = 0; p < npat; p++) {
      std::string s(m, 'A');
      for (auto &c : s) c = alpha[g() % na];
      if (wc)
        for (int r = 0; r < 1 + (int)(g() % 3); r++) s[g() % m] = "RYKMSWBDHVN"[g() % 11];
      pats.push_back(s);
    }
    auto pmatch = [&](char pc, uint8_t tc) { return pmo_scan_match((unsigned char)pc, tc, wc, 0) != 0; };
    for (int r = 0; r < 12; r++) {
      std::string q = pats[g() % npat];
      for (auto &c : q) if (!strchr("ACGT", c)) { const char *opt = "ACGT"; char pc = c; do { c = opt[g() % 4]; } while (!pmatch(pc, (uint8_t)c)); }
      const int ne = (int)(g() % (K + 2));
      for (int e = 0; e < ne && !q.empty(); e++) {
        const int op = (int)(g() % 3);
        const size_t at = g() % q.size();
        if (op == 0) q[at] = alpha[g() % na];
        else if (op == 1) q.insert(q.begin() + at, alpha[g() % na]);
        else q.erase(q.begin() + at);
      }
      // placed anywhere; a third of them flush against an end of sequence
      long at = 1 + (long)(g() % (n - 2));
      if (g() % 3 == 0) {
        std::vector<long> eos;
        for (long i = 0; i < n; i++) if (T[i] == '\n') eos.push_back(i);
        const long w = eos[g() % eos.size()];
        at = (g() % 2) ? w + 1 : w - (long)q.size();
      }
      for (size_t i = 0; i < q.size(); i++)
        if (at + (long)i > 0 && at + (long)i < n - 1 && T[at + i] != '\n') T[at + i] = (uint8_t)q[i];
    }
    // reference candidates
    std::vector<const char *> pp;
    for (auto &s : pats) pp.push_back(s.c_str());
    std::vector<pmo_hit> out(1 << 16);
    const long nr = pmo_wm_scan(T.data(), n, npat, pp.data(), K, 1, wc, 0, '\n', out.data(), (long)out.size());
    std::set<std::pair<long, int>> ref, lv;
    for (long i = 0; i < nr; i++) ref.insert({(long)out[i].pos, out[i].pid});
    n_ref += (long)ref.size();
    // packed text: codes / flags per character
    auto code_at = [&](long i) -> uint32_t { return (i >= 0 && i < n && T[i] != '\n' && T[i] != 'N') ? (uint32_t)code_of(T[i]) : (uint32_t)(i & 3); };
    auto flag_at = [&](long i) -> uint32_t {
      if (i < 0 || i >= n || T[i] == '\n') return 3u;
      if (T[i] == 'N') return 1u;
      return 0u;
    };
    for (int p = 0; p < npat; p++) {
      LvCtx t;
      uint64_t acc[4] = {0, 0, 0, 0};
      for (int j = 0; j < m; j++)
        for (const char *tc = "ACGT"; *tc; ++tc)
          if (pmatch(pats[p][j], (uint8_t)*tc)) acc[code_of((uint8_t)*tc)] |= 1ULL << (2 * j);
      lv_set_pattern(t, m, acc[0], acc[1], acc[2], acc[3], wc ? 0 : 1);   // what the host passes: plain patterns without -w
      // pair seeds: pieces pi < pj exact at text distance (pj - pi) L + e, |e| <= K, never < L
      for (long pos = 1; pos < n; pos++) {
        for (int pi = 0; pi < P; pi++) {
          // the index keys on the 2-bit codes alone: a flagged character (N, end of sequence) carries
          // an arbitrary code and can complete a piece by accident -- such a seed must verify to nothing
          bool okA = true;
          for (int x = 0; x < L && okA; x++)
            okA = pos + x < n && pmatch(pats[p][pi * L + x], (uint8_t)"ACTG"[code_at(pos + x)]);
          if (!okA) continue;
          bool paired = false;
          for (int pj = pi + 1; pj < P && !paired; pj++)
            for (int e = -K; e <= K && !paired; e++) {
              const int d = (pj - pi) * L + e;
              if (d < L) continue;
              bool okB = true;
              for (int x = 0; x < L && okB; x++)
                okB = pos + d + x < n && pmatch(pats[p][pj * L + x], (uint8_t)"ACTG"[code_at(pos + d + x)]);
              paired = okB;
            }
          if (!paired) continue;
          n_seeds++;
          const long s_nom = pos - pi * L;
          const long s0 = s_nom - K;
          for (int w = 0; w < 3; w++) {
            uint32_t c = 0, f = 0;
            for (int x = 0; x < 16; x++) {
              c |= code_at(s0 + 16 * w + x) << (2 * x);
              f |= flag_at(s0 + 16 * w + x) << (2 * x);
            }
            t.c[w] = c; t.f[w] = f;
          }
          const uint32_t ends = K == 1 ? run_lv<1>(t, pi * L, L) : (K == 2 ? run_lv<2>(t, pi * L, L) : run_lv<3>(t, pi * L, L));
          for (int d = -K; d <= K; d++)
            if ((ends >> (d + K)) & 1u) lv.insert({s_nom + m + d, p + 1});
        }
      }
    }
    if (ref != lv) {
      printf("MISMATCH it=%d K=%d m=%d L=%d n=%d: ref %zu lv %zu\n", it, K, m, L, n, ref.size(), lv.size());
      for (auto &x : ref) if (!lv.count(x)) printf("  only ref: pos %ld pat %d\n", x.first, x.second);
      for (auto &x : lv) if (!ref.count(x)) printf("  only lv : pos %ld pat %d\n", x.first, x.second);
      for (int p = 0; p < npat; p++) printf("  pat %d %s\n", p + 1, pats[p].c_str());
      std::string s(T.begin(), T.end());
      for (auto &c : s) if (c == '\n') c = '|';
      printf("  text %s\n", s.c_str());
      return 1;
    }
  }
  printf("lv_check: %d texts, %ld reference candidates reproduced from %ld pair seeds\n", iters, n_ref, n_seeds);
  return 0;
}
