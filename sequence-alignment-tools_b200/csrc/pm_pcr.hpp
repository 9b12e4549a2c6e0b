// pm_pcr.hpp -- the pairing rule of pcr_match (pcr_match.cc:945-1170) on the hit
// records pm_find_patterns returns: host code, no device work.
//
// The reference pairs batch by batch (-R hits per find_patterns call) and lets a
// hit near the end of a batch wait for the next one; with every hit at hand the
// rule is one pass in position order, which is what the reference does on any
// database with fewer hits than -R.  (The two differ only when one primer site
// lies inside its partner's site within -M of a batch end.)
#ifndef PM_PCR_HPP
#define PM_PCR_HPP

#include <stdint.h>

#include <algorithm>
#include <map>
#include <vector>

#include "../../include/pm_b200.h"

namespace pm {

// FastaFile::set_header_item (fasta_io.t:163-190): the entry whose key is the last <= pos - 1
inline long pcr_locate(const int64_t *seq_key, uint64_t n_seqs, int64_t pos) {
  long lo = 0, hi = (long)n_seqs;
  while (lo < hi) {
    const long mid = (lo + hi) / 2;
    if (seq_key[mid] <= pos - 1) lo = mid + 1; else hi = mid;
  }
  return lo - 1;
}

// element of the reference's sortedvector: compared by its key alone (sortedvector.t:38-44), so
// std::sort leaves equal keys in whatever order introsort leaves them -- the same order here,
// because it is the same std::sort on the same sequence of keys
struct KeyIdx {
  long key;
  uint64_t idx;
  bool operator<(const KeyIdx &o) const { return key < o.key; }
};

// filter_bitvec::find_patterns' order of reporting (filter_bitvec.cc:89-141), see pm_hit_order
inline void hit_order(const pm_hit *hits, uint64_t nhits, const pm_candidate *cands, uint64_t ncands, int k,
                      std::vector<uint64_t> &order) {
  order.clear();
  order.reserve(nhits);
  std::vector<bool> placed(nhits, false);
  if (cands && ncands && k > 0) {
    std::vector<std::pair<long, uint32_t> > emitted(ncands);   // (position, pattern): shift_and_inexact.cc:315-333
    for (uint64_t i = 0; i < ncands; i++) emitted[i] = std::make_pair((long)cands[i].pos, cands[i].pattern_id);
    std::sort(emitted.begin(), emitted.end());
    std::vector<KeyIdx> c(ncands);
    for (uint64_t i = 0; i < ncands; i++) { c[i].key = emitted[i].first; c[i].idx = emitted[i].second; }
    std::sort(c.begin(), c.end());                             // l.normalize(), filter_bitvec.cc:91
    std::map<std::pair<uint32_t, long>, uint64_t> by_head;     // (pattern, first candidate) -> hit
    for (uint64_t i = 0; i < nhits; i++) by_head[std::make_pair(hits[i].pattern_id, (long)hits[i].first)] = i;
    std::vector<bool> done(ncands, false);
    for (uint64_t i = 0; i < ncands; i++) {
      if (done[i]) continue;
      const uint32_t pid = (uint32_t)c[i].idx;
      const long firstpos = c[i].key;
      long pos = firstpos;
      done[i] = true;
      for (uint64_t j = i + 1; j < ncands && (done[j] || c[j].key <= pos + 2 * k + 1); j++)
        if (!done[j] && (uint32_t)c[j].idx == pid) { pos = c[j].key; done[j] = true; }
      std::map<std::pair<uint32_t, long>, uint64_t>::iterator it = by_head.find(std::make_pair(pid, firstpos));
      if (it != by_head.end() && !placed[it->second]) { order.push_back(it->second); placed[it->second] = true; }
    }
  }
  for (uint64_t i = 0; i < nhits; i++)
    if (!placed[i]) order.push_back(i);
}

inline void pcr_pairs(const pm_hit *hits, uint64_t nhits, uint32_t n, const uint32_t *patlen, const int64_t *seq_key,
                      uint64_t n_seqs, const pm_pcr_options &opt, std::vector<pm_pcr_pair> &out) {
  // the hits in the order find_patterns reports them, then l.normalize() (pcr_match.cc:958): std::sort
  // on the hit key alone.  A hit whose own alignment failed stays in the list (it takes part in
  // the sort) and is never paired (pcr_match.cc:1114).
  std::vector<uint64_t> order;
  hit_order(hits, nhits, opt.cands, opt.ncands, opt.k, order);
  std::vector<KeyIdx> sorted;
  for (uint64_t i : order) {
    const pm_hit &h = hits[i];
    if (h.pattern_id >= 1 && h.pattern_id <= 2 * n) { KeyIdx e; e.key = (long)h.end; e.idx = i; sorted.push_back(e); }
  }
  std::sort(sorted.begin(), sorted.end());
  std::vector<uint64_t> l(sorted.size());
  std::vector<bool> bogus(sorted.size(), false);
  for (size_t i = 0; i < sorted.size(); i++) {
    l[i] = sorted[i].idx;
    bogus[i] = !(hits[l[i]].editdist >= 0 && hits[l[i]].editdist <= opt.k);
  }
  std::map<uint32_t, std::vector<size_t> > m;  // pattern id -> positions in l
  for (size_t i = 0; i < l.size(); i++) m[hits[l[i]].pattern_id].push_back(i);
  std::vector<bool> done(l.size(), false);
  const long err = opt.indels ? opt.k : 1;
  std::vector<size_t> l1;
  for (size_t hi = 0; hi < l.size(); hi++) {
    const pm_hit &h = hits[l[hi]];
    if (bogus[hi]) { done[hi] = true; continue; }
    const unsigned long pid = h.pattern_id;
    const long pos = (long)h.end;
    // partner patterns, pcr_match.cc:968-997
    unsigned long pid1 = 0, pid2 = 0;
    if (pid <= n && pid % 2 == 1) pid1 = pid + 1;
    else if (pid > n && (pid - n) % 2 == 0) pid1 = pid - 1;
    if (opt.allorient) {
      if (pid <= n) {
        if (pid % 2 == 1) { pid2 = pid + n + 1; }
        else { pid1 = pid - 1; pid2 = pid + n - 1; }
      } else {
        if (pid % 2 == 0) { pid2 = pid - n - 1; }
        else { pid1 = pid + 1; pid2 = pid - n + 1; }
      }
    }
    // stretch window, pcr_match.cc:1019-1042 (types as there: the -d comparisons are unsigned)
    const unsigned long pair = (pid - ((pid > n) ? n : 0) + 1) / 2;
    long stretch_max = (long)opt.maxdist, stretch_min = (long)opt.mindist;
    if (opt.betweenlen) {
      long plen = 0;
      if (pid1 != 0) plen = (long)patlen[pid1];
      if (pid2 != 0 && (long)patlen[pid2] > plen) plen = (long)patlen[pid2];
      stretch_max += plen + (long)patlen[pid];
    }
    const bool sts = opt.sts_sizelb && opt.sts_sizeub && opt.deviation >= 0;
    if (sts) {
      const unsigned long lb = (unsigned long)opt.sts_sizelb[pair - 1], ub = (unsigned long)opt.sts_sizeub[pair - 1];
      if ((unsigned long)stretch_max > ub + opt.deviation) stretch_max = (long)(ub + opt.deviation);
      if ((unsigned long)stretch_min < lb - opt.deviation) stretch_min = (long)(lb - opt.deviation);
    }
    stretch_max += pos - (long)patlen[pid] + err;
    stretch_min += pos - (long)patlen[pid] - err;

    l1.clear();
    for (unsigned long q : {pid1, pid2}) {
      if (!q) continue;
      std::map<uint32_t, std::vector<size_t> >::iterator mit = m.find((uint32_t)q);
      if (mit == m.end()) continue;
      const std::vector<size_t> &v = mit->second;
      size_t j = 0, step = v.size();  // first hit of q ending at or after stretch_min
      while (step > 0) {
        const size_t half = step / 2;
        if ((long)hits[l[v[j + half]]].end < stretch_min) { j += half + 1; step -= half + 1; } else { step = half; }
      }
      for (; j < v.size(); j++) {
        if (done[v[j]]) continue;  // already had its turn (its key reads 0 in the reference)
        if ((long)hits[l[v[j]]].end > stretch_max) break;
        if (!bogus[v[j]]) l1.push_back(v[j]);
      }
    }
    for (size_t j : l1) {
      const pm_hit &h1 = hits[l[j]];
      const long length = (long)(h.end - h.start + 1);  // pattern_alignment::length()
      const long pe = (long)h.end, pe1 = (long)h1.end;
      const long ps = pe - length + 1, ps1 = (long)h1.start;
      const bool rc = pid > n, rc1 = h1.pattern_id > n;
      const unsigned long ind = pid - (rc ? n : 0), ind1 = h1.pattern_id - (rc1 ? n : 0);
      const unsigned long pind = (ind < ind1 ? ind : ind1) / 2 + 1;
      const long amplicon_len = opt.betweenlen ? ps1 - pe : pe1 - ps;
      // FastaFile::is_subseq(ps, pe1), fasta_io.t:203-213
      const long e0 = pcr_locate(seq_key, n_seqs, ps + 1);
      if (e0 < 0 || e0 != pcr_locate(seq_key, n_seqs, pe1)) continue;
      if (!(amplicon_len <= (long)opt.maxdist && amplicon_len >= (long)opt.mindist)) continue;
      if (sts) {
        const unsigned long lb = (unsigned long)opt.sts_sizelb[pind - 1];
        const int ub = (int)opt.sts_sizeub[pind - 1];
        if (!(((unsigned long)(amplicon_len + opt.deviation) >= lb) && (amplicon_len <= ub + opt.deviation))) continue;
      }
      pm_pcr_pair p;
      p.first = l[hi];
      p.second = l[j];
      p.amplicon_len = amplicon_len;
      p.pair = (uint32_t)pind;
      p.reserved = 0;
      out.push_back(p);
    }
    done[hi] = true;
  }
}

}  // namespace pm
#endif
