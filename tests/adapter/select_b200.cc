// select_b200.cc -- replaces the reference's select.cc in the adapter build: the same
// pick_pattern_index signature (select.h:8-19), every selection answered by a b200_match with
// the semantics of the index the reference would have built.  The automatic choice restates
// select.cc:32-141 (NOPRIMEGEN build): exact index for k = 0 (shift_and with wildcards),
// exact_halves for k = 1 when the primers are long enough for the database's alphabet and the
// exact-base constraints are shorter than half the primers on average, exact_bases from 6
// constrained characters on, filter_bitvec otherwise.  Hash-table selections (-N 6, a seed
// length with -x) are not part of the library and end the run with a message.
#include <stdlib.h>

#include "select.h"
#include "b200_match.h"

#ifndef B200_PCR_RULES
#define B200_PCR_RULES 0   /* 1 in the pcr_match build: editdist_alignment of every hit, also for k = 0 */
#endif

PatternMatch *pick_pattern_index(CharacterProducer *const &ff, int pmselect, int nmismatch,
                                 std::vector<std::pair<int, int> > *exact_const, std::vector<int> *patlen,
                                 int seedlen, bool wildcard, bool textn, bool indels, bool dna_mut, char eos,
                                 bool verbose) {
  long min_exact = 0, min_len = 0, longer_than_half = 0, sum_over_half = 0, min_inexact = MAXINT;
  if (exact_const && patlen) {
    min_exact = MAXINT;
    for (size_t i = 1; i < exact_const->size(); i++) {
      const long ex = std::max((*exact_const)[i].first, (*exact_const)[i].second);
      const long len = (*patlen)[i];
      min_exact = std::min(min_exact, ex);
      sum_over_half += ex - len / 2;
      longer_than_half += (ex - len / 2 >= 0) ? 1 : 0;
      min_inexact = std::min(min_inexact, len - ex);
    }
  }
  if (patlen) {
    min_len = MAXINT;
    for (size_t i = 1; i < patlen->size(); i++) min_len = std::min<long>(min_len, (*patlen)[i]);
  }
  if (min_inexact > min_len) min_inexact = min_len;
  if (exact_const && nmismatch >= min_inexact && nmismatch > 0) {
    timestampli("Fatal error: Number of edits >= Minimum number of inexact bases: ", min_inexact);
    exit(1);
  }
  if (dna_mut) {
    timestamp("Fatal error: b200_match does not score DNA mutations (-k .N)");
    exit(1);
  }
  int index = 0;
  if (pmselect == 0) {
    if (nmismatch == 0) {
      index = wildcard ? PM_INDEX_SHIFT_AND : PM_INDEX_KEYWORD_TREE;
    } else if (nmismatch == 1 && ((min_len >= 12 && ff->size() < 10) || (min_len >= 8 && ff->size() >= 10)) &&
               (longer_than_half <= 0 || sum_over_half <= 0)) {
      index = PM_INDEX_EXACT_HALVES;
    } else if (min_exact >= 6) {
      index = PM_INDEX_EXACT_BASES;
    } else if (seedlen > 0) {
      timestamp("Fatal error: b200_match has no hash-table index (seed length given)");
      exit(1);
    } else {
      index = PM_INDEX_FILTER_BITVEC;
    }
  } else if (pmselect >= 1 && pmselect <= 3) {
    index = PM_INDEX_KEYWORD_TREE;
  } else if (pmselect == 4) {
    index = PM_INDEX_SHIFT_AND;
  } else if (pmselect == 5) {
    index = PM_INDEX_FILTER_BITVEC;
  } else if (pmselect >= 7 && pmselect <= 10) {
    index = PM_INDEX_EXACT_BASES;
  } else if (pmselect >= 11 && pmselect <= 14) {
    index = PM_INDEX_EXACT_HALVES;
  } else {
    timestamp("Fatal error: pattern index selection not available in b200_match");
    exit(1);
  }
  if (verbose) timestamp("Using the B200 scan (libpm_b200)...");
  return new b200_match(index, (unsigned)nmismatch, eos, wildcard, textn, indels, (int)ff->size(), B200_PCR_RULES != 0);
}
