// pcr_match (B200 build) -- drop-in for the reference's second command line
// tool over the PatternMatch scan: primers come in PAIRS, both strands are
// always searched, and hits of the two primers of a pair that lie within
// -m / -M of each other on one sequence are reported as an amplicon.  Option
// letters, primer input (-p / -P / -F / -S), the pattern array with its
// per-pattern constraints, the pairing rules and the output format follow
// pcr_match.cc (options 77-262, alignformat 343-746, primers 764-913, pairing
// loop 921-1240).  The database scan runs on the GPU through
// include/pm_pattern_match.hpp; there is no CPU scan.
//
// The reference hands its hits over in batches (-R) and pairs batch by batch; a
// hit too close to the end of a batch waits for the next one.  Here the scan
// returns every hit at once and the pairing makes one pass in position order.
// The two agree except when one primer site lies INSIDE its partner's site and
// within -M of a batch end (there the reference's answer depends on -R).
//
// Not available here: -x (hash seeds), -D/-B (database index files: the FASTA
// file is read whole).
#include <getopt.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "pm_cli_common.hpp"
#include "pm_pattern_match.hpp"

using namespace pmb200;
using namespace pmcli;

namespace {

struct Options {
  std::string patterns, database;
  bool pattern_file = false, sts_pattern_file = false, fasta_pattern_file = false;
  bool rev_comp = false, allorient = false, ucdict = false, betweenlen = false;
  int nmismatch = 0;
  int eos_char = '\n';
  std::string alignformat =
      ">%h\\n %>T %>s ... %l ... %<e %<T\\n %>A  %!>s    %!l    %!<e  %<A\\n %>Q %>r%!>s    %!l    %!<e%<r %<Q %a%R\\n";
  bool wc = false, tn = false, indels = true;
  int node = 0, device = 0, dbind = 0;
  int maxdist = 2000, mindist = 0, deviation = -1;
  int stlen = 0, edlen = 0, fplen = 0, tplen = 0;
  unsigned long report_interval = 1000;
  std::string outfile;
};

void usage(const char *msg) {
  if (msg && *msg) std::cerr << msg << "\n\n";
  std::cerr << "Usage: pcr_match [options]\n\n"
               "  -i <sequence-database>  FASTA database. Required.\n"
               "  -p <sequences> | -P <file> | -F <fasta-file> | -S <UniSTS-file>   primer pairs (\"-\" = stdin)\n"
               "  -o <output-file>        appended to; default standard out\n"
               "  -k <n> / -K <n>         edits / substitutions permitted (0..3)\n"
               "  -r                      reverse complement the second primer of every pair (implied by -S)\n"
               "  -a                      output all primer-pair orientations\n"
               "  -s -e -5 -3 <#chars>    exact (or ~inexact) bases at start/end/5'/3'\n"
               "  -u                      uppercase primers\n"
               "  -w / -W                 IUPAC wildcards without / with text N\n"
               "  -E <int>                end-of-sequence character\n"
               "  -m <int> / -M <int>     minimum / maximum amplicon length (0 / 2000)\n"
               "  -d <int>                deviation from the reported amplicon length (-S)\n"
               "  -b                      ignore primers in amplicon length computation\n"
               "  -A <fmt>                alignment output format\n"
               "  -N <int>                index: 0 auto, 3 keyword tree, 4 bitvector, 5 inexact bitvector, 7-14\n"
               "  -G <int>                CUDA device ordinal\n";
  std::exit(1);
}

Options parse(int argc, char **argv) {
  Options o;
  int c;
  while ((c = getopt(argc, argv, "p:i:o:P:S:F:E:R:k:K:s:e:5:3:x:hrvVubaA:BD:wWN:M:m:d:G:")) != -1) {
    switch (c) {
      case 'p': o.patterns = optarg; o.pattern_file = false; break;
      case 'P': o.patterns = optarg; o.pattern_file = true; break;
      case 'S': o.patterns = optarg; o.sts_pattern_file = true; break;
      case 'F': o.patterns = optarg; o.fasta_pattern_file = true; break;
      case 'i': o.database = optarg; break;
      case 'o': o.outfile = optarg; break;
      case 'k': o.nmismatch = std::atoi(optarg); o.indels = true; break;
      case 'K': o.nmismatch = std::atoi(optarg); o.indels = false; break;
      case '3': o.tplen = tilde_int(optarg); break;
      case '5': o.fplen = tilde_int(optarg); break;
      case 's': o.stlen = tilde_int(optarg); break;
      case 'e': o.edlen = tilde_int(optarg); break;
      case 'R': o.report_interval = (unsigned long)std::atoi(optarg); break;
      case 'A': o.alignformat = optarg; break;
      case 'w': o.wc = true; o.tn = false; break;
      case 'W': o.wc = true; o.tn = true; break;
      case 'u': o.ucdict = true; break;
      case 'N': o.node = std::atoi(optarg); break;
      case 'G': o.device = std::atoi(optarg); break;
      case 'M': o.maxdist = std::atoi(optarg); break;
      case 'm': o.mindist = std::atoi(optarg); break;
      case 'd': o.deviation = std::atoi(optarg); break;
      case 'E': {
        int ec;
        if (!std::sscanf(optarg, "%i", &ec)) usage("Invalid end-of-sequence specification.\n");
        o.eos_char = (char)ec;
      } break;
      case 'b': o.betweenlen = true; break;
      case 'r': o.rev_comp = true; break;
      case 'a': o.allorient = true; break;
      case 'D':
        o.dbind = std::atoi(optarg);
        if (o.dbind < 0 || o.dbind > 4) usage("Invalid integer for fasta database indexing (-D).");
        break;
      case 'v': case 'V': case 'B': break;  // accepted, no effect here (the library has its own reader)
      case 'x': usage("Exact seeds (-x) are not available in the B200 build.");
      case 'h':
      default: usage(0);
    }
  }
  if (o.patterns.empty() || o.database.empty()) usage(0);
  if (o.nmismatch < 0) usage("Number of mismatches (-k) must be at least 0");
  return o;
}

std::string spaces(long v) {  // pcr_match.cc:264-268
  std::string ret = " ";
  while (v /= 10) ret += ' ';
  return ret;
}
std::string spaces(const std::string &s) { return std::string(s.size(), ' '); }

// one side of a reported pair
struct Side {
  long s, e, S, E;      // start / end in the sequence, start / end in the stream
  unsigned d;
  std::string p, patdef, q, r, R;
  Aligned al;
};

// alignformat, pcr_match.cc:343-746.  `a` = first side (%>x), `b` = second side (%<x).
void alignformat(std::ostream &os, const std::string &f, const Side &a, const Side &b, unsigned long i,
                 const StsEntry &sts, bool ppo, const std::string &h, const std::string &H, unsigned long fidx,
                 const std::string &amplicon, unsigned long ncount) {
  for (size_t pos = 0; pos < f.size(); pos++) {
    if (f[pos] == '%') {
      pos++;
      if (pos >= f.size()) { os << "%"; break; }
      bool widthonly = false;
      if (f[pos] == '!') { widthonly = true; pos++; }
      int dirn = 0;
      if (pos < f.size() && f[pos] == '>') dirn = 1;
      if (pos < f.size() && f[pos] == '<') dirn = -1;
      if (dirn != 0) pos++;
      if (pos >= f.size()) break;
      const Side *x = dirn > 0 ? &a : (dirn < 0 ? &b : 0);
      switch (f[pos]) {
        case 's': if (x) { if (!widthonly) os << x->s; else os << spaces(x->s); } break;
        case 'e': if (x) { if (!widthonly) os << x->e; else os << spaces(x->e); } break;
        case 'l':
          if (x) os << x->e - x->s;
          else if (!widthonly) os << b.e - a.s;
          else os << spaces(b.e - a.s);
          break;
        case 'S': if (x) os << x->S; break;
        case 'E': if (x) os << x->E; break;
        case 'i': os << i; break;
        case 'd': if (x) os << x->d; break;
        case 'p': if (x) os << x->p; break;
        case 'P': if (x) os << x->patdef; break;
        case 'I': os << sts.id; break;
        case 'L':
          if (sts.sizeub != sts.sizelb) {
            if (dirn > 0) os << sts.sizelb;
            else if (dirn < 0) os << sts.sizeub;
            else os << sts.sizelb << "-" << sts.sizeub;
          } else {
            os << sts.sizelb;
          }
          break;
        case 'D': {
          const int amplen = (int)(b.e - a.s);
          int deviance = 0;
          if ((unsigned long)amplen > sts.sizeub) deviance = (int)(amplen - sts.sizeub);
          else if ((unsigned long)amplen < sts.sizelb) deviance = (int)(sts.sizelb - amplen);
          os << deviance;
        } break;
        case 'a': os << sts.acc; break;
        case 'O': os << sts.species; break;
        case '&': os << sts.altacc; break;
        case 'X': os << sts.chrom; break;
        case 'q': if (x) os << x->q; break;
        case 'Q': if (x) { if (!widthonly) os << x->al.Q; else os << spaces(x->al.Q); } break;
        case 'r':
          if (x) os << x->r;
          else os << (ppo ? "F" : "R");
          break;
        case 'R':
          if (x) os << x->R;
          else os << (ppo ? "" : " REVERSE-STRAND");
          break;
        case 't': if (x) os << x->al.t; break;
        case 'T': if (x) os << x->al.T; break;
        case 'A': if (x) { if (!widthonly) os << x->al.A; else os << spaces(x->al.A); } break;
        case 'h': os << h; break;
        case 'H': os << H; break;
        case 'f': os << fidx; break;
        case '@': os << amplicon; break;
        case '*': os << (ppo ? amplicon : reverse_comp(amplicon)); break;
        case 'N': os << ncount; break;
        case '%': os << "%"; break;
        case '0':
          os << H << " " << a.s + 1 << ".." << b.e << '\t' << sts.id << '\t';
          if (sts.acc != "") {
            os << '\t' << sts.acc;
            if (sts.chrom != "") {
              os << '\t' << sts.chrom;
              if (sts.altacc != "") {
                os << '\t' << sts.altacc;
                if (sts.species != "") os << '\t' << sts.species;
              }
            }
          }
          break;
        default: os << f[pos];
      }
    } else if (f[pos] == '\\') {
      escape(os, f, pos);
    } else {
      os << f[pos];
    }
  }
}

}  // namespace

int main(int argc, char **argv) {
  Options opt = parse(argc, argv);
  std::ostream *out = &std::cout;
  std::ofstream ofs;
  if (!opt.outfile.empty()) {
    ofs.open(opt.outfile.c_str(), std::ios::out | std::ios::app | std::ios::ate);
    out = &ofs;
  }

  // primers, pcr_match.cc:764-836
  std::vector<std::string> patterns, patdefs;
  std::vector<StsEntry> stsarray(1);  // index 0 wasted, like the reference's
  {
    std::ifstream file;
    std::istream *is = &std::cin;
    if ((opt.pattern_file || opt.sts_pattern_file || opt.fasta_pattern_file) && opt.patterns != "-") {
      file.open(opt.patterns.c_str());
      is = &file;
    }
    if (opt.pattern_file) {
      read_words(*is, patterns);
    } else if (opt.sts_pattern_file) {
      StsEntry s;
      while (s.read(*is)) {
        if (s.forward_primer == "") break;
        stsarray.push_back(s);
        patterns.push_back(s.forward_primer);
        patterns.push_back(s.reverse_primer);
      }
    } else if (opt.fasta_pattern_file) {
      read_fasta_patterns(*is, patterns, patdefs);
    } else {
      std::istringstream ss(opt.patterns);
      read_words(ss, patterns);
    }
  }
  if (patterns.empty()) return 0;
  if (patterns.size() % 2 != 0) usage("Odd number of primers!");
  if (opt.ucdict)
    for (std::string &p : patterns)
      for (char &ch : p) ch = (char)toupper((unsigned char)ch);
  if (opt.rev_comp || opt.sts_pattern_file) {
    opt.rev_comp = true;
    for (size_t i = 1; i < patterns.size(); i += 2) patterns[i] = reverse_comp(patterns[i]);
  }

  // ids 1..n as given, n+1..2n reverse complements; -5 / -3 swap for the second primer of a pair
  const unsigned long n = patterns.size(), N1 = 2 * n;
  std::vector<std::string> patarray(N1 + 1);
  std::vector<std::pair<int, int> > patconst(N1 + 1);
  std::vector<long> patlen(N1 + 1, 0);
  for (unsigned long i = 1; i <= n; i++) {
    const std::string &p = patterns[i - 1];
    const int fplen = (i % 2 == 0) ? opt.tplen : opt.fplen, tplen = (i % 2 == 0) ? opt.fplen : opt.tplen;
    patarray[i] = p;
    patarray[i + n] = reverse_comp(p);
    patlen[i] = (long)p.length();
    patlen[i + n] = (long)patarray[i + n].length();
    pattern_constraints((int)p.length(), opt.stlen, opt.edlen, fplen, tplen, patconst[i], patconst[i + n]);
  }

  FastaStream ff;
  if (!ff.load(opt.database, false, (unsigned char)opt.eos_char, opt.dbind))
    fatal("Fatal error: can't read sequence database " + opt.database);

  b200_index *pm = pick_pattern_index(opt.node, opt.nmismatch, opt.wc, opt.tn, opt.indels, (char)opt.eos_char,
                                      opt.device);
  // a hit whose alignment at its own position fails is simply never paired (pcr_match.cc:1114)
  pm->tolerate_bogus(true);
  pm->alignment_rules(PM_ALIGN_PCR_MATCH);
  for (unsigned long i = 1; i <= N1; i++) pm->add_pattern(patarray[i], i, patconst[i].first, patconst[i].second);
  pm->init(ff);

  // every hit the scan reports
  std::vector<const pm_hit *> l;
  {
    pattern_hit_vector phv;
    while (pm->find_patterns(ff, phv, 1ul << 20) || !phv.empty()) {
      std::vector<const pm_hit *> &batch = pm->batch();
      for (const pm_hit *h : batch)   // every hit: one whose own alignment failed still takes part in the sort
        if (h->pattern_id) l.push_back(h);
      phv.clear();
      batch.clear();
    }
  }
  // the pairing itself is the library's (pm_pcr_pairs, csrc/pm_pcr.hpp)
  std::vector<pm_hit> hits;
  hits.reserve(l.size());
  for (const pm_hit *h : l) hits.push_back(*h);
  std::vector<uint32_t> plen32(patlen.begin(), patlen.end());
  std::vector<uint64_t> sizelb, sizeub;
  for (size_t i = 1; i < stsarray.size(); i++) { sizelb.push_back(stsarray[i].sizelb); sizeub.push_back(stsarray[i].sizeub); }
  pm_pcr_options po;
  std::memset(&po, 0, sizeof po);
  po.k = opt.nmismatch; po.indels = opt.indels; po.allorient = opt.allorient; po.betweenlen = opt.betweenlen;
  po.deviation = opt.sts_pattern_file ? opt.deviation : -1;
  po.mindist = opt.mindist; po.maxdist = opt.maxdist;
  po.sts_sizelb = opt.sts_pattern_file ? sizelb.data() : 0;
  po.sts_sizeub = opt.sts_pattern_file ? sizeub.data() : 0;
  // the candidates behind the hits (filter_bitvec only): the order of hits that end at one position
  pm->last_candidates(&po.cands, &po.ncands);
  pm_pcr_pair *pairs = 0;
  uint64_t npairs = 0;
  if (pm_pcr_pairs(hits.data(), hits.size(), (uint32_t)n, plen32.data(), ff.seq_keys(), ff.n_seqs(), &po, &pairs,
                   &npairs) != PM_OK)
    fatal(std::string("Fatal error: ") + pm_last_error());

  StsEntry null_sts;
  for (uint64_t pi = 0; pi < npairs; pi++) {
    const pm_hit &h = hits[pairs[pi].first];
    const unsigned long pid = h.pattern_id;
    const pm_hit &h1 = hits[pairs[pi].second];
    const unsigned long pidb = h1.pattern_id;
    const long length = (long)(h.end - h.start + 1), length1 = (long)(h1.end - h1.start + 1);
    const long pe = (long)h.end, pe1 = (long)h1.end;
    const long spe = (long)ff.get_seq_pos(pe), spe1 = (long)ff.get_seq_pos(pe1);
    const long sps = spe - length + 1, sps1 = spe1 - length1 + 1;
    const long ps = pe - length + 1, ps1 = pe1 - length1 + 1;
    bool rc = pid > n, rc1 = pidb > n;
    const unsigned long ind = pid - (rc ? n : 0), ind1 = pidb - (rc1 ? n : 0);
    const unsigned long pind = (ind < ind1 ? ind : ind1) / 2 + 1;
    const StsEntry &stsref = opt.sts_pattern_file ? stsarray[pind] : null_sts;
    if (opt.rev_comp) {
      if (ind % 2 == 0) rc = !rc;
      else if (ind1 % 2 == 0) rc1 = !rc1;
    }
    const long amplicon_len = (long)pairs[pi].amplicon_len;
    const std::string amplicon = amplicon_len > 0 ? ff.text(ps, ps + amplicon_len) : std::string();
    unsigned long ncount = 0;
    for (char ch : amplicon) ncount += (ch == 'N' || ch == 'n');
    Side a, b;
    a.s = sps; a.e = spe; a.S = ps; a.E = pe; a.d = (unsigned)h.editdist;
    b.s = sps1; b.e = spe1; b.S = ps1; b.E = pe1; b.d = (unsigned)h1.editdist;
    a.p = patarray[ind]; b.p = patarray[ind1];
    if (opt.fasta_pattern_file) { a.patdef = patdefs[ind - 1]; b.patdef = patdefs[ind1 - 1]; }
    a.q = patarray[pid]; b.q = patarray[pidb];
    a.r = rc ? "R" : "F"; b.r = rc1 ? "R" : "F";
    a.R = rc ? " REVCOMP" : ""; b.R = rc1 ? " REVCOMP" : "";
    a.al = render(h, patarray[pid], ff.text(h.start, h.end));
    b.al = render(h1, patarray[pidb], ff.text(h1.start, h1.end));
    alignformat(*out, opt.alignformat, a, b, pind, stsref, ind < ind1, ff.header(pe), ff.short_header(pe),
                ff.index(pe), amplicon, ncount);
  }
  pm_free(pairs);
  out->flush();
  delete pm;
  return 0;
}
