// primer_match (B200 build) -- drop-in for the reference's command line tool
// on the path this repository accelerates: option letters, pattern input,
// reverse complements, per-pattern constraints, hit and count output formats
// follow primer_match.cc (options 86-262, pattern array 943-1080, main loop
// 1100-1268, alignformat 352-748, countformat 750-838).  The database scan runs
// on the GPU through include/pm_pattern_match.hpp; there is no CPU scan.
//
// Not available here (the reference features outside the accelerated path):
// -T (translation), -x (hash seeds), -k .N (DNA mutation scoring), -D/-B/-I
// (database index files: the FASTA file is read whole).
#include <time.h>
#include <getopt.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
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
  bool pattern_file = false, fasta_pattern_file = false, sts_pattern_file = false;
  bool rev_comp = false, ucdict = false;
  int tplen = 0, fplen = 0, stlen = 0, edlen = 0;
  int eos_char = '\n';
  int nmismatch = 0, maxcount = 0;
  unsigned long report_interval = 1000;
  std::string alignformat = ">%h\\n %T %s %e %d\\n %A\\n %Q %i%R\\n";
  std::string countformat = "%i %r %q %c%+ ( %C )\\n";
  bool alignments = true, counts = false, aggregate = false;
  bool wc = false, tn = false, indels = true;
  int node = 0, device = 0, dbind = 0;
  std::string outfile;
};

void usage(const char *msg) {
  if (msg && *msg) std::cerr << msg << "\n\n";
  std::cerr << "Usage: primer_match [options]\n\n"
               "  -i <sequence-database>  FASTA database. Required.\n"
               "  -p <sequences> | -P <file> | -F <fasta-file> | -S <UniSTS-file>   primers (\"-\" = stdin)\n"
               "  -o <output-file>        appended to; default standard out\n"
               "  -k <n> / -K <n>         edits / substitutions permitted (0..3)\n"
               "  -r                      match reverse complements too\n"
               "  -s -e -5 -3 <#chars>    exact (or ~inexact) bases at start/end/5'/3'\n"
               "  -u                      uppercase primers and database\n"
               "  -w / -W                 IUPAC wildcards without / with text N\n"
               "  -E <int>                end-of-sequence character\n"
               "  -c -a -M <n> -C <fmt>   counts output\n"
               "  -A <fmt>                alignment output format\n"
               "  -R <int>                report interval\n"
               "  -N <int>                index: 0 auto, 3 keyword tree, 4 bitvector, 5 inexact bitvector\n"
               "  -G <int>                CUDA device ordinal\n";
  std::exit(1);
}

Options parse(int argc, char **argv) {
  Options o;
  int c;
  while ((c = getopt(argc, argv, "p:i:o:P:F:S:M:k:K:s:e:3:5:x:E:hrucavA:C:R:BN:D:IwWTG:")) != -1) {
    switch (c) {
      case 'p': o.patterns = optarg; o.pattern_file = false; break;
      case 'P': o.patterns = optarg; o.pattern_file = true; break;
      case 'F': o.patterns = optarg; o.fasta_pattern_file = true; break;
      case 'i': o.database = optarg; break;
      case 'o': o.outfile = optarg; break;
      case '3': o.tplen = tilde_int(optarg); break;
      case '5': o.fplen = tilde_int(optarg); break;
      case 's': o.stlen = tilde_int(optarg); break;
      case 'e': o.edlen = tilde_int(optarg); break;
      case 'k':
      case 'K':
        if (optarg[0] == '.') usage("DNA mutation scoring (-k .N) is not available in the B200 build.");
        o.nmismatch = std::atoi(optarg);
        o.indels = (c == 'k');
        break;
      case 'r': o.rev_comp = true; break;
      case 'c': o.counts = true; o.alignments = false; break;
      case 'M': o.maxcount = std::atoi(optarg); break;
      case 'A': if (*optarg) o.alignformat = optarg; o.alignments = true; break;
      case 'C': if (*optarg) o.countformat = optarg; o.counts = true; break;
      case 'u': o.ucdict = true; break;
      case 'a': o.aggregate = true; break;
      case 'w': o.wc = true; o.tn = false; break;
      case 'W': o.wc = true; o.tn = true; break;
      case 'R': o.report_interval = (unsigned long)std::atoi(optarg); break;
      case 'N': o.node = std::atoi(optarg); break;
      case 'G': o.device = std::atoi(optarg); break;
      case 'E': {
        int ec;
        if (!std::sscanf(optarg, "%i", &ec)) usage("Invalid end-of-sequence specification.\n");
        o.eos_char = (char)ec;
      } break;
      case 'D':   // primer_match.cc:246-247, 278
        o.dbind = std::atoi(optarg);
        if (o.dbind < 0 || o.dbind > 4) usage("Invalid integer for fasta database indexing (-D).");
        break;
      case 'v': case 'I': case 'B': break;  // accepted, no effect here (the library has its own reader)
      case 'S': o.patterns = optarg; o.sts_pattern_file = true; o.rev_comp = true; break;  // primer_match.cc:137-141
      case 'T': usage("Translation (-T) is not available in the B200 build.");
      case 'x': usage("Exact seeds (-x) are not available in the B200 build.");
      case 'h':
      default: usage(0);
    }
  }
  if (o.patterns.empty() || o.database.empty()) usage("No primers and/or no sequence database supplied.");
  if (o.nmismatch < 0) usage("Number of mismatches (-k) must be >= 0.");
  if (o.maxcount > 0 && !o.counts) usage("Can''t use maxcount (-M) without counts (-c or -C).");
  if (o.aggregate && !o.counts) usage("Can''t use aggregate (-a) without counts (-c or -C).");
  if (o.report_interval == 0) o.report_interval = 1;
  return o;
}

// alignformat, primer_match.cc:352-748 (the codes that do not need -S / -T / Tm)
void alignformat(std::ostream &os, const std::string &f, long s, long e, long five, long three, long S, long E,
                 unsigned long i, unsigned d, const std::string &p, const std::string &P, const std::string &q,
                 const Aligned &al, const char *r, const char *R, const std::string &h, const std::string &H,
                 unsigned long fidx, const StsEntry &sts) {
  unsigned mat = 0, del = 0, ins = 0, sub = 0, wcm = 0;
  for (char c : al.A) {
    switch (c) {
      case '|': mat++; break;
      case '^': del++; break;
      case 'v': ins++; break;
      case '*': sub++; break;
      case '+': wcm++; break;
    }
  }
  for (size_t pos = 0; pos < f.size(); pos++) {
    if (f[pos] == '%') {
      pos++;
      if (pos >= f.size()) { os << "%"; break; }
      switch (f[pos]) {
        case 's': os << s; break;
        case 'e': os << e; break;
        case 'l': os << e - s; break;
        case '5': os << five; break;
        case '3': os << three; break;
        case 'S': os << S; break;
        case 'E': os << E; break;
        case 'i': os << i; break;
        case 'd': os << d; break;
        case 'D': os << (long)p.length() - (s - e); break;
        case 'p': os << p; break;
        case 'P': os << P; break;
        case 'q': os << q; break;
        case 'Q': os << al.Q; break;
        case 'r': os << r; break;
        case 'R': os << R; break;
        case 't': os << al.t; break;
        case 'T': os << al.T; break;
        case 'U': os << (std::string(r) == "R" ? reverse_comp(al.t) : al.t); break;  // the text in primer orientation
        case 'F': os << -1; break;  // reading frame: -T only (default argument, primer_match.cc:378)
        case 'n': break;            // translated bases: -T only
        case 'A': os << al.A; break;
        case 'h': os << h; break;
        case 'H': os << H; break;
        case 'f': os << fidx; break;
        case 'L':
          if (sts.sizeub != sts.sizelb) os << sts.sizelb << "-" << sts.sizeub;
          else os << sts.sizelb;
          break;
        case '%': os << "%"; break;
        case '|': os << mat; break;
        case '^': os << del; break;
        case 'v': os << ins; break;
        case '*': os << sub; break;
        case '+': os << wcm; break;
        case '=': {
          const size_t len0 = al.T.length(), width0 = 50;
          size_t start0 = 0;
          for (size_t i0 = 0; i0 < len0; i0 += width0) {
            size_t nchars = width0;
            if (i0 + nchars > len0) nchars = len0 - i0;
            size_t end0 = start0 + nchars, edits = nchars;
            for (size_t j0 = 0; j0 < nchars; j0++) {
              if (al.A[i0 + j0] == '|' || al.A[i0 + j0] == '+') edits--;
              if (al.A[i0 + j0] == 'v') end0--;
            }
            os << " " << al.T.substr(i0, width0) << " " << start0 << " " << end0 << " " << edits << "\n"
               << " " << al.A.substr(i0, width0) << "\n"
               << " " << al.Q.substr(i0, width0) << " " << i << R << "\n";
            if (len0 - i0 > width0) os << std::endl;
            start0 = end0;
          }
        } break;
        default:
          if (!sts_code(os, f[pos], sts)) os << f[pos];
      }
    } else if (f[pos] == '\\') {
      escape(os, f, pos);
    } else {
      os << f[pos];
    }
  }
}

// countformat, primer_match.cc:750-838
void countformat(std::ostream &os, const std::string &f, unsigned long i, const std::string &p, const std::string &P,
                 const std::string &q, const char *r, const char *R, unsigned long c, const unsigned long *C,
                 unsigned k, bool gtmax, const StsEntry &sts) {
  for (size_t pos = 0; pos < f.size(); pos++) {
    if (f[pos] == '%') {
      pos++;
      if (pos >= f.size()) { os << "%"; break; }
      switch (f[pos]) {
        case 'i': os << i; break;
        case 'p': os << p; break;
        case 'P': os << P; break;
        case 'q': os << q; break;
        case 'r': os << r; break;
        case 'R': os << R; break;
        case 'c': os << c; break;
        case 'C':
          for (unsigned j = 0; j < k; j++) os << C[j] << " ";
          os << C[k];
          break;
        case '+': if (gtmax) os << "+"; break;
        case '%': os << "%"; break;
        case 'L':
          if (sts.sizeub != sts.sizelb) os << sts.sizelb << "-" << sts.sizeub;
          else os << sts.sizelb;
          break;
        default:
          if (!sts_code(os, f[pos], sts)) os << f[pos];
      }
    } else if (f[pos] == '\\') {
      escape(os, f, pos);
    } else {
      os << f[pos];
    }
  }
}

}  // namespace

static double now_s() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char **argv) {
  const double t_start = now_s();
  const bool timing = getenv("PM_TIMING") != 0;   // stage times to stderr
  Options opt = parse(argc, argv);
  std::ostream *out = &std::cout;
  std::ofstream ofs;
  if (!opt.outfile.empty()) {
    ofs.open(opt.outfile.c_str(), std::ios::out | std::ios::app | std::ios::ate);
    out = &ofs;
  }

  std::vector<std::string> patterns, patdefs;
  std::vector<StsEntry> stsarray(1);  // [(i + 1) / 2] for pattern i; index 0 wasted (primer_match.cc:955-989)
  const StsEntry null_sts;
  if (opt.sts_pattern_file) {  // primer_match.cc:901-916: both primers of every UniSTS line
    std::ifstream file;
    std::istream *is = &std::cin;
    if (opt.patterns != "-") { file.open(opt.patterns.c_str()); is = &file; }
    StsEntry e;
    while (e.read(*is)) {
      if (e.forward_primer == "") break;
      patterns.push_back(e.forward_primer);
      patterns.push_back(e.reverse_primer);
      stsarray.push_back(e);
    }
  } else if (opt.fasta_pattern_file) {
    if (opt.patterns == "-") read_fasta_patterns(std::cin, patterns, patdefs);
    else { std::ifstream f(opt.patterns.c_str()); read_fasta_patterns(f, patterns, patdefs); }
  } else if (opt.pattern_file) {
    if (opt.patterns == "-") read_words(std::cin, patterns);
    else { std::ifstream f(opt.patterns.c_str()); read_words(f, patterns); }
  } else {
    std::istringstream ss(opt.patterns);
    read_words(ss, patterns);
  }
  if (patterns.empty()) return 0;
  if (opt.ucdict)
    for (std::string &p : patterns)
      for (char &ch : p) ch = (char)toupper((unsigned char)ch);

  // pattern array with reverse complements and per-pattern constraints, primer_match.cc:943-1080
  const unsigned long n = patterns.size();
  const unsigned long N1 = (opt.rev_comp ? 2 : 1) * n;
  std::vector<std::string> patarray(N1 + 1);
  std::vector<std::pair<int, int> > patconst(N1 + 1);
  for (unsigned long i = 1; i <= n; i++) {
    const std::string &p = patterns[i - 1];
    const int len = (int)p.length();
    patarray[i] = p;
    std::pair<int, int> rev;
    pattern_constraints(len, opt.stlen, opt.edlen, opt.fplen, opt.tplen, patconst[i], rev);
    if (opt.rev_comp) {
      patarray[i + n] = reverse_comp(p);
      patconst[i + n] = rev;
    }
  }

  FastaStream ff;
  if (!ff.load(opt.database, opt.ucdict, (unsigned char)opt.eos_char, opt.dbind))
    fatal("Fatal error: can't read sequence database " + opt.database);

  b200_index *kt = pick_pattern_index(opt.node, opt.nmismatch, opt.wc, opt.tn, opt.indels, (char)opt.eos_char,
                                      opt.device);
  for (unsigned long i = 1; i <= N1; i++) kt->add_pattern(patarray[i], i, patconst[i].first, patconst[i].second);
  const double t_ctx = now_s();
  kt->init(ff);
  const double t_init = now_s();
  double t_first = 0;

  const unsigned K = (unsigned)opt.nmismatch;
  std::vector<unsigned long> patcount(opt.counts ? N1 * (K + 1) : 0, 0);
  std::vector<bool> maxpatcount(opt.counts && opt.maxcount > 0 ? N1 + 1 : 0, false);
  auto idx = [&](unsigned long i, unsigned k) { return (i - 1) * (K + 1) + k; };

  pattern_hit_vector l;
  bool more;
  while ((more = kt->find_patterns(ff, l, opt.report_interval)) || !l.empty()) {
    if (t_first == 0) t_first = now_s();
    std::vector<const pm_hit *> &batch = kt->batch();
    if (opt.alignments && !(K == 0 && !opt.wc)) ff.prefetch_texts(batch);   // one round trip for the batch's texts
    for (size_t hi = 0; hi < l.size(); hi++) {
      const pm_hit &h = *batch[hi];
      const unsigned long pid = h.pattern_id;
      if (!pid || (opt.maxcount > 0 && maxpatcount[pid])) continue;
      const unsigned ed = (unsigned)h.editdist;
      if (opt.alignments) {
        const long p = (long)h.end;
        const long length = (long)(h.end - h.start + 1);  // pattern_alignment::length()
        const long spe = (long)ff.get_seq_pos(p);
        const long sps = spe - length + 1;
        const long pe = p, ps = pe - length + 1;
        const bool rc = pid > n;
        const unsigned long ind = pid - (rc ? n : 0);
        const std::string &pat = patarray[pid];
        // matching_text: the pattern itself for exact_alignment (pattern_alignment.cc:25-43),
        // the text under the alignment otherwise
        std::string matching = (K == 0 && !opt.wc) ? pat : ff.text(h.start, h.end, (long)hi);
        const Aligned al = render(h, pat, matching);
        const std::string patdef = opt.fasta_pattern_file ? patdefs[ind - 1] : std::string();
        alignformat(*out, opt.alignformat, sps, spe, rc ? spe : sps, rc ? sps : spe, ps, pe, ind, ed, patarray[ind],
                    patdef, pat, al, rc ? "R" : "F", rc ? " REVCOMP" : "", ff.header(p), ff.short_header(p),
                    ff.index(p), opt.sts_pattern_file ? stsarray[(ind + 1) / 2] : null_sts);
      }
      if (opt.counts) {
        patcount[idx(pid, ed)]++;
        if (opt.maxcount > 0) {
          unsigned long count = 0;
          for (unsigned k = 0; k <= K; k++) count += patcount[idx(pid, k)];
          if (count >= (unsigned long)opt.maxcount) maxpatcount[pid] = true;
        }
      }
    }
    l.clear();
    batch.clear();
  }

  if (opt.counts) {  // primer_match.cc:1270-1327
    std::vector<unsigned long> counts(K + 1);
    for (unsigned long i = 1; i <= n; i++) {
      unsigned long total = 0;
      for (unsigned k = 0; k <= K; k++) { counts[k] = patcount[idx(i, k)]; total += counts[k]; }
      bool gtmax = opt.maxcount > 0 ? (bool)maxpatcount[i] : false;
      const std::string patdef = opt.fasta_pattern_file ? patdefs[i - 1] : std::string();
      const StsEntry &sts = opt.sts_pattern_file ? stsarray[(i + 1) / 2] : null_sts;
      if (!opt.aggregate)
        countformat(*out, opt.countformat, i, patarray[i], patdef, patarray[i], "F", "", total, counts.data(), K, gtmax,
                    sts);
      if (opt.rev_comp) {
        if (!opt.aggregate) {
          total = 0;
          for (unsigned k = 0; k <= K; k++) counts[k] = 0;
          gtmax = false;
        }
        for (unsigned k = 0; k <= K; k++) { counts[k] += patcount[idx(i + n, k)]; total += patcount[idx(i + n, k)]; }
        if (opt.maxcount > 0) gtmax = gtmax || maxpatcount[i + n];
        if (!opt.aggregate)
          countformat(*out, opt.countformat, i, patarray[i], patdef, patarray[i + n], "R", " REVCOMP", total,
                      counts.data(), K, gtmax, sts);
      }
      if (opt.aggregate)
        countformat(*out, opt.countformat, i, patarray[i], patdef, "", "", "", total, counts.data(), K, gtmax, sts);
    }
  }
  out->flush();
  if (timing) {
    const double t_end = now_s();
    if (t_first == 0) t_first = t_end;
    std::fprintf(stderr,
                 "PM_TIMING {\"context_s\": %.4f, \"init_and_read_s\": %.4f, \"scan_s\": %.4f, \"output_s\": %.4f, "
                 "\"total_s\": %.4f, \"stream_chars\": %llu}\n",
                 t_ctx - t_start, t_init - t_ctx, t_first - t_init, t_end - t_first, t_end - t_start,
                 (unsigned long long)ff.size());
  }
  delete kt;
  return 0;
}
