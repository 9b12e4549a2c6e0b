/*
 * pm_b200.h -- C ABI of the B200-native PrimerMatch database scan.
 *
 * This is the drop-in boundary for ONE path of
 * EdwardsLabProjects/sequence-alignment-tools: the PatternMatch plugin
 * interface (pattern_match.h:83-157) as primer_match drives it
 * (primer_match.cc:1100-1268).  Every entry point names the reference
 * interface it replaces.  Plain pointers and sizes only; no C++ or torch types.
 *
 * Threading: a pm_ctx is owned by one host thread at a time.  All device work
 * of a context is issued on the CUDA stream given to pm_set_stream (default:
 * the legacy default stream) and the blocking entry points synchronise that
 * stream only.
 *
 * Errors: every int-returning function returns PM_OK (0) or a negative
 * PM_E_* code; pm_last_error() gives the message of the calling thread's last
 * failure.  There is no CPU fallback: without a usable CUDA device
 * pm_create fails with PM_E_CUDA.
 */
#ifndef PM_B200_H
#define PM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PM_OK 0
#define PM_E_ARG (-1)         /* bad argument / call order                     */
#define PM_E_CUDA (-2)        /* CUDA runtime failure (message has the detail) */
#define PM_E_UNSUPPORTED (-3) /* valid for the reference, not built here       */
#define PM_E_NOMEM (-4)
#define PM_E_BOGUS (-5)       /* "Bogus hit returned to primer_match main()"
                                 (primer_match.cc:1249): the final alignment of
                                 a reported hit exceeded k; the reference
                                 exits(1) at this point                        */

/* Index selection: the values of primer_match's -N option this library
 * implements (select.cc:202-295).  All of them run on the same device scan;
 * the value selects the reference *semantics* of the hit list:
 *   keyword_tree / shift_and : every exact occurrence (k must be 0)
 *   filter_bitvec            : shift_and_inexact candidates, chained and
 *                              re-aligned (filter_bitvec.cc:126-236)
 *   exact_halves             : exact_halves over the exact index (k > 0; complete
 *                              only for k == 1, as in the reference)
 * PM_INDEX_AUTO follows pick_pattern_index (select.cc:100-141, a streamed FASTA
 * database, build with -DNOPRIMEGEN) for the selections available here:
 * keyword_tree when k==0 without wildcards, shift_and when k==0 with wildcards;
 * for k == 1 exact_halves when every pattern has at least 8 characters and the
 * exact-base constraints (-5/-3) are shorter than half the patterns on average
 * (resolved in pm_init, when the patterns are known); otherwise, for k > 0,
 * exact_bases when every pattern has at least 6 constrained characters, else
 * filter_bitvec.                                                                   */
#define PM_INDEX_AUTO 0
#define PM_INDEX_KEYWORD_TREE 3  /* select.cc:221-229, keyword_tree.t:427      */
#define PM_INDEX_SHIFT_AND 4     /* select.cc:230-233, shift_and.cc:204        */
#define PM_INDEX_FILTER_BITVEC 5 /* select.cc:234-237, filter_bitvec.cc:126    */
#define PM_INDEX_EXACT_BASES 9   /* select.cc:238-245 (-N 7..10), exact_bases.cc:68: exact
                                    occurrences of the constrained prefix / suffix (-s/-e/-5/-3),
                                    extended by primer_alignment; what pick_pattern_index takes
                                    when every pattern has >= 6 constrained characters   */
#define PM_INDEX_EXACT_HALVES 13 /* select.cc:246-261 (-N 11..14), exact_halves.cc:118:
                                    exact occurrences of the pattern halves, each
                                    extended by primer_alignment (primer_alignment.cc);
                                    what pick_pattern_index takes for k == 1           */

typedef struct pm_ctx pm_ctx;

/* Options of one scan: the arguments pick_pattern_index receives
 * (select.h:8-19) plus the device to run on.                                  */
typedef struct pm_options {
  int device;        /* CUDA device ordinal                                    */
  int index;         /* PM_INDEX_*  (-N)                                       */
  int k;             /* nmismatch   (-k / -K), 0..PM_MAX_K                     */
  int indels;        /* 1: -k (edits)  0: -K (substitutions only)              */
  int wildcards;     /* -w / -W: honour IUPAC codes                            */
  int text_n;        /* -W: a text 'N' matches                                 */
  unsigned char eos; /* end-of-sequence character (-E, default '\n')           */
} pm_options;

/* Which device scan produces the candidates (the hit list is the same):
 *   PLANES  bit-plane kernels (pm_scan.cuh): any alphabet, IUPAC, patterns <= 64
 *   INDEX   pair-seed index kernels (pm_index.cuh): text bytes in <= 4 match
 *           classes, patterns <= 32, pieces of >= 4 characters
 *   AUTO    INDEX when applicable, else PLANES                                 */
#define PM_SCAN_AUTO 0
#define PM_SCAN_PLANES 1
#define PM_SCAN_INDEX 2

#define PM_MAX_K 5   /* k <= 3: pair-seed index scan; 4, 5: bit-plane kernels */
#define PM_MAX_PATTERN_LEN 64
#define PM_MAX_CODES 72 /* >= PM_MAX_PATTERN_LEN + PM_MAX_K + 1 */

/* One hit after the final alignment, the numbers primer_match hands to
 * alignformat() (primer_match.cc:1158-1200).  Stream coordinates count
 * characters of the normalised database stream (eos, seq1, eos, seq2 ...),
 * exactly CharacterProducer::pos().                                           */
typedef struct pm_hit {
  int64_t start;       /* pattern_alignment::start()                           */
  int64_t end;         /* %E  pattern_alignment::end()                         */
  int64_t first;       /* pos of the first raw candidate of the hit's chain
                          (= end for the exact indexes): the order in which
                          the reference reports hits, and the ownership key
                          of pm_find_patterns' [begin, end)                    */
  uint32_t pattern_id; /* id passed to pm_add_pattern                          */
  int32_t editdist;    /* %d  pattern_alignment::editdist()                    */
  int32_t value;       /* editdist_alignment::value() (DP score)               */
  uint32_t ncodes;     /* alignment length                                     */
  uint8_t codes[PM_MAX_CODES]; /* alignment_code per column, pattern order
                                  (alignment_code.h:5-19)                      */
} pm_hit;

/* One raw candidate of the inner index (PatternMatch::find_patterns of
 * shift_and_inexact.cc:237, shift_and.cc:204, keyword_tree.t:427).            */
typedef struct pm_candidate {
  int64_t pos;         /* CharacterProducer::pos() after the last character    */
  uint32_t pattern_id;
  uint32_t reserved;
} pm_candidate;

const char *pm_last_error(void);
const char *pm_version(void);

/* new <index>(...) as pick_pattern_index does (select.cc:202-295).            */
int pm_create(pm_ctx **out, const pm_options *opt);
/* delete kt                                                                   */
void pm_destroy(pm_ctx *ctx);

/* Which tool's final alignment the hits get.  They differ for exact search (k = 0) with
 * wildcards only: primer_match makes an exact_wc_alignment (primer_match.cc:1137-1141), which
 * asks iupac_compatible(text, pattern) (pattern_alignment.cc:79); pcr_match makes an
 * editdist_alignment there too (pcr_match.cc:1096-1100), which asks iupac_compatible(pattern,
 * text) (pattern_alignment.cc:317).  A hit the chosen alignment rejects has editdist > k
 * (PM_E_BOGUS).  Default PM_ALIGN_PRIMER_MATCH.                                          */
#define PM_ALIGN_PRIMER_MATCH 0
#define PM_ALIGN_PCR_MATCH 1
int pm_set_alignment_rules(pm_ctx *ctx, int rules);

/* CharacterProducer::size() of the database the index will scan (before pm_init): 256 for a
 * streamed FASTA file (the default), the size of the character table for a .sqn / .sqz
 * database.  pick_pattern_index consults it when it chooses exact_halves for k == 1: primers
 * of at least 12 characters when the alphabet has fewer than 10, of at least 8 otherwise
 * (select.cc:120-123).  Only PM_INDEX_AUTO depends on it.                                */
int pm_set_alphabet_size(pm_ctx *ctx, int size);

/* Select the device scan (before pm_init).  PM_SCAN_INDEX on patterns it cannot
 * serve makes pm_init fail with PM_E_UNSUPPORTED.                              */
int pm_set_scan_path(pm_ctx *ctx, int path);

/* Device work of this context goes to `cuda_stream` (a cudaStream_t).         */
int pm_set_stream(pm_ctx *ctx, void *cuda_stream);

/* PatternMatch::add_pattern (pattern_match.h:118-121).  id must be > 0 and
 * unique; exact_start_bases / exact_end_bases are the per-pattern constraints
 * of primer_match -s/-e/-5/-3 (primer_match.cc:990-1017).                     */
int pm_add_pattern(pm_ctx *ctx, const char *pattern, uint32_t id, int exact_start_bases,
                   int exact_end_bases);

/* PatternMatch::init (pattern_match.h:132): builds the per-symbol match table
 * (the relation shift_and_inexact::computeu encodes in u_,
 * shift_and_inexact.cc:88-178) and uploads the pattern tables.  No pattern
 * may be added afterwards.                                                    */
int pm_init(pm_ctx *ctx);

/* The CharacterProducer the index scans.  `stream` is the normalised database
 * stream (see pm_normalize_fasta); `stream_offset` is the global position of
 * stream[0] (0 for a whole database; the shard's first position when a
 * database is partitioned over several contexts / GPUs, in which case the
 * shard must carry PM_SHARD_OVERLAP(ctx) extra characters to the left of the
 * range it is asked to report).  _host starts a chunked copy + pack pipeline
 * on a private copy stream and returns; the scan of the next pm_find_patterns
 * overlaps the rest of the transfer.  A page-locked `stream` is read directly
 * and must stay unchanged until that next blocking call returns; pageable
 * memory is staged through pinned buffers before the call returns.  _device
 * adopts a 4-byte-aligned device pointer that must stay valid until the next
 * pm_set_text_* / pm_destroy.  Both run the pack kernel (per-symbol match
 * bit-planes).                                                                */
int pm_set_text_host(pm_ctx *ctx, const uint8_t *stream, uint64_t n, int64_t stream_offset);
int pm_set_text_device(pm_ctx *ctx, const uint8_t *d_stream, uint64_t n, int64_t stream_offset);

/* IndexedFastaFile<Compressed<...>> (select.t:94-106; char_io.t:18-215): the text is
 * the `.sqz` bit stream `compress_seq -z true` writes (compress_seq.cc:745-870),
 * which primer_match opens when the file exists (-D 0) or is asked for (-D 4).
 * bits = pm_sqz_bits(table_size) bits per character, most significant bit first,
 * character i at bit i*bits; `table` is the content of the `.tbz` file.  n_chars =
 * characters of the stream (the `.seq` length); 0 = all the bytes hold (the writer
 * pads its last buffer with end-of-sequence characters, which find nothing).  A
 * shard (stream_offset > 0) must start on a byte boundary of the bit stream.
 * Same pipeline as pm_set_text_host with 8/bits times fewer bytes over PCIe; the
 * stream bytes are rebuilt on the device (unpack_sqz_kernel).                   */
int pm_set_text_sqz_host(pm_ctx *ctx, const uint8_t *sqz, uint64_t n_bytes, const uint8_t *table,
                         uint32_t table_size, uint64_t n_chars, int64_t stream_offset);
/* IndexedFastaFile<Normalized<...>> (select.t:29-44; char_io.t:216-283): the `.sqn` file
 * `compress_seq -n true` writes, one index into the `.tbl` table per character; what
 * primer_match opens first when it exists (-D 0 / -D 3).                              */
int pm_set_text_sqn_host(pm_ctx *ctx, const uint8_t *sqn, uint64_t n_chars, const uint8_t *table,
                         uint32_t table_size, int64_t stream_offset);
/* bits per character of a .sqz stream with this table size (char_io.t:58-61).  */
int pm_sqz_bits(uint32_t table_size);

/* Characters of left context a shard needs so that every hit it reports is
 * identical to a whole-database scan: longest pattern + 3k + 2.              */
int pm_shard_overlap(pm_ctx *ctx);

/* PatternMatch::find_patterns followed, per returned hit, by the final
 * alignment of primer_match.cc:1132-1153.  Reports the hits whose FIRST raw
 * candidate position (pm_candidate.pos of the head of its chain) lies in
 * [begin, end) of the global stream, sorted by (first, pattern_id): the order
 * the reference reports them in, up to hits that begin at one position (pm_hit_order
 * settles those).  The hits of exact_halves / exact_bases (first = end there) come in the
 * order of the half / exact-part occurrences they grew from, which is the reference's
 * (exact_halves.cc:113-116, exact_bases.cc:68-124).  *hits
 * points into pinned memory owned by the context, valid until the next call
 * on it.  Pass begin=0,end=INT64_MAX for the whole stream (what the reference
 * does).  PM_E_BOGUS (a hit whose final alignment has more than k errors, on
 * which primer_match stops) still delivers every hit: pcr_match keeps going
 * and skips those (pcr_match.cc:1114).                                         */
int pm_find_patterns(pm_ctx *ctx, int64_t begin, int64_t end, const pm_hit **hits,
                     uint64_t *nhits);

/* Same work, results left on the device: *d_hits is a device pointer to
 * *nhits pm_hit records.                                                      */
int pm_find_patterns_device(pm_ctx *ctx, int64_t begin, int64_t end, const pm_hit **d_hits,
                            uint64_t *nhits);

/* The inner index alone: raw candidates with pos in [begin, end), sorted by
 * (pattern_id, pos).  For tests and for callers that do their own filtering.  */
int pm_scan_candidates(pm_ctx *ctx, int64_t begin, int64_t end, const pm_candidate **cands,
                       uint64_t *ncands);

/* The raw candidates of the last pm_find_patterns / pm_find_patterns_device call on this
 * context (every candidate of the scanned range, sorted by (pattern_id, pos)): what
 * shift_and_inexact handed filter_bitvec (filter_bitvec.cc:89).  PM_E_UNSUPPORTED for the
 * exact_halves / exact_bases indexes, whose hits do not come from such a list.  *cands is
 * owned by the context, valid until the next call on it.                                  */
int pm_last_candidates(pm_ctx *ctx, const pm_candidate **cands, uint64_t *ncands);

/* The order in which filter_bitvec::find_patterns reports `hits` (filter_bitvec.cc:89-141):
 * the candidates as shift_and_inexact emits them (position, then pattern), std::sort on the
 * position alone, chains taken in that order.  order[i] = index into hits of the i-th hit
 * reported; hits without a chain in `cands` follow in their given order.  Host code.       */
int pm_hit_order(const pm_hit *hits, uint64_t nhits, const pm_candidate *cands, uint64_t ncands, int k,
                 uint64_t *order);

/* PatternMatch::reset (pattern_match.h:136): forget the text.                 */
int pm_reset(pm_ctx *ctx);

/* Counters of the last pm_find_patterns / pm_scan_candidates call.            */
typedef struct pm_stats {
  uint64_t text_len;        /* characters of the loaded shard                  */
  uint64_t n_patterns;
  uint64_t n_seeds;         /* edit mode: seed hits sent to the verifier       */
  uint64_t n_candidates;
  uint64_t n_chains;
  uint64_t n_hits;
  uint64_t n_bogus;
  uint64_t kernel_launches; /* kernels of this library launched by the call    */
  uint64_t scan_items;      /* (tile, pattern-chunk) blocks of the scan kernel */
  uint64_t scan_warp_steps; /* PLANES: (warp, pattern character) steps the dense
                               scan executed, each 8 words x 32 positions;
                               INDEX: level-1 survivor records sent to level 2 */
  uint32_t scan_path;       /* PM_SCAN_PLANES or PM_SCAN_INDEX                  */
  uint32_t index_q;         /* INDEX: piece length L of the pair index          */
  uint32_t index_nd;        /* INDEX: text distances looked up per position     */
  uint32_t index_pieces;    /* INDEX: pieces per pattern (k + 2)                */
  uint64_t index_lookups;   /* INDEX: level-1 lookups (positions x distances)   */
  uint64_t index_entries;   /* INDEX: (pair key, distance, pattern) entries     */
  uint64_t n_ambiguous;     /* INDEX: text characters re-examined on the raw text
                               (they match pattern symbols without being one of
                               the four text codes, e.g. a text 'R' under -w)   */
  uint64_t host_syncs;      /* times the call waited for the device: 2 for the
                               bit-vector and exact indexes (scan counters, then
                               every other count together with the records)    */
  float pack_ms;            /* device time of the last pack kernel             */
  float scan_ms;            /* device time of the dense scan kernel            */
  float total_ms;           /* device time of the whole call                   */
} pm_stats;
int pm_get_stats(pm_ctx *ctx, pm_stats *out);

/* StreamedFastaFile (fasta_io.t:448-751): raw FASTA bytes -> normalised
 * stream.  out must hold n+2 bytes.  seq_key[i] = stream index of the first
 * base of sequence i (the key get_seq_pos subtracts, fasta_io.t:194-202);
 * hdr_off/hdr_len = byte range of its header in raw.  Host-side, O(n).        */
int pm_normalize_fasta(const uint8_t *raw, uint64_t n, int upper_case, unsigned char eos,
                       uint8_t *out, uint64_t *out_len, int64_t *seq_key, int64_t *hdr_off,
                       int64_t *hdr_len, uint64_t max_seqs, uint64_t *n_seqs);

/* The same on the device (pm_fasta.cuh): `raw` = the bytes of a FASTA file in host
 * memory.  They are copied to the device and normalised there (a three-state machine
 * over the raw bytes, composed with a prefix scan), then packed as by
 * pm_set_text_device; stream_offset is 0.  Call after pm_init.  The sequence table
 * (what SortSeqFastaFile keeps for get_seq_pos, fasta_io.t:163-202) is returned by
 * pm_get_sequences: pointers into memory owned by the context, valid until the next
 * pm_set_text_* call.  pm_get_text copies characters of the loaded stream back to the
 * host (the alignment buffers primer_match prints with %t come from there).          */
int pm_set_text_fasta_host(pm_ctx *ctx, const uint8_t *raw, uint64_t n_raw, int upper_case);
int pm_get_sequences(pm_ctx *ctx, const int64_t **seq_key, const int64_t **hdr_off,
                     const int64_t **hdr_len, uint64_t *n_seqs, uint64_t *stream_len);
int pm_get_text(pm_ctx *ctx, uint64_t begin, uint64_t len, uint8_t *out);

/* pick_fasta_file (select.t:23-190) and the reader in one call: opens `database` the way
 * primer_match does for its -D value `dbind` -- 0: <db>.sqn if it exists, else <db>.sqz, else
 * <db>.seq, else the FASTA file itself; 1: the FASTA file; 2: <db>.seq; 3: <db>.sqn + <db>.tbl;
 * 4: <db>.sqz + <db>.tbz (the files compress_seq writes, with <db>.idb / <db>.hdr for the sequence
 * index) -- and reads it with a few threads into a ring of page-locked chunk buffers while the copy
 * engine moves the chunks that are full and the device normalises / unpacks / packs them: the
 * pinned, double-buffered host reader of the input path (replaces MapFileChars / BufferedFileChars,
 * char_io.h:150-190).  The file never sits in host memory as a whole.  pm_get_sequences and
 * pm_get_header serve what FastaFile::get_seq_pos / get_header_data serve (fasta_io.t:163-215).
 * Call after pm_init.  PM_READER_THREADS (environment) sets the reader threads (default 6).      */
int pm_set_text_database(pm_ctx *ctx, const char *database, int dbind, int upper_case);
/* Header line of sequence i (0-based, without '>' and the line end); valid until the next
 * pm_set_text_* call.                                                                         */
int pm_get_header(pm_ctx *ctx, uint64_t i, const char **text, uint64_t *len);

/* Characters of n ranges [begin[i], begin[i] + len[i]) of the loaded stream, concatenated into
 * `out`: one gather on the device and one copy, for the alignment buffers a tool prints next to a
 * batch of hits (pm_get_text costs a round trip per range).                                     */
int pm_gather_text(pm_ctx *ctx, const int64_t *begin, const int64_t *len, uint64_t n, uint8_t *out);

/* Page-locked host buffers for the database (the pinned reader side of the input
 * path): pm_set_text_host / _sqz_host / _fasta_host read such memory directly by DMA.
 * write_combined != 0 asks for write-combined pages: right for a buffer the CPU only
 * fills (file read) and never reads back.                                            */
int pm_host_alloc(uint64_t bytes, int write_combined, void **out);
int pm_host_free(void *p);

/* Diagnostic for the roofline: measured throughput of 3-input logic
 * instructions (LOP3) on `device`, in 1e12 thread-level operations per second. */
int pm_bench_int_peak(int device, double *lop3_tera_ops);

/* pcr_match's pairing of hits into amplicons (pcr_match.cc:945-1170), on the records
 * pm_find_patterns returned for the pattern array pcr_match builds (pcr_match.cc:800-913):
 * ids 1..n_primers the primers as given (pairs (1,2), (3,4) ...; the second of each pair
 * reverse complemented first under -r / -S), ids n_primers+1..2*n_primers their reverse
 * complements.  Host code; no device work.                                             */
typedef struct pm_pcr_options {
  int k;                /* -k / -K                                                       */
  int indels;           /* 1: -k, 0: -K                                                  */
  int allorient;        /* -a: every orientation of a pair, pcr_match.cc:977-997          */
  int betweenlen;       /* -b: amplicon length measured between the primers               */
  int deviation;        /* -d (with sizes below); < 0: none                               */
  int reserved;
  int64_t mindist;      /* -m                                                             */
  int64_t maxdist;      /* -M                                                             */
  const uint64_t *sts_sizelb; /* per pair: sts_entry::sizelb()/sizeub() of a UniSTS file  */
  const uint64_t *sts_sizeub; /* (-S), or NULL                                            */
  /* The raw candidates behind the hits (pm_last_candidates), or NULL.  The reference leaves
   * hits that end at one position in the order two std::sort calls on position alone leave
   * them -- filter_bitvec.cc:91 on the candidates, pcr_match.cc:958 on the hits -- and a hit
   * only pairs with partners that come later, so that order is part of its answer.  With the
   * candidates pm_pcr_pairs reproduces both sorts; without them the first is taken as stable. */
  const pm_candidate *cands;
  uint64_t ncands;
} pm_pcr_options;

typedef struct pm_pcr_pair {
  uint64_t first;       /* index into `hits`: the hit that found its partner (%>x codes)  */
  uint64_t second;      /* index into `hits`: the partner (%<x codes)                     */
  int64_t amplicon_len; /* pcr_match.cc:1151-1157                                         */
  uint32_t pair;        /* %i: 1-based primer pair                                        */
  uint32_t reserved;
} pm_pcr_pair;

/* Pairs in the order pcr_match prints them.  patlen[id] = length of pattern id (index 0
 * unused, 2*n_primers+1 entries); seq_key/n_seqs as pm_get_sequences returns them (a pair
 * must lie inside one FASTA entry, fasta_io.t:203-213).  *pairs is allocated with malloc;
 * release it with pm_free.                                                             */
int pm_pcr_pairs(const pm_hit *hits, uint64_t nhits, uint32_t n_primers, const uint32_t *patlen,
                 const int64_t *seq_key, uint64_t n_seqs, const pm_pcr_options *opt,
                 pm_pcr_pair **pairs, uint64_t *npairs);
void pm_free(void *p);

/* util.cc:374-381 reverse_comp (IUPAC aware, util.cc:319-348): out must hold
 * strlen(pattern)+1 bytes.                                                    */
int pm_reverse_comp(const char *pattern, char *out);

#ifdef __cplusplus
}
#endif
#endif /* PM_B200_H */
