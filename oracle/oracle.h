/*
 * oracle/oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C interface of the CPU restatement of Teloscope's scan core (oracle.cpp).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library; nothing under teloscope_b200/ links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement against outputs of
 * the unmodified reference binary (oracle/_ref/teloscope, built by oracle/Makefile from
 * /root/reference) committed under tests/golden/, and against the reference's own .tst goldens.
 *
 * Record layouts are deliberately the same as include/teloscope_b200.h so tests can compare
 * arrays directly.
 */
#ifndef TELO_ORACLE_H
#define TELO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint64_t pos;       /* absPos + offset in the segment                    */
    uint16_t len;       /* matchSize                                         */
    uint8_t is_forward; /* pattern orientation flag                          */
    uint8_t is_canonical;
    uint32_t pad_;
} orc_match;            /* 16 bytes */

typedef struct {
    uint64_t start;
    uint32_t len, counts, fwd_count, rev_count, can_count, noncan_count;
    uint32_t total_cov, fwd_cov, can_cov;
    uint8_t has_valid_or;
    char label;          /* 'p','q','b' */
    uint16_t pad_;
} orc_block;             /* 48 bytes */

typedef struct {
    uint64_t start;      /* absPos + windowStart */
    uint32_t size;       /* currentWindowSize    */
    uint32_t acgt[4];    /* A,C,G,T              */
    uint32_t can_cov, noncan_cov, fwd_cov, rev_cov;
    float gc, entropy;   /* reference float math (include/teloscope.h:199-214) */
} orc_window;            /* 56 bytes */

typedef struct {
    uint32_t window_size, step, terminal_limit;
    uint16_t max_match_dist, min_block_len, max_block_dist, min_block_counts;
    float min_block_density;
    uint8_t count_acgt;  /* outGC || outEntropy (src/teloscope.cpp:418) */
    uint8_t pad_[3];
} orc_params;

typedef struct orc_ctx orc_ctx;
typedef struct orc_result orc_result;

/* Pattern table producer (src/tools.cpp:201-283).  raw = comma separated seeds (already
 * upper-cased), canonical_fwd = lexicographically smaller of motif/revcomp.  Writes
 * "PATTERN\t<0|1>\n" lines (sorted, unique) into out; returns number of patterns or -1 if the
 * buffer is too small, -2 on an invalid nucleotide. */
int orc_expand_patterns(const char *raw_csv, int edit_distance, const char *canonical_fwd,
                        char *out, size_t out_len);

/* patterns: n NUL-terminated strings back to back; fwd[i] = orientation flag.
 * canonical_fwd/rev: the two canonical strings (include/teloscope.h:241-247). */
orc_ctx *orc_create(const char *patterns, const uint8_t *fwd, int n, const char *canonical_fwd,
                    const char *canonical_rev, const orc_params *params);
void orc_destroy(orc_ctx *);

/* Teloscope::scanSegment (src/teloscope.cpp:537-658).  seq must already be upper-cased by the
 * caller exactly as the reference's callers do (unmaskSequence). */
orc_result *orc_scan_segment(const orc_ctx *, const char *seq, uint64_t n, uint64_t abs_pos,
                             int tips_only);
void orc_result_free(orc_result *);

/* which: 0 canonicalMatches, 1 nonCanonicalMatches, 2 fwdMatches, 3 revMatches, 4 allMatches */
size_t orc_result_matches(const orc_result *, int which, const orc_match **out);
/* which: 0 terminalBlocks, 1 interstitialBlocks */
size_t orc_result_blocks(const orc_result *, int which, const orc_block **out);
size_t orc_result_windows(const orc_result *, const orc_window **out);

/* ReadTelomereFilter::matches (src/read-filter.cpp:10-45); params.min_block_len must already be
 * 42 unless -l was given; terminal_limit is overridden to UINT32_MAX/2 here. */
int orc_read_matches(const orc_ctx *, const char *seq, uint64_t n);

/* getGCContent / getShannonEntropy (include/teloscope.h:199-214) on given counters. */
void orc_window_floats(const uint32_t acgt[4], uint32_t size, float *gc, float *entropy);

/* Bulk helper for the CPU baseline: scans `count` segments laid out back to back (offsets has
 * count+1 entries) with `threads` std::threads, one job per segment like the reference's pool,
 * and returns a checksum of the results so the work cannot be optimised away. */
uint64_t orc_scan_many(const orc_ctx *, const char *buf, const uint64_t *offsets, int count,
                       int tips_only, int threads);

#ifdef __cplusplus
}
#endif
#endif
